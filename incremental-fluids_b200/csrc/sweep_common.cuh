// sweep_common.cuh -- PTX helpers of the wavefront kernels (sweep_kernels.cu).
#pragma once
#include "ifl_internal.cuh"

#include <cuda.h>

namespace ifl {

// Every in-kernel dependency wait is bounded in TIME (%globaltimer), not in polls: a poll of a
// shared-memory counter costs ~40 ns, an L2 / NVLink poll ~1 us and an mbarrier try_wait may
// suspend for microseconds, so a poll count means nothing.  Two seconds is far beyond any legal
// wait (with several ranks the upstream slab may start milliseconds later) and still returns the
// GPU promptly when the protocol is broken or a peer has died.
constexpr unsigned long long WATCHDOG_NS = 2000000000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Spin-loop bookkeeping: `n` counts polls, `t0` is latched on the first slow-path check; the
// clock is read every 256 polls only, so the fast path stays two instructions long.
struct Watch {
    unsigned n;
    unsigned long long t0;
    __device__ __forceinline__ Watch() : n(0), t0(0) {}
    __device__ __forceinline__ bool expired(volatile int *dead) {
        if ((++n & 255u) != 0) return false;
        if (*dead) return true;
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        return now - t0 > WATCHDOG_NS;
    }
};

int sweep_get_map(ifl_ctx *c, const Arr &a, int box_w, int box_h, CUtensorMap *out);

// ------------------------------------------------------------------ PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) { // same, on a shared-space address
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: every dependency wait in this file gives up after a (very long) poll
// budget, raises the watchdog flags and lets the kernel run to completion with garbage,
// so that a protocol bug can never hang the device.  `dead` is a CTA-wide shared flag.
__device__ __forceinline__ bool mbar_try(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity, volatile int *dead, SolveScalars *scal) {
    if (mbar_try(bar, parity)) return;
    if (*dead) return;
    Watch w;
    while (!mbar_try(bar, parity)) {
        if (w.expired(dead)) {
            *dead = 1;
            scal->watchdog = 1;
            return;
        }
    }
}
// Same, for the helper warps of the staircase engine: the try_wait carries a suspend-time hint, so a waiting warp
// re-polls every ~100 us instead of every few hundred ns (it is still woken when the phase completes).  Measured:
// two storer warps + loader waiting with the default limit cost the compute warp 5.6 cycles per step
// (profiles/r02_tri_experiments.txt, section 7).
__device__ __forceinline__ bool mbar_try_hint(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_hint(uint64_t *bar, unsigned parity, volatile int *dead, SolveScalars *scal) {
    if (mbar_try_hint(bar, parity)) return;
    if (*dead) return;
    const unsigned long long t0 = globaltimer_ns();
    while (!mbar_try_hint(bar, parity)) {
        if (*dead) return;
        if (globaltimer_ns() - t0 > WATCHDOG_NS) {
            *dead = 1;
            scal->watchdog = 1;
            return;
        }
    }
}
// For waits that are off the critical path (ring slots with stages of slack): non-blocking test + nanosleep, so the
// waiting warp issues nothing at all in between.
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, unsigned parity, volatile int *dead, SolveScalars *scal) {
    if (mbar_test(bar, parity)) return;
    if (*dead) return;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        __nanosleep(128);
        if (mbar_test(bar, parity)) return;
        if (*dead) return;
        if (globaltimer_ns() - t0 > WATCHDOG_NS) {
            *dead = 1;
            scal->watchdog = 1;
            return;
        }
    }
}
// TMA: one 2-D box global -> shared, completion counted in bytes on an mbarrier.
// Out-of-range rows (y = -1 for the first strip) are filled with zeros by the hardware.
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- thread-block cluster / distributed shared memory
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of THIS CTA's layout) in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_f64(uint32_t raddr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_remote_u32_release(uint32_t raddr, unsigned v) {
    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_remote_u32(uint32_t raddr) {
    unsigned v;
    asm volatile("ld.relaxed.cluster.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u32_acquire(uint32_t a) { // pairs with st_remote_u32_release
    unsigned v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// NCCL-LL style message: {lo, epoch, hi, epoch} in one 16-byte store / load.
__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned epoch) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch)
                 : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4 *src, unsigned epoch, double &v) {
    unsigned a, b, c, d;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
    v = __hiloint2double((int)c, (int)a);
    return b == epoch && d == epoch;
}

// The same message across a slab boundary: the producer is another GPU storing into this
// rank's hand-off array over NVLink, so both sides use system scope (8-byte halves stay
// atomic on the wire, which is all the protocol needs).
__device__ __forceinline__ void ll_store_sys(uint4 *dst, double v, unsigned epoch) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch)
                 : "memory");
}
__device__ __forceinline__ bool ll_load_sys(const uint4 *src, unsigned epoch, double &v) {
    unsigned a, b, c, d;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
    v = __hiloint2double((int)c, (int)a);
    return b == epoch && d == epoch;
}

// Shared-memory accessors on 32-bit shared-space byte addresses: the per-step address is
// `selected base + compile-time offset`, which ptxas folds into the instruction's
// immediate field, so a step spends one ISETP + one SEL on addressing.
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lds_f64_if(double &v, uint32_t a, bool pred) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p ld.shared.f64 %0, [%1];\n\t"
        "}"
        : "+d"(v)
        : "r"(a), "r"((unsigned)pred)
        : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// store only if `pred` (EDGE macro-steps: lanes outside the strip run the same code, unbranched)
template <bool ALWAYS>
__device__ __forceinline__ void sts_f64_p(uint32_t a, double v, bool pred) {
    if (ALWAYS) {
        sts_f64(a, v);
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %2, 0;\n\t"
            "@p st.shared.f64 [%0], %1;\n\t"
            "}" ::"r"(a),
            "d"(v), "r"((unsigned)pred)
            : "memory");
    }
}
__device__ __forceinline__ void sts_u32_volatile(uint32_t a, unsigned v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u32_volatile(uint32_t a) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// Bounded spin until the shared counter at `a` reaches `need`.  ACQUIRE = cluster-scope
// acquire loads (the counter is bumped by a remote CTA through DSMEM); ptxas implements
// those as LDS + CCTL.IVALL (an L1 invalidate), so the CTA-local paths use plain volatile
// loads: shared-memory accesses of one SM are ordered anyway.
template <bool ACQUIRE>
__device__ __forceinline__ unsigned counter_load(uint32_t a) {
    return ACQUIRE ? lds_u32_acquire(a) : lds_u32_volatile(a);
}
template <bool ACQUIRE = false>
__device__ __forceinline__ void wait_counter(uint32_t a, unsigned need, volatile int *dead, SolveScalars *scal) {
    if (counter_load<ACQUIRE>(a) >= need) return;
    Watch w;
    while (counter_load<ACQUIRE>(a) < need) {
        if (w.expired(dead)) {
            *dead = 1;
            scal->watchdog = 1;
            return;
        }
    }
}
__device__ __forceinline__ double sel_f64(bool pred, double a, double b) { // pred ? a : b, one select deep
    double r;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %3, 0;\n\t"
        "selp.f64 %0, %1, %2, p;\n\t"
        "}"
        : "=d"(r)
        : "d"(a), "d"(b), "r"((unsigned)pred));
    return r;
}

} // namespace ifl
