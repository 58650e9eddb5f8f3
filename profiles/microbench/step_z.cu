// Round-2 questions about the one-row-per-lane wavefront step (sweep_kernels.cu):
//   1. what exactly does "another warp is alive in the CTA" cost the recurrence warp, and does it
//      depend on WHERE that warp lives (same / other SM sub-partition) and HOW it waits
//      (nanosleep, mbarrier try_wait, spinning on shared memory)?
//   2. what does a per-step hand-off store by lane 31 cost (st.async + complete_tx, LL message
//      through st.shared::cluster, plain remote f64 store)?
//   3. how long does a hand-off take CTA -> CTA inside a cluster (ping-pong, half round trip):
//      LL message + polling, st.async + mbarrier try_wait?
// Build: nvcc -arch=sm_100a -fmad=false -O3 -o step_z step_z.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double sel_f64(bool pred, double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((unsigned)pred));
    return r;
}
__device__ __forceinline__ void sts_u32_volatile(uint32_t a, unsigned v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned lds_u32_volatile(uint32_t a) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}

constexpr int TILE_BYTES = 33 * 32 * 8;
constexpr int NT = 5, NST = 4;

// helper-warp behaviour
enum { H_NONE = 0, H_NANOSLEEP = 1, H_MBAR = 2, H_SPIN_SMEM = 3 };
// hand-off store of lane 31
enum { S_NONE = 0, S_ASYNC = 1, S_LL = 2, S_F64 = 3 };

struct Cfg {
    int helper_kind;  // H_*
    int helper_mask;  // bit w set: warp w is a helper (warp 0 is the compute warp)
    int sync;         // progress store + prefetched hand-off counter check every 8 steps (as in the product kernel)
    int mbar;         // per macro-step: syncwarp + arrive(done) and try_wait(full)
    int fence;        // fence.proxy.async before the arrive
};

template <int STORE>
__global__ void __launch_bounds__(512, 1) k_step(double *out, long long *cyc, int nmacro, int slot, Cfg cfg) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ unsigned counters[4];
    __shared__ __align__(16) double ring[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = cluster_ctarank();
    double *sm = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < NST * NT * 33 * 32; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13);
    if (threadIdx.x == 0) {
        counters[0] = 0;
        counters[1] = 1u << 30;
        counters[2] = 0;
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    }
    __syncthreads();
    cluster_sync_all();
    if (rank != 0) { // the receiving CTA of the hand-off stores just stays resident
        if (threadIdx.x == 0) {
            unsigned n = 0;
            while (lds_u32_volatile(smem_u32(&counters[2])) == 0 && ++n < (1u << 22)) __nanosleep(200);
        }
        __syncthreads();
        cluster_sync_all();
        return;
    }
    if (warp != 0) {
        if ((cfg.helper_mask >> warp) & 1) {
            if (cfg.helper_kind == H_NANOSLEEP) {
                unsigned n = 0;
                while (lds_u32_volatile(smem_u32(&counters[2])) == 0 && ++n < (1u << 22)) __nanosleep(20000);
            } else if (cfg.helper_kind == H_MBAR) {
                unsigned n = 0;
                while (!mbar_try(smem_u32(&bar[3]), 0) && ++n < (1u << 22)) {}
            } else if (cfg.helper_kind == H_SPIN_SMEM) {
                unsigned n = 0;
                while (lds_u32_volatile(smem_u32(&counters[2])) == 0 && ++n < (1u << 26)) {}
            }
        }
    } else {
        const uint32_t base0 = smem_u32(smem) + (uint32_t)((1 + lane) * 256);
        const uint32_t halo0 = smem_u32(smem) + (NST * NT - 1) * TILE_BYTES;
        const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
        const uint32_t r_ring = mapa(smem_u32(ring), 1), r_bar = mapa(smem_u32(&bar[2]), 1);
        double z = 0.5 + lane * 1e-3, c1 = 1e-3;
        double a = 1.0, cx = 1e-3, cy = 2e-3, pr = 0.999, halo = 0.25;
        unsigned seen = 0;
        const long long t0 = clock64();
        for (int m = 0; m < nmacro; m++) {
            const uint32_t sA = base0 + (uint32_t)((m % NST) * NT * TILE_BYTES) - (uint32_t)(8 * lane);
            const uint32_t sB = base0 + (uint32_t)(((m + NST - 1) % NST) * NT * TILE_BYTES) + 256u - (uint32_t)(8 * lane);
#pragma unroll
            for (int kk = 0; kk < 32; kk++) {
                if (cfg.sync) {
                    if (((kk + 5) % 8) == 0) seen = lds_u32_volatile(halo_cols_addr);
                    if (((kk + 1) % 8) == 0) {
                        const unsigned need = (unsigned)(32 * m + kk + 9);
                        unsigned n = 0;
                        if (seen < need)
                            while (lds_u32_volatile(halo_cols_addr) < need && ++n < 1000) {}
                    }
                }
                double up = __shfl_up_sync(0xffffffffu, z, 1);
                const uint32_t b = (lane > kk + 1) ? sB : sA;
                const uint32_t p = b + (uint32_t)(8 * (kk + 1));
                const double na = lds_f64(p);
                const double ncx = lds_f64(p + TILE_BYTES);
                const double ncy = lds_f64(p + 2 * TILE_BYTES - 256);
                const double npr = lds_f64(p + 3 * TILE_BYTES);
                const double nh = lds_f64(halo0 + (uint32_t)(8 * ((kk + 1) & 31)));
                up = sel_f64(lane == 0, halo, up);
                double t = a - c1 * z;
                t = t - cy * up;
                z = t * pr;
                c1 = cx;
                const uint32_t bs = (lane > kk) ? sB : sA;
                sts_f64(bs + (uint32_t)(8 * kk) + 4 * TILE_BYTES, z);
                if (STORE == S_ASYNC) {
                    if (lane == 31)
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                                         r_ring + (uint32_t)(8 * kk)),
                                     "l"(__double_as_longlong(z)), "r"(r_bar)
                                     : "memory");
                } else if (STORE == S_LL) {
                    if (lane == 31) {
                        const unsigned lo = (unsigned)__double2loint(z), hi = (unsigned)__double2hiint(z), tag = (unsigned)(m + 1);
                        asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r_ring + (uint32_t)(16 * kk)), "r"(lo), "r"(tag),
                                     "r"(hi), "r"(tag)
                                     : "memory");
                    }
                } else if (STORE == S_F64) {
                    if (lane == 31) asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(r_ring + (uint32_t)(8 * kk)), "d"(z) : "memory");
                }
                if (cfg.sync && ((kk + 2) % 8) == 0) sts_u32_volatile(progress_addr, (unsigned)(32 * m + kk));
                a = na;
                cx = ncx;
                cy = ncy;
                pr = npr;
                halo = nh;
            }
            if (cfg.mbar) {
                if (cfg.fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
                unsigned n = 0;
                while (!mbar_try(smem_u32(&bar[0]), (unsigned)(m & 1)) && ++n < 1000) {}
            }
        }
        const long long t1 = clock64();
        if (lane == 0) {
            cyc[slot] = t1 - t0;
            sts_u32_volatile(smem_u32(&counters[2]), 1u);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[3])) : "memory");
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa(smem_u32(&counters[2]), 1)), "r"(1u) : "memory");
        }
        out[lane] = z + c1;
    }
    __syncthreads();
    cluster_sync_all();
}

template <int STORE>
static void run(const char *name, double *out, long long *cyc, int slot, Cfg cfg, int threads) {
    const int nmacro = 128;
    const size_t smem = (size_t)NST * NT * TILE_BYTES;
    cudaFuncSetAttribute(k_step<STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2);
    lc.blockDim = dim3(threads);
    lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaLaunchKernelEx(&lc, k_step<STORE>, out, cyc, nmacro, slot, cfg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("  \"%s\": \"%s\",\n", name, cudaGetErrorString(e));
            return;
        }
    }
    long long c;
    cudaMemcpy(&c, cyc + slot, 8, cudaMemcpyDeviceToHost);
    printf("  \"%s\": %.1f,\n", name, (double)c / (nmacro * 32));
}

// ---- hand-off latency: two CTAs of one cluster play ping-pong, one thread each -----------------
// MODE 0: LL message {lo, tag, hi, tag} by st.shared::cluster.v4, receiver spins on its own shared memory
// MODE 1: st.async 8 bytes + complete_tx on the receiver's mbarrier, receiver sits in try_wait
template <int MODE>
__global__ void __launch_bounds__(32, 1) k_pingpong(long long *cyc, int rounds, int slot) {
    __shared__ __align__(16) uint4 msg[2];
    __shared__ __align__(8) uint64_t bar[1];
    __shared__ __align__(8) double val[2];
    const unsigned rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        msg[0] = make_uint4(0, 0, 0, 0);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        if (MODE == 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        const uint32_t r_msg = mapa(smem_u32(&msg[0]), rank ^ 1), r_bar = mapa(smem_u32(&bar[0]), rank ^ 1),
                       r_val = mapa(smem_u32(&val[0]), rank ^ 1);
        const long long t0 = clock64();
        for (int i = 1; i <= rounds; i++) {
            if (rank == 0) { // send first, then wait for the answer
                if (MODE == 0)
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r_msg), "r"(7u), "r"((unsigned)i), "r"(9u), "r"((unsigned)i) : "memory");
                else
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(r_val), "l"((long long)i), "r"(r_bar) : "memory");
            }
            if (MODE == 0) {
                unsigned a, b, c2, d, n = 0;
                do {
                    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "r"(smem_u32(&msg[0])) : "memory");
                } while ((b != (unsigned)i || d != (unsigned)i) && ++n < (1u << 20));
            } else {
                unsigned n = 0;
                while (!mbar_try(smem_u32(&bar[0]), (unsigned)((i - 1) & 1)) && ++n < (1u << 20)) {}
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(smem_u32(&bar[0])) : "memory"); // arm the next phase
            }
            if (rank == 1) { // answer
                if (MODE == 0)
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r_msg), "r"(7u), "r"((unsigned)i), "r"(9u), "r"((unsigned)i) : "memory");
                else
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(r_val), "l"((long long)i), "r"(r_bar) : "memory");
            }
        }
        const long long t1 = clock64();
        if (rank == 0) cyc[slot] = t1 - t0;
    }
    __syncthreads();
    cluster_sync_all();
}

template <int MODE>
static void run_pp(const char *name, long long *cyc, int slot) {
    const int rounds = 2000;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2);
    lc.blockDim = dim3(32);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaLaunchKernelEx(&lc, k_pingpong<MODE>, cyc, rounds, slot);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("  \"%s\": \"%s\",\n", name, cudaGetErrorString(e));
            return;
        }
    }
    long long c;
    cudaMemcpy(&c, cyc + slot, 8, cudaMemcpyDeviceToHost);
    printf("  \"%s\": %.1f,\n", name, (double)c / (2.0 * rounds));
}

int main(int argc, char **argv) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMemset(out, 0, 1 << 20);
    cudaMalloc(&cyc, 4096);
    const int part = argc > 1 ? atoi(argv[1]) : 0;
    printf("{\n");
    int s = 0;
    if (part == 0 || part == 1) {
        run_pp<0>("pingpong_ll_half_roundtrip_cyc", cyc, s++);
    }
    if (part == 0 || part == 2) {
        run_pp<1>("pingpong_stasync_half_roundtrip_cyc", cyc, s++);
    }
    if (part == 0 || part == 3) {
        const Cfg core = {H_NONE, 0, 0, 0, 0};
        run<S_NONE>("core_1warp", out, cyc, s++, core, 32);
        run<S_NONE>("core_exited_warps", out, cyc, s++, core, 192);
        const Cfg full = {H_NONE, 0, 1, 1, 1};
        run<S_NONE>("full_exited_warps", out, cyc, s++, full, 192);
        const Cfg full_nofence = {H_NONE, 0, 1, 1, 0};
        run<S_NONE>("full_nofence_exited_warps", out, cyc, s++, full_nofence, 192);
        const Cfg f_ns3 = {H_NANOSLEEP, 1 << 3, 1, 1, 1};
        run<S_NONE>("full_nanosleep_warp3", out, cyc, s++, f_ns3, 192);
        const Cfg f_ns4 = {H_NANOSLEEP, 1 << 4, 1, 1, 1};
        run<S_NONE>("full_nanosleep_warp4_same_smsp", out, cyc, s++, f_ns4, 192);
        const Cfg f_mb3 = {H_MBAR, 1 << 3, 1, 1, 1};
        run<S_NONE>("full_mbarwait_warp3", out, cyc, s++, f_mb3, 192);
        const Cfg f_mb4 = {H_MBAR, 1 << 4, 1, 1, 1};
        run<S_NONE>("full_mbarwait_warp4_same_smsp", out, cyc, s++, f_mb4, 192);
        const Cfg f_mb123 = {H_MBAR, (1 << 1) | (1 << 2) | (1 << 3), 1, 1, 1};
        run<S_NONE>("full_mbarwait_warps123", out, cyc, s++, f_mb123, 192);
        const Cfg f_mb4812 = {H_MBAR, (1 << 4) | (1 << 8) | (1 << 12), 1, 1, 1};
        run<S_NONE>("full_mbarwait_warps4_8_12_same_smsp", out, cyc, s++, f_mb4812, 512);
        const Cfg f_sp3 = {H_SPIN_SMEM, 1 << 3, 1, 1, 1};
        run<S_NONE>("full_spin_smem_warp3", out, cyc, s++, f_sp3, 192);
        const Cfg c_mb3 = {H_MBAR, 1 << 3, 0, 0, 0};
        run<S_NONE>("core_mbarwait_warp3", out, cyc, s++, c_mb3, 192);
        const Cfg s_mb3 = {H_MBAR, 1 << 3, 1, 0, 0};
        run<S_NONE>("sync_mbarwait_warp3", out, cyc, s++, s_mb3, 192);
        const Cfg m_mb3 = {H_MBAR, 1 << 3, 0, 1, 0};
        run<S_NONE>("mbar_nofence_mbarwait_warp3", out, cyc, s++, m_mb3, 192);
    }
    if (part == 0 || part == 4) {
        const Cfg f_mb3 = {H_MBAR, 1 << 3, 1, 1, 1};
        run<S_F64>("full_mbarwait_warp3_store_f64", out, cyc, s++, f_mb3, 192);
        run<S_LL>("full_mbarwait_warp3_store_ll", out, cyc, s++, f_mb3, 192);
    }
    if (part == 0 || part == 5) {
        const Cfg f_mb3 = {H_MBAR, 1 << 3, 1, 1, 1};
        run<S_ASYNC>("full_mbarwait_warp3_store_async", out, cyc, s++, f_mb3, 192);
    }
    printf("  \"unit\": \"SM cycles per step\"\n}\n");
    return 0;
}
