// Where do the cycles of one wavefront step go?  One warp runs the forward-substitution
// step of sweep_kernels.cu (lane t = row t, skewed one column per lane, fully unrolled
// 32-step macro-steps) with features switched on one at a time:
//   F_SHFL   upper neighbour through __shfl_up (else a register: pure 3-op chain)
//   F_SEL    lane 0 takes the hand-off value instead of the shuffle (FSEL pair)
//   F_LDS    operands a, cx, cy, precon, halo from shared memory, fetched one step ahead
//   F_STS    result stored to shared memory
//   F_ADDR   per-step address select between two stage bases (ISETP + SEL)
//   F_SYNC   progress store every 8 steps + hand-off counter poll every 8 steps
//   F_MBAR   per macro-step mbarrier try_wait + fence.proxy.async + __syncwarp + arrive
//   F_SPIN   a second warp on the same SM sub-partition spins on a global load
// Build: nvcc -arch=sm_100a -fmad=false -O3 -o step step.cu ; prints cycles per step.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { F_SHFL = 1, F_SEL = 2, F_LDS = 4, F_STS = 8, F_ADDR = 16, F_SYNC = 32, F_MBAR = 64, F_SPIN = 128, F_TWO = 256, F_PUB = 512, F_STORER = 1024, F_PUBSLEEP = 2048, F_IDLE = 4096, F_ALU = 8192, F_SELF = 16384, F_DOT = 32768 };

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

constexpr int TILE_BYTES = 33 * 32 * 8;
constexpr int NT = 5, NST = 4;

template <int F, int U>
__global__ void __launch_bounds__(160, 1) k_step(double *out, long long *cyc, int nmacro, volatile unsigned *spin, int slot) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ unsigned counters[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sm = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < NST * NT * 33 * 32; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13);
    if (threadIdx.x == 0) {
        counters[0] = 0;
        counters[1] = 1u << 30;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    }
    __syncthreads();
    if (warp == 4 && (F & F_SPIN)) { // same sub-partition as warp 0
        unsigned n = 0;
        while (*spin == 0 && n < (1u << 22)) n++;
        return;
    }
    if (warp == 3 && (F & (F_PUB | F_PUBSLEEP))) { // publisher-like: polls the progress counter in shared memory
        unsigned n = 0, seen = 0;
        while (n < (1u << 24)) {
            const unsigned v = lds_u32_volatile(smem_u32(&counters[0]));
            if (v >= (unsigned)(32 * nmacro - 40)) break;
            if (v != seen) { seen = v; out[64 + lane] = v; }
            if (F & F_PUBSLEEP) __nanosleep(100);
            n++;
        }
        return;
    }
    if (warp == 3 && (F & F_IDLE)) { // alive but asleep: no instructions, no memory traffic to speak of
        for (int i = 0; i < 40; i++) __nanosleep(1000000);
        return;
    }
    if (warp == 3 && (F & F_ALU)) { // alive, register-only integer loop on another sub-partition
        unsigned x = lane;
        for (int i = 0; i < 400000; i++) x = x * 1664525u + 1013904223u;
        out[128 + lane] = x;
        return;
    }
    if (warp == 2 && (F & F_STORER)) { // storer-like: drains a tile per macro-step (16 LDS.128 + 16 STG.128 per lane)
        unsigned n = 0, blk = 0;
        double2 *g = reinterpret_cast<double2 *>(out + 4096);
        while (n < (1u << 24)) {
            const unsigned v = lds_u32_volatile(smem_u32(&counters[0]));
            if (v >= (unsigned)(32 * nmacro - 40)) break;
            if (v >= 32 * (blk + 1)) {
                const double *tile = reinterpret_cast<const double *>(smem) + ((blk % NST) * NT + 4) * 33 * 32 + 32 + (lane & 15) * 2 + (lane >> 4) * 32;
                double2 v2[16];
#pragma unroll
                for (int i = 0; i < 16; i++) v2[i] = *reinterpret_cast<const double2 *>(tile + i * 64);
#pragma unroll
                for (int i = 0; i < 16; i++) g[(blk & 7) * 512 + i * 32 + lane] = v2[i];
                blk++;
            }
            n++;
        }
        return;
    }
    if (warp != 0) return;
    const uint32_t base0 = smem_u32(smem) + (uint32_t)((1 + lane) * 256);
    const uint32_t halo0 = smem_u32(smem) + (NST * NT - 1) * TILE_BYTES;
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    double z = 0.5 + lane * 1e-3, c1 = 1e-3;
    double a = 1.0, cx = 1e-3, cy = 2e-3, pr = 0.999, halo = 0.25;
    uint4 pre = make_uint4(0, 0, 0, 0);
    double dot = 0.0;
    const long long t0 = clock64();
    for (int m = 0; m < nmacro; m++) {
        const uint32_t sA = base0 + (uint32_t)((m % NST) * NT * TILE_BYTES) - (uint32_t)(8 * lane);
        const uint32_t sB = base0 + (uint32_t)(((m + NST - 1) % NST) * NT * TILE_BYTES) + 256u - (uint32_t)(8 * lane);
#pragma unroll 1
        for (int k0 = 0; k0 < 32; k0 += U)
#pragma unroll
        for (int ku = 0; ku < U; ku++) {
            const int kk = k0 + ku;
            if ((F & F_SYNC) && ((ku + 1) % 8) == 0) {
                unsigned n = 0;
                while (lds_u32_volatile(halo_cols_addr) < (unsigned)(32 * m + kk + 9) && ++n < 1000) {}
            }
            if ((F & F_SELF) && (kk % 8) == 0) {
                // the compute warp is its own publisher and poller: every 8 steps lanes 0..7 send the
                // 8 columns the last row has just finished as LL messages and pick up the 8 messages
                // they asked for 8 steps ago (validated by epoch, spun on only if missing)
                uint4 *llo = reinterpret_cast<uint4 *>(out + 8192) + ((32 * m + kk) & 1023);
                const uint4 *lli = reinterpret_cast<const uint4 *>(out + 65536) + ((32 * m + kk) & 1023);
                if (lane < 8) {
                    const double v = lds_f64(base0 - (uint32_t)((1 + lane) * 256) + 32 * 256 + 4 * TILE_BYTES + (uint32_t)(8 * ((kk + lane) & 31)));
                    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
                    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(llo + lane), "r"(lo), "r"(1u), "r"(hi), "r"(1u) : "memory");
                }
                // consume the prefetched messages
                bool ok = lane >= 8 || (pre.y == 0u && pre.w == 0u);
                unsigned nn = 0;
                while (!__all_sync(0xffffffffu, ok) && ++nn < 100) {
                    if (!ok) {
                        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pre.x), "=r"(pre.y), "=r"(pre.z), "=r"(pre.w) : "l"(lli + lane) : "memory");
                        ok = pre.y == 0u && pre.w == 0u;
                    }
                }
                if (lane < 8) sts_f64(halo0 + (uint32_t)(8 * ((kk + 8 + lane) & 31)), __hiloint2double((int)pre.z, (int)pre.x));
                // ask for the next group
                if (lane < 8)
                    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(pre.x), "=r"(pre.y), "=r"(pre.z), "=r"(pre.w) : "l"(lli + 8 + lane) : "memory");
            }
            double up = (F & F_SHFL) ? __shfl_up_sync(0xffffffffu, z, 1) : z;
            // operands of the next step
            double na = a, ncx = cx, ncy = cy, npr = pr, nh = halo;
            if (F & F_LDS) {
                uint32_t b = sA;
                if (F & F_ADDR) b = (lane > kk + 1) ? sB : sA;
                const uint32_t p = b + (uint32_t)(8 * (kk + 1));
                na = lds_f64(p);
                ncx = lds_f64(p + TILE_BYTES);
                ncy = lds_f64(p + 2 * TILE_BYTES - 256);
                npr = lds_f64(p + 3 * TILE_BYTES);
                nh = lds_f64(halo0 + (uint32_t)(8 * ((kk + 1) & 31)));
            }
            if (F & F_SEL) up = sel_f64(lane == 0, halo, up);
            double t = a - c1 * z;
            t = t - cy * up;
            z = t * pr;
            c1 = cx;
            if (F & F_DOT) dot += z * nh; // z.r folded by the compute warp itself (one more operand per step)
            if (F & F_STS) {
                uint32_t b = sA;
                if (F & F_ADDR) b = (lane > kk) ? sB : sA;
                sts_f64(b + (uint32_t)(8 * kk) + 4 * TILE_BYTES, z);
            }
            if ((F & F_SYNC) && ((ku + 2) % 8) == 0) sts_u32_volatile(progress_addr, (unsigned)(32 * m + kk));
            a = na; cx = ncx; cy = ncy; pr = npr; halo = nh;
        }
        if (F & F_MBAR) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
            unsigned ok = 0, n = 0;
            while (!ok && ++n < 1000) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"((unsigned)(m & 1)) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[slot] = t1 - t0;
    out[lane] = z + c1 + dot;
}

template <int F, int U = 32>
static void run(const char *name, double *out, long long *cyc, unsigned *spin, int slot) {
    const int nmacro = 128;
    const size_t smem = (size_t)NST * NT * TILE_BYTES;
    cudaFuncSetAttribute(k_step<F, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(spin, 0, 4);
        k_step<F, U><<<1, 160, smem>>>(out, cyc, nmacro, spin, slot);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    long long c;
    cudaMemcpy(&c, cyc + slot, 8, cudaMemcpyDeviceToHost);
    printf("  \"%s\": %.1f,\n", name, (double)c / (nmacro * 32));
}

// Two interleaved sub-strips in ONE warp (rows 0..31 and 32..63 of a 64-row strip): the
// second recurrence is independent of the first within a step (it consumes what the first
// produced 33 steps earlier, through a shared-memory ring), so its instructions fill the
// issue slots the first chain leaves empty.
constexpr int NST2 = 3;
template <int F>
__global__ void __launch_bounds__(192, 1) k_step2(double *out, long long *cyc, int nmacro, int slot) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ unsigned counters[4];
    __shared__ double ring[128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sm = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < 2 * NST2 * 4 * 33 * 32; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13);
    if (threadIdx.x < 128) ring[threadIdx.x] = 0.25;
    if (threadIdx.x == 0) {
        counters[0] = 0; counters[1] = 1u << 30; counters[2] = 0; counters[3] = 1u << 30;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    }
    __syncthreads();
    if (warp == 3 && (F & F_IDLE)) { for (int i = 0; i < 40; i++) __nanosleep(1000000); return; }
    if (warp != 0) return;
    constexpr int T4 = 4 * TILE_BYTES;
    const uint32_t baseA = smem_u32(smem) + (uint32_t)((1 + lane) * 256);
    const uint32_t baseB = baseA + NST2 * T4;
    const uint32_t halo0 = smem_u32(ring);
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    double zA = 0.5 + lane * 1e-3, c1A = 1e-3, zB = 0.4 + lane * 1e-3, c1B = 1e-3;
    double aA = 1.0, cxA = 1e-3, cyA = 2e-3, prA = 0.999, hA = 0.25;
    double aB = 1.0, cxB = 1e-3, cyB = 2e-3, prB = 0.999, hB = 0.25;
    const long long t0 = clock64();
    for (int m = 0; m < nmacro; m++) {
        const uint32_t sA = baseA + (uint32_t)((m % NST2) * T4) - (uint32_t)(8 * lane);
        const uint32_t sA2 = baseA + (uint32_t)(((m + NST2 - 1) % NST2) * T4) + 256u - (uint32_t)(8 * lane);
        const uint32_t sB = baseB + (uint32_t)((m % NST2) * T4) - (uint32_t)(8 * lane);
        const uint32_t sB2 = baseB + (uint32_t)(((m + NST2 - 1) % NST2) * T4) + 256u - (uint32_t)(8 * lane);
#pragma unroll
        for (int kk = 0; kk < 32; kk++) {
            if ((F & F_SYNC) && ((kk + 1) % 8) == 0) {
                unsigned n = 0;
                while (lds_u32_volatile(halo_cols_addr) < (unsigned)(32 * m + kk + 9) && ++n < 1000) {}
            }
            double upA = __shfl_up_sync(0xffffffffu, zA, 1);
            double upB = __shfl_up_sync(0xffffffffu, zB, 1);
            const uint32_t bA = (lane > kk + 1) ? sA2 : sA, bB = (lane > kk + 1) ? sB2 : sB;
            const uint32_t pA = bA + (uint32_t)(8 * (kk + 1)), pB = bB + (uint32_t)(8 * (kk + 1));
            const double naA = lds_f64(pA), ncxA = lds_f64(pA + TILE_BYTES), ncyA = lds_f64(pA + 2 * TILE_BYTES - 256), nprA = lds_f64(pA + 3 * TILE_BYTES);
            const double naB = lds_f64(pB), ncxB = lds_f64(pB + TILE_BYTES), ncyB = lds_f64(pB + 2 * TILE_BYTES - 256), nprB = lds_f64(pB + 3 * TILE_BYTES);
            const double nhA = lds_f64(halo0 + (uint32_t)(8 * ((kk + 1) & 31)) + 512);
            const double nhB = lds_f64(halo0 + (uint32_t)(8 * ((kk + 1) & 31)));   // what sub-strip A's last row wrote 33 steps ago
            upA = sel_f64(lane == 0, hA, upA);
            upB = sel_f64(lane == 0, hB, upB);
            double tA = aA - c1A * zA; tA = tA - cyA * upA; zA = tA * prA; c1A = cxA;
            double tB = aB - c1B * zB; tB = tB - cyB * upB; zB = tB * prB; c1B = cxB;
            const uint32_t qA = ((lane > kk) ? sA2 : sA) + (uint32_t)(8 * kk), qB = ((lane > kk) ? sB2 : sB) + (uint32_t)(8 * kk);
            sts_f64(qA, zA);
            sts_f64(qB, zB);
            if (lane == 31) sts_f64(halo0 + (uint32_t)(8 * ((kk + 1) & 31)), zA); // A's last row feeds B's first
            if ((F & F_SYNC) && ((kk + 2) % 8) == 0) sts_u32_volatile(progress_addr, (unsigned)(32 * m + kk));
            aA = naA; cxA = ncxA; cyA = ncyA; prA = nprA; hA = nhA;
            aB = naB; cxB = ncxB; cyB = ncyB; prB = nprB; hB = nhB;
        }
        if (F & F_MBAR) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
            unsigned ok = 0, n = 0;
            while (!ok && ++n < 1000) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"((unsigned)(m & 1)) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[slot] = t1 - t0;
    out[lane] = zA + c1A + zB + c1B;
}

template <int F>
static void run2(const char *name, double *out, long long *cyc, int slot) {
    const int nmacro = 128;
    const size_t smem = (size_t)2 * NST2 * 4 * TILE_BYTES;
    cudaFuncSetAttribute(k_step2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) {
        k_step2<F><<<1, 192, smem>>>(out, cyc, nmacro, slot);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    long long c;
    cudaMemcpy(&c, cyc + slot, 8, cudaMemcpyDeviceToHost);
    printf("  \"%s\": %.1f,\n", name, (double)c / (nmacro * 32));
}

int main() {
    double *out; long long *cyc; unsigned *spin;
    cudaMalloc(&out, 1 << 20); cudaMemset(out, 0, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&spin, 4);
    printf("{\n");
    run<0>("chain_3op_register", out, cyc, spin, 0);
    run<F_SHFL>("shfl", out, cyc, spin, 1);
    run<F_SHFL | F_SEL>("shfl_sel", out, cyc, spin, 2);
    run<F_SHFL | F_SEL | F_LDS>("shfl_sel_lds", out, cyc, spin, 3);
    run<F_SHFL | F_SEL | F_LDS | F_STS>("shfl_sel_lds_sts", out, cyc, spin, 4);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR>("shfl_sel_lds_sts_addr", out, cyc, spin, 5);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC>("plus_sync", out, cyc, spin, 6);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR>("plus_sync_mbar", out, cyc, spin, 7);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_SPIN>("plus_sync_mbar_spin", out, cyc, spin, 8);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_PUB>("plus_sync_mbar_publisher", out, cyc, spin, 11);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_PUBSLEEP>("plus_sync_mbar_publisher_sleep100", out, cyc, spin, 12);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_STORER>("plus_sync_mbar_storer", out, cyc, spin, 13);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_STORER | F_PUB>("plus_sync_mbar_storer_publisher", out, cyc, spin, 14);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_IDLE>("core_plus_idle_warp", out, cyc, spin, 15);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_ALU>("core_plus_alu_warp", out, cyc, spin, 16);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_PUB>("core_plus_publisher", out, cyc, spin, 17);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_IDLE>("plus_sync_mbar_idle_warp", out, cyc, spin, 18);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_IDLE>("plus_sync_idle_warp", out, cyc, spin, 19);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_MBAR | F_IDLE>("plus_mbar_idle_warp", out, cyc, spin, 20);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR, 8>("core_unroll8", out, cyc, spin, 21);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_IDLE, 8>("core_unroll8_idle_warp", out, cyc, spin, 22);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_IDLE, 8>("plus_sync_mbar_unroll8_idle_warp", out, cyc, spin, 23);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_IDLE, 16>("plus_sync_mbar_unroll16_idle_warp", out, cyc, spin, 24);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SYNC | F_MBAR | F_STORER | F_PUB, 8>("plus_sync_mbar_unroll8_storer_publisher", out, cyc, spin, 25);
    run2<0>("two_substrips_core", out, cyc, 30);
    run2<F_IDLE>("two_substrips_core_idle_warp", out, cyc, 31);
    run2<F_SYNC | F_MBAR | F_IDLE>("two_substrips_sync_mbar_idle_warp", out, cyc, 32);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SELF>("single_warp_self_publish_poll", out, cyc, spin, 40);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SELF | F_MBAR>("single_warp_self_publish_poll_mbar", out, cyc, spin, 41);
    run<F_SHFL | F_SEL | F_LDS | F_STS | F_ADDR | F_SELF | F_MBAR | F_DOT>("single_warp_self_publish_poll_mbar_dot", out, cyc, spin, 42);
    run<F_SHFL | F_LDS | F_STS>("shfl_lds_sts_nosel", out, cyc, spin, 9);
    run<F_LDS | F_STS>("lds_sts_noshfl", out, cyc, spin, 10);
    printf("  \"unit\": \"SM cycles per step\"\n}\n");
    return 0;
}
