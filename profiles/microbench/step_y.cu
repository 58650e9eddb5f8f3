// Prototype of the round-2 wavefront step ("Y" geometry): one warp sweeps a strip of 64 rows.
//   * lane t owns rows 2t (cell A) and 2t+1 (cell B) and runs 3 columns behind lane t-1;
//     B runs one column behind A.  A's upper neighbour is lane t-1's B of TWO steps ago, so the
//     shuffle that carries it is issued a whole step before its result is needed: the only
//     dependency carried from one step to the next is the in-register chain mul-sub-sub-mul.
//   * operands live in shared memory as [stage][operand][group of 4 lanes = 8 rows][16 columns],
//     boxes aligned to absolute multiples of 16 columns: lane t reads column (k - 3t) mod 16, which
//     makes the 16 lanes of a half-warp hit 16 different 8-byte bank slots.
// Features are switched on one at a time to see what each costs (SM cycles per step).
// Build: nvcc -arch=sm_100a -fmad=false -O3 -o step_y step_y.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { F_SYNC = 1, F_MBAR = 2, F_IDLE = 4, F_STORER = 8, F_NOSHFL = 16, F_STASYNC = 32, F_STLL = 64, F_ARRIVE = 128, F_NOSEL = 256 };

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

constexpr int BW = 16;               // columns per block
constexpr int NOP = 4;               // operand tiles per stage
constexpr int GT = 8 * BW * 8;       // bytes of one group tile (8 rows x 16 columns)
constexpr int OPB = 8 * GT;          // bytes of one operand tile (8 groups)
constexpr int STB = NOP * OPB;       // bytes per stage (32 KB)
constexpr int NST = 6;

struct Ops {
    double aA, cxA, cyA, pA, aB, cxB, cyB, pB, halo;
};

template <int F>
__global__ void __launch_bounds__(160, 1) k_y(double *out, long long *cyc, int nmacro, int slot) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ unsigned counters[4];
    __shared__ __align__(16) double halo_ring[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sm = reinterpret_cast<double *>(smem);
    for (int i = threadIdx.x; i < NST * STB / 8; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) halo_ring[i] = 0.25;
    if (threadIdx.x == 0) {
        counters[0] = 0;
        counters[1] = 1u << 30;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[2])));
    }
    __syncthreads();
    if (warp == 3 && (F & F_IDLE)) {
        for (int i = 0; i < 40; i++) __nanosleep(1000000);
        return;
    }
    if (warp == 2 && (F & F_STORER)) { // storer-like: drains one 64x16 tile per macro-step
        unsigned n = 0, blk = 0;
        double2 *g = reinterpret_cast<double2 *>(out + 4096);
        while (n < (1u << 24)) {
            const unsigned v = lds_u32_volatile(smem_u32(&counters[0]));
            if (v >= (unsigned)(BW * nmacro - 40)) break;
            if (v >= BW * (blk + 1)) {
                const double *tile = reinterpret_cast<const double *>(smem + (blk % NST) * STB) + lane * 2;
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
    const int g = lane >> 2, l = lane & 3;
    const int kt = (3 * lane) % BW;           // step of the macro-step at which this lane enters a new block
    const int lag = (3 * lane + BW - 1) / BW; // ceil(3t/16): blocks this lane is behind lane 0
    const uint32_t rowoff = smem_u32(smem) + (uint32_t)(g * GT + (2 * l) * BW * 8);
    const uint32_t halo0 = smem_u32(halo_ring);
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    const uint32_t r_halo = mapa(halo0, 0), r_bar = mapa(smem_u32(&bar[2]), 0);
    double zA = 0.5 + lane * 1e-3, zB = 0.4 + lane * 1e-3, c1A = 1e-3, c1B = 1e-3, upA = 0.3;
    Ops o;
    o.aA = o.aB = 1.0;
    o.cxA = o.cxB = 1e-3;
    o.cyA = o.cyB = 2e-3;
    o.pA = o.pB = 0.999;
    o.halo = 0.25;
    uint32_t pcur = rowoff, pprev = rowoff;
    unsigned gate_seen = 0;
    const long long t0 = clock64();
    for (int m = 0; m < nmacro; m++) {
        // block of this lane before / after its transition, and the one lanes 0/16 step into at the look-ahead
        const int b0 = m + NST * 4 - lag;
        const uint32_t before = rowoff + (uint32_t)(((b0) % NST) * STB) + (uint32_t)(BW * 8) - (uint32_t)(8 * kt);
        const uint32_t after = rowoff + (uint32_t)(((b0 + 1) % NST) * STB) - (uint32_t)(8 * kt);
        const uint32_t next = rowoff + (uint32_t)(((b0 + 2) % NST) * STB) - (uint32_t)(8 * BW);
        const uint32_t hbase = halo0 + (uint32_t)((m & 15) * BW * 8);
#pragma unroll
        for (int kk = 0; kk < BW; kk++) {
            if (F & F_SYNC) {
                if (((kk + 5) % 8) == 0) gate_seen = lds_u32_volatile(gate_addr);
                if (((kk + 1) % 8) == 0) {
                    const unsigned need = (unsigned)(BW * m + kk + 9);
                    unsigned n = 0;
                    if (gate_seen < need)
                        while (lds_u32_volatile(gate_addr) < need && ++n < 1000) {}
                }
            }
            // the upper neighbour of NEXT step's A: lane t-1's B as it stands now
            const double shf = (F & F_NOSHFL) ? zB : __shfl_up_sync(0xffffffffu, zB, 1);
            // operands of the next step
            Ops nx;
            const int j = kk + 1;
            uint32_t pn;
            if (j < BW)
                pn = ((kt <= j) ? after : before) + (uint32_t)(8 * j);
            else
                pn = ((kt == 0) ? next : after) + (uint32_t)(8 * j);
            nx.aA = lds_f64(pn);
            nx.cxA = lds_f64(pn + OPB);
            nx.cyA = lds_f64(pn + 2 * OPB);
            nx.pA = lds_f64(pn + 3 * OPB);
            const uint32_t pnb = pcur + BW * 8; // B of the next step: A's column of this step, one row down
            nx.aB = lds_f64(pnb);
            nx.cxB = lds_f64(pnb + OPB);
            nx.cyB = lds_f64(pnb + 2 * OPB);
            nx.pB = lds_f64(pnb + 3 * OPB);
            nx.halo = lds_f64(hbase + (uint32_t)(8 * j));
            // this step
            double tA = o.aA - c1A * zA;
            tA = tA - o.cyA * upA;
            const double nzA = tA * o.pA;
            double tB = o.aB - c1B * zB;
            tB = tB - o.cyB * zA;
            const double nzB = tB * o.pB;
            sts_f64(pcur, nzA);
            sts_f64(pprev + BW * 8, nzB);
            if (F & F_STASYNC) { // lane 31 hands B to the downstream strip: 8 bytes + complete_tx on its mbarrier
                if (lane == 31)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                                     r_halo + 2048u + (uint32_t)(8 * kk)),
                                 "l"(__double_as_longlong(nzB)), "r"(r_bar)
                                 : "memory");
            }
            if (F & F_STLL) { // same, as a flag-in-data message {lo, tag, hi, tag}
                if (lane == 31) {
                    const unsigned lo = (unsigned)__double2loint(nzB), hi = (unsigned)__double2hiint(nzB), tag = (unsigned)(m + 1);
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r_halo + 2048u + (uint32_t)(16 * kk)), "r"(lo),
                                 "r"(tag), "r"(hi), "r"(tag)
                                 : "memory");
                }
            }
            c1A = o.cxA;
            c1B = o.cxB;
            zA = nzA;
            zB = nzB;
            upA = (F & F_NOSEL) ? shf : sel_f64(lane == 0, nx.halo, shf);
            if ((F & F_SYNC) && ((kk + 2) % 8) == 0) sts_u32_volatile(progress_addr, (unsigned)(BW * m + kk));
            o = nx;
            pprev = pcur;
            pcur = pn;
        }
        if (F & F_ARRIVE) { // release the finished stage: non-blocking
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
        }
        if (F & F_MBAR) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
            unsigned ok = 0, n = 0;
            while (!ok && ++n < 1000) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok)
                             : "r"(smem_u32(&bar[0])), "r"((unsigned)(m & 1))
                             : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[slot] = t1 - t0;
    out[lane] = zA + c1A + zB + c1B + upA;
}

template <int F>
static void run(const char *name, double *out, long long *cyc, int slot) {
    const int nmacro = 256;
    const size_t smem = (size_t)NST * STB;
    cudaFuncSetAttribute(k_y<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) {
        k_y<F><<<1, 160, smem>>>(out, cyc, nmacro, slot);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%s: %s\n", name, cudaGetErrorString(e));
            return;
        }
    }
    long long c;
    cudaMemcpy(&c, cyc + slot, 8, cudaMemcpyDeviceToHost);
    printf("  \"%s\": %.1f,\n", name, (double)c / (nmacro * BW));
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMemset(out, 0, 1 << 20);
    cudaMalloc(&cyc, 1024);
    printf("{\n");
    run<F_NOSHFL | F_NOSEL>("y_core_noshfl", out, cyc, 0);
    run<F_NOSEL>("y_core_nosel", out, cyc, 1);
    run<0>("y_core", out, cyc, 2);
    run<F_IDLE>("y_core_idle", out, cyc, 3);
    run<F_SYNC | F_IDLE>("y_sync_idle", out, cyc, 4);
    run<F_SYNC | F_ARRIVE | F_IDLE>("y_sync_arrive_idle", out, cyc, 5);
    run<F_SYNC | F_MBAR | F_IDLE>("y_sync_mbar_idle", out, cyc, 6);
    run<F_SYNC | F_ARRIVE | F_STORER | F_IDLE>("y_sync_arrive_storer_idle", out, cyc, 7);
    run<F_SYNC | F_ARRIVE | F_STASYNC | F_IDLE>("y_sync_arrive_stasync_idle", out, cyc, 8);
    run<F_SYNC | F_ARRIVE | F_STLL | F_IDLE>("y_sync_arrive_stll_idle", out, cyc, 9);
    run<F_SYNC | F_ARRIVE | F_STLL | F_STORER | F_IDLE>("y_sync_arrive_stll_storer_idle", out, cyc, 10);
    printf("  \"unit\": \"SM cycles per step (2 cells per lane per step)\"\n}\n");
    return 0;
}
