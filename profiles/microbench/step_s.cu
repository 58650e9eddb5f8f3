// step_s.cu -- the staircase engine's steady-state macro-step (stair_kernels.cu, EDGE 0) on one warp with
// synthetic operands in shared memory and no helper warps: SM cycles per step, feature by feature
// (STAIR_EXP bits, see stair_kernels.cu).  profiles/microbench/step_s.sh builds and runs the variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -DSTAIR_EXP=<bits> -DIFL_STAIR_DEVICE_ONLY
//             -I ../../incremental-fluids_b200/csrc -I ../../include -o step_s step_s.cu
#include "stair_kernels.cu"

#include <cstdio>

using namespace ifl;
using namespace ifl::stair;

template <bool BWD>
__global__ void __launch_bounds__(32, 1) k_step(double *out, long long *cyc, int nmacro) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bell[NBELL];
    __shared__ unsigned counters[2];
    __shared__ int dead;
    typedef Geo<BWD> G;
    const int lane = threadIdx.x;
    double *sm = reinterpret_cast<double *>(smem);
    for (int i = lane; i < (NST * STAGE_BYTES + HRC * 8) / 8; i += 32) sm[i] = 1e-3 * ((i * 7) % 13) + 0.5;
    if (lane == 0) {
        counters[0] = 0;
        counters[1] = 1u << 30;
        dead = 0;
        for (int i = 0; i < NBELL; i++) mbar_init(&bell[i], 1);
    }
    __syncwarp();
    State s;
    s.zA = 0.5 + lane * 1e-3;
    s.zB = 0.4 + lane * 1e-3;
    s.cA = s.cB = 1e-3;
    s.kA = s.kB = 0.0;
    const int u = lane & 7, g = lane >> 3, rA = 2 * u;
    const uint32_t row0 = smem_u32(smem) + (uint32_t)(g * GT_BYTES + G::row_a(u) * ROWB + G::COL0 * 8);
    const uint32_t halo0 = smem_u32(smem + NST * STAGE_BYTES);
    auto bases_of = [&](int sA, int sB, int sN, uint32_t row, int r) {
        LaneBases lb;
        lb.A = row + (uint32_t)(sA * STAGE_BYTES) - (uint32_t)(G::DIR * r);
        lb.B = row + (uint32_t)(sB * STAGE_BYTES) + (uint32_t)(G::DIR * (BW - r));
        lb.N = row + (uint32_t)(sN * STAGE_BYTES) - (uint32_t)(G::DIR * (BW + r));
        return lb;
    };
    {
        const LaneBases la = bases_of(0, NST - 1, 1, row0, rA);
        const LaneBases lbb = bases_of(0, NST - 1, 1, row0 + (uint32_t)G::ROW_B, rA + 1);
        fetch<BWD>(s.qA, pos<BWD>(la, 0, rA) + (uint32_t)G::PAIR);
        fetch<BWD>(s.qB, pos<BWD>(lbb, -1, rA + 1) + (uint32_t)G::PAIR);
        fetch<BWD>(s.qB_n, pos<BWD>(lbb, 1, rA + 1) + (uint32_t)G::PAIR);
        s.h = lds_pair(halo0);
        s.qA_n = s.qA;
        s.h_n = s.h;
    }
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    int sm_ = 0;
    const long long t0 = clock64();
    for (int m = 4; m < nmacro + 4; m++) {
        const int a = sm_, b = a == 0 ? NST - 1 : a - 1, n = a == NST - 1 ? 0 : a + 1;
        const LaneBases la = bases_of(a, b, n, row0, rA);
        const LaneBases lbb = bases_of(a, b, n, row0 + (uint32_t)G::ROW_B, rA + 1);
        const uint32_t h_cur = halo0 + (uint32_t)(((BW * m) & (HRC - 1)) * 8);
        const uint32_t h_next = halo0 + (uint32_t)(((BW * (m + 1)) & (HRC - 1)) * 8);
        const uint32_t bell6 = smem_u32(&bell[(2 * m + NBELL - 8) & (NBELL - 1)]), bell14 = smem_u32(&bell[(2 * m + NBELL - 7) & (NBELL - 1)]);
        macro_step<BWD, 0>(la, lbb, h_cur, h_next, bell6, bell14, 0u, m, lane, rA, s, progress_addr, gate_addr, 1 << 28, 1 << 28, &dead, nullptr);
        sm_ = sm_ == NST - 1 ? 0 : sm_ + 1;
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = s.zA + s.zB + s.cA + s.cB + s.kA + s.kB;
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 64);
    const int nmacro = 256;
    const size_t smem = (size_t)NST * STAGE_BYTES + HRC * 8;
    for (int bwd = 0; bwd < 2; bwd++) {
        auto kern = bwd ? k_step<true> : k_step<false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        long long c = 0;
        for (int rep = 0; rep < 2; rep++) {
            kern<<<1, 32, smem>>>(out, cyc, nmacro);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("error: %s\n", cudaGetErrorString(e));
                return 1;
            }
        }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("STAIR_EXP=%d %s: %.1f cycles/step\n", STAIR_EXP, bwd ? "bwd" : "fwd", (double)c / (nmacro * BW));
    }
    return 0;
}
