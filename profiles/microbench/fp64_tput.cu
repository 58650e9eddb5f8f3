// FP64 issue throughput of one B200 SM without FMA contraction (-fmad=false is the product's
// arithmetic contract): cycles per warp-wide DMUL / DADD when 1, 2 or 4 warps (one per SM
// sub-partition) run independent chains.  Answers: is the FP64 pipe private to a sub-partition or
// shared by the SM, and how many cycles does one warp instruction occupy it?
// Build: nvcc -arch=sm_100a -fmad=false -O3 -o fp64_tput fp64_tput.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool ADD>
__global__ void k(double *out, long long *cyc, int iters, int active_warps) {
    const int warp = threadIdx.x >> 5;
    double a[CHAINS];
    for (int i = 0; i < CHAINS; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, d = 1e-12;
    __syncthreads();
    if (warp >= active_warps) return;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) a[i] = ADD ? a[i] + d : a[i] * m;
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) s += a[i];
    out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[warp] = t1 - t0;
}

template <int CHAINS, bool ADD>
static void run(const char *name, double *out, long long *cyc, int warps) {
    const int iters = 4096;
    for (int r = 0; r < 2; r++) {
        k<CHAINS, ADD><<<1, 128>>>(out, cyc, iters, warps);
        cudaDeviceSynchronize();
    }
    long long c[4];
    cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < warps; i++) mx = c[i] > mx ? c[i] : mx;
    printf("  \"%s\": %.2f,\n", name, (double)mx / ((double)iters * CHAINS));
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 64);
    printf("{\n");
    run<1, false>("dmul_1warp_1chain_cyc_per_instr", out, cyc, 1);
    run<8, false>("dmul_1warp_8chains_cyc_per_instr", out, cyc, 1);
    run<16, false>("dmul_1warp_16chains_cyc_per_instr", out, cyc, 1);
    run<16, true>("dadd_1warp_16chains_cyc_per_instr", out, cyc, 1);
    run<16, false>("dmul_2warps_16chains_cyc_per_instr_per_warp", out, cyc, 2);
    run<16, false>("dmul_4warps_16chains_cyc_per_instr_per_warp", out, cyc, 4);
    printf("  \"unit\": \"SM cycles per warp-wide FP64 instruction, as seen by each warp\"\n}\n");
    return 0;
}
