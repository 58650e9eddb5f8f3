// Latency microbenchmarks that size the wavefront kernels (sweep_kernels.cu):
// dependent FP64 DMUL/DADD chain, warp shuffle, shared-memory load, and the
// SM-to-SM "flag in data" hand-off through L2.  Build: nvcc -arch=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_chain(double *out, long long *cyc, int n, double a, double b) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { // mul, sub, sub, mul == one sweep cell
        double t = 1.0 - b * x;
        t = t - b * 0.5;
        x = t * a;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; cyc[0] = t1 - t0; }
}
__global__ void k_dadd(double *out, long long *cyc, int n, double a) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = x + a;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; cyc[1] = t1 - t0; }
}
__global__ void k_dmul(double *out, long long *cyc, int n, double a) {
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = x * a;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; cyc[2] = t1 - t0; }
}
__global__ void k_shfl(double *out, long long *cyc, int n, double a) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) x = __shfl_up_sync(0xffffffffu, x, 1);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; cyc[3] = t1 - t0; }
}
__global__ void k_lds(double *out, long long *cyc, int n) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i + 33) & 1023;
    __syncwarp();
    int j = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) j = idx[j];
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = j; cyc[4] = t1 - t0; }
}
// ping-pong between block 0 and block 1 (different SMs) with 16-byte {lo,ep,hi,ep} messages
__global__ void k_pingpong(uint4 *box, long long *cyc, int n) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    volatile uint4 *mine = box + me * 32, *peer = box + (1 - me) * 32;
    long long t0 = clock64();
    for (unsigned i = 1; i <= (unsigned)n; i++) {
        if (me == 0) {
            asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(peer), "r"(i), "r"(i), "r"(i), "r"(i) : "memory");
            unsigned a, b, c, d;
            do { asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(mine) : "memory"); } while (b != i || d != i);
        } else {
            unsigned a, b, c, d;
            do { asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(mine) : "memory"); } while (b != i || d != i);
            asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(peer), "r"(i), "r"(i), "r"(i), "r"(i) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) cyc[5] = t1 - t0;
}
int main() {
    double *out; long long *cyc; uint4 *box;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&box, 64 * sizeof(uint4));
    cudaMemset(box, 0, 64 * sizeof(uint4)); cudaMemset(cyc, 0, 64);
    const int n = 1 << 16;
    for (int rep = 0; rep < 2; rep++) {
        k_chain<<<1, 32>>>(out, cyc, n, 0.999, 0.001);
        k_dadd<<<1, 32>>>(out, cyc, n, 1e-9);
        k_dmul<<<1, 32>>>(out, cyc, n, 1.0000001);
        k_shfl<<<1, 32>>>(out, cyc, n, 1.0);
        k_lds<<<1, 32>>>(out, cyc, n);
        k_pingpong<<<2, 32>>>(box, cyc, 2000);
        cudaMemset(box, 0, 64 * sizeof(uint4));
    }
    long long h[8];
    cudaError_t e = cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"cuda\": \"%s\", \"clock_khz\": %d, \"cell_chain_cyc\": %.2f, \"dadd_cyc\": %.2f, \"dmul_cyc\": %.2f, "
           "\"shfl64_cyc\": %.2f, \"lds_cyc\": %.2f, \"l2_pingpong_roundtrip_cyc\": %.1f}\n",
           cudaGetErrorString(e), clk, (double)h[0] / n, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n,
           (double)h[4] / n, (double)h[5] / 2000);
    return 0;
}
