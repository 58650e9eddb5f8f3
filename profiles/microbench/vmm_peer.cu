// Feasibility probe: two processes, one GPU each, build ONE contiguous virtual range whose
// first half is backed by GPU 0 and second half by GPU 1 (cuMemCreate + POSIX fd export over
// SCM_RIGHTS + cuMemMap), then read the peer half from a kernel and measure a flag ping-pong.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x)                                                                      \
    do {                                                                           \
        CUresult r_ = (x);                                                         \
        if (r_ != CUDA_SUCCESS) {                                                  \
            const char *s_;                                                        \
            cuGetErrorString(r_, &s_);                                             \
            printf("rank %d: %s -> %s\n", g_rank, #x, s_);                         \
            exit(2);                                                               \
        }                                                                          \
    } while (0)
static int g_rank;

static void send_fd(int sock, int fd) {
    struct msghdr msg = {};
    char buf[CMSG_SPACE(sizeof(int))];
    memset(buf, 0, sizeof buf);
    char dummy = 'x';
    struct iovec io = {&dummy, 1};
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = buf;
    msg.msg_controllen = sizeof buf;
    struct cmsghdr *c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    if (sendmsg(sock, &msg, 0) < 0) { perror("sendmsg"); exit(3); }
}
static int recv_fd(int sock) {
    struct msghdr msg = {};
    char buf[CMSG_SPACE(sizeof(int))];
    char dummy;
    struct iovec io = {&dummy, 1};
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = buf;
    msg.msg_controllen = sizeof buf;
    if (recvmsg(sock, &msg, 0) < 0) { perror("recvmsg"); exit(3); }
    struct cmsghdr *c = CMSG_FIRSTHDR(&msg);
    int fd;
    memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}
static void sock_barrier(int sock) {
    char c = 'b';
    if (write(sock, &c, 1) != 1) exit(4);
    if (read(sock, &c, 1) != 1) exit(4);
}

__global__ void fill(double *p, size_t n, double v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v + (double)i;
}
__global__ void check(const double *p, size_t n, double v, int *bad) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (p[i] != v + (double)i) atomicAdd(bad, 1);
}
__global__ void sum(const double *p, size_t n, double *out) {
    double a = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += p[i];
    if (a == 123.456) *out = a;
}
// ping-pong: rank 0 writes k to peer flag, waits for own flag == k; rank 1 echoes.
__global__ void pingpong(volatile unsigned *mine, volatile unsigned *peer, int rank, int iters, long long *cycles) {
    long long t0 = clock64();
    for (int k = 1; k <= iters; k++) {
        if (rank == 0) {
            *peer = k;
            while (*mine != (unsigned)k) {}
        } else {
            while (*mine != (unsigned)k) {}
            *peer = k;
        }
    }
    *cycles = clock64() - t0;
}

int main() {
    int sv[2];
    socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
    pid_t pid = fork();
    g_rank = pid == 0 ? 1 : 0;
    const int sock = sv[g_rank];
    close(sv[1 - g_rank]);
    CK(cuInit(0));
    cudaSetDevice(g_rank);
    cudaFree(0);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = g_rank;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    const size_t half = 64u << 20;
    if (g_rank == 0) printf("granularity %zu\n", gran);
    CUmemGenericAllocationHandle h[2];
    CK(cuMemCreate(&h[g_rank], half, &prop, 0));
    int fd;
    CK(cuMemExportToShareableHandle(&fd, h[g_rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    int pfd;
    if (g_rank == 0) { send_fd(sock, fd); pfd = recv_fd(sock); } else { pfd = recv_fd(sock); send_fd(sock, fd); }
    CK(cuMemImportFromShareableHandle(&h[1 - g_rank], (void *)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    close(fd);
    close(pfd);
    CUdeviceptr va;
    CK(cuMemAddressReserve(&va, 2 * half, 0, 0, 0));
    CK(cuMemMap(va, half, 0, h[0], 0));
    CK(cuMemMap(va + half, half, 0, h[1], 0));
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = g_rank;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, 2 * half, &acc, 1));
    double *base = (double *)va;
    const size_t n = half / 8;
    fill<<<148, 256>>>(base + g_rank * n, n - 1024, 1000.0 * (g_rank + 1));
    cudaMemset(base + g_rank * n + n - 1024, 0, 8192);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("rank %d fill failed\n", g_rank); return 5; }
    sock_barrier(sock);
    int *bad;
    cudaMalloc(&bad, 4);
    cudaMemset(bad, 0, 4);
    check<<<148, 256>>>(base + (1 - g_rank) * n, n - 1024, 1000.0 * (2 - g_rank), bad);
    int hb = -1;
    cudaError_t e = cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("rank %d: peer-half check: %s, mismatches %d\n", g_rank, cudaGetErrorString(e), hb);
    // remote read bandwidth
    double *out;
    cudaMalloc(&out, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int which = 0; which < 2; which++) {
        sum<<<148 * 8, 256>>>(base + which * n, n, out);
        cudaEventRecord(a);
        for (int i = 0; i < 5; i++) sum<<<148 * 8, 256>>>(base + which * n, n, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("rank %d reads half %d: %.1f GB/s\n", g_rank, which, 5.0 * half / ms / 1e6);
    }
    sock_barrier(sock);
    // flag ping-pong through the last 8 KB of each half
    volatile unsigned *mine = (volatile unsigned *)(base + g_rank * n + n - 512);
    volatile unsigned *peer = (volatile unsigned *)(base + (1 - g_rank) * n + n - 512);
    long long *cyc;
    cudaMalloc(&cyc, 8);
    sock_barrier(sock);
    pingpong<<<1, 1>>>(mine, peer, g_rank, 1000, cyc);
    long long hc = 0;
    e = cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rank %d: ping-pong %s: %.0f cycles per round trip\n", g_rank, cudaGetErrorString(e), hc / 1000.0);
    sock_barrier(sock);
    if (g_rank == 0) { int st; wait(&st); }
    return 0;
}
