// dist.cu -- row-slab multi-GPU plumbing: one process per GPU, ONE address space.
//
// The grid is split into slabs of whole 64-row strips.  Instead of exchanging
// halo rows, every array lives in one virtual address range that all ranks map identically:
// the pages holding a slab are physical HBM of the slab's owner (cuMemCreate), exported as
// POSIX file descriptors, passed between the processes over a UNIX socket (SCM_RIGHTS) and
// mapped by every rank (cuMemMap).  A kernel then addresses the whole array exactly as on
// one GPU; it is launched over the rank's own rows only, and the few accesses that cross a
// slab boundary (stencil rows +-1, the MIC(0) upstream row, back-traced advection samples)
// become NVLink peer loads served by NVSwitch.  Cross-rank ordering is explicit:
//   * dist_barrier / dist_barrier_block: epoch flags stored into every peer's HBM;
//   * the wavefront sweeps keep their LL message hand-off, the last strip of a slab simply
//     publishes into the downstream rank's hand-off array (sweep_kernels.cu);
//   * reductions write per-block partials into one shared array and every rank folds ALL
//     partials in the same fixed order, so the solve scalars -- and with them every field --
//     are bit-identical on all ranks and identical to the one-GPU run (pcg_kernels.cu).
// Nothing here needs NCCL: the path has no bulk exchange step, only flags, 8-byte partials
// and one row of hand-off messages per slab boundary.
#include "ifl_internal.cuh"

#include <cuda.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <vector>

namespace ifl {

// ------------------------------------------------------------------ slab plan ----
// Strips of 64 rows (the unit of the triangular-solve engine, tri_kernels.cu; two strips of the
// one-row engine) are dealt to the ranks in contiguous runs: rank g owns strips
// [floor(S*g/G), floor(S*(g+1)/G)).  Rows of the wider arrays (v and phi have h+1 rows)
// follow the cell rows; everything past the last boundary belongs to the last rank.
static int slab_first_strip(int h, int world, int g) {
    const long long strips = (h + 63) / 64;
    return (int)(strips * g / world);
}
static int slab_row0(int h, int world, int g) { return g >= world ? (1 << 30) : 64 * slab_first_strip(h, world, g); }
static int rank_of_row(int h, int world, long long row) {
    int g = 0;
    while (g + 1 < world && row >= slab_row0(h, world, g + 1)) g++;
    return g;
}

// --------------------------------------------------------- driver entry points ----
struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
    CUresult (*MemUnmap)(CUdeviceptr, size_t);
    CUresult (*MemRelease)(CUmemGenericAllocationHandle);
    CUresult (*MemAddressFree)(CUdeviceptr, size_t);
    CUresult (*MemGetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
    bool ok;
};

static bool drv_sym(const char *name, void **fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
}

static Drv *drv() {
    static Drv d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        d.ok = drv_sym("cuMemCreate", (void **)&d.MemCreate) && drv_sym("cuMemExportToShareableHandle", (void **)&d.MemExport) &&
               drv_sym("cuMemImportFromShareableHandle", (void **)&d.MemImport) &&
               drv_sym("cuMemAddressReserve", (void **)&d.MemAddressReserve) && drv_sym("cuMemMap", (void **)&d.MemMap) &&
               drv_sym("cuMemSetAccess", (void **)&d.MemSetAccess) && drv_sym("cuMemUnmap", (void **)&d.MemUnmap) &&
               drv_sym("cuMemRelease", (void **)&d.MemRelease) && drv_sym("cuMemAddressFree", (void **)&d.MemAddressFree) &&
               drv_sym("cuMemGetAllocationGranularity", (void **)&d.MemGetGranularity);
    }
    return &d;
}

#define IFL_DRV(call)                                                                   \
    do {                                                                                \
        CUresult r_ = (call);                                                           \
        if (r_ != CUDA_SUCCESS) {                                                       \
            set_error("%s:%d %s -> CUresult %d", __FILE__, __LINE__, #call, (int)r_);   \
            return IFL_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

// ------------------------------------------------------------------ rendezvous ----
// Star over one UNIX stream socket per non-root rank: rank 0 binds `path`, the others
// connect (retrying while rank 0 is not there yet) and introduce themselves by rank.
struct Mapped {
    void *p;
    size_t size; // 0: plain cudaMalloc
};

struct DistState {
    int rank, world;
    int socks[MAX_WORLD]; // rank 0: socket to rank i; other ranks: socks[0] = socket to rank 0
    int listen_fd;
    char path[104];
    size_t gran;
    std::vector<Mapped> mem;
};

static const int RDV_TIMEOUT_S = 120;

static int sock_set_timeout(int fd) {
    struct timeval tv;
    tv.tv_sec = RDV_TIMEOUT_S;
    tv.tv_usec = 0;
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
    return 0;
}

static bool write_all(int fd, const void *buf, size_t n) {
    const char *p = (const char *)buf;
    while (n > 0) {
        ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}

static bool read_all(int fd, void *buf, size_t n) {
    char *p = (char *)buf;
    while (n > 0) {
        ssize_t k = recv(fd, p, n, 0);
        if (k < 0 && errno == EINTR) continue;
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}

// One message = 4 bytes (`has`), carrying a file descriptor as ancillary data when has != 0.
static bool send_fd_msg(int sock, int fd) {
    int has = fd >= 0 ? 1 : 0;
    struct msghdr msg;
    memset(&msg, 0, sizeof msg);
    struct iovec io;
    io.iov_base = &has;
    io.iov_len = sizeof has;
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    char ctl[CMSG_SPACE(sizeof(int))];
    if (has) {
        memset(ctl, 0, sizeof ctl);
        msg.msg_control = ctl;
        msg.msg_controllen = sizeof ctl;
        struct cmsghdr *cm = CMSG_FIRSTHDR(&msg);
        cm->cmsg_level = SOL_SOCKET;
        cm->cmsg_type = SCM_RIGHTS;
        cm->cmsg_len = CMSG_LEN(sizeof(int));
        memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    }
    for (;;) {
        ssize_t k = sendmsg(sock, &msg, MSG_NOSIGNAL);
        if (k < 0 && errno == EINTR) continue;
        return k == (ssize_t)sizeof has;
    }
}

static bool recv_fd_msg(int sock, int *fd) {
    int has = 0;
    struct msghdr msg;
    memset(&msg, 0, sizeof msg);
    struct iovec io;
    io.iov_base = &has;
    io.iov_len = sizeof has;
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof ctl);
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof ctl;
    ssize_t k;
    do {
        k = recvmsg(sock, &msg, MSG_WAITALL);
    } while (k < 0 && errno == EINTR);
    if (k != (ssize_t)sizeof has) return false;
    *fd = -1;
    if (has) {
        struct cmsghdr *cm = CMSG_FIRSTHDR(&msg);
        if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) return false;
        memcpy(fd, CMSG_DATA(cm), sizeof(int));
    }
    return true;
}

static int rdv_open(DistState *d, const char *path) {
    for (int i = 0; i < MAX_WORLD; i++) d->socks[i] = -1;
    d->listen_fd = -1;
    snprintf(d->path, sizeof d->path, "%s", path);
    struct sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    if (strlen(path) >= sizeof addr.sun_path) {
        set_error("rendezvous path too long: %s", path);
        return IFL_E_ARG;
    }
    strcpy(addr.sun_path, path);
    if (d->rank == 0) {
        unlink(path);
        d->listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
        if (d->listen_fd < 0 || bind(d->listen_fd, (struct sockaddr *)&addr, sizeof addr) != 0 ||
            listen(d->listen_fd, MAX_WORLD) != 0) {
            set_error("rendezvous: cannot listen on %s: %s", path, strerror(errno));
            return IFL_E_ARG;
        }
        sock_set_timeout(d->listen_fd);
        for (int i = 1; i < d->world; i++) {
            int fd = accept(d->listen_fd, nullptr, nullptr);
            if (fd < 0) {
                set_error("rendezvous: accept on %s: %s (%d of %d peers arrived)", path, strerror(errno), i - 1, d->world - 1);
                return IFL_E_ARG;
            }
            sock_set_timeout(fd);
            int r = -1;
            if (!read_all(fd, &r, sizeof r) || r <= 0 || r >= d->world || d->socks[r] >= 0) {
                set_error("rendezvous: bad hello (rank %d)", r);
                close(fd);
                return IFL_E_ARG;
            }
            d->socks[r] = fd;
        }
    } else {
        const time_t t0 = time(nullptr);
        int fd = -1;
        for (;;) {
            fd = socket(AF_UNIX, SOCK_STREAM, 0);
            if (fd >= 0 && connect(fd, (struct sockaddr *)&addr, sizeof addr) == 0) break;
            if (fd >= 0) close(fd);
            fd = -1;
            if (time(nullptr) - t0 > RDV_TIMEOUT_S) {
                set_error("rendezvous: rank %d could not reach rank 0 at %s", d->rank, path);
                return IFL_E_ARG;
            }
            usleep(20000);
        }
        sock_set_timeout(fd);
        if (!write_all(fd, &d->rank, sizeof d->rank)) {
            set_error("rendezvous: hello failed");
            close(fd);
            return IFL_E_ARG;
        }
        d->socks[0] = fd;
    }
    return IFL_OK;
}

static void rdv_close(DistState *d) {
    for (int i = 0; i < MAX_WORLD; i++)
        if (d->socks[i] >= 0) close(d->socks[i]);
    if (d->listen_fd >= 0) {
        close(d->listen_fd);
        unlink(d->path);
    }
}

// All-gather of one file descriptor per rank (-1 = none).  out[i] is a descriptor owned by
// the caller for every i != rank that sent one.
static int rdv_allgather_fd(DistState *d, int mine, int *out) {
    for (int i = 0; i < d->world; i++) out[i] = -1;
    bool ok = true;
    if (d->rank == 0) {
        for (int r = 1; r < d->world && ok; r++) ok = recv_fd_msg(d->socks[r], &out[r]);
        for (int r = 1; r < d->world && ok; r++)
            for (int i = 0; i < d->world && ok; i++)
                if (i != r) ok = send_fd_msg(d->socks[r], i == 0 ? mine : out[i]);
    } else {
        ok = send_fd_msg(d->socks[0], mine);
        for (int i = 0; i < d->world && ok; i++)
            if (i != d->rank) ok = recv_fd_msg(d->socks[0], &out[i]);
    }
    if (!ok) {
        set_error("rendezvous: descriptor exchange failed (%s)", strerror(errno));
        return IFL_E_ARG;
    }
    return IFL_OK;
}

// All-gather of a small POD (consistency checks, barrier).
static int rdv_allgather(DistState *d, const void *mine, void *all, size_t n) {
    char *a = (char *)all;
    memcpy(a + (size_t)d->rank * n, mine, n);
    bool ok = true;
    if (d->rank == 0) {
        for (int r = 1; r < d->world && ok; r++) ok = read_all(d->socks[r], a + (size_t)r * n, n);
        for (int r = 1; r < d->world && ok; r++) ok = write_all(d->socks[r], a, n * d->world);
    } else {
        ok = write_all(d->socks[0], mine, n) && read_all(d->socks[0], a, n * d->world);
    }
    if (!ok) {
        set_error("rendezvous: exchange failed (%s)", strerror(errno));
        return IFL_E_ARG;
    }
    return IFL_OK;
}

// ------------------------------------------------------------------- allocation ----
static size_t round_up_sz(size_t v, size_t m) { return (v + m - 1) / m * m; }

// Maps one virtual range of `size` bytes whose pages [first[g], first[g+1]) are rank g's HBM.
static int map_shared(ifl_ctx *c, void **out, size_t size, const size_t *first /* [world+1], in pages */) {
    DistState *d = c->dist;
    Drv *v = drv();
    const size_t gran = d->gran;
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    const size_t my_pages = first[d->rank + 1] - first[d->rank];
    CUmemGenericAllocationHandle mine = 0;
    int my_fd = -1;
    if (my_pages > 0) {
        IFL_DRV(v->MemCreate(&mine, my_pages * gran, &prop, 0));
        IFL_DRV(v->MemExport(&my_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    }
    int fds[MAX_WORLD];
    int rc = rdv_allgather_fd(d, my_fd, fds);
    if (my_fd >= 0) close(my_fd);
    if (rc != IFL_OK) return rc;
    CUdeviceptr va = 0;
    IFL_DRV(v->MemAddressReserve(&va, size, 0, 0, 0));
    for (int g = 0; g < d->world; g++) {
        const size_t pages = first[g + 1] - first[g];
        if (pages == 0) continue;
        CUmemGenericAllocationHandle h = mine;
        if (g != d->rank) {
            if (fds[g] < 0) {
                set_error("rank %d sent no memory handle for its %zu pages", g, pages);
                return IFL_E_ARG;
            }
            IFL_DRV(v->MemImport(&h, (void *)(uintptr_t)fds[g], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
            close(fds[g]);
        }
        IFL_DRV(v->MemMap(va + first[g] * gran, pages * gran, 0, h, 0));
        IFL_DRV(v->MemRelease(h)); // the mapping keeps the memory alive
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = c->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    IFL_DRV(v->MemSetAccess(va, size, &acc, 1));
    if (my_pages > 0) IFL_CUDA(cudaMemset((void *)(va + first[d->rank] * gran), 0, my_pages * gran));
    IFL_CUDA(cudaDeviceSynchronize());
    Mapped m;
    m.p = (void *)va;
    m.size = size;
    d->mem.push_back(m);
    *out = (void *)va;
    return dist_host_barrier(c); // nobody touches the range before every rank has zeroed its pages
}

int dist_alloc_rows(ifl_ctx *c, void **out, size_t bytes, size_t row_bytes, int h) {
    (void)h;
    if (c->world <= 1) {
        IFL_CUDA(cudaMalloc(out, bytes));
        IFL_CUDA(cudaMemset(*out, 0, bytes)); // uninitialised == zero page
        return IFL_OK;
    }
    DistState *d = c->dist;
    const size_t size = round_up_sz(bytes, d->gran), pages = size / d->gran;
    size_t first[MAX_WORLD + 1];
    // a page belongs to the rank that owns the row its first byte lies in
    int g = 0;
    first[0] = 0;
    for (size_t k = 0; k < pages; k++) {
        const int owner = rank_of_row(c->H, d->world, (long long)(k * d->gran / row_bytes));
        while (g < owner) first[++g] = k;
    }
    while (g < d->world) first[++g] = pages;
    return map_shared(c, out, size, first);
}

int dist_alloc_per_rank(ifl_ctx *c, void **out, size_t bytes, size_t *stride) {
    if (c->world <= 1) {
        IFL_CUDA(cudaMalloc(out, bytes));
        IFL_CUDA(cudaMemset(*out, 0, bytes));
        *stride = bytes;
        return IFL_OK;
    }
    DistState *d = c->dist;
    const size_t per = round_up_sz(bytes, d->gran), pages = per / d->gran;
    size_t first[MAX_WORLD + 1];
    for (int g = 0; g <= d->world; g++) first[g] = pages * g;
    *stride = per;
    return map_shared(c, out, per * d->world, first);
}

void dist_free_mem(ifl_ctx *c, void *p) {
    if (!p) return;
    if (c->world > 1 && c->dist) {
        DistState *d = c->dist;
        for (size_t i = 0; i < d->mem.size(); i++)
            if (d->mem[i].p == p) {
                Drv *v = drv();
                v->MemUnmap((CUdeviceptr)p, d->mem[i].size);
                v->MemAddressFree((CUdeviceptr)p, d->mem[i].size);
                d->mem.erase(d->mem.begin() + i);
                return;
            }
    }
    cudaFree(p);
}

// --------------------------------------------------------------------- barriers ----
__global__ void __launch_bounds__(32) k_dist_barrier(DistDev d, const SolveScalars *gate) {
    if (gate && gate->done) return; // `done` is bit-identical on all ranks: all skip or none
    dist_barrier_block(d);
}

int dist_barrier(ifl_ctx *c, bool gated) {
    if (c->world <= 1) return IFL_OK;
    ProfScope ps_(c, IFL_K_SCALAR);
    k_dist_barrier<<<1, 32, 0, c->stream>>>(c->ddev, gated ? c->scal : nullptr);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int dist_host_barrier(ifl_ctx *c) {
    if (c->world <= 1 || !c->dist) return IFL_OK;
    char all[MAX_WORLD], one = 1;
    return rdv_allgather(c->dist, &one, all, 1);
}

// Sum of one integer over the ranks (host side; every rank gets the same total).
int dist_host_sum(ifl_ctx *c, long long *v) {
    if (c->world <= 1 || !c->dist) return IFL_OK;
    long long all[MAX_WORLD];
    int rc = rdv_allgather(c->dist, v, all, sizeof(long long));
    if (rc != IFL_OK) return rc;
    long long t = 0;
    for (int g = 0; g < c->world; g++) t += all[g];
    *v = t;
    return IFL_OK;
}

// ------------------------------------------------------------------- life cycle ----
int dist_init(ifl_ctx *c, int rank, int world, const char *rendezvous) {
    c->rank = rank;
    c->world = world;
    c->ry0 = 0;
    c->ry1 = c->H;
    memset(&c->ddev, 0, sizeof c->ddev);
    c->ddev.world = 1;
    if (world <= 1) return IFL_OK;
    if (world > MAX_WORLD || rank < 0 || rank >= world || !rendezvous) {
        set_error("ifl_create_dist: rank %d of %d (at most %d ranks) needs a rendezvous path", rank, world, MAX_WORLD);
        return IFL_E_ARG;
    }
    if ((c->H + 63) / 64 < world) {
        set_error("ifl_create_dist: %d rows give fewer than %d strips of 64 rows", c->H, world);
        return IFL_E_ARG;
    }
    Drv *v = drv();
    if (!v->ok) {
        set_error("ifl_create_dist: the CUDA driver lacks the virtual memory management entry points");
        return IFL_E_CUDA;
    }
    DistState *d = new DistState();
    d->rank = rank;
    d->world = world;
    c->dist = d;
    int rc = rdv_open(d, rendezvous);
    if (rc != IFL_OK) return rc;
    // every rank must describe the same solver
    int mine[4] = {c->W, c->H, c->version, world}, all[4 * MAX_WORLD];
    if ((rc = rdv_allgather(d, mine, all, sizeof mine)) != IFL_OK) return rc;
    for (int g = 0; g < world; g++)
        if (memcmp(all + 4 * g, mine, sizeof mine) != 0) {
            set_error("ifl_create_dist: rank %d was created with a different (w, h, chapter, world)", g);
            return IFL_E_ARG;
        }
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    IFL_DRV(v->MemGetGranularity(&d->gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    c->ry0 = slab_row0(c->H, world, rank);
    c->ry1 = rank + 1 == world ? c->H : slab_row0(c->H, world, rank + 1);
    // barrier flags: one page per rank
    void *base = nullptr;
    size_t stride = 0;
    if ((rc = dist_alloc_per_rank(c, &base, 4096, &stride)) != IFL_OK) return rc;
    c->ddev.rank = rank;
    c->ddev.world = world;
    for (int g = 0; g < world; g++) c->ddev.flags_peer[g] = (unsigned long long *)((char *)base + stride * g);
    c->ddev.flags_local = c->ddev.flags_peer[rank];
    c->ddev.epoch = c->ddev.flags_local + 64;
    return IFL_OK;
}

void dist_free(ifl_ctx *c) {
    if (!c->dist) return;
    DistState *d = c->dist;
    dist_host_barrier(c); // peers may still be reading this rank's slabs
    Drv *v = drv();
    for (size_t i = 0; i < d->mem.size(); i++) {
        v->MemUnmap((CUdeviceptr)d->mem[i].p, d->mem[i].size);
        v->MemAddressFree((CUdeviceptr)d->mem[i].p, d->mem[i].size);
    }
    d->mem.clear();
    rdv_close(d);
    delete d;
    c->dist = nullptr;
}

} // namespace ifl

using namespace ifl;

extern "C" {

int ifl_dist_plan(int h, int world, int rank, int *row0, int *row1) {
    if (h < 2 || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || (h + 63) / 64 < world) {
        set_error("ifl_dist_plan: bad argument (h=%d world=%d rank=%d)", h, world, rank);
        return IFL_E_ARG;
    }
    if (row0) *row0 = slab_row0(h, world, rank);
    if (row1) *row1 = rank + 1 == world ? h : slab_row0(h, world, rank + 1);
    return IFL_OK;
}

int ifl_dist_info(const ifl_ctx *c, int *rank, int *world, int *row0, int *row1) {
    if (!c) return IFL_E_ARG;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (row0) *row0 = c->ry0;
    if (row1) *row1 = c->ry1;
    return IFL_OK;
}

int ifl_dist_barrier(ifl_ctx *c) {
    if (!c) return IFL_E_ARG;
    int rc = dist_barrier(c, false);
    if (rc != IFL_OK) return rc;
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return dist_host_barrier(c);
}

// Host-only self-test of the rendezvous (no CUDA): every rank creates a pipe, writes its
// rank into it, passes the read end around and checks what it reads from the others.
int ifl_dist_selftest(int rank, int world, const char *rendezvous) {
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || !rendezvous) return IFL_E_ARG;
    DistState d;
    d.rank = rank;
    d.world = world;
    int rc = rdv_open(&d, rendezvous);
    if (rc != IFL_OK) return rc;
    int pfd[2], fds[MAX_WORLD];
    if (pipe(pfd) != 0) return IFL_E_ARG;
    int tag = 1000 + rank;
    for (int g = 1; g < world; g++) // one copy per reader
        if (write(pfd[1], &tag, sizeof tag) != (ssize_t)sizeof tag) rc = IFL_E_ARG;
    if (rc == IFL_OK) rc = rdv_allgather_fd(&d, pfd[0], fds);
    for (int g = 0; g < world && rc == IFL_OK; g++) {
        if (g == rank) continue;
        int got = -1;
        if (fds[g] < 0 || read(fds[g], &got, sizeof got) != (ssize_t)sizeof got || got != 1000 + g) {
            set_error("selftest: rank %d read %d from rank %d's pipe", rank, got, g);
            rc = IFL_E_ARG;
        }
        if (fds[g] >= 0) close(fds[g]);
    }
    int vals[MAX_WORLD];
    if (rc == IFL_OK) rc = rdv_allgather(&d, &tag, vals, sizeof tag);
    for (int g = 0; g < world && rc == IFL_OK; g++)
        if (vals[g] != 1000 + g) rc = IFL_E_ARG;
    close(pfd[0]);
    close(pfd[1]);
    rdv_close(&d);
    return rc;
}

} // extern "C"
