// ifl_api.cu -- the C ABI of libifl_b200.so (include/ifl_b200.h): context lifetime,
// dense<->pitched data movement, and the per-chapter update() sequences.
// Everything here is host code; all arithmetic happens in the kernels.
#include "ifl_internal.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace ifl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---- per-class event timing.  Pairs live in a ring; when it is full the oldest pair is
// harvested (which waits for that launch to finish -- it bounds the queue depth only).
static void prof_harvest_one(ifl_ctx *c) {
    const int i = c->prof_head;
    float ms = 0.f;
    if (cudaEventSynchronize(c->prof_ev[2 * i + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) == cudaSuccess) {
        c->prof_ms[c->prof_cls[i]] += ms;
        c->prof_n[c->prof_cls[i]] += 1;
    }
    c->prof_head = (c->prof_head + 1) % PROF_RING;
    c->prof_count--;
}

void prof_begin(ifl_ctx *c, int cls) {
    if (c->prof_count == PROF_RING) prof_harvest_one(c);
    const int i = (c->prof_head + c->prof_count) % PROF_RING;
    c->prof_cls[i] = cls;
    cudaEventRecord(c->prof_ev[2 * i], c->stream);
}

void prof_end(ifl_ctx *c) {
    const int i = (c->prof_head + c->prof_count) % PROF_RING;
    cudaEventRecord(c->prof_ev[2 * i + 1], c->stream);
    c->prof_count++;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// All-zero array (uninitialised == zero page).  With row-slab
// multi-GPU the rows [ry0, ry1) are this rank's; v and phi (h+1 rows) follow the cell rows
// and the last rank takes the extra row.
static int alloc_arr(ifl_ctx *c, Arr &a, int w, int h) {
    a.w = w;
    a.h = h;
    a.pitch = round_up(w, TILE);
    a.rows = round_up(h, TILE) + TILE;
    a.p = nullptr;
    a.ry0 = c->ry0 < h ? c->ry0 : h;
    a.ry1 = (c->rank + 1 == c->world) ? h : (c->ry1 < h ? c->ry1 : h);
    return dist_alloc_rows(c, (void **)&a.p, a.bytes(), (size_t)a.pitch * sizeof(double), h);
}

static void free_arr(ifl_ctx *c, Arr &a) {
    if (a.p) dist_free_mem(c, a.p);
    a.p = nullptr;
}

__global__ void k_fill2d(Arr a, double value) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = a.ry0 + blockIdx.y;
    if (x < a.w && y < a.ry1) a.p[x + (size_t)y * a.pitch] = value;
}

static int fill_arr(const Arr &a, double value) {
    if (a.ry1 <= a.ry0) return IFL_OK;
    k_fill2d<<<dim3((a.w + 255) / 256, a.ry1 - a.ry0), 256>>>(a, value);
    IFL_CUDA(cudaGetLastError());
    return IFL_OK;
}

static int alloc_field(ifl_ctx *c, Field &f, int w, int h, double ox, double oy, bool solids) {
    f.w = w;
    f.h = h;
    f.ox = ox;
    f.oy = oy;
    int rc = alloc_arr(c, f.src, w, h);
    if (rc == IFL_OK) rc = alloc_arr(c, f.dst, w, h);
    if (rc != IFL_OK || !solids) return rc;
    // FluidQuantity ctor of chapters 4+ (v5:356-377): every cell fluid, volume 1
    if ((rc = alloc_arr(c, f.volume, w, h)) != IFL_OK) return rc;
    if ((rc = alloc_arr(c, f.normalX, w, h)) != IFL_OK) return rc;
    if ((rc = alloc_arr(c, f.normalY, w, h)) != IFL_OK) return rc;
    if ((rc = alloc_arr(c, f.phi, w + 1, h + 1)) != IFL_OK) return rc;
    if ((rc = fill_arr(f.volume, 1.0)) != IFL_OK) return rc;
    const size_t nb = (size_t)f.src.pitch * f.src.rows; // byte arrays share the doubles' row pitch (and slab split)
    if ((rc = dist_alloc_rows(c, (void **)&f.cell, nb, (size_t)f.src.pitch, h)) != IFL_OK) return rc; // 0 == CELL_FLUID
    if ((rc = dist_alloc_rows(c, (void **)&f.body, nb, (size_t)f.src.pitch, h)) != IFL_OK) return rc;
    if ((rc = dist_alloc_rows(c, (void **)&f.mask, nb, (size_t)f.src.pitch, h)) != IFL_OK) return rc;
    IFL_CUDA(cudaMalloc(&f.solid_list, (size_t)w * (f.src.ry1 - f.src.ry0 + 1) * sizeof(int)));
    IFL_CUDA(cudaMalloc(&f.solid_count, sizeof(int)));
    return IFL_OK;
}

static void free_field(ifl_ctx *c, Field &f) {
    Arr *arrs[] = {&f.src, &f.dst, &f.volume, &f.normalX, &f.normalY, &f.phi};
    for (int i = 0; i < 6; i++) free_arr(c, *arrs[i]);
    if (f.cell) dist_free_mem(c, f.cell);
    if (f.body) dist_free_mem(c, f.body);
    if (f.mask) dist_free_mem(c, f.mask);
    if (f.solid_list) cudaFree(f.solid_list);
    if (f.solid_count) cudaFree(f.solid_count);
    f.cell = f.body = f.mask = nullptr;
    f.solid_list = f.solid_count = nullptr;
}

static Arr *buf_arr(ifl_ctx *c, int buf) {
    switch (buf) {
    case IFL_BUF_D_SRC: return &c->fd[IFL_FIELD_D].src;
    case IFL_BUF_D_DST: return &c->fd[IFL_FIELD_D].dst;
    case IFL_BUF_U_SRC: return &c->fd[IFL_FIELD_U].src;
    case IFL_BUF_U_DST: return &c->fd[IFL_FIELD_U].dst;
    case IFL_BUF_V_SRC: return &c->fd[IFL_FIELD_V].src;
    case IFL_BUF_V_DST: return &c->fd[IFL_FIELD_V].dst;
    case IFL_BUF_T_SRC: return c->version >= 6 ? &c->fd[IFL_FIELD_T].src : nullptr;
    case IFL_BUF_T_DST: return c->version >= 6 ? &c->fd[IFL_FIELD_T].dst : nullptr;
    case IFL_BUF_R: return &c->r;
    case IFL_BUF_P: return &c->p;
    case IFL_BUF_Z: return &c->z;
    case IFL_BUF_S: return &c->s;
    case IFL_BUF_PRECON: return &c->precon;
    case IFL_BUF_ADIAG: return &c->aDiag;
    case IFL_BUF_APLUSX: return &c->aPlusX;
    case IFL_BUF_APLUSY: return &c->aPlusY;
    case IFL_BUF_UDENSITY: return c->version >= 7 ? &c->uDensity : nullptr;
    case IFL_BUF_VDENSITY: return c->version >= 7 ? &c->vDensity : nullptr;
    }
    return nullptr;
}

static bool pcg_chapter(const ifl_ctx *c) { return c->version >= 3; }

// The sticky watchdog word (SolveScalars::watchdog) is raised by any in-kernel dependency wait that
// timed out: a sweep hand-off, a TMA ring barrier, a rank barrier between stages.  The solves look
// at it in their own read-backs; everything else (the factorisation, the stage barriers of
// update()) is caught here, at the end of every entry point that launches such kernels.
int check_watchdog(ifl_ctx *c, const char *where) {
    int *h = reinterpret_cast<int *>(c->result_h + 7);
    IFL_CUDA(cudaMemcpyAsync(h, &c->scal->watchdog, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    if (*h) {
        cudaMemsetAsync(&c->scal->watchdog, 0, sizeof(int), c->stream);
        set_error("%s: a dependency wait (wavefront hand-off or rank barrier) timed out", where);
        return IFL_E_WATCHDOG;
    }
    return IFL_OK;
}

} // namespace ifl

using namespace ifl;

// every entry point runs on the context's device (a process may hold contexts on several GPUs)
#define CHECK_CTX(c)                                   \
    do {                                               \
        if (!(c)) {                                    \
            set_error("null context");                 \
            return IFL_E_ARG;                          \
        }                                              \
        if (cudaSetDevice((c)->device) != cudaSuccess) { \
            set_error("cudaSetDevice(%d) failed", (c)->device); \
            return IFL_E_CUDA;                         \
        }                                              \
    } while (0)
#define TRY(expr)                          \
    do {                                   \
        int rc_ = (expr);                  \
        if (rc_ != IFL_OK) return rc_;     \
    } while (0)

extern "C" {

const char *ifl_last_error(void) { return g_err; }

static int create_ctx(ifl_ctx **out, int w, int h, int version, int device, int rank, int world, const char *rendezvous) {
    if (!out || w < 2 || h < 2 || version < 1 || version > 8) {
        set_error("ifl_create: bad argument (w=%d h=%d version=%d)", w, h, version);
        return IFL_E_ARG;
    }
    if (version == 7 && h > w + 1) {
        // v7:694/700 index _vDensity (w columns per row) with _u's row stride w+1: on grids taller
        // than w+1 the reference itself reads past the end of the array (undefined behaviour),
        // so there is nothing to be bit-exact with.
        set_error("ifl_create: chapter 7 needs h <= w + 1 (the reference reads _vDensity out of bounds on taller grids)");
        return IFL_E_ARG;
    }
    if (world > 1 && version > 7) {
        set_error("ifl_create_dist: row-slab multi-GPU covers chapters 1-7 (the chapter-8 particle set is not sharded yet)");
        return IFL_E_ARG;
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("ifl_create: no CUDA device (%s); libifl_b200 has no CPU path", cudaGetErrorString(cudaGetLastError()));
        return IFL_E_CUDA;
    }
    IFL_CUDA(cudaSetDevice(device));
    ifl_ctx *c = (ifl_ctx *)calloc(1, sizeof(ifl_ctx));
    if (!c) return IFL_E_NOMEM;
    c->W = w;
    c->H = h;
    c->version = version;
    c->device = device;
    c->hx = 1.0 / (double)(w < h ? w : h); // v3:402
    int rc = IFL_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("cudaStreamCreate failed");
            rc = IFL_E_CUDA;
            break;
        }
        if ((rc = dist_init(c, rank, world, rendezvous)) != IFL_OK) break;
        // v3:404-406: d at cell centres, u/v on the staggered faces
        const bool solids = version >= 4;
        if ((rc = alloc_field(c, c->fd[IFL_FIELD_D], w, h, 0.5, 0.5, solids)) != IFL_OK) break;
        if ((rc = alloc_field(c, c->fd[IFL_FIELD_U], w + 1, h, 0.0, 0.5, solids)) != IFL_OK) break;
        if ((rc = alloc_field(c, c->fd[IFL_FIELD_V], w, h + 1, 0.5, 0.0, solids)) != IFL_OK) break;
        if (version >= 6) { // temperature field at ambient temperature, v6:921-941
            c->t_amb = 294.0;
            c->g = 9.81;
            if ((rc = alloc_field(c, c->fd[IFL_FIELD_T], w, h, 0.5, 0.5, true)) != IFL_OK) break;
            if ((rc = fill_arr(c->fd[IFL_FIELD_T].src, c->t_amb)) != IFL_OK) break;
        }
        if (version >= 7) {
            if ((rc = alloc_arr(c, c->uDensity, w + 1, h)) != IFL_OK) break;
            if ((rc = alloc_arr(c, c->vDensity, w, h + 1)) != IFL_OK) break;
        }
        if (solids) {
            if ((rc = alloc_arr(c, c->pe, w, h)) != IFL_OK) break;
            if ((rc = alloc_arr(c, c->fmask, w, h)) != IFL_OK) break;
            if ((rc = fill_arr(c->fmask, 1.0)) != IFL_OK) break;
            if (cudaMalloc(&c->bodies_d, MAX_BODIES * sizeof(BodyDev)) != cudaSuccess ||
                cudaMalloc(&c->ext_ready, sizeof(int)) != cudaSuccess) {
                set_error("ifl_create: body table allocation failed");
                rc = IFL_E_CUDA;
                break;
            }
        }
        Arr *cells[] = {&c->r, &c->p, &c->z, &c->s, &c->q, &c->precon, &c->aDiag, &c->aPlusX, &c->aPlusY, &c->cx, &c->cy};
        const int ncells = pcg_chapter(c) ? 11 : 2; // chapters 1-2 only own _r and _p (v2:219-220)
        for (int i = 0; i < ncells && rc == IFL_OK; i++) rc = alloc_arr(c, *cells[i], w, h);
        if (rc == IFL_OK && version == 3) rc = alloc_arr(c, c->s2, w, h); // ping-pong partner of s (pcg_kernels.cu: k_xpay_matvec)
        if (rc != IFL_OK) break;
        // two partial buffers (reductions alternate, pcg_kernels.cu); with several ranks they are
        // rank 0's memory and every rank folds all of them
        size_t npart = (size_t)((w + 255) / 256) * ((h + 15) / 16) + 64, pstride = 0;
        if (npart < MAX_PARTIALS) npart = MAX_PARTIALS;
        void *pbase = nullptr;
        if ((rc = dist_alloc_per_rank(c, &pbase, 2 * npart * sizeof(double), &pstride)) != IFL_OK) break;
        c->partials_buf[0] = (double *)pbase;
        c->partials_buf[1] = c->partials_buf[0] + npart;
        c->partials = c->partials_buf[0];
        if (cudaMalloc(&c->scal, sizeof(SolveScalars)) != cudaSuccess ||
            cudaMemset(c->scal, 0, sizeof(SolveScalars)) != cudaSuccess ||
            cudaMallocHost(&c->scal_h, 2 * sizeof(SolveScalars)) != cudaSuccess ||
            cudaMallocHost(&c->result_h, 8 * sizeof(double)) != cudaSuccess) {
            set_error("ifl_create: scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = IFL_E_CUDA;
            break;
        }
        c->ddev.watchdog = &c->scal->watchdog;
        if ((rc = sweep_init(c)) != IFL_OK) break;
        if (version >= 8 && (rc = flip_init(c)) != IFL_OK) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            set_error("ifl_create: %s", cudaGetErrorString(cudaGetLastError()));
            rc = IFL_E_CUDA;
            break;
        }
        rc = dist_host_barrier(c);
    } while (0);
    if (rc != IFL_OK) {
        ifl_destroy(c);
        return rc;
    }
    *out = c;
    return IFL_OK;
}

int ifl_create(ifl_ctx **out, int w, int h, int version, int device) {
    return create_ctx(out, w, h, version, device, 0, 1, nullptr);
}

int ifl_create_dist(ifl_ctx **out, int w, int h, int version, int device, int rank, int world, const char *rendezvous) {
    if (world < 1) {
        set_error("ifl_create_dist: world = %d", world);
        return IFL_E_ARG;
    }
    return create_ctx(out, w, h, version, device, rank, world, rendezvous);
}

int ifl_destroy(ifl_ctx *c) {
    if (!c) return IFL_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    dist_host_barrier(c); // peers may still be reading this rank's slabs
    for (int i = 0; i < 4; i++) free_field(c, c->fd[i]);
    free_arr(c, c->pe);
    free_arr(c, c->fmask);
    free_arr(c, c->uDensity);
    free_arr(c, c->vDensity);
    if (c->bodies_d) cudaFree(c->bodies_d);
    if (c->ext_ready) cudaFree(c->ext_ready);
    Arr *cells[] = {&c->r, &c->p, &c->z, &c->s, &c->q, &c->precon, &c->aDiag, &c->aPlusX, &c->aPlusY, &c->cx, &c->cy};
    for (int i = 0; i < 11; i++) free_arr(c, *cells[i]);
    free_arr(c, c->s2);
    if (c->partials_buf[0]) dist_free_mem(c, c->partials_buf[0]);
    if (c->scal) cudaFree(c->scal);
    if (c->scal_h) cudaFreeHost(c->scal_h);
    if (c->result_h) cudaFreeHost(c->result_h);
    if (c->prof_ev) {
        for (int i = 0; i < 2 * PROF_RING; i++)
            if (c->prof_ev[i]) cudaEventDestroy(c->prof_ev[i]);
        free(c->prof_ev);
        free(c->prof_cls);
    }
    sweep_free(c);
    flip_free(c);
    dist_free(c);
    if (c->stream) cudaStreamDestroy(c->stream);
    free(c);
    return IFL_OK;
}

int ifl_profile(ifl_ctx *c, int on) {
    CHECK_CTX(c);
    while (c->prof_count > 0) prof_harvest_one(c);
    if (on) {
        if (!c->prof_ev) {
            c->prof_ev = (cudaEvent_t *)calloc(2 * PROF_RING, sizeof(cudaEvent_t));
            c->prof_cls = (int *)calloc(PROF_RING, sizeof(int));
            if (!c->prof_ev || !c->prof_cls) return IFL_E_NOMEM;
            for (int i = 0; i < 2 * PROF_RING; i++) IFL_CUDA(cudaEventCreate(&c->prof_ev[i]));
        }
        memset(c->prof_ms, 0, sizeof c->prof_ms);
        memset(c->prof_n, 0, sizeof c->prof_n);
        c->prof_head = c->prof_count = 0;
    }
    c->prof_on = on ? 1 : 0;
    return IFL_OK;
}

int ifl_profile_read(ifl_ctx *c, double *ms, long long *launches) {
    CHECK_CTX(c);
    while (c->prof_count > 0) prof_harvest_one(c);
    for (int i = 0; i < IFL_K_COUNT_; i++) {
        if (ms) ms[i] = c->prof_ms[i];
        if (launches) launches[i] = c->prof_n[i];
    }
    return IFL_OK;
}

int ifl_debug_sweep_times(ifl_ctx *c, int arm, unsigned long long *out_ns, int capacity) {
    CHECK_CTX(c);
    if (arm) {
        c->sweep_times = c->sweep_times_buf;
        return IFL_OK;
    }
    c->sweep_times = nullptr;
    if (!out_ns || capacity < 16 * c->n_strips) {
        set_error("ifl_debug_sweep_times: need room for %d values", 16 * c->n_strips);
        return IFL_E_ARG;
    }
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    IFL_CUDA(cudaMemcpy(out_ns, c->sweep_times_buf, (size_t)16 * c->n_strips * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost));
    return c->tri_engine ? (c->H + 63) / 64 : c->n_strips; // strips of the engine that ran the triangular solves
}

long long ifl_launch_count(const ifl_ctx *c) { return c ? c->launches : 0; }
void *ifl_stream(const ifl_ctx *c) { return c ? (void *)c->stream : nullptr; }

int ifl_sync(ifl_ctx *c) {
    CHECK_CTX(c);
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

size_t ifl_buf_elems(const ifl_ctx *c, int buf) {
    Arr *a = c ? buf_arr(const_cast<ifl_ctx *>(c), buf) : nullptr;
    return (a && a->p) ? (size_t)a->w * a->h : 0;
}

// Dense host rows (w doubles) <-> pitched device rows; pageable or pinned host memory.
int ifl_upload(ifl_ctx *c, int buf, const double *host) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p || !host) {
        set_error("ifl_upload: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    if (buf == IFL_BUF_ADIAG || buf == IFL_BUF_APLUSX || buf == IFL_BUF_APLUSY) c->matrix_uniform = 0; // the caller's own matrix
    // several ranks: collective; every rank passes the whole dense array and copies its own rows
    if (a->ry1 > a->ry0)
        IFL_CUDA(cudaMemcpy2DAsync(a->p + (size_t)a->ry0 * a->pitch, (size_t)a->pitch * 8, host + (size_t)a->ry0 * a->w,
                                   (size_t)a->w * 8, (size_t)a->w * 8, a->ry1 - a->ry0, cudaMemcpyHostToDevice, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return dist_host_barrier(c);
}

int ifl_download(ifl_ctx *c, int buf, double *host) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p || !host) {
        set_error("ifl_download: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    // several ranks: collective; every rank receives the WHOLE array (the peers' slabs come over NVLink)
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    TRY(dist_host_barrier(c)); // the peers' kernels that produce their slabs have finished
    IFL_CUDA(cudaMemcpy2DAsync(host, (size_t)a->w * 8, a->p, (size_t)a->pitch * 8, (size_t)a->w * 8, a->h,
                               cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return dist_host_barrier(c); // nobody overwrites a slab while a peer still copies it
}

int ifl_fill(ifl_ctx *c, int buf, double value) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p) {
        set_error("ifl_fill: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    if (value == 0.0 && !signbit(value)) {
        IFL_CUDA(cudaMemsetAsync((char *)a->p + a->own_begin(), 0, a->own_end() - a->own_begin(), c->stream));
        return IFL_OK;
    }
    // rare path (tests): build one dense row block on the host
    const size_t n = (size_t)a->w * a->h;
    double *tmp = (double *)malloc(n * sizeof(double));
    if (!tmp) return IFL_E_NOMEM;
    for (size_t i = 0; i < n; i++) tmp[i] = value;
    int rc = ifl_upload(c, buf, tmp);
    free(tmp);
    return rc;
}

// ---- FluidQuantity ops ---------------------------------------------------------
static int check_field(ifl_ctx *c, int field) {
    if (field < 0 || field > IFL_FIELD_T || (field == IFL_FIELD_T && c->version < 6)) {
        set_error("bad field id %d for chapter %d", field, c->version);
        return IFL_E_ARG;
    }
    return IFL_OK;
}

int ifl_quantity_add_inflow(ifl_ctx *c, int field, double x0, double y0, double x1, double y1, double v) {
    CHECK_CTX(c);
    TRY(check_field(c, field));
    return launch_add_inflow(c, field, x0, y0, x1, y1, v);
}

int ifl_advect(ifl_ctx *c, int field, double timestep) {
    CHECK_CTX(c);
    TRY(check_field(c, field));
    TRY(dist_barrier(c));
    return launch_advect(c, field, timestep);
}

int ifl_max_timestep(ifl_ctx *c, double *result) { // FluidSolver::maxTimestep v1:310-328
    CHECK_CTX(c);
    if (!result) return IFL_E_ARG;
    if (c->world > 1) {
        set_error("ifl_max_timestep is a one-GPU entry point");
        return IFL_E_ARG;
    }
    TRY(launch_max_velocity(c));
    double *dev_out = &c->scal->beta; // scratch slot; no solve is in flight
    TRY(launch_finish_reduce(c, true, dev_out));
    IFL_CUDA(cudaMemcpyAsync(c->result_h, dev_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    const double max_velocity = c->result_h[0];
    const double dt = 2.0 * c->hx / max_velocity; // v1:324 (inf when the fluid is at rest)
    *result = dt < 1.0 ? dt : 1.0;               // std::min(maxTimestep, 1.0)  v1:327
    return IFL_OK;
}

int ifl_flip(ifl_ctx *c, int field) {
    CHECK_CTX(c);
    TRY(check_field(c, field));
    Field &f = c->fd[field];
    Arr t = f.src; // pointer swap, v3:105-107 (stream-ordered: kernels take Arr by value at launch)
    f.src = f.dst;
    f.dst = t;
    return IFL_OK;
}

// ---- chapters 4+: solid bodies ---------------------------------------------------------
static int need_solids(ifl_ctx *c, const char *what) {
    if (c->version < 4) {
        set_error("%s: solid bodies belong to chapters 4+", what);
        return IFL_E_ARG;
    }
    return IFL_OK;
}

int ifl_set_bodies(ifl_ctx *c, const ifl_body *bodies, int n) {
    CHECK_CTX(c);
    TRY(need_solids(c, "ifl_set_bodies"));
    if (n < 0 || n > MAX_BODIES || (n > 0 && !bodies)) {
        set_error("ifl_set_bodies: %d bodies (at most %d: FluidQuantity::_body is a uint8_t)", n, MAX_BODIES);
        return IFL_E_ARG;
    }
    BodyDev tmp[MAX_BODIES];
    for (int i = 0; i < n; i++) {
        tmp[i].kind = bodies[i].kind;
        tmp[i].pad = 0;
        tmp[i].posX = bodies[i].pos_x;
        tmp[i].posY = bodies[i].pos_y;
        tmp[i].scaleX = bodies[i].scale_x;
        tmp[i].scaleY = bodies[i].scale_y;
        tmp[i].theta = bodies[i].theta;
        tmp[i].velX = bodies[i].vel_x;
        tmp[i].velY = bodies[i].vel_y;
        tmp[i].velTheta = bodies[i].vel_theta;
        tmp[i].cosT = cos(bodies[i].theta); // host libm, as rotate() does (v4:58-62)
        tmp[i].sinT = sin(bodies[i].theta);
    }
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    if (n > 0) IFL_CUDA(cudaMemcpy(c->bodies_d, tmp, n * sizeof(BodyDev), cudaMemcpyHostToDevice));
    c->n_bodies = n;
    return IFL_OK;
}

int ifl_fill_solid_fields(ifl_ctx *c, int field) {
    CHECK_CTX(c);
    TRY(need_solids(c, "ifl_fill_solid_fields"));
    TRY(check_field(c, field));
    TRY(dist_barrier(c));
    return launch_fill_solid_fields(c, field);
}

int ifl_set_boundary_condition(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_solids(c, "ifl_set_boundary_condition"));
    TRY(dist_barrier(c));
    return launch_set_boundary_condition(c);
}

int ifl_extrapolate(ifl_ctx *c, int field) {
    CHECK_CTX(c);
    TRY(need_solids(c, "ifl_extrapolate"));
    TRY(check_field(c, field));
    TRY(dist_barrier(c));
    if (c->version >= 8) return flip_extrapolate(c, field); // + CELL_EMPTY cells, stack-ordered (v8:611-651)
    return launch_extrapolate(c, field);
}

// ---- chapters 6+: heat and variable density ----------------------------------------------
static int need_heat(ifl_ctx *c, const char *what) {
    if (c->version < 6) {
        set_error("%s: heat / buoyancy belong to chapters 6+", what);
        return IFL_E_ARG;
    }
    return IFL_OK;
}

int ifl_set_fluid_params(ifl_ctx *c, double rho_air, double rho_soot, double diffusion) {
    CHECK_CTX(c);
    TRY(need_heat(c, "ifl_set_fluid_params"));
    c->rho_air = rho_air;
    c->rho_soot = rho_soot;
    c->diffusion = diffusion;
    return IFL_OK;
}

double ifl_ambient_t(const ifl_ctx *c) { return c ? c->t_amb : 0.0; }

int ifl_build_heat_matrix(ifl_ctx *c, double timestep) {
    CHECK_CTX(c);
    TRY(need_heat(c, "ifl_build_heat_matrix"));
    TRY(dist_barrier(c));
    return launch_build_heat_matrix(c, timestep);
}

int ifl_add_buoyancy(ifl_ctx *c, double timestep) {
    CHECK_CTX(c);
    TRY(need_heat(c, "ifl_add_buoyancy"));
    TRY(dist_barrier(c));
    return launch_add_buoyancy(c, timestep);
}

int ifl_compute_densities(ifl_ctx *c) {
    CHECK_CTX(c);
    if (c->version < 7) {
        set_error("ifl_compute_densities: variable density belongs to chapters 7+");
        return IFL_E_ARG;
    }
    TRY(dist_barrier(c));
    return launch_compute_densities(c);
}

int ifl_add_inflow_t(ifl_ctx *c, double x, double y, double w, double h, double d, double t, double u, double v) {
    CHECK_CTX(c);
    TRY(need_heat(c, "ifl_add_inflow_t"));
    TRY(launch_add_inflow(c, IFL_FIELD_D, x, y, x + w, y + h, d)); // v6:1010-1015
    TRY(launch_add_inflow(c, IFL_FIELD_T, x, y, x + w, y + h, t));
    TRY(launch_add_inflow(c, IFL_FIELD_U, x, y, x + w, y + h, u));
    TRY(launch_add_inflow(c, IFL_FIELD_V, x, y, x + w, y + h, v));
    return IFL_OK;
}

// ---- chapter 8: FLIP -----------------------------------------------------------------------
static int need_flip(ifl_ctx *c, const char *what) {
    if (c->version < 8) {
        set_error("%s: particles belong to chapter 8", what);
        return IFL_E_ARG;
    }
    return IFL_OK;
}

long long ifl_particles_capacity(const ifl_ctx *c) { return (c && c->version >= 8) ? flip_particle_capacity(c) : 0; }
long long ifl_particles_count(const ifl_ctx *c) { return (c && c->version >= 8) ? flip_particle_count(c) : 0; }

int ifl_particles_init(ifl_ctx *c, int avg_per_cell) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_init"));
    return flip_particles_init(c, avg_per_cell);
}

int ifl_count_particles(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_count_particles"));
    return flip_count_particles(c);
}

int ifl_prune_particles(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_prune_particles"));
    return flip_prune_particles(c);
}

int ifl_seed_particles(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_seed_particles"));
    return flip_seed_particles(c);
}

int ifl_particles_to_grid(ifl_ctx *c, long long *count) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_to_grid"));
    return flip_particles_to_grid(c, count);
}

int ifl_particles_peek(ifl_ctx *c, int what, long long first, long long n, void *host) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_peek"));
    return flip_peek(c, what, first, n, host);
}

int ifl_particles_upload(ifl_ctx *c, long long count, const double *px, const double *py, const double *pd, const double *pt,
                         const double *pu, const double *pv) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_upload"));
    if (count > 0 && (!px || !py)) {
        set_error("ifl_particles_upload: null position array");
        return IFL_E_ARG;
    }
    const double *props[4] = {pd, pt, pu, pv};
    return flip_set_particles(c, count, px, py, props);
}

int ifl_particles_download(ifl_ctx *c, long long *count, double *px, double *py, double *pd, double *pt, double *pu,
                           double *pv) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_download"));
    double *props[4] = {pd, pt, pu, pv};
    return flip_get_particles(c, count, px, py, props);
}

int ifl_from_particles(ifl_ctx *c, int field) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_from_particles"));
    TRY(check_field(c, field));
    return launch_from_particles(c, field);
}

int ifl_grid_to_particles(ifl_ctx *c, double alpha) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_grid_to_particles"));
    return launch_grid_to_particles(c, alpha);
}

int ifl_quantity_copy(ifl_ctx *c, int field) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_quantity_copy"));
    TRY(check_field(c, field));
    return launch_copy(c, field);
}

int ifl_quantity_diff(ifl_ctx *c, int field, double alpha) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_quantity_diff"));
    TRY(check_field(c, field));
    return launch_diff(c, field, alpha, 0);
}

int ifl_quantity_undiff(ifl_ctx *c, int field, double alpha) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_quantity_undiff"));
    TRY(check_field(c, field));
    return launch_diff(c, field, alpha, 1);
}

int ifl_particles_advect(ifl_ctx *c, double timestep) {
    CHECK_CTX(c);
    TRY(need_flip(c, "ifl_particles_advect"));
    return launch_particles_advect(c, timestep);
}

// aux arrays of a FluidQuantity: doubles (volume, normals, phi) or bytes (cell, body)
static int aux_lookup(ifl_ctx *c, int field, int which, void **base, int *w, int *h, int *pitch, int *elsize) {
    if (c->version < 4 || check_field(c, field) != IFL_OK) {
        set_error("no solid-body arrays for field %d in chapter %d", field, c->version);
        return IFL_E_ARG;
    }
    Field &f = c->fd[field];
    const Arr *a = nullptr;
    switch (which) {
    case IFL_AUX_VOLUME: a = &f.volume; break;
    case IFL_AUX_NORMAL_X: a = &f.normalX; break;
    case IFL_AUX_NORMAL_Y: a = &f.normalY; break;
    case IFL_AUX_PHI: a = &f.phi; break;
    case IFL_AUX_CELL: *base = f.cell; break;
    case IFL_AUX_BODY: *base = f.body; break;
    default: set_error("bad aux id %d", which); return IFL_E_ARG;
    }
    if (a) {
        *base = a->p;
        *w = a->w;
        *h = a->h;
        *pitch = a->pitch;
        *elsize = 8;
    } else {
        *w = f.w;
        *h = f.h;
        *pitch = f.src.pitch;
        *elsize = 1;
    }
    return IFL_OK;
}

size_t ifl_aux_elems(const ifl_ctx *c, int field, int which) {
    void *b;
    int w, h, p, e;
    if (!c || aux_lookup(const_cast<ifl_ctx *>(c), field, which, &b, &w, &h, &p, &e) != IFL_OK) return 0;
    return (size_t)w * h;
}

int ifl_aux_download(ifl_ctx *c, int field, int which, void *host) {
    CHECK_CTX(c);
    void *b;
    int w, h, p, e;
    TRY(aux_lookup(c, field, which, &b, &w, &h, &p, &e));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    TRY(dist_host_barrier(c)); // several ranks: collective, like ifl_download
    IFL_CUDA(cudaMemcpy2DAsync(host, (size_t)w * e, b, (size_t)p * e, (size_t)w * e, h, cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return dist_host_barrier(c);
}

int ifl_aux_upload(ifl_ctx *c, int field, int which, const void *host) {
    CHECK_CTX(c);
    void *b;
    int w, h, p, e;
    TRY(aux_lookup(c, field, which, &b, &w, &h, &p, &e));
    if (c->world > 1) {
        set_error("ifl_aux_upload is a one-GPU test hook");
        return IFL_E_ARG;
    }
    IFL_CUDA(cudaMemcpy2DAsync(b, (size_t)p * e, host, (size_t)w * e, (size_t)w * e, h, cudaMemcpyHostToDevice, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    if (field == IFL_FIELD_D && which == IFL_AUX_CELL) { // keep the factorisation's fluid mask in step
        const uint8_t *cell = (const uint8_t *)host;
        double *tmp = (double *)malloc((size_t)w * h * sizeof(double));
        if (!tmp) return IFL_E_NOMEM;
        for (size_t i = 0; i < (size_t)w * h; i++) tmp[i] = cell[i] == CELL_FLUID ? 1.0 : 0.0;
        cudaError_t e2 = cudaMemcpy2D(c->fmask.p, (size_t)c->fmask.pitch * 8, tmp, (size_t)w * 8, (size_t)w * 8, h,
                                      cudaMemcpyHostToDevice);
        free(tmp);
        IFL_CUDA(e2);
    }
    return IFL_OK;
}

// ---- FluidSolver private hot-path methods ----------------------------------------
static int need_pcg(ifl_ctx *c, const char *what) {
    if (!pcg_chapter(c)) {
        set_error("%s: chapters 1-2 have no explicit matrix / PCG (use ifl_project_gs)", what);
        return IFL_E_ARG;
    }
    return IFL_OK;
}

static Arr *vec_arr(ifl_ctx *c, int buf) {
    switch (buf) {
    case IFL_BUF_R: case IFL_BUF_P: case IFL_BUF_Z: case IFL_BUF_S: return buf_arr(c, buf);
    }
    set_error("buffer id %d is not one of the PCG vectors r, p, z, s", buf);
    return nullptr;
}

int ifl_build_rhs(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(dist_barrier(c));
    return launch_build_rhs(c);
}

int ifl_build_pressure_matrix(ifl_ctx *c, double timestep, double density) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_build_pressure_matrix"));
    TRY(dist_barrier(c));
    return launch_build_matrix(c, timestep, density);
}

int ifl_build_preconditioner(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_build_preconditioner"));
    TRY(dist_barrier(c));
    TRY(launch_mic0_factor(c));
    return check_watchdog(c, "ifl_build_preconditioner");
}

int ifl_apply_preconditioner(ifl_ctx *c, int dst, int a) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_apply_preconditioner"));
    Arr *d = vec_arr(c, dst), *s = vec_arr(c, a);
    if (!d || !s || d == s) {
        if (d == s && d) set_error("ifl_apply_preconditioner: dst and a must differ");
        return IFL_E_ARG;
    }
    TRY(dist_barrier(c));
    TRY(launch_precon_forward(c, *d, *s, false));
    TRY(launch_precon_backward(c, *d, *s, false, false));
    return check_watchdog(c, "ifl_apply_preconditioner");
}

int ifl_matrix_vector_product(ifl_ctx *c, int dst, int b) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_matrix_vector_product"));
    Arr *d = vec_arr(c, dst), *s = vec_arr(c, b);
    if (!d || !s || d == s) {
        if (d == s && d) set_error("ifl_matrix_vector_product: dst and b must differ");
        return IFL_E_ARG;
    }
    TRY(dist_barrier(c));
    return launch_matvec(c, *d, *s, false);
}

static int fetch_result(ifl_ctx *c, bool is_max, double *result) {
    double *dev_out = &c->scal->beta; // scratch slot; no solve is in flight during granular calls
    TRY(launch_finish_reduce(c, is_max, dev_out));
    IFL_CUDA(cudaMemcpyAsync(c->result_h, dev_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    *result = c->result_h[0];
    return IFL_OK;
}

int ifl_dot_product(ifl_ctx *c, int a, int b, double *result) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_dot_product"));
    Arr *x = vec_arr(c, a), *y = vec_arr(c, b);
    if (!x || !y || !result) return IFL_E_ARG;
    TRY(launch_dot(c, *x, *y));
    return fetch_result(c, false, result);
}

int ifl_scaled_add(ifl_ctx *c, int dst, int a, int b, double s) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_scaled_add"));
    Arr *d = vec_arr(c, dst), *x = vec_arr(c, a), *y = vec_arr(c, b);
    if (!d || !x || !y) return IFL_E_ARG;
    return launch_scaled_add(c, *d, *x, *y, s);
}

int ifl_infinity_norm(ifl_ctx *c, int a, double *result) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_infinity_norm"));
    Arr *x = vec_arr(c, a);
    if (!x || !result) return IFL_E_ARG;
    TRY(launch_inf_norm(c, *x));
    return fetch_result(c, true, result);
}

int ifl_project(ifl_ctx *c, int limit, ifl_solve_info *info) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_project"));
    if (limit < 0) {
        set_error("ifl_project: negative limit");
        return IFL_E_ARG;
    }
    return pcg_project(c, limit, info);
}

int ifl_project_gs(ifl_ctx *c, int limit, double timestep, double density, ifl_solve_info *info) {
    CHECK_CTX(c);
    if (pcg_chapter(c)) {
        set_error("ifl_project_gs: Gauss-Seidel projection belongs to chapters 1-2");
        return IFL_E_ARG;
    }
    return gs_project(c, limit, timestep, density, info);
}

int ifl_apply_pressure(ifl_ctx *c, double timestep, double density) {
    CHECK_CTX(c);
    TRY(dist_barrier(c));
    return launch_apply_pressure(c, timestep, density);
}

// ---- FluidSolver public surface ------------------------------------------------------
int ifl_add_inflow(ifl_ctx *c, double x, double y, double w, double h, double d, double u, double v) {
    CHECK_CTX(c);
    TRY(launch_add_inflow(c, IFL_FIELD_D, x, y, x + w, y + h, d)); // v3:450-452
    TRY(launch_add_inflow(c, IFL_FIELD_U, x, y, x + w, y + h, u));
    TRY(launch_add_inflow(c, IFL_FIELD_V, x, y, x + w, y + h, v));
    return IFL_OK;
}

// FluidSolver::update of chapters 4-5 (v5:927-953, v4:869-895)
// Row-slab multi-GPU: `B` = dist_barrier, a no-op on one GPU, placed wherever the next stage
// reads rows of a neighbouring slab that the previous stage wrote (or overwrites rows a
// neighbour may still be reading).
#define B TRY(dist_barrier(c))
static int update_solids(ifl_ctx *c, double timestep, double density, ifl_solve_info *info) {
    const int fields[3] = {IFL_FIELD_D, IFL_FIELD_U, IFL_FIELD_V};
    B;
    for (int i = 0; i < 3; i++) TRY(launch_fill_solid_fields(c, fields[i]));
    B;
    TRY(launch_set_boundary_condition(c));
    B;
    TRY(launch_build_rhs(c));
    TRY(launch_build_matrix(c, timestep, density));
    B;
    TRY(launch_mic0_factor(c));
    TRY(pcg_project(c, 2000, info));
    TRY(launch_apply_pressure(c, timestep, density));
    B;
    for (int i = 0; i < 3; i++) TRY(launch_extrapolate(c, fields[i]));
    B;
    TRY(launch_set_boundary_condition(c));
    B;
    for (int i = 0; i < 3; i++) TRY(launch_advect(c, fields[i], timestep));
    for (int i = 0; i < 3; i++) TRY(ifl_flip(c, fields[i]));
    return IFL_OK;
}

static int copy_own_rows(ifl_ctx *c, const Arr &dst, const Arr &src) {
    IFL_CUDA(cudaMemcpyAsync((char *)dst.p + dst.own_begin(), (const char *)src.p + src.own_begin(),
                             src.own_end() - src.own_begin(), cudaMemcpyDeviceToDevice, c->stream));
    return IFL_OK;
}

// FluidSolver::update of chapters 6-7 (v7:995-1034, v6:966-1008): heat solve, buoyancy,
// (face densities,) pressure solve, advection of d, t, u, v.  infos[0] = heat solve,
// infos[1] = pressure solve.
static int update_heat(ifl_ctx *c, double timestep, ifl_solve_info *infos) {
    const int fields[4] = {IFL_FIELD_D, IFL_FIELD_T, IFL_FIELD_U, IFL_FIELD_V};
    ifl_solve_info local[2];
    if (!infos) infos = local;
    B;
    for (int i = 0; i < 4; i++) TRY(launch_fill_solid_fields(c, fields[i]));
    Arr &tsrc = c->fd[IFL_FIELD_T].src;
    TRY(copy_own_rows(c, c->r, tsrc)); // v7:1001
    B;
    TRY(launch_build_heat_matrix(c, timestep));
    B;
    TRY(launch_mic0_factor(c));
    TRY(pcg_project(c, 2000, &infos[0]));
    TRY(copy_own_rows(c, tsrc, c->p)); // v7:1005
    B;
    TRY(launch_extrapolate(c, IFL_FIELD_T));
    B;
    TRY(launch_add_buoyancy(c, timestep));
    B;
    TRY(launch_set_boundary_condition(c));
    B;
    TRY(launch_build_rhs(c));
    if (c->version >= 7) TRY(launch_compute_densities(c));
    B;
    TRY(launch_build_matrix(c, timestep, c->rho_air));
    B;
    TRY(launch_mic0_factor(c));
    TRY(pcg_project(c, 2000, &infos[1]));
    TRY(launch_apply_pressure(c, timestep, c->rho_air));
    B;
    TRY(launch_extrapolate(c, IFL_FIELD_D));
    TRY(launch_extrapolate(c, IFL_FIELD_U));
    TRY(launch_extrapolate(c, IFL_FIELD_V));
    B;
    TRY(launch_set_boundary_condition(c));
    B;
    for (int i = 0; i < 4; i++) TRY(launch_advect(c, fields[i], timestep));
    for (int i = 0; i < 4; i++) TRY(ifl_flip(c, fields[i]));
    return IFL_OK;
}
// FluidSolver::update of chapter 8 (v8:1350-1413): particles -> grid, the chapter-7 solves between copy() and
// diff(), grid -> particles, particle advection.  infos[0] = heat solve, infos[1] = pressure solve.
static int update_flip(ifl_ctx *c, double timestep, ifl_solve_info *infos) {
    const int fields[4] = {IFL_FIELD_D, IFL_FIELD_T, IFL_FIELD_U, IFL_FIELD_V};
    const double flip_alpha = 0.001; // _flipAlpha v8:1296
    ifl_solve_info local[2];
    if (!infos) infos = local;
    for (int i = 0; i < 4; i++) TRY(launch_fill_solid_fields(c, fields[i]));
    TRY(flip_particles_to_grid(c, nullptr));
    for (int i = 0; i < 4; i++) TRY(launch_copy(c, fields[i]));
    // the inflow lives INSIDE update in this chapter (v8:1368)
    TRY(ifl_add_inflow_t(c, 0.45, 0.2, 0.2, 0.05, 1.0, c->t_amb, 0.0, 0.0));
    Arr &tsrc = c->fd[IFL_FIELD_T].src;
    TRY(copy_own_rows(c, c->r, tsrc)); // v8:1370
    TRY(launch_build_heat_matrix(c, timestep));
    TRY(launch_mic0_factor(c));
    TRY(pcg_project(c, 2000, &infos[0]));
    TRY(copy_own_rows(c, tsrc, c->p)); // v8:1374
    TRY(flip_extrapolate(c, IFL_FIELD_T));
    TRY(launch_add_buoyancy(c, timestep));
    TRY(launch_set_boundary_condition(c));
    TRY(launch_build_rhs(c));
    TRY(launch_compute_densities(c));
    TRY(launch_build_matrix(c, timestep, c->rho_air));
    TRY(launch_mic0_factor(c));
    TRY(pcg_project(c, 2000, &infos[1]));
    TRY(launch_apply_pressure(c, timestep, c->rho_air));
    TRY(flip_extrapolate(c, IFL_FIELD_D));
    TRY(flip_extrapolate(c, IFL_FIELD_U));
    TRY(flip_extrapolate(c, IFL_FIELD_V));
    TRY(launch_set_boundary_condition(c));
    for (int i = 0; i < 4; i++) TRY(launch_diff(c, fields[i], flip_alpha, 0));
    TRY(launch_grid_to_particles(c, flip_alpha));
    for (int i = 0; i < 4; i++) TRY(launch_diff(c, fields[i], flip_alpha, 1));
    return launch_particles_advect(c, timestep);
}
#undef B

static int update_impl(ifl_ctx *c, double timestep, double density, ifl_solve_info *infos) {
    ifl_solve_info local;
    ifl_solve_info *info = infos ? infos : &local;
    if (c->version >= 8) return update_flip(c, timestep, infos);
    if (c->version >= 6) return update_heat(c, timestep, infos);
    if (c->version >= 4) return update_solids(c, timestep, density, info);
    // Row-slab multi-GPU: dist_barrier (a no-op on one GPU) sits wherever the next kernel reads
    // rows of a neighbouring slab that the previous kernels wrote; inside the solves the
    // reductions' folds and the sweeps' hand-off messages order the ranks.
    TRY(dist_barrier(c)); // u, v of the previous step (advection, inflow)
    TRY(launch_build_rhs(c));
    if (pcg_chapter(c)) { // v3:433-447
        TRY(launch_build_matrix(c, timestep, density));
        TRY(dist_barrier(c)); // aPlusX/aPlusY of the upstream slab's last row
        TRY(launch_mic0_factor(c));
        TRY(pcg_project(c, 600, info));
    } else { // v2:320-332, v1:284-297
        TRY(gs_project(c, 600, timestep, density, info));
    }
    TRY(launch_apply_pressure(c, timestep, density));
    TRY(dist_barrier(c)); // back-traced samples of u, v, d may lie in any slab
    TRY(launch_advect(c, IFL_FIELD_D, timestep));
    TRY(launch_advect(c, IFL_FIELD_U, timestep));
    TRY(launch_advect(c, IFL_FIELD_V, timestep));
    TRY(ifl_flip(c, IFL_FIELD_D));
    TRY(ifl_flip(c, IFL_FIELD_U));
    TRY(ifl_flip(c, IFL_FIELD_V));
    return IFL_OK;
}

int ifl_update(ifl_ctx *c, double timestep, double density, ifl_solve_info *infos) {
    CHECK_CTX(c);
    TRY(update_impl(c, timestep, density, infos));
    // one stream synchronisation per step: a timed-out stage barrier or factorisation sweep must not
    // go unnoticed (the solves check the same sticky word in their own read-backs)
    return check_watchdog(c, "ifl_update");
}

// slab != 0: the host buffers hold only this rank's rows (row ry0 of each array first)
static int update_host_impl(ifl_ctx *c, double timestep, double density, double *d, double *u, double *v,
                            ifl_solve_info *infos, int slab) {
    if (!d || !u || !v) {
        set_error("ifl_update_host: null host buffer");
        return IFL_E_ARG;
    }
    double *host[3] = {d, u, v};
    const int ids[3] = {IFL_FIELD_D, IFL_FIELD_U, IFL_FIELD_V};
    // every rank moves the rows of its own slab (one GPU: everything)
    for (int i = 0; i < 3; i++) {
        Arr &a = c->fd[ids[i]].src;
        const double *h = host[i] + (slab ? 0 : (size_t)a.ry0 * a.w);
        IFL_CUDA(cudaMemcpy2DAsync(a.p + (size_t)a.ry0 * a.pitch, (size_t)a.pitch * 8, h, (size_t)a.w * 8,
                                   (size_t)a.w * 8, a.ry1 - a.ry0, cudaMemcpyHostToDevice, c->stream));
    }
    TRY(ifl_update(c, timestep, density, infos));
    for (int i = 0; i < 3; i++) {
        Arr &a = c->fd[ids[i]].src;
        double *h = host[i] + (slab ? 0 : (size_t)a.ry0 * a.w);
        IFL_CUDA(cudaMemcpy2DAsync(h, (size_t)a.w * 8, a.p + (size_t)a.ry0 * a.pitch, (size_t)a.pitch * 8,
                                   (size_t)a.w * 8, a.ry1 - a.ry0, cudaMemcpyDeviceToHost, c->stream));
    }
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

int ifl_update_host(ifl_ctx *c, double timestep, double density, double *d, double *u, double *v,
                    ifl_solve_info *infos) {
    CHECK_CTX(c);
    return update_host_impl(c, timestep, density, d, u, v, infos, 0);
}

int ifl_update_host_slab(ifl_ctx *c, double timestep, double density, double *d, double *u, double *v,
                         ifl_solve_info *infos) {
    CHECK_CTX(c);
    return update_host_impl(c, timestep, density, d, u, v, infos, 1);
}

// Slab-local transfers (not collective): `host` holds the caller's rows only.
int ifl_upload_slab(ifl_ctx *c, int buf, const double *host) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p || !host) {
        set_error("ifl_upload_slab: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    if (buf == IFL_BUF_ADIAG || buf == IFL_BUF_APLUSX || buf == IFL_BUF_APLUSY) c->matrix_uniform = 0; // the caller's own matrix
    IFL_CUDA(cudaMemcpy2DAsync(a->p + (size_t)a->ry0 * a->pitch, (size_t)a->pitch * 8, host, (size_t)a->w * 8,
                               (size_t)a->w * 8, a->ry1 - a->ry0, cudaMemcpyHostToDevice, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

int ifl_download_slab(ifl_ctx *c, int buf, double *host) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p || !host) {
        set_error("ifl_download_slab: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    IFL_CUDA(cudaMemcpy2DAsync(host, (size_t)a->w * 8, a->p + (size_t)a->ry0 * a->pitch, (size_t)a->pitch * 8,
                               (size_t)a->w * 8, a->ry1 - a->ry0, cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

size_t ifl_slab_elems(const ifl_ctx *c, int buf) {
    Arr *a = c ? buf_arr(const_cast<ifl_ctx *>(c), buf) : nullptr;
    return (a && a->p) ? (size_t)a->w * (a->ry1 - a->ry0) : 0;
}

} // extern "C"
