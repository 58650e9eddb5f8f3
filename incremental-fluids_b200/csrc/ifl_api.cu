// ifl_api.cu -- the C ABI of libifl_b200.so (include/ifl_b200.h): context lifetime,
// dense<->pitched data movement, and the per-chapter update() sequences.
// Everything here is host code; all arithmetic happens in the kernels.
#include "ifl_internal.cuh"

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

static int alloc_arr(Arr &a, int w, int h) {
    a.w = w;
    a.h = h;
    a.pitch = round_up(w, TILE);
    a.rows = round_up(h, TILE) + TILE;
    a.p = nullptr;
    IFL_CUDA(cudaMalloc(&a.p, a.bytes()));
    IFL_CUDA(cudaMemset(a.p, 0, a.bytes())); // uninitialised == zero page
    return IFL_OK;
}

static void free_arr(Arr &a) {
    if (a.p) cudaFree(a.p);
    a.p = nullptr;
}

static int alloc_field(Field &f, int w, int h, double ox, double oy) {
    f.w = w;
    f.h = h;
    f.ox = ox;
    f.oy = oy;
    int rc = alloc_arr(f.src, w, h);
    if (rc == IFL_OK) rc = alloc_arr(f.dst, w, h);
    return rc;
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
    }
    return nullptr;
}

static bool pcg_chapter(const ifl_ctx *c) { return c->version >= 3; }

} // namespace ifl

using namespace ifl;

#define CHECK_CTX(c)                       \
    do {                                   \
        if (!(c)) {                        \
            set_error("null context");     \
            return IFL_E_ARG;              \
        }                                  \
    } while (0)
#define TRY(expr)                          \
    do {                                   \
        int rc_ = (expr);                  \
        if (rc_ != IFL_OK) return rc_;     \
    } while (0)

extern "C" {

const char *ifl_last_error(void) { return g_err; }

int ifl_create(ifl_ctx **out, int w, int h, int version, int device) {
    if (!out || w < 2 || h < 2 || version < 1 || version > 8) {
        set_error("ifl_create: bad argument (w=%d h=%d version=%d)", w, h, version);
        return IFL_E_ARG;
    }
    if (version > 3) {
        set_error("ifl_create: chapter %d (solid bodies and later) is not available in this build", version);
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
        // v3:404-406: d at cell centres, u/v on the staggered faces
        if ((rc = alloc_field(c->fd[IFL_FIELD_D], w, h, 0.5, 0.5)) != IFL_OK) break;
        if ((rc = alloc_field(c->fd[IFL_FIELD_U], w + 1, h, 0.0, 0.5)) != IFL_OK) break;
        if ((rc = alloc_field(c->fd[IFL_FIELD_V], w, h + 1, 0.5, 0.0)) != IFL_OK) break;
        Arr *cells[] = {&c->r, &c->p, &c->z, &c->s, &c->q, &c->precon, &c->aDiag, &c->aPlusX, &c->aPlusY, &c->cx, &c->cy};
        const int ncells = pcg_chapter(c) ? 11 : 2; // chapters 1-2 only own _r and _p (v2:219-220)
        for (int i = 0; i < ncells && rc == IFL_OK; i++) rc = alloc_arr(*cells[i], w, h);
        if (rc != IFL_OK) break;
        const size_t npart = (size_t)((w + 255) / 256) * ((h + 15) / 16) + 64;
        if (cudaMalloc(&c->partials, (npart > MAX_PARTIALS ? npart : MAX_PARTIALS) * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&c->scal, sizeof(SolveScalars)) != cudaSuccess ||
            cudaMemset(c->scal, 0, sizeof(SolveScalars)) != cudaSuccess ||
            cudaMallocHost(&c->scal_h, 2 * sizeof(SolveScalars)) != cudaSuccess ||
            cudaMallocHost(&c->result_h, 8 * sizeof(double)) != cudaSuccess) {
            set_error("ifl_create: scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = IFL_E_CUDA;
            break;
        }
        if ((rc = sweep_init(c)) != IFL_OK) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            set_error("ifl_create: %s", cudaGetErrorString(cudaGetLastError()));
            rc = IFL_E_CUDA;
        }
    } while (0);
    if (rc != IFL_OK) {
        ifl_destroy(c);
        return rc;
    }
    *out = c;
    return IFL_OK;
}

int ifl_destroy(ifl_ctx *c) {
    if (!c) return IFL_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 4; i++) {
        free_arr(c->fd[i].src);
        free_arr(c->fd[i].dst);
    }
    Arr *cells[] = {&c->r, &c->p, &c->z, &c->s, &c->q, &c->precon, &c->aDiag, &c->aPlusX, &c->aPlusY, &c->cx, &c->cy};
    for (int i = 0; i < 11; i++) free_arr(*cells[i]);
    if (c->partials) cudaFree(c->partials);
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
    return c->n_strips;
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
    IFL_CUDA(cudaMemcpy2DAsync(a->p, (size_t)a->pitch * 8, host, (size_t)a->w * 8, (size_t)a->w * 8, a->h,
                               cudaMemcpyHostToDevice, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

int ifl_download(ifl_ctx *c, int buf, double *host) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p || !host) {
        set_error("ifl_download: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    IFL_CUDA(cudaMemcpy2DAsync(host, (size_t)a->w * 8, a->p, (size_t)a->pitch * 8, (size_t)a->w * 8, a->h,
                               cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

int ifl_fill(ifl_ctx *c, int buf, double value) {
    CHECK_CTX(c);
    Arr *a = buf_arr(c, buf);
    if (!a || !a->p) {
        set_error("ifl_fill: bad buffer id %d", buf);
        return IFL_E_ARG;
    }
    if (value == 0.0 && !signbit(value)) {
        IFL_CUDA(cudaMemsetAsync(a->p, 0, a->bytes(), c->stream));
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
    return launch_advect(c, field, timestep);
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
    return launch_build_rhs(c);
}

int ifl_build_pressure_matrix(ifl_ctx *c, double timestep, double density) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_build_pressure_matrix"));
    return launch_build_matrix(c, timestep, density);
}

int ifl_build_preconditioner(ifl_ctx *c) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_build_preconditioner"));
    return launch_mic0_factor(c);
}

int ifl_apply_preconditioner(ifl_ctx *c, int dst, int a) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_apply_preconditioner"));
    Arr *d = vec_arr(c, dst), *s = vec_arr(c, a);
    if (!d || !s || d == s) {
        if (d == s && d) set_error("ifl_apply_preconditioner: dst and a must differ");
        return IFL_E_ARG;
    }
    TRY(launch_precon_forward(c, *d, *s, false));
    TRY(launch_precon_backward(c, *d, *s, false, false));
    IFL_CUDA(cudaMemcpyAsync(c->scal_h, c->scal, sizeof(SolveScalars), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    if (c->scal_h[0].watchdog) {
        set_error("ifl_apply_preconditioner: wavefront dependency watchdog fired");
        return IFL_E_WATCHDOG;
    }
    return IFL_OK;
}

int ifl_matrix_vector_product(ifl_ctx *c, int dst, int b) {
    CHECK_CTX(c);
    TRY(need_pcg(c, "ifl_matrix_vector_product"));
    Arr *d = vec_arr(c, dst), *s = vec_arr(c, b);
    if (!d || !s || d == s) {
        if (d == s && d) set_error("ifl_matrix_vector_product: dst and b must differ");
        return IFL_E_ARG;
    }
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

int ifl_update(ifl_ctx *c, double timestep, double density, ifl_solve_info *infos) {
    CHECK_CTX(c);
    ifl_solve_info local;
    ifl_solve_info *info = infos ? infos : &local;
    TRY(launch_build_rhs(c));
    if (pcg_chapter(c)) { // v3:433-447
        TRY(launch_build_matrix(c, timestep, density));
        TRY(launch_mic0_factor(c));
        TRY(pcg_project(c, 600, info));
    } else { // v2:320-332, v1:284-297
        TRY(gs_project(c, 600, timestep, density, info));
    }
    TRY(launch_apply_pressure(c, timestep, density));
    TRY(launch_advect(c, IFL_FIELD_D, timestep));
    TRY(launch_advect(c, IFL_FIELD_U, timestep));
    TRY(launch_advect(c, IFL_FIELD_V, timestep));
    TRY(ifl_flip(c, IFL_FIELD_D));
    TRY(ifl_flip(c, IFL_FIELD_U));
    TRY(ifl_flip(c, IFL_FIELD_V));
    return IFL_OK;
}

int ifl_update_host(ifl_ctx *c, double timestep, double density, double *d, double *u, double *v,
                    ifl_solve_info *infos) {
    CHECK_CTX(c);
    if (!d || !u || !v) {
        set_error("ifl_update_host: null host buffer");
        return IFL_E_ARG;
    }
    double *host[3] = {d, u, v};
    const int ids[3] = {IFL_FIELD_D, IFL_FIELD_U, IFL_FIELD_V};
    for (int i = 0; i < 3; i++) {
        Arr &a = c->fd[ids[i]].src;
        IFL_CUDA(cudaMemcpy2DAsync(a.p, (size_t)a.pitch * 8, host[i], (size_t)a.w * 8, (size_t)a.w * 8, a.h,
                                   cudaMemcpyHostToDevice, c->stream));
    }
    TRY(ifl_update(c, timestep, density, infos));
    for (int i = 0; i < 3; i++) {
        Arr &a = c->fd[ids[i]].src;
        IFL_CUDA(cudaMemcpy2DAsync(host[i], (size_t)a.w * 8, a.p, (size_t)a.pitch * 8, (size_t)a.w * 8, a.h,
                                   cudaMemcpyDeviceToHost, c->stream));
    }
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

} // extern "C"
