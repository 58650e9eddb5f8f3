// ifl_internal.cuh -- shared declarations of libifl_b200 (device context, array
// descriptors, exact-arithmetic helpers, deterministic block reductions).
//
// Arithmetic contract: every kernel in this library must reproduce the reference's
// double-precision results operation by operation.  The whole library is compiled
// with -fmad=false (no contraction), and helpers below restate std::min/std::max
// with the libstdc++ tie/NaN/-0.0 behaviour instead of fmin/fmax.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ifl_b200.h"
#include "solid_geometry.cuh" // BodyDev + SolidBox / SolidSphere geometry (shared with the host drop-in header)

namespace ifl {

// ------------------------------------------------------------------ layout ----
// Every array lives in HBM as a pitched 2-D block: logical w x h, physical
// pitch = round_up(w, 32) elements, rows = round_up(h, 32) + 32.  Pad cells are
// zero forever (kernels never store outside x<w, y<h), which lets tile loaders
// (cp.async.bulk needs 16-byte aligned 16-byte multiples) fetch full 32-wide tiles
// at any grid size.  The reference's dense x + y*w layout exists only on the host
// side of ifl_upload/ifl_download.
constexpr int TILE = 32;

struct Arr {
    double *p;
    int w, h, pitch, rows;
    // Row slab of this array that THIS rank computes: [ry0, ry1).  One GPU: [0, h).  With
    // row-slab multi-GPU (dist.cu) `p` addresses the whole array in a virtual range shared by
    // all ranks (each slab backed by its owner's HBM), kernels launch over the slab only and
    // reach the neighbours' rows through NVLink peer mappings.
    int ry0, ry1;
    __host__ __device__ size_t bytes() const { return (size_t)pitch * rows * sizeof(double); }
    // byte range [own_begin, own_end) of the rows this rank zero-fills / copies (the last
    // slab includes the pad rows below h)
    __host__ __device__ size_t own_begin() const { return (size_t)ry0 * pitch * sizeof(double); }
    __host__ __device__ size_t own_end() const { return (size_t)(ry1 == h ? rows : ry1) * pitch * sizeof(double); }
};

struct Field { // one FluidQuantity (v3:41-49; solid-body members v5:288-313)
    Arr src, dst;
    int w, h;
    double ox, oy;
    // chapters 4+: solid fields, same pitch as src (phi is (w+1) x (h+1) with its own pitch)
    Arr volume, normalX, normalY, phi;
    uint8_t *cell, *body, *mask; // byte arrays, row pitch == src.pitch
    int *solid_list;             // compacted indices (x + y*pitch) of interior non-fluid cells
    int *solid_count;            // device counter for solid_list
};

constexpr int MAX_BODIES = 256; // FluidQuantity::_body is uint8_t (v4:253)

// Device-resident scalars of one PCG / Gauss-Seidel solve (no host round trips
// inside the loop; the host only reads `done`/`iter` through a pinned mirror).
struct SolveScalars {
    double sigma;     // z.r          v3:358
    double alpha;     // sigma/(z.s)  v3:362
    double beta;      // sigmaNew/sigma v3:375
    double max_error; // |r|inf       v3:366
    int iter;         // zero-based iteration counter (what v3:368 prints)
    int done;         // 0 running, 1 converged, 2 initial-small
    // ---- everything above is reset at the start of a solve; the watchdog word is STICKY: set by
    // any dependency wait that timed out (sweep hand-off, rank barrier), cleared only by the host
    // after it has reported IFL_E_WATCHDOG (check_watchdog, ifl_api.cu)
    int watchdog;
    int pad;
};

constexpr int MAX_PARTIALS = 8192; // per-block partial results of one reduction

// ---- row-slab multi-GPU (dist.cu) ------------------------------------------------
constexpr int MAX_WORLD = 8;
#define IFL_MAX_DEVICES 64
// What a kernel needs for an in-kernel barrier across the ranks of one solver: every rank
// owns one flag word per peer (in its own HBM, peers store into it over NVLink).
struct DistDev {
    int rank, world;
    unsigned long long *epoch;                 // this rank's barrier counter (device memory)
    unsigned long long *flags_local;           // [MAX_WORLD] written by the peers
    unsigned long long *flags_peer[MAX_WORLD]; // flags_local of every rank (peer mappings)
    int *watchdog;                             // SolveScalars::watchdog of this rank
};
struct DistState; // host side (dist.cu)

} // namespace ifl

struct ifl_ctx {
    int W, H, version, device;
    double hx;
    cudaStream_t stream;
    ifl::Field fd[4]; // d, u, v, t
    ifl::Arr r, p, z, s, q, precon, aDiag, aPlusX, aPlusY, cx, cy;
    ifl::Arr s2;                       // chapter 3: the other half of the search direction's ping-pong pair (k_xpay_matvec)
    int fuse_xpay, xpay_pending, xpay_fused; // IFL_FUSE_XPAY; s = z + beta s waits for the next matvec; fused launches of this solve
    ifl::Arr pe;    // chapters 4+: precon with +0.0 at non-fluid cells (what the masked sweeps multiply by)
    ifl::Arr fmask; // chapters 4+: 1.0 at fluid cells of _d, 0.0 elsewhere (operand of the masked factorisation)
    ifl::Arr uDensity, vDensity; // chapter 7+: densities on the staggered faces (v7:598-599)
    double rho_air, rho_soot, diffusion; // chapters 6+ ctor arguments (v6:921)
    double t_amb, g;                      // 294.0, 9.81 (v6:932-933)
    ifl::BodyDev *bodies_d; // [MAX_BODIES]
    int n_bodies;
    int *ext_ready;         // extrapolate(): device counter of cells resolved in the last batch of rounds
    void *particles;        // chapter 8: ifl::ParticleSet (flip_kernels.cu)
    double *partials;          // [MAX_PARTIALS] block partials (sum or max)
    int n_partials;            // valid entries written by the last reducing kernel
    ifl::SolveScalars *scal;   // device
    ifl::SolveScalars *scal_h; // pinned host mirror
    double *result_h;          // pinned 8 doubles for scalar results
    // wavefront sweep plumbing (sweep_kernels.cu)
    unsigned long long *handoff; // [strips][pitch] x {value bits, epoch} (this rank's array)
    void *handoff_base;          // allocation holding every rank's hand-off array
    unsigned long long *ticket;  // strip ticket counter
    unsigned long long epoch;    // last used handoff epoch
    int n_strips;
    unsigned long long sweep_launches; // sweeps launched so far
    unsigned long long sweep_tickets;  // cluster tickets handed out by all previous sweeps
    int matrix_uniform;                // chapter 3: the stored matrix is the one buildPressureMatrix wrote (k_matvec may evaluate it)
    int matvec_uniform_allowed;        // IFL_MATVEC_UNIFORM=0 turns that off
    double matrix_scale;               // its `scale` (v3:223)
    int sweep_cluster;                 // thread-block cluster size of the sweep kernels
    int tri_cluster16;                 // tri_kernels.cu may use clusters of 16 when all strips are resident (IFL_TRI_CLUSTER16)
    int tri_engine;                    // triangular solves: 2 stair_kernels.cu (default), 1 tri_kernels.cu, 0 the one-row engine
    // overlap of k_axpy2_norm with the forward sweep (pcg_kernels.cu): the streaming kernel runs on a side
    // stream and counts finished blocks per 64-row band, the sweep's loader waits for its strip's band
    int overlap_axpy;                  // 0 off, 1 plain streaming kernel, >= 2 persistent streaming kernel (default 3)
    int sm_count;
    cudaStream_t side_stream;
    cudaEvent_t ev_alpha, ev_axpy;     // alpha is final (main -> side), r and |r| partials are final (side -> main)
    unsigned *band_count;              // [strips of 64 rows] blocks of k_axpy2_norm finished so far (monotonic)
    unsigned band_epoch;               // launches of the overlapped k_axpy2_norm so far
    int sweep_head_delay;              // SM cycles the head strip idles per macro-step (pace-setter, sweep_init)
    // row-slab multi-GPU: world == 1 unless the context came from ifl_create_dist
    int rank, world;
    int ry0, ry1;                // cell rows [ry0, ry1) owned by this rank (multiples of 32, ry1 clipped to H)
    ifl::DistState *dist;        // rendezvous sockets + peer mappings (null when world == 1)
    ifl::DistDev ddev;           // device-side barrier state (world == 1: barriers are no-ops)
    unsigned long long *handoff_down[2]; // hand-off array of the downstream rank: [0] forward sweeps, [1] backward
    double *partials_buf[2];     // reductions alternate between two partial buffers (see pcg_kernels.cu)
    int partials_sel;
    void *map_cache;                   // TMA tensor maps keyed by array base pointer
    unsigned long long *sweep_times_buf; // [strips][2] diagnostics buffer
    unsigned long long *sweep_times;     // == sweep_times_buf while ifl_debug_sweep_times is armed, else null
    long long launches;
    // per-kernel-class event timing (ifl_profile): ring of (start, stop, class)
    int prof_on;
    int prof_head, prof_count; // oldest pending pair, number of pending pairs
    cudaEvent_t *prof_ev;      // [2 * PROF_RING]
    int *prof_cls;             // [PROF_RING]
    double prof_ms[IFL_K_COUNT_];
    long long prof_n[IFL_K_COUNT_];
};

namespace ifl {

// -------------------------------------------------------------- error plumbing --
void set_error(const char *fmt, ...);
#define IFL_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            ifl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return IFL_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)
#define IFL_LAUNCHED(ctx)                                                                      \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) {                                                               \
            ifl::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return IFL_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------- profiling --
constexpr int PROF_RING = 4096;
void prof_begin(ifl_ctx *c, int cls);
void prof_end(ifl_ctx *c);
struct ProfScope { // brackets one kernel launch with events when profiling is on
    ifl_ctx *c;
    ProfScope(ifl_ctx *ctx, int cls) : c(ctx) {
        if (c->prof_on) prof_begin(c, cls);
    }
    ~ProfScope() {
        if (c->prof_on) prof_end(c);
    }
};

// ------------------------------------------------------------ exact arithmetic --
// std::min(a,b) == (b < a) ? b : a ; std::max(a,b) == (a < b) ? b : a   (libstdc++)
__host__ __device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__host__ __device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__host__ __device__ __forceinline__ int imin(int a, int b) { return (b < a) ? b : a; }
__host__ __device__ __forceinline__ int imax(int a, int b) { return (a < b) ? b : a; }
enum { CELL_FLUID = 0, CELL_SOLID = 1 }; // v4:70-73

// ------------------------------------------------- deterministic block reductions --
// Fixed-shape trees (xor-shuffle inside the warp, then a fixed loop over warps), so a
// given grid configuration always produces the same bits.
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = std_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// All threads must call; result valid in thread 0.  `red` = shared double[32].
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        t = red[0];
        for (int i = 1; i < nw; i++) t = IS_MAX ? std_max(t, red[i]) : (t + red[i]);
    }
    __syncthreads();
    return t;
}
#endif

// ---------------------------------------------------- barrier across the ranks --
// Called by every thread of a block with at least MAX_WORLD threads.  Rank r stores its
// new epoch into flags_local[r] of every rank (NVLink peer stores, release at system
// scope: everything the previous kernels of this stream wrote is visible first) and waits
// until all its own flag words have reached that epoch.  Epochs only grow, every rank runs
// the same sequence of barriers, so a fast rank may already be one barrier ahead.
#ifdef __CUDACC__
__device__ __forceinline__ void dist_barrier_block(const DistDev &d) {
    if (d.world <= 1) return;
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = *d.epoch + 1;
        *d.epoch = s_epoch;
    }
    __syncthreads();
    const unsigned long long e = s_epoch;
    if ((int)threadIdx.x < d.world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(d.flags_peer[threadIdx.x] + d.rank), "l"(e) : "memory");
        unsigned long long v = 0, t_first = 0;
        unsigned n = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(d.flags_local + threadIdx.x) : "memory");
            if (v >= e) break;
            if ((++n & 1023u) == 0) { // a peer died: raise the watchdog instead of hanging the GPU (2 s, by the clock)
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t_first == 0) t_first = now;
                if (now - t_first > 2000000000ull) {
                    *d.watchdog = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
}
#endif

// Reductions alternate between two partial buffers: with several ranks a fast rank may
// start the next reducing kernel while a slow one still folds the previous partials (the
// barrier sits at the START of the fold), but it cannot get two reductions ahead.
static inline double *partials_next(ifl_ctx *c) {
    c->partials_sel ^= 1;
    c->partials = c->partials_buf[c->partials_sel];
    return c->partials;
}

// ------------------------------------------------------------ kernel entry points --
// dist.cu
int dist_init(ifl_ctx *c, int rank, int world, const char *rendezvous);
void dist_free(ifl_ctx *c);
// Array memory: cudaMalloc on one GPU; with world > 1 one virtual range mapped by all ranks.
// rows_per_pitch describes the slab split (bytes per row); by_rank != 0 instead gives every
// rank `bytes` of its own at [base + rank*stride) and returns the stride.
int dist_alloc_rows(ifl_ctx *c, void **out, size_t bytes, size_t row_bytes, int h);
int dist_alloc_per_rank(ifl_ctx *c, void **out, size_t bytes, size_t *stride);
void dist_free_mem(ifl_ctx *c, void *p);
int dist_barrier(ifl_ctx *c, bool gated = false); // stream-ordered barrier across the ranks (no-op when world == 1); gated: skipped once scal->done
int dist_host_barrier(ifl_ctx *c); // host-side barrier through the rendezvous sockets
int dist_host_sum(ifl_ctx *c, long long *v); // host-side all-reduce (sum) of one integer
// ifl_api.cu
int check_watchdog(ifl_ctx *c, const char *where); // synchronises the stream; IFL_E_WATCHDOG (and clears the word) if a wait timed out
// pcg_kernels.cu
int launch_matvec(ifl_ctx *c, const Arr &dst, const Arr &b, bool with_dot);
int launch_dot(ifl_ctx *c, const Arr &a, const Arr &b);
int launch_scaled_add(ifl_ctx *c, const Arr &dst, const Arr &a, const Arr &b, double s);
int launch_inf_norm(ifl_ctx *c, const Arr &a);
int launch_finish_reduce(ifl_ctx *c, bool is_max, double *out_dev);
int pcg_project(ifl_ctx *c, int limit, ifl_solve_info *info);
// sweep_kernels.cu
int sweep_init(ifl_ctx *c);
void sweep_free(ifl_ctx *c);
int launch_mic0_factor(ifl_ctx *c);
int launch_precon_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated);
int launch_precon_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated);
int gs_project(ifl_ctx *c, int limit, double timestep, double density, ifl_solve_info *info);
// tri_kernels.cu: the two-rows-per-lane engine of the triangular solves (default; IFL_TRI=0 selects the
// one-row engine of sweep_kernels.cu for A/B measurements)
int launch_tri_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated, unsigned band_target = 0);
int launch_tri_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated);
// stair_kernels.cu: the staircase engine (cell B one column behind cell A), same contract
int launch_stair_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated, unsigned band_target = 0);
int launch_stair_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated);
// solid_kernels.cu (chapters 4+)
int launch_fill_solid_fields(ifl_ctx *c, int field);
int launch_set_boundary_condition(ifl_ctx *c);
int launch_extrapolate(ifl_ctx *c, int field);
// flip_kernels.cu (chapter 8)
int flip_init(ifl_ctx *c);
void flip_free(ifl_ctx *c);
int flip_set_particles(ifl_ctx *c, long long count, const double *posX, const double *posY, const double *const *props);
int flip_get_particles(ifl_ctx *c, long long *count, double *posX, double *posY, double *const *props);
// flip_book.cu: particle bookkeeping (v8:735-813) and the chapter-8 extrapolation (v8:478-651)
int flip_particles_init(ifl_ctx *c, int avg_per_cell);      // initParticles + gridToParticles(1.0) (ctors v8:877, 1306-1314)
int flip_count_particles(ifl_ctx *c);                       // countParticles v8:754
int flip_prune_particles(ifl_ctx *c);                       // pruneParticles v8:766
int flip_seed_particles(ifl_ctx *c);                        // seedParticles  v8:789
int flip_particles_to_grid(ifl_ctx *c, long long *count);   // particlesToGrid v8:916-927
int flip_extrapolate(ifl_ctx *c, int field);                // FluidQuantity::extrapolate v8:611-651
long long flip_particle_count(const ifl_ctx *c);
long long flip_particle_capacity(const ifl_ctx *c);
int flip_peek(ifl_ctx *c, int what, long long first, long long n, void *host); // test hook: raw slots incl. the stale tail
int launch_from_particles(ifl_ctx *c, int field);
int launch_grid_to_particles(ifl_ctx *c, double alpha);
int launch_copy(ifl_ctx *c, int field);
int launch_diff(ifl_ctx *c, int field, double alpha, int undo);
int launch_particles_advect(ifl_ctx *c, double timestep);
// assembly_kernels.cu
int launch_build_rhs(ifl_ctx *c);
int launch_build_matrix(ifl_ctx *c, double timestep, double density);
int launch_apply_pressure(ifl_ctx *c, double timestep, double density);
int launch_add_inflow(ifl_ctx *c, int field, double x0, double y0, double x1, double y1, double v);
int launch_build_heat_matrix(ifl_ctx *c, double timestep);
int launch_add_buoyancy(ifl_ctx *c, double timestep);
int launch_compute_densities(ifl_ctx *c);
// advect_kernels.cu
int launch_advect(ifl_ctx *c, int field, double timestep);
int launch_max_velocity(ifl_ctx *c);

} // namespace ifl
