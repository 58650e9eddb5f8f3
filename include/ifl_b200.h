/*
 * ifl_b200.h -- C ABI of libifl_b200.so, the B200 (sm_100a) implementation of the
 * per-timestep hot path of tunabrain/incremental-fluids.
 *
 * The reference has no FFI: its boundary is the C++ class surface used by main()
 * (FluidSolver::update / addInflow / toImage, FluidQuantity, SolidBody).  The
 * drop-in classes in incremental-fluids_b200/host/ keep that surface and forward
 * the bodies of the reference's PRIVATE hot-path methods to the entry points
 * below.  Citations: vN:L == /root/reference/N-<chapter>/Fluid.cpp:L.
 *
 * Conventions
 *  - one opaque ifl_ctx per FluidSolver; it owns every device buffer and one stream.
 *  - host code never sees device pointers: arrays are addressed by ifl_buf ids and
 *    moved with ifl_upload / ifl_download in the reference's dense row-major layout
 *    (index x + y*w, v3:113-119); the physical device pitch is hidden.
 *  - every function returns 0 on success, a negative IFL_E_* code otherwise
 *    (ifl_last_error() gives the text).  There is NO CPU fallback: without a CUDA
 *    device every compute entry point fails with IFL_E_CUDA.
 *  - plain C types only.
 */
#ifndef IFL_B200_H
#define IFL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ifl_ctx ifl_ctx;

enum {
    IFL_OK = 0,
    IFL_E_ARG = -1,      /* bad argument / op not available for this chapter */
    IFL_E_CUDA = -2,     /* CUDA runtime error or no device */
    IFL_E_WATCHDOG = -3, /* an in-kernel dependency wait ran out (never expected) */
    IFL_E_NOMEM = -4
};

/* Which chapter's update() semantics the context follows (1..8 == reference dirs). */
typedef enum {
    IFL_V1_MATRIXLESS = 1, /* GS project, Euler+bilinear advect   (1-matrixless)        */
    IFL_V2_ADVECTION = 2,  /* GS project, RK3+Catmull-Rom          (2-better-advection)  */
    IFL_V3_PCG = 3,        /* MIC(0)-PCG                           (3-conjugate-gradients)*/
    IFL_V4_SOLIDS = 4,
    IFL_V5_CURVED = 5,
    IFL_V6_HEAT = 6,
    IFL_V7_VARDENSITY = 7,
    IFL_V8_FLIP = 8
} ifl_version;

/* Buffer ids.  Sizes (in elements, reference layout): cell-sized = w*h;
 * u fields (w+1)*h; v fields w*(h+1).  All double unless noted. */
typedef enum {
    IFL_BUF_D_SRC = 0, /* FluidQuantity::_src of _d   v3:42  */
    IFL_BUF_D_DST,     /* FluidQuantity::_dst of _d   v3:43  */
    IFL_BUF_U_SRC,
    IFL_BUF_U_DST,
    IFL_BUF_V_SRC,
    IFL_BUF_V_DST,
    IFL_BUF_T_SRC, /* v6+ temperature v6:589 */
    IFL_BUF_T_DST,
    IFL_BUF_R,      /* FluidSolver::_r      v3:198 */
    IFL_BUF_P,      /* FluidSolver::_p      v3:199 */
    IFL_BUF_Z,      /* FluidSolver::_z      v3:200 */
    IFL_BUF_S,      /* FluidSolver::_s      v3:201 */
    IFL_BUF_PRECON, /* FluidSolver::_precon v3:202 */
    IFL_BUF_ADIAG,  /* FluidSolver::_aDiag  v3:204 */
    IFL_BUF_APLUSX, /* FluidSolver::_aPlusX v3:205 */
    IFL_BUF_APLUSY, /* FluidSolver::_aPlusY v3:206 */
    IFL_BUF_UDENSITY, /* FluidSolver::_uDensity v7:598, (w+1)*h */
    IFL_BUF_VDENSITY, /* FluidSolver::_vDensity v7:599, w*(h+1) */
    IFL_BUF_COUNT_
} ifl_buf;

/* Field ids for per-quantity ops (FluidQuantity instances of FluidSolver, v3:188-190). */
typedef enum { IFL_FIELD_D = 0, IFL_FIELD_U = 1, IFL_FIELD_V = 2, IFL_FIELD_T = 3 } ifl_field;

/* project() exit status; the host shim prints the reference's exact stdout line from it. */
typedef enum {
    IFL_SOLVE_CONVERGED = 0,     /* "Exiting solver after %d iterations, maximum error is %f"  v3:368 */
    IFL_SOLVE_EXCEEDED = 1,      /* "Exceeded budget of %d iterations, maximum error was %f"   v3:379 */
    IFL_SOLVE_INITIAL_SMALL = 2  /* early return, initial |r|inf < 1e-5                         v3:355 */
} ifl_solve_status;

typedef struct {
    int status;       /* ifl_solve_status */
    int iterations;   /* the zero-based `iter` the reference prints (v3:368), or `limit` when exceeded */
    double max_error; /* |r|inf (PCG, v3:366) or max |delta p| (Gauss-Seidel, v2:262) at exit */
} ifl_solve_info;

/* ---- lifetime -------------------------------------------------------------- */
/* FluidSolver ctor (v3:401-416): allocates fields + solver scratch on `device`,
 * all zero, hx = 1/min(w,h). */
int ifl_create(ifl_ctx **out, int w, int h, int version, int device);
/* FluidSolver dtor (v3:418-431). */
int ifl_destroy(ifl_ctx *ctx);
const char *ifl_last_error(void);

/* ---- row-slab multi-GPU (one process per GPU) --------------------------
 * Rank `rank` of `world` (<= 8) owns a slab of whole 32-row strips of every array and
 * runs on `device`.  All ranks map the slabs of all ranks into one address range (CUDA
 * virtual memory management, handles passed over the UNIX socket `rendezvous`, which
 * rank 0 binds), so kernels address neighbours' rows over NVLink; MIC(0) stays exact: the
 * wavefront crosses slab boundaries through the same hand-off messages it uses between
 * strips.  Results are bit-identical to the one-GPU context on every rank.
 * Collective calls (every rank, same order): create/destroy, upload/download/fill, every
 * compute entry point.  ifl_download returns the WHOLE array on every rank; ifl_upload and
 * ifl_update_host move only the caller's slab rows of the (whole-array) host buffers.
 * Chapters 1-7 (the chapter-8 particle set is not sharded yet). */
int ifl_create_dist(ifl_ctx **out, int w, int h, int version, int device, int rank, int world, const char *rendezvous);
/* Slab of `rank`: cell rows [row0, row1) (pure host arithmetic). */
int ifl_dist_plan(int h, int world, int rank, int *row0, int *row1);
int ifl_dist_info(const ifl_ctx *ctx, int *rank, int *world, int *row0, int *row1);
/* Device barrier across the ranks on the context's stream, then host synchronisation. */
int ifl_dist_barrier(ifl_ctx *ctx);
/* Host-only self-test of the rendezvous / descriptor passing (no CUDA involved). */
int ifl_dist_selftest(int rank, int world, const char *rendezvous);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long ifl_launch_count(const ifl_ctx *ctx);
/* The CUDA stream (cudaStream_t as void*) every kernel of this context is launched on. */
void *ifl_stream(const ifl_ctx *ctx);
int ifl_sync(ifl_ctx *ctx);

/* ---- measurement (bench.py's roofline leg) ------------------------------------ */
/* Kernel classes of the hot path; per-class device time is measured with CUDA events
 * recorded on the context's stream around every launch of that class. */
typedef enum {
    IFL_K_MATVEC = 0,   /* z = A s (+ z.s)                    v3:361-362 */
    IFL_K_AXPY2_NORM,   /* p += a s; r -= a z; |r|inf         v3:363-366 */
    IFL_K_PRECON_FWD,   /* forward substitution               v3:276-287 */
    IFL_K_PRECON_BWD,   /* backward substitution (+ z.r)      v3:289-303, 374 */
    IFL_K_XPAY,         /* s = z + b s                        v3:375 */
    IFL_K_SCALAR,       /* alpha / convergence test / beta    */
    IFL_K_FACTOR,       /* buildPreconditioner                v3:247-272 */
    IFL_K_ASSEMBLY,     /* buildRhs, buildPressureMatrix, applyPressure, addInflow */
    IFL_K_ADVECT,       /* FluidQuantity::advect              v2:170-183 */
    IFL_K_GS_SWEEP,     /* one Gauss-Seidel sweep             v2:239-268 */
    IFL_K_P2G,          /* FluidQuantity::fromParticles       v8:663-688 */
    IFL_K_G2P,          /* gridToParticles, diff, undiff      v8:904-911, 379-388 */
    IFL_K_COUNT_
} ifl_kernel_class;
/* Start (on != 0) / stop per-class event timing; starting resets the accumulators. */
int ifl_profile(ifl_ctx *ctx, int on);
/* Synchronises the stream and returns accumulated device milliseconds and launch counts
 * per class (arrays of IFL_K_COUNT_ entries). */
int ifl_profile_read(ifl_ctx *ctx, double *ms, long long *launches);

/* Diagnostics for the wavefront kernels: arm != 0 makes every following sweep record the
 * %globaltimer (ns) at which each 32-row strip started and finished; arm == 0 disarms,
 * copies the last sweep's [strips][16] table (slot 0 start, 1 end, 15 SM cycles of the compute
 * warp; slots 2..10 hold a timestamp every 1/8 of the strip only in builds with
 * -DIFL_SWEEP_DIAG=1, the probe costs the hot loop 3 %) into out_ns and returns the strip count. */
int ifl_debug_sweep_times(ifl_ctx *ctx, int arm, unsigned long long *out_ns, int capacity);

/* ---- data movement (backs FluidQuantity::at()/src(), toImage, test harness) ---- */
size_t ifl_buf_elems(const ifl_ctx *ctx, int buf);
int ifl_upload(ifl_ctx *ctx, int buf, const double *host);   /* dense host -> device   */
int ifl_download(ifl_ctx *ctx, int buf, double *host);       /* device -> dense host   */
int ifl_fill(ifl_ctx *ctx, int buf, double value);

/* ---- FluidQuantity ops ------------------------------------------------------ */
/* FluidQuantity::addInflow(x0,y0,x1,y1,v)  v2:188-205 (v1:141-151 when version==1),
 * including the `_h`-for-`_w` clamp of the x loop. */
int ifl_quantity_add_inflow(ifl_ctx *ctx, int field, double x0, double y0, double x1, double y1, double v);
/* FluidQuantity::advect(timestep,u,v)  v2:170-183 (v1:125-138): src -> dst, reads u._src, v._src. */
int ifl_advect(ifl_ctx *ctx, int field, double timestep);
/* FluidSolver::maxTimestep()  1-matrixless/Fluid.cpp:310-328: 2*hx / max |(u, v)| at the cell centres, at most 1. */
int ifl_max_timestep(ifl_ctx *ctx, double *result);
/* FluidQuantity::flip()  v3:105-107. */
int ifl_flip(ifl_ctx *ctx, int field);

/* ---- chapters 4+: solid bodies ------------------------------------------------- */
/* One SolidBody (v4:79-149) as plain data: kind 0 = SolidBox(x,y,sx,sy,theta,vx,vy,vtheta)
 * (v4:155), 1 = SolidSphere(x,y,s,theta,vx,vy,vtheta) (v4:207; scale_x == scale_y == s). */
typedef struct {
    int kind;
    double pos_x, pos_y, scale_x, scale_y, theta, vel_x, vel_y, vel_theta;
} ifl_body;
/* The solver holds its body list by reference and the caller moves the bodies between
 * updates (v4:960-961): call this after every SolidBody::update(). */
int ifl_set_bodies(ifl_ctx *ctx, const ifl_body *bodies, int n);
int ifl_fill_solid_fields(ifl_ctx *ctx, int field);   /* FluidQuantity::fillSolidFields v5:506-559 */
int ifl_set_boundary_condition(ifl_ctx *ctx);          /* FluidSolver::setBoundaryCondition v4:812-833 */
int ifl_extrapolate(ifl_ctx *ctx, int field);          /* FluidQuantity::extrapolate v4:551-587 */
/* Per-quantity solid arrays (dense host layout like ifl_upload): doubles for volume,
 * normals and phi ((w+1)*(h+1)); bytes for cell and body. */
typedef enum { IFL_AUX_VOLUME = 0, IFL_AUX_NORMAL_X, IFL_AUX_NORMAL_Y, IFL_AUX_PHI, IFL_AUX_CELL, IFL_AUX_BODY } ifl_aux;
size_t ifl_aux_elems(const ifl_ctx *ctx, int field, int which);
int ifl_aux_download(ifl_ctx *ctx, int field, int which, void *host);
int ifl_aux_upload(ifl_ctx *ctx, int field, int which, const void *host);

/* ---- chapters 6+: heat diffusion, buoyancy, variable density ------------------- */
/* FluidSolver(w,h,rhoAir,rhoSoot,diffusion,bodies) ctor arguments (v6:921); ambient
 * temperature 294 K and g = 9.81 are the reference's constants (v6:932-933). */
int ifl_set_fluid_params(ifl_ctx *ctx, double rho_air, double rho_soot, double diffusion);
double ifl_ambient_t(const ifl_ctx *ctx);                 /* FluidSolver::ambientT v6:1017 */
int ifl_build_heat_matrix(ifl_ctx *ctx, double timestep); /* buildHeatDiffusionMatrix v6:683-712 */
int ifl_add_buoyancy(ifl_ctx *ctx, double timestep);      /* addBuoyancy v6:881-895 */
int ifl_compute_densities(ifl_ctx *ctx);                  /* computeDensities v7:658-675 */
/* FluidSolver::addInflow(x,y,w,h,d,t,u,v) v6:1010-1015 */
int ifl_add_inflow_t(ifl_ctx *ctx, double x, double y, double w, double h, double d, double t, double u, double v);
/* For chapters 6+ ifl_update ignores `density` (uses ifl_set_fluid_params) and fills
 * infos[0] (heat solve) and infos[1] (pressure solve). */

/* ---- chapter 8: FLIP particle <-> grid transfers -------------------------------- */
/* The particle set is ParticleQuantities' SoA (v8:717-723): posX, posY and one property
 * array per registered quantity in the order d, t, u, v (v8:1309-1312); capacity w*h*12
 * (_MaxPerCell, v8:694, 869).  Counts are 64-bit: the reference's `int` overflows at 13378^2 cells. */
long long ifl_particles_capacity(const ifl_ctx *ctx);
long long ifl_particles_count(const ifl_ctx *ctx);
/* What the two constructors do (v8:877, 1306-1314): initParticles (v8:735-751) on the jittered grid with
 * `avg_per_cell` attempts per cell (_AvgPerCell, 4 as shipped), consuming the reference's frand stream
 * (v8:38-46) from its seed, then gridToParticles(1.0).  Call once, after the bodies have been set. */
int ifl_particles_init(ifl_ctx *ctx, int avg_per_cell);
/* particlesToGrid (v8:916-927): per quantity fromParticles + extrapolate, then countParticles,
 * pruneParticles, seedParticles (v8:754-813) with the same random stream, slot assignment and
 * swap-with-last order as the sequential loops.  *count (may be NULL) = what "Particle count: %d" prints. */
int ifl_particles_to_grid(ifl_ctx *ctx, long long *count);
int ifl_count_particles(ifl_ctx *ctx);  /* v8:754-763 */
int ifl_prune_particles(ifl_ctx *ctx);  /* v8:766-786 (needs ifl_count_particles) */
int ifl_seed_particles(ifl_ctx *ctx);   /* v8:789-813 (needs the counts after pruning) */
int ifl_particles_upload(ifl_ctx *ctx, long long count, const double *pos_x, const double *pos_y, const double *prop_d,
                         const double *prop_t, const double *prop_u, const double *prop_v);
int ifl_particles_download(ifl_ctx *ctx, long long *count, double *pos_x, double *pos_y, double *prop_d, double *prop_t,
                           double *prop_u, double *prop_v);
/* Test hook: raw slots [first, first + n) of posX (what = 0), posY (1), the four properties (2..5) -- including
 * the stale slots past the count, which the reference's seedParticles may read (v8:802) -- or the per-cell
 * counts as ints (6). */
int ifl_particles_peek(ifl_ctx *ctx, int what, long long first, long long n, void *host);
/* FluidQuantity::fromParticles (v8:663-688): P2G of one quantity, contributions summed in
 * ascending particle index like the reference; marks particle-free fluid cells CELL_EMPTY (2). */
int ifl_from_particles(ifl_ctx *ctx, int field);
/* ParticleQuantities::gridToParticles(alpha) v8:904-911 */
int ifl_grid_to_particles(ifl_ctx *ctx, double alpha);
int ifl_quantity_copy(ifl_ctx *ctx, int field);                   /* FluidQuantity::copy   v8:374 (_old == the *_DST buffer) */
int ifl_quantity_diff(ifl_ctx *ctx, int field, double alpha);     /* FluidQuantity::diff   v8:379 */
int ifl_quantity_undiff(ifl_ctx *ctx, int field, double alpha);   /* FluidQuantity::undiff v8:385 */
/* ParticleQuantities::advect(timestep, u, v) v8:931-939 */
int ifl_particles_advect(ifl_ctx *ctx, double timestep);

/* ---- FluidSolver private hot-path methods ----------------------------------- */
int ifl_build_rhs(ifl_ctx *ctx);                                         /* v3:208-217 */
int ifl_build_pressure_matrix(ifl_ctx *ctx, double timestep, double density); /* v3:222-244 */
int ifl_build_preconditioner(ifl_ctx *ctx);                              /* v3:247-272 */
int ifl_apply_preconditioner(ifl_ctx *ctx, int dst, int a);              /* v3:275-304 */
int ifl_matrix_vector_product(ifl_ctx *ctx, int dst, int b);             /* v3:315-332 */
int ifl_dot_product(ifl_ctx *ctx, int a, int b, double *result);         /* v3:307-312 */
int ifl_scaled_add(ifl_ctx *ctx, int dst, int a, int b, double s);       /* v3:335-338 */
int ifl_infinity_norm(ifl_ctx *ctx, int a, double *result);              /* v3:341-346 */
/* project(limit)  v3:349-380: MIC(0)-PCG, whole loop on the device. */
int ifl_project(ifl_ctx *ctx, int limit, ifl_solve_info *info);
/* project(limit, timestep)  v2:233-277 / v1:192-240: lexicographic Gauss-Seidel, warm-started _p. */
int ifl_project_gs(ifl_ctx *ctx, int limit, double timestep, double density, ifl_solve_info *info);
int ifl_apply_pressure(ifl_ctx *ctx, double timestep, double density);   /* v3:382-398 */

/* ---- FluidSolver public surface --------------------------------------------- */
/* FluidSolver::addInflow(x,y,w,h,d,u,v)  v3:449-453. */
int ifl_add_inflow(ifl_ctx *ctx, double x, double y, double w, double h, double d, double u, double v);
/* FluidSolver::update(timestep)  v3:433-447 (v2:320-332, v1:284-297): the whole
 * step stays on the device.  `density` is the ctor's density; infos[0] receives
 * the pressure solve's status (may be NULL). */
int ifl_update(ifl_ctx *ctx, double timestep, double density, ifl_solve_info *infos);
/* Same step with HOST buffers: uploads d,u,v (reference layout), runs update(),
 * downloads d,u,v.  This is the end-to-end path bench.py times as `e2e`. */
int ifl_update_host(ifl_ctx *ctx, double timestep, double density, double *d, double *u, double *v,
                    ifl_solve_info *infos);

/* The same with host buffers that hold only the caller's slab (ifl_slab_elems(buf) doubles,
 * the slab's first row first): what a multi-GPU host program keeps per rank.  On one GPU
 * the slab is the whole array and this equals ifl_update_host. */
int ifl_update_host_slab(ifl_ctx *ctx, double timestep, double density, double *d, double *u, double *v,
                         ifl_solve_info *infos);
size_t ifl_slab_elems(const ifl_ctx *ctx, int buf);
/* Slab-local ifl_upload / ifl_download (not collective; `host` = the caller's rows only). */
int ifl_upload_slab(ifl_ctx *ctx, int buf, const double *host);
int ifl_download_slab(ifl_ctx *ctx, int buf, double *host);

#ifdef __cplusplus
}
#endif
#endif /* IFL_B200_H */
