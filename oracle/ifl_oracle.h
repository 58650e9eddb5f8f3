/*
 * ifl_oracle.h -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * unmodified reference (oracle/_ref/libref_v*.so, built from /root/reference) by
 * tests/test_oracle_vs_reference.py, and against golden vectors generated from the
 * reference and committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
 * may load this library.  The product (libifl_b200.so) never links or calls it.
 *
 * Plain C, scalar, double precision, compiled with -O2 -ffp-contract=off (the
 * reference Makefile's -O2 emits no FMA; SURVEY "Facts established by probing").
 * All arrays are dense row-major, index x + y*w, exactly like the reference
 * (v3:113-119).  Citations: vN:L == /root/reference/N-<chapter>/Fluid.cpp:L.
 */
#ifndef IFL_ORACLE_H
#define IFL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* One FluidQuantity's geometry (v3:41-49): w x h samples, offset (ox, oy) in cell units. */
typedef struct {
    int w, h;
    double ox, oy;
} ofl_grid;

/* FluidQuantity::lerp(x,y) v2:133-145 and cerp(x,y) v2:150-167 */
double ofl_lerp(const double *f, ofl_grid g, double x, double y);
double ofl_cerp(const double *f, ofl_grid g, double x, double y);

/* FluidQuantity::advect v2:170-183 (mode 1: RK3 + Catmull-Rom) / v1:125-138 (mode 0: Euler + bilinear) */
void ofl_advect(int mode, double *dst, const double *src, ofl_grid g, const double *u, ofl_grid gu,
                const double *v, ofl_grid gv, double timestep, double hx);

/* FluidQuantity::addInflow v2:188-205 (smooth 1) / v1:141-151 (smooth 0) */
void ofl_add_inflow(double *src, ofl_grid g, double hx, double x0, double y0, double x1, double y1, double v,
                    int smooth);

/* FluidSolver::buildRhs v3:208-217 */
void ofl_build_rhs(double *r, const double *u, const double *v, int w, int h, double hx);
/* FluidSolver::buildPressureMatrix v3:222-244 */
void ofl_build_pressure_matrix(double *aDiag, double *aPlusX, double *aPlusY, int w, int h, double timestep,
                               double density, double hx);
/* FluidSolver::buildPreconditioner v3:247-272 */
void ofl_build_preconditioner(double *precon, const double *aDiag, const double *aPlusX, const double *aPlusY,
                              int w, int h);
/* FluidSolver::applyPreconditioner v3:275-304 */
void ofl_apply_preconditioner(double *dst, const double *a, const double *precon, const double *aPlusX,
                              const double *aPlusY, int w, int h);
/* v3:307-312, 315-332, 335-338, 341-346 */
double ofl_dot_product(const double *a, const double *b, int n);
void ofl_matrix_vector_product(double *dst, const double *b, const double *aDiag, const double *aPlusX,
                               const double *aPlusY, int w, int h);
void ofl_scaled_add(double *dst, const double *a, const double *b, double s, int n);
double ofl_infinity_norm(const double *a, int n);

/* FluidSolver::project(limit) v3:349-380.  status: 0 converged, 1 exceeded, 2 initial-small.
 * *iters is the zero-based `iter` the reference prints (or limit when exceeded). */
int ofl_project(int limit, double *p, double *r, double *z, double *s, const double *precon, const double *aDiag,
                const double *aPlusX, const double *aPlusY, int w, int h, int *iters, double *max_error);
/* FluidSolver::project(limit, timestep) v2:233-277 (Gauss-Seidel, warm-started p) */
int ofl_project_gs(int limit, double timestep, double density, double hx, double *p, const double *r, int w,
                   int h, int *iters, double *max_delta);
/* FluidSolver::applyPressure v3:382-398 */
void ofl_apply_pressure(double *u, double *v, const double *p, int w, int h, double timestep, double density,
                        double hx);

#ifdef __cplusplus
}
#endif
#endif
