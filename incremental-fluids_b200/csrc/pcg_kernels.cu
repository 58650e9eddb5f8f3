// pcg_kernels.cu -- the HBM-streaming half of the MIC(0)-PCG loop and its driver.
//
//   k_matvec       FluidSolver::matrixVectorProduct v3:315-332 (+ fused dotProduct(z,s) v3:362)
//   k_axpy2_norm   scaledAdd x2 + infinityNorm       v3:363-366
//   k_xpay         scaledAdd(s, z, s, beta)           v3:375
//   k_dot / k_inf_norm / k_scaled_add                 v3:307-312, 341-346, 335-338 (granular ABI)
//   k_scalar<...>  the three global scalars of an iteration (alpha, |r|inf test, beta)
//   pcg_project    FluidSolver::project(limit)        v3:349-380
//
// Layout: pitched arrays (ifl_internal.cuh).  Every thread owns an x-pair (one
// 16-byte double2 per array per row) and walks ROWS_PER_THREAD consecutive rows, so
// the stencil keeps the rows above/below in registers (each array is read once from
// HBM; x-neighbours come from warp shuffles).  Reductions are two-level and
// deterministic: fixed tree inside the block -> partials[block] -> one finishing
// block that sums the partials in a fixed order.
//
// The vector that receives A*s is a separate buffer `q` (the reference reuses _z,
// v3:361); this only renames storage and lets s = z + beta*s be fused into later
// kernels without a read-after-write hazard.  ifl_download(IFL_BUF_Z) is unaffected.
#include "ifl_internal.cuh"

#include <stddef.h>
#include <string.h>

namespace ifl {

constexpr int VEC_THREADS = 128;                // x-pairs per block -> 256 columns
constexpr int VEC_COLS = VEC_THREADS * 2;
constexpr int VEC_ROWS = 16;                    // rows walked by each thread

// Grid over the rows [ry0, ry1) this rank owns (one GPU: the whole array).  Blocks are
// numbered as in the whole-array grid (block row = y / VEC_ROWS; ry0 is a multiple of 32),
// so partials land in the same slots whatever the number of ranks and vec_blocks() -- the
// number of partials the finishing block folds -- always counts the whole array.
static dim3 vec_grid(const Arr &a) {
    return dim3((a.w + VEC_COLS - 1) / VEC_COLS, (a.ry1 - a.ry0 + VEC_ROWS - 1) / VEC_ROWS);
}
static int vec_blocks(const Arr &a) { return ((a.w + VEC_COLS - 1) / VEC_COLS) * ((a.h + VEC_ROWS - 1) / VEC_ROWS); }

// Chapters 6+ mask dotProduct / scaledAdd / infinityNorm with `cell == CELL_FLUID`
// (v6:781-826); matrixVectorProduct stays unmasked (v6:791).  `cell` is _d's byte array
// (row pitch == the vectors' pitch) or null for the unmasked chapters.
struct CellMask {
    const uint8_t *cell;
    __device__ __forceinline__ bool fluid(size_t i) const { return !cell || cell[i] == CELL_FLUID; }
};

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

// Chapter 3's matrix (v3:222-244) is a pure function of the cell's position: every existing neighbour contributes
// `scale` to the diagonal -- equal addends, so the sum depends on their NUMBER only: d[n] = ((0 + s) + s) ... -- and
// -scale to aPlusX / aPlusY.  UNIFORM evaluates those values in registers instead of reading 24 bytes per cell; the
// products and their order are the ones of the stored matrix (which buildPressureMatrix still writes: the factorisation
// and every accessor read it).  Off as soon as the caller uploads a matrix of their own.
struct UniformMatrix {
    double d[5];   // diagonal for 0..4 neighbours
    double off;    // -scale
};

// dst = A*b (5-point stencil in the reference's summation order: diag, left, up,
// right, down); optionally partial[block] = sum(dst*b) over the block's cells.
template <bool WITH_DOT, bool UNIFORM>
__global__ void __launch_bounds__(VEC_THREADS) k_matvec(Arr dst, Arr b, Arr aDiag, Arr aPlusX, Arr aPlusY,
                                                         double *__restrict__ partials,
                                                         const SolveScalars *__restrict__ gate, CellMask mk, UniformMatrix um) {
    if (gate && gate->done) return;
    __shared__ double red[32];
    const int W = dst.w, H = dst.h, pitch = dst.pitch;
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * VEC_COLS + threadIdx.x * 2;
    const int y0 = dst.ry0 + blockIdx.y * VEC_ROWS;
    const int y1 = imin(y0 + VEC_ROWS, dst.ry1);
    const bool in = x < pitch; // may load (pad columns are zero)
    const double2 zero2 = make_double2(0.0, 0.0);

    // rolling registers: b and aPlusY of the row above, b of this row and the row below
    double2 b_up = zero2, ay_up = zero2, b_c = zero2, b_dn;
    if (in) {
        if (y0 > 0) {
            b_up = ld2(b.p + x + (size_t)(y0 - 1) * pitch);
            if (!UNIFORM) ay_up = ld2(aPlusY.p + x + (size_t)(y0 - 1) * pitch);
        }
        b_c = ld2(b.p + x + (size_t)y0 * pitch);
    }
    if (UNIFORM) ay_up = make_double2(um.off, um.off); // only used for y > 0, where the cell above has a lower neighbour
    // diagonals of cells x, x+1 for 0, 1, 2 vertical neighbours (selected once per thread; the loop only picks by row)
    const int nx0 = (x > 0 ? 1 : 0) + (x < W - 1 ? 1 : 0), nx1 = 1 + (x + 1 < W - 1 ? 1 : 0); // x-neighbours of cells x, x+1
    double dx0[3], dx1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dx0[k] = nx0 == 0 ? um.d[k] : (nx0 == 1 ? um.d[k + 1] : um.d[k + 2]);
        dx1[k] = nx1 == 1 ? um.d[k + 1] : um.d[k + 2];
    }
    double acc = 0.0;
    // UNIFORM: one 16-byte load per row and thread is too little in flight to reach the HBM rate (measured 4.1 TB/s):
    // the row after next is requested one iteration early
    double2 b_dn2 = zero2;
    if (UNIFORM && in) b_dn2 = ld2(b.p + x + (size_t)(y0 + 1) * pitch);
    for (int y = y0; y < y1; y++) {
        const size_t row = (size_t)y * pitch;
        double2 ad = zero2, ax = zero2, ay = zero2;
        b_dn = zero2;
        if (in) {
            if (!UNIFORM) {
                ad = ld2(aDiag.p + x + row);
                ax = ld2(aPlusX.p + x + row);
                ay = ld2(aPlusY.p + x + row);
                b_dn = ld2(b.p + x + row + pitch); // rows >= H are zero pad (allocation has 32 spare rows)
            } else {
                b_dn = b_dn2;
                if (y + 1 < y1) b_dn2 = ld2(b.p + x + row + 2 * (size_t)pitch);
            }
        }
        if (UNIFORM) {
            const int ny = (y > 0 ? 1 : 0) + (y < H - 1 ? 1 : 0);
            ad = make_double2(ny == 2 ? dx0[2] : (ny == 1 ? dx0[1] : dx0[0]), ny == 2 ? dx1[2] : (ny == 1 ? dx1[1] : dx1[0]));
            ax = make_double2(um.off, um.off); // every use below is guarded by the neighbour's existence
            ay = ax;
        }
        // x-neighbours: b[x-1], aPlusX[x-1] from the lane to the left, b[x+2] from the right
        double b_l = __shfl_up_sync(0xffffffffu, b_c.y, 1);
        double ax_l = UNIFORM ? um.off : __shfl_up_sync(0xffffffffu, ax.y, 1);
        double b_r = __shfl_down_sync(0xffffffffu, b_c.x, 1);
        if (lane == 0 && x > 0 && in) {
            b_l = b.p[x - 1 + row];
            if (!UNIFORM) ax_l = aPlusX.p[x - 1 + row];
        }
        if (lane == 31 && x + 2 < pitch) b_r = b.p[x + 2 + row];

        // cell x
        double t0 = ad.x * b_c.x;
        if (x > 0) t0 += ax_l * b_l;
        if (y > 0) t0 += ay_up.x * b_up.x;
        if (x < W - 1) t0 += ax.x * b_c.y;
        if (y < H - 1) t0 += ay.x * b_dn.x;
        // cell x+1
        double t1 = ad.y * b_c.y;
        t1 += ax.x * b_c.x; // x+1 > 0 always
        if (y > 0) t1 += ay_up.y * b_up.y;
        if (x + 1 < W - 1) t1 += ax.y * b_r;
        if (y < H - 1) t1 += ay.y * b_dn.y;

        if (x + 1 < W) {
            st2(dst.p + x + row, make_double2(t0, t1));
            if (WITH_DOT) {
                if (mk.fluid(x + row)) acc += t0 * b_c.x;
                if (mk.fluid(x + 1 + row)) acc += t1 * b_c.y;
            }
        } else if (x < W) {
            dst.p[x + row] = t0;
            if (WITH_DOT && mk.fluid(x + row)) acc += t0 * b_c.x;
        }
        b_up = b_c;
        b_c = b_dn;
        ay_up = ay;
    }
    if (WITH_DOT) {
        const double s = block_reduce<false>(acc, red);
        if (threadIdx.x == 0) partials[(y0 / VEC_ROWS) * gridDim.x + blockIdx.x] = s;
    }
}

// Chapter 3, fused: s' = z + beta*s (v3:375 of the PREVIOUS iteration) ; q = A*s' ; partial[block] = sum(q*s')  (v3:361-362).
// The block forms s' for its 16 rows and for the row above and below them in registers (x-neighbours by shuffle), so the
// search direction is read once instead of written, read and re-read: 34 instead of 24 + 16 bytes per cell.  s' goes to the
// OTHER buffer of a ping-pong pair -- in place, a block could find its halo rows already overwritten by its neighbours --
// and the host swaps the two Arr records after the launch.  Same products, same order as k_scaled_add + k_matvec<UNIFORM>.
__global__ void __launch_bounds__(VEC_THREADS) k_xpay_matvec(Arr dst, Arr z, Arr s_old, Arr s_new, double *__restrict__ partials,
                                                              const SolveScalars *__restrict__ sc, UniformMatrix um) {
    if (sc->done) return;
    __shared__ double red[32];
    const double beta = sc->beta;
    const int W = dst.w, H = dst.h, pitch = dst.pitch;
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * VEC_COLS + threadIdx.x * 2;
    const int y0 = dst.ry0 + blockIdx.y * VEC_ROWS;
    const int y1 = imin(y0 + VEC_ROWS, dst.ry1);
    const bool in = x < pitch; // may load (pad columns are zero)
    const double2 zero2 = make_double2(0.0, 0.0);
    auto xpay2 = [&](size_t i) { // s' of cells i, i+1 (pad cells: 0 + 0*beta)
        const double2 zv = ld2(z.p + i), sv = ld2(s_old.p + i);
        return make_double2(zv.x + sv.x * beta, zv.y + sv.y * beta);
    };
    auto xpay1 = [&](size_t i) { return z.p[i] + s_old.p[i] * beta; };

    double2 b_up = zero2, b_c = zero2, b_dn = zero2;
    if (in) {
        if (y0 > 0) b_up = xpay2(x + (size_t)(y0 - 1) * pitch);
        b_c = xpay2(x + (size_t)y0 * pitch);
    }
    const int nx0 = (x > 0 ? 1 : 0) + (x < W - 1 ? 1 : 0), nx1 = 1 + (x + 1 < W - 1 ? 1 : 0); // x-neighbours of cells x, x+1
    double dx0[3], dx1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        dx0[k] = nx0 == 0 ? um.d[k] : (nx0 == 1 ? um.d[k + 1] : um.d[k + 2]);
        dx1[k] = nx1 == 1 ? um.d[k + 1] : um.d[k + 2];
    }
    const double off = um.off;
    double acc = 0.0;
    for (int y = y0; y < y1; y++) {
        const size_t row = (size_t)y * pitch;
        b_dn = zero2;
        if (in && y + 1 < H) b_dn = xpay2(x + row + pitch); // (row H is a zero pad row and never multiplied)
        const int ny = (y > 0 ? 1 : 0) + (y < H - 1 ? 1 : 0);
        const double2 ad = make_double2(ny == 2 ? dx0[2] : (ny == 1 ? dx0[1] : dx0[0]), ny == 2 ? dx1[2] : (ny == 1 ? dx1[1] : dx1[0]));
        double b_l = __shfl_up_sync(0xffffffffu, b_c.y, 1);
        double b_r = __shfl_down_sync(0xffffffffu, b_c.x, 1);
        if (lane == 0 && x > 0 && in) b_l = xpay1(x - 1 + row);
        if (lane == 31 && x + 2 < pitch) b_r = xpay1(x + 2 + row);

        double t0 = ad.x * b_c.x;
        if (x > 0) t0 += off * b_l;
        if (y > 0) t0 += off * b_up.x;
        if (x < W - 1) t0 += off * b_c.y;
        if (y < H - 1) t0 += off * b_dn.x;
        double t1 = ad.y * b_c.y;
        t1 += off * b_c.x;
        if (y > 0) t1 += off * b_up.y;
        if (x + 1 < W - 1) t1 += off * b_r;
        if (y < H - 1) t1 += off * b_dn.y;

        if (x + 1 < W) {
            st2(s_new.p + x + row, b_c);
            st2(dst.p + x + row, make_double2(t0, t1));
            acc += t0 * b_c.x;
            acc += t1 * b_c.y;
        } else if (x < W) {
            st2(s_new.p + x + row, b_c); // (the pad cell next to the last column receives 0 + 0*beta, as k_scaled_add writes it)
            dst.p[x + row] = t0;
            acc += t0 * b_c.x;
        }
        b_up = b_c;
        b_c = b_dn;
    }
    const double sum = block_reduce<false>(acc, red);
    if (threadIdx.x == 0) partials[(y0 / VEC_ROWS) * gridDim.x + blockIdx.x] = sum;
}

// p += alpha*s ; r += q*(-alpha) ; partial[block] = max|r|     v3:363-366
// band_count (may be null): one counter per band of 64 rows; every block adds 1 to its band's counter after its
// stores, so that the forward sweep -- launched concurrently on the main stream -- can start a strip as soon as
// the rows it reads are final (blocks are 16 rows tall and never straddle a band).
__global__ void __launch_bounds__(VEC_THREADS) k_axpy2_norm(Arr p, Arr s, Arr r, Arr q, const SolveScalars *sc,
                                                             double *__restrict__ partials, CellMask mk,
                                                             unsigned *__restrict__ band_count) {
    if (sc->done) { // the sweep does not wait either once the solve is over (it returns on the same flag)
        return;
    }
    __shared__ double red[32];
    const double alpha = sc->alpha;
    const double nalpha = -alpha;
    const int W = p.w, pitch = p.pitch;
    const int x = blockIdx.x * VEC_COLS + threadIdx.x * 2;
    const int y0 = p.ry0 + blockIdx.y * VEC_ROWS;
    const int y1 = imin(y0 + VEC_ROWS, p.ry1);
    double m = 0.0;
    if (x < W) {
#pragma unroll 4
        for (int y = y0; y < y1; y++) {
            const size_t i = x + (size_t)y * pitch;
            double2 pv = ld2(p.p + i), sv = ld2(s.p + i), rv = ld2(r.p + i), qv = ld2(q.p + i);
            const bool f0 = mk.fluid(i), f1 = mk.fluid(i + 1);
            if (f0) {
                pv.x = pv.x + sv.x * alpha;
                rv.x = rv.x + qv.x * nalpha;
                m = std_max(m, fabs(rv.x));
            }
            if (f1) {
                pv.y = pv.y + sv.y * alpha;
                rv.y = rv.y + qv.y * nalpha;
                m = std_max(m, fabs(rv.y)); // pad column (x+1 == W) holds 0 -> no effect
            }
            st2(p.p + i, pv);
            st2(r.p + i, rv);
        }
    }
    m = block_reduce<true>(m, red); // (ends with a __syncthreads: every thread's stores precede the signal below)
    if (threadIdx.x == 0) {
        partials[(y0 / VEC_ROWS) * gridDim.x + blockIdx.x] = m;
        if (band_count) {
            // every band counts 64 / VEC_ROWS slots per block column and launch; the last block row of a ragged grid
            // (H not a multiple of 64) also signs for the slots that do not exist
            const unsigned slots = (y0 + VEC_ROWS >= p.ry1) ? (unsigned)((64 - y0 % 64) / VEC_ROWS) : 1u;
            __threadfence();
            atomicAdd(&band_count[y0 / 64], slots);
        }
    }
}

// The same kernel as a PERSISTENT grid (a few blocks per SM, no shared memory to speak of) that walks the tiles of
// k_axpy2_norm's grid band by band, top to bottom: it leaves room on every SM for the forward sweep's CTAs (which
// need the SM's whole shared memory but only 224 threads), and finishes the bands in the order the sweep needs them.
// Partials land in the slots of the plain kernel, so the folded norm is the same.
__global__ void __launch_bounds__(VEC_THREADS) k_axpy2_norm_persistent(Arr p, Arr s, Arr r, Arr q, const SolveScalars *sc,
                                                                        double *__restrict__ partials, CellMask mk,
                                                                        unsigned *__restrict__ band_count, int gx, int gy,
                                                                        unsigned *__restrict__ ticket, unsigned ticket_base) {
    if (sc->done) return;
    __shared__ double red[32];
    __shared__ int s_t;
    const double alpha = sc->alpha;
    const double nalpha = -alpha;
    const int W = p.w, pitch = p.pitch;
    for (;;) {
        // blocks are handed out in order by a ticket, not striped over the CTAs: the forward sweep's CTAs already hold
        // many SMs when this kernel starts, and the CTAs of this grid that are not resident yet would otherwise keep a
        // share of EVERY band back until the resident ones had finished all of theirs
        if (threadIdx.x == 0) s_t = (int)(atomicAdd(ticket, 1u) - ticket_base);
        __syncthreads();
        const int t = s_t;
        if (t >= gx * gy) break; // (every CTA draws exactly one ticket too many: the host advances the base by total + grid)
        const int bx = t % gx, by = t / gx;
        const int x = bx * VEC_COLS + threadIdx.x * 2;
        const int y0 = p.ry0 + by * VEC_ROWS;
        const int y1 = imin(y0 + VEC_ROWS, p.ry1);
        double m = 0.0;
        if (x < W) {
#pragma unroll 4
            for (int y = y0; y < y1; y++) {
                const size_t i = x + (size_t)y * pitch;
                double2 pv = ld2(p.p + i), sv = ld2(s.p + i), rv = ld2(r.p + i), qv = ld2(q.p + i);
                const bool f0 = mk.fluid(i), f1 = mk.fluid(i + 1);
                if (f0) {
                    pv.x = pv.x + sv.x * alpha;
                    rv.x = rv.x + qv.x * nalpha;
                    m = std_max(m, fabs(rv.x));
                }
                if (f1) {
                    pv.y = pv.y + sv.y * alpha;
                    rv.y = rv.y + qv.y * nalpha;
                    m = std_max(m, fabs(rv.y));
                }
                st2(p.p + i, pv);
                st2(r.p + i, rv);
            }
        }
        m = block_reduce<true>(m, red); // (its barriers also keep s_t from being overwritten before every thread read it)
        if (threadIdx.x == 0) {
            partials[(y0 / VEC_ROWS) * gx + bx] = m;
            const unsigned slots = (y0 + VEC_ROWS >= p.ry1) ? (unsigned)((64 - y0 % 64) / VEC_ROWS) : 1u;
            __threadfence();
            atomicAdd(&band_count[y0 / 64], slots);
        }
    }
}

// dst = a + b*scale, scale either immediate or read from the device scalars (beta)
template <bool BETA_FROM_SCALARS>
__global__ void __launch_bounds__(VEC_THREADS) k_scaled_add(Arr dst, Arr a, Arr b, double scale,
                                                             const SolveScalars *sc, CellMask mk) {
    if (BETA_FROM_SCALARS) {
        if (sc->done) return;
        scale = sc->beta;
    }
    const int W = dst.w, pitch = dst.pitch;
    const int x = blockIdx.x * VEC_COLS + threadIdx.x * 2;
    const int y0 = dst.ry0 + blockIdx.y * VEC_ROWS;
    const int y1 = imin(y0 + VEC_ROWS, dst.ry1);
    if (x >= W) return;
#pragma unroll 4
    for (int y = y0; y < y1; y++) {
        const size_t i = x + (size_t)y * pitch;
        const double2 av = ld2(a.p + i), bv = ld2(b.p + i);
        if (!mk.cell) {
            st2(dst.p + i, make_double2(av.x + bv.x * scale, av.y + bv.y * scale));
        } else { // masked cells keep dst (which may alias a or b)
            if (mk.fluid(i)) dst.p[i] = av.x + bv.x * scale;
            if (mk.fluid(i + 1)) dst.p[i + 1] = av.y + bv.y * scale;
        }
    }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(VEC_THREADS) k_reduce2(Arr a, Arr b, double *__restrict__ partials, CellMask mk) {
    __shared__ double red[32];
    const int W = a.w, pitch = a.pitch;
    const int x = blockIdx.x * VEC_COLS + threadIdx.x * 2;
    const int y0 = a.ry0 + blockIdx.y * VEC_ROWS;
    const int y1 = imin(y0 + VEC_ROWS, a.ry1);
    double acc = 0.0;
    if (x < W) {
        for (int y = y0; y < y1; y++) {
            const size_t i = x + (size_t)y * pitch;
            const double2 av = ld2(a.p + i);
            const bool f0 = mk.fluid(i), f1 = mk.fluid(i + 1);
            if (IS_MAX) {
                if (f0) acc = std_max(acc, fabs(av.x));
                if (f1) acc = std_max(acc, fabs(av.y));
            } else {
                const double2 bv = ld2(b.p + i);
                if (f0) acc += av.x * bv.x;
                if (f1) acc += av.y * bv.y;
            }
        }
    }
    acc = block_reduce<IS_MAX>(acc, red);
    if (threadIdx.x == 0) partials[(y0 / VEC_ROWS) * gridDim.x + blockIdx.x] = acc;
}

// ---- scalar stage: one block folds the partials in a fixed order and updates the
// solve scalars.  MODE: 0 alpha = sigma/sum ; 1 convergence test on max ; 2 beta,
// sigma, iter++ ; 3 sigma = sum (prologue) ; 4 initial test (done=2 if small) ;
// 5 plain store of the reduced value to out[0].
enum { SC_ALPHA = 0, SC_CHECK = 1, SC_BETA = 2, SC_SIGMA0 = 3, SC_CHECK0 = 4, SC_STORE = 5 };

// With several ranks the partials of all slabs sit in one array (rank 0's HBM); the barrier
// makes every rank's blocks of the preceding reducing kernel visible, then every rank folds
// the SAME values in the SAME order: the scalars are bit-identical everywhere and equal to
// the one-GPU run's.  (`done` is such a scalar, so the early return is taken by all ranks.)
template <int MODE>
__global__ void __launch_bounds__(1024) k_scalar(const double *partials, int n, SolveScalars *sc, double *out,
                                                  DistDev dd) {
    constexpr bool IS_MAX = (MODE == SC_CHECK || MODE == SC_CHECK0);
    if (MODE != SC_STORE && MODE != SC_SIGMA0 && MODE != SC_CHECK0 && sc->done) return;
    dist_barrier_block(dd);
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double pv = __ldcg(partials + i); // L2 only: the peers' blocks were written over NVLink
        v = IS_MAX ? std_max(v, pv) : (v + pv);
    }
    v = block_reduce<IS_MAX>(v, red);
    if (threadIdx.x != 0) return;
    if (MODE == SC_ALPHA) {
        sc->alpha = sc->sigma / v; // v3:362
    } else if (MODE == SC_CHECK) {
        sc->max_error = v;
        if (v < 1e-5) sc->done = 1; // v3:367
    } else if (MODE == SC_BETA) {
        sc->beta = v / sc->sigma; // v3:375
        sc->sigma = v;            // v3:376
        sc->iter = sc->iter + 1;
    } else if (MODE == SC_SIGMA0) {
        sc->sigma = v; // v3:358
    } else if (MODE == SC_CHECK0) {
        sc->max_error = v;
        if (v < 1e-5) sc->done = 2; // v3:355
    } else {
        out[0] = v;
    }
}

// ------------------------------------------------------------------ launchers ----
static CellMask mask_of(ifl_ctx *c) {
    CellMask m;
    m.cell = c->version >= 6 ? c->fd[IFL_FIELD_D].cell : nullptr;
    return m;
}



int launch_matvec(ifl_ctx *c, const Arr &dst, const Arr &b, bool with_dot) {
    ProfScope ps_(c, IFL_K_MATVEC);
    dim3 g = vec_grid(dst);
    UniformMatrix um;
    memset(&um, 0, sizeof um);
    const bool uniform = c->matrix_uniform != 0;
    if (uniform) { // the sums buildPressureMatrix forms: equal addends, so only their number matters
        const double sc = c->matrix_scale;
        um.d[0] = 0.0;
        for (int n = 1; n <= 4; n++) um.d[n] = um.d[n - 1] + sc;
        um.off = -sc;
    }
    double *parts = with_dot ? partials_next(c) : nullptr;
    SolveScalars *gate = with_dot ? c->scal : nullptr;
    if (with_dot && uniform)
        k_matvec<true, true><<<g, VEC_THREADS, 0, c->stream>>>(dst, b, c->aDiag, c->aPlusX, c->aPlusY, parts, gate, mask_of(c), um);
    else if (with_dot)
        k_matvec<true, false><<<g, VEC_THREADS, 0, c->stream>>>(dst, b, c->aDiag, c->aPlusX, c->aPlusY, parts, gate, mask_of(c), um);
    else if (uniform)
        k_matvec<false, true><<<g, VEC_THREADS, 0, c->stream>>>(dst, b, c->aDiag, c->aPlusX, c->aPlusY, parts, gate, mask_of(c), um);
    else
        k_matvec<false, false><<<g, VEC_THREADS, 0, c->stream>>>(dst, b, c->aDiag, c->aPlusX, c->aPlusY, parts, gate, mask_of(c), um);
    if (with_dot) c->n_partials = vec_blocks(dst);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

// chapter 3: the pending s = z + beta*s of the previous iteration folded into q = A*s (ping-pong pair s / s2)
static int launch_xpay_matvec(ifl_ctx *c) {
    ProfScope ps_(c, IFL_K_MATVEC);
    UniformMatrix um;
    memset(&um, 0, sizeof um);
    const double sc = c->matrix_scale;
    for (int n = 1; n <= 4; n++) um.d[n] = um.d[n - 1] + sc;
    um.off = -sc;
    k_xpay_matvec<<<vec_grid(c->q), VEC_THREADS, 0, c->stream>>>(c->q, c->z, c->s, c->s2, partials_next(c), c->scal, um);
    c->n_partials = vec_blocks(c->q);
    IFL_LAUNCHED(c);
    Arr t = c->s; // the other buffer now holds s
    c->s = c->s2;
    c->s2 = t;
    return IFL_OK;
}

int launch_dot(ifl_ctx *c, const Arr &a, const Arr &b) {
    ProfScope ps_(c, IFL_K_SCALAR);
    k_reduce2<false><<<vec_grid(a), VEC_THREADS, 0, c->stream>>>(a, b, partials_next(c), mask_of(c));
    c->n_partials = vec_blocks(a);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_inf_norm(ifl_ctx *c, const Arr &a) {
    ProfScope ps_(c, IFL_K_SCALAR);
    k_reduce2<true><<<vec_grid(a), VEC_THREADS, 0, c->stream>>>(a, a, partials_next(c), mask_of(c));
    c->n_partials = vec_blocks(a);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_scaled_add(ifl_ctx *c, const Arr &dst, const Arr &a, const Arr &b, double s) {
    ProfScope ps_(c, IFL_K_XPAY);
    k_scaled_add<false><<<vec_grid(dst), VEC_THREADS, 0, c->stream>>>(dst, a, b, s, nullptr, mask_of(c));
    IFL_LAUNCHED(c);
    return IFL_OK;
}

__global__ void __launch_bounds__(1024) k_store_max(const double *partials, int n, double *out, DistDev dd) {
    dist_barrier_block(dd);
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) v = std_max(v, __ldcg(partials + i));
    v = block_reduce<true>(v, red);
    if (threadIdx.x == 0) out[0] = v;
}

// Folds the partials of the last reducing kernel into out_dev[0] (granular ABI ops).
int launch_finish_reduce(ifl_ctx *c, bool is_max, double *out_dev) {
    ProfScope ps_(c, IFL_K_SCALAR);
    if (is_max)
        k_store_max<<<1, 1024, 0, c->stream>>>(c->partials, c->n_partials, out_dev, c->ddev);
    else
        k_scalar<SC_STORE><<<1, 1024, 0, c->stream>>>(c->partials, c->n_partials, c->scal, out_dev, c->ddev);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

template <int MODE>
static int scalar_stage(ifl_ctx *c) {
    ProfScope ps_(c, IFL_K_SCALAR);
    k_scalar<MODE><<<1, 1024, 0, c->stream>>>(c->partials, c->n_partials, c->scal, nullptr, c->ddev);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

#define IFL_TRY(expr)            \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != IFL_OK) return rc_; \
    } while (0)

// One PCG iteration, enqueued without any host synchronisation (v3:361-376).
static int enqueue_iteration(ifl_ctx *c) {
    if (c->xpay_pending) { // chapter 3: s = z + beta s of the previous iteration rides on this matvec
        IFL_TRY(launch_xpay_matvec(c));
        c->xpay_fused++;
        c->xpay_pending = 0;
    } else {
        IFL_TRY(launch_matvec(c, c->q, c->s, true)); // q = A s ; partial q.s
    }
    IFL_TRY(scalar_stage<SC_ALPHA>(c));
    // Only while the sweep leaves at least half of the SMs free.  A sweep CTA takes its SM's whole shared-memory
    // carve-out and the streaming kernel's CTAs do not become resident next to it, so the streaming kernel runs on the
    // SMs the sweep does not use: 64 strips at 4096^2 leave 84 of 148 (the overlap hides most of the kernel), 128 strips
    // at 8192^2 leave 20 and the pair is twice as slow as the serial order (7.26 against 3.36 ms per iteration; 16384^2:
    // 27.0 against 11.4; profiles/r02_tri_experiments.txt section 11).  (>= 10: experiments, force.)
    const bool strips_resident = (c->ry1 - c->ry0 + 63) / 64 <= c->sm_count / 2 || c->overlap_axpy >= 10;
    if (c->overlap_axpy && c->tri_engine && !c->prof_on && strips_resident) {
        // p += alpha s, r -= alpha q, |r|inf on the side stream, the forward sweep concurrently on the main stream:
        // a strip starts when the 64-row band of r it reads is final (band counters).  The convergence test moves
        // behind the sweep; a converged solve has then run one forward sweep it did not need (its result is unused).
        IFL_CUDA(cudaEventRecord(c->ev_alpha, c->stream));
        IFL_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_alpha, 0));
        c->band_epoch++;
        const dim3 g = vec_grid(c->p);
        const unsigned target = c->band_epoch * g.x * (64 / VEC_ROWS);
        double *parts = partials_next(c);
        const int nparts = vec_blocks(c->p);
        // The streaming kernel is launched FIRST: it never waits for the sweep, so whichever of the two the
        // hardware makes resident first, the pair cannot deadlock (the sweep launched first on a full GPU can:
        // measured, profiles/r02_tri_experiments.txt).  Mode 3 (default): persistent form, 4 CTAs per SM walking
        // the bands top to bottom, so the first strips' bands complete first (717 vs 757 vs 771 ms per 600 iterations).
        if (c->overlap_axpy >= 2)
            k_axpy2_norm_persistent<<<c->sm_count * 4, VEC_THREADS, 0, c->side_stream>>>(
                c->p, c->s, c->r, c->q, c->scal, parts, mask_of(c), c->band_count, (int)g.x, (int)g.y,
                c->band_count + (c->H + 63) / 64, // the spare counter behind the bands: block tickets (zeroed with them)
                (c->band_epoch - 1) * (g.x * g.y + (unsigned)c->sm_count * 4));
        else
            k_axpy2_norm<<<g, VEC_THREADS, 0, c->side_stream>>>(c->p, c->s, c->r, c->q, c->scal, parts, mask_of(c), c->band_count);
        IFL_LAUNCHED(c);
        IFL_CUDA(cudaEventRecord(c->ev_axpy, c->side_stream));
        IFL_TRY(c->tri_engine == 2 ? launch_stair_forward(c, c->z, c->r, true, target) : launch_tri_forward(c, c->z, c->r, true, target));
        IFL_CUDA(cudaStreamWaitEvent(c->stream, c->ev_axpy, 0));
        c->partials = parts; // (the sweep launched no reduction in between)
        c->n_partials = nparts;
        IFL_TRY(scalar_stage<SC_CHECK>(c));
    } else {
        {
            ProfScope ps_(c, IFL_K_AXPY2_NORM);
            k_axpy2_norm<<<vec_grid(c->p), VEC_THREADS, 0, c->stream>>>(c->p, c->s, c->r, c->q, c->scal, partials_next(c),
                                                                        mask_of(c), nullptr);
            c->n_partials = vec_blocks(c->p);
            IFL_LAUNCHED(c);
        }
        IFL_TRY(scalar_stage<SC_CHECK>(c));
        IFL_TRY(launch_precon_forward(c, c->z, c->r, true));
    }
    IFL_TRY(launch_precon_backward(c, c->z, c->r, true, true)); // partial z.r
    IFL_TRY(scalar_stage<SC_BETA>(c));
    if (c->matrix_uniform && c->fuse_xpay && c->s2.p) {
        c->xpay_pending = 1; // folded into the next iteration's matvec (or flushed at the end of the solve)
    } else {
        ProfScope ps_(c, IFL_K_XPAY);
        k_scaled_add<true><<<vec_grid(c->s), VEC_THREADS, 0, c->stream>>>(c->s, c->z, c->s, 0.0, c->scal, mask_of(c));
        IFL_LAUNCHED(c);
    }
    return dist_barrier(c, true); // the next matvec reads the neighbours' boundary rows of s (fused: of z)
}

int pcg_project(ifl_ctx *c, int limit, ifl_solve_info *info) {
    cudaStream_t st = c->stream;
    // prologue v3:350-358
    // (the watchdog word is sticky: a factorisation sweep or a barrier that timed out BEFORE this
    // solve must still be seen by the first read-back below)
    IFL_CUDA(cudaMemsetAsync(c->scal, 0, offsetof(SolveScalars, watchdog), st));
    // band counters of the overlapped k_axpy2_norm: launches that were gated off at the end of the previous solve
    // never signalled, so every solve starts counting from zero (all of them have finished: the main stream waits
    // for the side stream before every convergence test, and drains at the end of a solve)
    IFL_CUDA(cudaMemsetAsync(c->band_count, 0, (size_t)((c->H + 63) / 64 + 1) * sizeof(unsigned), st));
    c->band_epoch = 0;
    c->xpay_pending = 0;
    c->xpay_fused = 0;
    IFL_CUDA(cudaMemsetAsync((char *)c->p.p + c->p.own_begin(), 0, c->p.own_end() - c->p.own_begin(), st));
    IFL_TRY(dist_barrier(c, false)); // the upstream slab's last row of cy (factorisation) is final
    IFL_TRY(launch_precon_forward(c, c->z, c->r, false));
    IFL_TRY(launch_precon_backward(c, c->z, c->r, true, false));
    IFL_TRY(scalar_stage<SC_SIGMA0>(c));
    IFL_CUDA(cudaMemcpyAsync((char *)c->s.p + c->s.own_begin(), (char *)c->z.p + c->z.own_begin(),
                             c->z.own_end() - c->z.own_begin(), cudaMemcpyDeviceToDevice, st));
    IFL_TRY(launch_inf_norm(c, c->r));
    IFL_TRY(scalar_stage<SC_CHECK0>(c));

    // The host looks at the solve scalars only through a pinned mirror and only
    // BETWEEN chunks of iterations, never inside one.  Every kernel of an iteration is
    // gated on scal->done, so a converged solve drains the rest of its chunk as empty
    // launches; the next chunk is already queued while the host waits, so the device
    // never idles on the host.
    cudaEvent_t ev[2];
    IFL_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    IFL_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int rc = IFL_OK;
    SolveScalars last;
    auto readback = [&](int sl) -> int {
        if (cudaMemcpyAsync(&c->scal_h[sl], c->scal, sizeof(SolveScalars), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(ev[sl], st) != cudaSuccess) {
            set_error("pcg_project: readback enqueue failed: %s", cudaGetErrorString(cudaGetLastError()));
            return IFL_E_CUDA;
        }
        return IFL_OK;
    };
    auto wait = [&](int sl) -> int {
        if (cudaEventSynchronize(ev[sl]) != cudaSuccess) {
            set_error("pcg_project: %s", cudaGetErrorString(cudaGetLastError()));
            return IFL_E_CUDA;
        }
        last = c->scal_h[sl];
        if (last.watchdog) {
            set_error("pcg_project: a dependency wait (wavefront hand-off or rank barrier) timed out");
            cudaMemsetAsync(&c->scal->watchdog, 0, sizeof(int), st);
            return IFL_E_WATCHDOG;
        }
        return IFL_OK;
    };
    rc = readback(0); // prologue result (initial |r|inf test)
    if (rc == IFL_OK) rc = wait(0);
    const int chunk = 16;
    int enq = 0, pending = 0, head = 0;
    while (rc == IFL_OK && !last.done) {
        while (rc == IFL_OK && pending < 2 && enq < limit) {
            const int n = imin(chunk, limit - enq);
            for (int i = 0; i < n && rc == IFL_OK; i++) rc = enqueue_iteration(c);
            enq += n;
            if (rc == IFL_OK) rc = readback((head + pending) & 1);
            pending++;
        }
        if (rc != IFL_OK || pending == 0) break;
        rc = wait(head);
        head ^= 1;
        pending--;
    }
    if (rc == IFL_OK && c->xpay_pending) { // the last iteration's s = z + beta s (a no-op launch once the solve has converged)
        ProfScope ps_(c, IFL_K_XPAY);
        k_scaled_add<true><<<vec_grid(c->s), VEC_THREADS, 0, st>>>(c->s, c->z, c->s, 0.0, c->scal, mask_of(c));
        IFL_LAUNCHED(c);
        c->xpay_pending = 0;
        pending = imax(pending, 1);
    }
    if (pending > 0 && cudaStreamSynchronize(st) != cudaSuccess && rc == IFL_OK) { // drain gated no-op launches
        set_error("pcg_project: %s", cudaGetErrorString(cudaGetLastError()));
        rc = IFL_E_CUDA;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (rc != IFL_OK) return rc;
    {   // every fused launch swapped the s / s2 records on the host; the launches enqueued after convergence were gated
        // off on the device, so an odd number of them leaves the records pointing the wrong way round
        const int executed = last.done == 1 ? imin(last.iter, c->xpay_fused) : (last.done == 2 ? 0 : c->xpay_fused);
        if ((c->xpay_fused - executed) & 1) {
            Arr t = c->s;
            c->s = c->s2;
            c->s2 = t;
        }
    }
    if (info) {
        info->max_error = last.max_error;
        if (last.done == 2) {
            info->status = IFL_SOLVE_INITIAL_SMALL;
            info->iterations = 0;
        } else if (last.done == 1) {
            info->status = IFL_SOLVE_CONVERGED;
            info->iterations = last.iter;
        } else {
            info->status = IFL_SOLVE_EXCEEDED;
            info->iterations = limit;
        }
    }
    return IFL_OK;
}

} // namespace ifl
