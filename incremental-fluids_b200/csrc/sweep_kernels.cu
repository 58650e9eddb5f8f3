// sweep_kernels.cu -- exact wavefront execution of the reference's raster-order
// recurrences (every cell depends on its left and upper neighbour, or right/lower):
//
//   KIND_FACTOR   FluidSolver::buildPreconditioner  v3:247-272   (MIC(0), tau .97, sigma .25)
//   KIND_FWD      applyPreconditioner, 1st loop      v3:276-287   (forward substitution)
//   KIND_BWD      applyPreconditioner, 2nd loop      v3:289-303   (backward substitution)
//                 + fused dotProduct(z, r)           v3:374
//   KIND_GS       one lexicographic Gauss-Seidel sweep of project(limit, timestep)
//                                                    v2:239-268 / v1:198-231
//
// Any schedule that honours the (x-1,y),(x,y-1) dependencies performs the same
// floating-point operations on the same operands as the raster loop, so results are
// bit-identical (hard part 1).  aPlusX*precon and aPlusY*precon are
// pre-multiplied into cx/cy by the factor kernel: the reference evaluates
// `_aPlusX[i]*_precon[i]*dst[i]` left to right (v3:281), so this is the same product.
//
// Execution model (B200: 148 SMs, 227 KB smem/SM, 1 CTA per SM):
//   * The padded grid is cut into strips of 32 rows.  One CTA owns one strip and
//     walks it in 32-column blocks.  Strips are handed out by an atomic ticket, so a
//     running CTA only ever waits on strips that are already running or finished
//     (no co-residency assumption, no deadlock).
//   * Inside the CTA three warps are specialised:
//       warp 0 (compute): lane t owns row t of the strip and is skewed t columns
//         behind lane t-1, i.e. the warp is one anti-diagonal.  The left neighbour is
//         the lane's own previous value (register), the upper neighbour arrives by
//         __shfl_up.  Operands are read from shared-memory tiles at skewed addresses
//         (row pitch 34 doubles -> conflict-free), one step ahead of their use, so the
//         per-step cost is the 4-deep dependent FP64 chain and nothing else.
//       warp 1 (loader): streams the strip's operand tiles HBM -> smem with
//         cp.async.bulk (TMA, one 256-byte row per copy) into an N-stage ring, and
//         polls the upstream strip's last row out of the hand-off buffer.
//       warp 2 (storer): drains finished result tiles smem -> HBM with 16-byte stores.
//     Stages are recycled through mbarriers (full / done / empty).
//   * Strip-to-strip hand-off of the swept variable uses NCCL-LL style 16-byte
//     messages {lo, epoch, hi, epoch}: the last lane publishes every value the moment
//     it is computed, the consumer validates both epochs, and no fence sits on the
//     critical path.  Epochs are unique per launch.
//   * Cells of the padded border (x >= w or y >= h) are swept too; they only ever see
//     +0.0 operands, which makes `t - c*z` an exact no-op for the real boundary cells
//     (the reference skips those terms, v3:280-283), so the inner loop has no
//     boundary predicates at all.  Pad results are never stored.
#include "ifl_internal.cuh"

#include <string.h>

namespace ifl {

enum { KIND_FWD = 0, KIND_BWD = 1, KIND_FACTOR = 2, KIND_GS = 3 };

constexpr int TP = 34;                          // tile row pitch in doubles (272 B: 16B aligned, conflict-free skew)
constexpr int TROWS = 33;                       // row 0 = upstream halo row, rows 1..32 = the strip
constexpr int TILE_DOUBLES = TROWS * TP;        // 1122
constexpr int TILE_BYTES = TILE_DOUBLES * 8;    // 8976 (multiple of 16)
constexpr int MAX_TILES = 6;
constexpr unsigned WATCHDOG_POLLS = 1u << 24;

struct TileDesc {
    double *p; // array base (pitched)
    int load;  // fetched HBM -> smem
    int row_shift; // fetch memory row (strip row + row_shift) instead of the strip row itself
    int halo;  // row 0 fetched from the upstream memory row as well
    int store; // drained smem -> HBM
};

struct SweepParams {
    TileDesc t[MAX_TILES];
    int nt;       // tiles per stage
    int nst;      // ring depth
    int swept;    // tile index of the swept variable (halo row comes from the hand-off buffer)
    int W, H, pitch, nbx, nby;
    int backward;
    uint4 *handoff; // [nby][nbx*32]
    unsigned epoch;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    SolveScalars *scal;
    int gated;         // skip when scal->done
    double *partials;  // KIND_BWD with dot: partial z.r per strip | KIND_GS: max |dp| per strip
    int with_dot;
    double scale;      // KIND_GS: timestep/(density*hx*hx)  v2:234
};

// ------------------------------------------------------------------ PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void ll_store(uint4 *dst, double v, unsigned epoch) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch)
                 : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4 *src, unsigned epoch, double &v) {
    unsigned a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
    v = __hiloint2double((int)c, (int)a);
    return b == epoch && d == epoch;
}

// ---------------------------------------------------------------- compute warp ----
// Per-lane state carried from column to column.
struct Carry {
    double zprev; // swept variable of the previous column (own row)
    double c1;    // FWD: cx of previous column | FACTOR: cx of previous column
    double c2;    // FACTOR: cy of previous column
    double acc;   // BWD: running z.r | GS: running max |p - newP|
};

// Operands of one cell, fetched one step ahead of their use.
struct Ops {
    double a, b, c, d, e, halo;
};

template <int KIND, bool DOT>
__device__ __forceinline__ void fetch(Ops &o, const double *p, const double *p_right, int lane) {
    // p -> tile 0, this lane's row, this step's column.  Tile k sits k*TILE_DOUBLES further.
    // p_right (KIND_GS only) -> tile 0, same row, next logical column (may sit in the next block).
    if (KIND == KIND_GS) {
        o.a = p[0];                      // p (old)            own cell, updated in place
        o.b = p_right[0];                // p (old)            right cell
        o.c = p[1 * TILE_DOUBLES];       // p (old)            lower cell (tile 1 = p fetched one row down)
        o.d = p[2 * TILE_DOUBLES];       // r                  own cell
        if (lane == 0) o.halo = p[-TP];  // p (new) of the upstream strip's last row
    } else if (KIND == KIND_FWD) {
        o.a = p[0];                      // a (rhs)            own cell
        o.b = p[1 * TILE_DOUBLES];       // cx                 own cell (carried to the next step)
        o.c = p[2 * TILE_DOUBLES - TP];  // cy                 upper cell (tile row 0 = halo row)
        o.d = p[3 * TILE_DOUBLES];       // precon             own cell
        if (lane == 0) o.halo = p[4 * TILE_DOUBLES - TP]; // z of the upstream strip's last row
    } else if (KIND == KIND_BWD) {
        o.a = p[0];                      // z (forward result) own cell, updated in place
        o.b = p[1 * TILE_DOUBLES];       // cx own
        o.c = p[2 * TILE_DOUBLES];       // cy own
        o.d = p[3 * TILE_DOUBLES];       // precon own
        if (DOT) o.e = p[4 * TILE_DOUBLES]; // r own
        if (lane == 0) o.halo = p[-TP];
    } else {
        o.a = p[0];                      // aDiag own
        o.b = p[1 * TILE_DOUBLES];       // aPlusX own
        o.c = p[2 * TILE_DOUBLES];       // aPlusY own
        o.d = p[1 * TILE_DOUBLES - TP];  // aPlusX upper
        o.e = p[2 * TILE_DOUBLES - TP];  // aPlusY upper
        if (lane == 0) o.halo = p[3 * TILE_DOUBLES - TP]; // precon of the upstream strip's last row
    }
}

// Per-lane constants of a Gauss-Seidel sweep.
struct GsConst {
    double scale;          // v2:234
    double d1, d2, d3, d4; // diag after 1..4 neighbour contributions: ((0+s)+s)+...  v2:251-263
    int ycnt;              // (y > 0) + (y < H-1) of this lane's row
    int yvalid;            // y < H
    int W;
};

// One cell.  `up` is the swept variable of the upper (upstream-row) neighbour, `c` the
// logical column.  Returns the new value of the swept variable; writes results into the tile.
template <int KIND, bool DOT>
__device__ __forceinline__ double cell(const Ops &o, Carry &cr, double up, double *p, int c, const GsConst &gs) {
    double znew;
    if (KIND == KIND_GS) {
        // Missing neighbours read +0.0 (left: initial carry, up: zeroed halo row, right:
        // predicated in fetch, down: zero pad row), and `off - scale*(+0.0)` is an exact
        // no-op, so the reference's four `if`s (v2:250-265) need no branches here.
        double off = 0.0;
        off = off - gs.scale * cr.zprev; // left  (already updated)  v2:252
        off = off - gs.scale * up;       // up    (already updated)  v2:256
        off = off - gs.scale * o.b;      // right (old)              v2:260
        off = off - gs.scale * o.c;      // down  (old)              v2:264
        const int cnt = gs.ycnt + (c > 0 ? 1 : 0) + (c < gs.W - 1 ? 1 : 0);
        const double diag = cnt == 4 ? gs.d4 : (cnt == 3 ? gs.d3 : (cnt == 2 ? gs.d2 : gs.d1));
        znew = (o.d - off) / diag;       // v2:267
        if (gs.yvalid && c < gs.W) cr.acc = std_max(cr.acc, fabs(o.a - znew)); // v2:269
        p[0] = znew;                     // v2:271
    } else if (KIND == KIND_FWD) {
        double t = o.a - cr.c1 * cr.zprev; // v3:281  t -= aPlusX[idx-1]*precon[idx-1]*dst[idx-1]
        t = t - o.c * up;                  // v3:283  t -= aPlusY[idx-w]*precon[idx-w]*dst[idx-w]
        znew = t * o.d;                    // v3:285
        p[4 * TILE_DOUBLES] = znew;
        cr.c1 = o.b;
    } else if (KIND == KIND_BWD) {
        double t = o.a - o.b * cr.zprev; // v3:297  t -= aPlusX[idx]*precon[idx]*dst[idx+1]
        t = t - o.c * up;                // v3:299  t -= aPlusY[idx]*precon[idx]*dst[idx+w]
        znew = t * o.d;                  // v3:301
        p[0] = znew;
        if (DOT) cr.acc += znew * o.e;   // v3:310 (partial)
    } else {
        const double tau = 0.97, sigma = 0.25; // v3:248-249
        double e = o.a;
        e = e - (cr.c1 * cr.c1 + tau * cr.c1 * cr.c2); // v3:256-258, px/py of the left cell
        const double pxu = o.d * up, pyu = o.e * up;     // v3:261-262
        e = e - (pyu * pyu + tau * pxu * pyu);           // v3:263
        if (e < sigma * o.a) e = o.a;                    // v3:266-267
        znew = 1.0 / sqrt(e);                            // v3:269
        const double cxo = o.b * znew, cyo = o.c * znew;
        p[3 * TILE_DOUBLES] = znew;
        p[4 * TILE_DOUBLES] = cxo;
        p[5 * TILE_DOUBLES] = cyo;
        cr.c1 = cxo;
        cr.c2 = cyo;
    }
    cr.zprev = znew;
    return znew;
}

// tile column of logical in-block column ci (backward sweeps walk the tile right to left)
template <int KIND>
__device__ __forceinline__ int tcol(int ci) {
    return (KIND == KIND_BWD) ? 31 - ci : ci;
}

// Tile-0 address of the in-macro-step column offset d (logical column 32m + d) in this
// lane's row.  d < 0 lies in block m-1, 0..31 in block m, >= 32 in block m+1.  EDGE 1:
// block m-1 does not exist (first macro-step), EDGE 2: block m does not exist (last
// macro-step); such positions are clamped to valid memory and their values never used.
template <int KIND, int EDGE>
__device__ __forceinline__ double *ptr_of(int d, double *s_prev, double *s_cur, double *s_next) {
    if (EDGE == 1) return (d < 0) ? s_cur + tcol<KIND>(0) : (d < 32 ? s_cur + tcol<KIND>(d) : s_next + tcol<KIND>(d - 32));
    if (EDGE == 2) return (d < 0) ? s_prev + tcol<KIND>(32 + d) : s_prev + tcol<KIND>(31);
    return (d < 0) ? s_prev + tcol<KIND>(32 + d) : (d < 32 ? s_cur + tcol<KIND>(d) : s_next + tcol<KIND>(d - 32));
}

// One macro-step = 32 steps of the skewed warp.  During macro-step m lane t works on
// logical columns 32m-t .. 32m-t+31, i.e. the tail of block m-1 (`s_prev`) and the
// head of block m (`s_cur`); the operands of each step are fetched one step early,
// which can reach into block m+1 (`s_next`, lane 0 only).  All three pointers already
// include this lane's tile-row offset.  EDGE: 0 interior, 1 first macro-step (lanes
// that have not entered the strip idle), 2 last macro-step (lanes that have left idle).
template <int KIND, bool DOT, int EDGE>
__device__ __forceinline__ void macro_step(double *s_prev, double *s_cur, double *s_next, uint64_t *full_next,
                                           unsigned parity_next, bool wait_next, int m, int lane, Carry &cr,
                                           Ops &ops, uint4 *handoff_row, bool publish, unsigned epoch,
                                           const GsConst &gs) {
    // Gauss-Seidel also reads the right neighbour, i.e. looks one column further ahead
    constexpr int WAIT_KK = (KIND == KIND_GS) ? 30 : 31;
#pragma unroll
    for (int kk = 0; kk < 32; kk++) {
        // ---- operands of step kk+1, issued before this step's stores
        Ops nxt = ops;
        if (kk == WAIT_KK && wait_next) mbar_wait(full_next, parity_next); // lane 0 is about to touch block m+1
        {
            const int d = kk + 1 - lane;
            const double *pn = ptr_of<KIND, EDGE>(d, s_prev, s_cur, s_next);
            const double *pr = (KIND == KIND_GS) ? ptr_of<KIND, EDGE>(d + 1, s_prev, s_cur, s_next) : pn;
            fetch<KIND, DOT>(nxt, pn, pr, lane);
            if (KIND == KIND_GS && !(32 * m + d + 1 < gs.W)) nxt.b = 0.0; // no right neighbour (v2:258)
        }
        // ---- this step
        const int d0 = kk - lane;
        const bool active = (EDGE == 0) ? true : (EDGE == 1 ? d0 >= 0 : d0 < 0);
        double *p = ptr_of<KIND, EDGE>(d0, s_prev, s_cur, s_next);
        double up = __shfl_up_sync(0xffffffffu, cr.zprev, 1);
        if (lane == 0) up = ops.halo;
        if (active) {
            const double z = cell<KIND, DOT>(ops, cr, up, p, 32 * m + d0, gs);
            if (publish && lane == 31) ll_store(handoff_row + 32 * m + d0, z, epoch);
        }
        ops = nxt;
    }
}

template <int KIND, bool DOT>
__device__ void compute_warp(const SweepParams &P, double *smem, uint64_t *full, uint64_t *done, int sj, int lane) {
    Carry cr;
    cr.zprev = 0.0;
    cr.c1 = 0.0;
    cr.c2 = 0.0;
    cr.acc = 0.0;
    const bool publish = sj + 1 < P.nby;
    uint4 *handoff_row = P.handoff + (size_t)sj * P.nbx * 32;
    const int nst = P.nst, nbx = P.nbx;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    double *row0 = smem + (1 + lane) * TP; // this lane's row in tile 0 of stage 0
    GsConst gs;
    gs.scale = P.scale;
    gs.d1 = 0.0 + P.scale;
    gs.d2 = gs.d1 + P.scale;
    gs.d3 = gs.d2 + P.scale;
    gs.d4 = gs.d3 + P.scale;
    {
        const int y = sj * 32 + lane; // KIND_GS sweeps forward
        gs.ycnt = (y > 0 ? 1 : 0) + (y < P.H - 1 ? 1 : 0);
        gs.yvalid = y < P.H ? 1 : 0;
        gs.W = P.W;
    }
    Ops ops;
    ops.a = ops.b = ops.c = ops.d = ops.e = ops.halo = 0.0;
    // operands of the very first step (lane 0: column 0; the others idle on column 0)
    mbar_wait(&full[0], 0);
    fetch<KIND, DOT>(ops, row0 + tcol<KIND>(0), row0 + tcol<KIND>(1), lane);
    if (KIND == KIND_GS && !(1 < P.W)) ops.b = 0.0;
    int sp = 0, sc = 0, sn = (nst > 1) ? 1 : 0; // stages of blocks m-1, m, m+1
    unsigned par_next = 0;                        // parity of full[sn] for block m+1
    for (int m = 0; m <= nbx; m++) {
        double *s_prev = row0 + sp * stage_doubles;
        double *s_cur = row0 + sc * stage_doubles;
        const bool has_next = m + 1 < nbx;
        double *s_next = has_next ? row0 + sn * stage_doubles : s_prev;
        if (m == 0)
            macro_step<KIND, DOT, 1>(s_prev, s_cur, s_next, &full[sn], par_next, has_next, m, lane, cr, ops, handoff_row,
                                     publish, P.epoch, gs);
        else if (m == nbx)
            macro_step<KIND, DOT, 2>(s_prev, s_cur, s_next, &full[sn], par_next, false, m, lane, cr, ops, handoff_row,
                                     publish, P.epoch, gs);
        else
            macro_step<KIND, DOT, 0>(s_prev, s_cur, s_next, &full[sn], par_next, has_next, m, lane, cr, ops, handoff_row,
                                     publish, P.epoch, gs);
        if (m >= 1) {
            // block m-1 is complete in smem: hand it to the storer (and, through it, the loader)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[sp]);
        }
        sp = sc;
        sc = sn;
        sn = sn + 1;
        if (sn == nst) {
            sn = 0;
            par_next ^= 1u;
        }
    }
    if (KIND == KIND_GS) {
        const double s = warp_max(cr.acc);
        if (lane == 0) P.partials[sj] = s;
    } else if (DOT) {
        const double s = warp_sum(cr.acc);
        if (lane == 0) P.partials[sj] = s;
    }
}

// ----------------------------------------------------------------- loader warp ----
__device__ void loader_warp(const SweepParams &P, double *smem, uint64_t *full, uint64_t *empty, int sj, int lane) {
    const int nst = P.nst;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    const bool bwd = P.backward != 0;
    const int ty = bwd ? (P.nby - 1 - sj) : sj;   // memory tile row of this strip
    const int y0 = ty * 32;
    const int my_row = bwd ? (y0 + 31 - lane) : (y0 + lane); // memory row behind tile row 1+lane
    const int halo_row = bwd ? (y0 + 32) : (y0 - 1);
    const bool has_up = sj > 0; // an upstream strip exists (halo rows are real data)
    // bytes that will land per stage
    unsigned bytes = 0;
    for (int k = 0; k < P.nt; k++)
        if (P.t[k].load) bytes += 32u * 256u + ((P.t[k].halo && has_up) ? 256u : 0u);
    const uint4 *up_row = P.handoff + (size_t)(sj - 1) * P.nbx * 32;
    unsigned polls = 0;

    for (int m = 0; m < P.nbx; m++) {
        const int st = m % nst;
        double *stage = smem + st * stage_doubles;
        if (m >= nst) mbar_wait(&empty[st], ((m / nst) - 1) & 1);
        const int tx = bwd ? (P.nbx - 1 - m) : m; // memory tile column
        const size_t col0 = (size_t)tx * 32;
        if (lane == 0) mbar_arrive_expect_tx(&full[st], bytes);
        __syncwarp();
        for (int k = 0; k < P.nt; k++) {
            if (!P.t[k].load) continue;
            double *tile = stage + k * TILE_DOUBLES;
            bulk_g2s(tile + (1 + lane) * TP, P.t[k].p + col0 + (size_t)(my_row + P.t[k].row_shift) * P.pitch, 256,
                     &full[st]);
            if (P.t[k].halo && has_up && lane == 0)
                bulk_g2s(tile, P.t[k].p + col0 + (size_t)halo_row * P.pitch, 256, &full[st]);
        }
        // swept variable of the upstream strip's last row -> tile row 0 (LL hand-off)
        if (has_up) {
            double v = 0.0;
            const uint4 *src = up_row + (size_t)m * 32 + lane;
            while (!ll_load(src, P.epoch, v)) {
                if (++polls > WATCHDOG_POLLS) {
                    P.scal->watchdog = 1;
                    break;
                }
            }
            stage[P.swept * TILE_DOUBLES + (bwd ? 31 - lane : lane)] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
    }
}

// ----------------------------------------------------------------- storer warp ----
__device__ void storer_warp(const SweepParams &P, double *smem, uint64_t *done, uint64_t *empty, int sj, int lane) {
    const int nst = P.nst;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    const bool bwd = P.backward != 0;
    const int ty = bwd ? (P.nby - 1 - sj) : sj;
    const int y0 = ty * 32;
    const int half = lane >> 4, l16 = lane & 15;
    for (int m = 0; m < P.nbx; m++) {
        const int st = m % nst;
        double *stage = smem + st * stage_doubles;
        mbar_wait(&done[st], (m / nst) & 1);
        const int tx = bwd ? (P.nbx - 1 - m) : m;
        const int x = tx * 32 + l16 * 2;
        for (int k = 0; k < P.nt; k++) {
            if (!P.t[k].store) continue;
            const double *tile = stage + k * TILE_DOUBLES;
            double *g = P.t[k].p;
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int lr = i * 2 + half; // lane-row inside the strip (tile row 1+lr)
                const int y = bwd ? (y0 + 31 - lr) : (y0 + lr);
                const double2 v = *reinterpret_cast<const double2 *>(tile + (1 + lr) * TP + l16 * 2);
                if (y < P.H) {
                    double *dst = g + x + (size_t)y * P.pitch;
                    if (x + 1 < P.W)
                        *reinterpret_cast<double2 *>(dst) = v;
                    else if (x < P.W)
                        dst[0] = v.x;
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
}

// ---------------------------------------------------------------------- kernel ----
template <int KIND, bool DOT>
__global__ void __launch_bounds__(96, 1) k_sweep(const SweepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[3 * 8]; // full[nst], done[nst], empty[nst] (nst <= 8)
    __shared__ int s_strip;
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = bars, *done = bars + 8, *empty = bars + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // ticket first (keeps the global count consistent even for gated launches)
        const unsigned long long tk = atomicAdd(P.ticket, 1ULL);
        s_strip = (int)(tk - P.ticket_base);
        for (int i = 0; i < P.nst; i++) {
            mbar_init(&full[i], 2);  // loader: expect_tx arrive + hand-off arrive
            mbar_init(&done[i], 1);  // compute warp
            mbar_init(&empty[i], 1); // storer warp
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (P.gated && P.scal->done) return;
    const int sj = s_strip;

    // tile row 0 of every tile starts as +0.0: the first strip has no upstream row
    for (int i = threadIdx.x; i < P.nst * P.nt * TP; i += blockDim.x) {
        const int tile = i / TP, col = i % TP;
        smem[tile * TILE_DOUBLES + col] = 0.0;
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == 0)
        compute_warp<KIND, DOT>(P, smem, full, done, sj, lane);
    else if (warp == 1)
        loader_warp(P, smem, full, empty, sj, lane);
    else
        storer_warp(P, smem, done, empty, sj, lane);
}

// ------------------------------------------------------------------- host side ----
int sweep_init(ifl_ctx *c) {
    const int nbx = (c->W + 31) / 32, nby = (c->H + 31) / 32;
    c->n_strips = nby;
    IFL_CUDA(cudaMalloc(&c->handoff, (size_t)nby * nbx * 32 * sizeof(uint4)));
    IFL_CUDA(cudaMemset(c->handoff, 0, (size_t)nby * nbx * 32 * sizeof(uint4)));
    IFL_CUDA(cudaMalloc(&c->ticket, sizeof(unsigned long long)));
    IFL_CUDA(cudaMemset(c->ticket, 0, sizeof(unsigned long long)));
    c->epoch = 0;
    return IFL_OK;
}

void sweep_free(ifl_ctx *c) {
    if (c->handoff) cudaFree(c->handoff);
    if (c->ticket) cudaFree(c->ticket);
    c->handoff = nullptr;
    c->ticket = nullptr;
}

template <int KIND, bool DOT>
static int launch_sweep(ifl_ctx *c, SweepParams &P) {
    P.W = c->W;
    P.H = c->H;
    P.pitch = c->r.pitch;
    P.nbx = (c->W + 31) / 32;
    P.nby = (c->H + 31) / 32;
    P.handoff = reinterpret_cast<uint4 *>(c->handoff);
    c->epoch++;
    P.epoch = (unsigned)(c->epoch & 0xffffffffu);
    if (P.epoch == 0) { // 0 is the value of never-written hand-off slots
        c->epoch++;
        P.epoch = 1;
    }
    P.ticket = reinterpret_cast<unsigned long long *>(c->ticket);
    P.ticket_base = c->sweep_launches * (unsigned long long)P.nby;
    c->sweep_launches++;
    P.scal = c->scal;
    const size_t smem = (size_t)P.nst * P.nt * TILE_BYTES;
    static bool attr_set[4][2] = {{false, false}, {false, false}, {false, false}, {false, false}};
    if (!attr_set[KIND][DOT]) {
        IFL_CUDA(cudaFuncSetAttribute(k_sweep<KIND, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr_set[KIND][DOT] = true;
    }
    ProfScope ps_(c, KIND == KIND_FWD ? IFL_K_PRECON_FWD : KIND == KIND_BWD ? IFL_K_PRECON_BWD : KIND == KIND_FACTOR ? IFL_K_FACTOR : IFL_K_GS_SWEEP);
    k_sweep<KIND, DOT><<<P.nby, 96, smem, c->stream>>>(P);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

static TileDesc tile(const Arr &a, int load, int halo, int store) {
    TileDesc t;
    t.p = a.p;
    t.load = load;
    t.row_shift = 0;
    t.halo = halo;
    t.store = store;
    return t;
}

int launch_mic0_factor(ifl_ctx *c) {
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.nt = 6;
    P.nst = 4;
    P.swept = 3;
    P.t[0] = tile(c->aDiag, 1, 0, 0);
    P.t[1] = tile(c->aPlusX, 1, 1, 0);
    P.t[2] = tile(c->aPlusY, 1, 1, 0);
    P.t[3] = tile(c->precon, 0, 0, 1);
    P.t[4] = tile(c->cx, 0, 0, 1);
    P.t[5] = tile(c->cy, 0, 0, 1);
    return launch_sweep<KIND_FACTOR, false>(c, P);
}

int launch_precon_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated) {
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.nt = 5;
    P.nst = 5;
    P.swept = 4;
    P.gated = gated ? 1 : 0;
    P.t[0] = tile(a, 1, 0, 0);
    P.t[1] = tile(c->cx, 1, 0, 0);
    P.t[2] = tile(c->cy, 1, 1, 0);
    P.t[3] = tile(c->precon, 1, 0, 0);
    P.t[4] = tile(dst, 0, 0, 1);
    return launch_sweep<KIND_FWD, false>(c, P);
}

int launch_precon_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.backward = 1;
    P.nst = 5;
    P.swept = 0;
    P.gated = gated ? 1 : 0;
    P.t[0] = tile(dst, 1, 0, 1);
    P.t[1] = tile(c->cx, 1, 0, 0);
    P.t[2] = tile(c->cy, 1, 0, 0);
    P.t[3] = tile(c->precon, 1, 0, 0);
    if (with_dot) {
        P.nt = 5;
        P.t[4] = tile(r_for_dot, 1, 0, 0);
        P.with_dot = 1;
        P.partials = c->partials;
        c->n_partials = (c->H + 31) / 32;
        return launch_sweep<KIND_BWD, true>(c, P);
    }
    P.nt = 4;
    return launch_sweep<KIND_BWD, false>(c, P);
}

// ------------------------------------------------------- Gauss-Seidel projection ----
// project(limit, timestep) of chapters 1-2 (v2:233-277): up to `limit` lexicographic
// sweeps over the warm-started _p, stopping when max |p - newP| < 1e-5.
__global__ void __launch_bounds__(1024) k_scalar_gs(const double *__restrict__ partials, int n, SolveScalars *sc) {
    if (sc->done) return;
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) v = std_max(v, partials[i]);
    v = block_reduce<true>(v, red);
    if (threadIdx.x != 0) return;
    sc->max_error = v;
    if (v < 1e-5)
        sc->done = 1; // v2:274 (iter keeps the zero-based index the reference prints)
    else
        sc->iter = sc->iter + 1;
}

static int enqueue_gs_sweep(ifl_ctx *c, double scale) {
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.nt = 3;
    P.nst = 5;
    P.swept = 0;
    P.gated = 1;
    P.scale = scale;
    P.t[0] = tile(c->p, 1, 0, 1);
    P.t[1] = tile(c->p, 1, 0, 0);
    P.t[1].row_shift = 1; // the row below, old values (v2:264)
    P.t[2] = tile(c->r, 1, 0, 0);
    P.partials = c->partials;
    c->n_partials = (c->H + 31) / 32;
    int rc = launch_sweep<KIND_GS, false>(c, P);
    if (rc != IFL_OK) return rc;
    ProfScope ps_(c, IFL_K_SCALAR);
    k_scalar_gs<<<1, 1024, 0, c->stream>>>(c->partials, c->n_partials, c->scal);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int gs_project(ifl_ctx *c, int limit, double timestep, double density, ifl_solve_info *info) {
    cudaStream_t st = c->stream;
    const double scale = timestep / (density * c->hx * c->hx); // v2:234
    IFL_CUDA(cudaMemsetAsync(c->scal, 0, sizeof(SolveScalars), st));
    cudaEvent_t ev[2];
    IFL_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    IFL_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int rc = IFL_OK;
    SolveScalars last;
    memset(&last, 0, sizeof last);
    const int chunk = 32;
    int enq = 0, pending = 0, head = 0;
    while (rc == IFL_OK && !last.done) {
        while (rc == IFL_OK && pending < 2 && enq < limit) {
            const int n = imin(chunk, limit - enq);
            for (int i = 0; i < n && rc == IFL_OK; i++) rc = enqueue_gs_sweep(c, scale);
            enq += n;
            const int sl = (head + pending) & 1;
            if (rc == IFL_OK &&
                (cudaMemcpyAsync(&c->scal_h[sl], c->scal, sizeof(SolveScalars), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                 cudaEventRecord(ev[sl], st) != cudaSuccess)) {
                set_error("gs_project: readback enqueue failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = IFL_E_CUDA;
            }
            pending++;
        }
        if (rc != IFL_OK || pending == 0) break;
        if (cudaEventSynchronize(ev[head]) != cudaSuccess) {
            set_error("gs_project: %s", cudaGetErrorString(cudaGetLastError()));
            rc = IFL_E_CUDA;
            break;
        }
        last = c->scal_h[head];
        head ^= 1;
        pending--;
        if (last.watchdog) {
            set_error("gs_project: wavefront dependency watchdog fired");
            rc = IFL_E_WATCHDOG;
        }
    }
    if (pending > 0 && cudaStreamSynchronize(st) != cudaSuccess && rc == IFL_OK) {
        set_error("gs_project: %s", cudaGetErrorString(cudaGetLastError()));
        rc = IFL_E_CUDA;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (rc != IFL_OK) return rc;
    if (info) {
        info->max_error = last.max_error;
        info->status = last.done ? IFL_SOLVE_CONVERGED : IFL_SOLVE_EXCEEDED;
        info->iterations = last.done ? last.iter : limit;
    }
    return IFL_OK;
}

} // namespace ifl
