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
//   * The padded grid is cut into strips of 32 rows.  One CTA owns one strip and walks it in
//     32-column blocks.  Strips are handed out by an atomic ticket (one per thread-block
//     cluster of 8 CTAs = 8 consecutive strips), so a running CTA only ever waits on strips
//     that are already running or finished (no co-residency assumption, no deadlock).
//   * Inside the CTA the warps are specialised so that the compute warp issues nothing but
//     the recurrence itself (warp 4 stays empty: it would share the compute warp's scheduler):
//       warp 0 (compute): lane t owns row t of the strip and is skewed t columns behind lane
//         t-1, i.e. the warp is one anti-diagonal.  The left neighbour is the lane's own
//         previous value (register), the upper neighbour arrives by __shfl_up.  Operands are
//         read from shared-memory tiles at skewed addresses (dense 32-double rows: the skew
//         itself staggers the banks), one step ahead of their use.  The forward and backward
//         solves update their swept tile in place.
//       warp 1 (loader): streams the strip's operand tiles HBM -> smem with 2-D tensor-map
//         TMA (cp.async.bulk.tensor, one 33x32 box per operand per block) into an N-stage ring.
//       warp 2 (storer): drains finished result tiles smem -> HBM with 16-byte stores and folds
//         dotProduct(z, r) for the backward solve.
//       warp 3 (publisher) / warp 5 (poller): the strip-to-strip hand-off, below.
//     Stages are recycled through mbarriers (full / done / empty).
//   * Strip-to-strip hand-off of the swept variable: NCCL-LL style 16-byte messages
//     {lo, tag, hi, tag}, 8 columns at a time, validated by the consumer's poller warp and
//     released to its compute warp through a shared counter; no fence sits on the critical
//     path.  Inside a cluster the publisher stores the messages straight into the downstream
//     CTA's shared memory (st.shared::cluster, tag = block number); between clusters -- and
//     between the slabs of two GPUs (dist.cu) -- they go through L2 / NVLink into a hand-off
//     array with a per-launch epoch as tag.
//   * Cells of the padded border (x >= w or y >= h) are swept too; they only ever see
//     +0.0 operands, which makes `t - c*z` an exact no-op for the real boundary cells
//     (the reference skips those terms, v3:280-283), so the inner loop has no
//     boundary predicates at all.  Pad results are never stored.
#include "ifl_internal.cuh"
#include "sweep_common.cuh"

// 1: the compute warp also records a timestamp every nbx/8 macro-steps (ifl_debug_sweep_times slots
// 2..10).  Off by default: even this one branch per macro-step shifts the hot loop's code generation.
#ifndef IFL_SWEEP_DIAG
#define IFL_SWEEP_DIAG 0
#endif

#include <cuda.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

namespace ifl {

enum { KIND_FWD = 0, KIND_BWD = 1, KIND_FACTOR = 2, KIND_GS = 3, KIND_FACTOR_M = 4 }; // _M: with solid cells (v5:715-744)

// A tile is one TMA box: 33 rows x 32 doubles, dense (256-byte rows).  Forward kinds
// fetch memory rows y0-1 .. y0+31 (tile row 0 = the upstream strip's last row, lane t
// owns tile row 1+t); the backward kind fetches y0 .. y0+32 (lane t owns tile row 31-t,
// tile row 32 = upstream).  Either way the upstream-row operand of a lane sits one tile
// row "before" its own row in sweep order: UP_OFF doubles away.
constexpr int TP = 32;
constexpr int TROWS = 33;
constexpr int TILE_DOUBLES = TROWS * TP;     // 1056
constexpr int TILE_BYTES = TILE_DOUBLES * 8; // 8448 (multiple of 128)
constexpr int MAX_TILES = 7;
constexpr int MAX_STAGES = 8;
constexpr int HG = 8; // hand-off granularity in columns (publisher and consumer side)
constexpr int HR = 16; // hand-off ring depth in blocks (512 columns): deep enough that back-pressure never binds

struct TileDesc {
    double *p;     // array base (pitched) -- used by the storer
    int load;      // fetched HBM -> smem by TMA
    int row_shift; // fetch memory rows shifted by this many rows (Gauss-Seidel: the row below)
    int store;     // drained smem -> HBM
    double *p2;    // optional second store target, written only where the value is non-zero
};

struct SweepParams {
    CUtensorMap map[MAX_TILES]; // one per loaded tile (must stay first: 64-byte aligned)
    TileDesc t[MAX_TILES];
    int nt;    // tiles per stage
    int nst;   // ring depth
    int W, H, pitch, nbx, nby;
    uint4 *handoff; // [nby][nbx*32]
    // row-slab multi-GPU: this launch sweeps the strips [sj_base, sj_base + nloc) of the nby
    // strips (sweep order).  The last of them publishes into the downstream rank's hand-off
    // array, the first polls messages that arrived over NVLink.
    int sj_base, nloc;
    uint4 *handoff_down;
    unsigned epoch;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    SolveScalars *scal;
    int gated;        // skip when scal->done
    double *partials; // KIND_BWD with dot: partial z.r per strip | KIND_GS: max |dp| per strip
    double scale;     // KIND_GS: timestep/(density*hx*hx)  v2:234
    int mask_tile;    // >= 0: results are stored only where this tile is non-zero (fluid cells), else -1
    int cs;           // thread-block cluster size (1 = no cluster): strips of one cluster hand off through DSMEM
    int head_delay;   // SM cycles the head strip (no upstream neighbour) idles per macro-step, see sweep_init
    unsigned long long *times; // diagnostics: [nby][2] globaltimer ns at strip start / end (or null)
};

// ---------------------------------------------------------------- compute warp ----
template <int KIND>
struct Geo {
    static constexpr bool BWD = (KIND == KIND_BWD);
    static constexpr int UP_OFF = BWD ? TP : -TP; // own row -> upstream row, in doubles
    __device__ static __forceinline__ int lane_row(int lane) { return BWD ? 31 - lane : 1 + lane; }
    // tile column of logical in-block column ci (backward sweeps walk the tile right to left)
    __device__ static __forceinline__ int tcol(int ci) { return BWD ? 31 - ci : ci; }
};

// Per-lane state carried from column to column.
struct Carry {
    double zprev; // swept variable of the previous column (own row)
    double c1;    // FWD / FACTOR: cx of the previous column
    double c2;    // FACTOR: cy of the previous column
    double acc;   // BWD: running z.r | GS: running max |p - newP|
};

// Operands of one cell, fetched one step ahead of their use.
struct Ops {
    double a, b, c, d, e, f, halo;
};

// p -> tile 0, this lane's row, this step's column (shared-space byte address); tile k
// sits k*TILE_BYTES further.  p_right (KIND_GS only) -> tile 0, same row, next logical
// column (may be in the next block).  ph (lane 0 only) -> the swept variable of the
// upstream strip's last row at this column.
template <int KIND, bool DOT>
__device__ __forceinline__ void fetch(Ops &o, uint32_t p, uint32_t p_right, uint32_t ph, int lane) {
    constexpr int UP = Geo<KIND>::UP_OFF * 8;
    if (KIND == KIND_GS) {
        o.a = lds_f64(p);                  // p (old)  own cell, updated in place
        o.b = lds_f64(p_right);            // p (old)  right cell
        o.c = lds_f64(p + 1 * TILE_BYTES); // p (old)  lower cell (tile 1 = p fetched one row down)
        o.d = lds_f64(p + 2 * TILE_BYTES); // r        own cell
    } else if (KIND == KIND_FWD) {
        o.a = lds_f64(p);                       // a (rhs)  own cell
        o.b = lds_f64(p + 1 * TILE_BYTES);      // cx       own cell (carried to the next step)
        o.c = lds_f64(p + 2 * TILE_BYTES + UP); // cy       upper cell
        o.d = lds_f64(p + 3 * TILE_BYTES);      // precon   own cell
    } else if (KIND == KIND_BWD) {
        o.a = lds_f64(p);                            // z (forward result) own cell, updated in place
        o.b = lds_f64(p + 1 * TILE_BYTES);           // cx own
        o.c = lds_f64(p + 2 * TILE_BYTES);           // cy own
        o.d = lds_f64(p + 3 * TILE_BYTES);           // precon own
    } else {
        o.a = lds_f64(p);                       // aDiag own
        o.b = lds_f64(p + 1 * TILE_BYTES);      // aPlusX own
        o.c = lds_f64(p + 2 * TILE_BYTES);      // aPlusY own
        o.d = lds_f64(p + 1 * TILE_BYTES + UP); // aPlusX upper
        o.e = lds_f64(p + 2 * TILE_BYTES + UP); // aPlusY upper
        if (KIND == KIND_FACTOR_M) o.f = lds_f64(p + 6 * TILE_BYTES); // 1.0 at fluid cells, else 0.0
    }
    o.halo = lds_f64(ph); // same address in every lane (broadcast); only lane 0 uses it
}

// Per-lane constants of a Gauss-Seidel sweep.
struct GsConst {
    double scale;          // v2:234
    double d1, d2, d3, d4; // diag after 1..4 neighbour contributions: ((0+s)+s)+...  v2:251-263
    int ycnt;              // (y > 0) + (y < H-1) of this lane's row
    int yvalid;            // y < H
    int W;
    int ncols;             // padded sweep width (32 * nbx)
    int cluster;           // hand-off counter is bumped remotely (DSMEM): needs acquire loads
};

// One cell.  `up` is the swept variable of the upper (upstream-row) neighbour, `c` the
// logical column.  Returns the new value of the swept variable; writes results into the tile.
// EDGE 0: interior macro-step.  EDGE 1 / 2: first / last macro-step, where lanes that have
// not entered (have left) the strip run the same instructions on aliased operands: their
// stores are predicated off and whatever they compute is never consumed -- a lane only
// reads the upper lane's value of the previous step, and lane t-1 is active at step k-1
// exactly when lane t is active at step k.  A lane's carried state is cleared at the step
// it enters the strip (`first`); that select sits on the left-neighbour path (4 FP64 ops
// per step), not on the shuffle path (shuffle + 3 ops) that sets the step time.
template <int KIND, bool DOT, int EDGE>
__device__ __forceinline__ double cell(const Ops &o, Carry &cr, double up, uint32_t p, int c, const GsConst &gs,
                                       bool active, bool first) {
    constexpr bool ALWAYS = EDGE == 0;
    double znew;
    if (EDGE == 1) {
        cr.zprev = sel_f64(first, 0.0, cr.zprev);
        cr.c1 = sel_f64(first, 0.0, cr.c1);
        if (KIND == KIND_FACTOR || KIND == KIND_FACTOR_M) cr.c2 = sel_f64(first, 0.0, cr.c2);
        if (KIND == KIND_GS) cr.acc = sel_f64(first, 0.0, cr.acc);
    }
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
        if (gs.yvalid && c < gs.W && (ALWAYS || active)) cr.acc = std_max(cr.acc, fabs(o.a - znew)); // v2:269
        sts_f64_p<ALWAYS>(p, znew, active); // v2:271
    } else if (KIND == KIND_FWD) {
        double t = o.a - cr.c1 * cr.zprev; // v3:281  t -= aPlusX[idx-1]*precon[idx-1]*dst[idx-1]
        t = t - o.c * up;                  // v3:283  t -= aPlusY[idx-w]*precon[idx-w]*dst[idx-w]
        znew = t * o.d;                    // v3:285
        sts_f64_p<ALWAYS>(p, znew, active); // in place: the rhs tile becomes the result tile
        cr.c1 = o.b;
    } else if (KIND == KIND_BWD) {
        double t = o.a - o.b * cr.zprev; // v3:297  t -= aPlusX[idx]*precon[idx]*dst[idx+1]
        t = t - o.c * up;                // v3:299  t -= aPlusY[idx]*precon[idx]*dst[idx+w]
        znew = t * o.d;                  // v3:301
        sts_f64_p<ALWAYS>(p, znew, active); // (z.r of v3:374 is accumulated by the storer warp)
    } else {
        const double tau = 0.97, sigma = 0.25; // v3:248-249
        double e = o.a;
        e = e - (cr.c1 * cr.c1 + tau * cr.c1 * cr.c2); // v3:256-258, px/py of the left cell
        const double pxu = o.d * up, pyu = o.e * up;     // v3:261-262
        e = e - (pyu * pyu + tau * pxu * pyu);           // v3:263
        if (e < sigma * o.a) e = o.a;                    // v3:266-267
        znew = 1.0 / sqrt(e);                            // v3:269
        double cxo = o.b * znew, cyo = o.c * znew;
        if (KIND == KIND_FACTOR_M && o.f == 0.0) {
            // non-fluid cell (v5:722-723 `continue`): the reference skips every term that
            // involves it; feeding +0.0 downstream makes those terms exact no-ops
            znew = 0.0;
            cxo = 0.0;
            cyo = 0.0;
        }
        sts_f64_p<ALWAYS>(p + 3 * TILE_BYTES, znew, active);
        sts_f64_p<ALWAYS>(p + 4 * TILE_BYTES, cxo, active);
        sts_f64_p<ALWAYS>(p + 5 * TILE_BYTES, cyo, active);
        cr.c1 = cxo;
        cr.c2 = cyo;
    }
    cr.zprev = znew;
    return znew;
}

// Per-lane tile-0 base pointers of one macro-step.  Logical position j (= kk + look-ahead,
// 0..33) of this lane lies in block m-1 when j < lane, in block m when j - lane < 32 and
// in block m+1 otherwise (lanes 0/1 only); each base already contains the lane's row and
// its skew, so the address is always `base + DIR*j` and j folds into the instruction's
// immediate offset.  Positions that do not exist (first / last macro-step) alias valid
// memory and are never used.
struct LaneBases { // shared-space byte addresses
    uint32_t A; // block m
    uint32_t B; // block m-1
    uint32_t N; // block m+1
};

template <int KIND>
__device__ __forceinline__ uint32_t pos(const LaneBases &lb, int j, int lane) {
    constexpr int DIR = Geo<KIND>::BWD ? -8 : 8;
    uint32_t base = (lane > j) ? lb.B : lb.A;
    if (j >= 32) base = (lane <= j - 32) ? lb.N : base; // j is a compile-time constant
    return base + (uint32_t)(DIR * j);
}

// One macro-step = 32 steps of the skewed warp.  During macro-step m lane t works on
// logical columns 32m-t .. 32m-t+31, i.e. the tail of block m-1 and the head of block m.
// Per step the only serial dependency is  z -> shuffle -> 3 FP64 ops -> z ; the shuffle
// is issued first and the operand fetch for the NEXT step (shared-memory loads at
// branch-free addresses) runs in its shadow.  EDGE: 0 interior, 1 first macro-step
// (lanes that have not entered the strip idle), 2 last macro-step (lanes that have left
// the strip idle).
template <int KIND, bool DOT, int EDGE>
__device__ __forceinline__ void macro_step(const LaneBases &lb, uint32_t h_cur, uint32_t h_next,
                                           uint64_t *full_next, unsigned parity_next, bool wait_next, int m, int lane,
                                           Carry &cr, Ops &ops, uint32_t progress_addr, uint32_t halo_cols_addr,
                                           bool has_up, const GsConst &gs, volatile int *dead, SolveScalars *scal,
                                           unsigned long long *times, int probe) {
    typedef Geo<KIND> G;
    constexpr int DIR = G::BWD ? -8 : 8;
    // Gauss-Seidel also reads the right neighbour, i.e. looks one column further ahead
    constexpr int WAIT_KK = (KIND == KIND_GS) ? 30 : 31;
    uint32_t p = pos<KIND>(lb, 0, lane);
    unsigned counter_seen = 0;
#pragma unroll
    for (int kk = 0; kk < 32; kk++) {
        if (kk == WAIT_KK && wait_next) mbar_wait(full_next, parity_next, dead, scal); // lane 0 is about to touch block m+1
        // lane 0 is about to fetch the first hand-off value of the next group of HG columns.
        // The counter is read four steps early so that its shared-memory latency is hidden;
        // only a consumer that has caught up with its producer falls into the polling loop.
        if (((kk + 5) % HG) == 0 && has_up && EDGE != 2 && !gs.cluster) counter_seen = lds_u32_volatile(halo_cols_addr);
        if (((kk + 1) % HG) == 0 && has_up && EDGE != 2) {
            const unsigned need = (unsigned)imin(32 * m + kk + 1 + HG, gs.ncols);
            if (gs.cluster)
                wait_counter<true>(halo_cols_addr, need, dead, scal);
            else if (counter_seen < need)
                wait_counter<false>(halo_cols_addr, need, dead, scal);
        }
        // ---- critical path first: the upper neighbour's value of the previous step
        double up = __shfl_up_sync(0xffffffffu, cr.zprev, 1);
        // ---- operands of step kk+1, in the shadow of the shuffle
        Ops nxt;
        const uint32_t pn = pos<KIND>(lb, kk + 1, lane);
        {
            const uint32_t pr = (KIND == KIND_GS) ? pos<KIND>(lb, kk + 2, lane) : pn;
            const uint32_t ph = (kk + 1 < 32) ? h_cur + (uint32_t)(DIR * (kk + 1)) : h_next + (uint32_t)(DIR * (kk + 1 - 32));
            fetch<KIND, DOT>(nxt, pn, pr, ph, lane);
            if (KIND == KIND_GS && !(32 * m + kk + 2 - lane < gs.W)) nxt.b = 0.0; // no right neighbour (v2:258)
        }
        // ---- this step
        const int d0 = kk - lane;
        const bool active = (EDGE == 0) ? true : (EDGE == 1 ? d0 >= 0 : d0 < 0);
        up = sel_f64(lane == 0, ops.halo, up);
        cell<KIND, DOT, EDGE>(ops, cr, up, p, 32 * m + d0, gs, active, EDGE == 1 && d0 == 0);
        // the strip's last row (lane 31) has just completed another group of HG columns
        if (((kk + 2) % HG) == 0) sts_u32_volatile(progress_addr, (unsigned)imax(32 * m + kk - 30, 0));
        ops = nxt;
        p = pn;
    }
}

template <int KIND, bool DOT>
__device__ void compute_warp(const SweepParams &P, double *smem, double *halo_s, uint64_t *full, uint64_t *done, int sj,
                             int lane, volatile int *dead, unsigned *counters) {
    typedef Geo<KIND> G;
    constexpr int DIR = G::BWD ? -8 : 8; // bytes per logical column step
    constexpr int COL0 = G::BWD ? 31 : 0; // tile column of logical in-block column 0
    Carry cr;
    cr.zprev = 0.0;
    cr.c1 = 0.0;
    cr.c2 = 0.0;
    cr.acc = 0.0;
    const bool has_up = sj > 0;
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    const int nst = P.nst, nbx = P.nbx;
    const uint32_t stage_bytes = (uint32_t)P.nt * TILE_BYTES;
    // this lane's row in tile 0 of stage 0 at logical column 0, and the hand-off row of stage 0
    const uint32_t row0 = smem_u32(smem) + (uint32_t)(G::lane_row(lane) * TP + COL0) * 8u;
    const uint32_t halo0 = smem_u32(halo_s) + (uint32_t)COL0 * 8u;
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
        gs.ncols = P.nbx * 32;
        gs.cluster = 0; // the hand-off counter is always bumped by this CTA's own poller warp
    }
    Ops ops;
    ops.a = ops.b = ops.c = ops.d = ops.e = ops.f = ops.halo = 0.0;
    // operands of the very first step (lane 0: column 0; the others idle on column 0)
    // Everything that does not depend on the upstream strip happens BEFORE the wait for its
    // first hand-off group: that wait sits on the critical path of the whole sweep.
    mbar_wait(&full[0], 0, dead, P.scal);
    fetch<KIND, DOT>(ops, row0, row0 + (uint32_t)DIR, halo0, lane);
    if (KIND == KIND_GS && !(1 < P.W)) ops.b = 0.0;
    if (has_up) {
        wait_counter<false>(halo_cols_addr, HG, dead, P.scal);
        ops.halo = lds_f64(halo0);
    }
    int sp = 0, sc = 0, sn = (nst > 1) ? 1 : 0; // stages of blocks m-1, m, m+1
    unsigned par_next = 0;                        // parity of full[sn] for block m+1
    const int probe = (nbx / 2) * 32; // diagnostics: hand-off timestamps for the group ending at this column
#if IFL_SWEEP_DIAG
    const int ck = (nbx + 7) / 8; // diagnostics: a timestamp every nbx/8 macro-steps
#endif
    for (int m = 0; m <= nbx; m++) {
#if IFL_SWEEP_DIAG
        if (P.times && lane == 0 && (m % ck) == 0 && m / ck < 12) {
            unsigned long long tt;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tt));
            P.times[16 * sj + 2 + m / ck] = tt;
        }
#endif
        if (!has_up && P.head_delay > 0) { // pace-setter: see sweep_init
            const long long t_ = clock64();
            while (clock64() - t_ < P.head_delay) {}
        }
        const bool has_next = m + 1 < nbx;
        const uint32_t s_prev = row0 + sp * stage_bytes;
        const uint32_t s_cur = row0 + sc * stage_bytes;
        const uint32_t s_next = row0 + sn * stage_bytes;
        const uint32_t h_cur = halo0 + (uint32_t)(m % HR) * 256u;
        const uint32_t h_next = halo0 + (uint32_t)((has_next ? m + 1 : m) % HR) * 256u;
        const uint32_t skew = (uint32_t)(DIR * lane), blk = (uint32_t)(DIR * 32);
        LaneBases lb;
        if (m == 0) { // no block m-1: idle lanes alias block 0
            lb.A = s_cur - skew;
            lb.B = lb.A;
            lb.N = has_next ? s_next - blk - skew : lb.A;
            macro_step<KIND, DOT, 1>(lb, h_cur, h_next, &full[sn], par_next, has_next, m, lane, cr, ops, progress_addr,
                                     halo_cols_addr, has_up, gs, dead, P.scal, P.times ? P.times + 16 * sj : nullptr, probe);
        } else if (m == nbx) { // no block m: idle lanes alias block m-1
            lb.B = s_prev + blk - skew;
            lb.A = s_prev - skew;
            lb.N = lb.A;
            macro_step<KIND, DOT, 2>(lb, h_cur, h_next, &full[sn], par_next, false, m, lane, cr, ops, progress_addr,
                                     halo_cols_addr, has_up, gs, dead, P.scal, P.times ? P.times + 16 * sj : nullptr, probe);
        } else {
            lb.A = s_cur - skew;
            lb.B = s_prev + blk - skew;
            lb.N = has_next ? s_next - blk - skew : lb.A;
            macro_step<KIND, DOT, 0>(lb, h_cur, h_next, &full[sn], par_next, has_next, m, lane, cr, ops, progress_addr,
                                     halo_cols_addr, has_up, gs, dead, P.scal, P.times ? P.times + 16 * sj : nullptr, probe);
        }
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
    }
}

// ----------------------------------------------------------------- loader warp ----
// Keeps the TMA ring `nst-3` blocks ahead of the compute warp.  The compute warp releases
// block j only at the end of macro-step j+1, and inside that macro-step it already waits
// for block j+2; a window deeper than nst-3 would wait for a stage that cannot drain.
template <int KIND>
__device__ void loader_warp(const SweepParams &P, double *smem, uint64_t *full, uint64_t *empty, int sj, int lane,
                            volatile int *dead) {
    typedef Geo<KIND> G;
    if (lane != 0) return;
    const int nst = P.nst, nbx = P.nbx;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    const int ty = G::BWD ? (P.nby - 1 - sj) : sj; // memory tile row of this strip
    const int box_y = G::BWD ? ty * 32 : ty * 32 - 1;
    int nload = 0;
    for (int k = 0; k < P.nt; k++) nload += P.t[k].load ? 1 : 0;
    const unsigned bytes = (unsigned)nload * TILE_BYTES;
    for (int m = 0; m < nbx; m++) {
        const int st = m % nst;
        if (m >= nst) mbar_wait(&empty[st], ((m / nst) - 1) & 1, dead, P.scal);
        mbar_arrive_expect_tx(&full[st], bytes);
        const int box_x = (G::BWD ? (nbx - 1 - m) : m) * 32;
        double *stage = smem + st * stage_doubles;
        for (int k = 0; k < P.nt; k++)
            if (P.t[k].load)
                tma_load_2d(stage + k * TILE_DOUBLES, &P.map[k], box_x, box_y + P.t[k].row_shift, &full[st]);
    }
}

// ----------------------------------------------------------------- poller warp ----
// Receives the swept variable of the upstream strip's last row (LL hand-off) into
// halo_s[stage] and releases the compute warp group by group (HG columns) through
// counters[1].  It runs independently of the TMA ring; it only has to stay less than a
// ring's worth of blocks ahead of the strip's own last row (counters[0]) so that a
// hand-off row is never overwritten while lane 0 may still read it.
template <int KIND>
__device__ void poller_warp(const SweepParams &P, double *halo_s, int sj, int lane, volatile int *dead,
                            unsigned *counters, const uint4 *ll_ring) {
    typedef Geo<KIND> G;
    const int nbx = P.nbx;
    // ll_ring != null: the upstream strip runs in the same thread-block cluster and stores its
    // messages {lo, tag, hi, tag} straight into this CTA's shared memory (tag = block + 1: a ring
    // slot is reused every HR blocks); otherwise they arrive through L2 / NVLink, one slot per column.
    const uint4 *up_row = P.handoff + (size_t)(sj - 1) * nbx * 32;
    const bool remote = sj == P.sj_base; // the upstream strip belongs to another rank
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    Watch watch;
    for (int m = 0; m < nbx; m++) {
        const int st = m % HR;
        if (m >= HR) wait_counter(progress_addr, (unsigned)(32 * (m - HR + 1)), dead, P.scal);
        const uint4 *src = up_row + (size_t)m * 32 + lane;
        const uint32_t src_s = ll_ring ? smem_u32(ll_ring + st * 32 + lane) : 0;
        const unsigned tag = ll_ring ? (unsigned)(m + 1) : P.epoch;
        bool have = false;
        unsigned published = 0; // columns of this block already released
        while (published < 32) {
            if (!have) {
                double v;
                bool ok;
                if (ll_ring) {
                    unsigned a, b, c2, d;
                    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "r"(src_s) : "memory");
                    v = __hiloint2double((int)c2, (int)a);
                    ok = b == tag && d == tag;
                } else {
                    ok = remote ? ll_load_sys(src, tag, v) : ll_load(src, tag, v);
                }
                if (ok) {
                    halo_s[st * 32 + G::tcol(lane)] = v;
                    have = true;
                    watch = Watch();
                } else if (watch.expired(dead)) {
                    *dead = 1;
                    P.scal->watchdog = 1;
                    have = true;
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, have);
            const unsigned lead = (mask == 0xffffffffu) ? 32u : (unsigned)(__ffs(~mask) - 1); // leading valid columns
            const unsigned groups = lead / HG * HG;
            if (groups > published) {
                __threadfence_block(); // halo_s values before the counter
                if (lane == 0) sts_u32_volatile(halo_cols_addr, (unsigned)(32 * m) + groups);
                if (P.times && lane == 0 && 32 * m + (int)published < (nbx / 2) * 32 && 32 * m + (int)groups >= (nbx / 2) * 32)
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(P.times[16 * sj + 13]));
                published = groups;
            }
        }
    }
}

// ----------------------------------------------------------------- storer warp ----
// Drains finished result tiles and, for the backward solve of the PCG loop, folds in
// dotProduct(z, r) (v3:374): each lane accumulates its cells in a fixed order, the warp
// sum goes to partials[strip].  Pad cells hold exact zeros and do not perturb the sum.
// MASKED (chapters 4+): a result is stored only where tile 3 (pe) is non-zero, i.e. at
// fluid cells; non-fluid cells keep their old value (v5:751-752).  The masked
// factorisation additionally writes precon itself where the value is non-zero (v5:741).
// All loads of a tile are issued before its stores so that the drain takes a few hundred
// cycles per block and never throttles the compute warp.
template <int KIND, bool DOT, bool MASKED>
__device__ void storer_warp(const SweepParams &P, double *smem, uint64_t *done, uint64_t *empty, int sj, int lane,
                            volatile int *dead) {
    typedef Geo<KIND> G;
    const int nst = P.nst;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    const int ty = G::BWD ? (P.nby - 1 - sj) : sj;
    const int y0 = ty * 32;
    const int half = lane >> 4, l16 = lane & 15;
    const int toff = (G::BWD ? half : 1 + half) * TP + l16 * 2; // this lane's first element inside a tile
    double acc = 0.0;
    for (int m = 0; m < P.nbx; m++) {
        const int st = m % nst;
        const double *stage = smem + st * stage_doubles;
        mbar_wait(&done[st], (m / nst) & 1, dead, P.scal);
        const int tx = G::BWD ? (P.nbx - 1 - m) : m;
        const int x = tx * 32 + l16 * 2;
        const bool x0 = x < P.W, x1 = x + 1 < P.W;
        for (int k = 0; k < P.nt; k++) {
            if (!P.t[k].store) continue;
            const double *tile = stage + k * TILE_DOUBLES + toff;
            double2 v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = *reinterpret_cast<const double2 *>(tile + i * 2 * TP);
            if (DOT && k == 0) { // KIND_BWD with dot: tile 4 holds r
                const double *rt = stage + 4 * TILE_DOUBLES + toff;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const double2 rv = *reinterpret_cast<const double2 *>(rt + i * 2 * TP);
                    acc += v[i].x * rv.x;
                    acc += v[i].y * rv.y;
                }
            }
            double *g = P.t[k].p + x + (size_t)(y0 + half) * P.pitch;
            if (!MASKED || KIND == KIND_FACTOR_M) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    double *dst = g + (size_t)(i * 2) * P.pitch;
                    if (y0 + half + i * 2 < P.H) {
                        if (x1)
                            *reinterpret_cast<double2 *>(dst) = v[i];
                        else if (x0)
                            dst[0] = v[i].x;
                    }
                }
                if (KIND == KIND_FACTOR_M && P.t[k].p2) { // precon: fluid cells only
                    double *g2 = P.t[k].p2 + x + (size_t)(y0 + half) * P.pitch;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        double *dst = g2 + (size_t)(i * 2) * P.pitch;
                        if (y0 + half + i * 2 < P.H) {
                            if (x0 && v[i].x != 0.0) dst[0] = v[i].x;
                            if (x1 && v[i].y != 0.0) dst[1] = v[i].y;
                        }
                    }
                }
            } else {
                const double *mt = stage + 3 * TILE_DOUBLES + toff; // pe: non-zero at fluid cells
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const double2 mk = *reinterpret_cast<const double2 *>(mt + i * 2 * TP);
                    double *dst = g + (size_t)(i * 2) * P.pitch;
                    if (y0 + half + i * 2 < P.H) {
                        if (x0 && mk.x != 0.0) dst[0] = v[i].x;
                        if (x1 && mk.y != 0.0) dst[1] = v[i].y;
                    }
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (DOT) {
        const double sum = warp_sum(acc);
        if (lane == 0) P.partials[sj] = sum;
    }
}

// -------------------------------------------------------------- publisher warp ----
// Forwards the strip's last row to the downstream strip, HG columns at a time, as soon as
// the compute warp's progress counter says they are final.
//   * downstream strip in the same thread-block cluster: 16-byte messages {lo, tag, hi, tag}
//     stored straight into that CTA's message ring (distributed shared memory), validated by its
//     poller warp in its own shared memory -- no fence, nothing goes through L2;
//   * otherwise: NCCL-LL style 16-byte messages {lo, epoch, hi, epoch} through L2, picked up
//     by the downstream CTA's poller warp.
// It holds each stage until its 32 columns have been sent (second arrival on done[]).
template <int KIND>
__device__ void publisher_warp(const SweepParams &P, double *smem, double *halo_s, uint64_t *done, int sj, int lane,
                               volatile int *dead, unsigned *counters, unsigned rank) {
    typedef Geo<KIND> G;
    const int nst = P.nst;
    const int stage_doubles = P.nt * TILE_DOUBLES;
    const int ncols = P.nbx * 32;
    int swept = (KIND == KIND_FACTOR || KIND == KIND_FACTOR_M) ? 3 : 0;
    const double *last_row = smem + swept * TILE_DOUBLES + G::lane_row(31) * TP; // tile row of the strip's last row
    const bool remote = sj + 1 == P.sj_base + P.nloc; // the downstream strip belongs to another rank
    uint4 *out = (remote ? P.handoff_down : P.handoff) + (size_t)sj * ncols;
    const uint32_t progress_addr = smem_u32(&counters[0]);
    // downstream strip in the same cluster (and on the same rank): its LL ring and progress counter
    const bool dsmem = P.cs > 1 && rank + 1 < (unsigned)P.cs && !remote;
    const uint32_t r_ll = dsmem ? mapa(smem_u32(halo_s + HR * 32), rank + 1) : 0;
    const uint32_t r_progress = dsmem ? mapa(smem_u32(&counters[0]), rank + 1) : 0;
    int down_progress = 0; // last value read from the downstream strip's own progress counter
    int sent = 0;        // columns forwarded so far
    int blk = 0;         // block `sent` lies in
    int st = 0;          // its stage
    const double *row = last_row;
    Watch watch;
    while (sent < ncols) {
        const int prog = (int)lds_u32_volatile(progress_addr);
        if (prog > sent) {
            while (sent < prog) {
                const int blk_end = (blk + 1) * 32;
                const int upto = prog < blk_end ? prog : blk_end; // stay inside one block
                const int c = sent + lane;
                if (dsmem) {
                    // the ring slot of block `blk` is free once the downstream strip's last row has
                    // left block blk - HR (its lane 0 is further ahead still)
                    while (blk >= HR && down_progress < 32 * (blk - HR + 1)) {
                        down_progress = (int)ld_remote_u32(r_progress);
                        if (watch.expired(dead)) {
                            *dead = 1;
                            P.scal->watchdog = 1;
                            break;
                        }
                    }
                    // flag-in-data message {lo, tag, hi, tag}, tag = block + 1: no fence, no counter
                    if (c < upto) {
                        const double v = row[G::tcol(c & 31)];
                        const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v), tag = (unsigned)(blk + 1);
                        asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(r_ll + (uint32_t)((blk % HR) * 32 + (c & 31)) * 16u),
                                     "r"(lo), "r"(tag), "r"(hi), "r"(tag)
                                     : "memory");
                    }
                } else {
                    if (c < upto) {
                        if (remote)
                            ll_store_sys(out + c, row[G::tcol(c & 31)], P.epoch);
                        else
                            ll_store(out + c, row[G::tcol(c & 31)], P.epoch);
                    }
                }
                sent = upto;
                if (sent == blk_end) { // all 32 columns of this block are out: the stage may drain
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&done[st]);
                    // refresh the downstream strip's progress off the critical path: the value is
                    // first needed when the next block's ring slot is checked, a block from now
                    if (dsmem && blk + 1 >= HR && blk + 1 < P.nbx) down_progress = (int)ld_remote_u32(r_progress); // (not after the last block: the downstream CTA may be gone)
                    blk++;
                    if (++st == nst) st = 0;
                    row = last_row + st * stage_doubles;
                }
            }
            if (P.times && lane == 0 && prog == (P.nbx / 2) * 32)
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(P.times[16 * sj + 12]));
            watch = Watch();
        } else if (watch.expired(dead)) {
            *dead = 1;
            P.scal->watchdog = 1;
            // release every stage so that the other warps can finish
            for (int b = sent >> 5; b < P.nbx; b++)
                if (lane == 0) mbar_arrive(&done[b % nst]);
            return;
        }
    }
}

// ---------------------------------------------------------------------- kernel ----
template <int KIND, bool DOT, bool MASKED>
__global__ void __launch_bounds__(192, 1) k_sweep(const __grid_constant__ SweepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[3 * MAX_STAGES]; // full[], done[], empty[]
    __shared__ int s_ticket;
    __shared__ int s_dead;
    __shared__ unsigned s_counters[2]; // [0] columns finished by the last row, [1] hand-off columns received
    double *smem = reinterpret_cast<double *>(smem_raw);
    double *halo_s = smem + P.nst * P.nt * TILE_DOUBLES; // [HR][32] hand-off ring
    uint64_t *full = bars, *done = bars + MAX_STAGES, *empty = bars + 2 * MAX_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = P.cs > 1 ? cluster_ctarank() : 0;

    // Strips are handed out in ticket order, one ticket per cluster (taken first, so the
    // global count stays consistent even for gated launches): a running CTA only ever
    // waits on strips that are already running or finished.
    if (threadIdx.x == 0) {
        if (rank == 0) s_ticket = (int)(atomicAdd(P.ticket, 1ULL) - P.ticket_base);
        s_dead = 0;
        s_counters[0] = 0;
        s_counters[1] = 0;
    }
    uint4 *ll_ring = reinterpret_cast<uint4 *>(halo_s + HR * 32); // [HR][32] messages from the cluster neighbour
    if (P.cs > 1)
        for (int i = threadIdx.x; i < HR * 32; i += blockDim.x) ll_ring[i] = make_uint4(0u, 0u, 0u, 0u); // tag 0 = empty
    __syncthreads();
    int ticket;
    if (P.cs > 1) {
        cluster_sync_all(); // every CTA of the cluster is resident, its counters are zero, rank 0's ticket is set
        ticket = (int)ld_remote_u32(mapa(smem_u32(&s_ticket), 0));
    } else {
        ticket = s_ticket;
    }
    const int sj = P.sj_base + ticket * P.cs + (int)rank;
    if (sj >= P.sj_base + P.nloc) return;    // padding CTA of the last cluster
    if (P.gated && P.scal->done) return;     // the solve has converged: nothing to do
    if (threadIdx.x == 0) {
        const bool publish = sj + 1 < P.nby;
        for (int i = 0; i < P.nst; i++) {
            mbar_init(&full[i], 1);                // loader's expect_tx arrival (+ TMA bytes)
            mbar_init(&done[i], publish ? 2 : 1);  // compute warp (+ publisher warp)
            mbar_init(&empty[i], 1);               // storer warp
        }
        fence_mbar_init();
    }
    // the very first strip has no upstream row: its hand-off rows read +0.0
    if (sj == 0)
        for (int i = threadIdx.x; i < HR * 32; i += blockDim.x) halo_s[i] = 0.0;
    __syncthreads();

    if (warp == 0) {
        unsigned long long t0 = 0;
        const long long c0 = clock64();
        if (P.times && lane == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        compute_warp<KIND, DOT>(P, smem, halo_s, full, done, sj, lane, &s_dead, s_counters);
        if (P.times && lane == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            P.times[16 * sj] = t0;
            P.times[16 * sj + 1] = t1;
            P.times[16 * sj + 15] = (unsigned long long)(clock64() - c0); // SM cycles spent by the compute warp
        }
    } else if (warp == 1) {
        loader_warp<KIND>(P, smem, full, empty, sj, lane, &s_dead);
    } else if (warp == 2) {
        storer_warp<KIND, DOT, MASKED>(P, smem, done, empty, sj, lane, &s_dead);
    } else if (warp == 3) {
        if (sj + 1 < P.nby) publisher_warp<KIND>(P, smem, halo_s, done, sj, lane, &s_dead, s_counters, rank);
    } else if (warp == 5 && sj > 0) {
        // (warp 4 stays idle: it would share the compute warp's scheduler, and a spinning
        // neighbour costs the recurrence ~12 cycles per step, profiles/microbench/step.cu)
        // first strip of a cluster: its upstream strip lives in another cluster (or on another
        // rank) and talks through L2; the others receive their messages in shared memory
        poller_warp<KIND>(P, halo_s, sj, lane, &s_dead, s_counters, (rank == 0 || sj == P.sj_base) ? nullptr : ll_ring);
    }
}

// ------------------------------------------------------------------- host side ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Tensor map of one pitched cell array: 2-D, double, box = box_w columns x box_h rows, no
// swizzle (the skewed access pattern is conflict-free on dense rows), zero OOB fill.
// Maps are cached per base pointer (flip() only swaps pointers).
struct MapCache {
    enum { N = 96 };
    void *key[N];
    int box_w[N], box_h[N];
    CUtensorMap map[N];
    int n;
};

int sweep_get_map(ifl_ctx *c, const Arr &a, int box_w, int box_h, CUtensorMap *out) {
    MapCache *mc = (MapCache *)c->map_cache;
    for (int i = 0; i < mc->n; i++)
        if (mc->key[i] == a.p && mc->box_w[i] == box_w && mc->box_h[i] == box_h) {
            *out = mc->map[i];
            return IFL_OK;
        }
    PFN_encodeTiled enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return IFL_E_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)a.pitch, (cuuint64_t)a.rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)a.pitch * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a.p, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %dx%d array", (int)r, a.w, a.h);
        return IFL_E_CUDA;
    }
    if (mc->n < MapCache::N) {
        mc->key[mc->n] = a.p;
        mc->box_w[mc->n] = box_w;
        mc->box_h[mc->n] = box_h;
        mc->map[mc->n] = m;
        mc->n++;
    }
    *out = m;
    return IFL_OK;
}

int sweep_init(ifl_ctx *c) {
    const int nbx = (c->W + 31) / 32, nby = (c->H + 31) / 32;
    c->n_strips = nby;
    {   // every rank owns a full [nby][nbx*32] hand-off array; a slab's last strip publishes into
        // the downstream rank's array (rank+1 for forward sweeps, rank-1 for backward ones)
        void *base = nullptr;
        size_t stride = 0;
        int rc = dist_alloc_per_rank(c, &base, (size_t)nby * nbx * 32 * sizeof(uint4), &stride);
        if (rc != IFL_OK) return rc;
        c->handoff_base = base;
        c->handoff = (unsigned long long *)((char *)base + stride * c->rank);
        c->handoff_down[0] = (unsigned long long *)((char *)base + stride * (c->rank + 1 < c->world ? c->rank + 1 : c->rank));
        c->handoff_down[1] = (unsigned long long *)((char *)base + stride * (c->rank > 0 ? c->rank - 1 : c->rank));
    }
    IFL_CUDA(cudaMalloc(&c->ticket, sizeof(unsigned long long)));
    IFL_CUDA(cudaMemset(c->ticket, 0, sizeof(unsigned long long)));
    IFL_CUDA(cudaMalloc(&c->sweep_times_buf, (size_t)nby * 16 * sizeof(unsigned long long)));
    // Strips are launched in thread-block clusters of 8: inside a cluster the hand-off messages
    // go straight into the downstream CTA's shared memory (flag-in-data, no fence), only every
    // 8th hand-off travels through L2.  Measured at 4096^2 (backward sweep): 645 us without
    // clusters, 583 / 563 / 555 us with clusters of 2 / 4 / 8.  (A first DSMEM version that
    // bumped a remote counter with a cluster-scope release store was SLOWER than L2, 825 us:
    // profiles/r01_c_dsmem_experiment.txt.)  IFL_SWEEP_CLUSTER=1|2|4|8 overrides.
    c->sweep_cluster = 8;
    if (const char *e = getenv("IFL_SWEEP_CLUSTER")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) c->sweep_cluster = v;
    }
    // tri_kernels.cu goes to the non-portable cluster size 16 when every strip of the sweep is resident at once
    // (4096^2: 4 clusters instead of 8, 440 instead of 453 us per sweep, 1.140 instead of 1.163 ms per PCG iteration;
    // profiles/r02_tri_experiments.txt section 10).  IFL_TRI_CLUSTER16=0 keeps 8.
    c->fuse_xpay = 1;
    if (const char *e = getenv("IFL_FUSE_XPAY")) c->fuse_xpay = atoi(e) != 0;
    c->matvec_uniform_allowed = 1;
    if (const char *e = getenv("IFL_MATVEC_UNIFORM")) c->matvec_uniform_allowed = atoi(e) != 0;
    c->tri_cluster16 = 1;
    if (const char *e = getenv("IFL_TRI_CLUSTER16")) c->tri_cluster16 = atoi(e) != 0;
    // Every strip with an upstream neighbour runs at the same pace (the hand-off checks make it ~3 %
    // slower than the head strip), so the lag a strip picks up while it starts -- cold instruction
    // cache, first TMA tiles, first hand-off -- is frozen for the whole sweep: a consumer that is not
    // faster than its producer never catches up (profiles/r01_g_sweep_timeline_4096.txt: 2.9 us per
    // strip, of which the hand-off itself explains ~2.1).  Letting the head strip idle a few cycles
    // per macro-step makes IT the pace-setter; all others then run into their hand-off waits and
    // settle at the minimal lag (skew + group + hand-off latency).  IFL_SWEEP_HEAD_DELAY overrides.
    c->sweep_head_delay = 0;
    if (const char *e = getenv("IFL_SWEEP_HEAD_DELAY")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 100000) c->sweep_head_delay = v;
    }
    // The triangular solves of the PCG loop run on the two-rows-per-lane engine (tri_kernels.cu);
    // IFL_TRI=0 keeps them on this file's one-row engine for A/B measurements.
    c->tri_engine = 1;
    if (const char *e = getenv("IFL_TRI")) c->tri_engine = atoi(e);
    // k_axpy2_norm overlapped with the forward sweep (pcg_kernels.cu, enqueue_iteration)
    c->overlap_axpy = 3;
    if (const char *e = getenv("IFL_OVERLAP_AXPY")) c->overlap_axpy = atoi(e);
    {
        cudaDeviceProp prop;
        IFL_CUDA(cudaGetDeviceProperties(&prop, c->device));
        c->sm_count = prop.multiProcessorCount;
    }
    IFL_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    IFL_CUDA(cudaEventCreateWithFlags(&c->ev_alpha, cudaEventDisableTiming));
    IFL_CUDA(cudaEventCreateWithFlags(&c->ev_axpy, cudaEventDisableTiming));
    {
        const size_t nb = (size_t)((c->H + 63) / 64 + 1) * sizeof(unsigned);
        IFL_CUDA(cudaMalloc(&c->band_count, nb));
        IFL_CUDA(cudaMemset(c->band_count, 0, nb));
        c->band_epoch = 0;
    }
    c->map_cache = calloc(1, sizeof(MapCache));
    if (!c->map_cache) return IFL_E_NOMEM;
    c->epoch = 0;
    return IFL_OK;
}

void sweep_free(ifl_ctx *c) {
    if (c->handoff_base) dist_free_mem(c, c->handoff_base);
    c->handoff_base = nullptr;
    if (c->ticket) cudaFree(c->ticket);
    if (c->band_count) cudaFree(c->band_count);
    if (c->ev_alpha) cudaEventDestroy(c->ev_alpha);
    if (c->ev_axpy) cudaEventDestroy(c->ev_axpy);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    c->band_count = nullptr;
    if (c->sweep_times_buf) cudaFree(c->sweep_times_buf);
    free(c->map_cache);
    c->handoff = nullptr;
    c->ticket = nullptr;
    c->map_cache = nullptr;
}

struct TileSpec {
    const Arr *a;
    int load, row_shift, store;
    const Arr *a2;       // optional second store target (non-zero values only)
    const Arr *store_to; // store target when it is not the array the tile was loaded from
};

template <int KIND, bool DOT>
static int launch_sweep(ifl_ctx *c, SweepParams &P, const TileSpec *spec, int nt, int nst) {
    P.nt = nt;
    P.nst = nst;
    for (int k = 0; k < nt; k++) {
        P.t[k].p = spec[k].store_to ? spec[k].store_to->p : spec[k].a->p;
        P.t[k].load = spec[k].load;
        P.t[k].row_shift = spec[k].row_shift;
        P.t[k].store = spec[k].store;
        P.t[k].p2 = spec[k].a2 ? spec[k].a2->p : nullptr;
        if (spec[k].load) {
            int rc = sweep_get_map(c, *spec[k].a, 32, TROWS, &P.map[k]);
            if (rc != IFL_OK) return rc;
        }
    }
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
    P.cs = c->sweep_cluster;
    {   // this rank's strips, in sweep order (backward sweeps start at the bottom strip)
        const int s0 = c->ry0 / 32, s1 = (c->ry1 + 31) / 32;
        P.nloc = s1 - s0;
        P.sj_base = (KIND == KIND_BWD) ? P.nby - s1 : s0;
        P.handoff_down = reinterpret_cast<uint4 *>(c->handoff_down[KIND == KIND_BWD ? 1 : 0]);
    }
    const int n_clusters = (P.nloc + P.cs - 1) / P.cs;
    P.ticket = c->ticket;
    P.ticket_base = c->sweep_tickets;
    c->sweep_tickets += (unsigned long long)n_clusters;
    c->sweep_launches++;
    P.scal = c->scal;
    P.head_delay = c->sweep_head_delay;
    P.times = c->sweep_times;
    const size_t smem = (size_t)nst * nt * TILE_BYTES + (size_t)HR * 32 * sizeof(double) + (P.cs > 1 ? (size_t)HR * 32 * sizeof(uint4) : 0);
    const bool masked = P.mask_tile >= 0;
    static bool attr_set[IFL_MAX_DEVICES][5][2][2]; // function attributes are per device
    if (!attr_set[c->device % IFL_MAX_DEVICES][KIND][DOT][masked]) {
        IFL_CUDA(masked ? cudaFuncSetAttribute(k_sweep<KIND, DOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024 - 1024)
                        : cudaFuncSetAttribute(k_sweep<KIND, DOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024 - 1024));
        IFL_CUDA(masked ? cudaFuncSetAttribute(k_sweep<KIND, DOT, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
                        : cudaFuncSetAttribute(k_sweep<KIND, DOT, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set[c->device % IFL_MAX_DEVICES][KIND][DOT][masked] = true;
    }
    ProfScope ps_(c, KIND == KIND_FWD ? IFL_K_PRECON_FWD : KIND == KIND_BWD ? IFL_K_PRECON_BWD : (KIND == KIND_FACTOR || KIND == KIND_FACTOR_M) ? IFL_K_FACTOR : IFL_K_GS_SWEEP);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(n_clusters * P.cs));
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = P.cs > 1 ? 1 : 0;
    const cudaError_t le = masked ? cudaLaunchKernelEx(&cfg, k_sweep<KIND, DOT, true>, P)
                                  : cudaLaunchKernelEx(&cfg, k_sweep<KIND, DOT, false>, P);
    if (le != cudaSuccess) {
        set_error("sweep launch (cluster %d) -> %s", P.cs, cudaGetErrorString(le));
        return IFL_E_CUDA;
    }
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_mic0_factor(ifl_ctx *c) {
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.mask_tile = -1;
    if (c->version >= 4) {
        // precon keeps its old value at non-fluid cells; pe is the +0.0-masked copy the solves use
        const TileSpec spec[7] = {{&c->aDiag, 1, 0, 0, nullptr}, {&c->aPlusX, 1, 0, 0, nullptr}, {&c->aPlusY, 1, 0, 0, nullptr},
                                  {&c->pe, 0, 0, 1, &c->precon}, {&c->cx, 0, 0, 1, nullptr},    {&c->cy, 0, 0, 1, nullptr},
                                  {&c->fmask, 1, 0, 0, nullptr}};
        return launch_sweep<KIND_FACTOR_M, false>(c, P, spec, 7, 3);
    }
    const TileSpec spec[6] = {{&c->aDiag, 1, 0, 0, nullptr}, {&c->aPlusX, 1, 0, 0, nullptr}, {&c->aPlusY, 1, 0, 0, nullptr},
                              {&c->precon, 0, 0, 1, nullptr}, {&c->cx, 0, 0, 1, nullptr},    {&c->cy, 0, 0, 1, nullptr}};
    return launch_sweep<KIND_FACTOR, false>(c, P, spec, 6, 4);
}

// chapters 4+ multiply by `pe` (precon with +0.0 at non-fluid cells) and store only fluid cells
static const Arr &precon_operand(ifl_ctx *c) { return c->version >= 4 ? c->pe : c->precon; }

// Ring depth of the two triangular solves: 5 stages, one CTA per SM.  Measured alternative
// (IFL_SWEEP_STAGES=3: 101 KB, two strips per SM for grids with more strips than SMs): slower,
// 1.72 vs 1.33 ms per forward sweep at 8192^2 -- a 3-stage ring leaves the TMA loads one
// macro-step (1.6 us) to land and the co-resident CTA does not hide the resulting stalls.
static int solve_stages(ifl_ctx *c) {
    (void)c;
    if (const char *e = getenv("IFL_SWEEP_STAGES")) {
        const int v = atoi(e);
        if (v >= 3 && v <= 5) return v;
    }
    return 5;
}

int launch_precon_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated) {
    if (c->tri_engine == 2) return launch_stair_forward(c, dst, a, gated);
    if (c->tri_engine) return launch_tri_forward(c, dst, a, gated);
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.gated = gated ? 1 : 0;
    P.mask_tile = c->version >= 4 ? 3 : -1;
    // the rhs tile is updated in place and drained into dst
    const TileSpec spec[4] = {{&a, 1, 0, 1, nullptr, &dst},
                              {&c->cx, 1, 0, 0, nullptr, nullptr},
                              {&c->cy, 1, 0, 0, nullptr, nullptr},
                              {&precon_operand(c), 1, 0, 0, nullptr, nullptr}};
    return launch_sweep<KIND_FWD, false>(c, P, spec, 4, solve_stages(c));
}

int launch_precon_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    if (c->tri_engine == 2) return launch_stair_backward(c, dst, r_for_dot, with_dot, gated);
    if (c->tri_engine) return launch_tri_backward(c, dst, r_for_dot, with_dot, gated);
    SweepParams P;
    memset(&P, 0, sizeof P);
    P.gated = gated ? 1 : 0;
    P.mask_tile = c->version >= 4 ? 3 : -1;
    const TileSpec spec[5] = {{&dst, 1, 0, 1, nullptr, nullptr},
                              {&c->cx, 1, 0, 0, nullptr, nullptr},
                              {&c->cy, 1, 0, 0, nullptr, nullptr},
                              {&precon_operand(c), 1, 0, 0, nullptr, nullptr},
                              {&r_for_dot, 1, 0, 0, nullptr, nullptr}};
    if (with_dot) {
        P.partials = partials_next(c);
        c->n_partials = (c->H + 31) / 32;
        return launch_sweep<KIND_BWD, true>(c, P, spec, 5, solve_stages(c));
    }
    return launch_sweep<KIND_BWD, false>(c, P, spec, 4, solve_stages(c));
}

// ------------------------------------------------------- Gauss-Seidel projection ----
// project(limit, timestep) of chapters 1-2 (v2:233-277): up to `limit` lexicographic
// sweeps over the warm-started _p, stopping when max |p - newP| < 1e-5.
__global__ void __launch_bounds__(1024) k_scalar_gs(const double *partials, int n, SolveScalars *sc, DistDev dd) {
    if (sc->done) return;
    dist_barrier_block(dd); // every rank's strips of this sweep are done; all ranks fold the same partials
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) v = std_max(v, __ldcg(partials + i));
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
    P.gated = 1;
    P.scale = scale;
    P.mask_tile = -1;
    // tile 1 is p again, fetched one row further down: the old values of the row below (v2:264)
    const TileSpec spec[3] = {{&c->p, 1, 0, 1, nullptr}, {&c->p, 1, 1, 0, nullptr}, {&c->r, 1, 0, 0, nullptr}};
    P.partials = partials_next(c);
    c->n_partials = (c->H + 31) / 32;
    int rc = launch_sweep<KIND_GS, false>(c, P, spec, 3, 5);
    if (rc != IFL_OK) return rc;
    ProfScope ps_(c, IFL_K_SCALAR);
    k_scalar_gs<<<1, 1024, 0, c->stream>>>(c->partials, c->n_partials, c->scal, c->ddev);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int gs_project(ifl_ctx *c, int limit, double timestep, double density, ifl_solve_info *info) {
    cudaStream_t st = c->stream;
    const double scale = timestep / (density * c->hx * c->hx); // v2:234
    IFL_CUDA(cudaMemsetAsync(c->scal, 0, offsetof(SolveScalars, watchdog), st)); // the watchdog word is sticky
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
            set_error("gs_project: a dependency wait (wavefront hand-off or rank barrier) timed out");
            cudaMemsetAsync(&c->scal->watchdog, 0, sizeof(int), st);
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
