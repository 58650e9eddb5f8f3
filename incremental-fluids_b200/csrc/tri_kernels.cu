// tri_kernels.cu -- the two triangular solves of FluidSolver::applyPreconditioner
// (v3:275-304; masked form v5:746-780), round-2 engine: TWO rows per lane.
//
//   forward   t = a[i] - cx[i-1]*z[i-1] - cy[i-w]*z[i-w] ; z[i] = t*precon[i]     v3:276-287
//   backward  t = z[i] - cx[i]*z[i+1]   - cy[i]*z[i+w]   ; z[i] = t*precon[i]     v3:289-303
//             (+ dotProduct(z, r) of v3:374 folded in by the storer warp)
//   cx = aPlusX*precon, cy = aPlusY*precon are written by the factorisation (sweep_kernels.cu);
//   the reference evaluates `_aPlusX[i]*_precon[i]*dst[i]` left to right (v3:281), so this is the
//   same product.  Any schedule that honours the (x-1,y),(x,y-1) dependencies performs the same
//   floating-point operations on the same operands as the raster loop: results are bit-identical.
//
// Why a second engine (the one-row engine of sweep_kernels.cu still runs the factorisation and
// Gauss-Seidel): a sweep is a wave that has to cross W + H cells; its duration is
//     (W + strips * (skew + hand-off)) * T_step.
// The one-row engine (32-row strips) measured T = 87 cycles and 66 steps of lag per strip at 4096^2.
// Here a lane owns two vertically adjacent rows: cell A (upper row) takes its upper neighbour from
// the lane above by shuffle, as before, and cell B (lower row) takes A straight from a register.
// A step is longer (one more mul-sub-mul chain) but a strip is 64 rows tall: half the strips, half
// the hand-offs, half the skew steps, and the per-step overheads (hand-off polling, ring barriers,
// address selects) are paid once for two rows.  Measured at 4096^2: T = 117 cycles, 48 steps (2.85 us)
// of lag per strip, 440 us per solve.
//
// Geometry
//   * strip = 64 rows, one CTA (1 per SM); lane t owns tile rows 1+2t, 2+2t (forward; tile row 0 is
//     the upstream strip's last row) or 63-2t, 62-2t (backward; tile row 64 is the upstream row) and
//     runs one column behind lane t-1: the warp is an anti-diagonal, 31 columns long.
//   * operands are staged by TMA in blocks of 16 columns x 65 rows (one box per operand per block:
//     rhs/z, cx, cy, precon), 6-stage ring.  Lanes 0..15 work in blocks m-1 and m, lanes 16..31 in
//     m-2 and m-1; a block is released (lane 0 arrives on done[] from inside the step stream, five steps
//     after lane 31 has left it -- no warp-wide fence).  Rows are dense (128 bytes), lane t
//     reads column (k - t) mod 16: the 16 lanes of a half-warp hit 16 different 8-byte bank slots.
//   * clusters of 16 CTAs (non-portable size) when every strip is resident at once, else of 8: inside a
//     cluster the hand-off goes through distributed shared memory, between clusters through L2.
//   * warps: 0 compute, 1 TMA loader, 2 storer (drains z, folds z.r reading r straight from global
//     memory -- it does not depend on the sweep, so its loads are issued before the wait), 3 publisher,
//     5 gatekeeper (strip-to-strip hand-off + TMA arrival -> one "gate" counter for the compute warp).
//   * NOBODY POLLS shared memory while the compute warp runs.  Measured (profiles/r02_tri_experiments.txt):
//     with a publisher spinning on the progress counter and a poller spinning on the message ring the
//     compute warp needs 129 cycles per step, without any helper warp 98 -- their shared-memory polls
//     queue up in front of the recurrence's own LDS / SHFL.  So every helper sleeps on an mbarrier:
//       - the compute warp rings a "bell" (mbarrier.arrive, ring of 16) every 8 columns its last row
//         completes; the publisher sleeps on it, then forwards those 8 values;
//       - inside a thread-block cluster the values travel as st.async (8 bytes + complete_tx) straight
//         into the downstream CTA's hand-off ring and count on one of ITS 64 group mbarriers, on which
//         its gatekeeper sleeps; between clusters and between GPUs they travel as NCCL-LL style
//         {lo, epoch, hi, epoch} messages through L2 / NVLink, which the gatekeeper polls in GLOBAL
//         memory (with a back-off);
//       - the gatekeeper also waits for the TMA ring (`full` barriers), so the compute warp never
//         executes an mbarrier wait: it reads one shared counter every 8 steps, four steps ahead of use.
#include "ifl_internal.cuh"
#include "sweep_common.cuh"

#include <cuda.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

// Timing experiments (profiles/tri_experiments.sh builds one library per value; never set in the
// product build; results are garbage, only strip 0's step time is of interest): bit 1 (2) no helper
// warps, gate preset to "everything is there"; bit 2 (4) no progress store / bell; bit 3 (8) no lane-0
// hand-off select; bit 4 (16) no gate checks; bit 5 (32) no done[] arrival; bit 7 (128) the storer does
// not drain (waits and releases only); bit 8 (256) the loader arrives on `full` without loading;
// bit 9 (512) no hand-off at all: no publisher, every strip runs as if it were the first.
#ifndef TRI_EXP
#define TRI_EXP 0
#endif

namespace ifl {
namespace tri {

constexpr int SR = 64;                        // rows per strip
constexpr int BW = 16;                        // columns per block
constexpr int TROWS = SR + 1;                 // tile rows (one upstream row)
constexpr int ROWB = BW * 8;                  // bytes per tile row
constexpr int TILE_BYTES = TROWS * ROWB;      // 8320 (multiple of 128)
constexpr int NT = 4;                         // tiles per stage: rhs/z, cx, cy, precon
constexpr int STAGE_BYTES = NT * TILE_BYTES;  // 33280
constexpr int NST = 6;                        // ring depth
constexpr int HG = 8;                         // hand-off granularity (columns)
constexpr int HRC = 512;                      // hand-off ring (columns): deep enough that back-pressure never binds
constexpr int NHB = HRC / HG;                 // hand-off group barriers (one per ring group)
constexpr int NBELL = 16;                     // bell ring (the publisher is never NST blocks = 12 groups behind)

struct TriParams {
    CUtensorMap map[NT]; // must stay first (64-byte aligned)
    double *dst;         // z: receives the swept tile
    const double *rdot;  // backward + dot: r
    int W, H, pitch, nbx, nby; // nbx blocks of 16 columns, nby strips of 64 rows
    uint4 *handoff;      // [nby][ncols] LL messages between strips of different clusters
    int sj_base, nloc;   // this rank's strips, in sweep order
    uint4 *handoff_down; // hand-off array of the downstream rank
    unsigned epoch;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    SolveScalars *scal;
    int gated;
    double *partials; // z.r per strip
    int cs;           // cluster size
    int head_delay;
    const unsigned *band_count; // forward sweep overlapped with k_axpy2_norm: finished blocks per 64-row band (or null)
    unsigned band_target;       // ... a strip may read its band of the rhs once the counter has reached this
    unsigned long long *times;
};

template <bool BWD>
struct Geo {
    static constexpr int DIR = BWD ? -8 : 8;       // bytes per logical column
    static constexpr int COL0 = BWD ? BW - 1 : 0;  // tile column of logical in-block column 0
    static constexpr int ROW_B = BWD ? -ROWB : ROWB; // cell A -> cell B (same column)
    static constexpr int UP_CY = BWD ? 0 : -ROWB;  // own cell -> the cell whose cy multiplies the upstream value
    __device__ static __forceinline__ int row_a(int lane) { return BWD ? SR - 1 - 2 * lane : 1 + 2 * lane; }
    __device__ static __forceinline__ int last_row() { return BWD ? 0 : SR; } // tile row of the strip's last row
    __device__ static __forceinline__ int tcol(int ci) { return BWD ? BW - 1 - ci : ci; }
};

struct Carry {
    double zA, zB; // swept variable of the previous column
    double cA, cB; // forward: cx of the previous column
};
struct Ops {
    double aA, xA, yA, pA, aB, xB, yB, pB, halo;
};

template <bool BWD>
__device__ __forceinline__ void fetch(Ops &o, uint32_t p, uint32_t ph) {
    typedef Geo<BWD> G;
    const uint32_t pb = p + (uint32_t)G::ROW_B;
    o.aA = lds_f64(p);
    o.xA = lds_f64(p + TILE_BYTES);
    o.yA = lds_f64(p + 2 * TILE_BYTES + (uint32_t)G::UP_CY);
    o.pA = lds_f64(p + 3 * TILE_BYTES);
    o.aB = lds_f64(pb);
    o.xB = lds_f64(pb + TILE_BYTES);
    o.yB = lds_f64(pb + 2 * TILE_BYTES + (uint32_t)G::UP_CY);
    o.pB = lds_f64(pb + 3 * TILE_BYTES);
    o.halo = lds_f64(ph); // same address in every lane (broadcast); only lane 0 uses it
}

// Per-lane tile-0 base pointers of one macro-step (16 steps).  With r = lane & 15 the lane moves
// from its "before" block into its "after" block at step kk == r; the address of logical step j is
// always `selected base + DIR*j`, j a compile-time constant.  N serves the look-ahead fetch (j == 16)
// of the lanes with r == 0.
struct LaneBases {
    uint32_t A, B, N;
};
template <bool BWD>
__device__ __forceinline__ uint32_t pos(const LaneBases &lb, int j, int r) {
    uint32_t base = (r > j) ? lb.B : lb.A;
    if (j >= BW) base = (r == 0) ? lb.N : lb.A;
    return base + (uint32_t)(Geo<BWD>::DIR * j);
}

// EDGE 0: every lane is inside the strip.  EDGE 1: first two macro-steps (lanes enter one by one),
// EDGE 2: last two (lanes leave).  Lanes outside run the same instructions on aliased operands with
// their stores predicated off; what they compute is never consumed (lane t-1 is inside at step k-1
// exactly when lane t is inside at step k).
template <bool BWD, int EDGE>
__device__ __forceinline__ void macro_step(const LaneBases &lb, uint32_t h_cur, uint32_t h_next, uint32_t bell6, uint32_t bell14,
                                           uint32_t done_addr, int m, int lane, int r, Carry &cr, Ops &ops, uint32_t progress_addr,
                                           uint32_t gate_addr, int ncols, volatile int *dead, SolveScalars *scal) {
    typedef Geo<BWD> G;
    uint32_t p = pos<BWD>(lb, 0, r);
    unsigned gate_seen = 0;
#pragma unroll
    for (int kk = 0; kk < BW; kk++) {
        // gate (hand-off values received AND operand blocks loaded, in columns): read four steps early,
        // tested when lane 0 is about to fetch the first column of the next group
        if (((kk + 5) % HG) == 0 && EDGE != 2 && !(TRI_EXP & 16)) gate_seen = lds_u32_volatile(gate_addr);
        // lane 31 left block m-3 at the end of macro-step m-1: hand it to the storer (and, through it, the loader; the
        // storer issues the proxy fence before the TMA may overwrite the stage).  The last store into it was issued
        // five steps (~600 cycles) ago by this same warp, so lane 0's arrive needs no warp-wide fence: a __syncwarp()
        // + arrive at the macro-step boundary drains every load in flight (measured on the staircase engine: 5.6 cycles
        // per step, profiles/r02_tri_experiments.txt section 7).
        if (kk == 4 && !(TRI_EXP & 32)) {
            if (lane == 0 && done_addr != 0) mbar_arrive_addr(done_addr);
        }
        if (((kk + 1) % HG) == 0 && EDGE != 2 && !(TRI_EXP & 16)) {
            const unsigned need = (unsigned)imin(BW * m + kk + 1 + HG, ncols);
            if (gate_seen < need) wait_counter<false>(gate_addr, need, dead, scal);
        }
        // ---- critical path first: the upstream lane's B of the previous step (lane 0: the hand-off value)
        double up = __shfl_up_sync(0xffffffffu, cr.zB, 1);
        // ---- operands of step kk+1, in the shadow of the shuffle
        Ops nxt;
        const uint32_t pn = pos<BWD>(lb, kk + 1, r);
        const uint32_t ph = (kk + 1 < BW) ? h_cur + (uint32_t)(8 * (kk + 1)) : h_next;
        fetch<BWD>(nxt, pn, ph);
        // ---- this step
        const int c = BW * m + kk - lane; // logical column of this lane
        const bool active = (EDGE == 0) ? true : (EDGE == 1 ? c >= 0 : c < ncols);
        if (EDGE == 1) {
            const bool first = c == 0;
            cr.zA = sel_f64(first, 0.0, cr.zA);
            cr.zB = sel_f64(first, 0.0, cr.zB);
            if (!BWD) {
                cr.cA = sel_f64(first, 0.0, cr.cA);
                cr.cB = sel_f64(first, 0.0, cr.cB);
            }
        }
        // (lane 0 takes the hand-off value.  Tried: a rotating shuffle with lane 31 carrying the hand-off value
        // in the shuffled register -- ptxas turns the predicated duplicate multiply into four FSELs on the
        // chain, slower than this one select pair.)
        if (!(TRI_EXP & 8)) up = sel_f64(lane == 0, ops.halo, up);
        double zA, zB;
        if (!BWD) {
            double t = ops.aA - cr.cA * cr.zA; // v3:281  t -= aPlusX[idx-1]*precon[idx-1]*dst[idx-1]
            t = t - ops.yA * up;               // v3:283  t -= aPlusY[idx-w]*precon[idx-w]*dst[idx-w]
            zA = t * ops.pA;                   // v3:285
            double u = ops.aB - cr.cB * cr.zB;
            u = u - ops.yB * zA; // the row above B is A
            zB = u * ops.pB;
            cr.cA = ops.xA;
            cr.cB = ops.xB;
        } else {
            double t = ops.aA - ops.xA * cr.zA; // v3:297  t -= aPlusX[idx]*precon[idx]*dst[idx+1]
            t = t - ops.yA * up;                // v3:299  t -= aPlusY[idx]*precon[idx]*dst[idx+w]
            zA = t * ops.pA;                    // v3:301
            double u = ops.aB - ops.xB * cr.zB;
            u = u - ops.yB * zA; // the row below B is A
            zB = u * ops.pB;
        }
        sts_f64_p<EDGE == 0>(p, zA, active); // in place: the rhs / z tile becomes the result tile
        sts_f64_p<EDGE == 0>(p + (uint32_t)G::ROW_B, zB, active);
        cr.zA = zA;
        cr.zB = zB;
        // the strip's last row (lane 31, cell B) has just completed another group of HG columns: publish
        // the count and ring the publisher's bell (lane 31 wrote those values itself: its arrive releases them)
        if (((kk + 2) % HG) == 0 && EDGE != 1 && !(TRI_EXP & 4)) {
            sts_u32_volatile(progress_addr, (unsigned)imin(BW * m + kk - 30, ncols));
            if (lane == 31) mbar_arrive_addr(kk < HG ? bell6 : bell14);
        }
        ops = nxt;
        p = pn;
    }
}

template <bool BWD>
__device__ void compute_warp(const TriParams &P, unsigned char *smem, double *halo_s, uint64_t *done, uint64_t *bell, int sj,
                             int lane, volatile int *dead, unsigned *counters) {
    typedef Geo<BWD> G;
    Carry cr;
    cr.zA = cr.zB = cr.cA = cr.cB = 0.0;
    const bool has_up = sj > 0;
    (void)has_up;
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    const int nbx = P.nbx, ncols = P.nbx * BW;
    const int q = lane >> 4, r = lane & 15;
    // this lane's row A in tile 0 of stage 0 at logical in-block column 0
    const uint32_t row0 = smem_u32(smem) + (uint32_t)(G::row_a(lane) * ROWB + G::COL0 * 8);
    const uint32_t halo0 = smem_u32(halo_s);
    // blocks of a lane in macro-step m: after its transition (bA = m - q), before (bA - 1), look-ahead of the
    // r == 0 lanes (bA + 1); blocks outside [0, nbx) are only touched by lanes outside the strip and alias
    // valid memory
    auto bases_of = [&](int sA, int sB, int sN) {
        LaneBases lb;
        lb.A = row0 + (uint32_t)(sA * STAGE_BYTES) - (uint32_t)(G::DIR * r);
        lb.B = row0 + (uint32_t)(sB * STAGE_BYTES) + (uint32_t)(G::DIR * (BW - r));
        lb.N = row0 + (uint32_t)(sN * STAGE_BYTES) - (uint32_t)(G::DIR * BW);
        return lb;
    };
    auto bases = [&](int m) { // general form (first / last macro-steps)
        const int bA = m - q;
        return bases_of(imin(imax(bA, 0), nbx - 1) % NST, imin(imax(bA - 1, 0), nbx - 1) % NST, imin(imax(bA + 1, 0), nbx - 1) % NST);
    };
    Ops ops;
    // everything that does not depend on the upstream strip happens BEFORE the wait for its first
    // hand-off group: that wait sits on the critical path of the whole sweep
    wait_counter<false>(gate_addr, (unsigned)imin(HG, ncols), dead, P.scal); // block 0 loaded, first hand-off group here
    fetch<BWD>(ops, pos<BWD>(bases(0), 0, r), halo0); // step 0: lane 0 at column 0, the others idle on valid memory
    const int nm = nbx + 2; // macro-steps: lane 31 finishes column ncols-1 at step ncols + 30
    int sm = 0;             // m % NST, kept incrementally (the per-macro-step bookkeeping is paid every 16 steps)
    auto run = [&](int m, auto edge_tag) {
        constexpr int EDGE = decltype(edge_tag)::value;
        if (!has_up && P.head_delay > 0) { // pace-setter, see sweep_init
            const long long t_ = clock64();
            while (clock64() - t_ < P.head_delay) {}
        }
        LaneBases lb;
        if (m >= 2 && m < nbx - 1) { // no block index leaves [0, nbx): stages by rotation
            int a = sm - q;
            a += a < 0 ? NST : 0;
            const int b = a == 0 ? NST - 1 : a - 1, n = a == NST - 1 ? 0 : a + 1;
            lb = bases_of(a, b, n);
        } else {
            lb = bases(m);
        }
        const uint32_t h_cur = halo0 + (uint32_t)(((BW * m) & (HRC - 1)) * 8);
        const uint32_t h_next = halo0 + (uint32_t)(((BW * (m + 1)) & (HRC - 1)) * 8);
        // groups completed by the last row at kk == 6 / 14 of this macro-step: 2m-4, 2m-3 (m >= 2)
        const uint32_t bell6 = smem_u32(&bell[(2 * m + NBELL - 4) & (NBELL - 1)]), bell14 = smem_u32(&bell[(2 * m + NBELL - 3) & (NBELL - 1)]);
        const uint32_t done_addr = m >= 3 ? smem_u32(&done[sm >= 3 ? sm - 3 : sm - 3 + NST]) : 0u; // block m-3, released at step 4
        macro_step<BWD, EDGE>(lb, h_cur, h_next, bell6, bell14, done_addr, m, lane, r, cr, ops, progress_addr, gate_addr, ncols, dead, P.scal);
        sm = sm == NST - 1 ? 0 : sm + 1;
    };
    // ONE loop with a three-way branch.  (Three loops -- entering / steady / leaving -- make the head strip 4 % faster,
    // 111.6 instead of 116.8 cycles per step, and the sweep 5 % SLOWER, 475 instead of 453 us: the downstream strips
    // then run into an empty gate more often, and every such stall costs more than the step it saved;
    // profiles/r02_tri_experiments.txt section 9, A/B on one box.)
    for (int m = 0; m < nm; m++) {
        if (m < 2)
            run(m, std::integral_constant<int, 1>());
        else if (m >= nbx)
            run(m, std::integral_constant<int, 2>());
        else
            run(m, std::integral_constant<int, 0>());
    }
    if (!(TRI_EXP & 32)) { // the last block (nbx-1 = nm-3)
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[(nbx - 1) % NST]);
    }
}

// ----------------------------------------------------------------- loader warp ----
template <bool BWD>
__device__ void loader_warp(const TriParams &P, unsigned char *smem, uint64_t *full, uint64_t *empty, int sj, int lane,
                            volatile int *dead) {
    if (lane != 0) return;
    const int nbx = P.nbx;
    const int ty = BWD ? (P.nby - 1 - sj) : sj; // memory strip of this CTA
    const int box_y = BWD ? ty * SR : ty * SR - 1;
    if (!BWD && P.band_count) {
        // the rhs rows of this strip are being written by k_axpy2_norm on another stream: wait until every block of
        // the band has signalled (release: __threadfence + atomicAdd; acquire here), then order the TMA reads after it
        Watch watch;
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.band_count + ty) : "memory");
            if (v >= P.band_target || P.scal->done) break;
            if (watch.expired(dead)) {
                *dead = 1;
                P.scal->watchdog = 1;
                break;
            }
            __nanosleep(200);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    for (int b = 0; b < nbx; b++) {
        const int st = b % NST;
        if (b >= NST) mbar_wait(&empty[st], (unsigned)(((b / NST) - 1) & 1), dead, P.scal);
        if (TRI_EXP & 256) {
            mbar_arrive(&full[st]);
            continue;
        }
        mbar_arrive_expect_tx(&full[st], (unsigned)STAGE_BYTES);
        const int box_x = (BWD ? (nbx - 1 - b) : b) * BW;
        unsigned char *stage = smem + (size_t)st * STAGE_BYTES;
        for (int k = 0; k < NT; k++) tma_load_2d(stage + k * TILE_BYTES, &P.map[k], box_x, box_y, &full[st]);
    }
}

// ------------------------------------------------------------- gatekeeper warp ----
// Releases the compute warp group by group (HG columns) through counters[1]: a group is released when
// the upstream strip's last-row values for it are in halo_s (indexed by logical column mod HRC) AND the
// operand block it lies in has landed.  `hb` != null: the upstream strip runs in the same cluster and
// sends st.async + complete_tx on hb[group % NHB] (armed here with expect_tx); otherwise the values
// arrive as LL messages in global memory (L2, or NVLink for the first strip of a rank).
template <bool BWD>
__device__ void gatekeeper_warp(const TriParams &P, double *halo_s, uint64_t *full, uint64_t *hb, int sj, int lane,
                                volatile int *dead, unsigned *counters) {
    const int ncols = P.nbx * BW;
    const bool has_up = sj > 0 && !(TRI_EXP & 512);
    const uint4 *up_row = P.handoff + (size_t)(has_up ? sj - 1 : 0) * ncols;
    const bool remote = sj == P.sj_base; // the upstream strip belongs to another rank
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    Watch watch;
    mbar_wait(&full[0], 0, dead, P.scal); // operand block 0
    if (has_up && !hb) {
        // LL messages in global memory.  Lane l owns columns l, l + 32, l + 64, ... and walks them on its own,
        // two polls in flight, so the L2 round trip (~0.4 us; a group of 8 columns arrives every ~0.5 us) is
        // pipelined across columns instead of being paid once per batch.  (Measured with a batch of 32 columns
        // polled to completion before the next: the first strip behind every L2 hand-off ran at 74 instead of
        // 62 ns per step, 50 us per sweep, profiles/r02_tri_experiments.txt.)  The gate is the contiguous
        // prefix of received columns, rounded down to whole groups.
        int next_c = lane;         // this lane's first column not yet received
        unsigned released = 0;     // columns released to the compute warp
        int loaded_blocks = 1;     // operand blocks known to have landed
        while (released < (unsigned)ncols) {
            // ring slots are reused every HRC columns: stay behind the strip's own last row
            const int limit = (int)lds_u32_volatile(progress_addr) + HRC;
            double v0 = 0.0, v1 = 0.0;
            bool ok0 = false, ok1 = false;
            const int c0 = next_c, c1 = next_c + 32;
            if (c0 < ncols && c0 < limit) ok0 = remote ? ll_load_sys(up_row + c0, P.epoch, v0) : ll_load(up_row + c0, P.epoch, v0);
            if (c1 < ncols && c1 < limit) ok1 = remote ? ll_load_sys(up_row + c1, P.epoch, v1) : ll_load(up_row + c1, P.epoch, v1);
            if (ok0) {
                halo_s[c0 % HRC] = v0;
                next_c = c1;
                if (ok1) {
                    halo_s[c1 % HRC] = v1;
                    next_c = c1 + 32;
                }
            }
            const unsigned prefix = __reduce_min_sync(0xffffffffu, (unsigned)imin(next_c, ncols));
            const unsigned groups = prefix / HG * HG;
            if (groups > released) {
                __threadfence_block(); // halo_s values before the counter
                __syncwarp();
                // The compute warp may touch every column below the gate, so the gate never passes the operand
                // blocks that have landed; it is published block by block -- a block further ahead only loads once
                // the compute warp, fed by the gate published so far, has released the stage it needs.
                while (released < groups) {
                    const unsigned upto = umin(groups, (unsigned)(loaded_blocks * BW));
                    if (upto > released) {
                        if (lane == 0) sts_u32_volatile(gate_addr, upto);
                        released = upto;
                    }
                    if (released < groups) {
                        mbar_wait(&full[loaded_blocks % NST], (unsigned)((loaded_blocks / NST) & 1), dead, P.scal);
                        loaded_blocks++;
                        if (*dead) break;
                    }
                }
                watch = Watch();
            } else if (watch.expired(dead)) {
                *dead = 1;
                P.scal->watchdog = 1;
                break;
            }
        }
        return;
    }
    for (int c0 = 0; c0 < ncols; c0 += HG) {
        const int g = c0 / HG;
        // a ring slot may be rewritten once the strip's own last row has passed the column it held
        if (has_up && c0 + HG > HRC) wait_counter(progress_addr, (unsigned)(c0 + HG - HRC), dead, P.scal);
        if (has_up) mbar_wait(&hb[g % NHB], (unsigned)((g / NHB) & 1), dead, P.scal);
        if (lane == 0) sts_u32_volatile(gate_addr, (unsigned)(c0 + HG));
        // ---- off the critical path: arm the group barrier for its next use (NHB groups from now) and make sure
        // the operand block of the NEXT group has landed (the TMA ring runs far ahead: this returns at once)
        if (has_up && lane == 0) mbar_arrive_expect_tx(&hb[g % NHB], HG * 8);
        if ((c0 + HG) % BW == 0 && c0 + HG < ncols) {
            const int b = (c0 + HG) / BW;
            mbar_wait(&full[b % NST], (unsigned)((b / NST) & 1), dead, P.scal);
        }
    }
}

// ----------------------------------------------------------------- storer warp ----
// Drains the result tile block by block (64 rows x 16 columns, 4 rows per instruction) and folds
// dotProduct(z, r) (v3:374; masked chapters: non-fluid z is +-0.0 and contributes nothing).  r does
// not depend on the sweep: its 16 loads per block are issued BEFORE the wait for the block.
// Two storer warps share the work (even / odd blocks): one block lasts ~1 us of sweep, and a storer that
// also folds the dot product needs ~1.4 us per block (r from L2, 64 FP64 ops, 16 stores, proxy fence),
// which throttled the whole ring (measured: 168 instead of 121 cycles per step).
template <bool BWD, bool DOT, bool MASKED>
__device__ void storer_warp(const TriParams &P, unsigned char *smem, uint64_t *done, uint64_t *empty, int sj, int lane,
                            volatile int *dead, int which) {
    const int nbx = P.nbx;
    const int ty = BWD ? (P.nby - 1 - sj) : sj;
    const int y0 = ty * SR;
    const int rs = lane >> 3, cp = (lane & 7) * 2;     // row within a group of 4, first of this lane's two columns
    const int trow0 = (BWD ? 0 : 1) + rs;              // tile row of memory row y0 + rs
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0; // four fixed interleaved partial sums (shorter dependency chain)
    constexpr int PF = 4; // r is pulled into L2 this many of this warp's blocks ahead of its use
    if (DOT)
        for (int b = which; b < 2 * PF && b < nbx; b += 2) {
            const int tx = BWD ? (nbx - 1 - b) : b;
            const double *g = P.rdot + tx * BW + (size_t)(y0 + 2 * lane) * P.pitch;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g + P.pitch));
        }
    for (int b = which; b < nbx; b += 2) {
        const int st = b % NST;
        const int tx = BWD ? (nbx - 1 - b) : b;
        const int x = tx * BW + cp;
        const bool x0 = x < P.W, x1 = x + 1 < P.W;
        double2 rv[16];
        if (DOT) {
            if (b + 2 * PF < nbx) { // one 128-byte line per row of block b + 2 PF: lane t takes rows 2t, 2t + 1
                const int txp = BWD ? (nbx - 1 - (b + 2 * PF)) : (b + 2 * PF);
                const double *gp = P.rdot + txp * BW + (size_t)(y0 + 2 * lane) * P.pitch;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + P.pitch));
            }
            const double *g = P.rdot + x + (size_t)(y0 + rs) * P.pitch;
#pragma unroll
            for (int i = 0; i < 16; i++) rv[i] = *reinterpret_cast<const double2 *>(g + (size_t)(4 * i) * P.pitch); // pad rows / columns exist and hold zeros
        }
        mbar_wait(&done[st], (unsigned)((b / NST) & 1), dead, P.scal);
        if (TRI_EXP & 128) {
            if (lane == 0) mbar_arrive(&empty[st]);
            continue;
        }
        const unsigned char *stage = smem + (size_t)st * STAGE_BYTES;
        const double *tile = reinterpret_cast<const double *>(stage) + trow0 * BW + cp;
        double2 v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = *reinterpret_cast<const double2 *>(tile + (4 * i) * BW);
        if (DOT) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                acc0 += v[i].x * rv[i].x;
                acc1 += v[i].y * rv[i].y;
                acc2 += v[i + 1].x * rv[i + 1].x;
                acc3 += v[i + 1].y * rv[i + 1].y;
            }
        }
        double *g = P.dst + x + (size_t)(y0 + rs) * P.pitch;
        if (!MASKED) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                double *d = g + (size_t)(4 * i) * P.pitch;
                if (y0 + rs + 4 * i < P.H) {
                    if (x1)
                        *reinterpret_cast<double2 *>(d) = v[i];
                    else if (x0)
                        d[0] = v[i].x;
                }
            }
        } else { // chapters 4+: non-fluid cells keep their old value (v5:751-752); pe is non-zero at fluid cells
            const double *mt = reinterpret_cast<const double *>(stage + 3 * TILE_BYTES) + trow0 * BW + cp;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const double2 mk = *reinterpret_cast<const double2 *>(mt + (4 * i) * BW);
                double *d = g + (size_t)(4 * i) * P.pitch;
                if (y0 + rs + 4 * i < P.H) {
                    if (x0 && mk.x != 0.0) d[0] = v[i].x;
                    if (x1 && mk.y != 0.0) d[1] = v[i].y;
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (DOT) {
        const double sum = warp_sum((acc0 + acc1) + (acc2 + acc3));
        if (lane == 0) P.partials[2 * sj + which] = sum;
    }
}

// -------------------------------------------------------------- publisher warp ----
// Forwards the strip's last row to the downstream strip, HG columns at a time: it sleeps on the bell
// the compute warp rings for every completed group, reads the 8 values from the result tile and sends
// them (st.async into the downstream CTA's ring inside a cluster, LL messages through L2 / NVLink
// otherwise).  It holds each stage until its 16 columns have been sent (second arrival on done[]).
template <bool BWD>
__device__ void publisher_warp(const TriParams &P, unsigned char *smem, double *halo_s, uint64_t *done, uint64_t *bell,
                               uint64_t *hb, int sj, int lane, volatile int *dead, unsigned *counters, unsigned rank) {
    typedef Geo<BWD> G;
    const int ncols = P.nbx * BW;
    const double *last_row = reinterpret_cast<const double *>(smem) + G::last_row() * BW; // tile 0 of stage 0
    const bool remote = sj + 1 == P.sj_base + P.nloc; // the downstream strip belongs to another rank
    uint4 *out = (remote ? P.handoff_down : P.handoff) + (size_t)sj * ncols;
    const bool dsmem = P.cs > 1 && rank + 1 < (unsigned)P.cs && !remote;
    const uint32_t r_halo = dsmem ? mapa(smem_u32(halo_s), rank + 1) : 0;
    const uint32_t r_hb = dsmem ? mapa(smem_u32(hb), rank + 1) : 0;
    const uint32_t r_progress = dsmem ? mapa(smem_u32(&counters[0]), rank + 1) : 0;
    int down_progress = 0; // last value read from the downstream strip's own progress counter
    Watch watch;
    for (int c0 = 0; c0 < ncols; c0 += HG) {
        const int g = c0 / HG;
        mbar_wait(&bell[g % NBELL], (unsigned)((g / NBELL) & 1), dead, P.scal);
        const int blk = c0 / BW;
        const double *row = last_row + (size_t)(blk % NST) * (STAGE_BYTES / 8);
        const int c = c0 + lane;
        double v = 0.0;
        if (lane < HG) v = row[G::tcol(c % BW)];
        if (dsmem) {
            // ring slot c % HRC (and its group barrier) is free once the downstream strip's last row has
            // passed column c - HRC; never asked after the downstream strip may have left
            while (c0 + HG > HRC && down_progress < c0 + HG - HRC) {
                down_progress = (int)ld_remote_u32(r_progress);
                if (watch.expired(dead)) {
                    *dead = 1;
                    P.scal->watchdog = 1;
                    break;
                }
            }
            if (lane < HG)
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                                 r_halo + (uint32_t)(c % HRC) * 8u),
                             "l"(__double_as_longlong(v)), "r"(r_hb + (uint32_t)(g % NHB) * 8u)
                             : "memory");
        } else if (lane < HG) {
            if (remote)
                ll_store_sys(out + c, v, P.epoch);
            else
                ll_store(out + c, v, P.epoch);
        }
        if ((c0 + HG) % BW == 0) { // all 16 columns of this block are out: the stage may drain
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[blk % NST]);
        }
    }
}

// ---------------------------------------------------------------------- kernel ----
template <bool BWD, bool DOT, bool MASKED>
__global__ void __launch_bounds__(224, 1) k_tri(const __grid_constant__ TriParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[3 * NST]; // full[], done[], empty[]
    __shared__ uint64_t bell[NBELL];   // rung by the compute warp for every group its last row completes
    __shared__ uint64_t hb[NHB];       // hand-off groups received from the cluster neighbour (tx bytes)
    __shared__ int s_ticket;
    __shared__ int s_dead;
    __shared__ unsigned s_counters[2]; // [0] columns finished by the last row, [1] hand-off columns received
    double *halo_s = reinterpret_cast<double *>(smem + (size_t)NST * STAGE_BYTES); // [HRC] upstream last-row values
    uint64_t *full = bars, *done = bars + NST, *empty = bars + 2 * NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = P.cs > 1 ? cluster_ctarank() : 0;

    if (threadIdx.x == 0) {
        if (rank == 0) s_ticket = (int)(atomicAdd(P.ticket, 1ULL) - P.ticket_base);
        s_dead = 0;
        s_counters[0] = 0;
        s_counters[1] = (TRI_EXP & 2) ? (1u << 30) : 0u;
        // the barriers a cluster neighbour may touch exist (and are armed) before the cluster barrier
        for (int i = 0; i < NHB; i++) {
            mbar_init(&hb[i], 1);
            mbar_arrive_expect_tx(&hb[i], HG * 8);
        }
        for (int i = 0; i < NBELL; i++) mbar_init(&bell[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    int ticket;
    if (P.cs > 1) {
        cluster_sync_all(); // every CTA of the cluster is resident, its barriers are armed, rank 0's ticket is set
        ticket = (int)ld_remote_u32(mapa(smem_u32(&s_ticket), 0));
    } else {
        ticket = s_ticket;
    }
    const int sj = P.sj_base + ticket * P.cs + (int)rank;
    if (sj >= P.sj_base + P.nloc) return; // padding CTA of the last cluster
    if (P.gated && P.scal->done) return;  // the solve has converged
    if (threadIdx.x == 0) {
        const bool publish = sj + 1 < P.nby && !(TRI_EXP & 512);
        for (int i = 0; i < NST; i++) {
            mbar_init(&full[i], 1);               // loader's expect_tx arrival (+ TMA bytes)
            mbar_init(&done[i], publish ? 2 : 1); // compute warp (+ publisher warp)
            mbar_init(&empty[i], 1);              // storer warp
        }
        fence_mbar_init();
    }
    if (sj == 0) // the very first strip has no upstream row: its hand-off values read +0.0
        for (int i = threadIdx.x; i < HRC; i += blockDim.x) halo_s[i] = 0.0;
    __syncthreads();

    if (warp == 0) {
        unsigned long long t0 = 0;
        const long long c0 = clock64();
        if (P.times && lane == 0) t0 = globaltimer_ns();
        compute_warp<BWD>(P, smem, halo_s, done, bell, sj, lane, &s_dead, s_counters);
        if (P.times && lane == 0) {
            P.times[16 * sj] = t0;
            P.times[16 * sj + 1] = globaltimer_ns();
            P.times[16 * sj + 15] = (unsigned long long)(clock64() - c0);
        }
    } else if (TRI_EXP & 2) {
        // (experiment: no helper warps)
    } else if (warp == 1) {
        loader_warp<BWD>(P, smem, full, empty, sj, lane, &s_dead);
    } else if (warp == 2 || warp == 6) { // (warp 4 would share the compute warp's scheduler)
        storer_warp<BWD, DOT, MASKED>(P, smem, done, empty, sj, lane, &s_dead, warp == 2 ? 0 : 1);
    } else if (warp == 3) {
        if (sj + 1 < P.nby && !(TRI_EXP & 512)) publisher_warp<BWD>(P, smem, halo_s, done, bell, hb, sj, lane, &s_dead, s_counters, rank);
    } else if (warp == 5) {
        // first strip of a cluster (and of a rank): its upstream strip talks through L2 / NVLink
        gatekeeper_warp<BWD>(P, halo_s, full, (rank == 0 || sj == P.sj_base) ? nullptr : hb, sj, lane, &s_dead, s_counters);
    }
}

} // namespace tri

// ------------------------------------------------------------------- host side ----
static const Arr &precon_operand(ifl_ctx *c) { return c->version >= 4 ? c->pe : c->precon; }

template <bool BWD, bool DOT>
static int launch_tri(ifl_ctx *c, const Arr &rhs, const Arr &dst, const Arr *rdot, bool gated, unsigned band_target = 0) {
    using namespace tri;
    TriParams P;
    memset(&P, 0, sizeof P);
    const Arr *ops[NT] = {&rhs, &c->cx, &c->cy, &precon_operand(c)};
    for (int k = 0; k < NT; k++) {
        int rc = sweep_get_map(c, *ops[k], BW, TROWS, &P.map[k]);
        if (rc != IFL_OK) return rc;
    }
    P.dst = dst.p;
    P.rdot = rdot ? rdot->p : nullptr;
    P.W = c->W;
    P.H = c->H;
    P.pitch = c->r.pitch;
    P.nbx = c->r.pitch / BW;
    P.nby = (c->H + SR - 1) / SR;
    P.handoff = reinterpret_cast<uint4 *>(c->handoff);
    c->epoch++;
    P.epoch = (unsigned)(c->epoch & 0xffffffffu);
    if (P.epoch == 0) { // 0 is the value of never-written hand-off slots
        c->epoch++;
        P.epoch = 1;
    }
    {   // this rank's strips, in sweep order (slabs are whole 64-row strips, dist.cu)
        const int s0 = c->ry0 / SR, s1 = (c->ry1 + SR - 1) / SR;
        P.nloc = s1 - s0;
        P.sj_base = BWD ? P.nby - s1 : s0;
        P.handoff_down = reinterpret_cast<uint4 *>(c->handoff_down[BWD ? 1 : 0]);
    }
    const size_t smem = (size_t)NST * STAGE_BYTES + (size_t)HRC * sizeof(double);
    const bool masked = c->version >= 4;
    static bool attr_set[IFL_MAX_DEVICES][2][2][2]; // function attributes are per device
    static int max16[IFL_MAX_DEVICES][2][2][2];     // co-resident clusters of 16 (non-portable size), per kernel
    auto kern = masked ? k_tri<BWD, DOT, true> : k_tri<BWD, DOT, false>;
    if (!attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked]) {
        IFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int n16 = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t q;
            memset(&q, 0, sizeof q);
            q.gridDim = dim3(16 * 64);
            q.blockDim = dim3(224);
            q.dynamicSmemBytes = smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 16;
            qa[0].val.clusterDim.y = qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&n16, kern, &q) != cudaSuccess) n16 = 0;
        }
        cudaGetLastError();
        max16[c->device % IFL_MAX_DEVICES][BWD][DOT][masked] = n16;
        attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked] = true;
    }
    // Clusters of 16 (one per GPC) halve the hand-offs that travel through L2 -- each costs ~3 us against 2.9 inside a
    // cluster -- as long as every strip is resident at once; otherwise the portable size 8, which packs more strips.
    P.cs = c->sweep_cluster;
    if (c->tri_cluster16 && P.cs == 8 && (P.nloc + 15) / 16 <= max16[c->device % IFL_MAX_DEVICES][BWD][DOT][masked]) P.cs = 16;
    const int n_clusters = (P.nloc + P.cs - 1) / P.cs;
    P.ticket = c->ticket;
    P.ticket_base = c->sweep_tickets;
    c->sweep_tickets += (unsigned long long)n_clusters;
    c->sweep_launches++;
    P.scal = c->scal;
    P.gated = gated ? 1 : 0;
    P.head_delay = c->sweep_head_delay;
    P.times = c->sweep_times;
    P.band_count = band_target ? c->band_count : nullptr;
    P.band_target = band_target;
    if (DOT) {
        P.partials = partials_next(c);
        c->n_partials = 2 * P.nby; // one per storer warp
    }
    ProfScope ps_(c, BWD ? IFL_K_PRECON_BWD : IFL_K_PRECON_FWD);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(n_clusters * P.cs));
    cfg.blockDim = dim3(224);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = P.cs > 1 ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, P);
    if (le != cudaSuccess) {
        set_error("triangular solve launch (cluster %d) -> %s", P.cs, cudaGetErrorString(le));
        return IFL_E_CUDA;
    }
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_tri_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated, unsigned band_target) {
    return launch_tri<false, false>(c, a, dst, nullptr, gated, band_target);
}

int launch_tri_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    if (with_dot) return launch_tri<true, true>(c, dst, dst, &r_for_dot, gated);
    return launch_tri<true, false>(c, dst, dst, nullptr, gated);
}

} // namespace ifl
