// stair_kernels.cu -- the two triangular solves of FluidSolver::applyPreconditioner
// (v3:275-304; masked form v5:746-780): the "staircase" engine.
//
//   forward   t = a[i] - cx[i-1]*z[i-1] - cy[i-w]*z[i-w] ; z[i] = t*precon[i]     v3:276-287
//   backward  t = z[i] - cx[i]*z[i+1]   - cy[i]*z[i+w]   ; z[i] = t*precon[i]     v3:289-303
//             (+ dotProduct(z, r) of v3:374 folded in by the storer warps)
//   cx = aPlusX*precon, cy = aPlusY*precon are written by the factorisation (sweep_kernels.cu);
//   the reference evaluates `_aPlusX[i]*_precon[i]*dst[i]` left to right (v3:281), so this is the
//   same product.  Any schedule that honours the (x-1,y),(x,y-1) dependencies performs the same
//   floating-point operations on the same operands as the raster loop: results are bit-identical.
//
// A sweep is a wave that has to cross W + H cells, and every cell is a mul-sub-sub-mul chain of
// FP64 operations (8.4 cycles each) that cannot be reassociated.  Its duration is
//     (W + strips * (skew + hand-off)) * T_step.
// STATUS: bit-exact and selectable (IFL_TRI=2), NOT the default.  Measured at 4096^2 it ties with the two-row
// engine of tri_kernels.cu (479-502 us per solve against 440-465): the step is shorter, 92-98 cycles against
// 117-126, but the strip skew doubles.  DESIGN.md section 4 item 10 and profiles/r02_tri_experiments.txt
// section 8 have the numbers and the breakdown of the step; the file stays because it is the better engine for
// wide, short slabs (W >> H, e.g. many GPUs) and because its staging scheme is what a faster step would need.
//
// The two-row engine gives a lane two vertically adjacent rows in the SAME column: cell B waits for cell A of
// the same step, so the loop-carried chain of a step is shuffle + 6 dependent FP64 operations (~65 cycles),
// with 31 steps of skew per 64-row strip.  Here cell B runs ONE COLUMN BEHIND cell A of its own lane: B takes
// A's value of the previous step from a register, A takes the B of the lane above (previous step) by shuffle,
// and the two chains of a step are independent.  The loop-carried cycle is B -> shuffle -> A -> B over two
// steps (~40 cycles per step) or a row's own mul-sub-sub-mul (34).  The price is a skew of two columns per
// lane, 63 steps per 64-row strip -- W + H steps in all, the length of the dependency chain itself.
//
// Geometry
//   * strip = 64 rows, one CTA (1 per SM).  Lane t = 8g + u owns rows 2t, 2t+1 of the strip (forward; the
//     backward sweep mirrors rows and columns); at step s its cell A is at column s - 2t, cell B at
//     s - 2t - 1.
//   * A rectangular operand block would have to stay in shared memory for 64 + 16 steps.  Instead a STAGE
//     is a staircase: four row groups of 16 rows (8 lanes), group g holding column block i - g of stage i
//     -- four TMA boxes of 17 rows x 16 columns per operand (row 0: the row above the group, for cy), with
//     different x coordinates.  Every lane then lives in stages m - 1 and m during macro-step m (16 steps),
//     a stage is released one macro-step after it was entered, and a 6-stage ring leaves four stages
//     (~2 us) of TMA prefetch.  Ring: 6 x 4 operands x 4 x 2176 B = 208,896 B.
//   * all lanes of a step read the same column parity, so 8-byte accesses would pair up on the banks;
//     operands are fetched as aligned PAIRS of columns (LDS.128, two steps per load, conflict-free for each
//     quarter-warp) and results stored the same way: 4.5 loads and one store per step for two cells.
//   * warps: 0 compute, 1 TMA loader, 2 + 6 storers (drain z, fold z.r), 3 publisher, 5 gatekeeper;
//     nobody polls shared memory while the compute warp runs (bells, st.async + complete_tx inside a
//     cluster, LL messages through L2 / NVLink between clusters and GPUs) -- see tri_kernels.cu, whose
//     helper-warp protocol this file keeps.
#include "ifl_internal.cuh"
#include "sweep_common.cuh"

#include <cuda.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

// Timing experiments (profiles/microbench/step_s.cu builds one binary per value; never set in the product
// build): bit 0 no gate checks, bit 1 no progress store / bell, bit 2 no lane-0 hand-off select, bit 3 no
// result stores, bit 4 no operand fetches, bit 5 no shuffle; in the engine (profiles/stair_experiments.sh): bit 6 the
// storers wait and release without draining, bit 7 the loader arrives without loading, bit 8 no hand-off (no publisher,
// every strip runs as if it were the first), bit 9 no gatekeeper either (gate preset; needs bits 7 and 8), bit 10 no
// done[] arrival (needs bit 11), bit 11 no storers / loader at all (needs bit 9).  Results are garbage, only the timing is of interest.
#ifndef STAIR_EXP
#define STAIR_EXP 0
#endif

namespace ifl {
namespace stair {

constexpr int SR = 64;                         // rows per strip
constexpr int BW = 16;                         // columns per block
constexpr int NG = 4;                          // row groups per strip
constexpr int GR = SR / NG;                    // rows per group (two per lane, 8 lanes)
constexpr int GROWS = GR + 1;                  // box rows (one upstream row)
constexpr int ROWB = BW * 8;                   // bytes per tile row
constexpr int GT_BYTES = GROWS * ROWB;         // 2176 (multiple of 128)
constexpr int TILE_BYTES = NG * GT_BYTES;      // 8704
constexpr int NT = 4;                          // tiles per stage: rhs/z, cx, cy, precon
constexpr int STAGE_BYTES = NT * TILE_BYTES;   // 34816
constexpr int NST = 6;                         // ring depth
constexpr int HG = 8;                          // hand-off granularity (columns)
constexpr int HRC = 512;                       // hand-off ring (columns)
constexpr int NHB = HRC / HG;                  // hand-off group barriers
constexpr int NBELL = 16;                      // bell ring
constexpr int LAG = 63;                        // steps between lane 0's cell A and lane 31's cell B in one column
constexpr int XS = NG - 1;                     // extra stages behind the last column block (groups 1..3 trail)

struct Params {
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
    double *partials; // z.r per strip and storer
    int cs;           // cluster size
    int head_delay;
    const unsigned *band_count; // forward sweep overlapped with k_axpy2_norm: finished blocks per 64-row band (or null)
    unsigned band_target;       // ... a strip may read its band of the rhs once the counter has reached this
    unsigned long long *times;
};

template <bool BWD>
struct Geo {
    static constexpr int DIR = BWD ? -8 : 8;         // bytes per logical column
    static constexpr int COL0 = BWD ? BW - 1 : 0;    // tile column of logical in-block column 0
    static constexpr int ROW_B = BWD ? -ROWB : ROWB; // cell A -> cell B (same column)
    static constexpr int UP_CY = BWD ? 0 : -ROWB;    // own cell -> the cell whose cy multiplies the upstream value
    static constexpr int PAIR = BWD ? -8 : 0;        // address of step j -> address of the aligned pair (j, j+1)
    // box row of cell A of lane u (of its group); the box holds memory rows [16g - 1, 16g + 15] (forward),
    // [48 - 16g, 64 - 16g] (backward, g counted in sweep order)
    __device__ static __forceinline__ int row_a(int u) { return BWD ? GR - 1 - 2 * u : 1 + 2 * u; }
    __device__ static __forceinline__ int last_row() { return BWD ? 0 : GR; } // box row of the strip's last row (group 3)
    __device__ static __forceinline__ int tcol(int ci) { return BWD ? BW - 1 - ci : ci; }
};

struct Pair {
    double lo, hi; // in address order
};
__device__ __forceinline__ Pair lds_pair(uint32_t a) {
    Pair v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.lo), "=d"(v.hi) : "r"(a) : "memory");
    return v;
}
template <bool ALWAYS>
__device__ __forceinline__ void sts_pair(uint32_t a, double lo, double hi, bool pred) {
    if (ALWAYS) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(lo), "d"(hi) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %3, 0;\n\t"
            "@p st.shared.v2.f64 [%0], {%1, %2};\n\t"
            "}" ::"r"(a),
            "d"(lo), "d"(hi), "r"((unsigned)pred)
            : "memory");
    }
}
// value of the first / second step of a pair (the backward sweep walks the tile towards lower addresses)
template <bool BWD>
__device__ __forceinline__ double first_of(const Pair &p) { return BWD ? p.hi : p.lo; }
template <bool BWD>
__device__ __forceinline__ double second_of(const Pair &p) { return BWD ? p.lo : p.hi; }

// the operands of one cell for two consecutive steps, and where they came from (the result pair goes back there)
struct Quad {
    Pair a, x, y, p;
    uint32_t addr;
};
template <bool BWD>
__device__ __forceinline__ void fetch(Quad &q, uint32_t pa) {
    q.addr = pa;
    q.a = lds_pair(pa);
    q.x = lds_pair(pa + TILE_BYTES);
    q.y = lds_pair(pa + 2 * TILE_BYTES + (uint32_t)Geo<BWD>::UP_CY);
    q.p = lds_pair(pa + 3 * TILE_BYTES);
}

// Per-cell tile-0 base pointers of one macro-step.  A cell with in-stage offset r (cell A of lane u: 2u, cell
// B: 2u + 1) is in stage m - 1 ("before") for the steps j < r of macro-step m and in stage m ("after") from
// step r on; steps 16, 17 (the look-ahead into the next macro-step) are still in stage m for j - 16 < r and in
// stage m + 1 ("next") otherwise.  The address of step j is always `selected base + DIR*j`, j a compile-time
// constant.
struct LaneBases {
    uint32_t A, B, N;
};
template <bool BWD>
__device__ __forceinline__ uint32_t pos(const LaneBases &lb, int j, int r) {
    uint32_t base;
    if (j < BW)
        base = (r > j) ? lb.B : lb.A;
    else
        base = (r > j - BW) ? lb.A : lb.N;
    return base + (uint32_t)(Geo<BWD>::DIR * j);
}

struct State {
    double zA, zB; // swept variable of the previous step (A: previous column of row A; B likewise)
    double cA, cB; // forward: cx of the previous column
    double kA, kB; // first result of the pair being assembled
    Quad qA, qA_n; // cell A: operands of steps (2i, 2i+1) / the next pair
    Quad qB, qB_n; // cell B: operands of steps (2i-1, 2i) / the next pair
    Pair h, h_n;   // hand-off values of the same steps (lane 0)
};

// EDGE 0: every cell is inside the strip.  EDGE 1: general (lanes enter one by one; also leaves, for grids
// narrower than the warp's skew), EDGE 2: lanes leave.  Cells outside run the same instructions on aliased
// operands with their stores predicated off; what they compute is never consumed (cell B of lane t-1 is inside
// at step k-1 exactly when cell A of lane t is inside at step k, and likewise A -> B inside a lane).
template <bool BWD, int EDGE>
__device__ __forceinline__ void macro_step(const LaneBases &la, const LaneBases &lbb, uint32_t h_cur, uint32_t h_next, uint32_t bell6,
                                           uint32_t bell14, uint32_t done_addr, int m, int lane, int rA, State &s, uint32_t progress_addr,
                                           uint32_t gate_addr, int ncols, int vcols, volatile int *dead, SolveScalars *scal) {
    typedef Geo<BWD> G;
    const int rB = rA + 1;
    unsigned gate_seen = 0;
#pragma unroll
    for (int kk = 0; kk < BW; kk++) {
        // gate (hand-off values received AND operand stages loaded, in columns of lane 0): read four steps early,
        // tested when the pair fetch is about to enter the next group of HG columns
        if ((kk % HG) == 2 && !(STAIR_EXP & 1)) gate_seen = lds_u32_volatile(gate_addr);
        // every cell left stage m-2 at the end of the previous macro-step: hand it to the storers (and, through them, the
        // loader; the storer issues the proxy fence before the TMA may overwrite the stage).  The last store into it was
        // issued six steps (~500 cycles) ago by this same warp, so lane 0's arrive needs no warp-wide fence -- a
        // __syncwarp() + arrive at the macro-step boundary drains every load in flight and cost 8.5 cycles per step.
        if (kk == 4 && !(STAIR_EXP & 1024)) {
            if (lane == 0 && done_addr != 0) mbar_arrive_addr(done_addr);
        }
        if ((kk % HG) == 6 && !(STAIR_EXP & 1)) {
            const unsigned need = (unsigned)imin(BW * m + kk + 2 + HG, vcols);
            if (gate_seen < need) wait_counter<false>(gate_addr, need, dead, scal);
        }
        // ---- critical path first: cell B of the lane above, previous step (lane 0: the hand-off value)
        double up = (STAIR_EXP & 32) ? s.zB : __shfl_up_sync(0xffffffffu, s.zB, 1);
        // ---- operands two steps ahead, in the shadow of the shuffle: cell A's pair on even steps, B's on odd
        if (STAIR_EXP & 16) {
        } else if ((kk & 1) == 0) {
            fetch<BWD>(s.qA_n, pos<BWD>(la, kk + 2, rA) + (uint32_t)G::PAIR);
            s.h_n = lds_pair(kk + 2 < BW ? h_cur + (uint32_t)(8 * (kk + 2)) : h_next + (uint32_t)(8 * (kk + 2 - BW)));
        } else {
            fetch<BWD>(s.qB_n, pos<BWD>(lbb, kk + 2, rB) + (uint32_t)G::PAIR);
        }
        // ---- this step's operands
        const bool even = (kk & 1) == 0;
        const double aA = even ? first_of<BWD>(s.qA.a) : second_of<BWD>(s.qA.a);
        const double xA = even ? first_of<BWD>(s.qA.x) : second_of<BWD>(s.qA.x);
        const double yA = even ? first_of<BWD>(s.qA.y) : second_of<BWD>(s.qA.y);
        const double pA = even ? first_of<BWD>(s.qA.p) : second_of<BWD>(s.qA.p);
        const double aB = even ? second_of<BWD>(s.qB.a) : first_of<BWD>(s.qB.a);
        const double xB = even ? second_of<BWD>(s.qB.x) : first_of<BWD>(s.qB.x);
        const double yB = even ? second_of<BWD>(s.qB.y) : first_of<BWD>(s.qB.y);
        const double pB = even ? second_of<BWD>(s.qB.p) : first_of<BWD>(s.qB.p);
        const double halo = even ? s.h.lo : s.h.hi;
        const int cA = BW * m + kk - 2 * lane; // logical column of cell A; cell B is at cA - 1
        bool actA = true, actB = true;
        if (EDGE == 1) {
            actA = cA >= 0 && cA < ncols;
            actB = cA >= 1 && cA <= ncols;
            const bool firstA = cA == 0, firstB = cA == 1;
            s.zA = sel_f64(firstA, 0.0, s.zA); // (cell B of this step is outside when firstA holds)
            s.zB = sel_f64(firstB, 0.0, s.zB);
            if (!BWD) {
                s.cA = sel_f64(firstA, 0.0, s.cA);
                s.cB = sel_f64(firstB, 0.0, s.cB);
            }
        } else if (EDGE == 2) {
            actA = cA < ncols;
            actB = cA <= ncols;
        }
        if (!(STAIR_EXP & 4)) up = sel_f64(lane == 0, halo, up);
        const double zA_prev = s.zA; // cell A one column back: the row above cell B's column
        double zA, zB;
        if (!BWD) {
            double t = aA - s.cA * s.zA; // v3:281  t -= aPlusX[idx-1]*precon[idx-1]*dst[idx-1]
            t = t - yA * up;             // v3:283  t -= aPlusY[idx-w]*precon[idx-w]*dst[idx-w]
            zA = t * pA;                 // v3:285
            double w = aB - s.cB * s.zB;
            w = w - yB * zA_prev;
            zB = w * pB;
            s.cA = xA;
            s.cB = xB;
        } else {
            double t = aA - xA * s.zA; // v3:297  t -= aPlusX[idx]*precon[idx]*dst[idx+1]
            t = t - yA * up;           // v3:299  t -= aPlusY[idx]*precon[idx]*dst[idx+w]
            zA = t * pA;               // v3:301
            double w = aB - xB * s.zB;
            w = w - yB * zA_prev;
            zB = w * pB;
        }
        s.zA = zA;
        s.zB = zB;
        // ---- results go back in place, pair by pair: the rhs / z tile becomes the result tile
        if (STAIR_EXP & 8) {
            s.kA += zA;
            s.kB += zB;
            if (even) s.qB = s.qB_n; else { s.qA = s.qA_n; s.h = s.h_n; }
        } else if (even) {
            s.kA = zA;
            if (BWD)
                sts_pair<EDGE == 0>(s.qB.addr, zB, s.kB, actB);
            else
                sts_pair<EDGE == 0>(s.qB.addr, s.kB, zB, actB);
            s.qB = s.qB_n;
        } else {
            s.kB = zB;
            if (BWD)
                sts_pair<EDGE == 0>(s.qA.addr, zA, s.kA, actA);
            else
                sts_pair<EDGE == 0>(s.qA.addr, s.kA, zA, actA);
            s.qA = s.qA_n;
            s.h = s.h_n;
        }
        // the strip's last row (lane 31, cell B) has just completed another group of HG columns: publish
        // the count and ring the publisher's bell (lane 31 wrote those values itself: its arrive releases them)
        if (((kk + 2) % HG) == 0 && EDGE != 1 && !(STAIR_EXP & 2)) {
            sts_u32_volatile(progress_addr, (unsigned)imin(BW * m + kk - (LAG - 1), ncols));
            if (lane == 31) mbar_arrive_addr(kk < HG ? bell6 : bell14);
        }
    }
}

template <bool BWD>
__device__ void compute_warp(const Params &P, unsigned char *smem, double *halo_s, uint64_t *done, uint64_t *bell, int sj, int lane,
                             volatile int *dead, unsigned *counters) {
    typedef Geo<BWD> G;
    State s;
    s.zA = s.zB = s.cA = s.cB = s.kA = s.kB = 0.0;
    const bool has_up = sj > 0;
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    const int nbx = P.nbx, ncols = nbx * BW;
    const int ns = nbx + XS, vcols = ns * BW; // stages, and the gate value that says "all of them are loaded"
    const int g = lane >> 3, u = lane & 7, rA = 2 * u;
    // this lane's row A in tile 0 of stage 0 at logical in-block column 0
    const uint32_t row0 = smem_u32(smem) + (uint32_t)(g * GT_BYTES + G::row_a(u) * ROWB + G::COL0 * 8);
    const uint32_t halo0 = smem_u32(halo_s);
    auto bases_of = [&](int sA, int sB, int sN, uint32_t row, int r) {
        LaneBases lb;
        lb.A = row + (uint32_t)(sA * STAGE_BYTES) - (uint32_t)(G::DIR * r);
        lb.B = row + (uint32_t)(sB * STAGE_BYTES) + (uint32_t)(G::DIR * (BW - r));
        lb.N = row + (uint32_t)(sN * STAGE_BYTES) - (uint32_t)(G::DIR * (BW + r));
        return lb;
    };
    auto clip = [&](int i) { return imin(imax(i, 0), ns - 1) % NST; }; // stages outside [0, ns) are only touched by cells outside the strip
    // everything that does not depend on the upstream strip happens BEFORE the wait for its first
    // hand-off group: that wait sits on the critical path of the whole sweep
    wait_counter<false>(gate_addr, (unsigned)imin(HG, ncols), dead, P.scal); // stage 0 loaded, first hand-off group here
    {
        const LaneBases la = bases_of(clip(0), clip(-1), clip(1), row0, rA);
        const LaneBases lbb = bases_of(clip(0), clip(-1), clip(1), row0 + (uint32_t)G::ROW_B, rA + 1);
        fetch<BWD>(s.qA, pos<BWD>(la, 0, rA) + (uint32_t)G::PAIR);       // steps 0, 1
        fetch<BWD>(s.qB, pos<BWD>(lbb, -1, rA + 1) + (uint32_t)G::PAIR); // steps -1, 0 (outside: never stored)
        fetch<BWD>(s.qB_n, pos<BWD>(lbb, 1, rA + 1) + (uint32_t)G::PAIR); // steps 1, 2
        s.h = lds_pair(halo0);
        s.qA_n = s.qA;
        s.h_n = s.h;
    }
    const int nm = nbx + NG; // macro-steps: lane 31's cell B finishes column ncols-1 at step ncols + 62
    int sm = 0;              // m % NST, kept incrementally
    // Three loops (lanes entering / steady state / lanes leaving) rather than one loop with a three-way branch: the
    // steady-state loop then has a single back edge and no register shuffling where the variants join.
    auto run = [&](int m, auto edge_tag) {
        constexpr int EDGE = decltype(edge_tag)::value;
        if (!has_up && P.head_delay > 0) { // pace-setter, see sweep_init
            const long long t_ = clock64();
            while (clock64() - t_ < P.head_delay) {}
        }
        int a, b, n;
        if (m >= 1 && m + 1 < ns) { // no stage index leaves [0, ns): by rotation
            a = sm;
            b = a == 0 ? NST - 1 : a - 1;
            n = a == NST - 1 ? 0 : a + 1;
        } else {
            a = clip(m);
            b = clip(m - 1);
            n = clip(m + 1);
        }
        const LaneBases la = bases_of(a, b, n, row0, rA);
        const LaneBases lbb = bases_of(a, b, n, row0 + (uint32_t)G::ROW_B, rA + 1);
        const uint32_t h_cur = halo0 + (uint32_t)(((BW * m) & (HRC - 1)) * 8);
        const uint32_t h_next = halo0 + (uint32_t)(((BW * (m + 1)) & (HRC - 1)) * 8);
        // groups completed by the last row at kk == 6 / 14 of this macro-step: 2m-8, 2m-7 (m >= 4)
        const uint32_t bell6 = smem_u32(&bell[(2 * m + NBELL - 8) & (NBELL - 1)]), bell14 = smem_u32(&bell[(2 * m + NBELL - 7) & (NBELL - 1)]);
        const uint32_t done_addr = m >= 2 ? smem_u32(&done[sm >= 2 ? sm - 2 : sm - 2 + NST]) : 0u; // stage m-2, released at step 4
        macro_step<BWD, EDGE>(la, lbb, h_cur, h_next, bell6, bell14, done_addr, m, lane, rA, s, progress_addr, gate_addr, ncols, vcols, dead, P.scal);
        sm = sm == NST - 1 ? 0 : sm + 1;
    };
    int m = 0;
    for (; m < imin(NG, nm); m++) run(m, std::integral_constant<int, 1>());
    for (; m < nbx; m++) run(m, std::integral_constant<int, 0>());
    for (; m < nm; m++) run(m, std::integral_constant<int, 2>());
    if (!(STAIR_EXP & 1024)) { // the last stage (ns-1 = nm-2)
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[(ns - 1) % NST]);
    }
}

// ----------------------------------------------------------------- loader warp ----
// Stage i: group g's box is column block i - g (skipped when that block does not exist).
template <bool BWD>
__device__ void loader_warp(const Params &P, unsigned char *smem, uint64_t *full, uint64_t *empty, int sj, int lane,
                            volatile int *dead) {
    if (lane != 0) return;
    const int nbx = P.nbx, ns = nbx + XS;
    const int ty = BWD ? (P.nby - 1 - sj) : sj; // memory strip of this CTA
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
    for (int i = 0; i < ns; i++) {
        const int st = i % NST;
        if (i >= NST) mbar_wait_sleep(&empty[st], (unsigned)(((i / NST) - 1) & 1), dead, P.scal);
        const int g_lo = imax(0, i - (nbx - 1)), g_hi = imin(NG - 1, i); // groups whose block i - g exists
        if (STAIR_EXP & 128) {
            mbar_arrive(&full[st]);
            continue;
        }
        mbar_arrive_expect_tx(&full[st], (unsigned)((g_hi - g_lo + 1) * NT * GT_BYTES));
        unsigned char *stage = smem + (size_t)st * STAGE_BYTES;
        for (int g = g_lo; g <= g_hi; g++) {
            const int b = i - g;
            const int box_x = (BWD ? (nbx - 1 - b) : b) * BW;
            const int box_y = BWD ? ty * SR + (SR - GR) - GR * g : ty * SR + GR * g - 1;
            for (int k = 0; k < NT; k++) tma_load_2d(stage + k * TILE_BYTES + g * GT_BYTES, &P.map[k], box_x, box_y, &full[st]);
        }
    }
}

// ------------------------------------------------------------- gatekeeper warp ----
// Releases the compute warp group by group (HG columns of lane 0) through counters[1]: a group is released
// when the upstream strip's last-row values for it are in halo_s (indexed by logical column mod HRC) AND
// the stage it lies in has landed; behind the last column the gate keeps counting, 16 per trailing stage,
// up to vcols.  `hb` != null: the upstream strip runs in the same cluster and sends st.async + complete_tx
// on hb[group % NHB] (armed here with expect_tx); otherwise the values arrive as LL messages in global
// memory (L2, or NVLink for the first strip of a rank).
// The gate never waits for a stage further ahead than the compute warp, fed by the gate published so far,
// can free: it is published stage by stage.
template <bool BWD>
__device__ void gatekeeper_warp(const Params &P, double *halo_s, uint64_t *full, uint64_t *hb, int sj, int lane,
                                volatile int *dead, unsigned *counters) {
    const int ncols = P.nbx * BW, ns = P.nbx + XS;
    const bool has_up = sj > 0 && !(STAIR_EXP & 256);
    const uint4 *up_row = P.handoff + (size_t)(has_up ? sj - 1 : 0) * ncols;
    const bool remote = sj == P.sj_base; // the upstream strip belongs to another rank
    const uint32_t progress_addr = smem_u32(&counters[0]), gate_addr = smem_u32(&counters[1]);
    Watch watch;
    int loaded = 1; // stages known to have landed
    mbar_wait_hint(&full[0], 0, dead, P.scal);
    if (has_up && !hb) {
        // LL messages in global memory.  Lane l owns columns l, l + 32, l + 64, ... and walks them on its own,
        // two polls in flight, so the L2 round trip is pipelined across columns (profiles/r02_tri_experiments.txt).
        // The gate is the contiguous prefix of received columns, rounded down to whole groups.
        int next_c = lane;     // this lane's first column not yet received
        unsigned released = 0; // columns released to the compute warp
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
                while (released < groups) {
                    const unsigned upto = umin(groups, (unsigned)(loaded * BW));
                    if (upto > released) {
                        if (lane == 0) sts_u32_volatile(gate_addr, upto);
                        released = upto;
                    }
                    if (released < groups) {
                        mbar_wait_hint(&full[loaded % NST], (unsigned)((loaded / NST) & 1), dead, P.scal);
                        loaded++;
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
    } else {
        for (int c0 = 0; c0 < ncols; c0 += HG) {
            const int g = c0 / HG;
            // a ring slot may be rewritten once the strip's own last row has passed the column it held
            if (has_up && c0 + HG > HRC) wait_counter(progress_addr, (unsigned)(c0 + HG - HRC), dead, P.scal);
            if (has_up) mbar_wait_hint(&hb[g % NHB], (unsigned)((g / NHB) & 1), dead, P.scal);
            if (lane == 0) sts_u32_volatile(gate_addr, (unsigned)(c0 + HG));
            // ---- off the critical path: arm the group barrier for its next use (NHB groups from now) and make sure
            // the stage of the NEXT group has landed (the TMA ring runs far ahead: this returns at once)
            if (has_up && lane == 0) mbar_arrive_expect_tx(&hb[g % NHB], HG * 8);
            if ((c0 + HG) % BW == 0 && c0 + HG < ncols) {
                const int b = (c0 + HG) / BW;
                mbar_wait_hint(&full[b % NST], (unsigned)((b / NST) & 1), dead, P.scal);
                loaded = b + 1;
            }
        }
    }
    // the trailing stages (row groups 1..3 finishing the last column blocks)
    __syncwarp();
    for (int b = imax(loaded, 1); b < ns && !*dead; b++) {
        mbar_wait_hint(&full[b % NST], (unsigned)((b / NST) & 1), dead, P.scal);
        if (lane == 0) sts_u32_volatile(gate_addr, (unsigned)((b + 1) * BW));
    }
}

// ----------------------------------------------------------------- storer warp ----
// Drains the result tile stage by stage (4 groups x 16 rows x 16 columns, 4 rows per instruction; group g
// of stage i is column block i - g) and folds dotProduct(z, r) (v3:374; masked chapters: non-fluid z is
// +-0.0 and contributes nothing).  r does not depend on the sweep: its loads are issued BEFORE the wait
// for the stage.  Two storer warps share the work (even / odd stages).
template <bool BWD, bool DOT, bool MASKED>
__device__ void storer_warp(const Params &P, unsigned char *smem, uint64_t *done, uint64_t *empty, int sj, int lane,
                            volatile int *dead, int which) {
    const int nbx = P.nbx, ns = nbx + XS;
    const int ty = BWD ? (P.nby - 1 - sj) : sj;
    const int y0 = ty * SR;
    const int rs = lane >> 3, cp = (lane & 7) * 2; // row within a group of 4, first of this lane's two columns
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0; // four fixed interleaved partial sums (shorter dependency chain)
    constexpr int PF = 4; // r is pulled into L2 this many of this warp's stages ahead of its use
    auto prefetch_r = [&](int i) { // the two 128-byte lines of rows 2t, 2t+1 in the column block their group has in stage i
        const int g = BWD ? NG - 1 - (2 * lane) / GR : (2 * lane) / GR;
        const int b = i - g;
        if (b < 0 || b >= nbx) return;
        const int tx = BWD ? (nbx - 1 - b) : b;
        const double *p = P.rdot + tx * BW + (size_t)(y0 + 2 * lane) * P.pitch;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + P.pitch));
    };
    if (DOT)
        for (int i = which; i < 2 * PF && i < ns; i += 2) prefetch_r(i);
    for (int i = which; i < ns; i += 2) {
        const int st = i % NST;
        // memory rows 4k .. 4k+3 of the strip (k = 0..15) belong to group k/4 (forward) or 3 - k/4 (backward)
        int xg[NG];   // global column of this lane's pair, per group-of-rows quarter (k / 4)
        bool vg[NG];  // that block exists
#pragma unroll
        for (int q = 0; q < NG; q++) {
            const int g = BWD ? NG - 1 - q : q;
            const int b = i - g;
            vg[q] = b >= 0 && b < nbx;
            xg[q] = (BWD ? (nbx - 1 - b) : b) * BW + cp;
        }
        double2 rv[16];
        if (DOT) {
            if (i + 2 * PF < ns) prefetch_r(i + 2 * PF);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                rv[k] = make_double2(0.0, 0.0);
                if (vg[k >> 2]) // pad rows / columns exist and hold zeros
                    rv[k] = *reinterpret_cast<const double2 *>(P.rdot + xg[k >> 2] + (size_t)(y0 + rs + 4 * k) * P.pitch);
            }
        }
        mbar_wait_sleep(&done[st], (unsigned)((i / NST) & 1), dead, P.scal);
        if (STAIR_EXP & 64) {
            if (lane == 0) mbar_arrive(&empty[st]);
            continue;
        }
        const unsigned char *stage = smem + (size_t)st * STAGE_BYTES;
        // box row of memory row (y0 + rs + 4k): forward 1 + (row - 16g), backward row - (48 - 16g) = rs + 4(k & 3)
        const int brow = (BWD ? 0 : 1) + rs;
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int g = BWD ? NG - 1 - (k >> 2) : (k >> 2);
            const double *tile = reinterpret_cast<const double *>(stage + g * GT_BYTES) + (brow + 4 * (k & 3)) * BW + cp;
            v[k] = *reinterpret_cast<const double2 *>(tile);
        }
        if (DOT) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                if (vg[k >> 2]) { // (a group without a block holds stale bytes)
                    acc0 += v[k].x * rv[k].x;
                    acc1 += v[k].y * rv[k].y;
                    acc2 += v[k + 1].x * rv[k + 1].x;
                    acc3 += v[k + 1].y * rv[k + 1].y;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int x = xg[k >> 2];
            const bool x0 = x < P.W, x1 = x + 1 < P.W;
            if (!vg[k >> 2] || y0 + rs + 4 * k >= P.H) continue;
            double *d = P.dst + x + (size_t)(y0 + rs + 4 * k) * P.pitch;
            if (!MASKED) {
                if (x1)
                    *reinterpret_cast<double2 *>(d) = v[k];
                else if (x0)
                    d[0] = v[k].x;
            } else { // chapters 4+: non-fluid cells keep their old value (v5:751-752); pe is non-zero at fluid cells
                const int g = BWD ? NG - 1 - (k >> 2) : (k >> 2);
                const double2 mk = *reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(stage + 3 * TILE_BYTES + g * GT_BYTES) +
                                                                       (brow + 4 * (k & 3)) * BW + cp);
                if (x0 && mk.x != 0.0) d[0] = v[k].x;
                if (x1 && mk.y != 0.0) d[1] = v[k].y;
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
// Forwards the strip's last row (group 3: column block b lives in stage b + 3) to the downstream strip, HG
// columns at a time: it sleeps on the bell the compute warp rings for every completed group, reads the 8
// values from the result tile and sends them (st.async into the downstream CTA's ring inside a cluster, LL
// messages through L2 / NVLink otherwise).  It holds each stage until its 16 columns have been sent
// (second arrival on done[]); the first three stages hold nothing of the last row.
template <bool BWD>
__device__ void publisher_warp(const Params &P, unsigned char *smem, double *halo_s, uint64_t *done, uint64_t *bell,
                               uint64_t *hb, int sj, int lane, volatile int *dead, unsigned *counters, unsigned rank) {
    typedef Geo<BWD> G;
    const int ncols = P.nbx * BW;
    const double *last_row = reinterpret_cast<const double *>(smem + XS * GT_BYTES) + G::last_row() * BW; // tile 0 of stage 0
    const bool remote = sj + 1 == P.sj_base + P.nloc; // the downstream strip belongs to another rank
    uint4 *out = (remote ? P.handoff_down : P.handoff) + (size_t)sj * ncols;
    const bool dsmem = P.cs > 1 && rank + 1 < (unsigned)P.cs && !remote;
    const uint32_t r_halo = dsmem ? mapa(smem_u32(halo_s), rank + 1) : 0;
    const uint32_t r_hb = dsmem ? mapa(smem_u32(hb), rank + 1) : 0;
    const uint32_t r_progress = dsmem ? mapa(smem_u32(&counters[0]), rank + 1) : 0;
    int down_progress = 0; // last value read from the downstream strip's own progress counter
    Watch watch;
    if (lane < XS) mbar_arrive(&done[lane]); // (NST > XS)
    for (int c0 = 0; c0 < ncols; c0 += HG) {
        const int g = c0 / HG;
        mbar_wait_hint(&bell[g % NBELL], (unsigned)((g / NBELL) & 1), dead, P.scal);
        const int stg = c0 / BW + XS;
        const double *row = last_row + (size_t)(stg % NST) * (STAGE_BYTES / 8);
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
            if (lane == 0) mbar_arrive(&done[stg % NST]);
        }
    }
}

// ---------------------------------------------------------------------- kernel ----
template <bool BWD, bool DOT, bool MASKED>
__global__ void __launch_bounds__(224, 1) k_stair(const __grid_constant__ Params P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[3 * NST]; // full[], done[], empty[]
    __shared__ uint64_t bell[NBELL];   // rung by the compute warp for every group its last row completes
    __shared__ uint64_t hb[NHB];       // hand-off groups received from the cluster neighbour (tx bytes)
    __shared__ int s_ticket;
    __shared__ int s_dead;
    __shared__ unsigned s_counters[2]; // [0] columns finished by the last row, [1] the gate
    double *halo_s = reinterpret_cast<double *>(smem + (size_t)NST * STAGE_BYTES); // [HRC] upstream last-row values
    uint64_t *full = bars, *done = bars + NST, *empty = bars + 2 * NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = P.cs > 1 ? cluster_ctarank() : 0;

    if (threadIdx.x == 0) {
        if (rank == 0) s_ticket = (int)(atomicAdd(P.ticket, 1ULL) - P.ticket_base);
        s_dead = 0;
        s_counters[0] = 0;
        s_counters[1] = (STAIR_EXP & 512) ? (1u << 30) : 0u;
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
        const bool publish = sj + 1 < P.nby && !(STAIR_EXP & 256);
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
    } else if (STAIR_EXP & 2048) {
    } else if (warp == 1) {
        loader_warp<BWD>(P, smem, full, empty, sj, lane, &s_dead);
    } else if (warp == 2 || warp == 6) { // (warp 4 would share the compute warp's scheduler)
        storer_warp<BWD, DOT, MASKED>(P, smem, done, empty, sj, lane, &s_dead, warp == 2 ? 0 : 1);
    } else if (warp == 3) {
        if (sj + 1 < P.nby && !(STAIR_EXP & 256)) publisher_warp<BWD>(P, smem, halo_s, done, bell, hb, sj, lane, &s_dead, s_counters, rank);
    } else if (warp == 5 && !(STAIR_EXP & 512)) {
        // first strip of a cluster (and of a rank): its upstream strip talks through L2 / NVLink
        gatekeeper_warp<BWD>(P, halo_s, full, (rank == 0 || sj == P.sj_base) ? nullptr : hb, sj, lane, &s_dead, s_counters);
    }
}

} // namespace stair

#ifndef IFL_STAIR_DEVICE_ONLY
// ------------------------------------------------------------------- host side ----
static const Arr &stair_precon_operand(ifl_ctx *c) { return c->version >= 4 ? c->pe : c->precon; }

template <bool BWD, bool DOT>
static int launch_stair(ifl_ctx *c, const Arr &rhs, const Arr &dst, const Arr *rdot, bool gated, unsigned band_target = 0) {
    using namespace stair;
    Params P;
    memset(&P, 0, sizeof P);
    const Arr *ops[NT] = {&rhs, &c->cx, &c->cy, &stair_precon_operand(c)};
    for (int k = 0; k < NT; k++) {
        int rc = sweep_get_map(c, *ops[k], BW, GROWS, &P.map[k]);
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
    P.cs = c->sweep_cluster;
    {   // this rank's strips, in sweep order (slabs are whole 64-row strips, dist.cu)
        const int s0 = c->ry0 / SR, s1 = (c->ry1 + SR - 1) / SR;
        P.nloc = s1 - s0;
        P.sj_base = BWD ? P.nby - s1 : s0;
        P.handoff_down = reinterpret_cast<uint4 *>(c->handoff_down[BWD ? 1 : 0]);
    }
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
    const size_t smem = (size_t)NST * STAGE_BYTES + (size_t)HRC * sizeof(double);
    const bool masked = c->version >= 4;
    static bool attr_set[IFL_MAX_DEVICES][2][2][2]; // function attributes are per device
    auto kern = masked ? k_stair<BWD, DOT, true> : k_stair<BWD, DOT, false>;
    if (!attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked]) {
        IFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked] = true;
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

int launch_stair_forward(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated, unsigned band_target) {
    return launch_stair<false, false>(c, a, dst, nullptr, gated, band_target);
}

int launch_stair_backward(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    if (with_dot) return launch_stair<true, true>(c, dst, dst, &r_for_dot, gated);
    return launch_stair<true, false>(c, dst, dst, nullptr, gated);
}

#endif // IFL_STAIR_DEVICE_ONLY

} // namespace ifl
