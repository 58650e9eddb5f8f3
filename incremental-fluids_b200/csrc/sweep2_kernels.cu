// sweep2_kernels.cu -- the two triangular solves of applyPreconditioner (v3:275-304,
// v5:746-780), which are ~80 % of a PCG iteration, as a TWO-COLUMNS-PER-STEP wavefront.
//
// Same CTA anatomy, hand-off protocol and exactness argument as sweep_kernels.cu (strips of
// 32 rows, ticket order, five specialised warps, LL messages between strips, +0.0 padding
// instead of boundary predicates).  What changes is the compute warp's schedule:
//
//   * lane t owns row t and handles the column PAIR j = k - t at step k (skew of two columns
//     per lane).  Inside a step the two cells are a serial 8-op FP64 chain (64 cycles
//     measured floor), but the upper neighbours of both cells were produced by lane t-1 one
//     whole step earlier, so the warp shuffle (24 cycles) and lane 0's hand-off select are
//     OFF the critical path.  The one-column schedule pays shuffle + 3 ops per column.
//   * operands are fetched as 16-byte pairs (all lanes have the same column parity), which
//     halves the shared-memory instructions of the in-order compute warp.
//   * blocks are 16 columns wide with a 12-stage ring of 4 tiles (the swept variable is
//     updated in place in its tile), i.e. the same 200 KB of shared memory now hold the
//     64-column skew window plus ~100 columns of TMA prefetch.
//
// Tiles of a stage: [0] swept variable (forward: a in, z out; backward: z in place),
// [1] cx = aPlusX*precon, [2] cy = aPlusY*precon, [3] precon (pe for chapters 4+).
#include "ifl_internal.cuh"
#include "sweep_common.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace ifl {

namespace s2 {

constexpr int BW = 16;                  // block width in columns
constexpr int PB = BW / 2;              // pairs per block
constexpr int ROW_B = BW * 8;           // tile row pitch in bytes (128)
constexpr int TROWS = 33;               // 32 strip rows + the upstream row
constexpr int TILE_B = TROWS * ROW_B;   // 4224
constexpr int NT = 4;                   // tiles per stage
constexpr int NST = 12;                 // ring depth
constexpr int STAGE_B = NT * TILE_B;    // 16896
constexpr int RING_COLS = 512;          // hand-off ring (columns)
constexpr int HG = 8;                   // hand-off granularity (columns)
constexpr int DEPTH = 4;                // lane 31 trails lane 0 by 31 pairs = DEPTH blocks (rounded up)

struct Params {
    CUtensorMap map[NT];
    double *store;  // global array receiving tile 0
    const double *r; // backward + dot: r (read by the storer straight from HBM)
    int W, H, pitch, nbx, nby; // nbx = number of 16-column blocks
    uint4 *handoff;            // [nby][nbx*16]
    unsigned epoch;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    SolveScalars *scal;
    int gated;
    double *partials;
    unsigned long long *times;
};

__device__ __forceinline__ double2 lds_v2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
template <bool ALWAYS>
__device__ __forceinline__ void sts_v2(uint32_t a, double x, double y, bool pred) {
    if (ALWAYS) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.u32 p, %3, 0;\n\t"
            "@p st.shared.v2.f64 [%0], {%1, %2};\n\t"
            "}" ::"r"(a),
            "d"(x), "d"(y), "r"((unsigned)pred)
            : "memory");
    }
}

// operands of one column pair, in sweep order (first = the column processed first)
struct Ops2 {
    double a1, a2, cx1, cx2, cy1, cy2, pr1, pr2, h1, h2;
};

template <bool BWD>
__device__ __forceinline__ void fetch2(Ops2 &o, uint32_t p, uint32_t ph) {
    // p: tile 0, this lane's row, this pair.  Backward sweeps walk memory right to left, so
    // the pair arrives as {second, first}.
    const double2 a = lds_v2(p), cx = lds_v2(p + TILE_B), pr = lds_v2(p + 3 * TILE_B);
    const double2 cy = lds_v2(p + 2 * TILE_B + (BWD ? 0 : -ROW_B)); // forward: cy of the UPPER cell (v3:283)
    const double2 h = lds_v2(ph); // hand-off ring is in sweep order for both directions (lane 0 only uses it)
    o.a1 = BWD ? a.y : a.x;   o.a2 = BWD ? a.x : a.y;
    o.cx1 = BWD ? cx.y : cx.x; o.cx2 = BWD ? cx.x : cx.y;
    o.cy1 = BWD ? cy.y : cy.x; o.cy2 = BWD ? cy.x : cy.y;
    o.pr1 = BWD ? pr.y : pr.x; o.pr2 = BWD ? pr.x : pr.y;
    o.h1 = h.x;
    o.h2 = h.y;
}

struct Carry2 {
    double zA, zB; // results of this lane's previous pair (zB is the left neighbour of the next cell)
    double c1;     // forward: cx of the previous column
};

// per-lane tile-0 byte addresses of the blocks this lane touches during one macro-step
struct Bases {
    uint32_t cur;  // block m - a      (pair slot = step - r,     for step >= r)
    uint32_t prev; // block m - a - 1  (pair slot = step - r + 8, for step <  r)
    uint32_t next; // block m - a + 1  (pair slot 0: only lanes with r == 0, when prefetching step 8)
};

template <bool BWD>
__device__ __forceinline__ uint32_t pos2(const Bases &b, int step, int r) { // step is a compile-time 0..8
    constexpr int DP = BWD ? -16 : 16; // bytes per pair in sweep direction
    uint32_t base = (step >= r) ? b.cur : b.prev;
    if (step == PB) base = (r == 0) ? b.next - (uint32_t)(DP * PB) : base;
    return base + (uint32_t)(DP * step);
}

template <bool BWD, bool EDGE>
__device__ __forceinline__ void macro_step2(const Bases &bs, int r, uint32_t halo_base, int m, int lane, int npairs,
                                            int ncols, Carry2 &cr, Ops2 &ops, uint64_t *full_next,
                                            unsigned parity_next, bool wait_next, uint32_t progress_addr,
                                            uint32_t halo_cols_addr, bool has_up, volatile int *dead, SolveScalars *scal) {
#pragma unroll
    for (int kk = 0; kk < PB; kk++) {
        const int j = PB * m + kk - lane; // this lane's pair
        // lane 0 is about to prefetch the first pair of block m+1 / of the next hand-off group
        if (kk == PB - 1 && wait_next) mbar_wait(full_next, parity_next, dead, scal);
        if (has_up && ((2 * (kk + 1)) % HG) == 0) {
            const int need = 2 * (PB * m + kk + 1) + HG;
            wait_counter(halo_cols_addr, (unsigned)imin(need, ncols), dead, scal);
        }
        // upper neighbours: lane t-1's previous pair (issued first, consumed 2 resp. 6 FP ops later)
        double upA = __shfl_up_sync(0xffffffffu, cr.zA, 1);
        double upB = __shfl_up_sync(0xffffffffu, cr.zB, 1);
        // operands of the next step
        Ops2 nxt;
        const uint32_t pn = pos2<BWD>(bs, kk + 1, r);
        {
            const int cn = 2 * (PB * m + kk + 1); // lane 0's next column
            const uint32_t ph = halo_base + (uint32_t)((cn & (RING_COLS - 1)) * 8);
            fetch2<BWD>(nxt, pn, ph);
        }
        upA = sel_f64(lane == 0, ops.h1, upA);
        upB = sel_f64(lane == 0, ops.h2, upB);
        // two cells, serial
        double zA, zB;
        if (!BWD) {
            double t = ops.a1 - cr.c1 * cr.zB; // v3:281
            t = t - ops.cy1 * upA;             // v3:283
            zA = t * ops.pr1;                  // v3:285
            t = ops.a2 - ops.cx1 * zA;
            t = t - ops.cy2 * upB;
            zB = t * ops.pr2;
        } else {
            double t = ops.a1 - ops.cx1 * cr.zB; // v3:297
            t = t - ops.cy1 * upA;               // v3:299
            zA = t * ops.pr1;                    // v3:301
            t = ops.a2 - ops.cx2 * zA;
            t = t - ops.cy2 * upB;
            zB = t * ops.pr2;
        }
        const bool active = !EDGE || (j >= 0 && j < npairs);
        const uint32_t p = pos2<BWD>(bs, kk, r);
        sts_v2<!EDGE>(p, BWD ? zB : zA, BWD ? zA : zB, active);
        if (EDGE) {
            cr.zA = sel_f64(active, zA, cr.zA);
            cr.zB = sel_f64(active, zB, cr.zB);
            cr.c1 = sel_f64(active, ops.cx2, cr.c1);
        } else {
            cr.zA = zA;
            cr.zB = zB;
            cr.c1 = ops.cx2;
        }
        // the strip's last row has completed another hand-off group
        if (((kk + 2) % (HG / 2)) == 0) sts_u32_volatile(progress_addr, (unsigned)imax(2 * (PB * m + kk - 30), 0));
        ops = nxt;
    }
}

template <bool BWD>
__device__ void compute_warp2(const Params &P, uint32_t smem, uint32_t halo_base, uint64_t *full, uint64_t *done,
                              int sj, int lane, volatile int *dead, unsigned *counters) {
    constexpr int DP = BWD ? -16 : 16;
    const int nbx = P.nbx, npairs = nbx * PB, ncols = nbx * BW;
    const bool has_up = sj > 0;
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    const int a = lane >> 3, r = lane & 7;
    // this lane's row in tile 0 of stage 0, at pair slot 0
    const uint32_t row0 = smem + (uint32_t)((BWD ? 31 - lane : 1 + lane) * ROW_B + (BWD ? ROW_B - 16 : 0));
    Carry2 cr;
    cr.zA = cr.zB = cr.c1 = 0.0;
    Ops2 ops;
    mbar_wait(&full[0], 0, dead, P.scal);
    if (has_up) wait_counter(halo_cols_addr, (unsigned)imin(HG, ncols), dead, P.scal);
    fetch2<BWD>(ops, row0, halo_base);
    // stage of block m - a (clamped to block 0 while the lane has not started)
    int blk = -a;
    int st_cur = 0;
    unsigned par_next = 0; // parity of full[] for block m+1
    int st_lead = 1 % NST; // stage of block m+1
    const int last_m = nbx + DEPTH - 1;
    for (int m = 0; m <= last_m; m++) {
        Bases bs;
        {
            const int bc = blk < 0 ? 0 : (blk >= nbx ? nbx - 1 : blk);
            const int bp = blk - 1 < 0 ? 0 : (blk - 1 >= nbx ? nbx - 1 : blk - 1);
            const int bn = blk + 1 < 0 ? 0 : (blk + 1 >= nbx ? nbx - 1 : blk + 1);
            bs.cur = row0 + (uint32_t)((bc % NST) * STAGE_B) - (uint32_t)(DP * r);
            bs.prev = row0 + (uint32_t)((bp % NST) * STAGE_B) + (uint32_t)(DP * (PB - r));
            bs.next = row0 + (uint32_t)((bn % NST) * STAGE_B);
        }
        const bool has_next = m + 1 < nbx;
        if (m >= DEPTH && m < nbx)
            macro_step2<BWD, false>(bs, r, halo_base, m, lane, npairs, ncols, cr, ops, &full[st_lead], par_next, has_next,
                                    progress_addr, halo_cols_addr, has_up, dead, P.scal);
        else
            macro_step2<BWD, true>(bs, r, halo_base, m, lane, npairs, ncols, cr, ops, &full[st_lead], par_next, has_next,
                                   progress_addr, halo_cols_addr, has_up, dead, P.scal);
        if (m >= DEPTH) { // lane 31 has left block m - DEPTH
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[(m - DEPTH) % NST]);
        }
        blk++;
        (void)st_cur;
        st_lead++;
        if (st_lead == NST) {
            st_lead = 0;
            par_next ^= 1u;
        }
    }
}

// ------------------------------------------------------------------- TMA loader ----
template <bool BWD>
__device__ void loader_warp2(const Params &P, unsigned char *smem, uint64_t *full, uint64_t *empty, int sj, int lane,
                             volatile int *dead) {
    if (lane != 0) return;
    const int nbx = P.nbx;
    const int ty = BWD ? (P.nby - 1 - sj) : sj;
    const int box_y = BWD ? ty * 32 : ty * 32 - 1;
    for (int b = 0; b < nbx; b++) {
        const int st = b % NST;
        if (b >= NST) mbar_wait(&empty[st], ((b / NST) - 1) & 1, dead, P.scal);
        mbar_arrive_expect_tx(&full[st], (unsigned)STAGE_B);
        const int box_x = (BWD ? (nbx - 1 - b) : b) * BW;
        for (int k = 0; k < NT; k++) tma_load_2d(smem + st * STAGE_B + k * TILE_B, &P.map[k], box_x, box_y, &full[st]);
    }
}

// ------------------------------------------------------------- hand-off poller ----
__device__ void poller_warp2(const Params &P, double *halo_s, int sj, int lane, volatile int *dead, unsigned *counters) {
    const int ncols = P.nbx * BW;
    const uint4 *up_row = P.handoff + (size_t)(sj - 1) * ncols;
    const uint32_t progress_addr = smem_u32(&counters[0]), halo_cols_addr = smem_u32(&counters[1]);
    unsigned polls = 0;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        // ring slot reuse: lane 31 of this strip (62+ columns behind lane 0) must have passed it
        if (c0 + 32 > RING_COLS) wait_counter(progress_addr, (unsigned)(c0 + 32 - RING_COLS), dead, P.scal);
        const int c = c0 + lane;
        const bool valid_col = c < ncols;
        const uint4 *src = up_row + c;
        bool have = !valid_col;
        const unsigned span = (unsigned)imin(32, ncols - c0);
        unsigned published = 0;
        while (published < span) {
            if (!have) {
                double v;
                if (ll_load(src, P.epoch, v)) {
                    halo_s[c & (RING_COLS - 1)] = v;
                    have = true;
                    polls = 0;
                } else if (++polls > WATCHDOG_POLLS || *dead) {
                    *dead = 1;
                    P.scal->watchdog = 1;
                    have = true;
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, have);
            const unsigned lead = (mask == 0xffffffffu) ? 32u : (unsigned)(__ffs(~mask) - 1);
            unsigned groups = lead / HG * HG;
            if (groups > span) groups = span;
            if (groups > published) {
                __threadfence_block();
                if (lane == 0) sts_u32_volatile(halo_cols_addr, (unsigned)c0 + groups);
                published = groups;
            }
        }
    }
}

// ----------------------------------------------------------------------- storer ----
template <bool BWD, bool DOT, bool MASKED>
__device__ void storer_warp2(const Params &P, unsigned char *smem, uint64_t *done, uint64_t *empty, int sj, int lane,
                             volatile int *dead) {
    const int ty = BWD ? (P.nby - 1 - sj) : sj;
    const int y0 = ty * 32;
    const int sub = lane >> 3, l8 = lane & 7; // 4 rows per instruction, 8 lanes x 16 bytes per row
    const int trow0 = (BWD ? 0 : 1) + sub;
    double acc = 0.0;
    for (int b = 0; b < P.nbx; b++) {
        const int st = b % NST;
        mbar_wait(&done[st], (b / NST) & 1, dead, P.scal);
        const unsigned char *stage = smem + st * STAGE_B;
        const int tx = BWD ? (P.nbx - 1 - b) : b;
        const int x = tx * BW + l8 * 2;
        const bool x0 = x < P.W, x1 = x + 1 < P.W;
        double2 v[8], mk[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int off = (trow0 + 4 * i) * ROW_B + l8 * 16;
            v[i] = *reinterpret_cast<const double2 *>(stage + off);
            if (MASKED) mk[i] = *reinterpret_cast<const double2 *>(stage + 3 * TILE_B + off);
        }
        if (DOT) { // dotProduct(z, r), v3:374 -- r straight from HBM (pad cells are zero on both sides)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int y = y0 + sub + 4 * i;
                const double2 rv = *reinterpret_cast<const double2 *>(P.r + x + (size_t)y * P.pitch);
                acc += v[i].x * rv.x;
                acc += v[i].y * rv.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int y = y0 + sub + 4 * i;
            if (y < P.H) {
                double *dst = P.store + x + (size_t)y * P.pitch;
                if (!MASKED) {
                    if (x1)
                        *reinterpret_cast<double2 *>(dst) = v[i];
                    else if (x0)
                        dst[0] = v[i].x;
                } else { // chapters 4+: non-fluid cells keep their old value (v5:751-752)
                    if (x0 && mk[i].x != 0.0) dst[0] = v[i].x;
                    if (x1 && mk[i].y != 0.0) dst[1] = v[i].y;
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

// -------------------------------------------------------------------- publisher ----
template <bool BWD>
__device__ void publisher_warp2(const Params &P, unsigned char *smem, uint64_t *done, int sj, int lane,
                                volatile int *dead, unsigned *counters) {
    const int ncols = P.nbx * BW;
    const unsigned char *last_row = smem + (BWD ? 0 : 32) * ROW_B; // lane 31's row of tile 0
    uint4 *out = P.handoff + (size_t)sj * ncols;
    const uint32_t progress_addr = smem_u32(&counters[0]);
    int sent = 0, blk = 0, st = 0;
    unsigned n = 0;
    while (sent < ncols) {
        const int prog = (int)lds_u32_volatile(progress_addr);
        if (prog > sent) {
            while (sent < prog) {
                const int blk_end = (blk + 1) * BW;
                const int upto = prog < blk_end ? prog : blk_end;
                const int c = sent + lane;
                if (c < upto) {
                    const int ci = c & (BW - 1);
                    const double z = *reinterpret_cast<const double *>(last_row + st * STAGE_B + (BWD ? BW - 1 - ci : ci) * 8);
                    ll_store(out + c, z, P.epoch);
                }
                sent = upto;
                if (sent == blk_end) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&done[st]);
                    blk++;
                    if (++st == NST) st = 0;
                }
            }
            n = 0;
        } else if (++n > WATCHDOG_POLLS || *dead) {
            *dead = 1;
            P.scal->watchdog = 1;
            for (int b = blk; b < P.nbx; b++)
                if (lane == 0) mbar_arrive(&done[b % NST]);
            return;
        }
    }
}

// ----------------------------------------------------------------------- kernel ----
template <bool BWD, bool DOT, bool MASKED>
__global__ void __launch_bounds__(160, 1) k_sweep2(const __grid_constant__ Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[3 * NST];
    __shared__ int s_strip;
    __shared__ int s_dead;
    __shared__ unsigned s_counters[2];
    double *halo_s = reinterpret_cast<double *>(smem_raw + NST * STAGE_B);
    uint64_t *full = bars, *done = bars + NST, *empty = bars + 2 * NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        s_strip = (int)(atomicAdd(P.ticket, 1ULL) - P.ticket_base); // ticket order, see sweep_kernels.cu
        s_dead = 0;
        s_counters[0] = 0;
        s_counters[1] = 0;
    }
    __syncthreads();
    const int sj = s_strip;
    if (P.gated && P.scal->done) return;
    if (threadIdx.x == 0) {
        const bool publish = sj + 1 < P.nby;
        for (int i = 0; i < NST; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], publish ? 2 : 1);
            mbar_init(&empty[i], 1);
        }
        fence_mbar_init();
    }
    if (sj == 0)
        for (int i = threadIdx.x; i < RING_COLS; i += blockDim.x) halo_s[i] = 0.0; // no upstream row: +0.0
    __syncthreads();

    if (warp == 0) {
        unsigned long long t0 = 0;
        const long long c0 = clock64();
        if (P.times && lane == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        compute_warp2<BWD>(P, smem_u32(smem_raw), smem_u32(halo_s), full, done, sj, lane, &s_dead, s_counters);
        if (P.times && lane == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            P.times[16 * sj] = t0;
            P.times[16 * sj + 1] = t1;
            P.times[16 * sj + 15] = (unsigned long long)(clock64() - c0);
        }
    } else if (warp == 1) {
        loader_warp2<BWD>(P, smem_raw, full, empty, sj, lane, &s_dead);
    } else if (warp == 2) {
        storer_warp2<BWD, DOT, MASKED>(P, smem_raw, done, empty, sj, lane, &s_dead);
    } else if (warp == 3) {
        if (sj + 1 < P.nby) publisher_warp2<BWD>(P, smem_raw, done, sj, lane, &s_dead, s_counters);
    } else if (sj > 0) {
        poller_warp2(P, halo_s, sj, lane, &s_dead, s_counters);
    }
}

} // namespace s2

// -------------------------------------------------------------------- host side ----
template <bool BWD, bool DOT>
static int launch2(ifl_ctx *c, const Arr &swept_in, const Arr &swept_out, const Arr &pre, const Arr *r, bool gated,
                   bool masked) {
    using namespace s2;
    Params P;
    memset(&P, 0, sizeof P);
    const Arr *in[NT] = {&swept_in, &c->cx, &c->cy, &pre};
    for (int k = 0; k < NT; k++) {
        int rc = sweep_get_map(c, *in[k], BW, &P.map[k]);
        if (rc != IFL_OK) return rc;
    }
    P.store = swept_out.p;
    P.r = r ? r->p : nullptr;
    P.W = c->W;
    P.H = c->H;
    P.pitch = c->r.pitch;
    P.nbx = (c->W + BW - 1) / BW;
    P.nby = (c->H + 31) / 32;
    P.handoff = reinterpret_cast<uint4 *>(c->handoff);
    c->epoch++;
    P.epoch = (unsigned)(c->epoch & 0xffffffffu);
    if (P.epoch == 0) {
        c->epoch++;
        P.epoch = 1;
    }
    P.ticket = c->ticket;
    P.ticket_base = c->sweep_tickets;
    c->sweep_tickets += (unsigned long long)P.nby;
    c->sweep_launches++;
    P.scal = c->scal;
    P.gated = gated ? 1 : 0;
    P.partials = c->partials;
    P.times = c->sweep_times;
    if (DOT) c->n_partials = P.nby;
    const size_t smem = (size_t)NST * STAGE_B + RING_COLS * sizeof(double);
    auto kern = masked ? k_sweep2<BWD, DOT, true> : k_sweep2<BWD, DOT, false>;
    static bool attr_set[IFL_MAX_DEVICES][2][2][2]; // function attributes are per device
    if (!attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked]) {
        IFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT][masked] = true;
    }
    ProfScope ps_(c, BWD ? IFL_K_PRECON_BWD : IFL_K_PRECON_FWD);
    kern<<<P.nby, 160, smem, c->stream>>>(P);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_precon_forward2(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated) {
    const bool masked = c->version >= 4;
    return launch2<false, false>(c, a, dst, masked ? c->pe : c->precon, nullptr, gated, masked);
}

int launch_precon_backward2(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    const bool masked = c->version >= 4;
    if (with_dot) return launch2<true, true>(c, dst, dst, masked ? c->pe : c->precon, &r_for_dot, gated, masked);
    return launch2<true, false>(c, dst, dst, masked ? c->pe : c->precon, nullptr, gated, masked);
}

} // namespace ifl
