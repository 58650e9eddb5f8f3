// sweep3_kernels.cu -- the two triangular solves of applyPreconditioner (v3:275-304) as
// ONE-WARP CTAs.  Same wavefront, same strips, same hand-off messages and the same
// arithmetic as sweep_kernels.cu (so the results are bit-identical); what changes is who
// does the chores.  On B200 the recurrence warp runs ~12 cycles per step slower as soon as
// any other warp of the CTA is alive, and polling shared counters that helper warps bump
// costs another ~8 (profiles/microbench/step.cu), so here the compute warp is alone on its
// SM and does everything itself, in the shadow of its own dependency stalls:
//   * lane 0 issues the TMA loads of the operand tiles two blocks ahead (mbarrier per stage)
//     and the TMA STORE of every finished result tile (the swept variable is updated in
//     place in its tile; cp.async.bulk.tensor shared -> global, clipped to W x H by the map);
//   * every 8 steps lanes 0..7 send the 8 values the strip's last row has just finished as
//     LL messages {lo, epoch, hi, epoch} to the downstream strip, and pick up the 8 upstream
//     messages they asked for 8 steps earlier (a global load that has long landed; only a
//     consumer that has caught up with its producer spins);
//   * the backward solve folds dotProduct(z, r) (v3:374) per lane, row by row.
// Chapters 1-3 (no solid cells); the masked variants, the factorisation and Gauss-Seidel stay
// with sweep_kernels.cu.  Opt-in (IFL_SWEEP_V3=1): measured slower than the five-warp engine at
// 4096^2 (688 vs 645 us per sweep), see DESIGN.md 4.
#include "ifl_internal.cuh"
#include "sweep_common.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace ifl {

int sweep_get_store_map(ifl_ctx *c, const Arr &a, CUtensorMap *out);

namespace s3 {

constexpr int TP = 32;                 // tile width in columns
constexpr int ROW_B = TP * 8;          // 256
constexpr int TILE_B = 33 * ROW_B;     // 8448: 32 strip rows + the upstream row
constexpr int NST = 5;                 // ring depth
constexpr int LOOK = 2;                // blocks loaded ahead of lane 0's block
constexpr int HG = 8;                  // hand-off granularity (columns)
constexpr int RING = 64;               // hand-off ring in shared memory (columns)
constexpr int MAXT = 5;

struct Params {
    CUtensorMap map[MAXT]; // [0] swept variable (forward: rhs in, z out; backward: z in place), [1] cx, [2] cy, [3] precon, [4] r
    CUtensorMap smap;      // store map of the swept result: 32 x 32 boxes, clipped to W x H
    int nt;
    int W, H, nbx, nby;
    uint4 *handoff, *handoff_down;
    int sj_base, nloc;
    unsigned epoch;
    unsigned long long *ticket;
    unsigned long long ticket_base;
    SolveScalars *scal;
    int gated;
    double *partials;
    unsigned long long *times;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, uint32_t smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
                 "r"(smem_src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

struct Ops {
    double a, cx, cy, pr, r, halo;
};
struct Carry {
    double z, c1, acc;
};
struct Bases {
    uint32_t A, B, N; // tile-0 byte addresses of blocks m, m-1, m+1 (lane row and skew folded in)
};

template <bool BWD>
__device__ __forceinline__ uint32_t pos(const Bases &b, int j, int lane) {
    constexpr int DIR = BWD ? -8 : 8;
    uint32_t base = (lane > j) ? b.B : b.A;
    if (j >= 32) base = (lane <= j - 32) ? b.N : base;
    return base + (uint32_t)(DIR * j);
}

template <bool BWD, bool DOT>
__device__ __forceinline__ void fetch(Ops &o, uint32_t p, uint32_t ph) {
    o.a = lds_f64(p);
    o.cx = lds_f64(p + 1 * TILE_B);
    o.cy = lds_f64(p + 2 * TILE_B + (BWD ? 0 : -ROW_B)); // forward: cy of the UPPER cell (v3:283)
    o.pr = lds_f64(p + 3 * TILE_B);
    if (DOT) o.r = lds_f64(p + 4 * TILE_B);
    o.halo = lds_f64(ph);
}

// Everything the warp needs to talk to its neighbours.  Pointers are per lane and advance
// with the macro-step, so that inside the unrolled loop every address is register + immediate.
struct Link {
    const uint4 *up;  // incoming messages: &up_row[32*m + lane]
    uint4 *down;      // outgoing messages: &down_row[32*(m-1) + lane]
    uint32_t ring;    // shared-memory hand-off ring (RING doubles = two blocks)
    uint32_t ring_w;  // &ring[(32*m & (RING-1)) + lane]: where lanes 0..7 drop column 32*m + lane (+ immediate)
    uint32_t pub;     // tile byte address of the last row's in-block column `lane` of block m-1
    unsigned epoch;
    bool has_up, publish;
};

__device__ __forceinline__ uint4 ll_read(const uint4 *p) { // system scope: the producer may be another GPU
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Cold path of consume_group: some of the 8 messages had not arrived yet.
__device__ __noinline__ uint4 poll_group(const uint4 *src, unsigned epoch, uint4 pre, bool ok, SolveScalars *scal) {
    unsigned n = 0;
    do {
        if (!ok) {
            pre = ll_read(src);
            ok = pre.y == epoch && pre.w == epoch;
            if (++n > WATCHDOG_POLLS) {
                scal->watchdog = 1;
                ok = true;
            }
        }
    } while (!__all_sync(0xffffffffu, ok));
    return pre;
}

// Lanes 0..7 hold in `pre` the messages of the 8 columns starting at in-block column `off`
// of block m (off == 32: first group of block m+1), loaded 8 steps ago.  Validate, spin on
// stragglers, drop the values into the ring, ask for the next group.
__device__ __forceinline__ void consume_group(const Link &L, uint4 &pre, int off, int lane, SolveScalars *scal) {
    const bool ok = lane >= HG || (pre.y == L.epoch && pre.w == L.epoch);
    if (!__all_sync(0xffffffffu, ok)) pre = poll_group(L.up + off, L.epoch, pre, ok, scal);
    const uint32_t w = (off < 32) ? L.ring_w + (uint32_t)(off * 8) : L.ring + ((L.ring_w - L.ring) ^ (32u * 8u));
    if (lane < HG) sts_f64(w, __hiloint2double((int)pre.z, (int)pre.x));
    __syncwarp();
    // (reads past the end of the row hit the next row of the hand-off array and are never used)
    if (lane < HG) pre = ll_read(L.up + off + HG);
}

// One macro-step (32 steps).  EDGE as in sweep_kernels.cu: 0 interior, 1 first, 2 last.
template <bool BWD, bool DOT, int EDGE>
__device__ __forceinline__ void macro_step(const Bases &bs, int m, int lane, Carry &cr, Ops &ops, const Link &L, uint4 &pre,
                                           uint64_t *full_next, unsigned parity_next, bool has_next, SolveScalars *scal) {
    uint32_t p = pos<BWD>(bs, 0, lane);
    const uint32_t ring_r = L.ring + (uint32_t)(((32 * m) & (RING - 1)) * 8); // lane 0's hand-off value of column 32m (+ imm)
#pragma unroll
    for (int kk = 0; kk < 32; kk++) {
        // lane 0 is about to fetch the hand-off value of column 32m+kk+1, the first of a new group
        if (((kk + 1) % HG) == 0 && L.has_up && EDGE != 2 && (kk != 31 || has_next)) consume_group(L, pre, kk + 1, lane, scal);
        if (kk == 31 && has_next) { // lane 0 is about to touch block m+1
            unsigned n = 0;
            while (!mbar_try(full_next, parity_next))
                if (++n > WATCHDOG_TRIES) {
                    scal->watchdog = 1;
                    break;
                }
        }
        double up = __shfl_up_sync(0xffffffffu, cr.z, 1);
        Ops nxt;
        const uint32_t pn = pos<BWD>(bs, kk + 1, lane);
        // ring position of column 32m+kk+1: RING is two blocks, so kk+1 == 32 is the other half
        fetch<BWD, DOT>(nxt, pn, (kk + 1 < 32) ? ring_r + (uint32_t)((kk + 1) * 8) : L.ring + ((ring_r - L.ring) ^ (32u * 8u)));
        const int d0 = kk - lane;
        const bool active = (EDGE == 0) ? true : (EDGE == 1 ? d0 >= 0 : d0 < 0);
        if (EDGE == 1) { // a lane enters the strip with a clean state (see sweep_kernels.cu cell())
            cr.z = sel_f64(d0 == 0, 0.0, cr.z);
            cr.c1 = sel_f64(d0 == 0, 0.0, cr.c1);
        }
        up = sel_f64(lane == 0, ops.halo, up);
        double t;
        if (!BWD)
            t = ops.a - cr.c1 * cr.z; // v3:281
        else
            t = ops.a - ops.cx * cr.z; // v3:297
        t = t - ops.cy * up;           // v3:283 / v3:299
        const double znew = t * ops.pr; // v3:285 / v3:301
        sts_f64_p<EDGE == 0>(p, znew, active);
        if (DOT) { // dotProduct(z, r), v3:374: every lane folds its row in sweep order
            const double term = znew * ops.r;
            cr.acc = (EDGE == 0 || active) ? cr.acc + term : cr.acc;
        }
        cr.z = znew;
        cr.c1 = ops.cx;
        // the last row has completed another group of HG columns, in-block columns kk-6 .. kk+1 of block m-1
        if (((kk + 2) % HG) == 0 && L.publish && EDGE != 1) {
            __syncwarp();
            if (lane < HG) {
                const double v = lds_f64(L.pub + (uint32_t)((BWD ? -8 : 8) * (kk - 6)));
                ll_store_sys(L.down + (kk - 6), v, L.epoch);
            }
        }
        ops = nxt;
        p = pn;
    }
}

template <bool BWD, bool DOT>
__global__ void __launch_bounds__(32, 1) k_sweep3(const __grid_constant__ Params P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[NST];
    const int lane = threadIdx.x;
    constexpr int DIR = BWD ? -8 : 8;
    constexpr int COL0 = BWD ? 31 : 0;

    int ticket = 0;
    if (lane == 0) ticket = (int)(atomicAdd(P.ticket, 1ULL) - P.ticket_base);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const int sj = P.sj_base + ticket;
    if (P.gated && P.scal->done) return;
    if (lane == 0) {
        for (int i = 0; i < NST; i++) mbar_init(&full[i], 1);
        fence_mbar_init();
    }
    const uint32_t stage_b = (uint32_t)P.nt * TILE_B;
    const uint32_t smem0 = smem_u32(smem_raw);
    double *ring_p = reinterpret_cast<double *>(smem_raw + NST * stage_b);
    for (int i = lane; i < RING; i += 32) ring_p[i] = 0.0; // the first strip has no upstream row: +0.0
    __syncwarp();

    const int nbx = P.nbx, ncols = nbx * 32;
    const int ty = BWD ? (P.nby - 1 - sj) : sj;
    const int box_y = BWD ? ty * 32 : ty * 32 - 1;
    Link L;
    L.has_up = sj > 0;
    L.publish = sj + 1 < P.nby;
    const uint4 *up_row = P.handoff + (size_t)(sj - 1) * ncols;
    uint4 *down_row = ((sj + 1 == P.sj_base + P.nloc) ? P.handoff_down : P.handoff) + (size_t)sj * ncols;
    L.ring = smem_u32(ring_p);
    asm volatile("mov.u32 %0, %1;" : "=r"(L.epoch) : "r"(P.epoch)); // a register, not a constant-bank reload per use
    L.up = up_row + lane;
    L.down = down_row + lane;
    L.ring_w = L.ring + (uint32_t)(lane * 8);
    L.pub = 0;

    auto load_block = [&](int b) { // lane 0 only
        const int st = b % NST;
        mbar_arrive_expect_tx(&full[st], (unsigned)stage_b);
        const int box_x = (BWD ? (nbx - 1 - b) : b) * 32;
        for (int k = 0; k < P.nt; k++) tma_load_2d(smem_raw + st * stage_b + k * TILE_B, &P.map[k], box_x, box_y, &full[st]);
    };
    if (lane == 0)
        for (int b = 0; b <= LOOK && b < nbx; b++) load_block(b);
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (L.has_up && lane < HG) pre = ll_read(up_row + lane);

    unsigned long long t0 = 0;
    const long long c0 = clock64();
    if (P.times && lane == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));

    Carry cr;
    cr.z = cr.c1 = cr.acc = 0.0;
    Ops ops;
    ops.r = 0.0;
    const uint32_t row0 = smem0 + (uint32_t)((BWD ? 31 - lane : 1 + lane) * ROW_B + COL0 * 8);
    {
        unsigned n = 0;
        while (!mbar_try(&full[0], 0))
            if (++n > WATCHDOG_TRIES) {
                P.scal->watchdog = 1;
                break;
            }
    }
    fetch<BWD, DOT>(ops, row0, L.ring);
    if (L.has_up) {
        consume_group(L, pre, 0, lane, P.scal);
        ops.halo = lds_f64(L.ring);
    }
    const uint32_t row31 = smem0 + (uint32_t)((BWD ? 0 : 32) * ROW_B); // tile row of the strip's last row, stage 0, tile 0
    const uint32_t store_row = smem0 + (uint32_t)(BWD ? 0 : ROW_B);    // first stored tile row
    for (int m = 0; m <= nbx; m++) {
        const int sp = (m + NST - 1) % NST, sc = m % NST, sn = (m + 1) % NST;
        const unsigned par_next = (unsigned)((m + 1) / NST) & 1u; // fill count of block m+1's stage
        if (m >= 1 && lane == 0 && m + LOOK < nbx) {
            // stage of block m+LOOK was last used by block m+LOOK-NST, stored at the end of macro-step
            // m+LOOK-NST+1; only the store of block m-2 may still be reading shared memory
            bulk_wait_read<1>();
            load_block(m + LOOK);
        }
        const bool has_next = m + 1 < nbx;
        const uint32_t skew = (uint32_t)(DIR * lane), blk = (uint32_t)(DIR * 32);
        const uint32_t s_prev = row0 + sp * stage_b, s_cur = row0 + sc * stage_b, s_next = row0 + sn * stage_b;
        Bases bs;
        L.up = up_row + 32 * m + lane;
        L.down = down_row + 32 * (m - 1) + lane;
        L.ring_w = L.ring + (uint32_t)((((32 * m) & (RING - 1)) + lane) * 8);
        L.pub = row31 + sp * stage_b + (uint32_t)((BWD ? 31 - lane : lane) * 8);
        if (m == 0) {
            bs.A = s_cur - skew;
            bs.B = bs.A;
            bs.N = has_next ? s_next - blk - skew : bs.A;
            macro_step<BWD, DOT, 1>(bs, m, lane, cr, ops, L, pre, &full[sn], par_next, has_next, P.scal);
        } else if (m == nbx) {
            bs.B = s_prev + blk - skew;
            bs.A = s_prev - skew;
            bs.N = bs.A;
            macro_step<BWD, DOT, 2>(bs, m, lane, cr, ops, L, pre, &full[sn], par_next, false, P.scal);
        } else {
            bs.A = s_cur - skew;
            bs.B = s_prev + blk - skew;
            bs.N = has_next ? s_next - blk - skew : bs.A;
            macro_step<BWD, DOT, 0>(bs, m, lane, cr, ops, L, pre, &full[sn], par_next, has_next, P.scal);
        }
        if (m >= 1) { // block m-1 is final: store its result tile (generic-proxy writes -> async-proxy read)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                const int b = m - 1;
                tma_store_2d(&P.smap, (BWD ? (nbx - 1 - b) : b) * 32, ty * 32, store_row + sp * stage_b);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    if (DOT) {
        const double sum = warp_sum(cr.acc);
        if (lane == 0) P.partials[sj] = sum;
    }
    if (P.times && lane == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        P.times[16 * sj] = t0;
        P.times[16 * sj + 1] = t1;
        P.times[16 * sj + 15] = (unsigned long long)(clock64() - c0);
    }
}

} // namespace s3

template <bool BWD, bool DOT>
static int launch3(ifl_ctx *c, const Arr &swept_in, const Arr &swept_out, const Arr *r, bool gated) {
    using namespace s3;
    Params P;
    memset(&P, 0, sizeof P);
    const Arr *in[MAXT] = {&swept_in, &c->cx, &c->cy, &c->precon, r};
    P.nt = DOT ? 5 : 4;
    for (int k = 0; k < P.nt; k++) {
        int rc = sweep_get_map(c, *in[k], 32, &P.map[k]);
        if (rc != IFL_OK) return rc;
    }
    int rc = sweep_get_store_map(c, swept_out, &P.smap);
    if (rc != IFL_OK) return rc;
    P.W = c->W;
    P.H = c->H;
    P.nbx = (c->W + 31) / 32;
    P.nby = (c->H + 31) / 32;
    P.handoff = reinterpret_cast<uint4 *>(c->handoff);
    {
        const int s0 = c->ry0 / 32, s1 = (c->ry1 + 31) / 32;
        P.nloc = s1 - s0;
        P.sj_base = BWD ? P.nby - s1 : s0;
        P.handoff_down = reinterpret_cast<uint4 *>(c->handoff_down[BWD ? 1 : 0]);
    }
    c->epoch++;
    P.epoch = (unsigned)(c->epoch & 0xffffffffu);
    if (P.epoch == 0) {
        c->epoch++;
        P.epoch = 1;
    }
    P.ticket = c->ticket;
    P.ticket_base = c->sweep_tickets;
    c->sweep_tickets += (unsigned long long)P.nloc;
    c->sweep_launches++;
    P.scal = c->scal;
    P.gated = gated ? 1 : 0;
    if (DOT) {
        P.partials = partials_next(c);
        c->n_partials = P.nby;
    }
    P.times = c->sweep_times;
    const size_t smem = (size_t)NST * P.nt * TILE_B + RING * sizeof(double);
    static bool attr_set[IFL_MAX_DEVICES][2][2]; // function attributes are per device
    if (!attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT]) {
        IFL_CUDA(cudaFuncSetAttribute(k_sweep3<BWD, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr_set[c->device % IFL_MAX_DEVICES][BWD][DOT] = true;
    }
    ProfScope ps_(c, BWD ? IFL_K_PRECON_BWD : IFL_K_PRECON_FWD);
    k_sweep3<BWD, DOT><<<P.nloc, 32, smem, c->stream>>>(P);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_precon_forward3(ifl_ctx *c, const Arr &dst, const Arr &a, bool gated) {
    return launch3<false, false>(c, a, dst, nullptr, gated);
}

int launch_precon_backward3(ifl_ctx *c, const Arr &dst, const Arr &r_for_dot, bool with_dot, bool gated) {
    if (with_dot) return launch3<true, true>(c, dst, dst, &r_for_dot, gated);
    return launch3<true, false>(c, dst, dst, nullptr, gated);
}

} // namespace ifl
