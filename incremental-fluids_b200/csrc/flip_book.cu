// flip_book.cu -- chapter 8 particle bookkeeping and its extrapolation, exactly as the reference runs them.
//
//   frand                 v8:38-46    31-bit LCG, function-static seed 0xBA5EBA11
//   initParticles         v8:735-751  jittered grid, _AvgPerCell attempts per cell, rejects land in bodies
//   countParticles        v8:754-763
//   pruneParticles        v8:766-786  swap-with-last removal in cells holding more than 12 particles
//   seedParticles         v8:789-813  up to 3 - count new particles per cell
//   particlesToGrid       v8:916-927
//   extrapolate (+ fillSolidMask, extrapolateAverage, extrapolateEmptyBorders)  v8:478-651
//
// All four bookkeeping loops are sequential in the reference and consume ONE random stream, so the
// order of everything matters for bit parity:
//   * frand: draw number d of the process is a pure function of d (LCG jump-ahead, O(log d)), so any
//     thread can produce any draw.  The context keeps the number of draws made so far.
//   * init / seed: the attempts are enumerated in the reference's raster order, every attempt owns two
//     draws whether it is accepted or not, an exclusive prefix sum over the accept flags gives each
//     accepted attempt the slot the sequential loop would have given it (float
//     addition `x + frand()`, rejection test on the particle whose index is the CELL index).
//   * prune: which particles go depends on the scan order, and particles moved in from the tail are
//     re-examined at the hole.  Only particles in overfull cells can ever be removed, so the sequential
//     scan is replayed by one warp over the compacted, index-ordered list of those candidates.
//   * the slots past _particleCount keep the reference's stale content (rejected attempts, moved
//     particles): quirk 12 may read them.
//   * extrapolate: the LIFO order decides which neighbours an EMPTY cell averages (v8:528-545).  Where no
//     two interior EMPTY cells touch, the result is order-independent and runs as dependency rounds; else
//     the stack is replayed by one thread on the compacted non-fluid set.  Both are exact.
#include "flip_internal.cuh"

#include <string.h>

namespace ifl {

// ------------------------------------------------------------------------ frand ----
constexpr unsigned LCG_A = 1103515245u, LCG_C = 12345u, LCG_SEED0 = 0xBA5EBA11u;

// state after `n` calls: the affine map s -> A s + C composed n times (mod 2^32; the reference masks to
// 31 bits after every step, which commutes with the composition because 2^31 divides 2^32)
__host__ __device__ inline unsigned lcg_state_after(unsigned long long n) {
    unsigned a = 1u, c = 0u, ca = LCG_A, cc = LCG_C;
    while (n) {
        if (n & 1ull) {
            a = a * ca;
            c = c * ca + cc;
        }
        cc = cc * ca + cc;
        ca = ca * ca;
        n >>= 1;
    }
    return (a * LCG_SEED0 + c) & 0x7FFFFFFFu;
}
// frand() number d (0-based) of the process
__device__ __forceinline__ float frand_draw(unsigned long long d) {
    const unsigned s = lcg_state_after(d + 1ull);
    return __uint_as_float((s >> 8) | 0x3F800000u) - 1.0f; // (double)f - 1.0 rounded back to float is exact
}
__device__ __forceinline__ float frand_next(unsigned &state) { // sequential form, for the replay kernels
    state = (state * LCG_A + LCG_C) & 0x7FFFFFFFu;
    return __uint_as_float((state >> 8) | 0x3F800000u) - 1.0f;
}

__device__ __forceinline__ bool point_in_body(const BodyDev *bodies, int nb, double x, double y, double hx) { // v8:725-731
    for (int i = 0; i < nb; i++)
        if (body_distance(bodies[i], x * hx, y * hx) < 0.0) return true;
    return false;
}

// -------------------------------------------------------------- initParticles ----
// attempt n: cell n / avg, draws draws0 + 2n (x) and draws0 + 2n + 1 (y); `x + frand()` is FLOAT addition
__device__ __forceinline__ void attempt_position(unsigned long long first_draw, int x, int y, double &px, double &py) {
    const float fx = (float)x + frand_draw(first_draw);
    const float fy = (float)y + frand_draw(first_draw + 1ull);
    px = (double)fx;
    py = (double)fy;
}

__global__ void __launch_bounds__(256) k_init_flags(long long n_att, int avg, int W, double hx, const BodyDev *bodies, int nb,
                                                    unsigned long long draws0, int *__restrict__ flags) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_att) return;
    const long long cell = n / avg;
    double px, py;
    attempt_position(draws0 + 2ull * (unsigned long long)n, (int)(cell % W), (int)(cell / W), px, py);
    flags[n] = point_in_body(bodies, nb, px, py, hx) ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_init_write(long long n_att, int avg, int W, unsigned long long draws0,
                                                    const int *__restrict__ flags, const long long *__restrict__ offsets,
                                                    const long long *__restrict__ total, double *__restrict__ posX,
                                                    double *__restrict__ posY, long long capacity) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_att) return;
    const bool last = n == n_att - 1;
    if (!flags[n] && !last) return;
    const long long cell = n / avg;
    double px, py;
    attempt_position(draws0 + 2ull * (unsigned long long)n, (int)(cell % W), (int)(cell / W), px, py);
    // a rejected attempt is overwritten by the next one (`idx--`, v8:745); only a rejected LAST attempt survives,
    // in the first slot past the final count
    const long long slot = flags[n] ? offsets[n] : *total;
    if (slot < capacity) {
        posX[slot] = px;
        posY[slot] = py;
    }
}

// -------------------------------------------------------------- countParticles ----
__global__ void __launch_bounds__(256) k_count(const double *__restrict__ posX, const double *__restrict__ posY, long long n,
                                               int W, int H, int *__restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ix = (int)posX[i], iy = (int)posY[i];
    if (ix >= 0 && iy >= 0 && ix < W && iy < H) atomicAdd(&counts[ix + iy * W], 1); // v8:760-761
}

// -------------------------------------------------------------- pruneParticles ----
constexpr int MAX_PER_CELL = 12, MIN_PER_CELL = 3; // v8:694-696

__global__ void __launch_bounds__(256) k_prune_flags(const double *__restrict__ posX, const double *__restrict__ posY,
                                                     long long n, int W, int H, const int *__restrict__ counts,
                                                     int *__restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ix = (int)posX[i], iy = (int)posY[i];
    flags[i] = (ix >= 0 && iy >= 0 && ix < W && iy < H && counts[ix + iy * W] > MAX_PER_CELL) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_compact(const int *__restrict__ flags, const long long *__restrict__ offsets,
                                                 long long n, unsigned *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[offsets[i]] = (unsigned)i;
}

// Replay of the reference's scan (v8:767-785) over the candidates only.  One warp: every lane follows the
// control flow, lanes 0..5 move the six arrays of a particle.
__global__ void __launch_bounds__(32) k_prune_replay(double *posX, double *posY, double *p0, double *p1, double *p2, double *p3,
                                                     int *counts, int W, int H, const unsigned *__restrict__ cand,
                                                     const long long *__restrict__ ncand_p, long long count,
                                                     long long *__restrict__ out_count) {
    const int lane = threadIdx.x;
    volatile double *vx = posX, *vy = posY;
    volatile int *vc = counts;
    double *arr = lane == 0 ? posX : lane == 1 ? posY : lane == 2 ? p0 : lane == 3 ? p1 : lane == 4 ? p2 : p3;
    const long long ncand = *ncand_p;
    long long E = count;
    bool done = false;
    for (long long k = 0; k < ncand && !done; k++) {
        const long long i = cand[k];
        if (i >= E) break; // `i < _particleCount` ended the reference's loop
        for (;;) {
            const int ix = (int)vx[i], iy = (int)vy[i];
            if (ix < 0 || iy < 0 || ix >= W || iy >= H) break; // `continue`
            const int idx = ix + iy * W;
            if (vc[idx] <= MAX_PER_CELL) break;
            const long long j = --E; // v8:777
            if (lane < 6) arr[i] = arr[j];
            if (lane == 0) vc[idx] = vc[idx] - 1;
            __threadfence_block();
            __syncwarp();
            if (i >= E) { // the hole WAS the last particle: `i--`, `i++`, `i < _particleCount` fails
                done = true;
                break;
            }
        }
    }
    if (lane == 0) *out_count = E;
}

// --------------------------------------------------------------- seedParticles ----
// per cell: attempts a = max(0, 3 - count) and whether they are accepted (quirk 12: the test looks at the
// particle whose INDEX is the cell index, after pruning; it is the same for all attempts of the cell)
__global__ void __launch_bounds__(256) k_seed_plan(const double *__restrict__ posX, const double *__restrict__ posY, int ncells,
                                                   const int *__restrict__ counts, double hx, const BodyDev *bodies, int nb,
                                                   int *__restrict__ att, int *__restrict__ acc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int a = imax(MIN_PER_CELL - counts[c], 0);
    att[c] = a;
    acc[c] = (a > 0 && !point_in_body(bodies, nb, posX[c], posY[c], hx)) ? a : 0;
}

struct Fields4 {
    Field f[4]; // registration order d, t, u, v (v8:1309-1312)
};

__global__ void __launch_bounds__(256) k_seed_write(int ncells, int W, const int *__restrict__ att, const int *__restrict__ acc,
                                                    const long long *__restrict__ att_off, const long long *__restrict__ acc_off,
                                                    const long long *__restrict__ totals, unsigned long long draws0,
                                                    long long count0, long long capacity, double *__restrict__ posX,
                                                    double *__restrict__ posY, double *__restrict__ p0, double *__restrict__ p1,
                                                    double *__restrict__ p2, double *__restrict__ p3, Fields4 q) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int a = att[c];
    if (a == 0) return;
    const bool ok = acc[c] != 0;
    const long long att_total = totals[0], acc_total = totals[1];
    const int x = c % W, y = c / W;
    for (int i = 0; i < a; i++) {
        const long long n = att_off[c] + i;
        double px, py;
        attempt_position(draws0 + 2ull * (unsigned long long)n, x, y, px, py);
        if (ok) {
            const long long j = count0 + acc_off[c] + i;
            posX[j] = px;
            posY[j] = py;
            p0[j] = field_lerp(q.f[0], px, py); // v8:806-807
            p1[j] = field_lerp(q.f[1], px, py);
            p2[j] = field_lerp(q.f[2], px, py);
            p3[j] = field_lerp(q.f[3], px, py);
        } else if (n == att_total - 1) { // a rejected attempt wrote _posX[j] before the test; the last one stays
            const long long j = count0 + acc_total;
            if (j < capacity) {
                posX[j] = px;
                posY[j] = py;
            }
        }
    }
}

// exact replay for the corner cases the parallel form does not cover (the set is smaller than the grid, so
// the quirk-12 test may look at a slot that is being written; or the capacity is reached, v8:793-794)
__global__ void k_seed_replay(int W, int H, const int *__restrict__ counts, double hx, const BodyDev *bodies, int nb,
                              unsigned state, long long count, long long capacity, double *posX, double *posY, double *p0,
                              double *p1, double *p2, double *p3, Fields4 q, long long *out) {
    long long draws = 0;
    for (int y = 0, idx = 0; y < H; y++)
        for (int x = 0; x < W; x++, idx++)
            for (int i = 0; i < MIN_PER_CELL - counts[idx]; i++) {
                if (count == capacity) goto finished;
                const long long j = count;
                posX[j] = (double)((float)x + frand_next(state));
                posY[j] = (double)((float)y + frand_next(state));
                draws += 2;
                if (point_in_body(bodies, nb, posX[idx], posY[idx], hx)) continue;
                p0[j] = field_lerp(q.f[0], posX[j], posY[j]);
                p1[j] = field_lerp(q.f[1], posX[j], posY[j]);
                p2[j] = field_lerp(q.f[2], posX[j], posY[j]);
                p3[j] = field_lerp(q.f[3], posX[j], posY[j]);
                count++;
            }
finished:
    out[0] = count;
    out[1] = draws;
}

// ------------------------------------------------------------------- host side ----
static ParticleSet *pset(ifl_ctx *c) { return (ParticleSet *)c->particles; }

static int read_scalars(ifl_ctx *c, int n) { // dev_scalars[0..n) -> host_scalars (synchronises the stream)
    ParticleSet *ps = pset(c);
    IFL_CUDA(cudaMemcpyAsync(ps->host_scalars, ps->dev_scalars, n * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

static Fields4 fields4(ifl_ctx *c) {
    Fields4 q;
    q.f[0] = c->fd[IFL_FIELD_D];
    q.f[1] = c->fd[IFL_FIELD_T];
    q.f[2] = c->fd[IFL_FIELD_U];
    q.f[3] = c->fd[IFL_FIELD_V];
    return q;
}

long long flip_particle_count(const ifl_ctx *c) { return c->particles ? ((ParticleSet *)c->particles)->count : 0; }
long long flip_particle_capacity(const ifl_ctx *c) { return c->particles ? ((ParticleSet *)c->particles)->capacity : 0; }

#define TRYB(expr)                     \
    do {                               \
        int rc_ = (expr);              \
        if (rc_ != IFL_OK) return rc_; \
    } while (0)

int flip_particles_init(ifl_ctx *c, int avg_per_cell) {
    ParticleSet *ps = pset(c);
    if (avg_per_cell < 1 || avg_per_cell > MAX_PER_CELL) {
        set_error("ifl_particles_init: %d particles per cell (1..12; _AvgPerCell is 4 in the reference, v8:698)", avg_per_cell);
        return IFL_E_ARG;
    }
    ps->avg_per_cell = avg_per_cell;
    const long long n_att = (long long)c->W * c->H * avg_per_cell;
    const unsigned blocks = (unsigned)((n_att + 255) / 256);
    k_init_flags<<<blocks, 256, 0, c->stream>>>(n_att, avg_per_cell, c->W, c->hx, c->bodies_d, c->n_bodies, ps->draws, ps->flags);
    IFL_LAUNCHED(c);
    TRYB(scan_exclusive(c, ps->flags, ps->flag_offsets, (size_t)n_att, ps->dev_scalars));
    k_init_write<<<blocks, 256, 0, c->stream>>>(n_att, avg_per_cell, c->W, ps->draws, ps->flags, ps->flag_offsets, ps->dev_scalars,
                                                ps->posX, ps->posY, ps->capacity);
    IFL_LAUNCHED(c);
    TRYB(read_scalars(c, 1));
    ps->count = ps->host_scalars[0];
    ps->draws += 2ull * (unsigned long long)n_att;
    ps->binned = false;
    // the four properties start at zero (addQuantity v8:893-894), then FluidSolver's ctor interpolates the initial
    // fields onto the particles: gridToParticles(1.0)  v8:1314
    for (int t = 0; t < 4; t++) IFL_CUDA(cudaMemsetAsync(ps->prop[t], 0, (size_t)ps->capacity * sizeof(double), c->stream));
    return launch_grid_to_particles(c, 1.0);
}

int flip_count_particles(ifl_ctx *c) {
    ParticleSet *ps = pset(c);
    IFL_CUDA(cudaMemsetAsync(ps->counts, 0, (size_t)c->W * c->H * sizeof(int), c->stream));
    if (ps->count > 0) {
        k_count<<<(unsigned)((ps->count + 255) / 256), 256, 0, c->stream>>>(ps->posX, ps->posY, ps->count, c->W, c->H, ps->counts);
        IFL_LAUNCHED(c);
    }
    ps->binned = false; // counts / list now serve the bookkeeping, not the P2G bins
    return IFL_OK;
}

int flip_prune_particles(ifl_ctx *c) {
    ParticleSet *ps = pset(c);
    if (ps->count == 0) return IFL_OK;
    const unsigned blocks = (unsigned)((ps->count + 255) / 256);
    k_prune_flags<<<blocks, 256, 0, c->stream>>>(ps->posX, ps->posY, ps->count, c->W, c->H, ps->counts, ps->flags);
    IFL_LAUNCHED(c);
    TRYB(scan_exclusive(c, ps->flags, ps->flag_offsets, (size_t)ps->count, ps->dev_scalars));
    k_compact<<<blocks, 256, 0, c->stream>>>(ps->flags, ps->flag_offsets, ps->count, ps->list);
    IFL_LAUNCHED(c);
    k_prune_replay<<<1, 32, 0, c->stream>>>(ps->posX, ps->posY, ps->prop[0], ps->prop[1], ps->prop[2], ps->prop[3], ps->counts,
                                            c->W, c->H, ps->list, ps->dev_scalars, ps->count, ps->dev_scalars + 1);
    IFL_LAUNCHED(c);
    TRYB(read_scalars(c, 2));
    ps->count = ps->host_scalars[1];
    ps->binned = false;
    return IFL_OK;
}

int flip_seed_particles(ifl_ctx *c) {
    ParticleSet *ps = pset(c);
    const int ncells = c->W * c->H;
    const unsigned blocks = (unsigned)((ncells + 255) / 256);
    int *att = ps->flags, *acc = ps->flags + ncells;
    long long *att_off = ps->flag_offsets, *acc_off = ps->flag_offsets + ncells;
    bool replay = ps->count < ncells; // quirk 12 could look at a slot past the set
    if (!replay) {
        k_seed_plan<<<blocks, 256, 0, c->stream>>>(ps->posX, ps->posY, ncells, ps->counts, c->hx, c->bodies_d, c->n_bodies, att, acc);
        IFL_LAUNCHED(c);
        TRYB(scan_exclusive(c, att, att_off, (size_t)ncells, ps->dev_scalars));
        TRYB(scan_exclusive(c, acc, acc_off, (size_t)ncells, ps->dev_scalars + 1));
        TRYB(read_scalars(c, 2));
        const long long att_total = ps->host_scalars[0], acc_total = ps->host_scalars[1];
        if (ps->count + acc_total > ps->capacity) {
            replay = true; // v8:793-794 cuts the loop short somewhere inside
        } else {
            if (att_total > 0) {
                k_seed_write<<<blocks, 256, 0, c->stream>>>(ncells, c->W, att, acc, att_off, acc_off, ps->dev_scalars, ps->draws,
                                                            ps->count, ps->capacity, ps->posX, ps->posY, ps->prop[0], ps->prop[1],
                                                            ps->prop[2], ps->prop[3], fields4(c));
                IFL_LAUNCHED(c);
            }
            ps->count += acc_total;
            ps->draws += 2ull * (unsigned long long)att_total;
        }
    }
    if (replay) {
        k_seed_replay<<<1, 1, 0, c->stream>>>(c->W, c->H, ps->counts, c->hx, c->bodies_d, c->n_bodies, lcg_state_after(ps->draws),
                                              ps->count, ps->capacity, ps->posX, ps->posY, ps->prop[0], ps->prop[1], ps->prop[2],
                                              ps->prop[3], fields4(c), ps->dev_scalars);
        IFL_LAUNCHED(c);
        TRYB(read_scalars(c, 2));
        ps->count = ps->host_scalars[0];
        ps->draws += (unsigned long long)ps->host_scalars[1];
    }
    ps->binned = false;
    return IFL_OK;
}

int flip_particles_to_grid(ifl_ctx *c, long long *count) { // v8:916-927
    const int fields[4] = {IFL_FIELD_D, IFL_FIELD_T, IFL_FIELD_U, IFL_FIELD_V};
    for (int t = 0; t < 4; t++) {
        TRYB(launch_from_particles(c, fields[t]));
        TRYB(flip_extrapolate(c, fields[t]));
    }
    TRYB(flip_count_particles(c));
    TRYB(flip_prune_particles(c));
    TRYB(flip_seed_particles(c));
    if (count) *count = pset(c)->count;
    return IFL_OK;
}

int flip_peek(ifl_ctx *c, int what, long long first, long long n, void *host) {
    ParticleSet *ps = pset(c);
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    if (what == 6) { // per-cell counts
        if (first < 0 || first + n > (long long)c->W * c->H) return IFL_E_ARG;
        IFL_CUDA(cudaMemcpy(host, ps->counts + first, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
        return IFL_OK;
    }
    if (what < 0 || what > 5 || first < 0 || first + n > ps->capacity) {
        set_error("ifl_particles_peek: bad range");
        return IFL_E_ARG;
    }
    const double *src = what == 0 ? ps->posX : what == 1 ? ps->posY : ps->prop[what - 2];
    IFL_CUDA(cudaMemcpy(host, src + first, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return IFL_OK;
}

// =============================================================== extrapolate (chapter 8) ====
// fillSolidMask v8:478-513.  Also counts the interior EMPTY cells that touch another interior EMPTY cell:
// zero of them means the result does not depend on the stack order.
__global__ void __launch_bounds__(256) k8_mask(Field f, int *__restrict__ touching) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= f.w || y >= f.h) return;
    const int pitch = f.src.pitch;
    const int idx = x + y * pitch;
    if (x == 0 || y == 0 || x == f.w - 1 || y == f.h - 1) {
        f.mask[idx] = 0xFF; // the border is not touched by the stack phase
        return;
    }
    unsigned m = 0;
    const int cell = f.cell[idx];
    if (cell == CELL_SOLID) {
        const double nx = f.normalX.p[idx], ny = f.normalY.p[idx];
        if (nx != 0.0 && f.cell[idx + sgn(nx)] != CELL_FLUID) m |= 1;
        if (ny != 0.0 && f.cell[idx + sgn(ny) * pitch] != CELL_FLUID) m |= 2;
    } else if (cell == CELL_EMPTY) {
        const int l = f.cell[idx - 1], r = f.cell[idx + 1], u = f.cell[idx - pitch], d = f.cell[idx + pitch];
        m = (l != CELL_FLUID && r != CELL_FLUID && u != CELL_FLUID && d != CELL_FLUID) ? 1 : 0;
        const bool il = x > 1, ir = x < f.w - 2, iu = y > 1, id = y < f.h - 2; // is that neighbour interior?
        if ((il && l == CELL_EMPTY) || (ir && r == CELL_EMPTY) || (iu && u == CELL_EMPTY) || (id && d == CELL_EMPTY))
            atomicAdd(touching, 1);
    }
    f.mask[idx] = (uint8_t)m;
}

__device__ __forceinline__ double x86_default_nan() { return __longlong_as_double((long long)0xFFF8000000000000ull); }

__device__ __forceinline__ double ext_normal(const Field &f, int idx) { // extrapolateNormal v8:515-523
    const int pitch = f.src.pitch;
    const double nx = f.normalX.p[idx], ny = f.normalY.p[idx];
    const double srcX = f.src.p[idx + sgn(nx)];
    const double srcY = f.src.p[idx + sgn(ny) * pitch];
    return (fabs(nx) * srcX + fabs(ny) * srcY) / (fabs(nx) + fabs(ny));
}
// extrapolateAverage v8:528-545; `fluid(i)` says whether neighbour i counts as CELL_FLUID right now
template <typename IsFluid>
__device__ __forceinline__ double ext_average(const Field &f, int idx, IsFluid fluid) {
    const int pitch = f.src.pitch;
    double value = 0.0;
    int count = 0;
    if (fluid(idx - 1)) { value += f.src.p[idx - 1]; count++; }
    if (fluid(idx + 1)) { value += f.src.p[idx + 1]; count++; }
    if (fluid(idx - pitch)) { value += f.src.p[idx - pitch]; count++; }
    if (fluid(idx + pitch)) { value += f.src.p[idx + pitch]; count++; }
    if (count == 0) return x86_default_nan(); // 0.0/0 on the reference's SSE2 divide: the default (negative) quiet NaN
    return value / count;
}

// ---- order-independent case: dependency rounds over the compacted interior non-fluid cells.
// mask bits: 1, 2 as in the reference; 0x80 = solved in this round, 0x40 = final (visible to dependants).
__global__ void __launch_bounds__(256) k8_list(Field f) {
    const int x = 1 + blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = 1 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= f.w - 1 || y >= f.h - 1) return;
    const int idx = x + y * f.src.pitch;
    if (f.cell[idx] != CELL_FLUID) f.solid_list[atomicAdd(f.solid_count, 1)] = idx;
}
__global__ void __launch_bounds__(256) k8_round(Field f, int n, int *resolved) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pitch = f.src.pitch;
    const int idx = f.solid_list[i];
    const unsigned m = f.mask[idx];
    if (m & 0x80) return;
    if (f.cell[idx] == CELL_SOLID) {
        const double nx = f.normalX.p[idx], ny = f.normalY.p[idx];
        if ((m & 1) && !(f.mask[idx + sgn(nx)] & 0x40)) return;
        if ((m & 2) && !(f.mask[idx + sgn(ny) * pitch] & 0x40)) return;
        f.src.p[idx] = ext_normal(f, idx);
    } else { // CELL_EMPTY without EMPTY interior neighbours: its fluid neighbours are the cells that are fluid NOW
        if ((m & 1) && !((f.mask[idx - 1] | f.mask[idx + 1] | f.mask[idx - pitch] | f.mask[idx + pitch]) & 0x40))
            return; // no fluid neighbour: the reference pushes it when the first neighbour has been processed
        f.src.p[idx] = ext_average(f, idx, [&](int j) { return f.cell[j] == CELL_FLUID; });
    }
    f.mask[idx] = (uint8_t)(m | 0x80);
    atomicAdd(resolved, 1);
}
__global__ void __launch_bounds__(256) k8_promote(Field f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int idx = f.solid_list[i];
    const unsigned m = f.mask[idx];
    if ((m & 0xC0) == 0x80 && m != 0xFF) f.mask[idx] = (uint8_t)(m | 0x40);
}

// ---- order-dependent case: replay of the explicit stack (v8:611-648) by one thread.
__global__ void __launch_bounds__(256) k8_ready_flags(Field f, int *__restrict__ flags) { // raster order over interior cells
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= f.w || y >= f.h) return;
    const int idx = x + y * f.src.pitch;
    const bool interior = x > 0 && y > 0 && x < f.w - 1 && y < f.h - 1;
    flags[x + y * f.w] = (interior && f.cell[idx] != CELL_FLUID && f.mask[idx] == 0) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k8_ready_compact(Field f, const int *__restrict__ flags, const long long *__restrict__ offsets) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= f.w || y >= f.h) return;
    const int d = x + y * f.w;
    if (flags[d]) f.solid_list[offsets[d]] = x + y * f.src.pitch;
}
__global__ void k8_stack_replay(Field f, const long long *__restrict__ n_ready) {
    const int pitch = f.src.pitch;
    int *stack = f.solid_list;
    long long top = *n_ready;
    while (top > 0) {
        const int idx = stack[--top];
        if (f.cell[idx] == CELL_EMPTY) {
            f.src.p[idx] = ext_average(f, idx, [&](int j) { return f.cell[j] == CELL_FLUID; });
            f.cell[idx] = CELL_FLUID; // v8:630
        } else {
            f.src.p[idx] = ext_normal(f, idx);
        }
        // freeSolidNeighbour v8:547-553, in the reference's order
        const int nbr[4] = {idx - 1, idx + 1, idx - pitch, idx + pitch};
        const bool pointing[4] = {f.normalX.p[idx - 1] > 0.0, f.normalX.p[idx + 1] < 0.0, f.normalY.p[idx - pitch] > 0.0,
                                  f.normalY.p[idx + pitch] < 0.0};
        const unsigned bit[4] = {1, 1, 2, 2};
        for (int k = 0; k < 4; k++)
            if (pointing[k] && f.cell[nbr[k]] == CELL_SOLID) {
                const unsigned m = f.mask[nbr[k]] & ~bit[k];
                f.mask[nbr[k]] = (uint8_t)m;
                if (m == 0) stack[top++] = nbr[k];
            }
        // freeEmptyNeighbour v8:558-565
        for (int k = 0; k < 4; k++)
            if (f.cell[nbr[k]] == CELL_EMPTY && f.mask[nbr[k]] == 1) {
                f.mask[nbr[k]] = 0;
                stack[top++] = nbr[k];
            }
    }
}

// extrapolateEmptyBorders v8:570-609: edges, then corners, then EMPTY -> FLUID everywhere
__global__ void __launch_bounds__(256) k8_border_edges(Field f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pitch = f.src.pitch;
    if (i >= 1 && i < f.w - 1) {
        const int t = i, b = i + (f.h - 1) * pitch;
        if (f.cell[t] == CELL_EMPTY) f.src.p[t] = f.src.p[t + pitch];
        if (f.cell[b] == CELL_EMPTY) f.src.p[b] = f.src.p[b - pitch];
    }
    if (i >= 1 && i < f.h - 1) {
        const int l = i * pitch, r = i * pitch + f.w - 1;
        if (f.cell[l] == CELL_EMPTY) f.src.p[l] = f.src.p[l + 1];
        if (f.cell[r] == CELL_EMPTY) f.src.p[r] = f.src.p[r - 1];
    }
}
__global__ void k8_border_corners(Field f) {
    const int pitch = f.src.pitch;
    const int tl = 0, tr = f.w - 1, bl = (f.h - 1) * pitch, br = (f.h - 1) * pitch + f.w - 1;
    if (f.cell[tl] == CELL_EMPTY) f.src.p[tl] = 0.5 * (f.src.p[tl + 1] + f.src.p[tl + pitch]);
    if (f.cell[tr] == CELL_EMPTY) f.src.p[tr] = 0.5 * (f.src.p[tr - 1] + f.src.p[tr + pitch]);
    if (f.cell[bl] == CELL_EMPTY) f.src.p[bl] = 0.5 * (f.src.p[bl + 1] + f.src.p[bl - pitch]);
    if (f.cell[br] == CELL_EMPTY) f.src.p[br] = 0.5 * (f.src.p[br - 1] + f.src.p[br - pitch]);
}
__global__ void __launch_bounds__(256) k8_empty_to_fluid(Field f) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= f.w || y >= f.h) return;
    const int idx = x + y * f.src.pitch;
    if (f.cell[idx] == CELL_EMPTY) f.cell[idx] = CELL_FLUID;
}

int flip_extrapolate(ifl_ctx *c, int field) {
    Field &f = c->fd[field];
    ParticleSet *ps = pset(c);
    if (f.w < 3 || f.h < 3) return IFL_OK;
    ProfScope scope(c, IFL_K_ASSEMBLY);
    const dim3 grid((f.w + 63) / 64, (f.h + 3) / 4);
    int *touching = c->ext_ready;
    IFL_CUDA(cudaMemsetAsync(touching, 0, sizeof(int), c->stream));
    IFL_CUDA(cudaMemsetAsync(f.solid_count, 0, sizeof(int), c->stream));
    k8_mask<<<grid, 256, 0, c->stream>>>(f, touching);
    IFL_LAUNCHED(c);
    int h_touch = 0;
    IFL_CUDA(cudaMemcpyAsync(&h_touch, touching, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    const char *force = getenv("IFL_FLIP_EXTRAPOLATE"); // "stack" / "rounds": test hook, both forms are exact where they apply
    const bool use_stack = (h_touch > 0 && !(force && strcmp(force, "rounds") == 0)) || (force && strcmp(force, "stack") == 0);
    if (!use_stack) {
        k8_list<<<dim3((f.w - 2 + 63) / 64, (f.h - 2 + 3) / 4), 256, 0, c->stream>>>(f);
        IFL_LAUNCHED(c);
        int n = 0;
        IFL_CUDA(cudaMemcpyAsync(&n, f.solid_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IFL_CUDA(cudaStreamSynchronize(c->stream));
        const int blocks = (n + 255) / 256, batch = 16;
        for (int guard = 0; n > 0 && guard < (f.w + f.h); guard += batch) {
            IFL_CUDA(cudaMemsetAsync(c->ext_ready, 0, sizeof(int), c->stream));
            for (int r = 0; r < batch; r++) {
                k8_round<<<blocks, 256, 0, c->stream>>>(f, n, c->ext_ready);
                IFL_LAUNCHED(c);
                k8_promote<<<blocks, 256, 0, c->stream>>>(f, n);
                IFL_LAUNCHED(c);
            }
            int resolved = 0;
            IFL_CUDA(cudaMemcpyAsync(&resolved, c->ext_ready, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            IFL_CUDA(cudaStreamSynchronize(c->stream));
            if (resolved == 0) break;
        }
    } else {
        if ((size_t)f.w * f.h > ps->flags_elems) {
            set_error("flip_extrapolate: field larger than the flag scratch");
            return IFL_E_ARG;
        }
        k8_ready_flags<<<grid, 256, 0, c->stream>>>(f, ps->flags);
        IFL_LAUNCHED(c);
        TRYB(scan_exclusive(c, ps->flags, ps->flag_offsets, (size_t)f.w * f.h, ps->dev_scalars));
        k8_ready_compact<<<grid, 256, 0, c->stream>>>(f, ps->flags, ps->flag_offsets);
        IFL_LAUNCHED(c);
        k8_stack_replay<<<1, 1, 0, c->stream>>>(f, ps->dev_scalars);
        IFL_LAUNCHED(c);
    }
    k8_border_edges<<<(imax(f.w, f.h) + 255) / 256, 256, 0, c->stream>>>(f);
    IFL_LAUNCHED(c);
    k8_border_corners<<<1, 1, 0, c->stream>>>(f);
    IFL_LAUNCHED(c);
    k8_empty_to_fluid<<<grid, 256, 0, c->stream>>>(f);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

} // namespace ifl
