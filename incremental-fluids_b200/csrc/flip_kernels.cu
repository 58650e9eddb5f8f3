// flip_kernels.cu -- chapter 8 (FLIP) particle <-> grid transfers.
//
//   from_particles (P2G)   FluidQuantity::fromParticles v8:663-688 + addSample v8:296-303
//   grid_to_particles      ParticleQuantities::gridToParticles v8:904-911
//   copy / diff / undiff   FluidQuantity::copy/diff/undiff v8:374-388
//   particles_advect       ParticleQuantities::advect v8:931-939 (rungeKutta3 v8:844-862,
//                          backProject v8:816-839)
//
// P2G parity.  The reference scatters particle i = 0..count-1 IN INDEX ORDER into the four
// surrounding nodes (`weight += k; src += k*value`), so every node's two sums are
// order-sensitive floating-point accumulations.  Atomics would make that order random.
// Instead the transfer is a gather:
//   1. bin particle indices by base cell (int)pos once per particle set (counting sort:
//      histogram -> CUB exclusive scan -> scatter -> per-cell insertion sort by index);
//   2. one thread per grid node walks the (at most 9) bins that can hold contributors,
//      always taking the smallest pending particle index, and accumulates exactly the
//      reference's two statements.  Result: bit-identical `src`, `weight` and cell flags.
// Particle SoA (posX, posY, one array per quantity) is kept as in the reference (v8:717-723);
// all particle kernels are coalesced one-thread-per-particle streams.
#include "flip_internal.cuh"

namespace ifl {

// ------------------------------------------------------------------ prefix sums ----
// Block-wise scan (1024 items per block), recursive scan of the block sums, add-back.  Plain kernels of
// this library: no CUB on the path.
constexpr int SCAN_ITEMS = 1024;
template <typename T>
__global__ void __launch_bounds__(256) k_scan_blocks(const T *__restrict__ in, long long *__restrict__ out, size_t n,
                                                     long long *__restrict__ block_sums) {
    __shared__ long long warp_tot[8];
    const size_t base = (size_t)blockIdx.x * SCAN_ITEMS + (size_t)threadIdx.x * 4;
    long long v[4], run = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = base + k < n ? (long long)in[base + k] : 0;
        run += v[k];
    }
    // inclusive scan of the per-thread totals: warp shuffle, then the 8 warp totals
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    long long before = 0;
    for (int w = 0; w < wid; w++) before += warp_tot[w];
    long long excl = before + incl - run;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = before + incl;
}
__global__ void __launch_bounds__(256) k_scan_add(long long *__restrict__ out, size_t n, const long long *__restrict__ block_offsets) {
    const size_t i = (size_t)blockIdx.x * SCAN_ITEMS + threadIdx.x;
    const long long add = block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t j = i + (size_t)k * 256;
        if (j < n) out[j] += add;
    }
}
// one block scans up to a few million block sums in place (exclusive) and stores the grand total
__global__ void __launch_bounds__(1024) k_scan_small(long long *__restrict__ a, size_t n, long long *__restrict__ total) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (size_t base = 0; base < n; base += 1024) {
        const size_t i = base + threadIdx.x;
        const long long v = i < n ? a[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        long long before = carry_s;
        for (int w = 0; w < wid; w++) before += warp_tot[w];
        if (i < n) a[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

int scan_exclusive(ifl_ctx *c, const int *in, long long *out, size_t n, long long *total_dev) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (n == 0) {
        if (total_dev) IFL_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(long long), c->stream));
        return IFL_OK;
    }
    const size_t nblocks = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    if (nblocks > ps->scan_tmp_elems) {
        set_error("scan_exclusive: %zu items exceed the scratch sized at flip_init", n);
        return IFL_E_ARG;
    }
    k_scan_blocks<int><<<(unsigned)nblocks, 256, 0, c->stream>>>(in, out, n, ps->scan_tmp);
    IFL_LAUNCHED(c);
    k_scan_small<<<1, 1024, 0, c->stream>>>(ps->scan_tmp, nblocks, total_dev);
    IFL_LAUNCHED(c);
    k_scan_add<<<(unsigned)nblocks, 256, 0, c->stream>>>(out, n, ps->scan_tmp);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

// ------------------------------------------------------------------- binning ----
__global__ void __launch_bounds__(256) k_bin_count(const double *__restrict__ posX, const double *__restrict__ posY,
                                                   long long n, int w, int h, int *__restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = imin(imax((int)posX[i], 0), w - 1), cy = imin(imax((int)posY[i], 0), h - 1);
    atomicAdd(&counts[cx + cy * w], 1);
}

__global__ void __launch_bounds__(256) k_bin_scatter(const double *__restrict__ posX, const double *__restrict__ posY,
                                                     long long n, int w, int h, const long long *__restrict__ offsets,
                                                     int *__restrict__ fill, unsigned *__restrict__ list) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = imin(imax((int)posX[i], 0), w - 1), cy = imin(imax((int)posY[i], 0), h - 1);
    const int c = cx + cy * w;
    list[offsets[c] + atomicAdd(&fill[c], 1)] = (unsigned)i;
}

// each bin ascending by particle index (bins hold a handful of particles: v8:694-698)
__global__ void __launch_bounds__(256) k_bin_sort(const long long *__restrict__ offsets, const int *__restrict__ counts,
                                                  int ncells, unsigned *__restrict__ list) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    unsigned *a = list + offsets[c];
    const int m = counts[c];
    for (int i = 1; i < m; i++) {
        const unsigned v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = v;
    }
}

// ----------------------------------------------------------------------- P2G ----
__global__ void __launch_bounds__(128) k_from_particles(Field f, double *__restrict__ weight, int weight_pitch,
                                                        const double *__restrict__ posX, const double *__restrict__ posY,
                                                        const double *__restrict__ prop, const long long *__restrict__ offsets,
                                                        const int *__restrict__ counts, const unsigned *__restrict__ list,
                                                        int W, int H) {
    const int gx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int gy = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (gx >= f.w || gy >= f.h) return;
    // bins that can hold contributors of node (gx, gy)
    const int bx0 = imax(gx - 1, 0), bx1 = imin(gx + 1, W - 1);
    const int by0 = imax(gy - 1, 0), by1 = imin(gy + 1, H - 1);
    long long cur[9], end[9];
    int nb = 0;
    for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
            const int c = bx + by * W;
            cur[nb] = offsets[c];
            end[nb] = offsets[c] + counts[c];
            nb++;
        }
    const double xmax = f.w - 1.5, ymax = f.h - 1.5;
    double wsum = 0.0, vsum = 0.0; // memset(_src), memset(weight)  v8:664-665
    for (;;) {
        int best = -1;
        unsigned bi = 0xffffffffu;
        for (int b = 0; b < nb; b++)
            if (cur[b] < end[b]) {
                const unsigned pi = list[cur[b]];
                if (best < 0 || pi < bi) {
                    bi = pi;
                    best = b;
                }
            }
        if (best < 0) break;
        cur[best]++;
        // v8:668-674
        double x = posX[bi] - f.ox;
        double y = posY[bi] - f.oy;
        x = std_max(0.5, std_min(xmax, x));
        y = std_max(0.5, std_min(ymax, y));
        const int ix = (int)x, iy = (int)y;
        if ((gx == ix || gx == ix + 1) && (gy == iy || gy == iy + 1)) {
            const double k = (1.0 - fabs(gx - x)) * (1.0 - fabs(gy - y)); // addSample v8:300
            wsum += k;
            vsum += k * prop[bi];
        }
    }
    const size_t idx = gx + (size_t)gy * f.src.pitch;
    weight[gx + (size_t)gy * weight_pitch] = wsum;
    if (wsum != 0.0) { // v8:682-687
        f.src.p[idx] = vsum / wsum;
    } else {
        f.src.p[idx] = vsum;
        if (f.cell[idx] == CELL_FLUID) f.cell[idx] = CELL_EMPTY;
    }
}

// ----------------------------------------------------------------------- G2P ----
// prop = prop*(1-alpha) + lerp(field)   v8:907-908, one launch per quantity (registration order)
__global__ void __launch_bounds__(256) k_grid_to_particles(Field f, double *__restrict__ prop,
                                                           const double *__restrict__ posX,
                                                           const double *__restrict__ posY, long long n, double alpha) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = prop[i];
    v *= 1.0 - alpha;
    v += field_lerp(f, posX[i], posY[i]);
    prop[i] = v;
}

// src -= (1-alpha)*old   /   src += (1-alpha)*old      v8:379-388
__global__ void __launch_bounds__(256) k_diff(Arr src, Arr old, double one_minus_alpha, int undo) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= src.w || y >= src.h) return;
    const size_t i = x + (size_t)y * src.pitch;
    if (undo)
        src.p[i] += one_minus_alpha * old.p[i];
    else
        src.p[i] -= one_minus_alpha * old.p[i];
}

// ----------------------------------------------------------- particle advection ----
__global__ void __launch_bounds__(256) k_particles_advect(double *__restrict__ posX, double *__restrict__ posY, long long n,
                                                          Field u, Field v, double timestep, double hx, int W, int H,
                                                          const BodyDev *__restrict__ bodies, int nb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = posX[i], y = posY[i];
    // rungeKutta3, forward in time  v8:844-862 (third stage not divided by hx)
    const double firstU = field_lerp(u, x, y) / hx;
    const double firstV = field_lerp(v, x, y) / hx;
    const double midX = x + 0.5 * timestep * firstU;
    const double midY = y + 0.5 * timestep * firstV;
    const double midU = field_lerp(u, midX, midY) / hx;
    const double midV = field_lerp(v, midX, midY) / hx;
    const double lastX = x + 0.75 * timestep * midU;
    const double lastY = y + 0.75 * timestep * midV;
    const double lastU = field_lerp(u, lastX, lastY);
    const double lastV = field_lerp(v, lastX, lastY);
    x += timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
    y += timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
    // backProject v8:816-839: only when more than one WORLD unit inside a body
    double d = 1e30;
    int closest = -1;
    for (int b = 0; b < nb; b++) {
        const double id = body_distance(bodies[b], x * hx, y * hx);
        if (id < d) {
            d = id;
            closest = b;
        }
    }
    if (d < -1.0) {
        x *= hx;
        y *= hx;
        body_closest_surface_point(bodies[closest], x, y);
        double nx, ny;
        body_distance_normal(bodies[closest], nx, ny, x, y);
        x -= nx * hx;
        y -= ny * hx;
        x /= hx;
        y /= hx;
    }
    posX[i] = std_max(std_min(x, W - 0.001), 0.0); // v8:936-937
    posY[i] = std_max(std_min(y, H - 0.001), 0.0);
}

// -------------------------------------------------------------------- host side ----
int flip_init(ifl_ctx *c) {
    ParticleSet *ps = (ParticleSet *)calloc(1, sizeof(ParticleSet));
    if (!ps) return IFL_E_NOMEM;
    c->particles = ps;
    ps->capacity = (long long)c->W * c->H * 12; // _MaxPerCell v8:694, 869 (64-bit: SURVEY quirk 14)
    ps->avg_per_cell = 4;                       // _AvgPerCell v8:698
    ps->draws = 0;
    if (ps->capacity > 0xffffffffLL) {
        set_error("ifl_create: %lld particle slots (w*h*12) exceed the 2^32 - 1 one rank can index; shard the grid over more GPUs",
                  ps->capacity);
        return IFL_E_ARG;
    }
    const size_t nb = (size_t)ps->capacity * sizeof(double);
    IFL_CUDA(cudaMalloc(&ps->posX, nb));
    IFL_CUDA(cudaMalloc(&ps->posY, nb));
    IFL_CUDA(cudaMemset(ps->posX, 0, nb)); // (`new double[]` of this size is a fresh zero mapping in the reference)
    IFL_CUDA(cudaMemset(ps->posY, 0, nb));
    for (int t = 0; t < 4; t++) {
        IFL_CUDA(cudaMalloc(&ps->prop[t], nb));
        IFL_CUDA(cudaMemset(ps->prop[t], 0, nb)); // addQuantity v8:893-894
    }
    const size_t ncells = (size_t)c->W * c->H;
    IFL_CUDA(cudaMalloc(&ps->counts, ncells * sizeof(int)));
    IFL_CUDA(cudaMalloc(&ps->offsets, ncells * sizeof(long long)));
    IFL_CUDA(cudaMalloc(&ps->fill, ncells * sizeof(int)));
    IFL_CUDA(cudaMalloc(&ps->list, (size_t)ps->capacity * sizeof(unsigned)));
    // flags / offsets of the bookkeeping passes: one per particle slot (>= one per init attempt, one per cell)
    ps->flags_elems = (size_t)ps->capacity;
    IFL_CUDA(cudaMalloc(&ps->flags, ps->flags_elems * sizeof(int)));
    IFL_CUDA(cudaMalloc(&ps->flag_offsets, ps->flags_elems * sizeof(long long)));
    ps->scan_tmp_elems = ps->flags_elems / SCAN_ITEMS + 2;
    IFL_CUDA(cudaMalloc(&ps->scan_tmp, ps->scan_tmp_elems * sizeof(long long)));
    IFL_CUDA(cudaMalloc(&ps->dev_scalars, 8 * sizeof(long long)));
    IFL_CUDA(cudaMallocHost(&ps->host_scalars, 8 * sizeof(long long)));
    ps->weight.w = c->W + 1;
    ps->weight.h = c->H + 1;
    ps->weight.pitch = (c->W + 1 + 31) / 32 * 32;
    ps->weight.rows = c->H + 1;
    IFL_CUDA(cudaMalloc(&ps->weight.p, ps->weight.bytes()));
    IFL_CUDA(cudaMemset(ps->weight.p, 0, ps->weight.bytes()));
    return IFL_OK;
}

void flip_free(ifl_ctx *c) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (!ps) return;
    void *ptrs[] = {ps->posX,   ps->posY,    ps->prop[0], ps->prop[1], ps->prop[2],  ps->prop[3],  ps->counts,
                    ps->offsets, ps->fill,   ps->list,    ps->scan_tmp, ps->weight.p, ps->flags,   ps->flag_offsets,
                    ps->dev_scalars};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (ps->host_scalars) cudaFreeHost(ps->host_scalars);
    free(ps);
    c->particles = nullptr;
}

int ensure_bins(ifl_ctx *c) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (ps->binned) return IFL_OK;
    const int ncells = c->W * c->H;
    const long long n = ps->count;
    IFL_CUDA(cudaMemsetAsync(ps->counts, 0, (size_t)ncells * sizeof(int), c->stream));
    IFL_CUDA(cudaMemsetAsync(ps->fill, 0, (size_t)ncells * sizeof(int), c->stream));
    if (n > 0) {
        k_bin_count<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(ps->posX, ps->posY, n, c->W, c->H, ps->counts);
        IFL_LAUNCHED(c);
    }
    int rc = scan_exclusive(c, ps->counts, ps->offsets, (size_t)ncells, nullptr);
    if (rc != IFL_OK) return rc;
    if (n > 0) {
        k_bin_scatter<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(ps->posX, ps->posY, n, c->W, c->H, ps->offsets, ps->fill,
                                                                         ps->list);
        IFL_LAUNCHED(c);
        k_bin_sort<<<(ncells + 255) / 256, 256, 0, c->stream>>>(ps->offsets, ps->counts, ncells, ps->list);
        IFL_LAUNCHED(c);
    }
    ps->binned = true;
    return IFL_OK;
}

int flip_set_particles(ifl_ctx *c, long long count, const double *posX, const double *posY, const double *const *props) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (count < 0 || count > ps->capacity) {
        set_error("particle count %lld exceeds the capacity w*h*12 = %lld (v8:869)", count, ps->capacity);
        return IFL_E_ARG;
    }
    const size_t nb = (size_t)count * sizeof(double);
    IFL_CUDA(cudaMemcpyAsync(ps->posX, posX, nb, cudaMemcpyHostToDevice, c->stream));
    IFL_CUDA(cudaMemcpyAsync(ps->posY, posY, nb, cudaMemcpyHostToDevice, c->stream));
    for (int t = 0; t < 4; t++)
        if (props && props[t]) IFL_CUDA(cudaMemcpyAsync(ps->prop[t], props[t], nb, cudaMemcpyHostToDevice, c->stream));
    // the slots past the uploaded set read as zero, like a fresh reference allocation (seedParticles may look
    // at them)
    const size_t tail = (size_t)(ps->capacity - count) * sizeof(double);
    if (tail) {
        IFL_CUDA(cudaMemsetAsync(ps->posX + count, 0, tail, c->stream));
        IFL_CUDA(cudaMemsetAsync(ps->posY + count, 0, tail, c->stream));
        for (int t = 0; t < 4; t++) IFL_CUDA(cudaMemsetAsync(ps->prop[t] + count, 0, tail, c->stream));
    }
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    ps->count = count;
    ps->binned = false;
    return IFL_OK;
}

int flip_get_particles(ifl_ctx *c, long long *count, double *posX, double *posY, double *const *props) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    const size_t nb = (size_t)ps->count * sizeof(double);
    if (count) *count = ps->count;
    if (posX) IFL_CUDA(cudaMemcpyAsync(posX, ps->posX, nb, cudaMemcpyDeviceToHost, c->stream));
    if (posY) IFL_CUDA(cudaMemcpyAsync(posY, ps->posY, nb, cudaMemcpyDeviceToHost, c->stream));
    for (int t = 0; t < 4; t++)
        if (props && props[t]) IFL_CUDA(cudaMemcpyAsync(props[t], ps->prop[t], nb, cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

static int quantity_slot(int field) { // registration order d, t, u, v  (v8:1309-1312)
    switch (field) {
    case IFL_FIELD_D: return 0;
    case IFL_FIELD_T: return 1;
    case IFL_FIELD_U: return 2;
    case IFL_FIELD_V: return 3;
    }
    return -1;
}

int launch_from_particles(ifl_ctx *c, int field) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    int rc = ensure_bins(c);
    if (rc != IFL_OK) return rc;
    Field &f = c->fd[field];
    ProfScope scope(c, IFL_K_P2G);
    k_from_particles<<<dim3((f.w + 31) / 32, (f.h + 3) / 4), 128, 0, c->stream>>>(
        f, ps->weight.p, ps->weight.pitch, ps->posX, ps->posY, ps->prop[quantity_slot(field)], ps->offsets, ps->counts,
        ps->list, c->W, c->H);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_grid_to_particles(ifl_ctx *c, double alpha) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (ps->count == 0) return IFL_OK;
    const int fields[4] = {IFL_FIELD_D, IFL_FIELD_T, IFL_FIELD_U, IFL_FIELD_V};
    ProfScope scope(c, IFL_K_G2P);
    for (int t = 0; t < 4; t++) {
        k_grid_to_particles<<<(unsigned)((ps->count + 255) / 256), 256, 0, c->stream>>>(c->fd[fields[t]], ps->prop[t], ps->posX,
                                                                            ps->posY, ps->count, alpha);
        IFL_LAUNCHED(c);
    }
    return IFL_OK;
}

int launch_copy(ifl_ctx *c, int field) { // FluidQuantity::copy v8:374-376 (_old lives in the dst slot)
    Field &f = c->fd[field];
    IFL_CUDA(cudaMemcpyAsync(f.dst.p, f.src.p, f.src.bytes(), cudaMemcpyDeviceToDevice, c->stream));
    return IFL_OK;
}

int launch_diff(ifl_ctx *c, int field, double alpha, int undo) {
    Field &f = c->fd[field];
    ProfScope scope(c, IFL_K_G2P);
    k_diff<<<dim3((f.w + 63) / 64, (f.h + 3) / 4), 256, 0, c->stream>>>(f.src, f.dst, 1.0 - alpha, undo);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_particles_advect(ifl_ctx *c, double timestep) {
    ParticleSet *ps = (ParticleSet *)c->particles;
    if (ps->count == 0) return IFL_OK;
    ProfScope scope(c, IFL_K_ADVECT);
    k_particles_advect<<<(unsigned)((ps->count + 255) / 256), 256, 0, c->stream>>>(ps->posX, ps->posY, ps->count, c->fd[IFL_FIELD_U],
                                                                       c->fd[IFL_FIELD_V], timestep, c->hx, c->W, c->H,
                                                                       c->bodies_d, c->n_bodies);
    IFL_LAUNCHED(c);
    ps->binned = false;
    return IFL_OK;
}

int flip_weight_download(ifl_ctx *c, double *host) { // dense (w+1) x (h+1), like _weight
    ParticleSet *ps = (ParticleSet *)c->particles;
    IFL_CUDA(cudaMemcpy2DAsync(host, (size_t)ps->weight.w * 8, ps->weight.p, (size_t)ps->weight.pitch * 8,
                               (size_t)ps->weight.w * 8, ps->weight.h, cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    return IFL_OK;
}

} // namespace ifl
