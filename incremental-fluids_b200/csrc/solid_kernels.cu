// solid_kernels.cu -- chapters 4+ (solid bodies): the glue between the hot-path stages
// that keeps a step device-resident.
//
//   body geometry          SolidBox / SolidSphere            v4:152-241 (v5:198-286)
//   fill_solid_fields      FluidQuantity::fillSolidFields    v5:506-559 (v4:449-480)
//   occupancy              marching-squares cell volume      v5:65-131
//   set_boundary_condition FluidSolver::setBoundaryCondition v4:812-833
//   extrapolate            FluidQuantity::extrapolate        v4:551-587 (+ fillSolidMask v4:508-529)
//
// Bit parity notes
//  * rotate() (v4:58-62) calls libm cos/sin; the host evaluates them once per body and
//    update (glibc cos is even and sin odd, exactly), the device only multiplies.
//  * setBoundaryCondition is a raster-order scatter with overlapping writes; the gather
//    below lets the LAST raster writer win (the cell to the right / below).
//  * extrapolate() is an explicit-stack fast-marching fill whose result does not depend
//    on the pop order (every cell is solved once, from neighbours that are final), so
//    it runs here as rounds over the compacted list of interior non-fluid cells.
#include "ifl_internal.cuh"
#include "solid_geometry.cuh"

namespace ifl {

// ----------------------------------------------------------------- occupancy ----
__device__ __forceinline__ double tri_occ(double out1, double in, double out2) { // v5:65-67
    return 0.5 * in * in / ((out1 - in) * (out2 - in));
}
__device__ __forceinline__ double trap_occ(double out1, double out2, double in1, double in2) { // v5:73-75
    return 0.5 * (-in1 / (out1 - in1) - in2 / (out2 - in2));
}
__device__ double occupancy(double d11, double d12, double d21, double d22) { // v5:95-131
    const double ds[4] = {d11, d12, d22, d21};
    unsigned b = 0;
    for (int i = 3; i >= 0; i--) b = (b << 1) | (ds[i] < 0.0 ? 1u : 0u);
    switch (b) {
    case 0x0: return 0.0;
    case 0x1: return tri_occ(d21, d11, d12);
    case 0x2: return tri_occ(d11, d12, d22);
    case 0x4: return tri_occ(d12, d22, d21);
    case 0x8: return tri_occ(d22, d21, d11);
    case 0xE: return 1.0 - tri_occ(-d21, -d11, -d12);
    case 0xD: return 1.0 - tri_occ(-d11, -d12, -d22);
    case 0xB: return 1.0 - tri_occ(-d12, -d22, -d21);
    case 0x7: return 1.0 - tri_occ(-d22, -d21, -d11);
    case 0x3: return trap_occ(d21, d22, d11, d12);
    case 0x6: return trap_occ(d11, d21, d12, d22);
    case 0x9: return trap_occ(d12, d22, d11, d21);
    case 0xC: return trap_occ(d11, d12, d21, d22);
    case 0x5: return tri_occ(d11, d12, d22) + tri_occ(d22, d21, d11);
    case 0xA: return tri_occ(d21, d11, d12) + tri_occ(d12, d22, d21);
    case 0xF: return 1.0;
    }
    return 0.0;
}

// ----------------------------------------------------------- fillSolidFields ----
// distance field on the (w+1) x (h+1) corner grid, v5:511-520
__global__ void __launch_bounds__(256) k_fill_phi(Arr phi, double ox, double oy, double hx, const BodyDev *bodies, int nb) {
    const int ix = blockIdx.x * 64 + (threadIdx.x & 63);
    const int iy = phi.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ix >= phi.w || iy >= phi.ry1) return;
    const double x = (ix + ox - 0.5) * hx;
    const double y = (iy + oy - 0.5) * hx;
    double d = body_distance(bodies[0], x, y);
    for (int i = 1; i < nb; i++) d = std_min(d, body_distance(bodies[i], x, y));
    phi.p[ix + (size_t)iy * phi.pitch] = d;
}

// per-cell body, volume, normal, cell type: v5:522-558 (curved) / v4:456-479 (binary)
__global__ void __launch_bounds__(256) k_fill_cells(Field f, double hx, const BodyDev *bodies, int nb, int curved,
                                                    double *fmask, int fmask_pitch) {
    const int ix = blockIdx.x * 64 + (threadIdx.x & 63);
    const int iy = f.src.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ix >= f.w || iy >= f.src.ry1) return;
    const double x = (ix + f.ox) * hx;
    const double y = (iy + f.oy) * hx;
    int body = 0;
    double d = body_distance(bodies[0], x, y);
    for (int i = 1; i < nb; i++) {
        const double id = body_distance(bodies[i], x, y);
        if (id < d) {
            body = i;
            d = id;
        }
    }
    const size_t idx = ix + (size_t)iy * f.src.pitch;
    f.body[idx] = (uint8_t)body;
    int cell;
    if (curved) {
        const size_t ip = ix + (size_t)iy * f.phi.pitch;
        double vol = 1.0 - occupancy(f.phi.p[ip], f.phi.p[ip + 1], f.phi.p[ip + f.phi.pitch], f.phi.p[ip + f.phi.pitch + 1]);
        if (vol < 0.01) vol = 0.0;
        f.volume.p[idx] = vol;
        cell = (vol == 0.0) ? CELL_SOLID : CELL_FLUID;
    } else {
        cell = (d < 0.0) ? CELL_SOLID : CELL_FLUID;
    }
    double nx, ny;
    body_distance_normal(bodies[body], nx, ny, x, y);
    f.normalX.p[idx] = nx;
    f.normalY.p[idx] = ny;
    f.cell[idx] = (uint8_t)cell;
    if (fmask) fmask[ix + (size_t)iy * fmask_pitch] = (cell == CELL_FLUID) ? 1.0 : 0.0;
}

int launch_fill_solid_fields(ifl_ctx *c, int field) {
    if (c->n_bodies == 0) return IFL_OK; // v5:507-508
    Field &f = c->fd[field];
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const int curved = c->version >= 5;
    if (curved) {
        dim3 g((f.phi.w + 63) / 64, (f.phi.ry1 - f.phi.ry0 + 3) / 4);
        k_fill_phi<<<g, 256, 0, c->stream>>>(f.phi, f.ox, f.oy, c->hx, c->bodies_d, c->n_bodies);
        IFL_LAUNCHED(c);
        int rc = dist_barrier(c); // a cell reads the corner row below it, which may be the next slab's
        if (rc != IFL_OK) return rc;
    }
    dim3 g((f.w + 63) / 64, (f.src.ry1 - f.src.ry0 + 3) / 4);
    const bool is_d = field == IFL_FIELD_D;
    k_fill_cells<<<g, 256, 0, c->stream>>>(f, c->hx, c->bodies_d, c->n_bodies, curved, is_d ? c->fmask.p : nullptr,
                                           c->fmask.pitch);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

// ------------------------------------------------------ setBoundaryCondition ----
// v4:812-833.  u face (x,y) is written by solid cell (x-1,y) [as its x+1 face] and then by
// solid cell (x,y) [as its x face]; both evaluate velocityX at the same point, the later
// raster cell wins.  Domain wall faces are zeroed afterwards.
__global__ void __launch_bounds__(256) k_set_bc_u(Field u, Field d, double hx, const BodyDev *bodies, int nb) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = u.src.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int W = d.w;
    if (x > W || y >= u.src.ry1) return;
    const size_t iu = x + (size_t)y * u.src.pitch;
    if (x == 0 || x == W) {
        u.src.p[iu] = 0.0;
        return;
    }
    if (nb == 0) return;
    const size_t ic = x + (size_t)y * d.src.pitch;
    int body = -1;
    if (d.cell[ic - 1] == CELL_SOLID) body = d.body[ic - 1];
    if (d.cell[ic] == CELL_SOLID) body = d.body[ic];
    if (body >= 0) u.src.p[iu] = body_velocity_x(bodies[body], x * hx, (y + 0.5) * hx);
}

__global__ void __launch_bounds__(256) k_set_bc_v(Field v, Field d, double hx, const BodyDev *bodies, int nb) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = v.src.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = d.h;
    if (x >= v.w || y >= v.src.ry1) return;
    const size_t iv = x + (size_t)y * v.src.pitch;
    if (y == 0 || y == H) {
        v.src.p[iv] = 0.0;
        return;
    }
    if (nb == 0) return;
    const size_t ic = x + (size_t)y * d.src.pitch;
    int body = -1;
    if (d.cell[ic - d.src.pitch] == CELL_SOLID) body = d.body[ic - d.src.pitch];
    if (d.cell[ic] == CELL_SOLID) body = d.body[ic];
    if (body >= 0) v.src.p[iv] = body_velocity_y(bodies[body], (x + 0.5) * hx, y * hx);
}

int launch_set_boundary_condition(ifl_ctx *c) {
    Field &u = c->fd[IFL_FIELD_U], &v = c->fd[IFL_FIELD_V], &d = c->fd[IFL_FIELD_D];
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    k_set_bc_u<<<dim3((u.w + 63) / 64, (u.src.ry1 - u.src.ry0 + 3) / 4), 256, 0, c->stream>>>(u, d, c->hx, c->bodies_d,
                                                                                            c->n_bodies);
    IFL_LAUNCHED(c);
    k_set_bc_v<<<dim3((v.w + 63) / 64, (v.src.ry1 - v.src.ry0 + 3) / 4), 256, 0, c->stream>>>(v, d, c->hx, c->bodies_d,
                                                                                            c->n_bodies);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

// ----------------------------------------------------------------- extrapolate ----
// fillSolidMask (v4:508-529) over interior cells + compaction of the interior non-fluid
// cells into solid_list.  Bit 0x80 of the mask marks "already solved".
__global__ void __launch_bounds__(256) k_ext_mask(Field f) {
    const int x = 1 + blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = imax(1, f.src.ry0) + blockIdx.y * 4 + (threadIdx.x >> 6); // interior rows of this rank's slab
    if (x >= f.w - 1 || y >= imin(f.h - 1, f.src.ry1)) return;
    const int pitch = f.src.pitch;
    const int idx = x + y * pitch;
    if (f.cell[idx] == CELL_FLUID) return;
    const double nx = f.normalX.p[idx], ny = f.normalY.p[idx];
    unsigned m = 0;
    if (nx != 0.0 && f.cell[idx + sgn(nx)] != CELL_FLUID) m |= 1;
    if (ny != 0.0 && f.cell[idx + sgn(ny) * pitch] != CELL_FLUID) m |= 2;
    f.mask[idx] = (uint8_t)m;
    f.solid_list[atomicAdd(f.solid_count, 1)] = idx;
}

// One round: every listed cell that is ready (both upstream neighbours final) and not
// yet solved gets its value (extrapolateNormal v4:531-540).  A cell's upstream neighbours
// are final when they are fluid, or non-fluid and marked solved in an EARLIER round
// (kernel boundaries order the rounds).  Cells outside the interior are never solved,
// exactly like in the reference, so their dependants wait forever.
__global__ void __launch_bounds__(256) k_ext_round(Field f, int n, int *resolved) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pitch = f.src.pitch;
    const int idx = f.solid_list[i];
    const unsigned m = f.mask[idx];
    if (m & 0x80) return; // solved
    const double nx = f.normalX.p[idx], ny = f.normalY.p[idx];
    const int ix = idx + sgn(nx), iy = idx + sgn(ny) * pitch;
    // upstream x neighbour must be final if the mask says we wait for it
    if ((m & 1) && !(f.mask[ix] & 0x40)) return;
    if ((m & 2) && !(f.mask[iy] & 0x40)) return;
    const double srcX = f.src.p[ix], srcY = f.src.p[iy];
    f.src.p[idx] = (fabs(nx) * srcX + fabs(ny) * srcY) / (fabs(nx) + fabs(ny));
    f.mask[idx] = (uint8_t)(m | 0x80); // solved now; becomes visible as "final" (0x40) next round
    atomicAdd(resolved, 1);
}

// promote "solved this round" (0x80) to "final" (0x40 | 0x80) between rounds
__global__ void __launch_bounds__(256) k_ext_promote(Field f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int idx = f.solid_list[i];
    const unsigned m = f.mask[idx];
    if ((m & 0xC0) == 0x80) f.mask[idx] = (uint8_t)(m | 0x40);
}

// Several ranks: every rank lists the interior non-fluid cells of its own slab; a round reads
// mask / value of cells one row into the neighbouring slabs, so the ranks meet at a barrier
// after every round and after every promotion, and the stop decision uses the number of
// cells resolved by ALL ranks (the same dependency rounds as on one GPU, hence the same values).
int launch_extrapolate(ifl_ctx *c, int field) {
    Field &f = c->fd[field];
    if (f.w < 3 || f.h < 3) return IFL_OK;
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    IFL_CUDA(cudaMemsetAsync(f.solid_count, 0, sizeof(int), c->stream));
    const int y_lo = imax(1, f.src.ry0), y_hi = imin(f.h - 1, f.src.ry1);
    if (y_hi > y_lo) {
        k_ext_mask<<<dim3((f.w - 2 + 63) / 64, (y_hi - y_lo + 3) / 4), 256, 0, c->stream>>>(f);
        IFL_LAUNCHED(c);
    }
    int n = 0;
    IFL_CUDA(cudaMemcpyAsync(&n, f.solid_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IFL_CUDA(cudaStreamSynchronize(c->stream));
    long long n_all = n;
    int rc = dist_host_sum(c, &n_all);
    if (rc != IFL_OK) return rc;
    if (n_all == 0) return IFL_OK;
    const int blocks = (n + 255) / 256;
    // rounds in batches; stop when a whole batch resolves nothing
    const int batch = 16;
    for (int guard = 0; guard < (f.w + f.h); guard += batch) {
        IFL_CUDA(cudaMemsetAsync(c->ext_ready, 0, sizeof(int), c->stream));
        for (int r = 0; r < batch; r++) {
            if ((rc = dist_barrier(c)) != IFL_OK) return rc; // masks promoted / cells listed by the neighbours
            if (n > 0) {
                k_ext_round<<<blocks, 256, 0, c->stream>>>(f, n, c->ext_ready);
                IFL_LAUNCHED(c);
            }
            if ((rc = dist_barrier(c)) != IFL_OK) return rc; // nobody still reads the flags promoted next
            if (n > 0) {
                k_ext_promote<<<blocks, 256, 0, c->stream>>>(f, n);
                IFL_LAUNCHED(c);
            }
        }
        int resolved = 0;
        IFL_CUDA(cudaMemcpyAsync(&resolved, c->ext_ready, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IFL_CUDA(cudaStreamSynchronize(c->stream));
        long long resolved_all = resolved;
        if ((rc = dist_host_sum(c, &resolved_all)) != IFL_OK) return rc;
        if (resolved_all == 0) break;
    }
    return dist_barrier(c);
}

} // namespace ifl
