// assembly_kernels.cu -- pressure-system assembly and application, inflow stamping.
//
//   build_rhs        FluidSolver::buildRhs            v3:208-217
//   build_matrix     FluidSolver::buildPressureMatrix v3:222-244
//   apply_pressure   FluidSolver::applyPressure       v3:382-398
//   add_inflow       FluidQuantity::addInflow         v2:188-205 / v1:141-151
//
// The reference writes these as scatter loops in raster order.  Here each output
// element is owned by one thread (gather form); the floating-point additions are
// issued in the order in which the reference's scatter would have ARRIVED at that
// element, so results are bit-identical.
// All kernels are HBM-streaming pointwise kernels: one thread per element,
// x fastest, 256-thread blocks covering a 128x2... rows of the pitched arrays.
#include "ifl_internal.cuh"

namespace ifl {

// r = -(1/hx) * (u[x+1,y] - u[x,y] + v[x,y+1] - v[x,y])            v3:213-214
__global__ void __launch_bounds__(256) k_build_rhs(Arr r, Arr u, Arr v, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = r.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= r.w || y >= r.ry1) return;
    const double ul = u.p[x + (size_t)y * u.pitch];
    const double ur = u.p[x + 1 + (size_t)y * u.pitch];
    const double vt = v.p[x + (size_t)y * v.pitch];
    const double vb = v.p[x + (size_t)(y + 1) * v.pitch];
    r.p[x + (size_t)y * r.pitch] = -scale * (ur - ul + vb - vt);
}

// Gather form of v3:227-243.  aDiag[idx] receives, in raster order of the scattering
// cell: +scale from the cell above (its y-branch), +scale from the cell to the left
// (its x-branch), then the cell's own x-branch and y-branch.
__global__ void __launch_bounds__(256) k_build_matrix(Arr aDiag, Arr aPlusX, Arr aPlusY, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = aDiag.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int w = aDiag.w, h = aDiag.h;
    if (x >= w || y >= aDiag.ry1) return;
    double diag = 0.0;
    if (y > 0) diag += scale;
    if (x > 0) diag += scale;
    double ax = 0.0, ay = 0.0;
    if (x < w - 1) {
        diag += scale;
        ax = -scale;
    }
    if (y < h - 1) {
        diag += scale;
        ay = -scale;
    }
    const size_t i = x + (size_t)y * aDiag.pitch;
    aDiag.p[i] = diag;
    aPlusX.p[i] = ax;
    aPlusY.p[i] = ay;
}

// u face (x,y), x in [0,W]: receives "+= scale*p[x-1,y]" (cell x-1, visited first)
// and then "-= scale*p[x,y]" (cell x).  v3:387-388; wall faces zeroed v3:394-395.
__global__ void __launch_bounds__(256) k_apply_pressure_u(Arr u, Arr p, double scale, int zero_walls) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = u.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int W = p.w;
    if (x > W || y >= u.ry1) return;
    const size_t iu = x + (size_t)y * u.pitch;
    double val = u.p[iu];
    if (x > 0) val += scale * p.p[x - 1 + (size_t)y * p.pitch];
    if (x < W) val -= scale * p.p[x + (size_t)y * p.pitch];
    if (zero_walls && (x == 0 || x == W)) val = 0.0;
    u.p[iu] = val;
}

// v face (x,y), y in [0,H]: "+= scale*p[x,y-1]" then "-= scale*p[x,y]".  v3:389-390, 396-397.
__global__ void __launch_bounds__(256) k_apply_pressure_v(Arr v, Arr p, double scale, int zero_walls) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = v.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = p.h;
    if (x >= v.w || y >= v.ry1) return; // v has H+1 rows; the last slab owns row H
    const size_t iv = x + (size_t)y * v.pitch;
    double val = v.p[iv];
    if (y > 0) val += scale * p.p[x + (size_t)(y - 1) * p.pitch];
    if (y < H) val -= scale * p.p[x + (size_t)y * p.pitch];
    if (zero_walls && (y == 0 || y == H)) val = 0.0;
    v.p[iv] = val;
}

// FluidQuantity::addInflow.  The reference clamps the x loop with _h, not _w
// (v2:195) -- kept.  smooth==0 is chapter 1's hard-edged
// stamp (v1:147-150), smooth==1 the cubic-pulse blob (v2:196-202).
__global__ void k_add_inflow(Arr src, int ix_lo, int ix_hi, int iy_lo, int iy_hi, double hx, double x0,
                             double y0, double x1, double y1, double v, int smooth) {
    const int x = ix_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = iy_lo + blockIdx.y;
    if (x >= ix_hi || y >= iy_hi) return;
    double vi = v;
    if (smooth) {
        const double lx = (2.0 * (x + 0.5) * hx - (x0 + x1)) / (x1 - x0);
        const double ly = (2.0 * (y + 0.5) * hx - (y0 + y1)) / (y1 - y0);
        const double l = sqrt(lx * lx + ly * ly); // length() v2:33
        const double xx = std_min(fabs(l), 1.0);   // cubicPulse() v2:42-45
        vi = (1.0 - xx * xx * (3.0 - 2.0 * xx)) * v;
    }
    const size_t i = x + (size_t)y * src.pitch;
    if (fabs(src.p[i]) < fabs(vi)) src.p[i] = vi;
}

// The same stamp on a w < h grid whose x range runs past the row (the `_h` clamp, quirk 2): the
// reference's dense index x + y*_w then lands in the following rows, cell (x mod w, y + x div w).
// One thread per TARGET cell applies its sources in the reference's raster order (ascending y), so the
// result is exact and race-free; targets past the last row (the reference writes out of bounds there)
// are dropped.
__global__ void k_add_inflow_wrap(Arr src, int ix_lo, int ix_hi, int iy_lo, int iy_hi, double hx, double x0, double y0,
                                  double x1, double y1, double v, int smooth) {
    const int tx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = src.ry0 + blockIdx.y;
    if (tx >= src.w || ty >= src.ry1) return;
    const size_t i = tx + (size_t)ty * src.pitch;
    double cur = src.p[i];
    bool touched = false;
    for (int j = (ix_hi - 1 - tx) / src.w; j >= 0; j--) { // source (tx + j*w, ty - j): ascending y
        const int x = tx + j * src.w, y = ty - j;
        if (x < ix_lo || x >= ix_hi || y < iy_lo || y >= iy_hi) continue;
        double vi = v;
        if (smooth) {
            const double lx = (2.0 * (x + 0.5) * hx - (x0 + x1)) / (x1 - x0);
            const double ly = (2.0 * (y + 0.5) * hx - (y0 + y1)) / (y1 - y0);
            const double l = sqrt(lx * lx + ly * ly);
            const double xx = std_min(fabs(l), 1.0);
            vi = (1.0 - xx * xx * (3.0 - 2.0 * xx)) * v;
        }
        if (fabs(cur) < fabs(vi)) {
            cur = vi;
            touched = true;
        }
    }
    if (touched) src.p[i] = cur;
}

// ---------------------------------------------------------- chapters 4+ variants ----
__device__ __forceinline__ double bvx(const BodyDev &b, double y) { return (b.posY - y) * b.velTheta + b.velX; } // v4:125
__device__ __forceinline__ double bvy(const BodyDev &b, double x) { return (x - b.posX) * b.velTheta + b.velY; } // v4:129

// buildRhs with solid cells (v4:614-628) and, for chapter 5+, fractional volumes and the
// solid-velocity blend terms (v5:654-683).
__global__ void __launch_bounds__(256) k_build_rhs_solid(Arr r, Field d, Field u, Field v, double hx, double scale,
                                                         const BodyDev *bodies, int nb, int curved) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = r.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int w = r.w, h = r.h;
    if (x >= w || y >= r.ry1) return;
    const size_t ic = x + (size_t)y * d.src.pitch;
    const size_t iu = x + (size_t)y * u.src.pitch;
    const size_t iv = x + (size_t)y * v.src.pitch;
    double out = 0.0;
    if (d.cell[ic] == CELL_FLUID) {
        if (!curved) {
            out = -scale * (u.src.p[iu + 1] - u.src.p[iu] + v.src.p[iv + v.src.pitch] - v.src.p[iv]);
        } else {
            const double uvr = u.volume.p[iu + 1], uvl = u.volume.p[iu];
            const double vvb = v.volume.p[iv + v.src.pitch], vvt = v.volume.p[iv];
            out = -scale * (uvr * u.src.p[iu + 1] - uvl * u.src.p[iu] + vvb * v.src.p[iv + v.src.pitch] - vvt * v.src.p[iv]);
            const double vol = d.volume.p[ic];
            if (nb > 0) {
                if (x > 0) out -= (uvl - vol) * bvx(bodies[d.body[ic - 1]], (y + 0.5) * hx);
                if (y > 0) out -= (vvt - vol) * bvy(bodies[d.body[ic - d.src.pitch]], (x + 0.5) * hx);
                if (x < w - 1) out += (uvr - vol) * bvx(bodies[d.body[ic + 1]], (y + 0.5) * hx);
                if (y < h - 1) out += (vvb - vol) * bvy(bodies[d.body[ic + d.src.pitch]], (x + 0.5) * hx);
            }
        }
    }
    r.p[x + (size_t)y * r.pitch] = out;
}

// buildPressureMatrix with solid cells, gather form of v5:694-712 (v4:652-670).  aDiag
// receives its face factors in the raster order of the scattering cell: the face shared
// with the cell above, with the cell to the left, then the cell's own right and lower
// faces.  With fractional volumes the four factors differ, so the order is significant.
__global__ void __launch_bounds__(256) k_build_matrix_solid(Arr aDiag, Arr aPlusX, Arr aPlusY, Field d, Field u, Field v,
                                                            double scale, int curved) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = aDiag.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int w = aDiag.w, h = aDiag.h;
    if (x >= w || y >= aDiag.ry1) return;
    const int cp = d.src.pitch;
    const size_t ic = x + (size_t)y * cp;
    double diag = 0.0, ax = 0.0, ay = 0.0;
    if (d.cell[ic] == CELL_FLUID) {
        const size_t iu = x + (size_t)y * u.src.pitch;
        const size_t iv = x + (size_t)y * v.src.pitch;
        if (y > 0 && d.cell[ic - cp] == CELL_FLUID) diag += curved ? scale * v.volume.p[iv] : scale;
        if (x > 0 && d.cell[ic - 1] == CELL_FLUID) diag += curved ? scale * u.volume.p[iu] : scale;
        if (x < w - 1 && d.cell[ic + 1] == CELL_FLUID) {
            const double factor = curved ? scale * u.volume.p[iu + 1] : scale;
            diag += factor;
            ax = -factor;
        }
        if (y < h - 1 && d.cell[ic + cp] == CELL_FLUID) {
            const double factor = curved ? scale * v.volume.p[iv + v.src.pitch] : scale;
            diag += factor;
            ay = -factor;
        }
    }
    const size_t i = x + (size_t)y * aDiag.pitch;
    aDiag.p[i] = diag;
    aPlusX.p[i] = ax;
    aPlusY.p[i] = ay;
}

// applyPressure with solid cells (v4:796-810): only fluid cells push on their faces.
__global__ void __launch_bounds__(256) k_apply_pressure_u_solid(Arr u, Arr p, Field d, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = u.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int W = p.w;
    if (x > W || y >= u.ry1) return;
    const size_t iu = x + (size_t)y * u.pitch;
    const size_t ic = x + (size_t)y * d.src.pitch;
    double val = u.p[iu];
    if (x > 0 && d.cell[ic - 1] == CELL_FLUID) val += scale * p.p[x - 1 + (size_t)y * p.pitch];
    if (x < W && d.cell[ic] == CELL_FLUID) val -= scale * p.p[x + (size_t)y * p.pitch];
    u.p[iu] = val;
}

__global__ void __launch_bounds__(256) k_apply_pressure_v_solid(Arr v, Arr p, Field d, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = v.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = p.h;
    if (x >= v.w || y >= v.ry1) return; // v has H+1 rows; the last slab owns row H
    const size_t iv = x + (size_t)y * v.pitch;
    const size_t ic = x + (size_t)y * d.src.pitch;
    double val = v.p[iv];
    if (y > 0 && d.cell[ic - d.src.pitch] == CELL_FLUID) val += scale * p.p[x + (size_t)(y - 1) * p.pitch];
    if (y < H && d.cell[ic] == CELL_FLUID) val -= scale * p.p[x + (size_t)y * p.pitch];
    v.p[iv] = val;
}

// ---------------------------------------------------------- chapters 6+ (heat, density) ----
// buildHeatDiffusionMatrix, gather form of v6:683-712: aDiag starts at 1.0 everywhere and
// collects +scale per fluid-fluid face in raster order of the scattering cell.
__global__ void __launch_bounds__(256) k_heat_matrix(Arr aDiag, Arr aPlusX, Arr aPlusY, Field d, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = aDiag.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int w = aDiag.w, h = aDiag.h;
    if (x >= w || y >= aDiag.ry1) return;
    const int cp = d.src.pitch;
    const size_t ic = x + (size_t)y * cp;
    double diag = 1.0, ax = 0.0, ay = 0.0;
    if (d.cell[ic] == CELL_FLUID) {
        if (y > 0 && d.cell[ic - cp] == CELL_FLUID) diag += scale;
        if (x > 0 && d.cell[ic - 1] == CELL_FLUID) diag += scale;
        if (x < w - 1 && d.cell[ic + 1] == CELL_FLUID) {
            diag += scale;
            ax = -scale;
        }
        if (y < h - 1 && d.cell[ic + cp] == CELL_FLUID) {
            diag += scale;
            ay = -scale;
        }
    }
    const size_t i = x + (size_t)y * aDiag.pitch;
    aDiag.p[i] = diag;
    aPlusX.p[i] = ax;
    aPlusY.p[i] = ay;
}

// addBuoyancy (v6:881-895), gather form: v face (x,y) receives the upper cell's half first
// (that cell is earlier in raster order), then the lower cell's half.  Not masked.
__global__ void __launch_bounds__(256) k_add_buoyancy(Arr v, Arr dsrc, Arr tsrc, double tg, double alpha, double tAmb) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = v.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = dsrc.h;
    if (x >= v.w || y >= v.ry1) return;
    const size_t iv = x + (size_t)y * v.pitch;
    double val = v.p[iv];
    if (y > 0) {
        const size_t i = x + (size_t)(y - 1) * dsrc.pitch;
        const double b = tg * (alpha * dsrc.p[i] - (tsrc.p[i] - tAmb) / tAmb);
        val += b * 0.5;
    }
    if (y < H) {
        const size_t i = x + (size_t)y * dsrc.pitch;
        const double b = tg * (alpha * dsrc.p[i] - (tsrc.p[i] - tAmb) / tAmb);
        val += b * 0.5;
    }
    v.p[iv] = val;
}

// computeDensities (v7:658-675): face density = sum of the halves of the adjacent cells
__device__ __forceinline__ double cell_density(const Arr &dsrc, const Arr &tsrc, int x, int y, double rhoAir, double tAmb,
                                               double alpha) {
    const size_t i = x + (size_t)y * dsrc.pitch;
    const double density = rhoAir * tAmb / tsrc.p[i] * (1.0 + alpha * dsrc.p[i]);
    return std_max(density, 0.05 * rhoAir);
}
__global__ void __launch_bounds__(256) k_density_u(Arr ud, Arr dsrc, Arr tsrc, double rhoAir, double tAmb, double alpha) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = ud.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int W = dsrc.w;
    if (x > W || y >= ud.ry1) return;
    double val = 0.0;
    if (x > 0) val += 0.5 * cell_density(dsrc, tsrc, x - 1, y, rhoAir, tAmb, alpha);
    if (x < W) val += 0.5 * cell_density(dsrc, tsrc, x, y, rhoAir, tAmb, alpha);
    ud.p[x + (size_t)y * ud.pitch] = val;
}
__global__ void __launch_bounds__(256) k_density_v(Arr vd, Arr dsrc, Arr tsrc, double rhoAir, double tAmb, double alpha) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = vd.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = dsrc.h;
    if (x >= vd.w || y >= vd.ry1) return;
    double val = 0.0;
    if (y > 0) val += 0.5 * cell_density(dsrc, tsrc, x, y - 1, rhoAir, tAmb, alpha);
    if (y < H) val += 0.5 * cell_density(dsrc, tsrc, x, y, rhoAir, tAmb, alpha);
    vd.p[x + (size_t)y * vd.pitch] = val;
}

// element `flat` of the reference's DENSE vDensity array (w columns per row)
__device__ __forceinline__ double vdens_flat(const Arr &vd, int flat) { return vd.p[(flat % vd.w) + (size_t)(flat / vd.w) * vd.pitch]; }

// buildPressureMatrix with variable density (v7:680-707), gather form.  The reference indexes
// _vDensity with _u->idx (row stride w+1 instead of w); kept.
__global__ void __launch_bounds__(256) k_build_matrix_density(Arr aDiag, Arr aPlusX, Arr aPlusY, Field d, Field u,
                                                              Field v, Arr ud, Arr vd, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = aDiag.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int w = aDiag.w, h = aDiag.h;
    if (x >= w || y >= aDiag.ry1) return;
    const int cp = d.src.pitch;
    const size_t ic = x + (size_t)y * cp;
    double diag = 0.0, ax = 0.0, ay = 0.0;
    if (d.cell[ic] == CELL_FLUID) {
        const size_t iu = x + (size_t)y * u.src.pitch;
        const size_t iv = x + (size_t)y * v.src.pitch;
        if (y > 0 && d.cell[ic - cp] == CELL_FLUID) diag += scale * v.volume.p[iv] / vdens_flat(vd, x + y * (w + 1));
        if (x > 0 && d.cell[ic - 1] == CELL_FLUID) diag += scale * u.volume.p[iu] / ud.p[x + (size_t)y * ud.pitch];
        if (x < w - 1 && d.cell[ic + 1] == CELL_FLUID) {
            const double factor = scale * u.volume.p[iu + 1] / ud.p[x + 1 + (size_t)y * ud.pitch];
            diag += factor;
            ax = -factor;
        }
        if (y < h - 1 && d.cell[ic + cp] == CELL_FLUID) {
            const double factor = scale * v.volume.p[iv + v.src.pitch] / vdens_flat(vd, x + (y + 1) * (w + 1));
            diag += factor;
            ay = -factor;
        }
    }
    const size_t i = x + (size_t)y * aDiag.pitch;
    aDiag.p[i] = diag;
    aPlusX.p[i] = ax;
    aPlusY.p[i] = ay;
}

// applyPressure with face densities (v7:891-906)
__global__ void __launch_bounds__(256) k_apply_pressure_u_density(Arr u, Arr p, Field d, Arr ud, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = u.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int W = p.w;
    if (x > W || y >= u.ry1) return;
    const size_t iu = x + (size_t)y * u.pitch;
    const size_t ic = x + (size_t)y * d.src.pitch;
    const double dens = ud.p[x + (size_t)y * ud.pitch];
    double val = u.p[iu];
    if (x > 0 && d.cell[ic - 1] == CELL_FLUID) val += scale * p.p[x - 1 + (size_t)y * p.pitch] / dens;
    if (x < W && d.cell[ic] == CELL_FLUID) val -= scale * p.p[x + (size_t)y * p.pitch] / dens;
    u.p[iu] = val;
}
__global__ void __launch_bounds__(256) k_apply_pressure_v_density(Arr v, Arr p, Field d, Arr vd, double scale) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = v.ry0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int H = p.h;
    if (x >= v.w || y >= v.ry1) return; // v has H+1 rows; the last slab owns row H
    const size_t iv = x + (size_t)y * v.pitch;
    const size_t ic = x + (size_t)y * d.src.pitch;
    const double dens = vd.p[x + (size_t)y * vd.pitch];
    double val = v.p[iv];
    if (y > 0 && d.cell[ic - d.src.pitch] == CELL_FLUID) val += scale * p.p[x + (size_t)(y - 1) * p.pitch] / dens;
    if (y < H && d.cell[ic] == CELL_FLUID) val -= scale * p.p[x + (size_t)y * p.pitch] / dens;
    v.p[iv] = val;
}

static dim3 grid2d(int w, int h) { return dim3((w + 63) / 64, (h + 3) / 4); }
// grid over the rows of `a` this rank owns (one GPU: all of them)
static dim3 grid_rows(const Arr &a) { return grid2d(a.w, a.ry1 - a.ry0); }

int launch_build_rhs(ifl_ctx *c) {
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double scale = 1.0 / c->hx;
    if (c->version >= 4)
        k_build_rhs_solid<<<grid_rows(c->r), 256, 0, c->stream>>>(c->r, c->fd[IFL_FIELD_D], c->fd[IFL_FIELD_U],
                                                                      c->fd[IFL_FIELD_V], c->hx, scale, c->bodies_d,
                                                                      c->n_bodies, c->version >= 5);
    else
        k_build_rhs<<<grid_rows(c->r), 256, 0, c->stream>>>(c->r, c->fd[IFL_FIELD_U].src, c->fd[IFL_FIELD_V].src, scale);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_build_matrix(ifl_ctx *c, double timestep, double density) {
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double scale = timestep / (density * c->hx * c->hx);
    if (c->version >= 7) {
        const double scale7 = timestep / (c->hx * c->hx); // v7:681
        k_build_matrix_density<<<grid_rows(c->aDiag), 256, 0, c->stream>>>(c->aDiag, c->aPlusX, c->aPlusY,
                                                                           c->fd[IFL_FIELD_D], c->fd[IFL_FIELD_U],
                                                                           c->fd[IFL_FIELD_V], c->uDensity, c->vDensity,
                                                                           scale7);
    } else if (c->version >= 4)
        k_build_matrix_solid<<<grid_rows(c->aDiag), 256, 0, c->stream>>>(c->aDiag, c->aPlusX, c->aPlusY,
                                                                         c->fd[IFL_FIELD_D], c->fd[IFL_FIELD_U],
                                                                         c->fd[IFL_FIELD_V], scale, c->version >= 5);
    else
        k_build_matrix<<<grid_rows(c->aDiag), 256, 0, c->stream>>>(c->aDiag, c->aPlusX, c->aPlusY, scale);
    // chapter 3: the matrix is a function of position and `scale` alone; k_matvec may evaluate it instead of reading it
    c->matrix_uniform = (c->version == 3 && c->matvec_uniform_allowed) ? 1 : 0;
    c->matrix_scale = scale;
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_apply_pressure(ifl_ctx *c, double timestep, double density) {
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double scale = timestep / (density * c->hx);
    Field &u = c->fd[IFL_FIELD_U], &v = c->fd[IFL_FIELD_V];
    if (c->version >= 7) {
        const double scale7 = timestep / c->hx; // v7:892
        k_apply_pressure_u_density<<<grid_rows(u.src), 256, 0, c->stream>>>(u.src, c->p, c->fd[IFL_FIELD_D], c->uDensity,
                                                                             scale7);
        IFL_LAUNCHED(c);
        k_apply_pressure_v_density<<<grid_rows(v.src), 256, 0, c->stream>>>(v.src, c->p, c->fd[IFL_FIELD_D], c->vDensity,
                                                                             scale7);
        IFL_LAUNCHED(c);
        return IFL_OK;
    }
    if (c->version >= 4) {
        k_apply_pressure_u_solid<<<grid_rows(u.src), 256, 0, c->stream>>>(u.src, c->p, c->fd[IFL_FIELD_D], scale);
        IFL_LAUNCHED(c);
        k_apply_pressure_v_solid<<<grid_rows(v.src), 256, 0, c->stream>>>(v.src, c->p, c->fd[IFL_FIELD_D], scale);
        IFL_LAUNCHED(c);
        return IFL_OK;
    }
    k_apply_pressure_u<<<grid_rows(u.src), 256, 0, c->stream>>>(u.src, c->p, scale, 1);
    IFL_LAUNCHED(c);
    k_apply_pressure_v<<<grid_rows(v.src), 256, 0, c->stream>>>(v.src, c->p, scale, 1);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_build_heat_matrix(ifl_ctx *c, double timestep) { // v6:683-712
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double scale = c->diffusion * timestep * 1.0 / (c->hx * c->hx);
    k_heat_matrix<<<grid_rows(c->aDiag), 256, 0, c->stream>>>(c->aDiag, c->aPlusX, c->aPlusY, c->fd[IFL_FIELD_D], scale);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_add_buoyancy(ifl_ctx *c, double timestep) { // v6:881-895
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double alpha = (c->rho_soot - c->rho_air) / c->rho_air;
    Field &v = c->fd[IFL_FIELD_V];
    k_add_buoyancy<<<grid_rows(v.src), 256, 0, c->stream>>>(v.src, c->fd[IFL_FIELD_D].src, c->fd[IFL_FIELD_T].src,
                                                            timestep * c->g, alpha, c->t_amb);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_compute_densities(ifl_ctx *c) { // v7:658-675
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    const double alpha = (c->rho_soot - c->rho_air) / c->rho_air;
    Arr &ds = c->fd[IFL_FIELD_D].src, &ts = c->fd[IFL_FIELD_T].src;
    k_density_u<<<grid_rows(c->uDensity), 256, 0, c->stream>>>(c->uDensity, ds, ts, c->rho_air, c->t_amb, alpha);
    IFL_LAUNCHED(c);
    k_density_v<<<grid_rows(c->vDensity), 256, 0, c->stream>>>(c->vDensity, ds, ts, c->rho_air, c->t_amb, alpha);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_add_inflow(ifl_ctx *c, int field, double x0, double y0, double x1, double y1, double v) {
    Field &f = c->fd[field];
    // v2:189-192 -- truncating casts, evaluated on the host exactly as the reference does
    const int ix0 = (int)(x0 / c->hx - f.ox);
    const int iy0 = (int)(y0 / c->hx - f.oy);
    const int ix1 = (int)(x1 / c->hx - f.ox);
    const int iy1 = (int)(y1 / c->hx - f.oy);
    const int xlo = imax(ix0, 0), xhi = imin(ix1, f.h); // sic: _h (v2:195)
    if (xhi > f.w) { // only on w < h grids: the stamp runs past the row and wraps into the following rows
        const int rlo = imax(iy0, 0), rhi = imin(iy1, f.h);
        if (xhi <= xlo || rhi <= rlo || f.src.ry1 <= f.src.ry0) return IFL_OK;
        ProfScope ps_(c, IFL_K_ASSEMBLY);
        k_add_inflow_wrap<<<dim3((f.w + 127) / 128, f.src.ry1 - f.src.ry0), 128, 0, c->stream>>>(
            f.src, xlo, xhi, rlo, rhi, c->hx, x0, y0, x1, y1, v, c->version >= 2 ? 1 : 0);
        IFL_LAUNCHED(c);
        return IFL_OK;
    }
    // rows: the reference's clamp, then this rank's slab
    const int ylo = imax(imax(iy0, 0), f.src.ry0), yhi = imin(imin(iy1, f.h), f.src.ry1);
    if (xhi <= xlo || yhi <= ylo) return IFL_OK;
    dim3 grid((xhi - xlo + 127) / 128, yhi - ylo);
    ProfScope ps_(c, IFL_K_ASSEMBLY);
    k_add_inflow<<<grid, 128, 0, c->stream>>>(f.src, xlo, xhi, ylo, yhi, c->hx, x0, y0, x1, y1, v,
                                              c->version >= 2 ? 1 : 0);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

} // namespace ifl
