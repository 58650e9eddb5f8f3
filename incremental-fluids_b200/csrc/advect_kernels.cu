// advect_kernels.cu -- semi-Lagrangian advection.
//
//   FluidQuantity::advect       v2:170-183 (v1:125-138)
//   FluidQuantity::rungeKutta3  v2:83-101   (third stage NOT divided by hx: SURVEY 3.5 q1)
//   FluidQuantity::euler        v1:68-74
//   FluidQuantity::lerp(x,y)    v2:133-145  (clamp [0, w-1.001], truncate, bilinear)
//   FluidQuantity::cerp(x,y)    v2:150-167  (4x4 Catmull-Rom, index clamp, min/max limiter)
//
// One thread per output cell, 32x8 cell blocks so that the 6 bilinear velocity
// gathers and the 16 Catmull-Rom taps of neighbouring threads fall into the same
// few cache lines (the back-trace displacement is smooth in space).  Every
// expression keeps the reference's operand order; the library is built with
// -fmad=false, so results are bit-identical to the CPU reference.
#include "ifl_internal.cuh"
#include "solid_geometry.cuh"

namespace ifl {

struct FieldView {
    const double *__restrict__ p;
    int w, h, pitch;
    double ox, oy;
};

__device__ __forceinline__ double at(const FieldView &f, int x, int y) {
    return __ldg(f.p + x + (size_t)y * f.pitch);
}

__device__ __forceinline__ double lerp1(double a, double b, double x) { return a * (1.0 - x) + b * x; } // v2:58

__device__ __forceinline__ double lerp2(const FieldView &f, double x, double y) {
    x = std_min(std_max(x - f.ox, 0.0), f.w - 1.001);
    y = std_min(std_max(y - f.oy, 0.0), f.h - 1.001);
    const int ix = (int)x;
    const int iy = (int)y;
    x -= ix;
    y -= iy;
    const double x00 = at(f, ix, iy), x10 = at(f, ix + 1, iy);
    const double x01 = at(f, ix, iy + 1), x11 = at(f, ix + 1, iy + 1);
    return lerp1(lerp1(x00, x10, x), lerp1(x01, x11, x), y);
}

// v2:66-80 -- the polynomial is written exactly as in the source, including the
// terms multiplied by 0.0 and 1.0 (they matter for signed zeros / non-finite data).
__device__ __forceinline__ double cerp1(double a, double b, double c, double d, double x) {
    const double xsq = x * x;
    const double xcu = xsq * x;
    const double minV = std_min(a, std_min(b, std_min(c, d)));
    const double maxV = std_max(a, std_max(b, std_max(c, d)));
    const double t = a * (0.0 - 0.5 * x + 1.0 * xsq - 0.5 * xcu) + b * (1.0 + 0.0 * x - 2.5 * xsq + 1.5 * xcu) +
                     c * (0.0 + 0.5 * x + 2.0 * xsq - 1.5 * xcu) + d * (0.0 + 0.0 * x - 0.5 * xsq + 0.5 * xcu);
    return std_min(std_max(t, minV), maxV);
}

__device__ __forceinline__ double cerp2(const FieldView &f, double x, double y) {
    x = std_min(std_max(x - f.ox, 0.0), f.w - 1.001);
    y = std_min(std_max(y - f.oy, 0.0), f.h - 1.001);
    const int ix = (int)x;
    const int iy = (int)y;
    x -= ix;
    y -= iy;
    const int x0 = imax(ix - 1, 0), x1 = ix, x2 = ix + 1, x3 = imin(ix + 2, f.w - 1);
    const int y0 = imax(iy - 1, 0), y1 = iy, y2 = iy + 1, y3 = imin(iy + 2, f.h - 1);
    const double q0 = cerp1(at(f, x0, y0), at(f, x1, y0), at(f, x2, y0), at(f, x3, y0), x);
    const double q1 = cerp1(at(f, x0, y1), at(f, x1, y1), at(f, x2, y1), at(f, x3, y1), x);
    const double q2 = cerp1(at(f, x0, y2), at(f, x1, y2), at(f, x2, y2), at(f, x3, y2), x);
    const double q3 = cerp1(at(f, x0, y3), at(f, x1, y3), at(f, x2, y3), at(f, x3, y3), x);
    return cerp1(q0, q1, q2, q3, y);
}

// Chapters 4+: advect only fluid cells (the others keep whatever _dst held: SURVEY 3.5
// quirk 4) and pull back-traced points that land inside a solid onto its surface
// (FluidQuantity::advect v5:468-485, backProject v5:455-466).
__global__ void __launch_bounds__(256) k_advect_solid(double *__restrict__ dst, int dst_pitch, FieldView self,
                                                       const uint8_t *__restrict__ cell, const uint8_t *__restrict__ body,
                                                       FieldView u, FieldView v, double timestep, double hx,
                                                       const BodyDev *__restrict__ bodies, int ry0, int ry1) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31);
    const int iy = ry0 + blockIdx.y * 8 + (threadIdx.x >> 5); // rows [ry0, ry1): this rank's slab
    if (ix >= self.w || iy >= ry1) return;
    if (cell[ix + (size_t)iy * self.pitch] != CELL_FLUID) return;
    double x = ix + self.ox;
    double y = iy + self.oy;
    const double firstU = lerp2(u, x, y) / hx;
    const double firstV = lerp2(v, x, y) / hx;
    const double midX = x - 0.5 * timestep * firstU;
    const double midY = y - 0.5 * timestep * firstV;
    const double midU = lerp2(u, midX, midY) / hx;
    const double midV = lerp2(v, midX, midY) / hx;
    const double lastX = x - 0.75 * timestep * midU;
    const double lastY = y - 0.75 * timestep * midV;
    const double lastU = lerp2(u, lastX, lastY);
    const double lastV = lerp2(v, lastX, lastY);
    x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
    y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
    // backProject
    const int rx = imin(imax((int)(x - self.ox), 0), self.w - 1);
    const int ry = imin(imax((int)(y - self.oy), 0), self.h - 1);
    const size_t ri = rx + (size_t)ry * self.pitch;
    if (cell[ri] != CELL_FLUID) {
        x = (x - self.ox) * hx;
        y = (y - self.oy) * hx;
        body_closest_surface_point(bodies[body[ri]], x, y);
        x = x / hx + self.ox;
        y = y / hx + self.oy;
    }
    dst[ix + (size_t)iy * dst_pitch] = cerp2(self, x, y);
}

template <bool RK3_CERP>
__global__ void __launch_bounds__(256) k_advect(double *__restrict__ dst, int dst_pitch, FieldView self, FieldView u,
                                                 FieldView v, double timestep, double hx, int ry0, int ry1) {
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31);
    const int iy = ry0 + blockIdx.y * 8 + (threadIdx.x >> 5); // rows [ry0, ry1): this rank's slab
    if (ix >= self.w || iy >= ry1) return;
    double x = ix + self.ox;
    double y = iy + self.oy;
    double out;
    if (RK3_CERP) {
        const double firstU = lerp2(u, x, y) / hx;
        const double firstV = lerp2(v, x, y) / hx;
        const double midX = x - 0.5 * timestep * firstU;
        const double midY = y - 0.5 * timestep * firstV;
        const double midU = lerp2(u, midX, midY) / hx;
        const double midV = lerp2(v, midX, midY) / hx;
        const double lastX = x - 0.75 * timestep * midU;
        const double lastY = y - 0.75 * timestep * midV;
        const double lastU = lerp2(u, lastX, lastY);
        const double lastV = lerp2(v, lastX, lastY);
        x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
        y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
        out = cerp2(self, x, y);
    } else { // chapter 1: forward Euler + bilinear, v1:68-74, v1:133-136
        const double uVel = lerp2(u, x, y) / hx;
        const double vVel = lerp2(v, x, y) / hx;
        x -= uVel * timestep;
        y -= vVel * timestep;
        out = lerp2(self, x, y);
    }
    dst[ix + (size_t)iy * dst_pitch] = out;
}

static FieldView view(const Field &f) {
    FieldView v;
    v.p = f.src.p;
    v.w = f.w;
    v.h = f.h;
    v.pitch = f.src.pitch;
    v.ox = f.ox;
    v.oy = f.oy;
    return v;
}

// FluidSolver::maxTimestep v1:310-328: max over the cell centres of |(u, v)| (bilinear samples), as block
// partials; the fold (a max: order-independent, exact) and 2*hx/max, min(.., 1.0) follow on the host side.
__global__ void __launch_bounds__(256) k_max_velocity(FieldView u, FieldView v, int W, int H, int ry0, int ry1,
                                                      double *__restrict__ partials) {
    __shared__ double red[32];
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y0 = ry0 + blockIdx.y * 16, y1 = imin(imin(y0 + 16, ry1), H);
    double m = 0.0;
    if (x < W)
        for (int y = y0; y < y1; y++) {
            const double uu = lerp2(u, x + 0.5, y + 0.5);
            const double vv = lerp2(v, x + 0.5, y + 0.5);
            m = std_max(m, sqrt(uu * uu + vv * vv));
        }
    m = block_reduce<true>(m, red);
    if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = m;
}

int launch_max_velocity(ifl_ctx *c) { // leaves the block partials in c->partials / c->n_partials
    ProfScope ps_(c, IFL_K_ADVECT);
    const Field &d = c->fd[IFL_FIELD_D];
    const int ry0 = d.src.ry0, ry1 = d.src.ry1;
    dim3 grid((c->W + 255) / 256, (ry1 - ry0 + 15) / 16); // as many partials as the PCG reductions: the buffer is sized for it
    k_max_velocity<<<grid, 256, 0, c->stream>>>(view(c->fd[IFL_FIELD_U]), view(c->fd[IFL_FIELD_V]), c->W, c->H, ry0, ry1,
                                               partials_next(c));
    c->n_partials = (int)(grid.x * grid.y);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

int launch_advect(ifl_ctx *c, int field, double timestep) {
    ProfScope ps_(c, IFL_K_ADVECT);
    Field &f = c->fd[field];
    const int ry0 = f.dst.ry0, ry1 = f.dst.ry1;
    dim3 grid_own((f.w + 31) / 32, (ry1 - ry0 + 7) / 8);
    FieldView self = view(f), u = view(c->fd[IFL_FIELD_U]), v = view(c->fd[IFL_FIELD_V]);
    if (c->version >= 4)
        k_advect_solid<<<grid_own, 256, 0, c->stream>>>(f.dst.p, f.dst.pitch, self, f.cell, f.body, u, v, timestep, c->hx,
                                                        c->bodies_d, ry0, ry1);
    else if (c->version >= 2)
        k_advect<true><<<grid_own, 256, 0, c->stream>>>(f.dst.p, f.dst.pitch, self, u, v, timestep, c->hx, ry0, ry1);
    else
        k_advect<false><<<grid_own, 256, 0, c->stream>>>(f.dst.p, f.dst.pitch, self, u, v, timestep, c->hx, ry0, ry1);
    IFL_LAUNCHED(c);
    return IFL_OK;
}

} // namespace ifl
