// solid_geometry.cuh -- SolidBox / SolidSphere (v4:152-241) as inline functions over plain BodyDev
// records.  rotate() (v4:58-62) uses the host-evaluated cos/sin of theta (glibc cos is even, sin is
// odd, exactly), so every expression below reproduces the reference bit for bit.
//
// Self-contained and free of CUDA headers: the kernels include it (through ifl_internal.cuh) and so does
// the host-side drop-in header (host/FluidSolver.hpp), whose SolidBody virtuals -- distance,
// closestSurfacePoint, distanceNormal (v4:116-118) -- evaluate exactly this code on the CPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define IFL_HD __host__ __device__
#else
#define IFL_HD
#endif

namespace ifl {

// One SolidBody (v4:79-241) as plain data; sin/cos of theta are evaluated on the host
// with libm so that rotate() (v4:58-62) is bit-identical to the reference.
struct BodyDev {
    int kind; // 0 SolidBox, 1 SolidSphere
    int pad;
    double posX, posY, scaleX, scaleY, theta, velX, velY, velTheta;
    double cosT, sinT;
};

IFL_HD inline double geo_max(double a, double b) { return (a < b) ? b : a; } // std::max

// ------------------------------------------------------------------ geometry ----
IFL_HD inline int nsgn(double v) { return v < 0.0 ? -1 : 1; }                  // v4:42-45
IFL_HD inline int sgn(double v) { return (0.0 < v ? 1 : 0) - (v < 0.0 ? 1 : 0); } // v4:38-40
IFL_HD inline double length2(double x, double y) { return sqrt(x * x + y * y); } // v4:47-49

// rotate(x, y, -theta): cos(-t) = cos t, sin(-t) = -sin t
IFL_HD inline void rot_neg(const BodyDev &b, double &x, double &y) {
    const double tx = x, ty = y, ns = -b.sinT;
    x = b.cosT * tx + ns * ty;
    y = -ns * tx + b.cosT * ty;
}
// rotate(x, y, +theta)
IFL_HD inline void rot_pos(const BodyDev &b, double &x, double &y) {
    const double tx = x, ty = y;
    x = b.cosT * tx + b.sinT * ty;
    y = -b.sinT * tx + b.cosT * ty;
}

IFL_HD inline double body_distance(const BodyDev &b, double x, double y) {
    if (b.kind == 0) { // SolidBox::distance v4:159-170
        x -= b.posX;
        y -= b.posY;
        rot_neg(b, x, y);
        const double dx = fabs(x) - b.scaleX * 0.5;
        const double dy = fabs(y) - b.scaleY * 0.5;
        if (dx >= 0.0 || dy >= 0.0) return length2(geo_max(dx, 0.0), geo_max(dy, 0.0));
        return geo_max(dx, dy);
    }
    return length2(x - b.posX, y - b.posY) - b.scaleX * 0.5; // SolidSphere::distance v4:210-212
}

IFL_HD inline void body_closest_surface_point(const BodyDev &b, double &x, double &y) {
    if (b.kind == 0) { // v4:172-187
        x -= b.posX;
        y -= b.posY;
        rot_neg(b, x, y);
        const double dx = fabs(x) - b.scaleX * 0.5;
        const double dy = fabs(y) - b.scaleY * 0.5;
        if (dx > dy)
            x = nsgn(x) * 0.5 * b.scaleX;
        else
            y = nsgn(y) * 0.5 * b.scaleY;
        rot_pos(b, x, y);
        x += b.posX;
        y += b.posY;
    } else { // v4:214-227 with globalToLocal / localToGlobal v4:91-105
        x -= b.posX;
        y -= b.posY;
        rot_neg(b, x, y);
        x /= b.scaleX;
        y /= b.scaleY;
        const double r = length2(x, y);
        if (r < 1e-4) {
            x = 0.5;
            y = 0.0;
        } else {
            x /= 2.0 * r;
            y /= 2.0 * r;
        }
        x *= b.scaleX;
        y *= b.scaleY;
        rot_pos(b, x, y);
        x += b.posX;
        y += b.posY;
    }
}

IFL_HD inline void body_distance_normal(const BodyDev &b, double &nx, double &ny, double x, double y) {
    if (b.kind == 0) { // v4:189-201
        x -= b.posX;
        y -= b.posY;
        rot_neg(b, x, y);
        if (fabs(x) - b.scaleX * 0.5 > fabs(y) - b.scaleY * 0.5) {
            nx = nsgn(x);
            ny = 0.0;
        } else {
            nx = 0.0;
            ny = nsgn(y);
        }
        rot_pos(b, nx, ny);
    } else { // v4:229-240 -- r is a float in the reference
        x -= b.posX;
        y -= b.posY;
        const float r = (float)length2(x, y);
        if (r < 1e-4) {
            nx = 1.0;
            ny = 0.0;
        } else {
            nx = x / r;
            ny = y / r;
        }
    }
}

IFL_HD inline double body_velocity_x(const BodyDev &b, double x, double y) { // v4:125-127
    (void)x;
    return (b.posY - y) * b.velTheta + b.velX;
}
IFL_HD inline double body_velocity_y(const BodyDev &b, double x, double y) { // v4:129-131
    (void)y;
    return (x - b.posX) * b.velTheta + b.velY;
}

} // namespace ifl
