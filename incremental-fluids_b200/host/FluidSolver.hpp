// FluidSolver.hpp -- drop-in host classes for tunabrain/incremental-fluids on top of
// libifl_b200.so.  Same public surface as the classes the reference defines inside each
// N-*/Fluid.cpp (there are no headers upstream), so a main() written against the
// reference compiles against this header unchanged:
//
//   FluidSolver(w, h, density)                       3-conjugate-gradients/Fluid.cpp:401
//   FluidSolver(w, h, density, bodies)               4-solid-boundaries/Fluid.cpp:836
//   void update(double timestep)                     v3:433 / v5:927
//   void addInflow(x, y, w, h, d, u, v)              v3:449
//   void toImage(unsigned char *rgba)                v3:455 / v5:961
//   FluidSolver(w, h, rhoAir, rhoSoot, diffusion, bodies)   6-heat/Fluid.cpp:921 (chapters 6-7)
//   void addInflow(x, y, w, h, d, t, u, v)           v6:1010
//   double ambientT()                                v6:1017
//   void toImage(unsigned char *rgba, bool renderHeat)      v6:1021
//   SolidBox(x,y,sx,sy,t,vx,vy,vt), SolidSphere(x,y,s,t,vx,vy,vt), SolidBody::update(dt)
//                                                    v4:155, v4:207, v4:140
//   SolidBody::distance / closestSurfacePoint / distanceNormal / velocity   v4:116-138 (host-callable virtuals)
//   double maxTimestep()                             1-matrixless/Fluid.cpp:310
//   FluidQuantity: src, at, lerp, cerp, cell, body, volume, addInflow, advect, flip, fillSolidFields,
//                  extrapolate, copy, diff, undiff, fromParticles     v3:41-185, v5:288-626, v8:271-688
//   ParticleQuantities: addQuantity, gridToParticles, particlesToGrid, advect   v8:865-939
//   chapter 8: FluidSolver(w, h, rhoAir, rhoSoot, diffusion, bodies) seeds the particles; update() carries the inflow (v8:1368)
//
// FluidQuantity and ParticleQuantities are VIEWS of a solver's device-resident state (the reference
// lets them own host arrays; here the arrays live in the solver's ifl_ctx), obtained from
// FluidSolver::quantity(IFL_FIELD_*) / FluidSolver::particles(); their methods forward to the same ABI
// entry points update() uses, so a caller can re-sequence a step exactly as the reference's update() does.
//
// The bodies of the reference's private hot-path methods live on the GPU; this header
// only forwards (one opaque ifl_ctx per solver) and prints the reference's own status
// lines ("Exiting solver after %d iterations, ...", v3:368, v3:379; "Particle count: %d", v8:926).
// Select the chapter with -DIFL_CHAPTER=1..8 (default 3), exactly as one would pick a reference directory.
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../include/ifl_b200.h"
#include "../csrc/solid_geometry.cuh" // the SAME geometry code the kernels run (plain C++ here)

#ifndef IFL_CHAPTER
#define IFL_CHAPTER 3
#endif
#ifndef IFL_AVG_PER_CELL
#define IFL_AVG_PER_CELL 4 // ParticleQuantities::_AvgPerCell, v8:698
#endif

class SolidBody { // v4:79-149
protected:
    double _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta;
    int _kind; // 0 box, 1 sphere: what the device evaluates (a user-defined shape has no device counterpart)
    SolidBody(int kind, double posX, double posY, double scaleX, double scaleY, double theta, double velX, double velY,
              double velTheta)
        : _posX(posX), _posY(posY), _scaleX(scaleX), _scaleY(scaleY), _theta(theta), _velX(velX), _velY(velY),
          _velTheta(velTheta), _kind(kind) {}

    ifl::BodyDev dev() const { // libm sin/cos per call, exactly like rotate() (v4:58-62)
        ifl::BodyDev b = {_kind, 0, _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta, cos(_theta), sin(_theta)};
        return b;
    }

public:
    virtual ~SolidBody() {}
    // v4:116-118 -- host-callable, same arithmetic as the device (csrc/solid_geometry.cuh)
    virtual double distance(double x, double y) const { return ifl::body_distance(dev(), x, y); }
    virtual void closestSurfacePoint(double &x, double &y) const { ifl::body_closest_surface_point(dev(), x, y); }
    virtual void distanceNormal(double &nx, double &ny, double x, double y) const { ifl::body_distance_normal(dev(), nx, ny, x, y); }
    double velocityX(double, double y) const { return (_posY - y) * _velTheta + _velX; } // v4:125
    double velocityY(double x, double) const { return (x - _posX) * _velTheta + _velY; } // v4:129
    void velocity(double &vx, double &vy, double x, double y) const { // v4:133
        vx = velocityX(x, y);
        vy = velocityY(x, y);
    }
    void update(double timestep) { // v4:140
        _posX += _velX * timestep;
        _posY += _velY * timestep;
        _theta += _velTheta * timestep;
    }
    ifl_body record() const {
        ifl_body b = {_kind, _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta};
        return b;
    }
};

class SolidBox : public SolidBody { // v4:152-201
public:
    SolidBox(double x, double y, double sx, double sy, double t, double vx, double vy, double vt)
        : SolidBody(0, x, y, sx, sy, t, vx, vy, vt) {}
};

class SolidSphere : public SolidBody { // v4:204-241
public:
    SolidSphere(double x, double y, double s, double t, double vx, double vy, double vt)
        : SolidBody(1, x, y, s, s, t, vx, vy, vt) {}
};

static inline void ifl_host_check(int rc, const char *what) {
    if (rc != IFL_OK) {
        fprintf(stderr, "libifl_b200: %s failed (%d): %s\n", what, rc, ifl_last_error());
        abort(); // the reference has no error channel either
    }
}

// One scalar field of the solver (v3:41-185, v5:288-626, v8:271-688) as a view of the device-resident
// arrays.  src() / at() / lerp() / cerp() work on a host copy that is refreshed by src(); the interpolation
// formulas are the reference's (lerp v2:133-145, cerp v2:150-167, the 1-D kernels v2:58, v2:66-80).
class FluidQuantity {
    ifl_ctx *_ctx;
    int _field, _w, _h;
    double _ox, _oy;
    std::vector<double> _host;
    std::vector<unsigned char> _bytes;

    static double lerp1(double a, double b, double x) { return a * (1.0 - x) + b * x; }
    static double cerp1(double a, double b, double c, double d, double x) {
        const double xsq = x * x, xcu = xsq * x;
        const double minV = std::min(a, std::min(b, std::min(c, d)));
        const double maxV = std::max(a, std::max(b, std::max(c, d)));
        const double t = a * (0.0 - 0.5 * x + 1.0 * xsq - 0.5 * xcu) + b * (1.0 + 0.0 * x - 2.5 * xsq + 1.5 * xcu) +
                         c * (0.0 + 0.5 * x + 2.0 * xsq - 1.5 * xcu) + d * (0.0 + 0.0 * x - 0.5 * xsq + 0.5 * xcu);
        return std::min(std::max(t, minV), maxV);
    }
    int srcBuf() const { return IFL_BUF_D_SRC + 2 * _field; }

public:
    FluidQuantity(ifl_ctx *ctx, int field, int w, int h, double ox, double oy)
        : _ctx(ctx), _field(field), _w(w), _h(h), _ox(ox), _oy(oy) {}

    int width() const { return _w; }
    int height() const { return _h; }
    // current values, reference layout x + y*w (downloads the field)
    const double *src() {
        _host.resize((size_t)_w * _h);
        ifl_host_check(ifl_download(_ctx, srcBuf(), &_host[0]), "ifl_download");
        return &_host[0];
    }
    double at(int x, int y) { // v3:113
        if (_host.empty()) src();
        return _host[x + (size_t)y * _w];
    }
    void upload(const double *values) { ifl_host_check(ifl_upload(_ctx, srcBuf(), values), "ifl_upload"); } // writes through src()
    double lerp(double x, double y) { // v2:133-145 on the host copy
        if (_host.empty()) src();
        x = std::min(std::max(x - _ox, 0.0), _w - 1.001);
        y = std::min(std::max(y - _oy, 0.0), _h - 1.001);
        const int ix = (int)x, iy = (int)y;
        x -= ix;
        y -= iy;
        const double *p = &_host[ix + (size_t)iy * _w];
        return lerp1(lerp1(p[0], p[1], x), lerp1(p[_w], p[_w + 1], x), y);
    }
    double cerp(double x, double y) { // v2:150-167 on the host copy
        if (_host.empty()) src();
        x = std::min(std::max(x - _ox, 0.0), _w - 1.001);
        y = std::min(std::max(y - _oy, 0.0), _h - 1.001);
        const int ix = (int)x, iy = (int)y;
        x -= ix;
        y -= iy;
        const int x0 = std::max(ix - 1, 0), x1 = ix, x2 = ix + 1, x3 = std::min(ix + 2, _w - 1);
        const int y0 = std::max(iy - 1, 0), y1 = iy, y2 = iy + 1, y3 = std::min(iy + 2, _h - 1);
        const double *h = &_host[0];
        const size_t w = (size_t)_w;
        const double q0 = cerp1(h[x0 + y0 * w], h[x1 + y0 * w], h[x2 + y0 * w], h[x3 + y0 * w], x);
        const double q1 = cerp1(h[x0 + y1 * w], h[x1 + y1 * w], h[x2 + y1 * w], h[x3 + y1 * w], x);
        const double q2 = cerp1(h[x0 + y2 * w], h[x1 + y2 * w], h[x2 + y2 * w], h[x3 + y2 * w], x);
        const double q3 = cerp1(h[x0 + y3 * w], h[x1 + y3 * w], h[x2 + y3 * w], h[x3 + y3 * w], x);
        return cerp1(q0, q1, q2, q3, y);
    }
    void addInflow(double x0, double y0, double x1, double y1, double v) { // v2:188-205
        ifl_host_check(ifl_quantity_add_inflow(_ctx, _field, x0, y0, x1, y1, v), "ifl_quantity_add_inflow");
        _host.clear();
    }
#if IFL_CHAPTER <= 7
    // v2:170-183 / v5:468-485: samples the SOLVER's u and v (the only pair there is), like update() does
    void advect(double timestep) { ifl_host_check(ifl_advect(_ctx, _field, timestep), "ifl_advect"); }
    void flip() { // v3:105
        ifl_host_check(ifl_flip(_ctx, _field), "ifl_flip");
        _host.clear();
    }
#endif
#if IFL_CHAPTER >= 4
    const unsigned char *cell() { // v4:287
        _bytes.resize((size_t)_w * _h);
        ifl_host_check(ifl_aux_download(_ctx, _field, IFL_AUX_CELL, &_bytes[0]), "ifl_aux_download");
        return &_bytes[0];
    }
    void fillSolidFields() { ifl_host_check(ifl_fill_solid_fields(_ctx, _field), "ifl_fill_solid_fields"); } // bodies: the solver's list
    void extrapolate() { // v4:551 / v8:611
        ifl_host_check(ifl_extrapolate(_ctx, _field), "ifl_extrapolate");
        _host.clear();
    }
#endif
#if IFL_CHAPTER >= 5
    double volume(int x, int y) { // v5:333
        std::vector<double> vol((size_t)_w * _h);
        ifl_host_check(ifl_aux_download(_ctx, _field, IFL_AUX_VOLUME, &vol[0]), "ifl_aux_download");
        return vol[x + (size_t)y * _w];
    }
#endif
#if IFL_CHAPTER >= 8
    void copy() { ifl_host_check(ifl_quantity_copy(_ctx, _field), "ifl_quantity_copy"); }                    // v8:374
    void diff(double alpha) { ifl_host_check(ifl_quantity_diff(_ctx, _field, alpha), "ifl_quantity_diff"); _host.clear(); }     // v8:379
    void undiff(double alpha) { ifl_host_check(ifl_quantity_undiff(_ctx, _field, alpha), "ifl_quantity_undiff"); _host.clear(); } // v8:385
    void fromParticles() { ifl_host_check(ifl_from_particles(_ctx, _field), "ifl_from_particles"); _host.clear(); }               // v8:663
#endif
};

#if IFL_CHAPTER >= 8
// ParticleQuantities (v8:692-940) as a view of the solver's device-resident particle set.
class ParticleQuantities {
    ifl_ctx *_ctx;

public:
    explicit ParticleQuantities(ifl_ctx *ctx) : _ctx(ctx) {}
    void addQuantity(FluidQuantity *) {} // the four quantities d, t, u, v are registered by the solver (v8:1309-1312)
    long long particleCount() const { return ifl_particles_count(_ctx); }
    void gridToParticles(double alpha) { ifl_host_check(ifl_grid_to_particles(_ctx, alpha), "ifl_grid_to_particles"); } // v8:904
    void particlesToGrid() {                                                                                             // v8:916
        long long n = 0;
        ifl_host_check(ifl_particles_to_grid(_ctx, &n), "ifl_particles_to_grid");
        printf("Particle count: %d\n", (int)n); // v8:926
    }
    void advect(double timestep) { ifl_host_check(ifl_particles_advect(_ctx, timestep), "ifl_particles_advect"); } // v8:931
};
#endif

class FluidSolver {
    ifl_ctx *_ctx;
    int _w, _h;
    double _density;
    const std::vector<const SolidBody *> *_bodies; // held by reference, like v4:612
    std::vector<double> _scratch;

    static void check(int rc, const char *what) { ifl_host_check(rc, what); }

    void create() {
        check(ifl_create(&_ctx, _w, _h, IFL_CHAPTER, 0), "ifl_create");
        _scratch.resize((size_t)_w * _h);
    }

    void syncBodies() {
        if (!_bodies) return;
        std::vector<ifl_body> rec;
        for (size_t i = 0; i < _bodies->size(); i++) rec.push_back((*_bodies)[i]->record());
        check(ifl_set_bodies(_ctx, rec.empty() ? NULL : &rec[0], (int)rec.size()), "ifl_set_bodies");
    }

public:
    FluidSolver(int w, int h, double density) : _ctx(NULL), _w(w), _h(h), _density(density), _bodies(NULL) { create(); }
    FluidSolver(int w, int h, double density, const std::vector<const SolidBody *> &bodies)
        : _ctx(NULL), _w(w), _h(h), _density(density), _bodies(&bodies) {
        create();
    }
    // chapters 6-8 (v6:921, v8:1289): air / soot densities and the heat diffusion coefficient; chapter 8 also
    // seeds the particles on the jittered grid and interpolates the initial fields onto them (v8:1306-1314)
    FluidSolver(int w, int h, double rhoAir, double rhoSoot, double diffusion, const std::vector<const SolidBody *> &bodies)
        : _ctx(NULL), _w(w), _h(h), _density(rhoAir), _bodies(&bodies) {
        create();
        check(ifl_set_fluid_params(_ctx, rhoAir, rhoSoot, diffusion), "ifl_set_fluid_params");
#if IFL_CHAPTER >= 8
        syncBodies();
        check(ifl_particles_init(_ctx, IFL_AVG_PER_CELL), "ifl_particles_init");
#endif
    }
    ~FluidSolver() { ifl_destroy(_ctx); }

    // the solver's quantities and particle set as reference-style objects (views, see the header comment)
    FluidQuantity quantity(int field) {
        syncBodies(); // the view's fillSolidFields / extrapolate / advect see the bodies as they are now (v4:612)
        const bool u = field == IFL_FIELD_U, v = field == IFL_FIELD_V;
        return FluidQuantity(_ctx, field, _w + (u ? 1 : 0), _h + (v ? 1 : 0), u ? 0.0 : 0.5, v ? 0.0 : 0.5); // v3:404-406
    }
#if IFL_CHAPTER >= 8
    ParticleQuantities particles() { return ParticleQuantities(_ctx); }
#endif
    double maxTimestep() { // v1:310-328
        double dt = 0.0;
        check(ifl_max_timestep(_ctx, &dt), "ifl_max_timestep");
        return dt;
    }

    void addInflow(double x, double y, double w, double h, double d, double t, double u, double v) { // v6:1010
        check(ifl_add_inflow_t(_ctx, x, y, w, h, d, t, u, v), "ifl_add_inflow_t");
    }
    double ambientT() { return ifl_ambient_t(_ctx); } // v6:1017

    // v6:1021-1060: smoke on the right half, black-body temperature ramp on the left half
    // (image is 2w x h) when renderHeat, else the chapter-5 picture
    void toImage(unsigned char *rgba, bool renderHeat) {
        const double *d = density();
        std::vector<double> vol((size_t)_w * _h), temp;
        check(ifl_aux_download(_ctx, IFL_FIELD_D, IFL_AUX_VOLUME, &vol[0]), "ifl_aux_download");
        if (renderHeat) {
            temp.resize((size_t)_w * _h);
            check(ifl_download(_ctx, IFL_BUF_T_SRC, &temp[0]), "ifl_download");
        }
        const double tAmb = ambientT();
        for (int y = 0; y < _h; y++)
            for (int x = 0; x < _w; x++) {
                const int i = x + y * _w;
                const int idxr = renderHeat ? 4 * (x + y * _w * 2 + _w) : 4 * i;
                const double volume = vol[i];
                double shade = (1.0 - d[i]) * volume;
                shade = std::min(std::max(shade, 0.0), 1.0);
                rgba[idxr + 0] = rgba[idxr + 1] = rgba[idxr + 2] = (unsigned char)(int)(shade * 255.0);
                rgba[idxr + 3] = 0xFF;
                if (renderHeat) {
                    const int idxl = 4 * (x + y * _w * 2);
#if IFL_CHAPTER >= 8
                    double t = fabs(temp[i] - tAmb) / 70.0; // v8:1452
#else
                    double t = (temp[i] - tAmb) / 700.0;
#endif
                    t = std::min(std::max(t, 0.0), 1.0);
                    const double r = 1.0 + volume * (std::min(t * 4.0, 1.0) - 1.0);
                    const double g = 1.0 + volume * (std::min(t * 2.0, 1.0) - 1.0);
                    const double b = 1.0 + volume * (std::max(std::min(t * 4.0 - 3.0, 1.0), 0.0) - 1.0);
                    rgba[idxl + 0] = (unsigned char)(int)(r * 255.0);
                    rgba[idxl + 1] = (unsigned char)(int)(g * 255.0);
                    rgba[idxl + 2] = (unsigned char)(int)(b * 255.0);
                    rgba[idxl + 3] = 0xFF;
                }
            }
    }

    static void report(const ifl_solve_info &info) { // the reference's own stdout lines
        const char *what = IFL_CHAPTER >= 3 ? "error" : "change";
        if (info.status == IFL_SOLVE_CONVERGED)
            printf("Exiting solver after %d iterations, maximum %s is %f\n", info.iterations, what, info.max_error);
        else if (info.status == IFL_SOLVE_EXCEEDED)
            printf("Exceeded budget of %d iterations, maximum %s was %f\n", info.iterations, what, info.max_error);
        else if (IFL_CHAPTER >= 6)
            printf("Initial guess sufficiently small\n"); // v6:835 (chapters 3-5 return silently, v3:355)
    }

    void update(double timestep) {
        syncBodies();
        ifl_solve_info info[2];
        check(ifl_update(_ctx, timestep, _density, info), "ifl_update");
#if IFL_CHAPTER >= 8
        printf("Particle count: %d\n", (int)ifl_particles_count(_ctx)); // particlesToGrid prints first (v8:926)
#endif
        report(info[0]);                     // chapters 6+: the heat solve (v6:974) ...
        if (IFL_CHAPTER >= 6) report(info[1]); // ... then the pressure solve (v6:990)
    }

    void addInflow(double x, double y, double w, double h, double d, double u, double v) {
        check(ifl_add_inflow(_ctx, x, y, w, h, d, u, v), "ifl_add_inflow");
    }

    // density field, reference layout (FluidQuantity::src() of _d)
    const double *density() {
        check(ifl_download(_ctx, IFL_BUF_D_SRC, &_scratch[0]), "ifl_download");
        return &_scratch[0];
    }

    void toImage(unsigned char *rgba) { // v3:455-465 (v5:961-972 multiplies by the fluid volume)
        const double *d = density();
        std::vector<double> vol;
        if (IFL_CHAPTER >= 5) {
            vol.resize((size_t)_w * _h);
            check(ifl_aux_download(_ctx, IFL_FIELD_D, IFL_AUX_VOLUME, &vol[0]), "ifl_aux_download");
        }
        std::vector<unsigned char> cell;
        if (IFL_CHAPTER == 4) {
            cell.resize((size_t)_w * _h);
            check(ifl_aux_download(_ctx, IFL_FIELD_D, IFL_AUX_CELL, &cell[0]), "ifl_aux_download");
        }
        for (int i = 0; i < _w * _h; i++) {
            int shade = IFL_CHAPTER >= 5 ? (int)((1.0 - d[i]) * vol[i] * 255.0) : (int)((1.0 - d[i]) * 255.0);
            shade = std::max(std::min(shade, 255), 0);
            if (IFL_CHAPTER == 4 && cell[i] == 1) shade = 0; // v4:907-908
            rgba[i * 4 + 0] = rgba[i * 4 + 1] = rgba[i * 4 + 2] = (unsigned char)shade;
            rgba[i * 4 + 3] = 0xFF;
        }
    }
};
