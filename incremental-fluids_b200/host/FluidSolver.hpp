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
//
// The bodies of the reference's private hot-path methods live on the GPU; this header
// only forwards (one opaque ifl_ctx per solver) and prints the reference's own status
// lines ("Exiting solver after %d iterations, ...", v3:368, v3:379).  Select the chapter
// with -DIFL_CHAPTER=1..7 (default 3), exactly as one would pick a reference directory.
#pragma once

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../include/ifl_b200.h"

#ifndef IFL_CHAPTER
#define IFL_CHAPTER 3
#endif

class SolidBody { // v4:79-149
protected:
    double _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta;
    int _kind;
    SolidBody(int kind, double posX, double posY, double scaleX, double scaleY, double theta, double velX, double velY,
              double velTheta)
        : _posX(posX), _posY(posY), _scaleX(scaleX), _scaleY(scaleY), _theta(theta), _velX(velX), _velY(velY),
          _velTheta(velTheta), _kind(kind) {}

public:
    virtual ~SolidBody() {}
    double velocityX(double, double y) const { return (_posY - y) * _velTheta + _velX; }
    double velocityY(double x, double) const { return (x - _posX) * _velTheta + _velY; }
    void update(double timestep) {
        _posX += _velX * timestep;
        _posY += _velY * timestep;
        _theta += _velTheta * timestep;
    }
    ifl_body record() const {
        ifl_body b = {_kind, _posX, _posY, _scaleX, _scaleY, _theta, _velX, _velY, _velTheta};
        return b;
    }
};

class SolidBox : public SolidBody { // v4:152-157
public:
    SolidBox(double x, double y, double sx, double sy, double t, double vx, double vy, double vt)
        : SolidBody(0, x, y, sx, sy, t, vx, vy, vt) {}
};

class SolidSphere : public SolidBody { // v4:204-209
public:
    SolidSphere(double x, double y, double s, double t, double vx, double vy, double vt)
        : SolidBody(1, x, y, s, s, t, vx, vy, vt) {}
};

class FluidSolver {
    ifl_ctx *_ctx;
    int _w, _h;
    double _density;
    const std::vector<const SolidBody *> *_bodies; // held by reference, like v4:612
    std::vector<double> _scratch;

    static void check(int rc, const char *what) {
        if (rc != IFL_OK) {
            fprintf(stderr, "libifl_b200: %s failed (%d): %s\n", what, rc, ifl_last_error());
            abort(); // the reference has no error channel either
        }
    }

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
    // chapters 6-7 (v6:921): air / soot densities and the heat diffusion coefficient
    FluidSolver(int w, int h, double rhoAir, double rhoSoot, double diffusion, const std::vector<const SolidBody *> &bodies)
        : _ctx(NULL), _w(w), _h(h), _density(rhoAir), _bodies(&bodies) {
        create();
        check(ifl_set_fluid_params(_ctx, rhoAir, rhoSoot, diffusion), "ifl_set_fluid_params");
    }
    ~FluidSolver() { ifl_destroy(_ctx); }

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
                    double t = (temp[i] - tAmb) / 700.0;
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
        report(info[0]);                     // chapters 6-7: the heat solve (v6:974) ...
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
