/*
 * oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference translation unit
 *     /root/reference/<N-name>/Fluid.cpp          (passed as -DREF_SRC="...")
 * into a shared object oracle/_ref/libref_v<N>.so and exposes its private
 * methods and arrays through a tiny name-keyed C API, so tests can call the
 * reference's own buildRhs()/project()/advect()/... at any grid size with the
 * reference's exact arithmetic.  No reference source is copied into this
 * repository: the file is #included from where it lies.
 *
 * Technique: pre-include every std header the reference
 * uses (include guards make its own #includes no-ops), then
 *   #define class struct      -> all members public (the reference never
 *                                writes "private:", privacy is class-default)
 *   #define main  ref_main    -> the shipped driver becomes a dead function
 *   #define printf(...)       -> solver status lines are captured, not printed
 * and include the file inside a per-version namespace.
 *
 * Build flags must match the reference Makefile (-O2, no -march=native, no
 * FMA contraction): see oracle/Makefile.
 */
#define _USE_MATH_DEFINES
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <string>
#include <stack>

#ifndef REF_VER
#error "compile with -DREF_VER=<1..8> -DREF_SRC=\"/root/reference/.../Fluid.cpp\""
#endif

/* lodepng is only referenced by the dead driver; give it a stub. */
#ifndef LODEPNG_H
#define LODEPNG_H
static unsigned lodepng_encode32_file(const char *, const unsigned char *, unsigned, unsigned) { return 0; }
#endif

static std::string g_log;
static int ref_capture(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_log += buf;
    return n;
}

#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_NS REF_CAT(ref_v, REF_VER)

/* The reference reads memory it never initialised (`new double[]` without a fill: _dst, _z, _s, _precon,
 * the particle arrays past _particleCount, ...).  At the shipped sizes glibc hands out
 * fresh zero pages for those blocks; the harness pins that behaviour at every size by making every
 * allocation of this library zero-filled (the library is built with hidden visibility, so these
 * replacements serve only the code in it). */
#include <new>
void *operator new[](size_t n) {
    void *p = calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new(size_t n) {
    void *p = calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void *p) noexcept { free(p); }
void operator delete(void *p) noexcept { free(p); }
void operator delete[](void *p, size_t) noexcept { free(p); }
void operator delete(void *p, size_t) noexcept { free(p); }

#define main ref_main
#define class struct
#define protected public /* SolidBody spells out "protected:" (v4:80) */
#define printf(...) ref_capture(__VA_ARGS__)
namespace REF_NS {
#include REF_SRC
}
#undef printf
#undef protected
#undef class
#undef main

using namespace REF_NS;

struct Handle {
    FluidSolver *s;
#if REF_VER >= 4
    std::vector<SolidBody *> bodies;
    std::vector<const SolidBody *> cbodies;
#endif
    int w, h;
};

static FluidQuantity *quantity(Handle *hd, const std::string &q) {
    if (q == "d") return hd->s->_d;
    if (q == "u") return hd->s->_u;
    if (q == "v") return hd->s->_v;
#if REF_VER >= 6
    if (q == "t") return hd->s->_t;
#endif
    return 0;
}

#define API extern "C" __attribute__((visibility("default")))

API int ref_version() { return REF_VER; }

/* params: v1-5 {density}; v6-8 {rhoAir, rhoSoot, diffusion}.
 * bodies: nbodies x 9 doubles {kind(0 box, 1 sphere), x, y, sx, sy, theta, vx, vy, vtheta}. */
API void *ref_create(int w, int h, const double *params, int nparams, const double *bodies, int nbodies) {
    Handle *hd = new Handle();
    hd->w = w;
    hd->h = h;
    (void)nparams; (void)bodies; (void)nbodies;
#if REF_VER >= 4
    for (int i = 0; i < nbodies; i++) {
        const double *b = bodies + 9*i;
        if (b[0] == 0.0)
            hd->bodies.push_back(new SolidBox(b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]));
        else
            hd->bodies.push_back(new SolidSphere(b[1], b[2], b[3], b[5], b[6], b[7], b[8]));
    }
    for (size_t i = 0; i < hd->bodies.size(); i++)
        hd->cbodies.push_back(hd->bodies[i]);
#endif
#if REF_VER <= 3
    hd->s = new FluidSolver(w, h, params[0]);
#elif REF_VER <= 5
    hd->s = new FluidSolver(w, h, params[0], hd->cbodies);
#else
    hd->s = new FluidSolver(w, h, params[0], params[1], params[2], hd->cbodies);
#endif
    return hd;
}

API void ref_destroy(void *p) {
    Handle *hd = (Handle *)p;
#if REF_VER < 8
    delete hd->s;
#else
    /* ~ParticleQuantities (v8:887-888) delete[]s its FluidQuantity pointers instead of its
     * property arrays -- undefined behaviour that aborts under glibc.  The reference's main()
     * never destroys its solver either; leak it. */
#endif
    /* SolidBody has a protected virtual destructor in the reference; leak the few bytes. */
    delete hd;
}

/* Returns a pointer INTO the reference object's storage (read/write). */
API void *ref_buf(void *p, const char *name, long *count, int *elsize) {
    Handle *hd = (Handle *)p;
    FluidSolver *s = hd->s;
    std::string n(name);
    long C = (long)hd->w*hd->h;
    *elsize = 8;
    *count = C;
    if (n == "r") return s->_r;
    if (n == "p") return s->_p;
#if REF_VER >= 3
    if (n == "z") return s->_z;
    if (n == "s") return s->_s;
    if (n == "precon") return s->_precon;
    if (n == "aDiag") return s->_aDiag;
    if (n == "aPlusX") return s->_aPlusX;
    if (n == "aPlusY") return s->_aPlusY;
#endif
#if REF_VER >= 7
    if (n == "uDensity") { *count = (long)(hd->w + 1)*hd->h; return s->_uDensity; }
    if (n == "vDensity") { *count = (long)hd->w*(hd->h + 1); return s->_vDensity; }
#endif
#if REF_VER >= 8
    if (n.compare(0, 3, "qs.") == 0) {
        ParticleQuantities *qs = s->_qs;
        std::string f = n.substr(3);
        *count = qs->_maxParticles;
        if (f == "posX") return qs->_posX;
        if (f == "posY") return qs->_posY;
        if (f == "weight") { *count = (long)(hd->w + 1)*(hd->h + 1); return qs->_weight; }
        if (f == "counts") { *count = C; *elsize = 4; return qs->_counts; }
        if (f.compare(0, 4, "prop") == 0) {
            size_t k = (size_t)atoi(f.c_str() + 4);
            if (k < qs->_properties.size()) return qs->_properties[k];
        }
        return 0;
    }
#endif
    size_t dot = n.find('.');
    if (dot == std::string::npos) return 0;
    FluidQuantity *q = quantity(hd, n.substr(0, dot));
    if (!q) return 0;
    std::string f = n.substr(dot + 1);
    *count = (long)q->_w*q->_h;
    if (f == "src") return q->_src;
#if REF_VER <= 7
    if (f == "dst") return q->_dst;
#else
    if (f == "old") return q->_old;
#endif
#if REF_VER >= 4
    if (f == "normalX") return q->_normalX;
    if (f == "normalY") return q->_normalY;
    if (f == "cell") { *elsize = 1; return q->_cell; }
    if (f == "body") { *elsize = 1; return q->_body; }
    if (f == "mask") { *elsize = 1; return q->_mask; }
#endif
#if REF_VER >= 5
    if (f == "volume") return q->_volume;
    if (f == "phi") { *count = (long)(q->_w + 1)*(q->_h + 1); return q->_phi; }
#endif
    return 0;
}

API const char *ref_log() { return g_log.c_str(); }
API void ref_log_clear() { g_log.clear(); }

static double *vec(Handle *hd, double id) {
    /* vector ids for the granular PCG ops: 0 r, 1 p, 2 z, 3 s */
    FluidSolver *s = hd->s;
    switch ((int)id) {
    case 0: return s->_r;
    case 1: return s->_p;
#if REF_VER >= 3
    case 2: return s->_z;
    case 3: return s->_s;
#endif
    }
    return 0;
}

/* Name-keyed dispatcher onto the reference's own (private) methods.
 * Returns 0 on success, -1 for an op this version does not have. */
API int ref_call(void *p, const char *opname, const double *a, int na, double *out) {
    Handle *hd = (Handle *)p;
    FluidSolver *s = hd->s;
    std::string op(opname);
    (void)na; (void)vec;

    if (op == "update") { s->update(a[0]); return 0; }
    if (op == "buildRhs") { s->buildRhs(); return 0; }
    if (op == "applyPressure") { s->applyPressure(a[0]); return 0; }
    if (op == "addInflow") {
#if REF_VER <= 5
        s->addInflow(a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
#else
        s->addInflow(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
#endif
        return 0;
    }
#if REF_VER <= 2
    if (op == "project") { s->project((int)a[0], a[1]); return 0; }
#else
    if (op == "project") { s->project((int)a[0]); return 0; }
    if (op == "buildPressureMatrix") { s->buildPressureMatrix(a[0]); return 0; }
    if (op == "buildPreconditioner") { s->buildPreconditioner(); return 0; }
    if (op == "applyPreconditioner") { s->applyPreconditioner(vec(hd, a[0]), vec(hd, a[1])); return 0; }
    if (op == "matrixVectorProduct") { s->matrixVectorProduct(vec(hd, a[0]), vec(hd, a[1])); return 0; }
    if (op == "dotProduct") { out[0] = s->dotProduct(vec(hd, a[0]), vec(hd, a[1])); return 0; }
    if (op == "scaledAdd") { s->scaledAdd(vec(hd, a[0]), vec(hd, a[1]), vec(hd, a[2]), a[3]); return 0; }
    if (op == "infinityNorm") { out[0] = s->infinityNorm(vec(hd, a[0])); return 0; }
#endif
#if REF_VER >= 4
    if (op == "setBoundaryCondition") { s->setBoundaryCondition(); return 0; }
    if (op == "bodiesUpdate") {
        for (size_t i = 0; i < hd->bodies.size(); i++) hd->bodies[i]->update(a[0]);
        return 0;
    }
    if (op == "bodyGeometry") { /* a = {index, x, y} -> out = {distance, normalX, normalY, closestX, closestY} (v4:116-118) */
        const SolidBody *b = hd->bodies[(int)a[0]];
        double x = a[1], y = a[2];
        out[0] = b->distance(x, y);
        b->distanceNormal(out[1], out[2], x, y);
        b->closestSurfacePoint(x, y);
        out[3] = x; out[4] = y;
        return 0;
    }
    if (op == "bodyState") { /* a[0] = index -> out[0..7] */
        SolidBody *b = hd->bodies[(int)a[0]];
        out[0] = b->_posX; out[1] = b->_posY; out[2] = b->_scaleX; out[3] = b->_scaleY;
        out[4] = b->_theta; out[5] = b->_velX; out[6] = b->_velY; out[7] = b->_velTheta;
        return 0;
    }
#endif
#if REF_VER >= 6
    if (op == "buildHeatDiffusionMatrix") { s->buildHeatDiffusionMatrix(a[0]); return 0; }
    if (op == "addBuoyancy") { s->addBuoyancy(a[0]); return 0; }
    if (op == "ambientT") { out[0] = s->ambientT(); return 0; }
#endif
#if REF_VER >= 7
    if (op == "computeDensities") { s->computeDensities(); return 0; }
#endif
#if REF_VER >= 8
    {
        ParticleQuantities *qs = s->_qs;
        if (op == "qs.particleCount") { out[0] = qs->_particleCount; return 0; }
        if (op == "qs.setParticleCount") { qs->_particleCount = (int)a[0]; return 0; }
        if (op == "qs.particlesToGrid") { qs->particlesToGrid(); return 0; }
        if (op == "qs.countParticles") { qs->countParticles(); return 0; }
        if (op == "qs.pruneParticles") { qs->pruneParticles(); return 0; }
        if (op == "qs.seedParticles") { qs->seedParticles(); return 0; }
        if (op == "qs.gridToParticles") { qs->gridToParticles(a[0]); return 0; }
        if (op == "qs.advect") { qs->advect(a[0], *s->_u, *s->_v); return 0; }
    }
#endif

    /* per-quantity ops: "<q>.<method>" */
    size_t dot = op.find('.');
    if (dot != std::string::npos) {
        FluidQuantity *q = quantity(hd, op.substr(0, dot));
        std::string m = op.substr(dot + 1);
        if (!q) return -1;
#if REF_VER <= 7
        if (m == "flip") { q->flip(); return 0; }
        if (m == "advect") {
#if REF_VER <= 3
            q->advect(a[0], *s->_u, *s->_v);
#else
            q->advect(a[0], *s->_u, *s->_v, hd->cbodies);
#endif
            return 0;
        }
#if REF_VER >= 2
        if (m == "cerp") { out[0] = q->cerp(a[0], a[1]); return 0; }
#endif
#endif
        if (m == "lerp") { out[0] = q->lerp(a[0], a[1]); return 0; }
        if (m == "addInflow") { q->addInflow(a[0], a[1], a[2], a[3], a[4]); return 0; }
#if REF_VER >= 4
        if (m == "fillSolidFields") { q->fillSolidFields(hd->cbodies); return 0; }
        if (m == "extrapolate") { q->extrapolate(); return 0; }
#endif
#if REF_VER >= 8
        if (m == "copy") { q->copy(); return 0; }
        if (m == "diff") { q->diff(a[0]); return 0; }
        if (m == "undiff") { q->undiff(a[0]); return 0; }
        if (m == "fromParticles") {
            ParticleQuantities *qs = s->_qs;
            q->fromParticles(qs->_weight, qs->_particleCount, qs->_posX, qs->_posY, qs->_properties[(int)a[0]]);
            return 0;
        }
#endif
    }
    return -1;
}
