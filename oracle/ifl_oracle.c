/*
 * ifl_oracle.c -- CPU restatement of the reference's hot path (see ifl_oracle.h).
 * TEST INFRASTRUCTURE ONLY; parity PINNED against the unmodified reference.
 *
 * The restatement is deliberately sequential and follows the reference's evaluation
 * order statement by statement; it exists so that the GPU box (which has no
 * /root/reference) can check libifl_b200.so against reference semantics.
 */
#include "ifl_oracle.h"

#include <math.h>
#include <string.h>

/* std::min / std::max as libstdc++ defines them (first argument wins ties and NaNs) */
static double dmin(double a, double b) { return (b < a) ? b : a; }
static double dmax(double a, double b) { return (a < b) ? b : a; }
static int imin_(int a, int b) { return (b < a) ? b : a; }
static int imax_(int a, int b) { return (a < b) ? b : a; }

static double at(const double *f, ofl_grid g, int x, int y) { return f[x + y * g.w]; }

/* 1-D helpers: v2:58-60 and v2:66-80 */
static double lerp1(double a, double b, double x) { return a * (1.0 - x) + b * x; }

static double cerp1(double a, double b, double c, double d, double x) {
    double xsq = x * x;
    double xcu = xsq * x;
    double lo = dmin(a, dmin(b, dmin(c, d)));
    double hi = dmax(a, dmax(b, dmax(c, d)));
    double t = a * (0.0 - 0.5 * x + 1.0 * xsq - 0.5 * xcu) + b * (1.0 + 0.0 * x - 2.5 * xsq + 1.5 * xcu) +
               c * (0.0 + 0.5 * x + 2.0 * xsq - 1.5 * xcu) + d * (0.0 + 0.0 * x - 0.5 * xsq + 0.5 * xcu);
    return dmin(dmax(t, lo), hi);
}

double ofl_lerp(const double *f, ofl_grid g, double x, double y) { /* v2:133-145 */
    int ix, iy;
    x = dmin(dmax(x - g.ox, 0.0), g.w - 1.001);
    y = dmin(dmax(y - g.oy, 0.0), g.h - 1.001);
    ix = (int)x;
    iy = (int)y;
    x -= ix;
    y -= iy;
    return lerp1(lerp1(at(f, g, ix, iy), at(f, g, ix + 1, iy), x),
                 lerp1(at(f, g, ix, iy + 1), at(f, g, ix + 1, iy + 1), x), y);
}

double ofl_cerp(const double *f, ofl_grid g, double x, double y) { /* v2:150-167 */
    int ix, iy, xs[4], ys[4], j;
    double q[4];
    x = dmin(dmax(x - g.ox, 0.0), g.w - 1.001);
    y = dmin(dmax(y - g.oy, 0.0), g.h - 1.001);
    ix = (int)x;
    iy = (int)y;
    x -= ix;
    y -= iy;
    xs[0] = imax_(ix - 1, 0); xs[1] = ix; xs[2] = ix + 1; xs[3] = imin_(ix + 2, g.w - 1);
    ys[0] = imax_(iy - 1, 0); ys[1] = iy; ys[2] = iy + 1; ys[3] = imin_(iy + 2, g.h - 1);
    for (j = 0; j < 4; j++)
        q[j] = cerp1(at(f, g, xs[0], ys[j]), at(f, g, xs[1], ys[j]), at(f, g, xs[2], ys[j]), at(f, g, xs[3], ys[j]), x);
    return cerp1(q[0], q[1], q[2], q[3], y);
}

void ofl_advect(int mode, double *dst, const double *src, ofl_grid g, const double *u, ofl_grid gu,
                const double *v, ofl_grid gv, double timestep, double hx) {
    int ix, iy, idx = 0;
    for (iy = 0; iy < g.h; iy++) {
        for (ix = 0; ix < g.w; ix++, idx++) {
            double x = ix + g.ox;
            double y = iy + g.oy;
            if (mode == 1) {
                /* rungeKutta3 v2:83-101; the third stage is NOT divided by hx */
                double firstU = ofl_lerp(u, gu, x, y) / hx;
                double firstV = ofl_lerp(v, gv, x, y) / hx;
                double midX = x - 0.5 * timestep * firstU;
                double midY = y - 0.5 * timestep * firstV;
                double midU = ofl_lerp(u, gu, midX, midY) / hx;
                double midV = ofl_lerp(v, gv, midX, midY) / hx;
                double lastX = x - 0.75 * timestep * midU;
                double lastY = y - 0.75 * timestep * midV;
                double lastU = ofl_lerp(u, gu, lastX, lastY);
                double lastV = ofl_lerp(v, gv, lastX, lastY);
                x -= timestep * ((2.0 / 9.0) * firstU + (3.0 / 9.0) * midU + (4.0 / 9.0) * lastU);
                y -= timestep * ((2.0 / 9.0) * firstV + (3.0 / 9.0) * midV + (4.0 / 9.0) * lastV);
                dst[idx] = ofl_cerp(src, g, x, y);
            } else {
                /* euler v1:68-74 + bilinear v1:133-136 */
                double uVel = ofl_lerp(u, gu, x, y) / hx;
                double vVel = ofl_lerp(v, gv, x, y) / hx;
                x -= uVel * timestep;
                y -= vVel * timestep;
                dst[idx] = ofl_lerp(src, g, x, y);
            }
        }
    }
}

void ofl_add_inflow(double *src, ofl_grid g, double hx, double x0, double y0, double x1, double y1, double v,
                    int smooth) {
    int ix0 = (int)(x0 / hx - g.ox);
    int iy0 = (int)(y0 / hx - g.oy);
    int ix1 = (int)(x1 / hx - g.ox);
    int iy1 = (int)(y1 / hx - g.oy);
    int x, y;
    for (y = imax_(iy0, 0); y < imin_(iy1, g.h); y++) {
        for (x = imax_(ix0, 0); x < imin_(ix1, g.h); x++) { /* sic: _h, v2:195 */
            double vi = v;
            if (smooth) {
                double lx = (2.0 * (x + 0.5) * hx - (x0 + x1)) / (x1 - x0);
                double ly = (2.0 * (y + 0.5) * hx - (y0 + y1)) / (y1 - y0);
                double l = sqrt(lx * lx + ly * ly);
                double c = dmin(fabs(l), 1.0);
                vi = (1.0 - c * c * (3.0 - 2.0 * c)) * v;
            }
            if (fabs(src[x + y * g.w]) < fabs(vi)) src[x + y * g.w] = vi;
        }
    }
}

void ofl_build_rhs(double *r, const double *u, const double *v, int w, int h, double hx) {
    double scale = 1.0 / hx;
    int x, y, idx = 0;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++, idx++)
            r[idx] = -scale * (u[x + 1 + y * (w + 1)] - u[x + y * (w + 1)] + v[x + (y + 1) * w] - v[x + y * w]);
}

void ofl_build_pressure_matrix(double *aDiag, double *aPlusX, double *aPlusY, int w, int h, double timestep,
                               double density, double hx) {
    double scale = timestep / (density * hx * hx);
    int x, y, idx = 0;
    memset(aDiag, 0, (size_t)w * h * sizeof(double));
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++, idx++) {
            if (x < w - 1) {
                aDiag[idx] += scale;
                aDiag[idx + 1] += scale;
                aPlusX[idx] = -scale;
            } else
                aPlusX[idx] = 0.0;
            if (y < h - 1) {
                aDiag[idx] += scale;
                aDiag[idx + w] += scale;
                aPlusY[idx] = -scale;
            } else
                aPlusY[idx] = 0.0;
        }
    }
}

void ofl_build_preconditioner(double *precon, const double *aDiag, const double *aPlusX, const double *aPlusY,
                              int w, int h) {
    const double tau = 0.97, sigma = 0.25;
    int x, y, idx = 0;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++, idx++) {
            double e = aDiag[idx];
            if (x > 0) {
                double px = aPlusX[idx - 1] * precon[idx - 1];
                double py = aPlusY[idx - 1] * precon[idx - 1];
                e = e - (px * px + tau * px * py);
            }
            if (y > 0) {
                double px = aPlusX[idx - w] * precon[idx - w];
                double py = aPlusY[idx - w] * precon[idx - w];
                e = e - (py * py + tau * px * py);
            }
            if (e < sigma * aDiag[idx]) e = aDiag[idx];
            precon[idx] = 1.0 / sqrt(e);
        }
    }
}

void ofl_apply_preconditioner(double *dst, const double *a, const double *precon, const double *aPlusX,
                              const double *aPlusY, int w, int h) {
    int x, y, idx = 0;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++, idx++) {
            double t = a[idx];
            if (x > 0) t -= aPlusX[idx - 1] * precon[idx - 1] * dst[idx - 1];
            if (y > 0) t -= aPlusY[idx - w] * precon[idx - w] * dst[idx - w];
            dst[idx] = t * precon[idx];
        }
    }
    for (y = h - 1; y >= 0; y--) {
        for (x = w - 1; x >= 0; x--) {
            double t;
            idx = x + y * w;
            t = dst[idx];
            if (x < w - 1) t -= aPlusX[idx] * precon[idx] * dst[idx + 1];
            if (y < h - 1) t -= aPlusY[idx] * precon[idx] * dst[idx + w];
            dst[idx] = t * precon[idx];
        }
    }
}

double ofl_dot_product(const double *a, const double *b, int n) {
    double result = 0.0;
    int i;
    for (i = 0; i < n; i++) result += a[i] * b[i];
    return result;
}

void ofl_matrix_vector_product(double *dst, const double *b, const double *aDiag, const double *aPlusX,
                               const double *aPlusY, int w, int h) {
    int x, y, idx = 0;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++, idx++) {
            double t = aDiag[idx] * b[idx];
            if (x > 0) t += aPlusX[idx - 1] * b[idx - 1];
            if (y > 0) t += aPlusY[idx - w] * b[idx - w];
            if (x < w - 1) t += aPlusX[idx] * b[idx + 1];
            if (y < h - 1) t += aPlusY[idx] * b[idx + w];
            dst[idx] = t;
        }
    }
}

void ofl_scaled_add(double *dst, const double *a, const double *b, double s, int n) {
    int i;
    for (i = 0; i < n; i++) dst[i] = a[i] + b[i] * s;
}

double ofl_infinity_norm(const double *a, int n) {
    double m = 0.0;
    int i;
    for (i = 0; i < n; i++) m = dmax(m, fabs(a[i]));
    return m;
}

int ofl_project(int limit, double *p, double *r, double *z, double *s, const double *precon, const double *aDiag,
                const double *aPlusX, const double *aPlusY, int w, int h, int *iters, double *max_error) {
    int n = w * h, iter;
    double maxError, sigma;
    memset(p, 0, (size_t)n * sizeof(double));
    ofl_apply_preconditioner(z, r, precon, aPlusX, aPlusY, w, h);
    memcpy(s, z, (size_t)n * sizeof(double));
    maxError = ofl_infinity_norm(r, n);
    *max_error = maxError;
    *iters = 0;
    if (maxError < 1e-5) return 2;
    sigma = ofl_dot_product(z, r, n);
    for (iter = 0; iter < limit; iter++) {
        double alpha, sigmaNew;
        ofl_matrix_vector_product(z, s, aDiag, aPlusX, aPlusY, w, h);
        alpha = sigma / ofl_dot_product(z, s, n);
        ofl_scaled_add(p, p, s, alpha, n);
        ofl_scaled_add(r, r, z, -alpha, n);
        maxError = ofl_infinity_norm(r, n);
        *max_error = maxError;
        if (maxError < 1e-5) {
            *iters = iter;
            return 0;
        }
        ofl_apply_preconditioner(z, r, precon, aPlusX, aPlusY, w, h);
        sigmaNew = ofl_dot_product(z, r, n);
        ofl_scaled_add(s, z, s, sigmaNew / sigma, n);
        sigma = sigmaNew;
    }
    *iters = limit;
    return 1;
}

int ofl_project_gs(int limit, double timestep, double density, double hx, double *p, const double *r, int w,
                   int h, int *iters, double *max_delta) {
    double scale = timestep / (density * hx * hx);
    double maxDelta = 0.0;
    int iter, x, y;
    for (iter = 0; iter < limit; iter++) {
        maxDelta = 0.0;
        for (y = 0; y < h; y++) {
            for (x = 0; x < w; x++) {
                int idx = x + y * w;
                double diag = 0.0, offDiag = 0.0, newP;
                if (x > 0) {
                    diag += scale;
                    offDiag -= scale * p[idx - 1];
                }
                if (y > 0) {
                    diag += scale;
                    offDiag -= scale * p[idx - w];
                }
                if (x < w - 1) {
                    diag += scale;
                    offDiag -= scale * p[idx + 1];
                }
                if (y < h - 1) {
                    diag += scale;
                    offDiag -= scale * p[idx + w];
                }
                newP = (r[idx] - offDiag) / diag;
                maxDelta = dmax(maxDelta, fabs(p[idx] - newP));
                p[idx] = newP;
            }
        }
        if (maxDelta < 1e-5) {
            *iters = iter;
            *max_delta = maxDelta;
            return 0;
        }
    }
    *iters = limit;
    *max_delta = maxDelta;
    return 1;
}

void ofl_apply_pressure(double *u, double *v, const double *p, int w, int h, double timestep, double density,
                        double hx) {
    double scale = timestep / (density * hx);
    int x, y, idx = 0;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++, idx++) {
            u[x + y * (w + 1)] -= scale * p[idx];
            u[x + 1 + y * (w + 1)] += scale * p[idx];
            v[x + y * w] -= scale * p[idx];
            v[x + (y + 1) * w] += scale * p[idx];
        }
    }
    for (y = 0; y < h; y++) u[y * (w + 1)] = u[w + y * (w + 1)] = 0.0;
    for (x = 0; x < w; x++) v[x] = v[x + h * w] = 0.0;
}
