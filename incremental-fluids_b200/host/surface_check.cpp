// surface_check.cpp -- exercises the parts of the reference's class surface that main() does not touch
// (FluidQuantity::src/at/lerp/cerp/cell/volume, SolidBody::distance/closestSurfacePoint/distanceNormal,
// FluidSolver::maxTimestep, ParticleQuantities) through the drop-in header and prints the results as
// hexadecimal doubles; tests/test_gpu_dropin.py compares them with the unmodified reference.
//     surface_v<N> [size=96]
#include <stdint.h>
#include <string.h>

#include "FluidSolver.hpp"

static void show(const char *name, double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    printf("%s=%016llx\n", name, (unsigned long long)b);
}

int main(int argc, char **argv) {
    const int size = argc > 1 ? atoi(argv[1]) : 96;
    std::vector<SolidBody *> bodies;
    bodies.push_back(new SolidBox(0.5, 0.6, 0.7, 0.1, M_PI * 0.25, 0.0, 0.0, 0.0));
    bodies.push_back(new SolidSphere(0.2, 0.3, 0.2, 0.3, 0.0, 0.0, 0.0));
    std::vector<const SolidBody *> cBodies;
    for (unsigned i = 0; i < bodies.size(); i++) cBodies.push_back(bodies[i]);
#if IFL_CHAPTER >= 6
    FluidSolver solver(size, size, 0.1, 0.25, 0.01, cBodies);
#else
    FluidSolver solver(size, size, 0.1, cBodies);
#endif
    // SolidBody virtuals on the host (v4:116-118)
    const double pts[4][2] = {{0.31, 0.42}, {0.5, 0.61}, {0.22, 0.33}, {0.9, 0.1}};
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 4; k++) {
            char name[64];
            double x = pts[k][0], y = pts[k][1], nx, ny;
            snprintf(name, sizeof name, "distance[%d][%d]", b, k);
            show(name, cBodies[b]->distance(x, y));
            cBodies[b]->distanceNormal(nx, ny, x, y);
            snprintf(name, sizeof name, "normalX[%d][%d]", b, k);
            show(name, nx);
            snprintf(name, sizeof name, "normalY[%d][%d]", b, k);
            show(name, ny);
            cBodies[b]->closestSurfacePoint(x, y);
            snprintf(name, sizeof name, "closestX[%d][%d]", b, k);
            show(name, x);
            snprintf(name, sizeof name, "closestY[%d][%d]", b, k);
            show(name, y);
        }
    // FluidQuantity views
#if IFL_CHAPTER >= 6
    solver.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, solver.ambientT(), 0.5, 3.0);
#else
    solver.addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.5, 3.0);
#endif
    FluidQuantity d = solver.quantity(IFL_FIELD_D), v = solver.quantity(IFL_FIELD_V);
    const double hx = 1.0 / size;
    const double sx[3] = {0.47 / hx, 0.52 / hx + 0.37, 0.58 / hx};
    const double sy[3] = {0.205 / hx, 0.21 / hx + 0.41, 0.22 / hx};
    for (int k = 0; k < 3; k++) {
        char name[64];
        snprintf(name, sizeof name, "d.lerp[%d]", k);
        show(name, d.lerp(sx[k], sy[k]));
        snprintf(name, sizeof name, "v.lerp[%d]", k);
        show(name, v.lerp(sx[k], sy[k]));
#if IFL_CHAPTER <= 7
        snprintf(name, sizeof name, "d.cerp[%d]", k);
        show(name, d.cerp(sx[k], sy[k]));
#endif
    }
    show("d.at", d.at((int)sx[1], (int)sy[1]));
    show("maxTimestep", solver.maxTimestep());
    d.fillSolidFields();
    const unsigned char *cell = d.cell();
    int solid = 0;
    for (int i = 0; i < size * size; i++) solid += cell[i] == 1;
    printf("solid_cells=%d\n", solid);
#if IFL_CHAPTER >= 5
    show("d.volume", d.volume((int)(0.5 / hx), (int)(0.6 / hx) - 4));
#endif
#if IFL_CHAPTER >= 8
    ParticleQuantities qs = solver.particles();
    printf("particles=%lld\n", qs.particleCount());
    solver.update(0.0025);
    printf("particles_after_update=%lld\n", qs.particleCount());
#endif
    return 0;
}
