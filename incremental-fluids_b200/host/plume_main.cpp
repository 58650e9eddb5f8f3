// plume_main.cpp -- the reference's main() (3-conjugate-gradients/Fluid.cpp:468-499,
// 5-curved-boundaries/Fluid.cpp:975-1016) written against the drop-in header; instead of
// PNG frames it prints an FNV-1a-64 of the RGBA image every frame.  Usage:
//     plume_v<N> [size=128] [frames=5]
#include <stdint.h>
#include <string.h>

#include "FluidSolver.hpp"

int main(int argc, char **argv) {
    const int size = argc > 1 ? atoi(argv[1]) : 128;
    const int frames = argc > 2 ? atoi(argv[2]) : 5;
    const double density = 0.1, timestep = IFL_CHAPTER == 8 ? 0.0025 : 0.005; // v8:1478
    unsigned char *image = new unsigned char[(size_t)size * size * 4 * 2];
    size_t image_bytes = (size_t)size * size * 4;
    (void)density;
#if IFL_CHAPTER >= 6
    // 6-heat/Fluid.cpp:1062-1098 (7-variable-density: densitySoot 1.0, box and inflow of v7:1099/1112)
    const double densityAir = 0.1, densitySoot = IFL_CHAPTER == 8 ? 0.25 : IFL_CHAPTER == 7 ? 1.0 : 0.1, diffusion = 0.01; // v8:1475
    const bool renderHeat = true;
    image_bytes *= 2;
    std::vector<SolidBody *> bodies;
#if IFL_CHAPTER >= 7
    bodies.push_back(new SolidBox(0.5, 0.6, 0.7, 0.1, M_PI * 0.25, 0.0, 0.0, 0.0)); // v7:1099, v8:1484
#else
    bodies.push_back(new SolidBox(0.3, 0.6, 0.1, 0.5, -M_PI * 0.05, 0.0, 0.0, 0.0));
#endif
    std::vector<const SolidBody *> cBodies;
    for (unsigned i = 0; i < bodies.size(); i++) cBodies.push_back(bodies[i]);
    FluidSolver *solver = new FluidSolver(size, size, densityAir, densitySoot, diffusion, cBodies);
#elif IFL_CHAPTER >= 4
    std::vector<SolidBody *> bodies;
    bodies.push_back(new SolidBox(0.5, 0.6, 0.7, 0.1, M_PI * 0.25, 0.0, 0.0, 0.0)); // v5:986
    std::vector<const SolidBody *> cBodies;
    for (unsigned i = 0; i < bodies.size(); i++) cBodies.push_back(bodies[i]);
    FluidSolver *solver = new FluidSolver(size, size, density, cBodies);
#else
    FluidSolver *solver = new FluidSolver(size, size, density);
#endif
    for (int f = 0; f < frames; f++) {
        for (int i = 0; i < (IFL_CHAPTER == 6 ? 10 : 4); i++) { // v6:1084 runs 10 updates per frame
#if IFL_CHAPTER == 8
            // the inflow lives inside update() in this chapter (v8:1368)
#elif IFL_CHAPTER == 7
            solver->addInflow(0.45, 0.2, 0.1, 0.05, 1.0, solver->ambientT(), 0.0, 0.0); // v7:1112
#elif IFL_CHAPTER == 6
            solver->addInflow(0.35, 0.9, 0.1, 0.05, 1.0, solver->ambientT() + 300.0, 0.0, 0.0); // v6:1086
#elif IFL_CHAPTER == 1
            solver->addInflow(0.45, 0.2, 0.1, 0.01, 1.0, 0.0, 3.0); // v1:362
#else
            solver->addInflow(0.45, 0.2, 0.15, 0.03, 1.0, 0.0, 3.0); // v3:485
#endif
            solver->update(timestep);
            fflush(stdout);
        }
#if IFL_CHAPTER >= 6
        solver->toImage(image, renderHeat);
#else
        solver->toImage(image);
#endif
        uint64_t hsh = 0xcbf29ce484222325ULL;
        for (size_t i = 0; i < image_bytes; i++) hsh = (hsh ^ image[i]) * 0x100000001b3ULL;
        printf("Frame%05d fnv64(rgba)=%016llx\n", f, (unsigned long long)hsh);
#if IFL_CHAPTER >= 4
        for (unsigned i = 0; i < bodies.size(); i++) bodies[i]->update(timestep);
#endif
    }
    delete solver;
    delete[] image;
    return 0;
}
