// flip_internal.cuh -- shared by flip_kernels.cu (particle <-> grid transfers, particle advection) and
// flip_book.cu (particle bookkeeping: init / count / prune / seed, and the stack-ordered extrapolation
// of chapter 8).
#pragma once
#include "ifl_internal.cuh"
#include "solid_geometry.cuh"

namespace ifl {

enum { CELL_EMPTY = 2 }; // v8:115-119

// ParticleQuantities (v8:692-940) as plain device arrays.  Counts, capacity and bin offsets are 64-bit:
// `_maxParticles = w*h*12` (v8:869) overflows the reference's `int` at 13378^2 cells.
// Particle INDICES inside the per-cell bins stay 32-bit unsigned, so one rank holds at most 2^32 - 1
// particles (12 per cell on 18918^2 cells, far beyond what fits one GPU's HBM); flip_init checks it.
struct ParticleSet {
    double *posX, *posY, *prop[4];
    long long count, capacity;
    int avg_per_cell;          // _AvgPerCell v8:698 (4 as shipped; BASELINE config 5 asks for 8)
    unsigned long long draws;  // frand() calls so far (v8:38-46): the LCG state is seed0 advanced `draws` times
    int *counts;               // particles per base cell (bins of the P2G gather == countParticles' _counts)
    long long *offsets;        // exclusive prefix sums of counts
    int *fill;
    unsigned *list;            // particle indices by cell, ascending inside a cell
    long long *scan_tmp;       // block sums of the scans
    size_t scan_tmp_elems;
    int *flags;                // per-attempt / per-particle / per-cell flags of the bookkeeping passes
    long long *flag_offsets;   // their exclusive prefix sums
    size_t flags_elems;
    long long *dev_scalars;    // [8] device scratch for counts returned to the host
    long long *host_scalars;   // pinned mirror
    bool binned;               // bins match the current positions
    Arr weight;                // (w+1) x (h+1) scratch like ParticleQuantities::_weight v8:874
};

__device__ __forceinline__ double lerp1p(double a, double b, double x) { return a * (1.0 - x) + b * x; } // v8:289

__device__ __forceinline__ double field_lerp(const Field &f, double x, double y) { // v8:390-402
    x = std_min(std_max(x - f.ox, 0.0), f.w - 1.001);
    y = std_min(std_max(y - f.oy, 0.0), f.h - 1.001);
    const int ix = (int)x, iy = (int)y;
    x -= ix;
    y -= iy;
    const double *p = f.src.p + ix + (size_t)iy * f.src.pitch;
    const double x00 = p[0], x10 = p[1], x01 = p[f.src.pitch], x11 = p[f.src.pitch + 1];
    return lerp1p(lerp1p(x00, x10, x), lerp1p(x01, x11, x), y);
}

// Exclusive prefix sums of `n` ints into 64-bit offsets (three plain kernels; deterministic).  total_dev
// (may be null) receives the grand total.
int scan_exclusive(ifl_ctx *c, const int *in, long long *out, size_t n, long long *total_dev);
int ensure_bins(ifl_ctx *c);

} // namespace ifl
