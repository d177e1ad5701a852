// Test harness (tests/test_tfmath_host_cpu.py): world2cell_fast (img_env_b200/csrc/tfmath.cuh, the division-free cell index
// the kernels use) against the reference's expression int(round(x / res)) (grid_map.cpp:40-44) on the HOST, over random
// coordinates and over coordinates placed within a few ulp of every rounding boundary.  Prints the number of mismatches.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define __forceinline__ inline
#include "tfmath.cuh"

int main(int argc, char** argv) {
    const double res_list[] = {(double)0.015f, (double)0.02f, (double)0.025f, (double)0.05f, (double)0.1f};
    unsigned long long bad = 0, n = 0;
    uint64_t s = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    for (double res : res_list) {
        const double inv = 1.0 / res;
        for (int k = 0; k < 2000000; k++) {          // anywhere on a map of up to 8192 cells (and a bit outside)
            const double x = (rnd() * 9000.0 - 400.0) * res;
            n++; bad += world2cell_fast(x, res, inv) != world2cell(x, res);
        }
        for (int c = -300; c < 8500; c++)            // around every half-integer cell coordinate: x = (c + 0.5) * res +- a few ulp
            for (int u = -6; u <= 6; u++) {
                double x = (c + 0.5) * res;
                for (int q = 0; q < (u < 0 ? -u : u); q++) x = nextafter(x, u < 0 ? -1e300 : 1e300);
                n++; bad += world2cell_fast(x, res, inv) != world2cell(x, res);
            }
    }
    printf("%llu %llu\n", bad, n);
    return bad != 0;
}
