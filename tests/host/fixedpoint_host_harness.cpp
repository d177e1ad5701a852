// Test harness (tests/test_fixedpoint_host_cpu.py): the 2^-32-cell fixed-point form of the view pixel -> world cell map
// (foot.cuh view_const_compute + the guard-band test of view.cuh, restated here in ten lines) against the reference's exact
// operation sequence (map2world, tf multiply, world2map: agent.cpp:388-393, grid_map.cpp:40-55) evaluated with the SAME
// tfmath.cuh functions the kernels use, for every pixel of the view raster over random robot poses.  Claim being tested
// (DESIGN.md section 3, H2): outside a guard band of 2^-19 cell around the rounding boundary the fixed-point cell index
// equals the exact one, so the kernels only fall back to the exact sequence inside the band.  Prints mismatches, pixels,
// and how many pixels fell inside the band.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define __forceinline__ inline
#include "tfmath.cuh"
#define FX_ONE 4294967296.0
#define FX_GUARD 8192u

int main(int argc, char** argv) {
    uint64_t s = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const int n_pose = argc > 2 ? atoi(argv[2]) : 20;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    unsigned long long bad = 0, n = 0, in_band = 0;
    const double res_list[] = {(double)0.015f, (double)0.02f, (double)0.025f};
    for (int p = 0; p < n_pose; p++) {
        const double res = res_list[p % 3];
        const double size = (p & 4) ? 4.0 : 6.0;
        const int vh = (int)(size / res), vw = vh;
        const Tf2 view_base = tf_from_pose(size / 2, size / 2, 3.14159);                  // agent.cpp:84-87
        const double span = (p & 8) ? 100.0 : 10.0;                                        // C4's 110 m map as well
        const Tf2 A = tf_mul(tf_from_pose(rnd() * span + 0.5, rnd() * span + 0.5, (rnd() * 2 - 1) * 3.14159), view_base);
        const long long ax = llrint(A.m00 * FX_ONE), bx = llrint(A.m01 * FX_ONE), cx = llrint((A.ox / res) * FX_ONE) + (1ll << 31);
        const long long ay = llrint(A.m10 * FX_ONE), by = llrint(A.m11 * FX_ONE), cy = llrint((A.oy / res) * FX_ONE) + (1ll << 31);
        for (int i = 0; i < vh; i++)
            for (int j = 0; j < vw; j++) {
                const long long tx = cx + (long long)i * ax + (long long)j * bx, ty = cy + (long long)i * ay + (long long)j * by;
                double wx, wy;
                tf_apply(A, i * res, j * res, wx, wy);
                const int ex = world2cell(wx, res), ey = world2cell(wy, res);
                const bool band = (unsigned)tx + FX_GUARD < 2 * FX_GUARD || (unsigned)ty + FX_GUARD < 2 * FX_GUARD;
                n++; in_band += band;
                if (!band) bad += (int)(tx >> 32) != ex || (int)(ty >> 32) != ey;
            }
    }
    printf("%llu %llu %llu\n", bad, n, in_band);
    return bad != 0;
}
