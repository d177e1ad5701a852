// Host-side, init/reset-time precomputation (plain C++, compiled with -ffp-contract=off):
//   * footprint lattices                      Agent::init_shape_circle/rectangle   agent.cpp:18-62
//   * view geometry, static FOV spans         Agent::init_view_map / view          agent.cpp:79-90, 373-386
//   * laser origin + ray end cells            agent.cpp:366-369, 414-430
//   * per-pixel highest/lowest touching ray   (closed form of agent.cpp:511-624, see view.cuh)
//   * OpenCV INTER_CUBIC tables               imgproc/resize.cpp (cv::resize) as called by yaml_env.py:433
//   * tf yaw extraction from a quaternion     img_env.cpp:180-184 (Matrix3x3(q).getRPY)
//   * RVO2 obstacle ring + BSP                RVOSimulator::addObstacle RVOSimulator.cpp:130-168,
//                                             KdTree::buildObstacleTreeRecursive KdTree.cpp:130-257
// The pose-independent parts of Agent::view are evaluated here once, with glibc's libm exactly as
// the node would, instead of per pixel per step on the device.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>
#include "state.cuh"

namespace ht {
// bounding circle of a footprint lattice for the stamping kernel's cell bitmap: centre of the bounding box (m) and the
// radius in cells of the farthest point from it, + 2 cells (rounding of the centre and of the points)
inline void bounding_circle(const std::vector<double>& pts, double res, double out[3]) {
    out[0] = out[1] = 0; out[2] = 2;
    if (pts.empty()) return;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (size_t k = 0; k + 1 < pts.size(); k += 2) { x0 = std::min(x0, pts[k]); x1 = std::max(x1, pts[k]); y0 = std::min(y0, pts[k + 1]); y1 = std::max(y1, pts[k + 1]); }
    out[0] = 0.5 * (x0 + x1); out[1] = 0.5 * (y0 + y1);
    double r = 0;
    for (size_t k = 0; k + 1 < pts.size(); k += 2) r = std::max(r, hypot(pts[k] - out[0], pts[k + 1] - out[1]));
    out[2] = ceil(r / res) + 2;
}


inline double f32(double x) { return (double)(float)x; }

inline void lattice_circle(double s0, double s1, double s2, std::vector<double>& out) {
    double resolution = 0.01;
    int bb = (int)ceil(s2 / resolution);
    for (int m = -bb; m <= bb; m++)
        for (int n = -bb; n <= bb; n++)
            if (sqrt(m * resolution * m * resolution + n * resolution * n * resolution) <= s2) {
                out.push_back(m * resolution + s0);
                out.push_back(n * resolution + s1);
            }
}
// Ring of a circle lattice.  A grid cell (side res) whose centre lies within r_in = r - cover - eps of the disc centre holds a
// lattice point for sure -- every point of the plane is within cover = 0.01 * sqrt(2) / 2 of a point of the 0.01 m lattice, which
// then lies strictly inside the cell's square (res / 2 > cover) and inside the disc -- so only the lattice points that can fall
// into OTHER cells have to be evaluated one by one: those farther than r_in - res * sqrt(2) / 2 from the centre.
// Returns r_in, or a negative number when the resolution does not allow the shortcut; ring = (x, y) pairs like the lattice.
inline double lattice_circle_ring(double s0, double s1, double s2, double res, std::vector<double>& ring) {
    const double cover = 0.01 * 0.70710678118654757;
    if (res * 0.5 - cover < 2e-4) return -1.0;
    const double r_in = s2 - cover - 5e-5;
    if (r_in <= 0) return -1.0;
    const double thr = r_in - res * 0.70710678118654757 - 5e-5;
    double resolution = 0.01;
    int bb = (int)ceil(s2 / resolution);
    for (int m = -bb; m <= bb; m++)
        for (int n = -bb; n <= bb; n++) {
            const double rad = sqrt(m * resolution * m * resolution + n * resolution * n * resolution);
            if (rad <= s2 && rad > thr) { ring.push_back(m * resolution + s0); ring.push_back(n * resolution + s1); }
        }
    return r_in;
}
inline void lattice_rect(double s0, double s1, double s2, double s3, std::vector<double>& out) {
    double resolution = 0.01;
    int x_min = (int)floor(s0 / resolution), x_max = (int)ceil(s1 / resolution);
    int y_min = (int)floor(s2 / resolution), y_max = (int)ceil(s3 / resolution);
    for (int m = x_min; m <= x_max; m++)
        for (int n = y_min; n <= y_max; n++) { out.push_back(m * resolution); out.push_back(n * resolution); }
}

// tf::Matrix3x3(q).getRPY -> yaw for a general quaternion (img_env.cpp:180-184, 230-236, 264-268)
inline double yaw_from_quaternion(double x, double y, double z, double w) {
    double d = x * x + y * y + z * z + w * w;
    double s = 2.0 / d;
    double xs = x * s, ys = y * s, zs = z * s;
    double wx = w * xs, wy = w * ys, wz = w * zs;
    double xx = x * xs, xy = x * ys, xz = x * zs;
    double yy = y * ys, yz = y * zs, zz = z * zs;
    (void)wx; (void)xx; (void)yz;
    double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy;
    if (fabs(m20) >= 1) return 0.0;
    double pitch = -asin(m20);
    return atan2(m10 / cos(pitch), m00 / cos(pitch));
}

struct TypeTables {
    RobotType t;
    std::vector<double> lattice;
    std::vector<short> ray_end;
    std::vector<short> spans;
    std::vector<unsigned short> khi, klo;
    std::vector<uint32_t> own_mask, tile_fov, edge_px, edge_tiles, dtab, ostat;
    std::vector<double> ring;      // circle robots: lattice points near the rim (lattice_circle_ring)
};

// desc: shape, size[4], sensor_cfg[2] (already float32-widened)
inline std::string build_type(const Cfg& c, const double* desc, double view_angle_begin, double view_angle_end,
                              double view_min_dist, double view_max_dist, TypeTables& T) {
    int shape = (int)desc[0];
    if (shape == 0) lattice_circle(desc[1], desc[2], desc[3], T.lattice);
    T.t.ring_n = -1; T.t.disc_cx = T.t.disc_cy = T.t.disc_rin = 0;
    if (shape == 0) {
        const double r_in = lattice_circle_ring(desc[1], desc[2], desc[3], c.res, T.ring);
        if (r_in > 0) { T.t.ring_n = (int)T.ring.size() / 2; T.t.disc_cx = desc[1]; T.t.disc_cy = desc[2]; T.t.disc_rin = r_in; }
        else T.ring.clear();
    }
    else if (shape == 1) lattice_rect(desc[1], desc[2], desc[3], desc[4], T.lattice);
    T.t.n_pts = (int)T.lattice.size() / 2;
    double sx = desc[5], sy = desc[6];
    T.t.sensor_x = sx; T.t.sensor_y = sy;
    // laser origin: base2view(sensor_base_) -> view_map_.world2map
    double ovx, ovy;
    tf_apply(c.base_view, sx, sy, ovx, ovy);
    T.t.org_x = world2cell(ovx, c.res); T.t.org_y = world2cell(ovy, c.res);
    if (T.t.org_x < 0 || T.t.org_x >= c.vh || T.t.org_y < 0 || T.t.org_y >= c.vw) return "sensor origin outside the view raster";
    // ray end cells
    double map_width = c.base_view.ox, map_height = c.base_view.oy;
    double max_range = sqrt(map_width * map_width + map_height * map_height);
    double angle_step = fabs(view_angle_end - view_angle_begin) / c.range_total;
    T.ray_end.resize(2 * (size_t)c.range_total);
    for (int i = 0; i < c.range_total; i++) {
        double cur_angle = view_angle_begin + angle_step * i;
        double x = max_range * cos(cur_angle), y = max_range * sin(cur_angle);
        double vx, vy;
        tf_apply(c.base_view, x, y, vx, vy);
        int ex = world2cell(vx, c.res), ey = world2cell(vy, c.res);
        if (ex < -32000 || ex > 32000 || ey < -32000 || ey > 32000) return "ray end cell out of range";
        T.ray_end[2 * i] = (short)ex; T.ray_end[2 * i + 1] = (short)ey;
    }
    // FOV spans (agent.cpp:379-386)
    T.spans.assign((size_t)c.vh * MAX_SPANS * 2, -1);
    for (int i = 0; i < c.vh; i++) {
        int nsp = 0; bool in = false;
        for (int j = 0; j <= c.vw; j++) {
            bool ok = false;
            if (j < c.vw) {
                double xv = i * c.res, yv = j * c.res, xb, yb;
                tf_apply(c.view_base, xv, yv, xb, yb);
                double view_angle = atan2(yb - sy, xb - sx);
                ok = !(view_angle <= view_angle_begin || view_angle >= view_angle_end || xb < view_min_dist || xb > view_max_dist);
            }
            if (ok && !in) {
                if (nsp >= MAX_SPANS) return "field of view needs more than 2 column spans per view row";
                T.spans[(i * MAX_SPANS + nsp) * 2] = (short)j; in = true;
            } else if (!ok && in) { T.spans[(i * MAX_SPANS + nsp) * 2 + 1] = (short)j; nsp++; in = false; }
        }
    }
    {   // 32x32 view tiles that contain at least one FOV pixel
        int th = (c.vh + 31) / 32, tw = c.vwb;
        T.tile_fov.assign(((size_t)th * tw + 31) / 32, 0u);
        for (int i = 0; i < c.vh; i++)
            for (int sp = 0; sp < MAX_SPANS; sp++) {
                int c0 = T.spans[(i * MAX_SPANS + sp) * 2], c1 = T.spans[(i * MAX_SPANS + sp) * 2 + 1];
                if (c0 < 0) continue;
                for (int wj = c0 >> 5; wj <= (c1 - 1) >> 5; wj++) { int t = (i >> 5) * tw + wj; T.tile_fov[t >> 5] |= 1u << (t & 31); }
            }
    }
    {   // FOV bounding box and FOV-edge pixels (in FOV with an 8-neighbour inside the raster that is not in FOV)
        auto in_fov = [&](int i, int j) {
            if (i < 0 || i >= c.vh || j < 0 || j >= c.vw) return false;
            for (int sp = 0; sp < MAX_SPANS; sp++) {
                int c0 = T.spans[(i * MAX_SPANS + sp) * 2], c1 = T.spans[(i * MAX_SPANS + sp) * 2 + 1];
                if (c0 >= 0 && j >= c0 && j < c1) return true;
            }
            return false;
        };
        T.t.fov_r0 = c.vh; T.t.fov_r1 = -1; T.t.fov_c0 = c.vw; T.t.fov_c1 = -1;
        for (int i = 0; i < c.vh; i++)
            for (int j = 0; j < c.vw; j++) {
                if (!in_fov(i, j)) continue;
                T.t.fov_r0 = std::min(T.t.fov_r0, i); T.t.fov_r1 = std::max(T.t.fov_r1, i);
                T.t.fov_c0 = std::min(T.t.fov_c0, j); T.t.fov_c1 = std::max(T.t.fov_c1, j);
                bool edge = false;
                for (int di = -1; di <= 1 && !edge; di++)
                    for (int dj = -1; dj <= 1; dj++) {
                        int ii = i + di, jj = j + dj;
                        if (ii < 0 || ii >= c.vh || jj < 0 || jj >= c.vw) continue;
                        if (!in_fov(ii, jj)) { edge = true; break; }
                    }
                if (edge || (i == T.t.org_x && j == T.t.org_y)) T.edge_px.push_back(((uint32_t)i << 16) | (uint32_t)j);   // + the laser origin (no predecessor)
            }
        // 16x16-pixel view tiles that hold such a pixel: a footprint record whose view-space bounding box (+3 px) touches one
        // of them contributes all of its cells to the raster, not only its candidate cells (view.cuh, phase B)
        const int etw = (c.vw + 15) / 16, eth = (c.vh + 15) / 16;
        // (one 64-bit column mask per tile row, as two u32 words: the view raster is at most 1022 pixels = 64 tiles wide)
        T.edge_tiles.assign((size_t)eth * 2, 0u);
        (void)etw;
        for (uint32_t ep : T.edge_px) { const int ti = (int)(ep >> 16) / 16, tj = (int)(ep & 0xFFFF) / 16; T.edge_tiles[2 * ti + (tj >> 5)] |= 1u << (tj & 31); }
    }
    // own footprint cells in the view raster: draw(view_map_, 100, "view_map", bbox_) agent.cpp:503
    size_t npx = (size_t)c.vh * c.vw;
    T.own_mask.assign((npx + 31) / 32, 0u);
    int r0 = c.vh, r1 = -1, c0 = c.vw, c1 = -1;
    for (int k = 0; k < T.t.n_pts; k++) {
        double vx, vy;
        tf_apply(c.base_view, T.lattice[2 * k], T.lattice[2 * k + 1], vx, vy);
        int cx = world2cell(vx, c.res), cy = world2cell(vy, c.res);
        if (cx >= 0 && cx < c.vh && cy >= 0 && cy < c.vw) {
            size_t q = (size_t)cx * c.vw + cy;
            T.own_mask[q >> 5] |= 1u << (q & 31);
        }
        r0 = std::min(r0, cx); r1 = std::max(r1, cx); c0 = std::min(c0, cy); c1 = std::max(c1, cy);
    }
    // conservative box where a set occupancy bit may be the robot's own stamp (superset is always safe)
    T.t.zone_r0 = r0 - 4; T.t.zone_r1 = r1 + 4; T.t.zone_c0 = c0 - 4; T.t.zone_c1 = c1 + 4;
    {   // the same thing in world cells: radius of the footprint around the robot position + margin
        double rmax = 0;
        for (int k = 0; k < T.t.n_pts; k++) rmax = std::max(rmax, sqrt(T.lattice[2 * k] * T.lattice[2 * k] + T.lattice[2 * k + 1] * T.lattice[2 * k + 1]));
        T.t.zone_rad = (int)ceil(rmax / c.res) + 5;
        double bc[3]; bounding_circle(T.lattice, c.res, bc);
        T.t.stamp_cx = bc[0]; T.t.stamp_cy = bc[1]; T.t.stamp_rad = (int)bc[2];
    }
    // per-pixel highest / lowest touching ray: walk every ray over its full static cell sequence
    T.khi.assign(npx, 0xFFFF); T.klo.assign(npx, 0xFFFF);
    std::vector<unsigned short> kcnt(npx, 0);
    for (int k = 0; k < c.range_total; k++) {
        int x1 = T.t.org_x, y1 = T.t.org_y, x2 = T.ray_end[2 * k], y2 = T.ray_end[2 * k + 1];
        int w = x2 - x1, h = y2 - y1;
        int dx = ((w > 0) << 1) - 1, dy = ((h > 0) << 1) - 1;
        w = abs(w); h = abs(h);
        int f, x, y;
        auto visit = [&](int cx, int cy) {
            size_t q = (size_t)cx * c.vw + cy;
            T.khi[q] = (unsigned short)k;                     // k ascending -> last write is the max
            if (T.klo[q] == 0xFFFF) T.klo[q] = (unsigned short)k;
            kcnt[q]++;
        };
        if (w > h) {
            f = 2 * h - w;
            for (x = x1, y = y1; x != x2; x += dx) {
                if (x < 0 || x >= c.vh || y < 0 || y >= c.vw) break;
                visit(x, y);
                if (f < 0) f += 2 * h; else { y += dy; f += (h - w) * 2; }
            }
        } else {
            f = 2 * w - h;
            for (x = x1, y = y1; y != y2; y += dy) {
                if (x < 0 || x >= c.vh || y < 0 || y >= c.vw) break;
                visit(x, y);
                if (f < 0) f += 2 * w; else { x += dx; f += (w - h) * 2; }
            }
        }
    }
    // The rays through a cell form a contiguous index interval in practice (monotone fan of digital lines); where that holds
    // (count == khi - klo + 1) bit 15 of klo is set and the kernel skips the per-ray touch test: on every ray through the cell
    // the step index is the Chebyshev distance to the origin (x-major rays visit one cell per x, y-major one per y).
    for (size_t q = 0; q < npx; q++)
        if (T.khi[q] != 0xFFFF && kcnt[q] == T.khi[q] - T.klo[q] + 1) T.klo[q] |= 0x8000;
    return "";
}

// Static-map tables for the observation kernel: cand = occupied cells (value < 250) that have a free cell within their 5x5
// neighbourhood or lie within 2 cells of the map border -- the only static cells a laser ray can hit first (view.cuh,
// phase B); crow / orow = per 32x32-cell block the mask of rows that hold a cand / occupied bit.
inline void static_planes(const std::vector<uint32_t>& occ, int H, int Wb, std::vector<uint32_t>& cand, std::vector<uint32_t>& crow,
                          std::vector<uint32_t>& orow) {
    const int Hc = (H + 31) / 32;
    cand.assign((size_t)H * Wb, 0u); crow.assign((size_t)Hc * Wb, 0u); orow.assign((size_t)Hc * Wb, 0u);
    for (int X = 0; X < H; X++)
        for (int bj = 0; bj < Wb; bj++) {
            const uint32_t w = occ[(size_t)X * Wb + bj];
            if (!w) continue;
            uint32_t interior = 0xffffffffu;
            for (int dr = -2; dr <= 2 && interior; dr++) {
                const int Xr = X + dr;
                if (Xr < 0 || Xr >= H) { interior = 0; break; }
                const uint32_t* rp = occ.data() + (size_t)Xr * Wb + bj;
                const uint32_t wc = rp[0], wl = bj > 0 ? rp[-1] : 0u, wr = bj + 1 < Wb ? rp[1] : 0u;
                interior &= wc & ((wc << 1) | (wl >> 31)) & ((wc << 2) | (wl >> 30)) & ((wc >> 1) | (wr << 31)) & ((wc >> 2) | (wr << 30));
            }
            const uint32_t cw = w & ~interior;
            cand[(size_t)X * Wb + bj] = cw;
            orow[(size_t)(X >> 5) * Wb + bj] |= 1u << (X & 31);
            if (cw) crow[(size_t)(X >> 5) * Wb + bj] |= 1u << (X & 31);
        }
}

// cv::resize INTER_CUBIC coefficient tables for a square src -> dst (OpenCV imgproc/resize.cpp)
inline void cubic_tables(int src, int dst, std::vector<short>& need_idx, std::vector<short>& tap, std::vector<short>& coef) {
    double inv_scale = (double)dst / src;
    double scale = 1. / inv_scale;
    std::vector<int> taps_src(4 * (size_t)dst);
    coef.resize(4 * (size_t)dst);
    std::vector<char> used(src, 0);
    for (int dx = 0; dx < dst; dx++) {
        float fx = (float)((dx + 0.5) * scale - 0.5);
        int sx = (int)floor(fx);
        fx -= sx;
        const float A = -0.75f;
        float cbuf[4];
        cbuf[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
        cbuf[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
        cbuf[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
        cbuf[3] = 1.f - cbuf[0] - cbuf[1] - cbuf[2];
        for (int k = 0; k < 4; k++) {
            int sxk = std::min(std::max(sx - 1 + k, 0), src - 1);
            taps_src[4 * dx + k] = sxk;
            long iv = lrintf(cbuf[k] * 2048.f);      // saturate_cast<short>(float): cvRound (half to even)
            coef[4 * dx + k] = (short)std::min(32767l, std::max(-32768l, iv));
            if (coef[4 * dx + k] != 0) used[sxk] = 1;
        }
    }
    std::vector<int> pos(src, -1);
    need_idx.clear();
    for (int i = 0; i < src; i++) if (used[i]) { pos[i] = (int)need_idx.size(); need_idx.push_back((short)i); }
    tap.resize(4 * (size_t)dst);
    for (int dx = 0; dx < dst; dx++)
        for (int k = 0; k < 4; k++) {
            int sxk = taps_src[4 * dx + k];
            // taps with a zero weight may point at an unused source line: alias them to any needed line
            tap[4 * dx + k] = (short)(pos[sxk] >= 0 ? pos[sxk] : 0);
        }
}

// float -> IEEE binary16, round to nearest even
inline uint16_t f32_to_f16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        m |= 0x800000u;
        int shift = 14 - e;
        uint32_t hm = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint32_t hm = m >> 13, rem = m & 0x1fffu;
    uint16_t h = (uint16_t)(sign | ((uint32_t)e << 10) | hm);
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) h++;
    return h;
}
// numpy: uint8.astype(float16) / 255.0 -> half division = float32 divide, round to half
inline void f16_lut(uint16_t* out) { for (int i = 0; i < 256; i++) out[i] = f32_to_f16((float)i / 255.0f); }

// ---------------- RVO2 obstacles ----------------
struct RvoObst { float px, py, dx, dy; int convex, next, prev; };
struct RvoNode { int obstacle, left, right, parent; };
struct F2 { float x, y; };
inline F2 f2(float x, float y) { F2 r; r.x = x; r.y = y; return r; }
inline F2 sub(F2 a, F2 b) { return f2(a.x - b.x, a.y - b.y); }
inline float fdet(F2 a, F2 b) { return a.x * b.y - a.y * b.x; }
inline float fleftOf(F2 a, F2 b, F2 c) { return fdet(sub(a, c), sub(b, a)); }
inline F2 fnormalize(F2 a) { float l = sqrtf(a.x * a.x + a.y * a.y); float inv = 1.0f / l; return f2(a.x * inv, a.y * inv); }

inline void rvo_add_obstacle(std::vector<RvoObst>& obs, const F2* v, int n) {
    const int obstacleNo = (int)obs.size();
    for (int i = 0; i < n; ++i) {
        RvoObst o; o.px = v[i].x; o.py = v[i].y; o.next = o.prev = -1;
        int me = (int)obs.size();
        if (i != 0) { o.prev = me - 1; obs[me - 1].next = me; }
        if (i == n - 1) { o.next = obstacleNo; }
        F2 d = fnormalize(sub(v[(i == n - 1 ? 0 : i + 1)], v[i]));
        o.dx = d.x; o.dy = d.y;
        if (n == 2) o.convex = 1;
        else o.convex = fleftOf(v[(i == 0 ? n - 1 : i - 1)], v[i], v[(i == n - 1 ? 0 : i + 1)]) >= 0.0f;
        obs.push_back(o);
        if (i == n - 1) obs[obstacleNo].prev = me;
    }
}
inline int rvo_build_tree(std::vector<RvoObst>& obs, std::vector<RvoNode>& nodes, const std::vector<int>& list) {
    const float EPS = 0.00001f;
    if (list.empty()) return -1;
    size_t optimalSplit = 0, minLeft = list.size(), minRight = list.size();
    auto P = [&](int i) { return f2(obs[i].px, obs[i].py); };
    for (size_t i = 0; i < list.size(); ++i) {
        size_t leftSize = 0, rightSize = 0;
        const int I1 = list[i], I2 = obs[I1].next;
        for (size_t j = 0; j < list.size(); ++j) {
            if (i == j) continue;
            const int J1 = list[j], J2 = obs[J1].next;
            const float j1LeftOfI = fleftOf(P(I1), P(I2), P(J1));
            const float j2LeftOfI = fleftOf(P(I1), P(I2), P(J2));
            if (j1LeftOfI >= -EPS && j2LeftOfI >= -EPS) ++leftSize;
            else if (j1LeftOfI <= EPS && j2LeftOfI <= EPS) ++rightSize;
            else { ++leftSize; ++rightSize; }
            if (std::make_pair(std::max(leftSize, rightSize), std::min(leftSize, rightSize)) >=
                std::make_pair(std::max(minLeft, minRight), std::min(minLeft, minRight))) break;
        }
        if (std::make_pair(std::max(leftSize, rightSize), std::min(leftSize, rightSize)) <
            std::make_pair(std::max(minLeft, minRight), std::min(minLeft, minRight))) {
            minLeft = leftSize; minRight = rightSize; optimalSplit = i;
        }
    }
    std::vector<int> leftObstacles(minLeft), rightObstacles(minRight);
    size_t leftCounter = 0, rightCounter = 0;
    const size_t i = optimalSplit;
    const int I1 = list[i], I2 = obs[I1].next;
    for (size_t j = 0; j < list.size(); ++j) {
        if (i == j) continue;
        const int J1 = list[j], J2 = obs[J1].next;
        const float j1LeftOfI = fleftOf(P(I1), P(I2), P(J1));
        const float j2LeftOfI = fleftOf(P(I1), P(I2), P(J2));
        if (j1LeftOfI >= -EPS && j2LeftOfI >= -EPS) leftObstacles[leftCounter++] = J1;
        else if (j1LeftOfI <= EPS && j2LeftOfI <= EPS) rightObstacles[rightCounter++] = J1;
        else {
            const float t = fdet(sub(P(I2), P(I1)), sub(P(J1), P(I1))) / fdet(sub(P(I2), P(I1)), sub(P(J1), P(J2)));
            F2 dj = sub(P(J2), P(J1));
            F2 splitpoint = f2(P(J1).x + t * dj.x, P(J1).y + t * dj.y);
            RvoObst n; n.px = splitpoint.x; n.py = splitpoint.y; n.prev = J1; n.next = J2; n.convex = 1;
            n.dx = obs[J1].dx; n.dy = obs[J1].dy;
            int id = (int)obs.size();
            obs.push_back(n);
            obs[J1].next = id; obs[J2].prev = id;
            if (j1LeftOfI > 0.0f) { leftObstacles[leftCounter++] = J1; rightObstacles[rightCounter++] = id; }
            else { rightObstacles[rightCounter++] = J1; leftObstacles[leftCounter++] = id; }
        }
    }
    int me = (int)nodes.size();
    nodes.push_back(RvoNode());
    nodes[me].obstacle = I1; nodes[me].parent = -1;
    int l = rvo_build_tree(obs, nodes, leftObstacles);
    nodes[me].left = l;
    if (l >= 0) nodes[l].parent = me;
    int r = rvo_build_tree(obs, nodes, rightObstacles);
    nodes[me].right = r;
    if (r >= 0) nodes[r].parent = me;
    return me;
}

}  // namespace ht
