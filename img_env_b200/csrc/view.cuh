// Observation stage: footprint stamping, collision codes, egocentric view raster, laser rays,
// laser_map reconstruction, 400->48 cubic resize, pedestrian observation packing.
// Reference: ImgEnv::view_ped/view_robot (img_env.cpp:594-674), Agent::draw (agent.cpp:285-327),
// PedAgent::draw_leg (agent.cpp:737-774), Agent::view (agent.cpp:356-509), Agent::bresenhamLine
// (agent.cpp:511-624), ImgEnv::get_states (img_env.cpp:547-587) and the Python post-processing
// yaml_env.py:392-481.  See DESIGN.md for how each phase maps to the reference and why the
// results are identical.
#pragma once
#include "state.cuh"
#include "kin.cuh"

#define VIEW_THREADS 256
#define FX_ONE 4294967296.0            // 2^32: fixed-point scale of cell coordinates
#define FX_GUARD 8192u                 // |frac - 0.5| below 2^-19 cells -> exact fp64 fallback

// ---------------------------------------------------------------------------------------------
// per-scene planes
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t plane_cells(const Cfg& c) { return ((size_t)c.H * c.W + 3) & ~(size_t)3; }

__device__ __forceinline__ void flag_or(uint8_t* flags, size_t ci, unsigned bits) {
    unsigned* w = reinterpret_cast<unsigned*>(flags + (ci & ~(size_t)3));
    atomicOr(w, bits << (8 * (ci & 3)));
}
__device__ __forceinline__ void flag_clear(uint8_t* flags, size_t ci, unsigned bits) {
    unsigned* w = reinterpret_cast<unsigned*>(flags + (ci & ~(size_t)3));
    atomicAnd(w, ~(bits << (8 * (ci & 3))));
}
// The value Agent::draw / Agent::view would read from robot `self`'s global_map_ at a cell:
// obs_map_ (static + reset objects) -> peds_map_ (pedestrians, value 1) -> other robots (value 2);
// a writer never overwrites 0/1/2 except the right-leg quirk (agent.cpp:767-770).
__device__ __forceinline__ int global_value(const Dev& d, int s, int self, int cx, int cy) {
    size_t ci = (size_t)cx * d.c.W + cy;
    size_t po = (size_t)s * plane_cells(d.c);
    int sv = d.grid[ci];
    unsigned f = d.flags[po + ci];
    if ((f & F_OBJ) && sv > 2) sv = 0;
    int v;
    if (f & F_RIGHT) v = 1;
    else if (f & F_LEFT) v = (sv == 0) ? 0 : 1;
    else if (f & F_CIRC) v = (sv <= 2) ? sv : 1;
    else v = sv;
    if (v > 2 && (f & F_ROBOT)) {
        if ((f & F_MULTI) || d.rmin[po + ci] != (unsigned short)self) v = 2;
    }
    return v;
}

// Reset objects (mode 4 stamp, 12 unstamp) are written point by point: they only change at reset. They live in
// base_occ as well as occ_all, and the block counts in coarse's low bits count them (no agent is stamped then).
__device__ __forceinline__ void stamp_object_cell(const Dev& d, int s, int cx, int cy, int mode) {
    if ((unsigned)cx >= (unsigned)d.c.H || (unsigned)cy >= (unsigned)d.c.W) return;
    const size_t ci = (size_t)cx * d.c.W + cy;
    const size_t po = (size_t)s * plane_cells(d.c);
    const size_t wi = (size_t)s * d.c.H * d.c.Wb + (size_t)cx * d.c.Wb + (cy >> 5);
    uint32_t* cz = d.coarse + (size_t)s * d.c.Hc * d.c.Wb + (size_t)(cx >> 5) * d.c.Wb + (cy >> 5);
    const uint32_t bit = 1u << (cy & 31);
    if (mode == 4) {
        flag_or(d.flags + po, ci, F_OBJ);
        atomicOr(d.occ_all + wi, bit);
        if (!(atomicOr(d.base_occ + wi, bit) & bit)) atomicAdd(cz, 1u);     // first setter of the bit maintains the block count
    } else {
        flag_clear(d.flags + po, ci, F_OBJ);
        if (d.static_occ[(size_t)cx * d.c.Wb + (cy >> 5)] & bit) return;
        atomicAnd(d.occ_all + wi, ~bit);
        if (atomicAnd(d.base_occ + wi, ~bit) & bit) atomicSub(cz, 1u);
    }
}

// Agent footprints.  The 0.01 m lattice puts 2-3 points into every 0.015 m cell, so a footprint is first reduced to
// its set of cells in a shared-memory bitmap (rows x 32-cell words aligned with occ_all's words).  The planes are
// then updated per group of 8 cells with fire-and-forget reductions only -- nothing waits for an L2 round trip
// except the robots' flag words, whose previous value tells whether another robot already stamped the cell:
//   stamp:   flags |= bits, rmin = id (plain store: with two robots on a cell F_MULTI makes rmin irrelevant),
//            occ_all |= mask, coarse |= 1<<31
//   unstamp: flags &= ~bits, occ_all = base_occ (every unstamping agent writes the same final word),
//            coarse &= ~(1<<31)
// mode: 0 robot(id), 1 ped body (circle), 2 left leg, 3 right leg; 8+ = unstamp of (mode-8)
#define STAMP_THREADS 64
struct StampBox { int cx0, wj0, nrow, wpr; };
inline __host__ __device__ int stamp_rad_cells(double ext, double res) { return (int)ceil(ext / res) + 1; }
inline __host__ __device__ int stamp_bitmap_words(int rad_cells) { return (2 * rad_cells + 1) * ((2 * rad_cells) / 32 + 2); }
__device__ __forceinline__ StampBox stamp_box(const Dev& d, double x, double y, int rad_cells) {
    const int ccx = world2cell(x, d.c.res), ccy = world2cell(y, d.c.res);
    StampBox b;
    b.cx0 = ccx - rad_cells; b.nrow = 2 * rad_cells + 1;
    b.wj0 = (ccy - rad_cells) >> 5; b.wpr = ((ccy + rad_cells) >> 5) - b.wj0 + 1;
    return b;
}
__device__ __forceinline__ void stamp_cells8(const Dev& d, int s, int cx, int cy0, unsigned m8, int mode, int id) {
    // cells (cx, cy0 .. cy0+7) selected by m8; cy0 is a multiple of 8: one occ_all word, one coarse block, <= 3 flag words
    if ((unsigned)cx >= (unsigned)d.c.H || cy0 < 0) return;
    if (cy0 + 8 > d.c.W) m8 &= cy0 >= d.c.W ? 0u : (0xFFu >> (cy0 + 8 - d.c.W));
    if (!m8) return;
    const size_t po = (size_t)s * plane_cells(d.c);
    const size_t ci0 = (size_t)cx * d.c.W + cy0;
    const size_t wi = (size_t)s * d.c.H * d.c.Wb + (size_t)cx * d.c.Wb + (cy0 >> 5);
    uint32_t* cz = d.coarse + (size_t)s * d.c.Hc * d.c.Wb + (size_t)(cx >> 5) * d.c.Wb + (cy0 >> 5);
    const int m = mode & 7;
    const unsigned fb = m == 0 ? F_ROBOT : m == 1 ? F_CIRC : m == 2 ? F_LEFT : F_RIGHT;
    // spread the 8 cell bits over the (at most 3) aligned 32-bit words of the flags plane they fall into
    const int a0 = (int)(ci0 & 3);
    const unsigned long long sel = (unsigned long long)m8 << a0;              // bit k = byte k of the 12-byte window at ci0 - a0
    unsigned wmask[3];
#pragma unroll
    for (int w = 0; w < 3; w++) {
        const unsigned n4 = (unsigned)(sel >> (4 * w)) & 0xFu;
        wmask[w] = ((n4 & 1u) | ((n4 & 2u) << 7) | ((n4 & 4u) << 14) | ((n4 & 8u) << 21)) * fb;
    }
    unsigned* fw = reinterpret_cast<unsigned*>(d.flags + po + (ci0 - a0));
    if (mode < 8) {
        if (m == 0) {
            unsigned old[3];
#pragma unroll
            for (int w = 0; w < 3; w++) old[w] = wmask[w] ? atomicOr(fw + w, wmask[w]) : 0u;
#pragma unroll
            for (int b = 0; b < 8; b++) if ((m8 >> b) & 1u) d.rmin[po + ci0 + b] = (unsigned short)id;
#pragma unroll
            for (int w = 0; w < 3; w++) {
                const unsigned again = old[w] & wmask[w];                        // another robot set F_ROBOT here before us
                if (again) atomicOr(fw + w, again * (F_MULTI / F_ROBOT));
            }
        } else {
#pragma unroll
            for (int w = 0; w < 3; w++) if (wmask[w]) atomicOr(fw + w, wmask[w]);
        }
        atomicOr(d.occ_all + wi, m8 << (cy0 & 31));
        atomicOr(cz, 0x80000000u);
    } else {
        const unsigned clr = m == 0 ? (F_MULTI / F_ROBOT + 1u) : 1u;            // robots also drop F_MULTI
#pragma unroll
        for (int w = 0; w < 3; w++) if (wmask[w]) atomicAnd(fw + w, ~(wmask[w] * clr));
        d.occ_all[wi] = d.base_occ[wi];
        atomicAnd(cz, 0x7FFFFFFFu);
    }
}
__device__ __forceinline__ void stamp_part(const Dev& d, int s, const Tf2& t, const double* pts, int n, int mode, int id,
                                           double offx, double offy, double ccx, double ccy, int rad_cells, uint32_t* bm) {
    double bwx, bwy;
    tf_apply(t, ccx + offx, ccy + offy, bwx, bwy);      // world position of the part's bounding-circle centre
    const StampBox bx = stamp_box(d, bwx, bwy, rad_cells);
    const int nw = bx.nrow * bx.wpr;
    for (int k = threadIdx.x; k < nw; k += STAMP_THREADS) bm[k] = 0u;
    __syncthreads();
    const bool leg = (mode & 7) == 2 || (mode & 7) == 3;
    for (int k = threadIdx.x; k < n; k += STAMP_THREADS) {
        double px = pts[2 * k], py = pts[2 * k + 1];
        if (leg) { px = px + offx; py = py + offy; }   // leg2base: identity rotation + leg origin (agent.cpp:815-837)
        double wx, wy;
        tf_apply(t, px, py, wx, wy);
        const int cx = world2cell_fast(wx, d.c.res, d.c.inv_res), cy = world2cell_fast(wy, d.c.res, d.c.inv_res);
        const int r = cx - bx.cx0, w = (cy >> 5) - bx.wj0;
        if ((unsigned)r < (unsigned)bx.nrow && (unsigned)w < (unsigned)bx.wpr) atomicOr(&bm[r * bx.wpr + w], 1u << (cy & 31));
        else stamp_cells8(d, s, cx, cy & ~7, 1u << (cy & 7), mode, id);   // cannot happen (the box bounds the footprint); still a valid write
    }
    __syncthreads();
    const unsigned inv_wpr = (65536u + bx.wpr - 1) / bx.wpr;     // word / wpr == (word * inv_wpr) >> 16 for word < 8192, wpr <= 12 (checked exhaustively)
    for (int it = threadIdx.x; it < 4 * nw; it += STAMP_THREADS) {
        const int word = it >> 2, byte = it & 3;
        const unsigned m8 = (bm[word] >> (8 * byte)) & 0xFFu;
        if (!m8) continue;
        const int r = bx.wpr <= 12 && nw < 8192 ? (int)(((unsigned)word * inv_wpr) >> 16) : word / bx.wpr, w = word - r * bx.wpr;
        stamp_cells8(d, s, bx.cx0 + r, (bx.wj0 + w) * 32 + 8 * byte, m8, mode, id);
    }
    __syncthreads();
}

// grid = n_scenes * (R + P) CTAs; `unstamp` selects the inverse operation.
__global__ void __launch_bounds__(STAMP_THREADS) k_stamp_agents(Dev d, const int* scene_ids, int unstamp) {
    extern __shared__ uint32_t bm[];     // stamp_bitmap_words(largest agent) words
    const int per = d.c.R + d.c.P;
    const int sl = blockIdx.x / per, a = blockIdx.x % per;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int add = unstamp ? 8 : 0;
    if (unstamp == 2 && a == 0 && threadIdx.x == 0) d.step_no[s] += 1;   // step_++ (img_env.cpp:518), after every reader of this step
    const bool is_robot = a < d.c.R;
    const int p = a - d.c.R;
    // An agent's stamp can only be read by a robot whose collision lattice or view raster reaches it: skip agents
    // farther than (view half-diagonal + own extent) from every (other) robot. The decision only depends on poses,
    // so stamp and unstamp agree.
    double x, y, yaw, ext;
    if (is_robot) { const int idx = s * d.c.R + a; x = RBF(d, RB_X, idx); y = RBF(d, RB_Y, idx); yaw = RBF(d, RB_YAW, idx); ext = d.types[d.type_of[a]].zone_rad * d.c.res; }
    else { const int idx = s * d.c.P + p; x = PDF(d, PD_X, idx); y = PDF(d, PD_Y, idx); yaw = PDF(d, PD_YAW, idx); ext = d.ped_ext[p]; }
    const double reach = d.c.cull_reach + ext;
    int rel = 0;
    for (int j = threadIdx.x; j < d.c.R; j += STAMP_THREADS) {
        if (is_robot && j == a) continue;
        const int idx = s * d.c.R + j;
        const double dx = RBF(d, RB_X, idx) - x, dy = RBF(d, RB_Y, idx) - y;
        rel |= dx * dx + dy * dy <= reach * reach;
    }
    if (!__syncthreads_or(rel)) return;
    const Tf2 t = tf_from_pose(x, y, yaw);
    if (is_robot) {
        const RobotType& ty = d.types[d.type_of[a]];
        stamp_part(d, s, t, d.lattice_xy + 2 * (size_t)ty.pts_off, ty.n_pts, 0 + add, a, 0, 0, ty.stamp_cx, ty.stamp_cy, ty.stamp_rad, bm);
    } else {
        const int idx = s * d.c.P + p;
        const int shape = d.ped_shape[p];
        const double* pc = d.ped_part + 6 * (size_t)p;
        if (shape == 0) {
            stamp_part(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p], d.ped_pts_n[2 * p], 1 + add, p, 0, 0, pc[0], pc[1], (int)pc[2], bm);
        } else if (shape == 2) {
            stamp_part(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p], d.ped_pts_n[2 * p], 2 + add, p,
                       PDF(d, PD_LLX, idx), PDF(d, PD_LLY, idx), pc[0], pc[1], (int)pc[2], bm);
            stamp_part(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p + 1], d.ped_pts_n[2 * p + 1], 3 + add, p,
                       PDF(d, PD_RLX, idx), PDF(d, PD_RLY, idx), pc[3], pc[4], (int)pc[5], bm);
        }   // rectangle pedestrians are never drawn (img_env.cpp:599-616 has no branch for them)
    }
}

// reset objects: grid = n_scenes * max_obs CTAs. Object lattices are generated on the fly
// (agent.cpp:18-62) because their sizes change at every reset.
__global__ void k_stamp_objects(Dev d, const int* scene_ids, int unstamp) {
    int sl = blockIdx.x / d.c.max_obs, o = blockIdx.x % d.c.max_obs;
    int s = scene_ids ? scene_ids[sl] : sl;
    if (o >= d.n_obs[s]) return;
    const double* ob = d.obs + ((size_t)s * d.c.max_obs + o) * 8;
    int shape = (int)ob[0];
    Tf2 t = tf_from_pose(ob[5], ob[6], ob[7]);
    const double resolution = 0.01;
    int mode = unstamp ? 12 : 4;
    if (shape == 0) {
        int bb = (int)ceil(ob[3] / resolution);
        int side = 2 * bb + 1;
        for (int k = threadIdx.x; k < side * side; k += blockDim.x) {
            int m = k / side - bb, n = k % side - bb;
            if (sqrt(m * resolution * m * resolution + n * resolution * n * resolution) <= ob[3]) {
                double px = m * resolution + ob[1], py = n * resolution + ob[2];
                double wx, wy;
                tf_apply(t, px, py, wx, wy);
                stamp_object_cell(d, s, world2cell(wx, d.c.res), world2cell(wy, d.c.res), mode);
            }
        }
    } else if (shape == 1) {
        int x_min = (int)floor(ob[1] / resolution), x_max = (int)ceil(ob[2] / resolution);
        int y_min = (int)floor(ob[3] / resolution), y_max = (int)ceil(ob[4] / resolution);
        int ny = y_max - y_min + 1, nx = x_max - x_min + 1;
        for (int k = threadIdx.x; k < nx * ny; k += blockDim.x) {
            int m = x_min + k / ny, n = y_min + k % ny;
            double wx, wy;
            tf_apply(t, m * resolution, n * resolution, wx, wy);
            stamp_object_cell(d, s, world2cell(wx, d.c.res), world2cell(wy, d.c.res), mode);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the per-robot observation kernel
// ---------------------------------------------------------------------------------------------
struct ViewShared {
    Tf2 base_world, view_world, world_base;
    long long ax, bx, cx, ay, by, cy;   // fixed-point (2^-32 cell) affine view pixel -> world cell
    int frozen;
    int red[VIEW_THREADS / 32];
    int coll_key;
    // inverse (world cell -> view pixel) search: pixel = inv * (cell - org), in double; world block range of the FOV
    double inv[4], org[2];
    int blk[4];              // first block row / col, number of block rows / cols covering the FOV's world bounding box
    int wbb[4];              // world bounding box of the FOV in cells (x0, x1, y0, y1), clamped to the map
    int zc[2];               // robot position in world cells
};

// python float floor division (Objects/floatobject.c float_floor_div) used by yaml_env.py:414-415
__device__ __forceinline__ double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod) { if ((wx < 0) != (mod < 0)) { div -= 1.0; } }
    double floordiv;
    if (div) { floordiv = floor(div); if (div - floordiv > 0.5) floordiv += 1.0; }
    else floordiv = copysign(0.0, vx / wx);
    return floordiv;
}

// does ray k touch view cell (pr,pc)?  Closed form of the integer line walk in
// agent.cpp:517-622: with w=|x2-x1|, h=|y2-y1| the i-th visited cell is
//   w > h : (x1 + dx*i, y1 + dy*floor((2*h*i + w) / (2*w))),  i < w
//   else  : (x1 + dx*floor((2*w*i + h) / (2*h)), y1 + dy*i),  i < h
// Returns the step index i or -1.
__device__ __forceinline__ int ray_touch(int ox, int oy, int ex, int ey, int pr, int pc) {
    int w = ex - ox, h = ey - oy;
    int dx = w > 0 ? 1 : -1, dy = h > 0 ? 1 : -1;
    w = abs(w); h = abs(h);
    int a = pr - ox, b = pc - oy;
    if (w > h) {
        int i = abs(a);
        if (i >= w || a != dx * i) return -1;
        int m = abs(b);
        if (b != dy * m) return -1;
        int num = 2 * h * i + w;
        if (num < 2 * w * m || num >= 2 * w * (m + 1)) return -1;
        return i;
    } else {
        int i = abs(b);
        if (i >= h || b != dy * i) return -1;
        int m = abs(a);
        if (a != dx * m) return -1;
        int num = 2 * w * i + h;
        if (num < 2 * h * m || num >= 2 * h * (m + 1)) return -1;
        return i;
    }
}

// position of the n-th (0-based) set bit of m; m has more than n bits set.  (The generic __fns intrinsic costs ~55 instructions.)
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (n >= c) { n -= c; pos = 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (n >= c) { n -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xFu);    if (n >= c) { n -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);    if (n >= c) { n -= c; pos += 2; m >>= 2; }
    return pos + (n >= (int)(m & 1u) ? 1 : 0);
}

// Rare continuation of the laser_map pixel rule (phase D): the pixel lies behind its top ray's hit and was not shadow
// written by it, so the rays below (kh-1 .. kl) decide, highest first.  Out of line to keep the kernel's hot loop small.
__device__ __noinline__ unsigned pixel_code_below(const unsigned* hitkey, const short* rend, int ox, int oy, int pr, int pc, int kh, int kl) {
    for (int kk = kh - 1; kk >= kl; kk--) {
        const int i = ray_touch(ox, oy, rend[2 * kk], rend[2 * kk + 1], pr, pc);
        if (i < 0) continue;
        const unsigned key2 = hitkey[kk];
        const int hp2 = (int)(key2 >> 22);
        if (i < hp2) return 3u;
        if (i == hp2) return 0u;
        if (pr != (int)((key2 >> 11) & 2047) && pc != (int)(key2 & 2047)) return 2u;
    }
    return 2u;
}
// Ray-hit candidates of one raster cell, resolved on the spot (only when the shared-memory lists are full).
__device__ __noinline__ void cell_rays_inline(unsigned* hitkey, const short* rend, int ox, int oy, int pr, int pc, int kl, int kh) {
    for (int k = kl; k <= kh; k++) {
        const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
        if (i >= 0) atomicMin(&hitkey[k], ((unsigned)i << 22) | ((unsigned)pr << 11) | (unsigned)pc);
    }
}

// Shared-memory plan of k_view (bytes for the canonical 400x400 view / 1000 rays / 48x48 outputs; 48 KB -> 4 CTAs per SM):
//   region A  occupancy raster 400*13*4 = 20.8 KB (x2 with lasers off: + the "known" plane)            phases B-D
//   region B  ray-hit candidate lists (BL_CAP + BL2_CAP)*4 = 13.3 KB                                     phases B-C
//             later: horizontal resize buffer ns*HB_COLS*4 = 9.2 KB + hit prefix counts 2 KB            phases D-F
//   hitkey    range_total*4 (first hit per ray), ray end cells range_total*4, needed-line indices, FOV spans vh*8,
//             list of occupied world blocks under the FOV INV_MAX_BLOCKS*4
//   (no 400x400 pixel buffer: the laser_map values are evaluated inside the horizontal resize pass)
#define BL_CAP 3072          // candidate cells kept in shared memory; further cells are resolved inline by their finder
#define BL2_CAP 256          // cells touched by many rays (close to the origin): processed warp-cooperatively
#define BL_HEAVY 24
#define NOHIT 0xFFFFFFFFu

#define HB_COLS 16           // output columns per vertical-pass block (bounds the horizontal buffer)
#define INV_EPS 0.004f       // band around a cell edge inside which the inverse rasterisation runs the exact forward map
#define INV_MAX_BLOCKS 512   // 32x32-cell world blocks under the FOV; more -> forward (tile) rasterisation
struct ViewLayout { size_t sh, regA, regB, hpre, hitkey, rays, need, spans, blocks, total; };
__host__ __device__ inline ViewLayout view_layout(const Cfg& c) {
    ViewLayout L;
    size_t off = 0;
    L.sh = off; off += (sizeof(ViewShared) + 15) & ~(size_t)15;
    size_t occ = (size_t)c.vh * c.vwb * 4 * (c.use_laser ? 1 : 2);
    L.regA = off; off += (occ + 15) & ~(size_t)15;
    size_t bl = (size_t)(BL_CAP + BL2_CAP) * 4, hb = (((size_t)c.ns * HB_COLS * 4 + 15) & ~(size_t)15) + ((size_t)c.range_total + 1) * 2;
    L.regB = off; L.hpre = off + (((size_t)c.ns * HB_COLS * 4 + 15) & ~(size_t)15); off += ((bl > hb ? bl : hb) + 15) & ~(size_t)15;
    L.hitkey = off; off += ((size_t)c.range_total * 4 + 15) & ~(size_t)15;
    L.rays = off; off += ((size_t)c.range_total * 4 + 15) & ~(size_t)15;
    L.need = off; off += ((size_t)c.ns * 2 + 15) & ~(size_t)15;
    L.spans = off; off += ((size_t)c.vh * 8 + 15) & ~(size_t)15;
    L.blocks = off; off += (size_t)INV_MAX_BLOCKS * 4;
    L.total = off + 16;
    return L;
}
inline size_t view_smem_bytes(const Cfg& c) { return view_layout(c).total; }

__device__ __forceinline__ unsigned hit_key(int i, int x, int y) { return ((unsigned)i << 22) | ((unsigned)x << 11) | (unsigned)y; }

template <bool DEBUG_FULL>
__global__ void __launch_bounds__(VIEW_THREADS) k_view(Dev d, const int* scene_ids, int is_reset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Cfg& c = d.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sl = blockIdx.x / c.R, r = blockIdx.x % c.R;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int idx = s * c.R + r;
    const RobotType ty = d.types[d.type_of[r]];
    const ViewLayout L = view_layout(c);
    const int vwb = c.vwb, vh = c.vh, vw = c.vw;

    ViewShared* sh = reinterpret_cast<ViewShared*>(smem_raw + L.sh);
    uint32_t* occ = reinterpret_cast<uint32_t*>(smem_raw + L.regA);
    uint32_t* known = occ + (size_t)vh * vwb;                               // only when !use_laser
    uint32_t* blist = reinterpret_cast<uint32_t*>(smem_raw + L.regB);
    uint32_t* blist2 = blist + BL_CAP;
    int* hbuf = reinterpret_cast<int*>(smem_raw + L.regB);
    unsigned short* hpre = reinterpret_cast<unsigned short*>(smem_raw + L.hpre);   // after phase C: hpre[k] = #rays < k with a hit
    short* spans = reinterpret_cast<short*>(smem_raw + L.spans);
    uint32_t* blocks = reinterpret_cast<uint32_t*>(smem_raw + L.blocks);
    unsigned* hitkey = reinterpret_cast<unsigned*>(smem_raw + L.hitkey);
    short* rend = reinterpret_cast<short*>(smem_raw + L.rays);
    short* need = reinterpret_cast<short*>(smem_raw + L.need);

    if (tid == 0) {
        double x = RBF(d, RB_X, idx), y = RBF(d, RB_Y, idx), yaw = RBF(d, RB_YAW, idx);
        sh->base_world = tf_from_pose(x, y, yaw);
        sh->view_world = tf_mul(sh->base_world, c.view_base);       // get_view_world(), agent.cpp:128-131
        sh->world_base = tf_inv(sh->base_world);
        const Tf2& A = sh->view_world;
        sh->ax = llrint(A.m00 * FX_ONE); sh->bx = llrint(A.m01 * FX_ONE);
        sh->cx = llrint((A.ox / c.res) * FX_ONE) + (1ll << 31);
        sh->ay = llrint(A.m10 * FX_ONE); sh->by = llrint(A.m11 * FX_ONE);
        sh->cy = llrint((A.oy / c.res) * FX_ONE) + (1ll << 31);
        // Agent::view early-out (agent.cpp:358-360): stale view_map_/hits_/is_collision_ are re-sent
        sh->frozen = (RBF(d, RB_COLL, idx) != 0.0) || (RBF(d, RB_ARR, idx) != 0.0);
        sh->coll_key = 0;      // reused as boundary-list counters below
        sh->red[0] = 0; sh->red[1] = 0; sh->red[2] = 0; sh->red[3] = 0; sh->red[4] = 0;
        {   // inverse map and the FOV's world bounding box (for the world->view rasterisation)
            const double det = A.m00 * A.m11 - A.m01 * A.m10;
            sh->inv[0] = A.m11 / det; sh->inv[1] = -A.m01 / det; sh->inv[2] = -A.m10 / det; sh->inv[3] = A.m00 / det;
            sh->org[0] = A.ox / c.res; sh->org[1] = A.oy / c.res;
            int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
            for (int k = 0; k < 4; k++) {
                const int ii = (k & 1) ? ty.fov_r1 : ty.fov_r0, jj = (k & 2) ? ty.fov_c1 : ty.fov_c0;
                const int cx = (int)((sh->cx + (long long)ii * sh->ax + (long long)jj * sh->bx) >> 32);
                const int cy = (int)((sh->cy + (long long)ii * sh->ay + (long long)jj * sh->by) >> 32);
                xmin = min(xmin, cx); xmax = max(xmax, cx); ymin = min(ymin, cy); ymax = max(ymax, cy);
            }
            xmin = max(xmin - 1, 0); ymin = max(ymin - 1, 0); xmax = min(xmax + 1, c.H - 1); ymax = min(ymax + 1, c.W - 1);
            sh->wbb[0] = xmin; sh->wbb[1] = xmax; sh->wbb[2] = ymin; sh->wbb[3] = ymax;
            const bool empty = xmin > xmax || ymin > ymax || ty.fov_r1 < ty.fov_r0;
            sh->blk[0] = xmin >> 5; sh->blk[1] = ymin >> 5;
            sh->blk[2] = empty ? 0 : (xmax >> 5) - (xmin >> 5) + 1; sh->blk[3] = empty ? 0 : (ymax >> 5) - (ymin >> 5) + 1;
            sh->zc[0] = world2cell(x, c.res); sh->zc[1] = world2cell(y, c.res);
        }
    }
    {   // static tables into shared memory
        // one ray end = 2 shorts, one row of FOV spans = 4 shorts: copied as 4- and 8-byte words
        const int* g_rend = reinterpret_cast<const int*>(d.ray_end + 2 * (size_t)ty.ray_off);
        for (int k = tid; k < c.range_total; k += VIEW_THREADS) reinterpret_cast<int*>(rend)[k] = __ldg(g_rend + k);
        for (int k = tid; k < c.ns; k += VIEW_THREADS) need[k] = d.need_idx[k];
        const int2* g_spans = reinterpret_cast<const int2*>(d.fov_spans + (size_t)ty.span_off);
        for (int k = tid; k < vh; k += VIEW_THREADS) reinterpret_cast<int2*>(spans)[k] = __ldg(g_spans + k);
        for (int k = tid; k < c.range_total; k += VIEW_THREADS) hitkey[k] = NOHIT;
    }
    __syncthreads();
    const bool frozen = DEBUG_FULL ? false : sh->frozen != 0;

    if (!frozen) {
        // ---- Phase A: collision code = code of the LAST colliding lattice point (agent.cpp:294-326)
        int best = 0;
        if (!DEBUG_FULL) {
            const double* pts = d.lattice_xy + 2 * (size_t)ty.pts_off;
            for (int k = tid; k < ty.n_pts; k += VIEW_THREADS) {
                double wx, wy;
                const double2 pt = __ldg(reinterpret_cast<const double2*>(pts) + k);
                tf_apply(sh->base_world, pt.x, pt.y, wx, wy);
                int cx = world2cell_fast(wx, c.res, c.inv_res), cy = world2cell_fast(wy, c.res, c.inv_res);
                if ((unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W) {
                    int v = global_value(d, s, r, cx, cy);
                    if (v <= 2) best = max(best, ((k + 1) << 2) | (v + 1));
                }
            }
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        // ---- Phase B: egocentric occupancy raster (agent.cpp:373-404), 1 bit per view cell:
        // zero <=> in FOV && in map && global value < 250.  FOV = static column spans per row; the pixel ->
        // world cell map is affine and evaluated in 2^-32-cell fixed point with an exact fp64 fallback inside
        // a guard band around the rounding boundary.
        // The raster is produced in 32x32-pixel tiles: a tile whose world footprint (bounding box of its four
        // corner cells + 1 cell margin) only touches 32x32-cell blocks with a zero occupancy count
        // (Dev::coarse, maintained by the stamp kernels) is all-free and is skipped.
        const uint32_t* occ_all = d.occ_all + (size_t)s * c.H * c.Wb;
        const uint32_t* coarse = d.coarse + (size_t)s * c.Hc * c.Wb;
        const unsigned H = c.H, W = c.W, Wb = c.Wb;
        const int ox = ty.org_x, oy = ty.org_y;
        const uint32_t* kpack = d.kpack + (size_t)ty.khi_off;
        int* n_list = &sh->red[0]; int* n_list2 = &sh->red[1];
        // a newly set raster cell goes straight to the ray-hit candidate lists (phase C)
        // every ray of [k0, k0+kstep, ...] within the cell's static ray interval that really passes through it
        // keeps the minimum step: hitkey[k] = min(step << 22 | cell)
        auto cell_rays = [&](unsigned cell, unsigned kp, int k0, int kstep) {      // cell = row << 16 | col
            const int pr = (int)(cell >> 16), pc = (int)(cell & 0xFFFFu);
            const int kh = kp & 0xFFFF, kl = kp >> 16;
            for (int k = kl + k0; k <= kh; k += kstep) {
                const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
                if (i >= 0) atomicMin(&hitkey[k], hit_key(i, pr, pc));
            }
        };
        auto push_cell = [&](int pr, int pc) {
            const unsigned kp = __ldg(kpack + pr * vw + pc);
            const int kh = kp & 0xFFFF;
            if (kh == 0xFFFF) return;                                      // no ray passes through this cell
            const bool heavy = kh - (int)(kp >> 16) + 1 > BL_HEAVY;
            const int p = atomicAdd(heavy ? n_list2 : n_list, 1);
            if (p < (heavy ? BL2_CAP : BL_CAP)) (heavy ? blist2 : blist)[p] = ((unsigned)pr << 16) | (unsigned)pc;
            else cell_rays_inline(hitkey, rend, ox, oy, pr, pc, (int)(kp >> 16), kh);   // list full: resolve this cell right here
        };
        const int n_trow = (vh + 31) >> 5, n_tiles = n_trow * vwb;
        // (1) World -> view ("inverse") rasterisation, used when lasers are on: only raster cells that can be the
        //     first hit of a ray matter, and every such cell has a free 8-neighbour in the view, hence its world
        //     cell has a free cell within its 5x5 neighbourhood (a view step moves at most 2 world cells), is
        //     within 2 cells of the map border, lies next to the robot's own stamp, or the view cell sits on the
        //     FOV edge.  So: walk the occupied world words under the FOV (32x32 blocks with a non-zero count),
        //     drop 5x5-interior cells with word-parallel bit operations, map each remaining cell back to its
        //     <= 4 candidate view pixels and keep those whose EXACT forward map returns that cell; FOV-edge
        //     pixels (static list) are evaluated forward.  Every bit set is a truly occupied view cell and the
        //     first hit of every ray is among them.
        // (2) Otherwise (lasers off, or a huge FOV): forward rasterisation in 32x32-pixel tiles, skipping tiles
        //     whose world footprint only touches blocks with a zero count.
        const bool use_inverse = c.use_laser && sh->blk[2] * sh->blk[3] <= INV_MAX_BLOCKS;
        for (int q = tid; q < vh * vwb; q += VIEW_THREADS) { occ[q] = 0u; if (!c.use_laser) known[q] = 0u; }
        if (use_inverse) {
            const int nbj = sh->blk[3], nb = sh->blk[2] * nbj;
            for (int t = tid; t < nb; t += VIEW_THREADS) {
                const int bi = sh->blk[0] + t / nbj, bj = sh->blk[1] + t % nbj;
                if (__ldg(coarse + (unsigned)bi * Wb + bj)) blocks[atomicAdd(&sh->red[3], 1)] = ((unsigned)bi << 16) | (unsigned)bj;
            }
            if (!DEBUG_FULL && lane == 0) atomicMax(&sh->coll_key, best);
            __syncthreads();
            // FOV-edge pixels: forward
            const uint32_t* edge = d.edge_px + ty.edge_off;
            for (int e = tid; e < ty.n_edge; e += VIEW_THREADS) {
                const unsigned ep = __ldg(edge + e);
                const int i = ep >> 16, j = ep & 0xFFFF;
                const long long tx = sh->cx + (long long)i * sh->ax + (long long)j * sh->bx;
                const long long tyy = sh->cy + (long long)i * sh->ay + (long long)j * sh->by;
                int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                if ((unsigned)tx + FX_GUARD < 2 * FX_GUARD || (unsigned)tyy + FX_GUARD < 2 * FX_GUARD) {
                    double wx, wy;
                    tf_apply(sh->view_world, i * c.res, j * c.res, wx, wy);
                    cx = world2cell(wx, c.res); cy = world2cell(wy, c.res);
                }
                if ((unsigned)cx < H && (unsigned)cy < W) {
                    bool o = (__ldg(occ_all + (unsigned)cx * Wb + ((unsigned)cy >> 5)) >> (cy & 31)) & 1u;
                    if (o && i >= ty.zone_r0 && i <= ty.zone_r1 && j >= ty.zone_c0 && j <= ty.zone_c1) o = global_value(d, s, r, cx, cy) < 250;
                    if (o && !(atomicOr(&occ[i * vwb + (j >> 5)], 1u << (j & 31)) & (1u << (j & 31)))) push_cell(i, j);
                }
            }
            // occupied world words -> candidate cells -> view pixels
            const int n_blocks = sh->red[3], n_items = n_blocks * 32;
            const int X0 = sh->wbb[0], X1 = sh->wbb[1], Y0 = sh->wbb[2], Y1 = sh->wbb[3];
            const int zx = sh->zc[0], zy = sh->zc[1], zr = ty.zone_rad;
            const float i00 = (float)sh->inv[0], i01 = (float)sh->inv[1], i10 = (float)sh->inv[2], i11 = (float)sh->inv[3];
            const int orgi0 = (int)floor(sh->org[0]), orgi1 = (int)floor(sh->org[1]);
            const float orgf0 = (float)(sh->org[0] - orgi0), orgf1 = (float)(sh->org[1] - orgi1);
            const float f00 = (float)sh->view_world.m00, f01 = (float)sh->view_world.m01, f10 = (float)sh->view_world.m10, f11 = (float)sh->view_world.m11;
            // each warp takes 32 words (one per lane), then expands their candidate bits over all lanes
            // batches of 32 words are handed out dynamically: dense world blocks make the work per batch very uneven
            for (;;) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sh->red[4], 32);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_items) break;
                const int item = base + lane;
                unsigned cand = 0; int X = 0, bj = 0;
                if (item < n_items) {
                    const unsigned bb = blocks[item >> 5];
                    X = (int)(bb >> 16) * 32 + (item & 31); bj = (int)(bb & 0xFFFF);
                    unsigned w = 0;
                    if (X >= X0 && X <= X1) w = __ldg(occ_all + (unsigned)X * Wb + bj);
                    if (w) {
                        unsigned interior = 0xffffffffu;
                        for (int dxr = -2; dxr <= 2 && interior; dxr++) {
                            const int Xr = X + dxr;
                            if (Xr < 0 || Xr >= (int)H) { interior = 0; break; }
                            const uint32_t* rp = occ_all + (unsigned)Xr * Wb + bj;
                            const unsigned wc = __ldg(rp), wl = bj > 0 ? __ldg(rp - 1) : 0u, wr = bj + 1 < (int)Wb ? __ldg(rp + 1) : 0u;
                            interior &= wc & ((wc << 1) | (wl >> 31)) & ((wc << 2) | (wl >> 30)) & ((wc >> 1) | (wr << 31)) & ((wc >> 2) | (wr << 30));
                        }
                        cand = w & ~interior;
                        if (abs(X - zx) <= zr) {      // next to the robot's own stamp: no interior filter
                            const int lo = max(zy - zr - bj * 32, 0), hi = min(zy + zr - bj * 32, 31);
                            if (lo <= hi) cand |= w & ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
                        }
                        const int lo = max(Y0 - bj * 32, 0), hi = min(Y1 - bj * 32, 31);    // only columns under the FOV's bounding box
                        cand = lo <= hi ? cand & ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
                    }
                }
                const int cnt = __popc(cand);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                const int total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - cnt;
                for (int k0 = 0; k0 < total; k0 += 32) {
                    const int k = k0 + lane;
                    int src = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const int e = __shfl_sync(0xffffffffu, excl, (src + step) & 31);
                        if (src + step < 32 && e <= k) src += step;
                    }
                    const int n = k - __shfl_sync(0xffffffffu, excl, src);
                    const unsigned m = __shfl_sync(0xffffffffu, cand, src);
                    const int cX = X - lane + src, cbj = bj;      // a batch is the 32 rows of one block: row = first row + lane, same word column
                    if (k >= total) continue;
                    const int cY = cbj * 32 + nth_set_bit(m, n);
                    const float u = (float)(cX - orgi0) - orgf0, v = (float)(cY - orgi1) - orgf1;      // cell - org, exact integer part
                    const float qi = i00 * u + i01 * v, qj = i10 * u + i11 * v;
                    // Pixel (i,j) maps to this cell iff M*((i,j) - q) lies in the unit square around the cell centre (M =
                    // rotation of view_world).  The float test decides all pixels farther than INV_EPS from the square's
                    // edge; only the others run the exact forward map.  |q| < 2^10 so the float error is < 1e-3 cell.
                    const int ia = (int)ceilf(qi - 0.72f), ja = (int)ceilf(qj - 0.72f);
                    // cheap part for the 2 x 2 window at once (M*d is linear: the four offsets share two products), ...
                    const float du0 = (float)ia - qi, dv0 = (float)ja - qj;
                    const float x00 = f00 * du0 + f01 * dv0, y00 = f10 * du0 + f11 * dv0;
                    float e4[4];
                    e4[0] = fmaxf(fabsf(x00), fabsf(y00));
                    e4[1] = fmaxf(fabsf(x00 + f01), fabsf(y00 + f11));
                    e4[2] = fmaxf(fabsf(x00 + f00), fabsf(y00 + f10));
                    e4[3] = fmaxf(fabsf(x00 + f00 + f01), fabsf(y00 + f10 + f11));
                    unsigned live = (e4[0] <= 0.5f + INV_EPS ? 1u : 0u) | (e4[1] <= 0.5f + INV_EPS ? 2u : 0u) |
                                    (e4[2] <= 0.5f + INV_EPS ? 4u : 0u) | (e4[3] <= 0.5f + INV_EPS ? 8u : 0u);
                    // ... then the (usually one) surviving pixel
                    while (live) {
                        const int t = __ffs(live) - 1; live &= live - 1;
                        const int i = ia + (t >> 1), j = ja + (t & 1);
                        const float em = t == 0 ? e4[0] : t == 1 ? e4[1] : t == 2 ? e4[2] : e4[3];
                        if ((unsigned)i >= (unsigned)vh || (unsigned)j >= (unsigned)vw) continue;
                        const int a0 = spans[i * 4 + 0], a1 = spans[i * 4 + 1], b0 = spans[i * 4 + 2], b1 = spans[i * 4 + 3];
                        if (!((j >= a0 && j < a1) || (j >= b0 && j < b1))) continue;
                        if (em > 0.5f - INV_EPS) {
                            const long long tx = sh->cx + (long long)i * sh->ax + (long long)j * sh->bx;
                            const long long tyy = sh->cy + (long long)i * sh->ay + (long long)j * sh->by;
                            int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                            if ((unsigned)tx + FX_GUARD < 2 * FX_GUARD || (unsigned)tyy + FX_GUARD < 2 * FX_GUARD) {
                                double wx, wy;
                                tf_apply(sh->view_world, i * c.res, j * c.res, wx, wy);
                                cx = world2cell(wx, c.res); cy = world2cell(wy, c.res);
                            }
                            if (cx != cX || cy != cY) continue;
                        }
                        if (i >= ty.zone_r0 && i <= ty.zone_r1 && j >= ty.zone_c0 && j <= ty.zone_c1 && !(global_value(d, s, r, cX, cY) < 250)) continue;
                        if (!(atomicOr(&occ[i * vwb + (j >> 5)], 1u << (j & 31)) & (1u << (j & 31)))) push_cell(i, j);
                    }
                }
            }
        } else {
        int* n_active = &sh->red[2];
        unsigned short* tile_list = reinterpret_cast<unsigned short*>(blist);     // region B is free until phase C
        for (int t = tid; t < n_tiles; t += VIEW_THREADS) {
            if (!((d.tile_fov[ty.tile_off + (t >> 5)] >> (t & 31)) & 1u)) continue;
            bool active = !c.use_laser;          // the "known" plane needs every FOV pixel
            if (!active) {
                const int ti = t / vwb, tj = t - ti * vwb;
                const int i0 = ti * 32, i1 = min(i0 + 31, vh - 1), j0 = tj * 32, j1 = min(j0 + 31, vw - 1);
                int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int ii = (k & 1) ? i1 : i0, jj = (k & 2) ? j1 : j0;
                    const int cx = (int)((sh->cx + (long long)ii * sh->ax + (long long)jj * sh->bx) >> 32);
                    const int cy = (int)((sh->cy + (long long)ii * sh->ay + (long long)jj * sh->by) >> 32);
                    xmin = min(xmin, cx); xmax = max(xmax, cx); ymin = min(ymin, cy); ymax = max(ymax, cy);
                }
                xmin = max(xmin - 1, 0); ymin = max(ymin - 1, 0); xmax = min(xmax + 1, (int)H - 1); ymax = min(ymax + 1, (int)W - 1);
                for (int bi = xmin >> 5; bi <= (xmax >> 5) && !active; bi++)
                    for (int bj = ymin >> 5; bj <= (ymax >> 5); bj++)
                        if (__ldg(coarse + (unsigned)bi * Wb + bj)) { active = true; break; }
            }
            if (active) { const int ti = t / vwb; tile_list[atomicAdd(n_active, 1)] = (unsigned short)((ti << 8) | (t - ti * vwb)); }
        }
        if (!DEBUG_FULL && lane == 0) atomicMax(&sh->coll_key, best);
        __syncthreads();
        {
            const int n_items = sh->red[2] * 32;
            const long long lbx = sh->cx + (long long)lane * sh->bx, lby = sh->cy + (long long)lane * sh->by;
            const long long bx32 = sh->bx * 32, by32 = sh->by * 32;
#pragma unroll 2
            for (int item = warp; item < n_items; item += VIEW_THREADS / 32) {
                const int t = tile_list[item >> 5];
                const int wj = t & 255;
                const int i = (t >> 8) * 32 + (item & 31);
                if (i >= vh) continue;
                const int a0 = spans[i * 4 + 0], a1 = spans[i * 4 + 1], b0 = spans[i * 4 + 2], b1 = spans[i * 4 + 3];
                const int j = wj * 32 + lane;
                const bool in_fov = (j >= a0 && j < a1) || (j >= b0 && j < b1);     // empty spans are (-1,-1)
                bool o = false, kn = false;
                if (in_fov) {
                    const long long tx = lbx + (long long)i * sh->ax + (long long)wj * bx32;
                    const long long tyy = lby + (long long)i * sh->ay + (long long)wj * by32;
                    int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                    const unsigned lx = (unsigned)tx, ly = (unsigned)tyy;
                    if (lx + FX_GUARD < 2 * FX_GUARD || ly + FX_GUARD < 2 * FX_GUARD) {
                        double wx, wy;   // exact path: map2world, tf multiply, world2map
                        tf_apply(sh->view_world, i * c.res, j * c.res, wx, wy);
                        cx = world2cell(wx, c.res); cy = world2cell(wy, c.res);
                    }
                    if ((unsigned)cx < H && (unsigned)cy < W) {
                        kn = true;
                        o = (__ldg(occ_all + (unsigned)cx * Wb + ((unsigned)cy >> 5)) >> (cy & 31)) & 1u;
                        if (o && i >= ty.zone_r0 && i <= ty.zone_r1 && j >= ty.zone_c0 && j <= ty.zone_c1)
                            o = global_value(d, s, r, cx, cy) < 250;   // exclude the robot's own stamp
                    }
                }
                const unsigned wo = __ballot_sync(0xffffffffu, o);
                if (!c.use_laser) { const unsigned wk = __ballot_sync(0xffffffffu, kn); if (lane == 0) known[i * vwb + wj] = wk; }
                if (lane == 0) occ[i * vwb + wj] = wo;
            }
        }
        }   // forward tile path
        __syncthreads();
        if (!DEBUG_FULL && tid == 0) {
            int code = sh->coll_key & 3;
            RBF(d, RB_COLL, idx) = (double)code;
        }

        // ---- Phase C: first occupied cell of every laser ray (agent.cpp:405-438, 511-624).
        // A ray's hit cell always has a free 8-neighbour (its predecessor on the ray), so only the boundary
        // cells of the raster can be hits.  For each boundary cell the (static) interval of ray indices whose
        // integer line walk passes through it is scanned with the closed-form touch test and the ray keeps
        // the minimum step (atomicMin on step<<22|cell).  Cells that do not fit the lists are resolved inline.
        if (c.use_laser) {
            if (!use_inverse)
            for (int q = tid; q < vh * 16 * ((vwb + 15) / 16); q += VIEW_THREADS) {
                const int wpr = 16 * ((vwb + 15) / 16);                  // words per row rounded up to 16: shift/mask indexing
                const int i = (wpr == 16) ? (q >> 4) : q / wpr, wj = (wpr == 16) ? (q & 15) : q - i * wpr;
                if (wj >= vwb) continue;
                const unsigned O = occ[i * vwb + wj];
                if (!O) continue;
                unsigned all8 = 0xffffffffu;
                for (int di = -1; di <= 1; di++) {
                    const int ii = i + di;
                    unsigned m = 0, ml = 0, mr = 0;      // outside the raster counts as free (conservative superset)
                    if (ii >= 0 && ii < vh) {
                        m = occ[ii * vwb + wj];
                        ml = wj > 0 ? occ[ii * vwb + wj - 1] : 0u;
                        mr = wj + 1 < vwb ? occ[ii * vwb + wj + 1] : 0u;
                    }
                    const unsigned left = (m << 1) | (ml >> 31), right = (m >> 1) | (mr << 31);
                    all8 &= left & right;
                    if (di != 0) all8 &= m;
                }
                unsigned bnd = O & ~all8;
                if (i == ox && (oy >> 5) == wj) bnd |= O & (1u << (oy & 31));   // an occupied origin hits every ray at step 0
                while (bnd) {
                    const int b = __ffs(bnd) - 1; bnd &= bnd - 1;
                    const int col = wj * 32 + b;
                    if (col >= vw) break;
                    push_cell(i, col);
                }
            }
            __syncthreads();
            const int nl = min(sh->red[0], BL_CAP), nl2 = min(sh->red[1], BL2_CAP);
            if (d.dbg_stats && tid == 0) {
                int* st = d.dbg_stats + 4 * (size_t)idx;
                st[0] = sh->red[2] + sh->red[3]; st[1] = sh->red[0]; st[2] = sh->red[1];
                st[3] = sh->red[0] > BL_CAP || sh->red[1] > BL2_CAP;
            }
            for (int q = tid; q < nl; q += VIEW_THREADS) { const unsigned cell = blist[q]; cell_rays(cell, __ldg(kpack + (cell >> 16) * vw + (cell & 0xFFFFu)), 0, 1); }
            for (int q = warp; q < nl2; q += VIEW_THREADS / 32) { const unsigned cell = blist2[q]; cell_rays(cell, __ldg(kpack + (cell >> 16) * vw + (cell & 0xFFFFu)), lane, 32); }
            __syncthreads();
            if (!DEBUG_FULL) {
                for (int k = tid; k < c.range_total; k += VIEW_THREADS) {
                    const unsigned key = hitkey[k];
                    double hit = 6;   // agent.cpp:513
                    if (key != NOHIT) {
                        const int hx = (key >> 11) & 2047, hy = key & 2047;
                        double x0 = ox * c.res, y0 = oy * c.res, xc = hx * c.res, yc = hy * c.res;
                        hit = sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc));
                    }
                    const float wire = (float)hit;                         // AgentState.laser is float32[]
                    d.o_laser[(size_t)idx * c.range_total + k] = c.laser_norm ? (float)((double)wire / c.laser_max) : wire;   // yaml_env.py:440-444
                }
                // prefix counts of the rays that hit something: phase D/F asks "any hit among rays [a, b]?" in O(1)
                const int per = (c.range_total + VIEW_THREADS - 1) / VIEW_THREADS;
                const int k0 = min(tid * per, c.range_total), k1 = min(k0 + per, c.range_total);
                int cnt = 0;
                for (int k = k0; k < k1; k++) cnt += hitkey[k] != NOHIT;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (lane == 31) sh->red[warp] = incl;      // the list counters in red[] are dead after the barrier above
                __syncthreads();
                int run = incl - cnt;
                for (int w = 0; w < warp; w++) run += sh->red[w];
                for (int k = k0; k < k1; k++) { hpre[k] = (unsigned short)run; run += hitkey[k] != NOHIT; }
                if (k1 == c.range_total) hpre[k1] = (unsigned short)run;
                __syncthreads();
            }
        }

        // ---- Phase D/E/F: laser_map reconstruction fused into the cubic resize.
        // D/E: the final view_map_ value of a pixel = last-writer-wins over the rays in index order, evaluated from
        //      the highest touching ray downwards (static tables), then the robot's own footprint (100, agent.cpp:503).
        //      dtab packs, per pixel the resize reads, the top ray, its step index there and the own-footprint bit:
        //      the top ray touches by construction, so the closed-form touch test only runs on fall-through.
        // F:   cv2.resize(INTER_CUBIC) 400->48 (yaml_env.py:433-434), OpenCV's own path: horizontal pass in int32 with
        //      11-bit weights, vertical pass as an fp32 FMA chain with weights * 2^-22, round-half-even, saturate; then
        //      float16(x)/255 via a host-built table.  Every source pixel with a non-zero weight belongs to exactly one
        //      output column, so the horizontal pass evaluates its (<= 4) pixels on the fly: no pixel buffer.
        const uint32_t* own_mask = d.own_mask + (size_t)ty.own_mask_off;
        const uint32_t* dtab = d.dtab + (size_t)ty.dtab_off;
        // value code of tap k of output column oc on needed row rr: 0 -> 0, 1 -> 100, 2 -> 200, 3 -> 255.
        // e is the tap's dtab entry; the source column is only looked up on the rare fall-through path.
        auto pixel_code = [&](unsigned e, int rr, const short* tp, int k) -> unsigned {
            unsigned code = 2u; bool own;
            if (c.use_laser) {
                own = e >> 31;
                const int kh = e & 0xFFF;
                if (kh != 0xFFF) {
                    const unsigned key = hitkey[kh];
                    const int hp = (int)(key >> 22), i0 = (e >> 12) & 0x3FF;
                    if (i0 < hp) code = 3u;
                    else if (i0 == hp) code = 0u;
                    else {
                        const int pr = need[rr], pc = need[tp[k]];
                        if (!(pr != (int)((key >> 11) & 2047) && pc != (int)(key & 2047))) {     // no shadow write: fall through to lower rays
                            const unsigned kp = __ldg(kpack + pr * vw + pc);
                            code = pixel_code_below(hitkey, rend, ox, oy, pr, pc, kh, (int)(kp >> 16));
                        }
                    }
                }
            } else {
                const int pr = need[rr], pc = need[tp[k]], full = pr * vw + pc;
                const bool o = (occ[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                const bool kn = (known[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                code = o ? 0u : (kn ? 3u : 2u);
                own = (own_mask[full >> 5] >> (full & 31)) & 1u;
            }
            if (code != 0u && own) code = 1u;
            return code;
        };
        if (DEBUG_FULL) {
            // whole 400x400 raster for the tests (generic per-pixel path)
            for (int rr = warp; rr < vh; rr += VIEW_THREADS / 32) {
                for (int cc = lane; cc < vw; cc += 32) {
                    const int pr = rr, pc = cc, full = pr * vw + pc;
                    int val = 200;
                    if (c.use_laser) {
                        const unsigned kp = __ldg(kpack + full);
                        const int kh = kp & 0xFFFF;
                        if (kh != 0xFFFF) {
                            const int kl = kp >> 16;
                            for (int k = kh; k >= kl; k--) {
                                const int i = ray_touch(ox, oy, rend[2 * k], rend[2 * k + 1], pr, pc);
                                if (i < 0) continue;
                                const unsigned key = hitkey[k];
                                const int hp = (int)(key >> 22);
                                if (i < hp) { val = 255; break; }
                                if (i == hp) { val = 0; break; }
                                if (pr != (int)((key >> 11) & 2047) && pc != (int)(key & 2047)) { val = 200; break; }   // shadow write (agent.cpp:557-558)
                            }
                        }
                    } else {
                        const bool o = (occ[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                        const bool kn = (known[pr * vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                        val = o ? 0 : (kn ? 255 : 200);
                    }
                    if (val != 0 && pr >= ty.zone_r0 && pr <= ty.zone_r1 && pc >= ty.zone_c0 && pc <= ty.zone_c1 &&
                        ((own_mask[full >> 5] >> (full & 31)) & 1u)) val = 100;
                    if (d.dbg_view) d.dbg_view[(size_t)idx * vh * vw + full] = (uint8_t)val;
                }
            }
        } else {
            const float scale = 1.f / (2048.f * 2048.f);
            for (int cb = 0; cb < c.img; cb += HB_COLS) {
                const int nc = min(HB_COLS, c.img - cb);
                // A warp covers a compact patch of 8 needed rows x 4 output columns (few rays cross it, so the whole warp often
                // takes the hit-free shortcut below); with 8 warps and 16 columns per block a thread keeps its output column
                // and walks down the needed rows in steps of 16.
                static_assert(VIEW_THREADS == 256 && HB_COLS == 16, "item mapping of the horizontal resize pass");
                const int ocl = (warp & 3) * 4 + (lane & 3);
                if (ocl < nc) {
                    const int oc = cb + ocl;
                    const short* tp = d.cubic_tap + 4 * oc;
                    const short4 cf = __ldg(reinterpret_cast<const short4*>(d.cubic_coef) + oc);
                    const int rr0 = (warp >> 2) * 8 + (lane >> 2);
                    const uint2* hsp = reinterpret_cast<const uint2*>(d.hstat) + (size_t)(ty.dtab_off >> 2) + oc;
                    const uint4* dtp = reinterpret_cast<const uint4*>(dtab) + oc;
                    for (int rr = rr0; rr < c.ns; rr += 16) {
                        uint4 e4 = make_uint4(0u, 0u, 0u, 0u);
                        if (c.use_laser) {
                            // static shortcut: when none of the top rays of this output's taps hit anything, all of its source
                            // pixels keep their hit-free value (free / own footprint / outside every ray) and the sum is a table entry
                            const uint2 hs = __ldg(hsp + rr * c.img);
                            const int kmin = hs.x & 0xFFFFu, kmax = hs.x >> 16;
                            if (kmax < kmin || hpre[kmax + 1] == hpre[kmin]) { hbuf[rr * HB_COLS + ocl] = (int)hs.y; continue; }
                            e4 = __ldg(dtp + rr * c.img);
                        }
                        int acc = 0;
                        if (cf.x) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.x, rr, tp, 0))) & 0xFFu) * cf.x;
                        if (cf.y) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.y, rr, tp, 1))) & 0xFFu) * cf.y;
                        if (cf.z) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.z, rr, tp, 2))) & 0xFFu) * cf.z;
                        if (cf.w) acc += (int)((0xFFC86400u >> (8 * pixel_code(e4.w, rr, tp, 3))) & 0xFFu) * cf.w;
                        hbuf[rr * HB_COLS + ocl] = acc;
                    }
                }
                __syncthreads();
                for (int q = tid; q < c.img * HB_COLS; q += VIEW_THREADS) {
                    const int orow = q / HB_COLS, ocl = q % HB_COLS;
                    if (ocl >= nc) continue;
                    const short4 tp = __ldg(reinterpret_cast<const short4*>(d.cubic_tap) + orow), cf = __ldg(reinterpret_cast<const short4*>(d.cubic_coef) + orow);
                    const float b0 = cf.x * scale, b1 = cf.y * scale, b2 = cf.z * scale, b3 = cf.w * scale;
                    const float s0 = (float)hbuf[tp.x * HB_COLS + ocl], s1 = (float)hbuf[tp.y * HB_COLS + ocl];
                    const float s2 = (float)hbuf[tp.z * HB_COLS + ocl], s3 = (float)hbuf[tp.w * HB_COLS + ocl];
                    const float v = fmaf(s0, b0, fmaf(s1, b1, fmaf(s2, b2, s3 * b3)));
                    int iv = __float2int_rn(v);
                    iv = min(255, max(0, iv));
                    d.o_sensor[(size_t)idx * c.img * c.img + orow * c.img + cb + ocl] = d.f16_lut[iv];
                }
                __syncthreads();
            }
        }
    }
    if (DEBUG_FULL) return;

    // ---- Phase G: state vector and the episode bookkeeping (img_env.cpp:547-587, yaml_env.py:316, 374-376, 467-471).
    //      The pedestrian observation does not depend on the raster: k_ped_obs below, on its own stream.
    if (tid == 0) {
        double st[5];
        robot_state_vec(RBF(d, RB_X, idx), RBF(d, RB_Y, idx), RBF(d, RB_YAW, idx), RBF(d, RB_GX, idx), RBF(d, RB_GY, idx),
                        RBF(d, RB_GYAW, idx), RBF(d, RB_L0V, idx), RBF(d, RB_L0W, idx), c.state_dim, st);
        float s0 = (float)st[0], s1 = (float)st[1];
        for (int k = 0; k < c.state_dim; k++) d.o_vec[(size_t)idx * c.state_dim + k] = (float)st[k];
        double dist = sqrt((double)s0 * (double)s0 + (double)s1 * (double)s1);   // yaml_env.py:467
        double prev = RBF(d, RB_PREVD, idx);
        d.o_stepd[idx] = isnan(prev) ? 0.f : (float)(prev - dist);
        RBF(d, RB_PREVD, idx) = dist;
        int coll = (int)RBF(d, RB_COLL, idx), arr = RBF(d, RB_ARR, idx) != 0.0;
        d.o_coll[idx] = (int8_t)coll;
        d.o_arr[idx] = (uint8_t)arr;
        RBF(d, RB_DONE, idx) = is_reset ? 0.0 : (double)min(1, min(coll, 1) + arr);   // yaml_env.py:316, 374-376
    }
}

// ---------------------------------------------------------------------------------------------
// Pedestrian observation of every robot (img_env.cpp:566-583 ped_info; yaml_env.py:392-466 _get_states /
// _draw_ped_map): pedestrians in the robot frame, nearest first (stable sort), ped_vector_states, ped_min_dists
// (NearbyPed persistence) and the 3 x img x img ped_maps where farther pedestrians overwrite nearer ones.
// It only reads poses, so it runs beside the stamp / view kernels on the library's side stream.
// ---------------------------------------------------------------------------------------------
#define PED_THREADS 128
inline size_t ped_smem_bytes(const Cfg& c) {
    return (((size_t)c.img * c.img * 4 + 15) & ~(size_t)15) + (size_t)((c.P + 1) & ~1) * 8 + (size_t)c.P * 16 + (size_t)c.P * 4 + 16;
}
__global__ void __launch_bounds__(PED_THREADS) k_ped_obs(Dev d, const int* scene_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Cfg& c = d.c;
    const int tid = threadIdx.x;
    const int sl = blockIdx.x / c.R, r = blockIdx.x % c.R;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int idx = s * c.R + r;
    const RobotType& ty = d.types[d.type_of[r]];
    int* winner = reinterpret_cast<int*>(smem_raw);
    double* pkey = reinterpret_cast<double*>(smem_raw + (((size_t)c.img * c.img * 4 + 15) & ~(size_t)15));
    float* pobs = reinterpret_cast<float*>(pkey + ((c.P + 1) & ~1));
    int* prank = reinterpret_cast<int*>(pobs + 4 * (size_t)c.P);
    const Tf2 world_base = tf_inv(tf_from_pose(RBF(d, RB_X, idx), RBF(d, RB_Y, idx), RBF(d, RB_YAW, idx)));
    for (int k = tid; k < c.img * c.img; k += PED_THREADS) winner[k] = -1;
    float* pvs = d.o_pvs + (size_t)idx * c.pvs_len;
    for (int k = tid; k < c.pvs_len; k += PED_THREADS) pvs[k] = k == 0 ? (float)c.P : 0.f;
    for (int j = tid; j < c.P; j += PED_THREADS) {
        int pi = s * c.P + j;
        double bx, by, bvx, bvy;
        tf_apply(world_base, PDF(d, PD_X, pi), PDF(d, PD_Y, pi), bx, by);
        tf_rotate(world_base, PDF(d, PD_VX, pi), PDF(d, PD_VY, pi), bvx, bvy);
        float px = (float)bx, py = (float)by;
        pobs[4 * j] = px; pobs[4 * j + 1] = py; pobs[4 * j + 2] = (float)bvx; pobs[4 * j + 3] = (float)bvy;
        pkey[j] = (double)px * (double)px + (double)py * (double)py;
    }
    __syncthreads();
    for (int j = tid; j < c.P; j += PED_THREADS) {   // stable rank == python's list.sort(key=...)
        const double kj = pkey[j];
        const int jb = j - (tid & 31);       // a warp ranks 32 consecutive pedestrians: ties only need care inside that group
        int rk = 0;
        for (int i = 0; i < jb; i++) rk += pkey[i] <= kj;
        for (int i = jb; i < min(jb + 32, c.P); i++) rk += (pkey[i] < kj) || (pkey[i] == kj && i < j);
        for (int i = jb + 32; i < c.P; i++) rk += pkey[i] < kj;
        prank[j] = rk;
    }
    __syncthreads();
    for (int j = tid; j < c.P; j += PED_THREADS) {
        int q = prank[j];
        double px = pobs[4 * j], py = pobs[4 * j + 1];
        double ped_r = d.ped_r_round[j];
        float f5 = (float)ped_r, f6 = (float)(ped_r + ty.size_last), f7 = (float)sqrt(px * px + py * py);
        if (q < c.max_ped) {
            float* o = pvs + 1 + (size_t)q * c.ped_vec_dim;
            o[0] = pobs[4 * j]; o[1] = pobs[4 * j + 1]; o[2] = pobs[4 * j + 2]; o[3] = pobs[4 * j + 3];
            o[4] = f5; o[5] = f6; o[6] = f7;
        }
        if (q == 0) {   // NearbyPed.set(i, ped_tmp[7] - ped_tmp[6]) in float32 (yaml_env.py:455-456)
            float md = f7 - f6;
            RBF(d, RB_MIND, idx) = (double)md;
        }
        if (px > 3 || px < -3 || py > 3 || py < -3) continue;
        double tmx = -px + 3, tmy = -py + 3;
        int x0 = (int)py_floordiv(tmx - c.ped_image_r, c.ped_res), x1 = (int)py_floordiv(tmx + c.ped_image_r, c.ped_res);
        int y0 = (int)py_floordiv(tmy - c.ped_image_r, c.ped_res), y1 = (int)py_floordiv(tmy + c.ped_image_r, c.ped_res);
        for (int jj = x0; jj < x1; jj++)
            for (int kk = y0; kk < y1; kk++) {
                if (jj < 0 || jj >= c.img || kk < 0 || kk >= c.img) continue;
                double ddx = (jj + 0.5) * c.ped_res - tmx, ddy = (kk + 0.5) * c.ped_res - tmy;
                if (ddx * ddx + ddy * ddy < c.ped_image_r * c.ped_image_r) atomicMax(&winner[jj * c.img + kk], (q << 16) | j);
            }
    }
    __syncthreads();
    if (tid == 0) d.o_mind[idx] = (float)RBF(d, RB_MIND, idx);
    float* pm = d.o_pmap + (size_t)idx * 3 * c.img * c.img;
    const int npm = c.img * c.img;
    for (int ch = 0; ch < 3; ch++)
        for (int cell = tid; cell < npm; cell += PED_THREADS) {
            int wv = winner[cell];
            float v = 0.f;
            if (wv >= 0) { int j = wv & 0xFFFF; v = ch == 0 ? 1.0f : pobs[4 * j + 1 + ch]; }
            pm[ch * npm + cell] = v;
        }
}
