// Footprint records: the cells every robot, pedestrian part and reset object covers, kept as small per-part bitmaps
// instead of map-sized per-scene planes.
//
// Reference semantics being reproduced: ImgEnv::view_ped / view_robot (img_env.cpp:594-674) clone the whole map once per
// robot and draw every other agent into the clone with Agent::draw (agent.cpp:285-327) / PedAgent::draw_leg (:737-774);
// reset objects are drawn into obs_map_ at reset (img_env.cpp:187).  A cell's final byte only depends on WHICH KINDS of
// things cover it (the writers never overwrite 0/1/2 except the right-leg quirk), so it is enough to know, per part, the
// set of cells round(T * p / res) over its 0.01 m lattice points p (agent.cpp:18-62) -- exactly what a record holds:
//
//   header  int4 { cx0, cy0, n, kind | id << 4 }   square box of n x n cells: rows cx0 .. cx0+n-1 (world x), columns
//                                                   cy0 .. cy0+n-1 (world y), stored as 32-cell words wj0 = cy0 >> 5 ..
//   occ     n * wpr u32 words (wpr words per row, aligned with the map's bit planes), bit = cell covered
//           (cells outside the map are never set)
//   cand    the same box: covered cells that have a cell NOT covered by this part within their 5x5 neighbourhood.
//           Only such cells can be the first hit of a laser ray (view.cuh, phase B), so the observation kernel usually
//           reads `cand` and skips the interior of every footprint.
//
// Records of robots and pedestrians are rebuilt every step by k_footprints (one warp per part, shared-memory bitmap, no
// global atomics, nothing to undo afterwards); records of reset objects by k_object_footprints at reset.  The observation
// kernel gathers the records whose box meets its field of view and composes them on the fly.
#pragma once
#include "state.cuh"

inline __host__ __device__ int stamp_rad_cells(double ext, double res) { return (int)ceil(ext / res) + 1; }
inline __host__ __device__ int stamp_bitmap_words(int rad_cells) { return (2 * rad_cells + 1) * ((2 * rad_cells) / 32 + 2); }

struct FootBox { int cx0, cy0, nrow, wj0, wpr; };
__device__ __forceinline__ FootBox foot_box(double x, double y, int rad_cells, double res) {
    const int ccx = world2cell(x, res), ccy = world2cell(y, res);
    FootBox b;
    b.cx0 = ccx - rad_cells; b.cy0 = ccy - rad_cells; b.nrow = 2 * rad_cells + 1;
    b.wj0 = b.cy0 >> 5; b.wpr = ((ccy + rad_cells) >> 5) - b.wj0 + 1;
    return b;
}
__device__ __forceinline__ int4 foot_pack(const FootBox& b, int kind, int id) { return make_int4(b.cx0, b.cy0, b.nrow, kind | (id << 4)); }
__device__ __forceinline__ int foot_nrow(const int4& h) { return h.z; }
__device__ __forceinline__ int foot_wj0(const int4& h) { return h.y >> 5; }
__device__ __forceinline__ int foot_wpr(const int4& h) { return ((h.y + h.z - 1) >> 5) - (h.y >> 5) + 1; }
__device__ __forceinline__ int foot_kind(const int4& h) { return h.w & 15; }
__device__ __forceinline__ int foot_id(const int4& h) { return h.w >> 4; }
__device__ __forceinline__ unsigned foot_flag(int kind) {
    return kind == FK_ROBOT ? F_ROBOT : kind == FK_CIRC ? F_CIRC : kind == FK_LEFT ? F_LEFT : kind == FK_RIGHT ? F_RIGHT : F_OBJ;
}
// does the part cover cell (cx, cy)?
__device__ __forceinline__ bool foot_covers(const int4& h, const uint32_t* occ, int cx, int cy) {
    const int r = cx - h.x;
    if ((unsigned)r >= (unsigned)h.z || (unsigned)(cy - h.y) >= (unsigned)h.z) return false;
    return (occ[r * foot_wpr(h) + (cy >> 5) - foot_wj0(h)] >> (cy & 31)) & 1u;
}
// The byte Agent::draw / Agent::view would read from the observer's global_map_ at a cell with static byte sv that is
// covered by the kinds in f (F_ROBOT = a robot other than the observer): obs_map_ (static + reset objects, which write 0
// except on 0/1/2) -> peds_map_ (circle: 1 except on 0/1/2; left leg: 1 except on 0; right leg: always 1, the draw_leg
// quirk agent.cpp:757-772) -> other robots (2 except on 0/1/2).  Order independent.
__device__ __forceinline__ int composed_value(int sv, unsigned f) {
    if ((f & F_OBJ) && sv > 2) sv = 0;
    int v;
    if (f & F_RIGHT) v = 1;
    else if (f & F_LEFT) v = (sv == 0) ? 0 : 1;
    else if (f & F_CIRC) v = (sv <= 2) ? sv : 1;
    else v = sv;
    if (v > 2 && (f & F_ROBOT)) v = 2;
    return v;
}

// cand = occ & ~(5x5-interior of occ), evaluated inside the record's own bitmap (outside the box nothing is covered)
__device__ __forceinline__ unsigned foot_cand_word(const uint32_t* bm, int nrow, int wpr, int r, int w) {
    const unsigned o = bm[r * wpr + w];
    if (!o) return 0u;
    unsigned interior = 0xffffffffu;
    for (int dr = -2; dr <= 2 && interior; dr++) {
        const int rr = r + dr;
        if (rr < 0 || rr >= nrow) { interior = 0; break; }
        const uint32_t* rp = bm + rr * wpr + w;
        const unsigned wc = rp[0], wl = w > 0 ? rp[-1] : 0u, wr = w + 1 < wpr ? rp[1] : 0u;
        interior &= wc & ((wc << 1) | (wl >> 31)) & ((wc << 2) | (wl >> 30)) & ((wc >> 1) | (wr << 31)) & ((wc >> 2) | (wr << 30));
    }
    return o & ~interior;
}

// ---------------------------------------------------------------------------------------------------------------------
// robots and pedestrians: grid = ceil(n_scenes * NPA / FOOT_WARPS) CTAs, one warp per part
// ---------------------------------------------------------------------------------------------------------------------
#define FOOT_WARPS 8
__global__ void __launch_bounds__(FOOT_WARPS * 32) k_footprints(Dev d, const int* scene_ids, int n_scenes, int bump_step) {
    extern __shared__ uint32_t foot_sm[];             // FOOT_WARPS * ag_cap words
    const Cfg& c = d.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = blockIdx.x * FOOT_WARPS + warp;
    if (gq >= n_scenes * c.NPA || (d.n_dev && gq >= *d.n_dev * c.NPA)) return;
    const int sl = gq / c.NPA, a = gq - sl * c.NPA;
    const int s = scene_ids ? scene_ids[sl] : sl;
    if (bump_step && a == 0 && lane == 0) d.step_no[s] += 1;      // step_++ (img_env.cpp:518): the dynamics stage of this step is done
    uint32_t* bm = foot_sm + (size_t)warp * c.ag_cap;
    int4* hdr = d.foot_hdr + (size_t)s * c.NP + a;
    uint32_t* out = d.foot_words + (size_t)s * c.scene_words + d.part_off[a];
    const int cap = (d.part_off[a + 1] - d.part_off[a]) >> 1;
    // which part: robot body, pedestrian body / left leg (part 0), right leg (part 1)
    const bool is_robot = a < c.R;
    const int p = is_robot ? 0 : (a - c.R) >> 1, leg = is_robot ? 0 : (a - c.R) & 1;
    double x, y, yaw, ext;
    int kind, id, n_pts, rad;
    const double* pts;
    double offx = 0, offy = 0, ccx, ccy;
    if (is_robot) {
        const int idx = s * c.R + a;
        const RobotType& ty = d.types[d.type_of[a]];
        x = RBF(d, RB_X, idx); y = RBF(d, RB_Y, idx); yaw = RBF(d, RB_YAW, idx); ext = ty.zone_rad * c.res;
        kind = FK_ROBOT; id = a; n_pts = ty.n_pts; pts = d.lattice_xy + 2 * (size_t)ty.pts_off; rad = ty.stamp_rad; ccx = ty.stamp_cx; ccy = ty.stamp_cy;
    } else {
        const int idx = s * c.P + p;
        const int shape = d.ped_shape[p];
        // rectangle pedestrians are never drawn (img_env.cpp:599-616 has no branch for them); circles have one part
        if (!(shape == 0 && leg == 0) && shape != 2) { if (lane == 0) *hdr = make_int4(0, 0, 0, 0); return; }
        x = PDF(d, PD_X, idx); y = PDF(d, PD_Y, idx); yaw = PDF(d, PD_YAW, idx); ext = d.ped_ext[p];
        const double* pc = d.ped_part + 6 * (size_t)p + 3 * leg;
        kind = shape == 0 ? FK_CIRC : (leg ? FK_RIGHT : FK_LEFT); id = p;
        n_pts = d.ped_pts_n[2 * p + leg]; pts = d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p + leg]; rad = (int)pc[2]; ccx = pc[0]; ccy = pc[1];
        if (shape == 2) {   // leg2base: identity rotation + leg origin (agent.cpp:815-837)
            offx = leg ? PDF(d, PD_RLX, idx) : PDF(d, PD_LLX, idx); offy = leg ? PDF(d, PD_RLY, idx) : PDF(d, PD_LLY, idx);
        }
    }
    // A part can only be read by a robot whose collision lattice or view raster reaches it: parts farther than
    // (view half-diagonal + own extent) from every (other) robot get an empty record.
    {
        const double reach = c.cull_reach + ext;
        bool rel = false;
        for (int j = lane; j < c.R; j += 32) {
            if (is_robot && j == a) continue;
            const int idx = s * c.R + j;
            const double dx = RBF(d, RB_X, idx) - x, dy = RBF(d, RB_Y, idx) - y;
            rel |= dx * dx + dy * dy <= reach * reach;
        }
        if (!__any_sync(0xffffffffu, rel)) { if (lane == 0) *hdr = make_int4(0, 0, 0, 0); return; }
    }
    const Tf2 t = tf_from_pose(x, y, yaw);
    double bwx, bwy;
    tf_apply(t, ccx + offx, ccy + offy, bwx, bwy);      // world position of the part's bounding-circle centre
    const FootBox bx = foot_box(bwx, bwy, rad, c.res);
    const int nw = bx.nrow * bx.wpr;                    // <= cap by construction of stamp_bitmap_words
    for (int k = lane; k < nw; k += 32) bm[k] = 0u;
    __syncwarp();
    for (int k = lane; k < n_pts; k += 32) {
        const double2 pt = __ldg(reinterpret_cast<const double2*>(pts) + k);
        double wx, wy;
        tf_apply(t, pt.x + offx, pt.y + offy, wx, wy);
        const int cx = world2cell_fast(wx, c.res, c.inv_res), cy = world2cell_fast(wy, c.res, c.inv_res);
        const int r = cx - bx.cx0, w = (cy >> 5) - bx.wj0;
        if ((unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W && (unsigned)r < (unsigned)bx.nrow && (unsigned)w < (unsigned)bx.wpr)
            atomicOr(&bm[r * bx.wpr + w], 1u << (cy & 31));
    }
    __syncwarp();
    for (int k = lane; k < nw; k += 32) {
        const int r = k / bx.wpr, w = k - r * bx.wpr;
        out[k] = bm[k];
        out[cap + k] = foot_cand_word(bm, bx.nrow, bx.wpr, r, w);
    }
    if (lane == 0) *hdr = foot_pack(bx, kind, id);
}

// ---------------------------------------------------------------------------------------------------------------------
// reset objects: grid = n_scenes * max_obs CTAs.  Lattices are generated on the fly (agent.cpp:18-62): sizes change at
// every reset.  obs record: shape, size[4], x, y, yaw.
// ---------------------------------------------------------------------------------------------------------------------
#define OBJ_THREADS 128
__host__ __device__ inline void object_bounds(const double* ob, double& ccx, double& ccy, double& rmax) {
    const double resolution = 0.01;
    if ((int)ob[0] == 0) { ccx = ob[1]; ccy = ob[2]; rmax = ob[3]; }
    else {
        const int x_min = (int)floor(ob[1] / resolution), x_max = (int)ceil(ob[2] / resolution);
        const int y_min = (int)floor(ob[3] / resolution), y_max = (int)ceil(ob[4] / resolution);
        ccx = 0.5 * (x_min + x_max) * resolution; ccy = 0.5 * (y_min + y_max) * resolution;
        rmax = 0.5 * hypot((double)(x_max - x_min), (double)(y_max - y_min)) * resolution;
    }
}
inline __host__ __device__ int object_rad_cells(double rmax, double res) { return (int)ceil(rmax / res) + 2; }

__global__ void __launch_bounds__(OBJ_THREADS) k_object_footprints(Dev d, const int* scene_ids) {
    extern __shared__ uint32_t foot_sm[];             // obj_cap words
    const Cfg& c = d.c;
    const int sl = blockIdx.x / c.max_obs, o = blockIdx.x % c.max_obs;
    if (d.n_dev && sl >= *d.n_dev) return;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int q = c.NPA + o;
    int4* hdr = d.foot_hdr + (size_t)s * c.NP + q;
    if (o >= d.n_obs[s]) { if (threadIdx.x == 0) *hdr = make_int4(0, 0, 0, 0); return; }
    uint32_t* out = d.foot_words + (size_t)s * c.scene_words + d.part_off[q];
    const int cap = c.obj_cap;
    const double* ob = d.obs + ((size_t)s * c.max_obs + o) * 8;
    const int shape = (int)ob[0];
    const Tf2 t = tf_from_pose(ob[5], ob[6], ob[7]);
    const double resolution = 0.01;
    double ccx, ccy, rmax;
    object_bounds(ob, ccx, ccy, rmax);
    double bwx, bwy;
    tf_apply(t, ccx, ccy, bwx, bwy);
    const FootBox bx = foot_box(bwx, bwy, min(object_rad_cells(rmax, c.res), c.obj_rad), c.res);   // (the host rejects larger objects)
    const int nw = bx.nrow * bx.wpr;
    for (int k = threadIdx.x; k < nw; k += OBJ_THREADS) foot_sm[k] = 0u;
    __syncthreads();
    auto put = [&](double px, double py) {
        double wx, wy;
        tf_apply(t, px, py, wx, wy);
        const int cx = world2cell(wx, c.res), cy = world2cell(wy, c.res);
        const int r = cx - bx.cx0, w = (cy >> 5) - bx.wj0;
        if ((unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W && (unsigned)r < (unsigned)bx.nrow && (unsigned)w < (unsigned)bx.wpr)
            atomicOr(&foot_sm[r * bx.wpr + w], 1u << (cy & 31));
    };
    if (shape == 0) {
        const int bb = (int)ceil(ob[3] / resolution);
        const int side = 2 * bb + 1;
        for (int k = threadIdx.x; k < side * side; k += OBJ_THREADS) {
            const int m = k / side - bb, n = k % side - bb;
            if (sqrt(m * resolution * m * resolution + n * resolution * n * resolution) <= ob[3]) put(m * resolution + ob[1], n * resolution + ob[2]);
        }
    } else if (shape == 1) {
        const int x_min = (int)floor(ob[1] / resolution), x_max = (int)ceil(ob[2] / resolution);
        const int y_min = (int)floor(ob[3] / resolution), y_max = (int)ceil(ob[4] / resolution);
        const int ny = y_max - y_min + 1, nx = x_max - x_min + 1;
        for (int k = threadIdx.x; k < nx * ny; k += OBJ_THREADS) put((x_min + k / ny) * resolution, (y_min + k % ny) * resolution);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nw; k += OBJ_THREADS) {
        const int r = k / bx.wpr, w = k - r * bx.wpr;
        out[k] = foot_sm[k];
        out[cap + k] = foot_cand_word(foot_sm, bx.nrow, bx.wpr, r, w);
    }
    if (threadIdx.x == 0) *hdr = foot_pack(bx, FK_OBJ, o);
}
