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
#include "kin.cuh"

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
// Pose-dependent constants of one robot's observation: the transforms Agent::view builds (get_base_world / get_view_world,
// agent.cpp:118-131), the pixel -> world-cell map in fixed point, its inverse and the world bounding box of the field of
// view.  Computed once per step by the robot's footprint warp (below) and bulk-copied (cp.async.bulk) into the observation
// kernel's shared memory, so that kernel starts without a serial fp64 prologue.
// ---------------------------------------------------------------------------------------------------------------------
#define FX_ONE 4294967296.0            // 2^32: fixed-point scale of cell coordinates
struct __align__(16) ViewConst {
    Tf2 base_world, view_world;
    long long ax, bx, cx, ay, by, cy;   // fixed-point (2^-32 cell) affine view pixel -> world cell, rounding offset folded in
    double inv[4], org[2];              // inverse (world cell -> view pixel): pixel = inv * (cell - org)
    int blk[4];                         // first block row / col, number of block rows / cols covering the FOV's world bounding box
    int wbb[4];                         // world bounding box of the FOV in cells (x0, x1, y0, y1), clamped to the map
    int4 own_hdr;                       // box of the observer's own footprint (header of the record k_footprints gives it)
    int frozen, pad[3];                 // Agent::view early-out (agent.cpp:358-360)
};
static_assert(sizeof(ViewConst) % 16 == 0, "ViewConst is moved with 16-byte bulk copies");

__device__ __forceinline__ void view_const_compute(const Dev& d, int idx, int r, ViewConst* k) {
    const Cfg& c = d.c;
    const RobotType& ty = d.types[d.type_of[r]];
    const double x = RBF(d, RB_X, idx), y = RBF(d, RB_Y, idx), yaw = RBF(d, RB_YAW, idx);
    const Tf2 B = tf_from_pose(x, y, yaw);
    const Tf2 A = tf_mul(B, c.view_base);                        // get_view_world(), agent.cpp:128-131
    k->base_world = B; k->view_world = A;
    const long long ax = llrint(A.m00 * FX_ONE), bx = llrint(A.m01 * FX_ONE), cx = llrint((A.ox / c.res) * FX_ONE) + (1ll << 31);
    const long long ay = llrint(A.m10 * FX_ONE), by = llrint(A.m11 * FX_ONE), cy = llrint((A.oy / c.res) * FX_ONE) + (1ll << 31);
    k->ax = ax; k->bx = bx; k->cx = cx; k->ay = ay; k->by = by; k->cy = cy;
    // stale view_map_/hits_/is_collision_ are re-sent for robots that collided or arrived
    k->frozen = (RBF(d, RB_COLL, idx) != 0.0) || (RBF(d, RB_ARR, idx) != 0.0);
    k->pad[0] = k->pad[1] = k->pad[2] = 0;
    {   // the box of the robot's own footprint (the record itself may be culled when no other robot is near)
        double bwx, bwy;
        tf_apply(B, ty.stamp_cx, ty.stamp_cy, bwx, bwy);
        k->own_hdr = foot_pack(foot_box(bwx, bwy, ty.stamp_rad, c.res), FK_ROBOT, r);
    }
    const double det = A.m00 * A.m11 - A.m01 * A.m10;
    k->inv[0] = A.m11 / det; k->inv[1] = -A.m01 / det; k->inv[2] = -A.m10 / det; k->inv[3] = A.m00 / det;
    k->org[0] = A.ox / c.res; k->org[1] = A.oy / c.res;
    int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
    for (int q = 0; q < 4; q++) {
        const int ii = (q & 1) ? ty.fov_r1 : ty.fov_r0, jj = (q & 2) ? ty.fov_c1 : ty.fov_c0;
        const int px = (int)((cx + (long long)ii * ax + (long long)jj * bx) >> 32);
        const int py = (int)((cy + (long long)ii * ay + (long long)jj * by) >> 32);
        xmin = min(xmin, px); xmax = max(xmax, px); ymin = min(ymin, py); ymax = max(ymax, py);
    }
    xmin = max(xmin - 1, 0); ymin = max(ymin - 1, 0); xmax = min(xmax + 1, c.H - 1); ymax = min(ymax + 1, c.W - 1);
    k->wbb[0] = xmin; k->wbb[1] = xmax; k->wbb[2] = ymin; k->wbb[3] = ymax;
    const bool empty = xmin > xmax || ymin > ymax || ty.fov_r1 < ty.fov_r0;
    k->blk[0] = xmin >> 5; k->blk[1] = ymin >> 5;
    k->blk[2] = empty ? 0 : (xmax >> 5) - (xmin >> 5) + 1; k->blk[3] = empty ? 0 : (ymax >> 5) - (ymin >> 5) + 1;
}

// mbarrier + 1-D bulk copy (TMA unit, global -> shared) helpers; sizes and addresses are multiples of 16 bytes
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Phase G of the observation: state vector and the episode bookkeeping that does not depend on the raster
// (img_env.cpp:547-587, yaml_env.py:467-471); the collision code / done flag follow in k_view.
__device__ __forceinline__ void view_state_vector(const Dev& d, int idx) {
    const Cfg& c = d.c;
    double st[5];
    robot_state_vec(RBF(d, RB_X, idx), RBF(d, RB_Y, idx), RBF(d, RB_YAW, idx), RBF(d, RB_GX, idx), RBF(d, RB_GY, idx),
                    RBF(d, RB_GYAW, idx), RBF(d, RB_L0V, idx), RBF(d, RB_L0W, idx), c.state_dim, st);
    float s0 = (float)st[0], s1 = (float)st[1];
    for (int k = 0; k < c.state_dim; k++) d.o_vec[(size_t)idx * c.state_dim + k] = (float)st[k];
    double dist = sqrt((double)s0 * (double)s0 + (double)s1 * (double)s1);   // yaml_env.py:467
    double prev = RBF(d, RB_PREVD, idx);
    d.o_stepd[idx] = isnan(prev) ? 0.f : (float)(prev - dist);
    RBF(d, RB_PREVD, idx) = dist;
    d.o_arr[idx] = (uint8_t)(RBF(d, RB_ARR, idx) != 0.0);
}

// one thread per robot (a serial fp64 routine: small CTAs spread it over the SMs); grid = ceil(n_scenes * R / VC_THREADS)
#define VC_THREADS 64
__global__ void __launch_bounds__(VC_THREADS) k_view_consts(Dev d, const int* scene_ids, int n_scenes, int debug_only) {
    const Cfg& c = d.c;
    const int g = blockIdx.x * VC_THREADS + threadIdx.x;
    const int sl = g / c.R, a = g - sl * c.R;
    if (sl >= n_scenes || (d.n_dev && sl >= *d.n_dev)) return;
    const int idx = (scene_ids ? scene_ids[sl] : sl) * c.R + a;
    view_const_compute(d, idx, a, d.vconst + idx);
    if (!debug_only) view_state_vector(d, idx);
}

// ---------------------------------------------------------------------------------------------------------------------
// robots and pedestrians: grid = ceil(n_scenes * NPA / FOOT_WARPS) CTAs, one warp per part
// ---------------------------------------------------------------------------------------------------------------------
#define FOOT_WARPS 8
__global__ void __launch_bounds__(FOOT_WARPS * 32, 4) k_footprints(Dev d, const int* scene_ids, int n_scenes, int flags) {      // flags: 1 = step_++, 2 = evaluate whole lattices (tests)
    extern __shared__ uint32_t foot_sm[];             // FOOT_WARPS * ag_cap words
    const Cfg& c = d.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = blockIdx.x * FOOT_WARPS + warp;
    if (gq >= n_scenes * c.NPA || (d.n_dev && gq >= *d.n_dev * c.NPA)) return;
    const int sl = gq / c.NPA, a = gq - sl * c.NPA;
    const int s = scene_ids ? scene_ids[sl] : sl;
    if ((flags & 1) && a == 0 && lane == 0) d.step_no[s] += 1;      // step_++ (img_env.cpp:518): the dynamics stage of this step is done
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
    int ring_n = -1; const double* ring = nullptr; double dcx = 0, dcy = 0, r_in = 0;      // circle parts: rim points + sure-covered disc
    if (is_robot) {
        const int idx = s * c.R + a;
        const RobotType& ty = d.types[d.type_of[a]];
        x = RBF(d, RB_X, idx); y = RBF(d, RB_Y, idx); yaw = RBF(d, RB_YAW, idx); ext = ty.zone_rad * c.res;
        kind = FK_ROBOT; id = a; n_pts = ty.n_pts; pts = d.lattice_xy + 2 * (size_t)ty.pts_off; rad = ty.stamp_rad; ccx = ty.stamp_cx; ccy = ty.stamp_cy;
        ring_n = ty.ring_n; ring = d.lattice_xy + 2 * (size_t)ty.ring_off; dcx = ty.disc_cx; dcy = ty.disc_cy; r_in = ty.disc_rin;
    } else {
        const int idx = s * c.P + p;
        const int shape = d.ped_shape[p];
        // rectangle pedestrians are never drawn (img_env.cpp:599-616 has no branch for them); circles have one part
        if (!(shape == 0 && leg == 0) && shape != 2) { if (lane == 0) *hdr = make_int4(0, 0, 0, 0); return; }
        x = PDF(d, PD_X, idx); y = PDF(d, PD_Y, idx); yaw = PDF(d, PD_YAW, idx); ext = d.ped_ext[p];
        const double* pc = d.ped_part + 6 * (size_t)p + 3 * leg;
        kind = shape == 0 ? FK_CIRC : (leg ? FK_RIGHT : FK_LEFT); id = p;
        n_pts = d.ped_pts_n[2 * p + leg]; pts = d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p + leg]; rad = (int)pc[2]; ccx = pc[0]; ccy = pc[1];
        ring_n = d.ped_ring_n[2 * p + leg]; ring = d.lattice_xy + 2 * (size_t)d.ped_ring_off[2 * p + leg];
        { const double* dz = d.ped_disc + 6 * (size_t)p + 3 * leg; dcx = dz[0]; dcy = dz[1]; r_in = dz[2]; }
        if (shape == 2) {   // leg2base: identity rotation + leg origin (agent.cpp:815-837)
            offx = leg ? PDF(d, PD_RLX, idx) : PDF(d, PD_LLX, idx); offy = leg ? PDF(d, PD_RLY, idx) : PDF(d, PD_LLY, idx);
        }
    }
    // A part can only be read by a robot whose collision lattice or view raster reaches it: parts farther than
    // (view half-diagonal + own extent) from every (other) robot get an empty record.
    {
        const double reach = c.cull_reach + ext;
        bool any = false;
        for (int j0 = 0; j0 < c.R && !any; j0 += 32) {      // 32 robots at a time, stop at the first batch with a robot in reach
            const int j = j0 + lane;
            bool rel = false;
            if (j < c.R && !(is_robot && j == a)) {
                const int idx = s * c.R + j;
                const double dx = RBF(d, RB_X, idx) - x, dy = RBF(d, RB_Y, idx) - y;
                rel = dx * dx + dy * dy <= reach * reach;
            }
            any = __any_sync(0xffffffffu, rel);
        }
        if (!any) { if (lane == 0) *hdr = make_int4(0, 0, 0, 0); return; }
    }
    const Tf2 t = tf_from_pose(x, y, yaw);
    double bwx, bwy;
    tf_apply(t, ccx + offx, ccy + offy, bwx, bwy);      // world position of the part's bounding-circle centre
    const FootBox bx = foot_box(bwx, bwy, rad, c.res);
    const int nw = bx.nrow * bx.wpr;                    // <= cap by construction of stamp_bitmap_words
    for (int k = lane; k < nw; k += 32) bm[k] = 0u;
    __syncwarp();
    if (ring_n >= 0 && !(flags & 2)) {
        // Circle lattice: every cell whose centre lies within r_in of the disc centre holds a lattice point for sure
        // (host_tables.h, lattice_circle_ring) -- those cells are set row by row, and only the rim points are evaluated.
        double wcx, wcy;
        tf_apply(t, dcx + offx, dcy + offy, wcx, wcy);
        for (int rr = lane; rr < bx.nrow; rr += 32) {
            const int X = bx.cx0 + rr;
            if ((unsigned)X >= (unsigned)c.H) continue;
            const double dx = X * c.res - wcx, h2 = r_in * r_in - dx * dx;
            if (h2 <= 0.0) continue;
            const double half = sqrt(h2);
            int y_lo = (int)ceil((wcy - half) * c.inv_res + 1e-9), y_hi = (int)floor((wcy + half) * c.inv_res - 1e-9);
            y_lo = max(y_lo, 0); y_hi = min(y_hi, c.W - 1);
            for (int w = 0; w < bx.wpr; w++) {
                const int lo = max(y_lo - (bx.wj0 + w) * 32, 0), hi = min(y_hi - (bx.wj0 + w) * 32, 31);
                if (lo <= hi) bm[rr * bx.wpr + w] = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
            }
        }
        __syncwarp();
        pts = ring; n_pts = ring_n;
    }
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
