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
__device__ __forceinline__ unsigned short rmin_min(unsigned short* p, unsigned short id) {
    unsigned short old = *p;
    while (old > id) {
        unsigned short assumed = old;
        old = atomicCAS(p, assumed, id);
        if (old == assumed) break;
    }
    return old;
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

// mode: 0 stamp robot(id), 1 stamp ped body (circle), 2 left leg, 3 right leg, 4 object,
//       8+ = unstamp of (mode-8)
__device__ __forceinline__ void stamp_cell(const Dev& d, int s, int cx, int cy, int mode, int id) {
    if ((unsigned)cx >= (unsigned)d.c.H || (unsigned)cy >= (unsigned)d.c.W) return;
    size_t ci = (size_t)cx * d.c.W + cy;
    size_t po = (size_t)s * plane_cells(d.c);
    uint32_t* occ = d.occ_all + (size_t)s * d.c.H * d.c.Wb + (size_t)cx * d.c.Wb + (cy >> 5);
    uint32_t bit = 1u << (cy & 31);
    if (mode < 8) {
        if (mode == 0) {
            unsigned short old = rmin_min(d.rmin + po + ci, (unsigned short)id);
            unsigned bits = F_ROBOT;
            if (old != RMIN_EMPTY && old != (unsigned short)id) bits |= F_MULTI;
            flag_or(d.flags + po, ci, bits);
        } else {
            flag_or(d.flags + po, ci, mode == 1 ? F_CIRC : mode == 2 ? F_LEFT : mode == 3 ? F_RIGHT : F_OBJ);
        }
        atomicOr(occ, bit);
    } else {
        int m = mode - 8;
        if (m == 0) { d.rmin[po + ci] = RMIN_EMPTY; flag_clear(d.flags + po, ci, F_ROBOT | F_MULTI); }
        else flag_clear(d.flags + po, ci, m == 1 ? F_CIRC : m == 2 ? F_LEFT : m == 3 ? F_RIGHT : F_OBJ);
        // restore the occupancy bit to static | object (dynamic stamps never survive a step)
        bool base = (d.static_occ[(size_t)cx * d.c.Wb + (cy >> 5)] & bit) != 0;
        if (m != 4) base = base || (d.flags[po + ci] & F_OBJ);
        if (!base) atomicAnd(occ, ~bit);
    }
}

__device__ __forceinline__ void stamp_points(const Dev& d, int s, const Tf2& t, const double* pts, int n, int mode, int id,
                                             double offx, double offy) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double bx = pts[2 * k], by = pts[2 * k + 1];
        if (mode == 2 || mode == 3 || mode == 10 || mode == 11) {   // leg2base: identity rotation + leg origin (agent.cpp:815-837)
            bx = bx + offx;
            by = by + offy;
        }
        double wx, wy;
        tf_apply(t, bx, by, wx, wy);
        stamp_cell(d, s, world2cell(wx, d.c.res), world2cell(wy, d.c.res), mode, id);
    }
}

// grid = n_scenes * (R + P) CTAs; `unstamp` selects the inverse operation.
__global__ void k_stamp_agents(Dev d, const int* scene_ids, int unstamp) {
    int per = d.c.R + d.c.P;
    int sl = blockIdx.x / per, a = blockIdx.x % per;
    int s = scene_ids ? scene_ids[sl] : sl;
    int add = unstamp ? 8 : 0;
    if (a < d.c.R) {
        int idx = s * d.c.R + a;
        Tf2 t = tf_from_pose(RBF(d, RB_X, idx), RBF(d, RB_Y, idx), RBF(d, RB_YAW, idx));
        const RobotType& ty = d.types[d.type_of[a]];
        stamp_points(d, s, t, d.lattice_xy + 2 * (size_t)ty.pts_off, ty.n_pts, 0 + add, a, 0, 0);
    } else {
        int p = a - d.c.R;
        int idx = s * d.c.P + p;
        Tf2 t = tf_from_pose(PDF(d, PD_X, idx), PDF(d, PD_Y, idx), PDF(d, PD_YAW, idx));
        int shape = d.ped_shape[p];
        if (shape == 0) {
            stamp_points(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p], d.ped_pts_n[2 * p], 1 + add, p, 0, 0);
        } else if (shape == 2) {
            stamp_points(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p], d.ped_pts_n[2 * p], 2 + add, p,
                         PDF(d, PD_LLX, idx), PDF(d, PD_LLY, idx));
            stamp_points(d, s, t, d.lattice_xy + 2 * (size_t)d.ped_pts_off[2 * p + 1], d.ped_pts_n[2 * p + 1], 3 + add, p,
                         PDF(d, PD_RLX, idx), PDF(d, PD_RLY, idx));
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
                stamp_cell(d, s, world2cell(wx, d.c.res), world2cell(wy, d.c.res), mode, 0);
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
            stamp_cell(d, s, world2cell(wx, d.c.res), world2cell(wy, d.c.res), mode, 0);
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

template <bool DEBUG_FULL>
__global__ void __launch_bounds__(VIEW_THREADS) k_view(Dev d, const int* scene_ids, int is_reset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Cfg& c = d.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sl = blockIdx.x / c.R, r = blockIdx.x % c.R;
    const int s = scene_ids ? scene_ids[sl] : sl;
    const int idx = s * c.R + r;
    const RobotType ty = d.types[d.type_of[r]];

    // ---- shared memory carve-up ----
    unsigned char* sp = smem_raw;
    ViewShared* sh = reinterpret_cast<ViewShared*>(sp); sp += (sizeof(ViewShared) + 15) & ~15;
    uint32_t* occ = reinterpret_cast<uint32_t*>(sp); sp += (size_t)c.vh * c.vwb * 4;
    uint32_t* known = reinterpret_cast<uint32_t*>(sp); if (!c.use_laser) sp += (size_t)c.vh * c.vwb * 4;
    unsigned short* hitpos = reinterpret_cast<unsigned short*>(sp); sp += ((size_t)c.range_total * 2 + 15) & ~15;
    short* hitx = reinterpret_cast<short*>(sp); sp += ((size_t)c.range_total * 2 + 15) & ~15;
    short* hity = reinterpret_cast<short*>(sp); sp += ((size_t)c.range_total * 2 + 15) & ~15;
    uint8_t* pix = sp; sp += ((size_t)c.ns * c.ns + 15) & ~15;
    int* hbuf = reinterpret_cast<int*>(sp); sp += (size_t)c.ns * c.img * 4;
    int* winner = reinterpret_cast<int*>(sp); sp += (size_t)c.img * c.img * 4;
    double* pkey = reinterpret_cast<double*>(sp); sp += (size_t)((c.P + 1) & ~1) * 8;
    float* pobs = reinterpret_cast<float*>(sp); sp += (size_t)c.P * 4 * 4;
    int* prank = reinterpret_cast<int*>(sp); sp += (size_t)c.P * 4;

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
        sh->coll_key = 0;
    }
    for (int k = tid; k < c.img * c.img; k += VIEW_THREADS) winner[k] = -1;
    __syncthreads();
    const bool frozen = DEBUG_FULL ? false : sh->frozen != 0;

    if (!frozen) {
        // ---- Phase A: collision code = code of the LAST colliding lattice point (agent.cpp:294-326)
        if (!DEBUG_FULL) {
            int best = 0;
            const double* pts = d.lattice_xy + 2 * (size_t)ty.pts_off;
            for (int k = tid; k < ty.n_pts; k += VIEW_THREADS) {
                double wx, wy;
                tf_apply(sh->base_world, pts[2 * k], pts[2 * k + 1], wx, wy);
                int cx = world2cell(wx, c.res), cy = world2cell(wy, c.res);
                if ((unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W) {
                    int v = global_value(d, s, r, cx, cy);
                    if (v <= 2) best = max(best, ((k + 1) << 2) | (v + 1));
                }
            }
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0) sh->red[warp] = best;
        }
        // zero the raster while the reduction settles
        for (int k = tid; k < c.vh * c.vwb; k += VIEW_THREADS) { occ[k] = 0; if (!c.use_laser) known[k] = 0; }
        __syncthreads();
        if (!DEBUG_FULL && tid == 0) {
            int best = 0;
            for (int k = 0; k < VIEW_THREADS / 32; k++) best = max(best, sh->red[k]);
            int code = best & 3;
            RBF(d, RB_COLL, idx) = (double)code;
            sh->coll_key = code;
        }

        // ---- Phase B: egocentric occupancy raster (agent.cpp:373-404), 1 bit per view cell.
        // zero <=> in FOV && in map && global value < 250.  The FOV test depends only on the pixel
        // (static spans); the pixel -> world cell map is affine: evaluated in 2^-32-cell fixed point
        // with an exact fp64 fallback inside a guard band around the rounding boundary.
        const short* spans = d.fov_spans + (size_t)ty.span_off;
        const uint32_t* occ_all = d.occ_all + (size_t)s * c.H * c.Wb;
        for (int i = warp; i < c.vh; i += VIEW_THREADS / 32) {
            long long rowx = sh->cx + (long long)i * sh->ax, rowy = sh->cy + (long long)i * sh->ay;
            for (int sp_i = 0; sp_i < MAX_SPANS; sp_i++) {
                int c0 = spans[(i * MAX_SPANS + sp_i) * 2], c1 = spans[(i * MAX_SPANS + sp_i) * 2 + 1];
                if (c0 < 0) continue;
                for (int j0 = c0 & ~31; j0 < c1; j0 += 32) {
                    int j = j0 + lane;
                    bool o = false, kn = false;
                    if (j >= c0 && j < c1) {
                        long long tx = rowx + (long long)j * sh->bx, tyy = rowy + (long long)j * sh->by;
                        int cx = (int)(tx >> 32), cy = (int)(tyy >> 32);
                        unsigned lx = (unsigned)tx, ly = (unsigned)tyy;
                        if (lx + FX_GUARD < 2 * FX_GUARD || ly + FX_GUARD < 2 * FX_GUARD) {
                            double wx, wy;   // exact path: map2world, tf multiply, world2map
                            tf_apply(sh->view_world, i * c.res, j * c.res, wx, wy);
                            cx = world2cell(wx, c.res); cy = world2cell(wy, c.res);
                        }
                        if ((unsigned)cx < (unsigned)c.H && (unsigned)cy < (unsigned)c.W) {
                            kn = true;
                            o = (occ_all[(size_t)cx * c.Wb + (cy >> 5)] >> (cy & 31)) & 1u;
                            if (o && i >= ty.zone_r0 && i <= ty.zone_r1 && j >= ty.zone_c0 && j <= ty.zone_c1)
                                o = global_value(d, s, r, cx, cy) < 250;   // exclude the robot's own stamp
                        }
                    }
                    unsigned wo = __ballot_sync(0xffffffffu, o);
                    unsigned wk = __ballot_sync(0xffffffffu, kn);
                    if (lane == 0) {
                        // a word may be shared by two spans of the same row: OR (same warp, sequential)
                        occ[i * c.vwb + (j0 >> 5)] |= wo;
                        if (!c.use_laser) known[i * c.vwb + (j0 >> 5)] |= wk;
                    }
                }
            }
        }
        __syncthreads();

        // ---- Phase C: laser rays (agent.cpp:405-438, 511-624): one thread per ray marches the bit raster
        const short* rend = d.ray_end + 2 * (size_t)ty.ray_off;
        if (c.use_laser) {
            for (int k = tid; k < c.range_total; k += VIEW_THREADS) {
                int x1 = ty.org_x, y1 = ty.org_y, x2 = rend[2 * k], y2 = rend[2 * k + 1];
                int w = x2 - x1, h = y2 - y1;
                int dx = w > 0 ? 1 : -1, dy = h > 0 ? 1 : -1;
                w = abs(w); h = abs(h);
                int hp = 0xFFFF, hx = -1, hy = -1;
                int x = x1, y = y1, f;
                if (w > h) {
                    f = 2 * h - w;
                    for (int i = 0; x != x2; x += dx, i++) {
                        if ((unsigned)x >= (unsigned)c.vh || (unsigned)y >= (unsigned)c.vw) break;
                        if ((occ[x * c.vwb + (y >> 5)] >> (y & 31)) & 1u) { hp = i; hx = x; hy = y; break; }
                        if (f < 0) f += 2 * h; else { y += dy; f += 2 * (h - w); }
                    }
                } else {
                    f = 2 * w - h;
                    for (int i = 0; y != y2; y += dy, i++) {
                        if ((unsigned)x >= (unsigned)c.vh || (unsigned)y >= (unsigned)c.vw) break;
                        if ((occ[x * c.vwb + (y >> 5)] >> (y & 31)) & 1u) { hp = i; hx = x; hy = y; break; }
                        if (f < 0) f += 2 * w; else { x += dx; f += 2 * (w - h); }
                    }
                }
                hitpos[k] = (unsigned short)hp; hitx[k] = (short)hx; hity[k] = (short)hy;
                if (!DEBUG_FULL) {
                    double hit = 6;   // agent.cpp:513
                    if (hp != 0xFFFF) {
                        double x0 = x1 * c.res, y0 = y1 * c.res, xc = hx * c.res, yc = hy * c.res;
                        hit = sqrt((x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc));
                    }
                    float wire = (float)hit;                         // AgentState.laser is float32[]
                    d.o_laser[(size_t)idx * c.range_total + k] =
                        c.laser_norm ? (float)((double)wire / c.laser_max) : wire;   // yaml_env.py:440-444
                }
            }
        }
        __syncthreads();

        // ---- Phase D/E: final view_map_ value of every pixel the cubic resize reads (or of the whole
        // raster in debug mode): last-writer-wins over rays in index order evaluated per pixel from the
        // highest ray downwards (static khi/klo tables), then the robot's own footprint (value 100,
        // agent.cpp:503) unless the cell is 0.
        const unsigned short* khi = d.khi + (size_t)ty.khi_off;
        const unsigned short* klo = d.klo + (size_t)ty.khi_off;
        const uint32_t* own_mask = d.own_mask + (size_t)ty.own_mask_off;
        const int npx = DEBUG_FULL ? c.vh * c.vw : c.ns * c.ns;
        for (int q = tid; q < npx; q += VIEW_THREADS) {
            int pr, pc;
            if (DEBUG_FULL) { pr = q / c.vw; pc = q % c.vw; }
            else { pr = d.need_idx[q / c.ns]; pc = d.need_idx[q % c.ns]; }
            int full = pr * c.vw + pc;
            int val = 200;
            if (c.use_laser) {
                int kh = khi[full];
                if (kh != 0xFFFF) {
                    int kl = klo[full];
                    for (int k = kh; k >= kl; k--) {
                        int i = ray_touch(ty.org_x, ty.org_y, rend[2 * k], rend[2 * k + 1], pr, pc);
                        if (i < 0) continue;
                        int hp = hitpos[k];
                        if (i < hp) { val = 255; break; }
                        if (i == hp) { val = 0; break; }
                        if (pr != hitx[k] && pc != hity[k]) { val = 200; break; }   // shadow write (agent.cpp:557-558)
                    }
                }
            } else {
                bool o = (occ[pr * c.vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                bool kn = (known[pr * c.vwb + (pc >> 5)] >> (pc & 31)) & 1u;
                val = o ? 0 : (kn ? 255 : 200);
            }
            if (val != 0 && ((own_mask[full >> 5] >> (full & 31)) & 1u)) val = 100;
            if (DEBUG_FULL) { if (d.dbg_view) d.dbg_view[(size_t)idx * c.vh * c.vw + full] = (uint8_t)val; }
            else pix[q] = (uint8_t)val;
        }
        __syncthreads();

        if (!DEBUG_FULL) {
            // ---- Phase F: cv2.resize(INTER_CUBIC) 400->48 (yaml_env.py:433-434), OpenCV's own path:
            // horizontal pass in int32 with 11-bit weights, vertical pass as an fp32 FMA chain with
            // weights * 2^-22, round-half-even, saturate; then float16(x)/255 via a host-built table.
            for (int q = tid; q < c.ns * c.img; q += VIEW_THREADS) {
                int rr = q / c.img, oc = q % c.img;
                const short* tp = d.cubic_tap + 4 * oc; const short* cf = d.cubic_coef + 4 * oc;
                const uint8_t* row = pix + rr * c.ns;
                hbuf[q] = row[tp[0]] * cf[0] + row[tp[1]] * cf[1] + row[tp[2]] * cf[2] + row[tp[3]] * cf[3];
            }
            __syncthreads();
            const float scale = 1.f / (2048.f * 2048.f);
            for (int q = tid; q < c.img * c.img; q += VIEW_THREADS) {
                int orow = q / c.img, oc = q % c.img;
                const short* tp = d.cubic_tap + 4 * orow; const short* cf = d.cubic_coef + 4 * orow;
                float b0 = cf[0] * scale, b1 = cf[1] * scale, b2 = cf[2] * scale, b3 = cf[3] * scale;
                float s0 = (float)hbuf[tp[0] * c.img + oc], s1 = (float)hbuf[tp[1] * c.img + oc];
                float s2 = (float)hbuf[tp[2] * c.img + oc], s3 = (float)hbuf[tp[3] * c.img + oc];
                float v = fmaf(s0, b0, fmaf(s1, b1, fmaf(s2, b2, s3 * b3)));
                int iv = __float2int_rn(v);
                iv = min(255, max(0, iv));
                d.o_sensor[(size_t)idx * c.img * c.img + q] = d.f16_lut[iv];
            }
        }
    }
    if (DEBUG_FULL) return;

    // ---- Phase G: state vector + pedestrian observation (img_env.cpp:547-587, yaml_env.py:392-481)
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
    float* pvs = d.o_pvs + (size_t)idx * c.pvs_len;
    for (int k = tid; k < c.pvs_len; k += VIEW_THREADS) pvs[k] = k == 0 ? (float)c.P : 0.f;
    for (int j = tid; j < c.P; j += VIEW_THREADS) {
        int pi = s * c.P + j;
        double bx, by, bvx, bvy;
        tf_apply(sh->world_base, PDF(d, PD_X, pi), PDF(d, PD_Y, pi), bx, by);
        tf_rotate(sh->world_base, PDF(d, PD_VX, pi), PDF(d, PD_VY, pi), bvx, bvy);
        float px = (float)bx, py = (float)by;
        pobs[4 * j] = px; pobs[4 * j + 1] = py; pobs[4 * j + 2] = (float)bvx; pobs[4 * j + 3] = (float)bvy;
        pkey[j] = (double)px * (double)px + (double)py * (double)py;
    }
    __syncthreads();
    for (int j = tid; j < c.P; j += VIEW_THREADS) {   // stable rank == python's list.sort(key=...)
        double kj = pkey[j];
        int rk = 0;
        for (int i = 0; i < c.P; i++) rk += (pkey[i] < kj) || (pkey[i] == kj && i < j);
        prank[j] = rk;
    }
    __syncthreads();
    for (int j = tid; j < c.P; j += VIEW_THREADS) {
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
    for (int q = tid; q < 3 * npm; q += VIEW_THREADS) {
        int ch = q / npm, cell = q % npm;
        int wv = winner[cell];
        float v = 0.f;
        if (wv >= 0) { int j = wv & 0xFFFF; v = ch == 0 ? 1.0f : pobs[4 * j + 1 + ch]; }
        pm[q] = v;
    }
}

inline size_t view_smem_bytes(const Cfg& c) {
    size_t b = (sizeof(ViewShared) + 15) & ~15;
    b += (size_t)c.vh * c.vwb * 4 * (c.use_laser ? 1 : 2);
    b += 3 * (((size_t)c.range_total * 2 + 15) & ~15);
    b += ((size_t)c.ns * c.ns + 15) & ~15;
    b += (size_t)c.ns * c.img * 4;
    b += (size_t)c.img * c.img * 4;
    b += (size_t)((c.P + 1) & ~1) * 8 + (size_t)c.P * 16 + (size_t)c.P * 4;
    return b + 64;
}
