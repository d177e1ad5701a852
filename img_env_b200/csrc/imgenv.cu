// libimgenv_b200.so — host side of the C ABI declared in include/imgenv.h.
// Replaces EnvService::{init_env,reset_env,step_env,end_ep} (img_env.cpp:716-755) and the Python
// _get_states post-processing (yaml_env.py:446-481).  Compiled for sm_100a only; there is no CPU
// fallback: every entry point fails loudly if CUDA is unavailable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <limits>
#include <string>
#include <vector>
#include "../../include/imgenv.h"
#include "state.cuh"
#include "kin.cuh"
#include "foot.cuh"
#include "view.cuh"
#include "dyn.cuh"
#include "rvotree.cuh"
#include "host_tables.h"
#include "sampler.h"
#include <random>

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct imgenv {
    Dev d;
    int device = 0;
    std::vector<void*> allocs;
    std::vector<double> robot_desc;   // [R][25] float32-widened
    std::vector<int> ped_shape;
    bool outputs_bound = false;
    int ped_yaw_mode = 0;
    unsigned solve_calls = 0;
    // Episode queue for device-side auto-reset (imgenv_autoreset_*): per scene a ring of pre-sampled ResetEnv records in device
    // memory.  The host sampler runs AHEAD of the simulation (same per-scene generator streams, hence the same episodes as
    // synchronous resets), so a reset selected by a device-side mask never waits for the host.
    struct AutoReset {
        bool on = false; int depth = 0; imgenv_sampler* smp = nullptr; int ignore_obstacle = 0;
        double* q_dbl = nullptr; int* q_int = nullptr; int* q_head = nullptr; int* q_tail = nullptr;   // device: [S][depth][dper], [S][depth][iper], [S], [S]
        int* ids_d = nullptr; int* n_dev = nullptr;
        std::vector<int> produced, seen;         // host: episodes uploaded per scene / consumed as last seen
        int* head_h = nullptr; cudaEvent_t ev_head = nullptr; bool head_pending = false; int head_age = 0;
        double* up_dbl = nullptr; int* up_int = nullptr; int up_cap = 0; cudaEvent_t ev_up = nullptr;   // pinned upload staging
    } ar;
    double* ar_up_dbl_d = nullptr; int* ar_up_int_d = nullptr;      // device landing area of a refill batch
    // reset staging (pinned host + device)
    // two sets used alternately: a reset only waits (on the set's event) for the reset before the previous one
    struct Stage { double* h = nullptr; double* d = nullptr; int* ih = nullptr; int* id = nullptr; float* fh = nullptr; float* fd = nullptr; cudaEvent_t ev = nullptr; };
    Stage stage[2]; int stage_i = 0;
    double* st_h = nullptr; double* st_d = nullptr; size_t st_doubles = 0;   // aliases of the set in use
    int* sti_h = nullptr; int* sti_d = nullptr; size_t st_ints = 0;
    float* stf_h = nullptr; float* stf_d = nullptr; size_t st_floats = 0;
    float* act_d = nullptr; uint8_t* alive_d = nullptr;
    size_t view_smem = 0, dyn_smem = 0, foot_smem = 0, obj_smem = 0;
    bool view_ctas5 = false;
    // Side stream, forked after the agents have moved and joined before the call returns to the caller's stream:
    // the pedestrian observation (k_ped_obs only reads poses) runs beside the stamp / view kernels.  SFM only: the
    // sequential, latency-bound quadtree update of step t follows it there and is joined before the next reader of
    // the tree (next step's forces, reset).
    cudaStream_t side = nullptr, side_tree = nullptr;      // pedestrian observation; SFM quadtree maintenance
    cudaEvent_t ev_moved = nullptr, ev_consts = nullptr, ev_ped = nullptr, ev_tree = nullptr; bool tree_pending = false;
    size_t ped_smem = 0;
    // optional per-kernel CUDA-event timing (bench.py roofline): 4 events per profiled step
    std::vector<cudaEvent_t> evs; int prof_max = 0, prof_n = 0;
};

template <class T> static int dalloc(imgenv* h, T** p, size_t n, int fill_byte = 0) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&q, bytes));
    CK(cudaMemset(q, fill_byte, bytes));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
template <class T> static int dupload(imgenv* h, const T** p, const std::vector<T>& v) {
    T* q = nullptr;
    if (dalloc(h, &q, v.size())) return -1;
    if (!v.empty()) CK(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *p = q;
    return 0;
}

// fn(i) for every listed scene; returns the first error message or nullptr
template <class F> static const char* for_scenes(int n, F fn) {
    for (int i = 0; i < n; i++) if (const char* e = fn(i)) return e;
    return nullptr;
}

extern "C" const char* imgenv_last_error(void) { return g_err.c_str(); }
extern "C" const char* imgenv_version(void) { return "img_env_b200 0.1 (sm_100a)"; }
// host-only helpers exported for CPU-side table checks (tests/test_tables.py)
extern "C" void imgenv_f16_lut(uint16_t* out256) { ht::f16_lut(out256); }
extern "C" int imgenv_cubic_tables(int src, int dst, short* need_idx, int* ns, short* tap, short* coef) {
    std::vector<short> n, t, c;
    ht::cubic_tables(src, dst, n, t, c);
    *ns = (int)n.size();
    memcpy(need_idx, n.data(), n.size() * 2); memcpy(tap, t.data(), t.size() * 2); memcpy(coef, c.data(), c.size() * 2);
    return 0;
}
extern "C" double imgenv_yaw_from_quaternion(double x, double y, double z, double w) { return ht::yaw_from_quaternion(x, y, z, w); }
extern "C" int imgenv_set_ped_yaw_mode(imgenv_t* h, int mode) { if (!h) return fail("null handle"); h->ped_yaw_mode = mode; return 0; }
// SpeedLimiter(msg) never assigns min_jerk (speed_limit.cpp:56-65): the node clamps with whatever its stack held.  The library
// defaults to min_jerk = max_jerk = msg.min_jerk; tests pin the value the reference build actually reads.  min_jerk[R][2] = lin, ang.
extern "C" int imgenv_debug_set_min_jerk(imgenv_t* h, const double* min_jerk) {
    if (!h || !min_jerk) return fail("imgenv_debug_set_min_jerk: null argument");
    const int R = h->d.c.R;
    std::vector<Limiter> lv(R), lw(R);
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(lv.data(), h->d.lim_v, sizeof(Limiter) * R, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lw.data(), h->d.lim_w, sizeof(Limiter) * R, cudaMemcpyDeviceToHost));
    for (int r = 0; r < R; r++) { lv[r].min_j = min_jerk[2 * r]; lw[r].min_j = min_jerk[2 * r + 1]; }
    CK(cudaMemcpy(const_cast<Limiter*>(h->d.lim_v), lv.data(), sizeof(Limiter) * R, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<Limiter*>(h->d.lim_w), lw.data(), sizeof(Limiter) * R, cudaMemcpyHostToDevice));
    return 0;
}

__global__ void k_init_state(Dev d) {
    size_t t0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t nh = (size_t)d.c.S * d.c.NP;
    for (size_t i = t0; i < nh; i += stride) d.foot_hdr[i] = make_int4(0, 0, 0, 0);      // no footprint record yet
    size_t nr = (size_t)d.c.S * d.c.R;
    for (size_t i = t0; i < nr; i += stride) {
        d.rb[(size_t)RB_PREVD * nr + i] = nan("");
        d.rb[(size_t)RB_MIND * nr + i] = INFINITY;
    }
    size_t na = (size_t)d.c.S * d.c.NA;
    if (d.c.scene_type == 1)
        for (size_t i = t0; i < na; i += stride) {
            double* rec = d.sfm + i * SFM_REC;
            rec[6] = d.sfm_vmax0[i % d.c.NA];   // Tagent(): vmax ~ N(1.2, 0.2) (ped_agent.cpp:43-44), setVmax for pedestrians (pedscene.h:66)
            rec[7] = -1; rec[8] = -1; rec[9] = 0;
            rec[10] = 1.0;     // every agent is inserted into the quadtree by addAgent (ped_scene.cpp:69-75)
        }
}

extern "C" int imgenv_destroy(imgenv_t* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d.rec_rb) cudaFree(h->d.rec_rb);
    if (h->d.rec_pd) cudaFree(h->d.rec_pd);
    if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
    if (h->side_tree) { cudaStreamSynchronize(h->side_tree); cudaStreamDestroy(h->side_tree); }
    if (h->ev_consts) cudaEventDestroy(h->ev_consts);
    if (h->ev_moved) cudaEventDestroy(h->ev_moved);
    if (h->ev_tree) cudaEventDestroy(h->ev_tree);
    if (h->ev_ped) cudaEventDestroy(h->ev_ped);
    if (h->ar.up_dbl) cudaFreeHost(h->ar.up_dbl);
    if (h->ar.up_int) cudaFreeHost(h->ar.up_int);
    if (h->ar.head_h) cudaFreeHost(h->ar.head_h);
    if (h->ar.ev_head) cudaEventDestroy(h->ar.ev_head);
    if (h->ar.ev_up) cudaEventDestroy(h->ar.ev_up);
    for (auto& g : h->stage) {
        if (g.h) cudaFreeHost(g.h);
        if (g.ih) cudaFreeHost(g.ih);
        if (g.fh) cudaFreeHost(g.fh);
        if (g.ev) cudaEventDestroy(g.ev);
    }
    delete h;
    return 0;
}

static int create_impl(imgenv* h, const imgenv_config* cfg, const uint8_t* grid, int32_t H, int32_t W, const double* robot_desc,
                       const double* ped_desc, const double* robot_size_last) {
    Cfg& c = h->d.c;
    using ht::f32;
    c.S = cfg->num_scenes; c.R = cfg->num_robots; c.P = cfg->num_peds;
    c.scene_type = c.P > 0 ? cfg->scene_type : 0;
    c.relation = cfg->relation_ped_robo;
    c.NA = (c.scene_type != 0 && c.scene_type != 4) ? c.P + (c.relation == 1 ? c.R : 0) : 0;
    dyn_pick_block(c);
    c.H = H; c.W = W; c.Wb = (W + 31) / 32; c.Hc = (H + 31) / 32;
    c.res = f32(cfg->view_resolution); c.inv_res = 1.0 / c.res;
    double vwid = f32(cfg->view_width), vhei = f32(cfg->view_height);
    c.vw = (int)(vwid / c.res); c.vh = (int)(vhei / c.res);       // agent.cpp:82-83
    c.vwb = (c.vw + 31) / 32;
    c.step_hz = f32(cfg->step_hz); c.control_hz = 0.05;            // agent.cpp:89
    c.state_dim = cfg->state_dim; c.use_laser = cfg->use_laser != 0; c.range_total = cfg->range_total;
    c.ktype = cfg->robot_ktype;
    c.beep_r = f32(cfg->beep_r); c.ped_ca_p = f32(cfg->ped_ca_p);
    c.view_max_dist = f32(cfg->view_max_dist);
    c.view_base = tf_from_pose(vhei / 2, vwid / 2, 3.14159);      // agent.cpp:84-87
    c.base_view = tf_inv(c.view_base);
    c.cull_reach = hypot(vhei / 2, vwid / 2) + 4 * c.res;          // view pixels lie within [-h/2, h/2] x [-w/2, w/2] of the base frame
    c.img = cfg->image_size;
    if (c.img < 1 || c.img > 255) return fail("imgenv_create: image_size must be in [1, 255]");
    c.img_inv = (unsigned)(((1ull << 32) + c.img - 1) / c.img);
    c.hb_shift = 0; while ((c.range_total >> c.hb_shift) >= 64) c.hb_shift++;
    c.max_ped = cfg->max_ped; c.ped_vec_dim = cfg->ped_vec_dim;
    c.pvs_len = 1 + c.max_ped * c.ped_vec_dim;
    c.ped_image_r = cfg->ped_image_r; c.ped_res = 6.0 / cfg->ped_image_size; c.laser_max = cfg->laser_max;
    c.laser_norm = cfg->laser_norm;
    c.max_obs = std::max(cfg->max_obstacles, 1); c.max_traj = std::max(cfg->max_traj, 1);
    c.seed = cfg->seed;
    const double max_obj_r = cfg->max_object_radius > 0 ? cfg->max_object_radius : 0.75;
    c.obj_rad = object_rad_cells(max_obj_r, c.res);
    c.obj_cap = stamp_bitmap_words(c.obj_rad);
    if (c.vh != c.vw) return fail("imgenv_create: only square view maps are supported");
    if (c.vh > 1022) return fail("imgenv_create: view raster larger than 1022x1022 cells is not supported");
    if (cfg->ped_image_size != cfg->image_size) return fail("imgenv_create: ped_image_size must equal image_size");
    if (c.ped_vec_dim != 7) return fail("imgenv_create: ped_vec_dim must be 7 (yaml_env.py:399-408)");
    if (c.state_dim < 3 || c.state_dim > 5) return fail("imgenv_create: state_dim must be 3, 4 or 5");
    if (c.R < 1 || c.R > 4096 || c.P < 0 || c.P > 4096 || c.S < 1) return fail("imgenv_create: bad S/R/P");
    if (c.range_total > 4000 || c.range_total < 1) return fail("imgenv_create: range_total out of range");   // 12-bit ray ids, 0xFFF = none
    {   // world blocks (32x32 cells) the field of view can span at worst (its diagonal, any rotation): the world->view
        // rasterisation keeps their list in shared memory; larger views use the forward rasteriser
        const int span = (int)ceil(hypot((double)c.vh, (double)c.vw)) / 32 + 2;
        c.inverse_ok = c.use_laser && span * span <= INV_MAX_BLOCKS;
    }
    if (c.scene_type < 0 || c.scene_type > 4) return fail("imgenv_create: unknown scene type");

    // ---- static tables ----
    std::vector<short> need_idx, tap, coef;
    ht::cubic_tables(c.vw, c.img, need_idx, tap, coef);
    c.ns = (int)need_idx.size();
    h->robot_desc.assign(robot_desc, robot_desc + 25 * (size_t)c.R);
    std::vector<int> type_of(c.R, -1);
    std::vector<ht::TypeTables> types;
    std::vector<Limiter> lv(c.R), lw(c.R);
    for (int r = 0; r < c.R; r++) {
        double* dsc = h->robot_desc.data() + 25 * (size_t)r;
        for (int k = 1; k < 7; k++) dsc[k] = f32(dsc[k]);
        for (int L = 0; L < 2; L++) {
            const double* q = dsc + 7 + 9 * L;
            Limiter m;
            m.has_v = q[0] != 0; m.has_a = q[1] != 0; m.has_j = q[2] != 0;
            m.min_v = f32(q[3]); m.max_v = f32(q[4]); m.min_a = f32(q[5]); m.max_a = f32(q[6]);
            // SpeedLimiter(msg) copies min_jerk into max_jerk and leaves min_jerk uninitialised
            // (speed_limit.cpp:56-65); we take min_jerk for both.
            m.min_j = f32(q[7]); m.max_j = f32(q[7]);
            (L == 0 ? lv : lw)[r] = m;
        }
        for (int t = 0; t < (int)types.size() && type_of[r] < 0; t++) {
            int r0 = -1;
            for (int q = 0; q < r; q++) if (type_of[q] == t) { r0 = q; break; }
            const double* a = h->robot_desc.data() + 25 * (size_t)r0;
            bool same = robot_size_last[r0] == robot_size_last[r];
            for (int k = 0; k < 7; k++) same = same && a[k] == dsc[k];
            if (same) type_of[r] = t;
        }
        if (type_of[r] < 0) {
            types.emplace_back();
            std::string e = ht::build_type(c, dsc, f32(cfg->view_angle_begin), f32(cfg->view_angle_end), f32(cfg->view_min_dist),
                                           f32(cfg->view_max_dist), types.back());
            if (!e.empty()) return fail("imgenv_create: " + e);
            types.back().t.size_last = robot_size_last[r];
            type_of[r] = (int)types.size() - 1;
        }
    }
    c.n_types = (int)types.size();
    std::vector<double> lattice; std::vector<short> ray_end, spans; std::vector<unsigned short> khi, klo; std::vector<uint32_t> own_mask, tile_fov, edge_px, edge_tiles, dtab, ostat;
    std::vector<uint16_t> lut(256); ht::f16_lut(lut.data());
    std::vector<RobotType> rts;
    for (auto& T : types) {
        T.t.pts_off = (int)lattice.size() / 2; lattice.insert(lattice.end(), T.lattice.begin(), T.lattice.end());
        T.t.ring_off = (int)lattice.size() / 2; lattice.insert(lattice.end(), T.ring.begin(), T.ring.end());
        // (ray ends, spans and need_idx are bulk-copied into shared memory in 16-byte units: every segment starts on one and is padded to one)
        T.t.ray_off = (int)ray_end.size() / 2; ray_end.insert(ray_end.end(), T.ray_end.begin(), T.ray_end.end());
        while (ray_end.size() % 8) ray_end.push_back(0);
        T.t.span_off = (int)spans.size(); spans.insert(spans.end(), T.spans.begin(), T.spans.end());
        while (spans.size() % 8) spans.push_back(-1);
        T.t.khi_off = (int)khi.size(); khi.insert(khi.end(), T.khi.begin(), T.khi.end()); klo.insert(klo.end(), T.klo.begin(), T.klo.end());
        T.t.own_mask_off = (int)own_mask.size(); own_mask.insert(own_mask.end(), T.own_mask.begin(), T.own_mask.end());
        T.t.tile_off = (int)tile_fov.size(); tile_fov.insert(tile_fov.end(), T.tile_fov.begin(), T.tile_fov.end());
        T.t.edge_off = (int)edge_px.size(); T.t.n_edge = (int)T.edge_px.size(); edge_px.insert(edge_px.end(), T.edge_px.begin(), T.edge_px.end());
        T.t.etile_off = (int)edge_tiles.size(); edge_tiles.insert(edge_tiles.end(), T.edge_tiles.begin(), T.edge_tiles.end());
        {   // per (needed row, output column, tap) table for the laser_map reconstruction fused into the horizontal
            // cubic pass (view.cuh phase D/F): one 16-byte load gives a thread the four source pixels of its output.
            T.dtab.assign((size_t)c.ns * c.img * 4, 0u);
            for (int rr = 0; rr < c.ns; rr++)
                for (int oc = 0; oc < c.img; oc++)
                    for (int k = 0; k < 4; k++) {
                        const int pr = need_idx[rr], pc = need_idx[tap[4 * oc + k]];
                        const size_t full = (size_t)pr * c.vw + pc;
                        uint32_t kh = T.khi[full], e;   // (klo carries the "exact interval" flag in bit 15; khi does not)
                        if (kh == 0xFFFF) e = 0xFFFu;
                        else {
                            const int w0 = abs((int)T.ray_end[2 * kh] - T.t.org_x), h0 = abs((int)T.ray_end[2 * kh + 1] - T.t.org_y);
                            const uint32_t itop = (uint32_t)(w0 > h0 ? abs(pr - T.t.org_x) : abs(pc - T.t.org_y));
                            e = kh | (itop << 12);
                        }
                        if ((T.own_mask[full >> 5] >> (full & 31)) & 1u) e |= 1u << 31;
                        T.dtab[((size_t)rr * c.img + oc) * 4 + k] = e;
                    }
            // Per OUTPUT pixel of the resize (its <= 16 source pixels belong to it alone: the scale 8.33 exceeds the 4 taps): the
            // interval of the source pixels' top rays and the float16 value the pixel takes when none of those rays hits
            // anything (free / own footprint / outside every ray), evaluated with the arithmetic of view.cuh phase F.
            const int npx = c.img * c.img;
            const size_t half = (size_t)npx + (npx + 1) / 2;
            const int n_seg = (npx + 7) / 8;
            T.ostat.assign(2 * half + n_seg, 0u);
            uint16_t* oval = reinterpret_cast<uint16_t*>(T.ostat.data() + npx);
            // second half: the all-shadow test of an output (interval of ALL rays through its source pixels | floor(smallest
            // Chebyshev distance of such a pixel to the laser origin / 4) << 24) and its value when every one of those rays is
            // stopped in front of the pixels (each source pixel "unknown" = 200, own footprint 100)
            uint32_t* oshad = T.ostat.data() + half;
            uint16_t* oshval = reinterpret_cast<uint16_t*>(T.ostat.data() + half + npx);
            const float scale = 1.f / (2048.f * 2048.f);
            for (int orow = 0; orow < c.img; orow++)
                for (int oc = 0; oc < c.img; oc++) {
                    uint32_t kmin = 0xFFFu, kmax = 0, imax = 0; bool any = false; float sv[4], ss[4];
                    uint32_t amin = 0xFFFu, amax = 0, imin = 1023;
                    for (int t = 0; t < 4; t++) {
                        const int rr = tap[4 * orow + t];
                        int sum = 0, sum_sh = 0;
                        for (int k = 0; k < 4; k++) {
                            const int w = coef[4 * oc + k];
                            if (w == 0) continue;
                            const uint32_t e = T.dtab[((size_t)rr * c.img + oc) * 4 + k], kh = e & 0xFFFu;
                            int val = 200;                                   // no ray passes: unknown
                            if (kh != 0xFFFu) {
                                val = 255;
                                if (coef[4 * orow + t] != 0) {
                                    kmin = std::min(kmin, kh); kmax = std::max(kmax, kh); any = true;
                                    const int pr = need_idx[rr], pc = need_idx[tap[4 * oc + k]];
                                    const uint32_t cheb = (uint32_t)std::max(abs(pr - T.t.org_x), abs(pc - T.t.org_y));
                                    imax = std::max(imax, cheb); imin = std::min(imin, cheb);
                                    amax = std::max(amax, kh); amin = std::min(amin, (uint32_t)(T.klo[(size_t)pr * c.vw + pc] & 0x7FFFu));
                                }
                            }
                            if (e >> 31) val = 100;                          // own footprint
                            sum += val * w; sum_sh += ((e >> 31) ? 100 : 200) * w;
                        }
                        sv[t] = (float)sum; ss[t] = (float)sum_sh;
                    }
                    if (!any) { kmin = 1; kmax = 0; }
                    const float b0 = coef[4 * orow + 0] * scale, b1 = coef[4 * orow + 1] * scale, b2 = coef[4 * orow + 2] * scale, b3 = coef[4 * orow + 3] * scale;
                    const float v = fmaf(sv[0], b0, fmaf(sv[1], b1, fmaf(sv[2], b2, sv[3] * b3)));
                    int iv = (int)lrintf(v);
                    iv = std::min(255, std::max(0, iv));
                    T.ostat[(size_t)orow * c.img + oc] = kmin | (kmax << 12) | (((imax + 3) / 4) << 24);
                    oval[(size_t)orow * c.img + oc] = lut[iv];
                    if (!any) { amin = 1; amax = 0; }
                    const float vs = fmaf(ss[0], b0, fmaf(ss[1], b1, fmaf(ss[2], b2, ss[3] * b3)));
                    const int is = std::min(255, std::max(0, (int)lrintf(vs)));
                    oshad[(size_t)orow * c.img + oc] = amin | (amax << 12) | ((imin / 4) << 24);
                    oshval[(size_t)orow * c.img + oc] = lut[is];
                }
            // third part: the same interval | farthest pixel per SEGMENT of 8 consecutive outputs
            for (int sg = 0; sg < n_seg; sg++) {
                uint32_t kmin = 1, kmax = 0, far4 = 0; bool any = false;
                for (int q = 8 * sg; q < std::min(8 * sg + 8, npx); q++) {
                    const uint32_t kk = T.ostat[q], a = kk & 0xFFFu, b = (kk >> 12) & 0xFFFu;
                    if (b < a) continue;
                    if (!any) { kmin = a; kmax = b; any = true; } else { kmin = std::min(kmin, a); kmax = std::max(kmax, b); }
                    far4 = std::max(far4, kk >> 24);
                }
                T.ostat[2 * half + sg] = kmin | (kmax << 12) | (far4 << 24);
            }
        }
        T.t.ostat_off = (int)ostat.size(); ostat.insert(ostat.end(), T.ostat.begin(), T.ostat.end());
        while (ostat.size() % 4) ostat.push_back(0u);      // (bulk-copied in 16-byte units)
        T.t.dtab_off = (int)dtab.size(); dtab.insert(dtab.end(), T.dtab.begin(), T.dtab.end());
        rts.push_back(T.t);
    }
    // pedestrians
    std::vector<int> pshape(c.P), poff(2 * (size_t)c.P, 0), pn(2 * (size_t)c.P, 0);
    std::vector<int> proff(2 * (size_t)std::max(c.P, 1), 0), prn(2 * (size_t)std::max(c.P, 1), -1);
    std::vector<double> pdisc(6 * (size_t)std::max(c.P, 1), 0.0);
    std::vector<double> psize(6 * (size_t)c.P), pmax(c.P), prr(c.P);
    std::vector<float> prw(c.P);
    std::vector<double> pext(std::max(c.P, 1), 0.0), ppart(6 * (size_t)std::max(c.P, 1), 0.0);
    for (int p = 0; p < c.P; p++) {
        const double* q = ped_desc + 8 * (size_t)p;
        pshape[p] = (int)q[0];
        for (int k = 0; k < 6; k++) psize[6 * p + k] = f32(q[1 + k]);
        pmax[p] = f32(q[7]);
        prw[p] = (float)q[3];                                     // PedInfo.r_ = sizes_[2] (img_env.cpp:582)
        char buf[64]; snprintf(buf, sizeof buf, "%.2f", (double)prw[p]); prr[p] = atof(buf);   // python round(rt.r_, 2)
        std::vector<double> a, b;
        if (pshape[p] == 0) ht::lattice_circle(psize[6 * p], psize[6 * p + 1], psize[6 * p + 2], a);
        else if (pshape[p] == 2) { ht::lattice_circle(0, 0, psize[6 * p + 2], a); ht::lattice_circle(0, 0, psize[6 * p + 5], b); }
        {   // stamp extent: body offset + radius, or leg offset (configured or gait +-0.3, agent.cpp:696-735) + leg radius
            const double* z = psize.data() + 6 * p;
            pext[p] = pshape[p] == 0 ? hypot(z[0], z[1]) + z[2]
                                     : hypot(std::max(std::max(fabs(z[0]), fabs(z[3])), 0.3), std::max(fabs(z[1]), fabs(z[4]))) + std::max(z[2], z[5]);
            pext[p] += 3 * c.res;
        }
        ht::bounding_circle(a, c.res, ppart.data() + 6 * p); ht::bounding_circle(b, c.res, ppart.data() + 6 * p + 3);
        poff[2 * p] = (int)lattice.size() / 2; pn[2 * p] = (int)a.size() / 2; lattice.insert(lattice.end(), a.begin(), a.end());
        poff[2 * p + 1] = (int)lattice.size() / 2; pn[2 * p + 1] = (int)b.size() / 2; lattice.insert(lattice.end(), b.begin(), b.end());
        for (int leg = 0; leg < 2; leg++) {      // rim points of the circle lattices (foot.cuh: the interior is filled analytically)
            const double* z = psize.data() + 6 * p;
            double s0, s1, s2;
            if (pshape[p] == 0 && leg == 0) { s0 = z[0]; s1 = z[1]; s2 = z[2]; }
            else if (pshape[p] == 2) { s0 = 0; s1 = 0; s2 = leg ? z[5] : z[2]; }
            else continue;
            std::vector<double> ring;
            const double r_in = ht::lattice_circle_ring(s0, s1, s2, c.res, ring);
            if (r_in <= 0) continue;
            proff[2 * p + leg] = (int)lattice.size() / 2; prn[2 * p + leg] = (int)ring.size() / 2;
            lattice.insert(lattice.end(), ring.begin(), ring.end());
            pdisc[6 * p + 3 * leg] = s0; pdisc[6 * p + 3 * leg + 1] = s1; pdisc[6 * p + 3 * leg + 2] = r_in;
        }
    }
    h->ped_shape = pshape;
    std::vector<uint8_t> g(grid, grid + (size_t)H * W);
    std::vector<uint32_t> socc((size_t)H * c.Wb, 0u);
    for (int i = 0; i < H; i++) for (int j = 0; j < W; j++) if (grid[(size_t)i * W + j] < 250) socc[(size_t)i * c.Wb + (j >> 5)] |= 1u << (j & 31);

    std::vector<uint32_t> scand, scrow, sorow;
    ht::static_planes(socc, H, c.Wb, scand, scrow, sorow);
    // footprint records: capacity of every part's bitmap and its offset inside a scene's slice (foot.cuh)
    c.NPA = c.R + 2 * c.P; c.NP = c.NPA + c.max_obs;
    if (c.NP > 32767) return fail("imgenv_create: more than 32767 footprint parts per scene (R + 2 P + max_obstacles)");
    std::vector<int> part_off(c.NP + 1, 0);
    c.ag_cap = 1;
    for (int q = 0; q < c.NP; q++) {
        int cap;
        if (q < c.R) cap = stamp_bitmap_words(rts[type_of[q]].stamp_rad);
        else if (q < c.NPA) cap = stamp_bitmap_words((int)ppart[3 * (q - c.R) + 2]);
        else cap = c.obj_cap;
        if (q < c.NPA) c.ag_cap = std::max(c.ag_cap, cap);
        part_off[q + 1] = part_off[q] + 2 * cap;
    }
    c.scene_words = part_off[c.NP];
    Dev& d = h->d;
    while (need_idx.size() % 8) need_idx.push_back(0);
    while (tap.size() % 8) tap.push_back(0);
    while (coef.size() % 8) coef.push_back(0);
#define UP(field, vec) if (dupload(h, &d.field, vec)) return -1;
    std::vector<uint32_t> kpack(khi.size());
    for (size_t k = 0; k < khi.size(); k++) kpack[k] = (uint32_t)khi[k] | ((uint32_t)klo[k] << 16);
    UP(kpack, kpack) UP(grid, g) UP(static_occ, socc) UP(types, rts) UP(type_of, type_of) UP(lattice_xy, lattice) UP(ray_end, ray_end)
    UP(fov_spans, spans) UP(own_mask, own_mask) UP(need_idx, need_idx) UP(cubic_tap, tap)
    UP(cubic_coef, coef) UP(f16_lut, lut) UP(lim_v, lv) UP(lim_w, lw) UP(ped_shape, pshape) UP(ped_size, psize)
    UP(tile_fov, tile_fov) UP(edge_px, edge_px) UP(edge_tiles, edge_tiles) UP(dtab, dtab) UP(ostat, ostat) UP(static_cand, scand) UP(static_crow, scrow)
    UP(static_orow, sorow) UP(part_off, part_off) UP(ped_maxspeed, pmax) UP(ped_r_round, prr) UP(ped_pts_off, poff)
    UP(ped_pts_n, pn) UP(ped_ext, pext) UP(ped_part, ppart) UP(ped_ring_off, proff) UP(ped_ring_n, prn) UP(ped_disc, pdisc)
#undef UP
    size_t S = c.S;
    d.max_verts = 16 * c.max_obs + 16;
#define AL(field, n) if (dalloc(h, &d.field, (size_t)(n))) return -1;
    AL(foot_hdr, S * c.NP) AL(foot_words, S * (size_t)c.scene_words) AL(vconst, S * c.R)
    AL(rb, (size_t)RB_FIELDS * S * c.R) AL(pd, (size_t)PD_FIELDS * S * c.P)
    AL(traj, S * c.P * c.max_traj * 3) AL(traj_v, c.scene_type == 4 ? S * c.P * c.max_traj * 3 : 1) AL(traj_len, S * c.P) AL(obs, S * c.max_obs * 8) AL(n_obs, S) AL(step_no, S)
    AL(rvo_pos, S * c.NA * 2) AL(rvo_vel, S * c.NA * 2) AL(rvo_nvel, S * c.NA * 2) AL(sfm_force, c.scene_type == 1 ? S * c.NA * 12 : 1)
    AL(rvo_verts, S * d.max_verts * 8) AL(rvo_nodes, S * d.max_verts * 4) AL(rvo_nodeseg, S * d.max_verts * 4) AL(counters, 32) AL(orca_cursor, 2)
    d.rvo_arena_len = (c.scene_type == 2 || c.scene_type == 3) ? 32 * d.max_verts : 1;
    AL(rvo_arena, S * (size_t)d.rvo_arena_len)
    AL(rvo_counts, S * 2) AL(sfm, S * c.NA * SFM_REC) AL(sfm_obs, S * c.max_obs * 4) AL(sfm_nobs, S)
    AL(sfm_wp, S * c.P * (1 + c.max_traj) * 3)
    {
        const bool sf = c.scene_type == 1;
        AL(qt_nodes, sf ? S : 1) AL(qt_box, sf ? S * QT_MAX_NODES * 4 : 1) AL(qt_child0, sf ? S * QT_MAX_NODES : 1) AL(qt_count, sf ? S * QT_MAX_NODES : 1)
        AL(qt_leaf, sf ? S * c.NA * QT_LEAVES : 1) AL(qt_hash, sf ? S * c.NA : 1) AL(sfm_newpos, sf ? S * c.NA * 4 : 1) AL(sfm_treepos, sf ? S * c.NA * 2 : 1)
    }
#undef AL
    {   // SFM start-up state of a fresh reference process (pedscene.h:58-83): per Tagent one N(1.2,0.2) draw from a
        // default-seeded std::default_random_engine (ped_agent.cpp:19,43-44), pedestrians placed at
        // rand()/2147483647*10 (glibc rand, seed 1) and inserted into the quadtree, robots inserted at (0,0).
        std::vector<double> vmax0(std::max(c.NA, 1), 1.2);
        if (c.scene_type == 1) {
            std::default_random_engine generator;
            std::vector<double> pos(2 * (size_t)c.NA, 0.0);
            struct random_data rd; char statebuf[128]; memset(&rd, 0, sizeof rd);
            initstate_r(1, statebuf, sizeof statebuf, &rd);
            for (int a = 0; a < c.P + c.R; a++) {
                std::normal_distribution<double> distribution(1.2, 0.2);
                double v = distribution(generator);
                if (a < c.NA) vmax0[a] = a < c.P ? pmax[a] : v;
                if (a < c.P) { int32_t r1, r2; random_r(&rd, &r1); random_r(&rd, &r2); pos[2 * a] = r1 / 2147483647.0 * 10.0; pos[2 * a + 1] = r2 / 2147483647.0 * 10.0; }
            }
            std::vector<int> n_nodes(1), child0(QT_MAX_NODES, -1), count(QT_MAX_NODES, 0), leaf((size_t)c.NA * QT_LEAVES, -1), hash(c.NA, -1);
            std::vector<double> box((size_t)QT_MAX_NODES * 4, 0.0);
            QTreeView t; t.n_nodes = n_nodes.data(); t.box = box.data(); t.child0 = child0.data(); t.count = count.data();
            t.leaf = leaf.data(); t.hash = hash.data(); t.pos = pos.data(); t.na = c.NA;
            qt_init(t);
            for (int a = 0; a < c.NA; a++) qt_add(t, a, 0);      // addPed ... addRobot order == agent index order
            for (size_t sc = 0; sc < S; sc++) {
                CK(cudaMemcpy(d.qt_nodes + sc, n_nodes.data(), 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(d.qt_box + sc * QT_MAX_NODES * 4, box.data(), box.size() * 8, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(d.qt_child0 + sc * QT_MAX_NODES, child0.data(), child0.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(d.qt_count + sc * QT_MAX_NODES, count.data(), count.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(d.qt_leaf + sc * c.NA * QT_LEAVES, leaf.data(), leaf.size() * 4, cudaMemcpyHostToDevice));
                CK(cudaMemcpy(d.qt_hash + sc * c.NA, hash.data(), hash.size() * 4, cudaMemcpyHostToDevice));
            }
        }
        if (dupload(h, &d.sfm_vmax0, vmax0)) return -1;
    }
    d.orca_nslabs = (c.scene_type == 2 || c.scene_type == 3) ? 16384 : 0;
    if (dalloc(h, &d.orca_pool, (size_t)std::max(d.orca_nslabs, 1) * ORCA_SLAB_BYTES)) return -1;
    if (dalloc(h, &h->act_d, S * c.R * 3)) return -1;
    if (dalloc(h, &h->alive_d, S * c.R)) return -1;
    {   // rvo root = -1 (no obstacle tree) until the first reset
        std::vector<int> rc(2 * S, 0);
        for (size_t s = 0; s < S; s++) rc[2 * s + 1] = -1;
        CK(cudaMemcpy(d.rvo_counts, rc.data(), rc.size() * 4, cudaMemcpyHostToDevice));
    }
    // reset staging: per scene doubles = obs 8*max_obs + robots 5*R + peds 5*P + traj 3*max_traj*P + sfm segs 4*max_obs
    h->st_doubles = S * ((size_t)8 * c.max_obs + 5 * c.R + 5 * c.P + 6 * (size_t)c.max_traj * c.P + 4 * c.max_obs) + 8;
    h->st_ints = S * ((size_t)5 + c.P) + S + 8;
    h->st_floats = 8;
    for (auto& g : h->stage) {
        CK(cudaMallocHost((void**)&g.h, h->st_doubles * 8)); if (dalloc(h, &g.d, h->st_doubles)) return -1;
        CK(cudaMallocHost((void**)&g.ih, h->st_ints * 4)); if (dalloc(h, &g.id, h->st_ints)) return -1;
        CK(cudaMallocHost((void**)&g.fh, h->st_floats * 4)); if (dalloc(h, &g.fd, h->st_floats)) return -1;
        CK(cudaEventCreateWithFlags(&g.ev, cudaEventDisableTiming));
    }

    CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));      // (stream priorities make no difference here: measured)
    CK(cudaStreamCreateWithFlags(&h->side_tree, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_moved, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_consts, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_ped, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_tree, cudaEventDisableTiming));
    h->ped_smem = ped_smem_bytes(c); h->d.pl = ped_layout(c);
    if (h->ped_smem > 200 * 1024) return fail("imgenv_create: too many pedestrians for the pedestrian observation kernel's shared memory");
    CK(cudaFuncSetAttribute(k_ped_obs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ped_smem));
    h->foot_smem = (size_t)FOOT_WARPS * c.ag_cap * 4; h->obj_smem = (size_t)c.obj_cap * 4;
    if (h->foot_smem > 200 * 1024 || h->obj_smem > 200 * 1024) return fail("imgenv_create: an agent / object footprint is too large for the footprint kernels");
    CK(cudaFuncSetAttribute(k_footprints, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->foot_smem));
    CK(cudaFuncSetAttribute(k_object_footprints, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->obj_smem));
    h->view_smem = view_smem_bytes(c); h->d.vl = view_layout(c);
    h->dyn_smem = dyn_smem_bytes(d);
    if (h->view_smem > 227 * 1024) return fail("imgenv_create: view kernel needs too much shared memory for this configuration");
    CK(cudaFuncSetAttribute(k_view<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->view_smem));
    CK(cudaFuncSetAttribute(k_view<false, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->view_smem));
    h->view_ctas5 = c.NP <= 256 && !getenv("IMGENV_VIEW_CTAS4");      // (view.cuh: occupancy variant, measured per workload)
    CK(cudaFuncSetAttribute(k_view<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->view_smem));
    CK(cudaFuncSetAttribute(k_view<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->view_smem));
    CK(cudaFuncSetAttribute(k_view<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->view_smem));
    if (h->dyn_smem > 200 * 1024) return fail("imgenv_create: too many solver agents per scene for the dynamics kernel's shared memory");
    if (h->dyn_smem > 48 * 1024) CK(cudaFuncSetAttribute(k_dyn_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dyn_smem));
    k_init_state<<<1184, 256>>>(d);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int imgenv_create(const imgenv_config* cfg, const uint8_t* grid, int32_t H, int32_t W, const double* robot_desc,
                             const double* ped_desc, const double* robot_size_last, int32_t device, imgenv_t** out) {
    if (!cfg || !grid || !out || !robot_desc || !robot_size_last) return fail("imgenv_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("imgenv_create: no CUDA device (this library has no CPU path)");
    CK(cudaSetDevice(device));
    imgenv* h = new imgenv();
    memset(&h->d, 0, sizeof(Dev));
    h->device = device;
    if (create_impl(h, cfg, grid, H, W, robot_desc, ped_desc, robot_size_last)) { std::string e = g_err; imgenv_destroy(h); g_err = e; return -1; }
    *out = h;
    return 0;
}

extern "C" int imgenv_bind_outputs(imgenv_t* h, const imgenv_outputs* o) {
    if (!h || !o) return fail("imgenv_bind_outputs: null argument");
    if (!o->vector_states || !o->sensor_maps || !o->is_collisions || !o->is_arrives || !o->lasers || !o->ped_vector_states ||
        !o->ped_maps || !o->step_ds || !o->ped_min_dists) return fail("imgenv_bind_outputs: every output pointer is required");
    Dev& d = h->d;
    d.o_vec = o->vector_states; d.o_sensor = o->sensor_maps; d.o_coll = o->is_collisions; d.o_arr = o->is_arrives;
    d.o_laser = o->lasers; d.o_pvs = o->ped_vector_states; d.o_pmap = o->ped_maps; d.o_stepd = o->step_ds; d.o_mind = o->ped_min_dists;
    h->outputs_bound = true;
    return 0;
}

// staged reset record layout (per listed scene), doubles:
//   obs[max_obs][8] | robots[R][5] x,y,yaw,gx,gy | peds[P][5] x,y,yaw,gx,gy | traj[P][max_traj][3] | segs[max_obs][4] | traj_v[P][max_traj][3]
// ints: scene_id, n_obs, n_segs, 0, 0 | traj_len[P]      (the RVO obstacle ring + BSP are built on the device: rvotree.cuh)
// Records come either from the staged list of an imgenv_reset call (queue_depth == 0: record sl of st / sti) or from the
// episode queues (queue_depth > 0: the listed scene's next queued record; an empty queue re-plays the newest record and is
// counted in counters[2]).
__global__ void k_apply_reset(Dev d, int n, const double* st, const int* sti, size_t dper, size_t iper,
                              int queue_depth, const int* ids, int* q_head, const int* q_tail) {
    const Cfg& c = d.c;
    int sl = blockIdx.x;
    if (sl >= n || (d.n_dev && sl >= *d.n_dev)) return;
    const double* D = st + dper * sl; const int* I = sti + iper * sl;
    if (queue_depth > 0) {
        const int sc = ids[sl];
        int e = q_head[sc];
        if (e >= q_tail[sc]) { e = q_tail[sc] - 1; if (threadIdx.x == 0) atomicAdd(d.counters + 2, 1ull); }
        const size_t slot = (size_t)sc * queue_depth + (size_t)(e % queue_depth);
        D = st + dper * slot; I = sti + iper * slot;
        __syncthreads();
        if (threadIdx.x == 0) q_head[sc] = e + 1;
    }
    int s = I[0];
    const double* obs = D; const double* rob = obs + 8 * (size_t)c.max_obs; const double* ped = rob + 5 * (size_t)c.R;
    const double* traj = ped + 5 * (size_t)c.P; const double* segs = traj + 3 * (size_t)c.max_traj * c.P;
    const double* trajv = segs + 4 * (size_t)c.max_obs;
    const int* tl = I + 5;
    for (int k = threadIdx.x; k < 8 * c.max_obs; k += blockDim.x) d.obs[(size_t)s * c.max_obs * 8 + k] = obs[k];
    for (int k = threadIdx.x; k < 4 * c.max_obs; k += blockDim.x) d.sfm_obs[(size_t)s * c.max_obs * 4 + k] = segs[k];
    if (threadIdx.x == 0) {
        d.n_obs[s] = I[1]; d.sfm_nobs[s] = I[2]; d.step_no[s] = 0;
    }
    for (int j = threadIdx.x; j < c.R; j += blockDim.x) {
        int idx = s * c.R + j;
        const double* q = rob + 5 * j;
        // init_pose (agent.cpp:133-142), set_goal (:144-154), flag reset (img_env.cpp:269-273)
        RBF(d, RB_X, idx) = q[0]; RBF(d, RB_Y, idx) = q[1]; RBF(d, RB_YAW, idx) = q[2];
        RBF(d, RB_L0V, idx) = 0; RBF(d, RB_L0W, idx) = 0;
        RBF(d, RB_GX, idx) = q[3]; RBF(d, RB_GY, idx) = q[4]; RBF(d, RB_GYAW, idx) = q[2];
        RBF(d, RB_COLL, idx) = 0; RBF(d, RB_ARR, idx) = 0; RBF(d, RB_DONE, idx) = 0;
        RBF(d, RB_PREVD, idx) = nan("");                                    // yaml_env.py:225
        if (c.relation == 1 && c.NA > 0) {                                  // setRobotPos(..., 0.0, 0.0) img_env.cpp:278-281
            int a = c.P + j;
            if (c.scene_type == 2 || c.scene_type == 3) {
                d.rvo_pos[((size_t)s * c.NA + a) * 2] = (float)q[0]; d.rvo_pos[((size_t)s * c.NA + a) * 2 + 1] = (float)q[1];
                d.rvo_vel[((size_t)s * c.NA + a) * 2] = 0.f; d.rvo_vel[((size_t)s * c.NA + a) * 2 + 1] = 0.f;
            } else if (c.scene_type == 1) {
                double* rec = d.sfm + ((size_t)s * c.NA + a) * SFM_REC;
                rec[0] = q[0]; rec[1] = q[1]; rec[2] = 1.0;
            }
        }
    }
    for (int p = threadIdx.x; p < c.P; p += blockDim.x) {
        int pi = s * c.P + p;
        const double* q = ped + 5 * p;
        PDF(d, PD_X, pi) = q[0]; PDF(d, PD_Y, pi) = q[1]; PDF(d, PD_YAW, pi) = q[2];
        PDF(d, PD_TIDX, pi) = 0;
        d.traj_len[pi] = tl[p];
        for (int k = 0; k < 3 * c.max_traj; k++) d.traj[(size_t)pi * c.max_traj * 3 + k] = traj[(size_t)p * c.max_traj * 3 + k];
        if (c.scene_type == 4) for (int k = 0; k < 3 * c.max_traj; k++) d.traj_v[(size_t)pi * c.max_traj * 3 + k] = trajv[(size_t)p * c.max_traj * 3 + k];
        if (c.scene_type == 2 || c.scene_type == 3) {                       // setPedPos (velocity is NOT reset by the node)
            d.rvo_pos[((size_t)s * c.NA + p) * 2] = (float)q[0]; d.rvo_pos[((size_t)s * c.NA + p) * 2 + 1] = (float)q[1];
        } else if (c.scene_type == 1) {
            double* rec = d.sfm + ((size_t)s * c.NA + p) * SFM_REC;
            rec[0] = q[0]; rec[1] = q[1]; rec[2] = 0.0;                     // setPosition(x, y, 0)
            // setWayPoint (pedscene.h:38-47): clearWaypoints, then goal (r=1) + trajectory (r = z);
            // addWaypoint leaves destination = waypoints.front() without popping it
            rec[7] = 0; rec[8] = -1; rec[9] = 0;
            double* wp = d.sfm_wp + (size_t)pi * (1 + c.max_traj) * 3;
            wp[0] = q[3]; wp[1] = q[4]; wp[2] = 1.0;
            for (int k = 0; k < tl[p]; k++) {
                const double* tk = traj + ((size_t)p * c.max_traj + k) * 3;
                wp[3 + 3 * k] = tk[0]; wp[4 + 3 * k] = tk[1]; wp[5 + 3 * k] = tk[2];
            }
        }
    }
}

static int launch_observe(imgenv* h, const int* d_scene_ids, int n_scenes, int is_reset, cudaStream_t st, cudaEvent_t* ev = nullptr, const int* n_dev = nullptr) {
    Dev d = h->d; const Cfg& c = d.c;
    d.n_dev = n_dev;
    // fork 1: the SFM quadtree maintenance (ped_scene.cpp:167-182 moveAgent) only needs the moved agents; a long sequential
    // kernel of S warps, started first on its own stream and joined at the next step / reset
    if (!is_reset && c.scene_type == 1) {
        CK(cudaEventRecord(h->ev_moved, st));
        CK(cudaStreamWaitEvent(h->side_tree, h->ev_moved, 0));
        k_sfm_tree<<<c.S, 32, 0, h->side_tree>>>(d);
        CK(cudaEventRecord(h->ev_tree, h->side_tree));
        h->tree_pending = true;
    }
    // per-robot pose constants + state vector: both observation kernels read them
    k_view_consts<<<(n_scenes * c.R + VC_THREADS - 1) / VC_THREADS, VC_THREADS, 0, st>>>(d, d_scene_ids, n_scenes, 0);
    // fork 2: pedestrian observation on the side stream, joined before the call returns control to the caller's stream
    CK(cudaEventRecord(h->ev_consts, st));
    CK(cudaStreamWaitEvent(h->side, h->ev_consts, 0));
    k_ped_obs<<<n_scenes * c.R, PED_THREADS, h->ped_smem, h->side>>>(d, d_scene_ids);
    CK(cudaEventRecord(h->ev_ped, h->side));
    k_footprints<<<(n_scenes * c.NPA + FOOT_WARPS - 1) / FOOT_WARPS, FOOT_WARPS * 32, h->foot_smem, st>>>(d, d_scene_ids, n_scenes, is_reset ? 0 : 1);
    if (ev) cudaEventRecord(ev[2], st);
    if (c.inverse_ok && h->view_ctas5) k_view<false, false, 5><<<n_scenes * c.R, VIEW_THREADS, h->view_smem, st>>>(d, d_scene_ids, is_reset);
    else if (c.inverse_ok) k_view<false, false><<<n_scenes * c.R, VIEW_THREADS, h->view_smem, st>>>(d, d_scene_ids, is_reset);
    else k_view<false, true><<<n_scenes * c.R, VIEW_THREADS, h->view_smem, st>>>(d, d_scene_ids, is_reset);
    if (ev) cudaEventRecord(ev[3], st);
    CK(cudaStreamWaitEvent(st, h->ev_ped, 0));      // join: every output of the call is ordered on the caller's stream
    if (h->tree_pending) {      // inside a CUDA graph capture every fork has to be joined before the capture ends
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) { CK(cudaStreamWaitEvent(st, h->ev_tree, 0)); h->tree_pending = false; }
    }
    CK(cudaGetLastError());
    return 0;
}

// One ResetEnv record in the staging layout k_apply_reset reads (see above): request arrays of ONE scene -> D[dper], I[iper].
static const char* pack_reset_record(imgenv* h, int s, int no, const double* obs, const double* robots, const double* peds, const int32_t* traj_len,
                                     const double* traj, const double* traj_v, int ignore_obstacle, double* D, int* I) {
    Dev& d = h->d; const Cfg& c = d.c;
    using ht::f32;
    const int sl = 0;

        if (s < 0 || s >= c.S) return "imgenv_reset: scene id out of range";
        double* o_obs = D; double* o_rob = o_obs + 8 * (size_t)c.max_obs; double* o_ped = o_rob + 5 * (size_t)c.R;
        double* o_traj = o_ped + 5 * (size_t)c.P; double* o_seg = o_traj + 3 * (size_t)c.max_traj * c.P;
        double* o_trajv = o_seg + 4 * (size_t)c.max_obs;
        if (no < 0 || no > c.max_obs) return "imgenv_reset: too many obstacles for max_obstacles";
        I[0] = s; I[1] = no;
        int nseg = 0;
        for (int k = 0; k < no; k++) {
            const double* q = obs + ((size_t)sl * c.max_obs + k) * 11;
            double* o = o_obs + 8 * k;
            int shape = (int)q[0];
            o[0] = shape; for (int m = 0; m < 4; m++) o[1 + m] = f32(q[1 + m]);
            o[5] = q[5]; o[6] = q[6]; o[7] = ht::yaw_from_quaternion(q[7], q[8], q[9], q[10]);   // img_env.cpp:180-185
            { double ccx, ccy, rmax; object_bounds(o, ccx, ccy, rmax);
              if (shape < 0 || shape > 1 || object_rad_cells(rmax, c.res) > c.obj_rad) return "imgenv_reset: reset object larger than max_object_radius (imgenv_config) or of unknown shape"; }
            // get_corners (agent.cpp:626-651) -> pedscene->addObs (img_env.cpp:188-192)
            Tf2 t = tf_from_pose(o[5], o[6], o[7]);
            double pax, pay, pbx, pby;
            if (shape == 0) { tf_apply(t, o[1] - o[3], o[2] - o[3], pax, pay); tf_apply(t, o[1] + o[3], o[2] + o[3], pbx, pby); }
            else { tf_apply(t, o[1], o[3], pax, pay); tf_apply(t, o[2], o[4], pbx, pby); }
            if (!ignore_obstacle) {
                double* sg = o_seg + 4 * nseg; sg[0] = pax; sg[1] = pay; sg[2] = pbx; sg[3] = pby; nseg++;   // pedscene.h:23-27
                // (the same two corners span the RVO polygon (ax,ay),(ax,by),(bx,by),(bx,ay), rvoscene.h:19-26: built by k_rvo_build)
            }
        }
        I[2] = nseg;
        for (int j = 0; j < c.R; j++) {
            const double* q = robots + ((size_t)sl * c.R + j) * 8;
            double* o = o_rob + 5 * j;
            o[0] = q[0]; o[1] = q[1]; o[2] = ht::yaw_from_quaternion(q[2], q[3], q[4], q[5]); o[3] = q[6]; o[4] = q[7];
        }
        for (int p = 0; p < c.P; p++) {
            const double* q = peds + ((size_t)sl * c.P + p) * 8;
            double* o = o_ped + 5 * p;
            o[0] = q[0]; o[1] = q[1]; o[2] = ht::yaw_from_quaternion(q[2], q[3], q[4], q[5]); o[3] = q[6]; o[4] = q[7];
            int tl = traj_len ? traj_len[(size_t)sl * c.P + p] : 0;
            if (tl < 1 || tl > c.max_traj) return "imgenv_reset: pedestrian trajectory length must be in [1, max_traj]";
            I[5 + p] = tl;
            for (int k = 0; k < 3 * tl; k++) o_traj[(size_t)p * c.max_traj * 3 + k] = traj[((size_t)sl * c.P + p) * c.max_traj * 3 + k];
            if (c.scene_type == 4) for (int k = 0; k < 3 * tl; k++) o_trajv[(size_t)p * c.max_traj * 3 + k] = traj_v[((size_t)sl * c.P + p) * c.max_traj * 3 + k];
        }
        return nullptr;
    }
static size_t reset_dper(const Cfg& c) { return (size_t)8 * c.max_obs + 5 * c.R + 5 * c.P + 6 * (size_t)c.max_traj * c.P + 4 * c.max_obs; }
static size_t reset_iper(const Cfg& c) { return (size_t)5 + c.P; }

extern "C" int imgenv_reset(imgenv_t* h, int32_t n, const int32_t* scene_ids, const int32_t* n_obs, const double* obs,
                            const double* robots, const double* peds, const int32_t* traj_len, const double* traj,
                            const double* traj_v, int32_t ignore_obstacle, void* stream) {
    if (!h) return fail("imgenv_reset: null handle");
    if (!h->outputs_bound) return fail("imgenv_reset: outputs not bound (imgenv_bind_outputs)");
    Dev& d = h->d; const Cfg& c = d.c;
    if (n < 1 || n > c.S) return fail("imgenv_reset: bad scene count");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (h->tree_pending) { CK(cudaStreamWaitEvent(st, h->ev_tree, 0)); h->tree_pending = false; }
    {   // staging is double-buffered: wait only until the reset that last used this set has consumed it
        imgenv::Stage& g = h->stage[h->stage_i ^= 1];
        CK(cudaEventSynchronize(g.ev));
        h->st_h = g.h; h->st_d = g.d; h->sti_h = g.ih; h->sti_d = g.id; h->stf_h = g.fh; h->stf_d = g.fd;
    }
    using ht::f32;
    if (c.scene_type == 4 && !traj_v) return fail("imgenv_reset: dataset replay needs traj_v");
    if (scene_ids) {   // two records for one scene would race in k_apply_reset / the object stamps
        std::vector<char> seen(c.S, 0);
        for (int sl = 0; sl < n; sl++) {
            const int s = scene_ids[sl];
            if (s < 0 || s >= c.S) return fail("imgenv_reset: scene id out of range");
            if (seen[s]) return fail("imgenv_reset: duplicate scene id");
            seen[s] = 1;
        }
    }
    const size_t dper = reset_dper(c), iper = reset_iper(c);
    memset(h->st_h, 0, dper * n * 8); memset(h->sti_h, 0, iper * n * 4);
    for (int sl = 0; sl < n; sl++) {
        const int sc = scene_ids ? scene_ids[sl] : sl;
        const char* e = pack_reset_record(h, sc, n_obs ? n_obs[sl] : 0, obs ? obs + (size_t)sl * c.max_obs * 11 : nullptr, robots + (size_t)sl * c.R * 8,
                                          peds ? peds + (size_t)sl * c.P * 8 : nullptr, traj_len ? traj_len + (size_t)sl * c.P : nullptr,
                                          traj ? traj + (size_t)sl * c.P * c.max_traj * 3 : nullptr, traj_v ? traj_v + (size_t)sl * c.P * c.max_traj * 3 : nullptr,
                                          ignore_obstacle, h->st_h + dper * sl, h->sti_h + iper * sl);
        if (e) return fail(e);
    }
    CK(cudaMemcpyAsync(h->st_d, h->st_h, dper * n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->sti_d, h->sti_h, iper * n * 4, cudaMemcpyHostToDevice, st));
    // the scene-id list lives at the head of each int record; build a compact list after the records
    int* ids_h = h->sti_h + iper * n;   // (st_ints has S*(...)+8 >= iper*n + n only if checked)
    if (iper * (size_t)n + n > h->st_ints) return fail("imgenv_reset: staging overflow");
    for (int sl = 0; sl < n; sl++) ids_h[sl] = h->sti_h[iper * sl];
    int* ids_d = h->sti_d + iper * n;
    CK(cudaMemcpyAsync(ids_d, ids_h, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_apply_reset<<<n, 128, 0, st>>>(d, n, h->st_d, h->sti_d, dper, iper, 0, nullptr, nullptr, nullptr);
    k_object_footprints<<<n * c.max_obs, OBJ_THREADS, h->obj_smem, st>>>(d, ids_d);     // obs.draw(obs_map_, 0, ...) img_env.cpp:187
    if (c.scene_type == 2 || c.scene_type == 3) k_rvo_build<<<n, 32, 0, st>>>(d, ids_d, ignore_obstacle);   // processObstacles (KdTree.cpp:119-128)
    if (launch_observe(h, ids_d, n, 1, st)) return -1;                 // view_agent(); get_states() img_env.cpp:285-286
    CK(cudaEventRecord(h->stage[h->stage_i].ev, st));                   // stream-ordered like imgenv_step: no host sync
    return 0;
}

// k_dyn_solve (if a solver runs), k_dyn_apply, k_view_consts, k_footprints, k_view, k_ped_obs (+ k_sfm_tree)
extern "C" int imgenv_launches_per_step(const imgenv_t* h) { return !h ? 5 : (h->d.c.scene_type == 1 ? 7 : (h->d.c.NA > 0 ? 6 : 5)); }

extern "C" int imgenv_step(imgenv_t* h, const float* d_actions, const uint8_t* d_alive, void* stream) {
    if (!h) return fail("imgenv_step: null handle");
    if (!h->outputs_bound) return fail("imgenv_step: outputs not bound (imgenv_bind_outputs)");
    if (!d_actions) return fail("imgenv_step: null actions");
    Dev& d = h->d; const Cfg& c = d.c;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t* ev = nullptr;
    if (h->prof_n < h->prof_max) { ev = h->evs.data() + 4 * (size_t)h->prof_n; h->prof_n++; }
    if (ev) cudaEventRecord(ev[0], st);
    {
        const int nblk = dyn_nblk(c);
        if (h->tree_pending) { CK(cudaStreamWaitEvent(st, h->ev_tree, 0)); h->tree_pending = false; }
        if (c.NA > 0) k_dyn_solve<<<c.S * nblk, c.dyn_threads, h->dyn_smem, st>>>(d, d_actions, d_alive, (int)(h->solve_calls++ & 1));
        k_dyn_apply<<<c.S * nblk, c.dyn_threads, 0, st>>>(d, d_actions, d_alive, h->ped_yaw_mode);

    }
    if (ev) cudaEventRecord(ev[1], st);
    return launch_observe(h, nullptr, c.S, 0, st, ev);
}

// Per-kernel device timing for bench.py: events are recorded on the launching stream around each of the
// four kernels of the next `max_steps` imgenv_step calls.
extern "C" int imgenv_profile_begin(imgenv_t* h, int max_steps) {
    if (!h) return fail("null handle");
    for (cudaEvent_t e : h->evs) cudaEventDestroy(e);
    h->evs.assign(4 * (size_t)max_steps, nullptr);
    for (auto& e : h->evs) CK(cudaEventCreate(&e));
    h->prof_max = max_steps; h->prof_n = 0;
    return 0;
}
// ms[3] = mean duration of the dynamics kernels, k_footprints, k_view (k_ped_obs runs beside them on the side stream); returns #steps
extern "C" int imgenv_profile_end(imgenv_t* h, float* ms) {
    if (!h) return fail("null handle");
    CK(cudaDeviceSynchronize());
    for (int k = 0; k < 3; k++) ms[k] = 0.f;
    for (int i = 0; i < h->prof_n; i++)
        for (int k = 0; k < 3; k++) { float t = 0; CK(cudaEventElapsedTime(&t, h->evs[4 * i + k], h->evs[4 * i + k + 1])); ms[k] += t; }
    int n = h->prof_n;
    for (int k = 0; k < 3 && n; k++) ms[k] /= n;
    for (cudaEvent_t e : h->evs) cudaEventDestroy(e);
    h->evs.clear(); h->prof_max = 0; h->prof_n = 0;
    return n;
}

// Clears is_collision_/is_arrive_ of every robot (what ImgEnv::_reset does at img_env.cpp:272-273) without
// moving anything, so a throughput run never benefits from Agent::view's early-out (agent.cpp:358-360).
__global__ void k_revive(Dev d) {
    size_t n = (size_t)d.c.S * d.c.R;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { d.rb[(size_t)RB_COLL * n + i] = 0; d.rb[(size_t)RB_ARR * n + i] = 0; d.rb[(size_t)RB_DONE * n + i] = 0; }
}
extern "C" int imgenv_revive(imgenv_t* h, void* stream) {
    if (!h) return fail("null handle");
    size_t n = (size_t)h->d.c.S * h->d.c.R;
    k_revive<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int imgenv_step_host(imgenv_t* h, const float* h_actions, const uint8_t* h_alive, void* stream) {
    if (!h || !h_actions) return fail("imgenv_step_host: null argument");
    const Cfg& c = h->d.c;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h->act_d, h_actions, (size_t)c.S * c.R * 3 * 4, cudaMemcpyHostToDevice, st));
    if (h_alive) CK(cudaMemcpyAsync(h->alive_d, h_alive, (size_t)c.S * c.R, cudaMemcpyHostToDevice, st));
    return imgenv_step(h, h->act_d, h_alive ? h->alive_d : nullptr, st);
}

// ---------------------------------------------------------------------------------------------------------
// Native episode sampler (EnvPos.reset, reset_helper.py:115-345) and the fused sample+reset entry point.
struct imgenv_sampler {
    sampler::Sampler s;
    std::vector<int32_t> n_obs, traj_len; std::vector<double> obs, robots, peds, traj;   // scratch of imgenv_reset_sampled
};
#define SAMPLER_HDR 8
#define SAMPLER_POSE_REC (3 + SAMPLER_MAX_MULTI * 6)
#define SAMPLER_AGENT_REC (1 + 2 * SAMPLER_POSE_REC)
#define SAMPLER_OBJ_REC 14
extern "C" int imgenv_sampler_create(const double* desc, int64_t n_desc, int32_t n_scenes, uint64_t seed, imgenv_sampler_t** out) {
    if (!desc || !out || n_desc < SAMPLER_HDR || n_scenes < 1) return fail("imgenv_sampler_create: bad arguments");
    imgenv_sampler* w = new imgenv_sampler();
    sampler::Sampler& S = w->s;
    S.R = (int)desc[0]; S.P = (int)desc[1]; const int nobj = (int)desc[2];
    S.circle_lo = desc[3]; S.circle_hi = desc[4]; S.target_min_dist = desc[5]; S.go_back = (int)desc[6];
    if (S.R < 0 || S.P < 0 || nobj < 0 || S.go_back < 0 || S.go_back > 2 ||
        n_desc != SAMPLER_HDR + (int64_t)(S.R + S.P) * SAMPLER_AGENT_REC + (int64_t)nobj * SAMPLER_OBJ_REC) {
        delete w;
        return fail("imgenv_sampler_create: descriptor size does not match its header");
    }
    const double* q = desc + SAMPLER_HDR;
    auto pose = [&](sampler::PoseSpec& ps) {
        ps.type = (int)q[0]; ps.n_multi = (int)q[1]; ps.len = (int)q[2];
        for (int m = 0; m < SAMPLER_MAX_MULTI; m++) for (int k = 0; k < 6; k++) ps.v[m][k] = q[3 + 6 * m + k];
        q += SAMPLER_POSE_REC;
        return ps.n_multi >= 0 && ps.n_multi <= SAMPLER_MAX_MULTI && (!(ps.type & sampler::T_MULTI) || ps.n_multi >= 1);
    };
    for (int i = 0; i < S.R + S.P; i++) {
        sampler::AgentSpec a; a.msize = *q++;
        if (!pose(a.begin) || !pose(a.target)) { delete w; return fail("imgenv_sampler_create: bad range_multi count"); }
        // configurations on which the reference's loop never terminates (reset_helper.py:222-305): a begin pose that is
        // not sampled ('range') next to a sampled target leaves reset_init True forever; 'view_plus' keeps a stale pose.
        const bool fixed_b = a.begin.type & (sampler::T_EQ_FIX | sampler::T_EQ_RAND_ANGLE), fixed_t = a.target.type & (sampler::T_EQ_FIX | sampler::T_EQ_RAND_ANGLE);
        if (!(fixed_b && fixed_t) && !(a.begin.type & sampler::T_RANGE)) {
            delete w;
            return fail("imgenv_sampler_create: begin_poses_type must be a 'range' type unless begin and target are both fixed "
                        "(the reference loops forever)");
        }
        if (!fixed_t && !(a.target.type & (sampler::T_RANGE | sampler::T_CIRCLE_FIX))) { delete w; return fail("imgenv_sampler_create: unknown target_poses_type"); }
        if (a.target.type & sampler::T_PLUS) { delete w; return fail("imgenv_sampler_create: 'view_plus' targets are not implemented by the reference"); }
        S.agents.push_back(a);
    }
    for (int i = 0; i < nobj; i++) {
        sampler::ObjectSpec o; o.shape = (int)q[0]; o.fixed = (int)q[1]; o.pose_len = (int)q[2];
        for (int k = 0; k < 6; k++) o.pose[k] = q[3 + k];
        for (int k = 0; k < 4; k++) o.size_range[k] = q[9 + k];
        q += SAMPLER_OBJ_REC;
        S.objects.push_back(o);
    }
    S.rng.resize(n_scenes);
    for (int i = 0; i < n_scenes; i++) S.rng[i].seed(seed + (uint64_t)i);
    *out = w;
    return 0;
}
extern "C" int imgenv_sampler_destroy(imgenv_sampler_t* w) { delete w; return 0; }
extern "C" int imgenv_sampler_seed(imgenv_sampler_t* w, int32_t scene, uint64_t seed) {
    if (!w || scene < 0 || scene >= (int)w->s.rng.size()) return fail("imgenv_sampler_seed: bad arguments");
    w->s.rng[scene].seed(seed);
    return 0;
}
extern "C" int imgenv_sampler_sample(imgenv_sampler_t* w, int32_t n, const int32_t* scene_ids, int32_t max_obs, int32_t max_traj,
                                     int32_t* n_obs, double* obs, double* robots, double* peds, int32_t* traj_len, double* traj) {
    if (!w || n < 1) return fail("imgenv_sampler_sample: bad arguments");
    const sampler::Sampler& S = w->s;
    if ((int)S.objects.size() > max_obs) return fail("imgenv_sampler_sample: more objects than max_obs");
    if (S.P > 0 && max_traj < 2) return fail("imgenv_sampler_sample: max_traj must be >= 2");
    std::vector<char> seen(S.rng.size(), 0);
    for (int i = 0; i < n; i++) {
        const int sc = scene_ids ? scene_ids[i] : i;
        if (sc < 0 || sc >= (int)S.rng.size()) return fail("imgenv_sampler_sample: scene id out of range");
        if (seen[sc]) return fail("imgenv_sampler_sample: duplicate scene id");
        seen[sc] = 1;
    }
    for_scenes(n, [&](int i) -> const char* {
        const int sc = scene_ids ? scene_ids[i] : i;
        n_obs[i] = (int)S.objects.size();
        sampler::sample_scene(S, w->s.rng[sc], obs + (size_t)i * max_obs * 11, robots + (size_t)i * S.R * 8, peds + (size_t)i * S.P * 8,
                              traj_len + (size_t)i * S.P, traj + (size_t)i * S.P * max_traj * 3, max_traj);
        return nullptr;
    });
    return 0;
}
// Debug: the raw double stream of one scene's generator (tests pin it against CPython's random module).
extern "C" int imgenv_sampler_draw(imgenv_sampler_t* w, int32_t scene, int32_t kind, double a, double b, double* out) {
    if (!w || scene < 0 || scene >= (int)w->s.rng.size()) return fail("imgenv_sampler_draw: bad arguments");
    sampler::PyRandom& r = w->s.rng[scene];
    *out = kind == 0 ? r.random() : kind == 1 ? r.uniform(a, b) : kind == 2 ? r.gauss(a, b) : (double)r.randint((int)a, (int)b);
    return 0;
}
extern "C" int imgenv_reset_sampled(imgenv_t* h, imgenv_sampler_t* w, int32_t n, const int32_t* scene_ids, int32_t ignore_obstacle, void* stream) {
    if (!h || !w) return fail("imgenv_reset_sampled: null handle");
    const Cfg& c = h->d.c;
    if (w->s.R != c.R || w->s.P != c.P) return fail("imgenv_reset_sampled: sampler and simulator disagree on the number of robots / pedestrians");
    if (c.scene_type == 4) return fail("imgenv_reset_sampled: dataset replay takes recorded trajectories (imgenv_reset)");
    if (n < 1 || n > c.S) return fail("imgenv_reset_sampled: bad scene count");
    const size_t mo = std::max(c.max_obs, 1), mt = std::max(c.max_traj, 1), P1 = std::max(c.P, 1);
    w->n_obs.resize(n); w->obs.assign((size_t)n * mo * 11, 0.0); w->robots.resize((size_t)n * c.R * 8); w->peds.resize((size_t)n * P1 * 8);
    w->traj_len.resize((size_t)n * P1); w->traj.assign((size_t)n * P1 * mt * 3, 0.0);
    if (imgenv_sampler_sample(w, n, scene_ids, c.max_obs, c.max_traj, w->n_obs.data(), w->obs.data(), w->robots.data(), w->peds.data(), w->traj_len.data(), w->traj.data())) return -1;
    return imgenv_reset(h, n, scene_ids, w->n_obs.data(), w->obs.data(), w->robots.data(), w->peds.data(), w->traj_len.data(), w->traj.data(), nullptr, ignore_obstacle, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Device-side auto-reset: episode queues + resets selected by a device mask.
// scene list of a device mask, in scene order; one warp
__global__ void __launch_bounds__(32) k_compact_mask(const uint8_t* mask, int S, int* ids, int* n_out) {
    int base = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + threadIdx.x;
        const bool on = s < S && mask[s] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) ids[base + __popc(m & ((1u << threadIdx.x) - 1u))] = s;
        base += __popc(m);
    }
    if (threadIdx.x == 0) *n_out = base;
}
// uploaded records -> their queue slots (I[3] = slot, I[4] = episodes of that scene produced so far)
__global__ void k_queue_scatter(int n, const double* up_dbl, const int* up_int, size_t dper, size_t iper, double* q_dbl, int* q_int, int* q_tail) {
    const int k = blockIdx.x;
    if (k >= n) return;
    const int* I = up_int + iper * k;
    const size_t slot = (size_t)I[3];
    for (size_t i = threadIdx.x; i < dper; i += blockDim.x) q_dbl[slot * dper + i] = up_dbl[dper * k + i];
    for (size_t i = threadIdx.x; i < iper; i += blockDim.x) q_int[slot * iper + i] = I[i];
    if (threadIdx.x == 0) atomicMax(q_tail + I[0], I[4]);
}
extern "C" int imgenv_autoreset_refill(imgenv_t* h, void* stream) {
    if (!h || !h->ar.on) return fail("imgenv_autoreset_refill: auto-reset is not enabled (imgenv_autoreset_enable)");
    imgenv::AutoReset& a = h->ar; const Cfg& c = h->d.c;
    cudaStream_t st = (cudaStream_t)stream;
    // Consumption counters travel back asynchronously.  A scene consumes at most one episode per call, so the queues cannot run
    // empty while the newest counters the host has seen are younger than `depth` calls: only when the host has run that far
    // ahead of the device does it wait for the copy (an RL loop whose policy reads the observations never gets there).
    if (a.head_pending) {
        a.head_age++;
        if (a.head_age >= a.depth - 2) CK(cudaEventSynchronize(a.ev_head));
        if (cudaEventQuery(a.ev_head) == cudaSuccess) {
            for (int s = 0; s < c.S; s++) a.seen[s] = a.head_h[s];
            a.head_pending = false;
        }
    }
    const size_t dper = reset_dper(c), iper = reset_iper(c);
    sampler::Sampler& S = a.smp->s;
    const size_t mo = std::max(c.max_obs, 1), mt = std::max(c.max_traj, 1), P1 = std::max(c.P, 1);
    std::vector<double>& obs = a.smp->obs; std::vector<double>& rob = a.smp->robots; std::vector<double>& ped = a.smp->peds; std::vector<double>& trj = a.smp->traj;
    std::vector<int32_t>& tl = a.smp->traj_len;
    obs.assign(mo * 11, 0.0); rob.resize((size_t)c.R * 8); ped.resize(P1 * 8); tl.resize(P1); trj.assign(P1 * mt * 3, 0.0);
    int s = 0;
    while (s < c.S) {
        // one upload batch: as many records as the pinned staging holds
        int k = 0;
        bool waited = false;
        for (; s < c.S; s++) {
            while (a.produced[s] - a.seen[s] < a.depth) {
                if (k == a.up_cap) break;
                if (!waited) { CK(cudaEventSynchronize(a.ev_up)); waited = true; }      // the previous batch has left the staging
                std::fill(obs.begin(), obs.end(), 0.0); std::fill(trj.begin(), trj.end(), 0.0);
                sampler::sample_scene(S, S.rng[s], obs.data(), rob.data(), ped.data(), tl.data(), trj.data(), c.max_traj);
                double* D = a.up_dbl + dper * k; int* I = a.up_int + iper * k;
                memset(D, 0, dper * 8); memset(I, 0, iper * 4);
                if (const char* e = pack_reset_record(h, s, (int)S.objects.size(), obs.data(), rob.data(), ped.data(), tl.data(), trj.data(), nullptr,
                                                      a.ignore_obstacle, D, I)) return fail(e);
                I[3] = s * a.depth + a.produced[s] % a.depth; I[4] = ++a.produced[s];
                k++;
            }
            if (k == a.up_cap && a.produced[s] - a.seen[s] < a.depth) break;
        }
        if (k > 0) {
            CK(cudaMemcpyAsync(h->ar_up_dbl_d, a.up_dbl, dper * k * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(h->ar_up_int_d, a.up_int, iper * k * 4, cudaMemcpyHostToDevice, st));
            k_queue_scatter<<<k, 128, 0, st>>>(k, h->ar_up_dbl_d, h->ar_up_int_d, dper, iper, a.q_dbl, a.q_int, a.q_tail);
            CK(cudaEventRecord(a.ev_up, st));
        }
        if (k < a.up_cap) break;
    }
    if (!a.head_pending) {      // consumption counters back to the host, for the next top-up
        CK(cudaMemcpyAsync(a.head_h, a.q_head, (size_t)c.S * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(a.ev_head, st));
        a.head_pending = true; a.head_age = 0;
    }
    CK(cudaGetLastError());
    return 0;
}
extern "C" int imgenv_autoreset_enable(imgenv_t* h, imgenv_sampler_t* w, int32_t depth, int32_t ignore_obstacle, void* stream) {
    if (!h || !w) return fail("imgenv_autoreset_enable: null argument");
    const Cfg& c = h->d.c;
    if (h->ar.on) return fail("imgenv_autoreset_enable: already enabled");
    if (w->s.R != c.R || w->s.P != c.P || (int)w->s.rng.size() != c.S) return fail("imgenv_autoreset_enable: the sampler does not match the simulator (robots / pedestrians / scenes)");
    if ((int)w->s.objects.size() > c.max_obs) return fail("imgenv_autoreset_enable: more objects than max_obstacles");
    if (c.scene_type == 4) return fail("imgenv_autoreset_enable: dataset replay takes recorded trajectories (imgenv_reset)");
    if (depth < 2 || depth > 64) return fail("imgenv_autoreset_enable: depth must be in [2, 64]");
    imgenv::AutoReset& a = h->ar;
    CK(cudaSetDevice(h->device));
    const size_t dper = reset_dper(c), iper = reset_iper(c), S = c.S;
    a.depth = depth; a.smp = w; a.ignore_obstacle = ignore_obstacle;
    if (dalloc(h, &a.q_dbl, S * depth * dper) || dalloc(h, &a.q_int, S * depth * iper) || dalloc(h, &a.q_head, S) || dalloc(h, &a.q_tail, S) ||
        dalloc(h, &a.ids_d, S + 8) || dalloc(h, &a.n_dev, 1)) return -1;
    const size_t rec_bytes = dper * 8 + iper * 4;
    a.up_cap = (int)std::max<size_t>(1, std::min<size_t>(S * depth, ((size_t)64 << 20) / rec_bytes));
    CK(cudaMallocHost((void**)&a.up_dbl, (size_t)a.up_cap * dper * 8));
    CK(cudaMallocHost((void**)&a.up_int, (size_t)a.up_cap * iper * 4));
    CK(cudaMallocHost((void**)&a.head_h, S * 4));
    if (dalloc(h, &h->ar_up_dbl_d, (size_t)a.up_cap * dper) || dalloc(h, &h->ar_up_int_d, (size_t)a.up_cap * iper)) return -1;
    CK(cudaEventCreateWithFlags(&a.ev_head, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&a.ev_up, cudaEventDisableTiming));
    a.produced.assign(S, 0); a.seen.assign(S, 0);
    a.on = true;
    return imgenv_autoreset_refill(h, stream);
}
// Resets the scenes whose mask byte is non-zero (device pointer, uint8 [S]) from their episode queues: EnvPos.reset + ResetEnv.srv
// without any host involvement in the selection.  Stream-ordered, no host synchronisation.  refill != 0 also tops the queues up
// (host sampler, non-blocking); pass 0 inside a CUDA graph capture and call imgenv_autoreset_refill between replays.
extern "C" int imgenv_reset_masked(imgenv_t* h, const uint8_t* d_mask, int32_t refill, void* stream) {
    if (!h || !d_mask) return fail("imgenv_reset_masked: null argument");
    if (!h->ar.on) return fail("imgenv_reset_masked: auto-reset is not enabled (imgenv_autoreset_enable)");
    if (!h->outputs_bound) return fail("imgenv_reset_masked: outputs not bound (imgenv_bind_outputs)");
    imgenv::AutoReset& a = h->ar; Dev d = h->d; const Cfg& c = d.c;
    cudaStream_t st = (cudaStream_t)stream;
    if (refill && imgenv_autoreset_refill(h, stream)) return -1;
    if (h->tree_pending) { CK(cudaStreamWaitEvent(st, h->ev_tree, 0)); h->tree_pending = false; }
    d.n_dev = a.n_dev;
    k_compact_mask<<<1, 32, 0, st>>>(d_mask, c.S, a.ids_d, a.n_dev);
    k_apply_reset<<<c.S, 128, 0, st>>>(d, c.S, a.q_dbl, a.q_int, reset_dper(c), reset_iper(c), a.depth, a.ids_d, a.q_head, a.q_tail);
    k_object_footprints<<<c.S * c.max_obs, OBJ_THREADS, h->obj_smem, st>>>(d, a.ids_d);
    if (c.scene_type == 2 || c.scene_type == 3) k_rvo_build<<<c.S, 32, 0, st>>>(d, a.ids_d, a.ignore_obstacle);
    if (launch_observe(h, a.ids_d, c.S, 1, st, nullptr, a.n_dev)) return -1;
    CK(cudaGetLastError());
    return 0;
}
// Diagnostic counters since creation: out4[0] = ORCA obstacle-neighbour table overflows (an agent had more than ORCA_OBST_CAP
// facing obstacle edges in range and kept the nearest; the reference keeps all of them), out4[1..3] reserved.
// Work counters of the observation kernel, summed over robots since creation -- only an instrumented build (-DVIEW_STATS=1)
// writes them: out16[0] robots observed, [1] footprint records near the FOV, [2] their bitmap words, [3] static candidate
// blocks, [4] candidate cells, [5] raster cells that updated rays, [6] outputs evaluated in full, [7] outputs settled by the
// all-shadow test, [8] robots that ran the collision lattice, [9] robots that ran the FOV-edge pixels, [10] heavy cells,
// [11] robots with any ray hit.
extern "C" int imgenv_debug_view_stats(imgenv_t* h, int64_t* out16, void* stream) {
    if (!h || !out16) return fail("imgenv_debug_view_stats: null argument");
    CK(cudaMemcpyAsync(out16, h->d.counters + 4, 128, cudaMemcpyDeviceToHost, (cudaStream_t)stream));      // (+ imgenv_debug_view_phases)
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
// Instrumented builds: SM cycles thread 0 of the observation CTAs spent between phase boundaries, summed over robots:
// out8[0] prologue, [1] gather, [2] phase B (+ A), [3] heavy cells + laser ranges, [4] D1, [5] D2, [6] dirty outputs.
extern "C" int imgenv_debug_view_phases(imgenv_t* h, int64_t* out8, void* stream) {
    if (!h || !out8) return fail("imgenv_debug_view_phases: null argument");
    CK(cudaMemcpyAsync(out8, h->d.counters + 20, 64, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
extern "C" int imgenv_debug_counters(imgenv_t* h, int64_t* out4, void* stream) {
    if (!h || !out4) return fail("imgenv_debug_counters: null argument");
    unsigned long long r[4];
    CK(cudaMemcpyAsync(r, h->d.counters, 32, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    for (int k = 0; k < 4; k++) out4[k] = (int64_t)r[k];
    return 0;
}
// The RVO obstacle set of one scene as built on the device (tests compare it with the host restatement of
// RVOSimulator::addObstacle / KdTree::buildObstacleTreeRecursive in host_tables.h and with the reference node's obstacles_):
// verts[max_verts][8], nodes[max_verts][4]; returns n_verts in *n and the root in *root.
extern "C" int imgenv_debug_rvo_tree(imgenv_t* h, int32_t scene, int32_t* n, int32_t* root, float* verts, int32_t* nodes, double* corners, void* stream) {
    if (!h || !n || !root || !verts || !nodes || !corners) return fail("imgenv_debug_rvo_tree: null argument");
    Dev& d = h->d;
    if (scene < 0 || scene >= d.c.S) return fail("imgenv_debug_rvo_tree: bad scene");
    cudaStream_t st = (cudaStream_t)stream;
    int cnt[2];
    CK(cudaMemcpyAsync(cnt, d.rvo_counts + 2 * scene, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(verts, d.rvo_verts + (size_t)scene * d.max_verts * 8, (size_t)d.max_verts * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nodes, d.rvo_nodes + (size_t)scene * d.max_verts * 4, (size_t)d.max_verts * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(corners, d.sfm_obs + (size_t)scene * d.c.max_obs * 4, (size_t)d.c.max_obs * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n = cnt[0]; *root = cnt[1];
    return 0;
}
// Host restatement of the same build (no CUDA): corners[n_obj][4] = ax, ay, bx, by.  Returns the vertex count or -1.
extern "C" int imgenv_host_rvo_tree(const double* corners, int32_t n_obj, int32_t max_verts, int32_t* root, float* verts, int32_t* nodes) {
    std::vector<ht::RvoObst> robst;
    for (int k = 0; k < n_obj; k++) {
        const double* q = corners + 4 * k;
        ht::F2 v[4] = {ht::f2((float)q[0], (float)q[1]), ht::f2((float)q[0], (float)q[3]), ht::f2((float)q[2], (float)q[3]), ht::f2((float)q[2], (float)q[1])};
        ht::rvo_add_obstacle(robst, v, 4);
    }
    std::vector<ht::RvoNode> nd;
    std::vector<int> list(robst.size());
    for (size_t k = 0; k < list.size(); k++) list[k] = (int)k;
    *root = ht::rvo_build_tree(robst, nd, list);
    if ((int)robst.size() > max_verts) return -1;
    for (size_t k = 0; k < robst.size(); k++) {
        float* v = verts + 8 * k;
        v[0] = robst[k].px; v[1] = robst[k].py; v[2] = robst[k].dx; v[3] = robst[k].dy; v[4] = (float)robst[k].convex;
        v[5] = (float)robst[k].next; v[6] = (float)robst[k].prev; v[7] = 0;
    }
    for (size_t k = 0; k < nd.size(); k++) { nodes[4 * k] = nd[k].obstacle; nodes[4 * k + 1] = nd[k].left; nodes[4 * k + 2] = nd[k].right; nodes[4 * k + 3] = nd[k].parent; }
    return (int)robst.size();
}
extern "C" int imgenv_end_episode(imgenv_t* h, int32_t) { return h ? 0 : fail("imgenv_end_episode: null handle"); }
extern "C" int imgenv_solver_agents(const imgenv_t* h) { return h ? h->d.c.NA : 0; }
extern "C" int imgenv_view_dims(const imgenv_t* h, int32_t* vh, int32_t* vw) {
    if (!h) return fail("null handle");
    *vh = h->d.c.vh; *vw = h->d.c.vw;
    return 0;
}
extern "C" int64_t imgenv_algorithmic_bytes_per_robot_step(const imgenv_t* h) {
    // outputs written once + action/alive read + robot state read-modify-write
    const Cfg& c = h->d.c;
    return (int64_t)c.img * c.img * 2 + 3ll * c.img * c.img * 4 + 4ll * c.range_total + 4ll * c.state_dim + 4ll * c.pvs_len + 10 + 13 + 2 * 104;
}

extern "C" int imgenv_get_internal(imgenv_t* h, double* robot, double* ped, double* solver) {
    if (!h) return fail("null handle");
    Dev& d = h->d; const Cfg& c = d.c;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    size_t nr = (size_t)c.S * c.R, np = (size_t)c.S * c.P;
    if (robot) {
        std::vector<double> rb((size_t)RB_FIELDS * nr);
        CK(cudaMemcpy(rb.data(), d.rb, rb.size() * 8, cudaMemcpyDeviceToHost));
        const int map[16] = {RB_X, RB_Y, RB_YAW, RB_GX, RB_GY, RB_GYAW, RB_L0V, RB_L0W, RB_L1V, RB_L1W, RB_VX, RB_VY, RB_COLL, RB_ARR, RB_BEEP, RB_PREVD};
        for (size_t i = 0; i < nr; i++) for (int k = 0; k < 16; k++) robot[16 * i + k] = rb[(size_t)map[k] * nr + i];
    }
    if (ped && np) {
        std::vector<double> pd((size_t)PD_FIELDS * np);
        CK(cudaMemcpy(pd.data(), d.pd, pd.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < np; i++) for (int k = 0; k < PD_FIELDS; k++) ped[(size_t)PD_FIELDS * i + k] = pd[(size_t)k * np + i];
    }
    if (solver && c.NA) {
        size_t na = (size_t)c.S * c.NA;
        if (c.scene_type == 2 || c.scene_type == 3) {
            std::vector<float> p(2 * na), v(2 * na);
            CK(cudaMemcpy(p.data(), d.rvo_pos, p.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(v.data(), d.rvo_vel, v.size() * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < na; i++) { solver[4 * i] = p[2 * i]; solver[4 * i + 1] = p[2 * i + 1]; solver[4 * i + 2] = v[2 * i]; solver[4 * i + 3] = v[2 * i + 1]; }
        } else if (c.scene_type == 1) {
            CK(cudaMemcpy(solver, d.sfm, na * SFM_REC * 8, cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

extern "C" int imgenv_set_internal(imgenv_t* h, const double* robot, const double* ped, const double* solver) {
    if (!h) return fail("null handle");
    Dev& d = h->d; const Cfg& c = d.c;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    size_t nr = (size_t)c.S * c.R, np = (size_t)c.S * c.P;
    if (robot) {
        std::vector<double> rb((size_t)RB_FIELDS * nr);
        CK(cudaMemcpy(rb.data(), d.rb, rb.size() * 8, cudaMemcpyDeviceToHost));
        const int map[16] = {RB_X, RB_Y, RB_YAW, RB_GX, RB_GY, RB_GYAW, RB_L0V, RB_L0W, RB_L1V, RB_L1W, RB_VX, RB_VY, RB_COLL, RB_ARR, RB_BEEP, RB_PREVD};
        for (size_t i = 0; i < nr; i++) for (int k = 0; k < 16; k++) rb[(size_t)map[k] * nr + i] = robot[16 * i + k];
        CK(cudaMemcpy(d.rb, rb.data(), rb.size() * 8, cudaMemcpyHostToDevice));
    }
    if (ped && np) {
        std::vector<double> pd((size_t)PD_FIELDS * np);
        for (size_t i = 0; i < np; i++) for (int k = 0; k < PD_FIELDS; k++) pd[(size_t)k * np + i] = ped[(size_t)PD_FIELDS * i + k];
        CK(cudaMemcpy(d.pd, pd.data(), pd.size() * 8, cudaMemcpyHostToDevice));
    }
    if (solver && c.NA) {
        size_t na = (size_t)c.S * c.NA;
        if (c.scene_type == 2 || c.scene_type == 3) {
            std::vector<float> p(2 * na), v(2 * na);
            for (size_t i = 0; i < na; i++) { p[2 * i] = (float)solver[4 * i]; p[2 * i + 1] = (float)solver[4 * i + 1]; v[2 * i] = (float)solver[4 * i + 2]; v[2 * i + 1] = (float)solver[4 * i + 3]; }
            CK(cudaMemcpy(d.rvo_pos, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d.rvo_vel, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
        } else if (c.scene_type == 1) {
            CK(cudaMemcpy(d.sfm, solver, na * SFM_REC * 8, cudaMemcpyHostToDevice));
        }
    }
    return 0;
}

extern "C" int imgenv_debug_view_maps2(imgenv_t* h, uint8_t* host_out, int32_t* stats_out, void* stream);
extern "C" int imgenv_debug_view_maps(imgenv_t* h, uint8_t* host_out, void* stream) { return imgenv_debug_view_maps2(h, host_out, nullptr, stream); }

// Same, plus per-robot kernel statistics (int32 [S][R][4]: active raster tiles, boundary cells, heavy cells,
// 1 if the per-ray marching fallback ran). Either output may be NULL.
extern "C" int imgenv_debug_view_maps2(imgenv_t* h, uint8_t* host_out, int32_t* stats_out, void* stream) {
    if (!h) return fail("imgenv_debug_view_maps: null argument");
    Dev d = h->d; const Cfg& c = d.c;
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)c.S * c.R * c.vh * c.vw;
    uint8_t* buf = nullptr;
    int* sbuf = nullptr;
    if (host_out) CK(cudaMalloc((void**)&buf, n));
    if (stats_out) { CK(cudaMalloc((void**)&sbuf, (size_t)c.S * c.R * 16)); CK(cudaMemsetAsync(sbuf, 0, (size_t)c.S * c.R * 16, st)); }
    d.dbg_view = buf; d.dbg_stats = sbuf;
    k_view_consts<<<(c.S * c.R + VC_THREADS - 1) / VC_THREADS, VC_THREADS, 0, st>>>(d, nullptr, c.S, 1);
    k_footprints<<<(c.S * c.NPA + FOOT_WARPS - 1) / FOOT_WARPS, FOOT_WARPS * 32, h->foot_smem, st>>>(d, nullptr, c.S, 0);
    if (c.inverse_ok) k_view<true, false><<<c.S * c.R, VIEW_THREADS, h->view_smem, st>>>(d, nullptr, 0);
    else k_view<true, true><<<c.S * c.R, VIEW_THREADS, h->view_smem, st>>>(d, nullptr, 0);
    cudaError_t e = cudaSuccess;
    if (host_out) e = cudaMemcpyAsync(host_out, buf, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && stats_out) e = cudaMemcpyAsync(stats_out, sbuf, (size_t)c.S * c.R * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (buf) cudaFree(buf);
    if (sbuf) cudaFree(sbuf);
    if (e != cudaSuccess) return fail(std::string("imgenv_debug_view_maps: ") + cudaGetErrorString(e));
    return 0;
}

// libpedsim quadtree state of one scene (tests): nodes[n][5] = x,y,w,h,child0 ; leaf[NA][4] ; hash[NA]
extern "C" int imgenv_sfm_tree_get(imgenv_t* h, int32_t scene, int32_t* n_nodes, double* nodes, int32_t* leaf, int32_t* hash) {
    if (!h || h->d.c.scene_type != 1) return fail("imgenv_sfm_tree_get: not an SFM handle");
    Dev& d = h->d; const Cfg& c = d.c;
    CK(cudaDeviceSynchronize());
    int n = 0;
    CK(cudaMemcpy(&n, d.qt_nodes + scene, 4, cudaMemcpyDeviceToHost));
    *n_nodes = n;
    std::vector<double> box((size_t)QT_MAX_NODES * 4); std::vector<int> ch(QT_MAX_NODES);
    CK(cudaMemcpy(box.data(), d.qt_box + (size_t)scene * QT_MAX_NODES * 4, box.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ch.data(), d.qt_child0 + (size_t)scene * QT_MAX_NODES, ch.size() * 4, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; k++) { for (int q = 0; q < 4; q++) nodes[5 * k + q] = box[4 * k + q]; nodes[5 * k + 4] = ch[k]; }
    CK(cudaMemcpy(leaf, d.qt_leaf + (size_t)scene * c.NA * QT_LEAVES, (size_t)c.NA * QT_LEAVES * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hash, d.qt_hash + (size_t)scene * c.NA, (size_t)c.NA * 4, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int imgenv_sfm_tree_set(imgenv_t* h, int32_t scene, int32_t n_nodes, const double* nodes, const int32_t* leaf, const int32_t* hash) {
    if (!h || h->d.c.scene_type != 1) return fail("imgenv_sfm_tree_set: not an SFM handle");
    if (n_nodes < 1 || n_nodes > QT_MAX_NODES) return fail("imgenv_sfm_tree_set: bad node count");
    Dev& d = h->d; const Cfg& c = d.c;
    CK(cudaDeviceSynchronize());
    std::vector<double> box((size_t)QT_MAX_NODES * 4, 0.0); std::vector<int> ch(QT_MAX_NODES, -1), cnt(QT_MAX_NODES, 0);
    for (int k = 0; k < n_nodes; k++) { for (int q = 0; q < 4; q++) box[4 * k + q] = nodes[5 * k + q]; ch[k] = (int)nodes[5 * k + 4]; }
    for (int a = 0; a < c.NA; a++) for (int k = 0; k < QT_LEAVES; k++) { int n = leaf[a * QT_LEAVES + k]; if (n >= 0 && n < n_nodes) cnt[n]++; }
    CK(cudaMemcpy(d.qt_nodes + scene, &n_nodes, 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.qt_box + (size_t)scene * QT_MAX_NODES * 4, box.data(), box.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.qt_child0 + (size_t)scene * QT_MAX_NODES, ch.data(), ch.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.qt_count + (size_t)scene * QT_MAX_NODES, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.qt_leaf + (size_t)scene * c.NA * QT_LEAVES, leaf, (size_t)c.NA * QT_LEAVES * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.qt_hash + (size_t)scene * c.NA, hash, (size_t)c.NA * 4, cudaMemcpyHostToDevice));
    return 0;
}

// The byte map robot `self` would see as its global_map_ (static + reset objects + pedestrians + other robots;
// self < 0: peds_map_, i.e. no robots; self == -2: obs_map_, i.e. no pedestrians either) for one scene,
// u8 [H][W] to host. Test / debugging aid: the product path never materialises a map; here the footprint
// records of the scene are painted into a scratch "who covers this cell" plane and composed with the static grid.
__global__ void k_debug_paint(Dev d, int s, int self, uint8_t* cover) {
    const int q = blockIdx.x;
    const int4 h = d.foot_hdr[(size_t)s * d.c.NP + q];
    const int nrow = foot_nrow(h), wpr = foot_wpr(h), kind = foot_kind(h);
    if (!nrow) return;
    if (kind == FK_ROBOT && (self < 0 || q == self)) return;
    if (self == -2 && kind != FK_OBJ) return;
    const uint32_t* occ = d.foot_words + (size_t)s * d.c.scene_words + d.part_off[q];
    for (int k = threadIdx.x; k < nrow * wpr * 32; k += blockDim.x) {
        const int w = k >> 5, b = k & 31;
        if (!((occ[w] >> b) & 1u)) continue;
        const int cx = h.x + w / wpr, cy = (foot_wj0(h) + w % wpr) * 32 + b;
        const size_t ci = (size_t)cx * d.c.W + cy;
        atomicOr(reinterpret_cast<unsigned*>(cover + (ci & ~(size_t)3)), foot_flag(kind) << (8 * (ci & 3)));
    }
}
__global__ void k_debug_compose(Dev d, const uint8_t* cover, uint8_t* out) {
    const size_t n = (size_t)d.c.H * d.c.W;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        out[q] = (uint8_t)composed_value(d.grid[q], cover[q]);
}
extern "C" int imgenv_debug_global_map(imgenv_t* h, int32_t scene, int32_t self, uint8_t* host_out, void* stream) {
    if (!h || !host_out) return fail("imgenv_debug_global_map: null argument");
    Dev d = h->d; const Cfg& c = d.c;
    if (scene < 0 || scene >= c.S || self >= c.R) return fail("imgenv_debug_global_map: bad scene / robot");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = ((size_t)c.H * c.W + 3) & ~(size_t)3;
    uint8_t* buf = nullptr; uint8_t* cover = nullptr;
    CK(cudaMalloc((void**)&buf, n));
    CK(cudaMalloc((void**)&cover, n));
    CK(cudaMemsetAsync(cover, 0, n, st));
    // the record of EVERY agent of the scene (no culling in a whole-map dump: infinite reach)
    Dev dd = d; dd.c.cull_reach = 1e30;
    int* ids = nullptr;
    CK(cudaMalloc((void**)&ids, 4));
    CK(cudaMemcpyAsync(ids, &scene, 4, cudaMemcpyHostToDevice, st));
    k_footprints<<<(c.NPA + FOOT_WARPS - 1) / FOOT_WARPS, FOOT_WARPS * 32, h->foot_smem, st>>>(dd, ids, 1, 0);
    k_debug_paint<<<c.NP, 128, 0, st>>>(dd, scene, self, cover);
    k_debug_compose<<<592, 256, 0, st>>>(dd, cover, buf);
    cudaError_t e = cudaMemcpyAsync(host_out, buf, (size_t)c.H * c.W, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(buf); cudaFree(cover); cudaFree(ids);
    if (e != cudaSuccess) return fail(std::string("imgenv_debug_global_map: ") + cudaGetErrorString(e));
    return 0;
}

// Footprint-record invariants for the tests: every record's candidate cells are a subset of its occupied cells, lie inside
// the map and inside the record's box.  out[4] = non-empty records, occupied cells, candidate cells, violations.
__global__ void k_debug_check_footprints(Dev d, unsigned long long* out) {
    const size_t q = blockIdx.x;                                 // s * NP + part
    const int part = (int)(q % d.c.NP);
    const int4 h = d.foot_hdr[q];
    const int nrow = foot_nrow(h), wpr = foot_wpr(h);
    if (!nrow) return;
    const int po = d.part_off[part], cap = (d.part_off[part + 1] - po) >> 1;
    const uint32_t* occ = d.foot_words + (q / d.c.NP) * (size_t)d.c.scene_words + po;
    unsigned long long no = 0, nc = 0, bad = 0;
    for (int k = threadIdx.x; k < nrow * wpr; k += blockDim.x) {
        const unsigned o = occ[k], cd = occ[cap + k];
        no += __popc(o); nc += __popc(cd); bad += (cd & ~o) != 0;
        const int cx = h.x + k / wpr, wj = foot_wj0(h) + k % wpr;
        for (unsigned m = o; m; m &= m - 1) {
            const int cy = wj * 32 + __ffs(m) - 1;
            bad += (unsigned)cx >= (unsigned)d.c.H || (unsigned)cy >= (unsigned)d.c.W || cy < h.y || cy >= h.y + h.z;
        }
    }
    if (nrow * wpr > cap) bad++;
    if (threadIdx.x == 0) atomicAdd(out + 0, 1ull);
    if (no) atomicAdd(out + 1, no);
    if (nc) atomicAdd(out + 2, nc);
    if (bad) atomicAdd(out + 3, bad);
}
// words / headers of the agents' records that differ between two builds
__global__ void k_debug_diff_footprints(Dev d, const uint32_t* words_b, const int4* hdr_b, unsigned long long* out) {
    const size_t q = blockIdx.x;                                 // s * NP + part
    const int part = (int)(q % d.c.NP);
    if (part >= d.c.NPA) return;
    const int4 h = d.foot_hdr[q], hb = hdr_b[q];
    unsigned long long bad = 0;
    if (threadIdx.x == 0 && (h.x != hb.x || h.y != hb.y || h.z != hb.z || h.w != hb.w)) bad++;
    const int nrow = foot_nrow(h), wpr = foot_wpr(h);
    const int po = d.part_off[part], cap = (d.part_off[part + 1] - po) >> 1;
    const size_t base = (q / d.c.NP) * (size_t)d.c.scene_words + po;
    for (int k = threadIdx.x; k < nrow * wpr; k += blockDim.x)
        bad += (d.foot_words[base + k] != words_b[base + k]) + (d.foot_words[base + cap + k] != words_b[base + cap + k]);
    if (bad) atomicAdd(out + 3, bad);
}
extern "C" int imgenv_debug_check_footprints(imgenv_t* h, int64_t* out4, void* stream) {
    if (!h || !out4) return fail("imgenv_debug_check_footprints: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* buf = nullptr;
    CK(cudaMalloc((void**)&buf, 32));
    CK(cudaMemsetAsync(buf, 0, 32, st));
    k_debug_check_footprints<<<h->d.c.S * h->d.c.NP, 64, 0, st>>>(h->d, buf);
    {   // the records as k_footprints builds them (circle parts: analytic interior + rim points) must equal the records the
        // whole lattices give: rebuild them the slow way and compare word for word (differences count as violations)
        const Cfg& c = h->d.c;
        const size_t nw = (size_t)c.S * c.scene_words, nh = (size_t)c.S * c.NP;
        uint32_t* wcopy = nullptr; int4* hcopy = nullptr;
        CK(cudaMalloc((void**)&wcopy, nw * 4)); CK(cudaMalloc((void**)&hcopy, nh * 16));
        const int grid = (c.S * c.NPA + FOOT_WARPS - 1) / FOOT_WARPS;
        k_footprints<<<grid, FOOT_WARPS * 32, h->foot_smem, st>>>(h->d, nullptr, c.S, 0);      // (as the step builds them)
        CK(cudaMemcpyAsync(wcopy, h->d.foot_words, nw * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(hcopy, h->d.foot_hdr, nh * 16, cudaMemcpyDeviceToDevice, st));
        k_footprints<<<grid, FOOT_WARPS * 32, h->foot_smem, st>>>(h->d, nullptr, c.S, 2);      // whole lattices
        k_debug_diff_footprints<<<c.S * c.NP, 64, 0, st>>>(h->d, wcopy, hcopy, buf);
        k_footprints<<<grid, FOOT_WARPS * 32, h->foot_smem, st>>>(h->d, nullptr, c.S, 0);
        cudaStreamSynchronize(st);
        cudaFree(wcopy); cudaFree(hcopy);
    }
    unsigned long long r[4];
    cudaError_t e = cudaMemcpyAsync(r, buf, 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(std::string("imgenv_debug_check_footprints: ") + cudaGetErrorString(e));
    for (int k = 0; k < 4; k++) out4[k] = (int64_t)r[k];
    return 0;
}

// Episode record (EpRes.msg: per step the pose and request speeds of every alive robot, pose and velocity of every
// pedestrian; img_env.cpp:355-357, 397-408, 527-545).  The obs_map image of the message is imgenv_debug_global_map(scene, -2).
extern "C" int imgenv_record_enable(imgenv_t* h, int32_t max_steps) {
    if (!h || max_steps < 0) return fail("imgenv_record_enable: bad argument");
    Dev& d = h->d; const Cfg& c = d.c;
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (d.rec_rb) { cudaFree(d.rec_rb); d.rec_rb = nullptr; }
    if (d.rec_pd) { cudaFree(d.rec_pd); d.rec_pd = nullptr; }
    d.rec_T = 0;
    if (max_steps == 0) return 0;
    CK(cudaMalloc((void**)&d.rec_rb, std::max<size_t>(1, (size_t)c.S * max_steps * c.R * 6) * 8));
    CK(cudaMalloc((void**)&d.rec_pd, std::max<size_t>(1, (size_t)c.S * max_steps * c.P * 5) * 8));
    d.rec_T = max_steps;
    return 0;
}
extern "C" int imgenv_record_fetch(imgenv_t* h, int32_t scene, int32_t* n_steps, double* robots, double* peds, void* stream) {
    if (!h || !n_steps) return fail("imgenv_record_fetch: null argument");
    Dev& d = h->d; const Cfg& c = d.c;
    if (d.rec_T <= 0) return fail("imgenv_record_fetch: recording is not enabled (imgenv_record_enable)");
    if (scene < 0 || scene >= c.S) return fail("imgenv_record_fetch: bad scene");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long n = 0;
    CK(cudaMemcpyAsync(&n, d.step_no + scene, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int T = (int)std::min<unsigned long long>(n, (unsigned long long)d.rec_T);
    *n_steps = T;
    if (T > 0 && robots) CK(cudaMemcpyAsync(robots, d.rec_rb + (size_t)scene * d.rec_T * c.R * 6, (size_t)T * c.R * 6 * 8, cudaMemcpyDeviceToHost, st));
    if (T > 0 && peds && c.P) CK(cudaMemcpyAsync(peds, d.rec_pd + (size_t)scene * d.rec_T * c.P * 5, (size_t)T * c.P * 5 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}
