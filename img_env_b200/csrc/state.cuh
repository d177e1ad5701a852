// Device-resident state of one handle (S scenes on one GPU) and the static per-config tables.
// Layout notes (DESIGN.md "Data layout in HBM"):
//   * one static occupancy grid (u8, H x W), its 1-bit "value < 250" plane and the 1-bit plane of its ray-hit
//     candidates (occupied cells with a free cell in their 5x5 neighbourhood), shared by all scenes (L2-resident),
//     instead of the reference's R full-map clones per step (img_env.cpp:623);
//   * per scene NO map-sized state: every robot, pedestrian part and reset object keeps a FOOTPRINT RECORD -- the
//     set of grid cells it covers as a small bitmap (box header + occupancy words + candidate words, foot.cuh).
//     Robots' and pedestrians' records are rebuilt every step (O(footprint)), objects' at reset; the observation
//     kernel composes the records near a robot on the fly;
//   * robots / pedestrians as SoA of doubles ([S][R], [S][P]).
#pragma once
#include <stdint.h>
#include "tfmath.cuh"

// who covers a cell (bit set), as the compositing rules of view_ped / view_robot / Agent::draw need it
#define F_OBJ 1u      // reset object wrote 0 here   (img_env.cpp:187)
#define F_RIGHT 2u    // right leg (overwrites 0)    (agent.cpp:757-772)
#define F_LEFT 4u     // left leg                    (agent.cpp:742-756)
#define F_CIRC 8u     // circle pedestrian           (img_env.cpp:599-601)
#define F_ROBOT 16u   // footprint of a robot other than the observer
// footprint record kinds (foot.cuh)
enum { FK_ROBOT = 0, FK_CIRC = 1, FK_LEFT = 2, FK_RIGHT = 3, FK_OBJ = 4 };

#define MAX_SPANS 2
#define RB_FIELDS 18   // doubles per robot
#define PD_FIELDS 20   // doubles per pedestrian

// per-robot SoA field ids (double)
enum { RB_X = 0, RB_Y, RB_YAW, RB_GX, RB_GY, RB_GYAW, RB_L0V, RB_L0W, RB_L1V, RB_L1W, RB_VX, RB_VY,
       RB_COLL, RB_ARR, RB_BEEP, RB_PREVD, RB_MIND, RB_DONE };
// per-ped SoA field ids (double)
enum { PD_X = 0, PD_Y, PD_YAW, PD_LX, PD_LY, PD_LYAW, PD_VX, PD_VY, PD_GAIT, PD_LGAIT, PD_REM,
       PD_LLX, PD_LLY, PD_LLZ, PD_RLX, PD_RLY, PD_RLZ, PD_TIDX, PD_PAD0, PD_PAD1 };

struct Limiter {
    int has_v, has_a, has_j;
    double min_v, max_v, min_a, max_a, min_j, max_j;
};

// One robot "type" = identical shape/sensor description (deduplicated on the host).
struct RobotType {
    int n_pts;            // footprint lattice points (agent.cpp:18-62)
    int pts_off;          // offset into lattice_xy (double2 units)
    int org_x, org_y;     // laser origin cell (agent.cpp:366-369)
    int ray_off;          // offset into ray_end (short2 units), range_total entries
    int span_off;         // offset into fov_spans (vh * MAX_SPANS * 2 shorts)
    int zone_r0, zone_r1, zone_c0, zone_c1;  // view-raster box that contains the robot's own footprint cells (own_mask)
    int khi_off;          // offset into kpack (vh*vw entries per type)
    int own_mask_off;     // offset into own_mask (vh*vw bits, u32 words)
    int tile_off;         // offset into tile_fov (u32 words): bit per 32x32 view tile that contains FOV pixels
    int edge_off, n_edge; // FOV pixels with an 8-neighbour outside the FOV (u32 row<<16|col), always evaluated forward
    int etile_off;        // offset into edge_tiles (u32 words, even): per row of 16x16 view tiles a 64-bit mask of the tiles that hold an edge pixel
    int fov_r0, fov_r1, fov_c0, fov_c1;   // bounding box (inclusive) of the FOV pixels in the view raster
    int dtab_off;         // offset into dtab (ns*img*4 u32)
    int ostat_off;        // offset into ostat (img*img uint2)
    double stamp_cx, stamp_cy; int stamp_rad;   // footprint lattice: centre (base frame, m) and radius (cells, incl. margin) of its bounding circle
    int ring_off, ring_n; // circle robots: lattice points near the rim (offset into lattice_xy, count; -1 = evaluate the whole lattice)
    double disc_cx, disc_cy, disc_rin;   // ... and the disc (base frame, m) every cell centre inside which is covered for sure
    int zone_rad;         // bound (world cells) on the distance from the robot's position to any cell of its footprint
    double size_last;     // python: robots[i].size[-1]
    double sensor_x, sensor_y;
};

struct Cfg {
    // geometry
    int S, R, P, NA;         // NA = solver agents = P + (relation ? R : 0)
    int dyn_threads, dyn_nblk;    // dynamics kernels: agents per CTA, CTAs per scene (dyn.cuh: dyn_pick_block)
    int H, W, Wb, Hc;        // grid rows, cols, u32 words per bit-plane row, 32-row blocks
    int vh, vw, vwb;         // view raster rows/cols, words per row
    double res, inv_res;     // view resolution == grid resolution (float32 widened), 1/res
    double cull_reach;       // farthest world distance from a robot's position to anything its observation can read
    double step_hz, control_hz;   // period (float32 widened), 0.05
    int state_dim, use_laser, range_total, ktype, scene_type, relation;
    int inverse_ok;               // lasers on and the FOV spans few enough world blocks: world->view rasterisation, no raster in shared memory
    double beep_r, ped_ca_p;
    double view_max_dist;
    Tf2 view_base, base_view;     // tf_view_base_, tf_base_view_
    // python side
    int img, ns;                  // output image side (48), needed source rows/cols (144)
    unsigned img_inv;             // ceil(2^32 / img): q / img == __umulhi(q, img_inv) for q < 2^16
    int hb_shift;                 // rays are grouped in blocks of 2^hb_shift (< 64 blocks) for the nearest-hit table
    int max_ped, ped_vec_dim, pvs_len;
    double ped_image_r, ped_res, laser_max;
    int laser_norm;
    int max_obs, max_traj;
    unsigned long long seed;
    int n_types;
    // footprint records (foot.cuh): parts per scene = R robots, 2 per pedestrian (body or left leg, right leg), max_obs objects
    int NPA, NP;                  // agent parts (R + 2P), all parts (NPA + max_obs)
    int ag_cap, obj_cap;          // bitmap capacity (u32 words) of the largest agent part / of an object
    int obj_rad;                  // largest object bounding radius in cells (incl. margin)
    int scene_words;              // u32 words of record bitmaps per scene
};

// shared-memory plans of k_view / k_ped_obs (byte offsets; view.cuh computes them once on the host)
struct ViewLayout { unsigned sh, regA, regB, hpre, hitkey, rays, need, spans, blocks, near, npre, chdr, coff, nhdr, noff, cword, cmeta, cpre, cwsum, seglist, total; };
struct PedLayout { unsigned winner, keys, dkeys, pobs, row, total, paint; int n_sort; };
struct ViewConst;
struct Dev {
    Cfg c;
    ViewLayout vl; PedLayout pl;
    // static, shared
    const uint8_t* grid;          // [H][W]
    const uint32_t* static_occ;   // [H][Wb]  bit = grid < 250
    const uint32_t* static_cand;  // [H][Wb]  static_occ cells that are within 2 cells of a free cell or of the map border
    const uint32_t* static_crow;  // [Hc][Wb] per 32x32-cell block: bit r = row r of the block has a static_cand bit
    const uint32_t* static_orow;  // [Hc][Wb] the same for static_occ
    const RobotType* types;       // [n_types]
    const int* type_of;           // [R]
    const double* lattice_xy;     // packed (x,y) pairs
    const short* ray_end;         // (x2,y2) pairs
    const short* fov_spans;       // per type: [vh][MAX_SPANS][2] (c0,c1 exclusive), -1 = none
    const uint32_t* kpack;        // per view pixel: highest ray through it | lowest << 16 | (rays form exactly that interval) << 31; 0xFFFF = none
    const uint32_t* own_mask;     // per type: bit per view pixel = own footprint cell
    const uint32_t* tile_fov;     // per type: bit per 32x32 view tile (row-major, vwb per row) with any FOV pixel
    const uint32_t* edge_px;      // per type: FOV-edge pixels
    const uint32_t* edge_tiles;   // per type: bit per 16x16 view tile near an edge pixel
    const uint32_t* dtab;         // per type [ns][img][4]: for tap k of output column oc on needed row rr: top ray (12b) | its step index there (10b) << 12 | own footprint << 31
    const uint32_t* ostat;        // per type: [img*img] u32 lowest (12 b) | highest (12 b) top ray over the source pixels of an output pixel |
                                  //           ceil(their largest Chebyshev distance to the laser origin / 4) << 24, then [img*img] u16: the
                                  //           output's float16 value when none of those rays hits anything in front of / on a source pixel
    const short* need_idx;        // [ns] source row/col index of the k-th needed row/col
    const short* cubic_tap;       // [img][4] index into need_idx space (0..ns-1) of the 4 taps
    const short* cubic_coef;      // [img][4] fixed-point weights (x2048)
    const uint16_t* f16_lut;      // [256] half(x/255)
    const Limiter* lim_v; const Limiter* lim_w;   // [R]
    // pedestrian footprints
    const int* ped_shape;         // [P]
    const double* ped_size;       // [P][6] (float32 widened)
    const double* ped_maxspeed;   // [P]
    const double* ped_r_round;    // [P] python round(r_,2)
    const int* ped_pts_off;       // [P][2] lattice offsets (body or left leg, right leg)
    const int* ped_pts_n;         // [P][2]
    const int* ped_ring_off;      // [P][2] rim points of the part's circle lattice (offset into lattice_xy)
    const int* ped_ring_n;        // [P][2] their number, -1 = evaluate the whole lattice
    const double* ped_disc;       // [P][2][3] disc centre (part frame, m) and the radius inside which every cell centre is covered
    const double* ped_part;       // [P][2][3] per stamped part (body / left leg, right leg): bounding circle centre x, y (m) and radius (cells)
    const double* ped_ext;        // [P] bound on the distance from the pedestrian position to any cell it stamps
    // per scene footprint records (foot.cuh)
    int4* foot_hdr;               // [S][NP] first row (world cell x), first 32-cell word column, rows | words per row << 16, kind | id << 4
    uint32_t* foot_words;         // [S][scene_words] per part: cap occupancy words, then cap candidate words
    const int* part_off;          // [NP + 1] word offset of a part's bitmaps inside its scene's slice (cap = (off[q+1] - off[q]) / 2)
    struct ViewConst* vconst;     // [S][R] pose-dependent constants of every robot's observation (foot.cuh), rewritten by k_footprints
    // dynamic state
    double* rb;                   // [RB_FIELDS][S*R]
    double* pd;                   // [PD_FIELDS][S*P]
    double* traj;                 // [S][P][max_traj][3]
    double* traj_v;               // [S][P][max_traj][3] (dataset replay)
    int* traj_len;                // [S][P]
    double* obs;                  // [S][max_obs][8]: shape, size[4], x, y, yaw
    int* n_obs;                   // [S]
    unsigned long long* step_no;  // [S]
    // optional episode record (EpRes, img_env.cpp:355-357, 397-408): rec_T > 0 enables it
    double* rec_rb;               // [S][rec_T][R][6] x, y, yaw, v, w (request values), alive
    double* rec_pd;               // [S][rec_T][P][5] x, y, yaw, vx, vy
    int rec_T;
    // solver state
    float* rvo_pos; float* rvo_vel;         // [S][NA][2]
    float* rvo_nvel;                        // [S][NA][2] new velocities (scratch between solve and apply)
    double* sfm_force;                      // [S][NA][12] SFM forces (scratch between solve and apply)
    float* rvo_verts;                       // [S][max_verts][8]: px,py,dx,dy,convex,next,prev,0
    int* rvo_nodes;                         // [S][max_verts][4]: obstacle edge, left child, right child, parent
    float* rvo_nodeseg;                     // [S][max_verts][4]: end points of every BSP node's edge
    int* rvo_arena; int rvo_arena_len;      // [S][rvo_arena_len] edge lists of the device-side BSP build (rvotree.cuh)
    unsigned long long* counters;           // [4] diagnostics: [0] ORCA obstacle-neighbour / line table overflows, [1] obstacle BSP build overflows
    unsigned char* orca_pool; int orca_nslabs; unsigned* orca_cursor;   // overflow slabs of the ORCA tables (orca.cuh), [2] cursors used alternately
    int* rvo_counts;                        // [S][2]: n_verts, root(-1 none)
    int max_verts;
    double* sfm;                            // [S][NA][12]
    double* sfm_obs;                        // [S][max_obs][4] segment ax,ay,bx,by
    int* sfm_nobs;                          // [S]
    double* sfm_wp;                         // [S][P][1+max_traj][3] waypoints x,y,r
    // libpedsim quadtree emulation (sfmtree.cuh), per scene
    int* qt_nodes;                          // [S]
    double* qt_box;                         // [S][QT_MAX_NODES][4]
    int* qt_child0; int* qt_count;          // [S][QT_MAX_NODES]
    int* qt_leaf;                           // [S][NA][4]
    int* qt_hash;                           // [S][NA]
    double* sfm_newpos;                     // [S][NA][4] post-move and pre-move positions handed to the tree update
    double* sfm_treepos;                    // [S][NA][2] scratch of the tree update: every agent's position "right now"
    const double* sfm_vmax0;                // [NA] initial vmax (Tagent(): N(1.2,0.2), setVmax for pedestrians)
    // outputs
    float* o_vec; uint16_t* o_sensor; int8_t* o_coll; uint8_t* o_arr; float* o_laser;
    float* o_pvs; float* o_pmap; float* o_stepd; float* o_mind;
    const int* n_dev;             // optional: number of listed scenes, on the device (masked resets launch for S scenes; blocks beyond *n_dev exit)
    uint8_t* dbg_view;            // optional [S][R][vh][vw]
    int* dbg_stats;               // optional [S][R][4]: active raster tiles, boundary cells, heavy cells, marching fallback
};

__host__ __device__ inline double& RBF(const Dev& d, int f, int idx) { return d.rb[(size_t)f * d.c.S * d.c.R + idx]; }
__host__ __device__ inline double& PDF(const Dev& d, int f, int idx) { return d.pd[(size_t)f * d.c.S * d.c.P + idx]; }
