/*
 * imgenv.h — C ABI of libimgenv_b200.so: the B200-native replacement for the img_env
 * per-step simulation hot path.
 *
 * What it replaces in the reference (DRL-Navigation/img_env):
 *   - the four ROS1 services of the scene node      src/img_env/src/img_env.cpp:716-755
 *       init_image_env  (InitEnv.srv)   -> imgenv_create
 *       reset_image_env (ResetEnv.srv)  -> imgenv_reset
 *       step_image_env  (StepEnv.srv)   -> imgenv_step
 *       ep_end_image_env(EndEp.srv)     -> imgenv_end_episode (no-op: EpRes logging is out of scope)
 *   - AND the Python post-processing of the reply   envs/env/yaml_env.py:392-481
 *       (_get_states/_draw_ped_map/_trans_cv2_sensor_map/_norm_lasers), so the nine ImageState
 *       fields (envs/state/state.py:4-28) are produced on the device.
 *
 * One handle = one GPU = S independent scenes sharing one configuration (the reference runs
 * one ROS node per scene, create_launch.py:25-34).  Plain pointers and sizes only; no torch
 * types.  All device work is enqueued on the caller's cudaStream_t (passed as void*; NULL =
 * legacy default stream) and is stream-ordered: no host synchronisation inside step or reset
 * (reset waits at most for the reset before the previous one to have consumed its pinned staging).
 * The handle is not thread-safe.  Every entry point returns 0 on success, <0 on error;
 * imgenv_last_error() returns the message of the last failure on the calling thread.
 */
#ifndef IMGENV_H_
#define IMGENV_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct imgenv imgenv_t;

enum { IMGENV_SCENE_EMPTY = 0, IMGENV_SCENE_PEDSCENE = 1, IMGENV_SCENE_RVO = 2, IMGENV_SCENE_ERVO = 3,
       IMGENV_SCENE_DATASET = 4 /* trajectory replay, ImgEnv::_step_ped_dataset img_env.cpp:361-386 */ };
enum { IMGENV_SHAPE_CIRCLE = 0, IMGENV_SHAPE_RECTANGLE = 1, IMGENV_SHAPE_LEG = 2 };
enum { IMGENV_KTYPE_DIFF = 0, IMGENV_KTYPE_OMNI = 1 };

/* InitEnv.srv scalars (float32 on the ROS wire: the library rounds them through float exactly
 * like the node sees them) + the cfg keys only yaml_env.py consumes. */
typedef struct imgenv_config {
    double view_resolution, view_width, view_height;   /* cfg view_map.{resolution,width,height} */
    double step_hz;                                     /* cfg control_hz: a PERIOD in seconds     */
    int32_t state_dim;                                  /* 3, 4 or 5 (agent.cpp:156-184)           */
    int32_t use_laser, range_total;
    double view_angle_begin, view_angle_end, view_min_dist, view_max_dist;
    double beep_r, ped_ca_p;                            /* never filled by the reference's Python: 0 */
    int32_t relation_ped_robo;
    /* python side (yaml_env.py:133-181) */
    int32_t image_size, ped_image_size;                 /* 48, 48 (square)                          */
    int32_t max_ped, ped_vec_dim;                       /* ped_vector_states is [1 + ped_vec_dim*max_ped] */
    double ped_image_r, laser_max;
    int32_t laser_norm;
    /* batch shape */
    int32_t num_scenes, num_robots, num_peds;           /* S, R, P per scene                        */
    int32_t scene_type;                                 /* IMGENV_SCENE_*  (cfg ped_sim.type)      */
    int32_t robot_ktype;                                /* IMGENV_KTYPE_*  (cfg robot_type)        */
    int32_t max_obstacles;                              /* reset objects per scene (cfg object.total) */
    int32_t max_traj;                                   /* waypoints per pedestrian (1 or 2 from reset_helper.py:337-342) */
    uint64_t seed;                                      /* beep Bernoulli draws (img_env.cpp:327 uses glibc rand()) */
    double max_object_radius;                           /* bound (m) on a reset object's bounding radius; 0 = 0.75 */
} imgenv_config;

/* Device pointers the library WRITES every reset/step (caller-owned, e.g. torch tensors; bound
 * once with imgenv_bind_outputs).  Rows of robots whose view is frozen (already collided /
 * arrived: agent.cpp:358-360) keep sensor_maps/lasers/is_collisions untouched, as the node
 * re-sends its stale buffers (img_env.cpp:553-565).  Layouts are C-contiguous. */
typedef struct imgenv_outputs {
    float*    vector_states;      /* [S,R,state_dim]   AgentState.state (float32 on the wire)     */
    uint16_t* sensor_maps;        /* [S,R,48,48] IEEE binary16: cubic-resized view_map / 255       */
    int8_t*   is_collisions;      /* [S,R] 0 none, 1 obstacle, 2 pedestrian, 3 robot               */
    uint8_t*  is_arrives;         /* [S,R]                                                          */
    float*    lasers;             /* [S,R,range_total]  (/laser_max when laser_norm)               */
    float*    ped_vector_states;  /* [S,R,1+ped_vec_dim*max_ped]                                    */
    float*    ped_maps;           /* [S,R,3,48,48]                                                  */
    float*    step_ds;            /* [S,R]                                                          */
    float*    ped_min_dists;      /* [S,R]  +inf until a pedestrian exists                          */
} imgenv_outputs;

/* robot_desc[R][25]: shape, size[4], sensor_cfg[2], speed_limiter_v[9], speed_limiter_w[9]
 *   (limiter = has_velocity, has_acceleration, has_jerk, min_v, max_v, min_a, max_a, min_j, max_j;
 *    Agent.msg / SpeedLimiter.msg, built by reset_helper.py:374-393)
 * ped_desc[P][8]:    shape, size[6], max_speed                       (reset_helper.py:395-412)
 * robot_size_last[R]: cfg robot.size[i][-1] as the python double yaml_env.py:407 adds to ped_r
 * grid: the occupancy PNG already decoded (cv2.imread GRAYSCALE) and resized to the view
 *   resolution as grid_map.cpp:28-38 does; H rows (world x), W cols (world y). */
int imgenv_create(const imgenv_config* cfg, const uint8_t* grid, int32_t H, int32_t W,
                  const double* robot_desc, const double* ped_desc, const double* robot_size_last,
                  int32_t device, imgenv_t** out);
int imgenv_destroy(imgenv_t* h);
int imgenv_bind_outputs(imgenv_t* h, const imgenv_outputs* out);

/* ResetEnv.srv for n scenes (host arrays, float64; poses as x,y,qx,qy,qz,qw like geometry_msgs/Pose).
 *   scene_ids[n]; n_obs[n]; obs[n][max_obstacles][11] = shape,size[4],x,y,q[4]
 *   robots[n][R][8] = x,y,q[4],goal_x,goal_y ;  peds[n][P][8] likewise
 *   traj_len[n][P] ; traj[n][P][max_traj][3]   (Agent.msg trajectory, reset_helper.py:337-342)
 *   traj_v[n][P][max_traj][3] or NULL          (Agent.msg trajectory_v; dataset replay only, reset_helper.py:417-434)
 * Runs the node's _reset (img_env.cpp:162-292) incl. view_agent + get_states, then the Python
 * _get_states; outputs of those scenes are written to the bound tensors. */
int imgenv_reset(imgenv_t* h, int32_t n, const int32_t* scene_ids, const int32_t* n_obs, const double* obs,
                 const double* robots, const double* peds, const int32_t* traj_len, const double* traj,
                 const double* traj_v, int32_t ignore_obstacle, void* stream);

/* Episode sampler: EnvPos.reset (envs/utils/reset_helper.py:115-345) in native code, drawing from a bit-exact
 * CPython `random` (MT19937) so that the same seed gives the poses the reference's Python sampler gives.
 * One generator per scene (seed + scene index).  desc (float64):
 *   [0] R  [1] P  [2] n_objects  [3..4] circle_ranges  [5] target_min_dist  [6] go_back 0 yes / 1 no / 2 random  [7] 0
 *   (R+P) agent records of 103: module size, then begin and target pose specs of 51 each:
 *       type bits (1 =='fix', 2 =='rand_angle', 4 'range', 8 'circle', 16 'fix' substring, 32 'multi', 64 'view',
 *       128 'plus', 256 'circle_fix'), n_multi, len of one range, 8 x 6 values (row 0 = the pose / range)
 *   n_objects records of 14: shape (0 circle / 1 rectangle), fixed, len(pose), pose[6], size_range[4], 0
 * Host-only: no CUDA call is made by the sampler itself. */
typedef struct imgenv_sampler imgenv_sampler_t;
int imgenv_sampler_create(const double* desc, int64_t n_desc, int32_t n_scenes, uint64_t seed, imgenv_sampler_t** out);
int imgenv_sampler_destroy(imgenv_sampler_t* s);
int imgenv_sampler_seed(imgenv_sampler_t* s, int32_t scene, uint64_t seed);
/* Samples n scenes into arrays laid out like imgenv_reset's arguments. */
int imgenv_sampler_sample(imgenv_sampler_t* s, int32_t n, const int32_t* scene_ids, int32_t max_obs, int32_t max_traj,
                          int32_t* n_obs, double* obs, double* robots, double* peds, int32_t* traj_len, double* traj);
/* One draw from a scene's generator: kind 0 random(), 1 uniform(a,b), 2 gauss(a,b), 3 randint(a,b). */
int imgenv_sampler_draw(imgenv_sampler_t* s, int32_t scene, int32_t kind, double a, double b, double* out);
/* EnvPos.reset + ResetEnv.srv for the listed scenes in one call (yaml_env.py:232-262 without the Python loop). */
int imgenv_reset_sampled(imgenv_t* h, imgenv_sampler_t* s, int32_t n, const int32_t* scene_ids, int32_t ignore_obstacle, void* stream);

/* Device-side auto-reset.  imgenv_autoreset_enable gives every scene a queue of `depth` PRE-SAMPLED episodes in device memory (the
 * host sampler runs ahead on the scenes' own generator streams, so each scene sees exactly the episode sequence synchronous
 * imgenv_reset_sampled calls would give it).  imgenv_reset_masked then resets the scenes selected by a DEVICE mask (uint8 [S]) from
 * their queues -- reset objects, RVO obstacle ring + BSP, poses, first observation -- stream-ordered and without any host
 * synchronisation, so an RL loop (NeverStopWrapper, envs/wrapper/base.py:193-211) never leaves the device.  refill != 0 also tops
 * the queues up (non-blocking); inside a CUDA graph pass 0 and call imgenv_autoreset_refill between replays.  A scene whose queue
 * ran empty replays its newest episode and is counted in imgenv_debug_counters out4[2]. */
int imgenv_autoreset_enable(imgenv_t* h, imgenv_sampler_t* s, int32_t depth, int32_t ignore_obstacle, void* stream);
int imgenv_autoreset_refill(imgenv_t* h, void* stream);
int imgenv_reset_masked(imgenv_t* h, const uint8_t* d_mask, int32_t refill, void* stream);

/* StepEnv.srv for all S scenes. d_actions[S][R][3] = v, w, v_y(beep) float32 DEVICE pointer;
 * d_alive[S][R] uint8 DEVICE pointer, or NULL to use the library's own dones bookkeeping
 * (yaml_env.py:319-331,373-377: alive = not (collided or arrived) after the previous call). */
int imgenv_step(imgenv_t* h, const float* d_actions, const uint8_t* d_alive, void* stream);
/* Same call with HOST buffers (pinned or pageable): copies in, steps, leaves outputs on device. */
int imgenv_step_host(imgenv_t* h, const float* h_actions, const uint8_t* h_alive, void* stream);
int imgenv_end_episode(imgenv_t* h, int32_t scene_id);

/* Internal state in/out for single-step parity tests. Host arrays.
 *   robot[S][R][16]: x,y,yaw, gx,gy,gyaw, last0 v,w, last1 v,w, vx,vy, is_collision,is_arrive,beep, prev_goal_dist(NaN=None)
 *   ped[S][P][20]:   x,y,yaw, lx,ly,lyaw, vx,vy, gait state,last_state,remaining, lleg xyz, rleg xyz, traj_idx, 0,0
 *   solver: RVO/ERVO [S][P+R'][4] float64 holding px,py,vx,vy ; SFM [S][P+R'][12]
 *           (p.xyz, v.xyz, vmax, dest, lastdest, deque_front, in_tree, 0)   R' = R if relation_ped_robo==1 else 0 */
int imgenv_get_internal(imgenv_t* h, double* robot, double* ped, double* solver);
int imgenv_set_internal(imgenv_t* h, const double* robot, const double* ped, const double* solver);
/* Debug raster: the 400x400 view_map_ of every robot as the node would send it (u8 [S][R][vh][vw]). */
int imgenv_debug_view_maps(imgenv_t* h, uint8_t* host_out, void* stream);
/* Same + per-robot kernel statistics int32 [S][R][4] (active raster tiles, boundary cells, heavy cells, marching
 * fallback taken). Either output may be NULL. */
int imgenv_debug_view_maps2(imgenv_t* h, uint8_t* host_out, int32_t* stats_out, void* stream);
/* Debug raster: the composited byte map of one scene, u8 [H][W] to host. self >= 0: robot self's
 * global_map_ (img_env.cpp:623-628); self == -1: peds_map_ (img_env.cpp:594-618); self == -2: obs_map_ (img_env.cpp:167-187). */
int imgenv_debug_global_map(imgenv_t* h, int32_t scene, int32_t self, uint8_t* host_out, void* stream);
/* Episode record (EpRes.msg; img_env.cpp:355-357, 397-408, 527-545): once enabled, every step stores per scene the pose and
 * request speeds of every robot (robots[t][R][6] = x, y, yaw, v, w, alive -- the node only appends alive robots) and the pose
 * and velocity of every pedestrian (peds[t][P][5] = x, y, yaw, vx, vy), up to max_steps steps after the scene's last reset.
 * max_steps = 0 disables and frees.  fetch copies the steps recorded since the last reset of `scene` to host arrays. */
int imgenv_record_enable(imgenv_t* h, int32_t max_steps);
int imgenv_record_fetch(imgenv_t* h, int32_t scene, int32_t* n_steps, double* robots, double* peds, void* stream);
/* Invariant check (tests) of the per-part footprint records the observation composes: out4 = non-empty records, occupied cells,
 * candidate cells, violations (a candidate cell that is not occupied, a cell outside the map or outside its record's box). */
int imgenv_debug_check_footprints(imgenv_t* h, int64_t* out4, void* stream);
/* Test hook: the node's SpeedLimiter(msg) leaves min_jerk unassigned (speed_limit.cpp:56-65) and clamps with whatever its
 * stack held; the library defaults to min_jerk = max_jerk = msg.min_jerk.  min_jerk[R][2] (linear, angular) overrides it. */
int imgenv_debug_set_min_jerk(imgenv_t* h, const double* min_jerk);
/* Pedestrian yaw the node reads from an unassigned local (img_env.cpp:346-349): 0 keep, 1 zero (this build of the node), 2 heading. */
int imgenv_set_ped_yaw_mode(imgenv_t* h, int mode);
/* Diagnostic counters since creation: out4[0] = times an ORCA agent had more facing obstacle edges in range than the solver's
 * per-agent table holds (it then keeps the nearest; the reference keeps all); out4[1] = obstacle BSP builds that ran out of space;
 * out4[2] = masked resets that found an empty episode queue.  Tests assert 0. */
int imgenv_debug_counters(imgenv_t* h, int64_t* out4, void* stream);
/* Work counters of the observation kernel summed over robots since creation; written only by an instrumented build of the
 * library (-DVIEW_STATS=1, tools/view_stats.py), all zero otherwise.  out16: [0] robots observed, [1] footprint records near
 * the field of view, [2] their bitmap words, [3] static candidate blocks, [4] candidate cells, [5] raster cells that updated
 * rays, [6] outputs evaluated in full, [7] outputs settled by the all-shadow test, [8] robots that ran the collision lattice,
 * [9] robots that ran the FOV-edge pixels, [10] heavy cells, [11] robots with any ray hit, [12] segments of 8 outputs listed. */
int imgenv_debug_view_stats(imgenv_t* h, int64_t* out16, void* stream);
/* Instrumented builds only: SM cycles between the phase boundaries of the observation CTAs (thread 0's clock), summed over
 * robots: out8[0] prologue, [1] gather, [2] phase B (+ collision lattice), [3] heavy cells + laser ranges, [4] segment
 * classification, [5] listed outputs, [6] dirty outputs, [7] 0. */
int imgenv_debug_view_phases(imgenv_t* h, int64_t* out8, void* stream);
/* Tests: the RVO obstacle set of one scene as the reset kernels built it on the device (vertex ring verts[max_verts][8] = px, py,
 * edge dir x, y, convex, next, prev, 0; BSP nodes[max_verts][4] = edge, left, right, parent; max_verts = 16 * max_obstacles + 16;
 * corners[max_obstacles][4] = the two rotated corners per reset object the ring was built from), and the host restatement of the
 * same construction (RVOSimulator.cpp:130-168, KdTree.cpp:119-257; no CUDA), returning the vertex count or -1. */
int imgenv_debug_rvo_tree(imgenv_t* h, int32_t scene, int32_t* n, int32_t* root, float* verts, int32_t* nodes, double* corners, void* stream);
int imgenv_host_rvo_tree(const double* corners, int32_t n_obj, int32_t max_verts, int32_t* root, float* verts, int32_t* nodes);
/* ped_min_dists persistence (NearbyPed, reset_helper.py:85-99) and dones are library state. */
int imgenv_solver_agents(const imgenv_t* h);   /* P + R' */
int imgenv_view_dims(const imgenv_t* h, int32_t* vh, int32_t* vw);
/* Number of kernel launches one imgenv_step enqueues (for bench.py's gpu_launches claim). */
int imgenv_launches_per_step(const imgenv_t* h);
/* Algorithmic HBM bytes one robot-step must move (SURVEY.md §8d formula for this config). */
int64_t imgenv_algorithmic_bytes_per_robot_step(const imgenv_t* h);
const char* imgenv_last_error(void);
const char* imgenv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* IMGENV_H_ */
