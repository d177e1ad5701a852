// TEST INFRASTRUCTURE — not part of the product. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load the library built
// from this file (oracle/_ref/libimgenv_ref.so).
//
// C-ABI driver around the UNMODIFIED reference node. The reference's own
// translation units (src/img_env/src/{img_env,agent,grid_map,speed_limit}.cpp,
// src/3rdparty/ervo_ros/src/*.cpp, src/3rdparty/pedsimros/src/ped_*.cpp) are
// compiled where they lie under /root/reference against the shim headers in
// oracle/ref_shim/ (ROS / tf / OpenCV are absent in this image). This file only
// (a) fills the InitEnv/ResetEnv/StepEnv request structs the way
// envs/env/yaml_env.py:183-209,223-247,319-331 does (float32 wire fields
// included), (b) calls EnvService::{init_env,reset_env,step_env}
// (img_env.cpp:726-749) and (c) copies the AgentState reply and the node's
// internal state in/out so a test can put the product into the same state.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <map>
#include <set>
#include <list>
#include <deque>
#include <stack>
#include <string>
#include <vector>
#include <memory>
#include <random>
#include <sstream>
#include <iostream>
#include <algorithm>
#define private public
#define protected public
#include "img_env.h"
#include <ervo_ros/Agent.h>
#include <ervo_ros/Obstacle.h>
#include <ervo_ros/KdTree.h>
#include <pedsimros/ped_tree.h>
#undef private
#undef protected

namespace {
struct Ref {
    EnvService svc;
    comn_pkg::InitEnv::Request init_req;
    comn_pkg::ResetEnv::Request reset_req;
    comn_pkg::StepEnv::Request step_req;
    std::vector<comn_pkg::AgentState> states;
    int R = 0, P = 0;
    std::string scene;
};
const char* kShape[] = {"circle", "rectangle", "leg"};

void fill_limiter(comn_pkg::SpeedLimiter& l, const double* d) {
    l.has_velocity_limits = d[0] != 0; l.has_acceleration_limits = d[1] != 0; l.has_jerk_limits = d[2] != 0;
    l.min_velocity = (float)d[3]; l.max_velocity = (float)d[4];
    l.min_acceleration = (float)d[5]; l.max_acceleration = (float)d[6];
    l.min_jerk = (float)d[7]; l.max_jerk = (float)d[8];
}
void fill_pose(geometry_msgs::Pose& p, const double* d) {  // x, y, qx, qy, qz, qw
    p.position.x = d[0]; p.position.y = d[1]; p.position.z = 0;
    p.orientation.x = d[2]; p.orientation.y = d[3]; p.orientation.z = d[4]; p.orientation.w = d[5];
}
RVO::RVOSimulator* rvo_of(Ref* r) {
    if (r->scene == "rvoscene") return static_cast<RVOScene*>(r->svc.ImgEnv_env.pedscene.get())->rvo_sim;
    if (r->scene == "ervoscene") return static_cast<ERVOScene*>(r->svc.ImgEnv_env.pedscene.get())->rvo_sim;
    return nullptr;
}
PedScene* sfm_of(Ref* r) {
    return r->scene == "pedscene" ? static_cast<PedScene*>(r->svc.ImgEnv_env.pedscene.get()) : nullptr;
}
}  // namespace

extern "C" {

void* ref_create() { return new Ref(); }
void ref_destroy(void* h) { delete static_cast<Ref*>(h); }

// scalars[20] in InitEnv.srv field order:
// view_resolution, view_width, view_height, step_hz, state_dim, is_show_gui(ignored->0), sleep_t(->0),
// window_height, show_image_height, is_draw_step, step_draw, use_laser, range_total, view_angle_begin,
// view_angle_end, view_min_dist, view_max_dist, beep_r, ped_ca_p, relation_ped_robo
// robot_desc[R][25]: shape, size[4], sensor_cfg[2], limiter_v[9], limiter_w[9]
// ped_desc[P][8]:   shape, size[6], max_speed
int ref_init(void* h, const double* sc, double global_resolution, const uint8_t* grid, int H, int W, int raw_h, int raw_w,
             int R, const double* robot_desc, const char* robot_ktype,
             int P, const double* ped_desc, const char* scene_type) {
    Ref* r = static_cast<Ref*>(h);
    comn_pkg::InitEnv::Request& q = r->init_req;
    q = comn_pkg::InitEnv::Request();
    q.view_resolution = (float)sc[0]; q.view_width = (float)sc[1]; q.view_height = (float)sc[2];
    q.step_hz = (float)sc[3]; q.state_dim = (int)sc[4]; q.is_show_gui = 0; q.sleep_t = 0;
    q.window_height = (uint32_t)sc[7]; q.show_image_height = (uint32_t)sc[8];
    q.is_draw_step = sc[9] != 0; q.step_draw = (uint32_t)sc[10];
    q.use_laser = sc[11] != 0; q.range_total = (uint32_t)sc[12];
    q.view_angle_begin = (float)sc[13]; q.view_angle_end = (float)sc[14];
    q.view_min_dist = (float)sc[15]; q.view_max_dist = (float)sc[16];
    q.beep_r = (float)sc[17]; q.ped_ca_p = (float)sc[18]; q.relation_ped_robo = (uint32_t)sc[19];
    q.env.map_file = "<host-decoded grid>";
    q.env.name = "oracle";
    q.env.global_resolution = (float)global_resolution;
    r->R = R; r->P = P; r->scene = scene_type;
    for (int i = 0; i < R; i++) {
        const double* d = robot_desc + 25 * i;
        comn_pkg::Agent a;
        a.ktype = robot_ktype; a.shape = kShape[(int)d[0]];
        int ns = a.shape == "circle" ? 3 : 4;
        for (int k = 0; k < ns; k++) a.size.push_back((float)d[1 + k]);
        a.sensor_cfg = {(float)d[5], (float)d[6]};
        fill_limiter(a.speed_limiter_v, d + 7);
        fill_limiter(a.speed_limiter_w, d + 16);
        a.name = "cool_robot" + std::to_string(i);
        q.env.robots.push_back(a);
    }
    for (int i = 0; i < P; i++) {
        const double* d = ped_desc + 8 * i;
        comn_pkg::Agent a;
        a.ktype = scene_type; a.shape = kShape[(int)d[0]];
        int ns = a.shape == "leg" ? 6 : (a.shape == "circle" ? 3 : 4);
        for (int k = 0; k < ns; k++) a.size.push_back((float)d[1 + k]);
        a.max_speed = (float)d[7];
        a.name = "cool_ped" + std::to_string(i);
        q.env.peds.push_back(a);
    }
    q.env.ped_scene_type = P > 0 ? scene_type : "";
    // The host decoded + resized the PNG with cv2 (as grid_map.cpp:30-36 would with OpenCV);
    // read_image() still computes W,H itself from the PNG size and both resolutions.
    cvshim::set_next_image(grid, H, W, raw_h, raw_w);
    comn_pkg::InitEnv::Response res;
    try { r->svc.init_env(q, res); } catch (std::exception& e) { fprintf(stderr, "ref_init: %s\n", e.what()); return -1; }
    GridMap& sm = r->svc.ImgEnv_env.EnvMap_maps_.static_map_;
    if (sm.img_height_ != H || sm.img_width_ != W) { fprintf(stderr, "ref_init: grid size mismatch %d %d\n", sm.img_height_, sm.img_width_); return -2; }
    return 0;
}

// obs[n][11]: shape, size[4], x, y, qx, qy, qz, qw ; robots[R][8]: x,y,q[4],gx,gy
// peds[P][8]: x,y,q[4],gx,gy ; traj_len[P]; traj[sum][3]; trajv_len[P] (may be null); trajv[sum][3]
int ref_reset(void* h, int n_obs, const double* obs, const double* robots, const double* peds,
              const int* traj_len, const double* traj, const int* trajv_len, const double* trajv, int ignore_obstacle) {
    Ref* r = static_cast<Ref*>(h);
    comn_pkg::ResetEnv::Request& q = r->reset_req;
    q = comn_pkg::ResetEnv::Request();
    q.ignore_obstacle = ignore_obstacle != 0;
    for (int i = 0; i < n_obs; i++) {
        const double* d = obs + 11 * i;
        comn_pkg::Agent a; a.name = "obstacle"; a.ktype = "obs"; a.shape = kShape[(int)d[0]];
        int ns = a.shape == "circle" ? 3 : 4;
        for (int k = 0; k < ns; k++) a.size.push_back((float)d[1 + k]);
        fill_pose(a.init_pose, d + 5);
        q.obstacles.push_back(a);
    }
    for (int i = 0; i < r->R; i++) {
        const double* d = robots + 8 * i;
        comn_pkg::Agent a = r->init_req.env.robots[i];
        fill_pose(a.init_pose, d); a.goal.x = d[6]; a.goal.y = d[7];
        q.robots.push_back(a);
    }
    size_t to = 0, tvo = 0;
    for (int i = 0; i < r->P; i++) {
        const double* d = peds + 8 * i;
        comn_pkg::Agent a = r->init_req.env.peds[i];
        fill_pose(a.init_pose, d); a.goal.x = d[6]; a.goal.y = d[7];
        for (int k = 0; k < traj_len[i]; k++, to++) { geometry_msgs::Point p; p.x = traj[3 * to]; p.y = traj[3 * to + 1]; p.z = traj[3 * to + 2]; a.trajectory.push_back(p); }
        if (trajv_len) for (int k = 0; k < trajv_len[i]; k++, tvo++) { geometry_msgs::Point p; p.x = trajv[3 * tvo]; p.y = trajv[3 * tvo + 1]; p.z = trajv[3 * tvo + 2]; a.trajectory_v.push_back(p); }
        q.peds.push_back(a);
    }
    r->step_req.robots = q.robots;  // yaml_env.py:232
    comn_pkg::ResetEnv::Response res;
    try { r->svc.reset_env(q, res); } catch (std::exception& e) { fprintf(stderr, "ref_reset: %s\n", e.what()); return -1; }
    r->states = res.robot_states;
    // EpRes logging vectors grow without bound in the node; irrelevant here.
    return 0;
}

// actions[R][3] = v, w, v_y(beep) as float32 (Agent.msg) ; alive[R]
int ref_step(void* h, const float* actions, const uint8_t* alive) {
    Ref* r = static_cast<Ref*>(h);
    for (int i = 0; i < r->R; i++) {
        comn_pkg::Agent& a = r->step_req.robots[i];
        a.alive = alive[i] != 0;
        a.v = alive[i] ? actions[3 * i] : 0.f; a.w = alive[i] ? actions[3 * i + 1] : 0.f; a.v_y = alive[i] ? actions[3 * i + 2] : 0.f;
    }
    comn_pkg::StepEnv::Response res;
    try { r->svc.step_env(r->step_req, res); } catch (std::exception& e) { fprintf(stderr, "ref_step: %s\n", e.what()); return -1; }
    r->states = res.robot_states;
    // keep the node's per-step EpRes log from growing during long CPU-baseline runs
    ImgEnv& e = r->svc.ImgEnv_env;
    for (auto& rr : e.eps_res_msg.robots_res) { rr.poses.clear(); rr.vs.clear(); rr.ws.clear(); }
    for (auto& pr : e.eps_res_msg.peds_res) { pr.poses.clear(); pr.vs.clear(); pr.v_ys.clear(); }
    return 0;
}

// Reply of the last reset/step, AgentState.msg fields the Python side reads (yaw_env.py:446-481).
void ref_get_states(void* h, uint8_t* view_maps, float* state, float* laser, int8_t* is_collision, uint8_t* is_arrive,
                    float* pedinfo /*[R][P][5] px,py,vx,vy,r_*/) {
    Ref* r = static_cast<Ref*>(h);
    for (int i = 0; i < r->R; i++) {
        const comn_pkg::AgentState& s = r->states[i];
        if (view_maps) memcpy(view_maps + (size_t)i * s.view_map.data.size(), s.view_map.data.data(), s.view_map.data.size());
        if (state) for (size_t k = 0; k < s.state.size(); k++) state[i * s.state.size() + k] = s.state[k];
        if (laser) for (size_t k = 0; k < s.laser.size(); k++) laser[i * s.laser.size() + k] = s.laser[k];
        if (is_collision) is_collision[i] = s.is_collision;
        if (is_arrive) is_arrive[i] = s.is_arrive;
        if (pedinfo) for (int j = 0; j < r->P; j++) {
            float* o = pedinfo + ((size_t)i * r->P + j) * 5;
            o[0] = s.pedinfo[j].px; o[1] = s.pedinfo[j].py; o[2] = s.pedinfo[j].vx; o[3] = s.pedinfo[j].vy; o[4] = s.pedinfo[j].r_;
        }
    }
}
int ref_view_size(void* h) { Ref* r = static_cast<Ref*>(h); return r->states.empty() ? 0 : (int)r->states[0].view_map.data.size(); }
int ref_laser_size(void* h) { Ref* r = static_cast<Ref*>(h); return r->states.empty() ? 0 : (int)r->states[0].laser.size(); }

// ---- internal state in/out (Appendix B of SURVEY.md) ----
// robot[R][16]: x,y,yaw, gx,gy,gyaw, last0 v,w, last1 v,w, vx,vy, is_collision, is_arrive, beep, 0
// ped[P][20]:   x,y,yaw, lx,ly,lyaw, vx,vy, state,last_state,remaining, lleg xyz, rleg xyz, traj_idx, 0, 0
void ref_get_internal(void* h, double* robot, double* ped) {
    Ref* r = static_cast<Ref*>(h);
    ImgEnv& e = r->svc.ImgEnv_env;
    for (int i = 0; i < r->R; i++) {
        Agent& a = e.robots_[i]; double* o = robot + 16 * i;
        o[0] = a.robot_pose_.x; o[1] = a.robot_pose_.y; o[2] = a.robot_pose_.z;
        o[3] = a.target_pose_.x; o[4] = a.target_pose_.y; o[5] = a.target_pose_.z;
        o[6] = a.last0_vw_.x; o[7] = a.last0_vw_.y; o[8] = a.last1_vw_.x; o[9] = a.last1_vw_.y;
        o[10] = a.vx; o[11] = a.vy; o[12] = a.is_collision_; o[13] = a.is_arrive_; o[14] = a.beep; o[15] = 0;
    }
    for (int i = 0; i < r->P; i++) {
        PedAgent& a = e.peds_[i]; double* o = ped + 20 * i;
        o[0] = a.robot_pose_.x; o[1] = a.robot_pose_.y; o[2] = a.robot_pose_.z;
        o[3] = a.last_robot_pose_.x; o[4] = a.last_robot_pose_.y; o[5] = a.last_robot_pose_.z;
        o[6] = a.PedAgent::vx; o[7] = a.PedAgent::vy; o[8] = a.state_; o[9] = a.last_state_; o[10] = a.remaining_dist_;
        o[11] = a.left_leg_.x; o[12] = a.left_leg_.y; o[13] = a.left_leg_.z;
        o[14] = a.right_leg_.x; o[15] = a.right_leg_.y; o[16] = a.right_leg_.z;
        o[17] = a.cur_traj_index_; o[18] = 0; o[19] = 0;
    }
}
// (min_jerk, max_jerk) of the linear and angular limiter as the node's robots hold them: SpeedLimiter(msg) never assigns
// min_jerk (speed_limit.cpp:56-65), so it is whatever the stack held.  out[R][4] = lin min, lin max, ang min, ang max.
void ref_get_jerk_limits(void* h, double* out) {
    Ref* r = static_cast<Ref*>(h);
    ImgEnv& e = r->svc.ImgEnv_env;
    for (int i = 0; i < r->R; i++) {
        Agent& a = e.robots_[i];
        out[4 * i] = a.limiter_lin_.min_jerk; out[4 * i + 1] = a.limiter_lin_.max_jerk;
        out[4 * i + 2] = a.limiter_ang_.min_jerk; out[4 * i + 3] = a.limiter_ang_.max_jerk;
    }
}
void ref_set_internal(void* h, const double* robot, const double* ped) {
    Ref* r = static_cast<Ref*>(h);
    ImgEnv& e = r->svc.ImgEnv_env;
    for (int i = 0; i < r->R && robot; i++) {
        Agent& a = e.robots_[i]; const double* o = robot + 16 * i;
        a.robot_pose_ = Point3d(o[0], o[1], o[2]);
        a.target_pose_ = Point3d(o[3], o[4], o[5]);
        {   // what set_goal() (agent.cpp:144-154) caches
            a.tf_target_world_.setOrigin(tf::Vector3(a.target_pose_.x, a.target_pose_.y, 0));
            tf::Quaternion q; q.setRPY(0, 0, a.target_pose_.z);
            a.tf_target_world_.setRotation(q);
            a.tf_world_target_ = a.tf_target_world_.inverse();
        }
        a.last0_vw_ = Point2d(o[6], o[7]); a.last1_vw_ = Point2d(o[8], o[9]);
        a.vx = o[10]; a.vy = o[11]; a.is_collision_ = (int)o[12]; a.is_arrive_ = o[13] != 0; a.beep = (int)o[14];
    }
    for (int i = 0; i < r->P && ped; i++) {
        PedAgent& a = e.peds_[i]; const double* o = ped + 20 * i;
        a.robot_pose_ = Point3d(o[0], o[1], o[2]); a.last_robot_pose_ = Point3d(o[3], o[4], o[5]);
        a.PedAgent::vx = o[6]; a.PedAgent::vy = o[7]; a.state_ = (int)o[8]; a.last_state_ = (int)o[9]; a.remaining_dist_ = o[10];
        a.left_leg_ = Point3d(o[11], o[12], o[13]); a.right_leg_ = Point3d(o[14], o[15], o[16]);
        a.cur_traj_index_ = (int)o[17];
    }
}

// ORCA / ERVO solver state: agents = peds then robots (if relation_ped_robo==1). a[n][4] = px,py,vx,vy (float32)
int ref_rvo_num_agents(void* h) { Ref* r = static_cast<Ref*>(h); RVO::RVOSimulator* s = rvo_of(r); return s ? (int)s->agents_.size() : 0; }
void ref_rvo_get(void* h, float* a) {
    Ref* r = static_cast<Ref*>(h); RVO::RVOSimulator* s = rvo_of(r); if (!s) return;
    for (size_t i = 0; i < s->agents_.size(); i++) {
        a[4 * i] = s->agents_[i]->position_.x(); a[4 * i + 1] = s->agents_[i]->position_.y();
        a[4 * i + 2] = s->agents_[i]->velocity_.x(); a[4 * i + 3] = s->agents_[i]->velocity_.y();
    }
}
void ref_rvo_set(void* h, const float* a) {
    Ref* r = static_cast<Ref*>(h); RVO::RVOSimulator* s = rvo_of(r); if (!s) return;
    for (size_t i = 0; i < s->agents_.size(); i++) {
        s->agents_[i]->position_ = RVO::Vector2(a[4 * i], a[4 * i + 1]);
        s->agents_[i]->velocity_ = RVO::Vector2(a[4 * i + 2], a[4 * i + 3]);
    }
}
// Obstacle vertex ring after processObstacles (incl. k-d split vertices): v[n][8] = px,py,dirx,diry,convex,next,prev,0
int ref_rvo_num_obstacles(void* h) { Ref* r = static_cast<Ref*>(h); RVO::RVOSimulator* s = rvo_of(r); return s ? (int)s->obstacles_.size() : 0; }
void ref_rvo_get_obstacles(void* h, float* v) {
    Ref* r = static_cast<Ref*>(h); RVO::RVOSimulator* s = rvo_of(r); if (!s) return;
    for (size_t i = 0; i < s->obstacles_.size(); i++) {
        RVO::Obstacle* o = s->obstacles_[i]; float* d = v + 8 * i;
        d[0] = o->point_.x(); d[1] = o->point_.y(); d[2] = o->unitDir_.x(); d[3] = o->unitDir_.y();
        d[4] = o->isConvex_; d[5] = (float)o->nextObstacle_->id_; d[6] = (float)o->prevObstacle_->id_; d[7] = 0;
    }
}

// SFM (libpedsim) state: agents = peds then robots. a[n][12] =
// p.xyz, v.xyz, vmax, dest(-1 none, else index into the agent's ORIGINAL waypoint list), lastdest, deque_rot, in_tree, 0
int ref_sfm_num_agents(void* h) { Ref* r = static_cast<Ref*>(h); PedScene* s = sfm_of(r); return s ? (int)(s->peds_sim_.size() + s->robots_sim_.size()) : 0; }
static Ped::Tagent* sfm_agent(PedScene* s, int i) {
    return i < (int)s->peds_sim_.size() ? s->peds_sim_[i] : s->robots_sim_[i - s->peds_sim_.size()];
}
static bool sfm_in_tree(Ped::Ttree* t, const Ped::Tagent* a) {
    if (t->isleaf) return t->agents.count(a) > 0;
    return sfm_in_tree(t->tree1, a) || sfm_in_tree(t->tree2, a) || sfm_in_tree(t->tree3, a) || sfm_in_tree(t->tree4, a);
}
void ref_sfm_get(void* h, double* a) {
    Ref* r = static_cast<Ref*>(h); PedScene* s = sfm_of(r); if (!s) return;
    int n = ref_sfm_num_agents(h);
    for (int i = 0; i < n; i++) {
        Ped::Tagent* t = sfm_agent(s, i); double* o = a + 12 * i;
        o[0] = t->p.x; o[1] = t->p.y; o[2] = t->p.z; o[3] = t->v.x; o[4] = t->v.y; o[5] = t->v.z; o[6] = t->vmax;
        // waypoints are identified by their coordinates' position in creation order: ids are global and
        // increase with creation, so within one agent the smallest id is waypoint 0 of the last setWayPoint.
        std::vector<int> ids; for (auto* w : t->waypoints) ids.push_back(w->getid());
        std::vector<int> sorted = ids; std::sort(sorted.begin(), sorted.end());
        auto idx_of = [&](Ped::Twaypoint* w) -> double {
            if (!w) return -1; for (size_t k = 0; k < sorted.size(); k++) if (sorted[k] == w->getid()) return (double)k; return -2; };
        o[7] = idx_of(t->destination); o[8] = idx_of(t->lastdestination);
        o[9] = ids.empty() ? 0 : (double)(std::find(sorted.begin(), sorted.end(), ids[0]) - sorted.begin());  // index of deque front
        o[10] = (r->svc.ImgEnv_env.relation_ped_robo == 1 || i < (int)s->peds_sim_.size()) ? (sfm_in_tree(s->pedscene_->tree, t) ? 1 : 0) : 0;
        o[11] = 0;
    }
}
void ref_sfm_set_pv(void* h, const double* a) {  // positions / velocities / vmax only
    Ref* r = static_cast<Ref*>(h); PedScene* s = sfm_of(r); if (!s) return;
    int n = ref_sfm_num_agents(h);
    for (int i = 0; i < n; i++) {
        Ped::Tagent* t = sfm_agent(s, i); const double* o = a + 12 * i;
        t->p.x = o[0]; t->p.y = o[1]; t->p.z = o[2]; t->v.x = o[3]; t->v.y = o[4]; t->v.z = o[5]; t->vmax = o[6];
    }
}
// Flattened quadtree: node[k][8] = x,y,w,h,isleaf,child1..4 packed as (first_child index, children are consecutive),n_agents,0
// plus per (node, agent) membership list. Returned sizes let the caller allocate.
static void sfm_tree_flat(PedScene* s, Ped::Ttree* t, std::vector<double>& nodes, std::vector<int>& member, int n_agents) {
    size_t me = nodes.size() / 8; nodes.resize(nodes.size() + 8);
    nodes[me * 8] = t->x; nodes[me * 8 + 1] = t->y; nodes[me * 8 + 2] = t->w; nodes[me * 8 + 3] = t->h; nodes[me * 8 + 4] = t->isleaf;
    nodes[me * 8 + 6] = (double)t->agents.size();
    if (t->isleaf) {
        for (int i = 0; i < n_agents; i++) if (t->agents.count(sfm_agent(s, i))) { member.push_back((int)me); member.push_back(i); }
        nodes[me * 8 + 5] = -1;
    } else {
        // children stored in order tree1..tree4, each subtree contiguous; record index of each child
        std::vector<double> idx;
        Ped::Ttree* c[4] = {t->tree1, t->tree2, t->tree3, t->tree4};
        double first = -1;
        for (int k = 0; k < 4; k++) { if (k == 0) first = (double)(nodes.size() / 8); sfm_tree_flat(s, c[k], nodes, member, n_agents); }
        nodes[me * 8 + 5] = first;
    }
}
int ref_sfm_tree(void* h, double* nodes, int max_nodes, int* member, int max_member, int* hash_node /*[n_agents] leaf index of treehash*/) {
    Ref* r = static_cast<Ref*>(h); PedScene* s = sfm_of(r); if (!s) return 0;
    int n = ref_sfm_num_agents(h);
    std::vector<double> nd; std::vector<int> mb;
    sfm_tree_flat(s, s->pedscene_->tree, nd, mb, n);
    int nn = (int)nd.size() / 8;
    if (nodes && nn <= max_nodes) memcpy(nodes, nd.data(), nd.size() * sizeof(double));
    if (member && (int)mb.size() / 2 <= max_member) memcpy(member, mb.data(), mb.size() * sizeof(int));
    (void)hash_node;
    return nn | ((int)(mb.size() / 2) << 16);
}

// Quadtree in the flat format of img_env_b200/csrc/sfmtree.cuh: nodes[n][5] = x,y,w,h,child0 (children tree1..tree4
// get consecutive ids), leaf[NA][4] = leaves holding agent a (-1 empty), hash[NA] = treehash[a]. Returns n.
int ref_sfm_tree2(void* h, double* nodes, int max_nodes, int* leaf, int* hash) {
    Ref* r = static_cast<Ref*>(h); PedScene* s = sfm_of(r); if (!s) return 0;
    int na = r->svc.ImgEnv_env.relation_ped_robo == 1 ? ref_sfm_num_agents(h) : (int)s->peds_sim_.size();
    std::vector<Ped::Ttree*> order; std::map<Ped::Ttree*, int> id;
    order.push_back(s->pedscene_->tree); id[s->pedscene_->tree] = 0;
    for (size_t k = 0; k < order.size(); k++) {
        Ped::Ttree* t = order[k];
        if (!t->isleaf) { Ped::Ttree* c[4] = {t->tree1, t->tree2, t->tree3, t->tree4}; for (int q = 0; q < 4; q++) { id[c[q]] = (int)order.size(); order.push_back(c[q]); } }
    }
    if ((int)order.size() > max_nodes) return -1;
    for (size_t k = 0; k < order.size(); k++) {
        Ped::Ttree* t = order[k]; double* o = nodes + 5 * k;
        o[0] = t->x; o[1] = t->y; o[2] = t->w; o[3] = t->h; o[4] = t->isleaf ? -1 : id[t->tree1];
    }
    for (int a = 0; a < na; a++) {
        Ped::Tagent* ag = sfm_agent(s, a);
        for (int q = 0; q < 4; q++) leaf[4 * a + q] = -1;
        int n = 0;
        for (size_t k = 0; k < order.size(); k++) if (order[k]->isleaf && order[k]->agents.count(ag) && n < 4) leaf[4 * a + n++] = (int)k;
        auto it = s->pedscene_->treehash.find(ag);
        hash[a] = it == s->pedscene_->treehash.end() ? -1 : id[it->second];
    }
    return (int)order.size();
}

// obs_map_ / peds_map_ of the node (for raster-level checks)
void ref_get_map(void* h, int which, uint8_t* out) {
    Ref* r = static_cast<Ref*>(h);
    EnvMap& m = r->svc.ImgEnv_env.EnvMap_maps_;
    GridMap& g = which == 0 ? m.static_map_ : (which == 1 ? m.obs_map_ : m.peds_map_);
    memcpy(out, g.map_.data, (size_t)g.img_height_ * g.img_width_);
}
int ref_bbox_size(void* h, int robot) { Ref* r = static_cast<Ref*>(h); return (int)r->svc.ImgEnv_env.robots_[robot].bbox_.size(); }

}  // extern "C"
