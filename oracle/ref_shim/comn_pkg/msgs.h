// TEST INFRASTRUCTURE shim: plain-struct equivalents of the generated comn_pkg
// message/service headers (schemas: src/comn_pkg/msg/*.msg, srv/*.srv). Field
// names and C++ types match what genmsg would emit (float32 -> float, etc.), so
// the float32 wire rounding of the reference is preserved.
#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include <geometry_msgs/Pose.h>
#include <sensor_msgs/Image.h>
namespace comn_pkg {
struct SpeedLimiter {
    uint8_t has_velocity_limits = 0, has_acceleration_limits = 0, has_jerk_limits = 0;
    float min_velocity = 0, max_velocity = 0, min_acceleration = 0, max_acceleration = 0, min_jerk = 0, max_jerk = 0;
};
struct Agent {
    std::string name, ktype;
    geometry_msgs::Pose init_pose;
    geometry_msgs::Point goal;
    std::string shape;
    std::vector<float> size;
    uint8_t alive = 0;
    float v = 0, w = 0, v_y = 0, max_speed = 0;
    std::string env_name;
    std::vector<float> sensor_cfg;
    std::vector<geometry_msgs::Point> trajectory, trajectory_v;
    SpeedLimiter speed_limiter_v, speed_limiter_w;
};
struct PedInfo { float px = 0, py = 0, vx = 0, vy = 0, r_ = 0, d_ = 0, co_r = 0, goal_x = 0, goal_y = 0, v_pref = 0, theta = 0; };
struct AgentState {
    sensor_msgs::Image view_map;
    std::vector<float> state, laser, hits_x, hits_y, angular_map;
    int8_t is_collision = 0;
    uint8_t is_arrive = 0;
    std::vector<PedInfo> pedinfo;
};
struct Env {
    std::string name, map_file;
    float global_resolution = 0;
    std::vector<Agent> robots, obstacles, peds;
    uint32_t env_id = 0;
    std::string env_name, ped_scene_type;
};
struct RobotRes {
    Agent info; std::string result;
    std::vector<geometry_msgs::Pose> poses;
    std::vector<float> vs, ws, v_ys;
};
struct EpRes {
    sensor_msgs::Image obs_map, ped_map;
    float resolution = 0; std::string env_name; float step_hz = 0;
    std::vector<RobotRes> robots_res, peds_res;
};
struct EnvsInfo {};
struct InitEnvRequest {
    float view_resolution = 0, view_width = 0, view_height = 0, step_hz = 0;
    int32_t state_dim = 0; uint8_t is_show_gui = 0; float sleep_t = 0;
    uint32_t window_height = 0, show_image_height = 0; uint8_t is_draw_step = 0; uint32_t step_draw = 0;
    uint8_t use_laser = 0; uint32_t range_total = 0;
    float view_angle_begin = 0, view_angle_end = 0, view_min_dist = 0, view_max_dist = 0, beep_r = 0, ped_ca_p = 0;
    uint32_t relation_ped_robo = 0;
    Env env;
};
struct InitEnvResponse {};
struct InitEnv { typedef InitEnvRequest Request; typedef InitEnvResponse Response; };
struct ResetEnvRequest { std::vector<Agent> obstacles, robots, peds; uint8_t is_test = 0; uint32_t env_id = 0; uint8_t ignore_obstacle = 0; };
struct ResetEnvResponse { std::vector<AgentState> robot_states; };
struct ResetEnv { typedef ResetEnvRequest Request; typedef ResetEnvResponse Response; };
struct StepEnvRequest { std::vector<Agent> robots; uint32_t env_id = 0; uint8_t is_test = 0; };
struct StepEnvResponse { std::vector<AgentState> robot_states; };
struct StepEnv { typedef StepEnvRequest Request; typedef StepEnvResponse Response; };
struct EndEpRequest { std::vector<std::string> robot_res; uint32_t env_id = 0; };
struct EndEpResponse {};
struct EndEp { typedef EndEpRequest Request; typedef EndEpResponse Response; };
}
