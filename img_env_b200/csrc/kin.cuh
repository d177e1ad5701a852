// Robot kinematics stage: SpeedLimiter::limit (speed_limit.cpp:92-173), Agent::cmd
// (agent.cpp:186-283), Agent::get_state (agent.cpp:156-184).  fp64 throughout: is_arrives must
// be bit-exact against the node (BASELINE.json north_star).
#pragma once
#include "state.cuh"

HD double lim_clamp(double x, double lo, double hi) { return fmin(fmax(lo, x), hi); }   // std::min(std::max(min,x),max)
HD int lim_sign(double x) { return x == 0 ? 0 : (int)(x / fabs(x)); }

// speed_limit.cpp:153-173
HD void limit_jerk(const Limiter& L, double& v, double v0, double v1, double dt) {
    if (L.has_j) {
        const double dv = v - v0;
        const double dv0 = v0 - v1;
        const double dt2 = 2. * dt * dt;
        const double da_min = L.min_j * dt2;
        const double da_max = L.max_j * dt2;
        const double da = lim_clamp(dv - dv0, da_min, da_max);
        v = v0 + dv0 + da;
    }
}
// speed_limit.cpp:115-151
HD void limit_acceleration(const Limiter& L, double& v, double v0, double dt) {
    const double tmp = v;
    if (L.has_a) {
        const int v_sign = lim_sign(v);
        const int v0_sign = lim_sign(v0);
        if (v_sign + v0_sign != 0) {
            const double dv_min = L.min_a * dt;
            const double dv_max = L.max_a * dt;
            double dv = v - v0;
            const int dv_sign = lim_sign(dv);
            if (dv_sign == v0_sign || dv_sign == v_sign)
                dv = dv_sign * lim_clamp(fabs(dv), dv_min, dv_max);
            else
                dv = dv_sign * fabs(lim_clamp(-fabs(dv), dv_min, dv_max));
            v = v0 + dv;
        } else {
            const double zero_dt = fabs(v0 / L.min_a);
            if (zero_dt >= dt)
                v = v0_sign * (fabs(v0) - fabs(L.min_a) * dt);
            else {
                const double v_dt = fabs(v / L.max_a);
                if (zero_dt + v_dt >= dt)
                    v = v_sign * fabs(L.max_a * (dt - zero_dt));
                else
                    v = tmp;
            }
        }
    }
}
// speed_limit.cpp:92-113
HD void limiter_apply(const Limiter& L, double& v, double v0, double v1, double dt) {
    limit_jerk(L, v, v0, v1, dt);
    limit_acceleration(L, v, v0, dt);
    if (L.has_v) v = lim_clamp(v, L.min_v, L.max_v);
}

struct RobotKin {
    double x, y, yaw, gx, gy, l0v, l0w, l1v, l1w, vx, vy;
    bool arrive;
};

// Agent::cmd, agent.cpp:186-283.  v, w, v_y arrive as float32 and are widened (Agent.msg).
HD void robot_cmd(RobotKin& r, const Limiter& Lv, const Limiter& Lw, int ktype, double step_hz, double control_hz,
                  double v, double w, double v_y) {
    limiter_apply(Lv, v, r.l0v, r.l1v, step_hz);
    limiter_apply(Lw, w, r.l0w, r.l1w, step_hz);
    r.l1v = r.l0v; r.l1w = r.l0w;
    r.l0v = v; r.l0w = w;
    bool is_arrive = false;
    double ox = r.x, oy = r.y, oz = r.yaw;
    double cur_control = 0;
    if (ktype == 0) {
        while (cur_control <= step_hz) {
            ox += v * control_hz * cos(oz);
            oy += v * control_hz * sin(oz);
            r.vx = v * cos(oz);
            r.vy = v * sin(oz);
            oz += w * control_hz;
            double cur_dist = sqrt((ox - r.gx) * (ox - r.gx) + (oy - r.gy) * (oy - r.gy));
            if (cur_dist <= 0.3) { is_arrive = true; break; }
            cur_control += control_hz;
        }
        double theta = r.yaw, dt = step_hz;
        if (w == 0) {
            r.x += v * dt * cos(theta);
            r.y += v * dt * sin(theta);
            r.yaw += w * dt;
        } else {
            double vw = v / w;
            r.x += -vw * sin(theta) + vw * sin(theta + w * dt);
            r.y += vw * cos(theta) - vw * cos(theta + w * dt);
            r.yaw += w * dt;
        }
    } else {
        while (cur_control <= step_hz) {
            double nx = ox + (v * control_hz * cos(oz) - v_y * control_hz * sin(oz));
            double ny = oy + (v * control_hz * sin(oz) + v_y * control_hz * cos(oz));
            ox = nx; oy = ny;
            oz += w * control_hz;
            double cur_dist = sqrt((ox - r.gx) * (ox - r.gx) + (oy - r.gy) * (oy - r.gy));
            if (cur_dist <= 0.3) { is_arrive = true; break; }
            cur_control += control_hz;
        }
        double theta = r.yaw, dt = step_hz;
        if (w == 0) {
            r.x += v * dt * cos(theta) - v_y * dt * sin(theta);
            r.y += v * dt * sin(theta) + v_y * dt * cos(theta);
            r.yaw += w * dt;
        } else {
            double vw = v / w;
            r.x += -vw * sin(theta) + vw * sin(theta + w * dt);
            r.y += vw * cos(theta) - vw * cos(theta + w * dt);
            double v_yw = v_y / w;
            r.x += -v_yw * cos(theta) + v_yw * cos(theta + w * dt);
            r.y += -v_yw * sin(theta) + v_yw * sin(theta + w * dt);
            r.yaw += w * dt;
        }
    }
    double cur_dist = sqrt((r.x - r.gx) * (r.x - r.gx) + (r.y - r.gy) * (r.y - r.gy));
    if (cur_dist <= 0.3) is_arrive = true;
    r.arrive = is_arrive;
}

// Agent::get_state, agent.cpp:156-184. out[0..dim) as doubles (the node then narrows to float32).
HD void robot_state_vec(double x, double y, double yaw, double gx, double gy, double gyaw, double l0v, double l0w,
                        int dim, double* out) {
    Tf2 target_world = tf_from_pose(gx, gy, gyaw);
    Tf2 world_target = tf_inv(target_world);
    Tf2 base_world = tf_from_pose(x, y, yaw);
    Tf2 target_base = tf_inv(tf_mul(world_target, base_world));
    out[0] = target_base.ox;
    out[1] = target_base.oy;
    if (dim == 3) out[2] = tf_yaw_of(target_base);
    else if (dim == 4) { out[2] = l0v; out[3] = l0w; }
    else if (dim == 5) { out[2] = tf_yaw_of(target_base); out[3] = l0v; out[4] = l0w; }
}
