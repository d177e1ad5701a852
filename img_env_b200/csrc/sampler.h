// Native episode sampler: the reference's EnvPos.reset (envs/utils/reset_helper.py:115-345)
// ported to C++ so that auto-resets of thousands of scenes do not run 120 us of Python each.
// It draws from a bit-exact re-implementation of CPython's `random` module (MT19937, random(), uniform(),
// gauss() with its cached second variate, randint() via getrandbits), so with the same seed it returns exactly
// the poses the reference's Python sampler returns (tests/test_sampler_cpu.py checks this draw for draw against
// img_env_b200.envs.reset_helper.EnvPos, which is golden-tested against the real reset_helper.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace sampler {

class PyRandom {   // CPython Modules/_randommodule.c + Lib/random.py
public:
    explicit PyRandom(uint64_t seed = 0) { this->seed(seed); }
    void seed(uint64_t a) {
        uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
        init_by_array(key, key[1] ? 2 : 1);
        has_gauss_ = false;
    }
    double random() {
        uint32_t a = genrand() >> 5, b = genrand() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double uniform(double a, double b) { return a + (b - a) * random(); }
    double gauss(double mu, double sigma) {
        double z;
        if (has_gauss_) { z = gauss_next_; has_gauss_ = false; }
        else {
            double x2pi = random() * (2.0 * 3.141592653589793);
            double g2rad = sqrt(-2.0 * log(1.0 - random()));
            z = cos(x2pi) * g2rad;
            gauss_next_ = sin(x2pi) * g2rad; has_gauss_ = true;
        }
        return mu + z * sigma;
    }
    int randint(int a, int b) {   // randrange(a, b+1) -> _randbelow_with_getrandbits
        uint32_t n = (uint32_t)(b - a + 1);
        int k = 0; for (uint32_t t = n; t; t >>= 1) k++;
        uint32_t r = genrand() >> (32 - k);
        while (r >= n) r = genrand() >> (32 - k);
        return a + (int)r;
    }
private:
    uint32_t mt_[624]; int idx_ = 625; bool has_gauss_ = false; double gauss_next_ = 0;
    void init_genrand(uint32_t s) { mt_[0] = s; for (int i = 1; i < 624; i++) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i; idx_ = 624; }
    void init_by_array(const uint32_t* key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = 624 > len ? 624 : len; k; k--) {
            mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= 624) { mt_[0] = mt_[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = 623; k; k--) {
            mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            i++;
            if (i >= 624) { mt_[0] = mt_[623]; i = 1; }
        }
        mt_[0] = 0x80000000u;
    }
    uint32_t genrand() {
        if (idx_ >= 624) {
            static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) { uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu); mt_[kk] = mt_[kk + 397] ^ (y >> 1) ^ mag01[y & 1u]; }
            for (; kk < 623; kk++) { uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu); mt_[kk] = mt_[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u]; }
            uint32_t y = (mt_[623] & 0x80000000u) | (mt_[0] & 0x7fffffffu); mt_[623] = mt_[396] ^ (y >> 1) ^ mag01[y & 1u];
            idx_ = 0;
        }
        uint32_t y = mt_[idx_++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
};

// pose-type bits (how the reference tests the yaml strings: == for 'fix'/'rand_angle', `in` for the rest)
enum { T_EQ_FIX = 1, T_EQ_RAND_ANGLE = 2, T_RANGE = 4, T_CIRCLE = 8, T_FIX = 16, T_MULTI = 32, T_VIEW = 64, T_PLUS = 128, T_CIRCLE_FIX = 256 };
#define SAMPLER_MAX_MULTI 8
struct PoseSpec { int type = 0; int n_multi = 0; int len = 0; double v[SAMPLER_MAX_MULTI][6]; };   // v[0] holds a plain pose / range
struct AgentSpec { PoseSpec begin, target; double msize = 0; };
struct ObjectSpec { int shape = 0, fixed = 0, pose_len = 0; double pose[6] = {0}, size_range[4] = {0}; };

struct Sampler {
    int R = 0, P = 0;
    double circle_lo = 0, circle_hi = 0, target_min_dist = 0;
    int go_back = 0;   // 0 yes, 1 no, 2 random
    std::vector<AgentSpec> agents;     // robots then peds
    std::vector<ObjectSpec> objects;
    std::vector<PyRandom> rng;         // one stream per scene
};

struct P3 { double x, y, a; bool set; };

inline void py_quat(double yaw, double* q) { double h = yaw / 2.0; q[0] = 0; q[1] = 0; q[2] = sin(h); q[3] = cos(h); }

// one scene: fills obs[n_obj][11], robots[R][8], peds[P][8], traj_len[P], traj[P][max_traj>=2][3]
inline void sample_scene(const Sampler& S, PyRandom& rnd, double* obs, double* robots, double* peds, int* traj_len, double* traj, int max_traj) {
    const int n = S.R + S.P;
    // ---- reset_obs (reset_helper.py:122-165)
    std::vector<double> obs_range;   // x, y, yaw, radius per object
    for (size_t i = 0; i < S.objects.size(); i++) {
        const ObjectSpec& o = S.objects[i];
        double radius = o.shape == 0 ? rnd.uniform(o.size_range[0], o.size_range[1]) : sqrt(o.size_range[0] * o.size_range[0] + o.size_range[2] * o.size_range[2]);
        double px, py, pa;
        if (o.fixed) { px = o.pose[0]; py = o.pose[1]; pa = o.pose_len == 2 ? 0.0 : o.pose[2]; }
        else {
            px = rnd.uniform(o.pose[0], o.pose[1]); py = rnd.uniform(o.pose[2], o.pose[3]);
            pa = o.pose_len == 4 ? rnd.uniform(-3.14, 3.14) : rnd.uniform(o.pose[4], o.pose[5]);
        }
        obs_range.push_back(px); obs_range.push_back(py); obs_range.push_back(pa); obs_range.push_back(radius);
        double* d = obs + 11 * i;
        d[0] = o.shape;
        if (o.shape == 0) { d[1] = 0; d[2] = 0; d[3] = radius; d[4] = 0; } else { for (int k = 0; k < 4; k++) d[1 + k] = o.size_range[k]; }
        d[5] = px; d[6] = py; py_quat(pa, d + 7);
    }
    auto free_obj = [&](double x, double y, double rad) {
        for (size_t i = 0; i < S.objects.size(); i++) {
            const double* p = &obs_range[4 * i];
            if (p[3] == 0.0) continue;
            if (sqrt((x - p[0]) * (x - p[0]) + (y - p[1]) * (y - p[1])) <= rad + p[3]) return false;
        }
        return true;
    };
    auto free_rp = [&](double x, double y, const std::vector<P3>& poses) {
        for (const P3& p : poses) { if (!p.set) continue; if (sqrt((x - p.x) * (x - p.x) + (y - p.y) * (y - p.y)) <= 1.0) return false; }
        return true;
    };
    auto random_pose = [&](const double* r, int len, double* out) {
        out[0] = rnd.uniform(r[0], r[1]); out[1] = rnd.uniform(r[2], r[3]);
        out[2] = len == 4 ? rnd.uniform(-3.14, 3.14) : rnd.uniform(r[4], r[5]);
    };
    for (;;) {   // EnvPos.reset: retry _reset_robot_ped until every pose is set
        std::vector<P3> init(n, P3{0, 0, 0, false}), target(n, P3{0, 0, 0, false});
        const double circle_range = rnd.uniform(S.circle_lo, S.circle_hi);
        for (int i = 0; i < n; i++) {
            const AgentSpec& a = S.agents[i];
            if (a.begin.type & T_EQ_FIX) init[i] = P3{a.begin.v[0][0], a.begin.v[0][1], a.begin.v[0][2], true};
            if (a.target.type & T_EQ_FIX) target[i] = P3{a.target.v[0][0], a.target.v[0][1], a.target.v[0][2], true};
            if (a.begin.type & T_EQ_RAND_ANGLE) init[i] = P3{a.begin.v[0][0], a.begin.v[0][1], rnd.uniform(a.begin.v[0][2], a.begin.v[0][3]), true};
            if (a.target.type & T_EQ_RAND_ANGLE) target[i] = P3{a.target.v[0][0], a.target.v[0][1], rnd.uniform(a.target.v[0][2], a.target.v[0][3]), true};
        }
        bool circle_ok = false;
        while (!circle_ok) {
            circle_ok = true;
            for (int i = 0; i < n; i++) {
                if (init[i].set && target[i].set) continue;
                const AgentSpec& a = S.agents[i];
                bool reset_init = true;
                while (reset_init) {
                    int goal_fail = 0, circle_fail = 0;
                    if (a.begin.type & T_RANGE) {
                        while (reset_init) {
                            double p[3];
                            const double* pr = a.begin.v[0]; int len = a.begin.len;
                            if (a.begin.type & T_CIRCLE) {
                                double ang = rnd.uniform(-3.14, 3.14);
                                if (a.begin.type & T_FIX) ang = -3.14 + (6.28 / n) * i;
                                p[0] = circle_range * cos(ang) + pr[0]; p[1] = circle_range * sin(ang) + pr[1]; p[2] = ang + 3.14;
                                p[0] += rnd.gauss(0, 0.5); p[1] += rnd.gauss(0, 0.5);
                            } else {
                                if (a.begin.type & T_MULTI) pr = a.begin.v[rnd.randint(0, a.begin.n_multi - 1)];
                                random_pose(pr, len, p);
                            }
                            if (free_rp(p[0], p[1], init) && free_obj(p[0], p[1], a.msize * 2)) { init[i] = P3{p[0], p[1], p[2], true}; reset_init = false; break; }
                            if (a.begin.type & T_CIRCLE) {
                                circle_fail++;
                                if (circle_fail > 50) {
                                    circle_ok = false;
                                    for (int j = 0; j < n; j++) if (S.agents[j].begin.type & T_CIRCLE) { init[j].set = false; target[j].set = false; }
                                }
                            }
                        }
                    }
                    if ((a.target.type & T_CIRCLE_FIX) && init[i].set) {
                        const double* pr = a.target.v[0]; double ang = init[i].a;
                        target[i] = P3{circle_range * cos(ang) + pr[0], circle_range * sin(ang) + pr[1], ang - 3.14, true};
                    }
                    if (a.target.type & T_RANGE) {
                        for (;;) {
                            double p[3] = {0, 0, 0};
                            const double* pr = a.target.v[0]; int len = a.target.len;
                            if ((a.target.type & T_CIRCLE) && init[i].set) {
                                double ang = init[i].a;
                                p[0] = circle_range * cos(ang) + pr[0]; p[1] = circle_range * sin(ang) + pr[1]; p[2] = ang - 3.14;
                                p[0] += rnd.gauss(0, 0.5); p[1] += rnd.gauss(0, 0.5);
                            }
                            if (a.target.type & T_MULTI) pr = a.target.v[rnd.randint(0, a.target.n_multi - 1)];
                            if (a.target.type & T_VIEW) {
                                if (!(a.target.type & T_PLUS)) {   // random_view (reset_helper.py:64-82)
                                    const double tv[4] = {2.5, 4.0, 2.5, 4.0};
                                    for (;;) {
                                        p[0] = rnd.uniform(init[i].x - tv[1], init[i].x + tv[1]); p[1] = rnd.uniform(init[i].y - tv[3], init[i].y + tv[3]);
                                        p[2] = rnd.uniform(-3.14, 3.14);
                                        if (init[i].x - tv[0] <= p[0] && p[0] <= init[i].x + tv[0] && init[i].y - tv[2] <= p[1] && p[1] <= init[i].y + tv[2]) continue;
                                        if (pr[0] <= p[0] && p[0] <= pr[1] && pr[2] <= p[1] && p[1] <= pr[3]) break;
                                    }
                                }
                            } else if (len == 4 || len == 6) random_pose(pr, len, p);
                            if ((init[i].x - p[0]) * (init[i].x - p[0]) + (init[i].y - p[1]) * (init[i].y - p[1]) > S.target_min_dist * S.target_min_dist &&
                                free_rp(p[0], p[1], target) && free_obj(p[0], p[1], a.msize * 2)) { target[i] = P3{p[0], p[1], p[2], true}; break; }
                            goal_fail++;
                            if (goal_fail > 50) { reset_init = true; break; }
                        }
                    }
                }
            }
        }
        bool ok = true;
        for (int i = 0; i < n; i++) ok = ok && init[i].set && target[i].set;
        if (!ok) continue;
        for (int i = 0; i < S.R; i++) {
            double* d = robots + 8 * i;
            d[0] = init[i].x; d[1] = init[i].y; py_quat(init[i].a, d + 2); d[6] = target[i].x; d[7] = target[i].y;
        }
        for (int k = 0; k < S.P; k++) {
            const int i = S.R + k;
            double* d = peds + 8 * k;
            d[0] = init[i].x; d[1] = init[i].y; py_quat(init[i].a, d + 2); d[6] = target[i].x; d[7] = target[i].y;
            double* t = traj + (size_t)k * max_traj * 3;
            t[0] = target[i].x; t[1] = target[i].y; t[2] = 0; traj_len[k] = 1;
            if (S.go_back == 0 || (S.go_back == 2 && rnd.random() > 0.5)) { t[3] = init[i].x; t[4] = init[i].y; t[5] = 0; traj_len[k] = 2; }
        }
        return;
    }
}

}  // namespace sampler
