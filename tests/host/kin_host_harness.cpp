// Test harness: the robot kinematics stage (img_env_b200/csrc/kin.cuh: SpeedLimiter::limit, Agent::cmd) called on the HOST
// -- the functions are __host__ __device__ -- for tests/test_kin_host_cpu.py.
// Input (stdin): ktype step_hz control_hz | 2 x (has_v has_a has_j min_v max_v min_a max_a min_j max_j) | n | n x (x y yaw gx gy
// l0v l0w l1v l1w vx vy  v w v_y).  Output: per record "x y yaw l0v l0w l1v l1w vx vy arrive" with 17 significant digits.
#include <cmath>
#include <cstdint>
#include <cstdio>
#define __forceinline__ inline
struct int4 { int x, y, z, w; };
#include "kin.cuh"

static int read_limiter(Limiter& L) {
    double h[3];
    if (scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf", &h[0], &h[1], &h[2], &L.min_v, &L.max_v, &L.min_a, &L.max_a, &L.min_j, &L.max_j) != 9) return 1;
    L.has_v = h[0] != 0; L.has_a = h[1] != 0; L.has_j = h[2] != 0;
    return 0;
}
int main() {
    int ktype, n; double step_hz, control_hz;
    Limiter Lv, Lw;
    if (scanf("%d %lf %lf", &ktype, &step_hz, &control_hz) != 3 || read_limiter(Lv) || read_limiter(Lw) || scanf("%d", &n) != 1) return 1;
    for (int k = 0; k < n; k++) {
        RobotKin r; double v, w, vy;
        if (scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", &r.x, &r.y, &r.yaw, &r.gx, &r.gy, &r.l0v, &r.l0w, &r.l1v, &r.l1w,
                  &r.vx, &r.vy, &v, &w, &vy) != 14) return 1;
        robot_cmd(r, Lv, Lw, ktype, step_hz, control_hz, v, w, vy);
        printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d\n", r.x, r.y, r.yaw, r.l0v, r.l0w, r.l1v, r.l1w, r.vx, r.vy, (int)r.arrive);
    }
    return 0;
}
