// Rigid 2-D transforms with the exact operation order of ROS tf's LinearMath as the reference
// uses it (tf::Transform / Matrix3x3 / Quaternion with roll = pitch = 0).  Every product and sum
// is a separately rounded fp64 operation (the library is compiled with --fmad=false and the host
// side with -ffp-contract=off) so cell indices round(x/res) come out identical to the x86 node.
// Reference call sites: agent.cpp:84-88 (init_view_map), :118-131 (get_base_world/get_view_world),
// :92-116 (base2view/view2base/base2world), :156-184 (get_state).
#pragma once
#include <math.h>
#ifndef __CUDACC__
#define __host__
#define __device__
#endif
#define HD __host__ __device__ __forceinline__

struct Tf2 {
    double m00, m01, m10, m11, ox, oy;
};

// tf::Quaternion::setRPY(0,0,yaw) -> (0,0,sin(yaw/2),cos(yaw/2)); Matrix3x3::setRotation(q)
HD Tf2 tf_from_pose(double x, double y, double yaw) {
    double h = yaw * 0.5;
    double cz = cos(h), sz = sin(h);
    double d = (sz * sz) + (cz * cz);   // x*x + y*y + z*z + w*w with x = y = 0
    double s = 2.0 / d;
    double zs = sz * s;
    double wz = cz * zs;
    double zz = sz * zs;
    Tf2 t;
    t.m00 = 1.0 - zz; t.m01 = 0.0 - wz; t.m10 = wz; t.m11 = 1.0 - zz;
    t.ox = x; t.oy = y;
    return t;
}
// Transform * Vector3(vx, vy, 0): row.dot(v) + origin
HD void tf_apply(const Tf2& t, double vx, double vy, double& rx, double& ry) {
    rx = (t.m00 * vx + t.m01 * vy) + t.ox;
    ry = (t.m10 * vx + t.m11 * vy) + t.oy;
}
// basis only (origin zeroed), as img_env.cpp:575-576 does for velocities
HD void tf_rotate(const Tf2& t, double vx, double vy, double& rx, double& ry) {
    rx = (t.m00 * vx + t.m01 * vy) + 0.0;
    ry = (t.m10 * vx + t.m11 * vy) + 0.0;
}
// Transform * Transform: basis via Matrix3x3 operator* (tdotx/tdoty), origin = a(b.origin)
HD Tf2 tf_mul(const Tf2& a, const Tf2& b) {
    Tf2 c;
    c.m00 = b.m00 * a.m00 + b.m10 * a.m01;
    c.m01 = b.m01 * a.m00 + b.m11 * a.m01;
    c.m10 = b.m00 * a.m10 + b.m10 * a.m11;
    c.m11 = b.m01 * a.m10 + b.m11 * a.m11;
    tf_apply(a, b.ox, b.oy, c.ox, c.oy);
    return c;
}
// Transform::inverse(): (B^T, B^T * (-o))
HD Tf2 tf_inv(const Tf2& a) {
    Tf2 c;
    c.m00 = a.m00; c.m01 = a.m10; c.m10 = a.m01; c.m11 = a.m11;
    double nx = -a.ox, ny = -a.oy;
    c.ox = c.m00 * nx + c.m01 * ny;
    c.oy = c.m10 * nx + c.m11 * ny;
    return c;
}
// Transform::getRotation() (trace method) followed by Matrix3x3(q).getRPY -> yaw, for a planar
// rotation (agent.cpp:165-168).
HD double tf_yaw_of(const Tf2& t) {
    const double m22 = 1.0;
    double trace = t.m00 + t.m11 + m22;
    double qz, qw;
    if (trace > 0.0) {
        double s = sqrt(trace + 1.0);
        qw = s * 0.5;
        s = 0.5 / s;
        qz = (t.m10 - t.m01) * s;
    } else {
        int i = t.m00 < t.m11 ? (t.m11 < m22 ? 2 : 1) : (t.m00 < m22 ? 2 : 0);
        if (i == 2) {
            double s = sqrt(m22 - t.m00 - t.m11 + 1.0);
            qz = s * 0.5;
            s = 0.5 / s;
            qw = (t.m10 - t.m01) * s;
        } else {
            // i in {0,1}: the quaternion has only x or y and w = (m[k][j]-m[j][k])*s = 0 for a planar
            // matrix; yaw of such a (numerically degenerate, |yaw| = pi) rotation:
            qz = 0.0; qw = 0.0;
            double mii = i == 0 ? t.m00 : t.m11, mjj = i == 0 ? t.m11 : m22, mkk = i == 0 ? m22 : t.m00;
            double s = sqrt(mii - mjj - mkk + 1.0);
            (void)s;
            return atan2(t.m10, t.m00);
        }
    }
    double d = (qz * qz) + (qw * qw);
    double s = 2.0 / d;
    double zs = qz * s;
    double wz = qw * zs;
    double zz = qz * zs;
    double m10 = 0.0 + wz, m00 = 1.0 - (0.0 + zz);
    // pitch = -asin(0) -> cos(pitch) = 1
    return atan2(m10 / 1.0, m00 / 1.0);
}
// GridMap::world2map (grid_map.cpp:40-44): int(round(x / res)), round half away from zero
HD int world2cell(double x, double res) { return (int)round(x / res); }

// Same result as world2cell() without the fp64 division in the common case: t = x * (1/res) differs from x/res
// by a few ulp (< 1e-11 cells for |t| < 1e4), so whenever t is farther than 1e-9 from a rounding boundary
// rint(t) == round(x/res); otherwise fall back to the exact expression.
HD int world2cell_fast(double x, double res, double inv_res) {
    const double t = x * inv_res;
    const double r = rint(t);
    if (fabs(t - r) > 0.499999999) return (int)round(x / res);
    return (int)r;
}
