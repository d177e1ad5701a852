// TEST INFRASTRUCTURE (oracle/_ref build only). Minimal stand-in for ROS tf's
// LinearMath (tf/LinearMath/{Vector3,Quaternion,Matrix3x3,Transform}.h, tf 1.13,
// bullet-derived, double precision). ROS is not installed in this image, so the
// arithmetic the reference calls is restated here operation-for-operation
// (same evaluation order, no FMA contraction). PARITY UNPINNED: no reference
// test pins this boundary.
#pragma once
#include <cmath>
#include <geometry_msgs/Point.h>
namespace tf {
typedef double tfScalar;
class Vector3 {
public:
    tfScalar m_floats[4];
    Vector3() { m_floats[0] = m_floats[1] = m_floats[2] = m_floats[3] = 0; }
    Vector3(tfScalar x, tfScalar y, tfScalar z) { m_floats[0] = x; m_floats[1] = y; m_floats[2] = z; m_floats[3] = 0; }
    tfScalar x() const { return m_floats[0]; }
    tfScalar y() const { return m_floats[1]; }
    tfScalar z() const { return m_floats[2]; }
    tfScalar getX() const { return m_floats[0]; }
    tfScalar getY() const { return m_floats[1]; }
    tfScalar getZ() const { return m_floats[2]; }
    tfScalar dot(const Vector3& v) const {
        return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2];
    }
    Vector3 operator-() const { return Vector3(-m_floats[0], -m_floats[1], -m_floats[2]); }
    tfScalar operator[](int i) const { return m_floats[i]; }
    tfScalar& operator[](int i) { return m_floats[i]; }
};
inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x() + b.x(), a.y() + b.y(), a.z() + b.z()); }

class Quaternion {
public:
    tfScalar m_floats[4];
    Quaternion() { m_floats[0] = m_floats[1] = m_floats[2] = m_floats[3] = 0; }
    Quaternion(tfScalar x, tfScalar y, tfScalar z, tfScalar w) { m_floats[0] = x; m_floats[1] = y; m_floats[2] = z; m_floats[3] = w; }
    void setValue(tfScalar x, tfScalar y, tfScalar z, tfScalar w) { m_floats[0] = x; m_floats[1] = y; m_floats[2] = z; m_floats[3] = w; }
    void setRPY(tfScalar roll, tfScalar pitch, tfScalar yaw) {
        tfScalar halfYaw = yaw * 0.5, halfPitch = pitch * 0.5, halfRoll = roll * 0.5;
        tfScalar cosYaw = cos(halfYaw), sinYaw = sin(halfYaw);
        tfScalar cosPitch = cos(halfPitch), sinPitch = sin(halfPitch);
        tfScalar cosRoll = cos(halfRoll), sinRoll = sin(halfRoll);
        setValue(sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw,
                 cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw,
                 cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw,
                 cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw);
    }
    tfScalar x() const { return m_floats[0]; }
    tfScalar y() const { return m_floats[1]; }
    tfScalar z() const { return m_floats[2]; }
    tfScalar w() const { return m_floats[3]; }
    tfScalar getX() const { return m_floats[0]; }
    tfScalar getY() const { return m_floats[1]; }
    tfScalar getZ() const { return m_floats[2]; }
    tfScalar getW() const { return m_floats[3]; }
    tfScalar dot(const Quaternion& q) const {
        return m_floats[0] * q.x() + m_floats[1] * q.y() + m_floats[2] * q.z() + m_floats[3] * q.m_floats[3];
    }
    tfScalar length2() const { return dot(*this); }
};

class Matrix3x3 {
public:
    Vector3 m_el[3];
    Matrix3x3() {}
    explicit Matrix3x3(const Quaternion& q) { setRotation(q); }
    Matrix3x3(tfScalar xx, tfScalar xy, tfScalar xz, tfScalar yx, tfScalar yy, tfScalar yz, tfScalar zx, tfScalar zy, tfScalar zz) {
        setValue(xx, xy, xz, yx, yy, yz, zx, zy, zz);
    }
    void setValue(tfScalar xx, tfScalar xy, tfScalar xz, tfScalar yx, tfScalar yy, tfScalar yz, tfScalar zx, tfScalar zy, tfScalar zz) {
        m_el[0] = Vector3(xx, xy, xz); m_el[1] = Vector3(yx, yy, yz); m_el[2] = Vector3(zx, zy, zz);
    }
    void setIdentity() { setValue(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    const Vector3& operator[](int i) const { return m_el[i]; }
    Vector3& operator[](int i) { return m_el[i]; }
    void setRotation(const Quaternion& q) {
        tfScalar d = q.length2();
        tfScalar s = tfScalar(2.0) / d;
        tfScalar xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
        tfScalar wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
        tfScalar xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
        tfScalar yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
        setValue(tfScalar(1.0) - (yy + zz), xy - wz, xz + wy,
                 xy + wz, tfScalar(1.0) - (xx + zz), yz - wx,
                 xz - wy, yz + wx, tfScalar(1.0) - (xx + yy));
    }
    void getRotation(Quaternion& q) const {
        tfScalar trace = m_el[0].x() + m_el[1].y() + m_el[2].z();
        tfScalar temp[4];
        if (trace > tfScalar(0.0)) {
            tfScalar s = sqrt(trace + tfScalar(1.0));
            temp[3] = (s * tfScalar(0.5));
            s = tfScalar(0.5) / s;
            temp[0] = ((m_el[2].y() - m_el[1].z()) * s);
            temp[1] = ((m_el[0].z() - m_el[2].x()) * s);
            temp[2] = ((m_el[1].x() - m_el[0].y()) * s);
        } else {
            int i = m_el[0].x() < m_el[1].y() ? (m_el[1].y() < m_el[2].z() ? 2 : 1) : (m_el[0].x() < m_el[2].z() ? 2 : 0);
            int j = (i + 1) % 3;
            int k = (i + 2) % 3;
            tfScalar s = sqrt(m_el[i][i] - m_el[j][j] - m_el[k][k] + tfScalar(1.0));
            temp[i] = s * tfScalar(0.5);
            s = tfScalar(0.5) / s;
            temp[3] = (m_el[k][j] - m_el[j][k]) * s;
            temp[j] = (m_el[j][i] + m_el[i][j]) * s;
            temp[k] = (m_el[k][i] + m_el[i][k]) * s;
        }
        q.setValue(temp[0], temp[1], temp[2], temp[3]);
    }
    void getEulerYPR(tfScalar& yaw, tfScalar& pitch, tfScalar& roll, unsigned int solution_number = 1) const {
        struct Euler { tfScalar yaw, pitch, roll; } euler_out, euler_out2;
        if (fabs(m_el[2].x()) >= 1) {
            euler_out.yaw = 0; euler_out2.yaw = 0;
            if (m_el[2].x() < 0) {
                tfScalar delta = atan2(m_el[0].y(), m_el[0].z());
                euler_out.pitch = M_PI / tfScalar(2.0); euler_out2.pitch = M_PI / tfScalar(2.0);
                euler_out.roll = delta; euler_out2.roll = delta;
            } else {
                tfScalar delta = atan2(-m_el[0].y(), -m_el[0].z());
                euler_out.pitch = -M_PI / tfScalar(2.0); euler_out2.pitch = -M_PI / tfScalar(2.0);
                euler_out.roll = delta; euler_out2.roll = delta;
            }
        } else {
            euler_out.pitch = -asin(m_el[2].x());
            euler_out2.pitch = M_PI - euler_out.pitch;
            euler_out.roll = atan2(m_el[2].y() / cos(euler_out.pitch), m_el[2].z() / cos(euler_out.pitch));
            euler_out2.roll = atan2(m_el[2].y() / cos(euler_out2.pitch), m_el[2].z() / cos(euler_out2.pitch));
            euler_out.yaw = atan2(m_el[1].x() / cos(euler_out.pitch), m_el[0].x() / cos(euler_out.pitch));
            euler_out2.yaw = atan2(m_el[1].x() / cos(euler_out2.pitch), m_el[0].x() / cos(euler_out2.pitch));
        }
        if (solution_number == 1) { yaw = euler_out.yaw; pitch = euler_out.pitch; roll = euler_out.roll; }
        else { yaw = euler_out2.yaw; pitch = euler_out2.pitch; roll = euler_out2.roll; }
    }
    void getRPY(tfScalar& roll, tfScalar& pitch, tfScalar& yaw, unsigned int solution_number = 1) const {
        getEulerYPR(yaw, pitch, roll, solution_number);
    }
    Matrix3x3 transpose() const {
        return Matrix3x3(m_el[0].x(), m_el[1].x(), m_el[2].x(),
                         m_el[0].y(), m_el[1].y(), m_el[2].y(),
                         m_el[0].z(), m_el[1].z(), m_el[2].z());
    }
    tfScalar tdotx(const Vector3& v) const { return m_el[0].x() * v.x() + m_el[1].x() * v.y() + m_el[2].x() * v.z(); }
    tfScalar tdoty(const Vector3& v) const { return m_el[0].y() * v.x() + m_el[1].y() * v.y() + m_el[2].y() * v.z(); }
    tfScalar tdotz(const Vector3& v) const { return m_el[0].z() * v.x() + m_el[1].z() * v.y() + m_el[2].z() * v.z(); }
};
inline Vector3 operator*(const Matrix3x3& m, const Vector3& v) { return Vector3(m[0].dot(v), m[1].dot(v), m[2].dot(v)); }
inline Matrix3x3 operator*(const Matrix3x3& m1, const Matrix3x3& m2) {
    return Matrix3x3(m2.tdotx(m1[0]), m2.tdoty(m1[0]), m2.tdotz(m1[0]),
                     m2.tdotx(m1[1]), m2.tdoty(m1[1]), m2.tdotz(m1[1]),
                     m2.tdotx(m1[2]), m2.tdoty(m1[2]), m2.tdotz(m1[2]));
}

class Transform {
public:
    Matrix3x3 m_basis;
    Vector3 m_origin;
    Transform() {}
    Transform(const Matrix3x3& b, const Vector3& c) : m_basis(b), m_origin(c) {}
    void setOrigin(const Vector3& o) { m_origin = o; }
    void setRotation(const Quaternion& q) { m_basis.setRotation(q); }
    const Vector3& getOrigin() const { return m_origin; }
    Vector3& getOrigin() { return m_origin; }
    const Matrix3x3& getBasis() const { return m_basis; }
    Quaternion getRotation() const { Quaternion q; m_basis.getRotation(q); return q; }
    Vector3 operator()(const Vector3& x) const {
        return Vector3(m_basis[0].dot(x) + m_origin.x(), m_basis[1].dot(x) + m_origin.y(), m_basis[2].dot(x) + m_origin.z());
    }
    Vector3 operator*(const Vector3& x) const { return (*this)(x); }
    Transform operator*(const Transform& t) const { return Transform(m_basis * t.m_basis, (*this)(t.m_origin)); }
    Transform inverse() const {
        Matrix3x3 inv = m_basis.transpose();
        return Transform(inv, inv * -m_origin);
    }
};
inline geometry_msgs::Quaternion createQuaternionMsgFromYaw(double yaw) {
    Quaternion q; q.setRPY(0.0, 0.0, yaw);
    geometry_msgs::Quaternion m; m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w();
    return m;
}
}  // namespace tf
