// TEST INFRASTRUCTURE — CPU restatement ("port") of the img_env per-step hot path.
// Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
// load liboracle_port.so. It is a plain, scalar, one-scene restatement that follows the reference
// line by line (full-map clones per robot, per-pixel atan2, cell-by-cell Bresenham ...), i.e. none
// of the restructuring the CUDA path does.  Each function cites the reference lines it follows.
// Pinned against the real reference: tests/test_oracle_cpu.py checks it against golden vectors
// generated from oracle/_ref (the unmodified node) and, when that library is present, step by step.
// Third-party arithmetic restated here because it is not under /root/reference: ROS tf LinearMath
// (tf 1.13, bullet-derived doubles) -> struct T2 below.  PARITY UNPINNED for that boundary only in the
// sense that no reference *test* pins it; it is pinned against oracle/_ref's shim of the same formulas.
//
// Exposes the same C ABI as oracle/ref_driver.cpp with the prefix port_.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <limits>
#include <string>
#include <vector>

namespace {

// ---------------- tf (Transform with roll = pitch = 0) ----------------
struct T2 { double a, b, c, d, ox, oy; };   // [[a b][c d]] + origin
T2 t2_pose(double x, double y, double yaw) {          // Quaternion::setRPY + Matrix3x3::setRotation
    double hy = yaw * 0.5, cz = cos(hy), sz = sin(hy);
    double len2 = sz * sz + cz * cz, s = 2.0 / len2;
    double zs = sz * s, wz = cz * zs, zz = sz * zs;
    return T2{1.0 - zz, 0.0 - wz, wz, 1.0 - zz, x, y};
}
void t2_map(const T2& t, double x, double y, double& rx, double& ry) {   // Transform * Vector3
    rx = (t.a * x + t.b * y) + t.ox; ry = (t.c * x + t.d * y) + t.oy;
}
T2 t2_mul(const T2& p, const T2& q) {                 // Transform * Transform
    T2 r; r.a = q.a * p.a + q.c * p.b; r.b = q.b * p.a + q.d * p.b; r.c = q.a * p.c + q.c * p.d; r.d = q.b * p.c + q.d * p.d;
    t2_map(p, q.ox, q.oy, r.ox, r.oy); return r;
}
T2 t2_inv(const T2& p) {                               // Transform::inverse
    T2 r; r.a = p.a; r.b = p.c; r.c = p.b; r.d = p.d;
    double nx = -p.ox, ny = -p.oy; r.ox = r.a * nx + r.b * ny; r.oy = r.c * nx + r.d * ny; return r;
}
double t2_yaw(const T2& t) {                           // getRotation() then Matrix3x3(q).getRPY
    double trace = t.a + t.d + 1.0, qz, qw;
    if (trace > 0.0) { double s = sqrt(trace + 1.0); qw = s * 0.5; s = 0.5 / s; qz = (t.c - t.b) * s; }
    else { double s = sqrt(1.0 - t.a - t.d + 1.0); qz = s * 0.5; s = 0.5 / s; qw = (t.c - t.b) * s; }
    double len2 = qz * qz + qw * qw, s = 2.0 / len2, zs = qz * s;
    return atan2((0.0 + qw * zs) / 1.0, (1.0 - (0.0 + qz * zs)) / 1.0);
}
double quat_yaw(double x, double y, double z, double w) {   // img_env.cpp:180-184
    double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
    double ys = y * s, zs = z * s;
    double wy = w * ys, wz = w * zs, xy = x * ys, xz = x * zs, yy = y * ys, zz = z * zs;
    double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy;
    if (fabs(m20) >= 1) return 0.0;
    double pitch = -asin(m20);
    return atan2(m10 / cos(pitch), m00 / cos(pitch));
}
int cell_of(double x, double res) { return (int)round(x / res); }   // grid_map.cpp:40-44

struct Grid {   // GridMap, grid_map.h
    int h = 0, w = 0; double res = 0; std::vector<uint8_t> m;
    bool inside(int r, int c) const { return r >= 0 && r < h && c >= 0 && c < w; }
    uint8_t& at(int r, int c) { return m[(size_t)r * w + c]; }
};
typedef std::vector<double> Pts;   // x0,y0,x1,y1,...
void shape_circle(double s0, double s1, double s2, Pts& out) {   // agent.cpp:18-30
    double resolution = 0.01; int bb = (int)ceil(s2 / resolution);
    for (int m = -bb; m <= bb; m++) for (int n = -bb; n <= bb; n++)
        if (sqrt(m * resolution * m * resolution + n * resolution * n * resolution) <= s2) { out.push_back(m * resolution + s0); out.push_back(n * resolution + s1); }
}
void shape_rect(double s0, double s1, double s2, double s3, Pts& out) {   // agent.cpp:51-62
    double resolution = 0.01;
    int x_min = (int)floor(s0 / resolution), x_max = (int)ceil(s1 / resolution), y_min = (int)floor(s2 / resolution), y_max = (int)ceil(s3 / resolution);
    for (int m = x_min; m <= x_max; m++) for (int n = y_min; n <= y_max; n++) { out.push_back(m * resolution); out.push_back(n * resolution); }
}

struct Lim { bool hv = false, ha = false, hj = false; double minv = 0, maxv = 0, mina = 0, maxa = 0, minj = 0, maxj = 0; };
double clampd(double x, double lo, double hi) { return std::min(std::max(lo, x), hi); }
int sgn(double x) { return x == 0 ? 0 : (int)(x / fabs(x)); }
void limit(const Lim& L, double& v, double v0, double v1, double dt) {   // speed_limit.cpp:92-173
    if (L.hj) { double dv = v - v0, dv0 = v0 - v1, dt2 = 2. * dt * dt; double da = clampd(dv - dv0, L.minj * dt2, L.maxj * dt2); v = v0 + dv0 + da; }
    if (L.ha) {
        const double tmp = v; int vs = sgn(v), v0s = sgn(v0);
        if (vs + v0s != 0) {
            double dv_min = L.mina * dt, dv_max = L.maxa * dt, dv = v - v0; int dvs = sgn(dv);
            if (dvs == v0s || dvs == vs) dv = dvs * clampd(fabs(dv), dv_min, dv_max);
            else dv = dvs * fabs(clampd(-fabs(dv), dv_min, dv_max));
            v = v0 + dv;
        } else {
            double zero_dt = fabs(v0 / L.mina);
            if (zero_dt >= dt) v = v0s * (fabs(v0) - fabs(L.mina) * dt);
            else { double v_dt = fabs(v / L.maxa); if (zero_dt + v_dt >= dt) v = vs * fabs(L.maxa * (dt - zero_dt)); else v = tmp; }
        }
    }
    if (L.hv) v = clampd(v, L.minv, L.maxv);
}

struct Robot {   // Agent, agent.h
    int shape = 0; double size[4] = {0, 0, 0, 0}; Pts bbox; double sx = 0, sy = 0; Lim lv, lw;
    double x = 0, y = 0, yaw = 0, gx = 0, gy = 0, gyaw = 0, l0v = 0, l0w = 0, l1v = 0, l1w = 0, vx = 0, vy = 0;
    int coll = 0; bool arrive = false; int beep = 0;
    Grid view; std::vector<double> hits;
};
struct Ped {     // PedAgent
    int shape = 0; double size[6] = {0, 0, 0, 0, 0, 0}; double max_speed = 0; Pts bbox, lbox, rbox;
    double x = 0, y = 0, yaw = 0, lx = 0, ly = 0, lyaw = 0, vx = 0, vy = 0;
    int state = 0, last_state = 0; double rem = 0, ll[3] = {0, 0, 0}, rl[3] = {0, 0, 0};
    int tidx = 0; std::vector<double> traj, trajv;   // xyz triples
};

// ---------------- RVO2 (ervo_ros) ----------------
struct F2 { float x, y; };
F2 mk(float x, float y) { return F2{x, y}; }
F2 operator+(F2 p, F2 q) { return mk(p.x + q.x, p.y + q.y); }
F2 operator-(F2 p, F2 q) { return mk(p.x - q.x, p.y - q.y); }
F2 operator-(F2 p) { return mk(-p.x, -p.y); }
float operator*(F2 p, F2 q) { return p.x * q.x + p.y * q.y; }
F2 operator*(float s, F2 p) { return mk(s * p.x, s * p.y); }
F2 operator*(F2 p, float s) { return mk(p.x * s, p.y * s); }
F2 operator/(F2 p, float s) { const float inv = 1.0f / s; return mk(p.x * inv, p.y * inv); }
float absSq(F2 p) { return p * p; }
float vabs(F2 p) { return std::sqrt(p * p); }
float det(F2 p, F2 q) { return p.x * q.y - p.y * q.x; }
F2 unit(F2 p) { return p / vabs(p); }
float sq(float a) { return a * a; }
float leftOf(F2 a, F2 b, F2 c) { return det(a - c, b - a); }
const float EPS = 0.00001f;
struct Obst { F2 p, dir; bool convex; int next, prev; };
struct Node { int obst, left, right; };
struct Line { F2 point, direction; };
struct RvoAgent { F2 pos = mk(0, 0), vel = mk(0, 0), pref = mk(0, 0), nv = mk(0, 0); float maxSpeed = 0; };

struct Rvo {
    std::vector<RvoAgent> ag; std::vector<Obst> ob; std::vector<Node> nodes; int root = -1; float dt = 0;
    void add_obstacle(const F2* v, int n) {   // RVOSimulator.cpp:130-168
        int first = (int)ob.size();
        for (int i = 0; i < n; i++) {
            Obst o; o.p = v[i]; o.next = o.prev = -1; int me = (int)ob.size();
            if (i != 0) { o.prev = me - 1; ob[me - 1].next = me; }
            if (i == n - 1) o.next = first;
            o.dir = unit(v[i == n - 1 ? 0 : i + 1] - v[i]);
            o.convex = n == 2 ? true : leftOf(v[i == 0 ? n - 1 : i - 1], v[i], v[i == n - 1 ? 0 : i + 1]) >= 0.0f;
            ob.push_back(o);
            if (i == n - 1) ob[first].prev = me;
        }
    }
    int build(const std::vector<int>& L) {     // KdTree.cpp:130-257
        if (L.empty()) return -1;
        size_t best = 0, minL = L.size(), minR = L.size();
        for (size_t i = 0; i < L.size(); i++) {
            size_t ls = 0, rs = 0; int I1 = L[i], I2 = ob[I1].next;
            for (size_t j = 0; j < L.size(); j++) {
                if (i == j) continue;
                int J1 = L[j], J2 = ob[J1].next;
                float a = leftOf(ob[I1].p, ob[I2].p, ob[J1].p), b = leftOf(ob[I1].p, ob[I2].p, ob[J2].p);
                if (a >= -EPS && b >= -EPS) ++ls; else if (a <= EPS && b <= EPS) ++rs; else { ++ls; ++rs; }
                if (std::make_pair(std::max(ls, rs), std::min(ls, rs)) >= std::make_pair(std::max(minL, minR), std::min(minL, minR))) break;
            }
            if (std::make_pair(std::max(ls, rs), std::min(ls, rs)) < std::make_pair(std::max(minL, minR), std::min(minL, minR))) { minL = ls; minR = rs; best = i; }
        }
        std::vector<int> lo(minL), ro(minR); size_t lc = 0, rc = 0;
        int I1 = L[best], I2 = ob[I1].next;
        for (size_t j = 0; j < L.size(); j++) {
            if (j == best) continue;
            int J1 = L[j], J2 = ob[J1].next;
            float a = leftOf(ob[I1].p, ob[I2].p, ob[J1].p), b = leftOf(ob[I1].p, ob[I2].p, ob[J2].p);
            if (a >= -EPS && b >= -EPS) lo[lc++] = J1;
            else if (a <= EPS && b <= EPS) ro[rc++] = J1;
            else {
                float t = det(ob[I2].p - ob[I1].p, ob[J1].p - ob[I1].p) / det(ob[I2].p - ob[I1].p, ob[J1].p - ob[J2].p);
                Obst n; n.p = ob[J1].p + t * (ob[J2].p - ob[J1].p); n.prev = J1; n.next = J2; n.convex = true; n.dir = ob[J1].dir;
                int id = (int)ob.size(); ob.push_back(n); ob[J1].next = id; ob[J2].prev = id;
                if (a > 0.0f) { lo[lc++] = J1; ro[rc++] = id; } else { ro[rc++] = J1; lo[lc++] = id; }
            }
        }
        int me = (int)nodes.size(); nodes.push_back(Node{I1, -1, -1});
        int l = build(lo); nodes[me].left = l;
        int r = build(ro); nodes[me].right = r;
        return me;
    }
    void process() { nodes.clear(); std::vector<int> L(ob.size()); for (size_t i = 0; i < L.size(); i++) L[i] = (int)i; root = build(L); }
    void query_obst(int node, F2 pos, float rangeSq, std::vector<std::pair<float, int>>& out) {   // KdTree.cpp:322-353 + Agent.cpp:820-838
        if (node < 0) return;
        int o1 = nodes[node].obst, o2 = ob[o1].next;
        float side = leftOf(ob[o1].p, ob[o2].p, pos);
        query_obst(side >= 0.0f ? nodes[node].left : nodes[node].right, pos, rangeSq, out);
        float dline = sq(side) / absSq(ob[o2].p - ob[o1].p);
        if (dline < rangeSq) {
            if (side < 0.0f) {
                F2 a = ob[o1].p, b = ob[o2].p; float r = ((pos - a) * (b - a)) / absSq(b - a);
                float dsq = r < 0.0f ? absSq(pos - a) : (r > 1.0f ? absSq(pos - b) : absSq(pos - (a + r * (b - a))));
                if (dsq < rangeSq) {
                    out.push_back(std::make_pair(dsq, o1)); size_t i = out.size() - 1;
                    while (i != 0 && dsq < out[i - 1].first) { out[i] = out[i - 1]; --i; }
                    out[i] = std::make_pair(dsq, o1);
                }
            }
            query_obst(side >= 0.0f ? nodes[node].right : nodes[node].left, pos, rangeSq, out);
        }
    }
    static bool lp1(const std::vector<Line>& L, size_t no, float radius, F2 opt, bool dirOpt, F2& res) {   // Agent.cpp:845-918
        float dp = L[no].point * L[no].direction, disc = sq(dp) + sq(radius) - absSq(L[no].point);
        if (disc < 0.0f) return false;
        float sd = std::sqrt(disc), tL = -dp - sd, tR = -dp + sd;
        for (size_t i = 0; i < no; i++) {
            float den = det(L[no].direction, L[i].direction), num = det(L[i].direction, L[no].point - L[i].point);
            if (std::fabs(den) <= EPS) { if (num < 0.0f) return false; else continue; }
            float t = num / den;
            if (den >= 0.0f) tR = std::min(tR, t); else tL = std::max(tL, t);
            if (tL > tR) return false;
        }
        if (dirOpt) { if (opt * L[no].direction > 0.0f) res = L[no].point + tR * L[no].direction; else res = L[no].point + tL * L[no].direction; }
        else { float t = L[no].direction * (opt - L[no].point);
            if (t < tL) res = L[no].point + tL * L[no].direction; else if (t > tR) res = L[no].point + tR * L[no].direction; else res = L[no].point + t * L[no].direction; }
        return true;
    }
    static size_t lp2(const std::vector<Line>& L, float radius, F2 opt, bool dirOpt, F2& res) {   // Agent.cpp:920-948
        if (dirOpt) res = opt * radius; else if (absSq(opt) > sq(radius)) res = unit(opt) * radius; else res = opt;
        for (size_t i = 0; i < L.size(); i++)
            if (det(L[i].direction, L[i].point - res) > 0.0f) { F2 tmp = res; if (!lp1(L, i, radius, opt, dirOpt, res)) { res = tmp; return i; } }
        return L.size();
    }
    static void lp3(const std::vector<Line>& L, size_t nObst, size_t begin, float radius, F2& res) {   // Agent.cpp:950-1001
        float distance = 0.0f;
        for (size_t i = begin; i < L.size(); i++)
            if (det(L[i].direction, L[i].point - res) > distance) {
                std::vector<Line> P(L.begin(), L.begin() + (ptrdiff_t)nObst);
                for (size_t j = nObst; j < i; j++) {
                    Line ln; float dt = det(L[i].direction, L[j].direction);
                    if (std::fabs(dt) <= EPS) { if (L[i].direction * L[j].direction > 0.0f) continue; else ln.point = 0.5f * (L[i].point + L[j].point); }
                    else ln.point = L[i].point + (det(L[j].direction, L[i].point - L[j].point) / dt) * L[i].direction;
                    ln.direction = unit(L[j].direction - L[i].direction); P.push_back(ln);
                }
                F2 tmp = res;
                if (lp2(P, radius, mk(-L[i].direction.y, L[i].direction.x), true, res) < P.size()) res = tmp;
                distance = det(L[i].direction, L[i].point - res);
            }
    }
    void new_velocity(int self, bool ervo, const std::vector<F2>& ps, const std::vector<float>& rs) {   // Agent.cpp:50-61, 437-793, 63-69
        const float radius = 0.5f, nDist = 0.5f, tH = 5.f, tHO = 5.f; const size_t maxN = 10;
        RvoAgent& A = ag[self];
        std::vector<std::pair<float, int>> on, an;
        query_obst(root, A.pos, sq(tHO * A.maxSpeed + radius), on);
        float rangeSq = sq(nDist);
        for (int o = 0; o < (int)ag.size(); o++) {   // (the node walks a k-d tree; same <=10 nearest list)
            if (o == self) continue;
            float dsq = absSq(A.pos - ag[o].pos);
            if (dsq < rangeSq) {
                if (an.size() < maxN) an.push_back(std::make_pair(dsq, o));
                size_t i = an.size() - 1;
                while (i != 0 && dsq < an[i - 1].first) { an[i] = an[i - 1]; --i; }
                an[i] = std::make_pair(dsq, o);
                if (an.size() == maxN) rangeSq = an.back().first;
            }
        }
        std::vector<Line> L; const float iTHO = 1.0f / tHO;
        for (size_t i = 0; i < on.size(); i++) {
            int o1 = on[i].second, o2 = ob[o1].next;
            F2 r1 = ob[o1].p - A.pos, r2 = ob[o2].p - A.pos;
            bool covered = false;
            for (size_t j = 0; j < L.size(); j++)
                if (det(iTHO * r1 - L[j].point, L[j].direction) - iTHO * radius >= -EPS && det(iTHO * r2 - L[j].point, L[j].direction) - iTHO * radius >= -EPS) { covered = true; break; }
            if (covered) continue;
            float d1 = absSq(r1), d2 = absSq(r2), rSq = sq(radius);
            F2 ov = ob[o2].p - ob[o1].p; float s = (-r1 * ov) / absSq(ov); float dL = absSq(-r1 - s * ov);
            Line ln;
            if (s < 0.0f && d1 <= rSq) { if (ob[o1].convex) { ln.point = mk(0, 0); ln.direction = unit(mk(-r1.y, r1.x)); L.push_back(ln); } continue; }
            else if (s > 1.0f && d2 <= rSq) { if (ob[o2].convex && det(r2, ob[o2].dir) >= 0.0f) { ln.point = mk(0, 0); ln.direction = unit(mk(-r2.y, r2.x)); L.push_back(ln); } continue; }
            else if (s >= 0.0f && s < 1.0f && dL <= rSq) { ln.point = mk(0, 0); ln.direction = -ob[o1].dir; L.push_back(ln); continue; }
            F2 lleg, rleg;
            if (s < 0.0f && dL <= rSq) {
                if (!ob[o1].convex) continue;
                o2 = o1; float leg1 = std::sqrt(d1 - rSq);
                lleg = mk(r1.x * leg1 - r1.y * radius, r1.x * radius + r1.y * leg1) / d1; rleg = mk(r1.x * leg1 + r1.y * radius, -r1.x * radius + r1.y * leg1) / d1;
            } else if (s > 1.0f && dL <= rSq) {
                if (!ob[o2].convex) continue;
                o1 = o2; float leg2 = std::sqrt(d2 - rSq);
                lleg = mk(r2.x * leg2 - r2.y * radius, r2.x * radius + r2.y * leg2) / d2; rleg = mk(r2.x * leg2 + r2.y * radius, -r2.x * radius + r2.y * leg2) / d2;
            } else {
                if (ob[o1].convex) { float leg1 = std::sqrt(d1 - rSq); lleg = mk(r1.x * leg1 - r1.y * radius, r1.x * radius + r1.y * leg1) / d1; } else lleg = -ob[o1].dir;
                if (ob[o2].convex) { float leg2 = std::sqrt(d2 - rSq); rleg = mk(r2.x * leg2 + r2.y * radius, -r2.x * radius + r2.y * leg2) / d2; } else rleg = ob[o1].dir;
            }
            int ln_ = ob[o1].prev; bool lf = false, rf = false;
            if (ob[o1].convex && det(lleg, -ob[ln_].dir) >= 0.0f) { lleg = -ob[ln_].dir; lf = true; }
            if (ob[o2].convex && det(rleg, ob[o2].dir) <= 0.0f) { rleg = ob[o2].dir; rf = true; }
            F2 lc = iTHO * (ob[o1].p - A.pos), rc = iTHO * (ob[o2].p - A.pos), cv = rc - lc;
            float t = (o1 == o2 ? 0.5f : ((A.vel - lc) * cv) / absSq(cv)), tl = ((A.vel - lc) * lleg), tr = ((A.vel - rc) * rleg);
            if ((t < 0.0f && tl < 0.0f) || (o1 == o2 && tl < 0.0f && tr < 0.0f)) { F2 u = unit(A.vel - lc); ln.direction = mk(u.y, -u.x); ln.point = lc + radius * iTHO * u; L.push_back(ln); continue; }
            else if (t > 1.0f && tr < 0.0f) { F2 u = unit(A.vel - rc); ln.direction = mk(u.y, -u.x); ln.point = rc + radius * iTHO * u; L.push_back(ln); continue; }
            const float INF = std::numeric_limits<float>::infinity();
            float dc = ((t < 0.0f || t > 1.0f || o1 == o2) ? INF : absSq(A.vel - (lc + t * cv)));
            float dl = ((tl < 0.0f) ? INF : absSq(A.vel - (lc + tl * lleg))), dr = ((tr < 0.0f) ? INF : absSq(A.vel - (rc + tr * rleg)));
            if (dc <= dl && dc <= dr) { ln.direction = -ob[o1].dir; ln.point = lc + radius * iTHO * mk(-ln.direction.y, ln.direction.x); L.push_back(ln); continue; }
            else if (dl <= dr) { if (lf) continue; ln.direction = lleg; ln.point = lc + radius * iTHO * mk(-ln.direction.y, ln.direction.x); L.push_back(ln); continue; }
            else { if (rf) continue; ln.direction = -rleg; ln.point = rc + radius * iTHO * mk(-ln.direction.y, ln.direction.x); L.push_back(ln); continue; }
        }
        size_t nObst = L.size(); const float iTH = 1.0f / tH;
        for (size_t i = 0; i < an.size(); i++) {
            const RvoAgent& O = ag[an[i].second];
            F2 rp = O.pos - A.pos, rv = A.vel - O.vel; float dsq = absSq(rp), cr = radius + radius, crSq = sq(cr);
            Line ln; F2 u;
            if (dsq > crSq) {
                F2 w = rv - iTH * rp; float wl2 = absSq(w), dp1 = w * rp;
                if (dp1 < 0.0f && sq(dp1) > crSq * wl2) { float wl = std::sqrt(wl2); F2 uw = w / wl; ln.direction = mk(uw.y, -uw.x); u = (cr * iTH - wl) * uw; }
                else { float leg = std::sqrt(dsq - crSq);
                    if (det(rp, w) > 0.0f) ln.direction = mk(rp.x * leg - rp.y * cr, rp.x * cr + rp.y * leg) / dsq;
                    else ln.direction = -mk(rp.x * leg + rp.y * cr, -rp.x * cr + rp.y * leg) / dsq;
                    float dp2 = rv * ln.direction; u = dp2 * ln.direction - rv; }
            } else { float its = 1.0f / dt; F2 w = rv - its * rp; float wl = vabs(w); F2 uw = w / wl; ln.direction = mk(uw.y, -uw.x); u = (cr * its - wl) * uw; }
            ln.point = A.vel + 0.5f * u; L.push_back(ln);
        }
        size_t fl = lp2(L, A.maxSpeed, A.pref, false, A.nv);
        if (fl < L.size()) lp3(L, nObst, fl, A.maxSpeed, A.nv);
        if (ervo) for (size_t b = 0; b < ps.size(); b++) { F2 ev = A.pos - ps[b]; if (vabs(ev) > rs[b] || vabs(ev) < 1e-4) continue; A.nv = A.nv + unit(ev); }
    }
    void do_step(bool ervo, const std::vector<F2>& ps, const std::vector<float>& rs) {   // RVOSimulator.cpp:180-199
        for (size_t i = 0; i < ag.size(); i++) new_velocity((int)i, ervo, ps, rs);
        for (auto& a : ag) { a.vel = a.nv; a.pos = a.pos + a.vel * dt; }
    }
};

// ---------------- libpedsim SFM ----------------
struct D3 { double x = 0, y = 0, z = 0; };
D3 d3(double x, double y, double z = 0) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
D3 operator*(double f, D3 a) { return d3(f * a.x, f * a.y, f * a.z); }
D3 operator*(D3 a, double f) { return d3(f * a.x, f * a.y, f * a.z); }
D3 operator/(D3 a, double dv) { return a * (1 / dv); }
double l2(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
double ln(D3 a) { if (a.x == 0 && a.y == 0 && a.z == 0) return 0; return sqrt(l2(a)); }
D3 nrm(D3 a) { double l = ln(a); if (l == 0) return D3(); return d3(a.x / l, a.y / l, a.z / l); }
struct SfmAgent { D3 p, v; double vmax = 1.2; int dest = -1, lastdest = -1, front = 0; bool in_tree = true; std::vector<D3> wp; D3 fd, fs, fo, fl, dirn; };
struct Sfm {
    std::vector<SfmAgent> ag; std::vector<std::pair<D3, D3>> obs;
    void forces(int self) {   // ped_agent.cpp:236-507
        SfmAgent& A = ag[self];
        A.dirn = D3();
        if (A.dest < 0 && !A.wp.empty()) { A.dest = A.front; A.front = (A.front + 1) % (int)A.wp.size(); }
        bool reached = false;
        if (A.dest >= 0) { D3 diff = d3(A.wp[A.dest].x - A.p.x, A.wp[A.dest].y - A.p.y); reached = ln(diff) < A.wp[A.dest].z; A.dirn = nrm(diff); }
        if (A.dest >= 0 && reached) { A.lastdest = A.dest; A.dest = -1; }
        A.fd = nrm(A.dirn) * A.vmax;
        const double pi = 3.14159265; int look = 0; D3 soc;
        for (int o = 0; o < (int)ag.size(); o++) {
            const SfmAgent& O = ag[o];
            if (!O.in_tree || o == self) continue;
            double dx = O.p.x - A.p.x, dy = O.p.y - A.p.y, dist2 = dx * dx + dy * dy;
            if (dist2 < 400) {
                double at2v = atan2(-A.dirn.x, -A.dirn.y), at2d = atan2(-dx, -dy), at2v2 = atan2(-O.v.x, -O.v.y);
                double s = at2d - at2v; if (s > pi) s -= 2 * pi; if (s < -pi) s += 2 * pi;
                double vv = at2v - at2v2; if (vv > pi) vv -= 2 * pi; if (vv < -pi) vv += 2 * pi;
                if (fabs(vv) > 2.5) { if (s < 0 && s > -0.3) look--; if (s > 0 && s < 0.3) look++; }
            }
            D3 diff = O.p - A.p;
            if (l2(diff) > 64.0) continue;
            D3 dd = nrm(diff), vd = A.v - O.v, iv = 2.0 * vd + dd; double il = ln(iv); D3 id = iv / il;
            double theta = atan2(dd.y, dd.x) - atan2(id.y, id.x);
            if (theta > M_PI) theta -= 2 * M_PI; else if (theta <= -M_PI) theta += 2 * M_PI;
            int ts = theta == 0 ? 0 : (int)(theta / fabs(theta));
            double B = 0.35 * il;
            double fva = -exp(-ln(diff) / B - (3 * B * theta) * (3 * B * theta)), faa = -ts * exp(-ln(diff) / B - (2 * B * theta) * (2 * B * theta));
            soc = soc + (fva * id + faa * d3(-id.y, id.x));
        }
        D3 lf;
        if (look < 0) { lf.x = 0.5f * A.dirn.y; lf.y = 0.5f * -A.dirn.x; }
        if (look > 0) { lf.x = 0.5f * -A.dirn.y; lf.y = 0.5f * A.dirn.x; }
        A.fl = lf; A.fs = soc;
        D3 md; double mds = INFINITY;
        for (auto& sg : obs) {
            D3 re = sg.second - sg.first, rp = A.p - sg.first; double lam = (rp.x * re.x + rp.y * re.y + rp.z * re.z) / l2(re);
            D3 cp = lam <= 0 ? sg.first : (lam >= 1 ? sg.second : sg.first + lam * re);
            D3 df = A.p - cp; double ds = l2(df); if (ds < mds) { mds = ds; md = df; }
        }
        A.fo = exp(-(sqrt(mds) - 0.2) / 0.8) * nrm(md);
    }
    void move(int self, double h) {   // ped_agent.cpp:519-571 + the never-split root leaf of pedscene.h:19
        SfmAgent& A = ag[self];
        D3 pd = A.p + A.v * h;
        for (auto& sg : obs) {
            double s1x = pd.x - A.p.x, s1y = pd.y - A.p.y, s2x = sg.second.x - sg.first.x, s2y = sg.second.y - sg.first.y;
            double s = (-s1y * (A.p.x - sg.first.x) + s1x * (A.p.y - sg.first.y)) / (-s2x * s1y + s1x * s2y);
            double t = (s2x * (A.p.y - sg.first.y) - s2y * (A.p.x - sg.first.x)) / (-s2x * s1y + s1x * s2y);
            if (s >= 0 && s <= 1 && t >= 0 && t <= 1) pd = d3(A.p.x + (t * s1x), A.p.y + (t * s1y)) - nrm(A.v * h) * 0.1;
        }
        A.p = pd;
        D3 a = 1.0 * A.fd + 2.1 * A.fs + 1.0 * A.fo + 1.0 * A.fl + D3();
        A.v = 0.5 * A.v + a * h;
        if (ln(A.v) > A.vmax) A.v = nrm(A.v) * A.vmax;
        if (A.p.x < 0 || A.p.x > 10 || A.p.y < 10 || A.p.y > 20) A.in_tree = false;
    }
};

struct Port {
    // config (float32-widened like the node sees it)
    double res = 0, vwid = 0, vhei = 0, step_hz = 0, ang0 = 0, ang1 = 0, dmin = 0, dmax = 0, beep_r = 0, ped_ca_p = 0;
    int state_dim = 3, use_laser = 1, range_total = 0, relation = 0, ktype = 0, scene = 0, vh = 0, vw = 0;
    Grid stat, obsm, pedm;
    std::vector<Robot> robots; std::vector<Ped> peds;
    T2 view_base, base_view;
    Rvo rvo; Sfm sfm;
    std::vector<double> obj_seg;
    int step_ = 0;
    // last reply
    std::vector<std::vector<float>> st_state, st_laser, st_ped; std::vector<int> st_coll, st_arr;
};

int draw(Port& P, Grid& g, const T2& bw, const Pts& bb, int value, int frame /*0 world,1 view*/) {   // agent.cpp:285-327
    int code = 0;
    for (size_t k = 0; k + 1 < bb.size(); k += 2) {
        double wx, wy;
        if (frame == 0) t2_map(bw, bb[k], bb[k + 1], wx, wy); else t2_map(P.base_view, bb[k], bb[k + 1], wx, wy);
        int r = cell_of(wx, g.res), c = cell_of(wy, g.res);
        if (g.inside(r, c)) {
            uint8_t& v = g.at(r, c);
            if (v == 0) code = 1; else if (v == 1) code = 2; else if (v == 2) code = 3; else if (value >= 0) v = (uint8_t)value;
        }
    }
    return code;
}
void draw_leg(Port& P, Grid& g, const Ped& p) {   // agent.cpp:737-774
    T2 bw = t2_pose(p.x, p.y, p.yaw);
    for (int leg = 0; leg < 2; leg++) {
        const Pts& bb = leg == 0 ? p.lbox : p.rbox; const double* off = leg == 0 ? p.ll : p.rl;
        for (size_t k = 0; k + 1 < bb.size(); k += 2) {
            double bx = bb[k] + off[0], by = bb[k + 1] + off[1], wx, wy;
            t2_map(bw, bx, by, wx, wy);
            int r = cell_of(wx, g.res), c = cell_of(wy, g.res);
            if (g.inside(r, c)) { uint8_t& v = g.at(r, c); if (leg == 0) { if (v != 0) v = 1; } else { if (v != 1) v = 1; } }
        }
    }
    (void)P;
}
double bresenham(int x1, int y1, int x2, int y2, Grid& src, Grid& dst) {   // agent.cpp:511-624
    double hit = 6; double x0 = x1 * dst.res, y0 = y1 * dst.res;
    int w = x2 - x1, h = y2 - y1, dx = ((w > 0) << 1) - 1, dy = ((h > 0) << 1) - 1; w = abs(w); h = abs(h);
    bool ended = false; int ex = -1, ey = -1, f, x = x1, y = y1;
    bool xmajor = w > h;
    f = xmajor ? 2 * h - w : 2 * w - h;
    while (xmajor ? x != x2 : y != y2) {
        if (!src.inside(x, y)) return hit;
        int cur = src.at(x, y);
        if (!ended) {
            if (cur != 0) dst.at(x, y) = 255;
            else if (ex == -1) { dst.at(x, y) = 0; ended = true; ex = x; ey = y; double cx = x * dst.res, cy = y * dst.res; hit = sqrt((x0 - cx) * (x0 - cx) + (y0 - cy) * (y0 - cy)); }
        } else if (x != ex && y != ey) dst.at(x, y) = 200;
        if (xmajor) { if (f < 0) f += 2 * h; else { y += dy; f += (h - w) * 2; } x += dx; }
        else { if (f < 0) f += 2 * w; else { x += dx; f += (w - h) * 2; } y += dy; }
    }
    return hit;
}
void view(Port& P, Robot& R, Grid& g) {   // agent.cpp:356-509
    if (R.coll || R.arrive) return;
    T2 bw = t2_pose(R.x, R.y, R.yaw);
    R.coll = draw(P, g, bw, R.bbox, -1, 0);
    double ovx, ovy; t2_map(P.base_view, R.sx, R.sy, ovx, ovy);
    int ox = cell_of(ovx, P.res), oy = cell_of(ovy, P.res);
    R.view.h = P.vh; R.view.w = P.vw; R.view.res = P.res; R.view.m.assign((size_t)P.vh * P.vw, 200);
    Grid laser = R.view;
    T2 vw = t2_mul(bw, P.view_base);
    for (int i = 0; i < P.vh; i++) for (int j = 0; j < P.vw; j++) {
        double xv = i * P.res, yv = j * P.res, xb, yb; t2_map(P.view_base, xv, yv, xb, yb);
        double ang = atan2(yb - R.sy, xb - R.sx);
        if (ang <= P.ang0 || ang >= P.ang1 || xb < P.dmin || xb > P.dmax) continue;
        double wx, wy; t2_map(vw, xv, yv, wx, wy);
        int r = cell_of(wx, g.res), c = cell_of(wy, g.res);
        if (g.inside(r, c)) R.view.at(i, j) = g.at(r, c) < 250 ? 0 : 255;
    }
    if (P.use_laser) {
        R.hits.clear();
        double mw = P.base_view.ox, mh = P.base_view.oy, max_range = sqrt(mw * mw + mh * mh);
        double step = fabs(P.ang1 - P.ang0) / P.range_total;
        for (int i = 0; i < P.range_total; i++) {
            double a = P.ang0 + step * i, x = max_range * cos(a), y = max_range * sin(a), vx, vy;
            t2_map(P.base_view, x, y, vx, vy);
            R.hits.push_back(bresenham(ox, oy, cell_of(vx, P.res), cell_of(vy, P.res), R.view, laser));
        }
        R.view = laser;
    }
    draw(P, R.view, bw, R.bbox, 100, 1);
}
void observe(Port& P) {   // view_ped + view_robot + get_states, img_env.cpp:547-674
    P.pedm = P.obsm;
    for (auto& p : P.peds) {
        if (p.shape == 0) draw(P, P.pedm, t2_pose(p.x, p.y, p.yaw), p.bbox, 1, 0);
        else if (p.shape == 2) draw_leg(P, P.pedm, p);
    }
    int R = (int)P.robots.size();
    for (int i = 0; i < R; i++) {
        Grid g = P.pedm;                                             // full-map clone per robot (img_env.cpp:623)
        for (int j = 0; j < R; j++) if (j != i) draw(P, g, t2_pose(P.robots[j].x, P.robots[j].y, P.robots[j].yaw), P.robots[j].bbox, 2, 0);
        view(P, P.robots[i], g);
    }
    P.st_state.assign(R, {}); P.st_laser.assign(R, {}); P.st_ped.assign(R, {}); P.st_coll.assign(R, 0); P.st_arr.assign(R, 0);
    for (int i = 0; i < R; i++) {
        Robot& r = P.robots[i];
        T2 tb = t2_inv(t2_mul(t2_inv(t2_pose(r.gx, r.gy, r.gyaw)), t2_pose(r.x, r.y, r.yaw)));   // agent.cpp:156-184
        std::vector<double> s = {tb.ox, tb.oy};
        if (P.state_dim == 3) s.push_back(t2_yaw(tb));
        else if (P.state_dim == 4) { s.push_back(r.l0v); s.push_back(r.l0w); }
        else if (P.state_dim == 5) { s.push_back(t2_yaw(tb)); s.push_back(r.l0v); s.push_back(r.l0w); }
        for (double v : s) P.st_state[i].push_back((float)v);
        for (double v : r.hits) P.st_laser[i].push_back((float)v);
        P.st_coll[i] = r.coll; P.st_arr[i] = r.arrive;
        T2 wb = t2_inv(t2_pose(r.x, r.y, r.yaw));
        for (auto& p : P.peds) {                                      // img_env.cpp:568-584
            double px, py; t2_map(wb, p.x, p.y, px, py);
            T2 rot = wb; rot.ox = 0; rot.oy = 0; double vx, vy; t2_map(rot, p.vx, p.vy, vx, vy);
            P.st_ped[i].push_back((float)px); P.st_ped[i].push_back((float)py); P.st_ped[i].push_back((float)vx); P.st_ped[i].push_back((float)vy);
            P.st_ped[i].push_back((float)p.size[2]);
        }
    }
}
void corners(const double* size, int shape, double x, double y, double yaw, double* out) {   // agent.cpp:626-651
    T2 t = t2_pose(x, y, yaw);
    if (shape == 0) { t2_map(t, size[0] - size[2], size[1] - size[2], out[0], out[1]); t2_map(t, size[0] + size[2], size[1] + size[2], out[2], out[3]); }
    else { t2_map(t, size[0], size[2], out[0], out[1]); t2_map(t, size[1], size[3], out[2], out[3]); }
}
double f32(double v) { return (double)(float)v; }
}  // namespace

extern "C" {
void* port_create() { return new Port(); }
void port_destroy(void* h) { delete static_cast<Port*>(h); }

int port_init(void* h, const double* sc, double, const uint8_t* grid, int H, int W, int, int, int R, const double* rd, const char* ktype,
              int Pn, const double* pd, const char* scene_type) {
    Port& P = *static_cast<Port*>(h);
    P.res = f32(sc[0]); P.vwid = f32(sc[1]); P.vhei = f32(sc[2]); P.step_hz = f32(sc[3]); P.state_dim = (int)sc[4];
    P.use_laser = sc[11] != 0; P.range_total = (int)sc[12]; P.ang0 = f32(sc[13]); P.ang1 = f32(sc[14]); P.dmin = f32(sc[15]); P.dmax = f32(sc[16]);
    P.beep_r = f32(sc[17]); P.ped_ca_p = f32(sc[18]); P.relation = (int)sc[19];
    P.ktype = std::string(ktype) == "omni" ? 1 : 0;
    std::string st = scene_type;
    P.scene = Pn == 0 ? 0 : (st == "pedscene" ? 1 : st == "rvoscene" ? 2 : st == "ervoscene" ? 3 : st == "dataset" ? 4 : 0);
    P.vw = (int)(P.vwid / P.res); P.vh = (int)(P.vhei / P.res);
    P.view_base = t2_pose(P.vhei / 2, P.vwid / 2, 3.14159); P.base_view = t2_inv(P.view_base);   // agent.cpp:79-90
    P.stat.h = H; P.stat.w = W; P.stat.res = P.res; P.stat.m.assign(grid, grid + (size_t)H * W); P.obsm = P.stat;
    P.robots.assign(R, Robot()); P.peds.assign(Pn, Ped());
    for (int i = 0; i < R; i++) {
        const double* d = rd + 25 * i; Robot& r = P.robots[i];
        r.shape = (int)d[0]; for (int k = 0; k < 4; k++) r.size[k] = f32(d[1 + k]); r.sx = f32(d[5]); r.sy = f32(d[6]);
        if (r.shape == 0) shape_circle(r.size[0], r.size[1], r.size[2], r.bbox); else if (r.shape == 1) shape_rect(r.size[0], r.size[1], r.size[2], r.size[3], r.bbox);
        for (int L = 0; L < 2; L++) { const double* q = d + 7 + 9 * L; Lim& m = L == 0 ? r.lv : r.lw;
            m.hv = q[0] != 0; m.ha = q[1] != 0; m.hj = q[2] != 0; m.minv = f32(q[3]); m.maxv = f32(q[4]); m.mina = f32(q[5]); m.maxa = f32(q[6]); m.minj = f32(q[7]); m.maxj = f32(q[7]); }
    }
    for (int i = 0; i < Pn; i++) {
        const double* d = pd + 8 * i; Ped& p = P.peds[i];
        p.shape = (int)d[0]; for (int k = 0; k < 6; k++) p.size[k] = f32(d[1 + k]); p.max_speed = f32(d[7]);
        if (p.shape == 0) shape_circle(p.size[0], p.size[1], p.size[2], p.bbox);
        else if (p.shape == 2) { shape_circle(0, 0, p.size[2], p.lbox); shape_circle(0, 0, p.size[5], p.rbox); }
    }
    int NA = (P.scene && P.scene != 4) ? Pn + (P.relation == 1 ? R : 0) : 0;
    if (P.scene == 2 || P.scene == 3) { P.rvo.ag.assign(NA, RvoAgent()); P.rvo.dt = (float)P.step_hz; for (int i = 0; i < NA; i++) P.rvo.ag[i].maxSpeed = i < Pn ? (float)P.peds[i].max_speed : 0.6f; }
    if (P.scene == 1) { P.sfm.ag.assign(NA, SfmAgent()); for (int i = 0; i < Pn; i++) P.sfm.ag[i].vmax = P.peds[i].max_speed; }
    return 0;
}

int port_reset(void* h, int n_obs, const double* obs, const double* robots, const double* peds, const int* traj_len, const double* traj,
               const int* trajv_len, const double* trajv, int ignore_obstacle) {   // img_env.cpp:162-292
    Port& P = *static_cast<Port*>(h);
    P.step_ = 0;
    P.obsm = P.stat; P.rvo.ob.clear(); P.sfm.obs.clear();
    for (int i = 0; i < n_obs; i++) {
        const double* d = obs + 11 * i; int shape = (int)d[0]; double size[4]; for (int k = 0; k < 4; k++) size[k] = f32(d[1 + k]);
        double yaw = quat_yaw(d[7], d[8], d[9], d[10]); Pts bb;
        if (shape == 0) shape_circle(size[0], size[1], size[2], bb); else shape_rect(size[0], size[1], size[2], size[3], bb);
        draw(P, P.obsm, t2_pose(d[5], d[6], yaw), bb, 0, 0);
        double c[4]; corners(size, shape, d[5], d[6], yaw, c);
        if (!ignore_obstacle) {
            P.sfm.obs.push_back(std::make_pair(d3(c[0], c[1]), d3(c[2], c[3])));
            F2 v[4] = {mk((float)c[0], (float)c[1]), mk((float)c[0], (float)c[3]), mk((float)c[2], (float)c[3]), mk((float)c[2], (float)c[1])};
            P.rvo.add_obstacle(v, 4);
        }
    }
    int Pn = (int)P.peds.size(), R = (int)P.robots.size(); size_t to = 0;
    for (int i = 0; i < Pn; i++) {
        const double* d = peds + 8 * i; Ped& p = P.peds[i];
        p.x = d[0]; p.y = d[1]; p.yaw = quat_yaw(d[2], d[3], d[4], d[5]); p.tidx = 0; p.traj.assign(traj + 3 * to, traj + 3 * (to + traj_len[i]));
        if (trajv_len && trajv) p.trajv.assign(trajv + 3 * to, trajv + 3 * (to + trajv_len[i]));
        to += traj_len[i];
        if (P.scene == 2 || P.scene == 3) P.rvo.ag[i].pos = mk((float)d[0], (float)d[1]);
        if (P.scene == 1) { SfmAgent& a = P.sfm.ag[i]; a.p = d3(d[0], d[1], 0); a.wp.clear(); a.wp.push_back(d3(d[6], d[7], 1)); for (int k = 0; k < traj_len[i]; k++) a.wp.push_back(d3(p.traj[3 * k], p.traj[3 * k + 1], p.traj[3 * k + 2])); a.dest = 0; a.lastdest = -1; a.front = 0; }
    }
    for (int i = 0; i < R; i++) {
        const double* d = robots + 8 * i; Robot& r = P.robots[i];
        r.x = d[0]; r.y = d[1]; r.yaw = quat_yaw(d[2], d[3], d[4], d[5]); r.l0v = r.l0w = 0; r.gx = d[6]; r.gy = d[7]; r.gyaw = r.yaw; r.coll = 0; r.arrive = false;
        if (P.relation == 1 && P.scene && P.scene != 4) {
            if (P.scene == 1) P.sfm.ag[Pn + i].p = d3(d[0], d[1], 1);
            else { P.rvo.ag[Pn + i].pos = mk((float)d[0], (float)d[1]); P.rvo.ag[Pn + i].vel = mk(0, 0); }
        }
    }
    if (P.scene == 2 || P.scene == 3) P.rvo.process();
    observe(P);
    return 0;
}

int port_step(void* h, const float* act, const uint8_t* alive) {   // img_env.cpp:304-359, 388-419, 421-525
    Port& P = *static_cast<Port*>(h);
    int Pn = (int)P.peds.size(), R = (int)P.robots.size();
    if (P.scene == 4) {   // _step_ped_dataset, img_env.cpp:361-386
        for (auto& p : P.peds) {
            int tl = (int)p.traj.size() / 3, ti = P.step_ >= tl ? tl - 1 : P.step_;
            double vx = p.trajv[3 * ti], vy = p.trajv[3 * ti + 1];
            p.lx = p.x; p.ly = p.y; p.lyaw = p.yaw;
            p.x = p.traj[3 * ti]; p.y = p.traj[3 * ti + 1]; p.yaw = atan2(vy, vx); p.vx = vx; p.vy = vy;
            if (p.shape == 2) {
                const double sl = 0.3; double md = sqrt((p.x - p.lx) * (p.x - p.lx) + (p.y - p.ly) * (p.y - p.ly));
                p.last_state = p.state; p.state = (int)((md + p.rem) / sl + p.last_state); p.rem = md + p.rem - (p.state - p.last_state) * sl; p.state %= 7;
                if (p.state == 0 || p.state == 4) { p.ll[0] = p.size[0]; p.ll[1] = p.size[1]; p.ll[2] = 0; p.rl[0] = p.size[3]; p.rl[1] = p.size[4]; p.rl[2] = 0; }
                else if (p.state == 1 || p.state == 3) { p.ll[0] = -sl / 2; p.rl[0] = sl / 2; } else if (p.state == 2) { p.ll[0] = -sl; p.rl[0] = sl; }
                else if (p.state == 5) { p.ll[0] = sl / 2; p.rl[0] = -sl / 2; } else if (p.state == 6) { p.ll[0] = sl; p.rl[0] = -sl; }
            }
        }
    } else if (P.scene) {
        std::vector<F2> goals, ps; std::vector<float> rs;
        for (auto& p : P.peds) {
            int tl = (int)p.traj.size() / 3;
            if (p.tidx < tl) { double gx = p.traj[3 * p.tidx], gy = p.traj[3 * p.tidx + 1]; if ((gx - p.x) * (gx - p.x) + (gy - p.y) * (gy - p.y) < 0.04) p.tidx++; }
            goals.push_back(mk((float)p.traj[3 * (p.tidx % tl)], (float)p.traj[3 * (p.tidx % tl) + 1]));
        }
        for (int j = 0; j < R; j++) {   // beep: ped_ca_p in {0,1} only (rand() draws are not reproducible across implementations)
            bool b = P.ped_ca_p >= 1.0 && (alive[j] ? act[3 * j + 2] : 0.f) > 0; P.robots[j].beep = b;
            ps.push_back(b ? mk((float)P.robots[j].x, (float)P.robots[j].y) : mk(0, 0)); rs.push_back(b ? (float)P.beep_r : 0.f);
        }
        if (P.scene == 2 || P.scene == 3) {
            for (int i = 0; i < Pn; i++) { F2 g = goals[i] - P.rvo.ag[i].pos; if (absSq(g) > 1.0f) g = unit(g); P.rvo.ag[i].pref = g; }
            P.rvo.do_step(P.scene == 3, ps, rs);
        } else { for (size_t i = 0; i < P.sfm.ag.size(); i++) P.sfm.forces((int)i); for (size_t i = 0; i < P.sfm.ag.size(); i++) P.sfm.move((int)i, P.step_hz); }
        for (int i = 0; i < Pn; i++) {
            Ped& p = P.peds[i]; p.lx = p.x; p.ly = p.y; p.lyaw = p.yaw;
            if (P.scene == 1) { p.x = P.sfm.ag[i].p.x; p.y = P.sfm.ag[i].p.y; p.vx = P.sfm.ag[i].v.x; p.vy = P.sfm.ag[i].v.y; }
            else { p.x = P.rvo.ag[i].pos.x; p.y = P.rvo.ag[i].pos.y; p.vx = P.rvo.ag[i].vel.x; p.vy = P.rvo.ag[i].vel.y; }
            p.yaw = 0.0;   // the node's uninitialised local (img_env.cpp:346-349) reads as 0.0 in the reference build
            if (p.shape == 2) {   // update_bbox, agent.cpp:696-735
                const double sl = 0.3; double md = sqrt((p.x - p.lx) * (p.x - p.lx) + (p.y - p.ly) * (p.y - p.ly));
                p.last_state = p.state; p.state = (int)((md + p.rem) / sl + p.last_state); p.rem = md + p.rem - (p.state - p.last_state) * sl; p.state %= 7;
                if (p.state == 0 || p.state == 4) { p.ll[0] = p.size[0]; p.ll[1] = p.size[1]; p.ll[2] = 0; p.rl[0] = p.size[3]; p.rl[1] = p.size[4]; p.rl[2] = 0; }
                else if (p.state == 1 || p.state == 3) { p.ll[0] = -sl / 2; p.rl[0] = sl / 2; } else if (p.state == 2) { p.ll[0] = -sl; p.rl[0] = sl; }
                else if (p.state == 5) { p.ll[0] = sl / 2; p.rl[0] = -sl / 2; } else if (p.state == 6) { p.ll[0] = sl; p.rl[0] = -sl; }
            }
        }
    }
    for (int j = 0; j < R; j++) {
        Robot& r = P.robots[j];
        if (alive[j]) {   // Agent::cmd, agent.cpp:186-283
            double v = act[3 * j], w = act[3 * j + 1], vy_ = act[3 * j + 2];
            limit(r.lv, v, r.l0v, r.l1v, P.step_hz); limit(r.lw, w, r.l0w, r.l1w, P.step_hz);
            r.l1v = r.l0v; r.l1w = r.l0w; r.l0v = v; r.l0w = w;
            bool arr = false; double ox = r.x, oy = r.y, oz = r.yaw, cc = 0; const double ch = 0.05;
            while (cc <= P.step_hz) {
                if (P.ktype == 0) { ox += v * ch * cos(oz); oy += v * ch * sin(oz); r.vx = v * cos(oz); r.vy = v * sin(oz); }
                else { double nx = ox + (v * ch * cos(oz) - vy_ * ch * sin(oz)), ny = oy + (v * ch * sin(oz) + vy_ * ch * cos(oz)); ox = nx; oy = ny; }
                oz += w * ch;
                if (sqrt((ox - r.gx) * (ox - r.gx) + (oy - r.gy) * (oy - r.gy)) <= 0.3) { arr = true; break; }
                cc += ch;
            }
            double th = r.yaw, dt = P.step_hz;
            if (w == 0) {
                if (P.ktype == 0) { r.x += v * dt * cos(th); r.y += v * dt * sin(th); }
                else { r.x += v * dt * cos(th) - vy_ * dt * sin(th); r.y += v * dt * sin(th) + vy_ * dt * cos(th); }
                r.yaw += w * dt;
            } else {
                double vw = v / w; r.x += -vw * sin(th) + vw * sin(th + w * dt); r.y += vw * cos(th) - vw * cos(th + w * dt);
                if (P.ktype == 1) { double yw = vy_ / w; r.x += -yw * cos(th) + yw * cos(th + w * dt); r.y += -yw * sin(th) + yw * sin(th + w * dt); }
                r.yaw += w * dt;
            }
            if (sqrt((r.x - r.gx) * (r.x - r.gx) + (r.y - r.gy) * (r.y - r.gy)) <= 0.3) arr = true;
            r.arrive = arr;
        }
        if (P.relation == 1 && P.scene && P.scene != 4) {
            if (P.scene == 1) P.sfm.ag[Pn + j].p = d3(r.x, r.y, 1);
            else { P.rvo.ag[Pn + j].pos = mk((float)r.x, (float)r.y); P.rvo.ag[Pn + j].vel = mk((float)r.vx, (float)r.vy); }
        }
    }
    P.step_ += 1;
    observe(P);
    return 0;
}

void port_get_states(void* h, uint8_t* view_maps, float* state, float* laser, int8_t* coll, uint8_t* arr, float* pedinfo) {
    Port& P = *static_cast<Port*>(h);
    for (size_t i = 0; i < P.robots.size(); i++) {
        if (view_maps && !P.robots[i].view.m.empty()) memcpy(view_maps + i * (size_t)P.vh * P.vw, P.robots[i].view.m.data(), (size_t)P.vh * P.vw);
        if (state) std::copy(P.st_state[i].begin(), P.st_state[i].end(), state + i * P.st_state[i].size());
        if (laser) std::copy(P.st_laser[i].begin(), P.st_laser[i].end(), laser + i * P.st_laser[i].size());
        if (coll) coll[i] = (int8_t)P.st_coll[i];
        if (arr) arr[i] = (uint8_t)P.st_arr[i];
        if (pedinfo) std::copy(P.st_ped[i].begin(), P.st_ped[i].end(), pedinfo + i * P.st_ped[i].size());
    }
}
int port_view_size(void* h) { Port& P = *static_cast<Port*>(h); return P.vh * P.vw; }
int port_laser_size(void* h) { Port& P = *static_cast<Port*>(h); return P.robots.empty() ? 0 : (int)P.st_laser[0].size(); }
void port_get_internal(void* h, double* rb, double* pd) {
    Port& P = *static_cast<Port*>(h);
    for (size_t i = 0; i < P.robots.size(); i++) { Robot& r = P.robots[i]; double* o = rb + 16 * i;
        double v[16] = {r.x, r.y, r.yaw, r.gx, r.gy, r.gyaw, r.l0v, r.l0w, r.l1v, r.l1w, r.vx, r.vy, (double)r.coll, (double)r.arrive, (double)r.beep, 0}; memcpy(o, v, sizeof v); }
    for (size_t i = 0; i < P.peds.size(); i++) { Ped& p = P.peds[i]; double* o = pd + 20 * i;
        double v[20] = {p.x, p.y, p.yaw, p.lx, p.ly, p.lyaw, p.vx, p.vy, (double)p.state, (double)p.last_state, p.rem, p.ll[0], p.ll[1], p.ll[2], p.rl[0], p.rl[1], p.rl[2], (double)p.tidx, 0, 0}; memcpy(o, v, sizeof v); }
}
void port_set_internal(void* h, const double* rb, const double* pd) {
    Port& P = *static_cast<Port*>(h);
    for (size_t i = 0; i < P.robots.size() && rb; i++) { Robot& r = P.robots[i]; const double* o = rb + 16 * i;
        r.x = o[0]; r.y = o[1]; r.yaw = o[2]; r.gx = o[3]; r.gy = o[4]; r.gyaw = o[5]; r.l0v = o[6]; r.l0w = o[7]; r.l1v = o[8]; r.l1w = o[9]; r.vx = o[10]; r.vy = o[11];
        r.coll = (int)o[12]; r.arrive = o[13] != 0; r.beep = (int)o[14]; }
    for (size_t i = 0; i < P.peds.size() && pd; i++) { Ped& p = P.peds[i]; const double* o = pd + 20 * i;
        p.x = o[0]; p.y = o[1]; p.yaw = o[2]; p.lx = o[3]; p.ly = o[4]; p.lyaw = o[5]; p.vx = o[6]; p.vy = o[7]; p.state = (int)o[8]; p.last_state = (int)o[9]; p.rem = o[10];
        p.ll[0] = o[11]; p.ll[1] = o[12]; p.ll[2] = o[13]; p.rl[0] = o[14]; p.rl[1] = o[15]; p.rl[2] = o[16]; p.tidx = (int)o[17]; }
}
int port_rvo_num_agents(void* h) { return (int)static_cast<Port*>(h)->rvo.ag.size(); }
void port_rvo_get(void* h, float* a) { Port& P = *static_cast<Port*>(h); for (size_t i = 0; i < P.rvo.ag.size(); i++) { a[4 * i] = P.rvo.ag[i].pos.x; a[4 * i + 1] = P.rvo.ag[i].pos.y; a[4 * i + 2] = P.rvo.ag[i].vel.x; a[4 * i + 3] = P.rvo.ag[i].vel.y; } }
void port_rvo_set(void* h, const float* a) { Port& P = *static_cast<Port*>(h); for (size_t i = 0; i < P.rvo.ag.size(); i++) { P.rvo.ag[i].pos = mk(a[4 * i], a[4 * i + 1]); P.rvo.ag[i].vel = mk(a[4 * i + 2], a[4 * i + 3]); } }
int port_rvo_num_obstacles(void* h) { return (int)static_cast<Port*>(h)->rvo.ob.size(); }
void port_rvo_get_obstacles(void* h, float* v) { Port& P = *static_cast<Port*>(h); for (size_t i = 0; i < P.rvo.ob.size(); i++) { Obst& o = P.rvo.ob[i]; float* d = v + 8 * i; d[0] = o.p.x; d[1] = o.p.y; d[2] = o.dir.x; d[3] = o.dir.y; d[4] = o.convex; d[5] = (float)o.next; d[6] = (float)o.prev; d[7] = 0; } }
int port_sfm_num_agents(void* h) { return (int)static_cast<Port*>(h)->sfm.ag.size(); }
void port_sfm_get(void* h, double* a) { Port& P = *static_cast<Port*>(h); for (size_t i = 0; i < P.sfm.ag.size(); i++) { SfmAgent& g = P.sfm.ag[i]; double* o = a + 12 * i;
    double v[12] = {g.p.x, g.p.y, g.p.z, g.v.x, g.v.y, g.v.z, g.vmax, (double)g.dest, (double)g.lastdest, (double)g.front, g.in_tree ? 1.0 : 0.0, 0}; memcpy(o, v, sizeof v); } }
void port_sfm_set(void* h, const double* a) { Port& P = *static_cast<Port*>(h); for (size_t i = 0; i < P.sfm.ag.size(); i++) { SfmAgent& g = P.sfm.ag[i]; const double* o = a + 12 * i;
    g.p = d3(o[0], o[1], o[2]); g.v = d3(o[3], o[4], o[5]); g.vmax = o[6]; g.dest = (int)o[7]; g.lastdest = (int)o[8]; g.front = (int)o[9]; g.in_tree = o[10] != 0; } }
void port_sfm_set_pv(void* h, const double* a) { port_sfm_set(h, a); }
void port_get_map(void* h, int which, uint8_t* out) { Port& P = *static_cast<Port*>(h); Grid& g = which == 0 ? P.stat : (which == 1 ? P.obsm : P.pedm); memcpy(out, g.m.data(), g.m.size()); }
}  // extern "C"
