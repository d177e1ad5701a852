// ORCA / emotional-ORCA pedestrian update for a block of agents of one scene.
//
// What the reference computes (RVO2 v2.0 as vendored in src/3rdparty/ervo_ros, driven by rvoscene.h:36-66 /
// ervoscene.h:13-22): per agent the <= 10 nearest agents within 0.5 m (Agent.cpp:50-61, 795-818), the obstacle edges
// within (5 maxSpeed + 0.5) m that face the agent, nearest first (KdTree.cpp:322-353, Agent.cpp:820-838), one velocity
// half-plane per edge / neighbour (Agent.cpp:437-761), a 2-D linear program over them with a 3-D fallback
// (Agent.cpp:845-1001), and for ERVO a unit "evacuation" velocity away from every beeping robot in range, added after the
// program and not re-clipped (Agent.cpp:63-69).  float32, no FMA contraction, like the x86 build of the reference.
//
// How it is organised here (one thread per agent, Cfg::dyn_threads agents per CTA, everything per-thread in SHARED memory,
// strided by the thread index so that a warp's accesses never conflict; no local-memory arrays):
//   * agent neighbours: the scene's agents are binned into a shared-memory spatial hash with 0.5 m cells
//     (= neighborDist); an agent walks the 3 x 3 cells around it and keeps its 10 nearest by (distance^2, index) --
//     the list RVO2's k-d tree walk produces, with exact distance ties resolved by index instead of tree order;
//   * obstacle neighbours: the scene's BSP is walked WITHOUT a stack (parent links; near side, node, far side -- the
//     reference's recursion order, so that equal distances, e.g. the two edges meeting in the nearest corner, keep the
//     reference's order), each facing edge in range goes into a list sorted by distance;
//   * half-planes are written to a per-thread line table as they are built; the linear programs read lines through an
//     accessor, and the 3-D fallback evaluates its projected lines on the fly from the table instead of building a
//     second array.
// Capacity: the shared-memory tables hold ORCA_FAST obstacle neighbours / lines per agent, which covers every ordinary
// scene; an agent that needs more takes a slab from a global pool (up to ORCA_OBST_CAP obstacle neighbours).  Only when the
// pool is exhausted or that cap is exceeded does an agent keep just its nearest edges; this is counted in Dev::counters[0]
// (imgenv_debug_counters) and the tests assert it stays zero.
#pragma once
#include "state.cuh"

#define RVO_EPS 0.00001f
#define ORCA_NEIGH_CAP 10
#ifndef ORCA_FAST
#define ORCA_FAST 12                    // obstacle neighbours / lines per agent kept in shared memory
#endif
#define ORCA_OBST_CAP 256               // obstacle neighbours per agent in total (beyond ORCA_FAST: in a pool slab)
#define ORCA_LINE_CAP (ORCA_NEIGH_CAP + ORCA_OBST_CAP)
#ifndef ORCA_NODE_CACHE
#define ORCA_NODE_CACHE 576             // BSP nodes (32 bytes each) staged in shared memory per CTA
#endif
#define ORCA_SLAB_BYTES ((ORCA_LINE_CAP - ORCA_FAST) * 16 + (ORCA_OBST_CAP - ORCA_FAST) * 8)
#ifndef DYN_MAX_THREADS
#define DYN_MAX_THREADS 224      // agents per CTA of the dynamics kernels at most (imgenv.cu picks the count per configuration)
#endif

struct V2 { float x, y; };
__device__ __forceinline__ V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator-(V2 a) { return v2(-a.x, -a.y); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ V2 scaled_inv(V2 a, float s) { const float inv = 1.0f / s; return v2(a.x * inv, a.y * inv); }   // Vector2::operator/
__device__ __forceinline__ float norm2(V2 a) { return dot(a, a); }
__device__ __forceinline__ float norm(V2 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ V2 unit(V2 a) { return scaled_inv(a, norm(a)); }
__device__ __forceinline__ V2 perp_ccw(V2 a) { return v2(-a.y, a.x); }
__device__ __forceinline__ float sq(float a) { return a * a; }
// > 0 when c lies to the left of the directed line a -> b
__device__ __forceinline__ float side_of(V2 a, V2 b, V2 c) { return cross(a - c, b - a); }
__device__ __forceinline__ float seg_dist2(V2 a, V2 b, V2 c) {
    const float r = dot(c - a, b - a) / norm2(b - a);
    if (r < 0.0f) return norm2(c - a);
    if (r > 1.0f) return norm2(c - b);
    return norm2(c - (a + r * (b - a)));
}

// One scene's obstacle polygons: vertex ring + BSP over its edges (built at reset: RVOSimulator::addObstacle,
// KdTree::buildObstacleTreeRecursive)
struct ObstacleSet {
    const float* verts;   // [n][8] px, py, edge direction x, y, convex, next, prev, 0
    const int* nodes;     // [n][4] edge (= its first vertex), left child, right child, parent
    const float* node_seg;// [n][4] the end points of the node's edge
    const int4* cache_nodes; const float4* cache_seg; int n_cached;   // the first nodes of both arrays, staged in shared memory
    int root;
    __device__ __forceinline__ V2 point(int i) const { return v2(verts[8 * i], verts[8 * i + 1]); }
    __device__ __forceinline__ V2 dir(int i) const { return v2(verts[8 * i + 2], verts[8 * i + 3]); }
    __device__ __forceinline__ bool convex(int i) const { return verts[8 * i + 4] != 0.f; }
    __device__ __forceinline__ int next(int i) const { return (int)verts[8 * i + 5]; }
    __device__ __forceinline__ int prev(int i) const { return (int)verts[8 * i + 6]; }
};

// ---- per-thread tables: the first ORCA_FAST entries in shared memory (element j of thread t at [j * threads + t]),
//      the rest in a slab taken from a global pool on first use ---------------------------------------------------------
struct OrcaPool { unsigned char* slabs; int n_slabs; unsigned* cursor; unsigned long long* overflow; };
struct OrcaScratch {
    float4* line;         // [ORCA_FAST] point.xy, direction.xy; the permitted side is the LEFT of the direction
    float* obst_d2;       // [ORCA_FAST]
    int* obst_id;         // [ORCA_FAST]
    float* nb_d2;         // [ORCA_NEIGH_CAP]
    int* nb_id;           // [ORCA_NEIGH_CAP]
    unsigned char* slab;  // nullptr until needed
    int stride;           // threads per CTA
    __device__ __forceinline__ bool need_slab(const OrcaPool& pool) {
        if (slab) return true;
        const unsigned k = atomicAdd(pool.cursor, 1u);
        if (k >= (unsigned)pool.n_slabs) return false;
        slab = pool.slabs + (size_t)k * ORCA_SLAB_BYTES;
        return true;
    }
    __device__ __forceinline__ float4* slab_line(int j) const { return reinterpret_cast<float4*>(slab) + (j - ORCA_FAST); }
    __device__ __forceinline__ float* slab_d2(int i) const { return reinterpret_cast<float*>(slab + (ORCA_LINE_CAP - ORCA_FAST) * 16) + 2 * (i - ORCA_FAST); }
    __device__ __forceinline__ float4 get_line(int j) const { return j < ORCA_FAST ? line[j * stride] : *slab_line(j); }
    __device__ __forceinline__ bool put_line(int j, float4 v, const OrcaPool& pool) {
        if (j < ORCA_FAST) { line[j * stride] = v; return true; }
        if (j >= ORCA_LINE_CAP || !need_slab(pool)) return false;
        *slab_line(j) = v; return true;
    }
    __device__ __forceinline__ float od2(int i) const { return i < ORCA_FAST ? obst_d2[i * stride] : slab_d2(i)[0]; }
    __device__ __forceinline__ int oid(int i) const { return i < ORCA_FAST ? obst_id[i * stride] : __float_as_int(slab_d2(i)[1]); }
    __device__ __forceinline__ void oput(int i, float d2, int id) {
        if (i < ORCA_FAST) { obst_d2[i * stride] = d2; obst_id[i * stride] = id; }
        else { float* q = slab_d2(i); q[0] = d2; q[1] = __int_as_float(id); }
    }
};
#define SLOT(j) ((j) * sc.stride)
__host__ __device__ inline size_t orca_scratch_bytes(int threads) {
    return (size_t)threads * (ORCA_FAST * 16 + ORCA_FAST * 8 + ORCA_NEIGH_CAP * 8);
}
__device__ __forceinline__ OrcaScratch orca_scratch(unsigned char* base, int tid, int threads) {
    OrcaScratch s;
    s.line = reinterpret_cast<float4*>(base) + tid;
    float* f = reinterpret_cast<float*>(base + (size_t)threads * ORCA_FAST * 16);
    s.obst_d2 = f + tid; s.obst_id = reinterpret_cast<int*>(f + threads * ORCA_FAST) + tid;
    f += 2 * threads * ORCA_FAST;
    s.nb_d2 = f + tid; s.nb_id = reinterpret_cast<int*>(f + threads * ORCA_NEIGH_CAP) + tid;
    s.slab = nullptr; s.stride = threads;
    return s;
}

// ---- spatial hash of the scene's agents (cell = neighborDist) ------------------------------------------------------------
struct AgentHash {
    unsigned short* head;     // [size] first agent of the bucket, 0xFFFF = empty
    unsigned short* next;     // [n_agents]
    int mask;                 // size - 1 (power of two)
};
#define HASH_CELL_INV 2.0f    // 1 / 0.5 m
__device__ __forceinline__ int hash_cell(float x) { return (int)floorf(x * HASH_CELL_INV); }
__device__ __forceinline__ int hash_bucket(int cx, int cy, int mask) { return (int)(((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u)) & mask; }
__host__ __device__ inline int agent_hash_size(int n_agents) { int s = 64; while (s < 2 * n_agents) s <<= 1; return s; }
// all threads of the CTA; pos = the scene's agent positions in shared memory
__device__ __forceinline__ void agent_hash_build(const AgentHash& h, const V2* pos, int n_agents, int tid, int n_threads) {
    for (int k = tid; k <= h.mask; k += n_threads) h.head[k] = 0xFFFFu;
    __syncthreads();
    for (int a = tid; a < n_agents; a += n_threads) {
        // push front with a 16-bit exchange emulated on the containing 32-bit word
        const int b = hash_bucket(hash_cell(pos[a].x), hash_cell(pos[a].y), h.mask);
        unsigned* w = reinterpret_cast<unsigned*>(h.head) + (b >> 1);
        const int sh = (b & 1) * 16;
        unsigned old = *w, assumed;
        do { assumed = old; old = atomicCAS(w, assumed, (assumed & ~(0xFFFFu << sh)) | ((unsigned)a << sh)); } while (old != assumed);
        h.next[a] = (unsigned short)((old >> sh) & 0xFFFFu);
    }
    __syncthreads();
}

// The <= 10 nearest agents within 0.5 m, sorted by (distance^2, index).  Returns their number.
__device__ __forceinline__ int gather_agent_neighbours(const OrcaScratch& sc, const AgentHash& h, const V2* pos, int self, float range) {
    const V2 p = pos[self];
    const int cx = hash_cell(p.x), cy = hash_cell(p.y);
    const float range2 = sq(range);
    int n = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const int qx = cx + dx, qy = cy + dy;
            for (int a = h.head[hash_bucket(qx, qy, h.mask)]; a != 0xFFFF; a = h.next[a]) {
                if (a == self) continue;
                const V2 q = pos[a];
                if (hash_cell(q.x) != qx || hash_cell(q.y) != qy) continue;      // another cell that shares the bucket
                const float d2 = norm2(p - q);
                if (!(d2 < range2)) continue;
                // sorted insertion; when the table is full the farthest entry drops out.  Ties: lower index first.
                if (n == ORCA_NEIGH_CAP) {
                    const float ld = sc.nb_d2[SLOT(n - 1)]; const int li = sc.nb_id[SLOT(n - 1)];
                    if (!(d2 < ld || (d2 == ld && a < li))) continue;
                } else n++;
                int i = n - 1;
                while (i > 0) {
                    const float pd = sc.nb_d2[SLOT(i - 1)]; const int pi = sc.nb_id[SLOT(i - 1)];
                    if (!(d2 < pd || (d2 == pd && a < pi))) break;
                    sc.nb_d2[SLOT(i)] = pd; sc.nb_id[SLOT(i)] = pi; --i;
                }
                sc.nb_d2[SLOT(i)] = d2; sc.nb_id[SLOT(i)] = a;
            }
        }
    return n;
}

// Facing obstacle edges within `range2`, nearest first; equal distances keep the order of the reference's recursive walk
// (near subtree, node, far subtree).  The walk needs no stack: every node knows its parent, and carries its edge's two end
// points so that a visit costs one round of independent loads.
__device__ __forceinline__ int gather_obstacle_neighbours(OrcaScratch& sc, const OrcaPool& pool, const ObstacleSet& ob, V2 p, float range2) {
    int n = 0;
    int node = ob.root, came_from = -2;        // -2: arrived from the parent; otherwise the child we return from
    while (node >= 0) {
        int4 nd; float4 sg;                      // (edge, left, right, parent), the edge's end points
        if (node < ob.n_cached) { nd = ob.cache_nodes[node]; sg = ob.cache_seg[node]; }
        else { nd = __ldg(reinterpret_cast<const int4*>(ob.nodes) + node); sg = __ldg(reinterpret_cast<const float4*>(ob.node_seg) + node); }
        const V2 a = v2(sg.x, sg.y), b = v2(sg.z, sg.w);
        const float side = side_of(a, b, p);
        const int near_child = side >= 0.0f ? nd.y : nd.z;
        const int far_child = side >= 0.0f ? nd.z : nd.y;
        if (came_from == -2 && near_child >= 0) { node = near_child; continue; }              // descend the near side first
        if (came_from == -2 || came_from == near_child) {
            // the near side is done: this node, then (if the edge's line is in range) the far side
            const float line_d2 = sq(side) / norm2(b - a);
            if (line_d2 < range2) {
                if (side < 0.0f) {          // the agent is on the right of the edge: it can see it
                    const float d2 = seg_dist2(a, b, p);
                    if (d2 < range2) {
                        bool room = n < ORCA_FAST || (n < ORCA_OBST_CAP && sc.need_slab(pool));
                        if (!room) {        // table (or pool) full: keep the nearest, and say so
                            atomicAdd(pool.overflow, 1ull);
                            if (d2 < sc.od2(n - 1)) { n--; room = true; }
                        }
                        if (room) {
                            int i = n++;
                            while (i > 0 && d2 < sc.od2(i - 1)) { sc.oput(i, sc.od2(i - 1), sc.oid(i - 1)); --i; }
                            sc.oput(i, d2, nd.x);
                        }
                    }
                }
                if (far_child >= 0) { came_from = -2; node = far_child; continue; }
            }
        }
        came_from = node; node = nd.w;                                                        // back to the parent
    }
    return n;
}

// ---- half-planes and the linear programs ------------------------------------------------------------------------------------
struct Line { V2 p, d; };
__device__ __forceinline__ Line make_line(V2 p, V2 d) { Line l; l.p = p; l.d = d; return l; }

// direct view of the thread's line table
struct TableLines {
    const OrcaScratch* t;
    __device__ __forceinline__ Line get(int j) const { const float4 v = t->get_line(j); return make_line(v2(v.x, v.y), v2(v.z, v.w)); }
};
// The line set of the relaxed (3-D) program for pivot line `piv`: the obstacle lines unchanged, every earlier agent line j
// replaced by the bisector between it and the pivot (Agent.cpp:960-987).  Evaluated on the fly.  A line the reference drops
// (parallel, same direction) becomes an inert line (zero direction): it is never violated and never clips.
struct RelaxedLines {
    const OrcaScratch* t; int n_obst; Line piv;
    __device__ __forceinline__ Line get(int j) const {
        const float4 v = t->get_line(j);
        const Line lj = make_line(v2(v.x, v.y), v2(v.z, v.w));
        if (j < n_obst) return lj;
        const float determinant = cross(piv.d, lj.d);
        V2 point;
        if (fabsf(determinant) <= RVO_EPS) {
            if (dot(piv.d, lj.d) > 0.0f) return make_line(v2(0.f, 0.f), v2(0.f, 0.f));
            point = 0.5f * (piv.p + lj.p);
        } else {
            point = piv.p + (cross(lj.d, piv.p - lj.p) / determinant) * piv.d;
        }
        return make_line(point, unit(lj.d - piv.d));
    }
};

// Best point on line `idx` inside the disc of radius `radius` and inside lines [0, idx).  false: infeasible.
template <class Lines>
__device__ __forceinline__ bool solve_on_line(const Lines& L, int idx, float radius, V2 target, bool target_is_direction, V2& result) {
    const Line me = L.get(idx);
    const float along = dot(me.p, me.d);
    const float disc = sq(along) + sq(radius) - norm2(me.p);
    if (disc < 0.0f) return false;
    const float root = sqrtf(disc);
    float t_lo = -along - root, t_hi = -along + root;
    for (int i = 0; i < idx; ++i) {
        const Line o = L.get(i);
        const float denom = cross(me.d, o.d);
        const float numer = cross(o.d, me.p - o.p);
        if (fabsf(denom) <= RVO_EPS) {       // (nearly) parallel
            if (numer < 0.0f) return false;
            continue;
        }
        const float t = numer / denom;
        if (denom >= 0.0f) t_hi = fminf(t_hi, t); else t_lo = fmaxf(t_lo, t);
        if (t_lo > t_hi) return false;
    }
    float t;
    if (target_is_direction) t = dot(target, me.d) > 0.0f ? t_hi : t_lo;
    else { t = dot(me.d, target - me.p); t = t < t_lo ? t_lo : (t > t_hi ? t_hi : t); }
    result = me.p + t * me.d;
    return true;
}
// Incremental 2-D program over lines [0, n).  Returns n on success, else the index of the first line that cannot be met.
template <class Lines>
__device__ __forceinline__ int solve_plane(const Lines& L, int n, float radius, V2 target, bool target_is_direction, V2& result) {
    if (target_is_direction) result = target * radius;
    else if (norm2(target) > sq(radius)) result = unit(target) * radius;
    else result = target;
    for (int i = 0; i < n; ++i) {
        const Line l = L.get(i);
        if (cross(l.d, l.p - result) > 0.0f) {            // result violates line i
            const V2 keep = result;
            if (!solve_on_line(L, i, radius, target, target_is_direction, result)) { result = keep; return i; }
        }
    }
    return n;
}
// Fallback when the 2-D program is infeasible from line `first_bad` on: minimise the largest violation of the agent lines
// (the obstacle lines stay hard).
__device__ __forceinline__ void solve_relaxed(const OrcaScratch* table, int n, int n_obst, int first_bad, float radius, V2& result) {
    const TableLines T{table};
    float worst = 0.0f;
    for (int i = first_bad; i < n; ++i) {
        const Line li = T.get(i);
        if (cross(li.d, li.p - result) > worst) {
            const RelaxedLines R{table, n_obst, li};
            const V2 keep = result;
            // the relaxed program holds ALL obstacle lines (also when line i is itself an obstacle line that the disc of
            // admissible speeds cannot meet) and the bisectors of the agent lines before i (Agent.cpp:960-987)
            const int n_proj = i > n_obst ? i : n_obst;
            if (solve_plane(R, n_proj, radius, perp_ccw(li.d), true, result) < n_proj) result = keep;
            worst = cross(li.d, li.p - result);
        }
    }
}

// Velocity half-plane induced by one obstacle edge (Agent.cpp:442-683), or nothing.  `covered`: an earlier line already
// keeps the agent away from the whole edge.
__device__ __forceinline__ bool obstacle_line(const ObstacleSet& ob, int e1, V2 pos, V2 vel, float radius, float inv_horizon,
                                              const OrcaScratch& sc, int n_lines, Line& out) {
    int v1 = e1, v2i = ob.next(e1);
    const V2 rel1 = ob.point(v1) - pos, rel2 = ob.point(v2i) - pos;
    for (int j = 0; j < n_lines; ++j) {
        const float4 q = sc.get_line(j);
        const V2 lp = v2(q.x, q.y), ld = v2(q.z, q.w);
        if (cross(inv_horizon * rel1 - lp, ld) - inv_horizon * radius >= -RVO_EPS &&
            cross(inv_horizon * rel2 - lp, ld) - inv_horizon * radius >= -RVO_EPS) return false;
    }
    const float d1 = norm2(rel1), d2 = norm2(rel2), r2 = sq(radius);
    const V2 edge = ob.point(v2i) - ob.point(v1);
    const float s = dot(-rel1, edge) / norm2(edge);          // parameter of the foot of the perpendicular on the edge
    const float dline = norm2(-rel1 - s * edge);
    // already touching: the constraint passes through the current velocity space origin
    if (s < 0.0f && d1 <= r2) {
        if (!ob.convex(v1)) return false;
        out = make_line(v2(0.f, 0.f), unit(perp_ccw(rel1)));
        return true;
    }
    if (s > 1.0f && d2 <= r2) {
        if (!(ob.convex(v2i) && cross(rel2, ob.dir(v2i)) >= 0.0f)) return false;
        out = make_line(v2(0.f, 0.f), unit(perp_ccw(rel2)));
        return true;
    }
    if (s >= 0.0f && s < 1.0f && dline <= r2) {
        out = make_line(v2(0.f, 0.f), -ob.dir(v1));
        return true;
    }
    // the two legs of the velocity obstacle (tangents from the agent to discs of its radius around the end points)
    auto left_tangent = [&](V2 rel, float dd) { const float leg = sqrtf(dd - r2); return scaled_inv(v2(rel.x * leg - rel.y * radius, rel.x * radius + rel.y * leg), dd); };
    auto right_tangent = [&](V2 rel, float dd) { const float leg = sqrtf(dd - r2); return scaled_inv(v2(rel.x * leg + rel.y * radius, -rel.x * radius + rel.y * leg), dd); };
    V2 left_leg, right_leg;
    if (s < 0.0f && dline <= r2) {            // seen obliquely: only the first end point matters
        if (!ob.convex(v1)) return false;
        v2i = v1;
        left_leg = left_tangent(rel1, d1); right_leg = right_tangent(rel1, d1);
    } else if (s > 1.0f && dline <= r2) {     // ... only the second
        if (!ob.convex(v2i)) return false;
        v1 = v2i;
        left_leg = left_tangent(rel2, d2); right_leg = right_tangent(rel2, d2);
    } else {
        left_leg = ob.convex(v1) ? left_tangent(rel1, d1) : -ob.dir(e1);
        right_leg = ob.convex(v2i) ? right_tangent(rel2, d2) : ob.dir(e1);
    }
    // a leg that points into the neighbouring edge is replaced by that edge's direction and never yields a constraint
    bool left_foreign = false, right_foreign = false;
    const int before = ob.prev(v1);
    if (ob.convex(v1) && cross(left_leg, -ob.dir(before)) >= 0.0f) { left_leg = -ob.dir(before); left_foreign = true; }
    if (ob.convex(v2i) && cross(right_leg, ob.dir(v2i)) <= 0.0f) { right_leg = ob.dir(v2i); right_foreign = true; }
    // project the current velocity on the truncated velocity obstacle: cut-off segment, left leg or right leg
    const V2 cut_l = inv_horizon * (ob.point(v1) - pos), cut_r = inv_horizon * (ob.point(v2i) - pos);
    const V2 cut = cut_r - cut_l;
    const bool single = v1 == v2i;
    const float t = single ? 0.5f : dot(vel - cut_l, cut) / norm2(cut);
    const float t_l = dot(vel - cut_l, left_leg), t_r = dot(vel - cut_r, right_leg);
    if ((t < 0.0f && t_l < 0.0f) || (single && t_l < 0.0f && t_r < 0.0f)) {      // nearest: the left cut-off corner
        const V2 w = unit(vel - cut_l);
        out = make_line(cut_l + radius * inv_horizon * w, v2(w.y, -w.x));
        return true;
    }
    if (t > 1.0f && t_r < 0.0f) {                                                // ... the right cut-off corner
        const V2 w = unit(vel - cut_r);
        out = make_line(cut_r + radius * inv_horizon * w, v2(w.y, -w.x));
        return true;
    }
    const float INF = __int_as_float(0x7f800000);
    const float dist_cut = (t < 0.0f || t > 1.0f || single) ? INF : norm2(vel - (cut_l + t * cut));
    const float dist_l = t_l < 0.0f ? INF : norm2(vel - (cut_l + t_l * left_leg));
    const float dist_r = t_r < 0.0f ? INF : norm2(vel - (cut_r + t_r * right_leg));
    V2 dirn, anchor;
    if (dist_cut <= dist_l && dist_cut <= dist_r) { dirn = -ob.dir(v1); anchor = cut_l; }
    else if (dist_l <= dist_r) { if (left_foreign) return false; dirn = left_leg; anchor = cut_l; }
    else { if (right_foreign) return false; dirn = -right_leg; anchor = cut_r; }
    out = make_line(anchor + radius * inv_horizon * perp_ccw(dirn), dirn);
    return true;
}

// Velocity half-plane induced by another agent (reciprocal: each takes half of the avoidance), Agent.cpp:689-761
__device__ __forceinline__ Line agent_line(V2 pos, V2 vel, V2 opos, V2 ovel, float radius, float inv_horizon, float time_step) {
    const V2 rel_p = opos - pos, rel_v = vel - ovel;
    const float d2 = norm2(rel_p);
    const float rr = radius + radius, rr2 = sq(rr);
    V2 dirn, u;
    if (d2 > rr2) {                              // not colliding: velocity obstacle truncated at the horizon
        const V2 w = rel_v - inv_horizon * rel_p;
        const float w2 = norm2(w);
        const float dp = dot(w, rel_p);
        if (dp < 0.0f && sq(dp) > rr2 * w2) {    // nearest point on the cut-off circle
            const float wl = sqrtf(w2);
            const V2 uw = scaled_inv(w, wl);
            dirn = v2(uw.y, -uw.x);
            u = (rr * inv_horizon - wl) * uw;
        } else {                                 // nearest point on a leg
            const float leg = sqrtf(d2 - rr2);
            if (cross(rel_p, w) > 0.0f) dirn = scaled_inv(v2(rel_p.x * leg - rel_p.y * rr, rel_p.x * rr + rel_p.y * leg), d2);
            else dirn = -scaled_inv(v2(rel_p.x * leg + rel_p.y * rr, -rel_p.x * rr + rel_p.y * leg), d2);
            u = dot(rel_v, dirn) * dirn - rel_v;
        }
    } else {                                     // colliding: leave the overlap within one time step
        const float inv_dt = 1.0f / time_step;
        const V2 w = rel_v - inv_dt * rel_p;
        const float wl = norm(w);
        const V2 uw = scaled_inv(w, wl);
        dirn = v2(uw.y, -uw.x);
        u = (rr * inv_dt - wl) * uw;
    }
    return make_line(vel + 0.5f * u, dirn);
}

// New velocity of agent `self` (RVOSimulator::doStep / ERVOSimulator::doStep for one agent; the caller applies
// Agent::update after every agent of the scene is done).  pos / vel: the scene's agents in shared memory.
__device__ __forceinline__ V2 orca_new_velocity(int self, const V2* pos, const V2* vel, const AgentHash& hash, V2 pref_velocity, float max_speed,
                                                float time_step, const ObstacleSet& ob, OrcaScratch& sc, const OrcaPool& pool, unsigned warp_mask,
                                                bool ervo, int n_beeps, const V2* beep_p, const float* beep_r) {
    const float radius = 0.5f, neighbour_dist = 0.5f, horizon = 5.f, horizon_obst = 5.f;   // rvoscene.h:53-66
    const V2 p = pos[self], v = vel[self];
    // Every stage below is a data-dependent loop; the lanes of the warp (warp_mask = those that own an agent) are brought
    // back together after each one, otherwise they drift apart for the rest of the kernel and execute one by one.
    const int n_obst_nb = gather_obstacle_neighbours(sc, pool, ob, p, sq(horizon_obst * max_speed + radius));
    __syncwarp(warp_mask);
    const int n_nb = gather_agent_neighbours(sc, hash, pos, self, neighbour_dist);
    __syncwarp(warp_mask);
    int n_lines = 0;
    const float inv_horizon_obst = 1.0f / horizon_obst;
    for (int i = 0; i < n_obst_nb; ++i) {
        Line l;
        if (obstacle_line(ob, sc.oid(i), p, v, radius, inv_horizon_obst, sc, n_lines, l)) {
            if (sc.put_line(n_lines, make_float4(l.p.x, l.p.y, l.d.x, l.d.y), pool)) n_lines++;
            else atomicAdd(pool.overflow, 1ull);
        }
    }
    __syncwarp(warp_mask);
    const int n_obst_lines = n_lines;
    const float inv_horizon = 1.0f / horizon;
    for (int i = 0; i < n_nb; ++i) {
        const int o = sc.nb_id[SLOT(i)];
        const Line l = agent_line(p, v, pos[o], vel[o], radius, inv_horizon, time_step);
        if (sc.put_line(n_lines, make_float4(l.p.x, l.p.y, l.d.x, l.d.y), pool)) n_lines++;
        else atomicAdd(pool.overflow, 1ull);
    }
    __syncwarp(warp_mask);
    V2 result;
    const int bad = solve_plane(TableLines{&sc}, n_lines, max_speed, pref_velocity, false, result);
    __syncwarp(warp_mask);
    if (bad < n_lines) solve_relaxed(&sc, n_lines, n_obst_lines, bad, max_speed, result);
    __syncwarp(warp_mask);
    if (ervo) {   // evacuation velocity: away from every beeping robot within its beep radius
        for (int b = 0; b < n_beeps; b++) {
            const V2 away = p - beep_p[b];
            const float dist = norm(away);
            if (dist > beep_r[b] || dist < 1e-4) continue;
            result = result + unit(away);
        }
    }
    return result;
}
