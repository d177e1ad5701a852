// Dynamics stage, one CTA per scene: pedestrian waypoints + beeps + solver step + gait
// (ImgEnv::_step_ped_normal, img_env.cpp:304-359), then robot kinematics and the hand-over of robot
// poses to the pedestrian solver (ImgEnv::_step_robot, img_env.cpp:388-419).
// Solvers: ORCA / ERVO in orca.cuh; SFM (libpedsim) below, following
//   Tscene::moveAgents ped_scene.cpp:167-182, Tagent::computeForces ped_agent.cpp:498-507,
//   desiredForce :236-306, socialForce :316-404, obstacleForce :411-429, lookaheadForce :439-480,
//   move :519-571, Twaypoint::getForce ped_waypoint.cpp:81-135, Tobstacle::closestPoint
//   ped_obstacle.cpp:90-105, Tvector::lineIntersection ped_vector.cpp, adapter pedscene.h:18-93.
#pragma once
#include "state.cuh"
#include "kin.cuh"
#include "orca.cuh"
#include "sfmtree.cuh"


struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z = 0) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(double f, V3 a) { return v3(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, double f) { return v3(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ V3 operator/(V3 a, double d) { double f = 1 / d; return v3(f * a.x, f * a.y, f * a.z); }   // scaled(1/divisor)
__device__ __forceinline__ double lensq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double len(V3 a) { if (a.x == 0 && a.y == 0 && a.z == 0) return 0; return sqrt(lensq(a)); }
__device__ __forceinline__ V3 normalized(V3 a) { double l = len(a); if (l == 0) return v3(0, 0, 0); return v3(a.x / l, a.y / l, a.z / l); }
__device__ __forceinline__ double dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// sfm state record: p.xyz, v.xyz, vmax, dest, lastdest, deque_front, in_tree, pad
#define SFM_REC 12

struct SfmForces { V3 desired, social, obstacle, lookahead; };

// Tagent::computeForces for agent `self` (reads the pre-move state of all agents)
__device__ __forceinline__ QTreeView qt_view(const Dev& d, int s, const double* pos) {
    QTreeView t;
    t.n_nodes = d.qt_nodes + s; t.box = d.qt_box + (size_t)s * QT_MAX_NODES * 4; t.child0 = d.qt_child0 + (size_t)s * QT_MAX_NODES;
    t.count = d.qt_count + (size_t)s * QT_MAX_NODES; t.leaf = d.qt_leaf + (size_t)s * d.c.NA * QT_LEAVES; t.hash = d.qt_hash + (size_t)s * d.c.NA;
    t.pos = pos; t.na = d.c.NA;
    return t;
}

__device__ inline SfmForces sfm_forces(const Dev& d, int s, int self, double* rec_self, const double* recs, int n_agents) {
    const Cfg& c = d.c;
    const QTreeView qt = qt_view(d, s, nullptr);
    SfmForces F;
    V3 p = v3(rec_self[0], rec_self[1], rec_self[2]);
    V3 vself = v3(rec_self[3], rec_self[4], rec_self[5]);
    double vmax = rec_self[6];
    // ---- desiredForce (waypoint bookkeeping mutates dest/lastdest/front) ----
    V3 desiredDirection = v3(0, 0, 0);
    int dest = (int)rec_self[7], lastdest = (int)rec_self[8], front = (int)rec_self[9];
    int nwp = 0;
    const double* wp = nullptr;
    if (self < c.P) { nwp = 1 + d.traj_len[s * c.P + self]; wp = d.sfm_wp + ((size_t)(s * c.P + self) * (1 + c.max_traj)) * 3; }
    if (dest < 0 && nwp > 0) { dest = front; front = (front + 1) % nwp; }   // pop_front + push_back (BEHAVIOR_CIRCULAR)
    bool reached = false;
    if (dest >= 0) {
        // both branches (temporary TYPE_POINT waypoint / TYPE_NORMAL) evaluate the same expression
        V3 diff = v3(wp[3 * dest] - p.x, wp[3 * dest + 1] - p.y, 0);
        reached = len(diff) < wp[3 * dest + 2];
        desiredDirection = normalized(diff);
    }
    if (dest >= 0 && reached) { lastdest = dest; dest = -1; }
    rec_self[7] = dest; rec_self[8] = lastdest; rec_self[9] = front;
    F.desired = normalized(desiredDirection) * vmax;
    // ---- neighbours: agents held by quadtree leaves that intersect the 20 m query box (sfmtree.cuh) ----
    // lookaheadForce
    const double pi = 3.14159265;
    int lookforwardcount = 0;
    V3 soc = v3(0, 0, 0);
    for (int o = 0; o < n_agents; o++) {
        const double* ro = recs + (size_t)o * SFM_REC;
        if (o == self) continue;
        if (!qt_visible(qt, o, p.x, p.y, 20.0)) continue;      // scene->getNeighbors(p.x, p.y, 20) (ped_agent.cpp:499-500)
        V3 op = v3(ro[0], ro[1], ro[2]), ov = v3(ro[3], ro[4], ro[5]);
        {
            double distancex = op.x - p.x, distancey = op.y - p.y;
            double dist2 = (distancex * distancex + distancey * distancey);
            if (dist2 < 400) {
                double at2v = atan2(-desiredDirection.x, -desiredDirection.y);
                double at2d = atan2(-distancex, -distancey);
                double at2v2 = atan2(-ov.x, -ov.y);
                double sa = at2d - at2v;
                if (sa > pi) sa -= 2 * pi;
                if (sa < -pi) sa += 2 * pi;
                double vv = at2v - at2v2;
                if (vv > pi) vv -= 2 * pi;
                if (vv < -pi) vv += 2 * pi;
                if (fabs(vv) > 2.5) {
                    if ((sa < 0) && (sa > -0.3)) lookforwardcount--;
                    if ((sa > 0) && (sa < 0.3)) lookforwardcount++;
                }
            }
        }
        {   // socialForce (Moussaid-Helbing 2009 constants)
            const double lambdaImportance = 2.0, gamma = 0.35, n = 2, n_prime = 3;
            V3 diff = op - p;
            if (lensq(diff) > 64.0) continue;
            V3 diffDirection = normalized(diff);
            V3 velDiff = vself - ov;
            V3 interactionVector = lambdaImportance * velDiff + diffDirection;
            double interactionLength = len(interactionVector);
            V3 interactionDirection = interactionVector / interactionLength;
            double angleThis = atan2(interactionDirection.y, interactionDirection.x);
            double angleOther = atan2(diffDirection.y, diffDirection.x);
            double theta = angleOther - angleThis;
            if (theta > M_PI) theta -= 2 * M_PI;
            else if (theta <= -M_PI) theta += 2 * M_PI;
            int thetaSign = (theta == 0) ? (0) : (int)(theta / fabs(theta));
            double B = gamma * interactionLength;
            double forceVelocityAmount = -exp(-len(diff) / B - (n_prime * B * theta) * (n_prime * B * theta));
            double forceAngleAmount = -thetaSign * exp(-len(diff) / B - (n * B * theta) * (n * B * theta));
            V3 forceVelocity = forceVelocityAmount * interactionDirection;
            V3 forceAngle = forceAngleAmount * v3(-interactionDirection.y, interactionDirection.x, 0);
            soc = soc + (forceVelocity + forceAngle);
        }
    }
    V3 lf = v3(0, 0, 0);
    if (lookforwardcount < 0) { lf.x = 0.5f * desiredDirection.y; lf.y = 0.5f * -desiredDirection.x; }
    if (lookforwardcount > 0) { lf.x = 0.5f * -desiredDirection.y; lf.y = 0.5f * desiredDirection.x; }
    F.lookahead = lf;
    F.social = soc;
    // ---- obstacleForce: closest obstacle segment only ----
    V3 minDiff = v3(0, 0, 0);
    double minDistanceSquared = INFINITY;
    int nob = d.sfm_nobs[s];
    for (int o = 0; o < nob; o++) {
        const double* sg = d.sfm_obs + ((size_t)s * c.max_obs + o) * 4;
        V3 startPoint = v3(sg[0], sg[1]), endPoint = v3(sg[2], sg[3]);
        V3 relativeEndPoint = endPoint - startPoint;
        V3 relativePoint = p - startPoint;
        double lambda = dot3(relativePoint, relativeEndPoint) / lensq(relativeEndPoint);
        V3 closest;
        if (lambda <= 0) closest = startPoint;
        else if (lambda >= 1) closest = endPoint;
        else closest = startPoint + lambda * relativeEndPoint;
        V3 diff = p - closest;
        double distanceSquared = lensq(diff);
        if (distanceSquared < minDistanceSquared) { minDistanceSquared = distanceSquared; minDiff = diff; }
    }
    double distance = sqrt(minDistanceSquared) - 0.2;    // agentRadius
    double forceAmount = exp(-distance / 0.8);            // obstacleForceSigma
    F.obstacle = forceAmount * normalized(minDiff);
    return F;
}

// Tagent::move (without the tree notification)
__device__ inline void sfm_move(const Dev& d, int s, double* rec, const SfmForces& F, double h) {
    const Cfg& c = d.c;
    V3 p = v3(rec[0], rec[1], rec[2]), v = v3(rec[3], rec[4], rec[5]);
    double vmax = rec[6];
    V3 p_desired = p + v * h;
    int nob = d.sfm_nobs[s];
    for (int o = 0; o < nob; o++) {
        const double* sg = d.sfm_obs + ((size_t)s * c.max_obs + o) * 4;
        double s1x = p_desired.x - p.x, s1y = p_desired.y - p.y;
        double s2x = sg[2] - sg[0], s2y = sg[3] - sg[1];
        double ss = (-s1y * (p.x - sg[0]) + s1x * (p.y - sg[1])) / (-s2x * s1y + s1x * s2y);
        double tt = (s2x * (p.y - sg[1]) - s2y * (p.x - sg[0])) / (-s2x * s1y + s1x * s2y);
        if (ss >= 0 && ss <= 1 && tt >= 0 && tt <= 1) {
            V3 inter = v3(p.x + (tt * s1x), p.y + (tt * s1y), 0);
            p_desired = inter - normalized(v * h) * 0.1;
        }
    }
    p = p_desired;
    V3 a = 1.0 * F.desired + 2.1 * F.social + 1.0 * F.obstacle + 1.0 * F.lookahead + v3(0, 0, 0);
    v = 0.5 * v + a * h;
    if (len(v) > vmax) v = normalized(v) * vmax;
    rec[0] = p.x; rec[1] = p.y; rec[2] = p.z; rec[3] = v.x; rec[4] = v.y; rec[5] = v.z;
}

__device__ __forceinline__ double beep_uniform(unsigned long long seed, unsigned long long step, int s, int r) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (step * 1315423911ull + (unsigned long long)s * 2654435761ull + r + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// PedAgent::update_bbox, agent.cpp:696-735
__device__ inline void ped_gait(const Dev& d, int pi, int p) {
    if (d.ped_shape[p] != 2) return;
    const double step_len = 0.3;
    const double* sz = d.ped_size + 6 * p;
    double x = PDF(d, PD_X, pi), y = PDF(d, PD_Y, pi), lx = PDF(d, PD_LX, pi), ly = PDF(d, PD_LY, pi);
    double move_dist = sqrt((x - lx) * (x - lx) + (y - ly) * (y - ly));
    int state = (int)PDF(d, PD_GAIT, pi);
    double rem = PDF(d, PD_REM, pi);
    int last_state = state;
    state = (int)((move_dist + rem) / step_len + last_state);
    rem = move_dist + rem - (state - last_state) * step_len;
    state %= 7;
    PDF(d, PD_LGAIT, pi) = last_state; PDF(d, PD_GAIT, pi) = state; PDF(d, PD_REM, pi) = rem;
    if (state == 0 || state == 4) {
        PDF(d, PD_LLX, pi) = sz[0]; PDF(d, PD_LLY, pi) = sz[1]; PDF(d, PD_LLZ, pi) = 0;
        PDF(d, PD_RLX, pi) = sz[3]; PDF(d, PD_RLY, pi) = sz[4]; PDF(d, PD_RLZ, pi) = 0;
    } else if (state == 1 || state == 3) { PDF(d, PD_LLX, pi) = -step_len / 2; PDF(d, PD_RLX, pi) = step_len / 2; }
    else if (state == 2) { PDF(d, PD_LLX, pi) = -step_len; PDF(d, PD_RLX, pi) = step_len; }
    else if (state == 5) { PDF(d, PD_LLX, pi) = step_len / 2; PDF(d, PD_RLX, pi) = -step_len / 2; }
    else if (state == 6) { PDF(d, PD_LLX, pi) = step_len; PDF(d, PD_RLX, pi) = -step_len; }
}

// The dynamics stage runs as two kernels so that a scene's agents spread over several CTAs:
//   k_dyn_solve : grid = S * nblk CTAs; every CTA loads the scene's agents (pre-step pos/vel) into shared
//                 memory and solves its slice of Cfg::dyn_threads agents (ORCA/ERVO new velocity, or the SFM forces),
//                 writing only scratch (new velocities / forces) and per-agent waypoint bookkeeping;
//   k_dyn_apply : grid = S * nblk CTAs; Agent::update / Tagent::move, pedestrian pose + gait, then the robots
//                 (limiter, cmd, arrival) and setRobotPos.
// actions [S][R][3] float32 (v, w, v_y); alive [S][R] or NULL (-> library dones)
__device__ __forceinline__ bool robot_alive(const Dev& d, const uint8_t* alive, int idx) {
    return alive ? alive[idx] != 0 : RBF(d, RB_DONE, idx) == 0.0;
}
__host__ __device__ __forceinline__ int dyn_nblk(const Cfg& c) { return c.dyn_nblk; }
// Agents per CTA and CTAs per scene: as few CTAs per scene as DYN_MAX_THREADS allows (every CTA of a scene stages the same
// agents, hash and BSP nodes), threads rounded up to whole warps.  C4: 400 agents -> 2 CTAs x 224; C3: 21 -> 1 x 32.
inline void dyn_pick_block(Cfg& c) {
    int n = c.NA > c.R ? c.NA : c.R;
    if (c.scene_type == 4 && c.P > n) n = c.P;
    if (n < 1) n = 1;
    c.dyn_nblk = (n + DYN_MAX_THREADS - 1) / DYN_MAX_THREADS;
    c.dyn_threads = ((n + c.dyn_nblk - 1) / c.dyn_nblk + 31) / 32 * 32;
}

// shared memory of k_dyn_solve: agents' pos + vel, beeps, spatial hash (head table + chain), per-thread ORCA tables
__host__ __device__ inline size_t dyn_scratch_offset(const Cfg& c) {
    size_t off = (size_t)c.NA * 16 + (size_t)c.R * 8 + ((size_t)c.R + (c.R & 1)) * 4 + ((size_t)agent_hash_size(c.NA) + c.NA) * 2;
    return (off + 15) & ~(size_t)15;
}
__host__ __device__ inline int dyn_node_cache(const Dev& d) { return d.max_verts < ORCA_NODE_CACHE ? d.max_verts : ORCA_NODE_CACHE; }
inline size_t dyn_smem_bytes(const Dev& d) {
    const Cfg& c = d.c;
    return dyn_scratch_offset(c) + ((c.scene_type == 2 || c.scene_type == 3) ? orca_scratch_bytes(c.dyn_threads) + (size_t)dyn_node_cache(d) * 32 : 0) + 16;
}

__global__ void __launch_bounds__(DYN_MAX_THREADS, 2) k_dyn_solve(Dev d, const float* actions, const uint8_t* alive, int parity) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const Cfg& c = d.c;
    const int nblk = dyn_nblk(c);
    const int s = blockIdx.x / nblk, blk = blockIdx.x % nblk, tid = threadIdx.x, DT = c.dyn_threads;
    if (!(c.P > 0 && c.scene_type != 0 && c.scene_type != 4)) return;
    // shared memory: the scene's agents (position, velocity), beeps, spatial hash, then the per-thread ORCA tables
    V2* pos = reinterpret_cast<V2*>(dsm);
    V2* vel = pos + c.NA;
    V2* beep_p = vel + c.NA;
    float* beep_r = reinterpret_cast<float*>(beep_p + c.R);
    AgentHash hash;
    hash.mask = agent_hash_size(c.NA) - 1;
    hash.head = reinterpret_cast<unsigned short*>(beep_r + c.R + (c.R & 1));
    hash.next = hash.head + hash.mask + 1;
    unsigned char* scratch = dsm + dyn_scratch_offset(c);
    int4* node_cache = reinterpret_cast<int4*>(scratch + orca_scratch_bytes(DT));
    float4* seg_cache = reinterpret_cast<float4*>(node_cache + dyn_node_cache(d));
    const unsigned long long step = d.step_no[s];
    // beeps (img_env.cpp:323-342): robots' PRE-step poses; recomputed identically by every CTA of the scene
    for (int j = tid; j < c.R; j += DT) {
        int idx = s * c.R + j;
        bool is_beep = false;
        float v_y = robot_alive(d, alive, idx) ? actions[(size_t)idx * 3 + 2] : 0.f;
        if (beep_uniform(c.seed, step, s, j) < c.ped_ca_p && (double)v_y > 0) {
            beep_p[j] = v2((float)RBF(d, RB_X, idx), (float)RBF(d, RB_Y, idx));
            beep_r[j] = (float)c.beep_r;
            is_beep = true;
        }
        if (!is_beep) { beep_p[j] = v2(0.f, 0.f); beep_r[j] = 0.f; }
        if (blk == 0) RBF(d, RB_BEEP, idx) = is_beep ? 1.0 : 0.0;
    }
    // ERVO only looks at robots that beep: keep those, in robot order (the evacuation velocities are summed in that order)
    __shared__ int s_nbeep;
    __syncthreads();
    if (tid < 32) {
        int base = 0;
        for (int j0 = 0; j0 < c.R; j0 += 32) {
            const int j = j0 + tid;
            const bool on = j < c.R && beep_r[j] > 0.f;
            const V2 bp = j < c.R ? beep_p[j] : v2(0.f, 0.f);
            const float br = j < c.R ? beep_r[j] : 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            __syncwarp();
            if (on) { const int k = base + __popc(m & ((1u << tid) - 1u)); beep_p[k] = bp; beep_r[k] = br; }
            base += __popc(m);
            __syncwarp();
        }
        if (tid == 0) s_nbeep = base;
    }
    __syncthreads();
    const int a = blk * DT + tid;
    if (c.scene_type == 2 || c.scene_type == 3) {
        for (int k = tid; k < c.NA; k += DT) {
            pos[k] = v2(d.rvo_pos[((size_t)s * c.NA + k) * 2], d.rvo_pos[((size_t)s * c.NA + k) * 2 + 1]);
            vel[k] = v2(d.rvo_vel[((size_t)s * c.NA + k) * 2], d.rvo_vel[((size_t)s * c.NA + k) * 2 + 1]);
        }
        ObstacleSet ob;
        ob.verts = d.rvo_verts + (size_t)s * d.max_verts * 8;
        ob.nodes = d.rvo_nodes + (size_t)s * d.max_verts * 4;
        ob.node_seg = d.rvo_nodeseg + (size_t)s * d.max_verts * 4;
        ob.root = d.rvo_counts[2 * s + 1];
        ob.n_cached = min(d.rvo_counts[2 * s], dyn_node_cache(d));
        ob.cache_nodes = node_cache; ob.cache_seg = seg_cache;
        for (int k = tid; k < ob.n_cached; k += DT) {     // the obstacle BSP is walked by every agent: stage it
            node_cache[k] = __ldg(reinterpret_cast<const int4*>(ob.nodes) + k);
            seg_cache[k] = __ldg(reinterpret_cast<const float4*>(ob.node_seg) + k);
        }
        __syncthreads();
        agent_hash_build(hash, pos, c.NA, tid, DT);
        const unsigned warp_mask = __ballot_sync(0xffffffffu, a < c.NA);
        if (a >= c.NA) return;
        if (blockIdx.x == 0 && tid == 0) d.orca_cursor[parity ^ 1] = 0u;       // the next call's slab cursor
        V2 pref = v2(0.f, 0.f);
        float maxSpeed = 0.6f;
        if (a < c.P) {
            int pi = s * c.P + a;
            // waypoint cycling (img_env.cpp:306-319, agent.cpp:823-843); an index past the end of
            // trajectory_ is an out-of-bounds read in the node -> treated as "not arrived"
            int ti = (int)PDF(d, PD_TIDX, pi), tl = d.traj_len[pi];
            const double* tr = d.traj + ((size_t)pi * c.max_traj) * 3;
            double x = PDF(d, PD_X, pi), y = PDF(d, PD_Y, pi);
            if (ti < tl) {
                double gx = tr[3 * ti], gy = tr[3 * ti + 1];
                if ((gx - x) * (gx - x) + (gy - y) * (gy - y) < 0.04) ti += 1;
            }
            PDF(d, PD_TIDX, pi) = ti;
            const double* g = tr + 3 * (ti % tl);
            V2 goalVector = v2((float)g[0], (float)g[1]) - pos[a];     // rvoscene.h:37-44
            if (norm2(goalVector) > 1.0f) goalVector = unit(goalVector);
            pref = goalVector;
            maxSpeed = (float)d.ped_maxspeed[a];
        }
        OrcaScratch sc = orca_scratch(scratch, tid, DT);
        OrcaPool pool;
        pool.slabs = d.orca_pool; pool.n_slabs = d.orca_nslabs; pool.cursor = d.orca_cursor + parity; pool.overflow = d.counters;
        V2 nv = orca_new_velocity(a, pos, vel, hash, pref, maxSpeed, (float)c.step_hz, ob, sc, pool, warp_mask, c.scene_type == 3, s_nbeep, beep_p, beep_r);
        d.rvo_nvel[((size_t)s * c.NA + a) * 2] = nv.x; d.rvo_nvel[((size_t)s * c.NA + a) * 2 + 1] = nv.y;
    } else if (c.scene_type == 1) {
        if (a >= c.NA) return;
        double* recs = d.sfm + (size_t)s * c.NA * SFM_REC;
        if (a < c.P) {   // waypoint cycling of the PedAgent wrapper still runs (unused by the SFM adapter's step)
            int pi = s * c.P + a;
            int ti = (int)PDF(d, PD_TIDX, pi), tl = d.traj_len[pi];
            const double* tr = d.traj + ((size_t)pi * c.max_traj) * 3;
            double x = PDF(d, PD_X, pi), y = PDF(d, PD_Y, pi);
            if (ti < tl) {
                double gx = tr[3 * ti], gy = tr[3 * ti + 1];
                if ((gx - x) * (gx - x) + (gy - y) * (gy - y) < 0.04) ti += 1;
            }
            PDF(d, PD_TIDX, pi) = ti;
        }
        SfmForces F = sfm_forces(d, s, a, recs + (size_t)a * SFM_REC, recs, c.NA);
        double* o = d.sfm_force + ((size_t)s * c.NA + a) * 12;
        o[0] = F.desired.x; o[1] = F.desired.y; o[2] = F.desired.z; o[3] = F.social.x; o[4] = F.social.y; o[5] = F.social.z;
        o[6] = F.obstacle.x; o[7] = F.obstacle.y; o[8] = F.obstacle.z; o[9] = F.lookahead.x; o[10] = F.lookahead.y; o[11] = F.lookahead.z;
    }
}

__global__ void __launch_bounds__(DYN_MAX_THREADS) k_dyn_apply(Dev d, const float* actions, const uint8_t* alive, int ped_yaw_mode) {
    const Cfg& c = d.c;
    const int nblk = dyn_nblk(c);
    const int s = blockIdx.x / nblk, blk = blockIdx.x % nblk, tid = threadIdx.x, DT = c.dyn_threads;
    const int a = blk * DT + tid;
    if (c.P > 0 && c.scene_type != 0 && a < c.NA) {
        if (c.scene_type == 2 || c.scene_type == 3) {   // Agent::update
            const size_t o = ((size_t)s * c.NA + a) * 2;
            V2 nv = v2(d.rvo_nvel[o], d.rvo_nvel[o + 1]);
            V2 np = v2(d.rvo_pos[o], d.rvo_pos[o + 1]) + nv * (float)c.step_hz;
            d.rvo_pos[o] = np.x; d.rvo_pos[o + 1] = np.y;
            d.rvo_vel[o] = nv.x; d.rvo_vel[o + 1] = nv.y;
            if (a < c.P) {   // getNewPosAndVel + set_position + update_bbox (img_env.cpp:344-358)
                int pi = s * c.P + a;
                PDF(d, PD_LX, pi) = PDF(d, PD_X, pi); PDF(d, PD_LY, pi) = PDF(d, PD_Y, pi); PDF(d, PD_LYAW, pi) = PDF(d, PD_YAW, pi);
                PDF(d, PD_X, pi) = (double)np.x; PDF(d, PD_Y, pi) = (double)np.y;
                PDF(d, PD_VX, pi) = (double)nv.x; PDF(d, PD_VY, pi) = (double)nv.y;
                if (ped_yaw_mode == 1) PDF(d, PD_YAW, pi) = 0.0;
                else if (ped_yaw_mode == 2) PDF(d, PD_YAW, pi) = atan2((double)nv.y, (double)nv.x);
                ped_gait(d, pi, a);
            }
        } else if (c.scene_type == 1) {
            double* rec = d.sfm + ((size_t)s * c.NA + a) * SFM_REC;
            const double* f = d.sfm_force + ((size_t)s * c.NA + a) * 12;
            SfmForces F;
            F.desired = v3(f[0], f[1], f[2]); F.social = v3(f[3], f[4], f[5]); F.obstacle = v3(f[6], f[7], f[8]); F.lookahead = v3(f[9], f[10], f[11]);
            // scene->moveAgent(this) runs inside Tagent::move, agent by agent: a leaf that splits while agent i moves re-sorts its
            // members by their CURRENT positions -- already moved for j < i, not yet for j > i.  k_sfm_tree therefore gets both
            // the pre-move position (robots: the pose setRobotPos wrote after the previous step) and the post-move one.
            double* tp = d.sfm_newpos + ((size_t)s * c.NA + a) * 4;
            tp[2] = rec[0]; tp[3] = rec[1];
            sfm_move(d, s, rec, F, c.step_hz);
            tp[0] = rec[0]; tp[1] = rec[1];
            if (a < c.P) {
                int pi = s * c.P + a;
                PDF(d, PD_LX, pi) = PDF(d, PD_X, pi); PDF(d, PD_LY, pi) = PDF(d, PD_Y, pi); PDF(d, PD_LYAW, pi) = PDF(d, PD_YAW, pi);
                PDF(d, PD_X, pi) = rec[0]; PDF(d, PD_Y, pi) = rec[1];
                PDF(d, PD_VX, pi) = rec[3]; PDF(d, PD_VY, pi) = rec[4];
                if (ped_yaw_mode == 1) PDF(d, PD_YAW, pi) = 0.0;
                else if (ped_yaw_mode == 2) PDF(d, PD_YAW, pi) = atan2(rec[4], rec[3]);
                ped_gait(d, pi, a);
            }
        }
    }
    if (c.scene_type == 4 && a < c.P) {   // ImgEnv::_step_ped_dataset (img_env.cpp:361-386): replay trajectory[step_]
        const int pi = s * c.P + a;
        const int tl = d.traj_len[pi];
        const unsigned long long st = d.step_no[s];
        const int ti = st >= (unsigned long long)tl ? tl - 1 : (int)st;
        const double* tp = d.traj + ((size_t)pi * c.max_traj + ti) * 3;
        const double* tv = d.traj_v + ((size_t)pi * c.max_traj + ti) * 3;
        PDF(d, PD_LX, pi) = PDF(d, PD_X, pi); PDF(d, PD_LY, pi) = PDF(d, PD_Y, pi); PDF(d, PD_LYAW, pi) = PDF(d, PD_YAW, pi);
        PDF(d, PD_X, pi) = tp[0]; PDF(d, PD_Y, pi) = tp[1]; PDF(d, PD_YAW, pi) = atan2(tv[1], tv[0]);
        PDF(d, PD_VX, pi) = tv[0]; PDF(d, PD_VY, pi) = tv[1];
        ped_gait(d, pi, a);
    }
    if (d.rec_T > 0 && a < c.P && d.step_no[s] < (unsigned long long)d.rec_T) {   // eps_res_msg.peds_res (img_env.cpp:355-357)
        const int pi = s * c.P + a;
        double* o = d.rec_pd + (((size_t)s * d.rec_T + (size_t)d.step_no[s]) * c.P + a) * 5;
        o[0] = PDF(d, PD_X, pi); o[1] = PDF(d, PD_Y, pi); o[2] = PDF(d, PD_YAW, pi); o[3] = PDF(d, PD_VX, pi); o[4] = PDF(d, PD_VY, pi);
    }
    // ---- robots (img_env.cpp:388-419). Robot j is handled by the thread that (if the robots are solver
    // agents) also updated solver agent P + j, so the solver write above is ordered before the overwrite below.
    const bool robots_in_solver = c.relation == 1 && c.NA > 0;
    const int j = robots_in_solver ? a - c.P : a;
    if (j >= 0 && j < c.R) {
        int idx = s * c.R + j;
        if (robot_alive(d, alive, idx)) {
            RobotKin rk;
            rk.x = RBF(d, RB_X, idx); rk.y = RBF(d, RB_Y, idx); rk.yaw = RBF(d, RB_YAW, idx);
            rk.gx = RBF(d, RB_GX, idx); rk.gy = RBF(d, RB_GY, idx);
            rk.l0v = RBF(d, RB_L0V, idx); rk.l0w = RBF(d, RB_L0W, idx); rk.l1v = RBF(d, RB_L1V, idx); rk.l1w = RBF(d, RB_L1W, idx);
            rk.vx = RBF(d, RB_VX, idx); rk.vy = RBF(d, RB_VY, idx);
            const float* ac = actions + (size_t)idx * 3;
            robot_cmd(rk, d.lim_v[j], d.lim_w[j], c.ktype, c.step_hz, c.control_hz, (double)ac[0], (double)ac[1], (double)ac[2]);
            RBF(d, RB_X, idx) = rk.x; RBF(d, RB_Y, idx) = rk.y; RBF(d, RB_YAW, idx) = rk.yaw;
            RBF(d, RB_L0V, idx) = rk.l0v; RBF(d, RB_L0W, idx) = rk.l0w; RBF(d, RB_L1V, idx) = rk.l1v; RBF(d, RB_L1W, idx) = rk.l1w;
            RBF(d, RB_VX, idx) = rk.vx; RBF(d, RB_VY, idx) = rk.vy;
            RBF(d, RB_ARR, idx) = rk.arrive ? 1.0 : 0.0;
        }
        if (d.rec_T > 0 && d.step_no[s] < (unsigned long long)d.rec_T) {   // eps_res_msg.robots_res (img_env.cpp:397-408; alive steps only)
            double* o = d.rec_rb + (((size_t)s * d.rec_T + (size_t)d.step_no[s]) * c.R + j) * 6;
            const float* ac = actions + (size_t)idx * 3;
            o[0] = RBF(d, RB_X, idx); o[1] = RBF(d, RB_Y, idx); o[2] = RBF(d, RB_YAW, idx); o[3] = (double)ac[0]; o[4] = (double)ac[1];
            o[5] = robot_alive(d, alive, idx) ? 1.0 : 0.0;
        }
        if (robots_in_solver) {   // setRobotPos
            int ag = c.P + j;
            if (c.scene_type == 2 || c.scene_type == 3) {
                d.rvo_pos[((size_t)s * c.NA + ag) * 2] = (float)RBF(d, RB_X, idx); d.rvo_pos[((size_t)s * c.NA + ag) * 2 + 1] = (float)RBF(d, RB_Y, idx);
                d.rvo_vel[((size_t)s * c.NA + ag) * 2] = (float)RBF(d, RB_VX, idx); d.rvo_vel[((size_t)s * c.NA + ag) * 2 + 1] = (float)RBF(d, RB_VY, idx);
            } else if (c.scene_type == 1) {
                double* rec = d.sfm + ((size_t)s * c.NA + ag) * SFM_REC;
                rec[0] = RBF(d, RB_X, idx); rec[1] = RBF(d, RB_Y, idx); rec[2] = 1.0;   // pedscene.h:53-56
            }
        }
    }
}

// Tscene::moveAgent for every agent in index order (the order Tscene::moveAgents moves them).  The walk is
// sequential per scene and data dependent, so one scene = one warp with one working lane (32 scenes in a warp
// would serialise their divergent walks); the in-tree flags are written back by all lanes.
__global__ void __launch_bounds__(32) k_sfm_tree(Dev d) {
    const int s = blockIdx.x;
    double* np = d.sfm_newpos + (size_t)s * d.c.NA * 4;            // [NA][4] new x, y, old x, y
    double* cur = d.sfm_treepos + (size_t)s * d.c.NA * 2;          // positions as Tagent::getPosition() would return them right now
    for (int a = threadIdx.x; a < d.c.NA; a += 32) { cur[2 * a] = np[4 * a + 2]; cur[2 * a + 1] = np[4 * a + 3]; }
    __syncwarp();
    const QTreeView t = qt_view(d, s, cur);
    if (threadIdx.x == 0) for (int a = 0; a < d.c.NA; a++) { cur[2 * a] = np[4 * a]; cur[2 * a + 1] = np[4 * a + 1]; qt_move(t, a); }
    __syncwarp();
    __threadfence_block();
    for (int a = threadIdx.x; a < d.c.NA; a += 32) d.sfm[((size_t)s * d.c.NA + a) * SFM_REC + 10] = qt_in_tree(t, a) ? 1.0 : 0.0;
}


