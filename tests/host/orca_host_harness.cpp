// Test harness: img_env_b200/csrc/orca.cuh compiled for the HOST (the CUDA qualifiers and the handful of intrinsics it uses are
// stubbed below), so that the ORCA / ERVO solver the GPU runs can be replayed on the CPU against the reference node
// (tests/test_orca_host_cpu.py).  Input (text, stdin): NA P dt ervo | NA x (px py vx vy) | n_verts root | verts[n][8] | nodes[n][4] |
// P x (goal_x goal_y max_speed) | n_beeps x (bx by br).  Output: one line "vx vy" per pedestrian.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { unsigned o = *p; if (o == c) *p = v; return o; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
using std::max;
using std::min;
#include "orca.cuh"

int main() {
    int NA, P, ervo; float dt;
    if (scanf("%d %d %f %d", &NA, &P, &dt, &ervo) != 4) return 1;
    std::vector<V2> pos(NA), vel(NA);
    for (int a = 0; a < NA; a++) if (scanf("%f %f %f %f", &pos[a].x, &pos[a].y, &vel[a].x, &vel[a].y) != 4) return 1;
    int nv, root;
    if (scanf("%d %d", &nv, &root) != 2) return 1;
    std::vector<float> verts(8 * (size_t)std::max(nv, 1)), seg(4 * (size_t)std::max(nv, 1));
    std::vector<int> nodes(4 * (size_t)std::max(nv, 1));
    for (int k = 0; k < 8 * nv; k++) if (scanf("%f", &verts[k]) != 1) return 1;
    for (int k = 0; k < 4 * nv; k++) if (scanf("%d", &nodes[k]) != 1) return 1;
    for (int k = 0; k < nv; k++) {      // the end points of every node's edge (rvotree.cuh writes the same into rvo_nodeseg)
        const int e = nodes[4 * k], nx = (int)verts[8 * e + 5];
        seg[4 * k] = verts[8 * e]; seg[4 * k + 1] = verts[8 * e + 1]; seg[4 * k + 2] = verts[8 * nx]; seg[4 * k + 3] = verts[8 * nx + 1];
    }
    std::vector<V2> goal(P); std::vector<float> vmax(P);
    for (int a = 0; a < P; a++) if (scanf("%f %f %f", &goal[a].x, &goal[a].y, &vmax[a]) != 3) return 1;
    int nb;
    if (scanf("%d", &nb) != 1) return 1;
    std::vector<V2> bp(std::max(nb, 1)); std::vector<float> br(std::max(nb, 1));
    for (int b = 0; b < nb; b++) if (scanf("%f %f %f", &bp[b].x, &bp[b].y, &br[b]) != 3) return 1;
    std::vector<unsigned short> head(agent_hash_size(NA)), next(NA);
    AgentHash hash; hash.mask = agent_hash_size(NA) - 1; hash.head = head.data(); hash.next = next.data();
    agent_hash_build(hash, pos.data(), NA, 0, 1);
    ObstacleSet ob; ob.verts = verts.data(); ob.nodes = nodes.data(); ob.node_seg = seg.data(); ob.root = nv > 0 ? root : -1;
    ob.n_cached = 0; ob.cache_nodes = nullptr; ob.cache_seg = nullptr;
    std::vector<unsigned char> scratch(orca_scratch_bytes(1) + 64), slabs((size_t)64 * ORCA_SLAB_BYTES);
    unsigned long long overflow = 0;
    for (int a = 0; a < P; a++) {
        unsigned cursor = 0;
        OrcaScratch sc = orca_scratch(scratch.data(), 0, 1);
        OrcaPool pool; pool.slabs = slabs.data(); pool.n_slabs = 64; pool.cursor = &cursor; pool.overflow = &overflow;
        V2 pref = goal[a] - pos[a];                                   // rvoscene.h:37-44, as k_dyn_solve does it
        if (norm2(pref) > 1.0f) pref = unit(pref);
        const V2 v = orca_new_velocity(a, pos.data(), vel.data(), hash, pref, vmax[a], dt, ob, sc, pool, 0xffffffffu, ervo != 0, nb, bp.data(), br.data());
        printf("%.9g %.9g\n", v.x, v.y);
    }
    if (overflow) { fprintf(stderr, "table overflow %llu\n", overflow); return 2; }
    return 0;
}
