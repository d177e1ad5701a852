// Reset-time build of a scene's RVO obstacle set ON THE DEVICE: the vertex ring of every reset object
// (RVOSimulator::addObstacle, RVOSimulator.cpp:130-168, fed by rvoscene.h:19-26 with the four corners of the box spanned by
// the object's two rotated corners) and the BSP over their edges (KdTree::buildObstacleTree / buildObstacleTreeRecursive,
// KdTree.cpp:119-257), float32 like the reference.  One warp per scene:
//   * the split edge of a node is the one that minimises (max(left, right), min(left, right)) over all edges, first such
//     edge in list order -- lanes take one candidate each, scan the other edges with the reference's early exit against the
//     best pair found so far, and the warp keeps the lexicographic minimum of (pair, position);
//   * the other edges are then distributed left / right in list order, edges that straddle the split line are cut and the
//     new vertex is appended to the ring (ids in list order: ballots give every lane its ordered output slots);
//   * the reference's recursion (node, left subtree, right subtree: node ids are pre-order) runs from an explicit frame stack;
//     edge lists live in a per-scene arena used as a stack.
// Output = the arrays the solver reads (orca.cuh): rvo_verts [n][8], rvo_nodes [n][4] (edge, left, right, parent),
// rvo_nodeseg [n][4], rvo_counts {n_verts, root}.  An arena / vertex overflow is counted in Dev::counters[1] and leaves the
// scene without obstacles for the solver (the tests assert the counter stays zero).
#pragma once
#include "state.cuh"
#include "orca.cuh"

#define RVO_STACK 128           // pending right subtrees along one root-to-leaf path
struct RvoFrame { int off, len, parent, is_right; };

__device__ __forceinline__ V2 rvo_pt(const float* verts, int i) { return v2(verts[8 * i], verts[8 * i + 1]); }

__global__ void __launch_bounds__(32) k_rvo_build(Dev d, const int* scene_ids, int ignore_obstacle) {
    const Cfg& c = d.c;
    if (d.n_dev && (int)blockIdx.x >= *d.n_dev) return;
    const int s = scene_ids ? scene_ids[blockIdx.x] : blockIdx.x;
    const int lane = threadIdx.x;
    float* verts = d.rvo_verts + (size_t)s * d.max_verts * 8;
    int* nodes = d.rvo_nodes + (size_t)s * d.max_verts * 4;
    float* nseg = d.rvo_nodeseg + (size_t)s * d.max_verts * 4;
    int* arena = d.rvo_arena + (size_t)s * d.rvo_arena_len;
    __shared__ RvoFrame stack[RVO_STACK];
    const float EPS = 0.00001f;
    // ---- vertex rings: object k -> vertices 4k .. 4k+3 = (ax, ay), (ax, by), (bx, by), (bx, ay)   (rvoscene.h:19-26)
    const int n_obj = ignore_obstacle ? 0 : d.sfm_nobs[s];
    int n_verts = 4 * n_obj;
    if (n_verts > d.max_verts || n_verts > d.rvo_arena_len) {
        if (lane == 0) { atomicAdd(d.counters + 1, 1ull); d.rvo_counts[2 * s] = 0; d.rvo_counts[2 * s + 1] = -1; }
        return;
    }
    for (int k = lane; k < n_obj; k += 32) {
        const double* sg = d.sfm_obs + ((size_t)s * c.max_obs + k) * 4;
        const V2 p[4] = {v2((float)sg[0], (float)sg[1]), v2((float)sg[0], (float)sg[3]), v2((float)sg[2], (float)sg[3]), v2((float)sg[2], (float)sg[1])};
        for (int i = 0; i < 4; i++) {
            float* o = verts + 8 * (4 * k + i);
            const V2 nx = p[(i + 1) & 3], pv = p[(i + 3) & 3];
            const V2 dir = unit(nx - p[i]);
            o[0] = p[i].x; o[1] = p[i].y; o[2] = dir.x; o[3] = dir.y;
            o[4] = side_of(pv, p[i], nx) >= 0.0f ? 1.f : 0.f;          // convex corner
            o[5] = (float)(4 * k + ((i + 1) & 3)); o[6] = (float)(4 * k + ((i + 3) & 3)); o[7] = 0.f;
        }
    }
    for (int k = lane; k < n_verts; k += 32) arena[k] = k;
    __syncwarp();
    int n_nodes = 0, top = n_verts, sp = 0, root = -1;
    bool failed = false;
    if (n_verts > 0) { if (lane == 0) stack[0] = RvoFrame{0, n_verts, -1, 0}; sp = 1; }
    __syncwarp();
    while (sp > 0 && !failed) {
        const RvoFrame f = stack[--sp];
        __syncwarp();
        const int* L = arena + f.off;
        const int m = f.len;
        top = f.off + f.len;                       // everything above this frame's list belongs to finished subtrees
        // ---- the split edge (KdTree.cpp:138-176)
        unsigned long long best = ~0ull;           // max(l, r) << 42 | min(l, r) << 21 | position
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            unsigned long long mine = ~0ull;
            if (i < m) {
                const int I1 = L[i], I2 = (int)verts[8 * I1 + 5];
                const V2 a = rvo_pt(verts, I1), b = rvo_pt(verts, I2);
                const unsigned long long bound = best >> 21;                       // the pair to beat
                unsigned l = 0, r = 0; bool alive = true;
                for (int j = 0; j < m; ++j) {
                    if (j == i) continue;
                    const int J1 = L[j], J2 = (int)verts[8 * J1 + 5];
                    const float s1 = side_of(a, b, rvo_pt(verts, J1)), s2 = side_of(a, b, rvo_pt(verts, J2));
                    if (s1 >= -EPS && s2 >= -EPS) ++l; else if (s1 <= EPS && s2 <= EPS) ++r; else { ++l; ++r; }
                    const unsigned long long pair = ((unsigned long long)max(l, r) << 21) | min(l, r);
                    if (pair >= bound) { alive = false; break; }                   // cannot beat the best any more
                }
                if (alive) mine = ((unsigned long long)max(l, r) << 42) | ((unsigned long long)min(l, r) << 21) | (unsigned)i;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, mine, o); mine = v < mine ? v : mine; }
            best = mine < best ? mine : best;
        }
        const int isplit = (int)(best & 0x1FFFFFu);
        const int I1 = L[isplit], I2 = (int)verts[8 * I1 + 5];
        const V2 a = rvo_pt(verts, I1), b = rvo_pt(verts, I2);
        // ---- this node (pre-order id), hooked into its parent
        const int me = n_nodes++;
        if (me >= d.max_verts) { failed = true; break; }
        if (lane == 0) {
            nodes[4 * me] = I1; nodes[4 * me + 1] = -1; nodes[4 * me + 2] = -1; nodes[4 * me + 3] = f.parent;
            nseg[4 * me] = a.x; nseg[4 * me + 1] = a.y; nseg[4 * me + 2] = b.x; nseg[4 * me + 3] = b.y;
            if (f.parent >= 0) nodes[4 * f.parent + 1 + f.is_right] = me;
        }
        if (f.parent < 0) root = me;
        // ---- distribute the other edges (KdTree.cpp:181-249): two passes, count then write, both in list order
        int cl = 0, cr = 0, cs = 0;
        for (int pass = 0; pass < 2; pass++) {
            int left_off = 0, right_off = 0;
            if (pass == 1) {
                if (top + cl + cr > d.rvo_arena_len || n_verts + cs > d.max_verts) { failed = true; break; }
                right_off = top; left_off = top + cr;      // right list BELOW the left one: the left subtree is built first and allocates above itself
            }
            int wl = 0, wr = 0, ws = 0;
            for (int j0 = 0; j0 < m; j0 += 32) {
                const int j = j0 + lane;
                int kind = 0;                       // 1 left, 2 right, 3 split with J1 left, 4 split with J1 right
                int J1 = 0, J2 = 0; float s1 = 0.f;
                if (j < m && j != isplit) {
                    J1 = L[j]; J2 = (int)verts[8 * J1 + 5];
                    s1 = side_of(a, b, rvo_pt(verts, J1));
                    const float s2 = side_of(a, b, rvo_pt(verts, J2));
                    if (s1 >= -EPS && s2 >= -EPS) kind = 1; else if (s1 <= EPS && s2 <= EPS) kind = 2; else kind = s1 > 0.0f ? 3 : 4;
                }
                const unsigned ml = __ballot_sync(0xffffffffu, kind == 1 || kind >= 3), mr = __ballot_sync(0xffffffffu, kind == 2 || kind >= 3);
                const unsigned ms = __ballot_sync(0xffffffffu, kind >= 3);
                const unsigned below = (1u << lane) - 1u;
                if (pass == 1 && kind) {
                    const int pl = left_off + wl + __popc(ml & below), pr = right_off + wr + __popc(mr & below);
                    if (kind == 1) arena[pl] = J1;
                    else if (kind == 2) arena[pr] = J1;
                    else {          // cut the edge J1 -> J2 where it crosses the split line; the new vertex follows J1
                        const int id = n_verts + ws + __popc(ms & below);
                        const V2 p1 = rvo_pt(verts, J1), p2 = rvo_pt(verts, J2);
                        const float t = cross(b - a, p1 - a) / cross(b - a, p1 - p2);
                        const V2 dj = p2 - p1;
                        float* o = verts + 8 * id;
                        o[0] = p1.x + t * dj.x; o[1] = p1.y + t * dj.y; o[2] = verts[8 * J1 + 2]; o[3] = verts[8 * J1 + 3];
                        o[4] = 1.f; o[5] = (float)J2; o[6] = (float)J1; o[7] = 0.f;
                        verts[8 * J1 + 5] = (float)id; verts[8 * J2 + 6] = (float)id;
                        if (kind == 3) { arena[pl] = J1; arena[pr] = id; } else { arena[pr] = J1; arena[pl] = id; }
                    }
                }
                wl += __popc(ml); wr += __popc(mr); ws += __popc(ms);
            }
            if (pass == 0) { cl = wl; cr = wr; cs = ws; }
        }
        if (failed) break;
        __syncwarp();
        n_verts += cs;
        // right subtree after the whole left subtree: push it first
        if (cr > 0) { if (sp >= RVO_STACK) { failed = true; break; } if (lane == 0) stack[sp] = RvoFrame{top, cr, me, 1}; sp++; }
        if (cl > 0) { if (sp >= RVO_STACK) { failed = true; break; } if (lane == 0) stack[sp] = RvoFrame{top + cr, cl, me, 0}; sp++; }
        __syncwarp();
    }
    if (lane == 0) {
        if (failed) { atomicAdd(d.counters + 1, 1ull); d.rvo_counts[2 * s] = 0; d.rvo_counts[2 * s + 1] = -1; }
        else { d.rvo_counts[2 * s] = n_verts; d.rvo_counts[2 * s + 1] = root; }
    }
}
