// Emulation of libpedsim's agent quadtree (Ttree, ped_tree.cpp:19-231; Tscene::placeAgent/moveAgent/
// getNeighbors, ped_scene.cpp:184-247) as flat per-scene arrays.  The tree only matters as *semantics*:
// which agents a getNeighbors(x, y, 20) query returns.  The reference's behaviour that has to be reproduced:
//   * the root box is x in [0,10], y in [10,20] (pedscene.h:19), leaves split when they hold > 8 agents and
//     never merge (cut() is never called);
//   * addAgent descends with inclusive comparisons in the order tree3, tree1, tree2, tree4, so an agent on a
//     split line lands in several leaves and treehash keeps the last one;
//   * moveAgent (called from Tagent::move) re-inserts an agent that left its leaf's box from the root and
//     THEN erases it from the old leaf - so an agent outside the root box that descends to the same leaf
//     vanishes from every query, and re-appears for one step when it descends to another leaf.
// Host (initial tree, built with the same code) and device (one thread per scene after the move phase).
#pragma once
#include "tfmath.cuh"

#define QT_MAX_NODES 1024
#define QT_LEAVES 4          // an agent can sit in up to 4 leaves (exactly on both split lines)

struct QTreeView {
    int* n_nodes;            // [1]
    double* box;             // [QT_MAX_NODES][4] x, y, w, h
    int* child0;             // [QT_MAX_NODES] index of tree1 (children tree1..tree4 are consecutive), -1 = leaf
    int* count;              // [QT_MAX_NODES] agents held (leaves only)
    int* leaf;               // [NA][QT_LEAVES] leaves holding the agent, -1 = empty slot
    int* hash;               // [NA] treehash
    const double* pos;       // [NA][2] current agent positions (x, y)
    int na;
};

HD bool qt_member(const QTreeView& t, int a, int n) {
    for (int k = 0; k < QT_LEAVES; k++) if (t.leaf[a * QT_LEAVES + k] == n) return true;
    return false;
}
HD void qt_erase(const QTreeView& t, int a, int n) {
    for (int k = 0; k < QT_LEAVES; k++) if (t.leaf[a * QT_LEAVES + k] == n) { t.leaf[a * QT_LEAVES + k] = -1; t.count[n]--; return; }
}
HD bool qt_in_tree(const QTreeView& t, int a) {
    for (int k = 0; k < QT_LEAVES; k++) if (t.leaf[a * QT_LEAVES + k] >= 0) return true;
    return false;
}
// Ttree::intersects
HD bool qt_intersects(const double* b, double px, double py, double pr) {
    return ((px + pr) > b[0]) && ((px - pr) < (b[0] + b[2])) && ((py + pr) > b[1]) && ((py - pr) < (b[1] + b[3]));
}
// Would getNeighbors(px, py, pr) return agent o?  (a leaf is reached iff its box intersects: child boxes nest)
HD bool qt_visible(const QTreeView& t, int o, double px, double py, double pr) {
    for (int k = 0; k < QT_LEAVES; k++) {
        int n = t.leaf[o * QT_LEAVES + k];
        if (n >= 0 && (n == 0 || qt_intersects(t.box + 4 * n, px, py, pr))) return true;
    }
    return false;
}

// Ttree::addAgent on the subtree rooted at `start`.  The reference recurses (addAgent -> split -> addAgent ...);
// here the same depth-first order is run from an explicit frame stack (no device recursion):
//   ADD(a, n)    insert agent a below node n (children visited in the reference's order tree3, tree1, tree2, tree4)
//   ERASE(a, n)  agents.erase(a) on the node that just split, after a was pushed down
//   CONT(n, m)   continue moving the members of the splitting node n, from agent index m (set order = index order)
#define QT_STACK 160
__host__ __device__ inline void qt_add(const QTreeView& t, int a0, int start) {
    int fk[QT_STACK], fa[QT_STACK], fn[QT_STACK];   // frame kind (0 ADD, 1 ERASE, 2 CONT), agent, node
    int sp = 0;
    fk[sp] = 0; fa[sp] = a0; fn[sp] = start; sp++;
    while (sp > 0) {
        --sp;
        const int kind = fk[sp], a = fa[sp], n = fn[sp];
        const double* b = t.box + 4 * n;
        const double cx = b[0] + b[2] / 2, cy = b[1] + b[3] / 2;
        if (kind == 1) { qt_erase(t, a, n); continue; }
        if (kind == 2) {
            int m = a;
            while (m < t.na && !qt_member(t, m, n)) m++;
            if (m >= t.na || sp + 6 >= QT_STACK) continue;
            const int c0 = t.child0[n];
            const double mx = t.pos[2 * m], my = t.pos[2 * m + 1];
            fk[sp] = 2; fa[sp] = m + 1; fn[sp] = n; sp++;
            fk[sp] = 1; fa[sp] = m; fn[sp] = n; sp++;
            if ((mx <= cx) && (my >= cy)) { fk[sp] = 0; fa[sp] = m; fn[sp] = c0 + 3; sp++; }
            if ((mx >= cx) && (my <= cy)) { fk[sp] = 0; fa[sp] = m; fn[sp] = c0 + 1; sp++; }
            if ((mx <= cx) && (my <= cy)) { fk[sp] = 0; fa[sp] = m; fn[sp] = c0 + 0; sp++; }
            if ((mx >= cx) && (my >= cy)) { fk[sp] = 0; fa[sp] = m; fn[sp] = c0 + 2; sp++; }
            continue;
        }
        const double px = t.pos[2 * a], py = t.pos[2 * a + 1];
        if (t.child0[n] < 0) {
            if (!qt_member(t, a, n)) {
                for (int k = 0; k < QT_LEAVES; k++) if (t.leaf[a * QT_LEAVES + k] < 0) { t.leaf[a * QT_LEAVES + k] = n; t.count[n]++; break; }
            }
            t.hash[a] = n;
            if (t.count[n] > 8 && *t.n_nodes + 4 <= QT_MAX_NODES && sp + 1 < QT_STACK) {
                // split (ped_tree.cpp:91-104): create the children, then move every held agent down
                const int c0 = *t.n_nodes; *t.n_nodes += 4;
                const double hw = b[2] / 2, hh = b[3] / 2;
                for (int k = 0; k < 4; k++) {
                    double* cb = t.box + 4 * (c0 + k);
                    cb[0] = (k == 1 || k == 2) ? b[0] + hw : b[0]; cb[1] = (k >= 2) ? b[1] + hh : b[1]; cb[2] = hw; cb[3] = hh;
                    t.child0[c0 + k] = -1; t.count[c0 + k] = 0;
                }
                t.child0[n] = c0;
                fk[sp] = 2; fa[sp] = 0; fn[sp] = n; sp++;
            }
        } else if (sp + 4 < QT_STACK) {
            const int c0 = t.child0[n];
            // pushed in reverse so that tree3 is processed first, then tree1, tree2, tree4
            if ((px <= cx) && (py >= cy)) { fk[sp] = 0; fa[sp] = a; fn[sp] = c0 + 3; sp++; }
            if ((px >= cx) && (py <= cy)) { fk[sp] = 0; fa[sp] = a; fn[sp] = c0 + 1; sp++; }
            if ((px <= cx) && (py <= cy)) { fk[sp] = 0; fa[sp] = a; fn[sp] = c0 + 0; sp++; }
            if ((px >= cx) && (py >= cy)) { fk[sp] = 0; fa[sp] = a; fn[sp] = c0 + 2; sp++; }
        }
    }
}
// Tscene::moveAgent -> Ttree::moveAgent (ped_tree.cpp:124-130)
__host__ __device__ inline void qt_move(const QTreeView& t, int a) {
    const int L = t.hash[a];
    if (L < 0) return;
    const double* b = t.box + 4 * L;
    const double px = t.pos[2 * a], py = t.pos[2 * a + 1];
    if ((px < b[0]) || (px > (b[0] + b[2])) || (py < b[1]) || (py > (b[1] + b[3]))) {
        qt_add(t, a, 0);        // scene->placeAgent(a)
        qt_erase(t, a, L);      // agents.erase(a) on the OLD leaf (no-op if it split meanwhile)
    }
}
HD void qt_init(const QTreeView& t) {
    *t.n_nodes = 1;
    t.box[0] = 0; t.box[1] = 10; t.box[2] = 10; t.box[3] = 10;     // Tscene(0,10,10,10), pedscene.h:19
    t.child0[0] = -1; t.count[0] = 0;
    for (int a = 0; a < t.na; a++) { t.hash[a] = -1; for (int k = 0; k < QT_LEAVES; k++) t.leaf[a * QT_LEAVES + k] = -1; }
}
