// ORCA / emotional ORCA pedestrian update, one thread per agent (BASELINE.json north_star).
// Reference: RVO2 v2.0 as vendored in src/3rdparty/ervo_ros:
//   Agent::computeNeighbors Agent.cpp:50-61, insertAgentNeighbor :795-818, insertObstacleNeighbor :820-838,
//   KdTree::queryObstacleTreeRecursive KdTree.cpp:322-353, Agent::computeNewVelocity Agent.cpp:437-793,
//   computeNewVelocityForERVO :72-434 + addEvacVelocity :63-69, linearProgram1/2/3 :845-1001,
//   Agent::update :840-843; adapters rvoscene.h:36-66, ervoscene.h:13-22.
// All arithmetic is float32 without FMA contraction like the x86 build of the reference.
// Agent neighbours: the reference walks a k-d tree; here every agent scans the scene's agents in
// index order with the same bounded sorted insertion, which yields the same <=10 nearest list
// (only exact distance ties could order differently).
#pragma once
#include "state.cuh"

#define RVO_EPS 0.00001f
#define ORCA_MAX_NEIGH 10
#define ORCA_MAX_OBST 48
#define ORCA_MAX_LINES (ORCA_MAX_NEIGH + ORCA_MAX_OBST)

struct V2 { float x, y; };
struct OLine { V2 point, direction; };
__device__ __forceinline__ V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator-(V2 a) { return v2(-a.x, -a.y); }
__device__ __forceinline__ float operator*(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
__device__ __forceinline__ V2 operator/(V2 a, float s) { const float inv = 1.0f / s; return v2(a.x * inv, a.y * inv); }
__device__ __forceinline__ float absSq(V2 a) { return a * a; }
__device__ __forceinline__ float vabs(V2 a) { return sqrtf(a * a); }
__device__ __forceinline__ float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ V2 normalize(V2 a) { return a / vabs(a); }
__device__ __forceinline__ float sqr(float a) { return a * a; }
__device__ __forceinline__ float leftOf(V2 a, V2 b, V2 c) { return det(a - c, b - a); }
__device__ __forceinline__ float distSqPointLineSegment(V2 a, V2 b, V2 c) {
    const float r = ((c - a) * (b - a)) / absSq(b - a);
    if (r < 0.0f) return absSq(c - a);
    else if (r > 1.0f) return absSq(c - b);
    else return absSq(c - (a + r * (b - a)));
}

struct ObstView {   // one scene's obstacle vertex ring + BSP
    const float* verts;   // [n][8] px,py,dx,dy,convex,next,prev,0
    const int* nodes;     // [n][3] obstacle,left,right
    int root;
    __device__ __forceinline__ V2 point(int i) const { return v2(verts[8 * i], verts[8 * i + 1]); }
    __device__ __forceinline__ V2 dir(int i) const { return v2(verts[8 * i + 2], verts[8 * i + 3]); }
    __device__ __forceinline__ bool convex(int i) const { return verts[8 * i + 4] != 0.f; }
    __device__ __forceinline__ int next(int i) const { return (int)verts[8 * i + 5]; }
    __device__ __forceinline__ int prev(int i) const { return (int)verts[8 * i + 6]; }
};

__device__ inline bool lp1(const OLine* lines, int lineNo, float radius, V2 optVelocity, bool directionOpt, V2& result) {
    const float dotProduct = lines[lineNo].point * lines[lineNo].direction;
    const float discriminant = sqr(dotProduct) + sqr(radius) - absSq(lines[lineNo].point);
    if (discriminant < 0.0f) return false;
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
    for (int i = 0; i < lineNo; ++i) {
        const float denominator = det(lines[lineNo].direction, lines[i].direction);
        const float numerator = det(lines[i].direction, lines[lineNo].point - lines[i].point);
        if (fabsf(denominator) <= RVO_EPS) {
            if (numerator < 0.0f) return false;
            else continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) return false;
    }
    if (directionOpt) {
        if (optVelocity * lines[lineNo].direction > 0.0f) result = lines[lineNo].point + tRight * lines[lineNo].direction;
        else result = lines[lineNo].point + tLeft * lines[lineNo].direction;
    } else {
        const float t = lines[lineNo].direction * (optVelocity - lines[lineNo].point);
        if (t < tLeft) result = lines[lineNo].point + tLeft * lines[lineNo].direction;
        else if (t > tRight) result = lines[lineNo].point + tRight * lines[lineNo].direction;
        else result = lines[lineNo].point + t * lines[lineNo].direction;
    }
    return true;
}

__device__ inline int lp2(const OLine* lines, int n, float radius, V2 optVelocity, bool directionOpt, V2& result) {
    if (directionOpt) result = optVelocity * radius;
    else if (absSq(optVelocity) > sqr(radius)) result = normalize(optVelocity) * radius;
    else result = optVelocity;
    for (int i = 0; i < n; ++i) {
        if (det(lines[i].direction, lines[i].point - result) > 0.0f) {
            const V2 tempResult = result;
            if (!lp1(lines, i, radius, optVelocity, directionOpt, result)) { result = tempResult; return i; }
        }
    }
    return n;
}

__device__ inline void lp3(const OLine* lines, int n, int numObstLines, int beginLine, float radius, V2& result, OLine* proj) {
    float distance = 0.0f;
    for (int i = beginLine; i < n; ++i) {
        if (det(lines[i].direction, lines[i].point - result) > distance) {
            int np = 0;
            for (int j = 0; j < numObstLines; ++j) proj[np++] = lines[j];
            for (int j = numObstLines; j < i; ++j) {
                OLine line;
                float determinant = det(lines[i].direction, lines[j].direction);
                if (fabsf(determinant) <= RVO_EPS) {
                    if (lines[i].direction * lines[j].direction > 0.0f) continue;
                    else line.point = 0.5f * (lines[i].point + lines[j].point);
                } else {
                    line.point = lines[i].point + (det(lines[j].direction, lines[i].point - lines[j].point) / determinant) * lines[i].direction;
                }
                line.direction = normalize(lines[j].direction - lines[i].direction);
                proj[np++] = line;
            }
            const V2 tempResult = result;
            if (lp2(proj, np, radius, v2(-lines[i].direction.y, lines[i].direction.x), true, result) < np) result = tempResult;
            distance = det(lines[i].direction, lines[i].point - result);
        }
    }
}

// One agent of RVOSimulator::doStep / ERVOSimulator::doStep. pos/vel: the scene's agents (shared memory).
// Returns the new velocity (the caller applies Agent::update after all agents are done).
__device__ inline V2 orca_new_velocity(int self, int n_agents, const V2* pos, const V2* vel, V2 prefVelocity,
                                       float maxSpeed, float timeStep, const ObstView& ob, bool ervo, int n_beeps,
                                       const V2* beep_p, const float* beep_r) {
    const float radius_ = 0.5f, neighborDist_ = 0.5f, timeHorizon_ = 5.f, timeHorizonObst_ = 5.f;   // rvoscene.h:53-66
    const V2 position_ = pos[self], velocity_ = vel[self];

    // ---- computeNeighbors: obstacles (BSP walk, KdTree.cpp:322-353) ----
    float od[ORCA_MAX_OBST]; int oi[ORCA_MAX_OBST]; int no = 0;
    {
        float rangeSq = sqr(timeHorizonObst_ * maxSpeed + radius_);
        // explicit stack of (node, stage): stage 0 = descend near side, 1 = after near side
        int stk_n[64]; unsigned char stk_s[64]; int sp = 0;
        if (ob.root >= 0) { stk_n[0] = ob.root; stk_s[0] = 0; sp = 1; }
        while (sp > 0) {
            int node = stk_n[sp - 1]; int stage = stk_s[sp - 1];
            const int o1 = ob.nodes[3 * node], o2 = ob.next(o1);
            const float agentLeftOfLine = leftOf(ob.point(o1), ob.point(o2), position_);
            if (stage == 0) {
                stk_s[sp - 1] = 1;
                int child = agentLeftOfLine >= 0.0f ? ob.nodes[3 * node + 1] : ob.nodes[3 * node + 2];
                if (child >= 0 && sp < 64) { stk_n[sp] = child; stk_s[sp] = 0; sp++; }
                continue;
            }
            sp--;
            const float distSqLine = sqr(agentLeftOfLine) / absSq(ob.point(o2) - ob.point(o1));
            if (distSqLine < rangeSq) {
                if (agentLeftOfLine < 0.0f) {
                    // insertObstacleNeighbor (Agent.cpp:820-838)
                    const float distSq = distSqPointLineSegment(ob.point(o1), ob.point(o2), position_);
                    if (distSq < rangeSq) {
                        if (no < ORCA_MAX_OBST) no++;
                        int i = no - 1;
                        while (i != 0 && distSq < od[i - 1]) { od[i] = od[i - 1]; oi[i] = oi[i - 1]; --i; }
                        od[i] = distSq; oi[i] = o1;
                    }
                }
                int child = agentLeftOfLine >= 0.0f ? ob.nodes[3 * node + 2] : ob.nodes[3 * node + 1];
                if (child >= 0 && sp < 64) { stk_n[sp] = child; stk_s[sp] = 0; sp++; }
            }
        }
    }
    // ---- computeNeighbors: agents (bounded sorted insertion, Agent.cpp:795-818) ----
    float nd[ORCA_MAX_NEIGH]; int ni[ORCA_MAX_NEIGH]; int nn = 0;
    {
        float rangeSq = sqr(neighborDist_);
        for (int a = 0; a < n_agents; a++) {
            if (a == self) continue;
            const float distSq = absSq(position_ - pos[a]);
            if (distSq < rangeSq) {
                if (nn < ORCA_MAX_NEIGH) nn++;
                int i = nn - 1;
                while (i != 0 && distSq < nd[i - 1]) { nd[i] = nd[i - 1]; ni[i] = ni[i - 1]; --i; }
                nd[i] = distSq; ni[i] = a;
                if (nn == ORCA_MAX_NEIGH) rangeSq = nd[nn - 1];
            }
        }
    }

    OLine lines[ORCA_MAX_LINES]; int nl = 0;
    const float invTimeHorizonObst = 1.0f / timeHorizonObst_;
    // ---- obstacle ORCA lines (Agent.cpp:442-683) ----
    for (int i = 0; i < no; ++i) {
        int obstacle1 = oi[i];
        int obstacle2 = ob.next(obstacle1);
        const V2 relativePosition1 = ob.point(obstacle1) - position_;
        const V2 relativePosition2 = ob.point(obstacle2) - position_;
        bool alreadyCovered = false;
        for (int j = 0; j < nl; ++j) {
            if (det(invTimeHorizonObst * relativePosition1 - lines[j].point, lines[j].direction) - invTimeHorizonObst * radius_ >= -RVO_EPS &&
                det(invTimeHorizonObst * relativePosition2 - lines[j].point, lines[j].direction) - invTimeHorizonObst * radius_ >= -RVO_EPS) {
                alreadyCovered = true;
                break;
            }
        }
        if (alreadyCovered) continue;
        const float distSq1 = absSq(relativePosition1);
        const float distSq2 = absSq(relativePosition2);
        const float radiusSq = sqr(radius_);
        const V2 obstacleVector = ob.point(obstacle2) - ob.point(obstacle1);
        const float s = (-relativePosition1 * obstacleVector) / absSq(obstacleVector);
        const float distSqLine = absSq(-relativePosition1 - s * obstacleVector);
        OLine line;
        if (s < 0.0f && distSq1 <= radiusSq) {
            if (ob.convex(obstacle1)) {
                line.point = v2(0.0f, 0.0f);
                line.direction = normalize(v2(-relativePosition1.y, relativePosition1.x));
                lines[nl++] = line;
            }
            continue;
        } else if (s > 1.0f && distSq2 <= radiusSq) {
            if (ob.convex(obstacle2) && det(relativePosition2, ob.dir(obstacle2)) >= 0.0f) {
                line.point = v2(0.0f, 0.0f);
                line.direction = normalize(v2(-relativePosition2.y, relativePosition2.x));
                lines[nl++] = line;
            }
            continue;
        } else if (s >= 0.0f && s < 1.0f && distSqLine <= radiusSq) {
            line.point = v2(0.0f, 0.0f);
            line.direction = -ob.dir(obstacle1);
            lines[nl++] = line;
            continue;
        }
        V2 leftLegDirection, rightLegDirection;
        if (s < 0.0f && distSqLine <= radiusSq) {
            if (!ob.convex(obstacle1)) continue;
            obstacle2 = obstacle1;
            const float leg1 = sqrtf(distSq1 - radiusSq);
            leftLegDirection = v2(relativePosition1.x * leg1 - relativePosition1.y * radius_, relativePosition1.x * radius_ + relativePosition1.y * leg1) / distSq1;
            rightLegDirection = v2(relativePosition1.x * leg1 + relativePosition1.y * radius_, -relativePosition1.x * radius_ + relativePosition1.y * leg1) / distSq1;
        } else if (s > 1.0f && distSqLine <= radiusSq) {
            if (!ob.convex(obstacle2)) continue;
            obstacle1 = obstacle2;
            const float leg2 = sqrtf(distSq2 - radiusSq);
            leftLegDirection = v2(relativePosition2.x * leg2 - relativePosition2.y * radius_, relativePosition2.x * radius_ + relativePosition2.y * leg2) / distSq2;
            rightLegDirection = v2(relativePosition2.x * leg2 + relativePosition2.y * radius_, -relativePosition2.x * radius_ + relativePosition2.y * leg2) / distSq2;
        } else {
            if (ob.convex(obstacle1)) {
                const float leg1 = sqrtf(distSq1 - radiusSq);
                leftLegDirection = v2(relativePosition1.x * leg1 - relativePosition1.y * radius_, relativePosition1.x * radius_ + relativePosition1.y * leg1) / distSq1;
            } else {
                leftLegDirection = -ob.dir(obstacle1);
            }
            if (ob.convex(obstacle2)) {
                const float leg2 = sqrtf(distSq2 - radiusSq);
                rightLegDirection = v2(relativePosition2.x * leg2 + relativePosition2.y * radius_, -relativePosition2.x * radius_ + relativePosition2.y * leg2) / distSq2;
            } else {
                rightLegDirection = ob.dir(obstacle1);
            }
        }
        const int leftNeighbor = ob.prev(obstacle1);
        bool isLeftLegForeign = false, isRightLegForeign = false;
        if (ob.convex(obstacle1) && det(leftLegDirection, -ob.dir(leftNeighbor)) >= 0.0f) {
            leftLegDirection = -ob.dir(leftNeighbor);
            isLeftLegForeign = true;
        }
        if (ob.convex(obstacle2) && det(rightLegDirection, ob.dir(obstacle2)) <= 0.0f) {
            rightLegDirection = ob.dir(obstacle2);
            isRightLegForeign = true;
        }
        const V2 leftCutoff = invTimeHorizonObst * (ob.point(obstacle1) - position_);
        const V2 rightCutoff = invTimeHorizonObst * (ob.point(obstacle2) - position_);
        const V2 cutoffVec = rightCutoff - leftCutoff;
        const float t = (obstacle1 == obstacle2 ? 0.5f : ((velocity_ - leftCutoff) * cutoffVec) / absSq(cutoffVec));
        const float tLeft = ((velocity_ - leftCutoff) * leftLegDirection);
        const float tRight = ((velocity_ - rightCutoff) * rightLegDirection);
        if ((t < 0.0f && tLeft < 0.0f) || (obstacle1 == obstacle2 && tLeft < 0.0f && tRight < 0.0f)) {
            const V2 unitW = normalize(velocity_ - leftCutoff);
            line.direction = v2(unitW.y, -unitW.x);
            line.point = leftCutoff + radius_ * invTimeHorizonObst * unitW;
            lines[nl++] = line;
            continue;
        } else if (t > 1.0f && tRight < 0.0f) {
            const V2 unitW = normalize(velocity_ - rightCutoff);
            line.direction = v2(unitW.y, -unitW.x);
            line.point = rightCutoff + radius_ * invTimeHorizonObst * unitW;
            lines[nl++] = line;
            continue;
        }
        const float INF = __int_as_float(0x7f800000);
        const float distSqCutoff = ((t < 0.0f || t > 1.0f || obstacle1 == obstacle2) ? INF : absSq(velocity_ - (leftCutoff + t * cutoffVec)));
        const float distSqLeft = ((tLeft < 0.0f) ? INF : absSq(velocity_ - (leftCutoff + tLeft * leftLegDirection)));
        const float distSqRight = ((tRight < 0.0f) ? INF : absSq(velocity_ - (rightCutoff + tRight * rightLegDirection)));
        if (distSqCutoff <= distSqLeft && distSqCutoff <= distSqRight) {
            line.direction = -ob.dir(obstacle1);
            line.point = leftCutoff + radius_ * invTimeHorizonObst * v2(-line.direction.y, line.direction.x);
            lines[nl++] = line;
            continue;
        } else if (distSqLeft <= distSqRight) {
            if (isLeftLegForeign) continue;
            line.direction = leftLegDirection;
            line.point = leftCutoff + radius_ * invTimeHorizonObst * v2(-line.direction.y, line.direction.x);
            lines[nl++] = line;
            continue;
        } else {
            if (isRightLegForeign) continue;
            line.direction = -rightLegDirection;
            line.point = rightCutoff + radius_ * invTimeHorizonObst * v2(-line.direction.y, line.direction.x);
            lines[nl++] = line;
            continue;
        }
    }
    const int numObstLines = nl;
    const float invTimeHorizon = 1.0f / timeHorizon_;
    // ---- agent ORCA lines (Agent.cpp:689-761) ----
    for (int i = 0; i < nn; ++i) {
        const int other = ni[i];
        const V2 relativePosition = pos[other] - position_;
        const V2 relativeVelocity = velocity_ - vel[other];
        const float distSq = absSq(relativePosition);
        const float combinedRadius = radius_ + radius_;
        const float combinedRadiusSq = sqr(combinedRadius);
        OLine line; V2 u;
        if (distSq > combinedRadiusSq) {
            const V2 w = relativeVelocity - invTimeHorizon * relativePosition;
            const float wLengthSq = absSq(w);
            const float dotProduct1 = w * relativePosition;
            if (dotProduct1 < 0.0f && sqr(dotProduct1) > combinedRadiusSq * wLengthSq) {
                const float wLength = sqrtf(wLengthSq);
                const V2 unitW = w / wLength;
                line.direction = v2(unitW.y, -unitW.x);
                u = (combinedRadius * invTimeHorizon - wLength) * unitW;
            } else {
                const float leg = sqrtf(distSq - combinedRadiusSq);
                if (det(relativePosition, w) > 0.0f) {
                    line.direction = v2(relativePosition.x * leg - relativePosition.y * combinedRadius, relativePosition.x * combinedRadius + relativePosition.y * leg) / distSq;
                } else {
                    line.direction = -v2(relativePosition.x * leg + relativePosition.y * combinedRadius, -relativePosition.x * combinedRadius + relativePosition.y * leg) / distSq;
                }
                const float dotProduct2 = relativeVelocity * line.direction;
                u = dotProduct2 * line.direction - relativeVelocity;
            }
        } else {
            const float invTimeStep = 1.0f / timeStep;
            const V2 w = relativeVelocity - invTimeStep * relativePosition;
            const float wLength = vabs(w);
            const V2 unitW = w / wLength;
            line.direction = v2(unitW.y, -unitW.x);
            u = (combinedRadius * invTimeStep - wLength) * unitW;
        }
        line.point = velocity_ + 0.5f * u;
        lines[nl++] = line;
    }
    V2 newVelocity;
    int lineFail = lp2(lines, nl, maxSpeed, prefVelocity, false, newVelocity);
    if (lineFail < nl) {
        OLine proj[ORCA_MAX_LINES];
        lp3(lines, nl, numObstLines, lineFail, maxSpeed, newVelocity, proj);
    }
    if (ervo) {   // addEvacVelocity, Agent.cpp:63-69 (added after the LP, not re-clipped)
        for (int b = 0; b < n_beeps; b++) {
            V2 evacVec = position_ - beep_p[b];
            if (vabs(evacVec) > beep_r[b] || vabs(evacVec) < 1e-4) continue;
            newVelocity = newVelocity + normalize(evacVec);
        }
    }
    return newVelocity;
}
