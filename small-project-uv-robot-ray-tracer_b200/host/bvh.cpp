// bvh.cpp -- binned-SAH BVH2 builder producing the same tree, node numbering and triIdx order as
// the reference's builder (/root/reference/bvh.cpp:13-220), so that the device traversal sees the
// same boxes in the same child order (needed for bit-exact closest hits, DESIGN.md "Parity").
//
// What is kept from the reference, because it decides the tree:
//   * centroid = (v0 + v1 + v2) * 0.3333f                                    (bvh.cpp:23)
//   * 8 bins per axis on the centroid bounds, bin = min(7, (int)((c - lo) * (8 / (hi - lo))))
//   * plane cost = nLeft * area(leftBox) + nRight * area(rightBox), area = ex*ey + ey*ez + ez*ex,
//     first strictly smaller cost wins, axes x,y,z then planes 1..7        (bvh.cpp:98-179)
//   * the right-hand box of plane p is grown from bins p..6 while the right-hand COUNT covers
//     bins p+1..7 (bvh.cpp:134-138: the index of the bin whose box is merged lags by one).
//     This looks like an off-by-one in the reference, but it selects the split planes, so it is
//     reproduced.
//   * a node is split iff bestCost < area(node) * triCount; no traversal-cost term  (bvh.cpp:52-54)
//   * children are allocated as an adjacent pair, left first; node 1 stays unused so pairs are
//     64-byte aligned; the first four levels are numbered depth-first from slot 2, and each of
//     the (up to 16) level-4 subtrees owns the slot range [base_i, base_i + 2*tris_i)  (bvh.cpp:31-42)
// What differs: plain scalar code instead of SSE lanes, explicit deferred-job list, the node
// array is allocated with the slack the numbering scheme really needs (2N + 64) and nodesUsed
// reports the true extent (the reference under-allocates: SURVEY App. B-3).
#include "precomp.h"
#include "../../include/uvrt.h"
#include <algorithm>

using namespace Tmpl8;

namespace {
inline float half_area(const float* lo, const float* hi)
{
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return ex * ey + ey * ez + ez * ex;
}
inline void grow(float* lo, float* hi, const float3_strict& p)
{
    lo[0] = p.x < lo[0] ? p.x : lo[0]; hi[0] = p.x > hi[0] ? p.x : hi[0];
    lo[1] = p.y < lo[1] ? p.y : lo[1]; hi[1] = p.y > hi[1] ? p.y : hi[1];
    lo[2] = p.z < lo[2] ? p.z : lo[2]; hi[2] = p.z > hi[2] ? p.z : hi[2];
}
// bin = min(BINS - 1, (int)((c - lo) * scale)) (bvh.cpp:117, 58).  For finite input the product lies in [0, 8];
// the lower clamp only matters for NaN / infinite coordinates, where the reference indexes out of bounds.
inline int BinOf(float c, float lo, float scale)
{
    const float x = (c - lo) * scale;
    if (!(x >= 0.0f)) return 0;
    return x >= (float)(BINS - 1) ? BINS - 1 : (int)x;
}
inline void reset(float* lo, float* hi)
{
    lo[0] = lo[1] = lo[2] = 1e30f;
    hi[0] = hi[1] = hi[2] = -1e30f;
}
} // namespace

BVH::BVH(Mesh* m) : mesh(m)
{
    nodeCapacity = (uint)m->triangleCount * 2u + 64u;
    void* p = nullptr;
    if (posix_memalign(&p, 64, sizeof(BVHNode) * (size_t)nodeCapacity) != 0) p = nullptr;
    bvhNode = (BVHNode*)p;
    triIdx = new uint[m->triangleCount > 0 ? m->triangleCount : 1];
    if (bvhNode) Build();
}

BVH::BVH(Mesh* m, uvrt_ctx* ctx) : mesh(m)
{
    nodeCapacity = (uint)m->triangleCount * 2u + 64u;
    void* p = nullptr;
    if (posix_memalign(&p, 64, sizeof(BVHNode) * (size_t)nodeCapacity) != 0) p = nullptr;
    bvhNode = (BVHNode*)p;
    triIdx = new uint[m->triangleCount > 0 ? m->triangleCount : 1];
    ok = false;
    if (!bvhNode || m->triangleCount <= 0) return;
    memset(bvhNode, 0, sizeof(BVHNode) * (size_t)nodeCapacity);
    uint32_t used = 0;
    // centroids are written back into the mesh, as Build() does (bvh.cpp:23 of the reference)
    ok = uvrt_build_bvh(ctx, m->triangles, m->triangleCount, bvhNode, (int)nodeCapacity, triIdx, &used, m->triangles) == UVRT_OK;
    nodesUsed = ok ? used : 0;
}

BVH::~BVH()
{
    free(bvhNode);
    delete[] triIdx;
}

void BVH::NodeBounds(uint nodeIdx, Bounds3& cb)
{
    BVHNode& node = bvhNode[nodeIdx];
    float lo[3], hi[3];
    reset(lo, hi);
    reset(cb.lo, cb.hi);
    for (uint i = 0; i < node.triCount; i++) {
        const Tri& t = mesh->triangles[triIdx[node.leftFirst + i]];
        grow(lo, hi, t.vertex0);
        grow(lo, hi, t.vertex1);
        grow(lo, hi, t.vertex2);
        grow(cb.lo, cb.hi, t.centroid);
    }
    node.aabbMin = make_float3_strict(lo[0], lo[1], lo[2]);
    node.aabbMax = make_float3_strict(hi[0], hi[1], hi[2]);
}

float BVH::BestSplit(const BVHNode& node, const Bounds3& cb, int& axis, int& plane) const
{
    float best = 1e30f;
    for (int a = 0; a < 3; a++) {
        const float lo = cb.lo[a], hi = cb.hi[a];
        if (lo == hi) continue;
        const float scale = BINS / (hi - lo);
        float binLo[BINS][3], binHi[BINS][3];
        uint count[BINS];
        for (int b = 0; b < BINS; b++) { reset(binLo[b], binHi[b]); count[b] = 0; }
        for (uint i = 0; i < node.triCount; i++) {
            const Tri& t = mesh->triangles[triIdx[node.leftFirst + i]];
            int b = BinOf(t.centroid[a], lo, scale);
            count[b]++;
            grow(binLo[b], binHi[b], t.vertex0);
            grow(binLo[b], binHi[b], t.vertex1);
            grow(binLo[b], binHi[b], t.vertex2);
        }
        // sweep from both ends; costL[p-1], costR[p-1] belong to the plane between bins p-1 and p
        float costL[BINS - 1], costR[BINS - 1];
        float lLo[3], lHi[3], rLo[3], rHi[3];
        reset(lLo, lHi);
        reset(rLo, rHi);
        int nL = 0, nR = 0;
        for (int i = 0; i < BINS - 1; i++) {
            nL += (int)count[i];
            for (int k = 0; k < 3; k++) {
                lLo[k] = binLo[i][k] < lLo[k] ? binLo[i][k] : lLo[k];
                lHi[k] = binHi[i][k] > lHi[k] ? binHi[i][k] : lHi[k];
            }
            costL[i] = nL * half_area(lLo, lHi);
            const int rb = BINS - 2 - i;   // box index lags the count index by one (see header)
            nR += (int)count[rb + 1];
            for (int k = 0; k < 3; k++) {
                rLo[k] = binLo[rb][k] < rLo[k] ? binLo[rb][k] : rLo[k];
                rHi[k] = binHi[rb][k] > rHi[k] ? binHi[rb][k] : rHi[k];
            }
            costR[rb] = nR * half_area(rLo, rHi);
        }
        for (int i = 0; i < BINS - 1; i++) {
            const float c = costL[i] + costR[i];
            if (c < best) { axis = a; plane = i + 1; best = c; }
        }
    }
    return best;
}

void BVH::Split(uint nodeIdx, int level, uint& nextFree, Bounds3 cb, std::vector<Job>* deferred)
{
    BVHNode& node = bvhNode[nodeIdx];
    int axis = 0, plane = 0;
    const float splitCost = BestSplit(node, cb, axis, plane);
    if (splitCost >= node.CalculateNodeCost()) return;
    // partition triIdx[leftFirst .. leftFirst+triCount) in place with the binning expression
    int i = (int)node.leftFirst;
    int j = i + (int)node.triCount - 1;
    const float lo = cb.lo[axis];
    const float scale = BINS / (cb.hi[axis] - lo);
    while (i <= j) {
        int b = BinOf(mesh->triangles[triIdx[i]].centroid[axis], lo, scale);
        if (b < plane) i++;
        else std::swap(triIdx[i], triIdx[j--]);
    }
    const uint nLeft = (uint)i - node.leftFirst;
    if (nLeft == 0 || nLeft == node.triCount) return;
    const uint left = nextFree++, right = nextFree++;
    bvhNode[left].leftFirst = node.leftFirst;
    bvhNode[left].triCount = nLeft;
    bvhNode[right].leftFirst = (uint)i;
    bvhNode[right].triCount = node.triCount - nLeft;
    node.leftFirst = left;
    node.triCount = 0;
    // children of level-3 nodes are the roots of the independent sub-builds
    const bool defer = deferred != nullptr && level == 3;
    const uint kids[2] = {left, right};
    for (uint k : kids) {
        Bounds3 ccb;
        NodeBounds(k, ccb);
        if (defer) deferred->push_back(Job{k, ccb});
        else Split(k, level + 1, nextFree, ccb, deferred);
    }
}

void BVH::Build()
{
    const int n = mesh->triangleCount;
    memset(bvhNode, 0, sizeof(BVHNode) * (size_t)nodeCapacity);
    nodesUsed = 0;
    if (n <= 0) return;
    Tri* tri = mesh->triangles;
    for (int i = 0; i < n; i++) {
        triIdx[i] = (uint)i;
        tri[i].centroid = (tri[i].vertex0 + tri[i].vertex1 + tri[i].vertex2) * 0.3333f;
    }
    bvhNode[0].leftFirst = 0;
    bvhNode[0].triCount = (uint)n;
    Bounds3 rootCb;
    NodeBounds(0, rootCb);
    uint nextFree = 2;   // slot 1 is skipped so that sibling pairs sit on 64-byte lines
    std::vector<Job> jobs;
    Split(0, 0, nextFree, rootCb, &jobs);
    // each deferred subtree numbers its nodes inside its own slot range
    std::vector<uint> base(jobs.size() + 1, nextFree);
    for (size_t k = 0; k < jobs.size(); k++) base[k + 1] = base[k] + bvhNode[jobs[k].node].triCount * 2;
    std::vector<uint> endSlot(jobs.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < (int)jobs.size(); k++) {
        uint next = base[k];
        Split(jobs[k].node, 99, next, jobs[k].cb, nullptr);
        endSlot[k] = next;
    }
    uint used = nextFree;
    for (size_t k = 0; k < jobs.size(); k++) used = std::max(used, endSlot[k]);
    nodesUsed = used;
}
