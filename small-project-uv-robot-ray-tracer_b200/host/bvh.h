// bvh.h -- host BVH2 with the reference's node format and numbering (bvh.h:11-50 of the
// reference), built by an own binned-SAH builder (bvh.cpp here).
#pragma once

namespace Tmpl8 { class Mesh; }
struct uvrt_ctx;

// 32-byte node, identical in memory to the reference's (bvh.h:11-21) and to cl/tools.cl:39-45
struct BVHNode {
    union { struct { Tmpl8::float3_strict aabbMin; uint leftFirst; }; float aabbMin4[4]; };
    union { struct { Tmpl8::float3_strict aabbMax; uint triCount; }; float aabbMax4[4]; };
    bool isLeaf() const { return triCount > 0; }
    float CalculateNodeCost() const
    {
        Tmpl8::float3_strict e = aabbMax - aabbMin;
        return (e.x * e.y + e.y * e.z + e.z * e.x) * triCount;
    }
};
static_assert(sizeof(BVHNode) == 32, "BVHNode must stay 32 bytes");

#define BINS 8 // SAH bins per axis (bvh.h:26 of the reference)

class BVH {
public:
    BVH() = default;
    explicit BVH(Tmpl8::Mesh* mesh);
    // Builds on the device through uvrt_build_bvh (same tree, numbering and triIdx order as Build()).
    // `ok` is false when the backend call failed; the arrays are then empty.
    BVH(Tmpl8::Mesh* mesh, uvrt_ctx* ctx);
    ~BVH();
    BVH(const BVH&) = delete;
    BVH& operator=(const BVH&) = delete;
    void Build();

    uint* triIdx = 0;
    // Number of node slots that hold the tree: highest used index + 1.  (The reference reports
    // 2N here, bvh.cpp:43, which is up to 30 short of its own tree -- SURVEY App. B-3.)
    uint nodesUsed = 0;
    BVHNode* bvhNode = 0;
    uint nodeCapacity = 0;   // allocated slots (2N + 64)
    bool ok = true;

private:
    struct Bounds3 { float lo[3], hi[3]; };
    struct Job { uint node; Bounds3 cb; };
    void NodeBounds(uint nodeIdx, Bounds3& centroidBounds);
    float BestSplit(const BVHNode& node, const Bounds3& cb, int& axis, int& plane) const;
    void Split(uint nodeIdx, int level, uint& nextFree, Bounds3 cb, std::vector<Job>* deferred);
    Tmpl8::Mesh* mesh = 0;
};
