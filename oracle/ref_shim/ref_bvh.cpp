// ref_bvh.cpp -- C entry point around the reference's unmodified BVH builder
// (compiled from /root/reference/bvh.cpp by oracle/build_ref.sh).  Test infrastructure only.
#include "precomp.h"

extern "C" {

// tris: n x 64 B (centroids are overwritten by the builder, bvh.cpp:23).
// nodesOut: capacity nodeCap >= 2n+64 nodes of 32 B; triIdxOut: n u32.
// Returns the number of node slots copied (2n+64, covering the overrun of App. B-3).
int ref_bvh_build(void* tris, int n, void* nodesOut, int nodeCap, unsigned* triIdxOut)
{
    Mesh mesh;
    mesh.triangles = (Tri*)tris;
    mesh.triangleCount = n;
    BVH* bvh = new BVH(&mesh);
    int slots = 2 * n + 64;
    if (slots > nodeCap) slots = nodeCap;
    memcpy(nodesOut, bvh->bvhNode, (size_t)slots * sizeof(BVHNode));
    memcpy(triIdxOut, bvh->triIdx, (size_t)n * sizeof(uint));
    free(bvh->bvhNode);
    delete[] bvh->triIdx;
    delete bvh;
    return slots;
}

int ref_bvh_nodes_used_reported(int n) { return 2 * n; } // bvh.cpp:43

} // extern "C"
