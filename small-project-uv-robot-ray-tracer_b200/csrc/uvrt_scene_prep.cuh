// uvrt_scene_prep.cuh -- the scene repack of uvrt_upload_scene on the device.
//
// Input: the reference's arrays exactly as RayTracer::Init hands them over (raytracer.cpp:24-30):
// Tri[nTris] (64 B, mesh.h:6-13), BVHNode[nNodes] (32 B, bvh.h:11-21; unused slots may hold anything)
// and triIdx[nTris].  Output: the traversal layout of DESIGN.md section 3 -- one 64-byte `pairs` record
// per reachable inner node in pre-order (left child first) and one 64-byte `wtris` record per leaf
// triangle in leaf order -- byte-identical to what the host repack produces.
//
// Two kernels instead of a level-by-level sweep (the host would have to learn the depth first):
//   k_prep_walk  one persistent grid walks the tree from the root through a global queue (thread t
//                owns queue entries t, t+T, ...; an entry is filled by the thread that processed the
//                parent, which is resident or already finished, so waiting cannot deadlock).  Per node:
//                validation (index range, reached twice, leaf span, triIdx values, depth), parent link,
//                and from every leaf a bottom-up walk that sums the subtree's inner nodes and triangle
//                slots; the second child to arrive at a parent carries on (atomic arrival counter).
//                The walk that reaches the root ends the kernel.
//   k_prep_emit  every reachable node finds its pre-order number by walking up to the root:
//                pair index  = #ancestors + sum of inner nodes in the left siblings along the path,
//                leaf slot   = sum of triangle slots in the left siblings along the path,
//                then writes its record(s): the children's boxes and references, or its triangles
//                with edge1 = v1 - v0, edge2 = v2 - v0 (the same fp32 subtractions as extend.cl:13).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uvrt_prep {

constexpr uint32_t kNone = 0xffffffffu;       // parent[]: not reached yet
constexpr uint32_t kRootParent = 0xfffffffeu;
constexpr unsigned long long kEmpty = ~0ull;  // queue entry not written yet
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kLastBit = 0x80000000u;

enum PrepError { PREP_OK = 0, PREP_NODE_RANGE = 1, PREP_TWICE = 2, PREP_LEAF_SPAN = 3, PREP_TRI_RANGE = 4, PREP_DEPTH = 5, PREP_TIMEOUT = 6 };

struct Status {                   // one per upload, read back by the host
    uint32_t err, errA, errB, errC;
    uint32_t done, reachable, maxDepth, tame;
    uint32_t nPairs, nLeaves, rootIsLeaf, nested;   // nested: every child box lies inside its parent's box
    unsigned long long nSlots;
    uint32_t tail, pad2;
};

struct RawNode { float mn[3]; uint32_t leftFirst; float mx[3]; uint32_t triCount; };

__device__ __forceinline__ void set_error(Status* st, uint32_t code, uint32_t a, uint32_t b, uint32_t c)
{
    if (atomicCAS(&st->err, 0u, code) == 0u) { st->errA = a; st->errB = b; st->errC = c; }
    __threadfence();
    atomicExch(&st->done, 1u);
}

__device__ __forceinline__ uint32_t ld_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__global__ void k_prep_init(unsigned long long* __restrict__ queue, Status* __restrict__ st, uint32_t* __restrict__ parent)
{
    Status z{};
    z.tame = 1;
    z.nested = 1;
    z.tail = 1;
    *st = z;
    parent[0] = kRootParent;
    queue[0] = 0ull;              // depth 0, node 0
}

__global__ void __launch_bounds__(128) k_prep_walk(const RawNode* __restrict__ nodes, uint32_t nNodes, const uint32_t* __restrict__ triIdx,
                                                   uint32_t nTris, int maxDepth, unsigned long long* __restrict__ queue,
                                                   uint32_t* __restrict__ parent, uint32_t* __restrict__ arrive,
                                                   uint32_t* __restrict__ subInner, uint32_t* __restrict__ subSlots,
                                                   Status* __restrict__ st)
{
    const uint32_t T = gridDim.x * blockDim.x;
    volatile unsigned long long* vq = queue;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += T) {
        unsigned long long e = vq[i];
        unsigned spins = 0;
        while (e == kEmpty) {
            if (ld_u32(&st->done)) return;
            if (++spins > (1u << 24)) { set_error(st, PREP_TIMEOUT, i, 0, 0); return; }
            __nanosleep(20);
            e = vq[i];
        }
        const uint32_t n = (uint32_t)e, depth = (uint32_t)(e >> 32);
        const RawNode nd = nodes[n];
        if (depth > ld_u32(&st->maxDepth)) atomicMax(&st->maxDepth, depth);   // rarely true: no hot-spot atomic
        if (nd.triCount == 0) {
            // inner node: claim both children and queue them
            const uint32_t c0 = nd.leftFirst, c1 = nd.leftFirst + 1;
            if (c1 >= nNodes || c1 < c0) { set_error(st, PREP_NODE_RANGE, c1 < c0 ? c0 : (c0 >= nNodes ? c0 : c1), 0, 0); return; }
            if ((int)depth + 1 >= maxDepth) { set_error(st, PREP_DEPTH, depth + 1, 0, 0); return; }
            // The two claims travel together (the level-to-level latency of this walk is what the upload waits
            // for); queue slots are reserved only by a node that owns both children, so `tail` never exceeds
            // the number of distinct claimed nodes + 1 <= nNodes even on a malformed (DAG / cyclic) array --
            // the validator must not write out of bounds on exactly the input it exists to reject.
            const uint32_t old0 = atomicCAS(&parent[c0], kNone, n);
            const uint32_t old1 = atomicCAS(&parent[c1], kNone, n);
            if (old0 != kNone) { set_error(st, PREP_TWICE, c0, 0, 0); return; }
            if (old1 != kNone) { set_error(st, PREP_TWICE, c1, 0, 0); return; }
            const uint32_t pos = atomicAdd(&st->tail, 2u);
            if (pos + 1u >= nNodes) { set_error(st, PREP_NODE_RANGE, c1, 0, 0); return; }   // cannot happen for claimed pairs; belt and braces
            __threadfence();
            const unsigned long long d1 = (unsigned long long)(depth + 1) << 32;
            vq[pos] = d1 | c0;
            vq[pos + 1] = d1 | c1;
            continue;
        }
        // leaf: validate, then carry the subtree totals upwards
        if ((unsigned long long)nd.leftFirst + nd.triCount > nTris) { set_error(st, PREP_LEAF_SPAN, n, nd.leftFirst, nd.leftFirst + nd.triCount); return; }
        for (uint32_t k = 0; k < nd.triCount; k++) {
            const uint32_t t = triIdx[nd.leftFirst + k];
            if (t >= nTris) { set_error(st, PREP_TRI_RANGE, t, 0, 0); return; }
        }
        uint32_t inner = 0, slots = nd.triCount, x = n;
        for (;;) {
            subInner[x] = inner;
            subSlots[x] = slots;
            const uint32_t p = ld_u32(&parent[x]);
            if (p == kRootParent) {
                st->nPairs = inner;
                st->nSlots = slots;
                st->rootIsLeaf = nodes[0].triCount > 0 ? 1u : 0u;
                __threadfence();
                atomicExch(&st->done, 1u);
                return;
            }
            __threadfence();
            if (atomicAdd(&arrive[p], 1u) == 0u) break;        // the sibling's walk carries on
            const uint32_t left = nodes[p].leftFirst;
            const uint32_t sib = x == left ? left + 1 : left;
            inner += 1u + ld_u32(&subInner[sib]);
            slots += ld_u32(&subSlots[sib]);
            x = p;
        }
    }
}

__device__ __forceinline__ bool coord_tame(float v)
{
    const float a = fabsf(v);
    return a == 0.0f || (a >= 6.6174449e-24f && a <= 1048576.0f);   // 0 or [2^-77, 2^20], see ray_is_tame()
}

__global__ void __launch_bounds__(256) k_prep_emit(const RawNode* __restrict__ nodes, const uint32_t* __restrict__ triIdx,
                                                   const float4* __restrict__ tris, const unsigned long long* __restrict__ queue,
                                                   uint32_t nReachable, const uint32_t* __restrict__ parent,
                                                   const uint32_t* __restrict__ subInner, const uint32_t* __restrict__ subSlots,
                                                   float4* __restrict__ pairs, float4* __restrict__ wtris, Status* __restrict__ st)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nReachable) return;
    const uint32_t n = (uint32_t)queue[i];
    const RawNode nd = nodes[n];
    // pre-order numbers from the path to the root
    uint32_t pid = 0, sb = 0, x = n;
    for (uint32_t p = parent[x]; p != kRootParent; p = parent[x]) {
        const uint32_t left = nodes[p].leftFirst;
        pid += 1u;
        if (x != left) { pid += subInner[left]; sb += subSlots[left]; }
        x = p;
    }
    if (nd.triCount > 0) {
        for (uint32_t k = 0; k < nd.triCount; k++) {
            const uint32_t t = triIdx[nd.leftFirst + k];
            const float4 v0 = tris[4ull * t], v1 = tris[4ull * t + 1], v2 = tris[4ull * t + 2];
            float4* w = wtris + 4ull * (sb + k);
            const uint32_t tag = t | (k + 1 == nd.triCount ? kLastBit : 0u);
            // the spare lanes carry the leaf's own box (every slot of a leaf repeats it) for the exact
            // verification step of the certified fast extend (uvrt_fast.cuh): min.z | max.z | (min.xy, max.xy)
            w[0] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(tag));
            w[1] = make_float4(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z), nd.mn[2]);
            w[2] = make_float4(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z), nd.mx[2]);
            w[3] = make_float4(nd.mn[0], nd.mn[1], nd.mx[0], nd.mx[1]);
        }
        return;
    }
    const uint32_t c0 = nd.leftFirst;
    const RawNode a = nodes[c0], b = nodes[c0 + 1];
    const uint32_t ref0 = a.triCount > 0 ? (kLeafBit | sb) : pid + 1u;
    const uint32_t ref1 = b.triCount > 0 ? (kLeafBit | (sb + subSlots[c0])) : pid + 1u + subInner[c0];
    // (min.x, min.y) (max.x, max.y) | (min.z, max.z) (ref, 0): 64-bit pairs for the packed fp32x2 pipe
    float4* p = pairs + 4ull * pid;
    p[0] = make_float4(a.mn[0], a.mn[1], a.mx[0], a.mx[1]);
    p[1] = make_float4(a.mn[2], a.mx[2], __uint_as_float(ref0), 0.0f);
    p[2] = make_float4(b.mn[0], b.mn[1], b.mx[0], b.mx[1]);
    p[3] = make_float4(b.mn[2], b.mx[2], __uint_as_float(ref1), 0.0f);
    bool tame = true;
#pragma unroll
    for (int k = 0; k < 3; k++)
        tame = tame && coord_tame(a.mn[k]) && coord_tame(a.mx[k]) && a.mn[k] <= a.mx[k] && coord_tame(b.mn[k]) && coord_tame(b.mx[k]) && b.mn[k] <= b.mx[k];
    if (!tame) st->tame = 0;
    // nested boxes make the exact slab results monotone along a path (uvrt_fast.cuh); the root's own box is never tested
    if (n != 0u) {
        bool nested = true;
#pragma unroll
        for (int k = 0; k < 3; k++)
            nested = nested && a.mn[k] >= nd.mn[k] && a.mx[k] <= nd.mx[k] && b.mn[k] >= nd.mn[k] && b.mx[k] <= nd.mx[k];
        if (!nested) st->nested = 0;
    }
}

} // namespace uvrt_prep
