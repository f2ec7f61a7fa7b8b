// uvrt_bvh_build.cuh -- binned-SAH BVH2 build on the device, producing the SAME tree, node numbering
// and triIdx order as the host builder (host/bvh.cpp) and therefore as the reference's
// /root/reference/bvh.cpp:13-220.  SURVEY.md section 8(f)-1: the CPU builder is the next wall after
// the wavefront path (0.1 s for the room, 6.4 s for the 10 M-triangle soup on 16 cores).
//
// What has to be reproduced exactly, and how it is done in parallel:
//   * split selection: per node and axis, 8 bins of triangle counts and vertex bounds (min/max are
//     exact and order-free, so shared-memory atomics give the same bins); the sweep over the 7 planes
//     is run by one thread with the host's fp32 expressions in the host's order (no FMA), including the
//     reference's lagging right-hand box (bvh.cpp:134-138).
//   * the in-place partition `while (i <= j) if (left(A[i])) i++; else swap(A[i], A[j--]);` leaves a
//     specific, unstable order.  Its result has a closed form (checked against the sequential loop on
//     2e5 random inputs, tests/test_host.py): with L = number of "left" elements,
//       - left elements of the prefix [0, L) stay;
//       - the k-th "hole" of the prefix (right element, from the left) receives the k-th left element
//         of the suffix counted from the right end;
//       - hole k's own element moves to `last` (k = 1) or to one slot before the (k-1)-th suffix-left;
//       - every right element of the suffix moves one slot towards the front, except the suffix's
//         first element, which (if right) moves to one slot before the last suffix-left (or to `last`).
//     Ranks come from block-wide prefix sums, so each element finds its destination independently.
//   * node numbering: children are allocated as an adjacent pair when a node splits, the left
//     subtree's descendants before the right's; the first four levels are numbered from slot 2 and
//     each level-4 subtree ("job") owns [base_i, base_i + 2*tris_i) (bvh.cpp:31-42).  The device build
//     numbers nodes arbitrarily (atomic counter), then computes subtree slot counts bottom-up and the
//     reference indices top-down, level by level.
// Work per level is dealt by node size: one thread for nodes of at most 16 triangles (runs the host
// loop as is), a 32-thread block up to 1,024, a 256-thread block up to 32,768, and above that the node
// is cut into 4,096-element chunks handled by a grid of blocks (k_huge_*: chunk bins merged with
// atomics, one decision thread per node, chunk counts -> scan -> ranks -> placement), so the top levels
// of a 10 M-triangle mesh run at memory speed instead of on one SM.  Leaves copy their segment to the
// final index array.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uvrt_bvh {

constexpr int kBins = 8;

struct BNode {            // 64 bytes, temporary numbering
    float bmin[3]; uint32_t first;
    float bmax[3]; uint32_t count;
    float cmin[3]; uint32_t left;     // temp index of the left child (right = left + 1); 0 = leaf
    float cmax[3]; uint32_t depth;
};

struct BAux {             // renumbering state per node
    uint32_t sFull;       // slots allocated inside Subdivide(node) when numbered recursively
    uint32_t alloc;       // value of the allocation pointer when Subdivide(node) starts
    uint32_t ref;         // index in the reference numbering
    uint32_t pad;
};

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }

// order-preserving float <-> uint (for atomicMin / atomicMax on floats)
__device__ __forceinline__ uint32_t enc(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float half_area(const float* lo, const float* hi)
{
    float ex = fs(hi[0], lo[0]), ey = fs(hi[1], lo[1]), ez = fs(hi[2], lo[2]);
    return fa(fa(fm(ex, ey), fm(ey, ez)), fm(ez, ex));
}

__device__ __forceinline__ int bin_of(float c, float lo, float scale)
{
    // (the clamps only matter for NaN / infinite coordinates; for finite input the product lies in [0, 8])
    const float x = fm(fs(c, lo), scale);
    if (!(x >= 0.0f)) return 0;
    return x >= (float)(kBins - 1) ? kBins - 1 : (int)x;
}

// centroid = (v0 + v1 + v2) * 0.3333f (bvh.cpp:23); also initialises the identity index array
__global__ void __launch_bounds__(256) k_centroids(float4* __restrict__ tris, uint32_t* __restrict__ idx, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tris[4ull * i], b = tris[4ull * i + 1], c = tris[4ull * i + 2];
    float4 ce = tris[4ull * i + 3];
    ce.x = fm(fa(fa(a.x, b.x), c.x), 0.3333f);
    ce.y = fm(fa(fa(a.y, b.y), c.y), 0.3333f);
    ce.z = fm(fa(fa(a.z, b.z), c.z), 0.3333f);
    tris[4ull * i + 3] = ce;
    idx[i] = (uint32_t)i;
}

// block-wide helpers -------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warpSums, uint32_t& total)
{
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += n;
    }
    __syncthreads();                       // warpSums may still be read from a previous call
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; k++) {
        uint32_t s = warpSums[k];
        before += (k < (int)w) ? s : 0u;
        tot += s;
    }
    total = tot;
    return before + inc - v;
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { float n = __shfl_xor_sync(0xffffffffu, v, o); v = n < v ? n : v; }
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { float n = __shfl_xor_sync(0xffffffffu, v, o); v = n > v ? n : v; }
    return v;
}

// Root: bounds of all triangles and of their centroids (UpdateNodeBounds, bvh.cpp:181-220).
__global__ void __launch_bounds__(256) k_root_bounds(const float4* __restrict__ tris, int n, uint32_t* __restrict__ acc /*12 encoded*/)
{
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v[3] = {tris[4ull * i], tris[4ull * i + 1], tris[4ull * i + 2]};
        float4 c = tris[4ull * i + 3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[0] = fminf(lo[0], v[k].x); hi[0] = fmaxf(hi[0], v[k].x);
            lo[1] = fminf(lo[1], v[k].y); hi[1] = fmaxf(hi[1], v[k].y);
            lo[2] = fminf(lo[2], v[k].z); hi[2] = fmaxf(hi[2], v[k].z);
        }
        clo[0] = fminf(clo[0], c.x); chi[0] = fmaxf(chi[0], c.x);
        clo[1] = fminf(clo[1], c.y); chi[1] = fmaxf(chi[1], c.y);
        clo[2] = fminf(clo[2], c.z); chi[2] = fmaxf(chi[2], c.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float a = warp_min(lo[k]), b = warp_max(hi[k]), c = warp_min(clo[k]), d = warp_max(chi[k]);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&acc[k], enc(a)); atomicMax(&acc[3 + k], enc(b));
            atomicMin(&acc[6 + k], enc(c)); atomicMax(&acc[9 + k], enc(d));
        }
    }
}

// ---- node lists of the next level, by size class -----------------------------------------------------
constexpr uint32_t kSmall = 16;        // <= kSmall: one thread (k_level_small)
constexpr uint32_t kMid = 1024;        // <= kMid: 32-thread block, <= kHuge: 256-thread block (k_level<>)
constexpr uint32_t kHuge = 32768;      // above: grid of chunk blocks (k_huge_*)
constexpr uint32_t kChunk = 4096;      // elements per chunk block of a huge node

struct HugeInfo { uint32_t id, first, count, chunk0; };

struct Lists {                          // next level's lists; counters: [0] next temp id, [1..4] list sizes
    uint32_t* small; uint32_t* mid; uint32_t* big; uint32_t* huge; HugeInfo* hugeInfo; uint32_t* counters;
};

__device__ __forceinline__ void enqueue_child(uint32_t id, uint32_t first, uint32_t count, const Lists& q)
{
    if (count <= kSmall) q.small[atomicAdd(&q.counters[1], 1u)] = id;
    else if (count <= kMid) q.mid[atomicAdd(&q.counters[2], 1u)] = id;
    else if (count <= kHuge) q.big[atomicAdd(&q.counters[3], 1u)] = id;
    else {
        const uint32_t s = atomicAdd(&q.counters[4], 1u);
        q.huge[s] = id;
        q.hugeInfo[s] = HugeInfo{id, first, count, 0u};
    }
}

__global__ void k_init_root(BNode* __restrict__ nodes, const uint32_t* __restrict__ acc, int n, Lists q)
{
    BNode r;
    for (int k = 0; k < 3; k++) { r.bmin[k] = dec(acc[k]); r.bmax[k] = dec(acc[3 + k]); r.cmin[k] = dec(acc[6 + k]); r.cmax[k] = dec(acc[9 + k]); }
    r.first = 0; r.count = (uint32_t)n; r.left = 0; r.depth = 0;
    nodes[0] = r;
    q.counters[0] = 1;      // temp ids: root = 0, children pairs follow
    q.counters[1] = q.counters[2] = q.counters[3] = q.counters[4] = 0;
    enqueue_child(0, 0, (uint32_t)n, q);
}

// Sweep over the 7 planes of every axis and the split decision (bvh.cpp:130-150, 52-54, 64-66): host
// arithmetic in host order.  cnt[3][8]; lo/hi[3][8][3] order-encoded (empty bins keep the sentinels).
// Returns 1 = split, 0 = leaf, 2 = leaf whose index segment the reference has already permuted: its
// in-place partition loop runs before the "one side is empty" check (bvh.cpp:56-66), and with every element
// on the right the loop leaves the segment rotated left by one.  (Only reachable when the costs degenerate
// to inf / NaN, e.g. coordinates beyond 1e19; with every element on the left nothing moves.)
__device__ __forceinline__ int decide_split(const uint32_t* cnt, const uint32_t* lo, const uint32_t* hi, const BNode& nd,
                                            int& axisOut, int& planeOut, uint32_t& Lout)
{
    float best = 1e30f;
    int axis = 0, plane = 0;
    for (int a = 0; a < 3; a++) {
        if (nd.cmin[a] == nd.cmax[a]) continue;
        float binLo[kBins][3], binHi[kBins][3];
        for (int b = 0; b < kBins; b++)
            for (int k = 0; k < 3; k++) {
                binLo[b][k] = cnt[a * kBins + b] ? dec(lo[(a * kBins + b) * 3 + k]) : 1e30f;
                binHi[b][k] = cnt[a * kBins + b] ? dec(hi[(a * kBins + b) * 3 + k]) : -1e30f;
            }
        float costL[kBins - 1], costR[kBins - 1];
        float lLo[3] = {1e30f, 1e30f, 1e30f}, lHi[3] = {-1e30f, -1e30f, -1e30f};
        float rLo[3] = {1e30f, 1e30f, 1e30f}, rHi[3] = {-1e30f, -1e30f, -1e30f};
        int nL = 0, nR = 0;
        for (int i = 0; i < kBins - 1; i++) {
            nL += (int)cnt[a * kBins + i];
            for (int k = 0; k < 3; k++) {
                lLo[k] = binLo[i][k] < lLo[k] ? binLo[i][k] : lLo[k];
                lHi[k] = binHi[i][k] > lHi[k] ? binHi[i][k] : lHi[k];
            }
            costL[i] = fm(__int2float_rn(nL), half_area(lLo, lHi));
            const int rb = kBins - 2 - i;         // box index lags the count index by one
            nR += (int)cnt[a * kBins + rb + 1];
            for (int k = 0; k < 3; k++) {
                rLo[k] = binLo[rb][k] < rLo[k] ? binLo[rb][k] : rLo[k];
                rHi[k] = binHi[rb][k] > rHi[k] ? binHi[rb][k] : rHi[k];
            }
            costR[rb] = fm(__int2float_rn(nR), half_area(rLo, rHi));
        }
        for (int i = 0; i < kBins - 1; i++) {
            const float c = fa(costL[i], costR[i]);
            if (c < best) { axis = a; plane = i + 1; best = c; }
        }
    }
    const float noSplit = fm(half_area(nd.bmin, nd.bmax), __uint2float_rn(nd.count));
    int split = !(best >= noSplit) ? 1 : 0;
    uint32_t L = 0;
    if (split) {
        for (int b = 0; b < plane; b++) L += cnt[axis * kBins + b];
        if (L == nd.count) split = 0;
        else if (L == 0) split = 2;
    }
    axisOut = axis; planeOut = plane; Lout = L;
    return split;
}

// segment of a leaf into the final index array; `rotated`: see decide_split
__device__ __forceinline__ uint32_t leaf_source(uint32_t first, uint32_t count, uint32_t i, bool rotated)
{
    return first + (rotated ? (i + 1 == count ? 0u : i + 1u) : i);
}

// per-thread bounds of the two children while placing: [lo hi clo chi] x {left, right}
struct ChildBounds {
    float b[24];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int k = 0; k < 24; k++) b[k] = ((k / 3) & 1) ? -1e30f : 1e30f;
    }
    __device__ __forceinline__ void add(bool left, const float4& v0, const float4& v1, const float4& v2, const float4& ce)
    {
        float* q = left ? b : b + 12;
        q[0] = fminf(q[0], fminf(fminf(v0.x, v1.x), v2.x)); q[3] = fmaxf(q[3], fmaxf(fmaxf(v0.x, v1.x), v2.x));
        q[1] = fminf(q[1], fminf(fminf(v0.y, v1.y), v2.y)); q[4] = fmaxf(q[4], fmaxf(fmaxf(v0.y, v1.y), v2.y));
        q[2] = fminf(q[2], fminf(fminf(v0.z, v1.z), v2.z)); q[5] = fmaxf(q[5], fmaxf(fmaxf(v0.z, v1.z), v2.z));
        q[6] = fminf(q[6], ce.x); q[9] = fmaxf(q[9], ce.x);
        q[7] = fminf(q[7], ce.y); q[10] = fmaxf(q[10], ce.y);
        q[8] = fminf(q[8], ce.z); q[11] = fmaxf(q[11], ce.z);
    }
    // block reduction into red[0][0..24)
    template <int THREADS>
    __device__ __forceinline__ void reduce(float (*red)[24])
    {
        const int tid = threadIdx.x;
#pragma unroll
        for (int k = 0; k < 24; k++) {
            const bool isMax = ((k % 12) / 3) & 1;
            const float v = isMax ? warp_max(b[k]) : warp_min(b[k]);
            if ((tid & 31) == 0) red[tid >> 5][k] = v;
        }
        __syncthreads();
        if (tid < 24) {
            const bool isMax = ((tid % 12) / 3) & 1;
            float v = red[0][tid];
            for (int w = 1; w < THREADS / 32; w++) v = isMax ? fmaxf(v, red[w][tid]) : fminf(v, red[w][tid]);
            red[0][tid] = v;
        }
        __syncthreads();
    }
};

// shared-memory bins of one block: counts and order-encoded vertex bounds per axis and bin
struct BlockBins {
    uint32_t cnt[3 * kBins];
    uint32_t lo[3 * kBins * 3], hi[3 * kBins * 3];
    template <int THREADS>
    __device__ __forceinline__ void clear()
    {
        for (int k = threadIdx.x; k < 3 * kBins; k += THREADS) cnt[k] = 0;
        for (int k = threadIdx.x; k < 3 * kBins * 3; k += THREADS) { lo[k] = 0xffffffffu; hi[k] = 0u; }
    }
    __device__ __forceinline__ void add(const BNode& nd, const float* scale, const float4& v0, const float4& v1, const float4& v2, const float4& ce)
    {
        const float tlo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
        const float thi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
        const float c[3] = {ce.x, ce.y, ce.z};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (nd.cmin[a] == nd.cmax[a]) continue;
            const int b = a * kBins + bin_of(c[a], nd.cmin[a], scale[a]);
            atomicAdd(&cnt[b], 1u);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // the bounds are monotone, so a plain look first saves most of the atomics (a stale value only
                // costs an atomic that changes nothing): after the first few triangles a bin's box rarely grows
                const uint32_t el = enc(tlo[k]), eh = enc(thi[k]);
                if (el < *(volatile uint32_t*)&lo[b * 3 + k]) atomicMin(&lo[b * 3 + k], el);
                if (eh > *(volatile uint32_t*)&hi[b * 3 + k]) atomicMax(&hi[b * 3 + k], eh);
            }
        }
    }
};

__device__ __forceinline__ void bin_scales(const BNode& nd, float* scale)
{
#pragma unroll
    for (int a = 0; a < 3; a++) scale[a] = nd.cmin[a] != nd.cmax[a] ? __fdiv_rn(8.0f, fs(nd.cmax[a], nd.cmin[a])) : 0.0f;
}

// One level of the build for nodes handled by one block each (THREADS = 32: up to kMid triangles,
// THREADS = 256: up to kHuge).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_level(const uint32_t* __restrict__ levelList, BNode* __restrict__ nodes,
                                                   const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                   uint32_t* __restrict__ dst, uint32_t* __restrict__ finalIdx,
                                                   uint32_t* __restrict__ rankScratch, uint32_t* __restrict__ holePos,
                                                   uint32_t* __restrict__ sleftPos, Lists q)
{
    __shared__ BlockBins bins;
    __shared__ uint32_t warpSums[THREADS / 32];
    __shared__ float red[THREADS / 32][24];
    __shared__ int sAxis, sPlane, sSplit;
    __shared__ uint32_t sL, sChild;

    const uint32_t nodeId = levelList[blockIdx.x];
    const BNode nd = nodes[nodeId];
    const uint32_t first = nd.first, count = nd.count;
    const int tid = threadIdx.x;

    // ---- 1. bins (bvh.cpp:108-129) ----
    bins.clear<THREADS>();
    __syncthreads();
    float scale[3];
    bin_scales(nd, scale);
    for (uint32_t i = tid; i < count; i += THREADS) {
        const uint32_t t = src[first + i];
        bins.add(nd, scale, __ldg(tris + 4ull * t), __ldg(tris + 4ull * t + 1), __ldg(tris + 4ull * t + 2), __ldg(tris + 4ull * t + 3));
    }
    __syncthreads();

    // ---- 2. sweep and decision, one thread ----
    if (tid == 0) {
        int axis, plane;
        uint32_t L;
        const int split = decide_split(bins.cnt, bins.lo, bins.hi, nd, axis, plane, L);
        sAxis = axis; sPlane = plane; sSplit = split; sL = L;
        if (split == 1) {
            const uint32_t child = atomicAdd(&q.counters[0], 2u);
            sChild = child;
            enqueue_child(child, first, L, q);
            enqueue_child(child + 1, first + L, count - L, q);
        }
    }
    __syncthreads();

    if (sSplit != 1) {
        // leaf: its segment is final
        for (uint32_t i = tid; i < count; i += THREADS) finalIdx[first + i] = src[leaf_source(first, count, i, sSplit == 2)];
        return;
    }

    // ---- 3. partition (bvh.cpp:56-63) through the closed form of the swap loop ----
    const int axis = sAxis, plane = sPlane;
    const uint32_t L = sL, last = first + count - 1;
    const float lo = nd.cmin[axis];
    const float sc = __fdiv_rn(8.0f, fs(nd.cmax[axis], lo));
    auto is_left = [&](uint32_t t) -> bool {
        const float4 ce = __ldg(tris + 4ull * t + 3);
        const float c = axis == 0 ? ce.x : (axis == 1 ? ce.y : ce.z);
        return bin_of(c, lo, sc) < plane;
    };
    // prefix [first, first+L): rank the holes from the left
    uint32_t carry = 0;
    for (uint32_t base = 0; base < L; base += THREADS) {
        const uint32_t i = base + tid;
        uint32_t flag = 0;
        if (i < L) flag = is_left(src[first + i]) ? 0u : 1u;
        uint32_t tot;
        const uint32_t r = block_exclusive_scan<THREADS>(flag, warpSums, tot) + carry;
        if (i < L) {
            rankScratch[first + i] = r;
            if (flag) holePos[first + r] = first + i;
        }
        carry += tot;
    }
    const uint32_t h = carry;
    // suffix (first+L .. last], walked from the right end: rank the left elements
    carry = 0;
    const uint32_t nSuf = count - L;
    for (uint32_t base = 0; base < nSuf; base += THREADS) {
        const uint32_t s = base + tid;               // back index
        uint32_t flag = 0;
        if (s < nSuf) flag = is_left(src[last - s]) ? 1u : 0u;
        uint32_t tot;
        const uint32_t r = block_exclusive_scan<THREADS>(flag, warpSums, tot) + carry;
        if (s < nSuf) {
            rankScratch[last - s] = r;
            if (flag) sleftPos[first + r] = last - s;
        }
        carry += tot;
    }
    __syncthreads();

    // placement + children's bounds (UpdateNodeBounds of both children)
    ChildBounds cb;
    cb.init();
    for (uint32_t i = tid; i < count; i += THREADS) {
        const uint32_t p = first + i;
        const uint32_t t = src[p];
        const float4 v0 = __ldg(tris + 4ull * t), v1 = __ldg(tris + 4ull * t + 1), v2 = __ldg(tris + 4ull * t + 2), ce = __ldg(tris + 4ull * t + 3);
        const float c = axis == 0 ? ce.x : (axis == 1 ? ce.y : ce.z);
        const bool left = bin_of(c, lo, sc) < plane;
        const uint32_t r = rankScratch[p];
        uint32_t d;
        if (i < L) {
            if (left) d = p;
            else d = (r == 0) ? last : sleftPos[first + r - 1] - 1;
        } else {
            if (left) d = holePos[first + r];
            else if (i == L) d = (h == 0) ? last : sleftPos[first + h - 1] - 1;
            else d = p - 1;
        }
        dst[d] = t;
        cb.add(left, v0, v1, v2, ce);
    }
    cb.reduce<THREADS>(red);
    if (tid < 2) {
        const float* qv = red[0] + 12 * tid;
        BNode c;
        for (int k = 0; k < 3; k++) { c.bmin[k] = qv[k]; c.bmax[k] = qv[3 + k]; c.cmin[k] = qv[6 + k]; c.cmax[k] = qv[9 + k]; }
        c.first = tid == 0 ? first : first + L;
        c.count = tid == 0 ? L : count - L;
        c.left = 0;
        c.depth = nd.depth + 1;
        nodes[sChild + tid] = c;
    }
    if (tid == 0) nodes[nodeId].left = sChild;
}

// ---- huge nodes: a grid of chunk blocks per node ------------------------------------------------------
struct HugeState {              // per huge node of the level ("slot")
    uint32_t cnt[3 * kBins];
    uint32_t lo[3 * kBins * 3], hi[3 * kBins * 3];
    uint32_t childAcc[24];      // order-encoded bounds of the two children
    uint32_t axis, plane, split, L, child, h, pad0, pad1;
};

// chunk0 per node (exclusive scan of the chunk counts), chunk -> slot map, state reset.  One block.
__global__ void __launch_bounds__(256) k_huge_prepare(HugeInfo* __restrict__ info, int nHuge, uint32_t* __restrict__ chunkMap,
                                                      HugeState* __restrict__ st, uint32_t* __restrict__ totalChunks)
{
    __shared__ uint32_t warpSums[8];
    uint32_t carry = 0;
    for (int base = 0; base < nHuge; base += 256) {
        const int s = base + threadIdx.x;
        const uint32_t nc = s < nHuge ? (info[s].count + kChunk - 1) / kChunk : 0u;
        uint32_t tot;
        const uint32_t c0 = block_exclusive_scan<256>(nc, warpSums, tot) + carry;
        if (s < nHuge) {
            info[s].chunk0 = c0;
            for (uint32_t k = 0; k < nc; k++) chunkMap[c0 + k] = (uint32_t)s;
            HugeState& h = st[s];
            for (int k = 0; k < 3 * kBins; k++) h.cnt[k] = 0;
            for (int k = 0; k < 3 * kBins * 3; k++) { h.lo[k] = 0xffffffffu; h.hi[k] = 0u; }
            for (int k = 0; k < 24; k++) h.childAcc[k] = ((k / 3) & 1) ? 0u : 0xffffffffu;
            h.split = 0;
        }
        carry += tot;
    }
    if (threadIdx.x == 0) *totalChunks = carry;
}

#define UVRT_HUGE_CHUNK_PROLOGUE()                                                                         \
    if (blockIdx.x >= *totalChunks) return;                                                                \
    const uint32_t slot = chunkMap[blockIdx.x];                                                            \
    const HugeInfo hi_ = info[slot];                                                                       \
    const uint32_t first = hi_.first, count = hi_.count;                                                   \
    const uint32_t c0 = (blockIdx.x - hi_.chunk0) * kChunk;            /* chunk range inside the node */   \
    const uint32_t c1 = c0 + kChunk < count ? c0 + kChunk : count;                                         \
    const int tid = threadIdx.x;

// 1. bins of a chunk, merged into the node's bins
__global__ void __launch_bounds__(256) k_huge_bins(const HugeInfo* __restrict__ info, const uint32_t* __restrict__ chunkMap,
                                                   const uint32_t* __restrict__ totalChunks, const BNode* __restrict__ nodes,
                                                   const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                   HugeState* __restrict__ st)
{
    __shared__ BlockBins bins;
    UVRT_HUGE_CHUNK_PROLOGUE()
    const BNode nd = nodes[hi_.id];
    bins.clear<256>();
    __syncthreads();
    float scale[3];
    bin_scales(nd, scale);
    for (uint32_t i = c0 + tid; i < c1; i += 256) {
        const uint32_t t = src[first + i];
        bins.add(nd, scale, __ldg(tris + 4ull * t), __ldg(tris + 4ull * t + 1), __ldg(tris + 4ull * t + 2), __ldg(tris + 4ull * t + 3));
    }
    __syncthreads();
    HugeState& h = st[slot];
    for (int k = tid; k < 3 * kBins; k += 256)
        if (bins.cnt[k]) atomicAdd(&h.cnt[k], bins.cnt[k]);
    for (int k = tid; k < 3 * kBins * 3; k += 256) {
        if (bins.lo[k] != 0xffffffffu) atomicMin(&h.lo[k], bins.lo[k]);
        if (bins.hi[k] != 0u) atomicMax(&h.hi[k], bins.hi[k]);
    }
}

// 2. decision, one thread per node
__global__ void __launch_bounds__(32) k_huge_decide(const HugeInfo* __restrict__ info, int nHuge, const BNode* __restrict__ nodes,
                                                    HugeState* __restrict__ st, Lists q)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nHuge) return;
    const HugeInfo hi_ = info[s];
    const BNode nd = nodes[hi_.id];
    HugeState& h = st[s];
    int axis, plane;
    uint32_t L;
    const int split = decide_split(h.cnt, h.lo, h.hi, nd, axis, plane, L);
    h.axis = (uint32_t)axis; h.plane = (uint32_t)plane; h.split = split == 1 ? 1u : 0u; h.L = L; h.h = 0;
    h.pad0 = split == 2 ? 1u : 0u;      // leaf with a rotated segment
    if (split == 1) {
        const uint32_t child = atomicAdd(&q.counters[0], 2u);
        h.child = child;
        enqueue_child(child, hi_.first, L, q);
        enqueue_child(child + 1, hi_.first + L, hi_.count - L, q);
    }
}

struct SplitTest {               // bin(centroid[axis]) < plane, the partition predicate
    int axis, plane; float lo, sc;
    __device__ __forceinline__ void set(const BNode& nd, uint32_t axis_, uint32_t plane_)
    {
        axis = (int)axis_; plane = (int)plane_; lo = nd.cmin[axis_];
        sc = __fdiv_rn(8.0f, fs(nd.cmax[axis_], nd.cmin[axis_]));
    }
    __device__ __forceinline__ bool left(const float4& ce) const
    {
        const float c = axis == 0 ? ce.x : (axis == 1 ? ce.y : ce.z);
        return bin_of(c, lo, sc) < plane;
    }
};

// 3. per chunk: holes of the prefix [0, L) and left elements of the suffix [L, count)
__global__ void __launch_bounds__(256) k_huge_count(const HugeInfo* __restrict__ info, const uint32_t* __restrict__ chunkMap,
                                                    const uint32_t* __restrict__ totalChunks, const BNode* __restrict__ nodes,
                                                    const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                    const HugeState* __restrict__ st, uint2* __restrict__ chunkCnt)
{
    __shared__ uint32_t sH, sS;
    UVRT_HUGE_CHUNK_PROLOGUE()
    const HugeState& h = st[slot];
    if (!h.split) return;
    SplitTest test;
    test.set(nodes[hi_.id], h.axis, h.plane);
    const uint32_t L = h.L;
    if (tid == 0) { sH = 0; sS = 0; }
    __syncthreads();
    uint32_t holes = 0, slefts = 0;
    for (uint32_t i = c0 + tid; i < c1; i += 256) {
        const bool left = test.left(__ldg(tris + 4ull * src[first + i] + 3));
        holes += (i < L && !left) ? 1u : 0u;
        slefts += (i >= L && left) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) { holes += __shfl_xor_sync(0xffffffffu, holes, o); slefts += __shfl_xor_sync(0xffffffffu, slefts, o); }
    if ((tid & 31) == 0) { atomicAdd(&sH, holes); atomicAdd(&sS, slefts); }
    __syncthreads();
    if (tid == 0) chunkCnt[blockIdx.x] = make_uint2(sH, sS);
}

// 4. per node: exclusive scan of the chunk counts -- holes left to right, suffix-lefts right to left
__global__ void __launch_bounds__(256) k_huge_scan(const HugeInfo* __restrict__ info, HugeState* __restrict__ st,
                                                   const uint2* __restrict__ chunkCnt, uint2* __restrict__ chunkOff)
{
    __shared__ uint32_t warpSums[8];
    const uint32_t slot = blockIdx.x;
    if (!st[slot].split) return;
    const HugeInfo hi_ = info[slot];
    const uint32_t nc = (hi_.count + kChunk - 1) / kChunk;
    uint32_t carryH = 0, carryS = 0;
    for (uint32_t base = 0; base < nc; base += 256) {
        const uint32_t k = base + threadIdx.x;
        const uint32_t cl = hi_.chunk0 + k, cr = hi_.chunk0 + nc - 1 - k;
        const uint32_t vh = k < nc ? chunkCnt[cl].x : 0u, vs = k < nc ? chunkCnt[cr].y : 0u;
        uint32_t totH, totS;
        const uint32_t eh = block_exclusive_scan<256>(vh, warpSums, totH) + carryH;
        const uint32_t es = block_exclusive_scan<256>(vs, warpSums, totS) + carryS;
        if (k < nc) { chunkOff[cl].x = eh; chunkOff[cr].y = es; }
        carryH += totH;
        carryS += totS;
    }
    if (threadIdx.x == 0) st[slot].h = carryH;
}

// 5. ranks: k-th hole from the left, k-th suffix-left from the right, and their positions
__global__ void __launch_bounds__(256) k_huge_rank(const HugeInfo* __restrict__ info, const uint32_t* __restrict__ chunkMap,
                                                   const uint32_t* __restrict__ totalChunks, const BNode* __restrict__ nodes,
                                                   const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                   const HugeState* __restrict__ st, const uint2* __restrict__ chunkCnt,
                                                   const uint2* __restrict__ chunkOff, uint32_t* __restrict__ rankScratch,
                                                   uint32_t* __restrict__ holePos, uint32_t* __restrict__ sleftPos)
{
    __shared__ uint32_t warpSums[8];
    UVRT_HUGE_CHUNK_PROLOGUE()
    const HugeState& h = st[slot];
    if (!h.split) return;
    SplitTest test;
    test.set(nodes[hi_.id], h.axis, h.plane);
    const uint32_t L = h.L;
    const uint2 off = chunkOff[blockIdx.x];
    const uint32_t sInChunk = chunkCnt[blockIdx.x].y;
    uint32_t carry = 0;                              // low 16 bits: holes so far, high 16: suffix-lefts so far
    for (uint32_t base = c0; base < c1; base += 256) {
        const uint32_t i = base + tid;
        uint32_t flag = 0;
        if (i < c1) {
            const bool left = test.left(__ldg(tris + 4ull * src[first + i] + 3));
            flag = (i < L && !left) ? 1u : ((i >= L && left) ? 0x10000u : 0u);
        }
        uint32_t tot;
        const uint32_t e = block_exclusive_scan<256>(flag, warpSums, tot) + carry;
        if (flag == 1u) {
            const uint32_t r = off.x + (e & 0xffffu);
            rankScratch[first + i] = r;
            holePos[first + r] = first + i;
        } else if (flag) {
            const uint32_t r = off.y + (sInChunk - 1u - (e >> 16));
            rankScratch[first + i] = r;
            sleftPos[first + r] = first + i;
        }
        carry += tot;
    }
}

// 6. placement through the closed form + children's bounds; an unsplit node's segment is final
__global__ void __launch_bounds__(256) k_huge_place(const HugeInfo* __restrict__ info, const uint32_t* __restrict__ chunkMap,
                                                    const uint32_t* __restrict__ totalChunks, const BNode* __restrict__ nodes,
                                                    const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                    uint32_t* __restrict__ dst, uint32_t* __restrict__ finalIdx,
                                                    HugeState* __restrict__ st, const uint32_t* __restrict__ rankScratch,
                                                    const uint32_t* __restrict__ holePos, const uint32_t* __restrict__ sleftPos)
{
    __shared__ float red[8][24];
    UVRT_HUGE_CHUNK_PROLOGUE()
    HugeState& h = st[slot];
    if (!h.split) {
        for (uint32_t i = c0 + tid; i < c1; i += 256) finalIdx[first + i] = src[leaf_source(first, count, i, h.pad0 != 0)];
        return;
    }
    SplitTest test;
    test.set(nodes[hi_.id], h.axis, h.plane);
    const uint32_t L = h.L, nh = h.h, last = first + count - 1;
    ChildBounds cb;
    cb.init();
    for (uint32_t i = c0 + tid; i < c1; i += 256) {
        const uint32_t p = first + i;
        const uint32_t t = src[p];
        const float4 v0 = __ldg(tris + 4ull * t), v1 = __ldg(tris + 4ull * t + 1), v2 = __ldg(tris + 4ull * t + 2), ce = __ldg(tris + 4ull * t + 3);
        const bool left = test.left(ce);
        uint32_t d;
        if (i < L) {
            if (left) d = p;
            else { const uint32_t r = rankScratch[p]; d = (r == 0) ? last : sleftPos[first + r - 1] - 1; }
        } else {
            if (left) d = holePos[first + rankScratch[p]];
            else if (i == L) d = (nh == 0) ? last : sleftPos[first + nh - 1] - 1;
            else d = p - 1;
        }
        dst[d] = t;
        cb.add(left, v0, v1, v2, ce);
    }
    cb.reduce<256>(red);
    if (tid < 24) {
        const bool isMax = ((tid % 12) / 3) & 1;
        const uint32_t e = enc(red[0][tid]);
        if (isMax) { if (e != 0u) atomicMax(&h.childAcc[tid], e); }
        else if (e != 0xffffffffu) atomicMin(&h.childAcc[tid], e);
    }
}

// 7. children of the split huge nodes
__global__ void __launch_bounds__(32) k_huge_finish(const HugeInfo* __restrict__ info, int nHuge, BNode* __restrict__ nodes,
                                                    const HugeState* __restrict__ st)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nHuge) return;
    const HugeState& h = st[s];
    if (!h.split) return;
    const HugeInfo hi_ = info[s];
    const uint32_t depth = nodes[hi_.id].depth;
    for (int side = 0; side < 2; side++) {
        const uint32_t* a = h.childAcc + 12 * side;
        BNode c;
        for (int k = 0; k < 3; k++) { c.bmin[k] = dec(a[k]); c.bmax[k] = dec(a[3 + k]); c.cmin[k] = dec(a[6 + k]); c.cmax[k] = dec(a[9 + k]); }
        c.first = side ? hi_.first + h.L : hi_.first;
        c.count = side ? hi_.count - h.L : h.L;
        c.left = 0;
        c.depth = depth + 1;
        nodes[h.child + side] = c;
    }
    nodes[hi_.id].left = h.child;
}
#undef UVRT_HUGE_CHUNK_PROLOGUE

// One level for small nodes (count <= kSmall): one thread per node runs the host algorithm as is --
// bins, sweep, the swap loop itself, children's bounds -- so nothing needs to be re-derived.
__global__ void __launch_bounds__(128) k_level_small(const uint32_t* __restrict__ list, int nList, BNode* __restrict__ nodes,
                                                     const float4* __restrict__ tris, const uint32_t* __restrict__ src,
                                                     uint32_t* __restrict__ dst, uint32_t* __restrict__ finalIdx, Lists q)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nList) return;
    const uint32_t nodeId = list[li];
    const BNode nd = nodes[nodeId];
    const uint32_t first = nd.first, count = nd.count;
    uint32_t idx[kSmall];
    for (uint32_t i = 0; i < count; i++) idx[i] = src[first + i];

    float best = 1e30f;
    int axis = 0, plane = 0;
    for (int a = 0; a < 3; a++) {
        const float lo = nd.cmin[a], hi = nd.cmax[a];
        if (lo == hi) continue;
        const float scale = __fdiv_rn(8.0f, fs(hi, lo));
        float binLo[kBins][3], binHi[kBins][3];
        int cnt[kBins];
        for (int b = 0; b < kBins; b++) {
            cnt[b] = 0;
            for (int k = 0; k < 3; k++) { binLo[b][k] = 1e30f; binHi[b][k] = -1e30f; }
        }
        for (uint32_t i = 0; i < count; i++) {
            const uint32_t t = idx[i];
            const float4 v0 = __ldg(tris + 4ull * t), v1 = __ldg(tris + 4ull * t + 1), v2 = __ldg(tris + 4ull * t + 2), ce = __ldg(tris + 4ull * t + 3);
            const float c = a == 0 ? ce.x : (a == 1 ? ce.y : ce.z);
            const int b = bin_of(c, lo, scale);
            cnt[b]++;
            binLo[b][0] = fminf(binLo[b][0], fminf(fminf(v0.x, v1.x), v2.x)); binHi[b][0] = fmaxf(binHi[b][0], fmaxf(fmaxf(v0.x, v1.x), v2.x));
            binLo[b][1] = fminf(binLo[b][1], fminf(fminf(v0.y, v1.y), v2.y)); binHi[b][1] = fmaxf(binHi[b][1], fmaxf(fmaxf(v0.y, v1.y), v2.y));
            binLo[b][2] = fminf(binLo[b][2], fminf(fminf(v0.z, v1.z), v2.z)); binHi[b][2] = fmaxf(binHi[b][2], fmaxf(fmaxf(v0.z, v1.z), v2.z));
        }
        float costL[kBins - 1], costR[kBins - 1];
        float lLo[3] = {1e30f, 1e30f, 1e30f}, lHi[3] = {-1e30f, -1e30f, -1e30f};
        float rLo[3] = {1e30f, 1e30f, 1e30f}, rHi[3] = {-1e30f, -1e30f, -1e30f};
        int nL = 0, nR = 0;
        for (int i = 0; i < kBins - 1; i++) {
            nL += cnt[i];
            for (int k = 0; k < 3; k++) {
                lLo[k] = binLo[i][k] < lLo[k] ? binLo[i][k] : lLo[k];
                lHi[k] = binHi[i][k] > lHi[k] ? binHi[i][k] : lHi[k];
            }
            costL[i] = fm(__int2float_rn(nL), half_area(lLo, lHi));
            const int rb = kBins - 2 - i;
            nR += cnt[rb + 1];
            for (int k = 0; k < 3; k++) {
                rLo[k] = binLo[rb][k] < rLo[k] ? binLo[rb][k] : rLo[k];
                rHi[k] = binHi[rb][k] > rHi[k] ? binHi[rb][k] : rHi[k];
            }
            costR[rb] = fm(__int2float_rn(nR), half_area(rLo, rHi));
        }
        for (int i = 0; i < kBins - 1; i++) {
            const float c = fa(costL[i], costR[i]);
            if (c < best) { axis = a; plane = i + 1; best = c; }
        }
    }
    const float noSplit = fm(half_area(nd.bmin, nd.bmax), __uint2float_rn(count));
    bool split = !(best >= noSplit);
    int i = 0, j = (int)count - 1;
    if (split) {
        const float lo = nd.cmin[axis];
        const float sc = __fdiv_rn(8.0f, fs(nd.cmax[axis], lo));
        while (i <= j) {
            const float4 ce = __ldg(tris + 4ull * idx[i] + 3);
            const float c = axis == 0 ? ce.x : (axis == 1 ? ce.y : ce.z);
            if (bin_of(c, lo, sc) < plane) i++;
            else { const uint32_t t = idx[i]; idx[i] = idx[j]; idx[j] = t; j--; }
        }
        if (i == 0 || i == (int)count) {
            // abandoned after the partition loop ran: the reference keeps the permuted order (bvh.cpp:56-66)
            for (uint32_t k = 0; k < count; k++) finalIdx[first + k] = idx[k];
            return;
        }
    }
    if (!split) {
        for (uint32_t k = 0; k < count; k++) finalIdx[first + k] = src[first + k];
        return;
    }
    const uint32_t L = (uint32_t)i;
    for (uint32_t k = 0; k < count; k++) dst[first + k] = idx[k];
    const uint32_t child = atomicAdd(&q.counters[0], 2u);
    for (int side = 0; side < 2; side++) {
        BNode c;
        for (int k = 0; k < 3; k++) { c.bmin[k] = 1e30f; c.bmax[k] = -1e30f; c.cmin[k] = 1e30f; c.cmax[k] = -1e30f; }
        const uint32_t b0 = side ? L : 0, b1 = side ? count : L;
        for (uint32_t k = b0; k < b1; k++) {
            const uint32_t t = idx[k];
            const float4 v0 = __ldg(tris + 4ull * t), v1 = __ldg(tris + 4ull * t + 1), v2 = __ldg(tris + 4ull * t + 2), ce = __ldg(tris + 4ull * t + 3);
            c.bmin[0] = fminf(c.bmin[0], fminf(fminf(v0.x, v1.x), v2.x)); c.bmax[0] = fmaxf(c.bmax[0], fmaxf(fmaxf(v0.x, v1.x), v2.x));
            c.bmin[1] = fminf(c.bmin[1], fminf(fminf(v0.y, v1.y), v2.y)); c.bmax[1] = fmaxf(c.bmax[1], fmaxf(fmaxf(v0.y, v1.y), v2.y));
            c.bmin[2] = fminf(c.bmin[2], fminf(fminf(v0.z, v1.z), v2.z)); c.bmax[2] = fmaxf(c.bmax[2], fmaxf(fmaxf(v0.z, v1.z), v2.z));
            c.cmin[0] = fminf(c.cmin[0], ce.x); c.cmax[0] = fmaxf(c.cmax[0], ce.x);
            c.cmin[1] = fminf(c.cmin[1], ce.y); c.cmax[1] = fmaxf(c.cmax[1], ce.y);
            c.cmin[2] = fminf(c.cmin[2], ce.z); c.cmax[2] = fmaxf(c.cmax[2], ce.z);
        }
        c.first = first + b0;
        c.count = b1 - b0;
        c.left = 0;
        c.depth = nd.depth + 1;
        nodes[child + side] = c;
        enqueue_child(child + side, c.first, c.count, q);
    }
    nodes[nodeId].left = child;
}

// ---- renumbering into the reference's node order ---------------------------------------------------
// bottom-up: slots allocated inside a recursively numbered subtree
// (the nodes of one level own a contiguous range of temp ids: they are all allocated while their
// parents' level is processed)
__global__ void __launch_bounds__(256) k_sizes(uint32_t idBegin, int n, const BNode* __restrict__ nodes, BAux* __restrict__ aux)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = idBegin + (uint32_t)i;
    const uint32_t l = nodes[id].left;
    aux[id].sFull = l ? 2u + aux[l].sFull + aux[l + 1].sFull : 0u;
}

// The first four levels (nodes of depth <= 3 are numbered serially, their grandchildren at depth 4
// are the roots of the independent jobs), by one thread -- at most 15 + 16 nodes.
__global__ void k_number_top(const BNode* __restrict__ nodes, BAux* __restrict__ aux, uint32_t* __restrict__ usedMax)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // depth-first walk mirroring bvh.cpp:46-96 with `depth == 3` deferral
    uint32_t stackId[40];
    int sp = 0;
    uint32_t nextFree = 2;
    uint32_t jobs[16];
    int nJobs = 0;
    aux[0].ref = 0;
    // explicit recursion: visit(node): if split { allocate pair; left: recurse or defer; right: recurse or defer }
    // pre-order with the right child pushed first reproduces "left subtree fully numbered before the right one"
    stackId[sp++] = 0;
    while (sp) {
        const uint32_t id = stackId[--sp];
        const BNode nd = nodes[id];
        if (!nd.left) continue;
        aux[nd.left].ref = nextFree;
        aux[nd.left + 1].ref = nextFree + 1;
        nextFree += 2;
        if (nd.depth == 3) {
            jobs[nJobs++] = nd.left;
            jobs[nJobs++] = nd.left + 1;
        } else {
            stackId[sp++] = nd.left + 1;
            stackId[sp++] = nd.left;
        }
    }
    // NOTE: the host walk allocates a node's pair BEFORE descending, and descends left first; the stack
    // order above visits the left child's subtree completely before the right child is popped.
    uint32_t base = nextFree, used = nextFree;
    for (int j = 0; j < nJobs; j++) {
        aux[jobs[j]].alloc = base;
        const uint32_t end = base + aux[jobs[j]].sFull;
        used = end > used ? end : used;
        base += nodes[jobs[j]].count * 2u;
    }
    *usedMax = used;
}

// top-down inside the jobs (depth >= 4): children pair at alloc, left subtree numbered first
__global__ void __launch_bounds__(256) k_number_level(uint32_t idBegin, int n, const BNode* __restrict__ nodes, BAux* __restrict__ aux)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = idBegin + (uint32_t)i;
    const BNode nd = nodes[id];
    if (nd.depth < 4 || !nd.left) return;
    const uint32_t a = aux[id].alloc;
    aux[nd.left].ref = a;
    aux[nd.left + 1].ref = a + 1;
    aux[nd.left].alloc = a + 2;
    aux[nd.left + 1].alloc = a + 2 + aux[nd.left].sFull;
}

// final node array in the reference layout (bvh.h:11-21)
__global__ void __launch_bounds__(256) k_emit(const BNode* __restrict__ nodes, const BAux* __restrict__ aux, int nTemp, float4* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nTemp) return;
    const BNode nd = nodes[i];
    const uint32_t r = aux[i].ref;
    const uint32_t leftFirst = nd.left ? aux[nd.left].ref : nd.first;
    const uint32_t triCount = nd.left ? 0u : nd.count;
    out[2ull * r] = make_float4(nd.bmin[0], nd.bmin[1], nd.bmin[2], __uint_as_float(leftFirst));
    out[2ull * r + 1] = make_float4(nd.bmax[0], nd.bmax[1], nd.bmax[2], __uint_as_float(triCount));
}

} // namespace uvrt_bvh
