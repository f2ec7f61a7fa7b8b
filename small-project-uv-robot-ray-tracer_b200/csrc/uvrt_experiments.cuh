// uvrt_experiments.cuh -- extend variants that were measured and REJECTED (profiles/r1_sweeps.md): persistent warps
// with a global ray queue (variant B, optional shared-memory hit table) and chunk-persistent warps (variant C).
// Not part of the product build: compiled only with -DUVRT_EXPERIMENTS (`make EXPERIMENTS=1`), selected with
// "extend_variant" 10..24 / 40..43.  Per-ray test sequences are those of variant A, so results are bit-identical.
#pragma once
#include "uvrt_kernels.cuh"

namespace uvrt {

// Variant B: persistent warps that pull rays from a global queue.  Lanes whose ray has finished
// are refilled together (one warp-aggregated atomicAdd on the queue head per refill) once fewer
// than REFILL lanes are still busy, so a warp is not held hostage by its longest ray.  Inside,
// every lane alternates between at most K inner-node steps and one leaf visit; the per-ray
// sequence of box tests, triangle tests and distance updates is exactly variant A's.
// HIST = 1 counts hits in a shared-memory table first (block-level pre-reduction of hot triangle
// IDs) and only spills colliding IDs and the final table to the global counters.
template <int HBITS>
struct HitTable {
    uint32_t tag[1 << HBITS];
    int cnt[1 << HBITS];
};

template <int HBITS>
__device__ __forceinline__ void hist_add(HitTable<HBITS>* tab, int* __restrict__ counts, uint32_t tri)
{
    uint32_t slot = (tri * 2654435761u) >> (32 - HBITS);
    uint32_t old = atomicCAS(&tab->tag[slot], 0xffffffffu, tri);
    if (old == 0xffffffffu || old == tri) atomicAdd(&tab->cnt[slot], 1);
    else atomicAdd(&counts[tri], 1);
}

template <int DIV, int STACK, int K, int REFILL, int HIST, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
k_extend_persist(int* __restrict__ counts, const float4* __restrict__ wtris, float4* __restrict__ rays,
                 const float4* __restrict__ pairs, uint32_t rootRef, uint32_t nRays, int sceneTame,
                 unsigned int* __restrict__ queueHead, const uint32_t* __restrict__ perm)
{
    constexpr int HBITS = 11;
    __shared__ HitTable<HIST ? HBITS : 1> tab;
    if (HIST) {
        for (int i = threadIdx.x; i < (1 << HBITS); i += THREADS) { tab.tag[i] = 0xffffffffu; tab.cnt[i] = 0; }
        __syncthreads();
    }
    uint32_t stack[STACK];
    const unsigned lane = threadIdx.x & 31u;
    RayCtx ray;
    ray.ox = ray.oy = ray.oz = ray.dx = ray.dy = ray.dz = 0.0f;
    ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
    ray.dist = kNoHit; ray.tri = 0;
    uint32_t cur = 0, rayIdx = 0xffffffffu;
    int sp = 0;
    bool busy = false, tame = false, drained = false;

    for (;;) {
        // ---- refill (warp-converged) ----
        unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (idle && !drained) {
            int leader = __ffs(idle) - 1;
            unsigned cnt = __popc(idle);
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(queueHead, cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!busy) {
                unsigned idx = base + __popc(idle & ((1u << lane) - 1u));
                if (base < nRays && idx < nRays) {
                    rayIdx = perm ? perm[idx] : idx;
                    load_ray(rays, rayIdx, ray);
                    tame = false;
                    if (DIV != DIV_IEEE && sceneTame && ray_is_tame(ray)) {
                        tame = true;
                        make_tame(ray);
                    }
                    cur = rootRef; sp = 0; busy = true;
                }
            }
            // base is warp-uniform: once the queue is past the end nobody asks again
            if (base >= nRays || nRays - base < cnt) drained = true;
        }
        unsigned act = __ballot_sync(0xffffffffu, busy);
        if (act == 0u) break;

        // ---- traverse until too few lanes are busy ----
        for (;;) {
#pragma unroll 1
            for (int k = 0; k < K && busy && !(cur & kLeafFlag); k++) {
                const float4* p = pairs + 4ull * cur;
                W4 ca = ldg256w(p), cb = ldg256w(p + 2);
                float t1, t2;
                bool h1, h2;
                if (DIV == DIV_IEEE || !tame) {
                    h1 = intersect_aabb<DIV_IEEE, -1>(ray, ca, t1);
                    h2 = intersect_aabb<DIV_IEEE, -1>(ray, cb, t2);
                } else {
                    h1 = intersect_aabb<DIV, -1>(ray, ca, t1);
                    h2 = intersect_aabb<DIV, -1>(ray, cb, t2);
                }
                uint32_t first, second;
                bool pushSecond;
                if (order_children(h1, h2, t1, t2, child_ref(ca), child_ref(cb), first, second, pushSecond)) {
                    cur = first;
                    if (pushSecond) stack[sp++] = second;
                } else {
                    if (sp == 0) busy = false; else cur = stack[--sp];
                }
            }
            if (busy && (cur & kLeafFlag)) {
                uint32_t slot = cur & ~kLeafFlag;
                uint32_t w;
                do {
                    const float4* t = wtris + 4ull * slot;
                    F8 ta = ldg256(t), tb = ldg256(t + 2);
                    w = __float_as_uint(ta.lo.w);
                    intersect_tri(ray, ta.lo, ta.hi, tb.lo);
                    slot++;
                } while (!(w & kLastFlag));
                if (sp == 0) busy = false; else cur = stack[--sp];
            }
            if (!busy && rayIdx != 0xffffffffu) {
                // the ray just finished: write back (extend.cl:26 writes in place) and count
                store_hit(rays, rayIdx, ray);
                if (ray.dist != kNoHit) {
                    if (HIST) hist_add<HBITS>(reinterpret_cast<HitTable<HBITS>*>(&tab), counts, ray.tri);
                    else atomicAdd(&counts[ray.tri], 1);
                }
                rayIdx = 0xffffffffu;
            }
            act = __ballot_sync(0xffffffffu, busy);
            if (act == 0u) break;
            if (!drained && __popc(act) < REFILL) break;
        }
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < (1 << HBITS); i += THREADS) {
            int c = tab.cnt[i];
            if (c) atomicAdd(&counts[tab.tag[i]], c);
        }
    }
}


// Variant C: persistent warps over PRIVATE chunks of the (binned) ray order.  Warp w owns rays
// [w*CH, (w+1)*CH) of the permutation; lanes whose ray has finished are refilled from the warp's own
// chunk (a warp-uniform cursor, no atomics) once fewer than `refill` lanes are busy.  Unlike variant B's
// global queue this keeps the rays of a warp neighbours in the binned order, so the coherence the
// binning created survives the refill; the tail where a warp waits for its longest ray is paid once per
// CH rays instead of once per 32.  Per-ray test sequence as in variant A.
template <int DIV, int STACK, int K, int CH, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
k_extend_chunk(int* __restrict__ counts, const float4* __restrict__ wtris, float4* __restrict__ rays,
               const float4* __restrict__ pairs, uint32_t rootRef, uint32_t nRays, int sceneTame,
               const uint32_t* __restrict__ perm, int refill)
{
    uint32_t stack[STACK];
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * THREADS + threadIdx.x) >> 5;
    uint32_t next = warp * (uint32_t)CH;
    const uint32_t end = min(next + (uint32_t)CH, nRays);
    if (next >= nRays) return;
    RayCtx ray;
    ray.ox = ray.oy = ray.oz = ray.dx = ray.dy = ray.dz = 0.0f;
    ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
    ray.dist = kNoHit; ray.tri = 0;
    uint32_t cur = 0, rayIdx = 0xffffffffu;
    int sp = 0;
    bool busy = false, tame = false;

    for (;;) {
        // ---- refill from the warp's own chunk (warp-converged) ----
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (idle && next < end) {
            if (!busy) {
                const uint32_t idx = next + __popc(idle & ((1u << lane) - 1u));
                if (idx < end) {
                    rayIdx = perm ? perm[idx] : idx;
                    load_ray(rays, rayIdx, ray);
                    tame = false;
                    if (DIV != DIV_IEEE && sceneTame && ray_is_tame(ray)) {
                        tame = true;
                        make_tame(ray);
                    }
                    cur = rootRef; sp = 0; busy = true;
                }
            }
            next = min(next + (uint32_t)__popc(idle), end);
        }
        unsigned act = __ballot_sync(0xffffffffu, busy);
        if (act == 0u) break;

        // ---- traverse until too few lanes are busy ----
        for (;;) {
#pragma unroll 1
            for (int k = 0; k < K && busy && !(cur & kLeafFlag); k++) {
                const float4* p = pairs + 4ull * cur;
                W4 ca = ldg256w(p), cb = ldg256w(p + 2);
                float t1, t2;
                bool h1, h2;
                if (DIV == DIV_IEEE || !tame) {
                    h1 = intersect_aabb<DIV_IEEE, -1>(ray, ca, t1);
                    h2 = intersect_aabb<DIV_IEEE, -1>(ray, cb, t2);
                } else {
                    h1 = intersect_aabb<DIV, -1>(ray, ca, t1);
                    h2 = intersect_aabb<DIV, -1>(ray, cb, t2);
                }
                uint32_t first, second;
                bool pushSecond;
                if (order_children(h1, h2, t1, t2, child_ref(ca), child_ref(cb), first, second, pushSecond)) {
                    cur = first;
                    if (pushSecond) stack[sp++] = second;
                } else {
                    if (sp == 0) busy = false; else cur = stack[--sp];
                }
            }
            if (busy && (cur & kLeafFlag)) {
                uint32_t slot = cur & ~kLeafFlag;
                uint32_t w;
                do {
                    const float4* t = wtris + 4ull * slot;
                    F8 ta = ldg256(t), tb = ldg256(t + 2);
                    w = __float_as_uint(ta.lo.w);
                    intersect_tri(ray, ta.lo, ta.hi, tb.lo);
                    slot++;
                } while (!(w & kLastFlag));
                if (sp == 0) busy = false; else cur = stack[--sp];
            }
            if (!busy && rayIdx != 0xffffffffu) {
                store_hit(rays, rayIdx, ray);
                if (ray.dist != kNoHit) atomicAdd(&counts[ray.tri], 1);
                rayIdx = 0xffffffffu;
            }
            act = __ballot_sync(0xffffffffu, busy);
            if (act == 0u) break;
            if (next < end && (int)__popc(act) < refill) break;
        }
    }
}

} // namespace uvrt
