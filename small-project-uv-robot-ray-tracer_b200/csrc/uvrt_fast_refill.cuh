// uvrt_fast_refill.cuh -- a REJECTED variant of the certified fast extend, compiled only with -DUVRT_EXPERIMENTS
// (make EXPERIMENTS=1) for A/B runs: "fast_cfg" 2.  Results: profiles/r2_fast_extend.md, profiles/r2_refill_*.jsonl
// (bit-identical to the exact kernel, 0 check mismatches; 3.3x slower on the room, 1.35-1.6x slower on the soups).
#pragma once
#include "uvrt_fast.cuh"

namespace uvrt {
// Why it loses (profiles/r2_fast_extend.md): the time added over k_extend_fast is 1.9 / 1.6-2.1 / 2.8 us per ray and warp
// on the room / 1 M soup / 10 M soup whatever the number of visits per ray -- the four latencies a lane's change of ray
// exposes (shared-memory draw, permutation entry, ray record, list append), each of which stalls the whole warp because
// divergent paths of a warp do not overlap, where the one-thread-per-ray kernel pays them once per 32 rays.  And the
// traversal itself gains little from fuller warps: a 32-byte gather costs one L1 wavefront per LANE, and the shipped
// kernel already runs at 65 % of that rate on the soup.
//
// The idea: for scenes where rays of a warp need very different numbers of visits (the 10 M-triangle
// soup: 310 inner visits per ray on average, from a handful to thousands), one thread per ray leaves most lanes of a warp
// idle most of the time (14.6 of 32 lanes per instruction, profiles/r2_extend_soup10m_metrics.json).  Here the grid is
// persistent and a LANE that finishes its ray takes the next one at once, while the other lanes of its warp keep
// traversing: the loop is flat (one inner-node or leaf step per round whatever ray a lane is on), nothing in it
// synchronises the warp, and the two things that used to happen per ray at the end -- the exact verification of the
// winner and the re-trace of uncertified rays -- move into two small follow-up kernels that run with full warps:
//
//   k_extend_fast_refill   conservative traversal only; appends (ray, winner slot, best, second) to `verify`,
//                          ineligible rays to `retry`
//   k_fast_verify          the certificate of fast_intersect(), per entry of `verify`: stores the hit and counts it,
//                          or appends the ray to `retry`
//   k_extend_retry         reference-order traversal of the rays in `retry`
//
// Rays are handed out in the binned order, a chunk of consecutive rays per warp at a time, so that the lanes of a warp
// stay on neighbouring rays: wq[warp] = (end << 32) | next of the warp's current chunk; a lane takes a ray with one
// shared-memory atomicAdd; the one lane that draws next == end fetches the next chunk from the global counter and
// publishes it, lanes that draw beyond it try again a round later (they never wait inside the round).
// The octant cannot be a template argument here (a lane's next ray may point elsewhere): node records use layout 1 and
// the near / far plane of each axis is picked by a PRMT selector held in a register, at the same instruction count.
struct RefillCtl { uint32_t next, nVerify, nRetry, pad; };     // zeroed before every launch
constexpr uint32_t kRetryVerifyOnly = 0x80000000u;              // retry entry: compare only ("fast_check")

struct FastSel { uint32_t nx, ny, nz, fx, fy, fz; };
// (PTX prmt and not __byte_perm, which masks a selector it cannot see with six more instructions per visit; the
// selectors used here are 0x4105 / 0x4325, no nibble has its replicate-sign bit set)
__device__ __forceinline__ float q_sel(uint32_t w, uint32_t sel)
{
    float r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=f"(r) : "r"(w), "r"(0x3Fu), "r"(sel));
    return r;
}

__device__ __forceinline__ bool fast_box_sel(const FastRay& fr, const FastSel& s, uint32_t wx, uint32_t wy, uint32_t wz, float dcull,
                                             float& tminOut)
{
    float nx, ny, fx, fy, nz, fz;
    upk2(fma2(pk2(q_sel(wx, s.nx), q_sel(wy, s.ny)), fr.sXY, fr.bXY), nx, ny);
    upk2(fma2(pk2(q_sel(wx, s.fx), q_sel(wy, s.fy)), fr.sXY, fr.bXY), fx, fy);
    upk2(fma2(pk2(q_sel(wz, s.nz), q_sel(wz, s.fz)), fr.sZZ, fr.bZZ), nz, fz);
    const float tmin = fmaxf(fmaxf(nx, ny), nz);
    const float tmax = fminf(fminf(fx, fy), fz);
    tminOut = tmin;
    return tmax >= tmin && tmin < dcull && tmax >= 0.0f;
}

__device__ __forceinline__ bool ray_in_grid_window_v(const RayCtx& r, const FastGrid& g)
{
    return r.ox >= g.lo[0] && r.ox <= g.hi[0] && r.oy >= g.lo[1] && r.oy <= g.hi[1] && r.oz >= g.lo[2] && r.oz <= g.hi[2];
}

template <int STACK, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
k_extend_fast_refill(const float4* __restrict__ wtris, const float4* __restrict__ rays, const uint4* __restrict__ qpairs,
                     const FastGrid grid, uint32_t nRays, const uint32_t* __restrict__ perm, RefillCtl* __restrict__ ctl,
                     uint4* __restrict__ verify, uint32_t* __restrict__ retry, uint32_t chunk, int checkMode,
                     FastStats* __restrict__ stats)
{
    __shared__ unsigned long long wq[THREADS / 32];
    __shared__ int wdone[THREADS / 32];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { wq[warp] = 0ull; wdone[warp] = 0; }
    __syncwarp();
    volatile int* done = &wdone[warp];

    uint32_t stack[STACK];
    int sp = 0;
    uint32_t cur = 0, bestSlot = 0, rayIdx = 0;
    float best = kNoHit, second = kNoHit, dcull = 3.0e38f;
    RayCtx ray;
    ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
    ray.ox = ray.oy = ray.oz = ray.dx = ray.dy = ray.dz = 0.0f; ray.dist = kNoHit; ray.tri = 0;
    FastRay fr;
    fr.sXY = fr.sZZ = fr.bXY = fr.bZZ = 0ull;
    FastSel sel;
    sel.nx = sel.ny = sel.nz = sel.fx = sel.fy = sel.fz = 0x4105u;
    // A lane between two rays takes three rounds to start the next one, so that neither of the two dependent loads
    // (permutation entry, ray record) is waited for inside a round the traversing lanes of the warp share:
    //   kDraw: take the next index, request the permutation entry     kLoad: request the ray record
    //   kSetup: eligibility, per-ray constants                        tracing: one inner-node or leaf step per round
    // (two variables, so that the test of the hot state is a plain predicate and not an entry of a jump table)
    enum { kDraw = 0, kLoad = 1, kSetup = 2 };
    int tracing = 0;
    int phase = kDraw;
    for (;;) {
        asm volatile("" : "+r"(tracing));      // keeps the compiler from merging the two into one switch
        if (tracing) {
            // nothing to fetch
        } else if (phase == kDraw) {
            if (*done) break;
            const unsigned long long old = atomicAdd(&wq[warp], 1ull);
            const uint32_t on = (uint32_t)old, oe = (uint32_t)(old >> 32);
            uint32_t k = on;
            bool got = on < oe;
            if (on == oe) {           // exactly one lane per published chunk draws next == end: it fetches the next chunk
                const uint32_t g = atomicAdd(&ctl->next, chunk);
                if (g >= nRays) { *done = 1; break; }
                const uint32_t ge = min(g + chunk, nRays);
                atomicExch(&wq[warp], ((unsigned long long)ge << 32) | (unsigned long long)(g + 1u));
                k = g;
                got = true;
            }
            if (got) {                // otherwise the chunk is being replaced: draw again next round
                rayIdx = perm ? ldg_u32_stream(perm + k) : k;
                phase = kLoad;
            }
        } else if (phase == kLoad) {
            const F8 r = ld256_stream(rays + 2ull * rayIdx);
            ray.dx = r.lo.x; ray.dy = r.lo.y; ray.dz = r.lo.z;
            ray.ox = r.lo.w; ray.oy = r.hi.x; ray.oz = r.hi.y;
            best = r.hi.z;            // the distance the ray arrives with, until kSetup has looked at it
            phase = kSetup;
        } else {
            // rays with a hit already recorded, rays that are not tame or start outside the origin window: exact kernel
            if (!(best == kNoHit && ray_is_tame(ray) && ray_in_grid_window_v(ray, grid))) {
                retry[atomicAdd(&ctl->nRetry, 1u)] = rayIdx;
                atomicAdd(&stats->fallbackIneligible, 1ull);
                phase = kDraw;
            } else {
                const float rx = __frcp_rn(ray.dx), ry = __frcp_rn(ray.dy), rz = __frcp_rn(ray.dz);
                const float sx = fm(fm(grid.step[0], 32768.0f), rx), sy = fm(fm(grid.step[1], 32768.0f), ry),
                            sz = fm(fm(grid.step[2], 32768.0f), rz);
                const float bx = __fmaf_rn(fs(grid.gmin[0], ray.ox), rx, -sx), by = __fmaf_rn(fs(grid.gmin[1], ray.oy), ry, -sy),
                            bz = __fmaf_rn(fs(grid.gmin[2], ray.oz), rz, -sz);
                fr.sXY = pk2(sx, sy); fr.sZZ = pk2(sz, sz);
                fr.bXY = pk2(bx, by); fr.bZZ = pk2(bz, bz);
                // low half-word = the box's lower plane, high half-word = its upper plane (layout 1)
                sel.nx = ray.dx < 0.0f ? 0x4325u : 0x4105u; sel.fx = ray.dx < 0.0f ? 0x4105u : 0x4325u;
                sel.ny = ray.dy < 0.0f ? 0x4325u : 0x4105u; sel.fy = ray.dy < 0.0f ? 0x4105u : 0x4325u;
                sel.nz = ray.dz < 0.0f ? 0x4325u : 0x4105u; sel.fz = ray.dz < 0.0f ? 0x4105u : 0x4325u;
                sp = 0; cur = 0; bestSlot = 0;
                best = second = kNoHit; dcull = 3.0e38f;
                tracing = 1;
            }
        }
        if (tracing) {
            bool finished = false;
            if (cur & kLeafFlag) {
                uint32_t slot = cur & ~kLeafFlag;
                uint32_t w;
                do {
                    const float4* t = wtris + 4ull * slot;
                    F8 ta = ldg256(t), tb = ldg256(t + 2);
                    w = __float_as_uint(ta.lo.w);
                    float tt;
                    if (tri_accept(ray, ta.lo, ta.hi, tb.lo, tt)) {
                        if (tt < best) {
                            second = best; best = tt; bestSlot = slot;
                            dcull = __fmaf_rn(best, 1.0f + 2.0f * kFastRel, 2.0f * kFastAbs);
                        } else
                            second = fminf(second, tt);
                    }
                    slot++;
                } while (!(w & kLastFlag));
                if (sp == 0) finished = true;
                else cur = stack[--sp];
            } else {
                uint4 ca, cb;
                asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(ca.x), "=r"(ca.y), "=r"(ca.z), "=r"(ca.w), "=r"(cb.x), "=r"(cb.y), "=r"(cb.z), "=r"(cb.w)
                    : "l"(qpairs + 2ull * cur));
                float t1, t2;
                const bool h1 = fast_box_sel(fr, sel, ca.x, ca.y, ca.z, dcull, t1);
                const bool h2 = fast_box_sel(fr, sel, cb.x, cb.y, cb.z, dcull, t2);
                uint32_t first, second_;
                bool pushSecond;
                if (order_children(h1, h2, t1, t2, ca.w, cb.w, first, second_, pushSecond)) {
                    cur = first;
                    if (pushSecond) stack[sp++] = second_;
                } else if (sp == 0)
                    finished = true;
                else
                    cur = stack[--sp];
            }
            if (finished) {
                tracing = 0;
                phase = kDraw;
                if (best != kNoHit)
                    verify[atomicAdd(&ctl->nVerify, 1u)] = make_uint4(rayIdx, bestSlot, __float_as_uint(best), __float_as_uint(second));
                else if (checkMode)   // certified miss: the ray record stays as it is; fast_check has it traced again
                    retry[atomicAdd(&ctl->nRetry, 1u)] = rayIdx | kRetryVerifyOnly;
            }
        }
    }
}

// The certificate of fast_intersect() for the winners k_extend_fast_refill found.
__global__ void __launch_bounds__(128) k_fast_verify(int* __restrict__ counts, const float4* __restrict__ wtris, float4* __restrict__ rays,
                                                     RefillCtl* __restrict__ ctl, const uint4* __restrict__ verify,
                                                     uint32_t* __restrict__ retry, FastStats* __restrict__ stats, int checkMode)
{
    const uint32_t n = ctl->nVerify;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint4 e = verify[j];
        const uint32_t i = e.x;
        const float best = __uint_as_float(e.z), second = __uint_as_float(e.w);
        RayCtx ray;
        {
            const F8 r = ld256_stream(rays + 2ull * i);
            ray.dx = r.lo.x; ray.dy = r.lo.y; ray.dz = r.lo.z;
            ray.ox = r.lo.w; ray.oy = r.hi.x; ray.oz = r.hi.y;
            ray.dist = r.hi.z;
            ray.tri = __float_as_uint(r.hi.w);
            ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
        }
        const float4* t = wtris + 4ull * e.y;
        const F8 ta = ldg256(t), tb = ldg256(t + 2);
        float tl;
        const bool reachable = leaf_box_exact<-1>(ray, pk2(tb.hi.x, tb.hi.y), pk2(tb.hi.z, tb.hi.w), pk2(ta.hi.w, tb.lo.w), tl);
        const float mplus = __fmaf_rn(best, 1.0f + kFastRel, kFastAbs);
        if (reachable && second > mplus && tl < mplus) {
            ray.dist = best;
            ray.tri = __float_as_uint(ta.lo.w) & ~kLastFlag;
            store_hit(rays, i, ray);
            atomicAdd(&counts[ray.tri], 1);
            if (checkMode) retry[atomicAdd(&ctl->nRetry, 1u)] = i | kRetryVerifyOnly;
        } else {
            retry[atomicAdd(&ctl->nRetry, 1u)] = i;
            atomicAdd(&stats->fallbackCert, 1ull);
        }
    }
}

// Reference-order traversal of the rays the fast path handed back.  Entries with kRetryVerifyOnly ("fast_check") are
// certified rays whose stored answer is compared with the reference-order answer; nothing is written for them.
template <int STACK>
__global__ void __launch_bounds__(128) k_extend_retry(int* __restrict__ counts, const float4* __restrict__ wtris, float4* __restrict__ rays,
                                                      const float4* __restrict__ pairs, const RefillCtl* __restrict__ ctl,
                                                      const uint32_t* __restrict__ retry, FastStats* __restrict__ stats)
{
    const uint32_t n = ctl->nRetry;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t e = retry[j];
        const uint32_t i = e & ~kRetryVerifyOnly;
        const bool verifyOnly = (e & kRetryVerifyOnly) != 0u;
        RayCtx ray;
        {
            const F8 r = ld256_stream(rays + 2ull * i);
            ray.dx = r.lo.x; ray.dy = r.lo.y; ray.dz = r.lo.z;
            ray.ox = r.lo.w; ray.oy = r.hi.x; ray.oz = r.hi.y;
            ray.dist = r.hi.z;
            ray.tri = __float_as_uint(r.hi.w);
            ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
        }
        const float fd = ray.dist;
        const uint32_t ft = ray.tri;
        if (verifyOnly) ray.dist = kNoHit;     // certified rays started without a hit; a miss leaves triID as it was
        if (ray_is_tame(ray)) {
            make_tame(ray);
            bvh_intersect<DIV_MARKSTEIN1, STACK, -1>(ray, pairs, wtris, 0u);
        } else
            bvh_intersect<DIV_IEEE, STACK, -1>(ray, pairs, wtris, 0u);
        if (verifyOnly) {
            if (__float_as_uint(fd) != __float_as_uint(ray.dist) || ft != ray.tri) atomicAdd(&stats->checkMismatch, 1ull);
        } else {
            store_hit(rays, i, ray);
            if (ray.dist != kNoHit) atomicAdd(&counts[ray.tri], 1);
        }
    }
}

} // namespace uvrt
