// uvrt_kernels.cuh -- sm_100a kernels of the wavefront hot path.
//
// Arithmetic contract (DESIGN.md "Parity"): every floating-point operation that the
// reference's kernels perform is executed here as the same IEEE-754 operation, in the same
// order, with round-to-nearest-even and WITHOUT fused multiply-adds.  That is what the
// explicit __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__d*_rn intrinsics below are for: nvcc never
// contracts them.  Where an FMA appears (the shared-reciprocal division) it computes a value
// that is proven equal to the IEEE quotient.
//
// Reference being restated: /root/reference/cl/{tools,generate,extend,accumulate,shade,reset}.cl
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uvrt {

constexpr uint32_t kLeafFlag = 0x80000000u;   // child reference: leaf (low bits = first triangle slot)
constexpr uint32_t kLastFlag = 0x80000000u;   // triangle slot: last triangle of its leaf
constexpr float kNoHit = 1e30f;

// ---- strict helpers -------------------------------------------------------------------------
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
// OpenCL C select forms (tools: SURVEY App. A)
__device__ __forceinline__ float clmin(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float clmax(float x, float y) { return x < y ? y : x; }

// 32-byte read-only load (sm_100: LDG.E.256).  One instruction fetches one whole 32-byte sector
// per lane, where two LDG.128 would each occupy the L1 data pipe for the same sector.
struct F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p)
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}
// The same 32 bytes as four 64-bit words: register pairs ready for the packed fp32x2 pipe.
typedef unsigned long long u64;
struct W4 { u64 w0, w1, w2, w3; };
__device__ __forceinline__ W4 ldg256w(const float4* p)
{
    W4 r;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.w0), "=l"(r.w1), "=l"(r.w2), "=l"(r.w3) : "l"(p));
    return r;
}
#ifdef UVRT_EXPERIMENTS
// cache-hint experiment ("fetch_mode" 4): nodes marked evict_last in L1
__device__ __forceinline__ W4 ldg256w_keep(const float4* p)
{
    W4 r;
    asm("ld.global.nc.L1::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.w0), "=l"(r.w1), "=l"(r.w2), "=l"(r.w3) : "l"(p));
    return r;
}
#endif
// streaming data (rays, permutation, results) is kept out of L1 ("fetch_mode" 3, the default)
__device__ __forceinline__ F8 ld256_stream(const float4* p)
{
    F8 r;
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32_stream(const uint32_t* p)
{
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// Packed binary32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each half is an ordinary IEEE round-to-nearest
// operation, so results are bit-identical to the scalar forms at half the issue slots.
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// plain (coherent) 32-byte load for data written by an earlier kernel of the same stream
__device__ __forceinline__ F8 ld256(const float4* p)
{
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}

// ---- RNG: tools.cl:2-4 ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t wang_hash(uint32_t s)
{
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s = s ^ (s >> 4);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15);
    return s;
}
__device__ __forceinline__ uint32_t random_int(uint32_t& s)
{
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}
__device__ __forceinline__ float random_float(uint32_t& s)
{
    return fm(__uint2float_rn(random_int(s)), 2.3283064365387e-10f);
}

struct RayRec {          // tools.cl:8-14, as two 16-byte halves
    float4 a;            // dir.x dir.y dir.z orig.x
    float4 b;            // orig.y orig.z dist triID(bits)
};

// generate.cl:8-40 for one work-item; returns the work-item's final RNG state
__device__ __forceinline__ uint32_t generate_ray(int gid, float lx, float ly, float lz, float lightLength,
                                                 uint32_t seedIn, RayRec& out)
{
    // generate.cl:13: the sum is evaluated in fp32 after the first (int) term; the final
    // float -> uint conversion saturates (cvt.rzi.u32.f32), SURVEY App. B-2
    float e = __int2float_rn((int)((uint32_t)gid * 17u + 1u));
    e = fa(e, fm(lx, 13.0f));
    e = fa(e, fm(ly, 7.0f));
    e = fa(e, fm(lz, 11.0f));
    e = fa(e, __uint2float_rn(seedIn >> 15));
    uint32_t seed = wang_hash(__float2uint_rz(e));

    float oy = fa(ly, fm(random_float(seed), lightLength));                 // generate.cl:16
    float diry = fs(fm(random_float(seed), 2.0f), 1.0f);                    // generate.cl:22
    double dy = (double)diry;
    double len = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(dy, dy)));             // generate.cl:23
    double x, z, d2;
    do {                                                                    // generate.cl:25-28
        float fx = fs(fm(random_float(seed), 2.0f), 1.0f);                  // x is drawn first
        float fz = fs(fm(random_float(seed), 2.0f), 1.0f);
        x = (double)fx;
        z = (double)fz;
        d2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(z, z));
        // WangHash(61) = 0, and xorshift32 never leaves 0: every draw is 0, (x, z) = (-1, -1) for ever and
        // the reference's loop never ends (its GPU hangs until the watchdog fires: SURVEY App. B, DESIGN.md
        // section 6).  Such a work-item keeps the values of its first pass: a ray straight down.
    } while (d2 > 1.0 && seed != 0u);
    double scale = __ddiv_rn(len, __dsqrt_rn(d2));                          // generate.cl:29
    out.a = make_float4(__double2float_rn(__dmul_rn(x, scale)), diry,
                        __double2float_rn(__dmul_rn(z, scale)), lx);
    out.b = make_float4(oy, lz, kNoHit, __uint_as_float(0u));
    return seed;
}

// ---- ray binning (queue re-ordering between generate and extend) --------------------------------
// Rays of one launch leave the lamp in random directions; a warp of 32 consecutive rays walks 32
// unrelated paths through the tree.  Before extend, rays are therefore ordered by a coarse key --
// (dir.y cell, origin slice along the lamp, azimuth cell), azimuth varying fastest; the cells are
// equal-probability for the lamp's uniform emission -- with a counting sort:
//   count   one atomicAdd per ray on a 2^16-entry table (almost never contended) hands the ray its
//           rank inside its bin; fused into generate while the ray is still in registers
//   scan    64 blocks scan 1,024 counters each (and clear them for the next launch)
//   scatter adds the prefix over the 64 block totals and writes the permutation
// Extend then visits rays through the permutation and writes results back to the ray's own slot,
// so the ray buffer keeps the reference's order and every per-ray result is unchanged.
struct BinDims { int nY, nT, nP; float y0, invLen; };
constexpr int kBinsPerScanBlock = 1024;

__device__ __forceinline__ uint32_t bin_key(const BinDims& d, float dx, float dy, float dz, float oy)
{
    int t = min(d.nT - 1, max(0, (int)((dy + 1.0f) * 0.5f * (float)d.nT)));
    int ph = min(d.nP - 1, max(0, (int)((atan2f(dz, dx) + 3.14159265f) * 0.159154943f * (float)d.nP)));
    int y = min(d.nY - 1, max(0, (int)((oy - d.y0) * d.invLen * (float)d.nY)));
    return (uint32_t)((t * d.nY + y) * d.nP + ph);
}

// for rays that were not produced by k_generate (uvrt_write)
__global__ void __launch_bounds__(256) k_bin_count(const float4* __restrict__ rays, uint32_t nRays, BinDims d,
                                                   unsigned int* __restrict__ binCount, uint2* __restrict__ keyRank)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRays) return;
    F8 r = ld256(rays + 2ull * i);
    uint32_t key = bin_key(d, r.lo.x, r.lo.y, r.lo.z, r.hi.x);
    uint32_t rank = atomicAdd(&binCount[key], 1u);
    keyRank[i] = make_uint2(key, rank);
}

// One block scans kBinsPerScanBlock counters: binStart = exclusive scan inside the block's range,
// blockTotal[blockIdx] = the range's sum; counters are cleared for the next launch.
__global__ void __launch_bounds__(256) k_bin_scan(uint4* __restrict__ binCount, uint4* __restrict__ binStart,
                                                  unsigned int* __restrict__ blockTotal)
{
    __shared__ unsigned int warpSums[8];
    const int v4 = blockIdx.x * (kBinsPerScanBlock / 4) + threadIdx.x;   // one uint4 per thread
    uint4 c = binCount[v4];
    unsigned int local = c.x + c.y + c.z + c.w;
    unsigned int v = local;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += n;
    }
    if (lane == 31) warpSums[w] = v;
    __syncthreads();
    unsigned int before = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) before += (k < (int)w) ? warpSums[k] : 0u;
    unsigned int run = before + v - local;
    uint4 o;
    o.x = run; run += c.x;
    o.y = run; run += c.y;
    o.z = run; run += c.z;
    o.w = run; run += c.w;
    binStart[v4] = o;
    binCount[v4] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 255) blockTotal[blockIdx.x] = run;
}

__global__ void __launch_bounds__(256) k_bin_scatter(const uint2* __restrict__ keyRank, const unsigned int* __restrict__ binStart,
                                                     const unsigned int* __restrict__ blockTotal, int nScanBlocks,
                                                     uint32_t nRays, uint32_t* __restrict__ perm)
{
    __shared__ unsigned int blockBase[64];
    __shared__ unsigned int first32Total;
    if (threadIdx.x < 64) {
        // exclusive prefix over the (at most 64) block totals, by two warps
        unsigned int t = (int)threadIdx.x < nScanBlocks ? blockTotal[threadIdx.x] : 0u;
        unsigned int v = t;
        const unsigned lane = threadIdx.x & 31u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (unsigned)o) v += n;
        }
        blockBase[threadIdx.x] = v - t;
        if (threadIdx.x == 31) first32Total = v;
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) blockBase[threadIdx.x] += first32Total;
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRays) return;
    uint2 kr = keyRank[i];
    perm[binStart[kr.x] + blockBase[kr.x / kBinsPerScanBlock] + kr.y] = i;
}

// BIN = 1 also takes the ray's slot in its bin (the count step of the counting sort)
template <int BIN>
__global__ void __launch_bounds__(256) k_generate(float4* __restrict__ rays, long long firstRay, long long nRays,
                                                  float lx, float ly, float lz, float lightLength, uint32_t seedIn,
                                                  BinDims d, unsigned int* __restrict__ binCount, uint2* __restrict__ keyRank)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRays) return;
    RayRec r;
    generate_ray((int)(firstRay + i), lx, ly, lz, lightLength, seedIn, r);
    rays[2 * i] = r.a;
    rays[2 * i + 1] = r.b;
    if (BIN) {
        uint32_t key = bin_key(d, r.a.x, r.a.y, r.a.z, r.b.x);
        uint32_t rank = atomicAdd(&binCount[key], 1u);
        keyRank[i] = make_uint2(key, rank);
    }
}

// SEED chain (generate.cl:39): one thread replays work-item 0 of consecutive launches
__global__ void k_seed_chain(const float* __restrict__ lightPos3, int nLaunches, float lightLength,
                             uint32_t seedIn, uint32_t* __restrict__ seedsOut)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t seed = seedIn;
    seedsOut[0] = seed;
    for (int i = 0; i < nLaunches; i++) {
        RayRec r;
        seed = generate_ray(0, lightPos3[3 * i], lightPos3[3 * i + 1], lightPos3[3 * i + 2], lightLength, seed, r);
        seedsOut[i + 1] = seed;
    }
}

// ---- extend ---------------------------------------------------------------------------------
// Division modes for the slab test (extend.cl:31-36 divides six times per box).
//   DIV_IEEE      : __fdiv_rn, the literal restatement.
//   DIV_MARKSTEIN2: r = RN(1/d) once per ray and axis, then q0 = n*r, two FMA residual
//                   corrections; the result equals RN(n/d) (Markstein 1990: with a correctly
//                   rounded reciprocal and a faithful q, one correction step rounds correctly;
//                   the first step makes q faithful).  Only used for rays whose direction and
//                   origin components keep every intermediate in the normal range, everything
//                   else takes DIV_IEEE (see ray_is_tame()).
//   DIV_MARKSTEIN1: one correction step.  Also equal to RN(n/d): the error of the value before
//                   the last rounding can only cross a rounding boundary for the finitely many
//                   significand pairs whose quotient lies within a few 2^-48 of a midpoint, and
//                   tools/prove_division.c enumerates and checks all of those.
enum DivMode { DIV_IEEE = 0, DIV_MARKSTEIN2 = 2, DIV_MARKSTEIN1 = 1 };

struct RayCtx {
    float ox, oy, oz, dx, dy, dz;
    // shared-reciprocal modes: (-o.x,-o.y) (-o.z,-o.z), RN(1/d) as (x,y) (z,z), (-d.x,-d.y) (-d.z,-d.z)
    u64 noXY, noZZ, rXY, rZZ, ndXY, ndZZ;
    float dist;
    uint32_t tri;
};

template <int DIV>
__device__ __forceinline__ float slab_q(float n, float d, float r)
{
    if (DIV == DIV_IEEE) {
        return __fdiv_rn(n, d);
    } else {
        float q = fm(n, r);
        float rem = __fmaf_rn(-d, q, n);
        q = __fmaf_rn(rem, r, q);
        if (DIV == DIV_MARKSTEIN2) {
            rem = __fmaf_rn(-d, q, n);
            q = __fmaf_rn(rem, r, q);
        }
        return q;
    }
}

// the same quotient for two lanes at once: q = (a - o) / d with no = -o, nd = -d, r = RN(1/d)
template <int DIV>
__device__ __forceinline__ u64 slab_q2(u64 a, u64 no, u64 nd, u64 r)
{
    u64 n = add2(a, no);              // a + (-o) == a - o
    u64 q = mul2(n, r);
    u64 rem = fma2(nd, q, n);
    q = fma2(rem, r, q);
    if (DIV == DIV_MARKSTEIN2) {
        rem = fma2(nd, q, n);
        q = fma2(rem, r, q);
    }
    return q;
}

// extend.cl:29-38.  A child record is four 64-bit words:
//   w0 = (min.x, min.y)   w1 = (max.x, max.y)   w2 = (min.z, max.z)   w3 = (reference, 0)
// OCT >= 0 (shared-reciprocal modes only): the ray's direction signs are known at compile time
// (bit 0/1/2 set = dir.x/y/z negative).  Division by a positive d is monotone, so with
// box.min <= box.max (checked at upload) min(t1, t2) is t1 for d > 0 and t2 for d < 0: the six
// per-axis min/max of extend.cl:32-36 disappear and tmin/tmax are one three-input max/min each.
template <int DIV, int OCT>
__device__ __forceinline__ bool intersect_aabb(const RayCtx& ray, const W4& c, float& tminOut)
{
    float tx1, ty1, tx2, ty2, tz1, tz2;
    if (DIV == DIV_IEEE) {
        float mnx, mny, mxx, mxy, mnz, mxz;
        upk2(c.w0, mnx, mny); upk2(c.w1, mxx, mxy); upk2(c.w2, mnz, mxz);
        tx1 = __fdiv_rn(fs(mnx, ray.ox), ray.dx); tx2 = __fdiv_rn(fs(mxx, ray.ox), ray.dx);
        ty1 = __fdiv_rn(fs(mny, ray.oy), ray.dy); ty2 = __fdiv_rn(fs(mxy, ray.oy), ray.dy);
        tz1 = __fdiv_rn(fs(mnz, ray.oz), ray.dz); tz2 = __fdiv_rn(fs(mxz, ray.oz), ray.dz);
        // NaNs are possible here (0/0): keep OpenCL's select forms exactly
        float tmin = clmin(tx1, tx2), tmax = clmax(tx1, tx2);
        tmin = clmax(tmin, clmin(ty1, ty2)); tmax = clmin(tmax, clmax(ty1, ty2));
        tmin = clmax(tmin, clmin(tz1, tz2)); tmax = clmin(tmax, clmax(tz1, tz2));
        tminOut = tmin;
        return tmax >= tmin && tmin < ray.dist && tmax > 0.0f;
    } else {
        upk2(slab_q2<DIV>(c.w0, ray.noXY, ray.ndXY, ray.rXY), tx1, ty1);
        upk2(slab_q2<DIV>(c.w1, ray.noXY, ray.ndXY, ray.rXY), tx2, ty2);
        upk2(slab_q2<DIV>(c.w2, ray.noZZ, ray.ndZZ, ray.rZZ), tz1, tz2);
        // tame rays produce no NaN; fminf/fmaxf then agree with the select forms up to the
        // sign of a zero, which no comparison below can observe
        float tmin, tmax;
        if (OCT >= 0) {
            const float nx = (OCT & 1) ? tx2 : tx1, fx = (OCT & 1) ? tx1 : tx2;
            const float ny = (OCT & 2) ? ty2 : ty1, fy = (OCT & 2) ? ty1 : ty2;
            const float nz = (OCT & 4) ? tz2 : tz1, fz = (OCT & 4) ? tz1 : tz2;
            tmin = fmaxf(fmaxf(nx, ny), nz);
            tmax = fminf(fminf(fx, fy), fz);
        } else {
            tmin = fminf(tx1, tx2); tmax = fmaxf(tx1, tx2);
            tmin = fmaxf(tmin, fminf(ty1, ty2)); tmax = fminf(tmax, fmaxf(ty1, ty2));
            tmin = fmaxf(tmin, fminf(tz1, tz2)); tmax = fminf(tmax, fmaxf(tz1, tz2));
        }
        tminOut = tmin;
        return tmax >= tmin && tmin < ray.dist && tmax > 0.0f;
    }
}

// extend.cl:60-78 without materialising the 1e30f "miss" distances: with d = hit ? tmin : 1e30f
// (tmin < dist <= 1e30f whenever hit), "d1 > d2" is h2 && (!h1 || tmin1 > tmin2).
// Returns false when both children are missed; otherwise `first` is the child to enter and
// `second` the one to push when pushSecond.
__device__ __forceinline__ bool order_children(bool h1, bool h2, float t1, float t2, uint32_t c1, uint32_t c2,
                                               uint32_t& first, uint32_t& second, bool& pushSecond)
{
    const bool swap = h2 && (!h1 || t1 > t2);
    first = swap ? c2 : c1;
    second = swap ? c1 : c2;
    pushSecond = h1 && h2;
    return h1 || h2;
}

__device__ __forceinline__ uint32_t child_ref(const W4& c) { return (uint32_t)c.w3; }

__device__ __forceinline__ void make_tame(RayCtx& ray)
{
    float rx = __frcp_rn(ray.dx), ry = __frcp_rn(ray.dy), rz = __frcp_rn(ray.dz);
    ray.noXY = pk2(-ray.ox, -ray.oy); ray.noZZ = pk2(-ray.oz, -ray.oz);
    ray.rXY = pk2(rx, ry);            ray.rZZ = pk2(rz, rz);
    ray.ndXY = pk2(-ray.dx, -ray.dy); ray.ndZZ = pk2(-ray.dz, -ray.dz);
}

// extend.cl:6-27 with edge1 = v1-v0 and edge2 = v2-v0 formed at upload time (the same IEEE
// subtractions, done once instead of once per test).  t0.w = triID | last-in-leaf flag.
__device__ __forceinline__ void intersect_tri(RayCtx& ray, const float4& t0, const float4& e1, const float4& e2)
{
    float hx = fs(fm(ray.dy, e2.z), fm(ray.dz, e2.y));
    float hy = fs(fm(ray.dz, e2.x), fm(ray.dx, e2.z));
    float hz = fs(fm(ray.dx, e2.y), fm(ray.dy, e2.x));
    float a = fa(fa(fm(e1.x, hx), fm(e1.y, hy)), fm(e1.z, hz));
    if (fabsf(a) < 0.00001f) return;
    float f = __frcp_rn(a);           // RN(1/a): the same value as the division 1 / a of extend.cl:17
    float sx = fs(ray.ox, t0.x), sy = fs(ray.oy, t0.y), sz = fs(ray.oz, t0.z);
    float u = fm(f, fa(fa(fm(sx, hx), fm(sy, hy)), fm(sz, hz)));
    if ((u < 0.0f) | (u > 1.0f)) return;
    float qx = fs(fm(sy, e1.z), fm(sz, e1.y));
    float qy = fs(fm(sz, e1.x), fm(sx, e1.z));
    float qz = fs(fm(sx, e1.y), fm(sy, e1.x));
    float v = fm(f, fa(fa(fm(ray.dx, qx), fm(ray.dy, qy)), fm(ray.dz, qz)));
    if ((v < 0.0f) | (fa(u, v) > 1.0f)) return;
    float t = fm(f, fa(fa(fm(e2.x, qx), fm(e2.y, qy)), fm(e2.z, qz)));
    if (t > 0.0001f && t < ray.dist) {
        ray.dist = t;
        ray.tri = __float_as_uint(t0.w) & ~kLastFlag;
    }
}

// A ray is "tame" when every intermediate of the shared-reciprocal quotient stays exactly
// representable, for every box of a scene whose coordinates are 0 or in [2^-77, 2^20] in magnitude
// (checked at upload): direction components in [2^-30, 2], origin components 0 or in [2^-77, 2^20].
// Then n = a - o is 0 or >= 2^-100 (a difference of two floats is a multiple of the smaller ulp),
// q0 = n*r >= 2^-101 is normal, the residual n - d*q0 is a multiple of 2^(e_n - 47) >= 2^-149 and so
// exact even when denormal (no flush-to-zero in this build), and |q| <= 2^51.
__device__ __forceinline__ bool ray_is_tame(const RayCtx& r)
{
    const float dlo = 9.31322574615478515625e-10f;  // 2^-30
    const float olo = 6.6174449e-24f;               // 2^-77
    float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
    bool d_ok = ax >= dlo && ay >= dlo && az >= dlo && ax <= 2.0f && ay <= 2.0f && az <= 2.0f;
    float px = fabsf(r.ox), py = fabsf(r.oy), pz = fabsf(r.oz);
    bool o_ok = (px == 0.0f || (px >= olo && px <= 1048576.0f)) && (py == 0.0f || (py >= olo && py <= 1048576.0f)) &&
                (pz == 0.0f || (pz >= olo && pz <= 1048576.0f));
    return d_ok && o_ok;
}

// extend.cl:40-81.  The traversal order, the tie rules (dist1 > dist2 swaps, so ties keep child
// 1 first; strict t < dist keeps the first-found triangle) and the distance culling are the
// reference's.  `pairs` holds, per inner node, its two children's boxes and references
// (2 x 32 bytes, see intersect_aabb); `wtris` holds leaf triangles in leaf order (4 x float4:
// v0+tag, edge1, edge2, pad -- two 32-byte sectors).
// Node fetch path: two LDG.256 through the LSU.  (UVRT_EXPERIMENTS builds also know "fetch_mode" 1 = four
// 16-byte texture fetches, 2 = second child through the texture unit, 4 = evict_last: all measured slower.)
#ifdef UVRT_EXPERIMENTS
__device__ __forceinline__ W4 tex_w4(cudaTextureObject_t tex, uint32_t texel)
{
    float4 a = tex1Dfetch<float4>(tex, (int)texel), b = tex1Dfetch<float4>(tex, (int)texel + 1);
    W4 r;
    r.w0 = pk2(a.x, a.y); r.w1 = pk2(a.z, a.w); r.w2 = pk2(b.x, b.y); r.w3 = pk2(b.z, b.w);
    return r;
}
#endif

template <int DIV, int STACK, int OCT, int FETCH = 0>
__device__ __forceinline__ void bvh_intersect(RayCtx& ray, const float4* __restrict__ pairs,
                                              const float4* __restrict__ wtris, uint32_t rootRef,
                                              cudaTextureObject_t tex = 0)
{
    uint32_t stack[STACK];
    int sp = 0;
    uint32_t cur = rootRef;
    for (;;) {
        if (cur & kLeafFlag) {
            uint32_t slot = cur & ~kLeafFlag;
            uint32_t w;
            do {
                const float4* t = wtris + 4ull * slot;
                F8 ta = ldg256(t), tb = ldg256(t + 2);
                w = __float_as_uint(ta.lo.w);
                intersect_tri(ray, ta.lo, ta.hi, tb.lo);
                slot++;
            } while (!(w & kLastFlag));
            if (sp == 0) break;
            cur = stack[--sp];
            // back to the loop head: leaf lanes and inner-node lanes run as two independent paths
            // until then, which lets the hardware overlap the triangle loads of the former with
            // the box tests of the latter (falling through into the inner step measured 13 % slower)
            continue;
        }
        const float4* p = pairs + 4ull * cur;
        W4 ca, cb;
#ifdef UVRT_EXPERIMENTS
        if (FETCH == 1) { ca = tex_w4(tex, cur * 4u); cb = tex_w4(tex, cur * 4u + 2u); }
        else if (FETCH == 2) { ca = ldg256w(p); cb = tex_w4(tex, cur * 4u + 2u); }
        else if (FETCH == 4) { ca = ldg256w_keep(p); cb = ldg256w_keep(p + 2); }
        else
#endif
        { ca = ldg256w(p); cb = ldg256w(p + 2); }
        float t1, t2;
        const bool h1 = intersect_aabb<DIV, OCT>(ray, ca, t1);
        const bool h2 = intersect_aabb<DIV, OCT>(ray, cb, t2);
        uint32_t first, second;
        bool pushSecond;
        if (order_children(h1, h2, t1, t2, child_ref(ca), child_ref(cb), first, second, pushSecond)) {
            cur = first;
            if (pushSecond) stack[sp++] = second;
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
}

template <int DIV, int STACK, int FETCH = 0>
__device__ __forceinline__ void trace_one(RayCtx& ray, const float4* __restrict__ pairs,
                                          const float4* __restrict__ wtris, uint32_t rootRef, bool sceneTame, bool binned,
                                          cudaTextureObject_t tex = 0)
{
    if (DIV == DIV_IEEE) {
        bvh_intersect<DIV_IEEE, STACK, -1>(ray, pairs, wtris, rootRef);
    } else {
        if (sceneTame && ray_is_tame(ray)) {
            make_tame(ray);
            // one specialised traversal loop per direction octant; binned rays make almost every
            // warp octant-pure, so the switch rarely diverges.  Unsorted rays would serialise all
            // eight loops in every warp: they take the generic loop (-1).
            const int oct = !binned ? -1 : (ray.dx < 0.0f ? 1 : 0) | (ray.dy < 0.0f ? 2 : 0) | (ray.dz < 0.0f ? 4 : 0);
            switch (oct) {
            case -1: bvh_intersect<DIV, STACK, -1, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 0: bvh_intersect<DIV, STACK, 0, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 1: bvh_intersect<DIV, STACK, 1, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 2: bvh_intersect<DIV, STACK, 2, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 3: bvh_intersect<DIV, STACK, 3, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 4: bvh_intersect<DIV, STACK, 4, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 5: bvh_intersect<DIV, STACK, 5, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            case 6: bvh_intersect<DIV, STACK, 6, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            default: bvh_intersect<DIV, STACK, 7, FETCH>(ray, pairs, wtris, rootRef, tex); break;
            }
        } else {
            bvh_intersect<DIV_IEEE, STACK, -1>(ray, pairs, wtris, rootRef);
        }
    }
}

__device__ __forceinline__ void load_ray(const float4* __restrict__ rays, long long i, RayCtx& ray)
{
    F8 r = ld256(rays + 2 * i);
    float4 a = r.lo, b = r.hi;
    ray.dx = a.x; ray.dy = a.y; ray.dz = a.z;
    ray.ox = a.w; ray.oy = b.x; ray.oz = b.y;
    ray.dist = b.z;
    ray.tri = __float_as_uint(b.w);
    ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
}

__device__ __forceinline__ void store_hit(float4* __restrict__ rays, long long i, const RayCtx& ray)
{
    // dist, triID occupy bytes 24..31 of the 32-byte ray record
    float2* p = reinterpret_cast<float2*>(rays + 2 * i + 1) + 1;
    *p = make_float2(ray.dist, __uint_as_float(ray.tri));
}

// Variant A: one thread per ray, the literal control flow of extend.cl:85-99.
template <int DIV, int STACK, int THREADS, int MINBLOCKS, int FETCH = 0>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) k_extend_simple(int* __restrict__ counts, const float4* __restrict__ wtris,
                                                       float4* __restrict__ rays, const float4* __restrict__ pairs,
                                                       uint32_t rootRef, long long nRays, int sceneTame,
                                                       const uint32_t* __restrict__ perm, cudaTextureObject_t tex = 0,
                                                       int genericOctant = 0, unsigned int* __restrict__ smCursor = nullptr)
{
    long long i;
#ifdef UVRT_EXPERIMENTS
    if (FETCH == 5) {
        // SM-affine chunks ("fetch_mode 5"): the binned ray order is cut into one contiguous range per SM
        // and a block takes the next 128-ray chunk of the range of the SM it happens to run on, so the
        // blocks resident on one SM work on neighbouring bins and share node lines in that SM's L1.  A
        // block whose SM's range is used up steals from the other SMs' ranges (every block takes exactly
        // one chunk, and there are as many chunks as blocks).
        __shared__ unsigned int sChunk;
        if (threadIdx.x == 0) {
            unsigned int smid, nsm;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
            const unsigned int C = gridDim.x;
            unsigned int chunk = 0xffffffffu;
            for (unsigned int v = 0; v < nsm && chunk == 0xffffffffu; v++) {
                const unsigned int s = (smid + v) % nsm;
                const unsigned int lo = (unsigned int)(((unsigned long long)s * C) / nsm), hi = (unsigned int)(((unsigned long long)(s + 1) * C) / nsm);
                if (hi == lo) continue;
                if (*reinterpret_cast<volatile unsigned int*>(&smCursor[s]) >= hi - lo) continue;
                const unsigned int c = atomicAdd(&smCursor[s], 1u);
                if (c < hi - lo) chunk = lo + c;
            }
            sChunk = chunk;
        }
        __syncthreads();
        if (sChunk == 0xffffffffu) return;
        i = (long long)sChunk * blockDim.x + threadIdx.x;
    } else
#endif
        i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRays) return;
    if (perm) i = FETCH >= 3 ? ldg_u32_stream(perm + i) : perm[i];
    RayCtx ray;
    if (FETCH >= 3) {
        F8 r = ld256_stream(rays + 2 * i);
        ray.dx = r.lo.x; ray.dy = r.lo.y; ray.dz = r.lo.z;
        ray.ox = r.lo.w; ray.oy = r.hi.x; ray.oz = r.hi.y;
        ray.dist = r.hi.z;
        ray.tri = __float_as_uint(r.hi.w);
        ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
    } else
        load_ray(rays, i, ray);
    trace_one<DIV, STACK, FETCH>(ray, pairs, wtris, rootRef, sceneTame != 0, perm != nullptr && !genericOctant, tex);
    store_hit(rays, i, ray);
    if (ray.dist != kNoHit) atomicAdd(&counts[ray.tri], 1);
}

#ifdef UVRT_EXPERIMENTS
} // namespace uvrt
#include "uvrt_experiments.cuh"   // rejected variants B / C (persistent warps), kept for A/B runs only
namespace uvrt {
#endif

// ---- cost probe (work sharing between GPUs) ------------------------------------------------------
// The first nRays rays of a launch at one lamp position, traversed in reference order with counters: inner-node
// visits and triangle tests per ray say how expensive the position is relative to the others (they differ by up to
// 1.4x on the room).  Every rank runs the same deterministic probe, so all ranks deal the launches alike without
// an exchange (RayTracer::PlanShards).  out[0] += inner visits, out[1] += triangle tests.
__global__ void __launch_bounds__(128) k_probe_cost(const float4* __restrict__ pairs, const float4* __restrict__ wtris, uint32_t rootRef,
                                                    int nRays, float lx, float ly, float lz, float lightLength, uint32_t seedIn,
                                                    unsigned long long* __restrict__ out)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int inner = 0, tests = 0;
    if (gid < nRays) {
        RayRec r;
        generate_ray(gid, lx, ly, lz, lightLength, seedIn, r);
        RayCtx ray;
        ray.dx = r.a.x; ray.dy = r.a.y; ray.dz = r.a.z; ray.ox = r.a.w; ray.oy = r.b.x; ray.oz = r.b.y;
        ray.dist = kNoHit; ray.tri = 0;
        ray.noXY = ray.noZZ = ray.rXY = ray.rZZ = ray.ndXY = ray.ndZZ = 0ull;
        uint32_t stack[64];
        int sp = 0;
        uint32_t cur = rootRef;
        for (;;) {
            if (cur & kLeafFlag) {
                uint32_t slot = cur & ~kLeafFlag, w;
                do {
                    const float4* t = wtris + 4ull * slot;
                    F8 ta = ldg256(t), tb = ldg256(t + 2);
                    w = __float_as_uint(ta.lo.w);
                    intersect_tri(ray, ta.lo, ta.hi, tb.lo);
                    tests++;
                    slot++;
                } while (!(w & kLastFlag));
            } else {
                inner++;
                const float4* p = pairs + 4ull * cur;
                W4 ca = ldg256w(p), cb = ldg256w(p + 2);
                float t1, t2;
                const bool h1 = intersect_aabb<DIV_IEEE, -1>(ray, ca, t1), h2 = intersect_aabb<DIV_IEEE, -1>(ray, cb, t2);
                uint32_t first, second;
                bool pushSecond;
                if (order_children(h1, h2, t1, t2, child_ref(ca), child_ref(cb), first, second, pushSecond)) {
                    cur = first;
                    if (pushSecond) stack[sp++] = second;
                    continue;
                }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inner += __shfl_down_sync(0xffffffffu, inner, o);
        tests += __shfl_down_sync(0xffffffffu, tests, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], (unsigned long long)inner);
        atomicAdd(&out[1], (unsigned long long)tests);
    }
}

// ---- per-triangle passes ------------------------------------------------------------------
// accumulate.cl:4-14
__global__ void __launch_bounds__(256) k_accumulate(double* __restrict__ photonMap, double* __restrict__ maxPhotonMap,
                                                    int* __restrict__ temp, float timeStep, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c = (double)temp[i];
    photonMap[i] = __dadd_rn(photonMap[i], __dmul_rn(c, (double)timeStep));
    double m = maxPhotonMap[i];
    maxPhotonMap[i] = m < c ? c : m;
    temp[i] = 0;
}

// accumulate.cl:4-14 replayed over the rows of a count matrix (one row per launch, in launch order): the same
// sequence of f64 operations per triangle as `rows` consecutive k_accumulate launches
__global__ void __launch_bounds__(256) k_fold_rows(double* __restrict__ photonMap, double* __restrict__ maxPhotonMap,
                                                   const int* __restrict__ matrix, const float* __restrict__ timeStep, int rows, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sum = photonMap[i], m = maxPhotonMap[i];
    for (int r = 0; r < rows; r++) {
        const double c = (double)matrix[(size_t)r * (size_t)n + i];
        sum = __dadd_rn(sum, __dmul_rn(c, (double)timeStep[r]));
        m = m < c ? c : m;
    }
    photonMap[i] = sum;
    maxPhotonMap[i] = m;
}

// shade.cl:23-41; `verts` is the reference-layout triangle array (64 B per triangle)
__global__ void __launch_bounds__(256) k_compute_dosage(const double* __restrict__ photonMap, float* __restrict__ dosage,
                                                        const float4* __restrict__ verts, int photonsPerLight,
                                                        float scaledPower, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v0 = __ldg(verts + 4ull * i), v1 = __ldg(verts + 4ull * i + 1), v2 = __ldg(verts + 4ull * i + 2);
    float ax = fs(v0.x, v1.x), ay = fs(v0.y, v1.y), az = fs(v0.z, v1.z);
    float bx = fs(v0.x, v2.x), by = fs(v0.y, v2.y), bz = fs(v0.z, v2.z);
    float cx = fs(fm(ay, bz), fm(az, by));
    float cy = fs(fm(az, bx), fm(ax, bz));
    float cz = fs(fm(ax, by), fm(ay, bx));
    float area = __fdiv_rn(__fsqrt_rn(fa(fa(fm(cx, cx), fm(cy, cy)), fm(cz, cz))), 2.0f);
    double num = __dmul_rn((double)scaledPower, photonMap[i]);
    float den = fm(area, __int2float_rn(photonsPerLight));
    dosage[i] = __double2float_rn(__ddiv_rn(num, (double)den));
}

// shade.cl:4-21
__device__ __forceinline__ void heatmap(float v, float& r, float& g, float& b)
{
    const float mid = 0.5f, hi = 0.75f, lo = 0.25f;
    if (v > mid) {
        if (v > hi) { r = 1.0f; g = __fdiv_rn(fs(1.0f, v), fs(1.0f, hi)); b = 0.0f; }
        else        { r = __fdiv_rn(fs(v, mid), fs(hi, mid)); g = 1.0f; b = 0.0f; }
    } else {
        if (v > lo) { r = 0.0f; g = 1.0f; b = __fdiv_rn(fs(mid, v), fs(mid, lo)); }
        else        { r = 0.0f; g = __fdiv_rn(v, lo); b = 1.0f; }
    }
}

// shade.cl:43-71: nine floats per triangle (the same colour on its three vertices)
__global__ void __launch_bounds__(256) k_dosage_to_color(const float* __restrict__ dosage, float* __restrict__ color,
                                                         float minValue, int thresholdView, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float maxValue = fm(minValue, 2.0f);
    float norm = __fdiv_rn(dosage[i], maxValue);
    float r, g, b;
    if (thresholdView && norm < 0.5f) { r = 0.0f; g = 0.0f; b = fm(norm, 2.0f); }
    else heatmap(norm, r, g, b);
    float* c = color + 9ull * i;
    c[0] = r; c[1] = g; c[2] = b;
    c[3] = r; c[4] = g; c[5] = b;
    c[6] = r; c[7] = g; c[8] = b;
}

// reset.cl:4-26
__global__ void __launch_bounds__(256) k_reset(double* __restrict__ photonMap, double* __restrict__ maxPhotonMap,
                                               int* __restrict__ temp, float* __restrict__ color, int resetColor, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    photonMap[i] = 0.0;
    maxPhotonMap[i] = 0.0;
    temp[i] = 0;
    if (!resetColor) return;
    float* c = color + 9ull * i;
#pragma unroll
    for (int k = 0; k < 9; k++) c[k] = 0.0f;
}

// ---- on-device self-test of the shared-reciprocal division -----------------------------------
// Draws (n, d) with the magnitudes ray_is_tame() admits (n down to 2^-100: denormal residuals) and counts quotients that differ from
// __fdiv_rn.  out[0] = samples, out[1] = mismatches of one-step, out[2] = mismatches of two-step.
__global__ void __launch_bounds__(256) k_selftest_division(unsigned long long* __restrict__ out, int iters, uint32_t salt)
{
    uint32_t s = wang_hash((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + salt) | 1u;
    unsigned long long bad1 = 0, bad2 = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t md = random_int(s) & 0x7fffffu, mn = random_int(s) & 0x7fffffu, e = random_int(s);
        int ed = -(int)(e % 31u), en = -100 + (int)((e >> 8) % 121u);
        if ((e >> 20) & 1u) md = ((e >> 21) & 1u) ? 0x7fffffu - (md & 0xffu) : (md & 0xffu);
        if ((e >> 22) & 1u) mn = ((e >> 23) & 1u) ? 0x7fffffu - (mn & 0xffu) : (mn & 0xffu);
        float d = __uint_as_float(((uint32_t)(ed + 127) << 23) | md | (((e >> 30) & 1u) << 31));
        float n = __uint_as_float(((uint32_t)(en + 127) << 23) | mn | ((e >> 31) << 31));
        float r = __frcp_rn(d);
        float q = __fdiv_rn(n, d);
        float q1 = slab_q<DIV_MARKSTEIN1>(n, d, r);
        float q2 = slab_q<DIV_MARKSTEIN2>(n, d, r);
        bad1 += (__float_as_uint(q1) != __float_as_uint(q));
        bad2 += (__float_as_uint(q2) != __float_as_uint(q));
    }
    atomicAdd(&out[0], (unsigned long long)iters);
    atomicAdd(&out[1], bad1);
    atomicAdd(&out[2], bad2);
}

} // namespace uvrt
