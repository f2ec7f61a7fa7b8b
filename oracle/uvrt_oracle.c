/*
 * uvrt_oracle.c -- CPU restatement of the reference's wavefront hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libuvrt.so, CUDA) never calls into this file and has no CPU fallback.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so
 * this port is pinned against the reference's OWN sources compiled here (oracle/_ref,
 * built by oracle/build_ref.sh from /root/reference/cl/[star].cl and bvh.cpp) and against the
 * known-answer vectors of SURVEY.md App. C (tests/test_oracle.py).
 *
 * Arithmetic contract: strict IEEE-754 binary32/binary64,
 * round-to-nearest-even, no FMA contraction (build with -ffp-contract=off), source
 * evaluation order; float->u32 of the seed expression saturates (App. B-2);
 * SEED is a launch argument read by all work-items, SEED_out = work-item 0's final
 * RNG state (App. B-1).
 *
 * All paths below cite /root/reference files.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- device structs: cl/tools.cl:8-14, 31-37, 39-45 ------------------------------- */
typedef struct { float dirx, diry, dirz, origx, origy, origz, dist; uint32_t triID; } orc_ray;      /* 32 B */
typedef struct { float v0x, v0y, v0z, d1, v1x, v1y, v1z, d2, v2x, v2y, v2z, d3, cx, cy, cz, d4; } orc_tri; /* 64 B */
typedef struct { float minx, miny, minz; int32_t leftFirst; float maxx, maxy, maxz; int32_t triCount; } orc_node; /* 32 B */

typedef struct {
    uint64_t rays, innerVisits, leafVisits, triTests, hits;
    uint32_t maxStack;
} orc_counters;

/* ---- RNG: cl/tools.cl:2-4 ---------------------------------------------------------- */
uint32_t orc_wang_hash(uint32_t s)
{
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s = s ^ (s >> 4);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15);
    return s;
}

static inline uint32_t random_int(uint32_t* s)
{
    *s ^= *s << 13;
    *s ^= *s >> 17;
    *s ^= *s << 5;
    return *s;
}

static inline float random_float(uint32_t* s)
{
    /* u32 -> f32 (RN) then one fp32 multiply; can return exactly 1.0f */
    return (float)random_int(s) * 2.3283064365387e-10f;
}

/* float -> u32 with saturation (what cvt.rzi.u32.f32 does) */
static inline uint32_t sat_u32(float f)
{
    if (!(f > 0.0f)) return 0u;          /* negative, zero, NaN */
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

/* ---- generate: cl/generate.cl:8-40 ------------------------------------------------- */
static inline uint32_t generate_one(orc_ray* out, int32_t tid, float lx, float ly, float lz,
                                    float lightLength, uint32_t seedIn)
{
    /* generate.cl:13 -- int term first, then fp32 adds left to right */
    float e = (float)(tid * 17 + 1);
    e = e + lx * 13.0f;
    e = e + ly * 7.0f;
    e = e + lz * 11.0f;
    e = e + (float)(seedIn >> 15);
    uint32_t seed = orc_wang_hash(sat_u32(e));

    orc_ray r;
    r.origx = lx;
    r.origy = ly + random_float(&seed) * lightLength;            /* generate.cl:16 */
    r.origz = lz;
    float diry = random_float(&seed) * 2.0f - 1.0f;              /* generate.cl:22 */
    double len = sqrt(1.0 - (double)diry * (double)diry);       /* generate.cl:23 */
    double x, z;
    /* generate.cl:25-28 -- operands of the vector literal are drawn left to right */
    do {
        float fx = random_float(&seed) * 2.0f - 1.0f;
        float fz = random_float(&seed) * 2.0f - 1.0f;
        x = (double)fx;
        z = (double)fz;
        /* WangHash(61) == 0 and xorshift32 never leaves 0: the reference's loop then never ends (a GPU
         * hang).  Decision: such a work-item keeps its first draw. */
    } while (x * x + z * z > 1.0 && seed != 0u);
    double scale = len / sqrt(x * x + z * z);                   /* generate.cl:29 */
    r.dirx = (float)(x * scale);
    r.diry = diry;
    r.dirz = (float)(z * scale);
    r.dist = 1e30f;
    r.triID = 0;
    if (out) *out = r;
    return seed;
}

/* Generates rays [firstRay, firstRay+nRays) of a launch into rays[0..nRays).
 * seedOut (optional) receives work-item 0's final RNG state (generate.cl:39). */
void orc_generate(orc_ray* rays, int64_t firstRay, int64_t nRays, float lx, float ly, float lz,
                  float lightLength, uint32_t seedIn, uint32_t* seedOut)
{
    int64_t i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < nRays; i++)
        generate_one(&rays[i], (int32_t)(firstRay + i), lx, ly, lz, lightLength, seedIn);
    if (seedOut) *seedOut = generate_one(NULL, 0, lx, ly, lz, lightLength, seedIn);
}

/* ---- extend: cl/extend.cl:6-99 ----------------------------------------------------- */
/* OpenCL min/max are select forms */
static inline float cl_min(float x, float y) { return y < x ? y : x; }
static inline float cl_max(float x, float y) { return x < y ? y : x; }

static inline void intersect_tri(orc_ray* ray, const orc_tri* t, uint32_t triID)
{
    /* extend.cl:6-27 (Moeller-Trumbore) */
    float e1x = t->v1x - t->v0x, e1y = t->v1y - t->v0y, e1z = t->v1z - t->v0z;
    float e2x = t->v2x - t->v0x, e2y = t->v2y - t->v0y, e2z = t->v2z - t->v0z;
    /* h = cross(dir, edge2) */
    float hx = ray->diry * e2z - ray->dirz * e2y;
    float hy = ray->dirz * e2x - ray->dirx * e2z;
    float hz = ray->dirx * e2y - ray->diry * e2x;
    float a = (e1x * hx + e1y * hy) + e1z * hz;
    if (fabsf(a) < 0.00001f) return;
    float f = 1.0f / a;
    float sx = ray->origx - t->v0x, sy = ray->origy - t->v0y, sz = ray->origz - t->v0z;
    float u = f * ((sx * hx + sy * hy) + sz * hz);
    if ((u < 0.0f) | (u > 1.0f)) return;
    /* q = cross(s, edge1) */
    float qx = sy * e1z - sz * e1y;
    float qy = sz * e1x - sx * e1z;
    float qz = sx * e1y - sy * e1x;
    float v = f * ((ray->dirx * qx + ray->diry * qy) + ray->dirz * qz);
    if ((v < 0.0f) | (u + v > 1.0f)) return;
    float tt = f * ((e2x * qx + e2y * qy) + e2z * qz);
    if (tt > 0.0001f && tt < ray->dist) {
        ray->dist = tt;
        ray->triID = triID;
    }
}

static inline float intersect_aabb(const orc_ray* ray, const orc_node* n)
{
    /* extend.cl:29-38 -- six true divisions */
    float tx1 = (n->minx - ray->origx) / ray->dirx, tx2 = (n->maxx - ray->origx) / ray->dirx;
    float tmin = cl_min(tx1, tx2), tmax = cl_max(tx1, tx2);
    float ty1 = (n->miny - ray->origy) / ray->diry, ty2 = (n->maxy - ray->origy) / ray->diry;
    tmin = cl_max(tmin, cl_min(ty1, ty2));
    tmax = cl_min(tmax, cl_max(ty1, ty2));
    float tz1 = (n->minz - ray->origz) / ray->dirz, tz2 = (n->maxz - ray->origz) / ray->dirz;
    tmin = cl_max(tmin, cl_min(tz1, tz2));
    tmax = cl_min(tmax, cl_max(tz1, tz2));
    if (tmax >= tmin && tmin < ray->dist && tmax > 0.0f) return tmin;
    return 1e30f;
}

#define ORC_STACK 128 /* reference uses 32 unchecked (extend.cl:43); deeper trees need more */

static void bvh_intersect(orc_ray* ray, const orc_tri* tri, const orc_node* nodes,
                          const uint32_t* triIdx, orc_counters* c)
{
    /* extend.cl:40-81 */
    const orc_node* node = &nodes[0];
    const orc_node* stack[ORC_STACK];
    uint32_t sp = 0;
    for (;;) {
        if (node->triCount > 0) {
            c->leafVisits++;
            for (uint32_t i = 0; i < (uint32_t)node->triCount; i++) {
                uint32_t id = triIdx[node->leftFirst + i];
                c->triTests++;
                intersect_tri(ray, &tri[id], id);
            }
            if (sp == 0) break;
            node = stack[--sp];
            continue;
        }
        c->innerVisits++;
        const orc_node* c1 = &nodes[node->leftFirst];
        const orc_node* c2 = &nodes[node->leftFirst + 1];
        float d1 = intersect_aabb(ray, c1);
        float d2 = intersect_aabb(ray, c2);
        if (d1 > d2) {
            float d = d1; d1 = d2; d2 = d;
            const orc_node* t = c1; c1 = c2; c2 = t;
        }
        if (d1 == 1e30f) {
            if (sp == 0) break;
            node = stack[--sp];
        } else {
            node = c1;
            if (d2 != 1e30f) {
                stack[sp++] = c2;
                if (sp > c->maxStack) c->maxStack = sp;
            }
        }
    }
}

/* extend.cl:85-99.  counters may be NULL.  nThreads<=0 -> all cores. */
void orc_extend(int32_t* tempPhotonMap, const orc_tri* tris, orc_ray* rays, const orc_node* nodes,
                const uint32_t* triIdx, int64_t nRays, int nThreads, orc_counters* counters)
{
    orc_counters tot;
    memset(&tot, 0, sizeof tot);
#ifdef _OPENMP
    if (nThreads <= 0) nThreads = omp_get_max_threads();
#else
    nThreads = 1;
#endif
#pragma omp parallel num_threads(nThreads)
    {
        orc_counters c;
        memset(&c, 0, sizeof c);
        int64_t i;
#pragma omp for schedule(dynamic, 4096)
        for (i = 0; i < nRays; i++) {
            orc_ray r = rays[i];
            bvh_intersect(&r, tris, nodes, triIdx, &c);
            rays[i].dist = r.dist;
            rays[i].triID = r.triID;
            if (r.dist != 1e30f) {
                c.hits++;
                __atomic_fetch_add(&tempPhotonMap[r.triID], 1, __ATOMIC_RELAXED);
            }
        }
#pragma omp critical
        {
            tot.innerVisits += c.innerVisits;
            tot.leafVisits += c.leafVisits;
            tot.triTests += c.triTests;
            tot.hits += c.hits;
            if (c.maxStack > tot.maxStack) tot.maxStack = c.maxStack;
        }
    }
    tot.rays = (uint64_t)nRays;
    if (counters) *counters = tot;
}

/* Brute-force closest hit over all triangles in index order (no BVH); used by tests
 * to bound slab-test false negatives. */
void orc_brute_force(const orc_tri* tris, int nTris, orc_ray* rays, int64_t nRays)
{
    int64_t i;
#pragma omp parallel for schedule(dynamic, 64)
    for (i = 0; i < nRays; i++) {
        orc_ray r = rays[i];
        for (int t = 0; t < nTris; t++) intersect_tri(&r, &tris[t], (uint32_t)t);
        rays[i] = r;
    }
}

/* ---- accumulate: cl/accumulate.cl:4-14 -------------------------------------------- */
void orc_accumulate(double* photonMap, double* maxPhotonMap, int32_t* temp, float timeStep, int n)
{
    for (int i = 0; i < n; i++) {
        photonMap[i] = photonMap[i] + (double)temp[i] * (double)timeStep;
        double t = (double)temp[i];
        maxPhotonMap[i] = maxPhotonMap[i] < t ? t : maxPhotonMap[i];
        temp[i] = 0;
    }
}

/* ---- shade: cl/shade.cl:23-41 ------------------------------------------------------ */
void orc_compute_dosage(const double* photonMap, float* dosage, const orc_tri* tris,
                        int photonsPerLight, float scaledPower, int n)
{
    for (int i = 0; i < n; i++) {
        const orc_tri* t = &tris[i];
        float ax = t->v0x - t->v1x, ay = t->v0y - t->v1y, az = t->v0z - t->v1z;
        float bx = t->v0x - t->v2x, by = t->v0y - t->v2y, bz = t->v0z - t->v2z;
        float cx = ay * bz - az * by;
        float cy = az * bx - ax * bz;
        float cz = ax * by - ay * bx;
        float area = sqrtf((cx * cx + cy * cy) + cz * cz) / 2.0f;
        /* numerator in fp64, denominator an fp32 product, quotient in fp64, narrowed */
        double num = (double)scaledPower * photonMap[i];
        float den = area * (float)photonsPerLight;
        dosage[i] = (float)(num / (double)den);
    }
}

/* shade.cl:4-21 */
static void heatmap(float v, float* rgb)
{
    const float mid = 0.5f, hi = 0.75f, lo = 0.25f;
    if (v > mid) {
        if (v > hi) { rgb[0] = 1.0f; rgb[1] = (1.0f - v) / (1.0f - hi); rgb[2] = 0.0f; }
        else        { rgb[0] = (v - mid) / (hi - mid); rgb[1] = 1.0f; rgb[2] = 0.0f; }
    } else {
        if (v > lo) { rgb[0] = 0.0f; rgb[1] = 1.0f; rgb[2] = (mid - v) / (mid - lo); }
        else        { rgb[0] = 0.0f; rgb[1] = v / lo; rgb[2] = 1.0f; }
    }
}

/* shade.cl:43-71 -- 9 floats per triangle */
void orc_dosage_to_color(const float* dosage, float* color9, float minValue, int thresholdView, int n)
{
    for (int i = 0; i < n; i++) {
        float maxValue = minValue * 2.0f;
        float norm = dosage[i] / maxValue;
        float rgb[3];
        if (thresholdView && norm < 0.5f) { rgb[0] = 0.0f; rgb[1] = 0.0f; rgb[2] = norm * 2.0f; }
        else heatmap(norm, rgb);
        for (int v = 0; v < 3; v++) {
            color9[i * 9 + v * 3 + 0] = rgb[0];
            color9[i * 9 + v * 3 + 1] = rgb[1];
            color9[i * 9 + v * 3 + 2] = rgb[2];
        }
    }
}

/* ---- reset: cl/reset.cl:4-26 ------------------------------------------------------- */
void orc_reset(double* photonMap, double* maxPhotonMap, int32_t* temp, float* color9, int resetColor, int n)
{
    for (int i = 0; i < n; i++) {
        photonMap[i] = 0.0;
        maxPhotonMap[i] = 0.0;
        temp[i] = 0;
        if (resetColor)
            for (int k = 0; k < 9; k++) color9[i * 9 + k] = 0.0f;
    }
}

/* ---- helpers ----------------------------------------------------------------------- */
uint64_t orc_fnv1a64(const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

/* FNV over (dist bits, triID) per ray, 8 B each */
uint64_t orc_fnv_hits(const orc_ray* rays, int64_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (int64_t i = 0; i < n; i++) {
        const unsigned char* b = (const unsigned char*)&rays[i].dist;
        for (int k = 0; k < 8; k++) { h ^= b[k]; h *= 1099511628211ull; }
    }
    return h;
}

/* bench.py's CPU legs pin the thread count explicitly: launchers such as torchrun export OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
