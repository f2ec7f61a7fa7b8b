/*
 * traversal_lab.c -- CPU laboratory for traversal strategies (design tool, not product, not oracle).
 *
 * Answers, on the host and with the reference's exact arithmetic, the questions a GPU sweep would
 * otherwise have to answer with box minutes:
 *   lab_exact_stats   the reference traversal (extend.cl:40-81) with extra counters: how many inner-node
 *                     visits are "dead" (the node was pushed with an entry distance that is no longer
 *                     below ray.dist when it is popped, so both children are culled), how inverted
 *                     (t <= entry distance of the leaf box) accepted triangle hits can get
 *   lab_fast          the certified fast traversal of csrc/uvrt_fast.cuh restated on the host: 15-bit
 *                     quantised conservative child boxes (or the exact boxes), distance culling with a
 *                     margin, any-order traversal, exact leaf-box verification of accepted hits and the
 *                     near-tie certificate; reports visits, tests, certificate failures and -- against the
 *                     exact traversal -- mismatching rays
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC tools/traversal_lab.c -o /tmp/liblab.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float dirx, diry, dirz, origx, origy, origz, dist; uint32_t triID; } ray_t;
typedef struct { float v0x, v0y, v0z, d1, v1x, v1y, v1z, d2, v2x, v2y, v2z, d3, cx, cy, cz, d4; } tri_t;
typedef struct { float minx, miny, minz; int32_t leftFirst; float maxx, maxy, maxz; int32_t triCount; } node_t;

typedef struct {
    uint64_t rays, innerVisits, leafVisits, triTests, hits;
    uint64_t deadInner, deadLeaf, pops;          /* popped entries whose push-time entry distance >= ray.dist */
    uint64_t accepted, inverted;                 /* accepted (ray, triangle) pairs; with t <= tmin of the leaf box */
    double maxInvAbs, maxInvRel;                 /* worst tmin_leaf - t (absolute, and relative to t) */
    uint64_t certFail, mismatch, rawMismatch, nearTie, boxReject, tminFail;
    uint64_t warpIters, warpLaneIters;           /* 32-ray groups: sum of max-iterations, sum of iterations */
} lab_stats;

static inline float cl_min(float x, float y) { return y < x ? y : x; }
static inline float cl_max(float x, float y) { return x < y ? y : x; }

/* Moeller-Trumbore of extend.cl:6-27 without the final distance comparison: returns 1 and *tt when the
 * triangle is hit at t > 1e-4 */
static inline int tri_accept(const ray_t* ray, const tri_t* t, float* tt)
{
    float e1x = t->v1x - t->v0x, e1y = t->v1y - t->v0y, e1z = t->v1z - t->v0z;
    float e2x = t->v2x - t->v0x, e2y = t->v2y - t->v0y, e2z = t->v2z - t->v0z;
    float hx = ray->diry * e2z - ray->dirz * e2y;
    float hy = ray->dirz * e2x - ray->dirx * e2z;
    float hz = ray->dirx * e2y - ray->diry * e2x;
    float a = (e1x * hx + e1y * hy) + e1z * hz;
    if (fabsf(a) < 0.00001f) return 0;
    float f = 1.0f / a;
    float sx = ray->origx - t->v0x, sy = ray->origy - t->v0y, sz = ray->origz - t->v0z;
    float u = f * ((sx * hx + sy * hy) + sz * hz);
    if ((u < 0.0f) | (u > 1.0f)) return 0;
    float qx = sy * e1z - sz * e1y;
    float qy = sz * e1x - sx * e1z;
    float qz = sx * e1y - sy * e1x;
    float v = f * ((ray->dirx * qx + ray->diry * qy) + ray->dirz * qz);
    if ((v < 0.0f) | (u + v > 1.0f)) return 0;
    float r = f * ((e2x * qx + e2y * qy) + e2z * qz);
    if (!(r > 0.0001f)) return 0;
    *tt = r;
    return 1;
}

/* extend.cl:29-38 split into its geometric part (tmax >= tmin && tmax > 0) and tmin */
static inline int box_exact(const ray_t* ray, const node_t* n, float* tminOut)
{
    float tx1 = (n->minx - ray->origx) / ray->dirx, tx2 = (n->maxx - ray->origx) / ray->dirx;
    float tmin = cl_min(tx1, tx2), tmax = cl_max(tx1, tx2);
    float ty1 = (n->miny - ray->origy) / ray->diry, ty2 = (n->maxy - ray->origy) / ray->diry;
    tmin = cl_max(tmin, cl_min(ty1, ty2));
    tmax = cl_min(tmax, cl_max(ty1, ty2));
    float tz1 = (n->minz - ray->origz) / ray->dirz, tz2 = (n->maxz - ray->origz) / ray->dirz;
    tmin = cl_max(tmin, cl_min(tz1, tz2));
    tmax = cl_min(tmax, cl_max(tz1, tz2));
    *tminOut = tmin;
    return tmax >= tmin && tmax > 0.0f;
}

#define LAB_STACK 128

/* ---- the reference traversal with extra counters ------------------------------------------------ */
static int exact_trace(ray_t* ray, const tri_t* tri, const node_t* nodes, const uint32_t* triIdx, lab_stats* c, int popCheck)
{
    const node_t* node = &nodes[0];
    const node_t* stack[LAB_STACK];
    float stackT[LAB_STACK];
    float nodeT = -1.0f;
    uint32_t sp = 0;
    int iters = 0;
    for (;;) {
        iters++;
        if (node->triCount > 0) {
            c->leafVisits++;
            for (uint32_t i = 0; i < (uint32_t)node->triCount; i++) {
                uint32_t id = triIdx[node->leftFirst + i];
                c->triTests++;
                float t;
                if (tri_accept(ray, &tri[id], &t)) {
                    c->accepted++;
                    float inv = nodeT - t;
                    if (inv >= 0.0f) {
                        c->inverted++;
                        if (inv > c->maxInvAbs) c->maxInvAbs = inv;
                        if (inv / t > c->maxInvRel) c->maxInvRel = inv / t;
                    }
                    if (t < ray->dist) { ray->dist = t; ray->triID = id; }
                }
            }
        } else {
            c->innerVisits++;
            const node_t* c1 = &nodes[node->leftFirst];
            const node_t* c2 = &nodes[node->leftFirst + 1];
            float t1, t2;
            int g1 = box_exact(ray, c1, &t1), g2 = box_exact(ray, c2, &t2);
            float d1 = (g1 && t1 < ray->dist) ? t1 : 1e30f, d2 = (g2 && t2 < ray->dist) ? t2 : 1e30f;
            if (d1 > d2) {
                float d = d1; d1 = d2; d2 = d;
                const node_t* t = c1; c1 = c2; c2 = t;
            }
            if (d1 != 1e30f) {
                node = c1;
                nodeT = d1;
                if (d2 != 1e30f) { stack[sp] = c2; stackT[sp] = d2; sp++; }
                continue;
            }
        }
        /* pop */
        for (;;) {
            if (sp == 0) return iters;
            --sp;
            node = stack[sp];
            nodeT = stackT[sp];
            c->pops++;
            if (nodeT >= ray->dist) {
                if (node->triCount > 0) c->deadLeaf++;
                else { c->deadInner++; if (popCheck) continue; }   /* provably a no-op visit (nested boxes) */
            }
            break;
        }
    }
}

void lab_exact_stats(const tri_t* tris, ray_t* rays, const node_t* nodes, const uint32_t* triIdx, int64_t nRays, int popCheck,
                     lab_stats* out)
{
    lab_stats tot;
    memset(&tot, 0, sizeof tot);
    const int64_t nGroups = (nRays + 31) / 32;
#pragma omp parallel
    {
        lab_stats c;
        memset(&c, 0, sizeof c);
        int64_t g;
#pragma omp for schedule(dynamic, 128)
        for (g = 0; g < nGroups; g++) {
            int mx = 0, sum = 0;
            for (int64_t i = g * 32; i < nRays && i < g * 32 + 32; i++) {
                ray_t r = rays[i];
                int it = exact_trace(&r, tris, nodes, triIdx, &c, popCheck);
                rays[i].dist = r.dist;
                rays[i].triID = r.triID;
                if (r.dist != 1e30f) c.hits++;
                if (it > mx) mx = it;
                sum += it;
            }
            c.warpIters += (uint64_t)mx;
            c.warpLaneIters += (uint64_t)sum;
        }
#pragma omp critical
        {
            uint64_t* a = (uint64_t*)&tot; const uint64_t* b = (const uint64_t*)&c;
            tot.innerVisits += c.innerVisits; tot.leafVisits += c.leafVisits; tot.triTests += c.triTests; tot.hits += c.hits;
            tot.deadInner += c.deadInner; tot.deadLeaf += c.deadLeaf; tot.pops += c.pops; tot.accepted += c.accepted;
            tot.inverted += c.inverted; tot.warpIters += c.warpIters; tot.warpLaneIters += c.warpLaneIters;
            if (c.maxInvAbs > tot.maxInvAbs) tot.maxInvAbs = c.maxInvAbs;
            if (c.maxInvRel > tot.maxInvRel) tot.maxInvRel = c.maxInvRel;
            (void)a; (void)b;
        }
    }
    tot.rays = (uint64_t)nRays;
    *out = tot;
}

/* ---- the certified fast traversal -------------------------------------------------------------- */
/* Quantised boxes: 15 bits per plane on a grid over the scene box, lo rounded down and hi rounded up, each
 * pushed one further step outwards (the arithmetic slack of the decode).  qbox[6*n..]: lo xyz, hi xyz. */
typedef struct { float gmin[3], step[3]; const uint16_t* q; } qscene;

void lab_quantise(const node_t* nodes, int nNodes, float* gminOut, float* stepOut, uint16_t* q)
{
    float lo[3] = {nodes[0].minx, nodes[0].miny, nodes[0].minz}, hi[3] = {nodes[0].maxx, nodes[0].maxy, nodes[0].maxz};
    for (int a = 0; a < 3; a++) {
        float ext = hi[a] - lo[a];
        float step = ext / 32760.0f;
        if (!(step > 0.0f)) step = 1e-6f;
        stepOut[a] = step;
        gminOut[a] = lo[a] - 3.0f * step;
    }
    for (int i = 0; i < nNodes; i++) {
        const float mn[3] = {nodes[i].minx, nodes[i].miny, nodes[i].minz}, mx[3] = {nodes[i].maxx, nodes[i].maxy, nodes[i].maxz};
        for (int a = 0; a < 3; a++) {
            double l = floor(((double)mn[a] - (double)gminOut[a]) / (double)stepOut[a]) - 1.0;
            double h = ceil(((double)mx[a] - (double)gminOut[a]) / (double)stepOut[a]) + 1.0;
            if (l < 0) l = 0; if (h > 32767) h = 32767;
            if (!(l >= 0)) l = 0; if (!(h <= 32767)) h = 32767;
            q[6 * (size_t)i + a] = (uint16_t)l;
            q[6 * (size_t)i + 3 + a] = (uint16_t)h;
        }
    }
}

typedef struct {
    float S[3], B[3];       /* t = fma(1 + q*2^-15, S, B) */
} qray;

static inline int box_quant(const qray* qr, const ray_t* ray, const uint16_t* q, float dcull, float* tminOut)
{
    float tmin = 0.0f, tmax = dcull;   /* folds "tmax > 0" (as >=) and "tmin < dcull" (as <=): conservative */
    for (int a = 0; a < 3; a++) {
        float fl = 1.0f + (float)q[a] * 3.0517578125e-05f, fh = 1.0f + (float)q[3 + a] * 3.0517578125e-05f;
        float t1 = fmaf(fl, qr->S[a], qr->B[a]), t2 = fmaf(fh, qr->S[a], qr->B[a]);
        float tn = t1 < t2 ? t1 : t2, tf = t1 < t2 ? t2 : t1;
        if (tn > tmin) tmin = tn;
        if (tf < tmax) tmax = tf;
    }
    *tminOut = tmin;
    return tmin <= tmax;
}

typedef struct { int quant; float dRel, dAbs; int popCheck; int order; } fast_cfg;

/* returns 1 when the certificate holds; result in ray */
static int fast_trace(ray_t* ray, const tri_t* tri, const node_t* nodes, const uint32_t* triIdx, const qscene* qs,
                      const fast_cfg* cfg, lab_stats* c, int* itersOut)
{
    qray qr;
    const float o[3] = {ray->origx, ray->origy, ray->origz}, d[3] = {ray->dirx, ray->diry, ray->dirz};
    for (int a = 0; a < 3; a++) {
        float r = 1.0f / d[a];
        float S = (qs->step[a] * 32768.0f) * r;
        qr.S[a] = S;
        qr.B[a] = fmaf(qs->gmin[a] - o[a], r, -S);
    }
    float best = 1e30f, second = 1e30f, bestTmin = 0.0f;
    uint32_t bestTri = 0;
    float dcull = 3e38f;
    uint32_t stack[LAB_STACK];
    float stackT[LAB_STACK];
    uint32_t sp = 0, cur = 0;
    int iters = 0;
    for (;;) {
        iters++;
        const node_t* node = &nodes[cur];
        if (node->triCount > 0) {
            c->leafVisits++;
            for (uint32_t i = 0; i < (uint32_t)node->triCount; i++) {
                uint32_t id = triIdx[node->leftFirst + i];
                c->triTests++;
                float t;
                if (!tri_accept(ray, &tri[id], &t)) continue;
                float tl;
                if (!box_exact(ray, node, &tl)) { c->boxReject++; continue; }   /* the reference never reaches this leaf */
                c->accepted++;
                if (t < best) { second = best; best = t; bestTri = id; bestTmin = tl; dcull = best * (1.0f + 2.0f * cfg->dRel) + 2.0f * cfg->dAbs; }
                else if (t < second) second = t;
            }
        } else {
            c->innerVisits++;
            uint32_t k1 = (uint32_t)node->leftFirst, k2 = k1 + 1;
            float t1, t2;
            int h1, h2;
            if (cfg->quant) {
                h1 = box_quant(&qr, ray, qs->q + 6 * (size_t)k1, dcull, &t1);
                h2 = box_quant(&qr, ray, qs->q + 6 * (size_t)k2, dcull, &t2);
            } else {
                h1 = box_exact(ray, &nodes[k1], &t1) && t1 < dcull;
                h2 = box_exact(ray, &nodes[k2], &t2) && t2 < dcull;
            }
            int swap = h2 && (!h1 || t1 > t2);
            if (swap) { uint32_t k = k1; k1 = k2; k2 = k; float t = t1; t1 = t2; t2 = t; int h = h1; h1 = h2; h2 = h; }
            if (h1) {
                cur = k1;
                if (h2) { stack[sp] = k2; stackT[sp] = t2; sp++; }
                continue;
            }
        }
        for (;;) {
            if (sp == 0) goto done;
            --sp;
            c->pops++;
            if (cfg->popCheck && !(stackT[sp] < dcull)) { c->deadInner++; continue; }
            cur = stack[sp];
            break;
        }
    }
done:
    *itersOut = iters;
    ray->dist = best;
    ray->triID = bestTri;
    if (best == 1e30f) return 1;
    int ok = 1;
    if (!(second > best * (1.0f + cfg->dRel) + cfg->dAbs)) { c->nearTie++; ok = 0; }
    if (!(bestTmin < best * (1.0f + cfg->dRel) + cfg->dAbs)) { c->tminFail++; ok = 0; }
    return ok;
}

/* rays: generated rays (dist = 1e30); ref: the same rays after the exact traversal (lab_exact_stats). */
void lab_fast(const tri_t* tris, ray_t* rays, const ray_t* ref, const node_t* nodes, int nNodes, const uint32_t* triIdx, int64_t nRays,
              int quant, float dRel, float dAbs, int popCheck, lab_stats* out)
{
    lab_stats tot;
    memset(&tot, 0, sizeof tot);
    qscene qs;
    uint16_t* q = (uint16_t*)malloc((size_t)nNodes * 12);
    lab_quantise(nodes, nNodes, qs.gmin, qs.step, q);
    qs.q = q;
    fast_cfg cfg = {quant, dRel, dAbs, popCheck, 0};
    const int64_t nGroups = (nRays + 31) / 32;
#pragma omp parallel
    {
        lab_stats c;
        memset(&c, 0, sizeof c);
        int64_t g;
#pragma omp for schedule(dynamic, 128)
        for (g = 0; g < nGroups; g++) {
            int mx = 0, sum = 0;
            for (int64_t i = g * 32; i < nRays && i < g * 32 + 32; i++) {
                ray_t r = rays[i];
                int it = 0;
                int ok = fast_trace(&r, tris, nodes, triIdx, &qs, &cfg, &c, &it);
                if (!ok) c.certFail++;
                if (r.dist != 1e30f) c.hits++;
                int same = memcmp(&r.dist, &ref[i].dist, 4) == 0 && r.triID == ref[i].triID;
                if (ok && !same) c.mismatch++;          /* a certified ray that differs from the reference: must be 0 */
                if (!same) c.rawMismatch++;             /* what the fast traversal alone (no fallback) would get wrong */
                rays[i].dist = r.dist;
                rays[i].triID = r.triID;
                if (it > mx) mx = it;
                sum += it;
            }
            c.warpIters += (uint64_t)mx;
            c.warpLaneIters += (uint64_t)sum;
        }
#pragma omp critical
        {
            tot.innerVisits += c.innerVisits; tot.leafVisits += c.leafVisits; tot.triTests += c.triTests; tot.hits += c.hits;
            tot.deadInner += c.deadInner; tot.pops += c.pops; tot.accepted += c.accepted; tot.certFail += c.certFail;
            tot.mismatch += c.mismatch; tot.rawMismatch += c.rawMismatch; tot.nearTie += c.nearTie; tot.boxReject += c.boxReject; tot.tminFail += c.tminFail;
            tot.warpIters += c.warpIters; tot.warpLaneIters += c.warpLaneIters;
        }
    }
    free(q);
    tot.rays = (uint64_t)nRays;
    *out = tot;
}

/* ---- lockstep warp model: what would postponing leaf tests buy? -----------------------------------
 * 32 consecutive rays form a warp.  Time advances in rounds; in a round the warp executes the inner-node
 * step (cost cInner) if any lane wants one and the leaf step (cost cLeaf) if any lane wants one -- two
 * divergent paths issued one after the other.  policy 0: a lane that reaches a leaf tests it in the next
 * round (today's kernels).  policy 1: a lane that reaches a leaf parks it (up to `queue` leaves) and keeps
 * traversing with its stale culling distance; the warp runs a leaf round only when at least `minLanes`
 * lanes have a parked leaf or nobody can do anything else.  Order is free in the certified scheme, so both are legal.
 * Returns the summed cost over all warps; out: rounds, inner steps, leaf steps, lane-steps. */
typedef struct {
    uint32_t stack[LAB_STACK];
    uint32_t sp, cur;          /* cur: node index or 0xffffffff */
    uint32_t pend[4];
    int npend, done;
    float best, dcull;
} lane_t;

static inline int lane_next_node(lane_t* L)
{
    if (L->sp == 0) { L->cur = 0xffffffffu; return 0; }
    L->cur = L->stack[--L->sp];
    return 1;
}

double lab_warp_model(const tri_t* tris, const ray_t* rays, const node_t* nodes, const uint32_t* triIdx, int64_t nRays, int policy,
                      int queue, int minLanes, double cInner, double cLeaf, double cVote, float dRel, float dAbs, uint64_t* out4)
{
    double total = 0;
    uint64_t rounds = 0, innerRounds = 0, leafRounds = 0, innerLaneSteps = 0, leafLaneSteps = 0;
    const int64_t nGroups = (nRays + 31) / 32;
    if (queue > 4) queue = 4;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total, rounds, innerRounds, leafRounds, innerLaneSteps, leafLaneSteps)
    for (int64_t g = 0; g < nGroups; g++) {
        lane_t L[32];
        ray_t R[32];
        int n = 0;
        for (int64_t i = g * 32; i < nRays && i < g * 32 + 32; i++, n++) {
            R[n] = rays[i];
            L[n].sp = 0; L[n].cur = 0; L[n].npend = 0; L[n].done = 0; L[n].best = 1e30f; L[n].dcull = 3e38f;
        }
        for (;;) {
            int wantInner = 0, wantLeaf = 0, blocked = 0, alive = 0;
            for (int k = 0; k < n; k++) {
                lane_t* l = &L[k];
                if (l->done) continue;
                alive++;
                int curIsLeaf = l->cur != 0xffffffffu && nodes[l->cur].triCount > 0;
                if (policy == 0) {
                    if (curIsLeaf) wantLeaf++; else if (l->cur != 0xffffffffu) wantInner++;
                } else {
                    if (l->cur != 0xffffffffu && !curIsLeaf) wantInner++;
                    else if (l->npend > 0 || curIsLeaf) { wantLeaf++; if (curIsLeaf || l->cur == 0xffffffffu) blocked++; }
                }
            }
            if (!alive) break;
            int doLeaf, doInner;
            if (policy == 0) { doLeaf = wantLeaf > 0; doInner = wantInner > 0; }
            else {
                int pendLanes = 0;
                for (int k = 0; k < n; k++) if (!L[k].done && (L[k].npend > 0 || (L[k].cur != 0xffffffffu && nodes[L[k].cur].triCount > 0))) pendLanes++;
                doLeaf = pendLanes >= minLanes || wantInner == 0;
                doInner = !doLeaf && wantInner > 0;
                total += cVote;
            }
            rounds++;
            if (doInner) {
                innerRounds++;
                total += cInner;
                for (int k = 0; k < n; k++) {
                    lane_t* l = &L[k];
                    if (l->done || l->cur == 0xffffffffu || nodes[l->cur].triCount > 0) continue;
                    innerLaneSteps++;
                    const node_t* nd = &nodes[l->cur];
                    uint32_t k1 = (uint32_t)nd->leftFirst, k2 = k1 + 1;
                    float t1, t2;
                    int h1 = box_exact(&R[k], &nodes[k1], &t1) && t1 < l->dcull, h2 = box_exact(&R[k], &nodes[k2], &t2) && t2 < l->dcull;
                    if (h2 && (!h1 || t1 > t2)) { uint32_t t = k1; k1 = k2; k2 = t; int h = h1; h1 = h2; h2 = h; }
                    if (h1) { l->cur = k1; if (h2) l->stack[l->sp++] = k2; }
                    else lane_next_node(l);
                    if (policy == 1) {
                        /* park leaves while there is room and something else to do */
                        while (l->cur != 0xffffffffu && nodes[l->cur].triCount > 0 && l->npend < queue) {
                            l->pend[l->npend++] = l->cur;
                            lane_next_node(l);
                        }
                    }
                }
            }
            if (doLeaf) {
                leafRounds++;
                total += cLeaf;
                for (int k = 0; k < n; k++) {
                    lane_t* l = &L[k];
                    if (l->done) continue;
                    uint32_t leaf = 0xffffffffu;
                    if (policy == 0) {
                        if (l->cur != 0xffffffffu && nodes[l->cur].triCount > 0) { leaf = l->cur; lane_next_node(l); }
                    } else {
                        if (l->npend > 0) { leaf = l->pend[0]; for (int q = 1; q < l->npend; q++) l->pend[q - 1] = l->pend[q]; l->npend--; }
                        else if (l->cur != 0xffffffffu && nodes[l->cur].triCount > 0) { leaf = l->cur; lane_next_node(l); }
                    }
                    if (leaf == 0xffffffffu) continue;
                    leafLaneSteps++;
                    const node_t* nd = &nodes[leaf];
                    for (uint32_t i = 0; i < (uint32_t)nd->triCount; i++) {
                        float t;
                        if (tri_accept(&R[k], &tris[triIdx[nd->leftFirst + i]], &t) && t < l->best) {
                            l->best = t;
                            l->dcull = t * (1.0f + 2.0f * dRel) + 2.0f * dAbs;
                        }
                    }
                }
            }
            for (int k = 0; k < n; k++) {
                lane_t* l = &L[k];
                if (!l->done && l->cur == 0xffffffffu && l->sp == 0 && l->npend == 0) l->done = 1;
                else if (!l->done && l->cur == 0xffffffffu && l->sp > 0) lane_next_node(l);
            }
        }
    }
    out4[0] = rounds; out4[1] = innerRounds; out4[2] = leafRounds; out4[3] = innerLaneSteps; out4[4] = leafLaneSteps;
    return total;
}


/* policy with refill: a warp owns `chunk` consecutive rays and puts the next one into a lane as soon as the lane's ray is
 * done (at a round boundary); leaves are parked as in policy 1 of lab_warp_model.  Same cost model. */
double lab_warp_model_refill(const tri_t* tris, const ray_t* rays, const node_t* nodes, const uint32_t* triIdx, int64_t nRays, int chunk,
                             int queue, int minLanes, double cInner, double cLeaf, double cVote, float dRel, float dAbs, uint64_t* out5)
{
    double total = 0;
    uint64_t rounds = 0, innerRounds = 0, leafRounds = 0, innerLaneSteps = 0, leafLaneSteps = 0;
    const int64_t nGroups = (nRays + chunk - 1) / chunk;
    if (queue > 4) queue = 4;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total, rounds, innerRounds, leafRounds, innerLaneSteps, leafLaneSteps)
    for (int64_t g = 0; g < nGroups; g++) {
        lane_t L[32];
        ray_t R[32];
        int64_t next = g * chunk;
        const int64_t end = (g + 1) * chunk < nRays ? (g + 1) * chunk : nRays;
        for (int k = 0; k < 32; k++) L[k].done = 1;
        for (;;) {
            /* refill */
            int alive = 0;
            for (int k = 0; k < 32; k++) {
                if (L[k].done && next < end) {
                    R[k] = rays[next++];
                    L[k].sp = 0; L[k].cur = 0; L[k].npend = 0; L[k].done = 0; L[k].best = 1e30f; L[k].dcull = 3e38f;
                }
                if (!L[k].done) alive++;
            }
            if (!alive) break;
            int wantInner = 0, pendLanes = 0;
            for (int k = 0; k < 32; k++) {
                lane_t* l = &L[k];
                if (l->done) continue;
                int curIsLeaf = l->cur != 0xffffffffu && nodes[l->cur].triCount > 0;
                if (l->cur != 0xffffffffu && !curIsLeaf) wantInner++;
                if (l->npend > 0 || curIsLeaf) pendLanes++;
            }
            const int doLeaf = pendLanes >= minLanes || wantInner == 0;
            const int doInner = !doLeaf && wantInner > 0;
            total += cVote;
            rounds++;
            if (doInner) {
                innerRounds++;
                total += cInner;
                for (int k = 0; k < 32; k++) {
                    lane_t* l = &L[k];
                    if (l->done || l->cur == 0xffffffffu || nodes[l->cur].triCount > 0) continue;
                    innerLaneSteps++;
                    const node_t* nd = &nodes[l->cur];
                    uint32_t k1 = (uint32_t)nd->leftFirst, k2 = k1 + 1;
                    float t1, t2;
                    int h1 = box_exact(&R[k], &nodes[k1], &t1) && t1 < l->dcull, h2 = box_exact(&R[k], &nodes[k2], &t2) && t2 < l->dcull;
                    if (h2 && (!h1 || t1 > t2)) { uint32_t t = k1; k1 = k2; k2 = t; int h = h1; h1 = h2; h2 = h; }
                    if (h1) { l->cur = k1; if (h2) l->stack[l->sp++] = k2; }
                    else lane_next_node(l);
                    while (l->cur != 0xffffffffu && nodes[l->cur].triCount > 0 && l->npend < queue) {
                        l->pend[l->npend++] = l->cur;
                        lane_next_node(l);
                    }
                }
            }
            if (doLeaf) {
                leafRounds++;
                total += cLeaf;
                for (int k = 0; k < 32; k++) {
                    lane_t* l = &L[k];
                    if (l->done) continue;
                    uint32_t leaf = 0xffffffffu;
                    if (l->npend > 0) { leaf = l->pend[0]; for (int q = 1; q < l->npend; q++) l->pend[q - 1] = l->pend[q]; l->npend--; }
                    else if (l->cur != 0xffffffffu && nodes[l->cur].triCount > 0) { leaf = l->cur; lane_next_node(l); }
                    if (leaf == 0xffffffffu) continue;
                    leafLaneSteps++;
                    const node_t* nd = &nodes[leaf];
                    for (uint32_t i = 0; i < (uint32_t)nd->triCount; i++) {
                        float t;
                        if (tri_accept(&R[k], &tris[triIdx[nd->leftFirst + i]], &t) && t < l->best) {
                            l->best = t;
                            l->dcull = t * (1.0f + 2.0f * dRel) + 2.0f * dAbs;
                        }
                    }
                }
            }
            for (int k = 0; k < 32; k++) {
                lane_t* l = &L[k];
                if (!l->done && l->cur == 0xffffffffu && l->sp == 0 && l->npend == 0) l->done = 1;
                else if (!l->done && l->cur == 0xffffffffu && l->sp > 0) lane_next_node(l);
            }
        }
    }
    out5[0] = rounds; out5[1] = innerRounds; out5[2] = leafRounds; out5[3] = innerLaneSteps; out5[4] = leafLaneSteps;
    return total;
}
