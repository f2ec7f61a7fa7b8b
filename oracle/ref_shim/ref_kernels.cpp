// ref_kernels.cpp -- drives the reference's OWN kernel sources (sed-transformed copies in
// oracle/_ref/gen/, never committed) one work-item at a time.  Test infrastructure only.
//
// Launch semantics fixed in SURVEY.md App. B: SEED is set before the launch and read by
// every work-item; work-item 0 runs LAST so that its write to SEED is the launch's output.
#include "clemu.h"
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

thread_local size_t uvrt_clemu_gid = 0;

namespace k_generate {
unsigned int WangHash(unsigned int s);
inline unsigned int WangHash(float f) { return WangHash(uvrt_sat_u32(f)); }
#include "gen/generate.cl.inc"
}
namespace k_extend {
#include "gen/extend.cl.inc"
}
namespace k_accumulate {
#include "gen/accumulate.cl.inc"
}
namespace k_shade {
#include "gen/shade.cl.inc"
}
namespace k_reset {
#include "gen/reset.cl.inc"
}

extern "C" {

void ref_generate(void* rays, long long firstRay, long long nRays, float lx, float ly, float lz,
                  float lightLength, unsigned seedIn, unsigned* seedOut)
{
    // rays points at slot 0 of the LAUNCH (the kernel indexes by global id)
    k_generate::Ray* base = (k_generate::Ray*)rays - firstRay;
    float3 lp(lx, ly, lz);
    k_generate::SEED = seedIn;
    long long i;
#pragma omp parallel for schedule(static)
    for (i = firstRay; i < firstRay + nRays; i++) {
        if (i == 0) continue;
        uvrt_clemu_gid = (size_t)i;
        k_generate::render(base, lp, lightLength);
    }
    // work-item 0 last; when it is outside the range its ray goes to a scratch slot
    k_generate::Ray scratch;
    uvrt_clemu_gid = 0;
    k_generate::render(firstRay == 0 && nRays > 0 ? base : &scratch, lp, lightLength);
    if (seedOut) *seedOut = k_generate::SEED;
}

void ref_extend(int* tempPhotonMap, void* tris, void* rays, void* nodes, unsigned* triIdx,
                long long nRays, int triangleCount, int nThreads)
{
#ifdef _OPENMP
    if (nThreads <= 0) nThreads = omp_get_max_threads();
#else
    nThreads = 1;
#endif
    long long i;
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nThreads)
    for (i = 0; i < nRays; i++) {
        uvrt_clemu_gid = (size_t)i;
        k_extend::render(tempPhotonMap, (k_extend::Triangle*)tris, (k_extend::Ray*)rays,
                         (k_extend::BVHNode*)nodes, triIdx, triangleCount);
    }
}

void ref_accumulate(double* photonMap, double* maxPhotonMap, int* temp, float timeStep, int n)
{
    for (int i = 0; i < n; i++) {
        uvrt_clemu_gid = (size_t)i;
        k_accumulate::render(photonMap, maxPhotonMap, temp, timeStep);
    }
}

void ref_compute_dosage(double* photonMap, float* dosage, void* tris, int photonsPerLight,
                        float scaledPower, int n)
{
    for (int i = 0; i < n; i++) {
        uvrt_clemu_gid = (size_t)i;
        k_shade::computeDosage(photonMap, dosage, (k_shade::Triangle*)tris, photonsPerLight, scaledPower);
    }
}

void ref_dosage_to_color(float* dosage, float* color9, float minValue, int thresholdView, int n)
{
    for (int i = 0; i < n; i++) {
        uvrt_clemu_gid = (size_t)i;
        k_shade::dosageToColor(dosage, (k_shade::TriangleColor*)color9, minValue, thresholdView);
    }
}

void ref_reset(double* photonMap, double* maxPhotonMap, int* temp, float* color9, int resetColor, int n)
{
    for (int i = 0; i < n; i++) {
        uvrt_clemu_gid = (size_t)i;
        k_reset::render(photonMap, maxPhotonMap, temp, (k_reset::TriangleColor*)color9, resetColor);
    }
}

void ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
