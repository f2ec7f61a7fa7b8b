// precomp.h -- shim that lets the reference's bvh.cpp / bvh.h compile UNMODIFIED under g++.
// Test infrastructure only (see oracle/build_ref.sh).  It supplies the handful of
// MSVC / Tmpl8 names bvh.cpp uses; Tri / Mesh restate the layout of the reference's
// mesh.h:6-34 without its GL / tinygltf members' behaviour.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
using namespace std;

typedef unsigned int uint;

// aggregate stand-in for Tmpl8's float3_strict (template/precomp.h:174-181): g++ refuses
// ctor-bearing members inside the anonymous structs of bvh.h:13-14
struct float3_strict {
    float x, y, z;
    float operator[](int n) const { return (&x)[n]; }
};
inline float3_strict operator+(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3_strict operator-(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3_strict operator*(const float3_strict& a, float s) { return float3_strict{a.x * s, a.y * s, a.z * s}; }

// MSVC exposes __m128::m128_f32 (bvh.cpp:141-142)
typedef __m128 uvrt_real_m128;
union uvrt_M128 {
    uvrt_real_m128 v;
    float m128_f32[4];
    uvrt_M128() = default;
    uvrt_M128(uvrt_real_m128 r) : v(r) {}
    operator uvrt_real_m128() const { return v; }
};
#define __m128 uvrt_M128
#define __declspec(x)

// The unmodified builder writes up to 30 nodes past the 2N it asks for:
// over-allocate and zero-fill so the complete tree exists.
inline void* _aligned_malloc(size_t bytes, size_t align)
{
    void* p = nullptr;
    size_t total = bytes + 8192;
    if (posix_memalign(&p, align, total)) return nullptr;
    memset(p, 0, total);
    return p;
}

namespace Tmpl8 {
struct alignas(64) Tri {
    union { float3_strict vertex0; __m128 v0; };
    union { float3_strict vertex1; __m128 v1; };
    union { float3_strict vertex2; __m128 v2; };
    union { float3_strict centroid; __m128 centroid4; };
};
}
class BVH;
namespace Tmpl8 {
class Mesh {
public:
    Tri* triangles;
    int triangleCount;
    float floorHeight;
    BVH* bvh = 0;
};
}
using namespace Tmpl8;
#include "bvh.h"
