// precomp.h -- the slice of the reference's template/precomp.h that the hot-path host classes
// (Mesh, BVH, RayTracer) need: vector types, uint, Timer.  Everything windowing / GL / OpenCL
// related in the reference's precomp.h is out of scope.
// Type layouts follow template/precomp.h:152-181 (float3 is 16 bytes, float3_strict is 12).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

typedef unsigned int uint;

namespace Tmpl8 {

struct float2 {
    float x, y;
};
inline float2 make_float2(float a, float b) { float2 f; f.x = a; f.y = b; return f; }

struct alignas(16) float3 {
    float3() = default;
    float3(float a, float b, float c) : x(a), y(b), z(c), dummy(0) {}
    float x, y, z, dummy;
    float operator[](int n) const { return (&x)[n]; }
};
inline float3 make_float3(float a, float b, float c) { return float3(a, b, c); }

// an aggregate (no constructors): it lives inside the anonymous structs of Tri and BVHNode
struct float3_strict {
    float x, y, z;
    float operator[](int n) const { return (&x)[n]; }
    float& operator[](int n) { return (&x)[n]; }
};
inline float3_strict make_float3_strict(float a, float b, float c) { return float3_strict{a, b, c}; }
inline float3_strict operator+(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3_strict operator-(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3_strict operator*(const float3_strict& a, float s) { return float3_strict{a.x * s, a.y * s, a.z * s}; }

// wall-clock timer (template/precomp.h:277-288)
struct Timer {
    Timer() { reset(); }
    float elapsed() const
    {
        return std::chrono::duration<float>(std::chrono::steady_clock::now() - start).count();
    }
    void reset() { start = std::chrono::steady_clock::now(); }
    std::chrono::steady_clock::time_point start;
};

// Directory that holds rooms/ and positions/ (the reference resolves both against the
// working directory; UVRT_ASSET_ROOT or SetAssetRoot() may point elsewhere).
const std::string& AssetRoot();
void SetAssetRoot(const std::string& dir);

} // namespace Tmpl8

#include "bvh.h"
#include "mesh.h"
#include "raytracer.h"
