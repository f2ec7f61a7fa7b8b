// clemu.h -- the slice of OpenCL C the reference's cl/*.cl files use, as C++.
// Test infrastructure only.  Strict IEEE: build with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

typedef unsigned int uint;
#define __kernel
#define __global

struct alignas(16) float3 {
    float x, y, z, pad_;
    float3() = default;
    float3(float a, float b, float c) : x(a), y(b), z(c), pad_(0) {}
    float3(float a) : x(a), y(a), z(a), pad_(0) {}           // OpenCL scalar widening
};
struct double2 {
    double x, y;
    double2() = default;
    double2(double a, double b) : x(a), y(b) {}
};
// vector literals become BRACED initialisers: operands are evaluated left to right
// (a function call would let g++ evaluate them right to left; SURVEY App. C pitfall)
#define make_f3(...) float3{__VA_ARGS__}
#define make_d2(...) double2{__VA_ARGS__}

inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline double2 operator*(double2 a, double s) { return double2(a.x * s, a.y * s); }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double dot(double2 a, double2 b) { return a.x * b.x + a.y * b.y; }
inline float length(float3 a) { return sqrtf(dot(a, a)); }
// OpenCL C 6.12.4: min(x,y) = y < x ? y : x ; max(x,y) = x < y ? y : x
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline double max(double x, double y) { return x < y ? y : x; }
inline float fabs(float x) { return fabsf(x); }
using std::sqrt;

extern thread_local size_t uvrt_clemu_gid;
inline size_t get_global_id(int) { return uvrt_clemu_gid; }
inline int atomic_inc(volatile int* p) { return __atomic_fetch_add(p, 1, __ATOMIC_RELAXED); }

// float -> u32 the way NVIDIA hardware converts (saturating)
inline unsigned int uvrt_sat_u32(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (unsigned int)f;
}
