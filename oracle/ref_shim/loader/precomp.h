// precomp.h (loader flavour) -- lets the reference's mesh.cpp, mesh.h, bvh.cpp, bvh.h and raytracer.h compile
// UNMODIFIED under g++ so that its own GLB / route loaders (tinygltf, tinyxml2 from /root/reference/lib) can be
// run next to this repo's readers.  Test infrastructure only (oracle/build_ref.sh);
// nothing of the reference is copied: the sources are compiled where they lie.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <immintrin.h>
using namespace std;

typedef unsigned int uint;

struct float2 { float x, y; };
// same members as template/precomp.h:174-181, as an aggregate (g++ refuses ctor-bearing members inside the
// anonymous structs of bvh.h:13-14)
struct float3_strict {
    float x, y, z;
    float operator[](int n) const { return (&x)[n]; }
};
inline float3_strict make_float3_strict(const float& a, const float& b, const float& c) { return float3_strict{a, b, c}; }
inline float3_strict operator+(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3_strict operator-(const float3_strict& a, const float3_strict& b) { return float3_strict{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3_strict operator*(const float3_strict& a, float s) { return float3_strict{a.x * s, a.y * s, a.z * s}; }

typedef __m128 uvrt_real_m128;
union uvrt_M128 {
    uvrt_real_m128 v;
    float m128_f32[4];
    uvrt_M128() = default;
    uvrt_M128(uvrt_real_m128 r) : v(r) {}
    operator uvrt_real_m128() const { return v; }
};
#define __m128 uvrt_M128
#define __declspec(x) alignas(64)      // mesh.h:6: __declspec(align(64)) struct Tri

inline void* _aligned_malloc(size_t bytes, size_t align)
{
    void* p = nullptr;
    size_t total = bytes + 8192;       // the builder overruns its own request
    if (posix_memalign(&p, align, total)) return nullptr;
    memset(p, 0, total);
    return p;
}

// OpenGL: Mesh::BindMesh (mesh.cpp:138-197) uploads to GL objects nobody reads here -- no-op stand-ins
typedef unsigned int GLenum;
enum { GL_ARRAY_BUFFER, GL_STATIC_DRAW, GL_DYNAMIC_DRAW, GL_FLOAT, GL_FALSE, GL_TEXTURE_2D, GL_UNPACK_ALIGNMENT, GL_TEXTURE_MIN_FILTER,
       GL_TEXTURE_MAG_FILTER, GL_LINEAR, GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T, GL_REPEAT, GL_RGBA, GL_RED, GL_RG, GL_RGB,
       GL_UNSIGNED_SHORT, GL_UNSIGNED_BYTE };
template <typename... A> inline void glGenVertexArrays(A...) {}
template <typename... A> inline void glGenBuffers(A...) {}
template <typename... A> inline void glBindVertexArray(A...) {}
template <typename... A> inline void glBindBuffer(A...) {}
template <typename... A> inline void glBufferData(A...) {}
template <typename... A> inline void glEnableVertexAttribArray(A...) {}
template <typename... A> inline void glVertexAttribPointer(A...) {}
template <typename... A> inline void glGenTextures(A...) {}
template <typename... A> inline void glBindTexture(A...) {}
template <typename... A> inline void glPixelStorei(A...) {}
template <typename... A> inline void glTexParameterf(A...) {}
template <typename... A> inline void glTexParameteri(A...) {}
template <typename... A> inline void glTexImage2D(A...) {}

// types raytracer.h only points at (or holds by value without using them here)
struct Kernel; struct Buffer; struct ShaderGL;
struct Timer {};

namespace Tmpl8 { class Mesh; }
using namespace Tmpl8;
#include "bvh.h"
#include "mesh.h"
#include "raytracer.h"
