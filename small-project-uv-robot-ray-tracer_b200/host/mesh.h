// mesh.h -- Tri / Mesh with the reference's public surface (mesh.h:6-34 of the reference).
// The GL object ids are kept as plain fields for source compatibility; nothing binds them
// (display is out of scope).
#pragma once

class BVH;

namespace Tmpl8 {

// 64-byte triangle, identical in memory to the reference's Tri and to cl/tools.cl:31-37
struct alignas(64) Tri {
    union { float3_strict vertex0; float v0[4]; };
    union { float3_strict vertex1; float v1[4]; };
    union { float3_strict vertex2; float v2[4]; };
    union { float3_strict centroid; float centroid4[4]; };
};
static_assert(sizeof(Tri) == 64, "Tri must stay 64 bytes");

class Mesh {
public:
    Mesh() = default;
    ~Mesh();
    Mesh(const Mesh&) = delete;
    Mesh& operator=(const Mesh&) = delete;

    void LoadMesh();             // rooms/<modelFile>.glb -> triangles, floorHeight, bvh
    void BindMesh();             // no-op: the reference uploads GL buffers here
    void DetermineFloorHeight();
    // Builds a mesh from caller-supplied triangles (n x 64 B, reference layout) instead of a file.
    void SetTriangles(const Tri* tris, int n, bool buildBvh = true);

    char modelFile[32] = "C046_1";
    // false: LoadMesh / SetTriangles leave `bvh` empty and RayTracer::Init builds it on the device
    // (uvrt_build_bvh) -- the reference always builds on the CPU inside LoadMesh (mesh.cpp:96).
    bool buildBvhOnLoad = true;
    // true: load every TRIANGLES primitive of every mesh the default scene instances, with the node
    // transforms applied (u8/u16/u32 or no indices).  false (default) = the reference: only
    // meshes[0].primitives[0], transforms ignored (mesh.cpp:28).
    bool loadWholeScene = false;

    Tri* triangles = 0;
    int triangleCount = 0;
    float* vertices = 0;   // 9 floats per triangle (de-indexed positions)
    int vertexCount = 0;   // number of floats in `vertices` (as in the reference)
    float* uvcoords = 0;   // 6 floats per triangle (zeros when the file has no TEXCOORD_0)
    unsigned int VAO = 0, VBO = 0, UVBuffer = 0, textureBuffer = 0;
    uint dosageBufferID = 0;
    bool loadedMesh = false;
    float floorHeight = 0;
    std::string lastError;   // why LoadMesh failed (the reference only prints)

    BVH* bvh = 0;

private:
    void Release();
    void LoadMeshImpl();
};

} // namespace Tmpl8
