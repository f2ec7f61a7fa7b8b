// ref_loader.cpp -- C entry points around the reference's OWN loaders: Mesh::LoadMesh / DetermineFloorHeight
// (mesh.cpp:5-136, tinygltf) and RayTracer::LoadRoute / SaveRoute / UpdatePhotonsPerLight (raytracer.cpp:61-64,
// 228-300, tinyxml2; those lines are cut out of raytracer.cpp by oracle/build_ref.sh into the git-ignored
// oracle/_ref/gen/ -- the rest of that file is OpenCL plumbing that cannot build here).  Test infrastructure only.
#include "precomp.h"
#include <unistd.h>

#include "gen/raytracer_route.inc"

extern "C" {

// Runs Mesh::LoadMesh in `root` (the loader opens "rooms/<modelFile>.glb" relative to the working directory).
// trisOut receives triangleCount x 64 bytes (malloc'ed, caller frees with refload_free).
int refload_mesh(const char* root, const char* modelFile, void** trisOut, int* triangleCount, float* floorHeight, unsigned* nodesUsed)
{
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd) || chdir(root) != 0) return -1;
    Mesh* mesh = new Mesh();
    mesh->triangles = 0;
    mesh->triangleCount = 0;
    strncpy(mesh->modelFile, modelFile, 31);
    mesh->LoadMesh();
    int rc = chdir(cwd);
    (void)rc;
    if (!mesh->triangles || mesh->triangleCount <= 0) return -2;
    size_t bytes = (size_t)mesh->triangleCount * sizeof(Tri);
    *trisOut = malloc(bytes);
    memcpy(*trisOut, mesh->triangles, bytes);
    *triangleCount = mesh->triangleCount;
    *floorHeight = mesh->floorHeight;
    *nodesUsed = mesh->bvh ? mesh->bvh->nodesUsed : 0;
    return 0;
}

void refload_free(void* p) { free(p); }

struct refload_route_params {
    int photonCount, maxIterations;
    float lightIntensity, minDosage, minPower, lightLength, lightHeight;
    int photonsPerLight, positions;
};

// RayTracer::LoadRoute in `root`; xyd receives (x, y, duration) triples.  saveAs != NULL: also SaveRoute under that name.
int refload_route(const char* root, const char* name, refload_route_params* out, float* xyd, int capacity, const char* saveAs)
{
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd) || chdir(root) != 0) return -1;
    RayTracer* rt = new RayTracer();
    char file[32] = "";
    strncpy(file, name, 31);
    rt->LoadRoute(file);
    if (saveAs) {
        char file2[32] = "";
        strncpy(file2, saveAs, 31);
        rt->SaveRoute(file2);
    }
    int rc = chdir(cwd);
    (void)rc;
    out->photonCount = rt->photonCount; out->maxIterations = rt->maxIterations; out->lightIntensity = rt->lightIntensity;
    out->minDosage = rt->minDosage; out->minPower = rt->minPower; out->lightLength = rt->lightLength; out->lightHeight = rt->lightHeight;
    out->photonsPerLight = rt->lightPositions.empty() ? 0 : rt->photonsPerLight;
    out->positions = (int)rt->lightPositions.size();
    for (int i = 0; i < out->positions && i < capacity; i++) {
        xyd[3 * i] = rt->lightPositions[i].position.x;
        xyd[3 * i + 1] = rt->lightPositions[i].position.y;
        xyd[3 * i + 2] = rt->lightPositions[i].duration;
    }
    return 0;
}

} // extern "C"
