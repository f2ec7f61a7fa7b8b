/*
 * uvrt_host.h -- flat C view of the host-side drop-in classes (Mesh, BVH, RayTracer in
 * small-project-uv-robot-ray-tracer_b200/host/, same public surface as the reference's
 * mesh.h / bvh.h / raytracer.h) for callers that cannot include C++ headers: the Python
 * tests and bench.py (ctypes).  C++ callers use the classes directly.
 *
 * A uvrt_sim owns one Tmpl8::Mesh and one Tmpl8::RayTracer, i.e. what MyApp holds in the
 * reference (myapp.h).  Every function returns 0 or a negative uvrt_status (uvrt.h).
 */
#ifndef UVRT_HOST_H
#define UVRT_HOST_H

#include "uvrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvrt_sim uvrt_sim;

/* Mirror of RayTracer's tunable / progress fields (raytracer.h:28-56 of the reference). */
typedef struct uvrt_sim_params {
    int photonCount;       /* aantal_fotonen */
    int maxIterations;     /* aantal_iteraties */
    float lightIntensity;  /* lamp_sterkte */
    float minDosage;       /* minimale_dosis */
    float minPower;        /* minimale_bestralingssterkte */
    float lightLength;     /* lamp_lengte */
    float lightHeight;     /* lamp_hoogte */
    int viewMode;          /* 0 dosage, 1 maxpower, 2 texture */
    int thresholdView;
    /* read-only progress */
    int photonsPerLight;
    int currIterations;
    int photonMapSize;
    uint32_t seedState;
    int finishedComputation;
} uvrt_sim_params;

/* assetRoot: directory holding rooms/ and positions/ (NULL or "": working directory, as in the
 * reference).  device: CUDA ordinal used by Init. */
int uvrt_sim_create(uvrt_sim** out, const char* assetRoot, int device);
void uvrt_sim_destroy(uvrt_sim* sim);
const char* uvrt_sim_last_error(const uvrt_sim* sim);

/* Mesh::LoadMesh (mesh.cpp:5-98): rooms/<modelFile>.glb -> Tri[], floor height, BVH.  CPU only. */
int uvrt_sim_load_mesh(uvrt_sim* sim, const char* modelFile);
/* deviceBvh != 0: load_mesh / set_triangles skip the CPU BVH build and uvrt_sim_init builds the tree
 * on the device (uvrt_build_bvh); the result is the same tree. */
int uvrt_sim_set_device_bvh(uvrt_sim* sim, int deviceBvh);
/* wholeScene != 0: load_mesh reads every triangle primitive of the default scene's node tree with the node
 * transforms applied; 0 (default) = the reference's loader: meshes[0].primitives[0] only (mesh.cpp:28). */
int uvrt_sim_set_whole_scene(uvrt_sim* sim, int wholeScene);
/* Mesh from caller-supplied triangles (n x 64 B, reference layout); builds the BVH.  CPU only. */
int uvrt_sim_set_triangles(uvrt_sim* sim, const void* tris, int n);
int uvrt_sim_mesh_info(const uvrt_sim* sim, int* triangleCount, float* floorHeight, unsigned* nodesUsed);
/* Borrowed pointers into the mesh: Tri[triangleCount], BVHNode[nodesUsed], triIdx[triangleCount]. */
int uvrt_sim_mesh_data(const uvrt_sim* sim, const void** tris, const void** nodes, const unsigned** triIdx);

/* RayTracer::LoadRoute / SaveRoute (raytracer.cpp:233-300): positions/<name>.xml.  CPU only. */
int uvrt_sim_load_route(uvrt_sim* sim, const char* name);
int uvrt_sim_save_route(uvrt_sim* sim, const char* name);
int uvrt_sim_get_params(const uvrt_sim* sim, uvrt_sim_params* p);
int uvrt_sim_set_params(uvrt_sim* sim, const uvrt_sim_params* p);   /* writable fields only */
/* lightPositions as (x, y, duration) triples */
int uvrt_sim_get_positions(const uvrt_sim* sim, float* xyd, int capacity, int* count);
int uvrt_sim_set_positions(uvrt_sim* sim, const float* xyd, int count);

/* RayTracer::Init (raytracer.cpp:12-59): loads positions/<routeName>.xml (NULL: keep the current
 * route and parameters), creates the backend context, uploads the scene.  Needs a GPU. */
int uvrt_sim_init(uvrt_sim* sim, const char* routeName);
int uvrt_sim_reset_dosage_map(uvrt_sim* sim);            /* RayTracer::ResetDosageMap */
int uvrt_sim_compute_dosage_map(uvrt_sim* sim);          /* RayTracer::ComputeDosageMap */
int uvrt_sim_compute_single(uvrt_sim* sim, float x, float y, float duration, int photons, int triangleCount);
int uvrt_sim_shade(uvrt_sim* sim);                       /* RayTracer::Shade */
/* One frame of MyApp::Tick (myapp.cpp:156-175): ComputeDosageMap + Shade + currIterations++ +
 * device sync.  *finished = 1 once currIterations >= maxIterations (nothing is computed then). */
int uvrt_sim_tick(uvrt_sim* sim, int* finished);
/* ResetDosageMap, then Tick until finished, then (sharded runs) Reduce + Shade.  The dose map is
 * copied into dose[0..triangleCount) when dose != NULL. */
int uvrt_sim_run(uvrt_sim* sim, float* dose, int capacity);
int uvrt_sim_calibrate(uvrt_sim* sim, float measurePower, float measureHeight, float measureDist, float* calibratedPower);
int uvrt_sim_read_dose(uvrt_sim* sim, float* dst, int capacity);
/* Work sharing over ranks and the cross-rank reduction (RayTracer::shardRank/shardCount/shardParts, Reduce):
 * every launch is cut into `parts` ray ranges (0 = chosen per run, 1 = whole launches), unit = launch * parts +
 * part goes to rank uvrt_host_shard_owner(unit, positions * parts, count); the ranks' integer counts meet in a
 * count matrix that uvrt_sim_reduce sums (one ncclAllReduce) and folds in launch order. */
int uvrt_sim_set_shard(uvrt_sim* sim, int rank, int count);
int uvrt_sim_set_shard_parts(uvrt_sim* sim, int parts);
int uvrt_sim_shard_parts(const uvrt_sim* sim);           /* the value in effect for the next run */
/* parts == 0 and cost-aware sharing on (default): ResetDosageMap probes the relative cost of every lamp position
 * (uvrt_probe_cost; deterministic, identical on every rank) and deals the whole launches of the run longest-first to
 * the least loaded rank (uvrt_host_plan_shards); off: the rotation of uvrt_host_shard_owner with halves for short runs. */
int uvrt_sim_set_cost_aware(uvrt_sim* sim, int on);
int uvrt_host_plan_shards(const double* launchCost, int launches, int ranks, int* ownerOut);
int uvrt_sim_reduce(uvrt_sim* sim);
/* The rank that traces unit `unit` (counted over the whole run) of a run with `unitsPerPass` units per pass on
 * `ranks` GPUs (RayTracer::ShardOwner). */
int uvrt_host_shard_owner(long long unit, int unitsPerPass, int ranks);
/* The device-side SEED (generate.cl:6) the next launch will see; it survives ResetDosageMap like the
 * reference's program-scope variable.  Setting it to 0 reproduces a freshly started application. */
int uvrt_sim_set_seed(uvrt_sim* sim, uint32_t seed);
/* SEED after a launch at (lx, ly, lz): work-item 0 of generate.cl:13-39 replayed on the host. */
uint32_t uvrt_host_seed_after_launch(float lx, float ly, float lz, float lightLength, uint32_t seedIn);
/* Result export: <basePath>.dose.f32 (float32 per triangle), .ply (per-vertex colours of dosageToColor),
 * .json (parameters).  Call after uvrt_sim_shade / uvrt_sim_run. */
int uvrt_sim_save_dosage_map(uvrt_sim* sim, const char* basePath);
/* Checkpoint / resume: photon and max maps, counters and the SEED chain.  load_checkpoint replaces
 * ResetDosageMap at the start of a resumed run; continue with uvrt_sim_tick. */
int uvrt_sim_save_checkpoint(uvrt_sim* sim, const char* path);
int uvrt_sim_load_checkpoint(uvrt_sim* sim, const char* path);
/* The backend context behind the RayTracer (valid after uvrt_sim_init), for uvrt_read & co. */
uvrt_ctx* uvrt_sim_ctx(uvrt_sim* sim);
int64_t uvrt_sim_rays_traced(const uvrt_sim* sim);

/* The BVH builder on its own (bvh.cpp): tris is n x 64 B (centroids are written), nodesOut holds
 * nodeCapacity >= 2n+64 slots of 32 B, triIdxOut n u32.  CPU only. */
int uvrt_host_build_bvh(void* tris, int n, void* nodesOut, int nodeCapacity, unsigned* triIdxOut, unsigned* nodesUsed);

#ifdef __cplusplus
}
#endif
#endif /* UVRT_HOST_H */
