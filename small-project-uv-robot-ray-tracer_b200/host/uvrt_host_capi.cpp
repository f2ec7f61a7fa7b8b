// uvrt_host_capi.cpp -- include/uvrt_host.h over the C++ host classes.
#include "precomp.h"
#include "../../include/uvrt_host.h"

using namespace Tmpl8;

struct uvrt_sim {
    Mesh mesh;
    RayTracer rt;
    std::string err;
};

namespace {
int sim_fail(uvrt_sim* s, int code, const std::string& why)
{
    if (s) s->err = why;
    return code;
}
int rt_status(uvrt_sim* s)
{
    if (s->rt.ok) return UVRT_OK;
    s->err = s->rt.lastError;
    return UVRT_ERR_CUDA;
}
void copy_name(char dst[32], const char* src)
{
    strncpy(dst, src, 31);
    dst[31] = 0;
}
} // namespace

extern "C" {

int uvrt_sim_create(uvrt_sim** out, const char* assetRoot, int device)
{
    if (!out) return UVRT_ERR_INVALID;
    uvrt_sim* s = new (std::nothrow) uvrt_sim();
    if (!s) return UVRT_ERR_NO_MEMORY;
    if (assetRoot) SetAssetRoot(assetRoot);
    s->rt.device = device;
    s->rt.saveRouteOnReset = false;   // a library caller decides when files are written
    *out = s;
    return UVRT_OK;
}

void uvrt_sim_destroy(uvrt_sim* s) { delete s; }

const char* uvrt_sim_last_error(const uvrt_sim* s) { return s ? s->err.c_str() : ""; }

int uvrt_sim_load_mesh(uvrt_sim* s, const char* modelFile)
{
    if (!s || !modelFile) return UVRT_ERR_INVALID;
    copy_name(s->mesh.modelFile, modelFile);
    s->mesh.LoadMesh();
    if (!s->mesh.loadedMesh) return sim_fail(s, UVRT_ERR_INVALID, s->mesh.lastError);
    return UVRT_OK;
}

int uvrt_sim_set_device_bvh(uvrt_sim* s, int deviceBvh)
{
    if (!s) return UVRT_ERR_INVALID;
    s->mesh.buildBvhOnLoad = deviceBvh == 0;
    return UVRT_OK;
}

int uvrt_sim_set_whole_scene(uvrt_sim* s, int wholeScene)
{
    if (!s) return UVRT_ERR_INVALID;
    s->mesh.loadWholeScene = wholeScene != 0;
    return UVRT_OK;
}

int uvrt_sim_set_triangles(uvrt_sim* s, const void* tris, int n)
{
    if (!s || !tris || n <= 0) return UVRT_ERR_INVALID;
    s->mesh.SetTriangles((const Tri*)tris, n, true);
    if (!s->mesh.loadedMesh)
        return s->mesh.lastError.empty() ? sim_fail(s, UVRT_ERR_NO_MEMORY, "SetTriangles failed") : sim_fail(s, UVRT_ERR_INVALID, s->mesh.lastError);
    return UVRT_OK;
}

int uvrt_sim_mesh_info(const uvrt_sim* s, int* triangleCount, float* floorHeight, unsigned* nodesUsed)
{
    if (!s || !s->mesh.loadedMesh) return UVRT_ERR_INVALID;
    if (triangleCount) *triangleCount = s->mesh.triangleCount;
    if (floorHeight) *floorHeight = s->mesh.floorHeight;
    if (nodesUsed) *nodesUsed = s->mesh.bvh ? s->mesh.bvh->nodesUsed : 0;
    return UVRT_OK;
}

int uvrt_sim_mesh_data(const uvrt_sim* s, const void** tris, const void** nodes, const unsigned** triIdx)
{
    if (!s || !s->mesh.loadedMesh || !s->mesh.bvh) return UVRT_ERR_INVALID;
    if (tris) *tris = s->mesh.triangles;
    if (nodes) *nodes = s->mesh.bvh->bvhNode;
    if (triIdx) *triIdx = s->mesh.bvh->triIdx;
    return UVRT_OK;
}

int uvrt_sim_load_route(uvrt_sim* s, const char* name)
{
    if (!s || !name) return UVRT_ERR_INVALID;
    char n[32];
    copy_name(n, name);
    s->rt.LoadRoute(n);
    return UVRT_OK;
}

int uvrt_sim_save_route(uvrt_sim* s, const char* name)
{
    if (!s || !name) return UVRT_ERR_INVALID;
    char n[32];
    copy_name(n, name);
    s->rt.SaveRoute(n);
    return UVRT_OK;
}

int uvrt_sim_get_params(const uvrt_sim* s, uvrt_sim_params* p)
{
    if (!s || !p) return UVRT_ERR_INVALID;
    const RayTracer& r = s->rt;
    p->photonCount = r.photonCount;
    p->maxIterations = r.maxIterations;
    p->lightIntensity = r.lightIntensity;
    p->minDosage = r.minDosage;
    p->minPower = r.minPower;
    p->lightLength = r.lightLength;
    p->lightHeight = r.lightHeight;
    p->viewMode = (int)r.viewMode;
    p->thresholdView = r.thresholdView ? 1 : 0;
    p->photonsPerLight = r.photonsPerLight;
    p->currIterations = r.currIterations;
    p->photonMapSize = r.photonMapSize;
    p->seedState = r.seedState;
    p->finishedComputation = r.finishedComputation ? 1 : 0;
    return UVRT_OK;
}

int uvrt_sim_set_params(uvrt_sim* s, const uvrt_sim_params* p)
{
    if (!s || !p) return UVRT_ERR_INVALID;
    RayTracer& r = s->rt;
    r.photonCount = p->photonCount;
    r.maxIterations = p->maxIterations;
    r.lightIntensity = p->lightIntensity;
    r.minDosage = p->minDosage;
    r.minPower = p->minPower;
    r.lightLength = p->lightLength;
    r.lightHeight = p->lightHeight;
    r.viewMode = (ViewMode)p->viewMode;
    r.thresholdView = p->thresholdView != 0;
    r.UpdatePhotonsPerLight();
    return UVRT_OK;
}

int uvrt_sim_get_positions(const uvrt_sim* s, float* xyd, int capacity, int* count)
{
    if (!s) return UVRT_ERR_INVALID;
    int n = (int)s->rt.lightPositions.size();
    if (count) *count = n;
    for (int i = 0; xyd && i < n && i < capacity; i++) {
        xyd[3 * i + 0] = s->rt.lightPositions[i].position.x;
        xyd[3 * i + 1] = s->rt.lightPositions[i].position.y;
        xyd[3 * i + 2] = s->rt.lightPositions[i].duration;
    }
    return UVRT_OK;
}

int uvrt_sim_set_positions(uvrt_sim* s, const float* xyd, int count)
{
    if (!s || count < 0 || (count && !xyd)) return UVRT_ERR_INVALID;
    s->rt.lightPositions.clear();
    for (int i = 0; i < count; i++) {
        LightPos lp;
        lp.position = make_float2(xyd[3 * i], xyd[3 * i + 1]);
        lp.duration = xyd[3 * i + 2];
        s->rt.lightPositions.push_back(lp);
    }
    s->rt.UpdatePhotonsPerLight();
    return UVRT_OK;
}

int uvrt_sim_init(uvrt_sim* s, const char* routeName)
{
    if (!s) return UVRT_ERR_INVALID;
    if (!s->mesh.loadedMesh) return sim_fail(s, UVRT_ERR_NO_SCENE, "init: load a mesh first");
    if (routeName) copy_name(s->rt.defaultRouteFile, routeName);
    else s->rt.defaultRouteFile[0] = 0;   // no file of that name: LoadRoute leaves everything as is
    s->rt.ok = true;
    s->rt.Init(&s->mesh);
    if (!s->rt.ok && !s->rt.ctx) return sim_fail(s, UVRT_ERR_NO_DEVICE, s->rt.lastError);
    return rt_status(s);
}

int uvrt_sim_reset_dosage_map(uvrt_sim* s)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    s->rt.ResetDosageMap();
    return rt_status(s);
}

int uvrt_sim_compute_dosage_map(uvrt_sim* s)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    s->rt.ComputeDosageMap();
    return rt_status(s);
}

int uvrt_sim_compute_single(uvrt_sim* s, float x, float y, float duration, int photons, int triangleCount)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    LightPos lp;
    lp.position = make_float2(x, y);
    lp.duration = duration;
    s->rt.ComputeSingleLightDosageMap(lp, photons, triangleCount);
    return rt_status(s);
}

int uvrt_sim_shade(uvrt_sim* s)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    s->rt.Shade();
    return rt_status(s);
}

// One frame of MyApp::Tick.  sync: the reference ends every frame with clFinish (myapp.cpp:165); a whole run
// (uvrt_sim_run) only synchronises once at its end, so the launches of consecutive passes keep overlapping.
static int sim_tick(uvrt_sim* s, int* finished, bool sync)
{
    RayTracer& rt = s->rt;
    if (!rt.finishedComputation) {
        rt.finishedComputation = rt.currIterations >= rt.maxIterations;
        if (!rt.finishedComputation) {
            rt.ComputeDosageMap();
            if (rt.shardCount <= 1 && sync) rt.Shade();   // a partial map is not worth shading; a run shades once at its end
            if (rt.viewMode == texture) rt.viewMode = dosage;
            rt.currIterations++;
            rt.progress = 100.0f * (float)rt.currIterations / (float)rt.maxIterations;
            if (sync && rt.ok && uvrt_sync(rt.ctx) != UVRT_OK) { rt.ok = false; rt.lastError = uvrt_last_error(rt.ctx); }
            rt.compTime += rt.timerClock.elapsed();
            rt.timerClock.reset();
        }
    }
    if (finished) *finished = rt.finishedComputation ? 1 : 0;
    return rt_status(s);
}

int uvrt_sim_tick(uvrt_sim* s, int* finished)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    return sim_tick(s, finished, true);
}

int uvrt_sim_run(uvrt_sim* s, float* dose, int capacity)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    RayTracer& rt = s->rt;
    rt.ResetDosageMap();
    int finished = 0, rc = UVRT_OK;
    while (!finished && rc == UVRT_OK) rc = sim_tick(s, &finished, false);
    if (rc != UVRT_OK) return rc;
    rt.Reduce();               // sharded runs: all-reduce + fold of the pending count-matrix rows; otherwise a no-op
    rt.Shade();
    if (dose) return uvrt_sim_read_dose(s, dose, capacity);
    if (rt.ok && uvrt_sync(rt.ctx) != UVRT_OK) { rt.ok = false; rt.lastError = uvrt_last_error(rt.ctx); }
    return rt_status(s);
}

int uvrt_sim_calibrate(uvrt_sim* s, float measurePower, float measureHeight, float measureDist, float* calibratedPower)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    s->rt.CalibratePower(measurePower, measureHeight, measureDist);
    if (calibratedPower) *calibratedPower = s->rt.calibratedPower;
    return rt_status(s);
}

int uvrt_sim_read_dose(uvrt_sim* s, float* dst, int capacity)
{
    if (!s || !s->rt.ctx || !dst) return UVRT_ERR_INVALID;
    if (capacity < s->mesh.triangleCount) return sim_fail(s, UVRT_ERR_INVALID, "read_dose: capacity too small");
    const float* d = s->rt.ReadDosageMap();
    if (!s->rt.ok) return rt_status(s);
    memcpy(dst, d, sizeof(float) * (size_t)s->mesh.triangleCount);
    return UVRT_OK;
}

int uvrt_sim_set_shard(uvrt_sim* s, int rank, int count)
{
    if (!s || count < 1 || rank < 0 || rank >= count) return UVRT_ERR_INVALID;
    s->rt.shardRank = rank;
    s->rt.shardCount = count;
    return UVRT_OK;
}

int uvrt_sim_set_shard_parts(uvrt_sim* s, int parts)
{
    if (!s || parts < 0 || parts > 64) return UVRT_ERR_INVALID;
    s->rt.shardParts = parts;
    return UVRT_OK;
}

int uvrt_sim_shard_parts(const uvrt_sim* s) { return s ? s->rt.AutoParts() : 0; }

int uvrt_sim_set_cost_aware(uvrt_sim* s, int on)
{
    if (!s) return UVRT_ERR_INVALID;
    s->rt.costAwareSharding = on != 0;
    return UVRT_OK;
}

int uvrt_host_plan_shards(const double* launchCost, int launches, int ranks, int* ownerOut)
{
    if (!launchCost || !ownerOut || launches < 0 || ranks < 1) return UVRT_ERR_INVALID;
    RayTracer::PlanShardsLPT(launchCost, launches, ranks, ownerOut);
    return UVRT_OK;
}

int uvrt_sim_set_seed(uvrt_sim* s, uint32_t seed)
{
    if (!s) return UVRT_ERR_INVALID;
    s->rt.seedState = seed;
    return UVRT_OK;
}

uint32_t uvrt_host_seed_after_launch(float lx, float ly, float lz, float lightLength, uint32_t seedIn)
{
    return RayTracer::SeedAfterLaunch(lx, ly, lz, lightLength, seedIn);
}

int uvrt_sim_reduce(uvrt_sim* s)
{
    if (!s || !s->rt.ctx) return UVRT_ERR_INVALID;
    s->rt.Reduce();
    return rt_status(s);
}

int uvrt_sim_save_dosage_map(uvrt_sim* s, const char* basePath)
{
    if (!s || !s->rt.ctx || !basePath) return UVRT_ERR_INVALID;
    if (!s->rt.SaveDosageMap(basePath)) return s->rt.ok ? sim_fail(s, UVRT_ERR_IO, s->rt.lastError) : rt_status(s);
    return UVRT_OK;
}

int uvrt_sim_save_checkpoint(uvrt_sim* s, const char* path)
{
    if (!s || !s->rt.ctx || !path) return UVRT_ERR_INVALID;
    if (!s->rt.SaveCheckpoint(path)) return s->rt.ok ? sim_fail(s, UVRT_ERR_IO, s->rt.lastError) : rt_status(s);
    return UVRT_OK;
}

int uvrt_sim_load_checkpoint(uvrt_sim* s, const char* path)
{
    if (!s || !s->rt.ctx || !path) return UVRT_ERR_INVALID;
    if (!s->rt.LoadCheckpoint(path)) return s->rt.ok ? sim_fail(s, UVRT_ERR_IO, s->rt.lastError) : rt_status(s);
    return UVRT_OK;
}

int uvrt_host_shard_owner(long long launch, int positions, int ranks) { return RayTracer::ShardOwner(launch, positions, ranks); }

uvrt_ctx* uvrt_sim_ctx(uvrt_sim* s) { return s ? s->rt.ctx : nullptr; }

int64_t uvrt_sim_rays_traced(const uvrt_sim* s) { return s ? s->rt.RaysTraced() : 0; }

int uvrt_host_build_bvh(void* tris, int n, void* nodesOut, int nodeCapacity, unsigned* triIdxOut, unsigned* nodesUsed)
{
    if (!tris || n <= 0 || !nodesOut || !triIdxOut) return UVRT_ERR_INVALID;
    Mesh m;
    m.SetTriangles((const Tri*)tris, n, true);
    if (!m.loadedMesh && !m.lastError.empty()) return UVRT_ERR_INVALID;   // non-finite coordinates
    if (!m.bvh || !m.bvh->bvhNode) return UVRT_ERR_NO_MEMORY;
    if ((int)m.bvh->nodesUsed > nodeCapacity) return UVRT_ERR_INVALID;
    memcpy(nodesOut, m.bvh->bvhNode, sizeof(BVHNode) * (size_t)m.bvh->nodesUsed);
    memcpy(triIdxOut, m.bvh->triIdx, sizeof(uint) * (size_t)n);
    memcpy(tris, m.triangles, sizeof(Tri) * (size_t)n);   // centroids
    if (nodesUsed) *nodesUsed = m.bvh->nodesUsed;
    return UVRT_OK;
}

} // extern "C"
