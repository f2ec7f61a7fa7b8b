// raytracer.cpp -- RayTracer on the libuvrt C ABI.
// Host sequencing follows /root/reference/raytracer.cpp:12-300: which stage runs when, the
// scalar arguments each stage gets (photons per light, scaled power, dose divisor) and the
// route-file schema.  The OpenCL plumbing (Kernel, Buffer, SetArgument) is replaced by uvrt_* calls.
#include "precomp.h"
#include "xml_min.h"
#include "../../include/uvrt.h"
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>

namespace Tmpl8 {

RayTracer::~RayTracer()
{
    if (ctx) uvrt_destroy(ctx);
    delete[] dosageMap;
}

bool RayTracer::Check(int rc, const char* what)
{
    if (rc == UVRT_OK) return true;
    ok = false;
    lastError = std::string(what) + ": " + (ctx ? uvrt_last_error(ctx) : uvrt_last_error(nullptr));
    std::cerr << "uvrt error " << rc << " in " << lastError << std::endl;
    return false;
}

void RayTracer::AddLamp()
{
    LightPos initLightPos;
    initLightPos.position = make_float2(0.0f, 0.0f);
    initLightPos.duration = 1;
    lightPositions.push_back(initLightPos);
    UpdatePhotonsPerLight();
}

void RayTracer::UploadScene()
{
    if (!ctx || !mesh || !mesh->bvh) return;
    Check(uvrt_upload_scene(ctx, mesh->triangles, mesh->triangleCount, mesh->bvh->bvhNode, (int)mesh->bvh->nodesUsed,
                            mesh->bvh->triIdx),
          "upload_scene");
}

void RayTracer::Init(Mesh* m)
{
    mesh = m;
    LoadRoute(defaultRouteFile);
    if (!ctx) {
        if (!Check(uvrt_create(&ctx, device), "uvrt_create")) return;
    }
    seedState = 0;
    launchCounter = 0;
    windowRows = windowFill = 0;
    if (!mesh || !mesh->loadedMesh) {
        ok = false;
        lastError = "Init: mesh not loaded";
        return;
    }
    if (!mesh->bvh) {
        // Mesh::buildBvhOnLoad == false: build the tree on the device now
        Timer t;
        mesh->bvh = new BVH(mesh, ctx);
        if (!mesh->bvh->ok) {
            Check(UVRT_ERR_CUDA, "build_bvh");
            delete mesh->bvh;
            mesh->bvh = 0;
            return;
        }
        std::cout << "BVH size: " << mesh->bvh->nodesUsed << " (device build, " << t.elapsed() * 1000.0f << " ms)" << std::endl;
    }
    UploadScene();
}

void RayTracer::UpdatePhotonsPerLight()
{
    // an even count, as in the reference (raytracer.cpp:63)
    if (lightPositions.empty()) { photonsPerLight = 0; return; }
    photonsPerLight = (photonCount / (int)lightPositions.size()) & ~1;
}

void RayTracer::ComputeDosageMap()
{
    if (!ok || lightPositions.empty()) return;
    for (LightPos& lightPosition : lightPositions) {
        ComputeSingleLightDosageMap(lightPosition, photonsPerLight, mesh->triangleCount);
    }
}

// Which rank traces unit u of a run with U units per pass.  Plain round-robin (u mod N) would hand a rank the
// same few positions in every pass whenever gcd(U, N) > 1 (12 positions, 8 ranks: three positions per rank),
// and positions differ in cost by up to 1.5x (24.9 - 36.6 node visits per ray on lange_route), so the slowest
// rank would set the pace.  The deal is therefore rotated by s ranks per pass, with the smallest s that makes
// U + s coprime to N: every position then visits every rank.
int RayTracer::ShardOwner(long long unit, int U, int N)
{
    if (N <= 1) return 0;
    if (U < 1) U = 1;
    auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
    int s = 0;
    while (gcd((long long)U + s, N) != 1) s++;
    const long long pass = unit / U;
    return (int)((unit + (long long)s * pass) % N);
}

// Ray ranges per launch of a sharded run.  Whole launches keep the rays of a launch together (the ray binning of
// extend works best on a full launch: halves cost 2.5 % more device time per ray, quarters 11 %, eighths 47 %,
// profiles/r2_split_sweep.jsonl), but a run of few launches then leaves the ranks unevenly loaded: 120 launches on
// 8 GPUs are 15 each, and positions differ in cost by up to 1.4x -- 8.6 % imbalance in the cost model of DESIGN.md
// section 4, 2.1 % with halves.  So: halves when a rank would get fewer than 24 whole launches, else whole launches.
int RayTracer::AutoParts() const
{
    if (shardParts > 0) return shardParts;
    if (!shardPlan.empty()) return 1;
    if (shardCount <= 1 || lightPositions.empty()) return 1;
    long long launches = (long long)lightPositions.size() * (maxIterations > 0 ? maxIterations : 1);
    return (launches < 24LL * shardCount && photonsPerLight / 2 >= (1 << 19)) ? 2 : 1;
}

// Longest-processing-time-first: launches in order of decreasing cost (ties: launch order), each to the rank with the
// least work so far (ties: lowest rank).  Deterministic, so every rank computes the same plan.
void RayTracer::PlanShardsLPT(const double* launchCost, int launches, int ranks, int* ownerOut)
{
    std::vector<int> order((size_t)launches);
    for (int k = 0; k < launches; k++) order[(size_t)k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return launchCost[a] > launchCost[b]; });
    std::vector<double> load((size_t)ranks, 0.0);
    for (int k : order) {
        int best = 0;
        for (int r = 1; r < ranks; r++)
            if (load[(size_t)r] < load[(size_t)best]) best = r;
        ownerOut[k] = best;
        load[(size_t)best] += launchCost[k];
    }
}

void RayTracer::PlanShards()
{
    shardPlan.clear();
    // sharded runs: both count-matrix buffers at their final size now, not in the middle of the run
    if (ok && shardCount > 1 && mesh) {
        const long long L = (long long)lightPositions.size();
        const long long launches = L * (maxIterations > 0 ? maxIterations : 1);
        const long long rows = launches < MaxWindowRows() ? launches : MaxWindowRows();
        if (rows > 0) Check(uvrt_matrix_reserve(ctx, (int)rows), "matrix_reserve");
    }
    if (!ok || shardCount <= 1 || shardParts > 0 || !costAwareSharding || lightPositions.empty() || !mesh) return;
    const size_t L = lightPositions.size();
    std::vector<float> key;
    key.reserve(3 * L + 3);
    for (const LightPos& lp : lightPositions) { key.push_back(lp.position.x); key.push_back(mesh->floorHeight + lightHeight); key.push_back(lp.position.y); }
    key.push_back(lightLength);
    key.push_back((float)mesh->triangleCount);
    if (key != planKey || positionCost.size() != L) {
        positionCost.assign(L, 0.0);
        for (size_t i = 0; i < L; i++) {
            double inner = 0, tests = 0;
            // issue-slot weights of one inner-node step and one triangle test of the extend kernel (profiles/r2_fast_extend.md)
            if (!Check(uvrt_probe_cost(ctx, key[3 * i], key[3 * i + 1], key[3 * i + 2], lightLength, 0u, 8192, &inner, &tests), "probe_cost")) {
                positionCost.clear();
                planKey.clear();
                ok = true;            // a failed probe only costs the plan, not the run
                return;
            }
            positionCost[i] = 44.0 * inner + 75.0 * tests + 50.0;
        }
        planKey = key;
    }
    const long long launches = (long long)L * (maxIterations > currIterations ? maxIterations - currIterations : 1);
    if (launches > (1 << 22)) return;
    std::vector<double> cost((size_t)launches);
    for (long long k = 0; k < launches; k++) cost[(size_t)k] = positionCost[(size_t)(k % (long long)L)];
    shardPlan.assign((size_t)launches, 0);
    // every window of the count matrix ends in an all-reduce, i.e. the ranks meet there: balance each window on its own
    const long long W = MaxWindowRows();
    for (long long k0 = 0; k0 < launches; k0 += W)
        PlanShardsLPT(cost.data() + k0, (int)std::min(W, launches - k0), shardCount, shardPlan.data() + k0);
}

// generate.cl:13-39 for work-item 0 (the only one that writes SEED): the seed expression in fp32 from left to
// right, float -> uint saturating, two draws for the origin offset and diry, then pairs of
// draws until the point lies inside the unit disc.  This library is built with -ffp-contract=off.
uint32_t RayTracer::SeedAfterLaunch(float lx, float ly, float lz, float /*lightLength*/, uint32_t seedIn)
{
    auto xorshift = [](uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
    auto uniform = [&](uint32_t& s) { return (float)xorshift(s) * 2.3283064365387e-10f; };
    float e = (float)(0 * 17 + 1);
    e = e + lx * 13.0f;
    e = e + ly * 7.0f;
    e = e + lz * 11.0f;
    e = e + (float)(seedIn >> 15);
    uint32_t s = !(e > 0.0f) ? 0u : (e >= 4294967296.0f ? 0xffffffffu : (uint32_t)e);
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s = s ^ (s >> 4);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15);
    uniform(s);                       // origin offset along the lamp (generate.cl:16)
    uniform(s);                       // diry (generate.cl:22)
    double x, z;
    do {                              // generate.cl:25-28; a work-item stuck at RNG state 0 keeps its first draw
        float fx = uniform(s) * 2.0f - 1.0f;
        float fz = uniform(s) * 2.0f - 1.0f;
        x = (double)fx;
        z = (double)fz;
    } while (x * x + z * z > 1.0 && s != 0u);
    return s;
}

// One int32 row per launch: at most 64 MiB and 256 rows at a time -- a window is one allocation and one all-reduce,
// and both should have the size they had in the caller's warm-up passes (a 268 MB first-time window cost 1.2 s of
// allocator and NCCL set-up in the middle of a timed run, profiles/r2_bench_n8_first.json).
long long RayTracer::MaxWindowRows() const
{
    long long maxRows = (64LL << 20) / (4LL * (mesh && mesh->triangleCount > 0 ? mesh->triangleCount : 1));
    if (maxRows < 1) maxRows = 1;
    if (maxRows > 256) maxRows = 256;
    return maxRows;
}

void RayTracer::BeginWindow()
{
    const long long L = (long long)lightPositions.size();
    // launches left in the run as planned (launchCounter restarts at ResetDosageMap); a caller that keeps going
    // beyond maxIterations gets one pass per window
    long long remaining = (long long)maxIterations * L - launchCounter;
    if (remaining < 1) remaining = L - launchCounter % L;
    const long long maxRows = MaxWindowRows();
    windowRows = (int)(remaining < maxRows ? remaining : maxRows);
    windowFill = 0;
    windowDurations.assign((size_t)windowRows, 0.0f);
    Check(uvrt_matrix_begin(ctx, windowRows), "matrix_begin");
}

void RayTracer::FoldWindow()
{
    if (windowFill > 0) Check(uvrt_matrix_fold(ctx, windowDurations.data(), windowFill, 1), "matrix_fold");
    windowRows = windowFill = 0;
}

// photonsPerLight and triangleCount are parameters because CalibratePower() uses its own values.
// (triangleCount only sized the reference's accumulate launch; the backend knows the scene size.)
void RayTracer::ComputeSingleLightDosageMap(LightPos lightPos, int photonsPerLight, int /*triangleCount*/)
{
    if (!ok) return;
    float3 lightposition = make_float3(lightPos.position.x, mesh->floorHeight + lightHeight, lightPos.position.y);
    if (shardCount <= 1) {
        if (!Check(uvrt_trace(ctx, lightposition.x, lightposition.y, lightposition.z, lightLength, lightPos.duration, 0,
                              photonsPerLight, seedState),
                   "trace"))
            return;
        raysTraced += photonsPerLight;
    } else {
        if (windowRows == 0) BeginWindow();
        if (!ok) return;
        const bool planned = !shardPlan.empty() && launchCounter < (long long)shardPlan.size();
        const int parts = planned ? 1 : AutoParts();
        const int U = (int)lightPositions.size() * parts;
        for (int j = 0; j < parts; j++) {
            if ((planned ? shardPlan[(size_t)launchCounter] : ShardOwner(launchCounter * parts + j, U, shardCount)) != shardRank) continue;
            const long long first = (long long)photonsPerLight * j / parts, last = (long long)photonsPerLight * (j + 1) / parts;
            if (last <= first) continue;
            if (!Check(uvrt_trace_row(ctx, windowFill, lightposition.x, lightposition.y, lightposition.z, lightLength, first,
                                      last - first, seedState),
                       "trace_row"))
                return;
            raysTraced += last - first;
        }
        windowDurations[(size_t)windowFill] = lightPos.duration;
        if (++windowFill == windowRows) FoldWindow();
    }
    seedState = SeedAfterLaunch(lightposition.x, lightposition.y, lightposition.z, lightLength, seedState);
    launchCounter++;
    // the reference's `int photonMapSize` (raytracer.h:56) overflows after 2^31 photons (64 default
    // iterations): count in 64 bits and let the int field saturate instead of wrapping
    photonMapSizeTotal += photonsPerLight;
    photonMapSize = photonMapSizeTotal > 0x7fffffffLL ? 0x7fffffff : (int)photonMapSizeTotal;
}

// Photon counts -> dose or irradiance -> heat-map colours (raytracer.cpp:93-120)
void RayTracer::Shade()
{
    if (!ok) return;
    if (viewMode == maxpower) {
        // photons of a single launch; x100: W/m^2 -> microW/cm^2
        if (!Check(uvrt_shade(ctx, 1, photonsPerLight, lightIntensity * 100), "shade")) return;
        Check(uvrt_color(ctx, minPower, thresholdView), "color");
    } else {
        // every photon carries 1/(photons per light) of one lamp's power; x0.1: J/m^2 -> mJ/cm^2
        long long perLight64 = lightPositions.empty() ? 0 : photonMapSizeTotal / (long long)lightPositions.size();
        int perLight = perLight64 > 0x7fffffffLL ? 0x7fffffff : (int)perLight64;
        if (!Check(uvrt_shade(ctx, 0, perLight, lightIntensity * 0.1f), "shade")) return;
        Check(uvrt_color(ctx, minDosage, thresholdView), "color");
    }
}

void RayTracer::Reduce()
{
    if (!ok) return;
    FoldWindow();
}

const float* RayTracer::ReadDosageMap()
{
    if (!ok || !mesh) return dosageMap;
    if (dosageMapSize < mesh->triangleCount) {
        delete[] dosageMap;
        dosageMapSize = mesh->triangleCount;
        dosageMap = new float[dosageMapSize];
    }
    Check(uvrt_read(ctx, UVRT_BUF_DOSE, dosageMap, sizeof(float) * (size_t)mesh->triangleCount), "read dose");
    return dosageMap;
}

void RayTracer::ResetDosageMap()
{
    startedComputation = true;
    compTime = 0;
    timerClock.reset();
    if (saveRouteOnReset) SaveRoute(defaultRouteFile);
    progress = 0;
    finishedComputation = false;
    currIterations = 0;
    launchCounter = 0;
    raysTraced = 0;
    windowRows = windowFill = 0;
    ClearBuffers(true);
    PlanShards();
}

void RayTracer::ClearBuffers(bool resetColor)
{
    photonMapSize = 0;
    photonMapSizeTotal = 0;
    if (!ok) return;
    // the reference also reallocates its 32*photonCount-byte ray buffer here (raytracer.cpp:137);
    // the backend sizes its ray buffer by the largest launch instead
    Check(uvrt_reset(ctx, resetColor ? 1 : 0), "reset");
}

// Scales lightIntensity so that the simulated irradiance on a 0.2 m square at measureDist matches
// a measured value (raytracer.cpp:151-227).
void RayTracer::CalibratePower(float measurePower, float measureHeight, float measureDist)
{
    if (!ok) return;
    measureHeight += mesh->floorHeight;
    LightPos singleLightPos;
    singleLightPos.position = make_float2(0.0f, 0.0f);
    singleLightPos.duration = 0;
    const float w = 0.1f;
    const float px = singleLightPos.position.x, pz = singleLightPos.position.y + measureDist;
    Tri square[2];
    memset(square, 0, sizeof square);
    square[0].vertex0 = make_float3_strict(px + w, measureHeight + w, pz);
    square[0].vertex1 = make_float3_strict(px - w, measureHeight + w, pz);
    square[0].vertex2 = make_float3_strict(px + w, measureHeight - w, pz);
    square[1].vertex0 = make_float3_strict(px - w, measureHeight - w, pz);
    square[1].vertex1 = make_float3_strict(px - w, measureHeight + w, pz);
    square[1].vertex2 = make_float3_strict(px + w, measureHeight - w, pz);
    BVHNode hostNode;
    memset(&hostNode, 0, sizeof hostNode);
    hostNode.leftFirst = 0;
    hostNode.triCount = 2;     // a single leaf; its box is never tested
    uint hostTriIdx[2] = {0, 1};
    if (!Check(uvrt_upload_scene(ctx, square, 2, &hostNode, 1, hostTriIdx), "upload calibration scene")) return;

    ClearBuffers(false);

    const int savedCount = shardCount;
    shardCount = 1;            // every rank calibrates on its own
    for (int i = 0; i < maxIterations; ++i) {
        ComputeSingleLightDosageMap(singleLightPos, photonCount, 2);
    }
    shardCount = savedCount;

    // power 1, so that measured / simulated irradiance is the calibrated power
    float two[2] = {0, 0};
    if (Check(uvrt_shade(ctx, 1, photonCount, 1.0f), "shade") &&
        Check(uvrt_read(ctx, UVRT_BUF_DOSE, two, sizeof two), "read dose")) {
        dosageMap[0] = two[0];
        dosageMap[1] = two[1];
        float avgPower = (dosageMap[0] + dosageMap[1]) / 2.0f;
        calibratedPower = 0.01f * (measurePower / avgPower);
        lightIntensity = calibratedPower;
    }
    UploadScene();             // back to the room (per-triangle buffers are zeroed)
    photonMapSize = 0;
    photonMapSizeTotal = 0;
    std::cout << "Done calibrating " << std::endl;
}

// ---- result export and checkpoints (no counterpart in the reference) -------------------------------
bool RayTracer::SaveDosageMap(const char* basePath)
{
    if (!ok || !mesh) return false;
    const int n = mesh->triangleCount;
    const float* dose = ReadDosageMap();
    if (!ok) return false;
    const std::string base(basePath);
    {
        std::ofstream f(base + ".dose.f32", std::ios::binary);
        if (!f) { lastError = "SaveDosageMap: cannot write " + base + ".dose.f32"; return false; }
        f.write((const char*)dose, sizeof(float) * (size_t)n);
    }
    std::vector<float> color((size_t)n * 9);
    if (!Check(uvrt_read(ctx, UVRT_BUF_COLOR, color.data(), color.size() * sizeof(float)), "read color")) return false;
    {
        std::ofstream f(base + ".ply", std::ios::binary);
        if (!f) { lastError = "SaveDosageMap: cannot write " + base + ".ply"; return false; }
        f << "ply\nformat binary_little_endian 1.0\ncomment uvrt dose map\n"
          << "element vertex " << (size_t)n * 3 << "\nproperty float x\nproperty float y\nproperty float z\n"
          << "property uchar red\nproperty uchar green\nproperty uchar blue\n"
          << "element face " << n << "\nproperty list uchar int vertex_indices\nend_header\n";
        std::vector<char> buf;
        buf.reserve((size_t)n * 3 * 15);
        auto to8 = [](float c) { c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c); return (unsigned char)(c * 255.0f + 0.5f); };
        for (int t = 0; t < n; t++) {
            const float3_strict* v[3] = {&mesh->triangles[t].vertex0, &mesh->triangles[t].vertex1, &mesh->triangles[t].vertex2};
            for (int k = 0; k < 3; k++) {
                const float xyz[3] = {v[k]->x, v[k]->y, v[k]->z};
                const unsigned char rgb[3] = {to8(color[(size_t)t * 9 + k * 3]), to8(color[(size_t)t * 9 + k * 3 + 1]), to8(color[(size_t)t * 9 + k * 3 + 2])};
                buf.insert(buf.end(), (const char*)xyz, (const char*)xyz + 12);
                buf.insert(buf.end(), (const char*)rgb, (const char*)rgb + 3);
            }
        }
        f.write(buf.data(), (std::streamsize)buf.size());
        buf.clear();
        for (int t = 0; t < n; t++) {
            const unsigned char three = 3;
            const int idx[3] = {3 * t, 3 * t + 1, 3 * t + 2};
            buf.push_back((char)three);
            buf.insert(buf.end(), (const char*)idx, (const char*)idx + 12);
        }
        f.write(buf.data(), (std::streamsize)buf.size());
    }
    {
        std::ofstream f(base + ".json", std::ios::binary);
        if (!f) { lastError = "SaveDosageMap: cannot write " + base + ".json"; return false; }
        f << "{\"room\": \"" << mesh->modelFile << "\", \"triangles\": " << n << ", \"floor_height\": " << uvrt_xml::fmt_float(mesh->floorHeight)
          << ", \"view\": \"" << (viewMode == maxpower ? "max_irradiance_uW_cm2" : "dose_mJ_cm2") << "\""
          << ", \"photon_count\": " << photonCount << ", \"photons_per_light\": " << photonsPerLight
          << ", \"iterations\": " << currIterations << ", \"max_iterations\": " << maxIterations
          << ", \"photons_traced\": " << photonMapSizeTotal << ", \"seed_state\": " << seedState
          << ", \"lamp_power\": " << uvrt_xml::fmt_float(lightIntensity) << ", \"lamp_length\": " << uvrt_xml::fmt_float(lightLength)
          << ", \"lamp_height\": " << uvrt_xml::fmt_float(lightHeight) << ", \"min_dose\": " << uvrt_xml::fmt_float(minDosage)
          << ", \"min_power\": " << uvrt_xml::fmt_float(minPower) << ", \"threshold_view\": " << (thresholdView ? "true" : "false")
          << ", \"route\": [";
        for (size_t i = 0; i < lightPositions.size(); i++)
            f << (i ? ", " : "") << "[" << uvrt_xml::fmt_float(lightPositions[i].position.x) << ", " << uvrt_xml::fmt_float(lightPositions[i].position.y)
              << ", " << uvrt_xml::fmt_float(lightPositions[i].duration) << "]";
        f << "], \"files\": {\"dose\": \"float32 x triangles\", \"ply\": \"3 vertices per triangle, uchar rgb\"}}\n";
    }
    return true;
}

namespace {
struct CheckpointHeader {
    char magic[8];            // "UVRTCKP2"
    int32_t triangles, positions, currIterations, photonsPerLight;
    int64_t launchCounter, photonMapSizeTotal, raysTraced;
    uint32_t seedState;
    int32_t shardCount;       // ranks of the run that wrote it (informational: the maps below are complete)
    float lightLength, lightHeight;
};
} // namespace

// The maps of a checkpoint are always COMPLETE: a sharded run folds its pending count-matrix rows first
// (Reduce(): a collective, so every rank must call SaveCheckpoint at the same launch; the ranks then hold
// identical maps and any one of them may write the file).  Loading it under any shard setup is therefore safe:
// later windows add the all-reduced counts to every rank's copy alike.
bool RayTracer::SaveCheckpoint(const char* path)
{
    if (!ok || !mesh) return false;
    Reduce();
    if (!ok) return false;
    const size_t n = (size_t)mesh->triangleCount;
    std::vector<double> maps(2 * n);
    if (!Check(uvrt_read(ctx, UVRT_BUF_SUM, maps.data(), n * 8), "read photon map")) return false;
    if (!Check(uvrt_read(ctx, UVRT_BUF_MAX, maps.data() + n, n * 8), "read max map")) return false;
    CheckpointHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "UVRTCKP2", 8);
    h.triangles = (int32_t)n; h.positions = (int32_t)lightPositions.size(); h.currIterations = currIterations;
    h.photonsPerLight = photonsPerLight; h.launchCounter = launchCounter; h.photonMapSizeTotal = photonMapSizeTotal;
    h.raysTraced = raysTraced; h.seedState = seedState; h.shardCount = shardCount;
    h.lightLength = lightLength; h.lightHeight = lightHeight;
    // write next to the target and rename: a crash mid-write must not destroy the previous checkpoint
    const std::string tmp = std::string(path) + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f) { lastError = std::string("SaveCheckpoint: cannot write ") + tmp; return false; }
        f.write((const char*)&h, sizeof h);
        f.write((const char*)maps.data(), (std::streamsize)(maps.size() * 8));
        f.flush();
        if (!f) { lastError = std::string("SaveCheckpoint: write failed: ") + tmp; std::remove(tmp.c_str()); return false; }
    }
    if (std::rename(tmp.c_str(), path) != 0) { lastError = std::string("SaveCheckpoint: cannot rename to ") + path; std::remove(tmp.c_str()); return false; }
    return true;
}

bool RayTracer::LoadCheckpoint(const char* path)
{
    if (!ok || !mesh) return false;
    const size_t n = (size_t)mesh->triangleCount;
    std::ifstream f(path, std::ios::binary);
    CheckpointHeader h;
    if (!f || !f.read((char*)&h, sizeof h) || memcmp(h.magic, "UVRTCKP2", 8) != 0) {
        lastError = std::string("LoadCheckpoint: not a checkpoint: ") + path;
        return false;
    }
    if ((size_t)h.triangles != n || (size_t)h.positions != lightPositions.size()) {
        lastError = "LoadCheckpoint: checkpoint belongs to another room or route";
        return false;
    }
    // the rays of the remaining launches depend on these: a mismatch would silently mix two different runs
    if (h.photonsPerLight != photonsPerLight || memcmp(&h.lightLength, &lightLength, 4) != 0 || memcmp(&h.lightHeight, &lightHeight, 4) != 0) {
        lastError = "LoadCheckpoint: checkpoint was written with other photon / lamp settings (photons per light " +
                    std::to_string(h.photonsPerLight) + " vs " + std::to_string(photonsPerLight) + ")";
        return false;
    }
    std::vector<double> maps(2 * n);
    if (!f.read((char*)maps.data(), (std::streamsize)(maps.size() * 8))) { lastError = "LoadCheckpoint: truncated file"; return false; }
    ClearBuffers(false);
    if (!Check(uvrt_write(ctx, UVRT_BUF_SUM, maps.data(), n * 8), "write photon map")) return false;
    if (!Check(uvrt_write(ctx, UVRT_BUF_MAX, maps.data() + n, n * 8), "write max map")) return false;
    currIterations = h.currIterations;
    launchCounter = h.launchCounter;
    photonMapSizeTotal = h.photonMapSizeTotal;
    photonMapSize = photonMapSizeTotal > 0x7fffffffLL ? 0x7fffffff : (int)photonMapSizeTotal;
    raysTraced = h.raysTraced;
    seedState = h.seedState;
    windowRows = windowFill = 0;
    startedComputation = true;
    finishedComputation = currIterations >= maxIterations;
    return true;
}

static std::string RoutePath(const char* fileName)
{
    return AssetRoot() + "positions/" + fileName + ".xml";
}

void RayTracer::SaveRoute(char fileName[32])
{
    using uvrt_xml::Element;
    Element root;
    root.name = "route";
    root.add("aantal_fotonen")->text = std::to_string(photonCount);
    root.add("aantal_iteraties")->text = std::to_string(maxIterations);
    root.add("lamp_sterkte")->text = uvrt_xml::fmt_float(lightIntensity);
    root.add("minimale_dosis")->text = uvrt_xml::fmt_float(minDosage);
    root.add("minimale_bestralingssterkte")->text = uvrt_xml::fmt_float(minPower);
    root.add("lamp_lengte")->text = uvrt_xml::fmt_float(lightLength);
    root.add("lamp_hoogte")->text = uvrt_xml::fmt_float(lightHeight);
    Element* route = root.add("route");
    for (size_t i = 0; i < lightPositions.size(); i++) {
        Element* e = route->add("lamp_positie_" + std::to_string(i));
        e->attrs.emplace_back("positie_x", uvrt_xml::fmt_float(lightPositions[i].position.x));
        e->attrs.emplace_back("positie_y", uvrt_xml::fmt_float(lightPositions[i].position.y));
        e->attrs.emplace_back("duration", uvrt_xml::fmt_float(lightPositions[i].duration));
    }
    std::string out;
    uvrt_xml::write(root, out, 0);
    std::ofstream f(RoutePath(fileName), std::ios::binary);
    if (f) f << out;
}

void RayTracer::LoadRoute(char fileName[32])
{
    std::ifstream f(RoutePath(fileName), std::ios::binary);
    if (!f) return;   // like the reference: a missing or broken file leaves the settings alone
    std::stringstream ss;
    ss << f.rdbuf();
    std::unique_ptr<uvrt_xml::Element> root = uvrt_xml::Reader(ss.str()).parse();
    if (!root) return;
    uvrt_xml::Element* e;
    if ((e = root->child("aantal_fotonen"))) e->int_text(&photonCount);
    if ((e = root->child("aantal_iteraties"))) e->int_text(&maxIterations);
    if ((e = root->child("lamp_sterkte"))) e->float_text(&lightIntensity);
    if ((e = root->child("minimale_dosis"))) e->float_text(&minDosage);
    if ((e = root->child("minimale_bestralingssterkte"))) e->float_text(&minPower);
    if ((e = root->child("lamp_lengte"))) e->float_text(&lightLength);
    if ((e = root->child("lamp_hoogte"))) e->float_text(&lightHeight);
    if ((e = root->child("route"))) {
        lightPositions.clear();
        // positions are looked up by consecutive index until one is missing (raytracer.cpp:285-297)
        for (int i = 0;; i++) {
            uvrt_xml::Element* p = e->child("lamp_positie_" + std::to_string(i));
            if (!p) break;
            LightPos lp;
            lp.position = make_float2(0.0f, 0.0f);
            lp.duration = 0.0f;
            p->float_attr("positie_x", &lp.position.x);
            p->float_attr("positie_y", &lp.position.y);
            p->float_attr("duration", &lp.duration);
            lightPositions.push_back(lp);
        }
    }
    UpdatePhotonsPerLight();
}

} // namespace Tmpl8
