// raytracer.cpp -- RayTracer on the libuvrt C ABI.
// Host sequencing follows /root/reference/raytracer.cpp:12-300: which stage runs when, the
// scalar arguments each stage gets (photons per light, scaled power, dose divisor) and the
// route-file schema.  The OpenCL plumbing (Kernel, Buffer, SetArgument) is replaced by uvrt_* calls.
#include "precomp.h"
#include "xml_min.h"
#include "../../include/uvrt.h"
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
    seedQueue.clear();
    seedQueueHead = 0;
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
    // SEED of every launch of the remaining passes in one device call, unless the queue built by
    // an earlier pass still matches the route
    const size_t L = lightPositions.size();
    std::vector<float> pos(3 * L);
    for (size_t i = 0; i < L; i++) {
        pos[3 * i + 0] = lightPositions[i].position.x;
        pos[3 * i + 1] = mesh->floorHeight + lightHeight;
        pos[3 * i + 2] = lightPositions[i].position.y;
    }
    bool queued = seedQueue.size() - seedQueueHead >= L &&
                  memcmp(&seedQueuePos[3 * seedQueueHead], pos.data(), sizeof(float) * 3 * L) == 0;
    if (!queued) {
        int passes = maxIterations - currIterations;
        if (passes < 1) passes = 1;
        std::vector<float> all(3 * L * (size_t)passes);
        for (int p = 0; p < passes; p++) memcpy(&all[3 * L * p], pos.data(), sizeof(float) * 3 * L);
        std::vector<uint32_t> seeds(L * (size_t)passes + 1);
        if (!Check(uvrt_seed_chain(ctx, all.data(), (int)(L * passes), lightLength, seedState, seeds.data()), "seed_chain")) return;
        seedQueue.assign(seeds.begin() + 1, seeds.end());
        seedQueuePos.swap(all);
        seedQueueHead = 0;
    }
    for (LightPos& lightPosition : lightPositions) {
        ComputeSingleLightDosageMap(lightPosition, photonsPerLight, mesh->triangleCount);
    }
}

// Which rank traces launch k of a run over a route of L positions.  Plain round-robin (k mod N) would
// hand a rank the same few positions in every pass whenever gcd(L, N) > 1 (L = 12, N = 8: three
// positions per rank), and positions differ in cost by up to 1.5x (24.9 - 36.6 node visits per ray on
// lange_route), so the slowest rank would set the pace.  The deal is therefore rotated by s ranks per
// pass, with the smallest s that makes L + s coprime to N: every position then visits every rank.
int RayTracer::ShardOwner(long long launch, int L, int N)
{
    if (N <= 1) return 0;
    if (L < 1) L = 1;
    auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
    int s = 0;
    while (gcd((long long)L + s, N) != 1) s++;
    const long long pass = launch / L;
    return (int)((launch + (long long)s * pass) % N);
}

uint32_t RayTracer::SeedAfter(const float3& lp)
{
    if (seedQueueHead < seedQueue.size()) {
        const float* q = &seedQueuePos[3 * seedQueueHead];
        if (!memcmp(&q[0], &lp.x, 4) && !memcmp(&q[1], &lp.y, 4) && !memcmp(&q[2], &lp.z, 4))
            return seedQueue[seedQueueHead++];
        seedQueue.clear();   // the route changed under us: fall back to one launch at a time
        seedQueueHead = 0;
    }
    float pos[3] = {lp.x, lp.y, lp.z};
    uint32_t seeds[2] = {seedState, seedState};
    Check(uvrt_seed_chain(ctx, pos, 1, lightLength, seedState, seeds), "seed_chain");
    return seeds[1];
}

// photonsPerLight and triangleCount are parameters because CalibratePower() uses its own values.
// (triangleCount only sized the reference's accumulate launch; the backend knows the scene size.)
void RayTracer::ComputeSingleLightDosageMap(LightPos lightPos, int photonsPerLight, int /*triangleCount*/)
{
    if (!ok) return;
    float3 lightposition = make_float3(lightPos.position.x, mesh->floorHeight + lightHeight, lightPos.position.y);
    const bool mine = shardCount <= 1 || ShardOwner(launchCounter, (int)lightPositions.size(), shardCount) == shardRank;
    if (mine) {
        if (!Check(uvrt_trace(ctx, lightposition.x, lightposition.y, lightposition.z, lightLength, lightPos.duration, 0,
                              photonsPerLight, seedState),
                   "trace"))
            return;
        raysTraced += photonsPerLight;
    }
    seedState = SeedAfter(lightposition);
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
    Check(uvrt_reduce(ctx), "reduce");
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
    ClearBuffers(true);
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
    char magic[8];            // "UVRTCKP1"
    int32_t triangles, positions, currIterations, photonsPerLight;
    int64_t launchCounter, photonMapSizeTotal, raysTraced;
    uint32_t seedState, pad;
};
} // namespace

bool RayTracer::SaveCheckpoint(const char* path)
{
    if (!ok || !mesh) return false;
    const size_t n = (size_t)mesh->triangleCount;
    std::vector<double> maps(2 * n);
    if (!Check(uvrt_read(ctx, UVRT_BUF_SUM, maps.data(), n * 8), "read photon map")) return false;
    if (!Check(uvrt_read(ctx, UVRT_BUF_MAX, maps.data() + n, n * 8), "read max map")) return false;
    CheckpointHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "UVRTCKP1", 8);
    h.triangles = (int32_t)n; h.positions = (int32_t)lightPositions.size(); h.currIterations = currIterations;
    h.photonsPerLight = photonsPerLight; h.launchCounter = launchCounter; h.photonMapSizeTotal = photonMapSizeTotal;
    h.raysTraced = raysTraced; h.seedState = seedState;
    std::ofstream f(path, std::ios::binary);
    if (!f) { lastError = std::string("SaveCheckpoint: cannot write ") + path; return false; }
    f.write((const char*)&h, sizeof h);
    f.write((const char*)maps.data(), (std::streamsize)(maps.size() * 8));
    return (bool)f;
}

bool RayTracer::LoadCheckpoint(const char* path)
{
    if (!ok || !mesh) return false;
    const size_t n = (size_t)mesh->triangleCount;
    std::ifstream f(path, std::ios::binary);
    CheckpointHeader h;
    if (!f || !f.read((char*)&h, sizeof h) || memcmp(h.magic, "UVRTCKP1", 8) != 0) {
        lastError = std::string("LoadCheckpoint: not a checkpoint: ") + path;
        return false;
    }
    if ((size_t)h.triangles != n || (size_t)h.positions != lightPositions.size()) {
        lastError = "LoadCheckpoint: checkpoint belongs to another room or route";
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
    seedQueue.clear();
    seedQueueHead = 0;
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
