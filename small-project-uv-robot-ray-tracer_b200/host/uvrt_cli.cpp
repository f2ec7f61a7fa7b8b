// uvrt_cli.cpp -- headless driver: what MyApp::Init + MyApp::Tick do in the reference
// (myapp.cpp:15-40, 156-175), without the window.  Loads rooms/<room>.glb and
// positions/<route>.xml, runs every iteration, prints the reference's progress line and
// optionally writes the per-triangle dose map (float32, triangle order) to a file.
#include "precomp.h"
#include "../../include/uvrt.h"
#include <fstream>

using namespace Tmpl8;

static void usage()
{
    printf("usage: uvrt_cli [--root DIR] [--room NAME] [--route NAME] [--iterations N] [--photons N]\n"
           "                [--device D] [--maxpower] [--out dose.f32] [--export BASE] [--variant V]\n"
           "                [--device-bvh] [--checkpoint FILE] [--resume FILE]\n"
           "  --export BASE      BASE.dose.f32, BASE.ply (per-vertex heat-map colours), BASE.json\n"
           "  --device-bvh       build the BVH on the GPU (uvrt_build_bvh) instead of on the host cores\n"
           "  --checkpoint FILE  rewrite FILE after every iteration; --resume FILE continues such a run\n"
           "  --json             one JSON line at the end: rays, wall and per-stage device times, Mrays/s\n");
}

int main(int argc, char** argv)
{
    std::string room = "testroomopt", route = "route", out, exportBase, checkpoint, resume;
    bool deviceBvh = false, jsonLine = false;
    int iterations = -1, photons = -1, device = 0, variant = -1;
    bool maxPower = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--root") SetAssetRoot(next());
        else if (a == "--room") room = next();
        else if (a == "--route") route = next();
        else if (a == "--iterations") iterations = atoi(next());
        else if (a == "--photons") photons = atoi(next());
        else if (a == "--device") device = atoi(next());
        else if (a == "--variant") variant = atoi(next());
        else if (a == "--maxpower") maxPower = true;
        else if (a == "--out") out = next();
        else if (a == "--export") exportBase = next();
        else if (a == "--checkpoint") checkpoint = next();
        else if (a == "--resume") resume = next();
        else if (a == "--device-bvh") deviceBvh = true;
        else if (a == "--json") jsonLine = true;
        else { usage(); return a == "--help" ? 0 : 2; }
    }
    Mesh mesh;
    mesh.buildBvhOnLoad = !deviceBvh;
    strncpy(mesh.modelFile, room.c_str(), 31);
    mesh.LoadMesh();
    if (!mesh.loadedMesh) { fprintf(stderr, "cannot load room: %s\n", mesh.lastError.c_str()); return 1; }

    RayTracer rayTracer;
    rayTracer.device = device;
    rayTracer.saveRouteOnReset = false;
    strncpy(rayTracer.defaultRouteFile, route.c_str(), 31);
    rayTracer.Init(&mesh);
    if (!rayTracer.ok) { fprintf(stderr, "%s\n", rayTracer.lastError.c_str()); return 1; }
    if (iterations > 0) rayTracer.maxIterations = iterations;
    if (photons > 0) { rayTracer.photonCount = photons; rayTracer.UpdatePhotonsPerLight(); }
    if (variant >= 0) uvrt_set_option(rayTracer.ctx, "extend_variant", variant);
    if (maxPower) rayTracer.viewMode = maxpower;

    if (jsonLine) uvrt_set_option(rayTracer.ctx, "stage_timing", 1);
    rayTracer.ResetDosageMap();
    if (!resume.empty()) {
        if (!rayTracer.LoadCheckpoint(resume.c_str())) { fprintf(stderr, "%s\n", rayTracer.lastError.c_str()); return 1; }
        rayTracer.Shade();     // a checkpoint of a finished run skips the loop below: the dose map must still be formed
    }
    while (rayTracer.ok) {
        rayTracer.finishedComputation = rayTracer.currIterations >= rayTracer.maxIterations;
        if (rayTracer.finishedComputation) break;
        rayTracer.ComputeDosageMap();
        rayTracer.Shade();
        if (rayTracer.viewMode == texture) rayTracer.viewMode = dosage;
        rayTracer.currIterations++;
        rayTracer.progress = 100.0f * (float)rayTracer.currIterations / (float)rayTracer.maxIterations;
        uvrt_sync(rayTracer.ctx);
        float time = rayTracer.timerClock.elapsed();
        rayTracer.compTime += time;
        std::cout << "Progress: " << rayTracer.progress << "% photon count: " << rayTracer.photonMapSize
                  << " delta time: " << time * 1000.0f << " total time: " << rayTracer.compTime * 1000.0f << std::endl;
        if (!checkpoint.empty() && !rayTracer.SaveCheckpoint(checkpoint.c_str())) break;
        rayTracer.timerClock.reset();
    }
    if (!rayTracer.ok) { fprintf(stderr, "%s\n", rayTracer.lastError.c_str()); return 1; }
    const float* dose = rayTracer.ReadDosageMap();
    double sum = 0;
    float mx = 0;
    for (int i = 0; i < mesh.triangleCount; i++) { sum += dose[i]; if (dose[i] > mx) mx = dose[i]; }
    printf("triangles %d  rays %lld  mean %s %.6g  max %.6g  (%.1f Mrays/s wall)\n", mesh.triangleCount,
           (long long)rayTracer.RaysTraced(), maxPower ? "irradiance" : "dose", sum / mesh.triangleCount, mx,
           rayTracer.RaysTraced() / (rayTracer.compTime * 1e6));
    if (jsonLine) {
        static const char* names[] = {"generate", "extend", "accumulate", "shade", "color", "reset", "bin"};
        char dev[128] = "";
        int sms = 0, ccMajor = 0, ccMinor = 0;
        uvrt_device_info(rayTracer.ctx, dev, sizeof dev, &sms, &ccMajor, &ccMinor);
        printf("{\"room\": \"%s\", \"route\": \"%s\", \"device\": \"%s\", \"triangles\": %d, \"positions\": %d, \"iterations\": %d, "
               "\"rays\": %lld, \"wall_ms\": %.3f, \"mrays_s_wall\": %.1f, \"kernel_launches\": %lld, \"stage_ms\": {",
               room.c_str(), route.c_str(), dev, mesh.triangleCount, (int)rayTracer.lightPositions.size(), rayTracer.currIterations,
               (long long)rayTracer.RaysTraced(), rayTracer.compTime * 1000.0f, rayTracer.RaysTraced() / (rayTracer.compTime * 1e6),
               (long long)uvrt_launch_count(rayTracer.ctx));
        for (int k = 0; k < 7; k++) {
            double ms = 0;
            int64_t n = 0;
            uvrt_stage_time(rayTracer.ctx, (uvrt_stage)k, &ms, &n);
            printf("%s\"%s\": %.3f", k ? ", " : "", names[k], ms);
        }
        printf("}, \"mean\": %.6g, \"max\": %.6g}\n", sum / mesh.triangleCount, mx);
    }
    if (!exportBase.empty() && !rayTracer.SaveDosageMap(exportBase.c_str())) { fprintf(stderr, "%s\n", rayTracer.lastError.c_str()); return 1; }
    if (!out.empty()) {
        std::ofstream f(out, std::ios::binary);
        f.write((const char*)dose, sizeof(float) * (size_t)mesh.triangleCount);
    }
    return 0;
}
