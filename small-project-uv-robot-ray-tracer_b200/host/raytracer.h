// raytracer.h -- the simulation driver with the reference's public surface
// (raytracer.h:5-59 of the reference: same methods, same tunable fields), re-hosted on the
// libuvrt C ABI (include/uvrt.h) instead of the OpenCL Kernel/Buffer wrapper.
#pragma once

struct uvrt_ctx;

namespace Tmpl8 {

struct LightPos {
    float2 position;
    float duration;
};

enum ViewMode { dosage, maxpower, texture };

class RayTracer {
public:
    RayTracer() = default;
    ~RayTracer();
    RayTracer(const RayTracer&) = delete;
    RayTracer& operator=(const RayTracer&) = delete;

    // ---- the reference's interface (raytracer.h:16-26) ----
    void Init(Mesh* mesh);
    void UpdatePhotonsPerLight();
    void ComputeDosageMap();
    void ComputeSingleLightDosageMap(LightPos lightPos, int photonsPerLight, int triangleCount);
    void Shade();
    void ResetDosageMap();
    void ClearBuffers(bool resetColor);
    void AddLamp();
    void CalibratePower(float measurePower, float measureHeight, float measureDist);
    void SaveRoute(char fileName[32]);
    void LoadRoute(char fileName[32]);

    float lightLength = 1.0f;
    float lightHeight = 0.8f;
    int maxPhotonCount = (1 << 26);
    int photonCount = (1 << 25);
    int maxIterations = 10;
    int currIterations = 0; // The number of computed iterations
    float lightIntensity = 450;
    float minDosage = 100, minPower = 1500;
    char defaultRouteFile[32] = "route";
    char newRouteFile[32] = "new_route";

    Mesh* mesh = 0;
    float* dosageMap = new float[2];
    std::vector<LightPos> lightPositions;
    int photonsPerLight = 0; // The number of photons per light of a single iteration
    float compTime = 0;
    float progressTextTimer = 0;
    float progress = 0;
    Timer timerClock;
    bool finishedComputation = true;
    ViewMode viewMode = texture;
    bool thresholdView = false;
    bool startedComputation = false;
    float calibratedPower = 0;
    int photonMapSize = 0;

    // ---- additions (no counterpart in the reference) ----
    int device = 0;              // CUDA device used by Init()
    uvrt_ctx* ctx = 0;           // backend context (replaces the Kernel*/Buffer* members)
    bool ok = true;              // false after a backend failure; see lastError
    std::string lastError;
    bool saveRouteOnReset = true; // the reference rewrites positions/<route>.xml on every run start
    // Device-side SEED of generate.cl:6 as seen by the next launch: starts at 0
    // and, like the reference's program-scope variable, survives ResetDosageMap().
    uint32_t seedState = 0;
    // Work sharing between GPUs.  Every launch of a run is cut into shardParts ray ranges (1 = whole launches);
    // unit u = launch * parts + part is traced by rank ShardOwner(u, positions * parts, shardCount) -- round-robin,
    // rotated once per pass so that every rank sees every lamp position.  A range passes its first global ray
    // id, so its rays are those of the unsplit launch.  Every rank advances seedState and photonMapSize for every
    // launch.  Sharded runs collect integer counts in a count matrix (one row per launch, uvrt.h) that Reduce()
    // sums over the ranks and folds in launch order: maps bit-identical to the single-GPU run for any split.
    static int ShardOwner(long long unit, int unitsPerPass, int ranks);
    int shardRank = 0, shardCount = 1;
    int shardParts = 0;          // 0: chosen per run (PlanShards / AutoParts)
    int AutoParts() const;
    // Cost-aware deal (default for shardParts == 0): lamp positions differ in cost by up to 1.4x, so ResetDosageMap
    // probes every position once per route (uvrt_probe_cost: node visits and triangle tests per ray of a small,
    // deterministic sample -- identical on every rank) and deals the WHOLE launches of the run longest-first to the
    // least loaded rank.  Falls back to the rotation of ShardOwner when switched off or when the run outgrows the plan.
    bool costAwareSharding = true;
    static void PlanShardsLPT(const double* launchCost, int launches, int ranks, int* ownerOut);
    long long launchCounter = 0;
    long long photonMapSizeTotal = 0;   // photonMapSize without the int overflow (see ComputeSingleLightDosageMap)
    // SEED after a launch at lightposition (generate.cl:13-39 for work-item 0), computed on the host: the chain
    // of a whole run needs no device round trip.  Bit-identical to the device (uvrt_seed_chain; GPU test).
    static uint32_t SeedAfterLaunch(float lx, float ly, float lz, float lightLength, uint32_t seedIn);
    // Sharded runs: sums the pending rows of the count matrix over the ranks (one ncclAllReduce; needs
    // uvrt_comm_init on ctx) and folds them into the photon / max maps.  Call after the last ComputeDosageMap()
    // and before Shade() / ReadDosageMap(); long runs also fold whenever the matrix window is full.
    void Reduce();
    // Copies the per-triangle dose (after Shade) into dosageMap, resized to triangleCount floats.
    const float* ReadDosageMap();
    int64_t RaysTraced() const { return raysTraced; }
    // Result export (the reference only ever shows the map in its GL window, myapp.cpp:180-205):
    //   <base>.dose.f32  per-triangle dose / irradiance, float32, triangle order (after Shade)
    //   <base>.ply       binary little-endian PLY, 3 vertices per triangle, per-vertex uchar RGB from
    //                    the colour buffer of dosageToColor (shade.cl:43-71), clamped to [0, 1]
    //   <base>.json      sidecar: room, parameters, route, rays, iterations, SEED
    bool SaveDosageMap(const char* basePath);
    // Checkpoint / resume of a run: photon map and max map (f64), iteration and launch counters, the
    // photon total and the device SEED chain.  A resumed run continues with the rays the uninterrupted
    // run would have traced, so the final maps are bit-identical.
    bool SaveCheckpoint(const char* path);
    bool LoadCheckpoint(const char* path);

private:
    bool Check(int rc, const char* what);
    void UploadScene();
    void PlanShards();
    long long MaxWindowRows() const;
    std::vector<int> shardPlan;            // owner of launch k of the run being traced; empty: ShardOwner
    std::vector<float> planKey;            // what positionCost was probed for
    std::vector<double> positionCost;
    // count-matrix window of a sharded run
    void BeginWindow();
    void FoldWindow();
    int windowRows = 0, windowFill = 0;
    std::vector<float> windowDurations;
    int dosageMapSize = 2;
    int64_t raysTraced = 0;
};

} // namespace Tmpl8
