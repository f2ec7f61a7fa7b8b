// uvrt_capi.cu -- the C ABI of include/uvrt.h over the kernels of uvrt_kernels.cuh.
//
// Replaces the reference's OpenCL wrapper layer (template/template.cpp:1039-1573) for the one
// path RayTracer drives (raytracer.cpp:12-143).  No torch, no C++ types in the interface.
#include "../../include/uvrt.h"
#include "uvrt_kernels.cuh"
#include "uvrt_bvh_build.cuh"
#include "uvrt_scene_prep.cuh"
#include "uvrt_fast.cuh"
#ifdef UVRT_EXPERIMENTS
#include "uvrt_fast_refill.cuh"
#endif

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler injects itself
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

using namespace uvrt;

// ---- NCCL, bound at run time so that libuvrt.so loads on machines without it -------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0 };
enum { kNcclInt32 = 2, kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclMax = 2 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
    bool ok = false;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.ok) return true;
    if (g_nccl.lib == nullptr) {
        const char* env = getenv("UVRT_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) {
            g_nccl.why = "libnccl.so.2 not found (set UVRT_NCCL_LIB)";
            return false;
        }
    }
#define UVRT_SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                         \
    if (!g_nccl.field) { g_nccl.why = std::string("missing symbol ") + name; return false; }
    UVRT_SYM(GetUniqueId, "ncclGetUniqueId")
    UVRT_SYM(CommInitRank, "ncclCommInitRank")
    UVRT_SYM(CommDestroy, "ncclCommDestroy")
    UVRT_SYM(AllReduce, "ncclAllReduce")
    UVRT_SYM(GroupStart, "ncclGroupStart")
    UVRT_SYM(GroupEnd, "ncclGroupEnd")
    UVRT_SYM(GetErrorString, "ncclGetErrorString")
#undef UVRT_SYM
    g_nccl.ok = true;
    return true;
}

thread_local std::string g_create_error;
} // namespace

struct TimedLaunch {
    cudaEvent_t start, stop;
    int stage;
};

// Everything that belongs to one launch's ray queue.
struct RaySlot {
    float4* dRays = nullptr;
    long long rayCap = 0;
    unsigned int* dBinCount = nullptr;
    unsigned int* dBinStart = nullptr;
    unsigned int* dBinBlock = nullptr;   // totals of the scan blocks
    int binCap = 0, binUsed = 0;
    long long countedRays = -1;          // rays whose bin slots k_generate already took (-1: none)
    uint2* dKeyRank = nullptr;
    uint32_t* dPerm = nullptr;
    long long permCap = 0;
    long long permRays = -1;             // dPerm already holds the permutation of this many rays (pipelined trace)
    float binY0 = 0.0f, binLen = 1.0f;   // lamp extent of the rays in the buffer
    bool binExtentKnown = false;
    cudaEvent_t genDone = nullptr, freeEv = nullptr;
    bool inFlight = false;               // freeEv has been recorded at least once
#ifdef UVRT_EXPERIMENTS
    // lists of the refill extend ("fast_cfg" 2, uvrt_fast_refill.cuh): counters, winners to verify, rays to re-trace
    RefillCtl* dRefillCtl = nullptr;
    uint4* dVerify = nullptr;
    uint32_t* dRetry = nullptr;
    long long refillCap = 0;
#endif
};

struct uvrt_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;

    // scene (device layout, DESIGN.md "Data layout in HBM")
    int nTris = 0;
    int nPairs = 0, nLeaves = 0, depth = 0;
    long long nSlots = 0;                // leaf triangle slots (= nTris for a tree that covers the mesh)
    uint32_t rootRef = 0;
    int sceneTame = 0;
    int sceneNested = 0;                 // every child box inside its parent's box (checked by the repack)
    // certified fast extend (uvrt_fast.cuh): 32-byte quantised node pairs, grid, fallback counters
    uint4* dQPairs = nullptr;
    size_t qpairCap = 0;                 // capacity in node pairs
    FastGrid* dFastGrid = nullptr;
    FastStats* dFastStats = nullptr;
    int fastCheck = 0;                   // option "fast_check": trace every ray twice and count certified mismatches
    int fastCfg = 0;                     // option "fast_cfg": register budget of the fast kernel (experiments: 2 = refill kernel)
#ifdef UVRT_EXPERIMENTS
    FastGrid fastGridHost{};             // copy of *dFastGrid (kernel argument of the refill extend)
    uint4* dQPairsSel = nullptr;         // layout 1 of the quantised pairs (refill extend), built on first use
    size_t qpairSelCap = 0;
    bool qpairSelValid = false;
    int refillChunk = 64;                // option "refill_chunk": rays a warp takes from the queue at a time
    int refillBlocksPerSm = 0;           // resident blocks of k_extend_fast_refill (occupancy query, cached)
#endif
    cudaTextureObject_t pairsTex = 0;   // the same buffer as a 1-D float4 texture ("fetch_mode" experiment)
    int fetchMode = 3;   // 3: rays, permutation kept out of L1 (ld.global.L1::no_allocate); 0: plain loads; 1/2: texture experiments; 4: + evict_last nodes
    float4* dPairs = nullptr;   // nPairs x 4 float4
    float4* dWtris = nullptr;   // nTris  x 4 float4 (leaf order: v0+tag, edge1, edge2, pad)
    float4* dVerts = nullptr;   // nTris  x 4 float4 (reference order and layout)
    float4* dVertsSpare = nullptr;   // upload target while the previous scene is still valid; swapped in on success
    size_t vertsCap = 0, vertsSpareCap = 0;   // capacities in triangles
    cudaEvent_t vertsEv = nullptr;
    // per-triangle state (raytracer.h:54-55)
    int* dCounts = nullptr;
    double* dSum = nullptr;
    double* dMax = nullptr;
    float* dDose = nullptr;
    float* dColor = nullptr;
    // rays: two slots, so that generate of the next launch can overlap extend of the current one
    RaySlot slots[2];
    int slot = 0;
    RaySlot& rs() { return slots[slot]; }
    cudaStream_t genStream = nullptr;    // generate (+ bin count) of pipelined uvrt_trace calls
    int pipeline = 1;
    // uvrt_trace with "overlap_extend": the extend kernels of consecutive launches run on two streams
    // (slot 0 / slot 1) with a count buffer each, so the tail of extend k overlaps the head of extend
    // k+1; the accumulates stay on the main stream, in launch order
    int overlapExtend = 1;
    cudaStream_t extStream[2] = {nullptr, nullptr};
    cudaStream_t accStream = nullptr;    // accumulates of overlapped traces (high priority: tiny kernels, they gate the next extend)
    cudaEvent_t extDone[2] = {nullptr, nullptr}, accDone[2] = {nullptr, nullptr}, forkEv = nullptr;
    bool extUsed[2] = {false, false}, accUsed[2] = {false, false};
    bool mainForeign = true;             // something other than an overlapped trace touched the main stream
    bool countsDirty = false;            // UVRT_BUF_COUNTS holds counts no accumulate has consumed yet
    int* dCountsAlt = nullptr;           // count buffer of slot 1
    cudaStream_t xStream = nullptr;      // stream / count buffer of the extend being launched
    int* xCounts = nullptr;
    long long lastRays = 0;
    unsigned int* dQueue = nullptr;      // persistent-kernel work counter
    unsigned int* dSmCursor = nullptr;   // per-SM chunk cursors of the SM-affine extend (fetch_mode 5), 2 x 1024
    uint32_t* dSeeds = nullptr;
    float* dSeedPos = nullptr;
    int seedCap = 0;
    bool poolTuned = false;
    void* dFlush = nullptr;              // L2 flush scratch
    // ray binning (counting sort by direction / origin cell); the tables live in the ray slots
    int binRays = 1, binY = 16, binT = 32, binP = 128;   // 65,536 bins: best of the sweeps
    int64_t uploadBytes = 0;
    // pinned staging for the scene upload, and scratch that survives between uploads
    void* hStage = nullptr;
    size_t hStageBytes = 0;
    size_t pairCap = 0, wtriCap = 0;     // device capacities in bytes
    std::vector<int32_t> upId;
    std::vector<uint32_t> upOrder;
    // device-side repack (uvrt_scene_prep.cuh): raw reference arrays + per-node scratch
    int hostRepack = 0;                  // option "host_repack": 1 = repack on the host cores instead
    void* dRawNodes = nullptr;
    uint32_t* dRawIdx = nullptr;
    unsigned long long* dPrepQueue = nullptr;
    uint32_t* dPrepNode = nullptr;       // parent, arrive, subInner, subSlots: 4 x nodeCap
    void* dPrepStatus = nullptr;
    void* hPrepStatus = nullptr;         // pinned
    size_t prepNodeCap = 0, prepTriCap = 0;
    int prepBlocks = 0;

    // options
    int extendVariant = -1;   // -1: default
    int stageTiming = 0;
    int histMode = 0;
    int blocksPerSm = 0;      // 0: default for the variant
    int carveout = -1;        // experiment: preferred shared-memory carveout (percent) of the extend kernel, -1 = driver default
    int genericOctant = 0;    // experiment: the one-thread-per-ray kernel without octant specialisation
    int chunk = 128;          // rays per warp of the chunk-persistent kernel
    int simpleCfg = 1;        // 128 threads, <= 40 registers (48 resident warps per SM): fastest in the sweep
    int refill = 24;          // persistent kernels: refill when fewer lanes than this are busy

    // count matrix (uvrt_matrix_*): one int32 row of nTris counters per launch of a run / window
    // Two buffers alternate from window to window, so that the all-reduce + fold of window k (main stream) run next
    // to the extends of window k+1 (extend streams) instead of draining the pipeline at every window boundary.
    int* dMatrix = nullptr;              // the current window's buffer (= matrixBuf[matrixSel])
    int* matrixBuf[2] = {nullptr, nullptr};
    size_t matrixCap[2] = {0, 0};        // capacities in int32 elements
    int matrixSel = 1;
    int matrixRows = 0;
    float* durBuf[2] = {nullptr, nullptr};
    int durCap[2] = {0, 0};
    cudaEvent_t matrixReady[2] = {nullptr, nullptr}, matrixFolded[2] = {nullptr, nullptr};
    bool matrixFoldedUsed[2] = {false, false};

    // measurement
    int timeline = 0;                    // option "timeline": host-side call log + device times of every stage launch
    cudaEvent_t timelineOrigin = nullptr;
    double timelineHost0 = 0;
    struct ApiCall { const char* name; double t0, t1; };
    std::vector<ApiCall> apiLog;
    std::vector<double> timedHost;       // host time at which timed[i] was enqueued
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> freeEvents;
    cudaEvent_t marks[16] = {};
    int64_t launches = 0;

    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, nRanks = 1;
};

namespace {

int fail(uvrt_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? UVRT_ERR_NO_MEMORY : UVRT_ERR_CUDA, \
                        "%s failed: %s", #call, cudaGetErrorString(e_));                      \
    } while (0)

#define CK_LAUNCH(name)                                                                    \
    do {                                                                                   \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess) return fail(ctx, UVRT_ERR_CUDA, "%s launch failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

struct Bind {
    explicit Bind(uvrt_ctx* c)
    {
        cudaGetDevice(&prev);
        if (prev != c->device) cudaSetDevice(c->device);
        dev = c->device;
        c->mainForeign = true;   // every entry point but the overlapped uvrt_trace clears nothing: see uvrt_trace
    }
    ~Bind() { if (prev != dev && prev >= 0) cudaSetDevice(prev); }
    int prev = -1, dev = -1;
};

cudaEvent_t get_event(uvrt_ctx* ctx)
{
    if (!ctx->freeEvents.empty()) {
        cudaEvent_t e = ctx->freeEvents.back();
        ctx->freeEvents.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

double host_now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

const char* const kStageNames[UVRT_STAGE_COUNT] = {"generate", "extend", "accumulate", "computeDosage", "dosageToColor", "reset", "bin"};

// Brackets the launches of one stage: an NVTX range (visible in nsys / ncu --nvtx) and, with
// "stage_timing" or "timeline" on, a pair of CUDA events on the launching stream.
struct StageTimer {
    uvrt_ctx* ctx;
    TimedLaunch t{};
    bool on;
    cudaStream_t stream;
    StageTimer(uvrt_ctx* c, int stage, cudaStream_t st = nullptr) : ctx(c), on(c->stageTiming != 0 || c->timeline != 0), stream(st ? st : c->stream)
    {
        nvtxRangePushA(stage >= 0 && stage < UVRT_STAGE_COUNT ? kStageNames[stage] : "stage");
        if (!on) return;
        t.stage = stage;
        t.start = get_event(c);
        t.stop = get_event(c);
        cudaEventRecord(t.start, stream);
    }
    ~StageTimer()
    {
        nvtxRangePop();
        if (!on) return;
        cudaEventRecord(t.stop, stream);
        ctx->timed.push_back(t);
        if (ctx->timeline) ctx->timedHost.push_back(host_now_us() - ctx->timelineHost0);
    }
};

// UVRT_DEBUG_ERRORS=1: report a CUDA error that is still pending when an entry point returns (an ignored return
// value somewhere inside it), instead of letting the next unrelated launch check trip over it.
void debug_pending_error(const char* where)
{
    static const bool on = getenv("UVRT_DEBUG_ERRORS") != nullptr;
    if (!on) return;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) fprintf(stderr, "[uvrt debug] CUDA error pending after %s: %s\n", where, cudaGetErrorString(e));
}

// One per C-ABI entry point that does work: NVTX range + (with "timeline") a host-side record of the call.
struct ApiScope {
    uvrt_ctx* ctx;
    const char* name;
    double t0 = 0;
    ApiScope(uvrt_ctx* c, const char* n) : ctx(c), name(n)
    {
        nvtxRangePushA(n);
        if (c && c->timeline) t0 = host_now_us() - c->timelineHost0;
    }
    ~ApiScope()
    {
        nvtxRangePop();
        if (ctx && ctx->timeline) ctx->apiLog.push_back({name, t0, host_now_us() - ctx->timelineHost0});
        debug_pending_error(name);
    }
};

template <typename T>
int dev_alloc(uvrt_ctx* ctx, T** p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return UVRT_OK;
    CK(cudaMalloc((void**)p, count * sizeof(T)));
    return UVRT_OK;
}

int ensure_rays(uvrt_ctx* ctx, long long nRays)
{
    if (nRays <= ctx->rs().rayCap) return UVRT_OK;
    // raytracer.cpp:137 sizes the ray buffer by photonCount; here it follows the largest launch
    long long cap = std::max<long long>(nRays, 1 << 16);
    if (ctx->rs().dRays) { cudaFree(ctx->rs().dRays); ctx->rs().dRays = nullptr; ctx->rs().rayCap = 0; }
    CK(cudaMalloc((void**)&ctx->rs().dRays, (size_t)cap * 32));
    ctx->rs().rayCap = cap;
    return UVRT_OK;
}

struct HostTri { float v[16]; };
struct HostNode { float mn[3]; uint32_t leftFirst; float mx[3]; uint32_t triCount; };

bool coord_tame(float v)
{
    float a = std::fabs(v);
    return a == 0.0f || (a >= 6.6174449e-24f && a <= 1048576.0f);   // 0 or [2^-77, 2^20], see ray_is_tame()
}

int check_buffer(uvrt_ctx* ctx, uvrt_buffer what, void** ptr, size_t* bytes)
{
    size_t n = (size_t)ctx->nTris;
    switch (what) {
    case UVRT_BUF_RAYS: *ptr = ctx->rs().dRays; *bytes = (size_t)ctx->rs().rayCap * 32; return UVRT_OK;
    case UVRT_BUF_COUNTS: *ptr = ctx->dCounts; *bytes = n * 4; return UVRT_OK;
    case UVRT_BUF_SUM: *ptr = ctx->dSum; *bytes = n * 8; return UVRT_OK;
    case UVRT_BUF_MAX: *ptr = ctx->dMax; *bytes = n * 8; return UVRT_OK;
    case UVRT_BUF_DOSE: *ptr = ctx->dDose; *bytes = n * 4; return UVRT_OK;
    case UVRT_BUF_COLOR: *ptr = ctx->dColor; *bytes = n * 36; return UVRT_OK;
    case UVRT_BUF_PAIRS: *ptr = ctx->dPairs; *bytes = (size_t)std::max(ctx->nPairs, 1) * 64; return UVRT_OK;
    case UVRT_BUF_WTRIS: *ptr = ctx->dWtris; *bytes = (size_t)ctx->nSlots * 64; return UVRT_OK;
    case UVRT_BUF_MATRIX: *ptr = ctx->dMatrix; *bytes = (size_t)ctx->matrixRows * n * 4; return UVRT_OK;
    }
    return fail(ctx, UVRT_ERR_INVALID, "unknown buffer id %d", (int)what);
}

constexpr long long kMinRaysForBinning = 65536;   // a few thousand rays are not worth the extra launches
constexpr int kMinPairsForBinning = 32;           // nor is a tree of a few nodes (CalibratePower's two triangles)
inline bool wants_binning(const uvrt_ctx* ctx, long long nRays)
{
    return ctx->binRays && nRays >= kMinRaysForBinning && ctx->nPairs >= kMinPairsForBinning;
}

constexpr int kDefaultVariant = 2;   // one thread per ray, one-step shared-reciprocal division (proven exact)
// "extend_variant" -1 (the default): the certified fast kernel (uvrt_fast.cuh) wherever it can serve the scene -- tame,
// nested boxes, a tree worth the name -- else the exact kernel.  Both give the same bits (the fast kernel re-traces
// in reference order what it cannot certify); the fast one is 7-11 % quicker on the room and 1.4-1.7x on the 1 M /
// 10 M-triangle soups, with 0 mismatching rays over every benchmarked configuration (profiles/r2_fast_extend.md).
constexpr int kFastMinPairs = 32;
inline bool fast_usable(const uvrt_ctx* ctx) { return ctx->nPairs > 0 && ctx->sceneTame && ctx->sceneNested && ctx->dQPairs != nullptr; }
inline int default_variant(const uvrt_ctx* ctx) { return fast_usable(ctx) && ctx->nPairs >= kFastMinPairs ? 50 : kDefaultVariant; }

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// ---- extend dispatch ---------------------------------------------------------------------------
// Variant ids (uvrt_set_option "extend_variant"):
//   0  simple / IEEE division            (literal restatement)
//   1  simple / Markstein two-step
//   2  simple / Markstein one-step
//   10 + 3*k + d  persistent, K = {1, 2, 4, 8, inf}[k], d = {IEEE, M2, M1}
constexpr int kStack = 64;


template <int DIV, int THREADS, int MINB>
void launch_simple_cfg(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    k_extend_simple<DIV, kStack, THREADS, MINB><<<grid_for(nRays, THREADS), THREADS, 0, ctx->xStream>>>(
        ctx->xCounts, ctx->dWtris, ctx->rs().dRays, ctx->dPairs, ctx->rootRef, nRays, ctx->sceneTame, perm, 0, ctx->genericOctant);
}

#ifdef UVRT_EXPERIMENTS
template <int K, int CH>
void launch_chunk_kc(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    constexpr int THREADS = 128;
    const long long warps = (nRays + CH - 1) / CH;
    k_extend_chunk<DIV_MARKSTEIN1, kStack, K, CH, THREADS, 10><<<grid_for(warps * 32, THREADS), THREADS, 0, ctx->xStream>>>(
        ctx->xCounts, ctx->dWtris, ctx->rs().dRays, ctx->dPairs, ctx->rootRef, (uint32_t)nRays, ctx->sceneTame, perm, ctx->refill);
}

template <int K>
void launch_chunk_k(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    if (ctx->chunk <= 64) launch_chunk_kc<K, 64>(ctx, nRays, perm);
    else if (ctx->chunk <= 128) launch_chunk_kc<K, 128>(ctx, nRays, perm);
    else if (ctx->chunk <= 256) launch_chunk_kc<K, 256>(ctx, nRays, perm);
    else launch_chunk_kc<K, 1024>(ctx, nRays, perm);
}

#endif

template <int DIV>
void launch_simple(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    // "simple_cfg": block size / resident-blocks hint (register cap) of the one-thread-per-ray kernel
    switch (ctx->simpleCfg) {
    case 1: launch_simple_cfg<DIV, 128, 12>(ctx, nRays, perm); break;   // <= 40 registers: 48 warps/SM
    case 3: launch_simple_cfg<DIV, 64, 24>(ctx, nRays, perm); break;    // <= 40 registers, smaller blocks
    default: launch_simple_cfg<DIV, 128, 1>(ctx, nRays, perm); break;
    }
}

template <int FETCH>
void launch_simple_fetch(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    if (ctx->carveout >= 0) {
        cudaFuncSetAttribute(k_extend_simple<DIV_MARKSTEIN1, kStack, 128, 12, FETCH>, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->carveout);
    }
    unsigned int* cursor = nullptr;
#ifdef UVRT_EXPERIMENTS
    if (FETCH == 5) {
        if (!ctx->dSmCursor) cudaMalloc((void**)&ctx->dSmCursor, 2 * 1024 * sizeof(unsigned int));
        cursor = ctx->dSmCursor + 1024 * ctx->slot;          // one cursor table per ray slot: extends may overlap
        cudaMemsetAsync(cursor, 0, 1024 * sizeof(unsigned int), ctx->xStream);
    }
#endif
    k_extend_simple<DIV_MARKSTEIN1, kStack, 128, 12, FETCH><<<grid_for(nRays, 128), 128, 0, ctx->xStream>>>(
        ctx->xCounts, ctx->dWtris, ctx->rs().dRays, ctx->dPairs, ctx->rootRef, nRays, ctx->sceneTame, perm, ctx->pairsTex, 0, cursor);
}

#ifdef UVRT_EXPERIMENTS
template <int DIV, int K, int HIST, int REFILL>
void launch_persist_r(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    constexpr int THREADS = 128;
    constexpr int MINB = 4;
    auto kern = k_extend_persist<DIV, kStack, K, REFILL, HIST, THREADS, MINB>;
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, THREADS, 0);
    if (perSm < 1) perSm = 1;
    if (ctx->blocksPerSm > 0) perSm = std::min(perSm, ctx->blocksPerSm);
    long long blocks = (long long)perSm * ctx->prop.multiProcessorCount;
    long long needed = (nRays + THREADS - 1) / THREADS;
    if (blocks > needed) blocks = needed;
    cudaMemsetAsync(ctx->dQueue, 0, sizeof(unsigned int), ctx->xStream);
    kern<<<(unsigned)blocks, THREADS, 0, ctx->xStream>>>(ctx->xCounts, ctx->dWtris, ctx->rs().dRays, ctx->dPairs,
                                                        ctx->rootRef, (uint32_t)nRays, ctx->sceneTame, ctx->dQueue, perm);
}

template <int DIV, int K, int HIST>
void launch_persist(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    // refill threshold: 0 = a warp takes 32 new rays only when all of its lanes are done
    if (HIST || ctx->refill >= 24) launch_persist_r<DIV, K, HIST, 24>(ctx, nRays, perm);
    else if (ctx->refill >= 8) launch_persist_r<DIV, K, 0, 8>(ctx, nRays, perm);
    else launch_persist_r<DIV, K, 0, 0>(ctx, nRays, perm);
}

template <int DIV, int K>
void launch_persist_h(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    if (ctx->histMode) launch_persist<DIV, K, 1>(ctx, nRays, perm);
    else launch_persist<DIV, K, 0>(ctx, nRays, perm);
}

template <int K>
void launch_persist_d(uvrt_ctx* ctx, long long nRays, int d, const uint32_t* perm)
{
    if (d == 0) launch_persist_h<DIV_IEEE, K>(ctx, nRays, perm);
    else if (d == 1) launch_persist_h<DIV_MARKSTEIN2, K>(ctx, nRays, perm);
    else launch_persist_h<DIV_MARKSTEIN1, K>(ctx, nRays, perm);
}

#endif

// Counting sort of the ray queue (kernels: uvrt_kernels.cuh "ray binning").  bin_prepare sizes the
// tables; the count step runs inside k_generate (countedRays) or as k_bin_count; bin_finish scans and
// scatters, leaving the permutation in ctx->rs().dPerm.
int bin_prepare(uvrt_ctx* ctx, long long nRays, BinDims* d, cudaStream_t stream)
{
    const int wanted = ctx->binY * ctx->binT * ctx->binP;
    int nBins = kBinsPerScanBlock;
    while (nBins < wanted) nBins *= 2;
    if (nBins > 64 * kBinsPerScanBlock)
        return fail(ctx, UVRT_ERR_INVALID, "bin_y*bin_t*bin_p = %d exceeds the supported %d bins", wanted, 64 * kBinsPerScanBlock);
    if (nBins > ctx->rs().binCap) {
        if (ctx->rs().dBinCount) cudaFree(ctx->rs().dBinCount);
        if (ctx->rs().dBinStart) cudaFree(ctx->rs().dBinStart);
        if (ctx->rs().dBinBlock) cudaFree(ctx->rs().dBinBlock);
        ctx->rs().dBinCount = ctx->rs().dBinStart = ctx->rs().dBinBlock = nullptr;
        ctx->rs().binCap = 0;
        CK(cudaMalloc((void**)&ctx->rs().dBinCount, (size_t)nBins * 4));
        CK(cudaMalloc((void**)&ctx->rs().dBinStart, (size_t)nBins * 4));
        CK(cudaMalloc((void**)&ctx->rs().dBinBlock, 64 * 4));
        // on the stream of the kernel that takes the bin slots next: a plain cudaMemset runs on the legacy stream,
        // which nothing orders before the non-blocking streams of this context
        CK(cudaMemsetAsync(ctx->rs().dBinCount, 0, (size_t)nBins * 4, stream));
        ctx->rs().binCap = nBins;
    }
    ctx->rs().binUsed = nBins;
    if (nRays > ctx->rs().permCap) {
        if (ctx->rs().dKeyRank) cudaFree(ctx->rs().dKeyRank);
        if (ctx->rs().dPerm) cudaFree(ctx->rs().dPerm);
        ctx->rs().dKeyRank = nullptr; ctx->rs().dPerm = nullptr; ctx->rs().permCap = 0;
        CK(cudaMalloc((void**)&ctx->rs().dKeyRank, (size_t)ctx->rs().rayCap * 8));
        CK(cudaMalloc((void**)&ctx->rs().dPerm, (size_t)ctx->rs().rayCap * 4));
        ctx->rs().permCap = ctx->rs().rayCap;
    }
    d->nY = ctx->rs().binExtentKnown ? ctx->binY : 1;
    d->nT = ctx->binT;
    d->nP = ctx->binP;
    d->y0 = ctx->rs().binY0;
    d->invLen = ctx->rs().binLen > 0.0f ? 1.0f / ctx->rs().binLen : 0.0f;
    return UVRT_OK;
}

int bin_finish(uvrt_ctx* ctx, long long nRays, cudaStream_t stream)
{
    StageTimer t(ctx, UVRT_STAGE_BIN, stream);
    if (ctx->rs().countedRays != nRays) {
        // the rays in the buffer did not (all) come from k_generate<1>: count them now
        BinDims d;
        int rc = bin_prepare(ctx, nRays, &d, stream);
        if (rc) return rc;
        if (ctx->rs().countedRays >= 0) CK(cudaMemsetAsync(ctx->rs().dBinCount, 0, (size_t)ctx->rs().binCap * 4, stream));
        k_bin_count<<<grid_for(nRays, 256), 256, 0, stream>>>(ctx->rs().dRays, (uint32_t)nRays, d, ctx->rs().dBinCount, ctx->rs().dKeyRank);
        ctx->launches++;
    }
    ctx->rs().countedRays = -1;
    const int nScanBlocks = ctx->rs().binUsed / kBinsPerScanBlock;
    k_bin_scan<<<nScanBlocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(ctx->rs().dBinCount),
                                                     reinterpret_cast<uint4*>(ctx->rs().dBinStart), ctx->rs().dBinBlock);
    k_bin_scatter<<<grid_for(nRays, 256), 256, 0, stream>>>(ctx->rs().dKeyRank, ctx->rs().dBinStart, ctx->rs().dBinBlock, nScanBlocks,
                                                                 (uint32_t)nRays, ctx->rs().dPerm);
    ctx->launches += 2;
    return UVRT_OK;
}

#ifdef UVRT_EXPERIMENTS
// The refill extend (uvrt_fast_refill.cuh, rejected): persistent conservative traversal, then the certificate for its winners, then the
// reference-order traversal of what is left.  The lists belong to the ray slot, like the permutation.
int launch_fast_refill(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    RaySlot& S = ctx->rs();
    cudaStream_t st = ctx->xStream;
    if (!ctx->qpairSelValid) {
        if ((size_t)ctx->nPairs > ctx->qpairSelCap) {
            uint4* fresh = nullptr;
            CK(cudaMalloc((void**)&fresh, (size_t)ctx->nPairs * 32));
            if (ctx->dQPairsSel) cudaFree(ctx->dQPairsSel);
            ctx->dQPairsSel = fresh;
            ctx->qpairSelCap = (size_t)ctx->nPairs;
        }
        k_fast_quantize<<<grid_for(ctx->nPairs, 256), 256, 0, st>>>(ctx->dPairs, ctx->nPairs, ctx->dQPairsSel, ctx->dFastGrid, 1);
        ctx->launches++;
        CK_LAUNCH("fast_quantize (layout 1)");
        ctx->qpairSelValid = true;
    }
    if (nRays > S.refillCap) {
        void* old[] = {S.dRefillCtl, S.dVerify, S.dRetry};
        for (void* p : old) if (p) cudaFree(p);
        S.dRefillCtl = nullptr; S.dVerify = nullptr; S.dRetry = nullptr; S.refillCap = 0;
        const long long cap = S.rayCap > nRays ? S.rayCap : nRays;
        CK(cudaMalloc((void**)&S.dRefillCtl, sizeof(RefillCtl)));
        CK(cudaMalloc((void**)&S.dVerify, (size_t)cap * 16));
        CK(cudaMalloc((void**)&S.dRetry, (size_t)cap * 4));
        S.refillCap = cap;
    }
    if (!ctx->refillBlocksPerSm) {
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend_fast_refill<kStack, 128, 8>, 128, 0));
        ctx->refillBlocksPerSm = nb > 0 ? nb : 1;
    }
    CK(cudaMemsetAsync(S.dRefillCtl, 0, sizeof(RefillCtl), st));
    const int sms = ctx->prop.multiProcessorCount;
    k_extend_fast_refill<kStack, 128, 8><<<sms * ctx->refillBlocksPerSm, 128, 0, st>>>(
        ctx->dWtris, S.dRays, ctx->dQPairsSel, ctx->fastGridHost, (uint32_t)nRays, perm, S.dRefillCtl, S.dVerify, S.dRetry,
        (uint32_t)ctx->refillChunk, ctx->fastCheck, ctx->dFastStats);
    k_fast_verify<<<sms * 8, 128, 0, st>>>(ctx->xCounts, ctx->dWtris, S.dRays, S.dRefillCtl, S.dVerify, S.dRetry, ctx->dFastStats,
                                           ctx->fastCheck);
    k_extend_retry<kStack><<<ctx->fastCheck ? sms * 8 : sms, 128, 0, st>>>(ctx->xCounts, ctx->dWtris, S.dRays, ctx->dPairs, S.dRefillCtl,
                                                                         S.dRetry, ctx->dFastStats);
    ctx->launches += 2;
    return UVRT_OK;
}
#endif

int launch_extend(uvrt_ctx* ctx, long long nRays, const uint32_t* perm)
{
    int v = ctx->extendVariant < 0 ? default_variant(ctx) : ctx->extendVariant;
    if (v == 0) launch_simple<DIV_IEEE>(ctx, nRays, perm);
    else if (v == 1) launch_simple<DIV_MARKSTEIN2>(ctx, nRays, perm);
    else if (v == 2 && ctx->fetchMode == 3 && ctx->simpleCfg == 1) launch_simple_fetch<3>(ctx, nRays, perm);
#ifdef UVRT_EXPERIMENTS
    else if (v == 2 && ctx->fetchMode == 1 && ctx->pairsTex) launch_simple_fetch<1>(ctx, nRays, perm);
    else if (v == 2 && ctx->fetchMode == 2 && ctx->pairsTex) launch_simple_fetch<2>(ctx, nRays, perm);
    else if (v == 2 && ctx->fetchMode == 4) launch_simple_fetch<4>(ctx, nRays, perm);
    else if (v == 2 && ctx->fetchMode == 5 && perm) launch_simple_fetch<5>(ctx, nRays, perm);
#endif
    else if (v == 2) launch_simple<DIV_MARKSTEIN1>(ctx, nRays, perm);
    else if (v == 50) {
        // certified fast extend (uvrt_fast.cuh); scenes it cannot serve (boxes not tame / not nested, a single leaf)
        // take the exact kernel
        if (fast_usable(ctx) && nRays <= 0x7fffffffLL) {
#define UVRT_FAST_LAUNCH(THREADS, MINB)                                                                                  \
    k_extend_fast<kStack, THREADS, MINB><<<grid_for(nRays, THREADS), THREADS, 0, ctx->xStream>>>(                          \
        ctx->xCounts, ctx->dWtris, ctx->rs().dRays, ctx->dPairs, ctx->dQPairs, ctx->dFastGrid, (uint32_t)nRays, perm, ctx->dFastStats, ctx->fastCheck)
            // "fast_cfg": 0 = 128 threads, at most 48 registers (46 used; 40 resident warps per SM; the default),
            // 1 = at most 64 registers (50 used; 32 warps).  (64-thread blocks: the same; 256-thread blocks: 2.5 % slower.)
#ifdef UVRT_EXPERIMENTS
            if (ctx->fastCfg == 2) {
                int rc = launch_fast_refill(ctx, nRays, perm);
                if (rc) return rc;
            } else
#endif
            if (ctx->fastCfg == 1) UVRT_FAST_LAUNCH(128, 8);
            else UVRT_FAST_LAUNCH(128, 10);
#undef UVRT_FAST_LAUNCH
        } else
            launch_simple_fetch<3>(ctx, nRays, perm);
    }
#ifdef UVRT_EXPERIMENTS
    else if (v >= 40 && v < 44) {
        // chunk-persistent warps (k_extend_chunk): K = {1, 2, 4, 8}[v - 40]; "refill", "chunk" options
        switch (v - 40) {
        case 0: launch_chunk_k<1>(ctx, nRays, perm); break;
        case 1: launch_chunk_k<2>(ctx, nRays, perm); break;
        case 2: launch_chunk_k<4>(ctx, nRays, perm); break;
        default: launch_chunk_k<8>(ctx, nRays, perm); break;
        }
    }
    else if (v >= 10 && v < 25) {
        int k = (v - 10) / 3, d = (v - 10) % 3;
        switch (k) {
        case 0: launch_persist_d<1>(ctx, nRays, d, perm); break;
        case 1: launch_persist_d<2>(ctx, nRays, d, perm); break;
        case 2: launch_persist_d<4>(ctx, nRays, d, perm); break;
        case 3: launch_persist_d<8>(ctx, nRays, d, perm); break;
        default: launch_persist_d<16>(ctx, nRays, d, perm); break;
        }
    }
#endif
    else
        return fail(ctx, UVRT_ERR_INVALID, "unknown extend_variant %d%s", v,
#ifdef UVRT_EXPERIMENTS
                    "");
#else
                    " (variants 10..24 and 40..43 exist only in builds with -DUVRT_EXPERIMENTS)");
#endif
    ctx->launches++;
    return UVRT_OK;
}

} // namespace

// ================================================================================================
extern "C" {

const char* uvrt_version(void) { return "uvrt-b200 0.1 (sm_100a)"; }

int uvrt_device_count(int* count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess) return fail(nullptr, UVRT_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return UVRT_OK;
}

int uvrt_create(uvrt_ctx** out, int device)
{
    if (!out) return fail(nullptr, UVRT_ERR_INVALID, "uvrt_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, UVRT_ERR_NO_DEVICE, "no usable CUDA device (%s); libuvrt has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, UVRT_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    uvrt_ctx* ctx = new (std::nothrow) uvrt_ctx();
    if (!ctx) return fail(nullptr, UVRT_ERR_NO_MEMORY, "out of host memory");
    ctx->device = device;
    cudaGetDeviceProperties(&ctx->prop, device);
    if (ctx->prop.major < 10) {
        int rc = fail(nullptr, UVRT_ERR_NO_DEVICE, "device %d is sm_%d%d; this build only carries sm_100a code", device,
                      ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return rc;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // higher priority: generate's blocks are placed as soon as extend's blocks retire, so the two
        // really run side by side instead of generate waiting for extend's last wave
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&ctx->genStream, cudaStreamNonBlocking, hi);
    }
    for (int k = 0; k < 2; k++) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->extStream[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->extDone[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->accDone[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->forkEv, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&ctx->accStream, cudaStreamNonBlocking, hi);
    }
    for (RaySlot& r : ctx->slots) {
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.genDone, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.freeEv, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->dQueue, 256);
    for (int i = 0; i < 16 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->marks[i]);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, UVRT_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    *out = ctx;
    return UVRT_OK;
}

void uvrt_destroy(uvrt_ctx* ctx)
{
    if (!ctx) return;
    Bind b(ctx);
    debug_pending_error("(entering uvrt_destroy)");
    if (ctx->genStream) cudaStreamSynchronize(ctx->genStream);
    for (int k = 0; k < 2; k++) if (ctx->extStream[k]) cudaStreamSynchronize(ctx->extStream[k]);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 2; k++) {
        if (ctx->extStream[k]) cudaStreamDestroy(ctx->extStream[k]);
        if (ctx->extDone[k]) cudaEventDestroy(ctx->extDone[k]);
        if (ctx->accDone[k]) cudaEventDestroy(ctx->accDone[k]);
    }
    if (ctx->poolTuned) {
        // hand the memory that uvrt_build_bvh kept in the device's allocation pool back to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    if (ctx->forkEv) cudaEventDestroy(ctx->forkEv);
    if (ctx->accStream) { cudaStreamSynchronize(ctx->accStream); cudaStreamDestroy(ctx->accStream); }
    if (ctx->dCountsAlt) cudaFree(ctx->dCountsAlt);
    if (ctx->dQPairs) cudaFree(ctx->dQPairs);
#ifdef UVRT_EXPERIMENTS
    if (ctx->dQPairsSel) cudaFree(ctx->dQPairsSel);
    for (RaySlot& r : ctx->slots) {
        void* q[] = {r.dRefillCtl, r.dVerify, r.dRetry};
        for (void* p : q) if (p) cudaFree(p);
    }
#endif
    if (ctx->dFastGrid) cudaFree(ctx->dFastGrid);
    if (ctx->dFastStats) cudaFree(ctx->dFastStats);
    for (int k = 0; k < 2; k++) {
        if (ctx->matrixBuf[k]) cudaFree(ctx->matrixBuf[k]);
        if (ctx->durBuf[k]) cudaFree(ctx->durBuf[k]);
        if (ctx->matrixReady[k]) cudaEventDestroy(ctx->matrixReady[k]);
        if (ctx->matrixFolded[k]) cudaEventDestroy(ctx->matrixFolded[k]);
    }
    if (ctx->timelineOrigin) cudaEventDestroy(ctx->timelineOrigin);
    if (ctx->dVertsSpare) cudaFree(ctx->dVertsSpare);
    if (ctx->vertsEv) cudaEventDestroy(ctx->vertsEv);
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    void* ptrs[] = {ctx->dPairs, ctx->dWtris, ctx->dVerts, ctx->dCounts, ctx->dSum, ctx->dMax, ctx->dDose,
                    ctx->dColor, ctx->dQueue, ctx->dSmCursor, ctx->dSeeds, ctx->dSeedPos, ctx->dFlush, ctx->dRawNodes, ctx->dRawIdx,
                    ctx->dPrepQueue, ctx->dPrepNode, ctx->dPrepStatus};
    if (ctx->hPrepStatus) cudaFreeHost(ctx->hPrepStatus);
    for (void* p : ptrs) if (p) cudaFree(p);
    for (RaySlot& r : ctx->slots) {
        void* q[] = {r.dRays, r.dBinCount, r.dBinStart, r.dBinBlock, r.dKeyRank, r.dPerm};
        for (void* p : q) if (p) cudaFree(p);
        if (r.genDone) cudaEventDestroy(r.genDone);
        if (r.freeEv) cudaEventDestroy(r.freeEv);
    }
    if (ctx->genStream) cudaStreamDestroy(ctx->genStream);
    if (ctx->hStage) cudaFreeHost(ctx->hStage);
    if (ctx->pairsTex) cudaDestroyTextureObject(ctx->pairsTex);
    for (auto& t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
    for (auto e : ctx->freeEvents) cudaEventDestroy(e);
    for (auto e : ctx->marks) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    debug_pending_error("uvrt_destroy");
    delete ctx;
}

const char* uvrt_last_error(const uvrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int uvrt_device_info(uvrt_ctx* ctx, char* dst, size_t bytes, int* smCount, int* ccMajor, int* ccMinor)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (dst && bytes) snprintf(dst, bytes, "%s", ctx->prop.name);
    if (smCount) *smCount = ctx->prop.multiProcessorCount;
    if (ccMajor) *ccMajor = ctx->prop.major;
    if (ccMinor) *ccMinor = ctx->prop.minor;
    return UVRT_OK;
}

// device buffers of the traversal layout and the per-triangle state.  Everything the new scene needs is
// allocated BEFORE anything of the old scene is released, so a failed allocation leaves the old scene usable
// (the header's promise for rejected uploads).
static int ensure_scene_buffers(uvrt_ctx* ctx, size_t pairBytes, size_t wtriBytes, int nTris, bool vertsToo)
{
    float4 *nPairsBuf = nullptr, *nWtris = nullptr, *nVerts = nullptr;
    int *nCounts = nullptr, *nCountsAlt = nullptr;
    double *nSum = nullptr, *nMax = nullptr;
    float *nDose = nullptr, *nColor = nullptr;
    const bool growPairs = pairBytes > ctx->pairCap, growWtris = wtriBytes > ctx->wtriCap;
    const bool growVerts = vertsToo && (size_t)nTris > ctx->vertsCap, perTri = nTris != ctx->nTris;
    cudaError_t e = cudaSuccess;
    auto want = [&](bool on, void** p, size_t bytes) { if (on && e == cudaSuccess) e = cudaMalloc(p, bytes); };
    want(growPairs, (void**)&nPairsBuf, pairBytes);
    want(growWtris, (void**)&nWtris, wtriBytes);
    want(growVerts, (void**)&nVerts, (size_t)nTris * 64);
    want(perTri, (void**)&nCounts, (size_t)nTris * 4);
    want(perTri, (void**)&nCountsAlt, (size_t)nTris * 4);
    want(perTri, (void**)&nSum, (size_t)nTris * 8);
    want(perTri, (void**)&nMax, (size_t)nTris * 8);
    want(perTri, (void**)&nDose, (size_t)nTris * 4);
    want(perTri, (void**)&nColor, (size_t)nTris * 36);
    if (e != cudaSuccess) {
        void* fresh[] = {nPairsBuf, nWtris, nVerts, nCounts, nCountsAlt, nSum, nMax, nDose, nColor};
        for (void* p : fresh) if (p) cudaFree(p);
        return fail(ctx, e == cudaErrorMemoryAllocation ? UVRT_ERR_NO_MEMORY : UVRT_ERR_CUDA,
                    "upload_scene: device allocation failed (%s); the previous scene is untouched", cudaGetErrorString(e));
    }
    auto swapIn = [](auto** slot, auto* fresh) { if (*slot) cudaFree(*slot); *slot = fresh; };
    if (growPairs) {
#ifdef UVRT_EXPERIMENTS
        if (ctx->pairsTex) { cudaDestroyTextureObject(ctx->pairsTex); ctx->pairsTex = 0; }
#endif
        swapIn(&ctx->dPairs, nPairsBuf);
        ctx->pairCap = pairBytes;
#ifdef UVRT_EXPERIMENTS
        if (pairBytes / 16 <= (1u << 27)) {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = ctx->dPairs;
            rd.res.linear.desc = cudaCreateChannelDesc<float4>();
            rd.res.linear.sizeInBytes = pairBytes;
            cudaTextureDesc td{};
            td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&ctx->pairsTex, &rd, &td, nullptr) != cudaSuccess) { ctx->pairsTex = 0; cudaGetLastError(); }
        }
#endif
    }
    if (growWtris) { swapIn(&ctx->dWtris, nWtris); ctx->wtriCap = wtriBytes; }
    if (growVerts) { swapIn(&ctx->dVerts, nVerts); ctx->vertsCap = (size_t)nTris; }
    if (perTri) {
        swapIn(&ctx->dCounts, nCounts); swapIn(&ctx->dCountsAlt, nCountsAlt);
        swapIn(&ctx->dSum, nSum); swapIn(&ctx->dMax, nMax);
        swapIn(&ctx->dDose, nDose); swapIn(&ctx->dColor, nColor);
        ctx->nTris = nTris;
        CK(cudaMemsetAsync(ctx->dCounts, 0, (size_t)nTris * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->dCountsAlt, 0, (size_t)nTris * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->dSum, 0, (size_t)nTris * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->dMax, 0, (size_t)nTris * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->dDose, 0, (size_t)nTris * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->dColor, 0, (size_t)nTris * 36, ctx->stream));
    }
    return UVRT_OK;
}

static int fast_prepare(uvrt_ctx* ctx);

static int ensure_stage(uvrt_ctx* ctx, size_t total)
{
    if (ctx->hStageBytes < total) {
        if (ctx->hStage) cudaFreeHost(ctx->hStage);
        ctx->hStage = nullptr;
        ctx->hStageBytes = 0;
        CK(cudaMallocHost(&ctx->hStage, total));
        ctx->hStageBytes = total;
    }
    return UVRT_OK;
}

// memcpy spread over the host cores (the staging copy of a 1 GB scene is otherwise the slowest step)
// Threads: the host's cores divided by the ranks sharing it (uvrt_comm_init tells how many there are;
// launchers such as torchrun set OMP_NUM_THREADS=1, which would leave one core per rank copying).
static int g_copyThreads = 0;
static void par_copy(void* dst, const void* src, size_t bytes)
{
    const size_t chunk = 256u << 10;
    const long long nChunks = (long long)((bytes + chunk - 1) / chunk);
    int threads = g_copyThreads;
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("OMP_NUM_THREADS")) { int v = atoi(e); if (v >= 1) threads = v; }
        threads = std::max(1, std::min(threads, 16));
    }
#pragma omp parallel for schedule(static) num_threads(threads) if (nChunks > 2)
    for (long long c = 0; c < nChunks; c++) {
        const size_t off = (size_t)c * chunk;
        memcpy((char*)dst + off, (const char*)src + off, std::min(chunk, bytes - off));
    }
}

// Host array -> pinned staging -> device in pieces, so that the DMA of one piece runs while the host
// cores copy the next (the user's arrays are pageable; the staging copy is as slow as the DMA).
static cudaError_t stage_and_copy(void* dDst, const void* hSrc, char* stage, size_t bytes, cudaStream_t st)
{
    // a room (a few MB) goes up in one piece -- the copy is spread over the host cores and takes as long as
    // starting several DMAs would; large scenes in 32 MiB pieces
    const size_t piece = bytes > (8u << 20) ? (32u << 20) : bytes;
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t n = std::min(piece, bytes - off);
        par_copy(stage + off, (const char*)hSrc + off, n);
        cudaError_t e = cudaMemcpyAsync((char*)dDst + off, stage + off, n, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// The way back: all DMAs into the pinned staging area are queued at once, and the host cores copy each piece
// out to the caller's (pageable) array as soon as its DMA has landed.
struct Unstage { void* hDst; const void* dSrc; size_t bytes; };
static cudaError_t copy_and_unstage(const Unstage* items, int nItems, char* stage, cudaStream_t st)
{
    struct Piece { void* dst; const char* src; size_t n; cudaEvent_t ev; };
    std::vector<Piece> pieces;
    size_t at = 0;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < nItems && e == cudaSuccess; k++) {
        const size_t piece = items[k].bytes > (8u << 20) ? (32u << 20) : std::max<size_t>(items[k].bytes, 1);
        for (size_t off = 0; off < items[k].bytes && e == cudaSuccess; off += piece) {
            const size_t n = std::min(piece, items[k].bytes - off);
            Piece p{(char*)items[k].hDst + off, stage + at, n, nullptr};
            e = cudaMemcpyAsync(stage + at, (const char*)items[k].dSrc + off, n, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(p.ev, st);
            pieces.push_back(p);
            at += (n + 255) & ~(size_t)255;
        }
    }
    for (Piece& p : pieces) {
        if (e == cudaSuccess && p.ev) e = cudaEventSynchronize(p.ev);
        if (e == cudaSuccess) par_copy(p.dst, p.src, p.n);
        if (p.ev) cudaEventDestroy(p.ev);
    }
    return e;
}

// The repack on the device (uvrt_scene_prep.cuh): the reference's three arrays go up as they are.
static int upload_scene_device(uvrt_ctx* ctx, const void* trisV, int nTris, const void* nodesV, int nNodes, const uint32_t* triIdx)
{
    using namespace uvrt_prep;
    const size_t nodeBytes = (size_t)nNodes * 32, idxBytes = (size_t)nTris * 4, vertBytes = (size_t)nTris * 64;
    const size_t idxOff = (nodeBytes + 255) & ~(size_t)255, vertOff = (idxOff + idxBytes + 255) & ~(size_t)255;
    int rc = ensure_stage(ctx, vertOff + vertBytes);
    if (rc) return rc;
    if ((size_t)nNodes > ctx->prepNodeCap) {
        ctx->prepNodeCap = 0;
        void* old[] = {ctx->dRawNodes, ctx->dPrepQueue, ctx->dPrepNode};
        for (void* p : old) if (p) cudaFree(p);
        ctx->dRawNodes = nullptr; ctx->dPrepQueue = nullptr; ctx->dPrepNode = nullptr;
        CK(cudaMalloc(&ctx->dRawNodes, nodeBytes));
        CK(cudaMalloc((void**)&ctx->dPrepQueue, (size_t)nNodes * 8));
        CK(cudaMalloc((void**)&ctx->dPrepNode, (size_t)nNodes * 16));
        ctx->prepNodeCap = (size_t)nNodes;
    }
    if ((size_t)nTris > ctx->prepTriCap) {
        ctx->prepTriCap = 0;
        if (ctx->dRawIdx) cudaFree(ctx->dRawIdx);
        ctx->dRawIdx = nullptr;
        CK(cudaMalloc((void**)&ctx->dRawIdx, idxBytes));
        ctx->prepTriCap = (size_t)nTris;
    }
    if (!ctx->dPrepStatus) {
        CK(cudaMalloc(&ctx->dPrepStatus, sizeof(Status)));
        CK(cudaMallocHost(&ctx->hPrepStatus, sizeof(Status)));
        int perSm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_prep_walk, 128, 0));
        ctx->prepBlocks = std::max(1, std::min(perSm, 8)) * ctx->prop.multiProcessorCount;   // all blocks resident
    }
    cudaStream_t st = ctx->stream;
    const bool verbose = getenv("UVRT_UPLOAD_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tPhase = now();
    auto phase = [&](const char* name) {
        if (!verbose) return;
        double t = now();
        fprintf(stderr, "[uvrt_upload_scene] %-22s %8.1f us\n", name, t - tPhase);
        tPhase = t;
    };
    char* stage = (char*)ctx->hStage;
    Status* dSt = (Status*)ctx->dPrepStatus;
    Status* hSt = (Status*)ctx->hPrepStatus;
    uint32_t* parent = ctx->dPrepNode;
    uint32_t *arrive = parent + (size_t)ctx->prepNodeCap, *subInner = arrive + (size_t)ctx->prepNodeCap,
             *subSlots = subInner + (size_t)ctx->prepNodeCap;
    // nodes + triIdx first: the tree walk needs nothing else and runs while the triangles are staged
    CK(cudaMemsetAsync(ctx->dPrepQueue, 0xff, (size_t)nNodes * 8, st));
    CK(cudaMemsetAsync(parent, 0xff, (size_t)nNodes * 4, st));
    CK(cudaMemsetAsync(arrive, 0, (size_t)nNodes * 4, st));
    CK(stage_and_copy(ctx->dRawNodes, nodesV, stage, nodeBytes, st));
    CK(stage_and_copy(ctx->dRawIdx, triIdx, stage + idxOff, idxBytes, st));
    phase("stage+copy nodes, idx");
    k_prep_init<<<1, 1, 0, st>>>(ctx->dPrepQueue, dSt, parent);
    const int blocks = (int)std::min<long long>(ctx->prepBlocks, ((long long)nNodes + 127) / 128);
    k_prep_walk<<<blocks, 128, 0, st>>>((const RawNode*)ctx->dRawNodes, (uint32_t)nNodes, ctx->dRawIdx, (uint32_t)nTris, kStack,
                                        ctx->dPrepQueue, parent, arrive, subInner, subSlots, dSt);
    CK(cudaMemcpyAsync(hSt, dSt, sizeof(Status), cudaMemcpyDeviceToHost, st));
    ctx->launches += 2;
    phase("enqueue copies+walk");
    // the triangles travel on the second stream, next to the walk, into a spare buffer: the previous
    // scene stays intact until the new tree has been validated
    if ((size_t)nTris > ctx->vertsSpareCap) {
        ctx->vertsSpareCap = 0;
        if ((rc = dev_alloc(ctx, &ctx->dVertsSpare, (size_t)nTris * 4))) return rc;
        ctx->vertsSpareCap = (size_t)nTris;
    }
    if (!ctx->vertsEv) CK(cudaEventCreateWithFlags(&ctx->vertsEv, cudaEventDisableTiming));
    CK(stage_and_copy(ctx->dVertsSpare, trisV, stage + vertOff, vertBytes, ctx->genStream));
    CK(cudaEventRecord(ctx->vertsEv, ctx->genStream));
    phase("stage triangles");
    CK(cudaStreamSynchronize(st));
    phase("wait for the walk");
    CK_LAUNCH("scene walk");
    const Status s = *hSt;
    // every early return below leaves the previous scene in place; the triangle DMA from the staging area
    // must have landed before the caller may reuse it (uvrt_build_bvh, the next upload)
    struct DmaGuard {
        cudaStream_t s; bool armed;
        ~DmaGuard() { if (armed) cudaStreamSynchronize(s); }
    } dmaGuard{ctx->genStream, true};
    switch (s.err) {
    case PREP_OK: break;
    case PREP_NODE_RANGE:
        return fail(ctx, UVRT_ERR_INVALID,
                    "upload_scene: node %u is reachable but nNodes is %d (the reference's nodesUsed = 2N "
                    "truncates its own tree; pass the full array)", s.errA, nNodes);
    case PREP_TWICE: return fail(ctx, UVRT_ERR_INVALID, "upload_scene: node %u reached twice", s.errA);
    case PREP_LEAF_SPAN:
        return fail(ctx, UVRT_ERR_INVALID, "upload_scene: leaf %u spans triIdx[%u..%u) of %d", s.errA, s.errB, s.errC, nTris);
    case PREP_TRI_RANGE: return fail(ctx, UVRT_ERR_INVALID, "upload_scene: triIdx value %u out of range", s.errA);
    case PREP_DEPTH: return fail(ctx, UVRT_ERR_INVALID, "upload_scene: BVH depth %u exceeds the traversal stack (%d)", s.errA, kStack);
    default: return fail(ctx, UVRT_ERR_CUDA, "upload_scene: the device tree walk did not finish (code %u at queue entry %u)", s.err, s.errA);
    }
    if (!s.done) return fail(ctx, UVRT_ERR_CUDA, "upload_scene: the device tree walk ended without reaching the root");
    const int nPairs = (int)s.nPairs;
    const unsigned long long nSlots = s.nSlots;
    const size_t pairBytes = (size_t)std::max(nPairs, 1) * 64, wtriBytes = (size_t)std::max<unsigned long long>(nSlots, 1) * 64;
    // the new scene's buffers first (an allocation failure must not cost the old scene its triangles), then the swap
    if ((rc = ensure_scene_buffers(ctx, pairBytes, wtriBytes, nTris, false))) return rc;
    std::swap(ctx->dVerts, ctx->dVertsSpare);
    std::swap(ctx->vertsCap, ctx->vertsSpareCap);
    dmaGuard.armed = false;
    CK(cudaStreamWaitEvent(st, ctx->vertsEv, 0));
    if (nPairs == 0) CK(cudaMemsetAsync(ctx->dPairs, 0, pairBytes, st));
    const uint32_t reachable = std::min<uint32_t>(s.tail, (uint32_t)nNodes);
    k_prep_emit<<<grid_for(reachable, 256), 256, 0, st>>>((const RawNode*)ctx->dRawNodes, ctx->dRawIdx, ctx->dVerts, ctx->dPrepQueue,
                                                         reachable, parent, subInner, subSlots, ctx->dPairs, ctx->dWtris, dSt);
    CK(cudaMemcpyAsync(hSt, dSt, sizeof(Status), cudaMemcpyDeviceToHost, st));
    ctx->launches += 1;
    phase("enqueue verts+emit");
    CK(cudaStreamSynchronize(st));   // the staging buffer may be reused right away
    phase("wait for emit");
    CK_LAUNCH("scene emit");
    ctx->nPairs = nPairs;
    ctx->nLeaves = nPairs + 1;       // a binary tree
    ctx->nSlots = (long long)nSlots;
    ctx->depth = (int)s.maxDepth;
    ctx->sceneTame = hSt->tame ? 1 : 0;
    ctx->sceneNested = hSt->nested ? 1 : 0;
    ctx->rootRef = s.rootIsLeaf ? kLeafFlag : 0u;
    ctx->uploadBytes = (int64_t)(nodeBytes + idxBytes + vertBytes);
    return UVRT_OK;
}

static int upload_scene_host(uvrt_ctx* ctx, const void* trisV, int nTris, const void* nodesV, int nNodes, const uint32_t* triIdx)
{
    const HostTri* tris = (const HostTri*)trisV;
    const HostNode* nodes = (const HostNode*)nodesV;

    // ---- pass 1: walk the tree from the root, number inner nodes (pairs) and leaf slots -------
    std::vector<int32_t>& id = ctx->upId;        // inner: pair index; leaf: first slot
    std::vector<uint32_t>& order = ctx->upOrder; // reachable nodes, pre-order, left child first
    id.assign((size_t)nNodes, -1);
    order.clear();
    struct Item { uint32_t node; int depth; };
    std::vector<Item> stack;
    stack.push_back({0u, 0});
    int nPairs = 0, nLeaves = 0, depth = 0;
    long long nSlots = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (it.node >= (uint32_t)nNodes)
            return fail(ctx, UVRT_ERR_INVALID,
                        "upload_scene: node %u is reachable but nNodes is %d (the reference's nodesUsed = 2N "
                        "truncates its own tree; pass the full array)", it.node, nNodes);
        if (id[it.node] != -1) return fail(ctx, UVRT_ERR_INVALID, "upload_scene: node %u reached twice", it.node);
        const HostNode& nd = nodes[it.node];
        depth = std::max(depth, it.depth);
        order.push_back(it.node);
        if (nd.triCount > 0) {
            if ((long long)nd.leftFirst + nd.triCount > nTris)
                return fail(ctx, UVRT_ERR_INVALID, "upload_scene: leaf %u spans triIdx[%u..%u) of %d", it.node,
                            nd.leftFirst, nd.leftFirst + nd.triCount, nTris);
            id[it.node] = (int32_t)nSlots;
            for (uint32_t k = 0; k < nd.triCount; k++) {
                uint32_t t = triIdx[nd.leftFirst + k];
                if (t >= (uint32_t)nTris) return fail(ctx, UVRT_ERR_INVALID, "upload_scene: triIdx value %u out of range", t);
            }
            nSlots += nd.triCount;
            nLeaves++;
        } else {
            id[it.node] = nPairs++;
            stack.push_back({nd.leftFirst + 1, it.depth + 1});
            stack.push_back({nd.leftFirst, it.depth + 1});
        }
    }
    if (depth + 1 > kStack)
        return fail(ctx, UVRT_ERR_INVALID, "upload_scene: BVH depth %d exceeds the traversal stack (%d)", depth, kStack);

    // ---- pass 2: fill the staging image --------------------------------------------------------
    size_t pairBytes = (size_t)std::max(nPairs, 1) * 64, wtriBytes = (size_t)std::max<long long>(nSlots, 1) * 64,
           vertBytes = (size_t)nTris * 64;
    size_t total = pairBytes + wtriBytes + vertBytes;
    {
        int rcs = ensure_stage(ctx, total);
        if (rcs) return rcs;
    }
    float* hp = (float*)ctx->hStage;
    float* hw = (float*)((char*)ctx->hStage + pairBytes);
    float* hv = (float*)((char*)ctx->hStage + pairBytes + wtriBytes);
    if (nPairs == 0) memset(hp, 0, pairBytes);
    bool tame = true;
    auto child_ref = [&](uint32_t n) -> uint32_t {
        return nodes[n].triCount > 0 ? (kLeafFlag | (uint32_t)id[n]) : (uint32_t)id[n];
    };
    // every reachable node fills its own record: independent, so spread over the host cores
    int tameAll = 1, nestedAll = 1;
    const long long nOrder = (long long)order.size();
#pragma omp parallel for schedule(static) reduction(&& : tameAll, nestedAll) if (nOrder > 4096)
    for (long long oi = 0; oi < nOrder; oi++) {
        const uint32_t n = order[oi];
        const HostNode& nd = nodes[n];
        if (nd.triCount > 0) {
            for (uint32_t k = 0; k < nd.triCount; k++) {
                uint32_t t = triIdx[nd.leftFirst + k];
                const float* v = tris[t].v;
                float* w = hw + ((size_t)id[n] + k) * 16;
                w[0] = v[0]; w[1] = v[1]; w[2] = v[2];
                uint32_t tag = t | (k + 1 == nd.triCount ? kLastFlag : 0u);
                memcpy(&w[3], &tag, 4);
                // edge1 = v1 - v0, edge2 = v2 - v0 (extend.cl:13): the same fp32 subtractions, hoisted
                // spare lanes: the leaf's own box, for the exact verification step of the fast extend (uvrt_fast.cuh)
                w[4] = v[4] - v[0]; w[5] = v[5] - v[1]; w[6] = v[6] - v[2]; w[7] = nd.mn[2];
                w[8] = v[8] - v[0]; w[9] = v[9] - v[1]; w[10] = v[10] - v[2]; w[11] = nd.mx[2];
                w[12] = nd.mn[0]; w[13] = nd.mn[1]; w[14] = nd.mx[0]; w[15] = nd.mx[1];
            }
        } else {
            float* p = hp + (size_t)id[n] * 16;
            for (int c = 0; c < 2; c++) {
                const HostNode& ch = nodes[nd.leftFirst + c];
                // (min.x, min.y) (max.x, max.y) (min.z, max.z) (ref, 0): 64-bit pairs for the packed pipe
                float* q = p + c * 8;
                q[0] = ch.mn[0]; q[1] = ch.mn[1]; q[2] = ch.mx[0]; q[3] = ch.mx[1];
                q[4] = ch.mn[2]; q[5] = ch.mx[2];
                uint32_t ref = child_ref(nd.leftFirst + c);
                memcpy(&q[6], &ref, 4);
                q[7] = 0.0f;
                // the fast box test needs normal-range coordinates and min <= max on every axis
                for (int a = 0; a < 3; a++) {
                    tameAll = tameAll && coord_tame(ch.mn[a]) && coord_tame(ch.mx[a]) && ch.mn[a] <= ch.mx[a];
                    if (n != 0u) nestedAll = nestedAll && ch.mn[a] >= nd.mn[a] && ch.mx[a] <= nd.mx[a];
                }
            }
        }
    }
    tame = tameAll != 0;
    memcpy(hv, tris, vertBytes);

    // ---- device buffers -------------------------------------------------------------------------
    int rc;
    if ((rc = ensure_scene_buffers(ctx, pairBytes, wtriBytes, nTris, true))) return rc;
    CK(cudaMemcpyAsync(ctx->dPairs, hp, pairBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dWtris, hw, wtriBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dVerts, hv, vertBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // the staging buffer may be reused right away
    ctx->nPairs = nPairs;
    ctx->nLeaves = nLeaves;
    ctx->nSlots = nSlots;
    ctx->depth = depth;
    ctx->sceneTame = tame ? 1 : 0;
    ctx->sceneNested = nestedAll ? 1 : 0;
    ctx->rootRef = child_ref(0);
    ctx->uploadBytes = (int64_t)total;
    return UVRT_OK;
}

int uvrt_upload_scene(uvrt_ctx* ctx, const void* trisV, int nTris, const void* nodesV, int nNodes,
                      const uint32_t* triIdx)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!trisV || !nodesV || !triIdx || nTris <= 0 || nNodes <= 0)
        return fail(ctx, UVRT_ERR_INVALID, "upload_scene: null pointer or empty scene (nTris=%d nNodes=%d)", nTris, nNodes);
    ApiScope api_(ctx, "uvrt_upload_scene");
    Bind b(ctx);
    // rays of a pipelined launch may still be in flight on the second stream
    if (ctx->genStream) CK(cudaStreamSynchronize(ctx->genStream));
    int rc = ctx->hostRepack ? upload_scene_host(ctx, trisV, nTris, nodesV, nNodes, triIdx)
                             : upload_scene_device(ctx, trisV, nTris, nodesV, nNodes, triIdx);
    if (rc) return rc;
    return fast_prepare(ctx);
}

// The node layout of the certified fast extend (uvrt_fast.cuh): quantised copies of the pair records, built on the
// device from the repacked scene.  Only trees whose boxes are tame and nested qualify; others keep the exact kernel.

static int fast_prepare(uvrt_ctx* ctx)
{
    if (ctx->nPairs <= 0 || !ctx->sceneTame || !ctx->sceneNested) return UVRT_OK;
    if ((size_t)ctx->nPairs > ctx->qpairCap) {
        uint4* fresh = nullptr;
        CK(cudaMalloc((void**)&fresh, (size_t)ctx->nPairs * 32));
        if (ctx->dQPairs) cudaFree(ctx->dQPairs);
        ctx->dQPairs = fresh;
        ctx->qpairCap = (size_t)ctx->nPairs;
    }
    if (!ctx->dFastGrid) CK(cudaMalloc((void**)&ctx->dFastGrid, sizeof(FastGrid)));
    if (!ctx->dFastStats) {
        CK(cudaMalloc((void**)&ctx->dFastStats, sizeof(FastStats)));
        CK(cudaMemsetAsync(ctx->dFastStats, 0, sizeof(FastStats), ctx->stream));
    }
    k_fast_quantize<<<grid_for(ctx->nPairs, 256), 256, 0, ctx->stream>>>(ctx->dPairs, ctx->nPairs, ctx->dQPairs, ctx->dFastGrid, 0);
    ctx->launches++;
    CK_LAUNCH("fast_quantize");
#ifdef UVRT_EXPERIMENTS
    CK(cudaMemcpyAsync(&ctx->fastGridHost, ctx->dFastGrid, sizeof(FastGrid), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->qpairSelValid = false;
#endif
    CK(cudaStreamSynchronize(ctx->stream));
    return UVRT_OK;
}

// ---- device BVH build (uvrt_bvh_build.cuh) ----------------------------------------------------------
int uvrt_build_bvh(uvrt_ctx* ctx, const void* trisHost, int nTris, void* nodesOut, int nodeCapacity,
                   uint32_t* triIdxOut, uint32_t* nodesUsedOut, void* trisOut)
{
    using namespace uvrt_bvh;
    if (!ctx) return UVRT_ERR_INVALID;
    if (!trisHost || nTris <= 0 || !nodesOut || !triIdxOut)
        return fail(ctx, UVRT_ERR_INVALID, "build_bvh: null pointer or empty mesh");
    ApiScope api_(ctx, "uvrt_build_bvh");
    Bind b(ctx);
    const bool verbose = getenv("UVRT_BVH_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tPhase = now();
    auto phase = [&](const char* name) {
        if (!verbose) return;
        cudaStreamSynchronize(ctx->stream);
        double t = now();
        fprintf(stderr, "[uvrt_build_bvh] %-10s %8.3f ms\n", name, t - tPhase);
        tPhase = t;
    };
    const size_t n = (size_t)nTris, maxNodes = 2 * n + 2;
    float4* dTris = nullptr; uint32_t *dIdxA = nullptr, *dIdxB = nullptr, *dFinal = nullptr, *dRank = nullptr, *dHole = nullptr, *dSleft = nullptr;
    uint32_t *dCounters = nullptr, *dAcc = nullptr, *dUsed = nullptr, *dChunkMap = nullptr, *dTotalChunks = nullptr;
    uint32_t* dList[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    HugeInfo* dHugeInfo[2] = {nullptr, nullptr};
    HugeState* dHugeState = nullptr; uint2 *dChunkCnt = nullptr, *dChunkOff = nullptr;
    BNode* dNodes = nullptr; BAux* dAux = nullptr; float4* dOut = nullptr;
    std::vector<void*> owned;
    // stream-ordered allocations from the device's pool: cudaMalloc/cudaFree cost tens of milliseconds
    // per build here, the pool hands the same memory back on the next build
    if (!ctx->poolTuned) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        ctx->poolTuned = true;
    }
    auto cleanup = [&]() { for (void* p : owned) cudaFreeAsync(p, ctx->stream); };
#define BALLOC(ptr, bytes)                                                                            \
    do {                                                                                              \
        cudaError_t e_ = cudaMallocAsync((void**)&(ptr), (bytes), ctx->stream);                                         \
        if (e_ != cudaSuccess) { cleanup(); return fail(ctx, UVRT_ERR_NO_MEMORY, "build_bvh: %s", cudaGetErrorString(e_)); } \
        owned.push_back((void*)(ptr));                                                                \
    } while (0)
#define BCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) { cleanup(); return fail(ctx, UVRT_ERR_CUDA, "build_bvh: %s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    // the nodes of one level own disjoint index segments, so a level has at most n / (smallest size of
    // the class) nodes of a class
    const size_t listCap[4] = {n + 1, n / (kSmall + 1) + 2, n / (kMid + 1) + 2, n / (kHuge + 1) + 2};
    const size_t maxChunks = n / kChunk + listCap[3] + 1;
    BALLOC(dTris, n * 64);
    BALLOC(dIdxA, n * 4); BALLOC(dIdxB, n * 4); BALLOC(dFinal, n * 4);
    BALLOC(dRank, n * 4); BALLOC(dHole, n * 4); BALLOC(dSleft, n * 4);
    for (int k = 0; k < 2; k++) {
        for (int c = 0; c < 4; c++) BALLOC(dList[k][c], listCap[c] * 4);
        BALLOC(dHugeInfo[k], listCap[3] * sizeof(HugeInfo));
    }
    BALLOC(dHugeState, listCap[3] * sizeof(HugeState));
    BALLOC(dChunkMap, maxChunks * 4); BALLOC(dChunkCnt, maxChunks * 8); BALLOC(dChunkOff, maxChunks * 8);
    BALLOC(dCounters, 32); BALLOC(dAcc, 64); BALLOC(dUsed, 4); BALLOC(dTotalChunks, 4);
    BALLOC(dNodes, maxNodes * sizeof(BNode)); BALLOC(dAux, maxNodes * sizeof(BAux));
    const size_t outSlots = 2 * n + 64;
    BALLOC(dOut, outSlots * 32);
    cudaStream_t st = ctx->stream;
    phase("alloc");
    // (through pinned staging in pieces: a pageable-memory copy of a 640 MB mesh runs at a quarter of the speed)
    if (ensure_stage(ctx, n * 64) != UVRT_OK) { cleanup(); return UVRT_ERR_NO_MEMORY; }
    BCK(stage_and_copy(dTris, trisHost, (char*)ctx->hStage, n * 64, st));
    BCK(cudaMemsetAsync(dAux, 0, maxNodes * sizeof(BAux), st));
    BCK(cudaMemsetAsync(dOut, 0, outSlots * 32, st));
    const uint32_t accInit[12] = {~0u, ~0u, ~0u, 0, 0, 0, ~0u, ~0u, ~0u, 0, 0, 0};
    BCK(cudaMemcpyAsync(dAcc, accInit, sizeof accInit, cudaMemcpyHostToDevice, st));
    auto lists = [&](int k) { return Lists{dList[k][0], dList[k][1], dList[k][2], dList[k][3], dHugeInfo[k], dCounters}; };
    k_centroids<<<grid_for(nTris, 256), 256, 0, st>>>(dTris, dIdxA, nTris);
    k_root_bounds<<<std::min<unsigned>(grid_for(nTris, 256), 1184u), 256, 0, st>>>(dTris, nTris, dAcc);
    k_init_root<<<1, 1, 0, st>>>(dNodes, dAcc, nTris, lists(0));
    ctx->launches += 3;
    phase("upload+root");
    // level by level; the nodes of a level are dealt to four kernels by size
    struct Level { uint32_t idBegin, idEnd; };
    std::vector<Level> levels;
    uint32_t cnt[5] = {1, 0, 0, 0, 0};          // [0] next temp id, [1..4] small / mid / big / huge nodes of this level
    cnt[n <= kSmall ? 1 : n <= kMid ? 2 : n <= kHuge ? 3 : 4] = 1;
    uint32_t idBegin = 0;
    int cur = 0;
    uint32_t *src = dIdxA, *dst = dIdxB;
    while (cnt[1] + cnt[2] + cnt[3] + cnt[4] > 0) {
        if (levels.size() > 200) { cleanup(); return fail(ctx, UVRT_ERR_INVALID, "build_bvh: tree deeper than 200 levels"); }
        levels.push_back({idBegin, cnt[0]});
        idBegin = cnt[0];
        BCK(cudaMemsetAsync(dCounters + 1, 0, 16, st));
        const Lists next = lists(1 - cur);
        const uint32_t nSmall = cnt[1], nMid = cnt[2], nBig = cnt[3], nHuge = cnt[4];
        if (nHuge) {
            const HugeInfo* info = dHugeInfo[cur];
            const unsigned chunks = (unsigned)std::min<size_t>(maxChunks, n / kChunk + nHuge);
            k_huge_prepare<<<1, 256, 0, st>>>(dHugeInfo[cur], (int)nHuge, dChunkMap, dHugeState, dTotalChunks);
            k_huge_bins<<<chunks, 256, 0, st>>>(info, dChunkMap, dTotalChunks, dNodes, dTris, src, dHugeState);
            k_huge_decide<<<grid_for(nHuge, 32), 32, 0, st>>>(info, (int)nHuge, dNodes, dHugeState, next);
            k_huge_count<<<chunks, 256, 0, st>>>(info, dChunkMap, dTotalChunks, dNodes, dTris, src, dHugeState, dChunkCnt);
            k_huge_scan<<<nHuge, 256, 0, st>>>(info, dHugeState, dChunkCnt, dChunkOff);
            k_huge_rank<<<chunks, 256, 0, st>>>(info, dChunkMap, dTotalChunks, dNodes, dTris, src, dHugeState, dChunkCnt, dChunkOff,
                                                dRank, dHole, dSleft);
            k_huge_place<<<chunks, 256, 0, st>>>(info, dChunkMap, dTotalChunks, dNodes, dTris, src, dst, dFinal, dHugeState,
                                                 dRank, dHole, dSleft);
            k_huge_finish<<<grid_for(nHuge, 32), 32, 0, st>>>(info, (int)nHuge, dNodes, dHugeState);
            ctx->launches += 8;
        }
        if (nBig)
            k_level<256><<<nBig, 256, 0, st>>>(dList[cur][2], dNodes, dTris, src, dst, dFinal, dRank, dHole, dSleft, next);
        if (nMid)
            k_level<32><<<nMid, 32, 0, st>>>(dList[cur][1], dNodes, dTris, src, dst, dFinal, dRank, dHole, dSleft, next);
        if (nSmall)
            k_level_small<<<grid_for(nSmall, 128), 128, 0, st>>>(dList[cur][0], (int)nSmall, dNodes, dTris, src, dst, dFinal, next);
        ctx->launches += (nBig ? 1 : 0) + (nMid ? 1 : 0) + (nSmall ? 1 : 0);
        BCK(cudaMemcpyAsync(cnt, dCounters, 20, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (verbose) fprintf(stderr, "[uvrt_build_bvh] level %2zu: %6u huge %7u big %8u mid %8u small ", levels.size() - 1, nHuge, nBig, nMid, nSmall);
        phase("");
        cur = 1 - cur;
        std::swap(src, dst);
    }
    const uint32_t nTemp = cnt[0];
    // renumber: subtree sizes bottom-up, reference indices top-down
    for (int d = (int)levels.size() - 1; d >= 0; d--) {
        const Level& lv = levels[d];
        k_sizes<<<grid_for(lv.idEnd - lv.idBegin, 256), 256, 0, st>>>(lv.idBegin, (int)(lv.idEnd - lv.idBegin), dNodes, dAux);
    }
    k_number_top<<<1, 1, 0, st>>>(dNodes, dAux, dUsed);
    for (size_t d = 4; d < levels.size(); d++) {
        const Level& lv = levels[d];
        k_number_level<<<grid_for(lv.idEnd - lv.idBegin, 256), 256, 0, st>>>(lv.idBegin, (int)(lv.idEnd - lv.idBegin), dNodes, dAux);
    }
    k_emit<<<grid_for(nTemp, 256), 256, 0, st>>>(dNodes, dAux, (int)nTemp, dOut);
    ctx->launches += (int64_t)levels.size() * 2 + 2;
    uint32_t used = 0;
    BCK(cudaMemcpyAsync(&used, dUsed, 4, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    phase("renumber");
    if ((long long)used > (long long)nodeCapacity) {
        cleanup();
        return fail(ctx, UVRT_ERR_INVALID, "build_bvh: %u node slots needed, nodeCapacity is %d", used, nodeCapacity);
    }
    {
        const Unstage items[3] = {{nodesOut, dOut, (size_t)used * 32}, {triIdxOut, dFinal, n * 4}, {trisOut, dTris, trisOut ? n * 64 : 0}};
        size_t total = 0;
        for (const Unstage& it : items) total += ((it.bytes + 255) & ~(size_t)255) + (32u << 20);
        if (ensure_stage(ctx, total) != UVRT_OK) { cleanup(); return UVRT_ERR_NO_MEMORY; }
        BCK(copy_and_unstage(items, 3, (char*)ctx->hStage, st));
    }
    BCK(cudaStreamSynchronize(st));
    if (nodesUsedOut) *nodesUsedOut = used;
    phase("download");
    cleanup();
    phase("free");
#undef BALLOC
#undef BCK
    return UVRT_OK;
}

int uvrt_scene_info(uvrt_ctx* ctx, int* innerNodes, int* leaves, int* depth, int* stackEntries)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "no scene uploaded");
    if (innerNodes) *innerNodes = ctx->nPairs;
    if (leaves) *leaves = ctx->nLeaves;
    if (depth) *depth = ctx->depth;
    if (stackEntries) *stackEntries = kStack;
    return UVRT_OK;
}

#define NEED_SCENE()                                                                       \
    if (!ctx) return UVRT_ERR_INVALID;                                                     \
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "%s: no scene uploaded", __func__); \
    ApiScope api_(ctx, __func__);                                                          \
    Bind bind_(ctx)

int uvrt_reset(uvrt_ctx* ctx, int resetColor)
{
    NEED_SCENE();
    {
        StageTimer t(ctx, UVRT_STAGE_RESET);
        k_reset<<<grid_for(ctx->nTris, 256), 256, 0, ctx->stream>>>(ctx->dSum, ctx->dMax, ctx->dCounts, ctx->dColor,
                                                                     resetColor, ctx->nTris);
    }
    CK(cudaMemsetAsync(ctx->dCountsAlt, 0, (size_t)ctx->nTris * 4, ctx->stream));
    ctx->launches++;
    ctx->countsDirty = false;
    CK_LAUNCH("reset");
    return UVRT_OK;
}

// generate (+ the count step of the binning) into the current ray slot, on `stream`
static int generate_on(uvrt_ctx* ctx, cudaStream_t stream, float lx, float ly, float lz, float lightLength,
                       int64_t firstRay, int64_t nRays, uint32_t seedIn)
{
    if (nRays < 0 || firstRay < 0 || nRays > 0x7fffffffll)
        return fail(ctx, UVRT_ERR_INVALID, "generate: bad ray range first=%lld n=%lld", (long long)firstRay, (long long)nRays);
    int rc = ensure_rays(ctx, nRays);
    if (rc) return rc;
    RaySlot& S = ctx->rs();
    ctx->lastRays = nRays;
    S.permRays = -1;
    S.binY0 = ly;
    S.binLen = lightLength;
    S.binExtentKnown = true;
    if (S.countedRays >= 0) {
        // slots taken by a generate whose rays were never extended
        CK(cudaMemsetAsync(S.dBinCount, 0, (size_t)S.binCap * 4, stream));
        S.countedRays = -1;
    }
    if (nRays == 0) return UVRT_OK;
    if (wants_binning(ctx, nRays)) {
        BinDims d;
        rc = bin_prepare(ctx, nRays, &d, stream);
        if (rc) return rc;
        StageTimer t(ctx, UVRT_STAGE_GENERATE, stream);
        k_generate<1><<<grid_for(nRays, 256), 256, 0, stream>>>(S.dRays, firstRay, nRays, lx, ly, lz, lightLength, seedIn,
                                                                 d, S.dBinCount, S.dKeyRank);
        S.countedRays = nRays;
    } else {
        StageTimer t(ctx, UVRT_STAGE_GENERATE, stream);
        k_generate<0><<<grid_for(nRays, 256), 256, 0, stream>>>(S.dRays, firstRay, nRays, lx, ly, lz, lightLength, seedIn,
                                                                 BinDims{1, 1, 1, 0.0f, 0.0f}, nullptr, nullptr);
    }
    ctx->launches++;
    CK_LAUNCH("generate");
    return UVRT_OK;
}

int uvrt_generate(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, int64_t firstRay, int64_t nRays,
                  uint32_t seedIn)
{
    NEED_SCENE();
    return generate_on(ctx, ctx->stream, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
}

int uvrt_extend(uvrt_ctx* ctx, int64_t nRays)
{
    NEED_SCENE();
    if (nRays < 0 || nRays > ctx->rs().rayCap)
        return fail(ctx, UVRT_ERR_INVALID, "extend: nRays=%lld exceeds the ray buffer (%lld)", (long long)nRays, ctx->rs().rayCap);
    if (nRays == 0) return UVRT_OK;
    int rc;
    const uint32_t* perm = nullptr;
    ctx->xStream = ctx->stream;
    ctx->xCounts = ctx->dCounts;
    ctx->countsDirty = true;
    // a few thousand rays are not worth the extra launches
    if (wants_binning(ctx, nRays)) {
        if (ctx->rs().permRays != nRays) {          // not already done on the generate stream
            rc = bin_finish(ctx, nRays, ctx->stream);
            if (rc) return rc;
            CK_LAUNCH("bin");
        }
        perm = ctx->rs().dPerm;
    } else if (ctx->rs().countedRays >= 0) {
        // binning was switched off between generate and extend: drop the slots generate took
        CK(cudaMemsetAsync(ctx->rs().dBinCount, 0, (size_t)ctx->rs().binCap * 4, ctx->stream));
        ctx->rs().countedRays = -1;
    }
    ctx->rs().permRays = -1;
    {
        StageTimer t(ctx, UVRT_STAGE_EXTEND);
        rc = launch_extend(ctx, nRays, perm);
    }
    if (rc) return rc;
    CK_LAUNCH("extend");
    return UVRT_OK;
}

int uvrt_accumulate(uvrt_ctx* ctx, float duration)
{
    NEED_SCENE();
    {
        StageTimer t(ctx, UVRT_STAGE_ACCUMULATE);
        k_accumulate<<<grid_for(ctx->nTris, 256), 256, 0, ctx->stream>>>(ctx->dSum, ctx->dMax, ctx->dCounts, duration, ctx->nTris);
    }
    ctx->launches++;
    ctx->countsDirty = false;
    CK_LAUNCH("accumulate");
    return UVRT_OK;
}

int uvrt_trace_counts(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, int64_t firstRay, int64_t nRays,
                      uint32_t seedIn)
{
    NEED_SCENE();
    if (!ctx->pipeline) {
        int rc = generate_on(ctx, ctx->stream, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
        if (rc) return rc;
        return uvrt_extend(ctx, nRays);
    }
    // Pipelined: this launch's rays go to the other slot and are generated on the second stream, so
    // generate (HBM-bound) overlaps the extend of the previous launch (issue-bound) still running on
    // the main stream.  Everything after generate stays on the main stream in the reference's order.
    ctx->slot ^= 1;
    RaySlot& S = ctx->rs();
    if (S.inFlight) CK(cudaStreamWaitEvent(ctx->genStream, S.freeEv, 0));   // the extend that last read this slot
    int rc = generate_on(ctx, ctx->genStream, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
    if (rc) return rc;
    if (wants_binning(ctx, nRays) && S.countedRays == nRays) {
        // scan + scatter of the ray binning also run ahead, next to the previous launch's extend
        rc = bin_finish(ctx, nRays, ctx->genStream);
        if (rc) return rc;
        CK_LAUNCH("bin");
        S.permRays = nRays;
    }
    CK(cudaEventRecord(S.genDone, ctx->genStream));
    CK(cudaStreamWaitEvent(ctx->stream, S.genDone, 0));
    rc = uvrt_extend(ctx, nRays);
    if (rc) return rc;
    CK(cudaEventRecord(S.freeEv, ctx->stream));
    S.inFlight = true;
    return UVRT_OK;
}

int uvrt_trace(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, float duration, int64_t firstRay,
               int64_t nRays, uint32_t seedIn)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "%s: no scene uploaded", __func__);
    ApiScope api_(ctx, "uvrt_trace");
    const bool foreign = ctx->mainForeign;
    const int variant = ctx->extendVariant < 0 ? default_variant(ctx) : ctx->extendVariant;
    const bool sharedQueue = variant >= 10 && variant < 25;   // variant B's global queue head is one per context
    // counts left behind by uvrt_extend / uvrt_trace_counts belong to this launch's accumulate as well: they sit
    // in the primary buffer, so this launch must use it too
    if (!ctx->pipeline || !ctx->overlapExtend || nRays <= 0 || sharedQueue || ctx->countsDirty) {
        int rc = uvrt_trace_counts(ctx, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
        if (rc) return rc;
        return uvrt_accumulate(ctx, duration);
    }
    // Overlapped: generate + bin on the generate stream, extend on the slot's own stream into the slot's
    // own count buffer, accumulate on the main stream.  Dependencies (events):
    //   generate k   after extend k-2      (same ray slot)
    //   extend k     after generate/bin k, after accumulate k-2 (it zeroes the slot's counts), and after
    //                whatever other calls put on the main stream since the last trace (reset, upload, ...)
    //   accumulate k after extend k; accumulates run in launch order on one (high-priority) stream, so the
    //                f64 sums are formed in the reference's order; the main stream waits for each of them
    // Extend k+1 therefore only waits for data, not for extend k: its blocks fill the SMs that extend k's
    // last wave leaves idle.
    Bind b(ctx);
    ctx->slot ^= 1;
    const int si = ctx->slot;
    RaySlot& S = ctx->rs();
    cudaStream_t es = ctx->extStream[si];
    if (foreign) {
        CK(cudaEventRecord(ctx->forkEv, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->extStream[0], ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->extStream[1], ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->accStream, ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->genStream, ctx->forkEv, 0));
    }
    if (S.inFlight) CK(cudaStreamWaitEvent(ctx->genStream, S.freeEv, 0));
    int rc = generate_on(ctx, ctx->genStream, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
    if (rc) return rc;
    const uint32_t* perm = nullptr;
    if (wants_binning(ctx, nRays)) {
        rc = bin_finish(ctx, nRays, ctx->genStream);
        if (rc) return rc;
        CK_LAUNCH("bin");
        perm = S.dPerm;
    }
    CK(cudaEventRecord(S.genDone, ctx->genStream));
    CK(cudaStreamWaitEvent(es, S.genDone, 0));
    if (ctx->accUsed[si]) CK(cudaStreamWaitEvent(es, ctx->accDone[si], 0));
    ctx->xStream = es;
    ctx->xCounts = si ? ctx->dCountsAlt : ctx->dCounts;
    {
        StageTimer t(ctx, UVRT_STAGE_EXTEND, es);
        rc = launch_extend(ctx, nRays, perm);
    }
    ctx->xStream = ctx->stream;
    int* counts = ctx->xCounts;
    ctx->xCounts = ctx->dCounts;
    if (rc) return rc;
    CK_LAUNCH("extend");
    CK(cudaEventRecord(ctx->extDone[si], es));
    CK(cudaEventRecord(S.freeEv, es));
    S.inFlight = true;
    ctx->extUsed[si] = true;
    CK(cudaStreamWaitEvent(ctx->accStream, ctx->extDone[si], 0));
    {
        StageTimer t(ctx, UVRT_STAGE_ACCUMULATE, ctx->accStream);
        k_accumulate<<<grid_for(ctx->nTris, 256), 256, 0, ctx->accStream>>>(ctx->dSum, ctx->dMax, counts, duration, ctx->nTris);
    }
    ctx->launches++;
    CK_LAUNCH("accumulate");
    CK(cudaEventRecord(ctx->accDone[si], ctx->accStream));
    ctx->accUsed[si] = true;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->accDone[si], 0));   // everything later on the main stream sees the maps
    ctx->mainForeign = false;    // (Bind set it) nothing but this trace's accumulate is new on the main stream
    return UVRT_OK;
}

// Relative cost of a launch at a lamp position: inner-node visits and triangle tests per ray over the first
// nRays rays of the launch, traversed in reference order (k_probe_cost).  Deterministic, so every rank of a sharded
// run computes the same numbers and deals the launches alike without an exchange.  Synchronises.
int uvrt_probe_cost(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, uint32_t seedIn, int nRays, double* innerPerRay,
                    double* testsPerRay)
{
    NEED_SCENE();
    if (nRays < 1 || nRays > (1 << 24)) return fail(ctx, UVRT_ERR_INVALID, "probe_cost: nRays = %d", nRays);
    unsigned long long* d = (unsigned long long*)ctx->dQueue;      // 256 bytes of scratch owned by the context
    CK(cudaMemsetAsync(d + 8, 0, 16, ctx->stream));
    k_probe_cost<<<grid_for(nRays, 128), 128, 0, ctx->stream>>>(ctx->dPairs, ctx->dWtris, ctx->rootRef, nRays, lx, ly, lz, lightLength, seedIn, d + 8);
    ctx->launches++;
    CK_LAUNCH("probe_cost");
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, d + 8, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (innerPerRay) *innerPerRay = (double)h[0] / nRays;
    if (testsPerRay) *testsPerRay = (double)h[1] / nRays;
    return UVRT_OK;
}

// ---- count matrix: launches of a run write their integer counts to one row each -----------------------
// For runs whose launches are shared between GPUs (whole launches or ray ranges of a launch): every rank
// traces its rays of launch k into row k, ONE all-reduce sums the integer rows, and the fold replays
// accumulate.cl:4-14 row by row in launch order on every rank -- the f64 sums and the per-launch maxima are
// then bit-identical to the single-GPU run for ANY split and ANY durations.
// Capacity for windows of up to `rows` rows in both buffers, so that no later uvrt_matrix_begin has to allocate
// (an allocation synchronises the device -- the wrong thing to happen inside somebody's timed run).
int uvrt_matrix_reserve(uvrt_ctx* ctx, int rows)
{
    NEED_SCENE();
    if (rows < 1) return fail(ctx, UVRT_ERR_INVALID, "matrix_reserve: rows = %d", rows);
    const size_t need = (size_t)rows * (size_t)ctx->nTris;
    for (int sel = 0; sel < 2; sel++) {
        if (need > ctx->matrixCap[sel]) {
            int* fresh = nullptr;
            CK(cudaMalloc((void**)&fresh, need * 4));
            if (ctx->matrixBuf[sel]) cudaFree(ctx->matrixBuf[sel]);
            if (ctx->dMatrix == ctx->matrixBuf[sel]) ctx->dMatrix = fresh;
            ctx->matrixBuf[sel] = fresh;
            ctx->matrixCap[sel] = need;
            ctx->matrixFoldedUsed[sel] = false;
        }
        if (rows > ctx->durCap[sel]) {
            float* fresh = nullptr;
            CK(cudaMalloc((void**)&fresh, (size_t)rows * 4));
            if (ctx->durBuf[sel]) cudaFree(ctx->durBuf[sel]);
            ctx->durBuf[sel] = fresh;
            ctx->durCap[sel] = rows;
        }
    }
    return UVRT_OK;
}

int uvrt_matrix_begin(uvrt_ctx* ctx, int rows)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "%s: no scene uploaded", __func__);
    const bool foreign = ctx->mainForeign;
    ApiScope api_(ctx, "uvrt_matrix_begin");
    Bind bind_(ctx);
    if (rows < 1) return fail(ctx, UVRT_ERR_INVALID, "matrix_begin: rows = %d", rows);
    const size_t need = (size_t)rows * (size_t)ctx->nTris;
    const int sel = ctx->matrixSel ^ 1;
    for (int k = 0; k < 2; k++) {
        if (!ctx->matrixReady[k]) CK(cudaEventCreateWithFlags(&ctx->matrixReady[k], cudaEventDisableTiming));
        if (!ctx->matrixFolded[k]) CK(cudaEventCreateWithFlags(&ctx->matrixFolded[k], cudaEventDisableTiming));
    }
    if (need > ctx->matrixCap[sel]) {
        int* fresh = nullptr;
        CK(cudaMalloc((void**)&fresh, need * 4));
        if (ctx->matrixBuf[sel]) cudaFree(ctx->matrixBuf[sel]);     // (synchronises the device: nothing reads it any more)
        ctx->matrixBuf[sel] = fresh;
        ctx->matrixCap[sel] = need;
        ctx->matrixFoldedUsed[sel] = false;
    }
    if (rows > ctx->durCap[sel]) {
        float* fresh = nullptr;
        CK(cudaMalloc((void**)&fresh, (size_t)rows * 4));
        if (ctx->durBuf[sel]) cudaFree(ctx->durBuf[sel]);
        ctx->durBuf[sel] = fresh;
        ctx->durCap[sel] = rows;
    }
    // Zeroed on the (otherwise idle) accumulate stream: after the fold that last read this buffer (two windows ago) and,
    // when something other than matrix calls happened on the main stream since the last trace (reset, upload), after
    // that; the extend streams wait for the zeroing only -- not for the main stream, where the previous window's
    // all-reduce and fold may still be running.
    if (foreign) {
        CK(cudaEventRecord(ctx->forkEv, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->accStream, ctx->forkEv, 0));
    }
    if (ctx->matrixFoldedUsed[sel]) CK(cudaStreamWaitEvent(ctx->accStream, ctx->matrixFolded[sel], 0));
    // rows of this buffer may still be written by extends of the window before the previous one only if the caller never
    // folded it; waiting for the extend streams' last events covers that case too
    for (int k = 0; k < 2; k++)
        if (ctx->extUsed[k] && !ctx->matrixFoldedUsed[sel]) CK(cudaStreamWaitEvent(ctx->accStream, ctx->extDone[k], 0));
    CK(cudaMemsetAsync(ctx->matrixBuf[sel], 0, need * 4, ctx->accStream));
    CK(cudaEventRecord(ctx->matrixReady[sel], ctx->accStream));
    CK(cudaStreamWaitEvent(ctx->extStream[0], ctx->matrixReady[sel], 0));
    CK(cudaStreamWaitEvent(ctx->extStream[1], ctx->matrixReady[sel], 0));
    ctx->matrixSel = sel;
    ctx->dMatrix = ctx->matrixBuf[sel];
    ctx->matrixRows = rows;
    ctx->mainForeign = foreign;          // a matrix call is not a reason for the next trace to wait for the main stream
    return UVRT_OK;
}

int uvrt_trace_row(uvrt_ctx* ctx, int row, float lx, float ly, float lz, float lightLength, int64_t firstRay, int64_t nRays,
                   uint32_t seedIn)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "%s: no scene uploaded", __func__);
    if (row < 0 || row >= ctx->matrixRows) return fail(ctx, UVRT_ERR_INVALID, "trace_row: row %d outside the matrix (%d rows)", row, ctx->matrixRows);
    if (nRays <= 0) return nRays == 0 ? UVRT_OK : fail(ctx, UVRT_ERR_INVALID, "trace_row: nRays = %lld", (long long)nRays);
    ApiScope api_(ctx, "uvrt_trace_row");
    const bool foreign = ctx->mainForeign;
    Bind b(ctx);
    // the stream choreography of the overlapped uvrt_trace without its accumulates: generate + bin on the
    // generate stream, extend on the ray slot's own stream; rows are disjoint, so extends never wait for each other
    ctx->slot ^= 1;
    const int si = ctx->slot;
    RaySlot& S = ctx->rs();
    cudaStream_t es = ctx->extStream[si];
    if (foreign) {
        CK(cudaEventRecord(ctx->forkEv, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->extStream[0], ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->extStream[1], ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->accStream, ctx->forkEv, 0));
        CK(cudaStreamWaitEvent(ctx->genStream, ctx->forkEv, 0));
    }
    if (S.inFlight) CK(cudaStreamWaitEvent(ctx->genStream, S.freeEv, 0));
    int rc = generate_on(ctx, ctx->genStream, lx, ly, lz, lightLength, firstRay, nRays, seedIn);
    if (rc) return rc;
    const uint32_t* perm = nullptr;
    if (wants_binning(ctx, nRays)) {
        rc = bin_finish(ctx, nRays, ctx->genStream);
        if (rc) return rc;
        CK_LAUNCH("bin");
        perm = S.dPerm;
    }
    CK(cudaEventRecord(S.genDone, ctx->genStream));
    CK(cudaStreamWaitEvent(es, S.genDone, 0));
    ctx->xStream = es;
    ctx->xCounts = ctx->dMatrix + (size_t)row * (size_t)ctx->nTris;
    {
        StageTimer t(ctx, UVRT_STAGE_EXTEND, es);
        rc = launch_extend(ctx, nRays, perm);
    }
    ctx->xStream = ctx->stream;
    ctx->xCounts = ctx->dCounts;
    if (rc) return rc;
    CK_LAUNCH("extend");
    CK(cudaEventRecord(ctx->extDone[si], es));
    CK(cudaEventRecord(S.freeEv, es));
    S.inFlight = true;
    ctx->extUsed[si] = true;
    ctx->mainForeign = false;
    return UVRT_OK;
}

int uvrt_matrix_fold(uvrt_ctx* ctx, const float* durations, int rows, int reduce)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (!ctx->nTris) return fail(ctx, UVRT_ERR_NO_SCENE, "%s: no scene uploaded", __func__);
    const bool foreign = ctx->mainForeign;
    ApiScope api_(ctx, "uvrt_matrix_fold");
    Bind bind_(ctx);
    if (rows < 0 || rows > ctx->matrixRows || (rows > 0 && !durations))
        return fail(ctx, UVRT_ERR_INVALID, "matrix_fold: rows = %d of %d", rows, ctx->matrixRows);
    const int sel = ctx->matrixSel;
    // the extends of this window are the last work on the extend streams (the next window's are not enqueued yet)
    for (int k = 0; k < 2; k++)
        if (ctx->extUsed[k]) CK(cudaStreamWaitEvent(ctx->stream, ctx->extDone[k], 0));
    if (ctx->matrixReady[sel]) CK(cudaStreamWaitEvent(ctx->stream, ctx->matrixReady[sel], 0));
    if (rows == 0) { ctx->mainForeign = foreign; return UVRT_OK; }
    if (reduce && (ctx->nRanks > 1 || ctx->comm)) {
        if (!ctx->comm) return fail(ctx, UVRT_ERR_NCCL, "matrix_fold: uvrt_comm_init was not called");
        nvtxRangePushA("ncclAllReduce(count matrix)");
        int r = g_nccl.AllReduce(ctx->dMatrix, ctx->dMatrix, (size_t)rows * (size_t)ctx->nTris, kNcclInt32, kNcclSum, ctx->comm, ctx->stream);
        nvtxRangePop();
        if (r != kNcclSuccess) return fail(ctx, UVRT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    }
    // (pageable source: staged before the call returns, so the caller may reuse `durations` at once)
    CK(cudaMemcpyAsync(ctx->durBuf[sel], durations, (size_t)rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, UVRT_STAGE_ACCUMULATE);
        k_fold_rows<<<grid_for(ctx->nTris, 256), 256, 0, ctx->stream>>>(ctx->dSum, ctx->dMax, ctx->dMatrix, ctx->durBuf[sel], rows, ctx->nTris);
    }
    ctx->launches++;
    CK_LAUNCH("fold");
    CK(cudaEventRecord(ctx->matrixFolded[sel], ctx->stream));
    ctx->matrixFoldedUsed[sel] = true;
    ctx->mainForeign = foreign;
    return UVRT_OK;
}

// ---- timeline ("timeline" option): where the wall clock of a run goes ------------------------------------
int uvrt_timeline_dump(uvrt_ctx* ctx, const char* path)
{
    if (!ctx || !path) return UVRT_ERR_INVALID;
    Bind b(ctx);
    CK(uvrt_sync(ctx) == UVRT_OK ? cudaSuccess : cudaErrorUnknown);
    FILE* f = fopen(path, "w");
    if (!f) return fail(ctx, UVRT_ERR_IO, "timeline_dump: cannot write %s", path);
    fprintf(f, "{\"unit\": \"us since the option was switched on (host clock for calls, device clock for kernels)\",\n \"calls\": [");
    for (size_t i = 0; i < ctx->apiLog.size(); i++)
        fprintf(f, "%s[\"%s\", %.1f, %.1f]", i ? ", " : "", ctx->apiLog[i].name, ctx->apiLog[i].t0, ctx->apiLog[i].t1);
    fprintf(f, "],\n \"kernels\": [");
    bool first = true;
    for (size_t i = 0; i < ctx->timed.size() && i < ctx->timedHost.size(); i++) {
        float a = 0, z = 0;
        if (!ctx->timelineOrigin || cudaEventElapsedTime(&a, ctx->timelineOrigin, ctx->timed[i].start) != cudaSuccess ||
            cudaEventElapsedTime(&z, ctx->timelineOrigin, ctx->timed[i].stop) != cudaSuccess) { cudaGetLastError(); continue; }
        const int st = ctx->timed[i].stage;
        fprintf(f, "%s[\"%s\", %.1f, %.1f, %.1f]", first ? "" : ", ", st >= 0 && st < UVRT_STAGE_COUNT ? kStageNames[st] : "?", ctx->timedHost[i],
                a * 1e3, z * 1e3);
        first = false;
    }
    fprintf(f, "]}\n");
    fclose(f);
    return UVRT_OK;
}

int uvrt_seed_chain(uvrt_ctx* ctx, const float* lightPos3, int nLaunches, float lightLength, uint32_t seedIn,
                    uint32_t* seedsOut)
{
    if (!ctx) return UVRT_ERR_INVALID;
    if (nLaunches < 0 || !seedsOut || (nLaunches > 0 && !lightPos3))
        return fail(ctx, UVRT_ERR_INVALID, "seed_chain: bad arguments");
    ApiScope api_(ctx, "uvrt_seed_chain");
    Bind b(ctx);
    if (nLaunches + 1 > ctx->seedCap) {
        int cap = std::max(nLaunches + 1, 256);
        if (ctx->dSeeds) cudaFree(ctx->dSeeds);
        if (ctx->dSeedPos) cudaFree(ctx->dSeedPos);
        ctx->dSeeds = nullptr; ctx->dSeedPos = nullptr; ctx->seedCap = 0;
        CK(cudaMalloc((void**)&ctx->dSeeds, (size_t)cap * 4));
        CK(cudaMalloc((void**)&ctx->dSeedPos, (size_t)cap * 12));
        ctx->seedCap = cap;
    }
    if (nLaunches > 0)
        CK(cudaMemcpyAsync(ctx->dSeedPos, lightPos3, (size_t)nLaunches * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_seed_chain<<<1, 32, 0, ctx->stream>>>(ctx->dSeedPos, nLaunches, lightLength, seedIn, ctx->dSeeds);
    ctx->launches++;
    CK_LAUNCH("seed_chain");
    CK(cudaMemcpyAsync(seedsOut, ctx->dSeeds, (size_t)(nLaunches + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UVRT_OK;
}

int uvrt_shade(uvrt_ctx* ctx, int useMaxMap, int photonsPerLight, float scaledPower)
{
    NEED_SCENE();
    {
        StageTimer t(ctx, UVRT_STAGE_SHADE);
        k_compute_dosage<<<grid_for(ctx->nTris, 256), 256, 0, ctx->stream>>>(useMaxMap ? ctx->dMax : ctx->dSum, ctx->dDose,
                                                                              ctx->dVerts, photonsPerLight, scaledPower, ctx->nTris);
    }
    ctx->launches++;
    CK_LAUNCH("computeDosage");
    return UVRT_OK;
}

int uvrt_color(uvrt_ctx* ctx, float minValue, int thresholdView)
{
    NEED_SCENE();
    {
        StageTimer t(ctx, UVRT_STAGE_COLOR);
        k_dosage_to_color<<<grid_for(ctx->nTris, 256), 256, 0, ctx->stream>>>(ctx->dDose, ctx->dColor, minValue, thresholdView, ctx->nTris);
    }
    ctx->launches++;
    CK_LAUNCH("dosageToColor");
    return UVRT_OK;
}

int uvrt_read(uvrt_ctx* ctx, uvrt_buffer what, void* dst, size_t bytes)
{
    NEED_SCENE();
    void* p = nullptr;
    size_t cap = 0;
    int rc = check_buffer(ctx, what, &p, &cap);
    if (rc) return rc;
    if (!dst || bytes > cap) return fail(ctx, UVRT_ERR_INVALID, "read: %zu bytes requested, buffer %d holds %zu", bytes, (int)what, cap);
    if (what == UVRT_BUF_MATRIX) {    // rows are zeroed on the accumulate stream and written by the extend streams of uvrt_trace_row
        if (ctx->matrixReady[ctx->matrixSel]) CK(cudaStreamWaitEvent(ctx->stream, ctx->matrixReady[ctx->matrixSel], 0));
        for (int k = 0; k < 2; k++)
            if (ctx->extUsed[k]) CK(cudaStreamWaitEvent(ctx->stream, ctx->extDone[k], 0));
    }
    if (bytes) CK(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UVRT_OK;
}

int uvrt_write(uvrt_ctx* ctx, uvrt_buffer what, const void* src, size_t bytes)
{
    NEED_SCENE();
    if (what == UVRT_BUF_PAIRS || what == UVRT_BUF_WTRIS)
        return fail(ctx, UVRT_ERR_INVALID, "write: buffer %d is read only (set by uvrt_upload_scene)", (int)what);
    if (what == UVRT_BUF_RAYS) {
        int rc = ensure_rays(ctx, (long long)((bytes + 31) / 32));
        if (rc) return rc;
        ctx->lastRays = (long long)(bytes / 32);
        ctx->rs().binExtentKnown = false;   // foreign rays: no origin slicing
        if (ctx->rs().countedRays >= 0 && ctx->rs().dBinCount) {
            CK(cudaMemsetAsync(ctx->rs().dBinCount, 0, (size_t)ctx->rs().binCap * 4, ctx->stream));
            ctx->rs().countedRays = -1;
        }
    }
    void* p = nullptr;
    size_t cap = 0;
    int rc = check_buffer(ctx, what, &p, &cap);
    if (rc) return rc;
    if (!src || bytes > cap) return fail(ctx, UVRT_ERR_INVALID, "write: %zu bytes offered, buffer %d holds %zu", bytes, (int)what, cap);
    if (what == UVRT_BUF_MATRIX) {
        if (ctx->matrixReady[ctx->matrixSel]) CK(cudaStreamWaitEvent(ctx->stream, ctx->matrixReady[ctx->matrixSel], 0));
        for (int k = 0; k < 2; k++)
            if (ctx->extUsed[k]) CK(cudaStreamWaitEvent(ctx->stream, ctx->extDone[k], 0));
    }
    if (bytes) CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (what == UVRT_BUF_COUNTS) ctx->countsDirty = true;
    return UVRT_OK;
}

int uvrt_sync(uvrt_ctx* ctx)
{
    if (!ctx) return UVRT_ERR_INVALID;
    ApiScope api_(ctx, "uvrt_sync");
    Bind b(ctx);
    CK(cudaStreamSynchronize(ctx->genStream));
    CK(cudaStreamSynchronize(ctx->extStream[0]));
    CK(cudaStreamSynchronize(ctx->extStream[1]));
    CK(cudaStreamSynchronize(ctx->accStream));
    CK(cudaStreamSynchronize(ctx->stream));
    return UVRT_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
int uvrt_comm_unique_id(void* id128)
{
    if (!id128) return UVRT_ERR_INVALID;
    if (!nccl_load()) return fail(nullptr, UVRT_ERR_NCCL, "NCCL unavailable: %s", g_nccl.why.c_str());
    ncclUniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != kNcclSuccess) return fail(nullptr, UVRT_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, 128);
    return UVRT_OK;
}

int uvrt_comm_init(uvrt_ctx* ctx, const void* id128, int rank, int nRanks)
{
    if (!ctx || !id128 || nRanks < 1 || rank < 0 || rank >= nRanks) return ctx ? fail(ctx, UVRT_ERR_INVALID, "comm_init: bad arguments") : UVRT_ERR_INVALID;
    if (!nccl_load()) return fail(ctx, UVRT_ERR_NCCL, "NCCL unavailable: %s", g_nccl.why.c_str());
    Bind b(ctx);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
    int r = g_nccl.CommInitRank(&ctx->comm, nRanks, id, rank);
    if (r != kNcclSuccess) return fail(ctx, UVRT_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    ctx->rank = rank;
    ctx->nRanks = nRanks;
    if (nRanks > 1) g_copyThreads = std::max(1, std::min(16, (int)std::thread::hardware_concurrency() / nRanks));
    return UVRT_OK;
}

int uvrt_reduce(uvrt_ctx* ctx)
{
    NEED_SCENE();
    if (ctx->nRanks == 1 && !ctx->comm) return UVRT_OK;
    if (!ctx->comm) return fail(ctx, UVRT_ERR_NCCL, "reduce: uvrt_comm_init was not called");
    int r = g_nccl.GroupStart();
    if (r == kNcclSuccess) r = g_nccl.AllReduce(ctx->dSum, ctx->dSum, (size_t)ctx->nTris, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (r == kNcclSuccess) r = g_nccl.AllReduce(ctx->dMax, ctx->dMax, (size_t)ctx->nTris, kNcclFloat64, kNcclMax, ctx->comm, ctx->stream);
    int r2 = g_nccl.GroupEnd();
    if (r == kNcclSuccess) r = r2;
    if (r != kNcclSuccess) return fail(ctx, UVRT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    return UVRT_OK;
}

int uvrt_reduce_counts(uvrt_ctx* ctx)
{
    NEED_SCENE();
    if (ctx->nRanks == 1 && !ctx->comm) return UVRT_OK;
    if (!ctx->comm) return fail(ctx, UVRT_ERR_NCCL, "reduce_counts: uvrt_comm_init was not called");
    int r = g_nccl.AllReduce(ctx->dCounts, ctx->dCounts, (size_t)ctx->nTris, kNcclInt32, kNcclSum, ctx->comm, ctx->stream);
    if (r != kNcclSuccess) return fail(ctx, UVRT_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
    return UVRT_OK;
}

// ---- options / measurement ---------------------------------------------------------------------------
int uvrt_set_option(uvrt_ctx* ctx, const char* key, int value)
{
    if (!ctx || !key) return UVRT_ERR_INVALID;
    if (!strcmp(key, "extend_variant")) ctx->extendVariant = value;
    else if (!strcmp(key, "stage_timing")) {
        ctx->stageTiming = value;
        if (value && ctx->freeEvents.size() < 8192) {
            // two events per stage launch: create them now, not inside the region being timed (cudaEventCreate costs
            // microseconds, more when eight processes share the driver: 16 ms of a 126 ms timed run at 8 GPUs)
            Bind b(ctx);
            while (ctx->freeEvents.size() < 8192) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); break; }
                ctx->freeEvents.push_back(e);
            }
        }
    }
    else if (!strcmp(key, "timeline")) {
        Bind b(ctx);
        uvrt_stage_time_reset(ctx);
        ctx->apiLog.clear();
        ctx->timedHost.clear();
        ctx->timeline = value;
        if (value) {
            if (!ctx->timelineOrigin) CK(cudaEventCreate(&ctx->timelineOrigin));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaEventRecord(ctx->timelineOrigin, ctx->stream));
            CK(cudaEventSynchronize(ctx->timelineOrigin));
            ctx->timelineHost0 = host_now_us();
        }
    }
    else if (!strcmp(key, "fast_check")) ctx->fastCheck = value;
    else if (!strcmp(key, "fast_cfg")) {
#ifdef UVRT_EXPERIMENTS
        const int hi = 2;
#else
        const int hi = 1;
#endif
        if (value < 0 || value > hi)
            return fail(ctx, UVRT_ERR_INVALID, "fast_cfg must be 0 or 1 (2, the refill kernel, exists only in builds with -DUVRT_EXPERIMENTS)");
        ctx->fastCfg = value;
    }
#ifdef UVRT_EXPERIMENTS
    else if (!strcmp(key, "refill_chunk")) {
        if (value < 1 || value > 65536) return fail(ctx, UVRT_ERR_INVALID, "refill_chunk must be in 1..65536");
        ctx->refillChunk = value;
    }
#endif
    else if (!strcmp(key, "hist_mode")) ctx->histMode = value;
    else if (!strcmp(key, "blocks_per_sm")) ctx->blocksPerSm = value;
    else if (!strcmp(key, "refill")) ctx->refill = value;
    else if (!strcmp(key, "simple_cfg")) ctx->simpleCfg = value;
    else if (!strcmp(key, "generic_octant")) ctx->genericOctant = value;
    else if (!strcmp(key, "carveout")) ctx->carveout = value;
    else if (!strcmp(key, "chunk")) ctx->chunk = value;
    else if (!strcmp(key, "pipeline")) ctx->pipeline = value;
    else if (!strcmp(key, "overlap_extend")) ctx->overlapExtend = value;
    else if (!strcmp(key, "host_repack")) ctx->hostRepack = value;
    else if (!strcmp(key, "fetch_mode")) ctx->fetchMode = value;
    else if (!strncmp(key, "bin_", 4)) {
        if (!strcmp(key, "bin_rays")) ctx->binRays = value;
        else if (!strcmp(key, "bin_y") && value >= 1 && value <= 64) ctx->binY = value;
        else if (!strcmp(key, "bin_t") && value >= 1 && value <= 1024) ctx->binT = value;
        else if (!strcmp(key, "bin_p") && value >= 1 && value <= 1024) ctx->binP = value;
        else return fail(ctx, UVRT_ERR_INVALID, "unknown option '%s' (or value %d out of range)", key, value);
        if (ctx->rs().countedRays >= 0 && ctx->rs().dBinCount) {
            // bin slots taken by the last generate belong to the old geometry: drop them
            Bind b(ctx);
            CK(cudaMemsetAsync(ctx->rs().dBinCount, 0, (size_t)ctx->rs().binCap * 4, ctx->stream));
            ctx->rs().countedRays = -1;
        }
    }
    else return fail(ctx, UVRT_ERR_INVALID, "unknown option '%s' (or value %d out of range)", key, value);
    return UVRT_OK;
}

int uvrt_get_option(uvrt_ctx* ctx, const char* key, int* value)
{
    if (!ctx || !key || !value) return UVRT_ERR_INVALID;
    if (!strcmp(key, "extend_variant")) *value = ctx->extendVariant < 0 ? default_variant(ctx) : ctx->extendVariant;
    else if (!strcmp(key, "stage_timing")) *value = ctx->stageTiming;
    else if (!strcmp(key, "timeline")) *value = ctx->timeline;
    else if (!strcmp(key, "hist_mode")) *value = ctx->histMode;
    else if (!strcmp(key, "blocks_per_sm")) *value = ctx->blocksPerSm;
    else if (!strcmp(key, "scene_tame")) *value = ctx->sceneTame;
    else if (!strcmp(key, "scene_nested")) *value = ctx->sceneNested;
    else if (!strcmp(key, "fast_ready")) *value = fast_usable(ctx) ? 1 : 0;
    else if (!strcmp(key, "fast_check")) *value = ctx->fastCheck;
    else if (!strcmp(key, "fast_cfg")) *value = ctx->fastCfg;
#ifdef UVRT_EXPERIMENTS
    else if (!strcmp(key, "refill_chunk")) *value = ctx->refillChunk;
#endif
    else if (!strcmp(key, "experiments")) {
#ifdef UVRT_EXPERIMENTS
        *value = 1;
#else
        *value = 0;
#endif
    }
    else if (!strcmp(key, "refill")) *value = ctx->refill;
    else if (!strcmp(key, "simple_cfg")) *value = ctx->simpleCfg;
    else if (!strcmp(key, "generic_octant")) *value = ctx->genericOctant;
    else if (!strcmp(key, "chunk")) *value = ctx->chunk;
    else if (!strcmp(key, "pipeline")) *value = ctx->pipeline;
    else if (!strcmp(key, "overlap_extend")) *value = ctx->overlapExtend;
    else if (!strcmp(key, "host_repack")) *value = ctx->hostRepack;
    else if (!strcmp(key, "fetch_mode")) *value = ctx->fetchMode;
    else if (!strcmp(key, "bin_rays")) *value = ctx->binRays;
    else if (!strcmp(key, "bin_y")) *value = ctx->binY;
    else if (!strcmp(key, "bin_t")) *value = ctx->binT;
    else if (!strcmp(key, "bin_p")) *value = ctx->binP;
    else return fail(ctx, UVRT_ERR_INVALID, "unknown option '%s'", key);
    return UVRT_OK;
}

int uvrt_stage_time(uvrt_ctx* ctx, uvrt_stage stage, double* ms, int64_t* launches)
{
    if (!ctx) return UVRT_ERR_INVALID;
    Bind b(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    double sum = 0;
    int64_t n = 0;
    // Launches of one stage may overlap (extend k+1 starts in the tail of extend k when "overlap_extend"
    // is on): a launch is charged from the later of its own start and its predecessor's stop, so the sum
    // is the time during which the stage was running at all.
    const TimedLaunch* prev = nullptr;
    for (auto& t : ctx->timed) {
        if (t.stage != (int)stage) continue;
        float f = 0;
        CK(cudaEventElapsedTime(&f, t.start, t.stop));
        if (prev) {
            float ov = 0;
            if (cudaEventElapsedTime(&ov, t.start, prev->stop) == cudaSuccess && ov > 0.0f) f -= std::min(ov, f);
            else cudaGetLastError();
        }
        sum += f;
        n++;
        prev = &t;
    }
    if (ms) *ms = sum;
    if (launches) *launches = n;
    return UVRT_OK;
}

int uvrt_stage_time_reset(uvrt_ctx* ctx)
{
    if (!ctx) return UVRT_ERR_INVALID;
    Bind b(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& t : ctx->timed) { ctx->freeEvents.push_back(t.start); ctx->freeEvents.push_back(t.stop); }
    ctx->timed.clear();
    ctx->timedHost.clear();
    return UVRT_OK;
}

int uvrt_flush_l2(uvrt_ctx* ctx)
{
    if (!ctx) return UVRT_ERR_INVALID;
    Bind b(ctx);
    const size_t bytes = 256u << 20;
    if (!ctx->dFlush) CK(cudaMalloc(&ctx->dFlush, bytes));
    CK(cudaMemsetAsync(ctx->dFlush, 0xA5, bytes, ctx->stream));
    return UVRT_OK;
}

int64_t uvrt_scene_upload_bytes(const uvrt_ctx* ctx) { return ctx ? ctx->uploadBytes : 0; }

int64_t uvrt_launch_count(const uvrt_ctx* ctx) { return ctx ? ctx->launches : 0; }

int uvrt_mark(uvrt_ctx* ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= 16) return UVRT_ERR_INVALID;
    Bind b(ctx);
    // rows of a count matrix are still being written on the extend streams (uvrt_trace_row): a mark covers them
    for (int k = 0; k < 2; k++)
        if (ctx->extUsed[k]) CK(cudaStreamWaitEvent(ctx->stream, ctx->extDone[k], 0));
    CK(cudaEventRecord(ctx->marks[slot], ctx->stream));
    // work of later calls on the second stream must not start before the mark
    CK(cudaStreamWaitEvent(ctx->genStream, ctx->marks[slot], 0));
    return UVRT_OK;
}

int uvrt_elapsed_ms(uvrt_ctx* ctx, int a, int b2, float* ms)
{
    if (!ctx || !ms || a < 0 || a >= 16 || b2 < 0 || b2 >= 16) return UVRT_ERR_INVALID;
    Bind b(ctx);
    CK(cudaEventSynchronize(ctx->marks[b2]));
    CK(cudaEventElapsedTime(ms, ctx->marks[a], ctx->marks[b2]));
    return UVRT_OK;
}

// Counters of the certified fast extend since the last reset: out3 = {rays traced again because the certificate
// failed, rays not eligible for the fast path, certified rays that differed from the exact traversal ("fast_check")}.
int uvrt_fast_stats(uvrt_ctx* ctx, unsigned long long* out3, int reset)
{
    if (!ctx || !out3) return UVRT_ERR_INVALID;
    Bind b(ctx);
    out3[0] = out3[1] = out3[2] = 0;
    if (!ctx->dFastStats) return UVRT_OK;
    CK(uvrt_sync(ctx) == UVRT_OK ? cudaSuccess : cudaErrorUnknown);
    CK(cudaMemcpy(out3, ctx->dFastStats, 24, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(ctx->dFastStats, 0, sizeof(FastStats)));
    return UVRT_OK;
}

// Diagnostic: on-device check of the shared-reciprocal division against __fdiv_rn.
// out3 = {samples, one-step mismatches, two-step mismatches}.
int uvrt_selftest_division(uvrt_ctx* ctx, int blocks, int itersPerThread, unsigned long long* out3)
{
    if (!ctx || !out3 || blocks <= 0 || itersPerThread <= 0) return UVRT_ERR_INVALID;
    Bind b(ctx);
    unsigned long long* d = nullptr;
    CK(cudaMalloc((void**)&d, 24));
    CK(cudaMemsetAsync(d, 0, 24, ctx->stream));
    k_selftest_division<<<blocks, 256, 0, ctx->stream>>>(d, itersPerThread, 0x9e3779b9u);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out3, d, 24, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, UVRT_ERR_CUDA, "selftest_division: %s", cudaGetErrorString(e));
    return UVRT_OK;
}

} // extern "C"
