/*
 * uvrt.h -- C ABI of libuvrt.so, the B200 (sm_100a) CUDA backend for the UV dose ray tracer's
 * wavefront hot path: generate -> extend -> accumulate -> computeDosage -> dosageToColor (+reset).
 *
 * This boundary replaces the OpenCL layer the reference's RayTracer drives
 * (Kernel / Buffer in /root/reference/template/template.cpp:1039-1573): every entry point
 * below names the reference call site it stands in for.  Plain pointers and sizes only; no
 * C++ or torch types.  All calls return UVRT_OK (0) or a negative uvrt_status and never
 * abort the process (the reference's FatalError -> exit(0), template.cpp:904-917, becomes an
 * error code + uvrt_last_error()).  There is no CPU fallback: without a usable CUDA device
 * uvrt_create() fails.
 *
 * Launches are asynchronous and ordered on the context's stream, like Kernel::Run on the
 * reference's single in-order queue (template.cpp:1568-1573); uvrt_read()/uvrt_sync()
 * synchronise (clFinish in myapp.cpp:165 / raytracer.cpp:202).
 *
 * Host data layouts are the reference's own:
 *   Tri      64 B  v0.xyz,pad,v1.xyz,pad,v2.xyz,pad,centroid.xyz,pad   (mesh.h:6-13, tools.cl:31-37)
 *   BVHNode  32 B  min.xyz,leftFirst,max.xyz,triCount                  (bvh.h:11-21, tools.cl:39-45)
 *   Ray      32 B  dir.xyz,orig.xyz,dist,triID                         (tools.cl:8-14)
 */
#ifndef UVRT_H
#define UVRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvrt_ctx uvrt_ctx;

typedef enum uvrt_status {
    UVRT_OK = 0,
    UVRT_ERR_INVALID = -1,   /* bad argument / bad scene data */
    UVRT_ERR_CUDA = -2,      /* CUDA runtime or kernel failure (message in uvrt_last_error) */
    UVRT_ERR_NO_SCENE = -3,  /* a stage was launched before uvrt_upload_scene */
    UVRT_ERR_NCCL = -4,      /* NCCL missing or failed */
    UVRT_ERR_NO_MEMORY = -5, /* host or device allocation failed */
    UVRT_ERR_NO_DEVICE = -6, /* no CUDA device / driver: the backend cannot run (no fallback) */
    UVRT_ERR_IO = -7         /* a file could not be read or written (host layer) */
} uvrt_status;

/* Device buffers addressable through uvrt_read / uvrt_write.  Names follow raytracer.h:54-55. */
typedef enum uvrt_buffer {
    UVRT_BUF_RAYS = 0,   /* rayBuffer:          nRays x 32 B (only the rays of the last launch)   */
    UVRT_BUF_COUNTS = 1, /* tempPhotonMapBuffer: nTris x i32  per-launch photon counts            */
    UVRT_BUF_SUM = 2,    /* photonMapBuffer:     nTris x f64  sum of count x duration             */
    UVRT_BUF_MAX = 3,    /* maxPhotonMapBuffer:  nTris x f64  max per-launch count                */
    UVRT_BUF_DOSE = 4,   /* dosageBuffer:        nTris x f32                                      */
    UVRT_BUF_COLOR = 5,  /* colorBuffer:         nTris x 9 x f32 (plain buffer instead of GL VBO) */
    UVRT_BUF_PAIRS = 6,  /* diagnostic, read only: traversal layout, 64 B per inner node   */
    UVRT_BUF_WTRIS = 7,  /* diagnostic, read only: leaf-ordered triangles, 64 B per slot                  */
    UVRT_BUF_MATRIX = 8  /* count matrix of uvrt_matrix_begin: rows x nTris x i32 (read: waits for the traces in
                            flight; write: lets a caller without NCCL exchange the rows by its own means)  */
} uvrt_buffer;

/* Stage ids for uvrt_stage_time */
typedef enum uvrt_stage {
    UVRT_STAGE_GENERATE = 0,
    UVRT_STAGE_EXTEND = 1,
    UVRT_STAGE_ACCUMULATE = 2,
    UVRT_STAGE_SHADE = 3,
    UVRT_STAGE_COLOR = 4,
    UVRT_STAGE_RESET = 5,
    UVRT_STAGE_BIN = 6,
    UVRT_STAGE_COUNT = 7
} uvrt_stage;

/* ---- context: replaces Kernel::InitCL / KillCL (template.cpp:1303-1455) ------------------- */
int uvrt_device_count(int* count);
int uvrt_create(uvrt_ctx** out, int device);
void uvrt_destroy(uvrt_ctx* ctx);
/* Message of the last failure on ctx (ctx == NULL: last failure of uvrt_create on this thread). */
const char* uvrt_last_error(const uvrt_ctx* ctx);
/* "sm_100 NVIDIA B200 148 SMs ..." */
int uvrt_device_info(uvrt_ctx* ctx, char* dst, size_t bytes, int* smCount, int* ccMajor, int* ccMinor);

/* ---- scene: replaces the Buffer creation + CopyToDevice of raytracer.cpp:24-30 and the
 *      scene swap of CalibratePower (raytracer.cpp:166-187, 214-224).
 * Copies (the host keeps ownership) and repacks into the device layout.  nNodes is
 * the number of valid 32-B slots at `nodes` (node 0 = root, node 1 unused, bvh.cpp:16); every
 * node reachable from the root must lie inside it (the reference uploads 2N, which truncates
 * its own tree -- SURVEY App. B-3 -- and is rejected here with UVRT_ERR_INVALID).
 * The arrays go up as they are; the tree is validated and repacked on the device.  A rejected upload
 * leaves the previously uploaded scene usable.  (Re)allocates and zeroes the per-triangle buffers when
 * nTris changes.  Synchronises. */
int uvrt_upload_scene(uvrt_ctx* ctx, const void* tris, int nTris, const void* nodes, int nNodes,
                      const uint32_t* triIdx);

/* BVH::Build (bvh.cpp:13-44) on the device: binned-SAH BVH2 with the reference builder's tree, node
 * numbering and triIdx order (bit-identical to host/bvh.cpp).  tris: nTris x 64 B (host); the
 * centroid lanes are written back when trisOut != NULL (the builder computes them, bvh.cpp:23).
 * nodesOut: nodeCapacity x 32 B (host; 2*nTris + 64 is always enough); triIdxOut: nTris x u32.
 * *nodesUsed = highest used node index + 1.  Synchronises.  Independent of the uploaded scene.
 * One caveat to "bit-identical": where a box or centroid extremum is attained by both -0.0 and +0.0 the host
 * builder keeps the zero it meets first in triangle order (p < lo ? p : lo), the device reduction the smaller
 * encoding (-0.0 for a minimum, +0.0 for a maximum).  Topology, triIdx and every non-zero bound are identical, and
 * no comparison or quotient of the traversal can tell the two zeros apart (a - o is the same number either way). */
int uvrt_build_bvh(uvrt_ctx* ctx, const void* tris, int nTris, void* nodesOut, int nodeCapacity,
                   uint32_t* triIdxOut, uint32_t* nodesUsed, void* trisOut);

/* ---- stages ---------------------------------------------------------------------------- */
/* reset.cl:4-26 via RayTracer::ClearBuffers (raytracer.cpp:133-143) */
int uvrt_reset(uvrt_ctx* ctx, int resetColor);
/* generate.cl:8-40.  Writes rays [firstRay, firstRay+nRays) of the launch to slots 0..nRays-1.
 * seedIn is the launch-start value of the reference's program-scope SEED. */
int uvrt_generate(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength,
                  int64_t firstRay, int64_t nRays, uint32_t seedIn);
/* extend.cl:85-99 on the first nRays slots of the ray buffer: closest hit written back in
 * place (dist, triID) and counts[triID]++ per hit. */
int uvrt_extend(uvrt_ctx* ctx, int64_t nRays);
/* accumulate.cl:4-14 */
int uvrt_accumulate(uvrt_ctx* ctx, float duration);
/* RayTracer::ComputeSingleLightDosageMap (raytracer.cpp:75-88): generate -> extend -> accumulate.
 * Asynchronous like Kernel::Run; consecutive calls overlap on the device (generate of the next launch and the
 * tail of the previous extend run next to the current extend), results are as if run back to back.
 * UVRT_BUF_COUNTS is zero again afterwards (accumulate.cl:13). */
int uvrt_trace(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, float duration,
               int64_t firstRay, int64_t nRays, uint32_t seedIn);
/* Same without the accumulate step (for launches split over several GPUs: the partial counts
 * are summed with uvrt_reduce_counts before uvrt_accumulate). */
int uvrt_trace_counts(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength,
                      int64_t firstRay, int64_t nRays, uint32_t seedIn);
/* SEED chain of nLaunches consecutive launches (generate.cl:13,39): seedsOut[0] = seedIn,
 * seedsOut[i+1] = work-item 0's final RNG state of launch i at lightPos3[3*i..].  Runs on the
 * device (one thread); synchronises. */
int uvrt_seed_chain(uvrt_ctx* ctx, const float* lightPos3, int nLaunches, float lightLength,
                    uint32_t seedIn, uint32_t* seedsOut);
/* shade.cl:23-41 computeDosage via RayTracer::Shade (raytracer.cpp:93-120).  useMaxMap selects
 * maxPhotonMapBuffer (viewMode == maxpower) instead of photonMapBuffer. */
int uvrt_shade(uvrt_ctx* ctx, int useMaxMap, int photonsPerLight, float scaledPower);
/* shade.cl:43-71 dosageToColor */
int uvrt_color(uvrt_ctx* ctx, float minValue, int thresholdView);

/* ---- data movement: Buffer::CopyFromDevice / CopyToDevice (template.cpp:1086-1107) ------- */
int uvrt_read(uvrt_ctx* ctx, uvrt_buffer what, void* dst, size_t bytes);       /* synchronises */
int uvrt_write(uvrt_ctx* ctx, uvrt_buffer what, const void* src, size_t bytes); /* synchronises */
int uvrt_sync(uvrt_ctx* ctx);                                                   /* clFinish */

/* ---- multi-GPU (no counterpart in the reference: one cl_context, one device) ------------- */
/* One process (or thread) per GPU.  id is NCCL's 128-byte unique id, created on rank 0 and
 * passed to the other ranks by the caller (torch.distributed / MPI / a file). */
int uvrt_comm_unique_id(void* id128);
int uvrt_comm_init(uvrt_ctx* ctx, const void* id128, int rank, int nRanks);
/* Sum of UVRT_BUF_SUM and max of UVRT_BUF_MAX over all ranks (in place, every rank gets the result). */
int uvrt_reduce(uvrt_ctx* ctx);
/* Sum of UVRT_BUF_COUNTS over all ranks (in place). */
int uvrt_reduce_counts(uvrt_ctx* ctx);

/* Relative cost of a launch at a lamp position, for sharing launches between GPUs evenly: inner-node visits and
 * triangle tests per ray over the first nRays rays of the launch (reference-order traversal with counters).
 * Deterministic: every rank gets the same numbers.  Synchronises; does not touch rays, counts or maps. */
int uvrt_probe_cost(uvrt_ctx* ctx, float lx, float ly, float lz, float lightLength, uint32_t seedIn, int nRays,
                    double* innerVisitsPerRay, double* triangleTestsPerRay);

/* Count matrix: the exchange format of runs whose launches are shared between GPUs, whole or cut into ray
 * ranges.  accumulate.cl:4-14 folds every launch's integer counts into an f64 sum and an
 * f64 per-launch maximum, so per-GPU f64 maps can only be combined exactly when no launch is split and the
 * count x duration products happen to add without rounding.  Instead every rank writes the counts of launch k
 * (its share of the rays) into row k of a rows x nTris int32 matrix, the rows are summed over the ranks with
 * ONE ncclAllReduce, and the fold replays accumulate row by row in launch order on every rank: photon map and
 * max map are then bit-identical to the single-GPU run for any split and any durations.
 *   uvrt_matrix_begin  sizes and zeroes the matrix (a run, or a window of a long run); two buffers alternate from window
 *                      to window, so the all-reduce + fold of one window run next to the extends of the next
 *   uvrt_trace_row     generate -> (bin) -> extend of rays [firstRay, firstRay + nRays) of a launch, counted in
 *                      `row`; asynchronous, consecutive calls overlap like uvrt_trace
 *   uvrt_matrix_fold   (reduce != 0 and a communicator exists: all-reduce of the first `rows` rows, then)
 *                      UVRT_BUF_SUM += count x durations[r], UVRT_BUF_MAX = max(., count) for r = 0 .. rows-1 */
int uvrt_matrix_reserve(uvrt_ctx* ctx, int rows);   /* optional: capacity for windows of up to `rows` rows, allocated now */
int uvrt_matrix_begin(uvrt_ctx* ctx, int rows);
int uvrt_trace_row(uvrt_ctx* ctx, int row, float lx, float ly, float lz, float lightLength,
                   int64_t firstRay, int64_t nRays, uint32_t seedIn);
int uvrt_matrix_fold(uvrt_ctx* ctx, const float* durations, int rows, int reduce);

/* ---- tuning and measurement ------------------------------------------------------------- */
/* Options (defaults are the measured best; everything else exists for A/B runs, see DESIGN.md section 4 and
 * profiles/r1_sweeps.md):
 *   "extend_variant"  kernel selection: 0/1/2 one thread per ray with IEEE / two-step / one-step (default) slab
 *                     division; 50 the certified fast extend (csrc/uvrt_fast.cuh: conservative inner-node tests
 *                     on 32-byte quantised node pairs, the winner verified with the reference's exact slab test,
 *                     rays without a certificate re-traced in reference order -- same bits as 0/1/2).  -1 (default)
 *                     is 50 wherever it can serve the scene (tame, nested boxes; uvrt_get_option "fast_ready"), else 2.
 *                     "fast_check" = 1 traces every ray both ways and counts disagreements (uvrt_fast_stats),
 *                     "fast_cfg" selects the register budget (0: 48, the default; 1: 64); read only:
 *                     "scene_nested", "fast_ready".  Builds with -DUVRT_EXPERIMENTS (`make EXPERIMENTS=1`, read-only
 *                     option "experiments") also carry the rejected variants of profiles/r1_sweeps.md and
 *                     profiles/r2_fast_extend.md: 10..24 persistent warps with a global queue, 40..43 chunk-persistent
 *                     warps, and "fast_cfg" 2, the refill variant of the fast extend ("refill_chunk": rays per chunk)
 *   "timeline"        1: log the host time of every call and the device start / stop of every stage launch
 *                     (uvrt_timeline_dump)
 *   "bin_rays"        1 (default): counting sort of the ray queue by direction / origin cell before extend;
 *                     "bin_y", "bin_t", "bin_p": the bin grid
 *   "pipeline"        1 (default): generate + bin of launch k+1 on a second stream next to extend k
 *   "overlap_extend"  1 (default): uvrt_trace alternates two extend streams / count buffers, so extend k+1 fills
 *                     the SMs that the last wave of extend k leaves idle
 *   "fetch_mode"      3 (default): rays, permutation, results bypass L1; 0 plain (experiment builds: 1/2 texture
 *                     path; 4 evict_last nodes; 5 SM-affine chunks)
 *   "host_repack"     0 (default): uvrt_upload_scene repacks the scene on the device; 1: on the host cores
 *   "stage_timing"    1: bracket every launch with CUDA events (uvrt_stage_time)
 *   "simple_cfg", "hist_mode", "refill", "chunk", "blocks_per_sm", "generic_octant", "carveout": sweep knobs
 *   read only: "scene_tame" (1 = every box coordinate is 0 or in [2^-77, 2^20]: fast slab division allowed) */
int uvrt_set_option(uvrt_ctx* ctx, const char* key, int value);
int uvrt_get_option(uvrt_ctx* ctx, const char* key, int* value);
/* Sum of event-timed durations and number of launches of a stage since the last
 * uvrt_stage_time_reset (needs "stage_timing" = 1).  Synchronises. */
int uvrt_stage_time(uvrt_ctx* ctx, uvrt_stage stage, double* ms, int64_t* launches);
int uvrt_stage_time_reset(uvrt_ctx* ctx);
/* Option "timeline" = 1 starts a log of every C-ABI call (host clock) and of every stage launch (host time of
 * the enqueue, device start / stop from CUDA events); uvrt_timeline_dump writes it as JSON (synchronises).
 * Every stage launch and API call is also an NVTX range (nvtx3, header only) for nsys / ncu --nvtx. */
int uvrt_timeline_dump(uvrt_ctx* ctx, const char* path);
/* Kernels launched by this context since creation. */
int64_t uvrt_launch_count(const uvrt_ctx* ctx);
/* Event timing on the context's stream (torch.cuda.Event only sees torch's stream). */
int uvrt_mark(uvrt_ctx* ctx, int slot);                       /* slot 0..15 */
int uvrt_elapsed_ms(uvrt_ctx* ctx, int slotStart, int slotStop, float* ms); /* synchronises */
/* Overwrites a scratch buffer larger than the L2 (256 MiB) on the context's stream, so that the
 * next launch starts from a cold L2 (benchmark hygiene). */
int uvrt_flush_l2(uvrt_ctx* ctx);
/* Bytes the last uvrt_upload_scene copied host -> device. */
int64_t uvrt_scene_upload_bytes(const uvrt_ctx* ctx);
/* Traversal statistics of the repacked scene: inner nodes, leaves, depth, stack bound. */
int uvrt_scene_info(uvrt_ctx* ctx, int* innerNodes, int* leaves, int* depth, int* stackEntries);

/* Certified fast extend ("extend_variant" 50, csrc/uvrt_fast.cuh): counters since the last reset.
 * out3 = {rays traced again in reference order because their certificate failed, rays not eligible for the fast
 * path (not tame, origin far outside the scene, a hit already recorded), certified rays whose answer differed
 * from the reference-order traversal -- counted only with option "fast_check" = 1, which traces every ray both
 * ways; must be 0}.  Synchronises. */
int uvrt_fast_stats(uvrt_ctx* ctx, unsigned long long* out3, int reset);

/* Diagnostic: compares the shared-reciprocal slab division used by the fast extend variants with
 * IEEE division on blocks*256*itersPerThread random operand pairs.
 * out3 = {samples, mismatches of the one-step form, mismatches of the two-step form}. */
int uvrt_selftest_division(uvrt_ctx* ctx, int blocks, int itersPerThread, unsigned long long* out3);

const char* uvrt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UVRT_H */
