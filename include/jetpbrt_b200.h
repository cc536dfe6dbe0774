/* jetpbrt_b200.h -- the drop-in boundary: C ABI of libjetpbrt_b200.so.
 *
 * The reference has no FFI; its seam for the hot path is the C++ call
 *     void FIntegrator::Render(const FScene*, FSampler*, FFilm*, int numthreads) const
 * (reference src/integrator.h:32, sole call site main.cc:156) on a scene prepared by
 * FScene::Preprocess (scene.cc:11-23), followed by FFilm::SaveAsImage (main.cc:160).
 * This header is what a maintainer of the reference would bind instead (INTEGRATION.md shows the
 * FIntegrator subclass that does it).  Plain pointers and sizes only; no C++ or torch types; every
 * function returns 0 on success or a negative jpbrt_status, never throws, and records a message
 * retrievable with jpbrt_last_error().  One host thread per context.
 *
 * There is NO CPU fallback: every compute entry point fails with JPBRT_ERR_CUDA when no sm_100
 * device (or no CUDA driver) is present.
 */
#ifndef JETPBRT_B200_H
#define JETPBRT_B200_H

#include <stddef.h>
#include <stdint.h>

#include "jetpbrt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum jpbrt_status {
    JPBRT_OK = 0,
    JPBRT_ERR_INVALID = -1,     /* bad argument / malformed scene description */
    JPBRT_ERR_CUDA = -2,        /* CUDA runtime error or no usable device */
    JPBRT_ERR_UNSUPPORTED = -3, /* scene uses a feature outside the hot-path scope */
    JPBRT_ERR_IO = -4
} jpbrt_status;

typedef struct jpbrt_ctx jpbrt_ctx;

/* Which FIntegrator::Li a pass evaluates (the reference picks one at compile time, main.cc:151-154).
 * Option "integrator" of jpbrt_set_option; the default is the one main.cc ships with. */
typedef enum jpbrt_integrator {
    JPBRT_INTEGRATOR_PATH = 0,           /* FPathIntegratorIteration (integrator.cc:316-403) -- the hot path */
    JPBRT_INTEGRATOR_PATH_RECURSIVE = 1, /* FPathIntegratorRecursive (integrator.cc:233-307): the same estimator and the
                                            same sampler draws, so it runs the same kernels; results equal the
                                            reference's up to the float evaluation order of the throughput */
    JPBRT_INTEGRATOR_WHITTED = 2,        /* FWhittedIntegrator (integrator.cc:115-220): direct light at every vertex,
                                            recursion through specular lobes only (a mirror is traced twice, as there) */
    JPBRT_INTEGRATOR_DEBUG = 3           /* FDebugIntegrator (integrator.h:44-58): the hit normal as the colour */
} jpbrt_integrator;

/* ------------------------------------------------------------------------------------------
 * The render trio (BASELINE.json north_star: upload_scene / render_pass / read_film).
 * ---------------------------------------------------------------------------------------- */

/* Replaces FScene::Preprocess (scene.cc:11-23) + the scene graph the integrator reads
 * (scene.h:139-149): computes shape normals/bounds, the world bound and the environment-light
 * radius, builds and flattens the BVH on the host, and copies everything to `device`.
 * The description is copied; the caller may free it afterwards. */
int jpbrt_upload_scene(const jpbrt_scene_desc* desc, int device, jpbrt_ctx** out_ctx);

/* Same with flags.  JPBRT_UPLOAD_GPU_BVH: build the BVH on the device (linear BVH: Morton sort + radix tree +
 * bottom-up refit, csrc/bvh_build.cuh) instead of the host's binned-SAH build -- replaces the recursive build of
 * FBVH_Node (bvh.h:59-92) for scenes whose build time matters; ~100x faster to build, slower to traverse.  Hits do
 * not depend on the tree.  jpbrt_upload_scene reads the default from the environment (JPBRT_BVH_BUILDER=gpu). */
#define JPBRT_UPLOAD_GPU_BVH 1u
int jpbrt_upload_scene_ex(const jpbrt_scene_desc* desc, int device, unsigned flags, jpbrt_ctx** out_ctx);

/* Replaces FIntegrator::DoRender (integrator.cc:82-111) for sample indices
 * [sample_begin, sample_begin + sample_count) of EVERY pixel: adds the raw radiance of each
 * sample (FPathIntegratorIteration::Li, integrator.cc:316-403) into the device film.
 * Asynchronous: work is queued on the context's stream.  The sampler is counter based: the
 * result depends only on (seed, pixel, sample index), not on how samples are split into passes
 * or across GPUs. */
int jpbrt_render_pass(jpbrt_ctx* ctx, int sample_begin, int sample_count, uint64_t seed);

/* Replaces FFilmView::AddColor(x, y, Clamp01(L)) (integrator.cc:108, film.h:64-68) and the film
 * readback.  If `finalize` != 0 writes clamp01(sum / spp_total), else the raw sums.  `rgb` is
 * host memory, width*height*3 floats, row 0 = top of the image (the reference's FFilm layout).
 * Synchronises the stream.  On a context with a communicator (jpbrt_comm_init) the films of all ranks are first
 * summed onto rank 0 (one ncclReduce); spp_total is then the total over all ranks, and ranks != 0 may pass rgb = NULL. */
int jpbrt_read_film(jpbrt_ctx* ctx, float* rgb, int spp_total, int finalize);

/* Zero the device film (FFilm::Clear, film.h:75-82). */
int jpbrt_clear_film(jpbrt_ctx* ctx);
/* Zero the work counters, the launch count and the stage timers reported by jpbrt_get_stats. */
int jpbrt_reset_stats(jpbrt_ctx* ctx);

void        jpbrt_destroy(jpbrt_ctx* ctx);
const char* jpbrt_last_error(const jpbrt_ctx* ctx); /* ctx may be NULL: last error of this thread */

/* One call = FIntegrator::Render (integrator.cc:35-80): upload + all passes + finalize + readback
 * into host memory.  `seconds_out` (optional) receives the wall time of the render part, the
 * interval the reference prints (integrator.cc:77-79). */
int jpbrt_render(const jpbrt_scene_desc* desc, int spp, uint64_t seed, int device, float* rgb, double* seconds_out);

/* Same with one of the reference's other integrators (jpbrt_integrator). */
int jpbrt_render_integrator(const jpbrt_scene_desc* desc, int integrator, int spp, uint64_t seed, int device, float* rgb,
                            double* seconds_out);

/* ------------------------------------------------------------------------------------------
 * Plumbing for multi-GPU (one process per GPU; the film sum is reduced by the caller with NCCL,
 * e.g. torch.distributed.reduce on a tensor aliasing this buffer) and for timing.
 * ---------------------------------------------------------------------------------------- */
void*  jpbrt_film_device_ptr(jpbrt_ctx* ctx);             /* float32[width*height*3] raw sums, device memory */
size_t jpbrt_film_num_floats(const jpbrt_ctx* ctx);
void*  jpbrt_stream(jpbrt_ctx* ctx);                      /* cudaStream_t all work is queued on */
int    jpbrt_synchronize(jpbrt_ctx* ctx);
/* Finalise on the device: out[i] = clamp01(film[i] / spp_total); out may alias the film. */
int    jpbrt_finalize_film_device(jpbrt_ctx* ctx, void* out_device, int spp_total);
/* Re-upload the flattened scene arrays from the host copy kept in the context (what a per-frame
 * caller pays for a changed scene); returns bytes copied through *bytes. */
int    jpbrt_reupload_scene(jpbrt_ctx* ctx, size_t* bytes);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU inside the boundary (replaces the thread pool of FIntegrator::Render, integrator.cc:53-74,
 * parallel.cc, across devices): the scene is replicated, the SAMPLE indices of every pixel are partitioned, and the raw
 * float32 films are summed onto rank 0 with ONE ncclReduce over NVLink; Clamp01(sum / spp_total) happens after the
 * reduce, as the reference clamps the mean (integrator.cc:108).  NCCL is bound at run time (dlopen libnccl.so.2: the
 * instance already loaded in the process, e.g. torch's, else the system's); without it these calls return
 * JPBRT_ERR_UNSUPPORTED and the single-GPU path is unaffected.
 *
 *   one process per GPU : rank 0 calls jpbrt_comm_unique_id and hands the bytes to the other ranks (any transport),
 *                         every rank calls jpbrt_comm_init on its context, renders its own sample range
 *                         (jpbrt_sample_partition) with jpbrt_render_pass, then calls jpbrt_read_film: a context with a
 *                         communicator reduces first; rank 0 receives the image, the other ranks pass rgb = NULL.
 *   one process, N GPUs : jpbrt_render_multi (ncclCommInitAll + grouped reduce).
 * ---------------------------------------------------------------------------------------- */
#define JPBRT_COMM_ID_BYTES 128
int  jpbrt_comm_unique_id(void* id, size_t bytes);
int  jpbrt_comm_init(jpbrt_ctx* ctx, const void* id, size_t bytes, int rank, int nranks);
int  jpbrt_comm_rank(const jpbrt_ctx* ctx);
int  jpbrt_comm_size(const jpbrt_ctx* ctx);
/* The reduce alone (stream-ordered, asynchronous): afterwards rank 0's device film holds the sum over all ranks and the
 * other ranks' films are zero.  jpbrt_read_film calls it if it has not run since the last pass. */
int  jpbrt_reduce_film(jpbrt_ctx* ctx);
/* Rank r's share of spp_total sample indices: contiguous, as even as possible. */
void jpbrt_sample_partition(int spp_total, int rank, int nranks, int* begin, int* count);
/* FIntegrator::Render with `ngpus` devices (0 .. ngpus-1) of this process for threads: upload to every device, render the
 * partitions concurrently, reduce, finalize, read back.  reduce_ms_out (optional): device time of the reduce on rank 0. */
int  jpbrt_render_multi(const jpbrt_scene_desc* desc, int integrator, int spp, uint64_t seed, int ngpus, float* rgb,
                        double* seconds_out, double* reduce_ms_out);

/* Options (value 0 / 1 unless noted):
 *   "integrator"       jpbrt_integrator
 *   "paths_in_flight"  capacity of the path pool = paths per wavefront (0 = automatic: the pass, at most 2^26)
 *   "band_pixels"      pixels per band of a wavefront (0 = 2^20); >= the frame disables banding
 *   "stage_timing"     per-stage CUDA events (jpbrt_stats.ms_*); disables CUDA-graph replay
 *   "count_traversal"  node / primitive test counters (jpbrt_stats.box_tests ...); uses the counting kernel variants
 *   "use_graph"        replay a wavefront as one CUDA graph (default 1)
 *   "sort_rays"        reorder the rays of bounces >= 1 by (origin cell, direction octant) before they are traced:
 *                      0 off, else cell bits per axis 1..6, +16 to include the octant
 *   traversal tunables "trav_blocks" (5 or 6 resident blocks per SM; other values are clamped), "refill_min" (idle lanes
 *                      that trigger a refill, 1..32, -1 automatic), "min_inner" (lanes at inner nodes below which a warp's
 *                      node phase ends, 0..32, -1 automatic), "node_format" (-1 automatic, 0 64-byte float BVH nodes,
 *                      1 32-byte nodes quantised to a 16-bit grid over the scene bounds -- one load per node step, boxes
 *                      rounded outwards; resident for trees of at most 2^20 nodes, automatic from 1,024 nodes)
 * The film is accumulated with float atomics: a render is reproducible up to float summation order (<= 1e-6 relative),
 * not bit for bit; the PATHS depend only on (seed, pixel, sample index). */
int jpbrt_set_option(jpbrt_ctx* ctx, const char* name, long long value);

typedef struct jpbrt_stats {
    uint64_t samples;          /* camera paths started */
    uint64_t extension_rays;   /* FScene::Intersect calls from Li (integrator.cc:327) */
    uint64_t shadow_rays;      /* FScene::Occluded calls (integrator.cc:367) */
    uint64_t shaded_vertices;
    uint64_t box_tests;        /* only with option "count_traversal" = 1 */
    uint64_t prim_tests;       /* only with option "count_traversal" = 1 */
    uint64_t shadow_box_tests;
    uint64_t shadow_prim_tests;
    uint64_t invalid_contributions; /* NaN/inf radiance dropped instead of being stored */
    uint64_t kernel_launches;  /* kernels of this library launched since the last jpbrt_reset_stats */
    double   ms_generate, ms_extend, ms_shade, ms_connect, ms_finalize; /* with option "stage_timing" = 1 */
    uint64_t n_nodes, n_prim_slots, scene_bytes;
    double   bvh_build_seconds;
    uint64_t paths_in_flight;  /* capacity of the path pool (paths per wavefront), 0 before the first pass */
    uint64_t bvh_builder;      /* 0 host binned SAH, 1 GPU linear BVH (JPBRT_UPLOAD_GPU_BVH), 2 host object-median rebuild (tree too deep) */
    double   bvh_device_seconds; /* GPU builder: box upload + sort/tree/refit kernels + tree download (part of bvh_build_seconds) */
    /* What the pipeline had to drop -- each is also part of invalid_contributions; all are 0 in a healthy render: */
    uint64_t dropped_rays;     /* rays that found the next wavefront queue full (Whitted ray trees), or were still queued after the last bounce */
    uint64_t stack_overflows;  /* far BVH children lost to a full traversal stack (tree deeper than 63 levels of pending subtrees) */
    uint64_t nee_dropped;      /* light samples that found no shadow-ray slot (pool smaller than vertices x lights) */
    uint64_t bvh_depth;        /* levels of the BVH (root = 1); trees deeper than 62 are rebuilt with median splits at upload */
    double   ms_reduce;        /* with "stage_timing": device time of the NCCL film reduce(s) on this rank */
    /* with "count_traversal": DISTINCT node / primitive records fetched per warp step (lanes of a warp that sit on the same
     * node share one fetch) -- what the memory system has to deliver, as opposed to the per-lane test counts above */
    uint64_t node_fetches, prim_fetches, shadow_node_fetches, shadow_prim_fetches;
    uint64_t node_bytes;       /* size of the BVH node record the production traversal kernels fetch per node step: 64 (float
                                * boxes) or 32 (boxes quantised to a 16-bit grid, option "node_format") */
} jpbrt_stats;
int jpbrt_get_stats(jpbrt_ctx* ctx, jpbrt_stats* out);

/* ------------------------------------------------------------------------------------------
 * Unit kernels: the device functions of the wavefront stages run on caller-supplied arrays, for
 * parity tests against the reference's CPU functions on identical inputs.
 * All pointers are HOST memory; arrays are copied in and out.  Layouts match oracle/oracle_api.h.
 * ---------------------------------------------------------------------------------------- */
/* FShape::Intersect on one shape (shape.h:291-327 / 399-435 / 487-526 / 199-221). rays8 = o,d,tmin,tmax */
int jpbrt_unit_intersect_shape(const jpbrt_shape* shape, int device, int n, const float* rays8,
                               int* hit, float* t, float* pos3, float* nrm3);
/* FScene::Intersect (scene.cc:25-33) through the flattened BVH: the `extend` stage's traversal. */
int jpbrt_unit_scene_intersect(jpbrt_ctx* ctx, int n, const float* rays8, int* prim, float* t, float* pos3, float* nrm3);
/* FScene::Occluded (scene.h:36-47): the `connect` stage's any-hit traversal. */
int jpbrt_unit_scene_occluded(jpbrt_ctx* ctx, int n, const float* pos3, const float* target3, int* occluded);
/* material->Scattering + FBSDF::Evalf/Pdf/Sample (material.cc, bsdf.h:285-302): the `shade` stage's BSDF code. */
int jpbrt_unit_bsdf(const jpbrt_material* mat, int device, int n,
                    const float* nrm3, const float* wo3, const float* wi3, const float* u2, const float* ulobe,
                    float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta);
/* The BSDF classes no FMaterial of the reference builds (jpbrt_bsdf_desc): FPhongSpecularReflection
 * (bsdf.h:557-633), FMicrofacetReflection with Beckmann / Trowbridge-Reitz and any Fresnel (bsdf.cc:35-78,
 * microfacet.cc:11-254), FMicrofacetTransmission (bsdf.cc:80-145): Evalf / Pdf / Sample in world space, frame = FFrame(nrm). */
int jpbrt_unit_bsdf_ex(const jpbrt_bsdf_desc* desc, int device, int n, const float* nrm3, const float* wo3, const float* wi3,
                       const float* u2, float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags);
/* FLight::Sample_Li (light.h): the `shade` stage's next-event sampling. */
int jpbrt_unit_light_sample(jpbrt_ctx* ctx, int light, int n, const float* pos3, const float* nrm3, const float* u2,
                            float* lpos3, float* wi3, float* pdf, float* Li3);
/* FPrimitive::GetLe (primitive.h:60-63). */
int jpbrt_unit_emitted(jpbrt_ctx* ctx, int n, const int* prim, const float* nrm3, const float* wo3, float* Le3);
/* FCamera::GenerateRay (camera.h:52-58): the `generate` stage. */
int jpbrt_unit_generate_rays(jpbrt_ctx* ctx, int n, const float* posfilm2, float* o3, float* d3);
/* The counter-based sampler's block (pixel, sample, block) -> 4 floats in [0,1). */
int jpbrt_unit_rng_block(int device, int n, const uint32_t* pixel, const uint32_t* sample, const uint32_t* block,
                         uint64_t seed, float* out4);
/* The raw generator under it: Philox4x32-10 on n (counter[4], key[2]) pairs -> 4 words each (for the published
 * known-answer vectors). */
int jpbrt_unit_philox_raw(int device, int n, const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4);
/* Scene facts computed at upload: out[0..2] world min, [3..5] world max, [6] environment radius. */
int jpbrt_scene_info(jpbrt_ctx* ctx, float* out7);

/* ------------------------------------------------------------------------------------------
 * Host-side scene construction (C++ class jetpbrt::Scene, host/scene.h) behind C handles:
 * the five BASELINE.json configurations and OBJ ingestion.  No GPU needed.
 * ---------------------------------------------------------------------------------------- */
typedef struct jpbrt_scene jpbrt_scene;
/* name: "cornell" (C1/C5), "bunny" (C2), "large" (C3), "glossy" (C4). scale: mesh-resolution
 * multiplier for bunny/large (1 = the configuration's full size). */
jpbrt_scene*            jpbrt_scene_builtin(const char* name, int width, int height, float scale);
const jpbrt_scene_desc* jpbrt_scene_get_desc(jpbrt_scene* s);
void                    jpbrt_scene_free(jpbrt_scene* s);

/* FFilm::SaveAsImage (film.cc:11-188): kind 0 = PPM, 1 = BMP, 2 = HDR (EImageType, film.h:15-20); `basename` without
 * extension.  Byte-identical to the reference's files wherever those are well defined: HDR for every pixel whose largest
 * channel is >= 1e-32 (below that the reference writes an uninitialised rgbe[4], film.cc:159-181; here: zeros), BMP whenever a row of
 * width*3 bytes needs no padding (width % 4 == 0 -- all BASELINE resolutions), PPM always -- including the reference's
 * quirk of streaming every channel as one raw byte under a "P3" header (film.cc:53-55).  Kind 3 writes the well-formed
 * text P3 instead.  For widths whose rows need padding the reference's BMP writer reads its padded buffer with the
 * unpadded stride (film.cc:139-141: a skewed, short file); this writer emits the valid padded BMP. */
int jpbrt_save_image(const char* basename, int kind, int width, int height, const float* rgb);

/* Wavefront OBJ -> triangles as LoadTriangleMesh delivers them (shape.cc:23-68, scene.cc:49-64): positions only, 9 floats
 * per triangle, transformed in the reference's order (z negated if flip_handedness, then * scale, then + offset3; offset3
 * may be NULL).  Backslashes in `filename` are accepted (the reference spells "scene\\bunny\\bunny.obj", main.cc:94).
 * Copies up to `capacity` triangles into tris9 (may be NULL) and returns the triangle count, or a negative status. */
long long jpbrt_load_obj_triangles(const char* filename, int flip_handedness, const float* offset3, float scale, float* tris9,
                                   long long capacity);

/* Number of CUDA devices visible (0 if none / no driver): lets callers skip instead of fail. */
int jpbrt_device_count(void);

/* Host-only view of what jpbrt_upload_scene would upload, for validating the BVH and the flattened
 * tables without a GPU.  `what`: 0 nodes, 1 slots, 2 slot_nrm (float4 records); 3 slot_ml (int2),
 * 4 prim_slot (int); 5 materials, 6 lights (float4 records); 7 the 32-byte quantised nodes (8 words each: six
 * min | max << 16 plane-index words, two child references; empty for trees of more than 2^20 nodes), 8 their grid
 * (origin xyz, cell xyz).  Copies up to `capacity` 32-bit words
 * into `out` (may be NULL) and returns the total number of 32-bit words, or a negative status. */
long long jpbrt_debug_flatten(const jpbrt_scene_desc* desc, int what, void* out, long long capacity);

/* The same tables (0 nodes, 1 slots, 2 slot_nrm, 3 slot_ml, 4 prim_slot) as a CONTEXT holds them -- i.e. of the tree
 * its builder (host SAH or GPU LBVH) actually produced; for validating a device-built BVH. */
long long jpbrt_debug_ctx_table(jpbrt_ctx* ctx, int what, void* out, long long capacity);

const char* jpbrt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JETPBRT_B200_H */
