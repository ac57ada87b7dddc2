// wavefront.cuh — buffers and launch parameters of the wavefront pipeline (shared by
// wavefront.cu and renderer.cpp-side host code).
//
// Pipeline per batch of paths:
//   k_primary   ray-gen + closest hit for camera rays (fused; misses with no env light die here)
//   k_shade     hit attributes, emitter hit / env miss MIS, Russian roulette, NEE shadow-ray
//               generation, BSDF / phase sampling, warp-ballot compaction of survivors
//   k_trace     one launch per bounce: closest hit for the compacted survivor queue + any-hit for the NEE rays
//               (marks the unoccluded ones)
//   k_settle    adds the marked NEE contributions (coalesced), resets the counters for the next bounce
//   k_resolve   per-sample clamp + per-pixel sum (renderer.cpp:77-84)
// All per-path state is SoA so that a warp's loads/stores are 128-byte transactions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200pt.h"
#include "device_scene.h"
#include "traverse.cuh"

namespace b200pt {

constexpr int kTileSize = 8;                  // pixels; tiles are dealt round-robin to ranks
constexpr int kTilePixels = kTileSize * kTileSize;

// One path segment waiting for its closest hit / shading.  Arrays of `capacity` elements.
struct PathQueue {
    float *ox, *oy, *oz;      // ray origin
    float *dx, *dy, *dz;      // ray direction (unit)
    float *tr, *tg, *tb;      // path throughput ("attenuation" in path.cpp)
    float *pdf;               // pdf of the direction sample that produced this ray (MIS at the next vertex)
    uint32_t *slot;           // sample slot inside the batch
    uint32_t *medium;         // volpath: medium the ray was scattered in, or kInvalid if it left a surface
    float *wx, *wy, *wz;      // volpath: `wo` of the last surface vertex (stale-wo behaviour of volpath.cpp)
    HitRec *hit;              // closest hit, written by k_primary / k_trace
};

struct ShadowQueue {
    float *ox, *oy, *oz, *dx, *dy, *dz, *tmax;
    float *cr, *cg, *cb;      // contribution if unoccluded (already multiplied by throughput)
    uint32_t *slot;
};

// Shading bins: a closest hit is filed under the BSDF model of the surface it found (bin = b200pt_bsdf_type; 0 = escaped
// ray or BSDF-less surface, 1 = area light), so that each shading launch runs ONE model (DeviceScene::integrator.shade_bins).
constexpr int kNumShadeBins = 8;
struct ShadeBins {
    uint32_t *lists;          // one list of queue positions per bin in use, `capacity` entries each; nullptr = no binning
    uint32_t capacity;
};
__host__ __device__ inline uint32_t BinListIndex(uint32_t bins_in_use, uint32_t bin) {
#if defined(__CUDA_ARCH__)
    return __popc(bins_in_use & ((1u << bin) - 1u));
#else
    return static_cast<uint32_t>(__builtin_popcount(bins_in_use & ((1u << bin) - 1u)));
#endif
}

enum KernelClass { kClassPrimary = 0, kClassExtend = 1, kClassShadow = 2, kClassShade = 3, kClassOther = 4, kClassTail = 5, kNumClasses = 6 };

struct ClassCounters {
    unsigned long long rays, node_visits, prim_tests;
};

// Per-sample radiance: ONE float4 (r, g, b, unused) per sample slot.  The adds are scattered (a path's slot has nothing to do
// with its queue position after the first bounce), and DRAM moves 32-byte sectors: three planes cost three sectors read and
// written per add (volumetric-caustic: k_settle 47 ms per 256-spp frame at 190 B of DRAM traffic per NEE ray,
// profiles/r02_counters_volumetric-caustic_1024x1024x256.json), one float4 costs one.
#ifdef __CUDACC__
__device__ __forceinline__ void RadianceAdd(float *radiance, uint32_t slot, float r, float g, float b) {
    float4 *p = reinterpret_cast<float4 *>(radiance) + slot;
    float4 v = *p;
    v.x += r, v.y += g, v.z += b;
    *p = v;
}
__device__ __forceinline__ void RadianceSet(float *radiance, uint32_t slot, float r, float g, float b) {
    reinterpret_cast<float4 *>(radiance)[slot] = make_float4(r, g, b, 0.0f);
}
__device__ __forceinline__ void RadianceAtomicAdd(float *radiance, uint32_t slot, float r, float g, float b) {
    float *p = radiance + 4ull * slot;
    atomicAdd(p, r), atomicAdd(p + 1, g), atomicAdd(p + 2, b);
}
#endif

struct Counters {             // device-resident
    uint32_t queue[2];        // entries in path queue 0 / 1 (zeroed per batch)
    uint32_t shadow;          // entries in the shadow queue
    uint32_t work_primary;    // next unclaimed ray of the persistent traversal loops
    uint32_t work_trace;
    uint32_t work_tail;       // next unclaimed entry of the tail kernel
    uint32_t tail_taken;      // != 0 once k_tail has taken over the batch's remaining paths (depth of the take-over + 1)
    uint32_t settle_ticket;   // CTAs of k_settle that have finished (the last one resets the counters)
    uint32_t work_packets;    // next unclaimed NEE ray of k_trace's packet loop (coherent shadow rays, traverse_packet.cuh)
    uint32_t bin_count[2][kNumShadeBins]; // entries of path queue 0 / 1 per shading bin (scenes with several BSDF models)
    ClassCounters cls[3];     // primary / extend / shadow traversal statistics (zeroed per render)
};
constexpr size_t kCountersPerBatchBytes = 36 + 2 * kNumShadeBins * sizeof(uint32_t); // the part of Counters that is zeroed for every batch

struct BatchParams {
    DCamera camera;
    uint32_t width, height, spp;
    float spp_inv;
    uint2 key;                // Philox key (seed)
    uint32_t tiles_x, num_tiles;
    uint32_t tile_rank, tile_world;
    uint32_t pixel_begin;     // first local pixel of this batch
    uint32_t pixel_count;     // local pixels in this batch
    uint32_t sample_begin;    // first sample index of this batch
    uint32_t sample_count;    // samples per pixel in this batch (slots = pixel_count * sample_count)
    // Visibility pre-pass (k_cull_tiles): the local pixels whose camera rays can reach geometry, in ascending order: the
    // job's pixel list.  nullptr = every local pixel.
    const uint32_t *active_pixels;
    // Progressive preview (renderer.cpp:97-138): one sample per pixel per call, every pixel with the same sub-pixel offset
    // (VdC_2, VdC_3 of frame_index + 1); the sample index of the random-number stream is the frame index.
    uint32_t progressive;
    float progressive_u, progressive_v;
};

// index in the job's pixel list -> local pixel (position in this rank's tile-ordered pixel buffer)
__host__ __device__ inline uint32_t JobPixelToLocal(const BatchParams &p, uint32_t job_pixel) {
    if (p.active_pixels == nullptr) return job_pixel;
    return p.active_pixels[job_pixel];
}

// local pixel -> image coordinates; false for padding pixels of edge tiles / tiles past the end.
__host__ __device__ inline bool LocalPixelToImage(const BatchParams &p, uint32_t local_pixel, uint32_t *i, uint32_t *j) {
    const uint32_t tile_local = local_pixel / kTilePixels, in_tile = local_pixel % kTilePixels;
    const uint32_t tile = tile_local * p.tile_world + p.tile_rank;
    if (tile >= p.num_tiles) return false;
    *i = (tile % p.tiles_x) * kTileSize + (in_tile % kTileSize);
    *j = (tile / p.tiles_x) * kTileSize + (in_tile / kTileSize);
    return *i < p.width && *j < p.height;
}

#if defined(__CUDACC__)
// Random-number counter (pixel, sample, depth) of the alpha tests along the ray of sample slot `slot`.
__device__ __forceinline__ uint3 SlotCounter(const BatchParams &bp, uint32_t slot, uint32_t depth) {
    uint32_t px = 0, py = 0;
    LocalPixelToImage(bp, JobPixelToLocal(bp, bp.pixel_begin + slot / bp.sample_count), &px, &py);
    return make_uint3(py * bp.width + px, bp.sample_begin + slot % bp.sample_count, depth);
}
#endif

struct LaunchConfig {
    int blocks;               // persistent grid: SMs x resident CTAs
    int threads;
    cudaStream_t stream;
    bool stats;
    int top_nodes;            // wide-BVH nodes staged in shared memory per CTA (80 B each), 0 = none
    int refill;               // idle lanes per warp that trigger a ray refill in the traversal loops
    int min_inner;            // binary layout: lanes still walking inner nodes below which a warp switches to its pending leaves
    int tri_min;              // wide layout: lanes holding triangles below which they are postponed in favour of inner nodes
    int shade_only;           // the one BSDF type every scattering surface of the scene has, or -1 (generic shading kernel)
    int packets;              // kPackets* bits: which ray sets of this launch walk the binary tree as warp packets
};
// Warp-packet traversal (traverse_packet.cuh) of the ray sets that are coherent by construction.
constexpr int kPacketsPrimary = 1;  // camera rays: 32 consecutive sample slots belong to one pixel
constexpr int kPacketsShadow = 2;   // NEE rays of this launch (first vertex, towards delta lights: neighbouring origins, one direction)

// kernel launchers (wavefront.cu)
// LaunchPrimary / LaunchTrace also file the traced queue's hits into the shading bins when `bins.lists` is set (one
// extra small kernel, k_bin_hits); they return the number of kernels launched.
int LaunchPrimary(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, PathQueue q, ShadeBins bins,
                   float *radiance, uint32_t capacity, Counters *counters);
// `depth` = index of the path vertex the rays leave from: with `bp` it keys the alpha-test random numbers (opacity masks).
// which < 0: no closest-hit rays this round (last bounce), only the NEE rays.
int LaunchTrace(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue q, int which,
                 ShadeBins bins, ShadowQueue sq, float *radiance, uint32_t capacity, Counters *counters);
// One launch per shading bin in use (or a single launch when the scene is not binned); returns the number of launches.
int LaunchShade(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue qin,
                int which_in, PathQueue qout, ShadeBins bins, ShadowQueue sq, float *radiance, Counters *counters, uint32_t capacity);
// Visibility pre-pass: masks[t] = the pixels of local tile t whose camera rays can reach one of the scene's cull boxes, offsets[t] =
// exclusive prefix sum of their counts, pixel_list = the ascending list of those pixels, counts = {pixels, non-empty tiles}.
void LaunchCullTiles(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t num_local_tiles,
                     unsigned long long *masks, uint32_t *offsets, uint32_t *pixel_list, uint32_t *counts);
// Path-at-a-time tail (tail_kernel.cu): launched after the traversal of bounce `depth - 1`; takes queue `which` over and
// finishes its paths if it holds at most `threshold` entries, else returns at once.  `depth` = the bounce shade would run.
void LaunchTail(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue q, int which,
                float *radiance, uint32_t capacity, Counters *counters, uint32_t threshold);
// Test hook (b200pt_debug_trace): rays[0..n) through the persistent traversal loop (or the per-lane one) of the scene's tree.
void LaunchDebugTrace(const LaunchConfig &lc, const DeviceScene &scene, const b200pt_debug_ray *rays, uint32_t n, bool any_hit, bool single,
                      bool raw_prim, bool packet, b200pt_debug_hit *out, uint32_t *work_counter);
// Test hook (b200pt_debug_eval, debug_eval.cu): leaf functions of the shading stage at caller-supplied inputs.
void LaunchDebugEval(cudaStream_t stream, const DeviceScene &scene, uint32_t what, uint32_t id, uint32_t n, const float *in, float *out);
// Test hook (b200pt_debug_render_replay, debug_eval.cu): a frame with the reference's loop shape and random-number stream.
void LaunchDebugReplay(cudaStream_t stream, const DeviceScene &scene, const BatchParams &bp, float *frame, uint32_t trace_pixel);
constexpr uint32_t kMaxTailDepth = 4096; // = kMaxRounds of the host loop
// Adds the contributions of the shadow rays the last k_trace marked unoccluded, then resets queue `which_queue` (>= 0), the
// shadow queue (reset_shadow) and the traversal work counter for the next bounce.
// unique_slots: at most one NEE ray per vertex (plain adds instead of atomics).
void LaunchSettle(const LaunchConfig &lc, Counters *counters, int which_queue, bool reset_shadow, ShadowQueue sq, float *radiance,
                  uint32_t capacity, bool unique_slots);
constexpr float kShadowUnoccluded = -1.0f; // written over ShadowQueue::tmax by k_trace (a real tmax is never negative)
void LaunchResolve(const LaunchConfig &lc, const BatchParams &bp, const float *radiance, uint32_t capacity, float *accum);
void LaunchFinalize(const LaunchConfig &lc, const BatchParams &bp, uint32_t num_local_pixels, const float *accum,
                    float *frame, float *tiles);
// Progressive preview: frame = (index * frame + sample) / (index + 1), sRGB copy with row 0 at the bottom (renderer.cpp:118-136).
void LaunchFinalizeProgressive(const LaunchConfig &lc, const BatchParams &bp, uint32_t num_local_pixels, const float *accum,
                               uint32_t frame_index, float *frame, float *frame_srgb);
void LaunchAssemble(const LaunchConfig &lc, uint32_t width, uint32_t height, uint32_t tile_world, uint32_t pixels_per_rank,
                    const float *gathered, float *frame);

} // namespace b200pt
