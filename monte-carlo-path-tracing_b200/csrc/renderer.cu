// renderer.cu — the C ABI (include/b200pt.h) and the host side of the wavefront loop.
//
// b200pt_create  replaces csrt::Renderer::Renderer (src/renderer/renderer.cpp:259-348): flatten, build the
//                BVH, derive the tables, upload once to HBM.
// b200pt_render  replaces csrt::Renderer::Draw (renderer.cpp:678-721): runs the wavefront rounds and
//                leaves width*height*3 linear floats in the caller's frame.
// There is no CPU fallback: every entry point that needs the GPU fails with B200PT_ECUDA when CUDA is unusable.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "b200pt.h"
#include "host_scene.hpp"
#include "host_util.hpp"
#include "traverse_wide.cuh"
#include "wavefront.cuh"

using namespace b200pt;

namespace {

// Upper bound of sample slots per batch.  Deep bounces keep the GPU full only when a batch carries many
// paths (Dragon 1024^2 x 256 spp: 16 Mi slots -> 166 ms, 256 Mi slots -> 75 ms), so the default takes what a
// 180 GB part affords: ~190 B of queue state per slot, at most 40 % of the free HBM.
constexpr uint64_t kDefaultPathsInFlight = 1ull << 28;
constexpr uint32_t kMaxRounds = 4096;                  // hard stop for max_depth = -1 scenes (RR ends paths long before)

double g_upload_alloc_ms = 0.0, g_upload_copy_ms = 0.0; // B200PT_VERBOSE_CREATE: where UploadScene's time goes

template <typename T>
struct DeviceArray {
    T *ptr = nullptr;
    size_t count = 0;
    ~DeviceArray() { Free(); }
    void Free() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    cudaError_t Alloc(size_t n) {
        Free();
        count = n;
        if (n == 0) return cudaSuccess;
        return cudaMalloc(&ptr, n * sizeof(T));
    }
    template <class A>
    cudaError_t Upload(const std::vector<T, A> &src) {
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = Alloc(src.size());
        const auto t1 = std::chrono::steady_clock::now();
        if (e != cudaSuccess || src.empty()) return e;
        e = cudaMemcpy(ptr, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
        g_upload_alloc_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        g_upload_copy_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
        return e;
    }
};

} // namespace

namespace {
constexpr int kMaxArenas = 8;
constexpr int kPollRing = 8;         // survivor-count polls in flight per arena
constexpr uint32_t kPollLag = 4;     // a poll is read this many rounds after it was enqueued, so the host never drains the GPU
constexpr uint32_t kFirstPollDepth = 4;

// One batch of paths in flight.
struct Arena {
    PathQueue queue[2]{};
    ShadowQueue shadow{};
    float *radiance = nullptr;       // one float4 (r, g, b, unused) per sample slot (RadianceAdd, wavefront.cuh)
    uint32_t *bin_lists = nullptr;   // shading bins (scenes with several BSDF models): `capacity` entries per bin in use
    Counters *counters = nullptr;
    cudaStream_t stream = nullptr;   // owned side stream (a single-arena render runs on the caller's stream instead)
    cudaEvent_t done = nullptr;
    cudaEvent_t poll_event[kPollRing]{};
    uint32_t *poll_count = nullptr;  // pinned, kPollRing entries
};
} // namespace

struct b200pt_context {
    int device = 0;
    std::string error;
    HostScene host;
    DeviceScene scene{};
    b200pt_stats stats{};

    DeviceArray<uint8_t> tree;       // binary nodes, then triangle vertices, in one allocation
    DeviceArray<WideNode> wide_nodes;
    DeviceArray<TriShade> tri_shade;
    DeviceArray<uint8_t> tri_bsdf_type;
    DeviceArray<AnalyticPrim> analytic;
    DeviceArray<DInstance> instances;
    DeviceArray<DBsdf> bsdfs;
    DeviceArray<DTexture> textures;
    DeviceArray<float> pixels;
    DeviceArray<DMedium> media;
    DeviceArray<DEmitter> emitters;
    DeviceArray<float> envmap_tables, kc_brdf, kc_albedo, cdf_area_light, light_tri_cdf;
    DeviceArray<uint32_t> map_area_light_instance, light_tri_ids;
    DeviceArray<float> cull_boxes;
    DeviceArray<float> fine_cull_boxes;
    DeviceArray<unsigned long long> tile_masks;  // visibility pre-pass: per local tile, the pixels that can see geometry
    DeviceArray<uint32_t> pixel_list;            // ... their local indices in ascending order, followed by {pixels, non-empty tiles}
    DeviceArray<uint32_t> tile_offsets;          // ... position of each tile's first surviving pixel in that list
    DeviceArray<uint32_t> cull_fine_begin;
    bool tile_cull = true;        // B200PT_TILE_CULL=0 turns the pre-pass off
    int shade_only = -1;          // see LaunchConfig::shade_only
    uint32_t tail_paths = 32768;  // B200PT_TAIL_PATHS: survivor count below which k_tail finishes a batch's paths (0 = never);
                                  // profiles/r01_sweep_tail_kernel.log: larger take-overs lose to the per-bounce launches

    // wavefront state: up to kMaxArenas independent batches in flight, each with private queues, counters and stream,
    // so that the latency-bound tail of one batch's launches is filled by another batch's work
    uint64_t capacity = 0;        // sample slots per arena as currently carved
    uint64_t max_capacity = 0;    // upper bound on slots over all arenas (create option / default)
    uint32_t shadow_per_vertex = 1;
    DeviceArray<float> wave;      // one allocation carved into the arenas' SoA queues
    Arena arenas[kMaxArenas];
    int num_arenas = 0;           // B200PT_ARENAS; 0 = automatic (profiles/r01_sweep_arenas_occupancy.log: 2 is best on Dragon, 4 on matpreview)
    DeviceArray<Counters> counters; // one per arena
    DeviceArray<float> accum;     // per local pixel RGB sums
    DeviceArray<float> frame;     // staging for b200pt_render (host frame)
    uint32_t *pinned_count = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_fork = nullptr;
    bool timing_pending = false;
    bool rendered_once = false;   // ev_end has been recorded at least once
    int num_sms = 148;
    // launch tunables (overridable through the environment for experiments: B200PT_TOP_NODES, B200PT_REFILL, B200PT_CTAS_PER_SM)
    int top_nodes = 0, refill = 20, ctas_per_sm = 4, min_inner = 8; // top_nodes = 0: no shared-memory staging (profiles/r01_sweep_sel3_topnodes.log)
    // B200PT_PACKETS (bit mask): 1 = camera rays as warp packets (default: Dragon 38.96 -> 37.45 ms, 1080p 257 -> 249 ms),
    // 2 = first-vertex NEE rays towards a single delta light as warp packets, 4 = NEE packets whatever the emitters are,
    // 8 = NEE packets at every depth.  2 / 4 / 8 are measured variants, all slower than per-lane ray replacement (Dragon k_trace
    // 28.4 -> 31.0 ms with 2: the NEE rays of neighbouring hits are not coherent enough, profiles/r02_sweep_packets.log).
    int packets = 1;

    bool nee_coherent = false;    // one emitter, of a delta kind, and no area lights: the NEE rays of neighbouring hits run in parallel
    int tri_min = 8;              // B200PT_TRI_MIN: triangle postponing threshold of the wide traversal (LaunchConfig::tri_min)
    uint32_t wide_top_nodes = 0;  // nodes at the head of the wide node array that were laid out breadth-first
    // B200PT_STATS_TIMING: (class, begin, end) per launch, resolved in b200pt_get_stats
    struct TimedLaunch {
        int cls;
        cudaEvent_t begin, end;
    };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
    uint64_t class_launches[kNumClasses] = {};

    cudaEvent_t NextEvent() {
        if (events_used == event_pool.size()) {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            event_pool.push_back(e);
        }
        return event_pool[events_used++];
    }

    ~b200pt_context() {
        for (Arena &a : arenas) {
            if (a.poll_count) cudaFreeHost(a.poll_count);
            if (a.done) cudaEventDestroy(a.done);
            for (cudaEvent_t e : a.poll_event)
                if (e) cudaEventDestroy(e);
            if (a.stream) cudaStreamDestroy(a.stream);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (pinned_count) cudaFreeHost(pinned_count);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (stream) cudaStreamDestroy(stream);
        for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
    }

    int Fail(int code, const std::string &msg) {
        error = msg;
        SetGlobalError(code, msg);
        return code;
    }
    int CudaFail(cudaError_t e, const char *what) {
        // same wording as the reference's CUDA path (renderer.cpp:696-710)
        return Fail(B200PT_ECUDA, std::string("CUDA error : \"") + cudaGetErrorString(e) + "\" (" + what + ").");
    }
};

namespace {

#define CU_CHECK(ctx, call)                                         \
    do {                                                            \
        const cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return (ctx)->CudaFail(e_, #call);   \
    } while (0)

int UploadScene(b200pt_context *c, const b200pt_scene_desc &desc) {
    HostScene &h = c->host;
    const size_t node_bytes = (h.nodes.size() * sizeof(BvhNode) + 255) & ~size_t(255), vert_bytes = h.tri_verts.size() * sizeof(TriVerts);
    CU_CHECK(c, c->tree.Alloc(node_bytes + vert_bytes));
    if (!h.nodes.empty()) CU_CHECK(c, cudaMemcpy(c->tree.ptr, h.nodes.data(), h.nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
    if (vert_bytes) CU_CHECK(c, cudaMemcpy(c->tree.ptr + node_bytes, h.tri_verts.data(), vert_bytes, cudaMemcpyHostToDevice));
    CU_CHECK(c, c->wide_nodes.Upload(h.wide_nodes));
    CU_CHECK(c, c->tri_shade.Upload(h.tri_shade));
    CU_CHECK(c, c->tri_bsdf_type.Upload(h.tri_bsdf_type));
    CU_CHECK(c, c->analytic.Upload(h.analytic));
    CU_CHECK(c, c->instances.Upload(h.instances));
    CU_CHECK(c, c->bsdfs.Upload(h.bsdfs));
    CU_CHECK(c, c->textures.Upload(h.textures));
    CU_CHECK(c, c->media.Upload(h.media));
    CU_CHECK(c, c->emitters.Upload(h.emitters));
    CU_CHECK(c, c->envmap_tables.Upload(h.envmap_tables));
    CU_CHECK(c, c->kc_brdf.Upload(h.kc_brdf_avg));
    CU_CHECK(c, c->kc_albedo.Upload(h.kc_albedo_avg));
    CU_CHECK(c, c->cdf_area_light.Upload(h.cdf_area_light));
    CU_CHECK(c, c->map_area_light_instance.Upload(h.map_area_light_instance));
    CU_CHECK(c, c->light_tri_cdf.Upload(h.light_tri_cdf));
    CU_CHECK(c, c->light_tri_ids.Upload(h.light_tri_ids));
    CU_CHECK(c, c->cull_boxes.Upload(h.cull_boxes));
    CU_CHECK(c, c->fine_cull_boxes.Upload(h.fine_cull_boxes));
    CU_CHECK(c, c->cull_fine_begin.Upload(h.cull_fine_begin));
    CU_CHECK(c, c->pixels.Alloc(desc.num_pixels));
    if (desc.num_pixels)
        CU_CHECK(c, cudaMemcpy(c->pixels.ptr, desc.pixels, desc.num_pixels * sizeof(float), cudaMemcpyHostToDevice));

    DeviceScene &s = c->scene;
    s.nodes = reinterpret_cast<BvhNode *>(c->tree.ptr), s.num_nodes = static_cast<uint32_t>(h.nodes.size());
    s.wide_nodes = c->wide_nodes.ptr, s.num_wide_nodes = static_cast<uint32_t>(h.wide_nodes.size());
    s.tri_verts = reinterpret_cast<TriVerts *>(c->tree.ptr + node_bytes), s.num_tris = static_cast<uint32_t>(h.tri_verts.size());
    s.tri_shade = c->tri_shade.ptr;
    s.tri_bsdf_type = c->tri_bsdf_type.ptr;
    s.analytic = c->analytic.ptr, s.num_analytic = static_cast<uint32_t>(h.analytic.size());
    s.instances = c->instances.ptr, s.num_instances = static_cast<uint32_t>(h.instances.size());
    s.bsdfs = c->bsdfs.ptr, s.num_bsdfs = static_cast<uint32_t>(h.bsdfs.size());
    s.textures = c->textures.ptr, s.num_textures = static_cast<uint32_t>(h.textures.size());
    s.pixels = c->pixels.ptr;
    s.media = c->media.ptr, s.num_media = static_cast<uint32_t>(h.media.size());
    s.emitters = c->emitters.ptr;
    s.envmap_tables = c->envmap_tables.ptr;
    s.kc_brdf_avg = c->kc_brdf.ptr, s.kc_albedo_avg = c->kc_albedo.ptr;
    s.cdf_area_light = c->cdf_area_light.ptr;
    s.map_area_light_instance = c->map_area_light_instance.ptr;
    s.light_tri_cdf = c->light_tri_cdf.ptr, s.light_tri_ids = c->light_tri_ids.ptr;
    memcpy(s.scene_bmin, h.scene_bmin, 12), memcpy(s.scene_bmax, h.scene_bmax, 12);
    s.cull_boxes = c->cull_boxes.ptr, s.num_cull_boxes = static_cast<uint32_t>(h.cull_boxes.size() / 6);
    s.cull_fine_begin = c->cull_fine_begin.ptr;
    s.fine_cull_boxes = c->fine_cull_boxes.ptr, s.num_fine_cull_boxes = static_cast<uint32_t>(h.fine_cull_boxes.size() / 6);
    if (getenv("B200PT_PIXEL_CULL") && atoi(getenv("B200PT_PIXEL_CULL")) == 0) s.num_fine_cull_boxes = 0; // experiments: tiles only
    s.integrator = h.integrator;

    // One BSDF type on every scattering surface (area lights and BSDF-less surfaces aside)?  Then the shading kernel
    // specialised to it runs.
    {
        int only = -1;
        bool mixed = false;
        for (const DInstance &in : h.instances) {
            if (in.id_bsdf == kInvalid) continue;
            const int type = static_cast<int>(h.bsdfs[in.id_bsdf].type);
            if (type == B200PT_BSDF_AREA_LIGHT) continue;
            if (only >= 0 && only != type) mixed = true;
            only = type;
        }
        c->shade_only = (mixed || getenv("B200PT_GENERIC_SHADE")) ? -1 : only;
    }
    {
        const DIntegrator &ig = s.integrator;
        bool delta = ig.num_emitters == 1 && ig.num_area_lights == 0;
        for (const DEmitter &e : h.emitters)
            if (e.type != B200PT_EMIT_POINT && e.type != B200PT_EMIT_SPOT && e.type != B200PT_EMIT_DIRECTIONAL && e.type != B200PT_EMIT_SUN) delta = false;
        c->nee_coherent = delta;
    }
    c->stats.num_bvh_nodes = h.wide_nodes.empty() ? h.nodes.size() : h.wide_nodes.size();
    c->stats.bvh_width = h.wide_nodes.empty() ? 2u : 8u;
    c->stats.bvh_depth = h.wide_depth;
    c->wide_top_nodes = h.wide_top_nodes;
    c->stats.num_triangles = h.tri_verts.size();
    c->stats.num_prims = h.tri_verts.size() + h.analytic.size();
    c->stats.bvh_build_ms = h.bvh_build_ms;
    c->stats.bvh_gpu_ms = h.bvh_gpu_ms;
    // the big host copies are not needed any more
    std::vector<BvhNode>().swap(h.nodes);
    std::vector<WideNode>().swap(h.wide_nodes);
    decltype(h.tri_verts)().swap(h.tri_verts);
    decltype(h.tri_shade)().swap(h.tri_shade);
    return B200PT_OK;
}

// Words of wavefront state per sample slot.
uint64_t WordsPerSlot(const b200pt_context *c) {
    const bool vol = c->scene.integrator.type == B200PT_INTEGRATOR_VOLPATH;
    const uint32_t shadow_per_vertex = c->scene.integrator.num_emitters + (c->scene.integrator.num_area_lights ? 1u : 0u);
    const uint64_t words_per_queue = 11 + (vol ? 4 : 0) + 4; // 10 floats + slot (+ medium, wo) + HitRec
    return 2 * words_per_queue + 11 * std::max(1u, shadow_per_vertex) + 4 + __builtin_popcount(c->scene.integrator.shade_bins);
}

// Carve the SoA queues of `arenas` arenas out of one allocation; each arena gets `wanted_per_arena` sample slots, or
// what fits in 40 % of the free HBM.  The allocation only ever grows; carving itself is pointer arithmetic.
int CarveWavefront(b200pt_context *c, uint64_t wanted_per_arena, int arenas) {
    const bool vol = c->scene.integrator.type == B200PT_INTEGRATOR_VOLPATH;
    c->shadow_per_vertex = c->scene.integrator.num_emitters + (c->scene.integrator.num_area_lights ? 1u : 0u);
    const uint64_t words_per_slot = WordsPerSlot(c);
    uint64_t capacity = std::max<uint64_t>((wanted_per_arena + 1023ull) & ~1023ull, 1024);
    if (c->wave.count < words_per_slot * capacity * arenas) {
        CU_CHECK(c, cudaDeviceSynchronize());
        c->wave.Free();
        size_t free_bytes = 0, total_bytes = 0;
        CU_CHECK(c, cudaMemGetInfo(&free_bytes, &total_bytes));
        const uint64_t fits = static_cast<uint64_t>(free_bytes * 0.4) / (words_per_slot * sizeof(float) * arenas);
        capacity = std::max<uint64_t>(std::min(capacity, fits) & ~1023ull, 1024);
        CU_CHECK(c, c->wave.Alloc(words_per_slot * capacity * arenas));
    }
    const uint64_t shadow_cap = capacity * std::max(1u, c->shadow_per_vertex);
    float *p = c->wave.ptr;
    auto take = [&](uint64_t n) {
        float *r = p;
        p += n;
        return r;
    };
    for (int a = 0; a < arenas; ++a) {
        Arena &ar = c->arenas[a];
        for (int k = 0; k < 2; ++k) {
            PathQueue &q = ar.queue[k];
            q.hit = reinterpret_cast<HitRec *>(take(4 * capacity)); // first: keeps 16-byte alignment
            q.ox = take(capacity), q.oy = take(capacity), q.oz = take(capacity);
            q.dx = take(capacity), q.dy = take(capacity), q.dz = take(capacity);
            q.tr = take(capacity), q.tg = take(capacity), q.tb = take(capacity);
            q.pdf = take(capacity);
            q.slot = reinterpret_cast<uint32_t *>(take(capacity));
            q.medium = nullptr, q.wx = q.wy = q.wz = nullptr;
            if (vol) {
                q.medium = reinterpret_cast<uint32_t *>(take(capacity));
                q.wx = take(capacity), q.wy = take(capacity), q.wz = take(capacity);
            }
        }
        ShadowQueue &sq = ar.shadow;
        sq.ox = take(shadow_cap), sq.oy = take(shadow_cap), sq.oz = take(shadow_cap);
        sq.dx = take(shadow_cap), sq.dy = take(shadow_cap), sq.dz = take(shadow_cap);
        sq.tmax = take(shadow_cap);
        sq.cr = take(shadow_cap), sq.cg = take(shadow_cap), sq.cb = take(shadow_cap);
        sq.slot = reinterpret_cast<uint32_t *>(take(shadow_cap));
        ar.radiance = take(4 * capacity); // 16-byte aligned: every array before it is a multiple of `capacity` (x 1024) floats
        const int bins_in_use = __builtin_popcount(c->scene.integrator.shade_bins);
        ar.bin_lists = bins_in_use ? reinterpret_cast<uint32_t *>(take(static_cast<uint64_t>(bins_in_use) * capacity)) : nullptr;
        ar.counters = c->counters.ptr + a;
    }
    c->capacity = capacity;
    return B200PT_OK;
}

struct ResolvedOpts {
    uint32_t width, height, spp, tile_rank, tile_world;
    uint64_t seed;
    bool counters, timing, tile_cull;
    bool progressive = false;     // b200pt_render_progressive_device: one sample per pixel, running mean into the frame
    uint32_t frame_index = 0;
    float *frame_srgb = nullptr;
};

int ResolveOpts(b200pt_context *c, const b200pt_render_opts *o, ResolvedOpts *r) {
    const b200pt_camera &cam = c->host.camera;
    r->width = (o && o->width) ? o->width : static_cast<uint32_t>(cam.width);
    r->height = (o && o->height) ? o->height : static_cast<uint32_t>(cam.height);
    r->spp = (o && o->spp) ? o->spp : cam.spp;
    r->seed = o ? o->seed : 0;
    r->tile_world = (o && o->tile_world) ? o->tile_world : 1;
    r->tile_rank = o ? o->tile_rank : 0;
    r->counters = o && (o->collect_stats & B200PT_STATS_COUNTERS);
    r->timing = o && (o->collect_stats & B200PT_STATS_TIMING);
    r->tile_cull = !(o && (o->flags & B200PT_RENDER_NO_TILE_CULL));
    if (r->width == 0 || r->height == 0 || r->spp == 0) return c->Fail(B200PT_EINVAL, "width, height and spp must be positive.");
    if (r->tile_rank >= r->tile_world) return c->Fail(B200PT_EINVAL, "tile_rank must be < tile_world.");
    if (static_cast<uint64_t>(r->width) * r->height > (1ull << 30)) return c->Fail(B200PT_EINVAL, "frame too large.");
    return B200PT_OK;
}

uint32_t PixelsPerRank(uint32_t width, uint32_t height, uint32_t world) {
    const uint32_t tiles = ((width + kTileSize - 1) / kTileSize) * ((height + kTileSize - 1) / kTileSize);
    return ((tiles + world - 1) / world) * kTilePixels;
}

// math.hpp:29-41 (GetVanDerCorputSequence<base>), including its float round trip of the index.
float VanDerCorputHost(uint32_t base, uint32_t index) {
    const float base_inv = 1.0f / base;
    float result = 0.0f, frac = base_inv;
    while (index > 0) {
        result += frac * (index % base);
        index = static_cast<uint32_t>(index * base_inv);
        frac *= base_inv;
    }
    return result;
}

// The wavefront loop.  Everything is enqueued on `stream`; the host only synchronises to poll the
// survivor count once RR is active (every 4 rounds), so shallow scenes run without host round trips.
int RenderOnStream(b200pt_context *c, const ResolvedOpts &ro, float *frame_dev, float *tiles_dev, cudaStream_t stream) {
    CU_CHECK(c, cudaSetDevice(c->device));
    BatchParams bp{};
    bp.camera = MakeCamera(c->host.camera, ro.width, ro.height);
    bp.width = ro.width, bp.height = ro.height, bp.spp = ro.spp;
    bp.spp_inv = 1.0f / ro.spp;
    bp.key = make_uint2(static_cast<uint32_t>(ro.seed), static_cast<uint32_t>(ro.seed >> 32) ^ 0xB200C0DEu);
    bp.tiles_x = (ro.width + kTileSize - 1) / kTileSize;
    bp.num_tiles = bp.tiles_x * ((ro.height + kTileSize - 1) / kTileSize);
    bp.tile_rank = ro.tile_rank, bp.tile_world = ro.tile_world;
    const uint32_t local_pixels = PixelsPerRank(ro.width, ro.height, ro.tile_world);
    if (ro.progressive) {
        bp.progressive = 1;
        bp.progressive_u = VanDerCorputHost(2, ro.frame_index + 1);
        bp.progressive_v = VanDerCorputHost(3, ro.frame_index + 1);
    }

    if (c->accum.count < 3ull * local_pixels) CU_CHECK(c, c->accum.Alloc(3ull * local_pixels));
    LaunchConfig lc{};
    lc.threads = 256;
    lc.stream = stream;
    lc.stats = ro.counters;
    lc.top_nodes = c->top_nodes;
    lc.refill = c->refill;
    lc.min_inner = c->min_inner;
    lc.tri_min = c->tri_min;
    lc.blocks = c->num_sms * c->ctas_per_sm;
    lc.shade_only = c->shade_only;
    lc.packets = c->packets & kPacketsPrimary;

    uint64_t launches = 0;
    c->timed.clear();
    c->events_used = 0;
    for (uint64_t &n : c->class_launches) n = 0;
    // One render in flight per handle: accum, the arenas and the counters are shared state, so work submitted on ANOTHER
    // stream is ordered after the previous render of this handle (a no-op on the same stream).
    if (c->rendered_once) CU_CHECK(c, cudaStreamWaitEvent(stream, c->ev_end, 0));
    c->rendered_once = true;
    CU_CHECK(c, cudaEventRecord(c->ev_begin, stream));

    // Visibility pre-pass: which of this rank's tiles can see geometry at all.  Escaped camera rays only carry radiance
    // when an environment map or a sun disc exists (path.cpp:24-35), so without those the other tiles are exactly black
    // and none of their samples needs a ray.  The job's pixel list shrinks to the active tiles (part of the timed render).
    const uint32_t local_tiles = local_pixels / kTilePixels;
    uint32_t job_pixels = local_pixels, active_tiles = local_tiles;
    const DIntegrator &ig = c->scene.integrator;
    if (ro.tile_cull && c->tile_cull && ig.id_envmap == kInvalid && ig.id_sun == kInvalid) {
        if (c->tile_masks.count < local_tiles) {
            CU_CHECK(c, c->tile_masks.Alloc(local_tiles));
            CU_CHECK(c, c->pixel_list.Alloc(static_cast<uint64_t>(local_pixels) + 2ull));
            CU_CHECK(c, c->tile_offsets.Alloc(local_tiles));
        }
        launches += 3;
        c->class_launches[kClassOther] += 3;
        LaunchCullTiles(lc, c->scene, bp, local_tiles, c->tile_masks.ptr, c->tile_offsets.ptr, c->pixel_list.ptr, c->pixel_list.ptr + local_pixels);
        CU_CHECK(c, cudaMemcpyAsync(c->pinned_count, c->pixel_list.ptr + local_pixels, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CU_CHECK(c, cudaStreamSynchronize(stream));
        job_pixels = c->pinned_count[0];
        active_tiles = c->pinned_count[1];
        bp.active_pixels = c->pixel_list.ptr;
    }
    c->stats.active_tiles = active_tiles;
    c->stats.active_pixels = job_pixels;
    c->stats.local_tiles = local_tiles;

    // Batches.  The job's pixels are cut into chunks, one arena works on one chunk at a time (all its sample batches in
    // order, so no two arenas ever add into the same pixel), chunks are dealt round-robin to the arenas.  Several arenas
    // in flight hide the tail of every launch — a traversal launch cannot end before its slowest ray (~0.2 ms on Dragon,
    // at every bounce) — behind another batch's work.  Per-class event timing needs serial launches: one arena then.
    const uint64_t job_slots = static_cast<uint64_t>(job_pixels) * ro.spp;
    // Binned scenes launch one (often small) shading kernel per BSDF model and bounce: more batches in flight pay off there.
    const int arenas_wanted = c->num_arenas > 0 ? c->num_arenas : (ig.shade_bins ? 4 : 2);
    const int S = (ro.timing || job_slots < (1ull << 21)) ? 1 : std::max(1, std::min(arenas_wanted, kMaxArenas));
    {
        const uint64_t total = std::min<uint64_t>(c->max_capacity, std::max<uint64_t>(job_slots, 1024));
        const int rc = CarveWavefront(c, (total + S - 1) / S, S);
        if (rc != B200PT_OK) return rc;
    }
    const uint32_t capacity = static_cast<uint32_t>(c->capacity); // per arena
    const uint32_t chunks_per_arena = std::max<uint32_t>(1, static_cast<uint32_t>((job_pixels + static_cast<uint64_t>(S) * capacity - 1) / (static_cast<uint64_t>(S) * capacity)));
    const uint32_t num_chunks = chunks_per_arena * S;
    const uint32_t pixels_per_chunk = std::max<uint32_t>(1, (job_pixels + num_chunks - 1) / num_chunks);
    const uint32_t samples_per_batch = std::max<uint32_t>(1, std::min<uint32_t>(ro.spp, capacity / pixels_per_chunk));
    // path.cpp:57-60: `depth < depth_rr || (depth < depth_max && rand < pdf_rr)` keeps a path going until
    // max(depth_rr, depth_max) — a scene with max_depth below rr_depth (parser default 5) still walks rr_depth - 1 vertices.
    const uint32_t max_rounds = std::min<uint32_t>(std::max(ig.depth_rr, ig.depth_max), kMaxRounds);

    CU_CHECK(c, cudaMemsetAsync(c->accum.ptr, 0, 3ull * local_pixels * sizeof(float), stream));
    CU_CHECK(c, cudaMemsetAsync(c->counters.ptr, 0, sizeof(Counters) * kMaxArenas, stream));
    if (S > 1) CU_CHECK(c, cudaEventRecord(c->ev_fork, stream));

    struct Run {              // progress of one arena through its chunks
        cudaStream_t stream = nullptr;
        uint32_t chunk = 0;   // next chunk of this arena (chunk index = chunk * S + arena)
        uint32_t sample_begin = 0;
        bool in_batch = false;
        uint32_t depth = 0;
        int which = 0;
        BatchParams bp{};
    };
    Run runs[kMaxArenas];
    for (int a = 0; a < S; ++a) {
        runs[a].stream = S > 1 ? c->arenas[a].stream : stream;
        runs[a].bp = bp;
        if (S > 1) CU_CHECK(c, cudaStreamWaitEvent(runs[a].stream, c->ev_fork, 0));
    }
    cudaError_t issue_error = cudaSuccess;
    // Issues the next piece of work of arena `a` (the start of a batch, or one bounce of the batch in flight);
    // false when the arena has nothing left.
    auto advance = [&](int a) -> bool {
        Arena &ar = c->arenas[a];
        Run &run = runs[a];
        LaunchConfig la = lc;
        la.stream = run.stream;
        const ShadeBins bins{ar.bin_lists, capacity};
        auto launch = [&](int cls, auto &&fn) { // one kernel launch, counted (and, with B200PT_STATS_TIMING, event-timed) under its class
            ++launches;
            ++c->class_launches[cls];
            if (ro.timing) {
                b200pt_context::TimedLaunch t{cls, c->NextEvent(), c->NextEvent()};
                cudaEventRecord(t.begin, la.stream);
                fn();
                cudaEventRecord(t.end, la.stream);
                c->timed.push_back(t);
            } else {
                fn();
            }
        };
        auto check = [&](cudaError_t e) {
            if (e != cudaSuccess && issue_error == cudaSuccess) issue_error = e;
        };
        if (!run.in_batch) {
            const uint64_t pixel_begin = (static_cast<uint64_t>(run.chunk) * S + a) * pixels_per_chunk;
            if (pixel_begin >= job_pixels) return false;
            run.bp.pixel_begin = static_cast<uint32_t>(pixel_begin);
            run.bp.pixel_count = std::min<uint32_t>(pixels_per_chunk, job_pixels - run.bp.pixel_begin);
            run.bp.sample_begin = run.sample_begin + (ro.progressive ? ro.frame_index : 0u);
            run.bp.sample_count = std::min(samples_per_batch, ro.spp - run.sample_begin);
            const uint64_t nslots = static_cast<uint64_t>(run.bp.pixel_count) * run.bp.sample_count;
            check(cudaMemsetAsync(ar.radiance, 0, nslots * 4 * sizeof(float), la.stream));
            check(cudaMemsetAsync(ar.counters, 0, kCountersPerBatchBytes, la.stream)); // queue lengths + work counters
            launch(kClassPrimary, [&] {
                const int n = LaunchPrimary(la, c->scene, run.bp, ar.queue[0], bins, ar.radiance, capacity, ar.counters);
                launches += n - 1;
                c->class_launches[kClassPrimary] += n - 1;
            });
            run.in_batch = true;
            run.depth = 1;
            run.which = 0;
            return true;
        }
        const uint32_t depth = run.depth;
        auto finish_batch = [&] {
            launch(kClassOther, [&] { LaunchSettle(la, ar.counters, -1, true, ar.shadow, ar.radiance, capacity, c->shadow_per_vertex <= 1); }); // NEE of the last bounce
            launch(kClassOther, [&] { LaunchResolve(la, run.bp, ar.radiance, capacity, c->accum.ptr); });
            run.in_batch = false;
            run.sample_begin += run.bp.sample_count;
            if (run.sample_begin >= ro.spp) {
                run.sample_begin = 0;
                ++run.chunk;
            }
        };
        // From the second bounce on, the tail kernel looks at the survivor queue first: once it is short enough it finishes
        // every remaining path in this one launch and the per-bounce kernels below find nothing to do.
        // Settle first: the NEE contributions of the previous bounce are added before anything of this bounce, whichever
        // kernel handles it (the order of float additions per sample is then the same with and without the tail kernel).
        if (depth > 1) launch(kClassOther, [&] { LaunchSettle(la, ar.counters, run.which ^ 1, true, ar.shadow, ar.radiance, capacity, c->shadow_per_vertex <= 1); });
        if (depth > 1 && c->tail_paths > 0)
            launch(kClassTail, [&] { LaunchTail(la, c->scene, run.bp, depth, ar.queue[run.which], run.which, ar.radiance, capacity, ar.counters, c->tail_paths); });
        launch(kClassShade, [&] {
            const int n = LaunchShade(la, c->scene, run.bp, depth, ar.queue[run.which], run.which, ar.queue[run.which ^ 1], bins, ar.shadow,
                                      ar.radiance, ar.counters, capacity);
            launches += n - 1;
            c->class_launches[kClassShade] += n - 1;
        });
        run.which ^= 1;
        // One traversal launch per bounce: closest hits of the survivors (queue `which`) + occlusion of the NEE rays.
        auto trace = [&](int extend_queue) {
            const bool nee_packets = (c->packets & kPacketsShadow) && (c->nee_coherent || (c->packets & 4)) && (depth == 1 || (c->packets & 8));
            la.packets = nee_packets ? kPacketsShadow : 0;
            launch(kClassExtend, [&] {
                const int n = LaunchTrace(la, c->scene, run.bp, depth, ar.queue[run.which], extend_queue, bins, ar.shadow, ar.radiance, capacity, ar.counters);
                launches += n - 1;
                c->class_launches[kClassExtend] += n - 1;
            });
        };
        if (depth >= max_rounds) {
            if (c->shadow_per_vertex > 0) trace(-1); // only the NEE rays of the last vertex are left
            finish_batch();
            return true;
        }
        if (depth >= kFirstPollDepth) {
            // Survivor count, polled with a lag: the count written by the shade kernel of round d is read while issuing
            // round d + kPollLag, when the copy has long completed, so the host never waits for the GPU to drain.  The
            // rounds issued in between on an empty queue are no-op launches.
            if (depth >= kFirstPollDepth + kPollLag) {
                const uint32_t old = (depth - kPollLag) % kPollRing;
                check(cudaEventSynchronize(ar.poll_event[old]));
                if (ar.poll_count[old] == 0) { // every ray of the rounds since then has been traced already
                    finish_batch();
                    return true;
                }
            }
            const uint32_t cur = depth % kPollRing;
            check(cudaMemcpyAsync(ar.poll_count + cur, &ar.counters->queue[run.which], sizeof(uint32_t), cudaMemcpyDeviceToHost, la.stream));
            check(cudaEventRecord(ar.poll_event[cur], la.stream));
        }
        trace(run.which);
        ++run.depth;
        return true;
    };
    for (bool any = true; any;) {
        any = false;
        for (int a = 0; a < S; ++a) any = advance(a) || any;
    }
    if (issue_error != cudaSuccess) return c->CudaFail(issue_error, "wavefront issue");
    if (S > 1)
        for (int a = 0; a < S; ++a) {
            CU_CHECK(c, cudaEventRecord(c->arenas[a].done, runs[a].stream));
            CU_CHECK(c, cudaStreamWaitEvent(stream, c->arenas[a].done, 0));
        }
    auto launch = [&](int cls, auto &&fn) {
        ++launches;
        ++c->class_launches[cls];
        fn();
    };
    if (ro.progressive)
        launch(kClassOther, [&] { LaunchFinalizeProgressive(lc, bp, local_pixels, c->accum.ptr, ro.frame_index, frame_dev, ro.frame_srgb); });
    else
        launch(kClassOther, [&] { LaunchFinalize(lc, bp, local_pixels, c->accum.ptr, frame_dev, tiles_dev); });
    CU_CHECK(c, cudaEventRecord(c->ev_end, stream));
    CU_CHECK(c, cudaGetLastError());
    c->timing_pending = true;
    c->stats.kernel_launches = launches;
    c->stats.samples = 0;
    // samples actually owned by this rank (edge tiles excluded)
    {
        uint64_t owned = 0;
        for (uint32_t t = ro.tile_rank; t < bp.num_tiles; t += ro.tile_world) {
            const uint32_t x0 = (t % bp.tiles_x) * kTileSize, y0 = (t / bp.tiles_x) * kTileSize;
            owned += static_cast<uint64_t>(std::min<uint32_t>(kTileSize, ro.width - x0)) * std::min<uint32_t>(kTileSize, ro.height - y0);
        }
        c->stats.samples = owned * ro.spp;
    }
    return B200PT_OK;
}

} // namespace

extern "C" {

static int CreateChecked(const b200pt_scene_desc *scene, const b200pt_create_opts *opts, b200pt_handle *out);

// No C++ exception crosses the C boundary (std::bad_alloc on a huge scene, std::system_error from the builder's threads): it
// comes back as an error code with the reference's "error when commit renderer" prefix.
int b200pt_create(const b200pt_scene_desc *scene, const b200pt_create_opts *opts, b200pt_handle *out) {
    if (!scene || !out) return SetGlobalError(B200PT_EINVAL, "b200pt_create: null argument");
    *out = nullptr;
    try {
        return CreateChecked(scene, opts, out);
    } catch (const std::exception &e) {
        return SetGlobalError(B200PT_EINVAL, std::string("error when commit renderer.\n\t") + e.what());
    } catch (...) {
        return SetGlobalError(B200PT_EINVAL, "error when commit renderer.\n\tunknown failure");
    }
}

static int CreateChecked(const b200pt_scene_desc *scene, const b200pt_create_opts *opts, b200pt_handle *out) {
    std::unique_ptr<b200pt_context> c(new b200pt_context());
    int device = opts ? opts->device : -1;
    cudaError_t e = cudaSuccess;
    if (device < 0) e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return SetGlobalError(B200PT_ECUDA, std::string("CUDA error : \"") + cudaGetErrorString(e) + "\" (no usable GPU; there is no CPU fallback).");
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    auto env_int = [](const char *name, int fallback, int lo, int hi) {
        const char *v = getenv(name);
        return v ? std::min(std::max(atoi(v), lo), hi) : fallback;
    };
    c->top_nodes = env_int("B200PT_TOP_NODES", c->top_nodes, 0, kWideTopNodesMax);
    c->tri_min = env_int("B200PT_TRI_MIN", c->tri_min, 0, 32);
    c->refill = env_int("B200PT_REFILL", c->refill, 1, 32);
    c->min_inner = env_int("B200PT_MIN_INNER", c->min_inner, 1, 32);
    c->ctas_per_sm = env_int("B200PT_CTAS_PER_SM", c->ctas_per_sm, 1, 16);
    c->tile_cull = env_int("B200PT_TILE_CULL", 1, 0, 1) != 0;
    c->tail_paths = static_cast<uint32_t>(env_int("B200PT_TAIL_PATHS", static_cast<int>(c->tail_paths), 0, 1 << 24));
    c->packets = env_int("B200PT_PACKETS", c->packets, 0, 15);

    std::string err;
    const auto t0 = std::chrono::steady_clock::now();
    const char *builder_env = getenv("B200PT_BVH_BUILDER"); // "lbvh" / "sah": overrides the create option (experiments)
    const bool gpu_lbvh = builder_env ? std::string(builder_env) == "lbvh" : (opts && (opts->flags & B200PT_CREATE_GPU_LBVH));
    const char *layout_env = getenv("B200PT_BVH_LAYOUT");   // "2" / "8": overrides the create option (experiments)
    const bool bvh8 = layout_env ? atoi(layout_env) == 8 : (opts && (opts->flags & B200PT_CREATE_BVH8));
    const bool verbose = getenv("B200PT_VERBOSE_CREATE") != nullptr; // wall time of the phases of b200pt_create on stderr
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    if (!BuildHostScene(*scene, opts ? opts->max_leaf_size : 0, gpu_lbvh, bvh8, &c->host, &err)) {
        // same prefix as renderer.cpp:343-346
        return SetGlobalError(B200PT_EINVAL, "error when commit renderer.\n\t" + err);
    }
    if (verbose) fprintf(stderr, "[b200pt create] %-28s %8.1f ms\n", "BuildHostScene (total)", since(t0));
    const auto t_upload = std::chrono::steady_clock::now();
    int rc = UploadScene(c.get(), *scene);
    if (rc != B200PT_OK) return rc;
    if (verbose)
        fprintf(stderr, "[b200pt create] %-28s %8.1f ms (cudaMalloc %.1f ms, cudaMemcpy %.1f ms, all creates of this process)\n", "UploadScene", since(t_upload),
                g_upload_alloc_ms, g_upload_copy_ms);
    uint64_t capacity = (opts && opts->max_paths_in_flight) ? opts->max_paths_in_flight : kDefaultPathsInFlight;
    capacity = std::max<uint64_t>(capacity, 1024);
    c->max_capacity = std::min<uint64_t>(capacity, 1ull << 28);
    c->num_arenas = env_int("B200PT_ARENAS", c->num_arenas, 0, kMaxArenas);
    if ((e = c->counters.Alloc(kMaxArenas)) != cudaSuccess) return c->CudaFail(e, "cudaMalloc counters");
    for (Arena &a : c->arenas) {
        if ((e = cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking)) != cudaSuccess) return c->CudaFail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming)) != cudaSuccess) return c->CudaFail(e, "cudaEventCreate");
        for (cudaEvent_t &pe : a.poll_event)
            if ((e = cudaEventCreateWithFlags(&pe, cudaEventDisableTiming)) != cudaSuccess) return c->CudaFail(e, "cudaEventCreate");
        if ((e = cudaMallocHost(&a.poll_count, sizeof(uint32_t) * kPollRing)) != cudaSuccess) return c->CudaFail(e, "cudaMallocHost");
    }
    if ((e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return c->CudaFail(e, "cudaEventCreate");
    if ((e = cudaMallocHost(&c->pinned_count, 2 * sizeof(uint32_t))) != cudaSuccess) return c->CudaFail(e, "cudaMallocHost");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return c->CudaFail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&c->ev_begin)) != cudaSuccess) return c->CudaFail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&c->ev_end)) != cudaSuccess) return c->CudaFail(e, "cudaEventCreate");
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return c->CudaFail(e, "scene upload");
    c->stats.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() - c->stats.bvh_build_ms;
    *out = c.release();
    return B200PT_OK;
}

void b200pt_destroy(b200pt_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    delete h;
}

int b200pt_render_device(b200pt_handle h, const b200pt_render_opts *opts, float *frame_dev, void *stream) {
    if (!h || !frame_dev) return SetGlobalError(B200PT_EINVAL, "b200pt_render_device: null argument");
    ResolvedOpts ro;
    int rc = ResolveOpts(h, opts, &ro);
    if (rc != B200PT_OK) return rc;
    return RenderOnStream(h, ro, frame_dev, nullptr, static_cast<cudaStream_t>(stream));
}

int b200pt_render_progressive_device(b200pt_handle h, const b200pt_render_opts *opts, uint32_t frame_index, float *frame_dev,
                                     float *frame_srgb_dev, void *stream) {
    if (!h || !frame_dev) return SetGlobalError(B200PT_EINVAL, "b200pt_render_progressive_device: null argument");
    ResolvedOpts ro;
    int rc = ResolveOpts(h, opts, &ro);
    if (rc != B200PT_OK) return rc;
    if (ro.tile_world != 1) return h->Fail(B200PT_EINVAL, "progressive rendering works on the whole frame (tile_world must be 1).");
    ro.spp = 1; // main.cpp:53-58: spp is forced to 1 when previewing
    ro.progressive = true;
    ro.frame_index = frame_index;
    ro.frame_srgb = frame_srgb_dev;
    return RenderOnStream(h, ro, frame_dev, nullptr, static_cast<cudaStream_t>(stream));
}

int b200pt_render(b200pt_handle h, const b200pt_render_opts *opts, float *frame_host) {
    if (!h || !frame_host) return SetGlobalError(B200PT_EINVAL, "b200pt_render: null argument");
    ResolvedOpts ro;
    int rc = ResolveOpts(h, opts, &ro);
    if (rc != B200PT_OK) return rc;
    const size_t n = static_cast<size_t>(ro.width) * ro.height * 3;
    CU_CHECK(h, cudaSetDevice(h->device));
    if (h->frame.count < n) CU_CHECK(h, h->frame.Alloc(n));
    if (ro.tile_world > 1) CU_CHECK(h, cudaMemcpyAsync(h->frame.ptr, frame_host, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    rc = RenderOnStream(h, ro, h->frame.ptr, nullptr, h->stream);
    if (rc != B200PT_OK) return rc;
    CU_CHECK(h, cudaMemcpyAsync(frame_host, h->frame.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return B200PT_OK;
}

uint64_t b200pt_tile_buffer_floats(uint32_t width, uint32_t height, uint32_t tile_world) {
    if (tile_world == 0) tile_world = 1;
    return 3ull * PixelsPerRank(width, height, tile_world);
}

int b200pt_render_tiles_device(b200pt_handle h, const b200pt_render_opts *opts, float *tiles_dev, void *stream) {
    if (!h || !tiles_dev) return SetGlobalError(B200PT_EINVAL, "b200pt_render_tiles_device: null argument");
    ResolvedOpts ro;
    int rc = ResolveOpts(h, opts, &ro);
    if (rc != B200PT_OK) return rc;
    return RenderOnStream(h, ro, nullptr, tiles_dev, static_cast<cudaStream_t>(stream));
}

int b200pt_assemble_tiles_device(b200pt_handle h, uint32_t width, uint32_t height, uint32_t tile_world, const float *gathered_dev,
                                 float *frame_dev, void *stream) {
    if (!h || !gathered_dev || !frame_dev || tile_world == 0)
        return SetGlobalError(B200PT_EINVAL, "b200pt_assemble_tiles_device: bad argument");
    CU_CHECK(h, cudaSetDevice(h->device));
    LaunchConfig lc{};
    lc.blocks = h->num_sms * 4, lc.threads = 256, lc.stream = static_cast<cudaStream_t>(stream), lc.stats = false;
    lc.top_nodes = h->top_nodes, lc.refill = h->refill, lc.min_inner = h->min_inner, lc.tri_min = h->tri_min;
    LaunchAssemble(lc, width, height, tile_world, PixelsPerRank(width, height, tile_world), gathered_dev, frame_dev);
    CU_CHECK(h, cudaGetLastError());
    return B200PT_OK;
}

int b200pt_get_stats(b200pt_handle h, b200pt_stats *out) {
    if (!h || !out) return SetGlobalError(B200PT_EINVAL, "b200pt_get_stats: null argument");
    CU_CHECK(h, cudaSetDevice(h->device));
    if (h->timing_pending) {
        CU_CHECK(h, cudaEventSynchronize(h->ev_end));
        float ms = 0.0f;
        CU_CHECK(h, cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end));
        h->stats.render_ms = ms;
        Counters per_arena[kMaxArenas];
        CU_CHECK(h, cudaMemcpy(per_arena, h->counters.ptr, sizeof(per_arena), cudaMemcpyDeviceToHost));
        Counters host_counters = per_arena[0];
        for (int a = 1; a < kMaxArenas; ++a)
            for (int k = 0; k < 3; ++k) {
                host_counters.cls[k].rays += per_arena[a].cls[k].rays;
                host_counters.cls[k].node_visits += per_arena[a].cls[k].node_visits;
                host_counters.cls[k].prim_tests += per_arena[a].cls[k].prim_tests;
            }
        b200pt_kernel_stats *ks[kNumClasses] = {&h->stats.primary, &h->stats.extend, &h->stats.shadow, &h->stats.shade, &h->stats.other, &h->stats.tail};
        for (int k = 0; k < kNumClasses; ++k) {
            *ks[k] = b200pt_kernel_stats{};
            ks[k]->launches = h->class_launches[k];
            if (k < 3) {
                ks[k]->rays = host_counters.cls[k].rays;
                ks[k]->node_visits = host_counters.cls[k].node_visits;
                ks[k]->prim_tests = host_counters.cls[k].prim_tests;
            }
        }
        const bool dump = getenv("B200PT_DUMP_TIMELINE") != nullptr; // every timed launch of the frame on stderr: class, start, duration
        static const char *const kClassNames[kNumClasses] = {"primary", "trace", "shadow", "shade", "other", "tail"};
        for (const b200pt_context::TimedLaunch &t : h->timed) {
            float t_ms = 0.0f;
            CU_CHECK(h, cudaEventElapsedTime(&t_ms, t.begin, t.end));
            ks[t.cls]->ms += t_ms;
            if (dump) {
                float at = 0.0f;
                cudaEventElapsedTime(&at, h->ev_begin, t.begin);
                fprintf(stderr, "[b200pt timeline] %-8s at %8.3f ms  %8.3f ms\n", kClassNames[t.cls], at, t_ms);
            }
        }
        h->timing_pending = false;
    }
    *out = h->stats;
    return B200PT_OK;
}

int b200pt_debug_trace(b200pt_handle h, const b200pt_debug_ray *rays_host, uint64_t n, uint32_t flags, b200pt_debug_hit *hits_host) {
    if (!h || (n && (!rays_host || !hits_host))) return SetGlobalError(B200PT_EINVAL, "b200pt_debug_trace: null argument");
    if (n == 0) return B200PT_OK;
    if (n > (1ull << 30)) return h->Fail(B200PT_EINVAL, "b200pt_debug_trace: too many rays.");
    CU_CHECK(h, cudaSetDevice(h->device));
    DeviceArray<b200pt_debug_ray> rays;
    DeviceArray<b200pt_debug_hit> hits;
    DeviceArray<uint32_t> counter;
    CU_CHECK(h, rays.Alloc(n));
    CU_CHECK(h, hits.Alloc(n));
    CU_CHECK(h, counter.Alloc(1));
    CU_CHECK(h, cudaMemcpyAsync(rays.ptr, rays_host, n * sizeof(b200pt_debug_ray), cudaMemcpyHostToDevice, h->stream));
    CU_CHECK(h, cudaMemsetAsync(counter.ptr, 0, sizeof(uint32_t), h->stream));
    LaunchConfig lc{};
    lc.blocks = h->num_sms * h->ctas_per_sm, lc.threads = 256, lc.stream = h->stream, lc.stats = false;
    lc.top_nodes = 0, lc.refill = h->refill, lc.min_inner = h->min_inner, lc.tri_min = h->tri_min;
    LaunchDebugTrace(lc, h->scene, rays.ptr, static_cast<uint32_t>(n), (flags & B200PT_DEBUG_ANY_HIT) != 0, (flags & B200PT_DEBUG_PER_LANE_LOOP) != 0,
                     (flags & B200PT_DEBUG_RAW_PRIM) != 0, (flags & B200PT_DEBUG_PACKET_LOOP) != 0, hits.ptr, counter.ptr);
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaMemcpyAsync(hits_host, hits.ptr, n * sizeof(b200pt_debug_hit), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return B200PT_OK;
}

int b200pt_debug_eval(b200pt_handle h, uint32_t what, uint32_t id, uint64_t n, const float *in_host, float *out_host) {
    if (!h || (n && (!in_host || !out_host))) return SetGlobalError(B200PT_EINVAL, "b200pt_debug_eval: null argument");
    if (n == 0) return B200PT_OK;
    const DeviceScene &s = h->scene;
    uint64_t limit = 0;
    switch (what) {
    case B200PT_EVAL_BSDF_EVALUATE: case B200PT_EVAL_BSDF_SAMPLE: limit = s.num_bsdfs; break;
    case B200PT_EVAL_EMITTER_SAMPLE: case B200PT_EVAL_EMITTER_DIR: limit = s.integrator.num_emitters; break;
    case B200PT_EVAL_MEDIUM_SAMPLE: case B200PT_EVAL_MEDIUM_EVALUATE: case B200PT_EVAL_PHASE_SAMPLE: case B200PT_EVAL_PHASE_EVALUATE: limit = s.num_media; break;
    case B200PT_EVAL_TEXTURE: limit = s.num_textures; break;
    case B200PT_EVAL_SURFACE: limit = 1; break;
    default: return h->Fail(B200PT_EINVAL, "b200pt_debug_eval: unknown function.");
    }
    if (id >= limit || n > (1ull << 26)) return h->Fail(B200PT_EINVAL, "b200pt_debug_eval: index out of range.");
    CU_CHECK(h, cudaSetDevice(h->device));
    DeviceArray<float> in, out;
    CU_CHECK(h, in.Alloc(n * B200PT_EVAL_IN));
    CU_CHECK(h, out.Alloc(n * B200PT_EVAL_OUT));
    CU_CHECK(h, cudaMemcpyAsync(in.ptr, in_host, n * B200PT_EVAL_IN * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    LaunchDebugEval(h->stream, s, what, id, static_cast<uint32_t>(n), in.ptr, out.ptr);
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaMemcpyAsync(out_host, out.ptr, n * B200PT_EVAL_OUT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return B200PT_OK;
}

int b200pt_debug_light_order(const b200pt_scene_desc *scene, uint32_t light, uint32_t *tri_ids, float *cdf, uint64_t capacity, uint64_t *count) {
    if (!scene || !count) return SetGlobalError(B200PT_EINVAL, "b200pt_debug_light_order: null argument");
    try {
        HostScene hs;
        std::string error;
        if (!BuildHostScene(*scene, 0, false, false, &hs, &error)) return SetGlobalError(B200PT_EINVAL, error);
        if (light >= hs.map_area_light_instance.size()) return SetGlobalError(B200PT_EINVAL, "b200pt_debug_light_order: no such area light.");
        const DInstance &inst = hs.instances[hs.map_area_light_instance[light]];
        *count = inst.analytic == kInvalid ? inst.light_tri_count : 0;
        // TriVerts::v1.w carries the triangle's number in the scene description (instances in order); the smallest one of the
        // instance is its first triangle
        uint32_t first = 0xffffffffu;
        std::vector<uint32_t> ids(*count);
        for (uint64_t k = 0; k < *count; ++k) {
            memcpy(&ids[k], &hs.tri_verts[hs.light_tri_ids[inst.light_tri_begin + k]].v1.w, 4);
            first = std::min(first, ids[k]);
        }
        for (uint64_t k = 0; k < *count && k < capacity; ++k) {
            if (tri_ids) tri_ids[k] = ids[k] - first;
            if (cdf) cdf[k] = hs.light_tri_cdf[inst.light_tri_begin + k];
        }
        return B200PT_OK;
    } catch (const std::exception &e) {
        return SetGlobalError(B200PT_ENOMEM, std::string("b200pt_debug_light_order: ") + e.what());
    }
}

int b200pt_debug_render_replay(b200pt_handle h, uint32_t width, uint32_t height, uint32_t spp, float *frame_host) {
    if (!h || !frame_host) return SetGlobalError(B200PT_EINVAL, "b200pt_debug_render_replay: null argument");
    if (h->scene.integrator.has_opacity)
        return h->Fail(B200PT_EINVAL, "b200pt_debug_render_replay: alpha-tested scenes draw their opacity numbers in the reference's own BVH order.");
    if (width == 0) width = static_cast<uint32_t>(h->host.camera.width);
    if (height == 0) height = static_cast<uint32_t>(h->host.camera.height);
    if (spp == 0) spp = h->host.camera.spp;
    if (width == 0 || height == 0 || spp == 0 || static_cast<uint64_t>(width) * height > (1ull << 24))
        return h->Fail(B200PT_EINVAL, "b200pt_debug_render_replay: bad frame size.");
    CU_CHECK(h, cudaSetDevice(h->device));
    BatchParams bp{};
    bp.camera = MakeCamera(h->host.camera, width, height);
    bp.width = width, bp.height = height, bp.spp = spp;
    bp.spp_inv = 1.0f / spp;
    bp.sample_count = 1;
    DeviceArray<float> frame;
    CU_CHECK(h, frame.Alloc(3ull * width * height));
    uint32_t trace_pixel = 0xffffffffu; // B200PT_REPLAY_TRACE="i,j": print the path vertices of that pixel (debug_eval.cu)
    if (const char *e = getenv("B200PT_REPLAY_TRACE")) {
        int ti = -1, tj = -1;
        if (sscanf(e, "%d,%d", &ti, &tj) == 2 && ti >= 0 && tj >= 0 && static_cast<uint32_t>(ti) < width && static_cast<uint32_t>(tj) < height)
            trace_pixel = static_cast<uint32_t>(tj) * width + static_cast<uint32_t>(ti);
    }
    LaunchDebugReplay(h->stream, h->scene, bp, frame.ptr, trace_pixel);
    CU_CHECK(h, cudaGetLastError());
    CU_CHECK(h, cudaMemcpyAsync(frame_host, frame.ptr, 3ull * width * height * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    return B200PT_OK;
}

const char *b200pt_last_error(b200pt_handle h) { return h ? h->error.c_str() : GlobalError(); }

int b200pt_get_kulla_conty(b200pt_handle h, float *brdf_avg, float *albedo_avg) {
    if (!brdf_avg || !albedo_avg) return SetGlobalError(B200PT_EINVAL, "b200pt_get_kulla_conty: null argument");
    if (!h) { // the tables depend on no scene: without a handle they are computed right here, on the host (no GPU needed)
        try {
            std::vector<float> brdf(kLutResolution * kLutResolution), albedo(kLutResolution);
            ComputeKullaContyTables(brdf.data(), albedo.data());
            memcpy(brdf_avg, brdf.data(), sizeof(float) * brdf.size());
            memcpy(albedo_avg, albedo.data(), sizeof(float) * albedo.size());
            return B200PT_OK;
        } catch (const std::exception &e) {
            return SetGlobalError(B200PT_ENOMEM, std::string("b200pt_get_kulla_conty: ") + e.what());
        }
    }
    // The tables are only built at create time when a conductor/dielectric BSDF reads them.
    bool built = false;
    for (float v : h->host.kc_albedo_avg) built = built || v != 0.0f;
    if (!built) ComputeKullaContyTables(h->host.kc_brdf_avg.data(), h->host.kc_albedo_avg.data());
    memcpy(brdf_avg, h->host.kc_brdf_avg.data(), sizeof(float) * kLutResolution * kLutResolution);
    memcpy(albedo_avg, h->host.kc_albedo_avg.data(), sizeof(float) * kLutResolution);
    return B200PT_OK;
}

int b200pt_get_envmap_tables(b200pt_handle h, float *out, uint64_t capacity_floats, uint64_t *num_floats, float *normalization) {
    if (!h || !num_floats) return SetGlobalError(B200PT_EINVAL, "b200pt_get_envmap_tables: null argument");
    *num_floats = h->host.envmap_tables.size();
    if (normalization) *normalization = h->host.envmap_normalization;
    if (out && capacity_floats >= h->host.envmap_tables.size() && !h->host.envmap_tables.empty())
        memcpy(out, h->host.envmap_tables.data(), h->host.envmap_tables.size() * sizeof(float));
    return B200PT_OK;
}

} // extern "C"
