// wavefront.cu — the sm_100a kernels of the wavefront path tracer.
//
// Replaces the per-pixel / per-sample loop DrawPixel (src/renderer/renderer.cpp:62-85) and the
// integrators ShadePath (src/renderer/integrators/path.cpp:8-236) and ShadeVolPath
// (src/renderer/integrators/volpath.cpp:8-485).  Instead of one thread walking one pixel's
// paths to completion (the reference's DispathRaysCuda megakernel, renderer.cpp:88-95), paths
// advance in lock step, one bounce per round, through queues of SoA ray records; survivors of
// Russian roulette / throughput cut-off are compacted with warp ballots so every traversal
// launch works on a dense queue.
#include <cuda_runtime.h>

#include <algorithm>

#include "shade_kernel.cuh"
#include "shading.cuh"
#include "traverse_packet.cuh"
#include "traverse_wide.cuh"
#include "wavefront.cuh"

namespace b200pt {

namespace {

constexpr int kThreads = 256;
#ifndef B200PT_TRACE_MIN_CTAS
#define B200PT_TRACE_MIN_CTAS 4   // resident CTAs per SM the traversal kernels are compiled for (register cap = 65536 / 256 / this)
#endif

__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Stage the top of the BVH (the first `bytes` of the node array, stored breadth-first) into shared memory with one
// TMA bulk copy (cp.async.bulk, completion signalled on an mbarrier).  Every ray walks these nodes, so they are
// served at shared-memory latency instead of L1/L2.
__device__ __forceinline__ void StageTop(const void *nodes, uint32_t bytes, void *top, uint64_t *bar) {
    if (bytes == 0) return;
    const uint32_t bar_addr = SmemAddr(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         SmemAddr(top)),
                     "l"(nodes), "r"(bytes), "r"(bar_addr)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar_addr)
                     : "memory");
    }
}

// Which tree the traversal kernels walk: the binary layout (cross-check / GPU-built trees), the compressed 8-wide layout
// (default), or the wide layout with its top nodes staged in shared memory.
enum TraversalLayout { kLayoutBinary = 0, kLayoutWide = 1, kLayoutWideTop = 2 };

// The persistent traversal loop of the layout the kernel was instantiated for.
template <bool MIXED, bool STATS, bool OPACITY, int LAYOUT, typename Fetch, typename Finish>
__device__ __forceinline__ void TraverseLayout(const DeviceScene &scene, const uint4 *top, uint32_t num_top, uint32_t num_rays, uint32_t *work_counter,
                                               int refill, int phase_lanes, uint2 key, Fetch fetch, Finish finish, TraversalCounters *tc,
                                               uint32_t *rays) {
    if constexpr (LAYOUT == kLayoutBinary)
        TraversePersistent<MIXED, STATS, OPACITY, false>(scene, nullptr, 0, num_rays, work_counter, refill, phase_lanes, key, fetch, finish, tc, rays);
    else
        TraversePersistentWide<MIXED, STATS, OPACITY, LAYOUT == kLayoutWideTop>(scene, top, num_top, num_rays, work_counter, refill, phase_lanes, key,
                                                                                 fetch, finish, tc, rays);
}

// Shading bin of a closest-hit record (bin = BSDF model of the surface; 0 = escaped / BSDF-less, 1 = area light).
__device__ __forceinline__ uint32_t ShadeBinOfHit(const DeviceScene &scene, uint32_t prim) {
    if (prim == kPrimMiss) return 0u;
    if (prim & kPrimAnalyticBit) {
        const uint32_t id_bsdf = scene.instances[scene.analytic[prim & ~kPrimAnalyticBit].inst].id_bsdf;
        return id_bsdf == kInvalid ? 0u : scene.bsdfs[id_bsdf].type;
    }
    return __ldg(scene.tri_bsdf_type + (prim & kPrimIndexMask));
}

// k_bin_hits: files every entry of a freshly traced path queue under the shading bin of its closest hit (one coalesced
// pass over the hit records).  Doing this inside the traversal kernels costs one atomic per RAY, because rays finish one
// lane at a time there (volumetric-caustic: +50 % traversal time).
// A CTA takes kBinChunk entries at a time: every thread loads its kBinPerThread hit records up front (independent loads in
// flight), ranks them per bin with warp votes and SHARED-memory counters, and the CTA reserves its slice of each bin list with
// ONE global atomic per bin and chunk.  (One global atomic per warp and bin — the first version — ran at the rate the L2
// serialises adds to the same three or four addresses: 41 ms per 256-spp volumetric-caustic frame at 6 % issue utilisation,
// profiles/r02_counters_volumetric-caustic_1024x1024x256.json.)
constexpr int kBinPerThread = 8;
constexpr uint32_t kBinChunk = kThreads * kBinPerThread;
__global__ void __launch_bounds__(kThreads) k_bin_hits(const __grid_constant__ DeviceScene scene, const HitRec *hits, int which,
                                                       ShadeBins bins, Counters *counters) {
    __shared__ uint32_t s_count[kNumShadeBins], s_base[kNumShadeBins];
    const uint32_t n = counters->queue[which];
    const uint32_t in_use = scene.integrator.shade_bins;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t chunk = blockIdx.x * kBinChunk; chunk < n; chunk += gridDim.x * kBinChunk) {
        if (threadIdx.x < kNumShadeBins) s_count[threadIdx.x] = 0;
        __syncthreads();
        uint32_t prim[kBinPerThread];
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const uint32_t i = chunk + k * kThreads + threadIdx.x;
            prim[k] = i < n ? hits[i].prim : kPrimMiss;
        }
        uint32_t packed_bin = 0, offset[kBinPerThread]; // 4 bits per entry
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const uint32_t i = chunk + k * kThreads + threadIdx.x;
            uint32_t bin = kNumShadeBins; // none
            // an escaped ray with nothing left to do (no environment map, no medium) is dropped here
            if (i < n && (prim[k] != kPrimMiss || (in_use & 1u))) bin = ShadeBinOfHit(scene, prim[k]);
            packed_bin |= bin << (4 * k);
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            offset[k] = 0;
            if (bin != kNumShadeBins) {
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (static_cast<int>(lane) == leader) base = atomicAdd(&s_count[bin], __popc(peers));
                offset[k] = __shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane) - 1u));
            }
        }
        __syncthreads();
        if (threadIdx.x < kNumShadeBins && s_count[threadIdx.x] != 0)
            s_base[threadIdx.x] = atomicAdd(&counters->bin_count[which][threadIdx.x], s_count[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const uint32_t bin = (packed_bin >> (4 * k)) & 15u;
            if (bin != kNumShadeBins)
                bins.lists[static_cast<uint64_t>(BinListIndex(in_use, bin)) * bins.capacity + s_base[bin] + offset[k]] = chunk + k * kThreads + threadIdx.x;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void FlushCounters(bool stats, const TraversalCounters &tc, uint32_t rays, int cls, Counters *c) {
    if (!stats) return;
    // one atomic per warp
    uint32_t nodes = tc.nodes, prims = tc.prims, r = rays;
    for (int o = 16; o > 0; o >>= 1) {
        nodes += __shfl_down_sync(0xffffffffu, nodes, o);
        prims += __shfl_down_sync(0xffffffffu, prims, o);
        r += __shfl_down_sync(0xffffffffu, r, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&c->cls[cls].node_visits, static_cast<unsigned long long>(nodes));
        atomicAdd(&c->cls[cls].prim_tests, static_cast<unsigned long long>(prims));
        atomicAdd(&c->cls[cls].rays, static_cast<unsigned long long>(r));
    }
}

// ---------------------------------------------------------------------------------------------
// k_primary: camera-ray generation (renderer.cpp:62-75) fused with the first closest hit.
// ---------------------------------------------------------------------------------------------
template <bool STATS, bool OPACITY, int LAYOUT>
__global__ void __launch_bounds__(kThreads, B200PT_TRACE_MIN_CTAS) k_primary(const __grid_constant__ DeviceScene scene,
                                                      const __grid_constant__ BatchParams bp, PathQueue q,
                                                      float *radiance, uint32_t capacity, Counters *counters, int max_top,
                                                      int refill, int phase_lanes, int packets) {
    extern __shared__ uint4 top[];
    __shared__ uint64_t bar;
    const uint32_t num_top = LAYOUT == kLayoutWideTop ? min(scene.num_wide_nodes, static_cast<uint32_t>(max_top)) : 0u;
    if (LAYOUT == kLayoutWideTop) StageTop(scene.wide_nodes, num_top * static_cast<uint32_t>(sizeof(WideNode)), top, &bar);
    const uint32_t nslots = bp.pixel_count * bp.sample_count;
    TraversalCounters tc[2];
    uint32_t rays[2] = {0, 0};
    Ray cam; // the camera ray of the slot this lane currently traces (the traversal shortens its own copy)
    auto fetch = [&](uint32_t slot, Ray *ray, uint3 *ctr, bool *) {
        uint32_t px, py;
        if (!LocalPixelToImage(bp, JobPixelToLocal(bp, bp.pixel_begin + slot / bp.sample_count), &px, &py)) return false;
        const uint32_t s = bp.sample_begin + slot % bp.sample_count;
        if (OPACITY) *ctr = make_uint3(py * bp.width + px, s, 0u);
        const float u = bp.progressive ? bp.progressive_u : s * bp.spp_inv, v = bp.progressive ? bp.progressive_v : VanDerCorput2(s + 1);
        const float x = 2.0f * (px + u) / static_cast<int>(bp.width) - 1.0f, y = 1.0f - 2.0f * (py + v) / static_cast<int>(bp.height);
        ray->o = mk3(bp.camera.eye);
        ray->d = Normalize(mk3(bp.camera.front) + x * mk3(bp.camera.view_dx) + y * mk3(bp.camera.view_dy));
        ray->tmin = kEpsilonDistance;
        ray->tmax = kMaxFloat;
        cam = *ray;
        return true;
    };
    auto finish = [&](uint32_t slot, const HitRec &hit, bool found, bool) {
        if (!found) { // path.cpp:24-35: an escaped camera ray sees the environment and the sun disc
            V3 L = mk3(0.0f);
            if (scene.integrator.id_envmap != kInvalid) L += EmitterEvaluateDir(scene, scene.emitters[scene.integrator.id_envmap], cam.d);
            if (scene.integrator.id_sun != kInvalid) L += EmitterEvaluateDir(scene, scene.emitters[scene.integrator.id_sun], cam.d);
            if (L.x != 0.0f || L.y != 0.0f || L.z != 0.0f) {
                RadianceSet(radiance, slot, L.x, L.y, L.z);
            }
            return;
        }
        const uint32_t idx = AppendCoalesced(&counters->queue[0]);
        q.ox[idx] = cam.o.x, q.oy[idx] = cam.o.y, q.oz[idx] = cam.o.z;
        q.dx[idx] = cam.d.x, q.dy[idx] = cam.d.y, q.dz[idx] = cam.d.z;
        q.tr[idx] = 1.0f, q.tg[idx] = 1.0f, q.tb[idx] = 1.0f;
        q.pdf[idx] = 0.0f;
        q.slot[idx] = slot;
        if (q.medium != nullptr) {
            q.medium[idx] = kInvalid;
            q.wx[idx] = -cam.d.x, q.wy[idx] = -cam.d.y, q.wz[idx] = -cam.d.z;
        }
        q.hit[idx] = hit;
    };
    if constexpr (LAYOUT == kLayoutBinary) {
        // 32 consecutive slots are 32 samples of one pixel (slot = pixel * sample_count + sample): one warp packet
        if (packets & kPacketsPrimary) {
            __shared__ int packet_stack[kThreads / 32][kPacketStack];
            TraversePacket<false, STATS, OPACITY>(
                scene, 0u, nslots, &counters->work_primary, bp.key, packet_stack[threadIdx.x >> 5],
                [&](uint32_t slot, Ray *ray, uint3 *ctr) { return fetch(slot, ray, ctr, nullptr); },
                [&](uint32_t slot, const HitRec &hit, bool found) { finish(slot, hit, found, false); }, &tc[0], &rays[0]);
            FlushCounters(STATS, tc[0], rays[0], kClassPrimary, counters);
            return;
        }
    }
    TraverseLayout<false, STATS, OPACITY, LAYOUT>(scene, top, num_top, nslots, &counters->work_primary, refill, phase_lanes, bp.key, fetch, finish, tc, rays);
    FlushCounters(STATS, tc[0], rays[0], kClassPrimary, counters);
}

// ---------------------------------------------------------------------------------------------
// k_trace: ONE persistent launch per bounce for both ray kinds that leave a path vertex — the closest hit of the
// compacted survivor queue (TLAS::Intersect, tlas.cpp:13-43) and the occlusion test of the NEE rays
// (TLAS::IntersectAny, tlas.cpp:44-76).  The two kinds are independent, so tracing them from one work counter
// halves the number of launch tails (a launch cannot end before its slowest ray: ~0.2 ms on Dragon, which dominated
// the deep bounces when they were two launches: profiles/r01_dram_traffic.json per-launch times).
// Ray index space: [0, n_extend) closest-hit rays of queue `which`, then [n_extend, n_extend + n_shadow) NEE rays.
// ---------------------------------------------------------------------------------------------

template <bool STATS, bool OPACITY, int LAYOUT>
__global__ void __launch_bounds__(kThreads, B200PT_TRACE_MIN_CTAS) k_trace(const __grid_constant__ DeviceScene scene,
                                                    const __grid_constant__ BatchParams bp, uint32_t depth, PathQueue q, int which,
                                                    ShadowQueue sq, float *radiance, uint32_t capacity, Counters *counters,
                                                    int max_top, int refill, int phase_lanes, int packets) {
    extern __shared__ uint4 top[];
    __shared__ uint64_t bar;
    const uint32_t n_extend = which >= 0 ? counters->queue[which] : 0u, n_shadow = counters->shadow;
    const uint32_t n = n_extend + n_shadow;
    if (blockIdx.x * blockDim.x >= n) return;
    const uint32_t num_top = LAYOUT == kLayoutWideTop ? min(scene.num_wide_nodes, static_cast<uint32_t>(max_top)) : 0u;
    if (LAYOUT == kLayoutWideTop) StageTop(scene.wide_nodes, num_top * static_cast<uint32_t>(sizeof(WideNode)), top, &bar);
    TraversalCounters tc[2];
    uint32_t rays[2] = {0, 0};
    auto fetch = [&](uint32_t i, Ray *ray, uint3 *ctr, bool *any) {
        ray->tmin = kEpsilonDistance;
        if (i < n_extend) {
            ray->o = mk3(q.ox[i], q.oy[i], q.oz[i]);
            ray->d = mk3(q.dx[i], q.dy[i], q.dz[i]);
            ray->tmax = kMaxFloat;
            *any = false;
            if (OPACITY) *ctr = SlotCounter(bp, q.slot[i], depth);
        } else {
            const uint32_t j = i - n_extend;
            ray->o = mk3(sq.ox[j], sq.oy[j], sq.oz[j]);
            ray->d = mk3(sq.dx[j], sq.dy[j], sq.dz[j]);
            ray->tmax = sq.tmax[j];
            *any = true;
            // several NEE rays of one vertex (one per emitter) share (pixel, sample, depth): the bits of the ray itself
            // separate them (the queue position would not be reproducible from run to run)
            if (OPACITY) *ctr = SlotCounter(bp, sq.slot[j], depth), ctr->y ^= (__float_as_uint(ray->tmax) ^ __float_as_uint(ray->d.x) * 0x85ebca6bu) * 0x9e3779b9u;
        }
        return true;
    };
    auto finish = [&](uint32_t i, const HitRec &hit, bool found, bool any) {
        if (!any) {
            q.hit[i] = hit;
        } else if (!found) {
            // Unoccluded: the light sample counts.  Rays end one lane at a time here, so reading the contribution and adding it
            // with three atomics cost seven scattered memory instructions at ~4 active lanes (7.7 % of this kernel's stall
            // samples, profiles/r01_k_trace_source.csv.gz).  The ray is only MARKED (one store); k_settle, the coalesced pass
            // that opens the next bounce, adds the marked contributions.
            sq.tmax[i - n_extend] = kShadowUnoccluded;
        }
    };
    if constexpr (LAYOUT == kLayoutBinary) {
        // Coherent NEE rays (the host says when: first vertex, delta lights) go first, as warp packets; the persistent loop then
        // only has the bounce rays left.  One launch all the same: a warp that finds no packet left moves on to the bounce rays.
        if ((packets & kPacketsShadow) && n_shadow > 0) {
            __shared__ int packet_stack[kThreads / 32][kPacketStack];
            TraversePacket<true, STATS, OPACITY>(
                scene, n_extend, n_shadow, &counters->work_packets, bp.key, packet_stack[threadIdx.x >> 5],
                [&](uint32_t i, Ray *ray, uint3 *ctr) {
                    bool any;
                    return fetch(i, ray, ctr, &any);
                },
                [&](uint32_t i, const HitRec &hit, bool found) { finish(i, hit, found, true); }, &tc[1], &rays[1]);
            if (n_extend > 0)
                TraversePersistent<false, STATS, OPACITY, false>(scene, nullptr, 0, n_extend, &counters->work_trace, refill, phase_lanes, bp.key, fetch, finish, tc, rays);
            FlushCounters(STATS, tc[0], rays[0], kClassExtend, counters);
            FlushCounters(STATS, tc[1], rays[1], kClassShadow, counters);
            return;
        }
    }
    TraverseLayout<true, STATS, OPACITY, LAYOUT>(scene, top, num_top, n, &counters->work_trace, refill, phase_lanes, bp.key, fetch, finish, tc, rays);
    FlushCounters(STATS, tc[0], rays[0], kClassExtend, counters);
    FlushCounters(STATS, tc[1], rays[1], kClassShadow, counters);
}

// ---------------------------------------------------------------------------------------------
// k_debug_trace: caller-supplied rays through the SAME traversal loops as k_primary / k_trace (b200pt_debug_trace, the test
// hook behind the pointwise parity tests of the Woop / box / analytic tests against the reference's TLAS::Intersect).
// `single`: the per-lane loop of the tail kernel instead of the persistent one.
// ---------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(kThreads, B200PT_TRACE_MIN_CTAS) k_debug_trace(const __grid_constant__ DeviceScene scene, const b200pt_debug_ray *rays_in,
                                                                                uint32_t n, bool any_hit, bool single, bool raw_prim, bool packet,
                                                                                b200pt_debug_hit *out, uint32_t *work_counter, int refill, int phase_lanes) {
    auto report = [&](uint32_t i, const HitRec &hit, bool found, bool any) {
        b200pt_debug_hit h;
        h.t = hit.t, h.u = hit.u, h.v = hit.v, h.prim = hit.prim;
        if (any) {
            h.prim = found ? 0u : kPrimMiss;
        } else if (!raw_prim && hit.prim != kPrimMiss && !(hit.prim & kPrimAnalyticBit)) { // leaf order -> index in the scene description
            const uint32_t original = __float_as_uint(__ldg(&scene.tri_verts[hit.prim & kPrimIndexMask].v1.w));
            h.prim = original | (hit.prim & kPrimInsideBit);
        }
        out[i] = h;
    };
    auto load = [&](uint32_t i, Ray *ray) {
        const b200pt_debug_ray r = rays_in[i];
        ray->o = mk3(r.o[0], r.o[1], r.o[2]), ray->d = mk3(r.d[0], r.d[1], r.d[2]);
        ray->tmin = r.tmin, ray->tmax = r.tmax;
    };
    if (single) {
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            Ray ray;
            load(i, &ray);
            HitRec hit;
            hit.t = ray.tmax, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
            const Rng rng(0, 0, 0, make_uint2(0u, 0u), kRngDomainClosest);
            const bool found = LAYOUT == kLayoutBinary ? TraverseSingle(scene, ray, any_hit, false, rng, &hit, false, nullptr)
                                                       : TraverseSingleWide(scene, ray, any_hit, false, rng, &hit, false, nullptr);
            report(i, hit, found, any_hit);
        }
        return;
    }
    extern __shared__ uint4 top[];
    TraversalCounters tc[2];
    uint32_t traced[2] = {0, 0};
    if constexpr (LAYOUT == kLayoutBinary) {
        if (packet) { // 32 consecutive caller rays per warp packet (traverse_packet.cuh)
            __shared__ int packet_stack[kThreads / 32][kPacketStack];
            auto fetch3 = [&](uint32_t i, Ray *ray, uint3 *) {
                load(i, ray);
                return true;
            };
            if (any_hit)
                TraversePacket<true, false, false>(scene, 0u, n, work_counter, make_uint2(0u, 0u), packet_stack[threadIdx.x >> 5], fetch3,
                                                   [&](uint32_t i, const HitRec &hit, bool found) { report(i, hit, found, true); }, &tc[0], &traced[0]);
            else
                TraversePacket<false, false, false>(scene, 0u, n, work_counter, make_uint2(0u, 0u), packet_stack[threadIdx.x >> 5], fetch3,
                                                    [&](uint32_t i, const HitRec &hit, bool found) { report(i, hit, found, false); }, &tc[0], &traced[0]);
            return;
        }
    }
    auto fetch = [&](uint32_t i, Ray *ray, uint3 *, bool *any) {
        load(i, ray);
        *any = any_hit;
        return true;
    };
    TraverseLayout<true, false, false, LAYOUT>(scene, top, 0u, n, work_counter, refill, phase_lanes, make_uint2(0u, 0u), fetch, report, tc, traced);
}

// ---------------------------------------------------------------------------------------------
// Small kernels
// ---------------------------------------------------------------------------------------------
// k_settle: between two bounces.  Adds the contributions of the NEE rays the last k_trace found unoccluded (marked in place,
// see k_trace) to their samples — a coalesced pass over the shadow queue — and then, in the last CTA to finish, resets the
// queue lengths and work counters for the next bounce (what a one-thread launch did before).
// UNIQUE: the scene sends one NEE ray per vertex, so no two entries of the queue share a sample slot: plain read-modify-write
// (atomics serialise in the L2: volumetric-caustic spent 51 ms per frame on them).  Otherwise the NEE rays of one vertex (one
// per emitter + one for the area lights) meet in one slot and are added atomically.
template <bool UNIQUE>
__global__ void __launch_bounds__(kThreads) k_settle(Counters *c, int which_queue, bool reset_shadow, ShadowQueue sq, float *radiance, uint32_t capacity) {
    const uint32_t n = c->shadow;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        if (sq.tmax[j] != kShadowUnoccluded) continue;
        const uint32_t slot = sq.slot[j];
        if (UNIQUE)
            RadianceAdd(radiance, slot, sq.cr[j], sq.cg[j], sq.cb[j]);
        else
            RadianceAtomicAdd(radiance, slot, sq.cr[j], sq.cg[j], sq.cb[j]);
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&c->settle_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        if (which_queue >= 0) {
            c->queue[which_queue] = 0;
            for (int b = 0; b < kNumShadeBins; ++b) c->bin_count[which_queue][b] = 0;
        }
        if (reset_shadow) c->shadow = 0;
        c->work_trace = 0;
        c->work_packets = 0;
        c->settle_ticket = 0;
    }
}

// Visibility pre-pass for camera rays.  Every camera ray of a screen rectangle starts at the eye and runs inside the pyramid
// spanned by the rectangle's four corner directions; a box that lies entirely outside one of the pyramid's side planes, or
// behind the eye, cannot be hit by any of them.  A rectangle that sees none of the scene's cull boxes (which cover all
// geometry) produces only escaped rays, i.e. exactly zero radiance when there is no environment / sun emitter, so its samples
// need no ray at all.  Conservative by construction: a pixel is dropped only on a proof of emptiness.
// Two levels: an 8x8 tile against the coarse boxes (<= 384: most of the screen ends here), then the fine boxes (<= 4096)
// that meet the tile's pyramid are short-listed in shared memory and every PIXEL of the tile is tested against the short
// list.  One warp per local tile; the result is a 64-bit pixel mask per tile.
struct Pyramid {
    V3 plane[5]; // inward normals of the four sides through the eye, and the view direction
};

__device__ __forceinline__ Pyramid MakePyramid(const BatchParams &bp, float i0, float i1, float j0, float j1) {
    const V3 front = mk3(bp.camera.front), dx = mk3(bp.camera.view_dx), dy = mk3(bp.camera.view_dy);
    const float x0 = 2.0f * i0 / static_cast<float>(bp.width) - 1.0f, x1 = 2.0f * i1 / static_cast<float>(bp.width) - 1.0f;
    const float y0 = 1.0f - 2.0f * j0 / static_cast<float>(bp.height), y1 = 1.0f - 2.0f * j1 / static_cast<float>(bp.height);
    const V3 corner[4] = {front + x0 * dx + y0 * dy, front + x1 * dx + y0 * dy, front + x1 * dx + y1 * dy, front + x0 * dx + y1 * dy};
    const V3 centre = corner[0] + corner[2];
    Pyramid py;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        V3 n = Cross(corner[k], corner[(k + 1) & 3]);
        if (Dot(n, centre) < 0.0f) n = -n; // inward
        py.plane[k] = n;
    }
    py.plane[4] = front;
    return py;
}

// Is the box (6 floats) entirely on the outer side of one of the pyramid's planes?
__device__ __forceinline__ bool BoxOutsidePyramid(const Pyramid &py, const V3 &eye, const float *bx) {
    const V3 lo = mk3(bx[0], bx[1], bx[2]), hi = mk3(bx[3], bx[4], bx[5]);
    const V3 c = 0.5f * (lo + hi) - eye, h = 0.5f * (hi - lo);
    bool outside = false;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const V3 n = py.plane[k];
        const float s = Dot(n, c), r = fabsf(n.x) * h.x + fabsf(n.y) * h.y + fabsf(n.z) * h.z;
        if (s + r < -1e-5f * (fabsf(s) + r)) outside = true; // the whole box is on the outer side
    }
    return outside;
}

constexpr uint32_t kCullShortList = 384; // fine boxes a tile may short-list before it gives up and keeps all its pixels

__global__ void __launch_bounds__(kThreads) k_cull_tiles(const __grid_constant__ BatchParams bp, const float *__restrict__ boxes,
                                                         uint32_t num_boxes, const float *__restrict__ fine_boxes, uint32_t num_fine,
                                                         const uint32_t *__restrict__ fine_begin, uint32_t num_local_tiles,
                                                         unsigned long long *masks) {
    __shared__ uint16_t short_all[kThreads / 32][kCullShortList];
    __shared__ uint32_t short_count[kThreads / 32];
    uint16_t *short_list = short_all[threadIdx.x >> 5];
    uint32_t *count_ptr = &short_count[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, num_warps = (gridDim.x * blockDim.x) >> 5;
    const V3 eye = mk3(bp.camera.eye);
    for (uint32_t t = warp; t < num_local_tiles; t += num_warps) {
        const uint32_t tile = t * bp.tile_world + bp.tile_rank;
        unsigned long long mask = 0ull;
        if (tile < bp.num_tiles) {
            const uint32_t ti = (tile % bp.tiles_x) * kTileSize, tj = (tile / bp.tiles_x) * kTileSize;
            const float i0 = static_cast<float>(ti), j0 = static_cast<float>(tj);
            const float i1 = fminf(i0 + kTileSize, static_cast<float>(bp.width)), j1 = fminf(j0 + kTileSize, static_cast<float>(bp.height));
            constexpr float kMargin = 0.5f; // pixels
            const Pyramid tile_pyramid = MakePyramid(bp, i0 - kMargin, i1 + kMargin, j0 - kMargin, j1 + kMargin);
            if (lane == 0) *count_ptr = 0u;
            __syncwarp();
            // coarse boxes, lanes over boxes; a lane whose coarse box meets the tile looks at the fine boxes under it
            bool visible = false;
            for (uint32_t b = lane; b < num_boxes; b += 32) {
                if (BoxOutsidePyramid(tile_pyramid, eye, boxes + 6 * b)) continue;
                visible = true;
                if (num_fine == 0) continue;
                for (uint32_t f = fine_begin[b]; f < fine_begin[b + 1]; ++f) {
                    if (BoxOutsidePyramid(tile_pyramid, eye, fine_boxes + 6 * f)) continue;
                    const uint32_t at = atomicAdd(count_ptr, 1u);
                    if (at < kCullShortList) short_list[at] = static_cast<uint16_t>(f);
                }
            }
            visible = __any_sync(0xffffffffu, visible);
            __syncwarp();
            const uint32_t count = *count_ptr;
            // the pixels of the tile that lie inside the image (lane k and k + 32 -> pixel (k % 8, k / 8))
            const unsigned in_lo = __ballot_sync(0xffffffffu, ti + (lane % kTileSize) < bp.width && tj + (lane / kTileSize) < bp.height);
            const unsigned in_hi = __ballot_sync(0xffffffffu, ti + (lane % kTileSize) < bp.width && tj + 4u + (lane / kTileSize) < bp.height);
            const unsigned long long inside = (static_cast<unsigned long long>(in_hi) << 32) | in_lo;
            if (visible && (num_fine == 0 || count > kCullShortList)) mask = inside; // pixel level off, or too many candidates
            else if (visible) {
                constexpr float kPixelMargin = 0.02f; // the samples of pixel i lie in [i, i + 1)
                unsigned lo_bits = 0u, hi_bits = 0u;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t k = lane + 32u * half, pi = ti + (k % kTileSize), pj = tj + (k / kTileSize);
                    bool sees = false;
                    if (pi < bp.width && pj < bp.height) {
                        const Pyramid pixel = MakePyramid(bp, pi - kPixelMargin, pi + 1.0f + kPixelMargin, pj - kPixelMargin, pj + 1.0f + kPixelMargin);
                        for (uint32_t q = 0; q < count && !sees; ++q) sees = !BoxOutsidePyramid(pixel, eye, fine_boxes + 6u * short_list[q]);
                    }
                    const unsigned ballot = __ballot_sync(0xffffffffu, sees);
                    if (half == 0) lo_bits = ballot;
                    else hi_bits = ballot;
                }
                mask = (static_cast<unsigned long long>(hi_bits) << 32) | lo_bits;
            }
            __syncwarp();
        }
        if (lane == 0) masks[t] = mask;
    }
}

// Exclusive prefix sums of the per-tile pixel counts by ONE CTA (at most a few 10^5 tiles): offsets[t] = position of tile t's first
// surviving pixel in the job's pixel list, counts[0] = pixels, counts[1] = tiles with at least one.
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned long long *masks, uint32_t n, uint32_t *offsets, uint32_t *counts) {
    __shared__ uint32_t warp_sums[32], warp_tiles[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t chunk = (n + blockDim.x - 1) / blockDim.x, begin = min(n, tid * chunk), end = min(n, begin + chunk);
    uint32_t mine = 0, tiles = 0;
    for (uint32_t i = begin; i < end; ++i) {
        mine += __popcll(masks[i]);
        tiles += masks[i] != 0ull;
    }
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= static_cast<uint32_t>(o)) incl += v;
    }
    for (int o = 16; o > 0; o >>= 1) tiles += __shfl_down_sync(0xffffffffu, tiles, o);
    if (lane == 31) warp_sums[warp] = incl;
    if (lane == 0) warp_tiles[warp] = tiles;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], wi = w, wt = warp_tiles[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= static_cast<uint32_t>(o)) wi += v;
        }
        for (int o = 16; o > 0; o >>= 1) wt += __shfl_down_sync(0xffffffffu, wt, o);
        warp_sums[lane] = wi - w; // exclusive
        if (lane == 31) counts[0] = wi;
        if (lane == 0) counts[1] = wt;
    }
    __syncthreads();
    uint32_t out = warp_sums[warp] + incl - mine;
    for (uint32_t i = begin; i < end; ++i) {
        offsets[i] = out;
        out += __popcll(masks[i]);
    }
}

// The job's pixel list: ascending local indices (tile * 64 + pixel in tile) of the set mask bits.  One warp per tile.
__global__ void __launch_bounds__(kThreads) k_list_pixels(const unsigned long long *masks, const uint32_t *offsets, uint32_t n, uint32_t *list) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, num_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = warp; t < n; t += num_warps) {
        const unsigned long long m = masks[t];
        if (m == 0ull) continue;
        const uint32_t base = offsets[t];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t k = lane + 32u * half;
            if ((m >> k) & 1ull) list[base + __popcll(m & ((1ull << k) - 1ull))] = t * kTilePixels + k;
        }
    }
}

// renderer.cpp:76-84: clamp each SAMPLE to <= 1 per channel (Q2), then sum the pixel's samples.
// One warp per pixel: lanes stride over the pixel's samples (coalesced), shuffle-reduce.
__global__ void __launch_bounds__(kThreads) k_resolve(const __grid_constant__ BatchParams bp, const float *radiance,
                                                      uint32_t capacity, float *accum) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, num_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t p = warp; p < bp.pixel_count; p += num_warps) {
        float r = 0.0f, g = 0.0f, b = 0.0f;
        const uint32_t base = p * bp.sample_count;
        for (uint32_t s = lane; s < bp.sample_count; s += 32) {
            const float4 v = reinterpret_cast<const float4 *>(radiance)[base + s];
            r += fminf(v.x, 1.0f);
            g += fminf(v.y, 1.0f);
            b += fminf(v.z, 1.0f);
        }
        for (int o = 16; o > 0; o >>= 1) {
            r += __shfl_down_sync(0xffffffffu, r, o);
            g += __shfl_down_sync(0xffffffffu, g, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            float *a = accum + 3ull * JobPixelToLocal(bp, bp.pixel_begin + p);
            a[0] += r, a[1] += g, a[2] += b;
        }
    }
}

// mean over spp (renderer.cpp:82-84) and scatter from tile order to the row-major frame.
__global__ void k_finalize(const __grid_constant__ BatchParams bp, uint32_t num_local_pixels, const float *accum, float *frame,
                           float *tiles) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < num_local_pixels; p += gridDim.x * blockDim.x) {
        const float r = accum[3ull * p] * bp.spp_inv, g = accum[3ull * p + 1] * bp.spp_inv, b = accum[3ull * p + 2] * bp.spp_inv;
        if (tiles != nullptr) {
            tiles[3ull * p] = r, tiles[3ull * p + 1] = g, tiles[3ull * p + 2] = b;
        }
        uint32_t i, j;
        if (frame != nullptr && LocalPixelToImage(bp, p, &i, &j)) {
            float *dst = frame + 3ull * (static_cast<uint64_t>(j) * bp.width + i);
            dst[0] = r, dst[1] = g, dst[2] = b;
        }
    }
}

// renderer.cpp:118-136: running mean over the preview frames + sRGB transfer into a bottom-up copy.
__global__ void k_finalize_progressive(const __grid_constant__ BatchParams bp, uint32_t num_local_pixels, const float *accum,
                                       uint32_t frame_index, float *frame, float *frame_srgb) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < num_local_pixels; p += gridDim.x * blockDim.x) {
        uint32_t i, j;
        if (!LocalPixelToImage(bp, p, &i, &j)) continue;
        const uint64_t offset = 3ull * (static_cast<uint64_t>(j) * bp.width + i),
                       offset_dest = 3ull * (static_cast<uint64_t>(bp.height - 1 - j) * bp.width + i);
        for (int c = 0; c < 3; ++c) {
            const float color = accum[3ull * p + c]; // one sample, already clamped to <= 1 by k_resolve
            const float mean = (frame_index * frame[offset + c] + color) / (frame_index + 1);
            frame[offset + c] = mean;
            if (frame_srgb != nullptr)
                frame_srgb[offset_dest + c] = mean <= 0.0031308f ? 12.92f * mean : 1.055f * powf(mean, 1.0f / 2.4f) - 0.055f;
        }
    }
}

// gathered = [tile_world][pixels_per_rank*3]: the all-gathered per-rank tile buffers.
__global__ void k_assemble(uint32_t width, uint32_t height, uint32_t tile_world, uint32_t pixels_per_rank, const float *gathered,
                           float *frame) {
    BatchParams bp{};
    bp.width = width, bp.height = height;
    bp.tiles_x = (width + kTileSize - 1) / kTileSize;
    bp.num_tiles = bp.tiles_x * ((height + kTileSize - 1) / kTileSize);
    bp.tile_world = tile_world;
    const uint64_t total = static_cast<uint64_t>(pixels_per_rank) * tile_world;
    for (uint64_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        bp.tile_rank = static_cast<uint32_t>(k / pixels_per_rank);
        const uint32_t p = static_cast<uint32_t>(k % pixels_per_rank);
        uint32_t i, j;
        if (LocalPixelToImage(bp, p, &i, &j)) {
            const float *src = gathered + 3ull * k;
            float *dst = frame + 3ull * (static_cast<uint64_t>(j) * width + i);
            dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
        }
    }
}

size_t TopSmemBytes(const LaunchConfig &lc) { return static_cast<size_t>(std::max(lc.top_nodes, 1)) * sizeof(WideNode); }

// Dynamic shared memory above 48 KB needs an opt-in per kernel function (only reached with B200PT_TOP_NODES > 614).
template <typename K>
void EnableSmem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWideTopNodesMax * sizeof(WideNode)));
}

} // namespace

// Picks the <STATS, OPACITY, LAYOUT> instantiation of a traversal kernel and launches it on the persistent grid.
// The layout follows the tree the scene was created with; lc.top_nodes > 0 selects the wide variant that stages the top
// of the tree in shared memory.
#define B200PT_LAUNCH_TRAVERSAL(kernel, ...)                                                                          \
    do {                                                                                                              \
        const bool opacity = scene.integrator.has_opacity != 0, wide = scene.num_wide_nodes > 0;                     \
        const bool top = wide && lc.top_nodes > 0;                                                                    \
        const size_t smem = top ? TopSmemBytes(lc) : 0;                                                               \
        const int phase_lanes = wide ? lc.tri_min : lc.min_inner;                                                     \
        auto go = [&](auto k) {                                                                                       \
            EnableSmem(k, smem);                                                                                      \
            k<<<lc.blocks, kThreads, smem, lc.stream>>>(__VA_ARGS__, lc.top_nodes, lc.refill, phase_lanes, lc.packets); \
        };                                                                                                            \
        auto pick_layout = [&](auto binary, auto wide_plain, auto wide_top) {                                         \
            if (!wide) go(binary);                                                                                    \
            else if (top) go(wide_top);                                                                               \
            else go(wide_plain);                                                                                      \
        };                                                                                                            \
        if (lc.stats && opacity) pick_layout(kernel<true, true, 0>, kernel<true, true, 1>, kernel<true, true, 2>);    \
        else if (lc.stats) pick_layout(kernel<true, false, 0>, kernel<true, false, 1>, kernel<true, false, 2>);       \
        else if (opacity) pick_layout(kernel<false, true, 0>, kernel<false, true, 1>, kernel<false, true, 2>);        \
        else pick_layout(kernel<false, false, 0>, kernel<false, false, 1>, kernel<false, false, 2>);                  \
    } while (0)

int LaunchPrimary(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, PathQueue q, ShadeBins bins,
                   float *radiance, uint32_t capacity, Counters *counters) {
    B200PT_LAUNCH_TRAVERSAL(k_primary, scene, bp, q, radiance, capacity, counters);
    if (bins.lists == nullptr) return 1;
    k_bin_hits<<<lc.blocks, kThreads, 0, lc.stream>>>(scene, q.hit, 0, bins, counters);
    return 2;
}

int LaunchTrace(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue q, int which,
                 ShadeBins bins, ShadowQueue sq, float *radiance, uint32_t capacity, Counters *counters) {
    B200PT_LAUNCH_TRAVERSAL(k_trace, scene, bp, depth, q, which, sq, radiance, capacity, counters);
    if (bins.lists == nullptr || which < 0) return 1;
    k_bin_hits<<<lc.blocks, kThreads, 0, lc.stream>>>(scene, q.hit, which, bins, counters);
    return 2;
}

int LaunchShade(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue qin,
                int which_in, PathQueue qout, ShadeBins bins, ShadowQueue sq, float *radiance, Counters *counters, uint32_t capacity) {
    const bool vol = scene.integrator.type == B200PT_INTEGRATOR_VOLPATH;
    // the kernel specialised to BSDF model `only` (kAnyBsdf: the generic one)
    auto variant = [](int only) {
        switch (only) {
        case 0: case B200PT_BSDF_AREA_LIGHT: return LaunchShadeVariant_0;
        case B200PT_BSDF_DIFFUSE: return LaunchShadeVariant_2;
        case B200PT_BSDF_ROUGH_DIFFUSE: return LaunchShadeVariant_3;
        case B200PT_BSDF_CONDUCTOR: return LaunchShadeVariant_4;
        case B200PT_BSDF_DIELECTRIC: return LaunchShadeVariant_5;
        case B200PT_BSDF_THIN_DIELECTRIC: return LaunchShadeVariant_6;
        case B200PT_BSDF_PLASTIC: return LaunchShadeVariant_7;
        default: return LaunchShadeVariant_any;
        }
    };
    const uint32_t in_use = scene.integrator.shade_bins;
    if (bins.lists == nullptr || in_use == 0) {
        variant(lc.shade_only)(vol, lc.blocks, lc.stream, scene, bp, depth, qin, which_in, qout, sq, radiance, counters, capacity, nullptr, 0);
        return 1;
    }
    int launches = 0;
    for (int bin = 0; bin < kNumShadeBins; ++bin) {
        if (!((in_use >> bin) & 1u)) continue;
        const uint32_t *list = bins.lists + static_cast<uint64_t>(BinListIndex(in_use, bin)) * bins.capacity;
        variant(bin)(vol, lc.blocks, lc.stream, scene, bp, depth, qin, which_in, qout, sq, radiance, counters, capacity, list, bin);
        ++launches;
    }
    return launches;
}

void LaunchDebugTrace(const LaunchConfig &lc, const DeviceScene &scene, const b200pt_debug_ray *rays, uint32_t n, bool any_hit, bool single,
                      bool raw_prim, bool packet, b200pt_debug_hit *out, uint32_t *work_counter) {
    if (scene.num_wide_nodes > 0)
        k_debug_trace<kLayoutWide><<<lc.blocks, kThreads, 0, lc.stream>>>(scene, rays, n, any_hit, single, raw_prim, false, out, work_counter, lc.refill, lc.tri_min);
    else
        k_debug_trace<kLayoutBinary><<<lc.blocks, kThreads, 0, lc.stream>>>(scene, rays, n, any_hit, single, raw_prim, packet, out, work_counter, lc.refill, lc.min_inner);
}

void LaunchCullTiles(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t num_local_tiles,
                     unsigned long long *masks, uint32_t *offsets, uint32_t *pixel_list, uint32_t *counts) {
    const int warps_per_cta = kThreads / 32;
    const int blocks = std::max(1, static_cast<int>(std::min<uint32_t>(lc.blocks, (num_local_tiles + warps_per_cta - 1) / warps_per_cta)));
    k_cull_tiles<<<blocks, kThreads, 0, lc.stream>>>(bp, scene.cull_boxes, scene.num_cull_boxes, scene.fine_cull_boxes, scene.num_fine_cull_boxes,
                                                    scene.cull_fine_begin, num_local_tiles, masks);
    k_scan_tiles<<<1, 1024, 0, lc.stream>>>(masks, num_local_tiles, offsets, counts);
    k_list_pixels<<<blocks, kThreads, 0, lc.stream>>>(masks, offsets, num_local_tiles, pixel_list);
}

void LaunchSettle(const LaunchConfig &lc, Counters *counters, int which_queue, bool reset_shadow, ShadowQueue sq, float *radiance,
                  uint32_t capacity, bool unique_slots) {
    if (unique_slots)
        k_settle<true><<<lc.blocks, kThreads, 0, lc.stream>>>(counters, which_queue, reset_shadow, sq, radiance, capacity);
    else
        k_settle<false><<<lc.blocks, kThreads, 0, lc.stream>>>(counters, which_queue, reset_shadow, sq, radiance, capacity);
}

void LaunchResolve(const LaunchConfig &lc, const BatchParams &bp, const float *radiance, uint32_t capacity, float *accum) {
    k_resolve<<<lc.blocks, kThreads, 0, lc.stream>>>(bp, radiance, capacity, accum);
}

void LaunchFinalize(const LaunchConfig &lc, const BatchParams &bp, uint32_t num_local_pixels, const float *accum, float *frame,
                    float *tiles) {
    const int blocks = static_cast<int>(std::min<uint64_t>(lc.blocks, (num_local_pixels + kThreads - 1) / kThreads));
    k_finalize<<<std::max(blocks, 1), kThreads, 0, lc.stream>>>(bp, num_local_pixels, accum, frame, tiles);
}

void LaunchFinalizeProgressive(const LaunchConfig &lc, const BatchParams &bp, uint32_t num_local_pixels, const float *accum,
                               uint32_t frame_index, float *frame, float *frame_srgb) {
    const int blocks = static_cast<int>(std::min<uint64_t>(lc.blocks, (num_local_pixels + kThreads - 1) / kThreads));
    k_finalize_progressive<<<std::max(blocks, 1), kThreads, 0, lc.stream>>>(bp, num_local_pixels, accum, frame_index, frame, frame_srgb);
}

void LaunchAssemble(const LaunchConfig &lc, uint32_t width, uint32_t height, uint32_t tile_world, uint32_t pixels_per_rank,
                    const float *gathered, float *frame) {
    k_assemble<<<lc.blocks, kThreads, 0, lc.stream>>>(width, height, tile_world, pixels_per_rank, gathered, frame);
}

} // namespace b200pt
