// tail_kernel.cu — k_tail: the last few thousand paths of a batch, one path per lane, to the end.
//
// The wavefront pipeline pays, per bounce, at least the latency of its slowest ray (a launch cannot end earlier: ~0.1-0.2 ms
// on Dragon) plus three launches.  That is nothing while a bounce carries millions of rays, and almost everything once
// Russian roulette and escaping rays have left a few thousand.  When the survivor queue of a batch has shrunk below a
// threshold, ONE launch of this kernel takes the queue over and walks every remaining path through all its remaining
// vertices — the reference's own loop structure (one thread, one path: ShadePath path.cpp:8-236) — with the same
// ShadeVertex code and the same counter-based random numbers as the wavefront kernels, so the samples are bit-identical
// to what the per-bounce launches would have produced (tests/test_gpu_parity.py::test_tail_kernel_is_bit_exact).
// Measured (profiles/r01_sweep_tail_kernel.log): the gain is small, because the tail is bounded by the LONGEST surviving
// path, whose vertices cost the same serial latency here as in per-bounce launches; it saves the launches and the
// per-bounce "slowest ray" maxima: Dragon on one of 8 ranks 7.22 -> 7.03 ms, matpreview -3 %, Dragon on one GPU +-0.
// Taking over earlier (more paths) loses: the generic shading code and lane-divergent traversal are much less efficient
// per path than the binned, compacted wavefront kernels.
#include "shade_kernel.cuh"
#include "traverse_wide.cuh"

namespace b200pt {

namespace {

constexpr int kTailThreads = 128;

// The per-lane traversal of whichever tree the scene was created with (a warp-uniform choice).
__device__ __forceinline__ bool TraceSingle(const DeviceScene &scene, const Ray &ray, bool any, bool opacity, const Rng &rng, HitRec *hit,
                                            bool stats, TraversalCounters *counters) {
    if (scene.num_wide_nodes > 0) return TraverseSingleWide(scene, ray, any, opacity, rng, hit, stats, counters);
    return TraverseSingle(scene, ray, any, opacity, rng, hit, stats, counters);
}

template <bool VOL>
__global__ void __launch_bounds__(kTailThreads) k_tail(const __grid_constant__ DeviceScene scene, const __grid_constant__ BatchParams bp,
                                                      uint32_t depth0, PathQueue q, int which, float *radiance, uint32_t capacity,
                                                      Counters *counters, uint32_t threshold, bool stats) {
    // Every CTA sees the same queue length (nothing modifies it during this launch) and takes the same decision.
    const uint32_t n = counters->queue[which];
    const uint32_t token = depth0 + 1;
    const uint32_t taken = counters->tail_taken;
    if (n == 0 || n > threshold || (taken != 0 && taken != token)) return;
    if (threadIdx.x == 0) counters->tail_taken = token; // the wavefront kernels of the later bounces find nothing to do
    const DIntegrator &ig = scene.integrator;
    const bool opacity = ig.has_opacity != 0;
    const uint32_t lane = threadIdx.x & 31u;
    TraversalCounters tc[2];
    uint32_t rays[2] = {0, 0};
    // The paths are spread over ALL warps of the launch, as few per warp as that allows: the tail is bound by latency, not by
    // lanes, and a warp walks the code of every one of its paths (traversal steps, leaves, shading) one after the other.
    const uint32_t total_warps = gridDim.x * (kTailThreads / 32);
    const uint32_t per_warp = min(32u, max(1u, (n + total_warps - 1u) / total_warps));
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counters->work_tail, per_warp);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        PathVertex v;
        bool alive = LoadPathVertex<VOL>(scene, q, i, lane < per_warp && i < n, &v);
        const uint32_t slot = v.slot;
        for (uint32_t depth = depth0; __any_sync(0xffffffffu, alive); ++depth) {
            PathNext next;
            V3 Ladd;
            const bool was_alive = alive;
            alive = ShadeVertex<VOL, kAnyBsdf>(
                scene, bp, depth, alive, v,
                [&](const ShadowCandidate &sc) { // next-event estimation: trace the shadow ray right away
                    if (!(sc.valid && (sc.c.x != 0.0f || sc.c.y != 0.0f || sc.c.z != 0.0f))) return;
                    Ray ray;
                    ray.o = sc.o, ray.d = sc.d, ray.tmin = kEpsilonDistance, ray.tmax = sc.tmax;
                    uint3 ctr = make_uint3(0, 0, 0);
                    if (opacity) { // the same alpha-test stream as k_trace gives this ray
                        ctr = SlotCounter(bp, slot, depth);
                        ctr.y ^= (__float_as_uint(ray.tmax) ^ __float_as_uint(ray.d.x) * 0x85ebca6bu) * 0x9e3779b9u;
                    }
                    HitRec unused;
                    ++rays[1];
                    if (!TraceSingle(scene, ray, true, opacity, Rng(ctr.x, ctr.y, ctr.z, bp.key, kRngDomainShadow), &unused, stats, &tc[1])) {
                        RadianceAdd(radiance, slot, sc.c.x, sc.c.y, sc.c.z);
                    }
                },
                &next, &Ladd);
            if (was_alive && (Ladd.x != 0.0f || Ladd.y != 0.0f || Ladd.z != 0.0f)) {
                RadianceAdd(radiance, slot, Ladd.x, Ladd.y, Ladd.z);
            }
            if (depth >= kMaxTailDepth) alive = false; // the host loop's hard stop (renderer.cu: kMaxRounds)
            if (alive) { // extend the path: closest hit of the sampled direction
                v.ray.o = next.o, v.ray.d = next.d, v.ray.tmin = kEpsilonDistance, v.ray.tmax = kMaxFloat;
                v.att = next.att, v.pdf_sample = next.pdf, v.ray_medium = next.medium, v.wo_prev = next.wo;
                uint3 ctr = make_uint3(0, 0, 0);
                if (opacity) ctr = SlotCounter(bp, slot, depth);
                ++rays[0];
                TraceSingle(scene, v.ray, false, opacity, Rng(ctr.x, ctr.y, ctr.z, bp.key, kRngDomainClosest), &v.hit, stats, &tc[0]);
                // same shortcut as LoadPathVertex: an escaped ray with no environment map to see is finished
                if (!VOL && v.hit.prim == kPrimMiss && ig.id_envmap == kInvalid) alive = false;
            }
        }
    }
    if (stats) {
        for (int k = 0; k < 2; ++k) {
            uint32_t nodes = tc[k].nodes, prims = tc[k].prims, r = rays[k];
            for (int o = 16; o > 0; o >>= 1) {
                nodes += __shfl_down_sync(0xffffffffu, nodes, o);
                prims += __shfl_down_sync(0xffffffffu, prims, o);
                r += __shfl_down_sync(0xffffffffu, r, o);
            }
            if (lane == 0) {
                ClassCounters &cc = counters->cls[k == 0 ? kClassExtend : kClassShadow];
                atomicAdd(&cc.node_visits, static_cast<unsigned long long>(nodes));
                atomicAdd(&cc.prim_tests, static_cast<unsigned long long>(prims));
                atomicAdd(&cc.rays, static_cast<unsigned long long>(r));
            }
        }
    }
}

} // namespace

void LaunchTail(const LaunchConfig &lc, const DeviceScene &scene, const BatchParams &bp, uint32_t depth, PathQueue q, int which,
                float *radiance, uint32_t capacity, Counters *counters, uint32_t threshold) {
    // enough CTAs for `threshold` paths at one path per WARP, capped at two waves of a full machine
    const int blocks = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(lc.blocks) * 2u, (threshold + kTailThreads / 32 - 1) / (kTailThreads / 32)));
    if (scene.integrator.type == B200PT_INTEGRATOR_VOLPATH)
        k_tail<true><<<std::max(blocks, 1), kTailThreads, 0, lc.stream>>>(scene, bp, depth, q, which, radiance, capacity, counters, threshold, lc.stats);
    else
        k_tail<false><<<std::max(blocks, 1), kTailThreads, 0, lc.stream>>>(scene, bp, depth, q, which, radiance, capacity, counters, threshold, lc.stats);
}

} // namespace b200pt
