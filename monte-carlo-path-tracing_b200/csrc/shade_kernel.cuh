// shade_kernel.cuh — k_shade, the shading stage of the wavefront pipeline, as a template over the integrator (path /
// volpath) and over the BSDF model it is specialised to.  Instantiated once per model by shade_variant.cu (one
// translation unit per model, compiled in parallel), launched through LaunchShade (wavefront.cu).
#pragma once
#include "shading.cuh"
#include "wavefront.cuh"

namespace b200pt {

constexpr int kShadeThreads = 256;

// Warp-ballot compaction: every lane of the warp must call this; lanes with flag get a unique
// slot in the queue whose length is *counter.
__device__ __forceinline__ uint32_t WarpAppend(bool flag, uint32_t *counter) {
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    if (ballot == 0) return 0;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ballot) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(ballot & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------------------
// k_shade
// ---------------------------------------------------------------------------------------------
struct ShadowCandidate {
    bool valid = false;
    V3 o = {0, 0, 0}, d = {0, 0, 0}, c = {0, 0, 0};
    float tmax = 0;
};

__device__ __forceinline__ void PushShadow(const ShadowCandidate &sc, uint32_t slot, ShadowQueue sq, Counters *counters) {
    const bool ok = sc.valid && (sc.c.x != 0.0f || sc.c.y != 0.0f || sc.c.z != 0.0f);
    const uint32_t idx = WarpAppend(ok, &counters->shadow);
    if (ok) {
        sq.ox[idx] = sc.o.x, sq.oy[idx] = sc.o.y, sq.oz[idx] = sc.o.z;
        sq.dx[idx] = sc.d.x, sq.dy[idx] = sc.d.y, sq.dz[idx] = sc.d.z;
        sq.tmax[idx] = sc.tmax;
        sq.cr[idx] = sc.c.x, sq.cg[idx] = sc.c.y, sq.cb[idx] = sc.c.z;
        sq.slot[idx] = slot;
    }
}

// Transmittance/pdf of a shadow segment through the medium at the shading point
// (volpath.cpp:276-285, 341-350, 397-402, 456-461).
__device__ __forceinline__ bool ShadowMediumAttenuation(const DMedium *medium, float distance, V3 *att) {
    *att = mk3(1.0f);
    if (medium == nullptr) return true;
    MediumRec mrec;
    mrec.distance = distance;
    MediumEvaluate(*medium, &mrec);
    if (!mrec.valid) return false;
    *att = mrec.att / mrec.pdf;
    return true;
}

// One path at the vertex it has just arrived at: exactly what a path-queue entry holds.
struct PathVertex {
    Ray ray;              // the segment that ends here (tmax is set from the hit record)
    V3 att;               // throughput
    V3 wo_prev;           // volpath: wo of the last surface vertex
    float pdf_sample;     // pdf of the direction sample that produced the segment
    uint32_t slot, ray_medium;
    HitRec hit;
#ifdef B200PT_RNG_REPLAY
    uint32_t replay_state; // test build (debug_eval.cu): the per-pixel LCG runs on from vertex to vertex, as in the reference
#endif
};

// The continuation a vertex produces (valid when ShadeVertex returns true).
struct PathNext {
    V3 o, d, att, wo;
    float pdf;
    uint32_t medium;
#ifdef B200PT_RNG_REPLAY
    uint32_t replay_state;
#endif
};

// One iteration of the reference's path loop (ShadePath path.cpp:8-236 / ShadeVolPath volpath.cpp:8-485) for the vertex `v`,
// minus the ray casts: radiance picked up at the vertex goes to *Ladd_out, every next-event-estimation candidate is handed
// to `emit_shadow` (called by ALL lanes of the warp the same number of times, with sc.valid = false where there is
// nothing to test, so that the caller may compact with warp ballots), the sampled continuation to *next.
// Returns whether the path continues.  Shared by the wavefront kernel k_shade and the path-at-a-time tail kernel k_tail.
template <bool VOL, int ONLY, typename EmitShadow>
__device__ __forceinline__ bool ShadeVertex(const DeviceScene &scene, const BatchParams &bp, uint32_t depth, bool alive, const PathVertex &v,
                                            EmitShadow emit_shadow, PathNext *next, V3 *Ladd_out) {
    const DIntegrator &ig = scene.integrator;
    Ray ray = v.ray;
    V3 att = v.att, wo = alive ? -ray.d : mk3(0.0f), wo_prev = v.wo_prev, Ladd = mk3(0.0f);
    float pdf_sample = v.pdf_sample;
    uint32_t ray_medium = v.ray_medium;
    const HitRec hit = v.hit;
#ifdef B200PT_RNG_REPLAY
    Rng rng(v.replay_state);
#else
    const uint32_t local_pixel = JobPixelToLocal(bp, bp.pixel_begin + v.slot / bp.sample_count);
    const uint32_t sample = bp.sample_begin + v.slot % bp.sample_count;
    uint32_t px = 0, py = 0;
    LocalPixelToImage(bp, local_pixel, &px, &py);
    Rng rng(py * bp.width + px, sample, depth, bp.key);
#endif

    // ---- the vertex this segment arrives at ----
    const bool has_hit = hit.prim != kPrimMiss;
    Surf surf;
    surf.inside = false, surf.inst = 0, surf.uv = {0, 0};
    surf.pos = surf.n = surf.t = surf.b = mk3(0.0f);
    const DBsdf *bsdf = nullptr;
    if (alive && has_hit) {
        ray.tmax = hit.t;
        if (hit.prim & kPrimAnalyticBit)
            surf = SurfAnalytic(scene, hit.prim & ~kPrimAnalyticBit, ray, hit.t);
        else
            surf = SurfTriangle(scene, hit.prim & kPrimIndexMask, hit.u, hit.v, (hit.prim & kPrimInsideBit) != 0);
        const uint32_t id_bsdf = scene.instances[surf.inst].id_bsdf;
        if (id_bsdf != kInvalid) bsdf = scene.bsdfs + id_bsdf;
    }

    // ---- participating medium along the segment (volpath.cpp:44-60, 119-137, 163-186) ----
    bool scattering = false;
    const DMedium *vertex_medium = nullptr; // medium of a scattering vertex
    V3 vertex_pos = surf.pos;
    if (VOL && alive) {
        uint32_t id_medium = ray_medium;
        if (id_medium == kInvalid && has_hit) {
            const bool inside = Dot(-ray.d, surf.n) > 0 ? surf.inside : !surf.inside;
            id_medium = inside ? scene.instances[surf.inst].id_medium_int : scene.instances[surf.inst].id_medium_ext;
        }
        if (id_medium != kInvalid) {
            MediumRec mrec;
            MediumSample(scene.media[id_medium], ray.tmax, rng, &mrec);
            if (mrec.valid) {
                att *= mrec.att / mrec.pdf;
                if (mrec.scattered) {
                    scattering = true;
                    vertex_pos = ray.o + ray.d * mrec.distance;
                    vertex_medium = scene.media + id_medium;
                    ray_medium = id_medium;
                    // volpath.cpp only refreshes `wo` at surface vertices (:233): a medium vertex keeps
                    // the wo that was used at the previous vertex.
                    wo = wo_prev;
                }
            }
        }
    }

    // ---- arrival at a surface / escape (path.cpp:24-53 for depth 1, :81-132 afterwards) ----
    if (alive && !scattering) {
        if (!has_hit) {
            // depth 1 misses are finished inside k_primary
            if (depth > 1 && ig.id_envmap != kInvalid) {
                const DEmitter &env = scene.emitters[ig.id_envmap];
                const V3 Le = EmitterEvaluateDir(scene, env, ray.d);
                const float pdf_direct = EmitterPdf(scene, env, ray.d), w = MisWeight(pdf_sample, pdf_direct);
                Ladd += w * att * Le;
            }
            alive = false;
        } else if (bsdf != nullptr) {
            if (surf.inside && !bsdf->twosided) {
                alive = false; // back of a one-sided surface absorbs
            } else if (bsdf->type == B200PT_BSDF_AREA_LIGHT) {
                if (depth == 1) {
                    if (!ig.hide_emitters) Ladd += TexColor(scene, bsdf->id_radiance, surf.uv);
                } else {
                    const float cos_theta_prime = Dot(-ray.d, surf.n);
                    if (cos_theta_prime >= kEpsilonFloat) {
                        const uint32_t light = scene.instances[surf.inst].area_light;
                        const float pdf_area = (__ldg(scene.cdf_area_light + light + 1) - __ldg(scene.cdf_area_light + light)) *
                                               scene.instances[surf.inst].pdf_area,
                                    pdf_direct = pdf_area * Sqr(ray.tmax) / cos_theta_prime,
                                    w = MisWeight(pdf_sample, pdf_direct);
                        Ladd += w * att * TexColor(scene, bsdf->id_radiance, surf.uv);
                    }
                }
                alive = false;
            }
        }
        if (alive) {
            wo = -ray.d;
            if (depth > 1 && depth - 1 >= ig.depth_rr) att *= ig.pdf_rr_rcp; // Q1
        }
    }

    // ---- loop condition of iteration `depth` (path.cpp:57-60) ----
    if (alive) {
        if (!(depth < ig.depth_rr || (depth < ig.depth_max && rng.Next() < ig.pdf_rr))) alive = false;
    }

    // ---- next-event estimation (path.cpp:138-236, volpath.cpp:247-485) ----
    const DMedium *nee_medium = nullptr;
    if (VOL && alive) {
        if (scattering) {
            nee_medium = vertex_medium;
        } else {
            const bool inside = Dot(wo, surf.n) > 0 ? surf.inside : !surf.inside;
            const uint32_t id_medium = inside ? scene.instances[surf.inst].id_medium_int : scene.instances[surf.inst].id_medium_ext;
            if (id_medium != kInvalid) nee_medium = scene.media + id_medium;
        }
    }
    for (uint32_t e = 0; e < ig.num_emitters; ++e) {
        ShadowCandidate sc;
        if (alive) {
            const DEmitter &em = scene.emitters[e];
            const float xi_1 = rng.Next(), xi_0 = rng.Next(); // path.cpp:149 argument order as GCC evaluates it (Q16)
            const EmitterRec erec = EmitterSample(scene, em, vertex_pos, xi_0, xi_1);
            bool ok = erec.valid;
            if (ok && !scattering && Dot(-erec.wi, surf.n) < kEpsilonFloat) ok = false;
            V3 medium_att = mk3(1.0f);
            if (ok && VOL && !ShadowMediumAttenuation(nee_medium, erec.distance, &medium_att)) ok = false;
            V3 f = mk3(0.0f);
            float pdf_f = 0.0f;
            if (ok) {
                if (scattering) {
                    PhaseRec prec;
                    prec.wi = erec.wi, prec.wo = wo;
                    PhaseEvaluate(*vertex_medium, &prec);
                    ok = prec.valid;
                    f = prec.att, pdf_f = prec.pdf;
                } else {
                    const BsdfRec brec = EvaluateRayPath<ONLY>(scene, erec.wi, wo, surf, bsdf);
                    ok = brec.valid;
                    f = brec.att, pdf_f = brec.pdf;
                }
            }
            if (ok) {
                const V3 Le = EmitterEvaluateRec(scene, em, erec);
                if (erec.harsh) {
                    sc.c = att * (Le * medium_att * f);
                } else {
                    const float pdf_direct = EmitterPdf(scene, em, -erec.wi);
                    if (pdf_direct > kEpsilonFloat)
                        sc.c = att * (MisWeight(pdf_direct, pdf_f) * Le * medium_att * f / pdf_direct);
                    else
                        ok = false;
                }
            }
            if (ok) {
                sc.valid = true;
                sc.o = vertex_pos, sc.d = -erec.wi;
                sc.tmax = erec.distance - kEpsilonDistance;
            }
        }
        emit_shadow(sc);
    }
    if (ig.num_area_lights != 0) {
        ShadowCandidate sc;
        if (alive) {
            const float xi_l = rng.Next();
            const uint32_t index_area_light = BinarySearch(ig.num_area_lights + 1, scene.cdf_area_light, xi_l) - 1; // Q4
            const uint32_t light_inst = __ldg(scene.map_area_light_instance + index_area_light);
            const float xi_2 = rng.Next(), xi_1 = rng.Next(), xi_0 = rng.Next(); // path.cpp:196 (Q16)
            const LightPoint lp = SampleInstance(scene, light_inst, xi_0, xi_1, xi_2);
            const V3 d_vec = vertex_pos - lp.pos;
            const float distance = Length(d_vec);
            const V3 wi = Normalize(d_vec);
            const float cos_theta_prime = Dot(wi, lp.n);
            bool ok = cos_theta_prime >= kEpsilonFloat;
            if (ok && !scattering && Dot(-wi, surf.n) < kEpsilonFloat) ok = false;
            V3 medium_att = mk3(1.0f);
            if (ok && VOL && !ShadowMediumAttenuation(nee_medium, distance, &medium_att)) ok = false;
            V3 f = mk3(0.0f);
            float pdf_f = 0.0f;
            if (ok) {
                if (scattering) {
                    PhaseRec prec;
                    prec.wi = wi, prec.wo = wo;
                    PhaseEvaluate(*vertex_medium, &prec);
                    ok = prec.valid;
                    f = prec.att, pdf_f = prec.pdf;
                } else {
                    const BsdfRec brec = EvaluateRayPath<ONLY>(scene, wi, wo, surf, bsdf);
                    ok = brec.valid;
                    f = brec.att, pdf_f = brec.pdf;
                }
            }
            if (ok) {
                const float pdf_area = (__ldg(scene.cdf_area_light + index_area_light + 1) -
                                        __ldg(scene.cdf_area_light + index_area_light)) *
                                       scene.instances[light_inst].pdf_area,
                            pdf_direct = pdf_area * Sqr(distance) / cos_theta_prime, w = MisWeight(pdf_direct, pdf_f);
                const DBsdf &light_bsdf = scene.bsdfs[scene.instances[light_inst].id_bsdf];
                const V3 Le = TexColor(scene, light_bsdf.id_radiance, lp.uv);
                sc.c = att * (w * (Le * medium_att * f / pdf_direct));
                sc.valid = true;
                sc.o = lp.pos, sc.d = wi; // traced from the light towards the shading point (path.cpp:201-202)
                sc.tmax = distance - kEpsilonDistance;
            }
        }
        emit_shadow(sc);
    }

    // ---- sample the continuation (path.cpp:66-79, volpath.cpp:98-117, 147-161) ----
    V3 next_d = mk3(0.0f);
    if (alive) {
        V3 wi, f;
        float pdf;
        bool ok;
        if (scattering) {
            PhaseRec prec;
            prec.wo = wo;
            PhaseSample(*vertex_medium, rng, &prec);
            ok = prec.valid;
            wi = prec.wi, f = prec.att, pdf = prec.pdf;
        } else {
            const BsdfRec brec = SampleRayPath<ONLY>(scene, wo, surf, bsdf, rng);
            ok = brec.valid;
            wi = brec.wi, f = brec.att, pdf = brec.pdf;
        }
        if (!ok) {
            alive = false;
        } else {
            att *= f / pdf;
            pdf_sample = pdf;
            if (MaxComp(att) < kEpsilon) alive = false;
            next_d = -wi;
        }
    }
    next->o = vertex_pos, next->d = next_d, next->att = att, next->wo = wo;
    next->pdf = pdf_sample;
    next->medium = scattering ? ray_medium : kInvalid;
#ifdef B200PT_RNG_REPLAY
    next->replay_state = rng.state;
#endif
    *Ladd_out = Ladd;
    return alive;
}

// Loads queue entry `i` (the dead-entry shortcut included); returns whether there is anything to shade.
template <bool VOL>
__device__ __forceinline__ bool LoadPathVertex(const DeviceScene &scene, const PathQueue &qin, uint32_t i, bool active, PathVertex *v) {
    bool alive = active;
    v->ray.o = v->ray.d = mk3(0.0f);
    v->ray.tmin = kEpsilonDistance, v->ray.tmax = kMaxFloat;
    v->att = v->wo_prev = mk3(0.0f);
    v->pdf_sample = 0.0f;
    v->slot = 0, v->ray_medium = kInvalid;
    v->hit.prim = kPrimMiss, v->hit.t = kMaxFloat, v->hit.u = v->hit.v = 0.0f;
    if (active) {
        v->hit = qin.hit[i];
        // A ray that left the scene contributes only through an environment map (path.cpp:81-93); without one
        // (and outside volpath, where the segment may still scatter) its queue entry is dead: skip the other 52 bytes.
        if (!VOL && v->hit.prim == kPrimMiss && scene.integrator.id_envmap == kInvalid) alive = false;
    }
    if (alive) {
        v->ray.o = mk3(qin.ox[i], qin.oy[i], qin.oz[i]);
        v->ray.d = mk3(qin.dx[i], qin.dy[i], qin.dz[i]);
        v->att = mk3(qin.tr[i], qin.tg[i], qin.tb[i]);
        v->pdf_sample = qin.pdf[i];
        v->slot = qin.slot[i];
        if (VOL) {
            v->ray_medium = qin.medium[i];
            v->wo_prev = mk3(qin.wx[i], qin.wy[i], qin.wz[i]);
        }
    }
    return alive;
}

// Resident CTAs per SM a variant is compiled for.  Left to the compiler the variants take 96-128 registers (2 CTAs, 25 %
// occupancy) and wait for their gathers; capped, they spill 100-200 bytes per thread and win all the same: 4 CTAs (64
// registers) for the path integrator (matpreview 105.0 -> 98.2 ms, lte-orb 111.9 -> 103.8, box 105.1 -> 96.2, material-testball
// 40.9 -> 37.4, Dragon 5.4 ms shade with 4 vs 6.6 with 3), 3 CTAs (80 registers) for volpath, whose vertex state is larger
// (volumetric-caustic 267.3 -> 242.7 ms with 3, 258.8 with 4): profiles/r02_sweep_shade_ctas_leaf_step.log.
#ifdef B200PT_SHADE_MIN_CTAS_OVERRIDE   // experiments: one bound for every variant of the translation unit
#define B200PT_SHADE_MIN_CTAS(VOL, ONLY) B200PT_SHADE_MIN_CTAS_OVERRIDE
#else
#define B200PT_SHADE_MIN_CTAS(VOL, ONLY) ((VOL) ? 3 : 4)
#endif
template <bool VOL, int ONLY>
__global__ void __launch_bounds__(kShadeThreads, B200PT_SHADE_MIN_CTAS(VOL, ONLY)) k_shade(const __grid_constant__ DeviceScene scene,
                                                    const __grid_constant__ BatchParams bp, uint32_t depth, PathQueue qin,
                                                    int which_in, PathQueue qout, ShadowQueue sq, float *radiance,
                                                    Counters *counters, uint32_t capacity, const uint32_t *bin_list,
                                                    int bin) {
    // bin_list == nullptr: every entry of queue `which_in`; else the entries whose surface falls in `bin` (their queue
    // positions, appended by k_bin_hits), so that the warps of this launch all run the same BSDF model.
    // Nothing to do once the tail kernel has taken the batch's remaining paths over.
    const uint32_t n = counters->tail_taken ? 0u : (bin_list != nullptr ? counters->bin_count[which_in][bin] : counters->queue[which_in]);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    // Warp-local compaction of the live entries.  From the second bounce on most queue entries of an open scene are dead
    // (Dragon: 55 % of the rays escaped, and without an environment map an escaped ray has nothing left to do): a warp
    // that shaded its 32 consecutive entries would run the whole vertex code at half its lanes.  Instead it scans
    // chunks of 32 hit records, keeps the positions of the live ones in a small per-warp list and shades 32 of them at a
    // time — no global atomics (a per-ray atomic inside the traversal kernel, or a separate binning pass, cost more than
    // the dead lanes did: profiles/README.md), and entries that are all live pass straight through.
    __shared__ uint32_t pending_all[kShadeThreads / 32][64];
    uint32_t *pending = pending_all[warp];
    const bool dead_possible = !VOL && scene.integrator.id_envmap == kInvalid; // the shortcut of LoadPathVertex
    uint32_t count = 0;                                                         // live positions waiting in `pending` (warp-uniform)
    for (uint32_t i0 = tid - lane; i0 < n || count > 0;) {
        while (count < 32u && i0 < n) {
            const uint32_t k = i0 + lane;
            bool live = k < n;
            uint32_t pos = 0;
            if (live) {
                pos = bin_list != nullptr ? bin_list[k] : k;
                if (dead_possible) live = qin.hit[pos].prim != kPrimMiss;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, live);
            if (live) pending[count + __popc(ballot & ((1u << lane) - 1u))] = pos;
            count += __popc(ballot);
            i0 += stride;
        }
        __syncwarp();
        const uint32_t take = min(count, 32u);
        const bool active = lane < take;
        const uint32_t i = active ? pending[count - take + lane] : 0u;
        count -= take;
        __syncwarp();
        PathVertex v;
        bool alive = LoadPathVertex<VOL>(scene, qin, i, active, &v);
        PathNext next;
        V3 Ladd;
        const uint32_t slot = v.slot;
        alive = ShadeVertex<VOL, ONLY>(scene, bp, depth, alive, v, [&](const ShadowCandidate &sc) { PushShadow(sc, slot, sq, counters); }, &next, &Ladd);
        const uint32_t out = WarpAppend(alive, &counters->queue[which_in ^ 1]);
        if (alive) {
            qout.ox[out] = next.o.x, qout.oy[out] = next.o.y, qout.oz[out] = next.o.z;
            qout.dx[out] = next.d.x, qout.dy[out] = next.d.y, qout.dz[out] = next.d.z;
            qout.tr[out] = next.att.x, qout.tg[out] = next.att.y, qout.tb[out] = next.att.z;
            qout.pdf[out] = next.pdf;
            qout.slot[out] = slot;
            if (VOL) {
                qout.medium[out] = next.medium;
                qout.wx[out] = next.wo.x, qout.wy[out] = next.wo.y, qout.wz[out] = next.wo.z;
            }
        }
        if (active && (Ladd.x != 0.0f || Ladd.y != 0.0f || Ladd.z != 0.0f)) {
            RadianceAdd(radiance, slot, Ladd.x, Ladd.y, Ladd.z);
        }
    }
}

// Defined by the shade_variant.cu translation units: ONLY = kAnyBsdf (-1), 0 (no BSDF code: escaped rays, BSDF-less
// surfaces, area lights) and one per BSDF model (B200PT_BSDF_DIFFUSE .. B200PT_BSDF_PLASTIC).
#define B200PT_DECLARE_SHADE_VARIANT(NAME)                                                                              \
    void NAME(bool vol, int blocks, cudaStream_t stream, const DeviceScene &scene, const BatchParams &bp, uint32_t depth,  \
              PathQueue qin, int which_in, PathQueue qout, ShadowQueue sq, float *radiance, Counters *counters,            \
              uint32_t capacity, const uint32_t *bin_list, int bin)
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_any);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_0);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_2);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_3);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_4);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_5);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_6);
B200PT_DECLARE_SHADE_VARIANT(LaunchShadeVariant_7);

} // namespace b200pt
