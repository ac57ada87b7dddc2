// debug_eval.cu — test hook b200pt_debug_eval: the device leaf functions of shading.cuh at caller-supplied inputs.
//
// Compiled with -DB200PT_RNG_REPLAY, which swaps the Philox generator for the reference's LCG (vecmath.cuh), so the
// sampling routines consume the same numbers as csrt::Bsdf::Sample / Medium::Sample / Medium::SamplePhase started from the
// same seed and can be compared pointwise (tests/test_gpu_pointwise.py against ref_eval of oracle/ref_glue.cpp).  The
// functions are the very ones k_shade inlines; nothing here is on the render path.
#define B200PT_RNG_REPLAY 1
#include <cstdio>

#include <cuda_runtime.h>

#include "b200pt.h"
#include "shade_kernel.cuh"
#include "traverse_wide.cuh"

namespace b200pt {

namespace {

__device__ __forceinline__ V3 In3(const float *p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void Out3(float *p, V3 v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }

// in: B200PT_EVAL_IN floats per item, out: B200PT_EVAL_OUT floats per item (layout: include/b200pt.h)
__global__ void k_debug_eval(const __grid_constant__ DeviceScene scene, uint32_t what, uint32_t id, uint32_t n, const float *in_all, float *out_all) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *in = in_all + static_cast<uint64_t>(i) * B200PT_EVAL_IN;
    float *out = out_all + static_cast<uint64_t>(i) * B200PT_EVAL_OUT;
    for (int k = 0; k < B200PT_EVAL_OUT; ++k) out[k] = 0.0f;
    Rng rng(__float_as_uint(in[18]));
    switch (what) {
    case B200PT_EVAL_BSDF_EVALUATE:
    case B200PT_EVAL_BSDF_SAMPLE: {
        BsdfRec rec;
        rec.wi = In3(in), rec.wo = In3(in + 3), rec.n = In3(in + 6), rec.t = In3(in + 9), rec.b = In3(in + 12);
        rec.uv = {in[15], in[16]};
        rec.inside = in[17] != 0.0f;
        if (what == B200PT_EVAL_BSDF_EVALUATE)
            BsdfEvaluate<kAnyBsdf>(scene, scene.bsdfs[id], &rec);
        else
            BsdfSample<kAnyBsdf>(scene, scene.bsdfs[id], rng, &rec);
        out[0] = rec.valid, out[1] = rec.pdf;
        Out3(out + 2, rec.att), Out3(out + 5, rec.wi);
        break;
    }
    case B200PT_EVAL_EMITTER_SAMPLE: {
        const DEmitter &e = scene.emitters[id];
        const EmitterRec rec = EmitterSample(scene, e, In3(in), in[3], in[4]);
        out[0] = rec.valid, out[1] = rec.harsh, out[2] = rec.distance;
        Out3(out + 3, rec.wi);
        if (rec.valid) {
            Out3(out + 6, EmitterEvaluateRec(scene, e, rec));
            out[9] = EmitterPdf(scene, e, -rec.wi);
        }
        break;
    }
    case B200PT_EVAL_EMITTER_DIR: {
        const DEmitter &e = scene.emitters[id];
        Out3(out, EmitterEvaluateDir(scene, e, In3(in)));
        out[3] = EmitterPdf(scene, e, In3(in));
        break;
    }
    case B200PT_EVAL_MEDIUM_SAMPLE:
    case B200PT_EVAL_MEDIUM_EVALUATE: {
        MediumRec rec;
        if (what == B200PT_EVAL_MEDIUM_SAMPLE) {
            MediumSample(scene.media[id], in[0], rng, &rec);
        } else {
            rec.distance = in[0];
            MediumEvaluate(scene.media[id], &rec);
        }
        out[0] = rec.valid, out[1] = rec.scattered, out[2] = rec.pdf, out[3] = rec.distance;
        Out3(out + 4, rec.att);
        break;
    }
    case B200PT_EVAL_PHASE_SAMPLE:
    case B200PT_EVAL_PHASE_EVALUATE: {
        PhaseRec rec;
        rec.wi = In3(in), rec.wo = In3(in + 3);
        if (what == B200PT_EVAL_PHASE_SAMPLE)
            PhaseSample(scene.media[id], rng, &rec);
        else
            PhaseEvaluate(scene.media[id], &rec);
        out[0] = rec.valid, out[1] = rec.pdf;
        Out3(out + 2, rec.att), Out3(out + 5, rec.wi);
        break;
    }
    case B200PT_EVAL_TEXTURE: {
        Out3(out, TexColor(scene, id, {in[0], in[1]}));
        break;
    }
    case B200PT_EVAL_SURFACE: { // hit attributes of a closest-hit record as the shading stage rebuilds them
        const uint32_t prim = __float_as_uint(in[0]);
        Ray ray;
        ray.o = In3(in + 3), ray.d = In3(in + 6), ray.tmin = kEpsilonDistance, ray.tmax = in[9];
        Surf s;
        if (prim & kPrimAnalyticBit)
            s = SurfAnalytic(scene, prim & ~kPrimAnalyticBit, ray, in[9]);
        else
            s = SurfTriangle(scene, prim & kPrimIndexMask, in[1], in[2], (prim & kPrimInsideBit) != 0);
        Out3(out, s.pos), Out3(out + 3, s.n), Out3(out + 6, s.t), Out3(out + 9, s.b);
        out[12] = s.uv.u, out[13] = s.uv.v, out[14] = s.inside, out[15] = __uint_as_float(s.inst);
        break;
    }
    default: break;
    }
    out[B200PT_EVAL_OUT - 1] = what == B200PT_EVAL_SURFACE ? out[B200PT_EVAL_OUT - 1] : __uint_as_float(rng.state);
}

// ---------------------------------------------------------------------------------------------
// k_debug_replay (b200pt_debug_render_replay): the reference's own loop shape and random-number stream — one thread per
// pixel, the samples of the pixel one after the other, ONE LCG per pixel seeded with Tea<4>(pixel_offset, 0) that runs on
// through every sample and every vertex (Renderer::DrawPixel renderer.cpp:62-85, RandomFloat math.hpp:57-63) — around the
// product's ShadeVertex and per-lane traversal.  Every sample then takes the same decisions as the same sample of the
// reference's --cpu run (up to the last-bit differences between CUDA's and glibc's sinf/cosf/expf/powf, which flip a
// decision once in ~10^5), so frames compare PER PIXEL far below the Monte Carlo noise
// (tests/test_gpu_replay.py).  Alpha-tested scenes are refused: the reference draws the numbers of its opacity tests
// inside its own BVH walk, whose visiting order the product's tree does not share.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t Tea4(uint32_t v0, uint32_t v1) { // math.hpp:43-54
    uint32_t s0 = 0;
    for (int n = 0; n < 4; ++n) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

__device__ __forceinline__ bool ReplayTrace(const DeviceScene &scene, const Ray &ray, bool any, HitRec *hit) {
    TraversalCounters unused;
    if (scene.num_wide_nodes > 0) return TraverseSingleWide(scene, ray, any, false, Rng(0u), hit, false, &unused);
    return TraverseSingle(scene, ray, any, false, Rng(0u), hit, false, &unused);
}

template <bool VOL>
__global__ void __launch_bounds__(64) k_debug_replay(const __grid_constant__ DeviceScene scene, const __grid_constant__ BatchParams bp, float *frame,
                                                             uint32_t trace_pixel) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inside = pixel < bp.width * bp.height;
    const uint32_t i = inside ? pixel % bp.width : 0u, j = inside ? pixel / bp.width : 0u;
    const DIntegrator &ig = scene.integrator;
    uint32_t seed = Tea4(pixel * 3u, 0u);
    V3 color = mk3(0.0f);
    const bool trace = pixel == trace_pixel; // B200PT_REPLAY_TRACE="i,j": the vertices of one pixel's paths, as ORACLE_TRACE_PIXEL prints them
    for (uint32_t k = 0; k < bp.spp; ++k) {
        if (trace) printf("[trace] sample %u\n", k);
        // renderer.cpp:66-75
        const float u = k * bp.spp_inv, v = VanDerCorput2(k + 1);
        const float x = 2.0f * (i + u) / static_cast<int>(bp.width) - 1.0f, y = 1.0f - 2.0f * (j + v) / static_cast<int>(bp.height);
        PathVertex pv;
        pv.ray.o = mk3(bp.camera.eye);
        pv.ray.d = Normalize(mk3(bp.camera.front) + x * mk3(bp.camera.view_dx) + y * mk3(bp.camera.view_dy));
        pv.ray.tmin = kEpsilonDistance, pv.ray.tmax = kMaxFloat;
        pv.att = mk3(1.0f), pv.wo_prev = -pv.ray.d, pv.pdf_sample = 0.0f;
        pv.slot = 0, pv.ray_medium = kInvalid;
        pv.replay_state = seed;
        V3 L = mk3(0.0f);
        bool alive = inside && ReplayTrace(scene, pv.ray, false, &pv.hit);
        if (inside && !alive) { // path.cpp:24-35
            if (ig.id_envmap != kInvalid) L += EmitterEvaluateDir(scene, scene.emitters[ig.id_envmap], pv.ray.d);
            if (ig.id_sun != kInvalid) L += EmitterEvaluateDir(scene, scene.emitters[ig.id_sun], pv.ray.d);
        }
        // the same loop as k_tail (tail_kernel.cu): every lane of the warp walks its path until all are done
        for (uint32_t depth = 1; __any_sync(0xffffffffu, alive); ++depth) {
            PathNext next;
            V3 Ladd;
            const bool was_alive = alive;
            if (trace && alive)
                printf("[trace] v%u state %08x t %.9g L %.9g %.9g %.9g att %.9g %.9g %.9g\n", depth, pv.replay_state, pv.hit.t, L.x, L.y, L.z, pv.att.x,
                       pv.att.y, pv.att.z);
            alive = ShadeVertex<VOL, kAnyBsdf>(
                scene, bp, depth, alive, pv,
                [&](const ShadowCandidate &sc) {
                    if (!(sc.valid && (sc.c.x != 0.0f || sc.c.y != 0.0f || sc.c.z != 0.0f))) return;
                    Ray ray;
                    ray.o = sc.o, ray.d = sc.d, ray.tmin = kEpsilonDistance, ray.tmax = sc.tmax;
                    HitRec unused;
                    if (!ReplayTrace(scene, ray, true, &unused)) L += sc.c;
                },
                &next, &Ladd);
            if (was_alive) {
                L += Ladd;
                pv.replay_state = next.replay_state;
            }
            if (depth >= kMaxTailDepth) alive = false;
            if (alive) {
                pv.ray.o = next.o, pv.ray.d = next.d, pv.ray.tmin = kEpsilonDistance, pv.ray.tmax = kMaxFloat;
                pv.att = next.att, pv.pdf_sample = next.pdf, pv.ray_medium = next.medium, pv.wo_prev = next.wo;
                ReplayTrace(scene, pv.ray, false, &pv.hit);
                if (!VOL && pv.hit.prim == kPrimMiss && ig.id_envmap == kInvalid) alive = false;
            }
        }
        seed = pv.replay_state;
        color += mk3(fminf(L.x, 1.0f), fminf(L.y, 1.0f), fminf(L.z, 1.0f)); // renderer.cpp:76-80 (Q2)
    }
    color = color * bp.spp_inv;
    if (inside) frame[3 * pixel] = color.x, frame[3 * pixel + 1] = color.y, frame[3 * pixel + 2] = color.z;
}

} // namespace

void LaunchDebugReplay(cudaStream_t stream, const DeviceScene &scene, const BatchParams &bp, float *frame, uint32_t trace_pixel) {
    const uint32_t pixels = bp.width * bp.height;
    if (scene.integrator.type == B200PT_INTEGRATOR_VOLPATH)
        k_debug_replay<true><<<(pixels + 63) / 64, 64, 0, stream>>>(scene, bp, frame, trace_pixel);
    else
        k_debug_replay<false><<<(pixels + 63) / 64, 64, 0, stream>>>(scene, bp, frame, trace_pixel);
}

void LaunchDebugEval(cudaStream_t stream, const DeviceScene &scene, uint32_t what, uint32_t id, uint32_t n, const float *in, float *out) {
    k_debug_eval<<<(n + 127) / 128, 128, 0, stream>>>(scene, what, id, n, in, out);
}

} // namespace b200pt
