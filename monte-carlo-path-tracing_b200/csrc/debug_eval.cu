// debug_eval.cu — test hook b200pt_debug_eval: the device leaf functions of shading.cuh at caller-supplied inputs.
//
// Compiled with -DB200PT_RNG_REPLAY, which swaps the Philox generator for the reference's LCG (vecmath.cuh), so the
// sampling routines consume the same numbers as csrt::Bsdf::Sample / Medium::Sample / Medium::SamplePhase started from the
// same seed and can be compared pointwise (tests/test_gpu_pointwise.py against ref_eval of oracle/ref_glue.cpp).  The
// functions are the very ones k_shade inlines; nothing here is on the render path.
#define B200PT_RNG_REPLAY 1
#include <cuda_runtime.h>

#include "b200pt.h"
#include "shading.cuh"

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

} // namespace

void LaunchDebugEval(cudaStream_t stream, const DeviceScene &scene, uint32_t what, uint32_t id, uint32_t n, const float *in, float *out) {
    k_debug_eval<<<(n + 127) / 128, 128, 0, stream>>>(scene, what, id, n, in, out);
}

} // namespace b200pt
