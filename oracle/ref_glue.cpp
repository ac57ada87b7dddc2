// ref_glue.cpp — TEST INFRASTRUCTURE ONLY.  C entry points into the UNMODIFIED reference.
//
// Compiled by oracle/Makefile together with the reference's own sources (where they lie under
// /root/reference) into oracle/_ref/libcsrt_ref_{mt,woop}.so.  Nothing in the product links this.
// It is the checker the parity tests and bench.py's cpu_baseline / --impl reference arm call:
//   ref_pack_from_xml : csrt::LoadConfig (src/parser/parser.cpp:94) + CLI overrides
//                       (apps/main.cpp:46-52) -> scene pack on disk
//   ref_render        : csrt::Renderer(config).Draw(frame) on the CPU backend
//                       (src/renderer/renderer.cpp:259, 678; DispathRaysCpu :142-253)
//   ref_* KAT helpers : direct calls of reference leaf functions for known-answer tests.
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <exception>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

// The pointwise checks (ref_eval) call Bsdf / Emitter / Medium / Texture objects exactly as csrt::Renderer built them; the
// renderer keeps them private, so THIS translation unit (test infrastructure, nothing else) reads the class with its
// members public.  The reference sources themselves are compiled unmodified; the class layout does not change.
#define private public
#include "csrt/renderer/renderer.hpp"
#undef private
#include "csrt/parser/parser.hpp"
#include "csrt/renderer/bsdfs/kulla_conty.hpp"
#include "csrt/rtcore/accel/bvh_builder.hpp"
#include "csrt/rtcore/scene.hpp"

#include "b200pt.h"
#include "csrc/host_util.hpp"
#include "host/csrt_glue.hpp"

namespace {

std::string g_error;

// The reference prints a progress line per 64-pixel patch to stderr (renderer.cpp:235-238).
class StderrSilencer {
public:
    StderrSilencer() {
        if (getenv("B200PT_REF_VERBOSE")) return;
        fflush(stderr);
        saved_ = dup(2);
        const int null_fd = open("/dev/null", O_WRONLY);
        if (null_fd >= 0) {
            dup2(null_fd, 2);
            close(null_fd);
        }
    }
    ~StderrSilencer() {
        if (saved_ >= 0) {
            fflush(stderr);
            dup2(saved_, 2);
            close(saved_);
        }
    }

private:
    int saved_ = -1;
};

double Seconds(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
}

} // namespace

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }

// 1 if this build uses Woop's watertight triangle test (src/rtcore/primitives/triangle.cpp:23-87).
int ref_is_watertight() {
#ifdef WATERTIGHT_TRIANGLES
    return 1;
#else
    return 0;
#endif
}

int ref_pack_from_xml(const char *xml_path, int width, int height, int spp, const char *out_pack) {
    try {
        csrt::RendererConfig cfg;
        {
            StderrSilencer quiet;
            cfg = csrt::LoadConfig(xml_path);
        }
        cfg.backend_type = csrt::BackendType::kCpu;
        // apps/main.cpp:46-52 — overrides are applied after fov_x was derived from the XML size (Q7).
        if (width > 0) cfg.camera.width = width;
        if (height > 0) cfg.camera.height = height;
        if (spp > 0) cfg.camera.spp = spp;
        std::unique_ptr<b200pt_scene> scene = b200pt_glue::FlattenConfig(cfg);
        if (b200pt_scene_save(&scene->desc, out_pack) != B200PT_OK) {
            g_error = b200pt::GlobalError();
            return -1;
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

int ref_render(const b200pt_scene_desc *desc, int width, int height, int spp, float *frame, double *build_seconds,
               double *render_seconds) {
    try {
        csrt::RendererConfig cfg = b200pt_glue::InflateScene(*desc);
        if (width > 0) cfg.camera.width = width;
        if (height > 0) cfg.camera.height = height;
        if (spp > 0) cfg.camera.spp = spp;
        StderrSilencer quiet;
        const auto t0 = std::chrono::steady_clock::now();
        csrt::Renderer renderer(cfg);
        const auto t1 = std::chrono::steady_clock::now();
        renderer.Draw(frame);
        const auto t2 = std::chrono::steady_clock::now();
        if (build_seconds) *build_seconds = Seconds(t0, t1);
        if (render_seconds) *render_seconds = Seconds(t1, t2);
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Persistent variant for timing: build once (scene commit + LBVH + LUTs), Draw many times.
void *ref_create(const b200pt_scene_desc *desc, int width, int height, int spp, double *build_seconds) {
    try {
        csrt::RendererConfig cfg = b200pt_glue::InflateScene(*desc);
        if (width > 0) cfg.camera.width = width;
        if (height > 0) cfg.camera.height = height;
        if (spp > 0) cfg.camera.spp = spp;
        StderrSilencer quiet;
        const auto t0 = std::chrono::steady_clock::now();
        csrt::Renderer *renderer = new csrt::Renderer(cfg);
        if (build_seconds) *build_seconds = Seconds(t0, std::chrono::steady_clock::now());
        return renderer;
    } catch (const std::exception &e) {
        g_error = e.what();
        return nullptr;
    }
}

// One csrt::Renderer::Draw (renderer.cpp:678): all hardware threads, returns wall seconds or -1.
double ref_draw(void *renderer, float *frame) {
    try {
        StderrSilencer quiet;
        const auto t0 = std::chrono::steady_clock::now();
        static_cast<csrt::Renderer *>(renderer)->Draw(frame);
        return Seconds(t0, std::chrono::steady_clock::now());
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1.0;
    }
}

void ref_destroy(void *renderer) { delete static_cast<csrt::Renderer *>(renderer); }

// ---- the reference's own CUDA backend (only in libcsrt_ref_cuda.so, built with -DENABLE_CUDA): BackendType::kCuda, i.e. its
// one-thread-per-pixel megakernel DispathRaysCuda (renderer.cpp:88-95) over managed memory, as `RayTracer --gpu` runs it. ----
int ref_has_cuda() {
#ifdef ENABLE_CUDA
    return 1;
#else
    return 0;
#endif
}

#ifdef ENABLE_CUDA
namespace {
struct CudaRef {
    csrt::Renderer *renderer = nullptr;
    float *frame = nullptr; // managed, as RayTracer::RayTracer allocates it (ray_tracer.cpp:133-136)
    size_t count = 0;
};
} // namespace

void *ref_create_cuda(const b200pt_scene_desc *desc, int width, int height, int spp, double *build_seconds) {
    try {
        csrt::RendererConfig cfg = b200pt_glue::InflateScene(*desc);
        cfg.backend_type = csrt::BackendType::kCuda;
        if (width > 0) cfg.camera.width = width;
        if (height > 0) cfg.camera.height = height;
        if (spp > 0) cfg.camera.spp = spp;
        StderrSilencer quiet;
        const auto t0 = std::chrono::steady_clock::now();
        CudaRef *r = new CudaRef();
        r->renderer = new csrt::Renderer(cfg);
        r->count = static_cast<size_t>(cfg.camera.width) * cfg.camera.height * 3;
        r->frame = csrt::MallocArray<float>(csrt::BackendType::kCuda, r->count);
        if (build_seconds) *build_seconds = Seconds(t0, std::chrono::steady_clock::now());
        return r;
    } catch (const std::exception &e) {
        g_error = e.what();
        return nullptr;
    }
}

// One csrt::Renderer::Draw on the CUDA backend (kernel launch + cudaDeviceSynchronize, renderer.cpp:692-711); returns its
// wall seconds or -1; the frame is then copied out of managed memory into frame_host (not timed).
double ref_draw_cuda(void *handle, float *frame_host) {
    try {
        CudaRef *r = static_cast<CudaRef *>(handle);
        StderrSilencer quiet;
        const auto t0 = std::chrono::steady_clock::now();
        r->renderer->Draw(r->frame);
        const double seconds = Seconds(t0, std::chrono::steady_clock::now());
        if (frame_host) memcpy(frame_host, r->frame, r->count * sizeof(float));
        return seconds;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1.0;
    }
}

// The preview path as the reference runs it: csrt::Renderer::Draw(index_frame, frame, frame_srgb) (renderer.cpp:97-138,
// 723-746), one call per displayed frame, running mean in managed memory.  Returns 0, or -1 with ref_last_error().
int ref_draw_progressive_cuda(void *handle, uint32_t num_frames, float *frame_host, float *frame_srgb_host) {
    try {
        CudaRef *r = static_cast<CudaRef *>(handle);
        float *srgb = csrt::MallocArray<float>(csrt::BackendType::kCuda, r->count);
        memset(r->frame, 0, r->count * sizeof(float));
        StderrSilencer quiet;
        for (uint32_t k = 0; k < num_frames; ++k) r->renderer->Draw(k, r->frame, srgb);
        if (frame_host) memcpy(frame_host, r->frame, r->count * sizeof(float));
        if (frame_srgb_host) memcpy(frame_srgb_host, srgb, r->count * sizeof(float));
        csrt::DeleteArray(csrt::BackendType::kCuda, srgb);
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

void ref_destroy_cuda(void *handle) {
    CudaRef *r = static_cast<CudaRef *>(handle);
    if (!r) return;
    csrt::DeleteArray(csrt::BackendType::kCuda, r->frame);
    delete r->renderer;
    delete r;
}
#endif

// ---- pointwise traversal: csrt::Scene (scene.cpp:118-141) + TLAS::Intersect / IntersectAny (tlas.cpp:13-76) on caller-supplied
// rays, without BSDFs (no opacity masks).  rays: 8 floats each (origin, direction, t_min, t_max). ----
namespace {
struct RefScene {
    std::unique_ptr<csrt::Scene> scene;
    std::vector<uint32_t> map_instance_bsdf;
};
} // namespace

struct ref_hit {
    float t;             // ray.t_max after the call
    uint32_t valid, inside, id_instance, id_primitive;
    float position[3], normal[3], texcoord[2], tangent[3], bitangent[3];
};

void *ref_scene_create(const b200pt_scene_desc *desc) {
    try {
        csrt::RendererConfig cfg = b200pt_glue::InflateScene(*desc);
        StderrSilencer quiet;
        RefScene *s = new RefScene();
        s->scene.reset(new csrt::Scene(csrt::BackendType::kCpu, cfg.instances));
        s->map_instance_bsdf.assign(cfg.instances.size() + 1, csrt::kInvalidId);
        return s;
    } catch (const std::exception &e) {
        g_error = e.what();
        return nullptr;
    }
}

void ref_scene_destroy(void *handle) { delete static_cast<RefScene *>(handle); }

int ref_trace(void *handle, uint64_t n, const float *rays, int any_hit, ref_hit *out) {
    try {
        RefScene *s = static_cast<RefScene *>(handle);
        const csrt::TLAS *tlas = s->scene->GetTlas();
        for (uint64_t i = 0; i < n; ++i) {
            const float *r = rays + 8 * i;
            csrt::Ray ray(csrt::Vec3(r[0], r[1], r[2]), csrt::Vec3(r[3], r[4], r[5])); // ray.cpp:18-47: dir_rcp, Woop k / shear
            ray.t_min = r[6], ray.t_max = r[7];
            uint32_t seed = 0;
            ref_hit h{};
            if (any_hit) {
                h.valid = tlas->IntersectAny(nullptr, s->map_instance_bsdf.data(), &seed, &ray) ? 1u : 0u;
                h.t = ray.t_max;
            } else {
                const csrt::Hit hit = tlas->Intersect(nullptr, s->map_instance_bsdf.data(), &seed, &ray);
                h.t = ray.t_max;
                h.valid = hit.valid, h.inside = hit.inside, h.id_instance = hit.id_instance, h.id_primitive = hit.id_primitve;
                h.position[0] = hit.position.x, h.position[1] = hit.position.y, h.position[2] = hit.position.z;
                h.normal[0] = hit.normal.x, h.normal[1] = hit.normal.y, h.normal[2] = hit.normal.z;
                h.texcoord[0] = hit.texcoord.u, h.texcoord[1] = hit.texcoord.v;
                h.tangent[0] = hit.tangent.x, h.tangent[1] = hit.tangent.y, h.tangent[2] = hit.tangent.z;
                h.bitangent[0] = hit.bitangent.x, h.bitangent[1] = hit.bitangent.y, h.bitangent[2] = hit.bitangent.z;
            }
            out[i] = h;
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// ---- pointwise leaf functions: the Bsdf / Emitter / Medium / Texture objects of a csrt::Renderer (as ref_create builds it)
// at caller-supplied inputs.  Same record layout as b200pt_debug_eval (include/b200pt.h): n x 32 floats in, n x 16 floats out. ----
int ref_eval(void *renderer_handle, uint32_t what, uint32_t id, uint64_t n, const float *in_all, float *out_all) {
    try {
        const csrt::Renderer *r = static_cast<const csrt::Renderer *>(renderer_handle);
        auto in3 = [](const float *p) { return csrt::Vec3(p[0], p[1], p[2]); };
        auto out3 = [](float *p, const csrt::Vec3 &v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; };
        for (uint64_t i = 0; i < n; ++i) {
            const float *in = in_all + i * B200PT_EVAL_IN;
            float *out = out_all + i * B200PT_EVAL_OUT;
            memset(out, 0, sizeof(float) * B200PT_EVAL_OUT);
            uint32_t seed;
            memcpy(&seed, in + 18, 4);
            switch (what) {
            case B200PT_EVAL_BSDF_EVALUATE:
            case B200PT_EVAL_BSDF_SAMPLE: {
                csrt::BsdfSampleRec rec;
                rec.wi = in3(in), rec.wo = in3(in + 3), rec.normal = in3(in + 6), rec.tangent = in3(in + 9), rec.bitangent = in3(in + 12);
                rec.texcoord = csrt::Vec2(in[15], in[16]);
                rec.inside = in[17] != 0.0f;
                if (what == B200PT_EVAL_BSDF_EVALUATE)
                    r->bsdfs_[id].Evaluate(&rec);
                else
                    r->bsdfs_[id].Sample(&seed, &rec);
                out[0] = rec.valid, out[1] = rec.pdf;
                out3(out + 2, rec.attenuation), out3(out + 5, rec.wi);
                break;
            }
            case B200PT_EVAL_EMITTER_SAMPLE: {
                const csrt::Emitter &e = r->emitters_[id];
                const csrt::EmitterSampleRec rec = e.Sample(in3(in), in[3], in[4]);
                out[0] = rec.valid, out[1] = rec.harsh, out[2] = rec.distance;
                out3(out + 3, rec.wi);
                if (rec.valid) {
                    out3(out + 6, e.Evaluate(rec));
                    out[9] = e.Pdf(-rec.wi);
                }
                break;
            }
            case B200PT_EVAL_EMITTER_DIR: {
                const csrt::Emitter &e = r->emitters_[id];
                out3(out, e.Evaluate(in3(in)));
                out[3] = e.Pdf(in3(in));
                break;
            }
            case B200PT_EVAL_MEDIUM_SAMPLE:
            case B200PT_EVAL_MEDIUM_EVALUATE: {
                csrt::MediumSampleRec rec;
                if (what == B200PT_EVAL_MEDIUM_SAMPLE) {
                    r->media_[id].Sample(in[0], &seed, &rec);
                } else {
                    rec.distance = in[0];
                    r->media_[id].Evaluate(&rec);
                }
                out[0] = rec.valid, out[1] = rec.scattered, out[2] = rec.pdf, out[3] = rec.distance;
                out3(out + 4, rec.attenuation);
                break;
            }
            case B200PT_EVAL_PHASE_SAMPLE:
            case B200PT_EVAL_PHASE_EVALUATE: {
                csrt::PhaseSampleRec rec;
                rec.wi = in3(in), rec.wo = in3(in + 3);
                if (what == B200PT_EVAL_PHASE_SAMPLE)
                    r->media_[id].SamplePhase(&seed, &rec);
                else
                    r->media_[id].EvaluatePhase(&rec);
                out[0] = rec.valid, out[1] = rec.pdf;
                out3(out + 2, rec.attenuation), out3(out + 5, rec.wi);
                break;
            }
            case B200PT_EVAL_TEXTURE: {
                out3(out, r->textures_[id].GetColor(csrt::Vec2(in[0], in[1])));
                break;
            }
            default:
                g_error = "ref_eval: unknown function";
                return -1;
            }
            memcpy(out + B200PT_EVAL_OUT - 1, &seed, 4);
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Instance ids with their BSDF, through the renderer's own scene: closest hit with the hit frame as Primitive::Intersect
// builds it (bump mapping included), for B200PT_EVAL_SURFACE.
int ref_trace_renderer(void *renderer_handle, uint64_t n, const float *rays, ref_hit *out) {
    try {
        const csrt::Renderer *r = static_cast<const csrt::Renderer *>(renderer_handle);
        const csrt::TLAS *tlas = r->scene_->GetTlas();
        for (uint64_t i = 0; i < n; ++i) {
            const float *q = rays + 8 * i;
            csrt::Ray ray(csrt::Vec3(q[0], q[1], q[2]), csrt::Vec3(q[3], q[4], q[5]));
            ray.t_min = q[6], ray.t_max = q[7];
            uint32_t seed = 0;
            const csrt::Hit hit = tlas->Intersect(r->bsdfs_, r->map_instance_bsdf_, &seed, &ray);
            ref_hit h{};
            h.t = ray.t_max;
            h.valid = hit.valid, h.inside = hit.inside, h.id_instance = hit.id_instance, h.id_primitive = hit.id_primitve;
            h.position[0] = hit.position.x, h.position[1] = hit.position.y, h.position[2] = hit.position.z;
            h.normal[0] = hit.normal.x, h.normal[1] = hit.normal.y, h.normal[2] = hit.normal.z;
            h.texcoord[0] = hit.texcoord.u, h.texcoord[1] = hit.texcoord.v;
            h.tangent[0] = hit.tangent.x, h.tangent[1] = hit.tangent.y, h.tangent[2] = hit.tangent.z;
            h.bitangent[0] = hit.bitangent.x, h.bitangent[1] = hit.bitangent.y, h.bitangent[2] = hit.bitangent.z;
            out[i] = h;
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// ---- known-answer helpers: reference leaf functions, called directly ----

uint32_t ref_tea4(uint32_t v0, uint32_t v1) { return csrt::Tea<4>(v0, v1); }

float ref_random_float(uint32_t *seed) { return csrt::RandomFloat(seed); }

float ref_van_der_corput2(uint32_t index) { return csrt::GetVanDerCorputSequence<2>(index); }
float ref_van_der_corput3(uint32_t index) { return csrt::GetVanDerCorputSequence<3>(index); }

float ref_mis_weight(float a, float b) { return csrt::MisWeight(a, b); }

void ref_sample_hemis_cos(float xi0, float xi1, float *vec3, float *pdf) {
    csrt::Vec3 v;
    csrt::SampleHemisCos(xi0, xi1, &v, pdf);
    vec3[0] = v.x, vec3[1] = v.y, vec3[2] = v.z;
}

void ref_kulla_conty(float *brdf_avg, float *albedo_avg) { csrt::ComputeKullaConty(brdf_avg, albedo_avg); }

// LBVH topology for `n` boxes: out_nodes[i] = {leaf, id_left, id_right, id_object}; returns node count.
uint32_t ref_build_bvh(uint32_t n, const float *aabb_min_max, const float *areas, uint32_t *out_nodes,
                       float *out_area, uint32_t capacity) {
    std::vector<csrt::AABB> aabbs(n);
    std::vector<float> area_list(areas, areas + n);
    for (uint32_t i = 0; i < n; ++i) {
        const float *b = aabb_min_max + 6 * i;
        aabbs[i] = csrt::AABB(csrt::Vec3(b[0], b[1], b[2]), csrt::Vec3(b[3], b[4], b[5]));
    }
    const std::vector<csrt::BvhNode> nodes = csrt::BvhBuilder::Build(aabbs, area_list);
    for (uint32_t i = 0; i < nodes.size() && i < capacity; ++i) {
        out_nodes[4 * i + 0] = nodes[i].leaf ? 1u : 0u;
        out_nodes[4 * i + 1] = nodes[i].id_left;
        out_nodes[4 * i + 2] = nodes[i].id_right;
        out_nodes[4 * i + 3] = nodes[i].id_object;
        out_area[i] = nodes[i].area;
    }
    return static_cast<uint32_t>(nodes.size());
}

} // extern "C"
