// host_scene.hpp — host-side staging of DeviceScene: everything the reference's
// Renderer/Scene constructors derive from a RendererConfig (renderer.cpp:259-348,
// scene.cpp:118-533), rebuilt here for a GPU-friendly layout.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "b200pt.h"
#include "device_scene.h"

namespace b200pt {

// std::vector whose resize() leaves trivially-constructible elements uninitialised: the big per-triangle arrays are filled by
// parallel loops right after being sized, and a value-initialising resize would first write (and page-fault) every byte of
// them on ONE thread — 40 % of the time of the loops that fill them.
template <class T>
struct NoInitAllocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = NoInitAllocator<U>;
    };
    template <class U, class... Args>
    void construct(U *p, Args &&...args) {
        if constexpr (sizeof...(Args) == 0)
            ::new (static_cast<void *>(p)) U; // default-initialisation: nothing for trivial types
        else
            ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T>
using RawVector = std::vector<T, NoInitAllocator<T>>;

struct HostScene {
    std::vector<BvhNode> nodes;          // binary layout (default, and what the GPU LBVH builder emits) ...
    std::vector<WideNode> wide_nodes;    // ... or the compressed 8-wide layout (B200PT_CREATE_BVH8); never both
    uint32_t wide_depth = 0, wide_top_nodes = 0;
    RawVector<TriVerts> tri_verts;
    RawVector<TriShade> tri_shade;
    std::vector<uint8_t> tri_bsdf_type;
    std::vector<AnalyticPrim> analytic;
    std::vector<DInstance> instances;
    std::vector<DBsdf> bsdfs;
    std::vector<DTexture> textures;
    std::vector<DMedium> media;
    std::vector<DEmitter> emitters;
    std::vector<float> envmap_tables;
    float envmap_normalization = 0.0f;
    std::vector<float> kc_brdf_avg, kc_albedo_avg;
    std::vector<float> cdf_area_light;
    std::vector<uint32_t> map_area_light_instance;
    std::vector<float> light_tri_cdf;
    std::vector<uint32_t> light_tri_ids;
    float scene_bmin[3], scene_bmax[3];
    std::vector<float> cull_boxes;   // 6 floats per box (DeviceScene::cull_boxes): coarse cover, rejects whole screen tiles
    std::vector<float> fine_cull_boxes; // finer cover (<= 4096 boxes + analytic primitives): rejects single pixels
    std::vector<uint32_t> cull_fine_begin; // per coarse box (+ 1): its range of fine boxes
    DIntegrator integrator{};
    b200pt_camera camera{};
    double bvh_build_ms = 0.0;  // whole BVH stage (boxes, build, flatten)
    double bvh_gpu_ms = 0.0;    // GPU LBVH only: Morton codes + sort + hierarchy + refit
};

// Returns false and sets *error on an inconsistent description.
// gpu_lbvh: build the BVH with the GPU LBVH builder (bvh_gpu.cu) instead of the host binned-SAH builder (binary layout).
// bvh8: collapse the host SAH tree into the compressed 8-wide layout instead of flattening it into the binary one.
bool BuildHostScene(const b200pt_scene_desc &desc, uint32_t max_leaf_size, bool gpu_lbvh, bool bvh8, HostScene *out, std::string *error);

// camera.cpp:26-37 for an arbitrary output size (the CLI may override width/height, Q7).
DCamera MakeCamera(const b200pt_camera &cam, uint32_t width, uint32_t height);

// kulla_conty.cpp:62-80, multi-threaded restatement (bit-identical tables).
void ComputeKullaContyTables(float *brdf_avg, float *albedo_avg);

} // namespace b200pt
