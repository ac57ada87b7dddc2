// device_scene.h — the scene as the kernels see it: flat, POD, HBM-resident arrays.
//
// Built on the host by scene_build.cpp from a b200pt_scene_desc (the flattened
// csrt::RendererConfig) and uploaded once.  Shared by host C++ and CUDA code.
//
// Data layout in HBM:
//   nodes      BVH2, 64 B per node = 4 x 16 B: the boxes of BOTH children + their links,
//              so one node fetch decides both subtrees.  Breadth-first order: the first
//              kTopNodes nodes (top of the tree) are also staged into shared memory.
//   tri_verts  48 B per triangle (3 x float4: world-space positions, .w = owning instance
//              / padding) in leaf order — the only triangle data traversal touches.
//   tri_shade  112 B per triangle (normals, tangents, uvs, instance id) read once per hit.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200PT_HD __host__ __device__
#else
#define B200PT_HD
#endif

namespace b200pt {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr uint32_t kPrimMiss = 0xFFFFFFFFu;
constexpr uint32_t kPrimAnalyticBit = 0x80000000u;
constexpr uint32_t kPrimInsideBit = 0x40000000u;  // triangle hit from its back side (det_inv < 0, triangle.cpp:117)
constexpr uint32_t kPrimIndexMask = 0x0FFFFFFFu;
constexpr int kLutResolution = 128;     // kulla_conty.hpp:10

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };

// 3x4 affine matrix (rows of a csrt::Mat4 whose last row is 0 0 0 1), row-major.
struct Affine { float m[12]; };

struct BvhNode {            // 64 B
    F4 c0xy;                // child0: min.x max.x min.y max.y
    F4 c1xy;                // child1: min.x max.x min.y max.y
    F4 cz;                  // c0.min.z c0.max.z c1.min.z c1.max.z
    int32_t child0, child1; // >=0 inner node index; <0 leaf: ~v = (first_tri << 3) | (count-1)
    int32_t pad0, pad1;
};

struct TriVerts {           // 48 B
    F4 v0, v1, v2;          // .w of v0 = bit pattern of instance id; .w of v1 = bit pattern of the triangle's index in
                            // the scene description (instances in order, triangles in order: the debug ray entry reports it)
};

// Compressed 8-wide BVH node (Ylitie, Karras, Laine 2017), 80 B = 5 x 16 B loads.  Child boxes are quantised to 8 bits per
// plane on the grid origin + q * 2^e of the node's own box (rounded outwards, so a quantised box contains the exact one).
// Children sit in octant slots: slot s holds the child whose centre lies furthest in direction (s&1 ? +x : -x, s&2 ? +y : -y,
// s&4 ? +z : -z), so a ray visits slots in descending order of (s ^ r), r = bits of the ray's positive direction components.
//   meta[s] = 0                              empty slot
//           = 0b001'00000 | (24 + s)         inner node: the k-th set bit of imask below s is child_base + k
//           = unary(count) << 5 | offset     leaf of 1..3 triangles tri_base + offset .. (offset < 24)
struct WideNode {
    float origin[3];
    uint8_t e[3];           // biased exponents of the grid spacing per axis: 2^(e - 127)
    uint8_t imask;          // slots that hold inner nodes
    uint32_t child_base;    // first inner child (children of one node are contiguous, in slot order)
    uint32_t tri_base;      // first triangle of this node's leaves
    uint8_t meta[8];
    uint8_t qlo_x[8], qlo_y[8], qlo_z[8], qhi_x[8], qhi_y[8], qhi_z[8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

struct TriShade {           // 112 B
    float n[3][3];          // per-vertex shading normals (world, unnormalised as in scene.cpp:261-281)
    float t[3][3];          // per-vertex tangents (after SetupMeshes orthonormalisation, scene.cpp:63-110)
    float uv[3][2];
    uint32_t inst;
    uint32_t pad[3];
};

enum AnalyticType : uint32_t { kSphere = 0, kDisk = 1, kCylinder = 2 };

struct AnalyticPrim {
    uint32_t type, inst;
    float radius, length;        // sphere radius (local), cylinder radius/length (world, scene.cpp:418-437)
    F3 center;                   // sphere centre (local)
    float bmin[3], bmax[3];      // world AABB (sphere.cpp:9-15, disk.cpp:9-15, cylinder.cpp:9-19)
    Affine to_world, to_local;   // to_local = to_world.Inverse() (recomputed per test by the reference)
    Affine normal_to_world;      // to_world.Transpose().Inverse(), 3x3 part used
};

struct DTexture {
    uint32_t type;
    int32_t width, height, channels;
    F3 color0, color1;
    Affine to_uv;
    uint64_t pixel_offset;
};

struct DBsdf {
    uint32_t type, twosided;
    uint32_t id_opacity, id_bump_map;
    uint32_t id_radiance, id_diffuse_reflectance, id_roughness_u, id_roughness_v;
    uint32_t id_specular_reflectance, id_specular_transmittance;
    uint32_t use_fast_approx, pad;
    float eta, eta_inv;          // dielectric (bsdf.cpp:176-177)
    float reflectivity_s;        // dielectric / plastic scalar reflectivity (bsdf.cpp:178-179, 188-189)
    float F_avg_s, F_avg_inv_s;  // dielectric / plastic average Fresnel (bsdf.cpp:160-161, 190)
    F3 reflectivity, edgetint, F_avg; // conductor (bsdf.cpp:151-155)
};

struct DMedium {                 // medium.cpp:6-39
    F3 sigma_s, sigma_t;
    float sampling_weight;
    uint32_t phase_type;
    F3 g;
};

struct DInstance {
    uint32_t id_bsdf, id_medium_int, id_medium_ext;
    uint32_t area_light;         // index into the area-light list or kInvalid (renderer.cpp:293-304)
    float pdf_area;              // 1 / (sum of primitive "areas"), scene.cpp:493-495 (Q3)
    uint32_t analytic;           // index into analytic[] for sphere/disk/cylinder instances, else kInvalid
    uint32_t light_tri_begin, light_tri_count; // range in light_tri_* for mesh instances that are area lights
};

struct DEmitter {
    uint32_t type, id_texture;
    F3 position, direction, radiance;
    float cutoff_angle, cos_cutoff_angle, uv_factor, beam_width, cos_beam_width, transition_width_rcp;
    Affine to_world, to_local;
    // env map (emitter.cpp:166-175 wiring, Q9): offsets into envmap_tables
    int32_t env_width, env_height;
    float env_normalization;
    uint32_t env_cdf_cols, env_cdf_rows, env_weight_rows;
};

struct DCamera {                 // camera.cpp:26-37
    F3 eye, front, view_dx, view_dy;
};

struct DIntegrator {
    uint32_t type, hide_emitters;
    float pdf_rr, pdf_rr_rcp;    // pdf_rr_rcp == pdf_rr (renderer.cpp:634, Q1)
    uint32_t depth_rr, depth_max;
    uint32_t num_emitters, num_area_lights;
    uint32_t id_sun, id_envmap;
    uint32_t has_opacity;        // some instance's BSDF carries an opacity texture: traversal runs its alpha-test variant
    uint32_t shade_bins;         // bit b set: shading bin b (= b200pt_bsdf_type) is in use; 0 = the scene is not binned
};

// Everything the kernels dereference.  Passed by value as a kernel parameter.
struct DeviceScene {
    const BvhNode *nodes;        uint32_t num_nodes;     // binary layout (default / the GPU LBVH builder), else empty
    const WideNode *wide_nodes;  uint32_t num_wide_nodes; // compressed 8-wide layout (B200PT_CREATE_BVH8); exactly one of the two is populated
    const TriVerts *tri_verts;   uint32_t num_tris;
    const TriShade *tri_shade;
    const uint8_t *tri_bsdf_type; // per triangle: b200pt_bsdf_type of its instance's BSDF (0 = none), the shading bin of a hit
    const AnalyticPrim *analytic; uint32_t num_analytic;
    const DInstance *instances;  uint32_t num_instances;
    const DBsdf *bsdfs;          uint32_t num_bsdfs;
    const DTexture *textures;    uint32_t num_textures;
    const float *pixels;
    const DMedium *media;        uint32_t num_media;
    const DEmitter *emitters;
    const float *envmap_tables;
    const float *kc_brdf_avg;    // 128*128
    const float *kc_albedo_avg;  // 128
    const float *cdf_area_light; // num_area_lights + 1, NOT normalised (Q4)
    const uint32_t *map_area_light_instance;
    const float *light_tri_cdf;  // per area-light mesh: inclusive prefix sums of triangle areas / total
    const uint32_t *light_tri_ids;
    float scene_bmin[3], scene_bmax[3];
    // A cut through the top of the BVH (plus the boxes of the analytic primitives): min.xyz max.xyz per box, together
    // they cover all geometry.  Screen tiles whose pyramid of camera rays misses every box trace nothing (k_cull_tiles).
    const float *cull_boxes;     uint32_t num_cull_boxes;
    const float *fine_cull_boxes; uint32_t num_fine_cull_boxes; // finer cover, grouped by coarse box: single pixels
    const uint32_t *cull_fine_begin; // num_cull_boxes + 1 entries: the fine boxes under coarse box b are [begin[b], begin[b + 1])
    DIntegrator integrator;
};

} // namespace b200pt
