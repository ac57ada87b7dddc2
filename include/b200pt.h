/*
 * b200pt.h — C ABI of the B200-native path-tracing integrator.
 *
 * Drop-in boundary.  The reference has no FFI; the seam this
 * ABI replaces is
 *
 *     csrt::RendererConfig  (include/csrt/renderer/renderer.hpp:18-28)
 *         -> csrt::Renderer::Renderer(const RendererConfig&)   (src/renderer/renderer.cpp:259)
 *         -> csrt::Renderer::Draw(float *frame)                (src/renderer/renderer.cpp:678)
 *
 * i.e. everything the reference does between "the XML parser has produced a
 * RendererConfig" and "frame[] holds w*h*3 linear-RGB floats".  Every struct
 * below is a plain-C mirror of the reference struct cited next to it; enum
 * values are numerically identical to the reference enums so that the glue
 * (oracle/ref_glue.cpp) is a field-by-field copy.
 *
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 * All functions return 0 on success and a negative B200PT_E* code on failure;
 * b200pt_last_error() returns the message (the C++ glue rethrows it as
 * csrt::MyException, matching renderer.cpp:696-710).
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PT_ABI_VERSION 1u
#define B200PT_INVALID_ID 0xFFFFFFFFu              /* csrt::kInvalidId, defs.hpp:22 */
#define B200PT_NO_OFFSET 0xFFFFFFFFFFFFFFFFull     /* "attribute array absent" */

/* error codes */
#define B200PT_OK 0
#define B200PT_EINVAL (-1)   /* bad argument / inconsistent scene description */
#define B200PT_ECUDA (-2)    /* CUDA runtime failure (message holds the code)  */
#define B200PT_ENOMEM (-3)
#define B200PT_EIO (-4)      /* scene-pack read/write failure */
#define B200PT_EUNSUPPORTED (-5)

/* csrt::TextureType, include/csrt/renderer/textures/texture.hpp:13-19 */
enum { B200PT_TEX_NONE = 0, B200PT_TEX_CONSTANT = 1, B200PT_TEX_CHECKERBOARD = 2, B200PT_TEX_BITMAP = 3 };
/* csrt::BsdfType, include/csrt/renderer/bsdfs/bsdf.hpp:17-27 */
enum {
    B200PT_BSDF_NONE = 0, B200PT_BSDF_AREA_LIGHT = 1, B200PT_BSDF_DIFFUSE = 2, B200PT_BSDF_ROUGH_DIFFUSE = 3,
    B200PT_BSDF_CONDUCTOR = 4, B200PT_BSDF_DIELECTRIC = 5, B200PT_BSDF_THIN_DIELECTRIC = 6, B200PT_BSDF_PLASTIC = 7
};
/* csrt::InstanceType, include/csrt/rtcore/instance.hpp:13-22 */
enum {
    B200PT_INST_NONE = 0, B200PT_INST_CUBE = 1, B200PT_INST_RECTANGLE = 2, B200PT_INST_MESHES = 3,
    B200PT_INST_SPHERE = 4, B200PT_INST_DISK = 5, B200PT_INST_CYLINDER = 6
};
/* csrt::EmitterType, include/csrt/renderer/emitters/emitter.hpp:19-28 */
enum {
    B200PT_EMIT_NONE = 0, B200PT_EMIT_POINT = 1, B200PT_EMIT_SPOT = 2, B200PT_EMIT_DIRECTIONAL = 3,
    B200PT_EMIT_SUN = 4, B200PT_EMIT_ENVMAP = 5, B200PT_EMIT_CONSTANT = 6
};
/* csrt::IntegratorType, include/csrt/renderer/integrators/integrator.hpp:10-14 */
enum { B200PT_INTEGRATOR_PATH = 0, B200PT_INTEGRATOR_VOLPATH = 1 };
/* csrt::PhaseFunctionType, include/csrt/renderer/medium/medium.hpp:12-16 */
enum { B200PT_PHASE_ISOTROPIC = 0, B200PT_PHASE_HG = 1 };

/* csrt::Camera::Info, include/csrt/renderer/camera.hpp:12-21 */
typedef struct b200pt_camera {
    uint32_t spp;
    int32_t width;
    int32_t height;
    float fov_x;      /* degrees; fov_y = fov_x*height/width (camera.cpp:30, quirk Q7) */
    float eye[3];
    float look_at[3];
    float up[3];
} b200pt_camera;

/* csrt::IntegratorInfo, include/csrt/renderer/integrators/integrator.hpp:16-27 */
typedef struct b200pt_integrator {
    uint32_t type;
    uint32_t hide_emitters;
    float pdf_rr;
    uint32_t depth_rr;
    uint32_t depth_max;
} b200pt_integrator;

/* csrt::TextureInfo, include/csrt/renderer/textures/texture.hpp:21-27.
 * Matrices are row-major 4x4 exactly as csrt::Mat4::rows (mat4.hpp). */
typedef struct b200pt_texture {
    uint32_t type;
    float color0[3];        /* constant colour, or checkerboard color0 */
    float color1[3];        /* checkerboard color1 */
    float to_uv[16];        /* checkerboard / bitmap to_uv */
    int32_t width, height, channels; /* bitmap only */
    uint32_t reserved;
    uint64_t pixel_offset;  /* bitmap only: first float of this bitmap in scene.pixels[] */
} b200pt_texture;

/* csrt::BsdfInfo, include/csrt/renderer/bsdfs/bsdf.hpp:40-58 (union flattened) */
typedef struct b200pt_bsdf {
    uint32_t type;
    uint32_t twosided;
    uint32_t id_opacity;
    uint32_t id_bump_map;
    uint32_t id_radiance;               /* area light  */
    uint32_t id_diffuse_reflectance;    /* diffuse, rough diffuse, plastic */
    uint32_t id_roughness_u;            /* conductor, (thin) dielectric; rough diffuse / plastic: roughness */
    uint32_t id_roughness_v;
    uint32_t id_specular_reflectance;   /* conductor, dielectrics, plastic */
    uint32_t id_specular_transmittance; /* dielectrics */
    float eta;                          /* dielectrics, plastic: int_ior/ext_ior */
    float reflectivity[3];              /* conductor */
    float edgetint[3];                  /* conductor */
    float area_light_weight;            /* AreaLightInfo::weight */
    uint32_t use_fast_approx;           /* RoughDiffuseInfo::use_fast_approx (ignored by the reference) */
} b200pt_bsdf;

/* csrt::MediumInfo, include/csrt/renderer/medium/medium.hpp:40-45 */
typedef struct b200pt_medium {
    uint32_t type;          /* 0 = homogeneous */
    float sigma_a[3];
    float sigma_s[3];
    uint32_t phase_type;
    float g[3];
} b200pt_medium;

/* csrt::InstanceInfo, include/csrt/rtcore/instance.hpp:24-51.  Mesh attribute
 * vectors of all instances live in the shared pools of b200pt_scene_desc;
 * offsets are in elements (vertices / triangles), B200PT_NO_OFFSET if the
 * reference vector is empty. */
typedef struct b200pt_instance {
    uint32_t type;
    uint32_t id_bsdf;
    uint32_t id_medium_int;
    uint32_t id_medium_ext;
    uint32_t flip_normals;   /* parsed, never used by the reference (Q14) */
    float to_world[16];
    float sphere_radius;
    float sphere_center[3];
    float cylinder_radius;
    float cylinder_p0[3];
    float cylinder_p1[3];
    uint32_t reserved;
    uint64_t num_vertices;
    uint64_t num_triangles;
    uint64_t position_offset;
    uint64_t normal_offset;
    uint64_t texcoord_offset;
    uint64_t tangent_offset;
    uint64_t bitangent_offset;
    uint64_t index_offset;
} b200pt_instance;

/* csrt::EmitterInfo, include/csrt/renderer/emitters/emitter.hpp:30-47 (union flattened) */
typedef struct b200pt_emitter {
    uint32_t type;
    float position[3];       /* point */
    float direction[3];      /* directional, sun */
    float radiance[3];       /* directional/sun/constant radiance; point/spot intensity */
    float cutoff_angle;      /* spot (radians) */
    float beam_width;        /* spot (radians) */
    float cos_cutoff_angle;  /* sun */
    uint32_t id_texture;     /* spot texture, sun texture, envmap radiance */
    float to_world[16];      /* spot, envmap */
} b200pt_emitter;

/* csrt::RendererConfig, include/csrt/renderer/renderer.hpp:18-28.
 * All pointers are HOST pointers owned by the caller; b200pt_create copies. */
typedef struct b200pt_scene_desc {
    uint32_t abi_version;    /* B200PT_ABI_VERSION */
    uint32_t reserved;
    b200pt_camera camera;
    b200pt_integrator integrator;

    uint64_t num_textures;  const b200pt_texture *textures;
    uint64_t num_pixels;    const float *pixels;             /* bitmap pool (renderer.cpp:371-392) */
    uint64_t num_bsdfs;     const b200pt_bsdf *bsdfs;
    uint64_t num_media;     const b200pt_medium *media;
    uint64_t num_instances; const b200pt_instance *instances;
    uint64_t num_emitters;  const b200pt_emitter *emitters;

    uint64_t num_positions;  const float *positions;    /* xyz */
    uint64_t num_normals;    const float *normals;      /* xyz */
    uint64_t num_texcoords;  const float *texcoords;    /* uv  */
    uint64_t num_tangents;   const float *tangents;     /* xyz */
    uint64_t num_bitangents; const float *bitangents;   /* xyz */
    uint64_t num_triangles;  const uint32_t *indices;   /* 3 per triangle, relative to the instance's vertex range */
} b200pt_scene_desc;

typedef struct b200pt_context *b200pt_handle;

typedef struct b200pt_create_opts {
    int32_t device;          /* CUDA ordinal; -1 = current device */
    uint32_t max_leaf_size;  /* BVH leaf size, 0 = default (4; the 8-wide layout holds at most 3 per leaf slot) */
    uint64_t max_paths_in_flight; /* wavefront capacity in paths, 0 = default */
    uint32_t flags;          /* B200PT_CREATE_* bit mask */
    uint32_t reserved;
} b200pt_create_opts;

/* Build the BVH on the GPU (LBVH: Morton codes, radix sort, Karras' parallel radix tree, bottom-up refit) instead of with
 * the host binned-SAH builder: the tree is ready in milliseconds instead of seconds, traversal is slower.
 * Replaces csrt::BvhBuilder::Build (src/rtcore/accel/bvh_builder.cpp:50-206), also an LBVH. */
#define B200PT_CREATE_GPU_LBVH 1u
/* Collapse the host SAH tree into the compressed 8-wide layout (eight quantised child boxes per 80-byte node, Ylitie et al.
 * 2017) instead of the default binary layout (two exact child boxes per 64-byte node).  Both find the same hits.  A ray
 * fetches 2.6x fewer nodes in the wide tree, but on B200 traversal is bound by instruction issue, not by fetches, and the
 * eight-box test costs more instructions than the 2.6 binary steps it replaces: measured 1.4x slower, so
 * the wide tree is the option and the cross-check, not the default. */
#define B200PT_CREATE_BVH8 2u

/* What to render.  width/height/spp = 0 take the value from the scene's camera
 * (the reference CLI overrides them after parsing: apps/main.cpp:46-52). */
typedef struct b200pt_render_opts {
    uint32_t width, height, spp;
    uint64_t seed;              /* Philox key */
    uint32_t tile_rank;         /* this process renders tiles t with t % tile_world == tile_rank */
    uint32_t tile_world;        /* 0 or 1 = whole frame */
    uint32_t collect_stats;     /* B200PT_STATS_* bit mask */
    uint32_t flags;             /* B200PT_RENDER_* bit mask */
} b200pt_render_opts;

/* Trace a camera ray for every sample even in screen tiles that provably see no geometry.  By default such tiles are
 * dropped by a conservative visibility pre-pass (only in scenes without environment / sun emitters, where an escaped
 * camera ray carries no radiance); the frame is bit-identical either way. */
#define B200PT_RENDER_NO_TILE_CULL 1u

/* Per kernel class of the wavefront pipeline. */
typedef struct b200pt_kernel_stats {
    double ms;                   /* summed device time of its launches (B200PT_STATS_TIMING) */
    uint64_t launches;
    uint64_t rays;               /* rays traced by it (B200PT_STATS_COUNTERS) */
    uint64_t node_visits;        /* 8-wide layout: 80-byte nodes fetched (8 child-box tests each);
                                    binary layout: child-box tests, 2 per 64-byte node fetched */
    uint64_t prim_tests;         /* ray-triangle / ray-analytic tests */
} b200pt_kernel_stats;

#define B200PT_STATS_COUNTERS 1u /* count rays / node visits / primitive tests (slows traversal kernels a little) */
#define B200PT_STATS_TIMING 2u   /* CUDA-event time per kernel class (no effect on the kernels themselves) */

typedef struct b200pt_stats {
    double render_ms;            /* device time of the last render (CUDA events on the render stream) */
    double upload_ms, bvh_build_ms;
    double bvh_gpu_ms;           /* GPU LBVH builder only: device time of Morton codes + sort + hierarchy + refit */
    uint64_t samples;            /* width*height*spp rendered by this rank */
    uint64_t kernel_launches;    /* kernels launched by the last render */
    uint64_t num_bvh_nodes, num_triangles, num_prims;
    uint32_t bvh_width;          /* 2 = binary layout (64-byte nodes, default), 8 = compressed wide layout (80-byte nodes) */
    uint32_t bvh_depth;          /* levels of the wide tree (0 for the binary layout) */
    uint64_t local_tiles, active_tiles; /* 8x8 tiles owned by this rank / those with a pixel that passed the visibility pre-pass */
    uint64_t active_pixels;      /* pixels of this rank that passed the visibility pre-pass (= the pixels camera rays are traced for) */
    b200pt_kernel_stats primary; /* k_primary: ray-gen + closest hit of camera rays */
    b200pt_kernel_stats extend;  /* k_trace  : closest hit of bounce rays + any-hit of NEE rays in ONE launch per bounce;
                                    ms / launches cover the whole launch, the counters only the bounce rays */
    b200pt_kernel_stats shadow;  /* counters of the NEE rays traced inside k_trace (ms = 0, launches = 0) */
    b200pt_kernel_stats shade;   /* k_shade  : shading, NEE generation, sampling, compaction */
    b200pt_kernel_stats other;   /* resets, resolve, finalize */
    b200pt_kernel_stats tail;    /* k_tail: one path per lane to the end, once a batch's survivor queue is short (its rays are
                                    counted under extend / shadow) */
} b200pt_stats;

/* ---- lifecycle: replaces csrt::Renderer ctor/dtor (renderer.cpp:259-369) ---- */
int b200pt_create(const b200pt_scene_desc *scene, const b200pt_create_opts *opts, b200pt_handle *out);
void b200pt_destroy(b200pt_handle h);

/* ---- render: replaces csrt::Renderer::Draw(float*) (renderer.cpp:678-721) ----
 * frame_host: width*height*3 floats, linear RGB, row 0 = top, caller-owned.
 * Blocking.  With tile_world > 1 only this rank's tiles are written (others untouched). */
int b200pt_render(b200pt_handle h, const b200pt_render_opts *opts, float *frame_host);

/* Same, but the frame stays in HBM: frame_dev is a DEVICE pointer to
 * width*height*3 floats; `stream` is a cudaStream_t (NULL = default stream).
 * Asynchronous with respect to the host. */
int b200pt_render_device(b200pt_handle h, const b200pt_render_opts *opts, float *frame_dev, void *stream);

/* Progressive preview: replaces csrt::Renderer::Draw(index_frame, frame, frame_srgb) (renderer.cpp:97-138, 723-746), the
 * call the reference's GLUT viewer makes once per displayed frame.  ONE sample per pixel with the sub-pixel offset
 * (VdC_2, VdC_3)(frame_index + 1); frame_dev (DEVICE, width*height*3, row 0 = top) holds the running mean over the frames
 * rendered so far and is updated in place: frame = (frame_index * frame + min(sample, 1)) / (frame_index + 1);
 * frame_srgb_dev (DEVICE, may be NULL) receives its sRGB-encoded copy with row 0 at the BOTTOM, as glDrawPixels wants it.
 * opts->spp is ignored (apps/main.cpp:53-58 forces 1).  Asynchronous with respect to the host. */
int b200pt_render_progressive_device(b200pt_handle h, const b200pt_render_opts *opts, uint32_t frame_index, float *frame_dev,
                                     float *frame_srgb_dev, void *stream);

/* Multi-GPU tile path: render this rank's interleaved tiles into
 * a compact DEVICE buffer of b200pt_tile_buffer_floats() floats (equal on every
 * rank, so one all-gather moves it), then scatter the gathered [world][floats]
 * buffer to the final frame. */
uint64_t b200pt_tile_buffer_floats(uint32_t width, uint32_t height, uint32_t tile_world);
int b200pt_render_tiles_device(b200pt_handle h, const b200pt_render_opts *opts, float *tiles_dev, void *stream);
int b200pt_assemble_tiles_device(b200pt_handle h, uint32_t width, uint32_t height, uint32_t tile_world,
                                 const float *gathered_dev, float *frame_dev, void *stream);

int b200pt_get_stats(b200pt_handle h, b200pt_stats *out);
const char *b200pt_last_error(b200pt_handle h);   /* h may be NULL: last error of a failed create/load */

/* ---- derived tables the reference computes on the host between config and kernel
 * (renderer.cpp:311-314, 571-611); exposed so tests can compare them. ---- */
int b200pt_get_kulla_conty(b200pt_handle h, float *brdf_avg_128x128, float *albedo_avg_128); /* h may be NULL (host only) */
int b200pt_get_envmap_tables(b200pt_handle h, float *out, uint64_t capacity_floats, uint64_t *num_floats,
                             float *normalization);

/* Test hook, HOST ONLY (needs no GPU): the sampling order of mesh area light `light` (index into the scene's list of
 * emissive instances, in instance order) as b200pt_create lays it out — the triangles of the instance, numbered as in its
 * own index list, in the order of the cumulative distribution that stands for BLAS::Sample's area-weighted BVH descent
 * (blas.cpp:79-98): the leaf order of the reference's Morton-built tree (bvh_builder.cpp:92-141), so that the same random
 * number picks the same triangle.  *count = triangles of the light (0 for a sphere / disk / cylinder light); tri_ids / cdf
 * (either may be NULL) receive min(*count, capacity) entries. */
int b200pt_debug_light_order(const b200pt_scene_desc *scene, uint32_t light, uint32_t *tri_ids, float *cdf, uint64_t capacity,
                             uint64_t *count);

/* ---- test hooks: pointwise access to the device leaf functions, so that tests can compare them with the reference's
 * functions at fixed inputs instead of through whole images.  Not needed by a renderer host. ---- */
typedef struct b200pt_debug_ray {
    float o[3], d[3];        /* origin, unit direction (csrt::Ray, include/csrt/rtcore/ray.hpp:9-27) */
    float tmin, tmax;        /* the reference constructs rays with t_min = 1e-4, t_max = FLT_MAX (ray.cpp:18-20) */
} b200pt_debug_ray;
typedef struct b200pt_debug_hit {
    float t;                 /* csrt::Ray::t_max after TLAS::Intersect (tlas.cpp:13-43); tmax of the input ray on a miss */
    uint32_t prim;           /* 0xFFFFFFFF miss | 0x80000000 + index of the sphere / disk / cylinder among the analytic
                                instances | index of the triangle in the scene description's enumeration (instances in
                                order, their triangles in order), + 0x40000000 when hit from its back side */
    float u, v;              /* triangle: barycentric weights of vertex 0 and 1 (triangle.cpp:64-87) */
} b200pt_debug_hit;
#define B200PT_DEBUG_ANY_HIT 1u      /* TLAS::IntersectAny (tlas.cpp:44-76): prim = 0 if occluded, 0xFFFFFFFF if not */
#define B200PT_DEBUG_PER_LANE_LOOP 2u /* the tail kernel's one-ray-per-lane loop instead of the persistent loop */
#define B200PT_DEBUG_PACKET_LOOP 8u  /* the warp-packet loop of the camera / first-vertex NEE rays (binary tree only): 32 consecutive rays walk together */
#define B200PT_DEBUG_RAW_PRIM 4u      /* prim as the kernels store it (leaf-order triangle index): input of B200PT_EVAL_SURFACE */
/* rays_host / hits_host: n elements each, HOST memory.  Runs the product's traversal kernels on the scene's tree. */
int b200pt_debug_trace(b200pt_handle h, const b200pt_debug_ray *rays_host, uint64_t n, uint32_t flags, b200pt_debug_hit *hits_host);

/* Leaf functions of the shading stage at caller-supplied inputs: `in_host` = n x B200PT_EVAL_IN floats, `out_host` = n x
 * B200PT_EVAL_OUT floats, `id` = index of the BSDF / emitter / medium / texture.  Sampling routines draw their random
 * numbers from the reference's LCG (math.hpp:57-63) started at the seed whose bit pattern is in[18]; out[15] returns the
 * state after the call (so the NUMBER of draws is checked too).
 *   what                    reference function                                    in                                   out
 *   BSDF_EVALUATE           Bsdf::Evaluate      (bsdf.cpp:213-236)                wi[0:3] wo[3:6] n[6:9] t[9:12]       valid pdf att[2:5]
 *                                                                                 b[12:15] uv[15:17] inside[17]
 *   BSDF_SAMPLE             Bsdf::Sample        (bsdf.cpp:188-211)                same, wi ignored, seed[18]           valid pdf att[2:5] wi[5:8]
 *   EMITTER_SAMPLE          Emitter::Sample / Evaluate(rec) / Pdf(-wi)            origin[0:3] xi_0[3] xi_1[4]          valid harsh distance wi[3:6]
 *                           (emitter.cpp:177-261)                                                                      Le[6:9] pdf[9]
 *   EMITTER_DIR             Emitter::Evaluate(dir) / Pdf(dir)                     dir[0:3]                             Le[0:3] pdf[3]
 *   MEDIUM_SAMPLE           Medium::Sample      (homogeneous.cpp:9-51)            max_distance[0] seed[18]             valid scattered pdf distance att[4:7]
 *   MEDIUM_EVALUATE         Medium::Evaluate    (homogeneous.cpp:53-82)           distance[0]                          same
 *   PHASE_SAMPLE            Medium::SamplePhase (medium.cpp:58-72)                wo[3:6] seed[18]                     valid pdf att[2:5] wi[5:8]
 *   PHASE_EVALUATE          Medium::EvaluatePhase (medium.cpp:74-87)              wi[0:3] wo[3:6]                      valid pdf att[2:5]
 *   TEXTURE                 Texture::GetColor   (texture.cpp)                     uv[0:2]                              rgb[0:3]
 *   SURFACE                 hit attributes of Primitive::Intersect (triangle.cpp:115-148, sphere.cpp:46-88, ...)
 *                           prim bits[0] (leaf-order index as the traversal kernels store it: NOT the b200pt_debug_hit
 *                           numbering) u[1] v[2] ray origin[3:6] dir[6:9] t[9]      pos[0:3] n[3:6] t[6:9] b[9:12] uv[12:14] inside[14] inst bits[15] */
enum {
    B200PT_EVAL_BSDF_EVALUATE = 0, B200PT_EVAL_BSDF_SAMPLE = 1, B200PT_EVAL_EMITTER_SAMPLE = 2, B200PT_EVAL_EMITTER_DIR = 3,
    B200PT_EVAL_MEDIUM_SAMPLE = 4, B200PT_EVAL_MEDIUM_EVALUATE = 5, B200PT_EVAL_PHASE_SAMPLE = 6, B200PT_EVAL_PHASE_EVALUATE = 7,
    B200PT_EVAL_TEXTURE = 8, B200PT_EVAL_SURFACE = 9
};
#define B200PT_EVAL_IN 32
#define B200PT_EVAL_OUT 16
int b200pt_debug_eval(b200pt_handle h, uint32_t what, uint32_t id, uint64_t n, const float *in_host, float *out_host);

/* Test hook ("exact mode"): renders the frame with the REFERENCE's loop shape and random-number stream — one thread per
 * pixel, its samples in order, one LCG per pixel seeded with Tea<4>(pixel_offset, 0) that runs on through every sample and
 * vertex (Renderer::DrawPixel renderer.cpp:62-85, RandomFloat math.hpp:57-63, the draws in the order GCC evaluates the
 * reference's call arguments) — around the product's own device functions (ShadeVertex, the per-lane traversal of the
 * scene's tree).  Each sample takes the decisions of the same sample of the reference's --cpu frame, so the two frames agree
 * PER PIXEL to float rounding, except where a last-bit difference (CUDA vs glibc libm, fused vs separate rounding) flips a
 * decision.  width / height / spp 0 = the scene's.  frame_host: width * height * 3 floats, rows top to bottom (the layout
 * of Renderer::Draw).  Slow by design (no wavefront, no sorting); refuses alpha-tested scenes (B200PT_EINVAL). */
int b200pt_debug_render_replay(b200pt_handle h, uint32_t width, uint32_t height, uint32_t spp, float *frame_host);

/* ---- scene packs: a lossless binary serialisation of b200pt_scene_desc, so a
 * scene parsed once by the reference's XML parser can travel without it. ---- */
typedef struct b200pt_scene b200pt_scene;
int b200pt_scene_load(const char *path, b200pt_scene **out);
int b200pt_scene_save(const b200pt_scene_desc *scene, const char *path);
const b200pt_scene_desc *b200pt_scene_get_desc(const b200pt_scene *s);
void b200pt_scene_free(b200pt_scene *s);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
