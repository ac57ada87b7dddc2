/*
 * pt_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's path-tracing hot path.
 *
 * Plain C over the flattened scene (include/b200pt.h).  It restates the reference ALGORITHM — same
 * per-pixel LCG stream, same draw order (GCC evaluates call arguments right to left),
 * same per-instance LBVH + TLAS and traversal order, same float expression order — so that its
 * frames can be compared with frames of the real reference build (oracle/_ref) value for value.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may call it; the product
 * (libb200pt.so) never links or loads it.
 *
 * Pinned by tests/test_oracle_pinning.py against (a) the golden frames in tests/golden/ produced by
 * the reference build (tests/golden/make_golden.py) and (b) live frames of oracle/_ref when present.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "b200pt.h"

/* ------------------------------------------------------------------------------------------- */
/* defs.hpp:22-29, math.hpp:15-27                                                              */
/* ------------------------------------------------------------------------------------------- */
#define kInvalidId 0xFFFFFFFFu
#define kEpsilonFloat 1.1920928955078125e-7f
#define kEpsilonDistance 1e-4f
#define kEpsilon 0.01f
#define kMaxFloat 3.402823466e+38f
#define kLowestFloat (-3.402823466e+38f)
#define kMaxUint 0xFFFFFFFFu
static const float kPi = 3.141592653589793f;
static const float k2Pi = 3.141592653589793f * 2.0f;
static const float kPiDiv2 = 3.141592653589793f * 0.5f;
static const float kPiDiv4 = 3.141592653589793f * 0.25f;
#define k1DivPi (1.0f / kPi)
#define k1Div2Pi (1.0f / k2Pi)
#define k1Div4Pi (1.0f / (4.0f * kPi))
#define kLutResolution 128

typedef struct { float u, v; } Vec2;
typedef struct { float x, y, z; } Vec3;
typedef struct { float x, y, z, w; } Vec4;
typedef struct { Vec4 rows[4]; } Mat4;

/* src/tensor/vec3.cpp — note: every division is a multiplication by the reciprocal */
static inline Vec3 v3(float x, float y, float z) { Vec3 r = {x, y, z}; return r; }
static inline Vec3 v3s(float s) { Vec3 r = {s, s, s}; return r; }
static inline Vec3 add(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 sub(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline Vec3 mul(Vec3 a, Vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Vec3 vdiv(Vec3 a, Vec3 b) { const float k0 = 1.0f / b.x, k1 = 1.0f / b.y, k2 = 1.0f / b.z; return v3(a.x * k0, a.y * k1, a.z * k2); }
static inline Vec3 muls(Vec3 a, float t) { return v3(a.x * t, a.y * t, a.z * t); }
static inline Vec3 smul(float t, Vec3 a) { return v3(t * a.x, t * a.y, t * a.z); }
static inline Vec3 divs(Vec3 a, float t) { const float k = 1.0f / t; return v3(a.x * k, a.y * k, a.z * k); }
static inline Vec3 adds(Vec3 a, float t) { return v3(a.x + t, a.y + t, a.z + t); }
static inline Vec3 sadd(float t, Vec3 a) { return v3(t + a.x, t + a.y, t + a.z); }
static inline Vec3 ssub(float t, Vec3 a) { return v3(t - a.x, t - a.y, t - a.z); }
static inline Vec3 sdiv(float t, Vec3 a) { const float k0 = 1.0f / a.x, k1 = 1.0f / a.y, k2 = 1.0f / a.z; return v3(t * k0, t * k1, t * k2); }
static inline Vec3 neg(Vec3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float length(Vec3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline Vec3 normalize(Vec3 a) { const float k = 1.0f / length(a); return muls(a, k); }
static inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x); }
static inline Vec3 vmin(Vec3 a, Vec3 b) { return v3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline Vec3 vmax(Vec3 a, Vec3 b) { return v3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline Vec3 vsqrt(Vec3 a) { return v3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
static inline float comp(Vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline void setcomp(Vec3 *a, int i, float v) { if (i == 0) a->x = v; else if (i == 1) a->y = v; else a->z = v; }
static inline float sqr(float t) { return t * t; }
static inline Vec3 sqr3(Vec3 a) { return mul(a, a); }
/* math.hpp:72-84 */
static inline float lerpf(float a, float b, float t) { return (1.0f - t) * a + t * b; }
static inline Vec3 lerp3(Vec3 a, Vec3 b, float t) { return add(smul(1.0f - t, a), smul(t, b)); }
static inline Vec3 bary3(const Vec3 *v, float a, float b, float c) { return add(add(smul(a, v[0]), smul(b, v[1])), smul(c, v[2])); }
static inline Vec2 bary2(const Vec2 *v, float a, float b, float c) {
    Vec2 r = {a * v[0].u + b * v[1].u + c * v[2].u, a * v[0].v + b * v[1].v + c * v[2].v};
    return r;
}

/* src/tensor/vec4.cpp:164-167, src/tensor/mat4.cpp */
static inline float dot4(Vec4 a, Vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline Vec4 v4(float x, float y, float z, float w) { Vec4 r = {x, y, z, w}; return r; }
static inline Vec4 mul4(Vec4 a, Vec4 b) { return v4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline Vec4 add4(Vec4 a, Vec4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline Vec4 sub4(Vec4 a, Vec4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline Vec4 smul4(float t, Vec4 a) { return v4(t * a.x, t * a.y, t * a.z, t * a.w); }
static Mat4 mat_identity(void) {
    Mat4 m = {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
    return m;
}
static Mat4 mat_load(const float *p) {
    Mat4 m;
    for (int r = 0; r < 4; ++r) m.rows[r] = v4(p[4 * r], p[4 * r + 1], p[4 * r + 2], p[4 * r + 3]);
    return m;
}
static Mat4 mat_transpose(const Mat4 *m) { /* mat4.cpp:102-108 */
    Mat4 t = {{{m->rows[0].x, m->rows[1].x, m->rows[2].x, m->rows[3].x},
               {m->rows[0].y, m->rows[1].y, m->rows[2].y, m->rows[3].y},
               {m->rows[0].z, m->rows[1].z, m->rows[2].z, m->rows[3].z},
               {m->rows[0].w, m->rows[1].w, m->rows[2].w, m->rows[3].w}}};
    return t;
}
static Mat4 mat_inverse(const Mat4 *m) { /* mat4.cpp:110-168 */
    const Vec4 *rows = m->rows;
    const float coef00 = rows[2].z * rows[3].w - rows[3].z * rows[2].w, coef02 = rows[1].z * rows[3].w - rows[3].z * rows[1].w,
                coef03 = rows[1].z * rows[2].w - rows[2].z * rows[1].w;
    const float coef04 = rows[2].y * rows[3].w - rows[3].y * rows[2].w, coef06 = rows[1].y * rows[3].w - rows[3].y * rows[1].w,
                coef07 = rows[1].y * rows[2].w - rows[2].y * rows[1].w;
    const float coef08 = rows[2].y * rows[3].z - rows[3].y * rows[2].z, coef10 = rows[1].y * rows[3].z - rows[3].y * rows[1].z,
                coef11 = rows[1].y * rows[2].z - rows[2].y * rows[1].z;
    const float coef12 = rows[2].x * rows[3].w - rows[3].x * rows[2].w, coef14 = rows[1].x * rows[3].w - rows[3].x * rows[1].w,
                coef15 = rows[1].x * rows[2].w - rows[2].x * rows[1].w;
    const float coef16 = rows[2].x * rows[3].z - rows[3].x * rows[2].z, coef18 = rows[1].x * rows[3].z - rows[3].x * rows[1].z,
                coef19 = rows[1].x * rows[2].z - rows[2].x * rows[1].z;
    const float coef20 = rows[2].x * rows[3].y - rows[3].x * rows[2].y, coef22 = rows[1].x * rows[3].y - rows[3].x * rows[1].y,
                coef23 = rows[1].x * rows[2].y - rows[2].x * rows[1].y;
    const Vec4 fac0 = {coef00, coef00, coef02, coef03}, fac1 = {coef04, coef04, coef06, coef07}, fac2 = {coef08, coef08, coef10, coef11},
               fac3 = {coef12, coef12, coef14, coef15}, fac4 = {coef16, coef16, coef18, coef19}, fac5 = {coef20, coef20, coef22, coef23};
    const Vec4 vec0 = {rows[1].x, rows[0].x, rows[0].x, rows[0].x}, vec1 = {rows[1].y, rows[0].y, rows[0].y, rows[0].y},
               vec2 = {rows[1].z, rows[0].z, rows[0].z, rows[0].z}, vec3 = {rows[1].w, rows[0].w, rows[0].w, rows[0].w};
    const Vec4 inv0 = add4(sub4(mul4(vec1, fac0), mul4(vec2, fac1)), mul4(vec3, fac2)),
               inv1 = add4(sub4(mul4(vec0, fac0), mul4(vec2, fac3)), mul4(vec3, fac4)),
               inv2 = add4(sub4(mul4(vec0, fac1), mul4(vec1, fac3)), mul4(vec3, fac5)),
               inv3 = add4(sub4(mul4(vec0, fac2), mul4(vec1, fac4)), mul4(vec2, fac5));
    const Vec4 sign_a = {+1.0f, -1.0f, +1.0f, -1.0f}, sign_b = {-1.0f, +1.0f, -1.0f, +1.0f};
    const Vec4 i0 = mul4(inv0, sign_a), i1 = mul4(inv1, sign_b), i2 = mul4(inv2, sign_a), i3 = mul4(inv3, sign_b);
    const Vec4 row0 = {i0.x, i1.x, i2.x, i3.x};
    const Vec4 dot0 = mul4(rows[0], row0);
    const float dot1 = (dot0.x + dot0.y) + (dot0.z + dot0.w);
    const float one_over_determinant = 1.0f / dot1;
    Mat4 r = {{smul4(one_over_determinant, i0), smul4(one_over_determinant, i1), smul4(one_over_determinant, i2),
               smul4(one_over_determinant, i3)}};
    return r;
}
static Vec4 mat_mul_vec(const Mat4 *m, Vec4 v) { return v4(dot4(m->rows[0], v), dot4(m->rows[1], v), dot4(m->rows[2], v), dot4(m->rows[3], v)); }
static Mat4 mat_mul(const Mat4 *a, const Mat4 *b) { /* mat4.cpp:181-195 */
    const Mat4 bt = mat_transpose(b);
    Mat4 r;
    for (int i = 0; i < 4; ++i) r.rows[i] = v4(dot4(a->rows[i], bt.rows[0]), dot4(a->rows[i], bt.rows[1]), dot4(a->rows[i], bt.rows[2]), dot4(a->rows[i], bt.rows[3]));
    return r;
}
static Vec3 TransformPoint(const Mat4 *m, Vec3 p) { /* mat4.cpp:265-268 + vec4.cpp:93-97 */
    const Vec4 r = mat_mul_vec(m, v4(p.x, p.y, p.z, 1.0f));
    const float k = 1.0f / r.w;
    return v3(r.x * k, r.y * k, r.z * k);
}
static Vec3 TransformVector(const Mat4 *m, Vec3 v) { /* mat4.cpp:270-273 + vec4.hpp:55: normalised! */
    const Vec4 r = mat_mul_vec(m, v4(v.x, v.y, v.z, 0.0f));
    return normalize(v3(r.x, r.y, r.z));
}
static Mat4 mat_translate(Vec3 t) {
    Mat4 m = {{{1, 0, 0, t.x}, {0, 1, 0, t.y}, {0, 0, 1, t.z}, {0, 0, 0, 1}}};
    return m;
}

/* ------------------------------------------------------------------------------------------- */
/* src/utils/math.cpp, include/csrt/utils/math.hpp                                             */
/* ------------------------------------------------------------------------------------------- */
uint32_t oracle_tea4(uint32_t v0, uint32_t v1) { /* math.hpp:43-54 */
    uint32_t s0 = 0;
    for (uint32_t n = 0; n < 4; ++n) {
        s0 += 0x9e3779b9;
        v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
        v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
    }
    return v0;
}
float oracle_random_float(uint32_t *seed) { /* math.hpp:57-63 */
    *seed = *seed * 1664525u + 1013904223u;
    return (float)(*seed & 0x00ffffff) / (float)(0x01000000u);
}
#define RandomFloat oracle_random_float
float oracle_van_der_corput2(uint32_t index) { /* math.hpp:29-41 */
    const float base_inv = 1.0f / 2;
    float result = 0.0f, frac = base_inv;
    while (index > 0) {
        result += frac * (index % 2);
        index = (uint32_t)(index * base_inv);
        frac *= base_inv;
    }
    return result;
}
float oracle_van_der_corput3(uint32_t index) { /* math.hpp:29-41 with base = 3 (the preview path's second offset) */
    const float base_inv = 1.0f / 3;
    float result = 0.0f, frac = base_inv;
    while (index > 0) {
        result += frac * (index % 3);
        index = (uint32_t)(index * base_inv);
        frac *= base_inv;
    }
    return result;
}
float oracle_mis_weight(float pdf1, float pdf2) { /* math.cpp:8-13 */
    pdf1 *= pdf1;
    pdf2 *= pdf2;
    return pdf1 / (pdf1 + pdf2);
}
#define MisWeight oracle_mis_weight
static Vec3 SampleConeUniform(float cos_cutoff, float xi_0, float xi_1) { /* math.cpp:15-22 */
    const float cos_theta = 1.0f - (1.0f - cos_cutoff) * xi_0, phi = 2.0f * kPi * xi_1;
    const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
    return v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
}
static Vec3 SampleSphereUniform(float xi_0, float xi_1) { /* math.cpp:24-29 */
    const float cos_theta = 1.0f - 2.0f * xi_0, phi = k2Pi * xi_1;
    const float sin_theta = sqrtf(1.0f - sqr(cos_theta));
    return v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
}
void oracle_sample_hemis_cos(float xi_0, float xi_1, float *vec, float *pdf) { /* math.cpp:31-38 */
    const float cos_theta = sqrtf(xi_0), phi = k2Pi * xi_1;
    const float sin_theta = sqrtf(1.0f - sqr(cos_theta));
    vec[0] = sin_theta * cosf(phi), vec[1] = sin_theta * sinf(phi), vec[2] = cos_theta;
    *pdf = k1DivPi * cos_theta;
}
static void SampleHemisCos(float xi_0, float xi_1, Vec3 *vec, float *pdf) { oracle_sample_hemis_cos(xi_0, xi_1, &vec->x, pdf); }
static uint32_t BinarySearch(uint32_t num, const float *cdf, float target) { /* math.cpp:40-55 */
    uint32_t begin = 0, end = num, middle;
    while (begin + 1 != end) {
        middle = (begin + end) >> 1;
        if (cdf[middle] < target) begin = middle;
        else if (cdf[middle] > target) end = middle;
        else return middle;
    }
    return end;
}
static int SolveQuadratic(float a, float b, float c, float *x0, float *x1) { /* math.cpp:57-99 */
    if (a == 0.0f) {
        if (b != 0.0f) { *x0 = *x1 = -c / b; return 1; }
        return 0;
    }
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0.0f) return 0;
    float temp;
    const float sqrt_discrim = sqrtf(discrim);
    if (b < 0.0f) temp = -0.5f * (b - sqrt_discrim);
    else temp = -0.5f * (b + sqrt_discrim);
    *x0 = temp / a;
    *x1 = c / temp;
    if (*x0 > *x1) { const float t = *x0; *x0 = *x1; *x1 = t; }
    return 1;
}
static void CartesianToSpherical(Vec3 vec, float *theta, float *phi, float *r) { /* math.cpp:102-119 */
    if (r != NULL) *r = length(vec);
    vec = normalize(vec);
    *theta = acosf(fminf(1.0f, fmaxf(-1.0f, vec.y)));
    if (vec.z == 0 && vec.x == 0) {
        *phi = 0;
    } else {
        *phi = atan2f(vec.z, vec.x);
        if (*phi < 0.0f) *phi += 2.0f * kPi;
    }
}
static Vec3 SphericalToCartesian(float theta, float phi, float r) { /* math.cpp:122-128 */
    const float sin_theta = sinf(theta);
    return v3(r * sinf(phi) * sin_theta, r * cosf(theta), r * cosf(phi) * sin_theta);
}
static Vec3 LocalToWorld(Vec3 local, Vec3 up) { /* math.cpp:130-146 */
    Vec3 C;
    if (sqrtf(sqr(up.x) + sqr(up.z)) > kEpsilonFloat) {
        const float len_inv = 1.0f / sqrtf(sqr(up.x) + sqr(up.z));
        C = v3(up.z * len_inv, 0, -up.x * len_inv);
    } else {
        const float len_inv = 1.0f / sqrtf(sqr(up.y) + sqr(up.z));
        C = v3(0, up.z * len_inv, -up.y * len_inv);
    }
    const Vec3 B = normalize(cross(C, up));
    return normalize(add(add(smul(local.x, B), smul(local.y, C)), smul(local.z, up)));
}
static Mat4 LocalToWorldMat(Vec3 up) { /* math.cpp:148-165 */
    Vec3 C;
    if (sqrtf(sqr(up.x) + sqr(up.z)) > kEpsilonFloat) {
        const float len_inv = 1.0f / sqrtf(sqr(up.x) + sqr(up.z));
        C = v3(-up.z * len_inv, 0, up.x * len_inv);
    } else {
        const float len_inv = 1.0f / sqrtf(sqr(up.y) + sqr(up.z));
        C = v3(0, -up.z * len_inv, up.y * len_inv);
    }
    const Vec3 B = normalize(cross(C, up));
    Mat4 m = {{{B.x, B.y, B.z, 0}, {C.x, C.y, C.z, 0}, {up.x, up.y, up.z, 0}, {0, 0, 0, 1}}};
    return m;
}

/* ------------------------------------------------------------------------------------------- */
/* Scene objects                                                                               */
/* ------------------------------------------------------------------------------------------- */
typedef struct { Vec3 min_, max_; } AABB;
static AABB aabb_empty(void) { AABB b = {{kMaxFloat, kMaxFloat, kMaxFloat}, {kLowestFloat, kLowestFloat, kLowestFloat}}; return b; }
static void aabb_add_point(AABB *b, Vec3 p) { b->min_ = vmin(p, b->min_); b->max_ = vmax(p, b->max_); }
static void aabb_add(AABB *b, const AABB *o) { b->min_ = vmin(o->min_, b->min_); b->max_ = vmax(o->max_, b->max_); }

typedef struct { /* ray.hpp:9-27 */
    float t_min, t_max;
    int k[3];
    Vec3 shear;
    Vec3 origin, dir, dir_rcp;
} Ray;

typedef struct { /* hit.hpp:9-30 */
    int valid, inside;
    uint32_t id_instance, id_primitve, id_medium_int, id_medium_ext;
    Vec2 texcoord;
    Vec3 position, normal, tangent, bitangent;
} Hit;

typedef struct { /* bvh_builder.hpp:11-25 */
    int leaf;
    uint32_t id, id_left, id_right, id_object;
    float area;
    AABB aabb;
} BvhNode;

enum { kPrimTriangle = 1, kPrimSphere, kPrimDisk, kPrimCylinder };
typedef struct { /* primitive.hpp:25-55 */
    uint32_t id;
    int type;
    /* triangle.hpp:12-19 */
    Vec2 texcoords[3];
    Vec3 positions[3], normals[3], tangents[3], bitangents[3];
    /* sphere / disk / cylinder */
    float radius, length;
    Vec3 center;
    Mat4 to_world;
} Primitive;

typedef struct {
    uint32_t type;
    Vec3 color0, color1;
    Mat4 to_uv;
    int width, height, channel;
    const float *data;
} Texture;

typedef struct { /* BsdfData, bsdf.hpp:60-79 */
    uint32_t type;
    int twosided;
    const Texture *opacity, *bump_map;
    const Texture *radiance, *diffuse_reflectance, *roughness, *roughness_u, *roughness_v, *specular_reflectance,
        *specular_transmittance;
    int use_fast_approx;
    Vec3 reflectivity3, edgetint, F_avg3;
    float reflectivity, eta, eta_inv, F_avg, F_avg_inv;
} Bsdf;

typedef struct { /* HomogeneousMediumData + PhaseFunctionData */
    float sampling_weight;
    Vec3 sigma_s, sigma_t;
    uint32_t phase_type;
    Vec3 g;
} Medium;

typedef struct {
    uint32_t type;
    Vec3 position, direction, radiance;
    float cutoff_angle, cos_cutoff_angle, uv_factor, beam_width, cos_beam_width, transition_width_rcp;
    const Texture *texture;
    Mat4 to_world, to_local;
    int width, height;
    float normalization;
    const float *cdf_cols, *cdf_rows, *weight_rows;
} Emitter;

typedef struct {
    uint32_t id, id_medium_int, id_medium_ext;
    const BvhNode *nodes;       /* BLAS nodes (blas.cpp:10-16) */
    const Primitive *primitives;
} Instance;

typedef struct {
    int watertight;
    /* camera.cpp:26-37 */
    int width, height;
    uint32_t spp;
    float spp_inv;
    Vec3 eye, front, view_dx, view_dy;
    /* IntegratorData, integrator.hpp:29-69 */
    uint32_t integrator_type;
    int hide_emitters;
    float pdf_rr, pdf_rr_rcp;
    uint32_t depth_rr, depth_max;
    uint32_t size_cdf_area_light, num_area_light, num_emitter, id_sun, id_envmap;
    Bsdf *bsdfs;
    Medium *media;
    Instance *instances;
    float *list_pdf_area_instance;
    Emitter *emitters;
    uint32_t *map_id_area_light_instance, *map_id_instance_area_light;
    float *cdf_area_light;
    uint32_t *map_instance_bsdf;
    BvhNode *nodes;            /* TLAS nodes first, then the BLASes (scene.cpp:499-508) */
    Primitive *primitives;
    Texture *textures;
    float *data_env_map, *brdf_avg, *albedo_avg;
    uint32_t num_instances;
    int has_tlas;
} Scene;

/* ------------------------------------------------------------------------------------------- */
/* Textures: textures/bitmap.cpp, checkboard.cpp, constant_texture.cpp, texture.cpp             */
/* ------------------------------------------------------------------------------------------- */
static Vec3 GetColorBitmap(const Texture *t, Vec2 texcoord) { /* bitmap.cpp:6-56 */
    const Vec3 uv = TransformPoint(&t->to_uv, v3(texcoord.u, texcoord.v, 0.0f));
    float x = uv.x * t->width, y = uv.y * t->height;
    while (x < 0) x += t->width;
    while (x > t->width - 1) x -= t->width;
    while (y < 0) y += t->height;
    while (y > t->height - 1) y -= t->height;
    const uint32_t x_0 = (uint32_t)x, y_0 = (uint32_t)y;
    const float t_x = x - x_0, t_y = y - y_0;
    const uint32_t x_1 = (t_x > 0.0f) ? x_0 + 1 : x_0, y_1 = (t_y > 0.0f) ? y_0 + 1 : y_0;
    if (t->channel == 1) {
        const float c00 = t->data[x_0 + t->width * y_0], c01 = t->data[x_0 + t->width * y_1], c10 = t->data[x_1 + t->width * y_0],
                    c11 = t->data[x_1 + t->width * y_1];
        const float c0 = lerpf(c00, c01, t_y), c1 = lerpf(c10, c11, t_y);
        return v3s(lerpf(c0, c1, t_x));
    }
    uint32_t o = (x_0 + t->width * y_0) * t->channel;
    const Vec3 c00 = v3(t->data[o], t->data[o + 1], t->data[o + 2]);
    o = (x_0 + t->width * y_1) * t->channel;
    const Vec3 c01 = v3(t->data[o], t->data[o + 1], t->data[o + 2]);
    o = (x_1 + t->width * y_0) * t->channel;
    const Vec3 c10 = v3(t->data[o], t->data[o + 1], t->data[o + 2]);
    o = (x_1 + t->width * y_1) * t->channel;
    const Vec3 c11 = v3(t->data[o], t->data[o + 1], t->data[o + 2]);
    const Vec3 c0 = lerp3(c00, c01, t_y), c1 = lerp3(c10, c11, t_y);
    return lerp3(c0, c1, t_x);
}
static Vec3 GetColorCheckerboard(const Texture *t, Vec2 texcoord) { /* checkboard.cpp:6-21 */
    Vec3 uv = TransformPoint(&t->to_uv, v3(texcoord.u, texcoord.v, 0.0f));
    while (uv.x > 1) uv.x -= 1;
    while (uv.x < 0) uv.x += 1;
    while (uv.y > 1) uv.y -= 1;
    while (uv.y < 0) uv.y += 1;
    const int x = 2 * (int)((int)(uv.x * 2) % 2) - 1, y = 2 * (int)((int)(uv.y * 2) % 2) - 1;
    return (x * y == 1) ? t->color0 : t->color1;
}
static Vec3 GetColor(const Texture *t, Vec2 texcoord) { /* texture.cpp:63-77 */
    switch (t->type) {
    case B200PT_TEX_CONSTANT: return t->color0;
    case B200PT_TEX_CHECKERBOARD: return GetColorCheckerboard(t, texcoord);
    case B200PT_TEX_BITMAP: return GetColorBitmap(t, texcoord);
    }
    return v3s(0);
}
static Vec2 GetGradient(const Texture *t, Vec2 texcoord) { /* texture.cpp:79-95, bitmap.cpp:58-68, checkboard.cpp:23-33 */
    Vec2 zero = {0, 0};
    if (t->type != B200PT_TEX_CHECKERBOARD && t->type != B200PT_TEX_BITMAP) return zero;
    const float delta = 1e-4f, norm = 1.0f / delta;
    const Vec2 tu = {texcoord.u + delta, texcoord.v + 0}, tv = {texcoord.u + 0, texcoord.v + delta};
    const float value = length(GetColor(t, texcoord)), value_u = length(GetColor(t, tu)), value_v = length(GetColor(t, tv));
    Vec2 g = {(value_u - value) * norm, (value_v - value) * norm};
    return g;
}
static int TextureIsTransparent(const Texture *t, Vec2 texcoord, uint32_t *seed) { /* texture.cpp:97-113 */
    switch (t->type) {
    case B200PT_TEX_CONSTANT: return t->color0.x < RandomFloat(seed); /* constant_texture.cpp:18-23 */
    case B200PT_TEX_CHECKERBOARD: return 0;
    case B200PT_TEX_BITMAP: { /* bitmap.cpp:70-100 */
        if (t->channel != 4) return 0;
        const Vec3 uv = TransformPoint(&t->to_uv, v3(texcoord.u, texcoord.v, 0.0f));
        float x = uv.x * t->width, y = uv.y * t->height;
        while (x < 0) x += t->width;
        while (x > t->width - 1) x -= t->width;
        while (y < 0) y += t->height;
        while (y > t->height - 1) y -= t->height;
        const uint32_t x_0 = (uint32_t)x, y_0 = (uint32_t)y;
        const float t_x = x - x_0, t_y = y - y_0;
        const uint32_t x_1 = (t_x > 0.0f) ? x_0 + 1 : x_0, y_1 = (t_y > 0.0f) ? y_0 + 1 : y_0;
        const float c00 = t->data[(x_0 + t->width * y_0) * 4 + 3], c01 = t->data[(x_0 + t->width * y_1) * 4 + 3],
                    c10 = t->data[(x_1 + t->width * y_0) * 4 + 3], c11 = t->data[(x_1 + t->width * y_1) * 4 + 3];
        const float c0 = lerpf(c00, c01, t_y), c1 = lerpf(c10, c11, t_y);
        return lerpf(c0, c1, t_x) < RandomFloat(seed);
    }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* BSDFs                                                                                       */
/* ------------------------------------------------------------------------------------------- */
typedef struct { /* bsdf.hpp:81-97 */
    int valid, inside;
    float pdf;
    Vec2 texcoord;
    Vec3 wi, wo, position, normal, tangent, bitangent, attenuation;
} BsdfSampleRec;

static Vec3 ToLocal(const BsdfSampleRec *r, Vec3 v) { return normalize(v3(dot(v, r->tangent), dot(v, r->bitangent), dot(v, r->normal))); }
static Vec3 ToWorld(const BsdfSampleRec *r, Vec3 v) { return normalize(add(add(smul(v.x, r->tangent), smul(v.y, r->bitangent)), smul(v.z, r->normal))); }
static Vec3 Reflect(Vec3 wi, Vec3 normal) { return normalize(sub(wi, smul(2.0f * dot(wi, normal), normal))); } /* ray.cpp:49-52 */
static int Refract(Vec3 wi, Vec3 normal, float eta_inv, Vec3 *wt) { /* ray.cpp:54-68 */
    const float cos_theta = fabsf(dot(wi, normal));
    const float k = 1.0f - sqr(eta_inv) * (1.0f - sqr(cos_theta));
    if (k < 0) return 0;
    *wt = normalize(add(smul(eta_inv, wi), smul(eta_inv * cos_theta - sqrtf(k), normal)));
    return 1;
}

/* microfacet.cpp — pow(x, 3) is std::pow(float, int), evaluated in double */
static void SampleGgx1(float xi_0, float xi_1, float roughness, Vec3 *vec, float *pdf) { /* :8-19 */
    const float alpha_2 = sqr(roughness);
    const float tan_theta_2 = alpha_2 * xi_0 / (1.0f - xi_0), phi = k2Pi * xi_1;
    const float cos_theta = 1.0f / sqrtf(1.0f + tan_theta_2), sin_theta = sqrtf(1.0f - sqr(cos_theta));
    *vec = v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
    *pdf = (float)(1.0f / (kPi * alpha_2 * pow((double)cos_theta, 3.0) * sqr(1.0f + tan_theta_2 / alpha_2)));
}
static void SampleGgx2(float xi_0, float xi_1, float roughness_u, float roughness_v, Vec3 *vec, float *pdf) { /* :21-37 */
    const float phi = (atanf(roughness_v / roughness_u * tanf(kPi + k2Pi * xi_1)) + kPi * floorf(2.0f * xi_1 + 0.5f));
    const float cos_phi = cosf(phi), sin_phi = sinf(phi), alpha_2 = 1.0f / (sqr(cos_phi / roughness_u) + sqr(sin_phi / roughness_v));
    const float tan_theta_2 = (float)(alpha_2 * xi_0 / (1.0 - xi_0));
    const float cos_theta = 1.0f / sqrtf(1.0f + tan_theta_2), sin_theta = sqrtf(1.0f - sqr(cos_theta));
    *vec = v3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
    *pdf = (float)(1.0f / (kPi * roughness_u * roughness_v * pow((double)cos_theta, 3.0) * sqr(1.0f + tan_theta_2 / alpha_2)));
}
static float PdfGgx1(float roughness, Vec3 vec) { /* :39-49 */
    const float cos_theta = vec.z;
    if (cos_theta <= 0.0f) return 0.0f;
    const float cos_theta_2 = sqr(cos_theta), tan_theta_2 = (1.0f - cos_theta_2) / cos_theta_2,
                cos_theta_3 = (float)pow((double)cos_theta, 3.0), alpha_2 = sqr(roughness);
    return alpha_2 / (kPi * cos_theta_3 * sqr(alpha_2 + tan_theta_2));
}
static float PdfGgx2(float roughness_u, float roughness_v, Vec3 vec) { /* :51-61 */
    const float cos_theta = vec.z;
    if (cos_theta <= 0.0f) return 0.0f;
    const float cos_theta_2 = sqr(cos_theta);
    return cos_theta / (kPi * roughness_u * roughness_v * sqr(sqr(vec.x / roughness_u) + sqr(vec.y / roughness_v) + cos_theta_2));
}
static float SmithG1Ggx1(float roughness, Vec3 v, Vec3 h) { /* :63-75 */
    const float N_dot_V = v.z;
    if (N_dot_V * h.z <= 0) return 0;
    const float cos_theta_2 = sqr(N_dot_V), tan_theta_2 = (1.0f - cos_theta_2) / cos_theta_2, alpha_2 = sqr(roughness);
    return 2.0f / (1.0f + sqrtf((float)(1.0 + alpha_2 * tan_theta_2)));
}
static float SmithG1Ggx2(float roughness_u, float roughness_v, Vec3 v, Vec3 h) { /* :77-85 */
    const float N_dot_V = v.z;
    if (N_dot_V * h.z <= 0) return 0;
    const float xy_alpha_2 = sqr(roughness_u * v.x) + sqr(roughness_v * v.y), tan_theta_2 = xy_alpha_2 / sqr(N_dot_V);
    return 2.0f / (1.0f + sqrtf(1.0f + tan_theta_2));
}
/* microfacet.hpp:24-29 */
static float FresnelSchlick1(float cos_theta, float r) { return (1.0f - r) * (float)pow((double)(1.0f - cos_theta), 5.0) + r; }
static Vec3 FresnelSchlick3(float cos_theta, Vec3 r) { return add(muls(ssub(1.0f, r), (float)pow((double)(1.0f - cos_theta), 5.0)), r); }

/* kulla_conty.cpp */
static float GetBrdfAvg(const float *buf, float cos_theta, float roughness) { /* :82-131 */
    const int R = kLutResolution;
    /* Quirk (found by tests/test_oracle_pointwise.py): kLutResolution is a uint32_t (kulla_conty.hpp:9), so the reference's
     * `offset_int2 >= kLutResolution - 1` compares UNSIGNED.  EvaluateDielectric passes negative cosines here (N_dot_O < 0 in its
     * transmission branch, dielectric.cpp:207-212; N_dot_I < 0 for light arriving from below): a column <= -1 wraps to a huge
     * value and takes the "last column" branch, i.e. a cosine below -1/128 reads the table at cosine 1.  Cosines in
     * (-1/128, 0) truncate to column 0 and extrapolate with a negative weight. */
    const float offset1 = roughness * R, offset2 = cos_theta * R;
    const int i1 = (int)offset1, i2 = (int)offset2;
    const int last_column = (uint32_t)i2 >= (uint32_t)(R - 1);
    if (i1 >= R - 1) {
        if (last_column) return buf[(R - 1) * R + R - 1];
        return lerpf(buf[(R - 1) * R + i2], buf[(R - 1) * R + i2 + 1], offset2 - i2);
    }
    if (last_column) return lerpf(buf[i1 * R + R - 1], buf[(i1 + 1) * R + R - 1], offset1 - i1);
    return lerpf(lerpf(buf[i1 * R + i2], buf[(i1 + 1) * R + i2], offset1 - i1),
                 lerpf(buf[i1 * R + i2 + 1], buf[(i1 + 1) * R + i2 + 1], offset1 - i1), offset2 - i2);
}
static float GetAlbedoAvg(const float *buf, float roughness) { /* :133-143 */
    const float offset = roughness * kLutResolution;
    const int i = (int)offset;
    if (i >= kLutResolution - 1) return buf[kLutResolution - 1];
    return lerpf(buf[i], buf[i + 1], offset - i);
}
static float IntegrateBRDF(Vec3 V, float roughness) { /* :13-37 */
    const uint32_t sample_count = 1024;
    const float step = 1.0f / 1024;
    const Vec3 N = {0.0f, 0.0f, 1.0f};
    float pdf_h, brdf_accum = 0.0f;
    Vec3 H, L;
    for (uint32_t i = 0; i < sample_count; ++i) {
        SampleGgx1(i * step, oracle_van_der_corput2(i), roughness, &H, &pdf_h);
        L = Reflect(V, H);
        const float G = SmithG1Ggx1(roughness, neg(V), H) * SmithG1Ggx1(roughness, L, H), N_dot_V = dot(N, neg(V)), N_Dot_L = dot(N, L),
                    N_dot_H = dot(N, H), H_dot_V = dot(H, neg(V));
        if (N_Dot_L > 0.0f && N_dot_H > 0.0f && H_dot_V > 0.0f) brdf_accum += (H_dot_V * G) / (N_dot_V * N_dot_H);
    }
    return fminf(brdf_accum * step, 1.0f);
}
static float IntegrateAlbedo(Vec3 V, float roughness, float brdf) { /* :39-58 */
    const uint32_t sample_count = 1024;
    const float step = 1.0f / 1024;
    const Vec3 N = {0.0f, 0.0f, 1.0f};
    float pdf_h, albedo_accum = 0.0f;
    Vec3 H, L;
    for (uint32_t i = 0; i < sample_count; ++i) {
        SampleGgx1(i * step, oracle_van_der_corput2(i), roughness, &H, &pdf_h);
        L = Reflect(V, H);
        const float N_Dot_L = dot(N, L), N_dot_H = dot(N, H), H_dot_V = dot(neg(V), H);
        if (N_Dot_L > 0.0f && N_dot_H > 0.0f && H_dot_V > 0.0f) albedo_accum += brdf * N_Dot_L;
    }
    return albedo_accum * 2.0f * step;
}
void oracle_kulla_conty(float *brdf_buffer, float *albedo_avg_buffer) { /* :62-80 */
    float step = 1.0f / kLutResolution, albedo_accum = 0.0f;
    for (int i = kLutResolution - 1; i >= 0; --i) {
        albedo_accum = 0.0f;
        float roughness = step * ((float)i + 0.5f);
        for (int j = kLutResolution - 1; j >= 0; --j) {
            const float N_dot_V = step * ((float)j + 0.5f);
            const Vec3 V = {-sqrtf(1.f - N_dot_V * N_dot_V), 0.0f, -N_dot_V};
            const float brdf_avg = IntegrateBRDF(V, roughness);
            brdf_buffer[i * kLutResolution + j] = brdf_avg;
            albedo_accum += IntegrateAlbedo(V, roughness, brdf_avg);
        }
        albedo_avg_buffer[i] = albedo_accum * step;
    }
}

/* diffuse.cpp */
static void EvaluateDiffuse(const Bsdf *d, BsdfSampleRec *rec) { /* :9-20 */
    rec->pdf = dot(rec->wo, rec->normal);
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const Vec3 albedo = GetColor(d->diffuse_reflectance, rec->texcoord);
    const float N_dot_I = dot(neg(rec->wi), rec->normal);
    rec->attenuation = muls(muls(albedo, k1DivPi), N_dot_I);
}
static void SampleDiffuse(const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :22-34 */
    Vec3 wi_local;
    const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed); /* right-to-left argument evaluation */
    SampleHemisCos(xi_0, xi_1, &wi_local, &rec->pdf);
    if (rec->pdf < kEpsilon) return;
    rec->wi = neg(ToWorld(rec, wi_local));
    rec->valid = 1;
    const Vec3 albedo = GetColor(d->diffuse_reflectance, rec->texcoord);
    const float N_dot_I = wi_local.z;
    rec->attenuation = muls(muls(albedo, k1DivPi), N_dot_I);
}

/* rough_diffuse.cpp:10-97 */
static void EvaluateOrenNayar(float rougness, Vec3 albedo, int use_fast_approx, BsdfSampleRec *rec) {
    const float conversion_factor = 0.70710678118f;
    const float sigma_2 = sqr(rougness * conversion_factor);
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi)), wo_local = ToLocal(rec, rec->wo);
    const float N_dot_I = wi_local.z, N_dot_O = wo_local.z, sin_theta_i = sqrtf(1.0f - N_dot_I * N_dot_I),
                sin_theta_o = sqrtf(1.0f - N_dot_O * N_dot_O);
    float phi_i, theta_i, phi_o, theta_o;
    CartesianToSpherical(wi_local, &theta_i, &phi_i, NULL);
    CartesianToSpherical(wo_local, &theta_o, &phi_o, NULL);
    float cos_phi_diff = cosf(phi_i) * cosf(phi_o) + sinf(phi_i) * sinf(phi_o);
    if (use_fast_approx) {
        float A = 1.0f - 0.5f * sigma_2 / (sigma_2 + 0.33f), B = 0.45f * sigma_2 / (sigma_2 + 0.09f);
        float sin_alpha, tan_beta;
        if (N_dot_I > N_dot_O) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / N_dot_I; }
        else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / N_dot_O; }
        rec->attenuation = muls(muls(muls(albedo, k1DivPi), N_dot_I), (A + B * fmaxf(cos_phi_diff, 0.0f) * sin_alpha * tan_beta));
    } else {
        float alpha = fmaxf(theta_i, theta_o), beta = fminf(theta_i, theta_o);
        float sin_alpha, sin_beta, tan_beta;
        if (N_dot_I > N_dot_O) { sin_alpha = sin_theta_o; sin_beta = sin_theta_i; tan_beta = sin_theta_i / N_dot_I; }
        else { sin_alpha = sin_theta_i; sin_beta = sin_theta_o; tan_beta = sin_theta_o / N_dot_O; }
        float tmp = sigma_2 / (sigma_2 + 0.09f), tmp2 = 4.0f * k1DivPi * k1DivPi * alpha * beta, tmp3 = 2.0f * beta * k1DivPi;
        float C1 = 1.0f - 0.5f * sigma_2 / (sigma_2 + 0.33f), C2 = 0.45f * tmp, C3 = 0.125f * tmp * tmp2 * tmp2,
              C4 = 0.17f * sigma_2 / (sigma_2 + 0.13f);
        if (cos_phi_diff > 0) C2 *= sin_alpha;
        else C2 = (float)(C2 * (sin_alpha - pow((double)tmp3, 3.0)));
        float tan_half = (sin_alpha + sin_beta) / (sqrtf(fmaxf(0.0f, 1.0f - sqr(sin_alpha))) + sqrtf(fmaxf(0.0f, 1.0f - sqr(sin_beta))));
        Vec3 sngl_scat = muls(albedo, (C1 + cos_phi_diff * C2 * tan_beta + (1.0f - fabsf(cos_phi_diff)) * C3 * tan_half)),
             dbl_scat = muls(sqr3(albedo), (C4 * (1.0f - cos_phi_diff * sqr(tmp3))));
        rec->attenuation = muls(muls(add(sngl_scat, dbl_scat), k1DivPi), N_dot_I);
    }
}
static void SampleRoughDiffuse(const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :99-115 */
    Vec3 wi;
    const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
    SampleHemisCos(xi_0, xi_1, &wi, &rec->pdf);
    if (rec->pdf < kEpsilon) return;
    rec->wi = neg(normalize(add(add(smul(wi.x, rec->tangent), smul(wi.y, rec->bitangent)), smul(wi.z, rec->normal))));
    rec->valid = 1;
    const float alpha = GetColor(d->roughness, rec->texcoord).x;
    const Vec3 albedo = GetColor(d->diffuse_reflectance, rec->texcoord);
    EvaluateOrenNayar(alpha, albedo, d->use_fast_approx, rec);
}
static void EvaluateRoughDiffuse(const Bsdf *d, BsdfSampleRec *rec) { /* :117-129 */
    rec->pdf = dot(rec->wo, rec->normal);
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const float alpha = GetColor(d->roughness, rec->texcoord).x;
    const Vec3 albedo = GetColor(d->diffuse_reflectance, rec->texcoord);
    EvaluateOrenNayar(alpha, albedo, d->use_fast_approx, rec);
}

/* conductor.cpp */
static Vec3 ConductorMultipleScatter(const Scene *s, const Bsdf *d, float N_dot_I, float N_dot_O, float roughness) { /* :14-28 */
    const float brdf_i = GetBrdfAvg(s->brdf_avg, N_dot_I, roughness), brdf_o = GetBrdfAvg(s->brdf_avg, N_dot_O, roughness),
                albedo_avg = GetAlbedoAvg(s->albedo_avg, roughness), f_ms = (1.0f - brdf_i) * (1.0f - brdf_o) / (kPi * (1.0f - albedo_avg));
    const Vec3 f_add = vdiv(muls(sqr3(d->F_avg3), albedo_avg), ssub(1.0f, muls(d->F_avg3, 1.0f - albedo_avg)));
    return muls(smul(f_ms, f_add), N_dot_I);
}
static void SampleConductor(const Scene *s, const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :34-77 */
    Vec3 h_local = v3s(0);
    float D = 0;
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x, alpha_v = GetColor(d->roughness_v, rec->texcoord).x;
    const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
    SampleGgx2(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const Vec3 h_world = ToWorld(rec, h_local);
    const float H_dot_O = dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * H_dot_O);
    if (rec->pdf < kEpsilon) return;
    rec->wi = neg(Reflect(neg(rec->wo), h_world));
    const float N_dot_I = dot(neg(rec->wi), rec->normal);
    if (N_dot_I < kEpsilonFloat) return;
    rec->valid = 1;
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi)), wo_local = ToLocal(rec, rec->wo);
    const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local),
                H_dot_I = dot(neg(rec->wi), h_world), N_dot_O = wo_local.z;
    const Vec3 F = FresnelSchlick3(H_dot_I, d->reflectivity3);
    rec->attenuation = divs(muls(muls(F, D), G), 4.0f * N_dot_O);
    if (alpha_u == alpha_v) rec->attenuation = add(rec->attenuation, ConductorMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u));
    rec->attenuation = mul(rec->attenuation, GetColor(d->specular_reflectance, rec->texcoord));
}
static void EvaluateConductor(const Scene *s, const Bsdf *d, BsdfSampleRec *rec) { /* :79-119 */
    const float N_dot_O = dot(rec->wo, rec->normal);
    if (N_dot_O < kEpsilonFloat) return;
    const Vec3 h_world = normalize(add(neg(rec->wi), rec->wo)), h_local = ToLocal(rec, h_world);
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x, alpha_v = GetColor(d->roughness_v, rec->texcoord).x,
                D = PdfGgx2(alpha_u, alpha_v, h_local), H_dot_O = dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * H_dot_O);
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi)), wo_local = ToLocal(rec, rec->wo);
    const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local),
                H_dot_I = dot(neg(rec->wi), h_world);
    const Vec3 F = FresnelSchlick3(H_dot_I, d->reflectivity3);
    rec->attenuation = divs(muls(muls(F, D), G), 4.0f * N_dot_O);
    if (alpha_u == alpha_v) {
        const float N_dot_I = dot(neg(rec->wi), rec->normal);
        rec->attenuation = add(rec->attenuation, ConductorMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u));
    }
    rec->attenuation = mul(rec->attenuation, GetColor(d->specular_reflectance, rec->texcoord));
}

/* dielectric.cpp — abs() has float semantics (oracle/_ref is built with -include math.h -include stdlib.h, Q13) */
static float DielectricMultipleScatter(const Scene *s, const Bsdf *d, float N_dot_I, float N_dot_O, float roughness, int inside, int reflect) { /* :14-38 */
    const float brdf_i = GetBrdfAvg(s->brdf_avg, N_dot_I, roughness), brdf_o = GetBrdfAvg(s->brdf_avg, N_dot_O, roughness),
                albedo_avg = GetAlbedoAvg(s->albedo_avg, roughness), f_ms = (1.0f - brdf_i) * (1.0f - brdf_o) / (kPi * (1.0f - albedo_avg));
    const float F_avg = inside ? d->F_avg_inv : d->F_avg, eta = inside ? d->eta_inv : d->eta;
    const float f_add = (float)(pow((double)F_avg, 2.0) * albedo_avg / (1.0f - F_avg * (1.0f - albedo_avg))),
                ratio_trans = (float)(((1.0f - d->F_avg) * (1.0f - d->F_avg_inv) * pow((double)eta, 2.0) /
                                       ((1.0f - d->F_avg) + (1.0f - d->F_avg_inv) * pow((double)eta, 2.0))));
    const float ret = f_ms * f_add * N_dot_I;
    return reflect ? (1.0f - ratio_trans) * ret : ratio_trans * ret;
}
static void SampleDielectric(const Scene *s, const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :44-145 */
    const float scale = 1.2f - 0.2f * sqrtf(fabsf(dot(neg(rec->wo), rec->normal)));
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x * scale, alpha_v = GetColor(d->roughness_v, rec->texcoord).x * scale;
    Vec3 h_local = v3s(0);
    float D = 0;
    const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
    SampleGgx2(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const Vec3 h_world = ToWorld(rec, h_local);
    float H_dot_O = dot(rec->wo, h_world);
    if (H_dot_O < kEpsilonFloat) return;
    float eta = d->eta, eta_inv = d->eta_inv;
    if (!rec->inside) { float temp = eta_inv; eta_inv = eta; eta = temp; }
    Vec3 wt = v3s(0);
    const int full_reflect = !Refract(neg(rec->wo), h_world, eta, &wt);
    float F = FresnelSchlick1(H_dot_O, d->reflectivity);
    const Vec3 wo_local = ToLocal(rec, rec->wo);
    if (full_reflect || RandomFloat(seed) < F) {
        rec->wi = neg(Reflect(neg(rec->wo), h_world));
        const float N_dot_I = dot(neg(rec->wi), rec->normal);
        if (N_dot_I < kEpsilonFloat) return;
        rec->pdf = F * D / (4.0f * H_dot_O);
        if (rec->pdf < kEpsilon) return;
        const Vec3 wi_local = ToLocal(rec, neg(rec->wi));
        const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local), N_dot_O = wo_local.z;
        rec->attenuation = v3s((F * D * G) / (4.0f * N_dot_O));
        if (alpha_u == alpha_v) rec->attenuation = adds(rec->attenuation, DielectricMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u, rec->inside, 1));
        rec->attenuation = mul(rec->attenuation, GetColor(d->specular_reflectance, rec->texcoord));
    } else {
        rec->wi = neg(wt);
        Vec3 wi_local = ToLocal(rec, neg(rec->wi));
        wi_local.z = -wi_local.z;
        const float N_dot_I = wi_local.z;
        if (N_dot_I < kEpsilonFloat) return;
        const float H_dot_I = -dot(wt, h_world);
        if (H_dot_I < kEpsilonFloat) return;
        H_dot_O = -H_dot_O;
        F = FresnelSchlick1(H_dot_I, d->reflectivity);
        rec->pdf = ((1.0f - F) * D) * fabsf(H_dot_O / sqr(eta_inv * H_dot_I + H_dot_O));
        if (rec->pdf < kEpsilon) return;
        const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local), N_dot_O = wo_local.z;
        rec->attenuation = v3s(((fabsf(H_dot_I) * fabsf(H_dot_O)) * ((1.0f - F) * G * D)) / fabsf(N_dot_O * sqr(eta_inv * H_dot_I + H_dot_O)));
        if (alpha_u == alpha_v) rec->attenuation = adds(rec->attenuation, DielectricMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u, !rec->inside, 0));
        rec->attenuation = muls(rec->attenuation, sqr(eta));
        rec->attenuation = mul(rec->attenuation, GetColor(d->specular_transmittance, rec->texcoord));
    }
    rec->valid = 1;
}
static void EvaluateDielectric(const Scene *s, const Bsdf *d, BsdfSampleRec *rec) { /* :147-224 */
    float eta = d->eta, eta_inv = d->eta_inv;
    if (rec->inside) { float temp = eta_inv; eta_inv = eta; eta = temp; }
    const float N_dot_O = dot(rec->wo, rec->normal);
    const int relfect = N_dot_O > 0.0f;
    const Vec3 h_world = relfect ? normalize(add(neg(rec->wi), rec->wo)) : neg(normalize(add(smul(eta_inv, neg(rec->wi)), rec->wo))),
               h_local = ToLocal(rec, h_world);
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x, alpha_v = GetColor(d->roughness_v, rec->texcoord).x,
                D = PdfGgx2(alpha_u, alpha_v, h_local), H_dot_I = dot(neg(rec->wi), h_world), H_dot_O = dot(rec->wo, h_world),
                F = FresnelSchlick1(H_dot_I, d->reflectivity);
    rec->pdf = relfect ? (F * D) / (4.0f * H_dot_O) : (((1.0f - F) * D) * fabsf(H_dot_O / sqr(eta_inv * H_dot_I + H_dot_O)));
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi));
    if (relfect) {
        const Vec3 wo_local = ToLocal(rec, rec->wo);
        const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local);
        rec->attenuation = v3s((F * D * G) / (4.0f * N_dot_O));
        if (alpha_u == alpha_v) {
            const float N_dot_I = dot(neg(rec->wi), rec->normal);
            rec->attenuation = adds(rec->attenuation, DielectricMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u, rec->inside, 1));
        }
        rec->attenuation = mul(rec->attenuation, GetColor(d->specular_reflectance, rec->texcoord));
    } else {
        const Vec3 wo_local = ToLocal(rec, neg(rec->wo));
        const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local);
        rec->attenuation = v3s(((fabsf(H_dot_I) * fabsf(H_dot_O)) * ((1.0f - F) * G * D)) / fabsf(N_dot_O * sqr(eta_inv * H_dot_I + H_dot_O)));
        if (alpha_u == alpha_v) {
            const float N_dot_I = dot(rec->normal, neg(rec->wi));
            rec->attenuation = adds(rec->attenuation, DielectricMultipleScatter(s, d, N_dot_I, N_dot_O, alpha_u, rec->inside, 0));
        }
        rec->attenuation = muls(rec->attenuation, sqr(eta));
        rec->attenuation = mul(rec->attenuation, GetColor(d->specular_transmittance, rec->texcoord));
    }
}

/* thin_dielectric.cpp */
static void SampleThinDielectric(const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :11-69 */
    Vec3 h_local = v3s(0);
    float D = 0;
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x, alpha_v = GetColor(d->roughness_v, rec->texcoord).x;
    const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
    SampleGgx2(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const Vec3 h_world = ToWorld(rec, h_local);
    const float H_dot_O = dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * H_dot_O);
    if (rec->pdf < kEpsilon) return;
    rec->wi = neg(Reflect(neg(rec->wo), h_world));
    const float N_dot_I = dot(neg(rec->wi), rec->normal);
    if (N_dot_I < kEpsilonFloat) return;
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi)), wo_local = ToLocal(rec, rec->wo);
    const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local),
                H_dot_I = dot(neg(rec->wi), h_world), N_dot_O = wo_local.z;
    float F = FresnelSchlick1(H_dot_I, d->reflectivity);
    if (F < 1.0f) F *= 2.0f / (1.0f + F);
    if (RandomFloat(seed) < F) {
        rec->pdf *= F;
        if (rec->pdf < kEpsilon) return;
        rec->attenuation = mul(v3s((F * D * G) / (4.0f * N_dot_O)), GetColor(d->specular_reflectance, rec->texcoord));
    } else {
        rec->pdf *= 1.0f - F;
        if (rec->pdf < kEpsilon) return;
        rec->attenuation = mul(v3s(((1.0f - F) * D * G) / (4.0f * N_dot_O)), GetColor(d->specular_transmittance, rec->texcoord));
        rec->wi = rec->wo;
    }
    rec->valid = 1;
}
static void EvaluateThinDielectric(const Bsdf *d, BsdfSampleRec *rec) { /* :71-124 */
    int reflect = 1;
    Vec3 wo = rec->wo;
    float N_dot_O = dot(rec->wo, rec->normal);
    if (fabsf(N_dot_O) < kEpsilonFloat) return;
    Vec3 wo_local = ToLocal(rec, rec->wo);
    if (N_dot_O < 0.0f) {
        reflect = 0;
        N_dot_O = -N_dot_O;
        wo_local.z = -wo_local.z;
        wo = ToWorld(rec, wo_local);
    }
    const Vec3 h_world = normalize(add(neg(rec->wi), wo)), h_local = ToLocal(rec, h_world);
    const float alpha_u = GetColor(d->roughness_u, rec->texcoord).x, alpha_v = GetColor(d->roughness_v, rec->texcoord).x,
                D = PdfGgx2(alpha_u, alpha_v, h_local), H_dot_I = dot(neg(rec->wi), h_world), H_dot_O = dot(rec->wo, h_world);
    float F = FresnelSchlick1(H_dot_I, d->reflectivity);
    if (F < 1.0f) F *= 2.0f / (1.0f + F);
    rec->pdf = reflect ? (F * D) / (4.0f * H_dot_O) : ((1.0f - F) * D) / (4.0f * H_dot_O);
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const Vec3 wi_local = ToLocal(rec, neg(rec->wi));
    const float G = SmithG1Ggx2(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx2(alpha_u, alpha_v, wo_local, h_local);
    if (reflect) rec->attenuation = mul(v3s((F * D * G) / (4.0f * N_dot_O)), GetColor(d->specular_reflectance, rec->texcoord));
    else rec->attenuation = mul(v3s(((1.0f - F) * D * G) / (4.0f * N_dot_O)), GetColor(d->specular_transmittance, rec->texcoord));
}

/* plastic.cpp */
static void SamplePlastic(const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) { /* :11-95 */
    const Vec3 kd = GetColor(d->diffuse_reflectance, rec->texcoord), ks = GetColor(d->specular_reflectance, rec->texcoord);
    float weight_spec = (ks.x + ks.y + ks.z) / ((kd.x + kd.y + kd.z) + (ks.x + ks.y + ks.z));
    const float N_dot_O = dot(rec->wo, rec->normal), kr_o = FresnelSchlick1(N_dot_O, d->reflectivity);
    float kr_i = kr_o, pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * (1.0f - weight_spec);
    pdf_spec = pdf_spec / (pdf_spec + pdf_diff);
    pdf_diff = 1.0f - pdf_spec;
    Vec3 h_local = v3s(0), h_world = v3s(0);
    float D = 0;
    const float alpha = GetColor(d->roughness, rec->texcoord).x;
    float N_dot_I = 0;
    if (RandomFloat(seed) < pdf_spec) {
        const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
        SampleGgx1(xi_0, xi_1, alpha, &h_local, &D);
        h_world = ToWorld(rec, h_local);
        rec->wi = neg(Reflect(neg(rec->wo), h_world));
        N_dot_I = dot(neg(rec->wi), rec->normal);
        if (N_dot_I < kEpsilonFloat) return;
        kr_i = FresnelSchlick1(N_dot_I, d->reflectivity);
        pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * weight_spec;
        pdf_spec = pdf_spec / (pdf_spec + pdf_diff), pdf_diff = 1.0f - pdf_spec;
        const float H_dot_O = dot(rec->wo, h_world);
        pdf_spec *= D / (4.0f * H_dot_O);
        pdf_diff *= dot(neg(rec->wi), rec->normal);
    } else {
        Vec3 wi_local = v3s(0);
        float pdf_diff_local = 0.0f;
        const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
        SampleHemisCos(xi_0, xi_1, &wi_local, &pdf_diff_local);
        rec->wi = neg(ToWorld(rec, wi_local));
        N_dot_I = dot(neg(rec->wi), rec->normal);
        kr_i = FresnelSchlick1(N_dot_I, d->reflectivity);
        pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * weight_spec;
        pdf_spec = pdf_spec / (pdf_spec + pdf_diff), pdf_diff = 1.0f - pdf_spec;
        h_world = normalize(add(neg(rec->wi), rec->wo)), h_local = ToLocal(rec, h_world);
        D = PdfGgx1(alpha, h_local);
        const float H_dot_O = dot(rec->wo, h_world);
        pdf_spec = (float)(pdf_spec * (D / (4.0 * H_dot_O)));
        pdf_diff *= pdf_diff_local;
    }
    rec->pdf = pdf_spec + pdf_diff;
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    if (pdf_spec > kEpsilon) {
        const Vec3 wi_local = ToLocal(rec, neg(rec->wi)), wo_local = ToLocal(rec, rec->wo);
        const float H_dot_I = dot(neg(rec->wi), h_world), F = FresnelSchlick1(H_dot_I, d->reflectivity),
                    G = (SmithG1Ggx1(alpha, wo_local, h_local) * SmithG1Ggx1(alpha, wi_local, h_local));
        Vec3 spec = v3s((F * D * G) / (4.0f * N_dot_O));
        rec->attenuation = add(rec->attenuation, mul(spec, ks));
    }
    if (pdf_diff > kEpsilon) {
        Vec3 diff = muls(muls(kd, k1DivPi), N_dot_I);
        diff = muls(diff, ((1.0f - kr_i) * (1.0f - kr_o)) / (1.0f - d->F_avg));
        rec->attenuation = add(rec->attenuation, diff);
    }
}
static void EvaluatePlastic(const Bsdf *d, BsdfSampleRec *rec) { /* :97-153 */
    const float N_dot_O = dot(rec->wo, rec->normal);
    if (N_dot_O < kEpsilonFloat) return;
    const Vec3 kd = GetColor(d->diffuse_reflectance, rec->texcoord), ks = GetColor(d->specular_reflectance, rec->texcoord);
    float weight_spec = (ks.x + ks.y + ks.z) / ((kd.x + kd.y + kd.z) + (ks.x + ks.y + ks.z));
    const float N_dot_I = dot(neg(rec->wi), rec->normal), kr_i = FresnelSchlick1(N_dot_I, d->reflectivity);
    float pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * (1.0f - weight_spec);
    pdf_spec = pdf_spec / (pdf_spec + pdf_diff);
    pdf_diff = 1.0f - pdf_spec;
    const Vec3 h_world = normalize(add(neg(rec->wi), rec->wo)), h_local = ToLocal(rec, h_world);
    const float alpha = GetColor(d->roughness, rec->texcoord).x, D = PdfGgx1(alpha, h_local), H_dot_O = dot(rec->wo, h_world);
    pdf_spec *= D / (4.0f * H_dot_O);
    const Vec3 wo_local = ToLocal(rec, rec->wo);
    pdf_diff *= wo_local.z;
    rec->pdf = pdf_spec + pdf_diff;
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    if (pdf_spec > kEpsilon) {
        const Vec3 wi_local = ToLocal(rec, neg(rec->wi));
        const float H_dot_I = dot(neg(rec->wi), h_world), F = FresnelSchlick1(H_dot_I, d->reflectivity),
                    G = (SmithG1Ggx1(alpha, wo_local, h_local) * SmithG1Ggx1(alpha, wi_local, h_local));
        Vec3 spec = v3s((F * D * G) / (4.0f * N_dot_O));
        rec->attenuation = add(rec->attenuation, mul(spec, ks));
    }
    if (pdf_diff > kEpsilon) {
        Vec3 diff = muls(muls(kd, k1DivPi), N_dot_I);
        const float kr_o = FresnelSchlick1(N_dot_O, d->reflectivity);
        diff = muls(diff, ((1.0f - kr_i) * (1.0f - kr_o)) / (1.0f - d->F_avg));
        rec->attenuation = add(rec->attenuation, diff);
    }
}

/* bsdf.cpp:188-276 */
static void BsdfSample(const Scene *s, const Bsdf *d, uint32_t *seed, BsdfSampleRec *rec) {
    switch (d->type) {
    case B200PT_BSDF_DIFFUSE: SampleDiffuse(d, seed, rec); break;
    case B200PT_BSDF_ROUGH_DIFFUSE: SampleRoughDiffuse(d, seed, rec); break;
    case B200PT_BSDF_CONDUCTOR: SampleConductor(s, d, seed, rec); break;
    case B200PT_BSDF_DIELECTRIC: SampleDielectric(s, d, seed, rec); break;
    case B200PT_BSDF_THIN_DIELECTRIC: SampleThinDielectric(d, seed, rec); break;
    case B200PT_BSDF_PLASTIC: SamplePlastic(d, seed, rec); break;
    }
}
static void BsdfEvaluate(const Scene *s, const Bsdf *d, BsdfSampleRec *rec) {
    switch (d->type) {
    case B200PT_BSDF_DIFFUSE: EvaluateDiffuse(d, rec); break;
    case B200PT_BSDF_ROUGH_DIFFUSE: EvaluateRoughDiffuse(d, rec); break;
    case B200PT_BSDF_CONDUCTOR: EvaluateConductor(s, d, rec); break;
    case B200PT_BSDF_DIELECTRIC: EvaluateDielectric(s, d, rec); break;
    case B200PT_BSDF_THIN_DIELECTRIC: EvaluateThinDielectric(d, rec); break;
    case B200PT_BSDF_PLASTIC: EvaluatePlastic(d, rec); break;
    }
}
static Vec3 ApplyBumpMapping(const Bsdf *d, Vec3 normal, Vec3 tangent, Vec3 bitangent, Vec2 texcoord) { /* :238-254 */
    if (d->bump_map == NULL) return normal;
    const Vec2 gradient = GetGradient(d->bump_map, texcoord);
    return normalize(add(sub(smul(-gradient.u, tangent), smul(gradient.v, bitangent)), normal));
}
static Vec3 GetRadiance(const Bsdf *d, Vec2 texcoord) { /* :256-266 */
    if (d->type == B200PT_BSDF_AREA_LIGHT) return GetColor(d->radiance, texcoord);
    return v3s(0);
}
static int BsdfIsTransparent(const Bsdf *d, Vec2 texcoord, uint32_t *seed) { return d->opacity && TextureIsTransparent(d->opacity, texcoord, seed); }

/* ------------------------------------------------------------------------------------------- */
/* Ray, AABB, primitives, BLAS/TLAS                                                            */
/* ------------------------------------------------------------------------------------------- */
static Ray MakeRay(const Scene *s, Vec3 origin, Vec3 dir) { /* ray.cpp:18-47 */
    Ray r;
    r.origin = origin, r.dir = dir, r.t_min = kEpsilonDistance, r.t_max = kMaxFloat;
    r.dir_rcp = v3(1.0f / (dir.x != 0 ? dir.x : kEpsilonDistance), 1.0f / (dir.y != 0 ? dir.y : kEpsilonDistance),
                   1.0f / (dir.z != 0 ? dir.z : kEpsilonDistance));
    r.k[0] = r.k[1] = r.k[2] = 0;
    r.shear = v3s(0);
    if (s->watertight) {
        r.k[2] = (fabs(dir.x) > fabs(dir.y) && fabs(dir.x) > fabs(dir.z)) ? 0 : (fabs(dir.y) > fabs(dir.z) ? 1 : 2);
        r.k[0] = r.k[2] + 1;
        if (r.k[0] == 3) r.k[0] = 0;
        r.k[1] = r.k[0] + 1;
        if (r.k[1] == 3) r.k[1] = 0;
        if (comp(dir, r.k[2]) < 0.0f) { int temp = r.k[0]; r.k[0] = r.k[1]; r.k[1] = temp; }
        r.shear = v3(comp(dir, r.k[0]) / comp(dir, r.k[2]), comp(dir, r.k[1]) / comp(dir, r.k[2]), 1.0f / comp(dir, r.k[2]));
    }
    return r;
}
static int AabbIntersect(const AABB *b, const Ray *ray) { /* aabb.cpp:29-48 */
    const Vec3 t_min = mul(sub(b->min_, ray->origin), ray->dir_rcp), t_max = mul(sub(b->max_, ray->origin), ray->dir_rcp);
    float t_enter = ray->t_min, t_exit = ray->t_max;
    for (int i = 0; i < 3; ++i) {
        if (comp(ray->dir_rcp, i) > 0) {
            t_enter = fmaxf(t_enter, comp(t_min, i));
            t_exit = fminf(t_exit, comp(t_max, i));
        } else {
            t_enter = fmaxf(t_enter, comp(t_max, i));
            t_exit = fminf(t_exit, comp(t_min, i));
        }
    }
    return t_enter <= t_exit;
}
static Hit HitInvalid(void) { /* hit.cpp:9-15 */
    Hit h;
    memset(&h, 0, sizeof(h));
    h.id_instance = h.id_primitve = h.id_medium_int = h.id_medium_ext = kInvalidId;
    return h;
}
static Hit HitFull(uint32_t id, int inside, Vec2 texcoord, Vec3 position, Vec3 normal, Vec3 tangent, Vec3 bitangent) { /* hit.cpp:26-35 */
    Hit h = HitInvalid();
    h.valid = 1, h.inside = inside, h.id_primitve = id, h.texcoord = texcoord, h.position = position, h.normal = normal,
    h.tangent = tangent, h.bitangent = bitangent;
    return h;
}
static Hit HitSample(uint32_t id, Vec2 texcoord, Vec3 position, Vec3 normal) { /* hit.cpp:17-24 */
    Hit h = HitInvalid();
    h.valid = 1, h.id_primitve = id, h.texcoord = texcoord, h.position = position, h.normal = normal;
    return h;
}

static int IntersectTriangle(const Scene *s, const Primitive *p, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* triangle.cpp:19-148 */
    float t, u, v, w, det_inv;
    if (s->watertight) {
        const Vec3 A = sub(p->positions[0], ray->origin), B = sub(p->positions[1], ray->origin), C = sub(p->positions[2], ray->origin);
        const float Ax = comp(A, ray->k[0]) - ray->shear.x * comp(A, ray->k[2]), Ay = comp(A, ray->k[1]) - ray->shear.y * comp(A, ray->k[2]);
        const float Bx = comp(B, ray->k[0]) - ray->shear.x * comp(B, ray->k[2]), By = comp(B, ray->k[1]) - ray->shear.y * comp(B, ray->k[2]);
        const float Cx = comp(C, ray->k[0]) - ray->shear.x * comp(C, ray->k[2]), Cy = comp(C, ray->k[1]) - ray->shear.y * comp(C, ray->k[2]);
        float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
        if (U == 0.0f || V == 0.0f || W == 0.0f) {
            double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
            U = (float)(CxBy - CyBx);
            double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
            V = (float)(AxCy - AyCx);
            double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
            W = (float)(BxAy - ByAx);
        }
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
        const float det = U + V + W;
        if (det == 0.0f) return 0;
        const float Az = ray->shear.z * comp(A, ray->k[2]), Bz = ray->shear.z * comp(B, ray->k[2]), Cz = ray->shear.z * comp(C, ray->k[2]);
        const float T = U * Az + V * Bz + W * Cz;
        det_inv = 1.0f / det;
        t = T * det_inv;
        if (t > ray->t_max || t < ray->t_min) return 0;
        u = U * det_inv, v = V * det_inv, w = W * det_inv;
    } else {
        const Vec3 v0v1 = sub(p->positions[1], p->positions[0]), v0v2 = sub(p->positions[2], p->positions[0]);
        const Vec3 P = cross(ray->dir, v0v2);
        det_inv = 1.0f / dot(v0v1, P);
        const Vec3 T = sub(ray->origin, p->positions[0]);
        v = dot(T, P) * det_inv;
        if (v < 0.0f || v > 1.0f) return 0;
        const Vec3 Q = cross(T, v0v1);
        w = dot(ray->dir, Q) * det_inv;
        if (w < 0.0f || (v + w) > 1.0f) return 0;
        t = dot(v0v2, Q) * det_inv;
        if (t > ray->t_max || t < ray->t_min) return 0;
        u = 1.0f - v - w;
    }
    const Vec2 texcoord = bary2(p->texcoords, u, v, w);
    if (bsdf != NULL && BsdfIsTransparent(bsdf, texcoord, seed)) return 0;
    ray->t_max = t;
    if (hit != NULL) {
        const int inside = det_inv < 0;
        const Vec3 position = bary3(p->positions, u, v, w);
        Vec3 normal = normalize(bary3(p->normals, u, v, w)), tangent = normalize(bary3(p->tangents, u, v, w)),
             bitangent = normalize(bary3(p->bitangents, u, v, w));
        if (bsdf != NULL) {
            normal = ApplyBumpMapping(bsdf, normal, tangent, bitangent, texcoord);
            bitangent = normalize(cross(normal, tangent));
            tangent = normalize(cross(bitangent, normal));
        }
        if (inside) { normal = neg(normal); bitangent = neg(bitangent); }
        *hit = HitFull(p->id, inside, texcoord, position, normal, tangent, bitangent);
    }
    return 1;
}
static Hit SampleTriangle(const Primitive *p, float xi_0, float xi_1) { /* triangle.cpp:150-160 */
    const float temp = sqrtf(1.0f - xi_0);
    const float u = 1.0f - temp, v = temp * xi_1, w = 1.0f - u - v;
    const Vec2 texcoord = bary2(p->texcoords, w, u, v);
    const Vec3 position = bary3(p->positions, w, u, v), normal = normalize(bary3(p->normals, w, u, v));
    return HitSample(p->id, texcoord, position, normal);
}

static int IntersectSphere(const Primitive *p, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* sphere.cpp:17-88 */
    const Mat4 to_local = mat_inverse(&p->to_world);
    const Vec3 ray_origin = sub(TransformPoint(&to_local, ray->origin), p->center), ray_direction = TransformVector(&to_local, ray->dir);
    const float a = dot(ray_direction, ray_direction), b = 2.0f * dot(ray_direction, ray_origin), c = dot(ray_origin, ray_origin) - sqr(p->radius);
    float t_near = 0.0f, t_far = 0.0f;
    if (!SolveQuadratic(a, b, c, &t_near, &t_far) || t_far < kEpsilonDistance) return 0;
    float t = t_near < kEpsilonDistance ? t_far : t_near;
    const Vec3 position_local = add(ray_origin, smul(t, ray_direction)), position = TransformPoint(&p->to_world, add(position_local, p->center));
    t = length(sub(position, ray->origin));
    if (t > ray->t_max || t < ray->t_min) return 0;
    float theta, phi;
    CartesianToSpherical(position_local, &theta, &phi, NULL);
    const Vec2 texcoord = {phi * k1Div2Pi, theta * k1DivPi};
    if (bsdf != NULL && BsdfIsTransparent(bsdf, texcoord, seed)) return 0;
    ray->t_max = t;
    if (hit != NULL) {
        const int inside = c < 0.0f;
        const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
        const Vec3 normal_local = normalize(position_local);
        Vec3 normal = TransformVector(&normal_to_world, normal_local);
        const float epsilon_jitter = 0.01f * kPi;
        float theta_prime = theta + epsilon_jitter;
        const int flip_bitangent = theta_prime > kPi;
        if (flip_bitangent) theta_prime = theta - epsilon_jitter;
        const Vec3 position_prime = TransformPoint(&p->to_world, SphericalToCartesian(theta_prime, phi, 1));
        Vec3 bitangent = normalize(sub(position_prime, position));
        if (flip_bitangent) bitangent = neg(bitangent);
        Vec3 tangent = normalize(cross(bitangent, normal));
        bitangent = normalize(cross(normal, tangent));
        if (bsdf != NULL) {
            normal = ApplyBumpMapping(bsdf, normal, tangent, bitangent, texcoord);
            bitangent = normalize(cross(normal, tangent));
            tangent = normalize(cross(bitangent, normal));
        }
        if (inside) { normal = neg(normal); bitangent = neg(bitangent); }
        *hit = HitFull(p->id, inside, texcoord, position, normal, tangent, bitangent);
    }
    return 1;
}
static Hit SampleSphere(const Primitive *p, float xi_0, float xi_1) { /* sphere.cpp:90-105 */
    const float cos_theta = 1.0f - 2.0f * xi_0;
    const Vec2 texcoord = {xi_1, acosf(cos_theta) * k1DivPi};
    const float sin_theta = sqrtf(1.0f - sqr(cos_theta)), phi = k2Pi * xi_1;
    const Vec3 normal_local = {sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta}, position_local = add(p->center, smul(p->radius, normal_local));
    const Vec3 position = TransformPoint(&p->to_world, position_local);
    const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
    const Vec3 normal = TransformVector(&normal_to_world, normal_local);
    return HitSample(p->id, texcoord, position, normal);
}

static int IntersectDisk(const Primitive *p, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* disk.cpp:17-110 */
    const Mat4 to_local = mat_inverse(&p->to_world);
    const Vec3 ray_origin = TransformPoint(&to_local, ray->origin), ray_direction = TransformVector(&to_local, ray->dir);
    const float t_z = -ray_origin.z / ray_direction.z;
    if (t_z < kEpsilonFloat) return 0;
    const Vec3 position_local = add(ray_origin, smul(t_z, ray_direction));
    if (length(position_local) > 0.5f) return 0;
    const Vec3 position = TransformPoint(&p->to_world, position_local);
    const float t = length(sub(position, ray->origin));
    if (t > ray->t_max || t < ray->t_min) return 0;
    float theta, phi, r;
    CartesianToSpherical(position_local, &theta, &phi, &r);
    const Vec2 texcoord = {r, phi * k1Div2Pi};
    if (bsdf != NULL && BsdfIsTransparent(bsdf, texcoord, seed)) return 0;
    ray->t_max = t;
    if (hit != NULL) {
        const int inside = ray_direction.z > 0;
        const float epsilon_jitter = 0.01f * kPi;
        float r_prime = r + epsilon_jitter;
        const int flip_bitangent = r_prime > r;
        if (flip_bitangent) r_prime = r - epsilon_jitter;
        float phi_prime = phi + epsilon_jitter;
        const int flip_tangent = phi_prime > kPi;
        if (flip_tangent) phi_prime = phi - epsilon_jitter;
        const Vec3 v0v1_local = sub(SphericalToCartesian(theta, phi, r_prime), position_local),
                   v0v2_local = sub(SphericalToCartesian(theta, phi_prime, r), position_local);
        const Vec2 delta_uv_1 = {r_prime - texcoord.u, texcoord.v - texcoord.v}, delta_uv_2 = {texcoord.u - texcoord.u, phi_prime * k1Div2Pi - texcoord.v};
        const float norm = 1.0f / (delta_uv_2.u * delta_uv_1.v - delta_uv_1.u * delta_uv_2.v);
        Vec3 tangent = normalize(muls(sub(smul(delta_uv_1.v, v0v2_local), smul(delta_uv_2.v, v0v1_local)), norm)),
             bitangent = normalize(muls(sub(smul(delta_uv_2.u, v0v1_local), smul(delta_uv_1.u, v0v2_local)), norm)), normal = {0, 0, 1};
        if (flip_bitangent) bitangent = neg(bitangent);
        if (flip_tangent) tangent = neg(tangent);
        bitangent = normalize(cross(normal, tangent));
        tangent = normalize(cross(bitangent, normal));
        if (bsdf != NULL) {
            normal = ApplyBumpMapping(bsdf, normal, tangent, bitangent, texcoord);
            bitangent = normalize(cross(normal, tangent));
            tangent = normalize(cross(bitangent, normal));
        }
        const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
        normal = TransformVector(&normal_to_world, normal);
        tangent = TransformVector(&p->to_world, tangent);
        bitangent = TransformVector(&p->to_world, bitangent);
        if (inside) { normal = neg(normal); bitangent = neg(bitangent); }
        *hit = HitFull(p->id, inside, texcoord, position, normal, tangent, bitangent);
    }
    return 1;
}
static Hit SampleDisk(const Primitive *p, float xi_0, float xi_1) { /* disk.cpp:112-141 */
    const float r1 = 2.0f * xi_0 - 1.0f, r2 = 2.0f * xi_1 - 1.0f;
    float phi, r;
    if (r1 == 0.0f && r2 == 0.0f) { r = phi = 0; }
    else if (sqr(r1) > sqr(r2)) { r = r1; phi = kPiDiv4 * (r2 / r1); }
    else { r = r2; phi = kPiDiv2 - (r1 / r2) * kPiDiv4; }
    const Vec2 xy = {r * cosf(phi), r * sinf(phi)};
    const Vec2 texcoord = {r, phi * k1Div2Pi};
    const Vec3 position = TransformPoint(&p->to_world, v3(xy.u * 0.5f, xy.v * 0.5f, 0));
    const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
    const Vec3 normal = TransformVector(&normal_to_world, v3(0, 0, 1));
    return HitSample(p->id, texcoord, position, normal);
}

static int IntersectCylinder(const Primitive *p, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* cylinder.cpp:21-90 */
    const Mat4 to_local = mat_inverse(&p->to_world);
    const Vec3 ray_origin = TransformPoint(&to_local, ray->origin), ray_direction = TransformVector(&to_local, ray->dir);
    const float a = sqr(ray_direction.x) + sqr(ray_direction.y), b = 2.0f * (ray_direction.x * ray_origin.x + ray_direction.y * ray_origin.y),
                c = sqr(ray_origin.x) + sqr(ray_origin.y) - sqr(p->radius);
    float t_near = 0.0f, t_far = 0.0f;
    if (!SolveQuadratic(a, b, c, &t_near, &t_far) || t_far < kEpsilonDistance) return 0;
    const float z_near = ray_origin.z + ray_direction.z * t_near, z_far = ray_origin.z + ray_direction.z * t_far;
    float t = 0;
    if (kEpsilonDistance < t_near && 0.0f <= z_near && z_near <= p->length) t = t_near;
    else if (0.0 <= z_far && z_far <= p->length) t = t_far;
    else return 0;
    const Vec3 position_local = add(ray_origin, smul(t, ray_direction));
    const Vec2 texcoord = {atan2f(position_local.y, position_local.x) * k1Div2Pi, position_local.z / p->length};
    if (bsdf != NULL && BsdfIsTransparent(bsdf, texcoord, seed)) return 0;
    const Vec3 position = TransformPoint(&p->to_world, position_local);
    t = length(sub(position, ray->origin));
    if (t > ray->t_max || t < ray->t_min) return 0;
    ray->t_max = t;
    if (hit != NULL) {
        const int inside = c < 0.0f;
        const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
        const Vec3 normal_local = normalize(v3(position_local.x, position_local.y, 0.0f));
        Vec3 normal = TransformVector(&normal_to_world, normal_local), tangent = TransformVector(&normal_to_world, v3(0, 0, 1)),
             bitangent = normalize(cross(normal, tangent));
        if (bsdf != NULL) {
            normal = ApplyBumpMapping(bsdf, normal, tangent, bitangent, texcoord);
            bitangent = normalize(cross(normal, tangent));
            tangent = normalize(cross(bitangent, normal));
        }
        if (inside) { normal = neg(normal); bitangent = neg(bitangent); }
        *hit = HitFull(p->id, inside, texcoord, position, normal, tangent, bitangent);
    }
    return 1;
}
static Hit SampleCylinder(const Primitive *p, float xi_0, float xi_1) { /* cylinder.cpp:92-105 */
    const float phi = k2Pi * xi_0, z = xi_1 * p->length;
    const Vec2 texcoord = {xi_0, xi_1};
    const Vec3 position = TransformPoint(&p->to_world, v3(cosf(phi) * p->radius, sinf(phi) * p->radius, z));
    const Mat4 tr = mat_transpose(&p->to_world), normal_to_world = mat_inverse(&tr);
    const Vec3 normal = TransformVector(&normal_to_world, v3(cosf(phi), sinf(phi), 0));
    return HitSample(p->id, texcoord, position, normal);
}

static int PrimitiveIntersect(const Scene *s, const Primitive *p, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* primitive.cpp:84-104 */
    switch (p->type) {
    case kPrimTriangle: return IntersectTriangle(s, p, bsdf, seed, ray, hit);
    case kPrimSphere: return IntersectSphere(p, bsdf, seed, ray, hit);
    case kPrimDisk: return IntersectDisk(p, bsdf, seed, ray, hit);
    case kPrimCylinder: return IntersectCylinder(p, bsdf, seed, ray, hit);
    }
    return 0;
}
static Hit PrimitiveSample(const Primitive *p, float xi_0, float xi_1) { /* primitive.cpp:106-122 */
    switch (p->type) {
    case kPrimTriangle: return SampleTriangle(p, xi_0, xi_1);
    case kPrimSphere: return SampleSphere(p, xi_0, xi_1);
    case kPrimDisk: return SampleDisk(p, xi_0, xi_1);
    case kPrimCylinder: return SampleCylinder(p, xi_0, xi_1);
    }
    return HitInvalid();
}

static void BlasIntersect(const Scene *s, const Instance *in, const Bsdf *bsdf, uint32_t *seed, Ray *ray, Hit *hit) { /* blas.cpp:18-47 */
    uint32_t stack[65];
    stack[0] = 0;
    int ptr = 0;
    while (ptr >= 0) {
        const BvhNode *node = in->nodes + stack[ptr];
        --ptr;
        while (AabbIntersect(&node->aabb, ray)) {
            if (node->leaf) {
                PrimitiveIntersect(s, in->primitives + node->id_object, bsdf, seed, ray, hit);
                break;
            } else {
                ++ptr;
                stack[ptr] = node->id_right;
                node = in->nodes + node->id_left;
            }
        }
    }
}
static int BlasIntersectAny(const Scene *s, const Instance *in, const Bsdf *bsdf, uint32_t *seed, Ray *ray) { /* blas.cpp:49-77 */
    uint32_t stack[65];
    stack[0] = 0;
    int ptr = 0;
    while (ptr >= 0) {
        const BvhNode *node = in->nodes + stack[ptr];
        --ptr;
        while (AabbIntersect(&node->aabb, ray)) {
            if (node->leaf) {
                if (PrimitiveIntersect(s, in->primitives + node->id_object, bsdf, seed, ray, NULL)) return 1;
                else break;
            } else {
                ++ptr;
                stack[ptr] = node->id_right;
                node = in->nodes + node->id_left;
            }
        }
    }
    return 0;
}
static Hit InstanceSample(const Instance *in, float xi_0, float xi_1, float xi_2) { /* instance.cpp:56-60, blas.cpp:79-98 */
    const BvhNode *node = in->nodes;
    float thresh = node->area * xi_0;
    while (!node->leaf) {
        if (thresh < in->nodes[node->id_left].area) {
            node = in->nodes + node->id_left;
        } else {
            thresh -= in->nodes[node->id_left].area;
            node = in->nodes + node->id_right;
        }
    }
    return PrimitiveSample(in->primitives + node->id_object, xi_1, xi_2);
}
static const Bsdf *InstanceBsdf(const Scene *s, uint32_t id) { return s->map_instance_bsdf[id] != kInvalidId ? s->bsdfs + s->map_instance_bsdf[id] : NULL; }

static Hit TlasIntersect(const Scene *s, uint32_t *seed, Ray *ray) { /* tlas.cpp:13-43, instance.cpp:25-44 */
    uint32_t stack[65];
    stack[0] = 0;
    int ptr = 0;
    Hit hit = HitInvalid();
    while (ptr >= 0) {
        const BvhNode *node = s->nodes + stack[ptr];
        --ptr;
        while (AabbIntersect(&node->aabb, ray)) {
            if (node->leaf) {
                const Instance *in = s->instances + node->id_object;
                Ray ray_local = *ray;
                Hit hit_local = HitInvalid();
                BlasIntersect(s, in, InstanceBsdf(s, in->id), seed, &ray_local, &hit_local);
                if (hit_local.valid && ray_local.t_max <= ray->t_max) {
                    *ray = ray_local;
                    hit = hit_local;
                    hit.id_instance = in->id;
                    hit.id_medium_int = in->id_medium_int;
                    hit.id_medium_ext = in->id_medium_ext;
                }
                break;
            } else {
                ++ptr;
                stack[ptr] = node->id_right;
                node = s->nodes + node->id_left;
            }
        }
    }
    return hit;
}
static int TlasIntersectAny(const Scene *s, uint32_t *seed, Ray *ray) { /* tlas.cpp:44-76, instance.cpp:46-54 */
    uint32_t stack[65];
    stack[0] = 0;
    int ptr = 0;
    while (ptr >= 0) {
        const BvhNode *node = s->nodes + stack[ptr];
        --ptr;
        while (AabbIntersect(&node->aabb, ray)) {
            if (node->leaf) {
                const Instance *in = s->instances + node->id_object;
                if (BlasIntersectAny(s, in, InstanceBsdf(s, in->id), seed, ray)) return 1;
                else break;
            } else {
                ++ptr;
                stack[ptr] = node->id_right;
                node = s->nodes + node->id_left;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Emitters                                                                                    */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int valid, harsh; float distance; Vec3 wi; } EmitterSampleRec; /* emitter.hpp:49-55 */

static EmitterSampleRec EmitterSample(const Emitter *e, Vec3 origin, float xi_0, float xi_1) { /* emitter.cpp:177-203 */
    EmitterSampleRec rec = {0, 1, kMaxFloat, {0, 0, 0}};
    switch (e->type) {
    case B200PT_EMIT_POINT: { /* point_light.cpp:8-19 */
        const Vec3 vec = sub(origin, e->position);
        rec.valid = 1, rec.harsh = 1, rec.distance = length(vec), rec.wi = normalize(vec);
        break;
    }
    case B200PT_EMIT_SPOT: { /* spot_light.cpp:8-24 */
        const Vec3 vec = sub(origin, e->position);
        const Vec3 wi = normalize(vec), dir_local = TransformVector(&e->to_local, wi);
        if (dir_local.z >= e->cos_cutoff_angle) { rec.valid = 1, rec.harsh = 1, rec.distance = length(vec), rec.wi = wi; }
        break;
    }
    case B200PT_EMIT_DIRECTIONAL: /* directional_light.cpp:8-18 */
        rec.valid = 1, rec.harsh = 1, rec.distance = kMaxFloat, rec.wi = e->direction;
        break;
    case B200PT_EMIT_SUN: { /* sun.cpp:8-18 */
        const Vec3 dir_local = SampleConeUniform(e->cos_cutoff_angle, xi_0, xi_1);
        rec.valid = 1, rec.harsh = 1, rec.distance = kMaxFloat, rec.wi = LocalToWorld(dir_local, e->direction);
        break;
    }
    case B200PT_EMIT_ENVMAP: { /* envmap.cpp:70-88 */
        uint32_t row = BinarySearch(e->height + 1, e->cdf_rows, xi_0) - 1;
        const float *cdf_col = e->cdf_cols + row * (e->width + 1);
        uint32_t col = BinarySearch(e->width + 1, cdf_col, xi_1) - 1;
        Vec3 vec_local = SphericalToCartesian(row * kPi / e->height, col * k2Pi / e->width, 1), vec = TransformVector(&e->to_world, vec_local);
        rec.valid = 1, rec.harsh = 0, rec.distance = kMaxFloat, rec.wi = vec;
        break;
    }
    case B200PT_EMIT_CONSTANT: /* constant_light.cpp:8-19 */
        rec.valid = 1, rec.harsh = 0, rec.distance = kMaxFloat, rec.wi = SampleSphereUniform(xi_0, xi_1);
        break;
    }
    return rec;
}
static Vec2 LatLong(Vec3 dir) {
    float phi = 0, theta = 0;
    CartesianToSpherical(dir, &theta, &phi, NULL);
    Vec2 t = {phi * k1Div2Pi, theta * k1DivPi};
    return t;
}
static Vec3 EmitterEvaluateRec(const Emitter *e, const EmitterSampleRec *rec) { /* emitter.cpp:205-231 */
    switch (e->type) {
    case B200PT_EMIT_POINT: return v3s(0);
    case B200PT_EMIT_SPOT: { /* spot_light.cpp:26-44 */
        const Vec3 dir = TransformVector(&e->to_local, rec->wi);
        Vec3 fall_off = {1.0f, 1.0f, 1.0f};
        if (e->texture != NULL) {
            const Vec2 texcoord = {0.5f + 0.5f * dir.x / (dir.z * e->uv_factor), 0.5f + 0.5f * dir.y / (dir.z * e->uv_factor)};
            fall_off = mul(fall_off, GetColor(e->texture, texcoord));
        }
        if (dir.z < e->cos_beam_width) fall_off = muls(fall_off, (e->cutoff_angle - acosf(dir.z)) * e->transition_width_rcp);
        return muls(mul(e->radiance, fall_off), sqr(1.0f / rec->distance));
    }
    case B200PT_EMIT_DIRECTIONAL:
    case B200PT_EMIT_SUN:
    case B200PT_EMIT_CONSTANT: return e->radiance;
    case B200PT_EMIT_ENVMAP: return GetColor(e->texture, LatLong(neg(TransformVector(&e->to_local, rec->wi)))); /* envmap.cpp:90-98 */
    }
    return v3s(0);
}
static Vec3 EmitterEvaluateDir(const Emitter *e, Vec3 look_dir) { /* emitter.cpp:233-249 */
    switch (e->type) {
    case B200PT_EMIT_SUN: return GetColor(e->texture, LatLong(look_dir));
    case B200PT_EMIT_ENVMAP: return GetColor(e->texture, LatLong(TransformVector(&e->to_local, look_dir)));
    case B200PT_EMIT_CONSTANT: return e->radiance;
    }
    return v3s(0);
}
static float EmitterPdf(const Emitter *e, Vec3 look_dir) { /* emitter.cpp:251-261, envmap.cpp:109-133 */
    if (e->type == B200PT_EMIT_CONSTANT) return k1Div4Pi;
    if (e->type != B200PT_EMIT_ENVMAP) return 0;
    const Vec3 dir = TransformVector(&e->to_local, look_dir);
    float phi = 0, theta = 0;
    CartesianToSpherical(dir, &theta, &phi, NULL);
    const Vec2 texcoord = {phi * k1Div2Pi, theta * k1DivPi};
    const Vec3 color = GetColor(e->texture, texcoord);
    const float lum = 0.2126f * color.x + 0.7152f * color.y + 0.0722f * color.z;
    const float row = fminf(fmaxf(texcoord.u * e->height, 0), e->height - 1);
    const int row_int = (int)row;
    const float t = row - row_int;
    if (t == 0) return lum * e->weight_rows[row_int] * e->normalization / fmaxf(fabsf(sinf(theta)), 1e-4f);
    return lum * lerpf(e->weight_rows[row_int], e->weight_rows[row_int + 1], t) * e->normalization / fmaxf(fabsf(sinf(theta)), 1e-4f);
}

/* ------------------------------------------------------------------------------------------- */
/* Media                                                                                       */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int valid, scattered; float pdf, distance; Vec3 attenuation; } MediumSampleRec; /* medium.hpp:54-61 */
typedef struct { int valid; float pdf; Vec3 wi, wo, attenuation; } PhaseSampleRec;             /* medium.hpp:26-33 */
static MediumSampleRec MediumRecInit(void) { MediumSampleRec r = {0, 0, 1.0f, 0, {1.0f, 1.0f, 1.0f}}; return r; }

static void MediumSample(const Medium *m, float max_distance, uint32_t *seed, MediumSampleRec *rec) { /* homogeneous.cpp:9-51 */
    float xi_0 = RandomFloat(seed);
    if (xi_0 < m->sampling_weight) {
        xi_0 /= m->sampling_weight;
        const int channel = (int)(RandomFloat(seed) * 3);
        rec->distance = -logf(1.0f - xi_0) / comp(m->sigma_t, channel);
        if (rec->distance < max_distance) {
            for (int dim = 0; dim < 3; ++dim) rec->pdf += comp(m->sigma_t, dim) * expf(-comp(m->sigma_t, dim) * rec->distance);
            rec->pdf *= m->sampling_weight * (1.0f / 3.0f);
            rec->scattered = 1;
        }
    }
    if (!rec->scattered) {
        rec->distance = max_distance;
        rec->pdf = 0;
        for (int dim = 0; dim < 3; ++dim) rec->pdf += expf(-comp(m->sigma_t, dim) * rec->distance);
        rec->pdf = m->sampling_weight * (1.0f / 3.0f) * rec->pdf + (1.0f - m->sampling_weight);
    }
    for (int dim = 0; dim < 3; ++dim) {
        setcomp(&rec->attenuation, dim, expf(-comp(m->sigma_t, dim) * rec->distance));
        if (comp(rec->attenuation, dim) > kEpsilonFloat) rec->valid = 1;
    }
    if (rec->scattered) rec->attenuation = mul(rec->attenuation, m->sigma_s);
}
static void MediumEvaluate(const Medium *m, MediumSampleRec *rec) { /* homogeneous.cpp:53-82 */
    for (int dim = 0; dim < 3; ++dim) {
        setcomp(&rec->attenuation, dim, expf(-comp(m->sigma_t, dim) * rec->distance));
        if (comp(rec->attenuation, dim) > kEpsilonFloat) rec->valid = 1;
    }
    if (!rec->valid) return;
    if (rec->scattered) {
        for (int dim = 0; dim < 3; ++dim) rec->pdf += comp(m->sigma_t, dim) * comp(rec->attenuation, dim);
        rec->pdf *= m->sampling_weight * (1.0f / 3.0f);
        rec->attenuation = mul(rec->attenuation, m->sigma_s);
    } else {
        for (int dim = 0; dim < 3; ++dim) rec->pdf += comp(rec->attenuation, dim);
        rec->pdf = m->sampling_weight * (1.0f / 3.0f) * rec->pdf + (1.0f - m->sampling_weight);
    }
}
static void HgValue(Vec3 g, float cos_theta, PhaseSampleRec *rec) { /* henyey_greenstein.cpp:28-33, 49-55 */
    const Vec3 temp = add(sadd(1.0f, sqr3(g)), smul(2.0f * cos_theta, g));
    rec->attenuation = vdiv(smul(k1Div4Pi, ssub(1.0f, sqr3(g))), mul(temp, vsqrt(temp)));
    rec->pdf = 0;
    for (int dim = 0; dim < 3; ++dim) rec->pdf += comp(rec->attenuation, dim);
    rec->pdf *= (1.0f / 3.0f);
}
static void PhaseSample(const Medium *m, uint32_t *seed, PhaseSampleRec *rec) { /* medium.cpp:76-87 */
    if (m->phase_type == B200PT_PHASE_ISOTROPIC) { /* isotropic.cpp:9-15 */
        rec->valid = 1;
        rec->attenuation = v3s(k1Div4Pi);
        rec->pdf = k1Div4Pi;
        const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
        rec->wi = SampleSphereUniform(xi_0, xi_1);
        return;
    }
    /* henyey_greenstein.cpp:9-43 */
    const Vec3 g = m->g;
    const int channel = (int)(RandomFloat(seed) * 3);
    float cos_theta = 0;
    if (fabsf(comp(g, channel)) < kEpsilonFloat) {
        cos_theta = 1.0f - 2.0f * RandomFloat(seed);
    } else {
        const float gc = comp(g, channel);
        const float sqr_term = (1.0f - sqr(gc)) / (1.0f - gc + 2.0f * gc * RandomFloat(seed));
        cos_theta = (1.0f + sqr(gc) - sqr(sqr_term)) / (2.0f * gc);
    }
    HgValue(g, cos_theta, rec);
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
    const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - sqr(cos_theta)));
    const float phi = k2Pi * RandomFloat(seed);
    rec->wi = v3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
    rec->wi = neg(LocalToWorld(rec->wi, rec->wo));
}
static void PhaseEvaluate(const Medium *m, PhaseSampleRec *rec) { /* medium.cpp:63-74 */
    if (m->phase_type == B200PT_PHASE_ISOTROPIC) { /* isotropic.cpp:17-22 */
        rec->valid = 1;
        rec->attenuation = v3s(k1Div4Pi);
        rec->pdf = k1Div4Pi;
        return;
    }
    HgValue(m->g, dot(neg(rec->wi), rec->wo), rec); /* henyey_greenstein.cpp:45-60 */
    if (rec->pdf < kEpsilon) return;
    rec->valid = 1;
}

/* ------------------------------------------------------------------------------------------- */
/* Integrators                                                                                 */
/* ------------------------------------------------------------------------------------------- */
static BsdfSampleRec RecInit(void) { BsdfSampleRec r; memset(&r, 0, sizeof(r)); return r; }

static BsdfSampleRec EvaluateRayPath(const Scene *s, Vec3 wi, Vec3 wo, const Hit *hit, const Bsdf *bsdf) { /* path.cpp:238-266 */
    BsdfSampleRec rec = RecInit();
    rec.wi = wi, rec.wo = wo, rec.texcoord = hit->texcoord, rec.position = hit->position;
    if (bsdf) {
        rec.inside = hit->inside, rec.normal = hit->normal, rec.tangent = hit->tangent, rec.bitangent = hit->bitangent;
        if (dot(neg(wi), hit->normal) < 0.0f) { rec.inside = !rec.inside; rec.normal = neg(rec.normal); }
        BsdfEvaluate(s, bsdf, &rec);
    } else {
        rec.pdf = 1, rec.attenuation = v3s(1), rec.valid = 1;
    }
    return rec;
}
static BsdfSampleRec SampleRayPath(const Scene *s, Vec3 wo, const Hit *hit, const Bsdf *bsdf, uint32_t *seed) { /* path.cpp:268-296 */
    BsdfSampleRec rec = RecInit();
    rec.wo = wo, rec.texcoord = hit->texcoord, rec.position = hit->position;
    if (bsdf != NULL) {
        rec.inside = hit->inside, rec.normal = hit->normal, rec.tangent = hit->tangent, rec.bitangent = hit->bitangent;
        if (dot(wo, hit->normal) < 0.0f) { rec.inside = !rec.inside; rec.normal = neg(rec.normal); }
        BsdfSample(s, bsdf, seed, &rec);
    } else {
        rec.wi = wo, rec.pdf = 1.0f, rec.attenuation = v3s(1.0f), rec.valid = 1;
    }
    return rec;
}

/* Shared body of EvaluateDirectLightPath (path.cpp:138-236) and the two EvaluateDirectLightVolPath
 * overloads (volpath.cpp:247-375 surface, :377-485 medium).  `medium` = medium used for the shadow
 * segment transmittance (NULL: none), `phase_medium` != NULL selects the medium-vertex variant. */
static Vec3 EvaluateDirectLight(const Scene *s, const Hit *hit, Vec3 position, Vec3 wo, uint32_t *seed, int volpath,
                                const Medium *medium, const Medium *phase_medium) {
    Vec3 L = v3s(0);
    for (uint32_t i = 0; i < s->num_emitter; ++i) {
        const Emitter *emitter = s->emitters + i;
        const float xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
        EmitterSampleRec rec = EmitterSample(emitter, position, xi_0, xi_1);
        Ray ray_test = MakeRay(s, position, neg(rec.wi));
        ray_test.t_max = rec.distance - kEpsilonDistance;
        if (TlasIntersectAny(s, seed, &ray_test)) continue;
        if (phase_medium == NULL && dot(neg(rec.wi), hit->normal) < kEpsilonFloat) continue;
        Vec3 medium_attenuation = {1.0f, 1.0f, 1.0f};
        if (volpath && medium != NULL) {
            MediumSampleRec medium_rec = MediumRecInit();
            medium_rec.distance = rec.distance;
            MediumEvaluate(medium, &medium_rec);
            if (!medium_rec.valid) continue;
            medium_attenuation = divs(medium_rec.attenuation, medium_rec.pdf);
        }
        Vec3 f;
        float pdf_f;
        if (phase_medium != NULL) {
            PhaseSampleRec phase_rec;
            memset(&phase_rec, 0, sizeof(phase_rec));
            phase_rec.wi = rec.wi, phase_rec.wo = wo;
            PhaseEvaluate(phase_medium, &phase_rec);
            if (!phase_rec.valid) continue;
            f = phase_rec.attenuation, pdf_f = phase_rec.pdf;
        } else {
            const BsdfSampleRec rec1 = EvaluateRayPath(s, rec.wi, wo, hit, InstanceBsdf(s, hit->id_instance));
            if (!rec1.valid) continue;
            f = rec1.attenuation, pdf_f = rec1.pdf;
        }
        const Vec3 radiance = EmitterEvaluateRec(emitter, &rec);
        if (rec.harsh) {
            if (volpath) L = add(L, mul(mul(radiance, medium_attenuation), f));
            else L = add(L, mul(radiance, f));
        } else {
            const float pdf_direct = EmitterPdf(emitter, neg(rec.wi));
            if (pdf_direct > kEpsilonFloat) {
                const float weight_direct = MisWeight(pdf_direct, pdf_f);
                if (volpath) L = add(L, divs(mul(mul(smul(weight_direct, radiance), medium_attenuation), f), pdf_direct));
                else L = add(L, mul(smul(weight_direct, radiance), divs(f, pdf_direct)));
            }
        }
    }
    if (s->num_area_light != 0) {
        const uint32_t index_area_light = BinarySearch(s->size_cdf_area_light, s->cdf_area_light, RandomFloat(seed)) - 1,
                       id_area_light_instance = s->map_id_area_light_instance[index_area_light];
        const float xi_2 = RandomFloat(seed), xi_1 = RandomFloat(seed), xi_0 = RandomFloat(seed);
        const Hit hit_pre = InstanceSample(s->instances + id_area_light_instance, xi_0, xi_1, xi_2);
        const Vec3 d_vec = sub(position, hit_pre.position);
        const float distance = length(d_vec);
        Ray ray_test = MakeRay(s, hit_pre.position, normalize(d_vec));
        ray_test.t_max = distance - kEpsilonDistance;
        if (TlasIntersectAny(s, seed, &ray_test)) return L;
        const Vec3 wi = normalize(d_vec);
        const float cos_theta_prime = dot(wi, hit_pre.normal);
        if (cos_theta_prime < kEpsilonFloat) return L;
        if (phase_medium == NULL && dot(neg(wi), hit->normal) < kEpsilonFloat) return L;
        Vec3 medium_attenuation = {1.0f, 1.0f, 1.0f};
        if (volpath && medium != NULL) {
            MediumSampleRec medium_rec = MediumRecInit();
            medium_rec.distance = distance;
            MediumEvaluate(medium, &medium_rec);
            if (!medium_rec.valid) return L;
            medium_attenuation = divs(medium_rec.attenuation, medium_rec.pdf);
        }
        Vec3 f;
        float pdf_f;
        if (phase_medium != NULL) {
            PhaseSampleRec phase_rec;
            memset(&phase_rec, 0, sizeof(phase_rec));
            phase_rec.wi = wi, phase_rec.wo = wo;
            PhaseEvaluate(phase_medium, &phase_rec);
            if (!phase_rec.valid) return L;
            f = phase_rec.attenuation, pdf_f = phase_rec.pdf;
        } else {
            const BsdfSampleRec rec = EvaluateRayPath(s, wi, wo, hit, InstanceBsdf(s, hit->id_instance));
            if (!rec.valid) return L;
            f = rec.attenuation, pdf_f = rec.pdf;
        }
        const float pdf_area = (s->cdf_area_light[index_area_light + 1] - s->cdf_area_light[index_area_light]) *
                               s->list_pdf_area_instance[id_area_light_instance],
                    pdf_direct = pdf_area * sqr(distance) / cos_theta_prime, weight_direct = MisWeight(pdf_direct, pdf_f);
        const Bsdf *bsdf_pre = s->bsdfs + s->map_instance_bsdf[id_area_light_instance];
        const Vec3 radiance = GetRadiance(bsdf_pre, hit_pre.texcoord);
        if (volpath) L = add(L, smul(weight_direct, divs(mul(mul(radiance, medium_attenuation), f), pdf_direct)));
        else L = add(L, mul(smul(weight_direct, radiance), divs(f, pdf_direct)));
    }
    return L;
}

/* Debugging aid for the exact-mode comparison (tools/replay_trace.py): ORACLE_TRACE_PIXEL="i,j" prints, for that pixel, the
 * state of every path vertex (LCG state, hit distance, radiance so far, throughput) to stderr. */
static __thread int g_trace = 0;
static void TraceVertex(uint32_t depth, uint32_t seed, float t, Vec3 L, Vec3 att) {
    if (g_trace) fprintf(stderr, "[trace] v%u state %08x t %.9g L %.9g %.9g %.9g att %.9g %.9g %.9g\n", depth, seed, t, L.x, L.y, L.z, att.x, att.y, att.z);
}

static Vec3 ShadePath(const Scene *s, Vec3 eye, Vec3 look_dir, uint32_t *seed) { /* path.cpp:8-136 */
    Vec3 L = v3s(0);
    Ray ray = MakeRay(s, eye, look_dir);
    Hit hit = HitInvalid();
    if (s->has_tlas) hit = TlasIntersect(s, seed, &ray);
    if (!hit.valid) {
        if (s->id_envmap != kInvalidId) L = add(L, EmitterEvaluateDir(s->emitters + s->id_envmap, look_dir));
        if (s->id_sun != kInvalidId) L = add(L, EmitterEvaluateDir(s->emitters + s->id_sun, look_dir));
        return L;
    }
    const Bsdf *bsdf = InstanceBsdf(s, hit.id_instance);
    if (bsdf != NULL) {
        if (hit.inside && !bsdf->twosided) return v3s(0);
        else if (bsdf->type == B200PT_BSDF_AREA_LIGHT) {
            if (s->hide_emitters) return v3s(0);
            else return GetRadiance(bsdf, hit.texcoord);
        }
    }
    Vec3 attenuation = v3s(1), wo = neg(look_dir);
    TraceVertex(1, *seed, ray.t_max, L, attenuation);
    for (uint32_t depth = 1; depth < s->depth_rr || (depth < s->depth_max && RandomFloat(seed) < s->pdf_rr); ++depth) {
        L = add(L, mul(attenuation, EvaluateDirectLight(s, &hit, hit.position, wo, seed, 0, NULL, NULL)));
        BsdfSampleRec rec = SampleRayPath(s, wo, &hit, bsdf, seed);
        if (!rec.valid) break;
        attenuation = mul(attenuation, divs(rec.attenuation, rec.pdf));
        if (fmaxf(fmaxf(attenuation.x, attenuation.y), attenuation.z) < kEpsilon) break;
        ray = MakeRay(s, rec.position, neg(rec.wi));
        hit = TlasIntersect(s, seed, &ray);
        if (!hit.valid) {
            if (s->id_envmap != kInvalidId) {
                const Vec3 radiance = EmitterEvaluateDir(s->emitters + s->id_envmap, neg(rec.wi));
                const float pdf_direct = EmitterPdf(s->emitters + s->id_envmap, neg(rec.wi)), weight_bsdf = MisWeight(rec.pdf, pdf_direct);
                L = add(L, mul(smul(weight_bsdf, attenuation), radiance));
            }
            break;
        }
        bsdf = InstanceBsdf(s, hit.id_instance);
        if (bsdf != NULL) {
            if (hit.inside && !bsdf->twosided) break;
            else if (bsdf->type == B200PT_BSDF_AREA_LIGHT) {
                const float cos_theta_prime = dot(rec.wi, hit.normal);
                if (cos_theta_prime < kEpsilonFloat) break;
                const uint32_t id_instance_area_light = s->map_id_instance_area_light[hit.id_instance];
                const float pdf_area = (s->cdf_area_light[id_instance_area_light + 1] - s->cdf_area_light[id_instance_area_light]) *
                                       s->list_pdf_area_instance[hit.id_instance],
                            pdf_direct = pdf_area * sqr(ray.t_max) / cos_theta_prime, weight_bsdf = MisWeight(rec.pdf, pdf_direct);
                const Vec3 radiance = GetRadiance(bsdf, hit.texcoord), L_dir = mul(smul(weight_bsdf, attenuation), radiance);
                L = add(L, L_dir);
                break;
            }
        }
        wo = rec.wi;
        TraceVertex(depth + 1, *seed, ray.t_max, L, attenuation);
        if (depth >= s->depth_rr) attenuation = muls(attenuation, s->pdf_rr_rcp);
    }
    return L;
}

static const Medium *HitMedium(const Scene *s, const Hit *hit, Vec3 w) { /* volpath.cpp:44-45, 163-166, 252-256 */
    const int inside = dot(w, hit->normal) > 0 ? hit->inside : !hit->inside;
    const uint32_t id_medium = inside ? hit->id_medium_int : hit->id_medium_ext;
    return id_medium != kInvalidId ? s->media + id_medium : NULL;
}

static Vec3 ShadeVolPath(const Scene *s, Vec3 eye, Vec3 look_dir, uint32_t *seed) { /* volpath.cpp:8-245 */
    Vec3 L = v3s(0);
    Ray ray = MakeRay(s, eye, look_dir);
    Hit hit = HitInvalid();
    if (s->has_tlas) hit = TlasIntersect(s, seed, &ray);
    if (!hit.valid) {
        if (s->id_envmap != kInvalidId) L = add(L, EmitterEvaluateDir(s->emitters + s->id_envmap, look_dir));
        if (s->id_sun != kInvalidId) L = add(L, EmitterEvaluateDir(s->emitters + s->id_sun, look_dir));
        return L;
    }
    Vec3 attenuation = v3s(1), wo = neg(look_dir);
    int scattering = 0;
    Vec3 medium_hit_position = v3s(0);
    const Medium *medium_hit_medium = NULL;
    {
        const Medium *medium = HitMedium(s, &hit, wo);
        if (medium != NULL) {
            MediumSampleRec medium_rec = MediumRecInit();
            MediumSample(medium, ray.t_max, seed, &medium_rec);
            if (medium_rec.valid) {
                attenuation = mul(attenuation, divs(medium_rec.attenuation, medium_rec.pdf));
                if (medium_rec.scattered) {
                    scattering = 1;
                    medium_hit_position = add(ray.origin, muls(ray.dir, medium_rec.distance));
                    medium_hit_medium = medium;
                }
            }
        }
    }
    const Bsdf *bsdf = NULL;
    if (!scattering) {
        bsdf = InstanceBsdf(s, hit.id_instance);
        if (bsdf != NULL) {
            if (hit.inside && !bsdf->twosided) return v3s(0);
            else if (bsdf->type == B200PT_BSDF_AREA_LIGHT) {
                if (s->hide_emitters) return v3s(0);
                else return GetRadiance(bsdf, hit.texcoord);
            }
        }
    }
    Vec3 wi = v3s(0);
    float pdf_sample = 0;
    TraceVertex(1, *seed, ray.t_max, L, attenuation);
    for (uint32_t depth = 1; depth < s->depth_rr || (depth < s->depth_max && RandomFloat(seed) < s->pdf_rr); ++depth) {
        if (scattering) {
            L = add(L, mul(attenuation, EvaluateDirectLight(s, NULL, medium_hit_position, wo, seed, 1, medium_hit_medium, medium_hit_medium)));
            PhaseSampleRec phase_rec;
            memset(&phase_rec, 0, sizeof(phase_rec));
            phase_rec.wo = wo;
            PhaseSample(medium_hit_medium, seed, &phase_rec);
            if (!phase_rec.valid) break;
            wi = phase_rec.wi;
            attenuation = mul(attenuation, divs(phase_rec.attenuation, phase_rec.pdf));
            pdf_sample = phase_rec.pdf;
            if (fmaxf(fmaxf(attenuation.x, attenuation.y), attenuation.z) < kEpsilon) break;
            ray = MakeRay(s, medium_hit_position, neg(wi));
            hit = TlasIntersect(s, seed, &ray);
            MediumSampleRec medium_rec = MediumRecInit();
            MediumSample(medium_hit_medium, ray.t_max, seed, &medium_rec);
            if (medium_rec.valid) {
                attenuation = mul(attenuation, divs(medium_rec.attenuation, medium_rec.pdf));
                if (medium_rec.scattered) {
                    scattering = 1;
                    medium_hit_position = add(ray.origin, muls(ray.dir, medium_rec.distance));
                } else {
                    scattering = 0;
                }
            } else {
                scattering = 0;
            }
        } else {
            L = add(L, mul(attenuation, EvaluateDirectLight(s, &hit, hit.position, wo, seed, 1, HitMedium(s, &hit, wo), NULL)));
            BsdfSampleRec rec = SampleRayPath(s, wo, &hit, bsdf, seed);
            if (!rec.valid) break;
            wi = rec.wi;
            pdf_sample = rec.pdf;
            attenuation = mul(attenuation, divs(rec.attenuation, pdf_sample));
            if (fmaxf(fmaxf(attenuation.x, attenuation.y), attenuation.z) < kEpsilon) break;
            ray = MakeRay(s, rec.position, neg(wi));
            hit = TlasIntersect(s, seed, &ray);
            const Medium *medium = HitMedium(s, &hit, wi);
            if (medium != NULL) {
                MediumSampleRec medium_rec = MediumRecInit();
                MediumSample(medium, ray.t_max, seed, &medium_rec);
                if (medium_rec.valid) {
                    attenuation = mul(attenuation, divs(medium_rec.attenuation, medium_rec.pdf));
                    if (medium_rec.scattered) {
                        scattering = 1;
                        medium_hit_position = add(ray.origin, muls(ray.dir, medium_rec.distance));
                        medium_hit_medium = medium;
                    }
                }
            }
        }
        if (!scattering) {
            if (!hit.valid) {
                if (s->id_envmap != kInvalidId) {
                    const Vec3 radiance = EmitterEvaluateDir(s->emitters + s->id_envmap, neg(wi));
                    const float pdf_direct = EmitterPdf(s->emitters + s->id_envmap, neg(wi)), weight_bsdf = MisWeight(pdf_sample, pdf_direct);
                    L = add(L, mul(smul(weight_bsdf, attenuation), radiance));
                }
                break;
            }
            bsdf = InstanceBsdf(s, hit.id_instance);
            if (bsdf != NULL) {
                if (hit.inside && !bsdf->twosided) break;
                else if (bsdf->type == B200PT_BSDF_AREA_LIGHT) {
                    const float cos_theta_prime = dot(wi, hit.normal);
                    if (cos_theta_prime < kEpsilonFloat) break;
                    const uint32_t id_instance_area_light = s->map_id_instance_area_light[hit.id_instance];
                    const float pdf_area = (s->cdf_area_light[id_instance_area_light + 1] - s->cdf_area_light[id_instance_area_light]) *
                                           s->list_pdf_area_instance[hit.id_instance],
                                pdf_direct = pdf_area * sqr(ray.t_max) / cos_theta_prime, weight_bsdf = MisWeight(pdf_sample, pdf_direct);
                    const Vec3 radiance = GetRadiance(bsdf, hit.texcoord), L_dir = mul(smul(weight_bsdf, attenuation), radiance);
                    L = add(L, L_dir);
                    break;
                }
            }
            wo = wi;
            if (depth >= s->depth_rr) attenuation = muls(attenuation, s->pdf_rr_rcp);
        }
        TraceVertex(depth + 1, *seed, ray.t_max, L, attenuation);
    }
    return L;
}

/* renderer.cpp:62-85 */
static void DrawPixel(const Scene *s, uint32_t i, uint32_t j, float *frame) {
    const uint32_t pixel_offset = (j * s->width + i) * 3;
    uint32_t seed = oracle_tea4(pixel_offset, 0);
    Vec3 color = v3s(0), temp;
    int trace_i = -1, trace_j = -1;
    const char *trace_env = getenv("ORACLE_TRACE_PIXEL");
    if (trace_env != NULL) sscanf(trace_env, "%d,%d", &trace_i, &trace_j);
    g_trace = (int)i == trace_i && (int)j == trace_j;
    for (uint32_t k = 0; k < s->spp; ++k) {
        if (g_trace) fprintf(stderr, "[trace] sample %u\n", k);
        const float u = k * s->spp_inv, v = oracle_van_der_corput2(k + 1), x = 2.0f * (i + u) / s->width - 1.0f,
                    y = 1.0f - 2.0f * (j + v) / s->height;
        const Vec3 look_dir = normalize(add(add(s->front, smul(x, s->view_dx)), smul(y, s->view_dy)));
        temp = s->integrator_type == B200PT_INTEGRATOR_VOLPATH ? ShadeVolPath(s, s->eye, look_dir, &seed) : ShadePath(s, s->eye, look_dir, &seed);
        temp.x = fminf(temp.x, 1.0f);
        temp.y = fminf(temp.y, 1.0f);
        temp.z = fminf(temp.z, 1.0f);
        color = add(color, temp);
    }
    color = muls(color, s->spp_inv);
    frame[pixel_offset] = color.x, frame[pixel_offset + 1] = color.y, frame[pixel_offset + 2] = color.z;
}

/* ------------------------------------------------------------------------------------------- */
/* LBVH (bvh_builder.cpp)                                                                      */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    const AABB *aabbs;
    const float *areas;
    uint32_t *map_id;
    uint64_t *mortons;
    BvhNode *nodes;
    uint32_t num_nodes;
} Builder;

static uint32_t ExpandBits(uint32_t v) { /* :16-22 */
    v = (v * ((0x1ul << 16) + 1)) & 0xFF0000FFul;
    v = (v * ((0x1ul << 8) + 1)) & 0x0F00F00Ful;
    v = (v * ((0x1ul << 4) + 1)) & 0xC30C30C3ul;
    v = (v * ((0x1ul << 2) + 1)) & 0x49249249ul;
    return v;
}
static int LeadingZeros64(uint64_t n) { /* :24-35 */
    int count = 0;
    for (int i = 0; i < 64; ++i) {
        if ((n >> (63 - i)) & 0x1) break;
        else ++count;
    }
    return count;
}
static uint32_t GetMorton3D(Vec3 v) { /* :39-48 */
    const float x = fminf(fmaxf(v.x * 1024.0f, 0.0f), 1023.0f), y = fminf(fmaxf(v.y * 1024.0f, 0.0f), 1023.0f),
                z = fminf(fmaxf(v.z * 1024.0f, 0.0f), 1023.0f);
    const uint32_t xx = ExpandBits((uint32_t)x), yy = ExpandBits((uint32_t)y), zz = ExpandBits((uint32_t)z);
    return xx * 4 + yy * 2 + zz;
}
static const uint64_t *g_sort_keys;
static int CompareByMorton(const void *a, const void *b) {
    const uint64_t ka = g_sort_keys[*(const uint32_t *)a], kb = g_sort_keys[*(const uint32_t *)b];
    return ka < kb ? -1 : (ka > kb ? 1 : 0);
}
static uint32_t FindSplit(const Builder *b, uint32_t first, uint32_t last) { /* :172-206 */
    const uint64_t first_code = b->mortons[b->map_id[first]], last_code = b->mortons[b->map_id[last - 1]];
    if (first_code == last_code) return (first + last) >> 1;
    const int common_prefix = LeadingZeros64(first_code ^ last_code);
    uint32_t split = first, step = last - first;
    do {
        step = (step + 1) >> 1;
        uint32_t new_split = split + step;
        if (new_split < last) {
            const uint64_t split_code = b->mortons[b->map_id[new_split]];
            const int split_prefix = LeadingZeros64(first_code ^ split_code);
            if (split_prefix > common_prefix) split = new_split;
        }
    } while (step > 1);
    return split;
}
static uint32_t BuildTopDown(Builder *b, uint32_t begin, uint32_t end) { /* :143-170 */
    const uint32_t id_node = b->num_nodes;
    if (begin + 1 > end) return kInvalidId;
    BvhNode *node = &b->nodes[b->num_nodes++];
    memset(node, 0, sizeof(*node));
    node->id = id_node, node->id_left = node->id_right = node->id_object = kInvalidId;
    if (begin + 1 == end) {
        node->leaf = 1, node->id_object = b->map_id[begin], node->aabb = b->aabbs[b->map_id[begin]], node->area = b->areas[b->map_id[begin]];
        return id_node;
    }
    node->leaf = 0, node->aabb = aabb_empty(), node->area = 0;
    const uint32_t middle = FindSplit(b, begin, end) + 1;
    const uint32_t left = BuildTopDown(b, begin, middle), right = BuildTopDown(b, middle, end);
    node = &b->nodes[id_node];
    node->id_left = left, node->id_right = right;
    node->area = b->nodes[left].area + b->nodes[right].area;
    node->aabb.min_ = vmin(b->nodes[left].aabb.min_, b->nodes[right].aabb.min_);
    node->aabb.max_ = vmax(b->nodes[left].aabb.max_, b->nodes[right].aabb.max_);
    return id_node;
}
/* Builds 2n-1 nodes into out_nodes (caller-allocated); returns the node count. :92-141 */
static uint32_t BuildLinearBvh(uint32_t n, const AABB *aabbs, const float *areas, BvhNode *out_nodes) {
    Builder b;
    b.aabbs = aabbs, b.areas = areas, b.nodes = out_nodes, b.num_nodes = 0;
    b.map_id = (uint32_t *)malloc(sizeof(uint32_t) * n);
    b.mortons = (uint64_t *)malloc(sizeof(uint64_t) * n);
    for (uint32_t i = 0; i < n; ++i) b.map_id[i] = i;
    AABB all = aabb_empty();
    for (uint32_t i = 0; i < n; ++i) aabb_add(&all, &aabbs[i]);
    const Vec3 aabb_size = sub(all.max_, all.min_);
    for (uint32_t i = 0; i < n; ++i) {
        const Vec3 center = muls(add(aabbs[i].min_, aabbs[i].max_), 0.5f);
        const Vec3 position_relative = vdiv(sub(center, all.min_), aabb_size);
        b.mortons[i] = GetMorton3D(position_relative);
        b.mortons[i] = (b.mortons[i] << 32) | (uint64_t)i;
    }
    g_sort_keys = b.mortons; /* keys are unique (morton << 32 | index), so any comparison sort gives std::sort's order */
    qsort(b.map_id, n, sizeof(uint32_t), CompareByMorton);
    BuildTopDown(&b, 0, n);
    free(b.map_id);
    free(b.mortons);
    return b.num_nodes;
}
uint32_t oracle_build_bvh(uint32_t n, const float *aabb_min_max, const float *areas, uint32_t *out_nodes, float *out_area, uint32_t capacity) {
    AABB *aabbs = (AABB *)malloc(sizeof(AABB) * n);
    for (uint32_t i = 0; i < n; ++i) {
        aabbs[i].min_ = v3(aabb_min_max[6 * i], aabb_min_max[6 * i + 1], aabb_min_max[6 * i + 2]);
        aabbs[i].max_ = v3(aabb_min_max[6 * i + 3], aabb_min_max[6 * i + 4], aabb_min_max[6 * i + 5]);
    }
    BvhNode *nodes = (BvhNode *)malloc(sizeof(BvhNode) * (2 * (size_t)n));
    const uint32_t count = BuildLinearBvh(n, aabbs, areas, nodes);
    for (uint32_t i = 0; i < count && i < capacity; ++i) {
        out_nodes[4 * i] = nodes[i].leaf ? 1u : 0u, out_nodes[4 * i + 1] = nodes[i].id_left, out_nodes[4 * i + 2] = nodes[i].id_right,
        out_nodes[4 * i + 3] = nodes[i].id_object;
        out_area[i] = nodes[i].area;
    }
    free(nodes);
    free(aabbs);
    return count;
}

/* ------------------------------------------------------------------------------------------- */
/* Scene commit (scene.cpp, renderer.cpp:259-676, bsdf.cpp:112-186, emitter.cpp:122-175, ...)  */
/* ------------------------------------------------------------------------------------------- */
static float AverageFresnelDielectric(float eta) { /* bsdf.cpp:12-39 */
    if (eta < 1.0) return -1.4399f * sqr(eta) + 0.7099f * eta + 0.6681f + 0.0636f / eta;
    float inv_eta = 1.0f / eta, inv_eta_2 = inv_eta * inv_eta, inv_eta_3 = inv_eta_2 * inv_eta, inv_eta_4 = inv_eta_3 * inv_eta,
          inv_eta_5 = inv_eta_4 * inv_eta;
    return 0.919317f - 3.4793f * inv_eta + 6.75335f * inv_eta_2 - 7.80989f * inv_eta_3 + 4.98554f * inv_eta_4 - 1.36881f * inv_eta_5;
}
static Vec3 AverageFresnelConductor(Vec3 r, Vec3 e) { /* bsdf.cpp:41-54 */
    Vec3 acc = add(v3s(0.087237f), smul(0.0230685f, e));
    acc = sub(acc, mul(smul(0.0864902f, e), e));
    acc = add(acc, mul(mul(smul(0.0774594f, e), e), e));
    acc = add(acc, smul(0.782654f, r));
    acc = sub(acc, mul(smul(0.136432f, r), r));
    acc = add(acc, mul(mul(smul(0.278708f, r), r), r));
    acc = add(acc, mul(smul(0.19744f, e), r));
    acc = add(acc, mul(mul(smul(0.0360605f, e), e), r));
    acc = sub(acc, mul(mul(smul(0.2586f, e), r), r));
    return acc;
}

typedef struct {
    Primitive *prims;
    BvhNode *nodes;
    uint32_t num_prims, num_nodes;
} InstanceGeometry;

static void SetupMeshTriangles(const b200pt_scene_desc *d, const b200pt_instance *in, int builtin, InstanceGeometry *g, float **areas_out, AABB **aabbs_out) {
    /* scene.cpp:200-245 built-in rectangle / cube */
    static const float rect_uv[] = {0, 0, 1, 0, 1, 1, 0, 1}, rect_pos[] = {-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0},
                       rect_nrm[] = {0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1};
    static const uint32_t rect_idx[] = {0, 1, 2, 2, 3, 0};
    static const float cube_uv[] = {0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0,
                                    0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0};
    static const float cube_pos[] = {1,  -1, -1, 1,  -1, 1,  -1, -1, 1,  -1, -1, -1, 1,  1,  -1, -1, 1,  -1, -1, 1,  1,  1,  1,  1,
                                     1,  -1, -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1,  1,  -1, -1, 1,
                                     -1, -1, 1,  -1, 1,  1,  -1, 1,  -1, -1, -1, -1, 1,  1,  -1, 1,  -1, -1, -1, -1, -1, -1, 1,  -1};
    static const float cube_nrm[] = {0,  -1, 0, 0,  -1, 0, 0,  -1, 0, 0,  -1, 0, 0, 1, 0,  0, 1, 0,  0, 1, 0,  0, 1, 0,
                                     1,  0,  0, 1,  0,  0, 1,  0,  0, 1,  0,  0, 0, 0, 1,  0, 0, 1,  0, 0, 1,  0, 0, 1,
                                     -1, 0,  0, -1, 0,  0, -1, 0,  0, -1, 0,  0, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1};
    static const uint32_t cube_idx[] = {0,  1,  2,  3,  0,  2,  4,  5,  6,  7,  4,  6,  8,  9,  10, 11, 8,  10,
                                        12, 13, 14, 15, 12, 14, 16, 17, 18, 19, 16, 18, 20, 21, 22, 23, 20, 22};
    const float *positions = NULL, *normals = NULL, *texcoords = NULL, *tangents = NULL, *bitangents = NULL;
    const uint32_t *indices = NULL;
    uint64_t nv = 0, nt = 0;
    if (builtin == B200PT_INST_RECTANGLE) {
        positions = rect_pos, normals = rect_nrm, texcoords = rect_uv, indices = rect_idx, nv = 4, nt = 2;
    } else if (builtin == B200PT_INST_CUBE) {
        positions = cube_pos, normals = cube_nrm, texcoords = cube_uv, indices = cube_idx, nv = 24, nt = 12;
    } else {
        nv = in->num_vertices, nt = in->num_triangles;
        if (in->position_offset != B200PT_NO_OFFSET) positions = d->positions + 3 * in->position_offset;
        if (in->normal_offset != B200PT_NO_OFFSET) normals = d->normals + 3 * in->normal_offset;
        if (in->texcoord_offset != B200PT_NO_OFFSET) texcoords = d->texcoords + 2 * in->texcoord_offset;
        if (in->tangent_offset != B200PT_NO_OFFSET) tangents = d->tangents + 3 * in->tangent_offset;
        if (in->bitangent_offset != B200PT_NO_OFFSET) bitangents = d->bitangents + 3 * in->bitangent_offset;
        indices = d->indices + 3 * in->index_offset;
    }
    /* scene.cpp:247-281: bake to_world */
    const Mat4 to_world = mat_load(in->to_world);
    Vec3 *pos = (Vec3 *)malloc(sizeof(Vec3) * nv), *nrm = NULL, *tan = NULL, *bit = NULL;
    for (uint64_t i = 0; i < nv; ++i) pos[i] = TransformPoint(&to_world, v3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]));
    if (normals) {
        const Mat4 tr = mat_transpose(&to_world), normal_to_world = mat_inverse(&tr);
        nrm = (Vec3 *)malloc(sizeof(Vec3) * nv);
        for (uint64_t i = 0; i < nv; ++i) nrm[i] = TransformVector(&normal_to_world, v3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
    }
    if (tangents) {
        tan = (Vec3 *)malloc(sizeof(Vec3) * nv);
        for (uint64_t i = 0; i < nv; ++i) tan[i] = TransformVector(&to_world, v3(tangents[3 * i], tangents[3 * i + 1], tangents[3 * i + 2]));
    }
    if (bitangents) {
        bit = (Vec3 *)malloc(sizeof(Vec3) * nv);
        for (uint64_t i = 0; i < nv; ++i) bit[i] = TransformVector(&to_world, v3(bitangents[3 * i], bitangents[3 * i + 1], bitangents[3 * i + 2]));
    }
    /* scene.cpp:15-111 SetupMeshes */
    g->num_prims = (uint32_t)nt;
    g->prims = (Primitive *)calloc(nt, sizeof(Primitive));
    float *areas = (float *)malloc(sizeof(float) * nt);
    AABB *aabbs = (AABB *)malloc(sizeof(AABB) * nt);
    for (uint32_t i = 0; i < nt; ++i) {
        Primitive *p = &g->prims[i];
        p->id = i, p->type = kPrimTriangle;
        const uint32_t idx[3] = {indices[3 * i], indices[3 * i + 1], indices[3 * i + 2]};
        if (!texcoords) {
            p->texcoords[0].u = 0, p->texcoords[0].v = 0, p->texcoords[1].u = 1, p->texcoords[1].v = 0, p->texcoords[2].u = 1, p->texcoords[2].v = 1;
        } else {
            for (int j = 0; j < 3; ++j) p->texcoords[j].u = texcoords[2 * idx[j]], p->texcoords[j].v = texcoords[2 * idx[j] + 1];
        }
        for (int j = 0; j < 3; ++j) p->positions[j] = pos[idx[j]];
        const Vec3 v0v1 = sub(p->positions[1], p->positions[0]), v0v2 = sub(p->positions[2], p->positions[0]);
        const Vec3 normal_geom = cross(v0v1, v0v2);
        areas[i] = length(normal_geom);
        if (!nrm) {
            const Vec3 normal = normalize(normal_geom);
            for (int j = 0; j < 3; ++j) p->normals[j] = normal;
        } else {
            for (int j = 0; j < 3; ++j) p->normals[j] = nrm[idx[j]];
        }
        if (!tan && !bit) {
            const Vec2 d01 = {p->texcoords[1].u - p->texcoords[0].u, p->texcoords[1].v - p->texcoords[0].v},
                       d02 = {p->texcoords[2].u - p->texcoords[0].u, p->texcoords[2].v - p->texcoords[0].v};
            const float r = 1.0f / (d01.v * d02.u - d01.u * d02.v);
            const Vec3 tangent = normalize(muls(sub(smul(d01.v, v0v2), smul(d02.v, v0v1)), r));
            for (int j = 0; j < 3; ++j) {
                p->bitangents[j] = normalize(cross(p->normals[j], tangent));
                p->tangents[j] = normalize(cross(p->bitangents[j], p->normals[j]));
            }
        } else if (!tan) {
            for (int j = 0; j < 3; ++j) {
                p->bitangents[j] = bit[idx[j]];
                p->tangents[j] = normalize(cross(p->bitangents[j], p->normals[j]));
                p->bitangents[j] = normalize(cross(p->normals[j], p->tangents[j]));
            }
        } else {
            for (int j = 0; j < 3; ++j) {
                p->tangents[j] = tan[idx[j]];
                p->bitangents[j] = normalize(cross(p->normals[j], p->tangents[j]));
                p->tangents[j] = normalize(cross(p->bitangents[j], p->normals[j]));
            }
        }
        aabbs[i] = aabb_empty(); /* triangle.cpp:9-15 */
        for (int j = 0; j < 3; ++j) aabb_add_point(&aabbs[i], p->positions[j]);
    }
    free(pos), free(nrm), free(tan), free(bit);
    *areas_out = areas, *aabbs_out = aabbs;
}

typedef struct {
    Scene scene;
    uint32_t num_bsdfs, num_textures;
} OracleScene;

static void FreeScene(OracleScene *os) {
    Scene *s = &os->scene;
    free(s->bsdfs), free(s->media), free(s->instances), free(s->list_pdf_area_instance), free(s->emitters);
    free(s->map_id_area_light_instance), free(s->map_id_instance_area_light), free(s->cdf_area_light), free(s->map_instance_bsdf);
    free(s->nodes), free(s->primitives), free(s->textures), free(s->data_env_map), free(s->brdf_avg), free(s->albedo_avg);
    free(os);
}

static OracleScene *CommitScene(const b200pt_scene_desc *d, int width, int height, int spp, int watertight) {
    OracleScene *os = (OracleScene *)calloc(1, sizeof(OracleScene));
    Scene *s = &os->scene;
    s->watertight = watertight;
    /* camera.cpp:26-37 */
    s->width = width > 0 ? width : d->camera.width;
    s->height = height > 0 ? height : d->camera.height;
    s->spp = spp > 0 ? (uint32_t)spp : d->camera.spp;
    s->spp_inv = 1.0f / s->spp;
    {
        const Vec3 eye = v3(d->camera.eye[0], d->camera.eye[1], d->camera.eye[2]), look_at = v3(d->camera.look_at[0], d->camera.look_at[1], d->camera.look_at[2]),
                   up_in = v3(d->camera.up[0], d->camera.up[1], d->camera.up[2]);
        const float fov_y = d->camera.fov_x * s->height / s->width;
        s->eye = eye;
        s->front = normalize(sub(look_at, eye));
        const Vec3 right = normalize(cross(s->front, up_in)), up = normalize(cross(right, s->front));
        const float to_rad = 0.01745329251994329576923690768489f;
        s->view_dx = muls(right, tanf((0.5f * d->camera.fov_x) * to_rad));
        s->view_dy = muls(up, tanf((0.5f * fov_y) * to_rad));
    }

    /* ---- Scene::CommitPrimitives (scene.cpp:153-198) ---- */
    const uint32_t ni = (uint32_t)d->num_instances;
    s->num_instances = ni;
    InstanceGeometry *geo = (InstanceGeometry *)calloc(ni ? ni : 1, sizeof(InstanceGeometry));
    uint64_t total_prims = 0, total_nodes = 0;
    for (uint32_t i = 0; i < ni; ++i) {
        const b200pt_instance *in = &d->instances[i];
        InstanceGeometry *g = &geo[i];
        float *areas = NULL;
        AABB *aabbs = NULL;
        const Mat4 to_world = mat_load(in->to_world);
        if (in->type == B200PT_INST_MESHES || in->type == B200PT_INST_RECTANGLE || in->type == B200PT_INST_CUBE) {
            SetupMeshTriangles(d, in, in->type == B200PT_INST_MESHES ? 0 : (int)in->type, g, &areas, &aabbs);
        } else {
            g->num_prims = 1;
            g->prims = (Primitive *)calloc(1, sizeof(Primitive));
            areas = (float *)malloc(sizeof(float));
            aabbs = (AABB *)malloc(sizeof(AABB));
            Primitive *p = g->prims;
            p->id = 0;
            aabbs[0] = aabb_empty();
            if (in->type == B200PT_INST_SPHERE) { /* scene.cpp:326-372, sphere.cpp:9-15 */
                p->type = kPrimSphere, p->radius = in->sphere_radius, p->center = v3(in->sphere_center[0], in->sphere_center[1], in->sphere_center[2]);
                p->to_world = to_world;
                aabb_add_point(&aabbs[0], TransformPoint(&to_world, adds(p->center, p->radius)));
                aabb_add_point(&aabbs[0], TransformPoint(&to_world, adds(p->center, -p->radius)));
                const Vec3 center_world = TransformPoint(&to_world, p->center), boundary_local = add(p->center, v3(p->radius, 0.0f, 0.0f)),
                           boundary_world = TransformPoint(&to_world, boundary_local);
                const float radius_world = length(sub(center_world, boundary_world));
                areas[0] = 4.0f * kPi * sqr(radius_world);
            } else if (in->type == B200PT_INST_DISK) { /* scene.cpp:374-416, disk.cpp:9-15 */
                p->type = kPrimDisk, p->to_world = to_world;
                aabb_add_point(&aabbs[0], TransformPoint(&to_world, v3(-0.5f, -0.5f, 0)));
                aabb_add_point(&aabbs[0], TransformPoint(&to_world, v3(0.5f, 0.5f, 0)));
                const Vec3 center_world = TransformPoint(&to_world, v3s(0)), boundary_world = TransformPoint(&to_world, v3(0.5f, 0, 0));
                const float radius_world = length(sub(center_world, boundary_world));
                areas[0] = kPi * sqr(radius_world);
            } else { /* cylinder: scene.cpp:418-472, cylinder.cpp:9-19 */
                p->type = kPrimCylinder;
                const Vec3 p0 = v3(in->cylinder_p0[0], in->cylinder_p0[1], in->cylinder_p0[2]), p1 = v3(in->cylinder_p1[0], in->cylinder_p1[1], in->cylinder_p1[2]);
                Mat4 m = LocalToWorldMat(normalize(sub(p1, p0)));
                const Mat4 tr = mat_translate(p0);
                m = mat_mul(&tr, &m);
                m = mat_mul(&to_world, &m);
                p->to_world = m;
                p->length = length(sub(TransformPoint(&m, v3(0, 0, length(sub(p1, p0)))), TransformPoint(&m, v3(0, 0, 0))));
                p->radius = length(sub(TransformPoint(&m, v3(in->cylinder_radius, 0, 0)), TransformPoint(&m, v3(0, 0, 0))));
                aabb_add_point(&aabbs[0], TransformPoint(&m, v3(p->radius, p->radius, 0)));
                aabb_add_point(&aabbs[0], TransformPoint(&m, v3(-p->radius, -p->radius, 0)));
                aabb_add_point(&aabbs[0], TransformPoint(&m, v3(p->radius, p->radius, p->length)));
                aabb_add_point(&aabbs[0], TransformPoint(&m, v3(-p->radius, -p->radius, p->length)));
                areas[0] = k2Pi * sqr(p->radius);
            }
        }
        g->nodes = (BvhNode *)malloc(sizeof(BvhNode) * (2 * (size_t)g->num_prims));
        g->num_nodes = BuildLinearBvh(g->num_prims, aabbs, areas, g->nodes);
        free(areas), free(aabbs);
        total_prims += g->num_prims, total_nodes += g->num_nodes;
    }

    /* ---- Scene::CommitInstances (scene.cpp:474-533) ---- */
    s->list_pdf_area_instance = (float *)malloc(sizeof(float) * (ni ? ni : 1));
    BvhNode *tlas = (BvhNode *)malloc(sizeof(BvhNode) * (2 * (size_t)(ni ? ni : 1)));
    uint32_t tlas_nodes = 0;
    if (ni > 0) {
        AABB *aabbs = (AABB *)malloc(sizeof(AABB) * ni);
        float *areas = (float *)malloc(sizeof(float) * ni);
        for (uint32_t i = 0; i < ni; ++i) {
            aabbs[i] = geo[i].nodes[0].aabb, areas[i] = geo[i].nodes[0].area;
            s->list_pdf_area_instance[i] = 1.0f / areas[i];
        }
        tlas_nodes = BuildLinearBvh(ni, aabbs, areas, tlas);
        free(aabbs), free(areas);
    }
    s->has_tlas = 1; /* the reference always allocates a TLAS (scene.cpp:527-528) */
    s->nodes = (BvhNode *)malloc(sizeof(BvhNode) * (tlas_nodes + total_nodes + 1));
    s->primitives = (Primitive *)malloc(sizeof(Primitive) * (total_prims + 1));
    memcpy(s->nodes, tlas, sizeof(BvhNode) * tlas_nodes);
    free(tlas);
    s->instances = (Instance *)calloc(ni ? ni : 1, sizeof(Instance));
    {
        uint64_t node_off = tlas_nodes, prim_off = 0;
        for (uint32_t i = 0; i < ni; ++i) {
            memcpy(s->nodes + node_off, geo[i].nodes, sizeof(BvhNode) * geo[i].num_nodes);
            memcpy(s->primitives + prim_off, geo[i].prims, sizeof(Primitive) * geo[i].num_prims);
            s->instances[i].id = i;
            s->instances[i].id_medium_int = d->instances[i].id_medium_int;
            s->instances[i].id_medium_ext = d->instances[i].id_medium_ext;
            s->instances[i].nodes = s->nodes + node_off;
            s->instances[i].primitives = s->primitives + prim_off;
            node_off += geo[i].num_nodes, prim_off += geo[i].num_prims;
            free(geo[i].nodes), free(geo[i].prims);
        }
    }
    free(geo);

    /* ---- renderer.cpp:271-304: maps and the un-normalised area-light CDF (Q4) ---- */
    s->map_instance_bsdf = (uint32_t *)malloc(sizeof(uint32_t) * (ni ? ni : 1));
    s->map_id_instance_area_light = (uint32_t *)malloc(sizeof(uint32_t) * (ni ? ni : 1));
    s->map_id_area_light_instance = (uint32_t *)malloc(sizeof(uint32_t) * (ni ? ni : 1));
    s->cdf_area_light = (float *)malloc(sizeof(float) * (ni + 1));
    s->cdf_area_light[0] = 0;
    uint32_t num_area_light = 0;
    for (uint32_t i = 0; i < ni; ++i) {
        s->map_instance_bsdf[i] = d->instances[i].id_bsdf;
        s->map_id_instance_area_light[i] = kInvalidId;
        if (d->instances[i].id_bsdf < d->num_bsdfs && d->bsdfs[d->instances[i].id_bsdf].type == B200PT_BSDF_AREA_LIGHT) {
            s->map_id_area_light_instance[num_area_light] = i;
            s->cdf_area_light[num_area_light + 1] = d->bsdfs[d->instances[i].id_bsdf].area_light_weight + s->cdf_area_light[num_area_light];
            s->map_id_instance_area_light[i] = num_area_light;
            ++num_area_light;
        }
    }
    s->num_area_light = num_area_light, s->size_cdf_area_light = num_area_light + 1;

    /* ---- textures (renderer.cpp:371-431) ---- */
    os->num_textures = (uint32_t)d->num_textures;
    s->textures = (Texture *)calloc(d->num_textures ? d->num_textures : 1, sizeof(Texture));
    for (uint64_t i = 0; i < d->num_textures; ++i) {
        const b200pt_texture *t = &d->textures[i];
        Texture *o = &s->textures[i];
        o->type = t->type;
        o->color0 = v3(t->color0[0], t->color0[1], t->color0[2]), o->color1 = v3(t->color1[0], t->color1[1], t->color1[2]);
        o->to_uv = mat_load(t->to_uv);
        o->width = t->width, o->height = t->height, o->channel = t->channels;
        o->data = d->pixels + t->pixel_offset;
    }
    const Texture *T = s->textures;
#define TEX(id) ((id) == kInvalidId ? NULL : T + (id))

    /* ---- Kulla-Conty LUT (renderer.cpp:311-314); only read by conductors/dielectrics ---- */
    s->brdf_avg = (float *)calloc(kLutResolution * kLutResolution, sizeof(float));
    s->albedo_avg = (float *)calloc(kLutResolution, sizeof(float));
    int need_lut = 0;
    for (uint64_t i = 0; i < d->num_bsdfs; ++i)
        if (d->bsdfs[i].type == B200PT_BSDF_CONDUCTOR || d->bsdfs[i].type == B200PT_BSDF_DIELECTRIC) need_lut = 1;
    if (need_lut) oracle_kulla_conty(s->brdf_avg, s->albedo_avg);

    /* ---- BSDFs (bsdf.cpp:112-186) ---- */
    os->num_bsdfs = (uint32_t)d->num_bsdfs;
    s->bsdfs = (Bsdf *)calloc(d->num_bsdfs ? d->num_bsdfs : 1, sizeof(Bsdf));
    for (uint64_t i = 0; i < d->num_bsdfs; ++i) {
        const b200pt_bsdf *b = &d->bsdfs[i];
        Bsdf *o = &s->bsdfs[i];
        o->type = b->type, o->twosided = b->twosided != 0;
        o->opacity = TEX(b->id_opacity), o->bump_map = TEX(b->id_bump_map);
        o->F_avg = 1.0f, o->F_avg_inv = 1.0f, o->reflectivity = 1.0f, o->eta = 1.0f, o->eta_inv = 1.0f;
        o->use_fast_approx = 0; /* never copied from the info struct by the reference (bsdf.cpp:136-141) */
        switch (b->type) {
        case B200PT_BSDF_AREA_LIGHT: o->radiance = TEX(b->id_radiance); break;
        case B200PT_BSDF_DIFFUSE: o->diffuse_reflectance = TEX(b->id_diffuse_reflectance); break;
        case B200PT_BSDF_ROUGH_DIFFUSE: o->diffuse_reflectance = TEX(b->id_diffuse_reflectance), o->roughness = TEX(b->id_roughness_u); break;
        case B200PT_BSDF_CONDUCTOR:
            o->roughness_u = TEX(b->id_roughness_u), o->roughness_v = TEX(b->id_roughness_v), o->specular_reflectance = TEX(b->id_specular_reflectance);
            o->reflectivity3 = v3(b->reflectivity[0], b->reflectivity[1], b->reflectivity[2]);
            o->edgetint = v3(b->edgetint[0], b->edgetint[1], b->edgetint[2]);
            o->F_avg3 = AverageFresnelConductor(o->reflectivity3, o->edgetint);
            break;
        case B200PT_BSDF_DIELECTRIC:
            o->F_avg = AverageFresnelDielectric(b->eta);
            o->F_avg_inv = AverageFresnelDielectric(1.0f / b->eta);
            /* fallthrough */
        case B200PT_BSDF_THIN_DIELECTRIC:
            o->twosided = 1;
            o->roughness_u = TEX(b->id_roughness_u), o->roughness_v = TEX(b->id_roughness_v);
            o->specular_reflectance = TEX(b->id_specular_reflectance), o->specular_transmittance = TEX(b->id_specular_transmittance);
            o->eta = b->eta, o->eta_inv = 1.0f / b->eta;
            o->reflectivity = (sqr(b->eta - 1.0f) / sqr(b->eta + 1.0f));
            break;
        case B200PT_BSDF_PLASTIC:
            o->roughness = TEX(b->id_roughness_u), o->diffuse_reflectance = TEX(b->id_diffuse_reflectance), o->specular_reflectance = TEX(b->id_specular_reflectance);
            o->reflectivity = (sqr(b->eta - 1.0f) / sqr(b->eta + 1.0f));
            o->F_avg = AverageFresnelDielectric(b->eta);
            break;
        }
    }

    /* ---- media (medium.cpp:6-39) ---- */
    s->media = (Medium *)calloc(d->num_media ? d->num_media : 1, sizeof(Medium));
    for (uint64_t i = 0; i < d->num_media; ++i) {
        const b200pt_medium *m = &d->media[i];
        Medium *o = &s->media[i];
        const Vec3 sigma_a = v3(m->sigma_a[0], m->sigma_a[1], m->sigma_a[2]), sigma_s = v3(m->sigma_s[0], m->sigma_s[1], m->sigma_s[2]);
        o->sigma_s = sigma_s, o->sigma_t = add(sigma_a, sigma_s);
        const Vec3 albedo = vdiv(sigma_s, add(sigma_a, sigma_s));
        o->sampling_weight = 0.0f;
        for (int dim = 0; dim < 3; ++dim)
            if (comp(albedo, dim) > o->sampling_weight && comp(o->sigma_t, dim) > 0) o->sampling_weight = comp(albedo, dim);
        if (o->sampling_weight > 0 && o->sampling_weight < 0.5f) o->sampling_weight = 0.5f;
        o->phase_type = m->phase_type, o->g = v3(m->g[0], m->g[1], m->g[2]);
    }

    /* ---- emitters (emitter.cpp:122-175, renderer.cpp:522-620, envmap.cpp:20-68) ---- */
    s->num_emitter = (uint32_t)d->num_emitters;
    s->id_sun = s->id_envmap = kInvalidId;
    s->emitters = (Emitter *)calloc(d->num_emitters ? d->num_emitters : 1, sizeof(Emitter));
    for (uint64_t i = 0; i < d->num_emitters; ++i) {
        const b200pt_emitter *e = &d->emitters[i];
        Emitter *o = &s->emitters[i];
        o->type = e->type;
        o->position = v3(e->position[0], e->position[1], e->position[2]);
        o->direction = v3(e->direction[0], e->direction[1], e->direction[2]);
        o->radiance = v3(e->radiance[0], e->radiance[1], e->radiance[2]);
        o->to_world = mat_load(e->to_world);
        o->to_local = mat_inverse(&o->to_world);
        o->texture = TEX(e->id_texture);
        switch (e->type) {
        case B200PT_EMIT_SPOT:
            o->cutoff_angle = e->cutoff_angle, o->cos_cutoff_angle = cosf(e->cutoff_angle), o->uv_factor = tanf(e->cutoff_angle);
            o->beam_width = e->beam_width, o->cos_beam_width = cosf(e->beam_width);
            o->transition_width_rcp = 1.0f / (e->cutoff_angle - e->beam_width);
            o->position = TransformPoint(&o->to_world, v3(0, 0, 0));
            break;
        case B200PT_EMIT_SUN:
            o->cos_cutoff_angle = e->cos_cutoff_angle;
            s->id_sun = (uint32_t)i;
            break;
        case B200PT_EMIT_ENVMAP: {
            const Texture *radiance = o->texture;
            const int w = radiance->width, h = radiance->height;
            const float width_inv = 1.0f / w, height_inv = 1.0f / h;
            float *cdf_rows = (float *)calloc(h + 1, sizeof(float)), *weight_rows = (float *)calloc(h, sizeof(float)),
                  *cdf_cols = (float *)calloc((size_t)(w + 1) * h, sizeof(float));
            float sum_row = 0.0f;
            cdf_rows[0] = 0;
            for (int y = 0; y < h; ++y) {
                float sum_col = 0.0f;
                cdf_cols[0] = 0;
                for (int x = 0; x < w; ++x) {
                    const Vec2 tc = {x * width_inv, y * height_inv};
                    const Vec3 rgb = GetColor(radiance, tc);
                    sum_col += 0.2126f * rgb.x + 0.7152f * rgb.y + 0.0722f * rgb.z;
                    cdf_cols[(size_t)y * (w + 1) + (x + 1)] = sum_col;
                }
                cdf_cols[(size_t)y * (w + 1) + w] = 1.0f;
                const float normalization_col = 1.0f / sum_col;
                for (int x = 1; x < w; ++x) cdf_cols[(size_t)y * (w + 1) + w - x] *= normalization_col;
                const float weight = sinf((y + 0.5f) * kPi / h);
                weight_rows[y] = weight;
                sum_row += sum_col * weight;
                cdf_rows[y + 1] = sum_row;
            }
            cdf_rows[h] = 1.0f;
            const float normalization_row = 1.0f / sum_row;
            for (int y = 1; y < h; ++y) cdf_rows[h - y] *= normalization_row;
            const float normalization = (float)(1.0 / (sum_row * (k2Pi * width_inv) * (kPi * height_inv)));
            /* packing renderer.cpp:584-605 vs wiring emitter.cpp:166-175 (Q9) */
            free(s->data_env_map);
            s->data_env_map = (float *)malloc(sizeof(float) * ((h + 1) + h + (size_t)(w + 1) * h));
            memcpy(s->data_env_map, cdf_rows, sizeof(float) * (h + 1));
            memcpy(s->data_env_map + (h + 1), weight_rows, sizeof(float) * h);
            memcpy(s->data_env_map + (h + 1) + h, cdf_cols, sizeof(float) * (size_t)(w + 1) * h);
            free(cdf_rows), free(weight_rows), free(cdf_cols);
            o->width = w, o->height = h, o->normalization = normalization;
            o->cdf_cols = s->data_env_map;
            o->cdf_rows = s->data_env_map + h + 1;
            o->weight_rows = s->data_env_map + (h + 1) + h;
            s->id_envmap = (uint32_t)i;
            break;
        }
        case B200PT_EMIT_CONSTANT:
            s->id_envmap = (uint32_t)i;
            break;
        }
    }

    /* ---- integrator (renderer.cpp:622-676) ---- */
    s->integrator_type = d->integrator.type;
    s->hide_emitters = d->integrator.hide_emitters != 0;
    s->pdf_rr = d->integrator.pdf_rr;
    s->pdf_rr_rcp = d->integrator.pdf_rr; /* Q1 */
    s->depth_rr = d->integrator.depth_rr, s->depth_max = d->integrator.depth_max;
    return os;
}

/* ------------------------------------------------------------------------------------------- */
/* Public entry points                                                                         */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    const Scene *scene;
    float *frame;
    volatile int *next_row;
    pthread_mutex_t *mutex;
} Worker;

static void *RenderRows(void *arg) {
    Worker *w = (Worker *)arg;
    for (;;) {
        pthread_mutex_lock(w->mutex);
        const int row = (*w->next_row)++;
        pthread_mutex_unlock(w->mutex);
        if (row >= w->scene->height) break;
        for (int i = 0; i < w->scene->width; ++i) DrawPixel(w->scene, (uint32_t)i, (uint32_t)row, w->frame);
    }
    return NULL;
}

/* Renders width*height*3 floats (0 = take the value from the scene's camera).  watertight = 1 selects Woop's
 * triangle test (-DWATERTIGHT_TRIANGLES), 0 Moeller-Trumbore.  Per-pixel seeds make the frame independent of
 * num_threads.  Returns 0 on success. */
int oracle_render(const b200pt_scene_desc *desc, int width, int height, int spp, int watertight, int num_threads, float *frame) {
    if (!desc || !frame) return -1;
    OracleScene *os = CommitScene(desc, width, height, spp, watertight);
    if (num_threads < 1) num_threads = 1;
    if (num_threads > 256) num_threads = 256;
    pthread_t threads[256];
    Worker workers[256];
    pthread_mutex_t mutex = PTHREAD_MUTEX_INITIALIZER;
    volatile int next_row = 0;
    for (int t = 0; t < num_threads; ++t) {
        workers[t].scene = &os->scene, workers[t].frame = frame, workers[t].next_row = &next_row, workers[t].mutex = &mutex;
        pthread_create(&threads[t], NULL, RenderRows, &workers[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    FreeScene(os);
    return 0;
}

/* Progressive preview frame: the body of DispathRaysCuda(camera, integrator, index_frame, frame, frame_srgb)
 * (renderer.cpp:97-138), which the reference only runs under CUDA + GLUT (ENABLE_VIEWER), restated for the CPU on top of the
 * pinned pieces (Tea<4>, VdC, ShadePath / ShadeVolPath).  The loop itself has no bit-exact pin: against the reference's CUDA
 * build of this call it agrees statistically (profiles/r01_progressive_pin.log), as the reference's own CPU and CUDA backends
 * do with each other; everything it calls is pinned bit for bit.  frame: running mean, updated in place; frame_srgb: sRGB
 * copy with row 0 at the bottom (may be NULL).  Single-threaded: meant for small frames. */
int oracle_render_progressive(const b200pt_scene_desc *desc, int width, int height, int watertight, uint32_t index_frame, float *frame,
                              float *frame_srgb) {
    if (!desc || !frame) return -1;
    OracleScene *os = CommitScene(desc, width, height, 1, watertight);
    const Scene *s = &os->scene;
    const float u = oracle_van_der_corput2(index_frame + 1), v = oracle_van_der_corput3(index_frame + 1);
    for (uint32_t j = 0; j < (uint32_t)s->height; ++j)
        for (uint32_t i = 0; i < (uint32_t)s->width; ++i) {
            const float x = 2.0f * (i + u) / s->width - 1.0f, y = 1.0f - 2.0f * (j + v) / s->height;
            const Vec3 look_dir = normalize(add(add(s->front, smul(x, s->view_dx)), smul(y, s->view_dy)));
            const uint32_t pixel_offset = (j * s->width + i) * 3, offset_dest = ((s->height - 1 - j) * s->width + i) * 3;
            uint32_t seed = oracle_tea4(pixel_offset, index_frame);
            const Vec3 L = s->integrator_type == B200PT_INTEGRATOR_VOLPATH ? ShadeVolPath(s, s->eye, look_dir, &seed) : ShadePath(s, s->eye, look_dir, &seed);
            const float color[3] = {fminf(L.x, 1.0f), fminf(L.y, 1.0f), fminf(L.z, 1.0f)};
            for (int c = 0; c < 3; ++c) {
                frame[pixel_offset + c] = (index_frame * frame[pixel_offset + c] + color[c]) / (index_frame + 1);
                if (frame_srgb != NULL)
                    frame_srgb[offset_dest + c] = frame[pixel_offset + c] <= 0.0031308f ? 12.92f * frame[pixel_offset + c]
                                                                                       : 1.055f * powf(frame[pixel_offset + c], 1.0f / 2.4f) - 0.055f;
            }
        }
    FreeScene(os);
    return 0;
}

/* One radiance sample for an arbitrary camera ray and LCG state (unit-test hook for ShadePath/ShadeVolPath). */
int oracle_shade(const b200pt_scene_desc *desc, int watertight, const float *eye, const float *dir, uint32_t *seed, float *rgb) {
    OracleScene *os = CommitScene(desc, 0, 0, 0, watertight);
    const Scene *s = &os->scene;
    const Vec3 L = s->integrator_type == B200PT_INTEGRATOR_VOLPATH ? ShadeVolPath(s, v3(eye[0], eye[1], eye[2]), v3(dir[0], dir[1], dir[2]), seed)
                                                                 : ShadePath(s, v3(eye[0], eye[1], eye[2]), v3(dir[0], dir[1], dir[2]), seed);
    rgb[0] = L.x, rgb[1] = L.y, rgb[2] = L.z;
    FreeScene(os);
    return 0;
}

/* ---- pointwise entries: the restated leaf functions at caller-supplied inputs, same record layout as b200pt_debug_eval
 * (include/b200pt.h) and ref_eval (oracle/ref_glue.cpp).  tests/test_oracle_pointwise.py pins them on the reference's own
 * functions without a GPU; tests/test_gpu_pointwise.py compares the device functions with the reference's. ---- */
void *oracle_scene_create(const b200pt_scene_desc *desc, int watertight) { return CommitScene(desc, 0, 0, 0, watertight); }
void oracle_scene_destroy(void *handle) { if (handle) FreeScene((OracleScene *)handle); }

int oracle_eval(void *handle, uint32_t what, uint32_t id, uint64_t n, const float *in_all, float *out_all) {
    const Scene *s = &((OracleScene *)handle)->scene;
    for (uint64_t i = 0; i < n; ++i) {
        const float *in = in_all + i * B200PT_EVAL_IN;
        float *out = out_all + i * B200PT_EVAL_OUT;
        memset(out, 0, sizeof(float) * B200PT_EVAL_OUT);
        uint32_t seed;
        memcpy(&seed, in + 18, 4);
#define IN3(p) v3((p)[0], (p)[1], (p)[2])
#define OUT3(p, v) ((p)[0] = (v).x, (p)[1] = (v).y, (p)[2] = (v).z)
        switch (what) {
        case B200PT_EVAL_BSDF_EVALUATE:
        case B200PT_EVAL_BSDF_SAMPLE: {
            BsdfSampleRec rec = RecInit();
            rec.wi = IN3(in), rec.wo = IN3(in + 3), rec.normal = IN3(in + 6), rec.tangent = IN3(in + 9), rec.bitangent = IN3(in + 12);
            rec.texcoord.u = in[15], rec.texcoord.v = in[16];
            rec.inside = in[17] != 0.0f;
            if (what == B200PT_EVAL_BSDF_EVALUATE) BsdfEvaluate(s, s->bsdfs + id, &rec);
            else BsdfSample(s, s->bsdfs + id, &seed, &rec);
            out[0] = (float)rec.valid, out[1] = rec.pdf;
            OUT3(out + 2, rec.attenuation), OUT3(out + 5, rec.wi);
            break;
        }
        case B200PT_EVAL_EMITTER_SAMPLE: {
            const Emitter *e = s->emitters + id;
            const EmitterSampleRec rec = EmitterSample(e, IN3(in), in[3], in[4]);
            out[0] = (float)rec.valid, out[1] = (float)rec.harsh, out[2] = rec.distance;
            OUT3(out + 3, rec.wi);
            if (rec.valid) {
                const Vec3 Le = EmitterEvaluateRec(e, &rec);
                OUT3(out + 6, Le);
                out[9] = EmitterPdf(e, neg(rec.wi));
            }
            break;
        }
        case B200PT_EVAL_EMITTER_DIR: {
            const Vec3 Le = EmitterEvaluateDir(s->emitters + id, IN3(in));
            OUT3(out, Le);
            out[3] = EmitterPdf(s->emitters + id, IN3(in));
            break;
        }
        case B200PT_EVAL_MEDIUM_SAMPLE:
        case B200PT_EVAL_MEDIUM_EVALUATE: {
            MediumSampleRec rec = MediumRecInit();
            if (what == B200PT_EVAL_MEDIUM_SAMPLE) {
                MediumSample(s->media + id, in[0], &seed, &rec);
            } else {
                rec.distance = in[0];
                MediumEvaluate(s->media + id, &rec);
            }
            out[0] = (float)rec.valid, out[1] = (float)rec.scattered, out[2] = rec.pdf, out[3] = rec.distance;
            OUT3(out + 4, rec.attenuation);
            break;
        }
        case B200PT_EVAL_PHASE_SAMPLE:
        case B200PT_EVAL_PHASE_EVALUATE: {
            PhaseSampleRec rec;
            memset(&rec, 0, sizeof(rec));
            rec.wi = IN3(in), rec.wo = IN3(in + 3);
            if (what == B200PT_EVAL_PHASE_SAMPLE) PhaseSample(s->media + id, &seed, &rec);
            else PhaseEvaluate(s->media + id, &rec);
            out[0] = (float)rec.valid, out[1] = rec.pdf;
            OUT3(out + 2, rec.attenuation), OUT3(out + 5, rec.wi);
            break;
        }
        case B200PT_EVAL_TEXTURE: {
            Vec2 uv = {in[0], in[1]};
            const Vec3 c = GetColor(s->textures + id, uv);
            OUT3(out, c);
            break;
        }
        default: return -1;
        }
#undef IN3
#undef OUT3
        memcpy(out + B200PT_EVAL_OUT - 1, &seed, 4);
    }
    return 0;
}

/* struct ref_hit of oracle/ref_glue.cpp */
typedef struct {
    float t;
    uint32_t valid, inside, id_instance, id_primitive;
    float position[3], normal[3], texcoord[2], tangent[3], bitangent[3];
} OracleHit;

/* TlasIntersect / TlasIntersectAny with the scene's BSDFs (bump maps, opacity masks drawing from `seed` = 0 per ray). */
int oracle_trace(void *handle, uint64_t n, const float *rays, int any_hit, OracleHit *out) {
    const Scene *s = &((OracleScene *)handle)->scene;
    for (uint64_t i = 0; i < n; ++i) {
        const float *q = rays + 8 * i;
        Ray ray = MakeRay(s, v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]));
        ray.t_min = q[6], ray.t_max = q[7];
        uint32_t seed = 0;
        OracleHit h;
        memset(&h, 0, sizeof(h));
        if (any_hit) {
            h.valid = (uint32_t)TlasIntersectAny(s, &seed, &ray);
            h.t = ray.t_max;
        } else {
            const Hit hit = TlasIntersect(s, &seed, &ray);
            h.t = ray.t_max;
            h.valid = (uint32_t)hit.valid, h.inside = (uint32_t)hit.inside, h.id_instance = hit.id_instance, h.id_primitive = hit.id_primitve;
            h.position[0] = hit.position.x, h.position[1] = hit.position.y, h.position[2] = hit.position.z;
            h.normal[0] = hit.normal.x, h.normal[1] = hit.normal.y, h.normal[2] = hit.normal.z;
            h.texcoord[0] = hit.texcoord.u, h.texcoord[1] = hit.texcoord.v;
            h.tangent[0] = hit.tangent.x, h.tangent[1] = hit.tangent.y, h.tangent[2] = hit.tangent.z;
            h.bitangent[0] = hit.bitangent.x, h.bitangent[1] = hit.bitangent.y, h.bitangent[2] = hit.bitangent.z;
        }
        out[i] = h;
    }
    return 0;
}
