// shading.cuh — textures, BSDFs, emitters, media and hit-attribute reconstruction (device).
//
// Restates, function by function, the closed-form evaluation/sampling routines the reference
// dispatches through Bsdf/Emitter/Medium/Texture (src/renderer/{bsdfs,emitters,medium,textures}),
// including the behaviours listed in SURVEY.md §8a-Q that change the image.  Conventions
// (path.cpp): `wi` = direction of light propagation arriving at the point, `wo` = direction
// from the point back towards the previous vertex; the next ray leaves along -wi.
#pragma once
#include "b200pt.h"
#include "textures.cuh"
#include "traverse.cuh"

namespace b200pt {

// ---------------------------------------------------------------------------------------------
// Surface point
// ---------------------------------------------------------------------------------------------
struct Surf {
    bool inside;
    uint32_t inst;
    V2 uv;
    V3 pos, n, t, b;
};

// bsdf.cpp:238-254
__device__ __forceinline__ V3 ApplyBump(const DeviceScene &s, const DBsdf &bsdf, V3 n, V3 t, V3 b, V2 uv) {
    if (bsdf.id_bump_map == kInvalid) return n;
    const V2 g = TexGradient(s, bsdf.id_bump_map, uv);
    return Normalize(-g.u * t - g.v * b + n);
}

__device__ __forceinline__ void FinishFrame(const DeviceScene &s, const DBsdf *bsdf, Surf *h) {
    if (bsdf != nullptr) { // triangle.cpp:127-133 and the identical blocks of sphere/disk/cylinder
        h->n = ApplyBump(s, *bsdf, h->n, h->t, h->b, h->uv);
        h->b = Normalize(Cross(h->n, h->t));
        h->t = Normalize(Cross(h->b, h->n));
    }
}

// triangle.cpp:115-146
__device__ __forceinline__ Surf SurfTriangle(const DeviceScene &s, uint32_t tri, float u, float v, bool inside) {
    const float w = 1.0f - u - v;
    const float4 *vp = reinterpret_cast<const float4 *>(s.tri_verts + tri);
    const float4 p0 = __ldg(vp), p1 = __ldg(vp + 1), p2 = __ldg(vp + 2);
    const float4 *sp = reinterpret_cast<const float4 *>(s.tri_shade + tri);
    float a[28];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float4 q = __ldg(sp + i);
        a[4 * i] = q.x, a[4 * i + 1] = q.y, a[4 * i + 2] = q.z, a[4 * i + 3] = q.w;
    }
    Surf h;
    h.inside = inside;
    h.inst = __float_as_uint(a[24]);
    h.uv = {u * a[18] + v * a[20] + w * a[22], u * a[19] + v * a[21] + w * a[23]};
    h.pos = {u * p0.x + v * p1.x + w * p2.x, u * p0.y + v * p1.y + w * p2.y, u * p0.z + v * p1.z + w * p2.z};
    const V3 n0 = mk3(a[0], a[1], a[2]), n1 = mk3(a[3], a[4], a[5]), n2 = mk3(a[6], a[7], a[8]);
    const V3 t0 = mk3(a[9], a[10], a[11]), t1 = mk3(a[12], a[13], a[14]), t2 = mk3(a[15], a[16], a[17]);
    h.n = Normalize(u * n0 + v * n1 + w * n2);
    h.t = Normalize(u * t0 + v * t1 + w * t2);
    const uint32_t id_bsdf = s.instances[h.inst].id_bsdf;
    const DBsdf *bsdf = id_bsdf != kInvalid ? s.bsdfs + id_bsdf : nullptr;
    if (bsdf != nullptr && bsdf->id_bump_map != kInvalid) {
        // per-vertex bitangents as SetupMeshes leaves them (scene.cpp:63-110)
        h.b = Normalize(u * Normalize(Cross(n0, t0)) + v * Normalize(Cross(n1, t1)) + w * Normalize(Cross(n2, t2)));
    } else {
        h.b = Normalize(Cross(h.n, h.t));
    }
    FinishFrame(s, bsdf, &h);
    if (inside) {
        h.n = -h.n;
        h.b = -h.b;
    }
    return h;
}

// sphere.cpp:17-88, disk.cpp:17-110, cylinder.cpp:21-90: attributes of the hit found by IntersectAnalytic.
__device__ __forceinline__ Surf SurfAnalytic(const DeviceScene &s, uint32_t index, const Ray &ray, float t_world) {
    const AnalyticPrim &p = s.analytic[index];
    Surf h;
    h.inst = p.inst;
    const uint32_t id_bsdf = s.instances[p.inst].id_bsdf;
    const DBsdf *bsdf = id_bsdf != kInvalid ? s.bsdfs + id_bsdf : nullptr;
    const V3 o_l = XformPoint(p.to_local, ray.o), d_l = XformVector(p.to_local, ray.d);
    if (p.type == kSphere) {
        const V3 ro = o_l - mk3(p.center);
        const float a = Dot(d_l, d_l), b = 2.0f * Dot(d_l, ro), c = Dot(ro, ro) - Sqr(p.radius);
        float t_near = 0.0f, t_far = 0.0f;
        SolveQuadratic(a, b, c, &t_near, &t_far);
        const float t = t_near < kEpsilonDistance ? t_far : t_near;
        const V3 pos_l = ro + t * d_l;
        h.pos = XformPoint(p.to_world, pos_l + mk3(p.center));
        float theta, phi;
        CartesianToSpherical(pos_l, &theta, &phi, nullptr);
        h.uv = {phi * k1Div2Pi, theta * k1DivPi};
        h.inside = c < 0.0f;
        const V3 n_l = Normalize(pos_l);
        h.n = XformVector(p.normal_to_world, n_l);
        constexpr float eps = 0.01f * kPi;
        float theta_p = theta + eps;
        const bool flip = theta_p > kPi;
        if (flip) theta_p = theta - eps;
        const V3 pos_p = XformPoint(p.to_world, SphericalToCartesian(theta_p, phi, 1));
        h.b = Normalize(pos_p - h.pos);
        if (flip) h.b = -h.b;
        h.t = Normalize(Cross(h.b, h.n));
        h.b = Normalize(Cross(h.n, h.t));
        FinishFrame(s, bsdf, &h);
    } else if (p.type == kDisk) {
        const float t_z = -o_l.z / d_l.z;
        const V3 pos_l = o_l + t_z * d_l;
        h.pos = XformPoint(p.to_world, pos_l);
        float theta, phi, r;
        CartesianToSpherical(pos_l, &theta, &phi, &r);
        h.uv = {r, phi * k1Div2Pi};
        h.inside = d_l.z > 0;
        constexpr float eps = 0.01f * kPi;
        float r_p = r + eps;
        const bool flip_b = r_p > r;
        if (flip_b) r_p = r - eps;
        float phi_p = phi + eps;
        const bool flip_t = phi_p > kPi;
        if (flip_t) phi_p = phi - eps;
        const V3 e1 = SphericalToCartesian(theta, phi, r_p) - pos_l, e2 = SphericalToCartesian(theta, phi_p, r) - pos_l;
        const V2 d1 = {r_p - h.uv.u, h.uv.v - h.uv.v}, d2 = {h.uv.u - h.uv.u, phi_p * k1Div2Pi - h.uv.v};
        const float norm = 1.0f / (d2.u * d1.v - d1.u * d2.v);
        V3 tangent = Normalize((d1.v * e2 - d2.v * e1) * norm), bitangent = Normalize((d2.u * e1 - d1.u * e2) * norm);
        V3 normal = mk3(0, 0, 1);
        if (flip_b) bitangent = -bitangent;
        if (flip_t) tangent = -tangent;
        bitangent = Normalize(Cross(normal, tangent));
        tangent = Normalize(Cross(bitangent, normal));
        h.n = normal, h.t = tangent, h.b = bitangent;
        FinishFrame(s, bsdf, &h);
        h.n = XformVector(p.normal_to_world, h.n);
        h.t = XformVector(p.to_world, h.t);
        h.b = XformVector(p.to_world, h.b);
    } else {
        const float a = Sqr(d_l.x) + Sqr(d_l.y), b = 2.0f * (d_l.x * o_l.x + d_l.y * o_l.y),
                    c = Sqr(o_l.x) + Sqr(o_l.y) - Sqr(p.radius);
        float t_near = 0.0f, t_far = 0.0f;
        SolveQuadratic(a, b, c, &t_near, &t_far);
        const float z_near = o_l.z + d_l.z * t_near;
        const float t = (kEpsilonDistance < t_near && 0.0f <= z_near && z_near <= p.length) ? t_near : t_far;
        const V3 pos_l = o_l + t * d_l;
        h.uv = {atan2f(pos_l.y, pos_l.x) * k1Div2Pi, pos_l.z / p.length};
        h.pos = XformPoint(p.to_world, pos_l);
        h.inside = c < 0.0f;
        const V3 n_l = Normalize(mk3(pos_l.x, pos_l.y, 0.0f));
        h.n = XformVector(p.normal_to_world, n_l);
        h.t = XformVector(p.normal_to_world, mk3(0, 0, 1));
        h.b = Normalize(Cross(h.n, h.t));
        FinishFrame(s, bsdf, &h);
    }
    (void)t_world;
    if (h.inside) {
        h.n = -h.n;
        h.b = -h.b;
    }
    return h;
}

// ---------------------------------------------------------------------------------------------
// Area-light point sampling: Instance::Sample (instance.cpp:56-60) -> BLAS::Sample (blas.cpp:79-98)
// -> Sample{Triangle,Sphere,Disk,Cylinder}.  The reference descends its LBVH by sub-tree area; a CDF
// over the same triangle areas selects triangles with the same probabilities.
// ---------------------------------------------------------------------------------------------
struct LightPoint {
    V3 pos, n;
    V2 uv;
};

__device__ __forceinline__ LightPoint SampleInstance(const DeviceScene &s, uint32_t inst_id, float xi_0, float xi_1,
                                                     float xi_2) {
    const DInstance &inst = s.instances[inst_id];
    LightPoint lp;
    if (inst.analytic == kInvalid) {
        const float *cdf = s.light_tri_cdf + inst.light_tri_begin;
        uint32_t lo = 0, hi = inst.light_tri_count - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(cdf + mid) <= xi_0)
                lo = mid + 1;
            else
                hi = mid;
        }
        const uint32_t tri = __ldg(s.light_tri_ids + inst.light_tri_begin + lo);
        // triangle.cpp:150-160
        const float temp = sqrtf(1.0f - xi_1);
        const float u = 1.0f - temp, v = temp * xi_2, w = 1.0f - u - v;
        const float4 *vp = reinterpret_cast<const float4 *>(s.tri_verts + tri);
        const float4 p0 = __ldg(vp), p1 = __ldg(vp + 1), p2 = __ldg(vp + 2);
        const TriShade &sh = s.tri_shade[tri];
        lp.uv = {w * sh.uv[0][0] + u * sh.uv[1][0] + v * sh.uv[2][0], w * sh.uv[0][1] + u * sh.uv[1][1] + v * sh.uv[2][1]};
        lp.pos = {w * p0.x + u * p1.x + v * p2.x, w * p0.y + u * p1.y + v * p2.y, w * p0.z + u * p1.z + v * p2.z};
        lp.n = Normalize(mk3(w * sh.n[0][0] + u * sh.n[1][0] + v * sh.n[2][0], w * sh.n[0][1] + u * sh.n[1][1] + v * sh.n[2][1],
                             w * sh.n[0][2] + u * sh.n[1][2] + v * sh.n[2][2]));
        return lp;
    }
    const AnalyticPrim &p = s.analytic[inst.analytic];
    if (p.type == kSphere) { // sphere.cpp:90-105
        const float cos_theta = 1.0f - 2.0f * xi_1;
        lp.uv = {xi_2, acosf(cos_theta) * k1DivPi};
        const float sin_theta = sqrtf(1.0f - Sqr(cos_theta)), phi = k2Pi * xi_2;
        const V3 n_l = mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta);
        lp.pos = XformPoint(p.to_world, mk3(p.center) + p.radius * n_l);
        lp.n = XformVector(p.normal_to_world, n_l);
    } else if (p.type == kDisk) { // disk.cpp:112-141
        const float r1 = 2.0f * xi_1 - 1.0f, r2 = 2.0f * xi_2 - 1.0f;
        float phi, r;
        if (r1 == 0.0f && r2 == 0.0f) {
            r = phi = 0;
        } else if (Sqr(r1) > Sqr(r2)) {
            r = r1;
            phi = kPiDiv4 * (r2 / r1);
        } else {
            r = r2;
            phi = kPiDiv2 - (r1 / r2) * kPiDiv4;
        }
        lp.uv = {r, phi * k1Div2Pi};
        lp.pos = XformPoint(p.to_world, mk3(r * cosf(phi) * 0.5f, r * sinf(phi) * 0.5f, 0.0f));
        lp.n = XformVector(p.normal_to_world, mk3(0, 0, 1));
    } else { // cylinder.cpp:92-105
        const float phi = k2Pi * xi_1, z = xi_2 * p.length;
        lp.uv = {xi_1, xi_2};
        lp.pos = XformPoint(p.to_world, mk3(cosf(phi) * p.radius, sinf(phi) * p.radius, z));
        lp.n = XformVector(p.normal_to_world, mk3(cosf(phi), sinf(phi), 0.0f));
    }
    return lp;
}

// ---------------------------------------------------------------------------------------------
// BSDFs
// ---------------------------------------------------------------------------------------------
struct BsdfRec { // bsdf.hpp:81-97
    bool valid = false;
    bool inside = false;
    float pdf = 0;
    V2 uv = {0, 0};
    V3 wi = {0, 0, 0}, wo = {0, 0, 0}, pos = {0, 0, 0}, n = {0, 0, 0}, t = {0, 0, 0}, b = {0, 0, 0}, att = {0, 0, 0};
    __device__ __forceinline__ V3 ToLocal(V3 v) const { return Normalize(mk3(Dot(v, t), Dot(v, b), Dot(v, n))); }
    __device__ __forceinline__ V3 ToWorld(V3 v) const { return Normalize(v.x * t + v.y * b + v.z * n); }
};

// microfacet.cpp:21-37 (pdf: pow(cos,3) evaluated as a product)
__device__ __forceinline__ void SampleGgx(float xi_0, float xi_1, float ru, float rv, V3 *vec, float *pdf) {
    const float phi = atanf(rv / ru * tanf(kPi + k2Pi * xi_1)) + kPi * floorf(2.0f * xi_1 + 0.5f);
    float sin_phi, cos_phi;
    sincosf(phi, &sin_phi, &cos_phi);
    const float alpha_2 = 1.0f / (Sqr(cos_phi / ru) + Sqr(sin_phi / rv));
    const float tan_theta_2 = alpha_2 * xi_0 / (1.0f - xi_0);
    const float cos_theta = 1.0f / sqrtf(1.0f + tan_theta_2), sin_theta = sqrtf(1.0f - Sqr(cos_theta));
    *vec = {sin_theta * cos_phi, sin_theta * sin_phi, cos_theta};
    *pdf = 1.0f / (kPi * ru * rv * (cos_theta * cos_theta * cos_theta) * Sqr(1.0f + tan_theta_2 / alpha_2));
}
// microfacet.cpp:8-19
__device__ __forceinline__ void SampleGgx(float xi_0, float xi_1, float r, V3 *vec, float *pdf) {
    const float alpha_2 = Sqr(r);
    const float tan_theta_2 = alpha_2 * xi_0 / (1.0f - xi_0), phi = k2Pi * xi_1;
    const float cos_theta = 1.0f / sqrtf(1.0f + tan_theta_2), sin_theta = sqrtf(1.0f - Sqr(cos_theta));
    float s, c;
    sincosf(phi, &s, &c);
    *vec = {sin_theta * c, sin_theta * s, cos_theta};
    *pdf = 1.0f / (kPi * alpha_2 * (cos_theta * cos_theta * cos_theta) * Sqr(1.0f + tan_theta_2 / alpha_2));
}
// microfacet.cpp:39-49
__device__ __forceinline__ float PdfGgx(float r, V3 vec) {
    const float cos_theta = vec.z;
    if (cos_theta <= 0.0f) return 0.0f;
    const float cos_theta_2 = Sqr(cos_theta), tan_theta_2 = (1.0f - cos_theta_2) / cos_theta_2,
                cos_theta_3 = cos_theta * cos_theta * cos_theta, alpha_2 = Sqr(r);
    return alpha_2 / (kPi * cos_theta_3 * Sqr(alpha_2 + tan_theta_2));
}
// microfacet.cpp:51-61
__device__ __forceinline__ float PdfGgx(float ru, float rv, V3 vec) {
    const float cos_theta = vec.z;
    if (cos_theta <= 0.0f) return 0.0f;
    return cos_theta / (kPi * ru * rv * Sqr(Sqr(vec.x / ru) + Sqr(vec.y / rv) + Sqr(cos_theta)));
}
// microfacet.cpp:63-75
__device__ __forceinline__ float SmithG1Ggx(float r, V3 v, V3 h) {
    const float n_dot_v = v.z;
    if (n_dot_v * h.z <= 0) return 0;
    const float cos_theta_2 = Sqr(n_dot_v), tan_theta_2 = (1.0f - cos_theta_2) / cos_theta_2, alpha_2 = Sqr(r);
    return 2.0f / (1.0f + sqrtf(1.0f + alpha_2 * tan_theta_2));
}
// microfacet.cpp:77-85
__device__ __forceinline__ float SmithG1Ggx(float ru, float rv, V3 v, V3 h) {
    const float n_dot_v = v.z;
    if (n_dot_v * h.z <= 0) return 0;
    const float xy_alpha_2 = Sqr(ru * v.x) + Sqr(rv * v.y), tan_theta_2 = xy_alpha_2 / Sqr(n_dot_v);
    return 2.0f / (1.0f + sqrtf(1.0f + tan_theta_2));
}
// microfacet.hpp:24-29
__device__ __forceinline__ float Pow5(float x) {
    const float x2 = x * x;
    return x2 * x2 * x;
}
__device__ __forceinline__ float FresnelSchlick(float cos_theta, float r) { return (1.0f - r) * Pow5(1.0f - cos_theta) + r; }
__device__ __forceinline__ V3 FresnelSchlick(float cos_theta, V3 r) { return (1.0f - r) * Pow5(1.0f - cos_theta) + r; }

// kulla_conty.cpp:82-131
__device__ __forceinline__ float GetBrdfAvg(const float *buf, float cos_theta, float roughness) {
    constexpr int R = kLutResolution;
    // Quirk found by the pointwise tests: kLutResolution is a uint32_t in the reference (kulla_conty.hpp:9), so its
    // `offset_int2 >= kLutResolution - 1` compares UNSIGNED.  EvaluateDielectric passes negative cosines (N_dot_O < 0 in the
    // transmission branch, dielectric.cpp:207-212; N_dot_I < 0 for light arriving from below): a column <= -1 wraps to a
    // huge value and takes the "last column" branch — a cosine below -1/128 reads the table at cosine 1.  Cosines in
    // (-1/128, 0) truncate to column 0 and extrapolate with a negative weight.
    const float offset1 = roughness * R, offset2 = cos_theta * R;
    const int i1 = static_cast<int>(offset1), i2 = static_cast<int>(offset2);
    const bool last_column = static_cast<uint32_t>(i2) >= static_cast<uint32_t>(R - 1);
    if (i1 >= R - 1) {
        if (last_column) return __ldg(buf + (R - 1) * R + R - 1);
        return Lerp(__ldg(buf + (R - 1) * R + i2), __ldg(buf + (R - 1) * R + i2 + 1), offset2 - i2);
    }
    if (last_column) return Lerp(__ldg(buf + i1 * R + R - 1), __ldg(buf + (i1 + 1) * R + R - 1), offset1 - i1);
    return Lerp(Lerp(__ldg(buf + i1 * R + i2), __ldg(buf + (i1 + 1) * R + i2), offset1 - i1),
                Lerp(__ldg(buf + i1 * R + i2 + 1), __ldg(buf + (i1 + 1) * R + i2 + 1), offset1 - i1), offset2 - i2);
}
// kulla_conty.cpp:133-143
__device__ __forceinline__ float GetAlbedoAvg(const float *buf, float roughness) {
    constexpr int R = kLutResolution;
    const float offset = roughness * R;
    const int i = static_cast<int>(offset);
    if (i >= R - 1) return __ldg(buf + R - 1);
    return Lerp(__ldg(buf + i), __ldg(buf + i + 1), offset - i);
}

// diffuse.cpp:9-34
__device__ __forceinline__ void EvaluateDiffuse(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    rec->pdf = Dot(rec->wo, rec->n); // Q5: not cos/pi, and of wo
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    const V3 albedo = TexColor(s, d.id_diffuse_reflectance, rec->uv);
    rec->att = albedo * k1DivPi * Dot(-rec->wi, rec->n);
}
__device__ __forceinline__ void SampleDiffuse(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    V3 wi_local;
    const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
    SampleHemisCos(xi_0, xi_1, &wi_local, &rec->pdf);
    if (rec->pdf < kEpsilon) return; // Q6
    rec->wi = -rec->ToWorld(wi_local);
    rec->valid = true;
    const V3 albedo = TexColor(s, d.id_diffuse_reflectance, rec->uv);
    rec->att = albedo * k1DivPi * wi_local.z;
}

// rough_diffuse.cpp:10-97 (Oren-Nayar; use_fast_approx is always false in the reference, see scene_build.cpp)
__device__ __forceinline__ void OrenNayar(float roughness, V3 albedo, bool fast, BsdfRec *rec) {
    constexpr float conversion_factor = 0.70710678118f;
    const float sigma_2 = Sqr(roughness * conversion_factor);
    const V3 wi_l = rec->ToLocal(-rec->wi), wo_l = rec->ToLocal(rec->wo);
    const float n_dot_i = wi_l.z, n_dot_o = wo_l.z, sin_theta_i = sqrtf(1.0f - n_dot_i * n_dot_i),
                sin_theta_o = sqrtf(1.0f - n_dot_o * n_dot_o);
    float phi_i, theta_i, phi_o, theta_o;
    CartesianToSpherical(wi_l, &theta_i, &phi_i, nullptr);
    CartesianToSpherical(wo_l, &theta_o, &phi_o, nullptr);
    const float cos_phi_diff = cosf(phi_i) * cosf(phi_o) + sinf(phi_i) * sinf(phi_o);
    if (fast) {
        const float A = 1.0f - 0.5f * sigma_2 / (sigma_2 + 0.33f), B = 0.45f * sigma_2 / (sigma_2 + 0.09f);
        float sin_alpha, tan_beta;
        if (n_dot_i > n_dot_o) {
            sin_alpha = sin_theta_o;
            tan_beta = sin_theta_i / n_dot_i;
        } else {
            sin_alpha = sin_theta_i;
            tan_beta = sin_theta_o / n_dot_o;
        }
        rec->att = albedo * k1DivPi * n_dot_i * (A + B * fmaxf(cos_phi_diff, 0.0f) * sin_alpha * tan_beta);
    } else {
        const float alpha = fmaxf(theta_i, theta_o), beta = fminf(theta_i, theta_o);
        float sin_alpha, sin_beta, tan_beta;
        if (n_dot_i > n_dot_o) {
            sin_alpha = sin_theta_o;
            sin_beta = sin_theta_i;
            tan_beta = sin_theta_i / n_dot_i;
        } else {
            sin_alpha = sin_theta_i;
            sin_beta = sin_theta_o;
            tan_beta = sin_theta_o / n_dot_o;
        }
        const float tmp = sigma_2 / (sigma_2 + 0.09f), tmp2 = 4.0f * k1DivPi * k1DivPi * alpha * beta,
                    tmp3 = 2.0f * beta * k1DivPi;
        const float C1 = 1.0f - 0.5f * sigma_2 / (sigma_2 + 0.33f);
        float C2 = 0.45f * tmp;
        const float C3 = 0.125f * tmp * tmp2 * tmp2, C4 = 0.17f * sigma_2 / (sigma_2 + 0.13f);
        if (cos_phi_diff > 0)
            C2 *= sin_alpha;
        else
            C2 *= sin_alpha - tmp3 * tmp3 * tmp3;
        const float tan_half = (sin_alpha + sin_beta) /
                               (sqrtf(fmaxf(0.0f, 1.0f - Sqr(sin_alpha))) + sqrtf(fmaxf(0.0f, 1.0f - Sqr(sin_beta))));
        const V3 sngl = albedo * (C1 + cos_phi_diff * C2 * tan_beta + (1.0f - fabsf(cos_phi_diff)) * C3 * tan_half),
                 dbl = Sqr(albedo) * (C4 * (1.0f - cos_phi_diff * Sqr(tmp3)));
        rec->att = (sngl + dbl) * k1DivPi * n_dot_i;
    }
}
// rough_diffuse.cpp:99-129
__device__ __forceinline__ void SampleRoughDiffuse(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    V3 wi;
    const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
    SampleHemisCos(xi_0, xi_1, &wi, &rec->pdf);
    if (rec->pdf < kEpsilon) return;
    rec->wi = -Normalize(wi.x * rec->t + wi.y * rec->b + wi.z * rec->n);
    rec->valid = true;
    OrenNayar(TexColor(s, d.id_roughness_u, rec->uv).x, TexColor(s, d.id_diffuse_reflectance, rec->uv), d.use_fast_approx != 0, rec);
}
__device__ __forceinline__ void EvaluateRoughDiffuse(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    rec->pdf = Dot(rec->wo, rec->n);
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    OrenNayar(TexColor(s, d.id_roughness_u, rec->uv).x, TexColor(s, d.id_diffuse_reflectance, rec->uv), d.use_fast_approx != 0, rec);
}

// conductor.cpp:14-28
__device__ __forceinline__ V3 ConductorMultiScatter(const DeviceScene &s, const DBsdf &d, float n_dot_i, float n_dot_o, float roughness) {
    const float brdf_i = GetBrdfAvg(s.kc_brdf_avg, n_dot_i, roughness), brdf_o = GetBrdfAvg(s.kc_brdf_avg, n_dot_o, roughness),
                albedo_avg = GetAlbedoAvg(s.kc_albedo_avg, roughness),
                f_ms = (1.0f - brdf_i) * (1.0f - brdf_o) / (kPi * (1.0f - albedo_avg));
    const V3 F_avg = mk3(d.F_avg);
    const V3 f_add = Sqr(F_avg) * albedo_avg / (1.0f - F_avg * (1.0f - albedo_avg));
    return f_ms * f_add * n_dot_i;
}
// conductor.cpp:34-77
__device__ __forceinline__ void SampleConductor(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    V3 h_local = mk3(0.0f);
    float D = 0;
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x;
    const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
    SampleGgx(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const V3 h_world = rec->ToWorld(h_local);
    const float h_dot_o = Dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * h_dot_o);
    if (rec->pdf < kEpsilon) return;
    rec->wi = -Reflect(-rec->wo, h_world);
    const float n_dot_i = Dot(-rec->wi, rec->n);
    if (n_dot_i < kEpsilonFloat) return;
    rec->valid = true;
    const V3 wi_local = rec->ToLocal(-rec->wi), wo_local = rec->ToLocal(rec->wo);
    const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local),
                h_dot_i = Dot(-rec->wi, h_world), n_dot_o = wo_local.z;
    const V3 F = FresnelSchlick(h_dot_i, mk3(d.reflectivity));
    rec->att = (F * D * G) / (4.0f * n_dot_o);
    if (alpha_u == alpha_v) rec->att += ConductorMultiScatter(s, d, n_dot_i, n_dot_o, alpha_u);
    rec->att *= TexColor(s, d.id_specular_reflectance, rec->uv);
}
// conductor.cpp:79-119
__device__ __forceinline__ void EvaluateConductor(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    const float n_dot_o = Dot(rec->wo, rec->n);
    if (n_dot_o < kEpsilonFloat) return;
    const V3 h_world = Normalize(-rec->wi + rec->wo), h_local = rec->ToLocal(h_world);
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x,
                D = PdfGgx(alpha_u, alpha_v, h_local), h_dot_o = Dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * h_dot_o);
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    const V3 wi_local = rec->ToLocal(-rec->wi), wo_local = rec->ToLocal(rec->wo);
    const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local),
                h_dot_i = Dot(-rec->wi, h_world);
    const V3 F = FresnelSchlick(h_dot_i, mk3(d.reflectivity));
    rec->att = (F * D * G) / (4.0f * n_dot_o);
    if (alpha_u == alpha_v) rec->att += ConductorMultiScatter(s, d, Dot(-rec->wi, rec->n), n_dot_o, alpha_u);
    rec->att *= TexColor(s, d.id_specular_reflectance, rec->uv);
}

// dielectric.cpp:14-38
__device__ __forceinline__ float DielectricMultiScatter(const DeviceScene &s, const DBsdf &d, float n_dot_i, float n_dot_o,
                                                        float roughness, bool inside, bool reflect) {
    const float brdf_i = GetBrdfAvg(s.kc_brdf_avg, n_dot_i, roughness), brdf_o = GetBrdfAvg(s.kc_brdf_avg, n_dot_o, roughness),
                albedo_avg = GetAlbedoAvg(s.kc_albedo_avg, roughness),
                f_ms = (1.0f - brdf_i) * (1.0f - brdf_o) / (kPi * (1.0f - albedo_avg));
    const float F_avg = inside ? d.F_avg_inv_s : d.F_avg_s, eta = inside ? d.eta_inv : d.eta;
    const float f_add = (F_avg * F_avg) * albedo_avg / (1.0f - F_avg * (1.0f - albedo_avg)),
                ratio_trans = ((1.0f - d.F_avg_s) * (1.0f - d.F_avg_inv_s) * (eta * eta) /
                               ((1.0f - d.F_avg_s) + (1.0f - d.F_avg_inv_s) * (eta * eta)));
    const float ret = f_ms * f_add * n_dot_i;
    return reflect ? (1.0f - ratio_trans) * ret : ratio_trans * ret;
}
// dielectric.cpp:44-145 — abs() has float semantics (Q13)
__device__ __forceinline__ void SampleDielectric(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    const float scale = 1.2f - 0.2f * sqrtf(fabsf(Dot(-rec->wo, rec->n)));
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x * scale, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x * scale;
    V3 h_local = mk3(0.0f);
    float D = 0;
    const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
    SampleGgx(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const V3 h_world = rec->ToWorld(h_local);
    float h_dot_o = Dot(rec->wo, h_world);
    if (h_dot_o < kEpsilonFloat) return;
    float eta = d.eta, eta_inv = d.eta_inv;
    if (!rec->inside) {
        const float temp = eta_inv;
        eta_inv = eta;
        eta = temp;
    }
    V3 wt = mk3(0.0f);
    const bool full_reflect = !Refract(-rec->wo, h_world, eta, &wt);
    float F = FresnelSchlick(h_dot_o, d.reflectivity_s);
    const V3 wo_local = rec->ToLocal(rec->wo);
    if (full_reflect || rng.Next() < F) {
        rec->wi = -Reflect(-rec->wo, h_world);
        const float n_dot_i = Dot(-rec->wi, rec->n);
        if (n_dot_i < kEpsilonFloat) return;
        rec->pdf = F * D / (4.0f * h_dot_o);
        if (rec->pdf < kEpsilon) return;
        const V3 wi_local = rec->ToLocal(-rec->wi);
        const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local),
                    n_dot_o = wo_local.z;
        float att = (F * D * G) / (4.0f * n_dot_o);
        if (alpha_u == alpha_v) att += DielectricMultiScatter(s, d, n_dot_i, n_dot_o, alpha_u, rec->inside, true);
        rec->att = att * TexColor(s, d.id_specular_reflectance, rec->uv);
    } else {
        rec->wi = -wt;
        V3 wi_local = rec->ToLocal(-rec->wi);
        wi_local.z = -wi_local.z;
        const float n_dot_i = wi_local.z;
        if (n_dot_i < kEpsilonFloat) return;
        const float h_dot_i = -Dot(wt, h_world);
        if (h_dot_i < kEpsilonFloat) return;
        h_dot_o = -h_dot_o;
        F = FresnelSchlick(h_dot_i, d.reflectivity_s);
        rec->pdf = ((1.0f - F) * D) * fabsf(h_dot_o / Sqr(eta_inv * h_dot_i + h_dot_o));
        if (rec->pdf < kEpsilon) return;
        const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local),
                    n_dot_o = wo_local.z;
        float att = ((fabsf(h_dot_i) * fabsf(h_dot_o)) * ((1.0f - F) * G * D)) / fabsf(n_dot_o * Sqr(eta_inv * h_dot_i + h_dot_o));
        if (alpha_u == alpha_v) att += DielectricMultiScatter(s, d, n_dot_i, n_dot_o, alpha_u, !rec->inside, false);
        att *= Sqr(eta);
        rec->att = att * TexColor(s, d.id_specular_transmittance, rec->uv);
    }
    rec->valid = true;
}
// dielectric.cpp:147-224
__device__ __forceinline__ void EvaluateDielectric(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    float eta = d.eta, eta_inv = d.eta_inv;
    if (rec->inside) {
        const float temp = eta_inv;
        eta_inv = eta;
        eta = temp;
    }
    const float n_dot_o = Dot(rec->wo, rec->n);
    const bool reflect = n_dot_o > 0.0f;
    const V3 h_world = reflect ? Normalize(-rec->wi + rec->wo) : -Normalize(eta_inv * (-rec->wi) + rec->wo),
             h_local = rec->ToLocal(h_world);
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x,
                D = PdfGgx(alpha_u, alpha_v, h_local), h_dot_i = Dot(-rec->wi, h_world), h_dot_o = Dot(rec->wo, h_world),
                F = FresnelSchlick(h_dot_i, d.reflectivity_s);
    rec->pdf = reflect ? (F * D) / (4.0f * h_dot_o) : (((1.0f - F) * D) * fabsf(h_dot_o / Sqr(eta_inv * h_dot_i + h_dot_o)));
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    const V3 wi_local = rec->ToLocal(-rec->wi);
    if (reflect) {
        const V3 wo_local = rec->ToLocal(rec->wo);
        const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local);
        float att = (F * D * G) / (4.0f * n_dot_o);
        if (alpha_u == alpha_v) att += DielectricMultiScatter(s, d, Dot(-rec->wi, rec->n), n_dot_o, alpha_u, rec->inside, true);
        rec->att = att * TexColor(s, d.id_specular_reflectance, rec->uv);
    } else {
        const V3 wo_local = rec->ToLocal(-rec->wo);
        const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local);
        float att = ((fabsf(h_dot_i) * fabsf(h_dot_o)) * ((1.0f - F) * G * D)) / fabsf(n_dot_o * Sqr(eta_inv * h_dot_i + h_dot_o));
        if (alpha_u == alpha_v) att += DielectricMultiScatter(s, d, Dot(rec->n, -rec->wi), n_dot_o, alpha_u, rec->inside, false);
        att *= Sqr(eta);
        rec->att = att * TexColor(s, d.id_specular_transmittance, rec->uv);
    }
}

// thin_dielectric.cpp:11-69
__device__ __forceinline__ void SampleThinDielectric(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    V3 h_local = mk3(0.0f);
    float D = 0;
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x;
    const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
    SampleGgx(xi_0, xi_1, alpha_u, alpha_v, &h_local, &D);
    const V3 h_world = rec->ToWorld(h_local);
    const float h_dot_o = Dot(rec->wo, h_world);
    rec->pdf = D / (4.0f * h_dot_o);
    if (rec->pdf < kEpsilon) return;
    rec->wi = -Reflect(-rec->wo, h_world);
    const float n_dot_i = Dot(-rec->wi, rec->n);
    if (n_dot_i < kEpsilonFloat) return;
    const V3 wi_local = rec->ToLocal(-rec->wi), wo_local = rec->ToLocal(rec->wo);
    const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local),
                h_dot_i = Dot(-rec->wi, h_world), n_dot_o = wo_local.z;
    float F = FresnelSchlick(h_dot_i, d.reflectivity_s);
    if (F < 1.0f) F *= 2.0f / (1.0f + F);
    if (rng.Next() < F) {
        rec->pdf *= F;
        if (rec->pdf < kEpsilon) return;
        rec->att = mk3((F * D * G) / (4.0f * n_dot_o)) * TexColor(s, d.id_specular_reflectance, rec->uv);
    } else {
        rec->pdf *= 1.0f - F;
        if (rec->pdf < kEpsilon) return;
        rec->att = mk3(((1.0f - F) * D * G) / (4.0f * n_dot_o)) * TexColor(s, d.id_specular_transmittance, rec->uv);
        rec->wi = rec->wo;
    }
    rec->valid = true;
}
// thin_dielectric.cpp:71-124
__device__ __forceinline__ void EvaluateThinDielectric(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    bool reflect = true;
    V3 wo = rec->wo;
    float n_dot_o = Dot(rec->wo, rec->n);
    if (fabsf(n_dot_o) < kEpsilonFloat) return;
    V3 wo_local = rec->ToLocal(rec->wo);
    if (n_dot_o < 0.0f) {
        reflect = false;
        n_dot_o = -n_dot_o;
        wo_local.z = -wo_local.z;
        wo = rec->ToWorld(wo_local);
    }
    const V3 h_world = Normalize(-rec->wi + wo), h_local = rec->ToLocal(h_world);
    const float alpha_u = TexColor(s, d.id_roughness_u, rec->uv).x, alpha_v = TexColor(s, d.id_roughness_v, rec->uv).x,
                D = PdfGgx(alpha_u, alpha_v, h_local), h_dot_i = Dot(-rec->wi, h_world), h_dot_o = Dot(rec->wo, h_world);
    float F = FresnelSchlick(h_dot_i, d.reflectivity_s);
    if (F < 1.0f) F *= 2.0f / (1.0f + F);
    rec->pdf = reflect ? (F * D) / (4.0f * h_dot_o) : ((1.0f - F) * D) / (4.0f * h_dot_o);
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    const V3 wi_local = rec->ToLocal(-rec->wi);
    const float G = SmithG1Ggx(alpha_u, alpha_v, wi_local, h_local) * SmithG1Ggx(alpha_u, alpha_v, wo_local, h_local);
    if (reflect)
        rec->att = mk3((F * D * G) / (4.0f * n_dot_o)) * TexColor(s, d.id_specular_reflectance, rec->uv);
    else
        rec->att = mk3(((1.0f - F) * D * G) / (4.0f * n_dot_o)) * TexColor(s, d.id_specular_transmittance, rec->uv);
}

// plastic.cpp:11-95
__device__ __forceinline__ void SamplePlastic(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    const V3 kd = TexColor(s, d.id_diffuse_reflectance, rec->uv), ks = TexColor(s, d.id_specular_reflectance, rec->uv);
    const float weight_spec = (ks.x + ks.y + ks.z) / ((kd.x + kd.y + kd.z) + (ks.x + ks.y + ks.z));
    const float n_dot_o = Dot(rec->wo, rec->n), kr_o = FresnelSchlick(n_dot_o, d.reflectivity_s);
    float kr_i = kr_o, pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * (1.0f - weight_spec);
    pdf_spec = pdf_spec / (pdf_spec + pdf_diff);
    pdf_diff = 1.0f - pdf_spec;
    V3 h_local = mk3(0.0f), h_world = mk3(0.0f);
    float D = 0;
    const float alpha = TexColor(s, d.id_roughness_u, rec->uv).x;
    float n_dot_i = 0;
    if (rng.Next() < pdf_spec) {
        const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
        SampleGgx(xi_0, xi_1, alpha, &h_local, &D);
        h_world = rec->ToWorld(h_local);
        rec->wi = -Reflect(-rec->wo, h_world);
        n_dot_i = Dot(-rec->wi, rec->n);
        if (n_dot_i < kEpsilonFloat) return;
        kr_i = FresnelSchlick(n_dot_i, d.reflectivity_s);
        pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * weight_spec;
        pdf_spec = pdf_spec / (pdf_spec + pdf_diff), pdf_diff = 1.0f - pdf_spec;
        const float h_dot_o = Dot(rec->wo, h_world);
        pdf_spec *= D / (4.0f * h_dot_o);
        pdf_diff *= Dot(-rec->wi, rec->n);
    } else {
        V3 wi_local = mk3(0.0f);
        float pdf_diff_local = 0.0f;
        const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
        SampleHemisCos(xi_0, xi_1, &wi_local, &pdf_diff_local);
        rec->wi = -rec->ToWorld(wi_local);
        n_dot_i = Dot(-rec->wi, rec->n);
        kr_i = FresnelSchlick(n_dot_i, d.reflectivity_s);
        pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * weight_spec;
        pdf_spec = pdf_spec / (pdf_spec + pdf_diff), pdf_diff = 1.0f - pdf_spec;
        h_world = Normalize(-rec->wi + rec->wo), h_local = rec->ToLocal(h_world);
        D = PdfGgx(alpha, h_local);
        const float h_dot_o = Dot(rec->wo, h_world);
        pdf_spec *= D / (4.0f * h_dot_o);
        pdf_diff *= pdf_diff_local;
    }
    rec->pdf = pdf_spec + pdf_diff;
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    if (pdf_spec > kEpsilon) {
        const V3 wi_local = rec->ToLocal(-rec->wi), wo_local = rec->ToLocal(rec->wo);
        const float h_dot_i = Dot(-rec->wi, h_world), F = FresnelSchlick(h_dot_i, d.reflectivity_s),
                    G = SmithG1Ggx(alpha, wo_local, h_local) * SmithG1Ggx(alpha, wi_local, h_local);
        rec->att += mk3((F * D * G) / (4.0f * n_dot_o)) * ks;
    }
    if (pdf_diff > kEpsilon) {
        V3 diff = kd * k1DivPi * n_dot_i;
        diff *= ((1.0f - kr_i) * (1.0f - kr_o)) / (1.0f - d.F_avg_s);
        rec->att += diff;
    }
}
// plastic.cpp:97-153
__device__ __forceinline__ void EvaluatePlastic(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    const float n_dot_o = Dot(rec->wo, rec->n);
    if (n_dot_o < kEpsilonFloat) return;
    const V3 kd = TexColor(s, d.id_diffuse_reflectance, rec->uv), ks = TexColor(s, d.id_specular_reflectance, rec->uv);
    const float weight_spec = (ks.x + ks.y + ks.z) / ((kd.x + kd.y + kd.z) + (ks.x + ks.y + ks.z));
    const float n_dot_i = Dot(-rec->wi, rec->n), kr_i = FresnelSchlick(n_dot_i, d.reflectivity_s);
    float pdf_spec = kr_i * weight_spec, pdf_diff = (1.0f - kr_i) * (1.0f - weight_spec);
    pdf_spec = pdf_spec / (pdf_spec + pdf_diff);
    pdf_diff = 1.0f - pdf_spec;
    const V3 h_world = Normalize(-rec->wi + rec->wo), h_local = rec->ToLocal(h_world);
    const float alpha = TexColor(s, d.id_roughness_u, rec->uv).x, D = PdfGgx(alpha, h_local), h_dot_o = Dot(rec->wo, h_world);
    pdf_spec *= D / (4.0f * h_dot_o);
    const V3 wo_local = rec->ToLocal(rec->wo);
    pdf_diff *= wo_local.z;
    rec->pdf = pdf_spec + pdf_diff;
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    if (pdf_spec > kEpsilon) {
        const V3 wi_local = rec->ToLocal(-rec->wi);
        const float h_dot_i = Dot(-rec->wi, h_world), F = FresnelSchlick(h_dot_i, d.reflectivity_s),
                    G = SmithG1Ggx(alpha, wo_local, h_local) * SmithG1Ggx(alpha, wi_local, h_local);
        rec->att += mk3((F * D * G) / (4.0f * n_dot_o)) * ks;
    }
    if (pdf_diff > kEpsilon) {
        V3 diff = kd * k1DivPi * n_dot_i;
        const float kr_o = FresnelSchlick(n_dot_o, d.reflectivity_s);
        diff *= ((1.0f - kr_i) * (1.0f - kr_o)) / (1.0f - d.F_avg_s);
        rec->att += diff;
    }
}

// bsdf.cpp:188-236 — kAreaLight has no case: the record stays invalid.
// ONLY >= 0 specialises a shading kernel to ONE BSDF type (the queue entries it is given are binned by the type of the
// surface they hit, wavefront.cu): the other models are not compiled in, which keeps the code a warp runs through short
// and contiguous (the generic kernel is 0.7 MB of code and stalls on instruction fetch when lanes scatter over it).
constexpr int kAnyBsdf = -1;

template <int ONLY>
__device__ __forceinline__ void BsdfSample(const DeviceScene &s, const DBsdf &d, Rng &rng, BsdfRec *rec) {
    switch (ONLY == kAnyBsdf ? d.type : static_cast<uint32_t>(ONLY)) {
    case B200PT_BSDF_DIFFUSE: SampleDiffuse(s, d, rng, rec); break;
    case B200PT_BSDF_ROUGH_DIFFUSE: SampleRoughDiffuse(s, d, rng, rec); break;
    case B200PT_BSDF_CONDUCTOR: SampleConductor(s, d, rng, rec); break;
    case B200PT_BSDF_DIELECTRIC: SampleDielectric(s, d, rng, rec); break;
    case B200PT_BSDF_THIN_DIELECTRIC: SampleThinDielectric(s, d, rng, rec); break;
    case B200PT_BSDF_PLASTIC: SamplePlastic(s, d, rng, rec); break;
    }
}
template <int ONLY>
__device__ __forceinline__ void BsdfEvaluate(const DeviceScene &s, const DBsdf &d, BsdfRec *rec) {
    switch (ONLY == kAnyBsdf ? d.type : static_cast<uint32_t>(ONLY)) {
    case B200PT_BSDF_DIFFUSE: EvaluateDiffuse(s, d, rec); break;
    case B200PT_BSDF_ROUGH_DIFFUSE: EvaluateRoughDiffuse(s, d, rec); break;
    case B200PT_BSDF_CONDUCTOR: EvaluateConductor(s, d, rec); break;
    case B200PT_BSDF_DIELECTRIC: EvaluateDielectric(s, d, rec); break;
    case B200PT_BSDF_THIN_DIELECTRIC: EvaluateThinDielectric(s, d, rec); break;
    case B200PT_BSDF_PLASTIC: EvaluatePlastic(s, d, rec); break;
    }
}

// path.cpp:238-266
template <int ONLY>
__device__ __forceinline__ BsdfRec EvaluateRayPath(const DeviceScene &s, V3 wi, V3 wo, const Surf &hit, const DBsdf *bsdf) {
    BsdfRec rec;
    rec.wi = wi, rec.wo = wo, rec.uv = hit.uv, rec.pos = hit.pos;
    if (bsdf) {
        rec.inside = hit.inside;
        rec.n = hit.n, rec.t = hit.t, rec.b = hit.b;
        if (Dot(-wi, hit.n) < 0.0f) {
            rec.inside = !rec.inside;
            rec.n = -rec.n;
        }
        BsdfEvaluate<ONLY>(s, *bsdf, &rec);
    } else {
        rec.pdf = 1;
        rec.att = mk3(1.0f);
        rec.valid = true;
    }
    return rec;
}
// path.cpp:268-296
template <int ONLY>
__device__ __forceinline__ BsdfRec SampleRayPath(const DeviceScene &s, V3 wo, const Surf &hit, const DBsdf *bsdf, Rng &rng) {
    BsdfRec rec;
    rec.wo = wo, rec.uv = hit.uv, rec.pos = hit.pos;
    if (bsdf != nullptr) {
        rec.inside = hit.inside;
        rec.n = hit.n, rec.t = hit.t, rec.b = hit.b;
        if (Dot(wo, hit.n) < 0.0f) {
            rec.inside = !rec.inside;
            rec.n = -rec.n;
        }
        BsdfSample<ONLY>(s, *bsdf, rng, &rec);
    } else {
        rec.wi = wo;
        rec.pdf = 1.0f;
        rec.att = mk3(1.0f);
        rec.valid = true;
    }
    return rec;
}

// ---------------------------------------------------------------------------------------------
// Emitters (emitters/*.cpp)
// ---------------------------------------------------------------------------------------------
struct EmitterRec { // emitter.hpp:49-55
    bool valid = false, harsh = true;
    float distance = kMaxFloat;
    V3 wi = {0, 0, 0};
};

__device__ __forceinline__ V2 DirToLatLong(V3 dir) {
    float phi = 0, theta = 0;
    CartesianToSpherical(dir, &theta, &phi, nullptr);
    return {phi * k1Div2Pi, theta * k1DivPi};
}

// emitter.cpp:177-203 and the per-type Sample functions
__device__ __forceinline__ EmitterRec EmitterSample(const DeviceScene &s, const DEmitter &e, V3 origin, float xi_0, float xi_1) {
    EmitterRec rec;
    switch (e.type) {
    case B200PT_EMIT_POINT: { // point_light.cpp:8-19
        const V3 vec = origin - mk3(e.position);
        rec = {true, true, Length(vec), Normalize(vec)};
        break;
    }
    case B200PT_EMIT_SPOT: { // spot_light.cpp:8-24
        const V3 vec = origin - mk3(e.position);
        const V3 wi = Normalize(vec), dir_local = XformVector(e.to_local, wi);
        if (dir_local.z >= e.cos_cutoff_angle) rec = {true, true, Length(vec), wi};
        break;
    }
    case B200PT_EMIT_DIRECTIONAL: // directional_light.cpp:8-18
        rec = {true, true, kMaxFloat, mk3(e.direction)};
        break;
    case B200PT_EMIT_SUN: // sun.cpp:8-18
        rec = {true, true, kMaxFloat, LocalToWorld(SampleConeUniform(e.cos_cutoff_angle, xi_0, xi_1), mk3(e.direction))};
        break;
    case B200PT_EMIT_ENVMAP: { // envmap.cpp:70-88 with the table wiring of emitter.cpp:166-175 (Q9)
        const float *cdf_rows = s.envmap_tables + e.env_cdf_rows, *cdf_cols = s.envmap_tables + e.env_cdf_cols;
        const uint32_t row = BinarySearch(e.env_height + 1, cdf_rows, xi_0) - 1;
        const uint32_t col = BinarySearch(e.env_width + 1, cdf_cols + row * (e.env_width + 1), xi_1) - 1;
        const V3 vec_local = SphericalToCartesian(row * kPi / e.env_height, col * k2Pi / e.env_width, 1);
        rec = {true, false, kMaxFloat, XformVector(e.to_world, vec_local)};
        break;
    }
    case B200PT_EMIT_CONSTANT: // constant_light.cpp:8-19
        rec = {true, false, kMaxFloat, SampleSphereUniform(xi_0, xi_1)};
        break;
    }
    return rec;
}

// emitter.cpp:205-231
__device__ __forceinline__ V3 EmitterEvaluateRec(const DeviceScene &s, const DEmitter &e, const EmitterRec &rec) {
    switch (e.type) {
    case B200PT_EMIT_POINT: return mk3(0.0f); // Q11: point_light.cpp:21-25
    case B200PT_EMIT_SPOT: { // spot_light.cpp:26-44
        const V3 dir = XformVector(e.to_local, rec.wi);
        V3 fall_off = mk3(1.0f);
        if (e.id_texture != kInvalid) {
            const V2 uv = {0.5f + 0.5f * dir.x / (dir.z * e.uv_factor), 0.5f + 0.5f * dir.y / (dir.z * e.uv_factor)};
            fall_off *= TexColor(s, e.id_texture, uv);
        }
        if (dir.z < e.cos_beam_width) fall_off *= (e.cutoff_angle - acosf(dir.z)) * e.transition_width_rcp;
        return mk3(e.radiance) * fall_off * Sqr(1.0f / rec.distance);
    }
    case B200PT_EMIT_DIRECTIONAL:
    case B200PT_EMIT_SUN:
    case B200PT_EMIT_CONSTANT: return mk3(e.radiance);
    case B200PT_EMIT_ENVMAP: { // envmap.cpp:90-98: looked up at -dir
        const V3 dir = XformVector(e.to_local, rec.wi);
        return TexColor(s, e.id_texture, DirToLatLong(-dir));
    }
    }
    return mk3(0.0f);
}

// emitter.cpp:233-249
__device__ __forceinline__ V3 EmitterEvaluateDir(const DeviceScene &s, const DEmitter &e, V3 look_dir) {
    switch (e.type) {
    case B200PT_EMIT_SUN: return TexColor(s, e.id_texture, DirToLatLong(look_dir)); // sun.cpp:26-32
    case B200PT_EMIT_ENVMAP: return TexColor(s, e.id_texture, DirToLatLong(XformVector(e.to_local, look_dir)));
    case B200PT_EMIT_CONSTANT: return mk3(e.radiance);
    }
    return mk3(0.0f);
}

// emitter.cpp:251-261, envmap.cpp:109-133, constant_light.cpp:33-36
__device__ __forceinline__ float EmitterPdf(const DeviceScene &s, const DEmitter &e, V3 look_dir) {
    if (e.type == B200PT_EMIT_CONSTANT) return k1Div4Pi;
    if (e.type != B200PT_EMIT_ENVMAP) return 0;
    const V3 dir = XformVector(e.to_local, look_dir);
    float phi = 0, theta = 0;
    CartesianToSpherical(dir, &theta, &phi, nullptr);
    const V2 uv = {phi * k1Div2Pi, theta * k1DivPi};
    const V3 color = TexColor(s, e.id_texture, uv);
    const float lum = 0.2126f * color.x + 0.7152f * color.y + 0.0722f * color.z;
    const float *weight_rows = s.envmap_tables + e.env_weight_rows;
    const float row = fminf(fmaxf(uv.u * e.env_height, 0), e.env_height - 1); // Q9: u, not v
    const int row_int = static_cast<int>(row);
    const float t = row - row_int;
    const float w = (t == 0) ? __ldg(weight_rows + row_int) : Lerp(__ldg(weight_rows + row_int), __ldg(weight_rows + row_int + 1), t);
    return lum * w * e.env_normalization / fmaxf(fabsf(sinf(theta)), 1e-4f);
}

// ---------------------------------------------------------------------------------------------
// Media (medium/*.cpp)
// ---------------------------------------------------------------------------------------------
struct MediumRec { // medium.hpp:54-61 (Q10: pdf starts at 1 and the branches use +=)
    bool valid = false, scattered = false;
    float pdf = 1.0f, distance = 0;
    V3 att = {1.0f, 1.0f, 1.0f};
};

// homogeneous.cpp:9-51
__device__ __forceinline__ void MediumSample(const DMedium &m, float max_distance, Rng &rng, MediumRec *rec) {
    float xi_0 = rng.Next();
    const V3 sigma_t = mk3(m.sigma_t);
    if (xi_0 < m.sampling_weight) {
        xi_0 /= m.sampling_weight;
        const int channel = static_cast<int>(rng.Next() * 3);
        rec->distance = -logf(1.0f - xi_0) / Comp(sigma_t, channel);
        if (rec->distance < max_distance) {
            rec->pdf += sigma_t.x * expf(-sigma_t.x * rec->distance);
            rec->pdf += sigma_t.y * expf(-sigma_t.y * rec->distance);
            rec->pdf += sigma_t.z * expf(-sigma_t.z * rec->distance);
            rec->pdf *= m.sampling_weight * (1.0f / 3.0f);
            rec->scattered = true;
        }
    }
    if (!rec->scattered) {
        rec->distance = max_distance;
        rec->pdf = 0;
        rec->pdf += expf(-sigma_t.x * rec->distance);
        rec->pdf += expf(-sigma_t.y * rec->distance);
        rec->pdf += expf(-sigma_t.z * rec->distance);
        rec->pdf = m.sampling_weight * (1.0f / 3.0f) * rec->pdf + (1.0f - m.sampling_weight);
    }
    rec->att = {expf(-sigma_t.x * rec->distance), expf(-sigma_t.y * rec->distance), expf(-sigma_t.z * rec->distance)};
    if (rec->att.x > kEpsilonFloat || rec->att.y > kEpsilonFloat || rec->att.z > kEpsilonFloat) rec->valid = true;
    if (rec->scattered) rec->att *= mk3(m.sigma_s);
}
// homogeneous.cpp:53-82
__device__ __forceinline__ void MediumEvaluate(const DMedium &m, MediumRec *rec) {
    const V3 sigma_t = mk3(m.sigma_t);
    rec->att = {expf(-sigma_t.x * rec->distance), expf(-sigma_t.y * rec->distance), expf(-sigma_t.z * rec->distance)};
    if (rec->att.x > kEpsilonFloat || rec->att.y > kEpsilonFloat || rec->att.z > kEpsilonFloat) rec->valid = true;
    if (!rec->valid) return;
    if (rec->scattered) {
        rec->pdf += sigma_t.x * rec->att.x;
        rec->pdf += sigma_t.y * rec->att.y;
        rec->pdf += sigma_t.z * rec->att.z;
        rec->pdf *= m.sampling_weight * (1.0f / 3.0f);
        rec->att *= mk3(m.sigma_s);
    } else {
        rec->pdf += rec->att.x;
        rec->pdf += rec->att.y;
        rec->pdf += rec->att.z;
        rec->pdf = m.sampling_weight * (1.0f / 3.0f) * rec->pdf + (1.0f - m.sampling_weight);
    }
}

struct PhaseRec { // medium.hpp:26-33
    bool valid = false;
    float pdf = 0;
    V3 wi = {0, 0, 0}, wo = {0, 0, 0}, att = {0, 0, 0};
};

__device__ __forceinline__ void HgValue(V3 g, float cos_theta, PhaseRec *rec) {
    const V3 temp = 1.0f + Sqr(g) + 2.0f * cos_theta * g;
    rec->att = k1Div4Pi * (1.0f - Sqr(g)) / (temp * Sqrt3(temp));
    rec->pdf = (rec->att.x + rec->att.y + rec->att.z) * (1.0f / 3.0f);
}
// medium.cpp:76-87, isotropic.cpp:9-15, henyey_greenstein.cpp:9-43
__device__ __forceinline__ void PhaseSample(const DMedium &m, Rng &rng, PhaseRec *rec) {
    if (m.phase_type == B200PT_PHASE_ISOTROPIC) {
        rec->valid = true;
        rec->att = mk3(k1Div4Pi);
        rec->pdf = k1Div4Pi;
        const float xi_1 = rng.Next(), xi_0 = rng.Next(); // drawn in the reference's order (Q16: GCC evaluates call arguments right to left)
        rec->wi = SampleSphereUniform(xi_0, xi_1);
        return;
    }
    const V3 g = mk3(m.g);
    const int channel = static_cast<int>(rng.Next() * 3);
    const float gc = Comp(g, channel);
    float cos_theta;
    if (fabsf(gc) < kEpsilonFloat) {
        cos_theta = 1.0f - 2.0f * rng.Next();
    } else {
        const float sqr_term = (1.0f - Sqr(gc)) / (1.0f - gc + 2.0f * gc * rng.Next());
        cos_theta = (1.0f + Sqr(gc) - Sqr(sqr_term)) / (2.0f * gc);
    }
    HgValue(g, cos_theta, rec);
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
    const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - Sqr(cos_theta)));
    const float phi = k2Pi * rng.Next();
    rec->wi = -LocalToWorld(mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta), rec->wo);
}
// medium.cpp:63-74, isotropic.cpp:17-22, henyey_greenstein.cpp:45-60
__device__ __forceinline__ void PhaseEvaluate(const DMedium &m, PhaseRec *rec) {
    if (m.phase_type == B200PT_PHASE_ISOTROPIC) {
        rec->valid = true;
        rec->att = mk3(k1Div4Pi);
        rec->pdf = k1Div4Pi;
        return;
    }
    HgValue(mk3(m.g), Dot(-rec->wi, rec->wo), rec);
    if (rec->pdf < kEpsilon) return;
    rec->valid = true;
}

} // namespace b200pt
