// vecmath.cuh — small float3 algebra + the sampling helpers of the reference
// (src/tensor/vec3.cpp, src/utils/math.cpp), device side.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "device_scene.h"

namespace b200pt {

constexpr float kPi = 3.141592653589793f;
constexpr float k2Pi = 3.141592653589793f * 2.0f;
constexpr float kPiDiv2 = 3.141592653589793f * 0.5f;
constexpr float kPiDiv4 = 3.141592653589793f * 0.25f;
constexpr float k1DivPi = 1.0f / kPi;
constexpr float k1Div2Pi = 1.0f / k2Pi;
constexpr float k1Div4Pi = 1.0f / (4.0f * kPi);
constexpr float kEpsilonFloat = 1.1920928955078125e-7f; // defs.hpp:24
constexpr float kEpsilonDistance = 1e-4f;               // defs.hpp:25
constexpr float kEpsilon = 0.01f;                       // defs.hpp:26
constexpr float kMaxFloat = 3.402823466e+38f;

struct V3 {
    float x, y, z;
};
struct V2 {
    float u, v;
};

__device__ __forceinline__ V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 mk3(float s) { return V3{s, s, s}; }
__device__ __forceinline__ V3 mk3(const F3 &f) { return V3{f.x, f.y, f.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator/(V3 a, V3 b) { return {a.x * (1.0f / b.x), a.y * (1.0f / b.y), a.z * (1.0f / b.z)}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) {
    const float k = 1.0f / s;
    return {a.x * k, a.y * k, a.z * k};
}
__device__ __forceinline__ V3 operator+(V3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
__device__ __forceinline__ V3 operator+(float s, V3 a) { return {a.x + s, a.y + s, a.z + s}; }
__device__ __forceinline__ V3 operator-(float s, V3 a) { return {s - a.x, s - a.y, s - a.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 &operator+=(V3 &a, V3 b) {
    a.x += b.x, a.y += b.y, a.z += b.z;
    return a;
}
__device__ __forceinline__ V3 &operator*=(V3 &a, V3 b) {
    a.x *= b.x, a.y *= b.y, a.z *= b.z;
    return a;
}
__device__ __forceinline__ V3 &operator*=(V3 &a, float s) {
    a.x *= s, a.y *= s, a.z *= s;
    return a;
}
__device__ __forceinline__ float Dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 Cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float Length(V3 a) { return sqrtf(Dot(a, a)); }
__device__ __forceinline__ V3 Normalize(V3 a) { return a * (1.0f / Length(a)); }
__device__ __forceinline__ V3 Sqrt3(V3 a) { return {sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
__device__ __forceinline__ float Sqr(float x) { return x * x; }
__device__ __forceinline__ V3 Sqr(V3 a) { return a * a; }
__device__ __forceinline__ float MaxComp(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
__device__ __forceinline__ float Lerp(float a, float b, float t) { return (1.0f - t) * a + t * b; }
__device__ __forceinline__ V3 Lerp(V3 a, V3 b, float t) { return (1.0f - t) * a + t * b; }
__device__ __forceinline__ float Comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// forward declarations used by the transform helpers
__device__ __forceinline__ V3 Normalize(V3 a);

__device__ __forceinline__ V3 XformPoint(const Affine &m, V3 p) {
    return {m.m[0] * p.x + m.m[1] * p.y + m.m[2] * p.z + m.m[3], m.m[4] * p.x + m.m[5] * p.y + m.m[6] * p.z + m.m[7],
            m.m[8] * p.x + m.m[9] * p.y + m.m[10] * p.z + m.m[11]};
}
// TransformVector (mat4.cpp:270-273) returns Vec4::direction(), i.e. the NORMALISED transformed vector
// (vec4.hpp:55) — normals, local ray directions and env-map directions all go through this.
__device__ __forceinline__ V3 XformVector(const Affine &m, V3 v) {
    return Normalize(V3{m.m[0] * v.x + m.m[1] * v.y + m.m[2] * v.z, m.m[4] * v.x + m.m[5] * v.y + m.m[6] * v.z,
                        m.m[8] * v.x + m.m[9] * v.y + m.m[10] * v.z});
}

// math.cpp:8-13
__device__ __forceinline__ float MisWeight(float pdf1, float pdf2) {
    pdf1 *= pdf1;
    pdf2 *= pdf2;
    return pdf1 / (pdf1 + pdf2);
}

// math.cpp:15-22
__device__ __forceinline__ V3 SampleConeUniform(float cos_cutoff, float xi_0, float xi_1) {
    const float cos_theta = 1.0f - (1.0f - cos_cutoff) * xi_0, phi = 2.0f * kPi * xi_1;
    const float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
    float s, c;
    sincosf(phi, &s, &c);
    return {sin_theta * c, sin_theta * s, cos_theta};
}

// math.cpp:24-29
__device__ __forceinline__ V3 SampleSphereUniform(float xi_0, float xi_1) {
    const float cos_theta = 1.0f - 2.0f * xi_0, phi = k2Pi * xi_1;
    const float sin_theta = sqrtf(1.0f - Sqr(cos_theta));
    float s, c;
    sincosf(phi, &s, &c);
    return {sin_theta * c, sin_theta * s, cos_theta};
}

// math.cpp:31-38
__device__ __forceinline__ void SampleHemisCos(float xi_0, float xi_1, V3 *vec, float *pdf) {
    const float cos_theta = sqrtf(xi_0), phi = k2Pi * xi_1;
    const float sin_theta = sqrtf(1.0f - Sqr(cos_theta));
    float s, c;
    sincosf(phi, &s, &c);
    *vec = {sin_theta * c, sin_theta * s, cos_theta};
    *pdf = k1DivPi * cos_theta;
}

// math.cpp:40-55
__device__ __forceinline__ uint32_t BinarySearch(uint32_t num, const float *cdf, float target) {
    uint32_t begin = 0, end = num, middle;
    while (begin + 1 != end) {
        middle = (begin + end) >> 1;
        const float c = __ldg(cdf + middle);
        if (c < target)
            begin = middle;
        else if (c > target)
            end = middle;
        else
            return middle;
    }
    return end;
}

// math.cpp:57-99
__device__ __forceinline__ bool SolveQuadratic(float a, float b, float c, float *x0, float *x1) {
    if (a == 0.0f) {
        if (b != 0.0f) {
            *x0 = *x1 = -c / b;
            return true;
        }
        return false;
    }
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0.0f) return false;
    float temp;
    const float sqrt_discrim = sqrtf(discrim);
    if (b < 0.0f)
        temp = -0.5f * (b - sqrt_discrim);
    else
        temp = -0.5f * (b + sqrt_discrim);
    *x0 = temp / a;
    *x1 = c / temp;
    if (*x0 > *x1) {
        const float t = *x0;
        *x0 = *x1;
        *x1 = t;
    }
    return true;
}

// math.cpp:102-119 — y-up convention: theta from +y, phi = atan2(z, x)
__device__ __forceinline__ void CartesianToSpherical(V3 vec, float *theta, float *phi, float *r) {
    if (r != nullptr) *r = Length(vec);
    vec = Normalize(vec);
    *theta = acosf(fminf(1.0f, fmaxf(-1.0f, vec.y)));
    if (vec.z == 0 && vec.x == 0) {
        *phi = 0;
    } else {
        *phi = atan2f(vec.z, vec.x);
        if (*phi < 0.0f) *phi += 2.0f * kPi;
    }
}

// math.cpp:122-128
__device__ __forceinline__ V3 SphericalToCartesian(float theta, float phi, float r) {
    const float sin_theta = sinf(theta);
    return {r * sinf(phi) * sin_theta, r * cosf(theta), r * cosf(phi) * sin_theta};
}

// math.cpp:130-146
__device__ __forceinline__ V3 LocalToWorld(V3 local, V3 up) {
    V3 C;
    if (sqrtf(Sqr(up.x) + Sqr(up.z)) > kEpsilonFloat) {
        const float len_inv = 1.0f / sqrtf(Sqr(up.x) + Sqr(up.z));
        C = {up.z * len_inv, 0, -up.x * len_inv};
    } else {
        const float len_inv = 1.0f / sqrtf(Sqr(up.y) + Sqr(up.z));
        C = {0, up.z * len_inv, -up.y * len_inv};
    }
    const V3 B = Normalize(Cross(C, up));
    return Normalize(local.x * B + local.y * C + local.z * up);
}

// ray.cpp:49-69
__device__ __forceinline__ V3 Reflect(V3 wi, V3 normal) { return Normalize(wi - 2.0f * Dot(wi, normal) * normal); }
__device__ __forceinline__ bool Refract(V3 wi, V3 normal, float eta_inv, V3 *wt) {
    const float cos_theta = fabsf(Dot(wi, normal));
    const float k = 1.0f - Sqr(eta_inv) * (1.0f - Sqr(cos_theta));
    if (k < 0) return false;
    *wt = Normalize(eta_inv * wi + (eta_inv * cos_theta - sqrtf(k)) * normal);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG: the state is (pixel, sample, dimension) — nothing touches HBM.
// Replaces the per-pixel LCG of the reference (math.hpp:57-63), so parity is statistical.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 Philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

#if !defined(B200PT_RNG_REPLAY)
struct Rng {
    uint32_t c0, c1, c2, domain;
    uint2 key;
    uint4 buf;
    int idx;
    // (pixel, sample, depth) identify the path vertex; each vertex owns 2^16 blocks of 4 floats.
    // `dom` separates independent consumers of the same vertex (shading / alpha tests of the closest-hit ray / of the shadow ray).
    __device__ __forceinline__ Rng(uint32_t pixel, uint32_t sample, uint32_t depth, uint2 k, uint32_t dom = 0x5eedu)
        : c0(pixel), c1(sample), c2(depth << 16), domain(dom), key(k), buf(make_uint4(0, 0, 0, 0)), idx(4) {}
    __device__ __forceinline__ float Next() {
        if (idx == 4) {
            buf = Philox4x32_10(make_uint4(c0, c1, c2, domain), key);
            ++c2;
            idx = 0;
        }
        const uint32_t v = idx == 0 ? buf.x : (idx == 1 ? buf.y : (idx == 2 ? buf.z : buf.w));
        ++idx;
        return static_cast<float>(v >> 8) * (1.0f / 16777216.0f); // 24-bit mantissa, as math.hpp:60-62
    }
    // One number per (stream, primitive), independent of how many were drawn before: the alpha tests inside traversal.
    __device__ __forceinline__ float ForPrimitive(uint32_t prim) const {
        const uint4 r = Philox4x32_10(make_uint4(c0, c1, c2 ^ 0x80000000u, domain ^ (prim * 0x9E3779B9u)), key);
        return static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);
    }
};
#else
// Test build only (csrc/debug_eval.cu, -DB200PT_RNG_REPLAY): the reference's own generator — the per-pixel LCG RandomFloat
// (include/csrt/utils/math.hpp:57-63) — behind the same interface, so that the sampling routines can be compared with the
// reference's POINTWISE from a common seed (the device code draws its numbers in the order GCC evaluates the reference's
// call arguments, right to left).  The inline namespace gives every function that takes an Rng a different mangled
// name from the product's, so the two definitions never meet.
inline namespace lcg_replay {
struct Rng {
    uint32_t state;
    __device__ __forceinline__ explicit Rng(uint32_t seed) : state(seed) {}
    __device__ __forceinline__ Rng(uint32_t, uint32_t, uint32_t, uint2, uint32_t = 0) : state(0) {}
    __device__ __forceinline__ float Next() {
        state = state * 1664525u + 1013904223u;
        return static_cast<float>(state & 0x00FFFFFFu) / static_cast<float>(0x01000000u);
    }
    __device__ __forceinline__ float ForPrimitive(uint32_t prim) const {
        const uint32_t s = (state ^ prim) * 1664525u + 1013904223u;
        return static_cast<float>(s & 0x00FFFFFFu) / static_cast<float>(0x01000000u);
    }
};
} // namespace lcg_replay
#endif

// math.hpp:29-41
__device__ __forceinline__ float VanDerCorput2(uint32_t index) {
    return static_cast<float>(__brev(index) >> 8) * (1.0f / 16777216.0f);
}

} // namespace b200pt
