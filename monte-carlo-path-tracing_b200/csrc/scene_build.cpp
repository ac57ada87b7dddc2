// scene_build.cpp — b200pt_scene_desc -> HostScene (host side of b200pt_create).
//
// Restates, for a GPU layout, what the reference derives on the host between the
// parsed config and the first ray:
//   geometry bake + per-triangle attributes   src/rtcore/scene.cpp:15-111, 247-324
//   analytic primitives and their "areas"     src/rtcore/scene.cpp:326-472 (Q3)
//   per-instance 1/area                       src/rtcore/scene.cpp:493-495
//   area-light maps + un-normalised CDF       src/renderer/renderer.cpp:271-304 (Q4)
//   BSDF / medium / emitter derived fields    bsdfs/bsdf.cpp:112-186, medium/medium.cpp:6-39,
//                                             emitters/emitter.cpp:122-175
//   env-map tables                            emitters/envmap.cpp:20-68, renderer.cpp:571-611 (Q9)
//   Kulla-Conty LUTs                          bsdfs/kulla_conty.cpp:13-80
// The acceleration structure is NOT the reference's per-instance LBVH: one binned-SAH
// BVH2 over all world-space triangles, 2 child boxes per 64-byte node, BFS-ordered top.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <numeric>
#include <queue>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "bvh_gpu.hpp"
#include "bvh_wide.hpp"
#include "host_scene.hpp"

namespace b200pt {

namespace {

constexpr float kPi = 3.141592653589793f;
constexpr float k2Pi = kPi * 2.0f;
constexpr float k1DivPi = 1.0f / kPi;
constexpr float kFltMax = 3.402823466e+38f;
constexpr uint32_t kBvh2StackSize = 64; // = kStackSize of traverse.cuh

struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline float Dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 Cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, -a.x * b.z + a.z * b.x, a.x * b.y - a.y * b.x}; }
inline float Length(V3 a) { return sqrtf(Dot(a, a)); }
inline V3 Normalize(V3 a) { return a * (1.0f / Length(a)); }
inline V3 Min(V3 a, V3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline V3 Max(V3 a, V3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline V3 Load3(const float *p) { return {p[0], p[1], p[2]}; }

struct M4 {
    float m[4][4];
};

M4 Identity() {
    M4 r{};
    for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0f;
    return r;
}
M4 LoadM4(const float *p) {
    M4 r;
    memcpy(r.m, p, sizeof(r.m));
    return r;
}
M4 Transpose(const M4 &a) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
M4 Mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
// General 4x4 inverse by cofactor expansion (same quantity as Mat4::Inverse, mat4.cpp:110-168).
M4 Inverse(const M4 &a) {
    const float *s = &a.m[0][0];
    float inv[16];
    inv[0] = s[5] * s[10] * s[15] - s[5] * s[11] * s[14] - s[9] * s[6] * s[15] + s[9] * s[7] * s[14] + s[13] * s[6] * s[11] - s[13] * s[7] * s[10];
    inv[4] = -s[4] * s[10] * s[15] + s[4] * s[11] * s[14] + s[8] * s[6] * s[15] - s[8] * s[7] * s[14] - s[12] * s[6] * s[11] + s[12] * s[7] * s[10];
    inv[8] = s[4] * s[9] * s[15] - s[4] * s[11] * s[13] - s[8] * s[5] * s[15] + s[8] * s[7] * s[13] + s[12] * s[5] * s[11] - s[12] * s[7] * s[9];
    inv[12] = -s[4] * s[9] * s[14] + s[4] * s[10] * s[13] + s[8] * s[5] * s[14] - s[8] * s[6] * s[13] - s[12] * s[5] * s[10] + s[12] * s[6] * s[9];
    inv[1] = -s[1] * s[10] * s[15] + s[1] * s[11] * s[14] + s[9] * s[2] * s[15] - s[9] * s[3] * s[14] - s[13] * s[2] * s[11] + s[13] * s[3] * s[10];
    inv[5] = s[0] * s[10] * s[15] - s[0] * s[11] * s[14] - s[8] * s[2] * s[15] + s[8] * s[3] * s[14] + s[12] * s[2] * s[11] - s[12] * s[3] * s[10];
    inv[9] = -s[0] * s[9] * s[15] + s[0] * s[11] * s[13] + s[8] * s[1] * s[15] - s[8] * s[3] * s[13] - s[12] * s[1] * s[11] + s[12] * s[3] * s[9];
    inv[13] = s[0] * s[9] * s[14] - s[0] * s[10] * s[13] - s[8] * s[1] * s[14] + s[8] * s[2] * s[13] + s[12] * s[1] * s[10] - s[12] * s[2] * s[9];
    inv[2] = s[1] * s[6] * s[15] - s[1] * s[7] * s[14] - s[5] * s[2] * s[15] + s[5] * s[3] * s[14] + s[13] * s[2] * s[7] - s[13] * s[3] * s[6];
    inv[6] = -s[0] * s[6] * s[15] + s[0] * s[7] * s[14] + s[4] * s[2] * s[15] - s[4] * s[3] * s[14] - s[12] * s[2] * s[7] + s[12] * s[3] * s[6];
    inv[10] = s[0] * s[5] * s[15] - s[0] * s[7] * s[13] - s[4] * s[1] * s[15] + s[4] * s[3] * s[13] + s[12] * s[1] * s[7] - s[12] * s[3] * s[5];
    inv[14] = -s[0] * s[5] * s[14] + s[0] * s[6] * s[13] + s[4] * s[1] * s[14] - s[4] * s[2] * s[13] - s[12] * s[1] * s[6] + s[12] * s[2] * s[5];
    inv[3] = -s[1] * s[6] * s[11] + s[1] * s[7] * s[10] + s[5] * s[2] * s[11] - s[5] * s[3] * s[10] - s[9] * s[2] * s[7] + s[9] * s[3] * s[6];
    inv[7] = s[0] * s[6] * s[11] - s[0] * s[7] * s[10] - s[4] * s[2] * s[11] + s[4] * s[3] * s[10] + s[8] * s[2] * s[7] - s[8] * s[3] * s[6];
    inv[11] = -s[0] * s[5] * s[11] + s[0] * s[7] * s[9] + s[4] * s[1] * s[11] - s[4] * s[3] * s[9] - s[8] * s[1] * s[7] + s[8] * s[3] * s[5];
    inv[15] = s[0] * s[5] * s[10] - s[0] * s[6] * s[9] - s[4] * s[1] * s[10] + s[4] * s[2] * s[9] + s[8] * s[1] * s[6] - s[8] * s[2] * s[5];
    const float det = s[0] * inv[0] + s[1] * inv[4] + s[2] * inv[8] + s[3] * inv[12];
    const float k = 1.0f / det;
    M4 r;
    for (int i = 0; i < 16; ++i) (&r.m[0][0])[i] = inv[i] * k;
    return r;
}
V3 TransformPoint(const M4 &m, V3 p) {
    return {m.m[0][0] * p.x + m.m[0][1] * p.y + m.m[0][2] * p.z + m.m[0][3],
            m.m[1][0] * p.x + m.m[1][1] * p.y + m.m[1][2] * p.z + m.m[1][3],
            m.m[2][0] * p.x + m.m[2][1] * p.y + m.m[2][2] * p.z + m.m[2][3]};
}
// mat4.cpp:270-273: TransformVector returns Vec4::direction() = the NORMALISED vector (vec4.hpp:55).
V3 TransformVector(const M4 &m, V3 v) {
    return Normalize(V3{m.m[0][0] * v.x + m.m[0][1] * v.y + m.m[0][2] * v.z, m.m[1][0] * v.x + m.m[1][1] * v.y + m.m[1][2] * v.z,
                        m.m[2][0] * v.x + m.m[2][1] * v.y + m.m[2][2] * v.z});
}
Affine ToAffine(const M4 &m) {
    Affine a;
    memcpy(a.m, m.m, sizeof(a.m));
    return a;
}
M4 Translate(V3 v) {
    M4 r = Identity();
    r.m[0][3] = v.x, r.m[1][3] = v.y, r.m[2][3] = v.z;
    return r;
}
// math.cpp:148-165 (the Mat4 flavour of LocalToWorld used for cylinders)
M4 LocalToWorldMat(V3 up) {
    V3 C;
    if (sqrtf(up.x * up.x + up.z * up.z) > 1.1920929e-7f) {
        const float len_inv = 1.0f / sqrtf(up.x * up.x + up.z * up.z);
        C = {-up.z * len_inv, 0, up.x * len_inv};
    } else {
        const float len_inv = 1.0f / sqrtf(up.y * up.y + up.z * up.z);
        C = {0, -up.z * len_inv, up.y * len_inv};
    }
    const V3 B = Normalize(Cross(C, up));
    M4 r = Identity();
    r.m[0][0] = B.x, r.m[0][1] = B.y, r.m[0][2] = B.z;
    r.m[1][0] = C.x, r.m[1][1] = C.y, r.m[1][2] = C.z;
    r.m[2][0] = up.x, r.m[2][1] = up.y, r.m[2][2] = up.z;
    return r;
}

F3 ToF3(V3 v) { return {v.x, v.y, v.z}; }

// ---------------------------------------------------------------------------------------------
// Geometry
// ---------------------------------------------------------------------------------------------
// Runs fn(t, num_threads) on num_threads threads (the calling thread is thread 0).
template <typename Fn>
void ParallelRun(unsigned num_threads, Fn fn) {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < num_threads; ++t) pool.emplace_back([&fn, t, num_threads] { fn(t, num_threads); });
    fn(0u, num_threads);
    for (std::thread &th : pool) th.join();
}

// B200PT_VERBOSE_CREATE=1: wall time of every phase of BuildHostScene on stderr.
struct PhaseTimer {
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    bool on = getenv("B200PT_VERBOSE_CREATE") != nullptr;
    void operator()(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200pt create] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    }
};

unsigned BuildThreads() {
    static const unsigned threads = [] {
        const char *env = getenv("B200PT_BUILD_THREADS"); // experiments
        const unsigned want = env ? static_cast<unsigned>(atoi(env)) : std::thread::hardware_concurrency();
        return std::max(1u, std::min(32u, want));
    }();
    return threads;
}

// fn(i) for i in [0, n), in contiguous chunks over the build threads.
template <typename Fn>
void ParallelFor(size_t n, Fn fn) {
    const unsigned threads = n < 8192 ? 1u : BuildThreads();
    ParallelRun(threads, [&](unsigned t, unsigned nt) {
        const size_t b = n * t / nt, e = n * (t + 1) / nt;
        for (size_t i = b; i < e; ++i) fn(i);
    });
}

struct RawTriangle {
    V3 p[3], n[3], t[3];
    float uv[3][2];
    uint32_t inst;
    float area; // |e1 x e2| : twice the true area, as the reference uses (scene.cpp:48-49, Q3)
};

struct MeshView {
    const float *positions = nullptr, *normals = nullptr, *texcoords = nullptr, *tangents = nullptr,
                *bitangents = nullptr;
    const uint32_t *indices = nullptr;
    uint64_t num_vertices = 0, num_triangles = 0;
};

// scene.cpp:200-245: the built-in rectangle and cube meshes.
const float kRectUv[] = {0, 0, 1, 0, 1, 1, 0, 1};
const float kRectPos[] = {-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0};
const float kRectNrm[] = {0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1};
const uint32_t kRectIdx[] = {0, 1, 2, 2, 3, 0};
const float kCubeUv[] = {0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0,
                         0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0};
const float kCubePos[] = {1,  -1, -1, 1,  -1, 1,  -1, -1, 1,  -1, -1, -1, 1,  1,  -1, -1, 1,  -1, -1, 1,  1,  1,  1,  1,
                          1,  -1, -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1,  1,  -1, 1,  1,  1,  1,  -1, 1,  1,  -1, -1, 1,
                          -1, -1, 1,  -1, 1,  1,  -1, 1,  -1, -1, -1, -1, 1,  1,  -1, 1,  -1, -1, -1, -1, -1, -1, 1,  -1};
const float kCubeNrm[] = {0,  -1, 0, 0,  -1, 0, 0,  -1, 0, 0,  -1, 0, 0, 1, 0,  0, 1, 0,  0, 1, 0,  0, 1, 0,
                          1,  0,  0, 1,  0,  0, 1,  0,  0, 1,  0,  0, 0, 0, 1,  0, 0, 1,  0, 0, 1,  0, 0, 1,
                          -1, 0,  0, -1, 0,  0, -1, 0,  0, -1, 0,  0, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1};
const uint32_t kCubeIdx[] = {0,  1,  2,  3,  0,  2,  4,  5,  6,  7,  4,  6,  8,  9,  10, 11, 8,  10,
                             12, 13, 14, 15, 12, 14, 16, 17, 18, 19, 16, 18, 20, 21, 22, 23, 20, 22};

struct Box {
    V3 lo{kFltMax, kFltMax, kFltMax}, hi{-kFltMax, -kFltMax, -kFltMax};
    void Grow(V3 p) { lo = Min(lo, p), hi = Max(hi, p); }
    void Grow(const Box &b) { lo = Min(lo, b.lo), hi = Max(hi, b.hi); }
    float HalfArea() const {
        const V3 d = hi - lo;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
};

// scene.cpp:247-324 (bake to_world) + scene.cpp:15-111 (SetupMeshes).  Also leaves every triangle's box and box centre (what the BVH
// builder reads) while the triangle is still in cache.
bool AppendMesh(const MeshView &mesh, const M4 &to_world, uint32_t inst, RawVector<RawTriangle> *tris, std::vector<Box> *boxes,
                std::vector<V3> *centers, float *area_sum, std::string *error) {
    if (mesh.num_triangles == 0 || mesh.indices == nullptr) {
        *error = "cannot find vertex index info when adding instance to scene.";
        return false;
    }
    if (mesh.num_vertices == 0 || mesh.positions == nullptr) {
        *error = "cannot find vertex position info when adding instance to scene.";
        return false;
    }
    const uint64_t nv = mesh.num_vertices;
    RawVector<V3> pos(nv), nrm, tan, bit;
    ParallelFor(nv, [&](size_t i) { pos[i] = TransformPoint(to_world, Load3(mesh.positions + 3 * i)); });
    if (mesh.normals) {
        const M4 normal_to_world = Inverse(Transpose(to_world));
        nrm.resize(nv);
        ParallelFor(nv, [&](size_t i) { nrm[i] = TransformVector(normal_to_world, Load3(mesh.normals + 3 * i)); });
    }
    if (mesh.tangents) {
        tan.resize(nv);
        ParallelFor(nv, [&](size_t i) { tan[i] = TransformVector(to_world, Load3(mesh.tangents + 3 * i)); });
    }
    if (mesh.bitangents) {
        bit.resize(nv);
        ParallelFor(nv, [&](size_t i) { bit[i] = TransformVector(to_world, Load3(mesh.bitangents + 3 * i)); });
    }
    for (uint64_t k = 0; k < 3 * mesh.num_triangles; ++k)
        if (mesh.indices[k] >= nv) {
            *error = "mesh index out of range.";
            return false;
        }
    const size_t first_triangle = tris->size();
    tris->resize(first_triangle + mesh.num_triangles); // within the capacity BuildHostScene reserved
    boxes->resize(first_triangle + mesh.num_triangles);
    centers->resize(first_triangle + mesh.num_triangles);
    ParallelFor(mesh.num_triangles, [&](size_t f) {
        RawTriangle tri;
        tri.inst = inst;
        uint32_t idx[3];
        for (int j = 0; j < 3; ++j) idx[j] = mesh.indices[3 * f + j];
        if (!mesh.texcoords) {
            tri.uv[0][0] = 0, tri.uv[0][1] = 0, tri.uv[1][0] = 1, tri.uv[1][1] = 0, tri.uv[2][0] = 1, tri.uv[2][1] = 1;
        } else {
            for (int j = 0; j < 3; ++j) tri.uv[j][0] = mesh.texcoords[2 * idx[j]], tri.uv[j][1] = mesh.texcoords[2 * idx[j] + 1];
        }
        for (int j = 0; j < 3; ++j) tri.p[j] = pos[idx[j]];
        const V3 v0v1 = tri.p[1] - tri.p[0], v0v2 = tri.p[2] - tri.p[0];
        const V3 normal_geom = Cross(v0v1, v0v2);
        tri.area = Length(normal_geom);
        if (nrm.empty()) {
            const V3 n = Normalize(normal_geom);
            for (int j = 0; j < 3; ++j) tri.n[j] = n;
        } else {
            for (int j = 0; j < 3; ++j) tri.n[j] = nrm[idx[j]];
        }
        if (tan.empty() && bit.empty()) {
            const float d01u = tri.uv[1][0] - tri.uv[0][0], d01v = tri.uv[1][1] - tri.uv[0][1];
            const float d02u = tri.uv[2][0] - tri.uv[0][0], d02v = tri.uv[2][1] - tri.uv[0][1];
            const float r = 1.0f / (d01v * d02u - d01u * d02v);
            const V3 tangent = Normalize((d01v * v0v2 - d02v * v0v1) * r);
            for (int j = 0; j < 3; ++j) {
                const V3 b = Normalize(Cross(tri.n[j], tangent));
                tri.t[j] = Normalize(Cross(b, tri.n[j]));
            }
        } else if (tan.empty()) {
            for (int j = 0; j < 3; ++j) tri.t[j] = Normalize(Cross(bit[idx[j]], tri.n[j]));
        } else {
            for (int j = 0; j < 3; ++j) {
                const V3 b = Normalize(Cross(tri.n[j], tan[idx[j]]));
                tri.t[j] = Normalize(Cross(b, tri.n[j]));
            }
        }
        (*tris)[first_triangle + f] = tri;
        Box box;
        for (int j = 0; j < 3; ++j) box.Grow(tri.p[j]);
        (*boxes)[first_triangle + f] = box;
        (*centers)[first_triangle + f] = (box.lo + box.hi) * 0.5f;
    });
    float area_total = 0.0f; // summed in triangle order, as scene.cpp:48-49 does (Q3: the float sum is what the pdf uses)
    for (uint64_t f = 0; f < mesh.num_triangles; ++f) area_total += (*tris)[first_triangle + f].area;
    *area_sum = area_total;
    return true;
}

// ---------------------------------------------------------------------------------------------
// BVH: binned SAH, two child boxes per node.
// ---------------------------------------------------------------------------------------------
// Four floats in one SSE register (plain loops elsewhere): the builder's inner loop is min / max of box corners.
#if defined(__SSE2__)
struct alignas(16) F4 {
    float v[4];
};
inline F4 Min4(const F4 &a, const F4 &b) {
    F4 r;
    _mm_store_ps(r.v, _mm_min_ps(_mm_load_ps(a.v), _mm_load_ps(b.v)));
    return r;
}
inline F4 Max4(const F4 &a, const F4 &b) {
    F4 r;
    _mm_store_ps(r.v, _mm_max_ps(_mm_load_ps(a.v), _mm_load_ps(b.v)));
    return r;
}
#else
struct alignas(16) F4 {
    float v[4];
};
inline F4 Min4(const F4 &a, const F4 &b) {
    F4 r;
    for (int k = 0; k < 4; ++k) r.v[k] = a.v[k] < b.v[k] ? a.v[k] : b.v[k];
    return r;
}
inline F4 Max4(const F4 &a, const F4 &b) {
    F4 r;
    for (int k = 0; k < 4; ++k) r.v[k] = a.v[k] > b.v[k] ? a.v[k] : b.v[k];
    return r;
}
#endif
constexpr F4 kF4Max{{kFltMax, kFltMax, kFltMax, kFltMax}}, kF4Lowest{{-kFltMax, -kFltMax, -kFltMax, -kFltMax}};

struct BuildNode {
    Box box;
    int32_t left = -1, right = -1; // children (BuildNode indices) or -1 for leaf
    uint32_t first = 0, count = 0; // leaf range in the permuted triangle order
};

// Binned SAH (32 bins per axis, all three axes in one pass over the node's references).  The references are kept in an
// array that is partitioned in place, so every pass is a sequential scan; the bins also carry the bounds of the centres,
// so a child's box and centre box come out of the parent's bins and no node is scanned twice.  Large nodes are binned and
// partitioned by all threads together (the first levels of the tree hold most of the work), the subtrees below
// kParallelNode references are then built one per thread.
// Storage for up to `capacity` BuildNodes that is NOT touched until a node is handed out (a binary tree over n triangles may
// need 2n - 1 nodes, a SAH tree with 2-4 triangles per leaf uses about half: 66 MB of page faults and constructors on Dragon).
class BuildNodePool {
public:
    explicit BuildNodePool(size_t capacity) : data_(static_cast<BuildNode *>(malloc(std::max<size_t>(capacity, 1) * sizeof(BuildNode)))) {}
    ~BuildNodePool() { free(data_); }
    BuildNodePool(const BuildNodePool &) = delete;
    BuildNodePool &operator=(const BuildNodePool &) = delete;
    int32_t Allocate() {
        const int32_t id = used_.fetch_add(1);
        new (data_ + id) BuildNode();
        return id;
    }
    BuildNode &operator[](size_t i) { return data_[i]; }
    const BuildNode &operator[](size_t i) const { return data_[i]; }
    size_t size() const { return static_cast<size_t>(used_.load()); }

private:
    BuildNode *data_;
    std::atomic<int32_t> used_{0};
};

class BvhBuilder {
public:
    BvhBuilder(const std::vector<Box> &boxes, const std::vector<V3> &centers, uint32_t max_leaf, float traversal_cost)
        : max_leaf_(max_leaf), traversal_cost_(traversal_cost), nodes_(boxes.size() * 2 + 1) { // at most 2n-1 nodes over n leaves
        const size_t n = boxes.size();
        refs_.resize(n);
        (void)centers; // = (lo + hi) * 0.5f, recomputed with the same rounding where needed (Ref::Centre)
        ParallelFor(n, [&](size_t i) {
            Ref r;
            const uint32_t id = static_cast<uint32_t>(i);
            r.lo = F4{{boxes[i].lo.x, boxes[i].lo.y, boxes[i].lo.z, 0.0f}}, r.hi = F4{{boxes[i].hi.x, boxes[i].hi.y, boxes[i].hi.z, 0.0f}};
            memcpy(&r.lo.v[3], &id, 4);
            refs_[i] = r;
        });
    }

    int32_t Build() {
        const uint32_t n = static_cast<uint32_t>(refs_.size());
        const unsigned threads = BuildThreads();
        // bounds of the root
        std::vector<Box> part_box(threads), part_cbox(threads);
        ParallelRun(threads, [&](unsigned t, unsigned nt) {
            const size_t b = static_cast<size_t>(n) * t / nt, e = static_cast<size_t>(n) * (t + 1) / nt;
            for (size_t i = b; i < e; ++i) {
                const Ref &rf = refs_[i];
                part_box[t].Grow(Box{{rf.lo.v[0], rf.lo.v[1], rf.lo.v[2]}, {rf.hi.v[0], rf.hi.v[1], rf.hi.v[2]}});
                part_cbox[t].Grow(V3{rf.Centre(0), rf.Centre(1), rf.Centre(2)});
            }
        });
        Task root{0, n, 0, Box(), Box(), nodes_.Allocate()};
        for (unsigned t = 0; t < threads; ++t) root.box.Grow(part_box[t]), root.cbox.Grow(part_cbox[t]);
        const int32_t root_id = root.id;
        PhaseTimer phase;
        phase("  root bounds");
        // large nodes: one at a time, all threads on each
        std::vector<Task> large{root}, small;
        scratch_.resize(threads > 1 ? n : 0);
        while (!large.empty()) {
            const Task task = large.back();
            large.pop_back();
            if (threads == 1 || task.end - task.begin < kParallelNode) {
                small.push_back(task);
                continue;
            }
            Task l, r;
            if (SplitNode(task, threads, &l, &r)) large.push_back(l), large.push_back(r);
        }
        phase("  large nodes");
        // the subtrees below: biggest first, one per thread
        std::sort(small.begin(), small.end(), [](const Task &a, const Task &b) { return a.end - a.begin > b.end - b.begin; });
        std::atomic<size_t> next{0};
        ParallelRun(std::min<size_t>(threads, std::max<size_t>(small.size(), 1)), [&](unsigned, unsigned) {
            std::unique_ptr<Bins> bins(new Bins());
            for (size_t k = next.fetch_add(1); k < small.size(); k = next.fetch_add(1)) Recurse(small[k], *bins);
        });
        phase("  subtrees");
        scratch_.clear();
        scratch_.shrink_to_fit();
        order_.resize(n);
        ParallelFor(n, [&](size_t i) { order_[i] = refs_[i].id(); });
        return root_id;
    }
    const BuildNodePool &nodes() const { return nodes_; }
    const std::vector<uint32_t> &order() const { return order_; }

private:
    static constexpr int kBins = 32;
    static constexpr int kMedianSplitDepth = 32; // + ceil(log2(2^28 triangles)) = 60 levels at most < kStackSize (64)
    static constexpr uint32_t kParallelNode = 1u << 16;

    struct Ref {   // 32 bytes: box corners padded to four floats, the triangle's index in the pad of `lo`
        F4 lo, hi;
        uint32_t id() const {
            uint32_t i;
            memcpy(&i, &lo.v[3], 4);
            return i;
        }
        float Centre(int axis) const { return (lo.v[axis] + hi.v[axis]) * 0.5f; } // = centers[i] of the caller, same rounding
    };
    struct Task {
        uint32_t begin, end;
        int depth;
        Box box, cbox; // bounds of the boxes / of the centres of [begin, end)
        int32_t id;    // node slot
    };
    struct Bin {
        F4 lo = kF4Max, hi = kF4Lowest;   // box of the member boxes
        F4 clo = kF4Max, chi = kF4Lowest; // box of the member centres
        Box box() const { return Box{{lo.v[0], lo.v[1], lo.v[2]}, {hi.v[0], hi.v[1], hi.v[2]}}; }
        Box cbox() const { return Box{{clo.v[0], clo.v[1], clo.v[2]}, {chi.v[0], chi.v[1], chi.v[2]}}; }
    };
    struct Bins {
        Bin b[3][kBins];
        uint32_t count[3][kBins] = {};
        uint32_t mask[3] = {0, 0, 0}; // non-empty bins per axis: small nodes touch a few of the 96 bins, and only those are swept / reset
        void Merge(const Bins &o) {
            for (int a = 0; a < 3; ++a) {
                for (int k = 0; k < kBins; ++k) {
                    b[a][k].lo = Min4(b[a][k].lo, o.b[a][k].lo), b[a][k].hi = Max4(b[a][k].hi, o.b[a][k].hi);
                    b[a][k].clo = Min4(b[a][k].clo, o.b[a][k].clo), b[a][k].chi = Max4(b[a][k].chi, o.b[a][k].chi);
                    count[a][k] += o.count[a][k];
                }
                mask[a] |= o.mask[a];
            }
        }
        void Reset() {
            for (int a = 0; a < 3; ++a) {
                for (uint32_t m = mask[a]; m != 0; m &= m - 1) {
                    const int k = __builtin_ctz(m);
                    b[a][k] = Bin();
                    count[a][k] = 0;
                }
                mask[a] = 0;
            }
        }
    };
    struct Split {
        int axis = -1, bin = -1;
        float cost = kFltMax;
    };

    static int BinOf(float c, float lo, float scale) {
        const int b = static_cast<int>((c - lo) * scale);
        return std::min(std::max(b, 0), kBins - 1);
    }

    void BinRange(const Task &t, uint32_t begin, uint32_t end, Bins *bins) const {
        const V3 ext = t.cbox.hi - t.cbox.lo;
        float lo[3], scale[3];
        bool use[3];
        for (int a = 0; a < 3; ++a) {
            const float e = (&ext.x)[a];
            use[a] = e > 0.0f;
            lo[a] = (&t.cbox.lo.x)[a], scale[a] = use[a] ? kBins / e : 0.0f;
        }
        for (uint32_t i = begin; i < end; ++i) {
            const Ref &r = refs_[i];
            F4 c; // lane 3 of `lo` holds the index bits (a denormal as a float: arithmetic on it would trap to microcode)
            for (int k = 0; k < 3; ++k) c.v[k] = (r.lo.v[k] + r.hi.v[k]) * 0.5f;
            c.v[3] = 0.0f;
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                const int k = BinOf(c.v[a], lo[a], scale[a]);
                Bin &bin = bins->b[a][k];
                bin.lo = Min4(bin.lo, r.lo), bin.hi = Max4(bin.hi, r.hi);
                bin.clo = Min4(bin.clo, c), bin.chi = Max4(bin.chi, c);
                ++bins->count[a][k];
                bins->mask[a] |= 1u << k;
            }
        }
    }

    // Cheapest of the 3 x 31 bin boundaries.  Only boundaries right after a NON-EMPTY bin are evaluated: the boundaries inside a
    // run of empty bins separate the same two sets at the same cost, and the first of them (the one evaluated here) is the
    // one a sweep over all 31 would keep.
    Split BestSplit(const Task &t, const Bins &bins) const {
        const V3 ext = t.cbox.hi - t.cbox.lo;
        Split best;
        for (int axis = 0; axis < 3; ++axis) {
            if (!((&ext.x)[axis] > 0.0f)) continue;
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            auto half_area = [](const F4 &lo, const F4 &hi) { // Box::HalfArea, same expression
                const float dx = hi.v[0] - lo.v[0], dy = hi.v[1] - lo.v[1], dz = hi.v[2] - lo.v[2];
                return dx * dy + dy * dz + dz * dx;
            };
            F4 lo = kF4Max, hi = kF4Lowest;
            uint32_t cnt = 0;
            for (uint32_t m = bins.mask[axis]; m != 0;) { // suffixes, from the last non-empty bin down
                const int b = 31 - __builtin_clz(m);
                m &= ~(1u << b);
                lo = Min4(lo, bins.b[axis][b].lo), hi = Max4(hi, bins.b[axis][b].hi);
                cnt += bins.count[axis][b];
                right_area[b] = half_area(lo, hi);
                right_cnt[b] = cnt;
            }
            lo = kF4Max, hi = kF4Lowest;
            cnt = 0;
            for (uint32_t m = bins.mask[axis]; m != 0;) {
                const int b = __builtin_ctz(m);
                m &= m - 1;
                if (m == 0) break; // the last non-empty bin: nothing to its right
                const int next = __builtin_ctz(m);
                lo = Min4(lo, bins.b[axis][b].lo), hi = Max4(hi, bins.b[axis][b].hi);
                cnt += bins.count[axis][b];
                const float cost = half_area(lo, hi) * cnt + right_area[next] * right_cnt[next];
                if (cost < best.cost) best.cost = cost, best.axis = axis, best.bin = b;
            }
        }
        return best;
    }

    // Leaf or split?  Returns true with the split to use (axis < 0: halve the list), false for a leaf.
    bool Decide(const Task &t, const Bins &bins, Split *split) {
        const uint32_t n = t.end - t.begin;
        nodes_[t.id].box = t.box;
        *split = n > 1 ? BestSplit(t, bins) : Split();
        const float leaf_cost = t.box.HalfArea() * n;
        // SAH: splitting costs one traversal step (traversal_cost_ triangle tests) on top of the children's expected tests
        if (n == 1 || (n <= max_leaf_ && (split->axis < 0 || split->cost + traversal_cost_ * t.box.HalfArea() >= leaf_cost))) {
            nodes_[t.id].first = t.begin;
            nodes_[t.id].count = n;
            return false;
        }
        // all centroids coincide, or the tree is getting deep (peeling off a sliver per level): split the list in halves,
        // which bounds the depth by kMedianSplitDepth + log2(n) and keeps the traversal stacks (traverse.cuh) safe
        if (t.depth >= kMedianSplitDepth) split->axis = -1;
        return true;
    }

    void ChildTasks(const Task &t, const Split &split, const Bins &bins, uint32_t mid, Task *l, Task *r) {
        *l = Task{t.begin, mid, t.depth + 1, Box(), Box(), nodes_.Allocate()};
        *r = Task{mid, t.end, t.depth + 1, Box(), Box(), nodes_.Allocate()};
        if (split.axis >= 0) {
            for (uint32_t m = bins.mask[split.axis]; m != 0; m &= m - 1) {
                const int b = __builtin_ctz(m);
                Task *side = b <= split.bin ? l : r;
                side->box.Grow(bins.b[split.axis][b].box()), side->cbox.Grow(bins.b[split.axis][b].cbox());
            }
        } else {
            for (uint32_t i = t.begin; i < t.end; ++i) {
                Task *side = i < mid ? l : r;
                const Ref &rf = refs_[i];
                side->box.Grow(Box{{rf.lo.v[0], rf.lo.v[1], rf.lo.v[2]}, {rf.hi.v[0], rf.hi.v[1], rf.hi.v[2]}});
                side->cbox.Grow(V3{rf.Centre(0), rf.Centre(1), rf.Centre(2)});
            }
        }
        nodes_[t.id].left = l->id;
        nodes_[t.id].right = r->id;
    }

    // One large node with all threads: false when it became a leaf.
    bool SplitNode(const Task &t, unsigned threads, Task *l, Task *r) {
        const uint32_t n = t.end - t.begin;
        std::vector<Bins> part(threads);
        ParallelRun(threads, [&](unsigned k, unsigned nt) {
            BinRange(t, t.begin + static_cast<uint32_t>(static_cast<uint64_t>(n) * k / nt), t.begin + static_cast<uint32_t>(static_cast<uint64_t>(n) * (k + 1) / nt), &part[k]);
        });
        for (unsigned k = 1; k < threads; ++k) part[0].Merge(part[k]);
        Split split;
        if (!Decide(t, part[0], &split)) return false;
        uint32_t mid = t.begin + n / 2;
        if (split.axis >= 0) {
            // stable parallel partition through the scratch array: count per chunk, then scatter
            const float lo = (&t.cbox.lo.x)[split.axis], sc = kBins / ((&t.cbox.hi.x)[split.axis] - lo);
            std::vector<uint32_t> left_count(threads + 1, 0);
            auto chunk = [&](unsigned k, unsigned nt, uint32_t *b, uint32_t *en) {
                *b = t.begin + static_cast<uint32_t>(static_cast<uint64_t>(n) * k / nt);
                *en = t.begin + static_cast<uint32_t>(static_cast<uint64_t>(n) * (k + 1) / nt);
            };
            ParallelRun(threads, [&](unsigned k, unsigned nt) {
                uint32_t b, en, c = 0;
                chunk(k, nt, &b, &en);
                for (uint32_t i = b; i < en; ++i) c += BinOf(refs_[i].Centre(split.axis), lo, sc) <= split.bin;
                left_count[k + 1] = c;
            });
            for (unsigned k = 0; k < threads; ++k) left_count[k + 1] += left_count[k];
            mid = t.begin + left_count[threads];
            ParallelRun(threads, [&](unsigned k, unsigned nt) {
                uint32_t b, en;
                chunk(k, nt, &b, &en);
                uint32_t lpos = t.begin + left_count[k], rpos = mid + (b - t.begin) - left_count[k];
                for (uint32_t i = b; i < en; ++i) {
                    if (BinOf(refs_[i].Centre(split.axis), lo, sc) <= split.bin) scratch_[lpos++] = refs_[i];
                    else scratch_[rpos++] = refs_[i];
                }
            });
            ParallelRun(threads, [&](unsigned k, unsigned nt) {
                uint32_t b, en;
                chunk(k, nt, &b, &en);
                std::copy(scratch_.begin() + b, scratch_.begin() + en, refs_.begin() + b);
            });
        }
        ChildTasks(t, split, part[0], mid, l, r);
        return true;
    }

    // `bins`: this thread's scratch, all empty on entry and on return.
    void Recurse(const Task &t, Bins &bins) {
        if (t.end - t.begin > 1) BinRange(t, t.begin, t.end, &bins);
        Split split;
        if (!Decide(t, bins, &split)) {
            bins.Reset();
            return;
        }
        const uint32_t n = t.end - t.begin;
        uint32_t mid = t.begin + n / 2;
        if (split.axis >= 0) {
            const float lo = (&t.cbox.lo.x)[split.axis], sc = kBins / ((&t.cbox.hi.x)[split.axis] - lo);
            auto it = std::partition(refs_.begin() + t.begin, refs_.begin() + t.end,
                                     [&](const Ref &r) { return BinOf(r.Centre(split.axis), lo, sc) <= split.bin; });
            mid = static_cast<uint32_t>(it - refs_.begin());
        }
        Task l, r;
        ChildTasks(t, split, bins, mid, &l, &r);
        bins.Reset();
        Recurse(l, bins);
        Recurse(r, bins);
    }

    uint32_t max_leaf_;
    float traversal_cost_;
    std::vector<Ref> refs_, scratch_;
    std::vector<uint32_t> order_;
    BuildNodePool nodes_;
};

int32_t EncodeLeaf(uint32_t first, uint32_t count) { return ~static_cast<int32_t>((first << 3) | (count - 1)); }

void SetChildBox(BvhNode *node, int which, const Box &b) {
    if (which == 0) {
        node->c0xy = {b.lo.x, b.hi.x, b.lo.y, b.hi.y};
        node->cz.x = b.lo.z, node->cz.y = b.hi.z;
    } else {
        node->c1xy = {b.lo.x, b.hi.x, b.lo.y, b.hi.y};
        node->cz.z = b.lo.z, node->cz.w = b.hi.z;
    }
}

// Leaves with more than 8 triangles cannot be encoded; the builder never makes them for max_leaf <= 8.
// Returns the number of inner-node levels (what FlatBvhDepth would count on the result).
uint32_t FlattenBvh(const BuildNodePool &bn, int32_t root, uint32_t top_nodes, std::vector<BvhNode> *out) {
    out->clear();
    if (root < 0) return 0;
    const Box empty; // inverted box: never hit
    if (bn[root].left < 0) { // whole scene is one leaf
        BvhNode n{};
        SetChildBox(&n, 0, bn[root].box);
        SetChildBox(&n, 1, empty);
        n.child0 = EncodeLeaf(bn[root].first, bn[root].count);
        n.child1 = EncodeLeaf(0, 1);
        out->push_back(n);
        return 1;
    }
    // Output order: breadth-first for the first `top_nodes` inner nodes (the part staged in
    // shared memory), depth-first below so that subtrees stay contiguous in HBM/L2.
    std::vector<int32_t> out_index(bn.size(), -1);
    std::vector<int32_t> order;
    order.reserve(bn.size());
    std::vector<uint32_t> level(bn.size(), 0); // of inner nodes, root = 1
    uint32_t deepest = 1;
    level[root] = 1;
    auto visit = [&](int32_t id, int32_t child, std::vector<int32_t> *to) {
        if (bn[child].left < 0) return;
        level[child] = level[id] + 1;
        deepest = std::max(deepest, level[child]);
        to->push_back(child);
    };
    std::vector<int32_t> frontier{root};
    size_t head = 0;
    while (head < frontier.size() && order.size() < top_nodes) {
        const int32_t id = frontier[head++];
        out_index[id] = static_cast<int32_t>(order.size());
        order.push_back(id);
        visit(id, bn[id].left, &frontier);
        visit(id, bn[id].right, &frontier);
    }
    std::vector<int32_t> stack;
    for (size_t i = frontier.size(); i-- > head;) stack.push_back(frontier[i]);
    while (!stack.empty()) {
        const int32_t id = stack.back();
        stack.pop_back();
        out_index[id] = static_cast<int32_t>(order.size());
        order.push_back(id);
        visit(id, bn[id].right, &stack);
        visit(id, bn[id].left, &stack);
    }
    out->resize(order.size());
    ParallelFor(order.size(), [&](size_t i) {
        const BuildNode &src = bn[order[i]];
        BvhNode n{};
        const BuildNode &l = bn[src.left], &r = bn[src.right];
        SetChildBox(&n, 0, l.box);
        SetChildBox(&n, 1, r.box);
        n.child0 = l.left >= 0 ? out_index[src.left] : EncodeLeaf(l.first, l.count);
        n.child1 = r.left >= 0 ? out_index[src.right] : EncodeLeaf(r.first, r.count);
        (*out)[i] = n;
    });
    return deepest;
}

// Number of inner-node levels of a flattened tree = the most entries a traversal can have on its stack, plus one.
uint32_t FlatBvhDepth(const std::vector<BvhNode> &nodes) {
    if (nodes.empty()) return 0;
    uint32_t deepest = 0;
    std::vector<std::pair<int32_t, uint32_t>> stack{{0, 1u}};
    while (!stack.empty()) {
        const auto [id, depth] = stack.back();
        stack.pop_back();
        deepest = std::max(deepest, depth);
        if (nodes[id].child0 >= 0) stack.push_back({nodes[id].child0, depth + 1});
        if (nodes[id].child1 >= 0) stack.push_back({nodes[id].child1, depth + 1});
    }
    return deepest;
}

// ---------------------------------------------------------------------------------------------
// Textures on the host (needed for the env-map tables)
// ---------------------------------------------------------------------------------------------
// bitmap.cpp:6-56, identity to_uv assumed by the caller where noted.
V3 GetColorBitmapHost(const DTexture &t, const float *pixels, float u, float v) {
    const float *m = t.to_uv.m;
    const float ux = m[0] * u + m[1] * v + m[2] * 0.0f + m[3], uy = m[4] * u + m[5] * v + m[6] * 0.0f + m[7];
    float x = ux * t.width, y = uy * t.height;
    while (x < 0) x += t.width;
    while (x > t.width - 1) x -= t.width;
    while (y < 0) y += t.height;
    while (y > t.height - 1) y -= t.height;
    const uint32_t x0 = static_cast<uint32_t>(x), y0 = static_cast<uint32_t>(y);
    const float tx = x - x0, ty = y - y0;
    const uint32_t x1 = tx > 0.0f ? x0 + 1 : x0, y1 = ty > 0.0f ? y0 + 1 : y0;
    const float *data = pixels + t.pixel_offset;
    auto lerp = [](float a, float b, float w) { return (1.0f - w) * a + w * b; };
    if (t.channels == 1) {
        const float c00 = data[x0 + t.width * y0], c01 = data[x0 + t.width * y1], c10 = data[x1 + t.width * y0],
                    c11 = data[x1 + t.width * y1];
        const float c = lerp(lerp(c00, c01, ty), lerp(c10, c11, ty), tx);
        return {c, c, c};
    }
    auto px = [&](uint32_t xx, uint32_t yy) {
        const uint32_t o = (xx + t.width * yy) * t.channels;
        return V3{data[o], data[o + 1], data[o + 2]};
    };
    const V3 c00 = px(x0, y0), c01 = px(x0, y1), c10 = px(x1, y0), c11 = px(x1, y1);
    const V3 c0 = (1.0f - ty) * c00 + ty * c01, c1 = (1.0f - ty) * c10 + ty * c11;
    return (1.0f - tx) * c0 + tx * c1;
}

// envmap.cpp:20-68 + packing of renderer.cpp:584-605.
bool BuildEnvMapTables(const DTexture &tex, const float *pixels, std::vector<float> *packed, float *normalization,
                       std::string *error) {
    const int width = tex.width, height = tex.height;
    const float width_inv = 1.0f / width, height_inv = 1.0f / height;
    std::vector<float> cdf_rows(height + 1), weight_rows(height), cdf_cols(static_cast<size_t>(width + 1) * height);
    float sum_row = 0.0f;
    cdf_rows[0] = 0;
    for (int y = 0; y < height; ++y) {
        float sum_col = 0.0f;
        cdf_cols[0] = 0;
        for (int x = 0; x < width; ++x) {
            const V3 rgb = GetColorBitmapHost(tex, pixels, x * width_inv, y * height_inv);
            sum_col += 0.2126f * rgb.x + 0.7152f * rgb.y + 0.0722f * rgb.z;
            cdf_cols[static_cast<size_t>(y) * (width + 1) + (x + 1)] = sum_col;
        }
        cdf_cols[static_cast<size_t>(y) * (width + 1) + width] = 1.0f;
        const float normalization_col = 1.0f / sum_col;
        for (int x = 1; x < width; ++x) cdf_cols[static_cast<size_t>(y) * (width + 1) + width - x] *= normalization_col;
        const float weight = sinf((y + 0.5f) * kPi / height);
        weight_rows[y] = weight;
        sum_row += sum_col * weight;
        cdf_rows[y + 1] = sum_row;
    }
    cdf_rows[height] = 1.0f;
    const float normalization_row = 1.0f / sum_row;
    for (int y = 1; y < height; ++y) cdf_rows[height - y] *= normalization_row;
    if (!std::isfinite(sum_row)) {
        *error = "The environment map contains an invalid floating point value (nan/inf).";
        return false;
    }
    *normalization = static_cast<float>(1.0 / (sum_row * (k2Pi * width_inv) * (kPi * height_inv)));
    packed->clear();
    packed->insert(packed->end(), cdf_rows.begin(), cdf_rows.end());
    packed->insert(packed->end(), weight_rows.begin(), weight_rows.end());
    packed->insert(packed->end(), cdf_cols.begin(), cdf_cols.end());
    return true;
}

// bsdf.cpp:12-56
float AverageFresnelDielectric(float eta) {
    if (eta < 1.0) {
        return -1.4399f * (eta * eta) + 0.7099f * eta + 0.6681f + 0.0636f / eta;
    }
    const float inv_eta = 1.0f / eta, inv_eta_2 = inv_eta * inv_eta, inv_eta_3 = inv_eta_2 * inv_eta,
                inv_eta_4 = inv_eta_3 * inv_eta, inv_eta_5 = inv_eta_4 * inv_eta;
    return 0.919317f - 3.4793f * inv_eta + 6.75335f * inv_eta_2 - 7.80989f * inv_eta_3 + 4.98554f * inv_eta_4 -
           1.36881f * inv_eta_5;
}
float AverageFresnelConductor(float r, float e) {
    return 0.087237f + 0.0230685f * e - 0.0864902f * e * e + 0.0774594f * e * e * e + 0.782654f * r -
           0.136432f * r * r + 0.278708f * r * r * r + 0.19744f * e * r + 0.0360605f * e * e * r - 0.2586f * e * r * r;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Kulla-Conty tables: kulla_conty.cpp:13-80 (+ microfacet.cpp:8-19, 63-75; ray.cpp:49-52)
// ---------------------------------------------------------------------------------------------
namespace {

float VanDerCorput2(uint32_t index) { // math.hpp:29-41
    const float base_inv = 1.0f / 2;
    float result = 0.0f, frac = base_inv;
    while (index > 0) {
        result += frac * (index % 2);
        index = static_cast<uint32_t>(index * base_inv);
        frac *= base_inv;
    }
    return result;
}

void SampleGgxIso(float xi_0, float xi_1, float roughness, V3 *h) { // microfacet.cpp:8-19 (direction only)
    const float alpha_2 = roughness * roughness;
    const float tan_theta_2 = alpha_2 * xi_0 / (1.0f - xi_0), phi = k2Pi * xi_1;
    const float cos_theta = 1.0f / sqrtf(1.0f + tan_theta_2), sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    *h = {sin_theta * cosf(phi), sin_theta * sinf(phi), cos_theta};
}

float SmithG1GgxIso(float roughness, V3 v, V3 h) { // microfacet.cpp:63-75
    const float n_dot_v = v.z;
    if (n_dot_v * h.z <= 0) return 0;
    const float cos_theta_2 = n_dot_v * n_dot_v, tan_theta_2 = (1.0f - cos_theta_2) / cos_theta_2,
                alpha_2 = roughness * roughness;
    return 2.0f / (1.0f + sqrtf(static_cast<float>(1.0 + static_cast<double>(alpha_2 * tan_theta_2))));
}

V3 Reflect(V3 wi, V3 n) { return Normalize(wi - (2.0f * Dot(wi, n)) * n); } // ray.cpp:49-52

float IntegrateBrdf(V3 V, float roughness) {
    constexpr uint32_t sample_count = 1024;
    constexpr float step = 1.0f / sample_count;
    const V3 N = {0, 0, 1};
    float accum = 0.0f;
    for (uint32_t i = 0; i < sample_count; ++i) {
        V3 H;
        SampleGgxIso(i * step, VanDerCorput2(i), roughness, &H);
        const V3 L = Reflect(V, H);
        const V3 mV = {-V.x, -V.y, -V.z};
        const float G = SmithG1GgxIso(roughness, mV, H) * SmithG1GgxIso(roughness, L, H), n_dot_v = Dot(N, mV),
                    n_dot_l = Dot(N, L), n_dot_h = Dot(N, H), h_dot_v = Dot(H, mV);
        if (n_dot_l > 0.0f && n_dot_h > 0.0f && h_dot_v > 0.0f) accum += (h_dot_v * G) / (n_dot_v * n_dot_h);
    }
    return fminf(accum * step, 1.0f);
}

float IntegrateAlbedo(V3 V, float roughness, float brdf) {
    constexpr uint32_t sample_count = 1024;
    constexpr float step = 1.0f / sample_count;
    const V3 N = {0, 0, 1};
    float accum = 0.0f;
    for (uint32_t i = 0; i < sample_count; ++i) {
        V3 H;
        SampleGgxIso(i * step, VanDerCorput2(i), roughness, &H);
        const V3 L = Reflect(V, H);
        const V3 mV = {-V.x, -V.y, -V.z};
        const float n_dot_l = Dot(N, L), n_dot_h = Dot(N, H), h_dot_v = Dot(mV, H);
        if (n_dot_l > 0.0f && n_dot_h > 0.0f && h_dot_v > 0.0f) accum += brdf * n_dot_l;
    }
    return accum * 2.0f * step;
}

} // namespace

// The tables depend on nothing but the constants above: computed once per process (0.1-0.4 s), copied afterwards.
void ComputeKullaContyTablesOnce(float *brdf_avg, float *albedo_avg);
void ComputeKullaContyTables(float *brdf_avg, float *albedo_avg) {
    static std::once_flag once;
    static std::vector<float> brdf, albedo;
    std::call_once(once, [] {
        brdf.assign(kLutResolution * kLutResolution, 0.0f), albedo.assign(kLutResolution, 0.0f);
        ComputeKullaContyTablesOnce(brdf.data(), albedo.data());
    });
    std::copy(brdf.begin(), brdf.end(), brdf_avg);
    std::copy(albedo.begin(), albedo.end(), albedo_avg);
}

void ComputeKullaContyTablesOnce(float *brdf_avg, float *albedo_avg) {
    const float step = 1.0f / kLutResolution;
    auto row = [&](int i) {
        float albedo_accum = 0.0f;
        const float roughness = step * (static_cast<float>(i) + 0.5f);
        for (int j = kLutResolution - 1; j >= 0; --j) {
            const float n_dot_v = step * (static_cast<float>(j) + 0.5f);
            const V3 V = {-sqrtf(1.f - n_dot_v * n_dot_v), 0.0f, -n_dot_v};
            const float b = IntegrateBrdf(V, roughness);
            brdf_avg[i * kLutResolution + j] = b;
            albedo_accum += IntegrateAlbedo(V, roughness, b);
        }
        albedo_avg[i] = albedo_accum * step;
    };
    ParallelRun(BuildThreads(), [&](unsigned t, unsigned num_threads) {
        for (int i = static_cast<int>(t); i < kLutResolution; i += static_cast<int>(num_threads)) row(i);
    });
}

DCamera MakeCamera(const b200pt_camera &cam, uint32_t width, uint32_t height) {
    const V3 eye = Load3(cam.eye), look_at = Load3(cam.look_at), up_in = Load3(cam.up);
    const float fov_y = cam.fov_x * static_cast<int>(height) / static_cast<int>(width);
    const V3 front = Normalize(look_at - eye);
    const V3 right = Normalize(Cross(front, up_in));
    const V3 up = Normalize(Cross(right, front));
    const float to_rad = 0.01745329251994329576923690768489f;
    DCamera c;
    c.eye = ToF3(eye);
    c.front = ToF3(front);
    c.view_dx = ToF3(right * tanf((0.5f * cam.fov_x) * to_rad));
    c.view_dy = ToF3(up * tanf((0.5f * fov_y) * to_rad));
    return c;
}

bool BuildHostScene(const b200pt_scene_desc &d, uint32_t max_leaf_size, bool gpu_lbvh, bool bvh8, HostScene *hs, std::string *error) {
    const bool wide = bvh8 && !gpu_lbvh; // the GPU builder emits the binary layout directly
    if (d.abi_version != B200PT_ABI_VERSION) {
        *error = "b200pt_scene_desc.abi_version mismatch";
        return false;
    }
    if (max_leaf_size == 0) max_leaf_size = 4;
    max_leaf_size = std::min(max_leaf_size, 8u);
    hs->camera = d.camera;

    PhaseTimer whole;
    // ---- textures ----
    for (uint64_t i = 0; i < d.num_textures; ++i) {
        const b200pt_texture &t = d.textures[i];
        DTexture o{};
        o.type = t.type;
        o.width = t.width, o.height = t.height, o.channels = t.channels;
        o.color0 = {t.color0[0], t.color0[1], t.color0[2]};
        o.color1 = {t.color1[0], t.color1[1], t.color1[2]};
        memcpy(o.to_uv.m, t.to_uv, sizeof(o.to_uv.m));
        o.pixel_offset = t.pixel_offset;
        if (t.type == B200PT_TEX_BITMAP) {
            const uint64_t n = static_cast<uint64_t>(t.width) * t.height * t.channels;
            if (t.width <= 0 || t.height <= 0 || t.channels <= 0 || t.pixel_offset + n > d.num_pixels) {
                *error = "bitmap texture " + std::to_string(i) + " exceeds the pixel pool.";
                return false;
            }
        } else if (t.type != B200PT_TEX_CONSTANT && t.type != B200PT_TEX_CHECKERBOARD) {
            *error = "unknow texture type."; // renderer.cpp:421
            return false;
        }
        hs->textures.push_back(o);
    }
    auto check_texture = [&](uint32_t id, bool allow_invalid) {
        if (id == kInvalid && allow_invalid) return true;
        if (id >= d.num_textures) {
            *error = "cannot find texture (id " + std::to_string(id) + ")."; // renderer.cpp:441-446
            return false;
        }
        return true;
    };

    // ---- BSDFs (bsdf.cpp:112-186) ----
    bool need_kulla_conty = false;
    for (uint64_t i = 0; i < d.num_bsdfs; ++i) {
        const b200pt_bsdf &b = d.bsdfs[i];
        DBsdf o{};
        o.type = b.type;
        o.twosided = b.twosided;
        o.id_opacity = b.id_opacity, o.id_bump_map = b.id_bump_map;
        o.id_radiance = b.id_radiance, o.id_diffuse_reflectance = b.id_diffuse_reflectance;
        o.id_roughness_u = b.id_roughness_u, o.id_roughness_v = b.id_roughness_v;
        o.id_specular_reflectance = b.id_specular_reflectance;
        o.id_specular_transmittance = b.id_specular_transmittance;
        // The reference never copies RoughDiffuseInfo::use_fast_approx into its BsdfData (bsdf.cpp:136-141):
        // the field keeps the zero the union was initialised with, i.e. the full Oren-Nayar model is always used.
        o.use_fast_approx = 0;
        o.eta = b.eta, o.eta_inv = 1.0f / b.eta;
        o.F_avg_s = 1.0f, o.F_avg_inv_s = 1.0f, o.reflectivity_s = 1.0f;
        bool ok = check_texture(b.id_opacity, true) && check_texture(b.id_bump_map, true);
        switch (b.type) {
        case B200PT_BSDF_AREA_LIGHT:
            ok = ok && check_texture(b.id_radiance, false);
            break;
        case B200PT_BSDF_DIFFUSE:
            ok = ok && check_texture(b.id_diffuse_reflectance, false);
            break;
        case B200PT_BSDF_ROUGH_DIFFUSE:
            ok = ok && check_texture(b.id_diffuse_reflectance, false) && check_texture(b.id_roughness_u, false);
            break;
        case B200PT_BSDF_CONDUCTOR:
            ok = ok && check_texture(b.id_roughness_u, false) && check_texture(b.id_roughness_v, false) &&
                 check_texture(b.id_specular_reflectance, false);
            o.reflectivity = {b.reflectivity[0], b.reflectivity[1], b.reflectivity[2]};
            o.edgetint = {b.edgetint[0], b.edgetint[1], b.edgetint[2]};
            o.F_avg = {AverageFresnelConductor(b.reflectivity[0], b.edgetint[0]),
                       AverageFresnelConductor(b.reflectivity[1], b.edgetint[1]),
                       AverageFresnelConductor(b.reflectivity[2], b.edgetint[2])};
            need_kulla_conty = true;
            break;
        case B200PT_BSDF_DIELECTRIC:
            o.F_avg_s = AverageFresnelDielectric(b.eta);
            o.F_avg_inv_s = AverageFresnelDielectric(1.0f / b.eta);
            need_kulla_conty = true;
            [[fallthrough]]; // bsdf.cpp:157-181
        case B200PT_BSDF_THIN_DIELECTRIC:
            ok = ok && check_texture(b.id_roughness_u, false) && check_texture(b.id_roughness_v, false) &&
                 check_texture(b.id_specular_reflectance, false) && check_texture(b.id_specular_transmittance, false);
            o.twosided = 1;
            o.reflectivity_s = ((b.eta - 1.0f) * (b.eta - 1.0f)) / ((b.eta + 1.0f) * (b.eta + 1.0f));
            break;
        case B200PT_BSDF_PLASTIC:
            ok = ok && check_texture(b.id_roughness_u, false) && check_texture(b.id_diffuse_reflectance, false) &&
                 check_texture(b.id_specular_reflectance, false);
            o.reflectivity_s = ((b.eta - 1.0f) * (b.eta - 1.0f)) / ((b.eta + 1.0f) * (b.eta + 1.0f));
            o.F_avg_s = AverageFresnelDielectric(b.eta);
            break;
        default:
            *error = "unknow BSDF type."; // renderer.cpp:484
            return false;
        }
        if (!ok) return false;
        hs->bsdfs.push_back(o);
    }

    // ---- media (medium.cpp:6-39) ----
    for (uint64_t i = 0; i < d.num_media; ++i) {
        const b200pt_medium &m = d.media[i];
        DMedium o{};
        o.sigma_s = {m.sigma_s[0], m.sigma_s[1], m.sigma_s[2]};
        o.sigma_t = {m.sigma_a[0] + m.sigma_s[0], m.sigma_a[1] + m.sigma_s[1], m.sigma_a[2] + m.sigma_s[2]};
        o.sampling_weight = 0.0f;
        for (int dim = 0; dim < 3; ++dim) {
            const float sigma_t = m.sigma_a[dim] + m.sigma_s[dim];
            const float albedo = m.sigma_s[dim] * (1.0f / sigma_t);
            if (albedo > o.sampling_weight && sigma_t > 0) o.sampling_weight = albedo;
        }
        if (o.sampling_weight > 0 && o.sampling_weight < 0.5f) o.sampling_weight = 0.5f;
        o.phase_type = m.phase_type;
        o.g = {m.g[0], m.g[1], m.g[2]};
        hs->media.push_back(o);
    }

    // ---- instances + geometry ----
    RawVector<RawTriangle> tris;
    std::vector<Box> boxes; // per triangle: its box and the centre of the box, filled by AppendMesh
    std::vector<V3> centers;
    {
        uint64_t total = 0; // one allocation for all meshes (176 B per triangle: growing by doubling copies it all again and again)
        for (uint64_t i = 0; i < d.num_instances; ++i)
            total += d.instances[i].type == B200PT_INST_MESHES ? d.instances[i].num_triangles : (d.instances[i].type == B200PT_INST_CUBE ? 12u : 2u);
        tris.reserve(total), boxes.reserve(total), centers.reserve(total);
    }
    std::vector<float> inst_area(d.num_instances, 0.0f);
    hs->instances.resize(d.num_instances);
    Box scene_box;
    for (uint64_t i = 0; i < d.num_instances; ++i) {
        const b200pt_instance &in = d.instances[i];
        DInstance &o = hs->instances[i];
        o = DInstance{};
        o.id_bsdf = in.id_bsdf;
        o.id_medium_int = in.id_medium_int, o.id_medium_ext = in.id_medium_ext;
        o.area_light = kInvalid, o.analytic = kInvalid;
        if (in.id_bsdf != kInvalid && in.id_bsdf >= d.num_bsdfs) o.id_bsdf = kInvalid; // renderer.cpp:277
        if ((in.id_medium_int != kInvalid && in.id_medium_int >= d.num_media) ||
            (in.id_medium_ext != kInvalid && in.id_medium_ext >= d.num_media)) {
            *error = "instance " + std::to_string(i) + " refers to a medium that does not exist.";
            return false;
        }
        const M4 to_world = LoadM4(in.to_world);
        MeshView mesh;
        bool is_mesh = false;
        switch (in.type) {
        case B200PT_INST_RECTANGLE:
            mesh.positions = kRectPos, mesh.normals = kRectNrm, mesh.texcoords = kRectUv, mesh.indices = kRectIdx;
            mesh.num_vertices = 4, mesh.num_triangles = 2;
            is_mesh = true;
            break;
        case B200PT_INST_CUBE:
            mesh.positions = kCubePos, mesh.normals = kCubeNrm, mesh.texcoords = kCubeUv, mesh.indices = kCubeIdx;
            mesh.num_vertices = 24, mesh.num_triangles = 12;
            is_mesh = true;
            break;
        case B200PT_INST_MESHES: {
            const uint64_t nv = in.num_vertices;
            auto in_range = [&](uint64_t off, uint64_t pool) { return off == B200PT_NO_OFFSET || off + nv <= pool; };
            if (!in_range(in.position_offset, d.num_positions) || !in_range(in.normal_offset, d.num_normals) ||
                !in_range(in.texcoord_offset, d.num_texcoords) || !in_range(in.tangent_offset, d.num_tangents) ||
                !in_range(in.bitangent_offset, d.num_bitangents) ||
                in.index_offset + in.num_triangles > d.num_triangles) {
                *error = "instance " + std::to_string(i) + ": mesh ranges exceed the attribute pools.";
                return false;
            }
            if (in.position_offset != B200PT_NO_OFFSET) mesh.positions = d.positions + 3 * in.position_offset;
            if (in.normal_offset != B200PT_NO_OFFSET) mesh.normals = d.normals + 3 * in.normal_offset;
            if (in.texcoord_offset != B200PT_NO_OFFSET) mesh.texcoords = d.texcoords + 2 * in.texcoord_offset;
            if (in.tangent_offset != B200PT_NO_OFFSET) mesh.tangents = d.tangents + 3 * in.tangent_offset;
            if (in.bitangent_offset != B200PT_NO_OFFSET) mesh.bitangents = d.bitangents + 3 * in.bitangent_offset;
            mesh.indices = d.indices + 3 * in.index_offset;
            mesh.num_vertices = mesh.positions ? nv : 0;
            mesh.num_triangles = in.num_triangles;
            is_mesh = true;
            break;
        }
        case B200PT_INST_SPHERE: { // scene.cpp:326-372, sphere.cpp:9-15
            AnalyticPrim p{};
            p.type = kSphere, p.inst = static_cast<uint32_t>(i);
            p.radius = in.sphere_radius;
            p.center = {in.sphere_center[0], in.sphere_center[1], in.sphere_center[2]};
            p.to_world = ToAffine(to_world);
            p.to_local = ToAffine(Inverse(to_world));
            p.normal_to_world = ToAffine(Inverse(Transpose(to_world)));
            const V3 c = Load3(in.sphere_center), r{in.sphere_radius, in.sphere_radius, in.sphere_radius};
            Box b;
            b.Grow(TransformPoint(to_world, c + r));
            b.Grow(TransformPoint(to_world, c - r));
            memcpy(p.bmin, &b.lo, 12), memcpy(p.bmax, &b.hi, 12);
            const V3 center_world = TransformPoint(to_world, c),
                     boundary_world = TransformPoint(to_world, c + V3{in.sphere_radius, 0.0f, 0.0f});
            const float radius_world = Length(center_world - boundary_world);
            inst_area[i] = 4.0f * kPi * (radius_world * radius_world);
            o.analytic = static_cast<uint32_t>(hs->analytic.size());
            hs->analytic.push_back(p);
            scene_box.Grow(b);
            break;
        }
        case B200PT_INST_DISK: { // scene.cpp:374-416, disk.cpp:9-15
            AnalyticPrim p{};
            p.type = kDisk, p.inst = static_cast<uint32_t>(i);
            p.to_world = ToAffine(to_world);
            p.to_local = ToAffine(Inverse(to_world));
            p.normal_to_world = ToAffine(Inverse(Transpose(to_world)));
            Box b;
            b.Grow(TransformPoint(to_world, V3{-0.5f, -0.5f, 0}));
            b.Grow(TransformPoint(to_world, V3{0.5f, 0.5f, 0}));
            memcpy(p.bmin, &b.lo, 12), memcpy(p.bmax, &b.hi, 12);
            const float radius_world = Length(TransformPoint(to_world, V3{0, 0, 0}) - TransformPoint(to_world, V3{0.5f, 0, 0}));
            inst_area[i] = kPi * (radius_world * radius_world);
            o.analytic = static_cast<uint32_t>(hs->analytic.size());
            hs->analytic.push_back(p);
            scene_box.Grow(b);
            break;
        }
        case B200PT_INST_CYLINDER: { // scene.cpp:418-472, cylinder.cpp:9-19
            AnalyticPrim p{};
            p.type = kCylinder, p.inst = static_cast<uint32_t>(i);
            const V3 p0 = Load3(in.cylinder_p0), p1 = Load3(in.cylinder_p1);
            M4 m = LocalToWorldMat(Normalize(p1 - p0));
            m = Mul(Translate(p0), m);
            m = Mul(to_world, m);
            p.length = Length(TransformPoint(m, V3{0, 0, Length(p1 - p0)}) - TransformPoint(m, V3{0, 0, 0}));
            p.radius = Length(TransformPoint(m, V3{in.cylinder_radius, 0, 0}) - TransformPoint(m, V3{0, 0, 0}));
            p.to_world = ToAffine(m);
            p.to_local = ToAffine(Inverse(m));
            p.normal_to_world = ToAffine(Inverse(Transpose(m)));
            Box b;
            b.Grow(TransformPoint(m, V3{p.radius, p.radius, 0}));
            b.Grow(TransformPoint(m, V3{-p.radius, -p.radius, 0}));
            b.Grow(TransformPoint(m, V3{p.radius, p.radius, p.length}));
            b.Grow(TransformPoint(m, V3{-p.radius, -p.radius, p.length}));
            memcpy(p.bmin, &b.lo, 12), memcpy(p.bmax, &b.hi, 12);
            inst_area[i] = k2Pi * (p.radius * p.radius);
            o.analytic = static_cast<uint32_t>(hs->analytic.size());
            hs->analytic.push_back(p);
            scene_box.Grow(b);
            break;
        }
        default:
            *error = "unknow instance type."; // scene.cpp:184
            return false;
        }
        if (is_mesh && !AppendMesh(mesh, to_world, static_cast<uint32_t>(i), &tris, &boxes, &centers, &inst_area[i], error)) return false;
        o.pdf_area = 1.0f / inst_area[i];
    }

    whole("textures .. geometry");
    // ---- area lights (renderer.cpp:271-304) ----
    hs->cdf_area_light.assign(1, 0.0f);
    for (uint64_t i = 0; i < d.num_instances; ++i) {
        const uint32_t id_bsdf = hs->instances[i].id_bsdf;
        if (id_bsdf != kInvalid && d.bsdfs[id_bsdf].type == B200PT_BSDF_AREA_LIGHT) {
            hs->instances[i].area_light = static_cast<uint32_t>(hs->map_area_light_instance.size());
            hs->map_area_light_instance.push_back(static_cast<uint32_t>(i));
            hs->cdf_area_light.push_back(d.bsdfs[id_bsdf].area_light_weight + hs->cdf_area_light.back());
        }
    }

    // ---- BVH over all triangles ----
    PhaseTimer phase;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t nt = tris.size();
    if (gpu_lbvh) // (the host builder leaves the bounds of all triangles in its root)
        for (size_t i = 0; i < nt; ++i) scene_box.Grow(boxes[i]);
    phase("triangle boxes");
    std::vector<uint32_t> order;
    std::vector<Bvh2Node> binary; // the SAH tree as handed to the wide collapse / the cull-box cut
    int32_t binary_root = -1;
    if (nt > 0) {
        if (nt >= (1u << 28)) {
            *error = "too many triangles for the leaf encoding.";
            return false;
        }
        bool built = false;
        if (gpu_lbvh && nt > 64) {
            // GPU LBVH (bvh_gpu.cu): Morton sort, radix tree, refit, leaves of <= max_leaf triangles and the traversal
            // layout are all produced on the device; the host only permutes the triangle attributes afterwards.
            std::vector<float> flat(6 * nt);
            for (size_t i = 0; i < nt; ++i) {
                memcpy(&flat[6 * i], &boxes[i].lo, 12);
                memcpy(&flat[6 * i + 3], &boxes[i].hi, 12);
            }
            if (!BuildLbvhGpuFlat(flat.data(), static_cast<uint32_t>(nt), &scene_box.lo.x, &scene_box.hi.x, max_leaf_size, &hs->nodes, &order,
                                  &hs->bvh_gpu_ms, error))
                return false;
            // A radix tree over 63-bit Morton keys + index tie-break can be ~95 levels deep on clustered input; the
            // traversal stack holds kStackSize entries.  Such a tree is rebuilt with the (depth-bounded) SAH builder.
            built = FlatBvhDepth(hs->nodes) < kBvh2StackSize;
            if (!built) hs->nodes.clear();
        }
        if (!built) {
            const char *ct_env = getenv("B200PT_SAH_TRAVERSAL_COST"); // tuning knob, default 1 triangle test per node step
            BvhBuilder builder(boxes, centers, wide ? std::min(max_leaf_size, kWideMaxLeaf) : max_leaf_size,
                               ct_env ? static_cast<float>(atof(ct_env)) : 1.0f);
            phase("builder setup");
            binary_root = builder.Build();
            order = builder.order();
            if (!gpu_lbvh && binary_root >= 0) scene_box.Grow(builder.nodes()[binary_root].box);
            phase("SAH build");
            if (wide) {
                const BuildNodePool &bn = builder.nodes();
                binary.resize(bn.size());
                for (size_t i = 0; i < bn.size(); ++i) {
                    memcpy(binary[i].lo, &bn[i].box.lo, 12), memcpy(binary[i].hi, &bn[i].box.hi, 12);
                    binary[i].left = bn[i].left, binary[i].right = bn[i].right;
                    binary[i].first = bn[i].first, binary[i].count = bn[i].count;
                }
                const char *top_env = getenv("B200PT_WIDE_TOP_TARGET");
                WideBuildInfo info;
                if (!BuildWideBvh(binary, binary_root, top_env ? static_cast<uint32_t>(atoi(top_env)) : 2048u, &order, &hs->wide_nodes, &info, error))
                    return false;
                hs->wide_depth = info.depth, hs->wide_top_nodes = info.top_nodes;
            } else {
                const uint32_t levels = FlattenBvh(builder.nodes(), binary_root, 1024, &hs->nodes);
                phase("flatten");
                if (levels >= kBvh2StackSize) {
                    *error = "BVH too deep for the traversal stack.";
                    return false;
                }
            }
        }
    }
    phase("tree done");
    hs->tri_verts.resize(nt);
    hs->tri_shade.resize(nt);
    ParallelFor(nt, [&](size_t i) {
        const RawTriangle &t = tris[order[i]];
        TriVerts &v = hs->tri_verts[i];
        float inst_bits;
        memcpy(&inst_bits, &t.inst, 4);
        v.v0 = {t.p[0].x, t.p[0].y, t.p[0].z, inst_bits};
        float id_bits; // index of the triangle in the scene description (instances in order): what the debug ray entry reports
        memcpy(&id_bits, &order[i], 4);
        v.v1 = {t.p[1].x, t.p[1].y, t.p[1].z, id_bits};
        v.v2 = {t.p[2].x, t.p[2].y, t.p[2].z, 0.0f};
        TriShade &s = hs->tri_shade[i];
        memset(&s, 0, sizeof(s));
        for (int j = 0; j < 3; ++j) {
            s.n[j][0] = t.n[j].x, s.n[j][1] = t.n[j].y, s.n[j][2] = t.n[j].z;
            s.t[j][0] = t.t[j].x, s.t[j][1] = t.t[j].y, s.t[j][2] = t.t[j].z;
            s.uv[j][0] = t.uv[j][0], s.uv[j][1] = t.uv[j][1];
        }
        s.inst = t.inst;
    });
    phase("attribute permutation");
    hs->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    memcpy(hs->scene_bmin, &scene_box.lo, 12);
    memcpy(hs->scene_bmax, &scene_box.hi, 12);

    // ---- covers of the geometry by boxes, for the screen-space visibility pre-pass (wavefront.cu: k_cull_tiles): best-first cuts
    // through the BVH (always open the box with the largest surface area) + the analytic primitives.  A coarse cut (<= 384 boxes)
    // rejects whole 8x8 tiles, a fine one (<= 4096) single pixels.
    {
        struct CutBox {
            float lo[3], hi[3];
            int32_t link;
            float Area() const {
                const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                return dx * dy + dy * dz + dz * dx;
            }
            bool operator<(const CutBox &o) const { return Area() < o.Area(); } // max-heap on the area
        };
        // children of an inner node of whichever binary tree exists: the flattened layout, or the SAH tree behind the wide layout
        auto children = [&](int32_t node, CutBox *a, CutBox *b) {
            if (!binary.empty()) {
                const Bvh2Node &l = binary[binary[node].left], &r = binary[binary[node].right];
                *a = {{l.lo[0], l.lo[1], l.lo[2]}, {l.hi[0], l.hi[1], l.hi[2]}, l.left >= 0 ? binary[node].left : -1};
                *b = {{r.lo[0], r.lo[1], r.lo[2]}, {r.hi[0], r.hi[1], r.hi[2]}, r.left >= 0 ? binary[node].right : -1};
                return;
            }
            const BvhNode &n = hs->nodes[node];
            *a = {{n.c0xy.x, n.c0xy.z, n.cz.x}, {n.c0xy.y, n.c0xy.w, n.cz.y}, n.child0};
            *b = {{n.c1xy.x, n.c1xy.z, n.cz.z}, {n.c1xy.y, n.c1xy.w, n.cz.w}, n.child1};
        };
        // Opens boxes best-first (largest area) until `max_boxes` are held; every box keeps the index of the seed it descends from.
        struct Tagged {
            CutBox box;
            uint32_t seed;
            bool operator<(const Tagged &o) const { return box.Area() < o.box.Area(); }
        };
        auto open_best_first = [&](std::vector<Tagged> seeds, size_t max_boxes) {
            std::vector<Tagged> done; // leaves: cannot be opened
            std::priority_queue<Tagged> open;
            auto add = [&](const Tagged &t) {
                if (t.box.hi[0] < t.box.lo[0] || t.box.hi[1] < t.box.lo[1] || t.box.hi[2] < t.box.lo[2]) return; // the empty box of a one-leaf tree
                if (t.box.link >= 0) open.push(t);
                else done.push_back(t);
            };
            for (const Tagged &t : seeds) add(t);
            CutBox a, b;
            while (!open.empty() && open.size() + done.size() < max_boxes) {
                const Tagged top = open.top();
                open.pop();
                children(top.box.link, &a, &b);
                add(Tagged{a, top.seed});
                add(Tagged{b, top.seed});
            }
            for (; !open.empty(); open.pop()) done.push_back(open.top());
            return done;
        };
        std::vector<Tagged> roots;
        if (!binary.empty() && binary[binary_root].left < 0) { // the whole scene is one leaf
            const Bvh2Node &n = binary[binary_root];
            if (n.hi[0] >= n.lo[0] && n.hi[1] >= n.lo[1] && n.hi[2] >= n.lo[2]) // (not the empty box of an empty tree)
                roots.push_back({{{n.lo[0], n.lo[1], n.lo[2]}, {n.hi[0], n.hi[1], n.hi[2]}, -1}, 0u});
        } else if (!binary.empty() || !hs->nodes.empty()) {
            CutBox a, b;
            children(binary.empty() ? 0 : binary_root, &a, &b);
            roots.push_back({a, 0u});
            roots.push_back({b, 0u});
        }
        // coarse cover; then every coarse box is opened further into its share of the fine cover (grouped by coarse box, so
        // that a tile only looks at the fine boxes under the coarse boxes it meets)
        std::vector<Tagged> coarse = open_best_first(roots, 384);
        for (size_t i = 0; i < coarse.size(); ++i) coarse[i].seed = static_cast<uint32_t>(i);
        std::vector<Tagged> fine = open_best_first(coarse, 4096);
        std::stable_sort(fine.begin(), fine.end(), [](const Tagged &x, const Tagged &y) { return x.seed < y.seed; });
        auto put = [](const CutBox &c, std::vector<float> *out) { out->insert(out->end(), {c.lo[0], c.lo[1], c.lo[2], c.hi[0], c.hi[1], c.hi[2]}); };
        size_t f = 0;
        for (size_t i = 0; i < coarse.size(); ++i) {
            put(coarse[i].box, &hs->cull_boxes);
            hs->cull_fine_begin.push_back(static_cast<uint32_t>(f));
            for (; f < fine.size() && fine[f].seed == i; ++f) put(fine[f].box, &hs->fine_cull_boxes);
        }
        for (const AnalyticPrim &p : hs->analytic) { // an analytic primitive is its own coarse and fine box
            hs->cull_fine_begin.push_back(static_cast<uint32_t>(hs->fine_cull_boxes.size() / 6));
            for (std::vector<float> *out : {&hs->cull_boxes, &hs->fine_cull_boxes})
                out->insert(out->end(), {p.bmin[0], p.bmin[1], p.bmin[2], p.bmax[0], p.bmax[1], p.bmax[2]});
        }
        hs->cull_fine_begin.push_back(static_cast<uint32_t>(hs->fine_cull_boxes.size() / 6));
    }

    // ---- triangle CDFs of mesh area lights (stand for the area-weighted BVH descent of blas.cpp:79-98) ----
    // BLAS::Sample walks the instance's LBVH from the root with thresh = area * xi_0, left when thresh < area(left): a CDF over the
    // LEAVES IN TREE ORDER, and the leaves of a Morton-built tree (bvh_builder.cpp:92-141) stand in the order of their sort keys
    // (30-bit Morton code of the box centre relative to the instance's bounds) << 32 | triangle index.  The CDF is laid out in that
    // order, so the same xi_0 picks the same triangle here as in the reference (what the exact-mode comparison needs; any
    // order gives the same distribution).
    auto expand_bits = [](uint32_t v) { // bvh_builder.cpp:16-22
        v = (v * 0x00010001u) & 0xFF0000FFu;
        v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u;
        v = (v * 0x00000005u) & 0x49249249u;
        return v;
    };
    for (uint32_t light = 0; light < hs->map_area_light_instance.size(); ++light) {
        const uint32_t inst = hs->map_area_light_instance[light];
        DInstance &o = hs->instances[inst];
        if (o.analytic != kInvalid) continue;
        o.light_tri_begin = static_cast<uint32_t>(hs->light_tri_ids.size());
        std::vector<uint32_t> members; // positions in leaf order
        Box all;
        for (size_t i = 0; i < nt; ++i)
            if (tris[order[i]].inst == inst) {
                members.push_back(static_cast<uint32_t>(i));
                all.Grow(boxes[order[i]]);
            }
        const V3 size = all.hi - all.lo;
        std::vector<std::pair<uint64_t, uint32_t>> keyed(members.size());
        for (size_t k = 0; k < members.size(); ++k) {
            const Box &b = boxes[order[members[k]]];
            const V3 c = (b.lo + b.hi) * 0.5f;
            // Vec3 / Vec3 of the reference multiplies by the reciprocal (vec3.cpp:128-133); a flat axis gives 0 * inf = NaN -> cell 0
            const float rel[3] = {(c.x - all.lo.x) * (1.0f / size.x), (c.y - all.lo.y) * (1.0f / size.y), (c.z - all.lo.z) * (1.0f / size.z)};
            uint32_t q[3];
            for (int a = 0; a < 3; ++a) q[a] = static_cast<uint32_t>(fminf(fmaxf(rel[a] * 1024.0f, 0.0f), 1023.0f)); // :39-48 (0/0 -> 0)
            const uint64_t morton = expand_bits(q[0]) * 4u + expand_bits(q[1]) * 2u + expand_bits(q[2]);
            keyed[k] = {(morton << 32) | order[members[k]], members[k]}; // scene-wide triangle index: same order as the instance-local one
        }
        std::sort(keyed.begin(), keyed.end());
        double total = 0.0;
        for (const auto &kv : keyed) {
            hs->light_tri_ids.push_back(kv.second);
            total += tris[order[kv.second]].area;
            hs->light_tri_cdf.push_back(static_cast<float>(total));
        }
        o.light_tri_count = static_cast<uint32_t>(hs->light_tri_ids.size()) - o.light_tri_begin;
        for (uint32_t k = 0; k < o.light_tri_count; ++k)
            hs->light_tri_cdf[o.light_tri_begin + k] = static_cast<float>(hs->light_tri_cdf[o.light_tri_begin + k] / total);
        if (o.light_tri_count) hs->light_tri_cdf[o.light_tri_begin + o.light_tri_count - 1] = 1.0f;
    }

    // ---- emitters (emitter.cpp:122-175, renderer.cpp:522-620) ----
    uint32_t id_sun = kInvalid, id_envmap = kInvalid;
    for (uint64_t i = 0; i < d.num_emitters; ++i) {
        const b200pt_emitter &e = d.emitters[i];
        DEmitter o{};
        o.type = e.type;
        o.id_texture = e.id_texture;
        o.position = {e.position[0], e.position[1], e.position[2]};
        o.direction = {e.direction[0], e.direction[1], e.direction[2]};
        o.radiance = {e.radiance[0], e.radiance[1], e.radiance[2]};
        const M4 to_world = LoadM4(e.to_world);
        o.to_world = ToAffine(to_world);
        o.to_local = ToAffine(Inverse(to_world));
        switch (e.type) {
        case B200PT_EMIT_POINT:
        case B200PT_EMIT_DIRECTIONAL:
            break;
        case B200PT_EMIT_SPOT:
            o.cutoff_angle = e.cutoff_angle;
            o.cos_cutoff_angle = cosf(e.cutoff_angle);
            o.uv_factor = tanf(e.cutoff_angle);
            o.beam_width = e.beam_width;
            o.cos_beam_width = cosf(e.beam_width);
            o.transition_width_rcp = 1.0f / (e.cutoff_angle - e.beam_width);
            o.position = ToF3(TransformPoint(to_world, V3{0, 0, 0}));
            if (!check_texture(e.id_texture, true)) return false;
            break;
        case B200PT_EMIT_SUN:
            if (!check_texture(e.id_texture, false)) return false;
            o.cos_cutoff_angle = e.cos_cutoff_angle;
            id_sun = static_cast<uint32_t>(i);
            break;
        case B200PT_EMIT_ENVMAP: {
            if (!check_texture(e.id_texture, false)) return false;
            const DTexture &tex = hs->textures[e.id_texture];
            if (tex.type != B200PT_TEX_BITMAP) {
                *error = "radiance texture '" + std::to_string(e.id_texture) + "' for emitter '" + std::to_string(i) +
                         "' is not a bitmap."; // renderer.cpp:575-580
                return false;
            }
            if (!BuildEnvMapTables(tex, d.pixels, &hs->envmap_tables, &hs->envmap_normalization, error)) return false;
            o.env_width = tex.width, o.env_height = tex.height;
            o.env_normalization = hs->envmap_normalization;
            // InitEnvMap (emitter.cpp:166-175) against the packing [rows | weights | cols] (Q9)
            o.env_cdf_cols = 0;
            o.env_cdf_rows = tex.height + 1;
            o.env_weight_rows = (tex.height + 1) + tex.height;
            id_envmap = static_cast<uint32_t>(i);
            break;
        }
        case B200PT_EMIT_CONSTANT:
            id_envmap = static_cast<uint32_t>(i);
            break;
        default:
            *error = "unknow emitter type."; // renderer.cpp:563
            return false;
        }
        hs->emitters.push_back(o);
    }

    // ---- integrator (renderer.cpp:622-676) ----
    DIntegrator &ig = hs->integrator;
    ig.type = d.integrator.type;
    if (ig.type != B200PT_INTEGRATOR_PATH && ig.type != B200PT_INTEGRATOR_VOLPATH) {
        *error = "unknow integrator type"; // renderer.cpp:664
        return false;
    }
    ig.hide_emitters = d.integrator.hide_emitters;
    ig.pdf_rr = d.integrator.pdf_rr;
    ig.pdf_rr_rcp = d.integrator.pdf_rr; // Q1: renderer.cpp:634 stores pdf_rr, not its reciprocal
    ig.depth_rr = d.integrator.depth_rr;
    ig.depth_max = d.integrator.depth_max;
    ig.num_emitters = static_cast<uint32_t>(d.num_emitters);
    ig.num_area_lights = static_cast<uint32_t>(hs->map_area_light_instance.size());
    ig.id_sun = id_sun;
    ig.id_envmap = id_envmap;
    // Shading bins: with two or more BSDF models in the scene, hits are filed by model and shaded model by model.
    {
        uint32_t bins = 0, models = 0;
        for (const DInstance &in : hs->instances) bins |= 1u << (in.id_bsdf == kInvalid ? 0u : hs->bsdfs[in.id_bsdf].type);
        for (uint32_t t = B200PT_BSDF_DIFFUSE; t <= B200PT_BSDF_PLASTIC; ++t) models += (bins >> t) & 1u;
        // escaped rays still need shading when an environment map lights them or a medium may scatter them first
        if (id_envmap != kInvalid || ig.type == B200PT_INTEGRATOR_VOLPATH) bins |= 1u;
        // One model only: binning is still worth it when many traced rays escape (their dead queue entries then never reach
        // the shading kernel), i.e. in open scenes without an environment map; B200PT_BIN_SINGLE=0/1 overrides.
        const char *bin_single = getenv("B200PT_BIN_SINGLE");
        const bool single_binned = bin_single ? atoi(bin_single) != 0 : false;
        ig.shade_bins = (models >= 2 || (models == 1 && single_binned)) ? bins : 0u;
        hs->tri_bsdf_type.resize(hs->tri_shade.size());
        for (size_t i = 0; i < hs->tri_shade.size(); ++i) {
            const uint32_t id_bsdf = hs->instances[hs->tri_shade[i].inst].id_bsdf;
            hs->tri_bsdf_type[i] = static_cast<uint8_t>(id_bsdf == kInvalid ? 0u : hs->bsdfs[id_bsdf].type);
        }
    }
    ig.has_opacity = 0;
    for (const DInstance &in : hs->instances)
        if (in.id_bsdf != kInvalid && hs->bsdfs[in.id_bsdf].id_opacity != kInvalid) ig.has_opacity = 1;

    whole("BVH .. integrator");
    // ---- Kulla-Conty LUTs (renderer.cpp:311-314); only read by conductor/dielectric BSDFs ----
    hs->kc_brdf_avg.assign(kLutResolution * kLutResolution, 0.0f);
    hs->kc_albedo_avg.assign(kLutResolution, 0.0f);
    if (need_kulla_conty) ComputeKullaContyTables(hs->kc_brdf_avg.data(), hs->kc_albedo_avg.data());
    whole("Kulla-Conty tables");
    return true;
}

} // namespace b200pt
