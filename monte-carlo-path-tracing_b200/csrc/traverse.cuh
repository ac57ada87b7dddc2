// traverse.cuh — BVH closest-hit / any-hit traversal and ray-primitive tests.
//
// Replaces TLAS::Intersect / IntersectAny (src/rtcore/accel/tlas.cpp:13-76),
// BLAS::Intersect / IntersectAny (src/rtcore/accel/blas.cpp:18-77), AABB::Intersect
// (src/rtcore/accel/aabb.cpp:29-48) and the primitive tests of src/rtcore/primitives/*.
// Geometry is in world space in one BVH2 (the reference's instances carry no transform of
// their own: to_world is baked into the triangles, scene.cpp:261-281), so the two-level
// walk collapses into one.
#pragma once
#include "textures.cuh"
#include "vecmath.cuh"

namespace b200pt {

struct Ray {
    V3 o, d;
    float tmin, tmax;
};

struct HitRec {       // 16 B, what the closest-hit kernel writes per ray
    float t;
    uint32_t prim;    // triangle index (| kPrimInsideBit) | kPrimAnalyticBit+index | kPrimMiss
    float u, v;       // triangle: barycentric weights of vertex 0 and 1; .u sign bit unused
};

// Per-ray constants of Woop's watertight test (ray.cpp:24-46) and of the slab test (ray.cpp:21-22).
struct RayPre {
    V3 idir, ood;     // 1/d (with the reference's 1e-4 substitute for zero components), o/d
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

__device__ __forceinline__ RayPre Precompute(const Ray &r) {
    RayPre p;
    p.idir = {1.0f / (r.d.x != 0 ? r.d.x : kEpsilonDistance), 1.0f / (r.d.y != 0 ? r.d.y : kEpsilonDistance),
              1.0f / (r.d.z != 0 ? r.d.z : kEpsilonDistance)};
    p.ood = r.o * p.idir;
    const float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
    p.kz = (ax > ay && ax > az) ? 0 : (ay > az ? 1 : 2);
    p.kx = p.kz + 1;
    if (p.kx == 3) p.kx = 0;
    p.ky = p.kx + 1;
    if (p.ky == 3) p.ky = 0;
    const float dz = Comp(r.d, p.kz);
    if (dz < 0.0f) {
        const int t = p.kx;
        p.kx = p.ky;
        p.ky = t;
    }
    p.Sx = Comp(r.d, p.kx) / dz;
    p.Sy = Comp(r.d, p.ky) / dz;
    p.Sz = 1.0f / dz;
    return p;
}

// a[k] for a per-lane k in {0,1,2} as two SEL instructions.  Written in PTX because nvcc turns the C++ conditional
// into divergent branches (BSSY/BRA/BSYNC per component: profiles/r01_full_first8.txt, 160 instructions per test).
__device__ __forceinline__ float Sel3(const V3 &a, int k) {
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.s32 p0, %4, 0;\n\tsetp.eq.s32 p1, %4, 1;\n\tselp.f32 %0, %2, %3, p1;\n\t"
        "selp.f32 %0, %1, %0, p0;\n\t}"
        : "=f"(r)
        : "f"(a.x), "f"(a.y), "f"(a.z), "r"(k));
    return r;
}

// Woop, Benthin, Wald 2013 — triangle.cpp:23-87.  Products and differences are kept un-fused
// (__fmul_rn/__fsub_rn) so shared edges evaluate identically from both sides.
__device__ __forceinline__ bool IntersectTriangleWoop(const Ray &ray, const RayPre &pre, const float4 &p0, const float4 &p1,
                                                      const float4 &p2, float *t_out, float *u_out, float *v_out,
                                                      bool *inside) {
    const V3 A = {p0.x - ray.o.x, p0.y - ray.o.y, p0.z - ray.o.z};
    const V3 B = {p1.x - ray.o.x, p1.y - ray.o.y, p1.z - ray.o.z};
    const V3 C = {p2.x - ray.o.x, p2.y - ray.o.y, p2.z - ray.o.z};
    const float Akz = Sel3(A, pre.kz), Bkz = Sel3(B, pre.kz), Ckz = Sel3(C, pre.kz);
    const float Ax = __fsub_rn(Sel3(A, pre.kx), __fmul_rn(pre.Sx, Akz)), Ay = __fsub_rn(Sel3(A, pre.ky), __fmul_rn(pre.Sy, Akz));
    const float Bx = __fsub_rn(Sel3(B, pre.kx), __fmul_rn(pre.Sx, Bkz)), By = __fsub_rn(Sel3(B, pre.ky), __fmul_rn(pre.Sy, Bkz));
    const float Cx = __fsub_rn(Sel3(C, pre.kx), __fmul_rn(pre.Sx, Ckz)), Cy = __fsub_rn(Sel3(C, pre.ky), __fmul_rn(pre.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) { // exact edge hits: redo in double (triangle.cpp:50-63)
        U = static_cast<float>(__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx)));
        V = static_cast<float>(__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx)));
        W = static_cast<float>(__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax)));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = pre.Sz * Akz, Bz = pre.Sz * Bkz, Cz = pre.Sz * Ckz;
    // un-fused like the reference build (g++ for baseline x86-64 has no fma to contract into): t is then BIT-EQUAL to
    // TLAS::Intersect's (tests/test_gpu_traversal.py)
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float det_inv = 1.0f / det;
    const float t = T * det_inv;
    if (t > ray.tmax || t < ray.tmin) return false;
    *t_out = t;
    *u_out = U * det_inv;
    *v_out = V * det_inv;
    *inside = det_inv < 0;
    return true;
}

// aabb.cpp:29-48 for a world-space box.
__device__ __forceinline__ bool IntersectBox(const float *bmin, const float *bmax, const Ray &ray, const RayPre &pre) {
    const float tx0 = (bmin[0] - ray.o.x) * pre.idir.x, tx1 = (bmax[0] - ray.o.x) * pre.idir.x;
    const float ty0 = (bmin[1] - ray.o.y) * pre.idir.y, ty1 = (bmax[1] - ray.o.y) * pre.idir.y;
    const float tz0 = (bmin[2] - ray.o.z) * pre.idir.z, tz1 = (bmax[2] - ray.o.z) * pre.idir.z;
    const float t_enter = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), ray.tmin));
    const float t_exit = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), ray.tmax));
    return t_enter <= t_exit;
}

// Distance-only halves of sphere.cpp:17-44, disk.cpp:17-40, cylinder.cpp:21-60.
// The hit attributes (normal, uv, frame) are rebuilt in the shading stage from the hit point.
// `uv_out` (optional) receives the texcoord the reference hands to Bsdf::IsTransparent (sphere.cpp:39-41, disk.cpp:38-40,
// cylinder.cpp:46-49).
__device__ __forceinline__ bool IntersectAnalytic(const AnalyticPrim &p, const Ray &ray, float *t_out, V2 *uv_out = nullptr) {
    const V3 o_l = XformPoint(p.to_local, ray.o), d_l = XformVector(p.to_local, ray.d);
    if (p.type == kSphere) {
        const V3 ro = o_l - mk3(p.center);
        const float a = Dot(d_l, d_l), b = 2.0f * Dot(d_l, ro), c = Dot(ro, ro) - Sqr(p.radius);
        float t_near = 0.0f, t_far = 0.0f;
        if (!SolveQuadratic(a, b, c, &t_near, &t_far) || t_far < kEpsilonDistance) return false;
        float t = t_near < kEpsilonDistance ? t_far : t_near;
        const V3 pos_l = ro + t * d_l, pos = XformPoint(p.to_world, pos_l + mk3(p.center));
        t = Length(pos - ray.o);
        if (t > ray.tmax || t < ray.tmin) return false;
        if (uv_out != nullptr) {
            float theta, phi;
            CartesianToSpherical(pos_l, &theta, &phi, nullptr);
            *uv_out = {phi * k1Div2Pi, theta * k1DivPi};
        }
        *t_out = t;
        return true;
    } else if (p.type == kDisk) {
        const float t_z = -o_l.z / d_l.z;
        if (t_z < kEpsilonFloat) return false;
        const V3 pos_l = o_l + t_z * d_l;
        if (Length(pos_l) > 0.5f) return false;
        const V3 pos = XformPoint(p.to_world, pos_l);
        const float t = Length(pos - ray.o);
        if (t > ray.tmax || t < ray.tmin) return false;
        if (uv_out != nullptr) {
            float theta, phi, r;
            CartesianToSpherical(pos_l, &theta, &phi, &r);
            *uv_out = {r, phi * k1Div2Pi};
        }
        *t_out = t;
        return true;
    } else {
        const float a = Sqr(d_l.x) + Sqr(d_l.y), b = 2.0f * (d_l.x * o_l.x + d_l.y * o_l.y),
                    c = Sqr(o_l.x) + Sqr(o_l.y) - Sqr(p.radius);
        float t_near = 0.0f, t_far = 0.0f;
        if (!SolveQuadratic(a, b, c, &t_near, &t_far) || t_far < kEpsilonDistance) return false;
        const float z_near = o_l.z + d_l.z * t_near, z_far = o_l.z + d_l.z * t_far;
        float t = 0;
        if (kEpsilonDistance < t_near && 0.0f <= z_near && z_near <= p.length)
            t = t_near;
        else if (0.0f <= z_far && z_far <= p.length)
            t = t_far;
        else
            return false;
        const V3 pos_l = o_l + t * d_l;
        const V3 pos = XformPoint(p.to_world, pos_l);
        t = Length(pos - ray.o);
        if (t > ray.tmax || t < ray.tmin) return false;
        if (uv_out != nullptr) *uv_out = {atan2f(pos_l.y, pos_l.x) * k1Div2Pi, pos_l.z / p.length};
        *t_out = t;
        return true;
    }
}

// Stochastic alpha test of a candidate hit: Bsdf::IsTransparent (bsdf.cpp:272-276) as the reference calls it from inside
// every primitive test (triangle.cpp:116-118, sphere.cpp:42, disk.cpp:41, cylinder.cpp:50).  A transparent verdict makes
// the primitive invisible to THIS ray only; the random number is a function of the ray's own Philox stream AND the
// primitive (the reference advances the per-pixel LCG here), so the verdict does not depend on the order in which a
// traversal happens to meet the candidates: both tree layouts and both loop shapes take the same decisions.
__device__ __forceinline__ bool OpacityRejects(const DeviceScene &scene, uint32_t inst, V2 uv, const Rng &rng, uint32_t prim) {
    const uint32_t id_bsdf = scene.instances[inst].id_bsdf;
    if (id_bsdf == kInvalid) return false;
    const uint32_t id_opacity = scene.bsdfs[id_bsdf].id_opacity;
    if (id_opacity == kInvalid) return false;
    return TexIsTransparent(scene, id_opacity, uv, rng.ForPrimitive(prim));
}

// Random-number domains (4th Philox counter word) of the alpha tests inside closest-hit and any-hit traversal; the
// shading stage uses Rng's default domain.
constexpr uint32_t kRngDomainClosest = 0x0a1fa001u, kRngDomainShadow = 0x0a1fa002u;

struct TraversalCounters {
    uint32_t nodes = 0, prims = 0;
};

#ifndef B200PT_SPECULATIVE
#define B200PT_SPECULATIVE 0   // measured variant, see profiles/README.md
#endif
constexpr int kStackSize = 64;
constexpr int kSentinel = 0x7FFFFFFF;
constexpr int kTopNodes = 512;     // default number of BVH nodes staged in shared memory (32 KB per CTA)
constexpr int kTopNodesMax = 2048; // upper bound (128 KB)

// One BVH2 node = 4 x 16 B loads; `top` is the shared-memory copy of nodes [0, num_top).
// TOP = false compiles the shared-memory path out (plain read-only loads, all of the SM's unified L1 left to the cache).
template <bool TOP>
__device__ __forceinline__ void LoadNode(const BvhNode *__restrict__ nodes, const float4 *top, int num_top, int index,
                                         float4 *n0, float4 *n1, float4 *nz, int *c0, int *c1) {
    const float4 *src = (TOP && index < num_top) ? top + index * 4 : reinterpret_cast<const float4 *>(nodes + index);
    if (TOP && index < num_top) {
        *n0 = src[0], *n1 = src[1], *nz = src[2];
        const float4 links = src[3];
        *c0 = __float_as_int(links.x), *c1 = __float_as_int(links.y);
    } else {
        *n0 = __ldg(src), *n1 = __ldg(src + 1), *nz = __ldg(src + 2);
        const float4 links = __ldg(src + 3);
        *c0 = __float_as_int(links.x), *c1 = __float_as_int(links.y);
    }
}

// Warp-aggregated queue append usable from divergent code: the currently active lanes reserve
// consecutive slots with one atomic.
__device__ __forceinline__ uint32_t AppendCoalesced(uint32_t *counter) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// Persistent-threads traversal (Aila & Laine style "while-while" with per-lane ray replacement).
//
// Every lane owns at most one ray.  A lane whose ray has terminated immediately fetches the next
// ray index from a global work counter (one warp-aggregated atomic for all idle lanes), so lanes
// never idle behind the longest ray of their warp; inside an iteration all lanes first walk inner
// nodes (inner while), then all lanes that reached a leaf intersect its triangles, which keeps the
// two divergent phases apart.  Incoherent bounce rays ran at 5.6 active lanes per instruction with
// one-ray-per-thread launches (profiles/r01_extend_baseline.txt); this loop is the fix.
//
//   fetch(index, &ray, &ctr, &any) -> bool   builds ray `index` (false: nothing to trace for this index) and says whether it
//                                            is an occlusion ray (`any`: the first hit ends it) or a closest-hit ray; with
//                                            OPACITY it also returns the (pixel, sample, depth) counter of the ray's
//                                            random-number stream
//   finish(index, hit, found, any)           consumes the result (closest hit record, or occlusion flag)
// OPACITY compiles the stochastic alpha test into the primitive tests (scenes whose BSDFs carry opacity textures).
// MIXED = false promises that no ray is an occlusion ray (camera rays) and compiles the per-lane flag out.
// `counters` / `rays_traced` have two entries: [0] closest-hit rays, [1] occlusion rays.
template <bool MIXED, bool STATS, bool OPACITY, bool TOP, typename Fetch, typename Finish>
__device__ __forceinline__ void TraversePersistent(const DeviceScene &scene, const float4 *top, int num_top, uint32_t num_rays,
                                                   uint32_t *work_counter, int refill_threshold, int min_inner_lanes, uint2 key, Fetch fetch, Finish finish,
                                                   TraversalCounters *counters, uint32_t *rays_traced) {
    // The top of the traversal stack lives in a register: a pop hands `tos` over at once and re-loads the next one in the
    // background, so the local-memory load is off the chain  pop -> node address -> node fetch  that bounds a step.
    // stack[0] is a permanent sentinel, stack[1 + i] holds entry i below the top (Pop / Push below).
    int stack[kStackSize + 1];
    int sp = 0, cur = kSentinel, tos = kSentinel;
    stack[0] = kSentinel;
    auto Push = [&](int node) {
        stack[sp] = tos; // sp == 0: rewrites the sentinel with itself
        tos = node;
        ++sp;
    };
    auto Pop = [&]() {
        const int node = tos; // kSentinel when the stack is empty
        sp = max(sp - 1, 0);
        tos = stack[sp];
        return node;
    };
#if B200PT_SPECULATIVE
    int postponed = 0; // a leaf link (< 0) waiting to be intersected, 0 = none
#endif
    bool has = false, exhausted = false, found = false, any = false;
    uint32_t index = 0;
    Ray ray;
    ray.o = ray.d = mk3(0.0f);
    ray.tmin = ray.tmax = 0.0f;
    RayPre pre = Precompute(ray);
    HitRec hit;
    hit.t = 0.0f, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
    Rng rng(0, 0, 0, key, kRngDomainClosest);

    for (;;) {
        // Refill idle lanes in groups: the fetch path (ray load, 1/d, shear constants) is long, so it is run
        // when at least `refill_threshold` lanes are idle (or nothing else is left to do), not lane by lane.
        const unsigned idle = __ballot_sync(0xffffffffu, !has && !exhausted);
        const bool refill = __popc(idle) >= refill_threshold || __all_sync(0xffffffffu, !has);
        if (refill && !has && !exhausted) {
            index = AppendCoalesced(work_counter);
            if (index >= num_rays) {
                exhausted = true;
            } else if (uint3 ctr = make_uint3(0, 0, 0); fetch(index, &ray, &ctr, &any)) {
                if (!MIXED) any = false;
                if (OPACITY) rng = Rng(ctr.x, ctr.y, ctr.z, key, any ? kRngDomainShadow : kRngDomainClosest);
                pre = Precompute(ray);
                hit.t = ray.tmax, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
                found = false;
                has = true;
                sp = 0, tos = kSentinel;
                cur = scene.num_nodes ? 0 : kSentinel;
                ++rays_traced[any];
                // Analytic primitives (spheres, disks, cylinders) are few: tested linearly up front.
                for (uint32_t i = 0; i < scene.num_analytic; ++i) {
                    const AnalyticPrim &p = scene.analytic[i];
                    if (STATS) ++counters[any].nodes;
                    if (!IntersectBox(p.bmin, p.bmax, ray, pre)) continue;
                    if (STATS) ++counters[any].prims;
                    float t;
                    V2 uv = {0.0f, 0.0f};
                    if (IntersectAnalytic(p, ray, &t, OPACITY ? &uv : nullptr)) {
                        if (OPACITY && OpacityRejects(scene, p.inst, uv, rng, kPrimAnalyticBit | i)) continue;
                        found = true;
                        if (any) {
                            cur = kSentinel;
                            break;
                        }
                        ray.tmax = t;
                        hit.t = t;
                        hit.prim = kPrimAnalyticBit | i;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, !has)) break;
        if (has) {
            // ---- inner nodes: descend until this lane holds a leaf or runs out of work ----
            while (static_cast<unsigned>(cur) < static_cast<unsigned>(kSentinel)) {
                float4 n0, n1, nz;
                int child0, child1;
                LoadNode<TOP>(scene.nodes, top, num_top, cur, &n0, &n1, &nz, &child0, &child1);
                if (STATS) counters[any].nodes += 2;
                const float c0lox = fmaf(n0.x, pre.idir.x, -pre.ood.x), c0hix = fmaf(n0.y, pre.idir.x, -pre.ood.x);
                const float c0loy = fmaf(n0.z, pre.idir.y, -pre.ood.y), c0hiy = fmaf(n0.w, pre.idir.y, -pre.ood.y);
                const float c0loz = fmaf(nz.x, pre.idir.z, -pre.ood.z), c0hiz = fmaf(nz.y, pre.idir.z, -pre.ood.z);
                const float c1lox = fmaf(n1.x, pre.idir.x, -pre.ood.x), c1hix = fmaf(n1.y, pre.idir.x, -pre.ood.x);
                const float c1loy = fmaf(n1.z, pre.idir.y, -pre.ood.y), c1hiy = fmaf(n1.w, pre.idir.y, -pre.ood.y);
                const float c1loz = fmaf(nz.z, pre.idir.z, -pre.ood.z), c1hiz = fmaf(nz.w, pre.idir.z, -pre.ood.z);
                const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), ray.tmin));
                const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), ray.tmax));
                const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), ray.tmin));
                const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), ray.tmax));
                const bool hit0 = c0min <= c0max, hit1 = c1min <= c1max;
                // Straight-line selection of the next node (three short predicated groups, no three-way branch: the lanes that
                // missed both boxes used to pop in a pass of their own at 2.6 active lanes, profiles/r02_k_trace_source.csv.gz).
                const bool both = hit0 && hit1, none = !(hit0 || hit1);
                const bool swap = c1min < c0min;
                if (both) Push(swap ? child0 : child1);
                const int next = (hit1 && (swap || !hit0)) ? child1 : child0; // the nearer of two hits, or the only one
                cur = none ? tos : next;
                if (none) {
                    sp = max(sp - 1, 0);
                    tos = stack[sp];
                }
                // Most lanes of an incoherent warp reach their next leaf within a few steps while a few stragglers
                // keep descending; once fewer than `min_inner_lanes` lanes are still walking inner nodes the warp
                // leaves the inner phase so the waiting lanes can intersect their leaves (stragglers resume later).
#if B200PT_SPECULATIVE
                // Speculative descent (Aila & Laine): the first leaf a lane reaches is put aside and the lane keeps walking
                // inner nodes with the rest of the warp; it only waits once it holds a second one.
                if (cur < 0 && postponed == 0) {
                    postponed = cur;
                    cur = Pop();
                }
#endif
                if (__popc(__activemask()) < min_inner_lanes) break;
            }
            // ---- leaves ----
            auto process_leaf = [&](int link) {
                const uint32_t leaf = static_cast<uint32_t>(~link);
                const uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
                const float4 *verts = reinterpret_cast<const float4 *>(scene.tri_verts + first);
                for (uint32_t j = 0; j < count; ++j) {
                    const float4 p0 = __ldg(verts + 3 * j), p1 = __ldg(verts + 3 * j + 1), p2 = __ldg(verts + 3 * j + 2);
                    if (STATS) ++counters[any].prims;
                    float t, u, v;
                    bool inside;
                    if (IntersectTriangleWoop(ray, pre, p0, p1, p2, &t, &u, &v, &inside)) {
                        if (OPACITY) { // triangle.cpp:115-118: texcoord = Lerp(texcoords, u, v, w)
                            const float *tc = &scene.tri_shade[first + j].uv[0][0];
                            const float w = 1.0f - u - v;
                            const V2 uv = {u * __ldg(tc) + v * __ldg(tc + 2) + w * __ldg(tc + 4), u * __ldg(tc + 1) + v * __ldg(tc + 3) + w * __ldg(tc + 5)};
                            if (OpacityRejects(scene, __float_as_uint(p0.w), uv, rng, __float_as_uint(p1.w))) continue;
                        }
                        found = true;
                        if (any) {
                            cur = kSentinel;
                            break;
                        }
                        ray.tmax = t;
                        hit.t = t;
                        hit.prim = (first + j) | (inside ? kPrimInsideBit : 0u);
                        hit.u = u;
                        hit.v = v;
                    }
                }
            };
#if B200PT_SPECULATIVE
            // first the leaf that was put aside while the lane kept descending, then the one it stopped at
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                int link = 0;
                if (pass == 0) {
                    link = postponed;
                    postponed = 0;
                } else if (cur < 0) {
                    link = cur;
                    cur = Pop();
                }
                if (link < 0) process_leaf(link);
            }
#else
            if (cur < 0) {
                const int link = cur;
                cur = Pop();
                process_leaf(link);
            }
#endif
            if (cur == kSentinel) {
                finish(index, hit, found, any);
                has = false;
            }
        }
    }
}

// One ray per lane, walked to the end by its own lane: the plain traversal loop, for the path-at-a-time tail kernel
// (k_tail), where a lane owns a whole path and there is no queue to refill from.  Same node / leaf / primitive tests and
// the same order of visits as TraversePersistent, so both find the same hit.  `opacity` / `stats` are warp-uniform
// run-time switches here (the tail handles a few thousand paths; it does not need a variant per scene kind).
// Returns whether something was hit (closest-hit rays: the record is in *hit_out).
__device__ __forceinline__ bool TraverseSingle(const DeviceScene &scene, Ray ray, bool any, bool opacity, Rng rng, HitRec *hit_out,
                                               bool stats, TraversalCounters *counters) {
    int stack[kStackSize];
    int sp = 0;
    const RayPre pre = Precompute(ray);
    HitRec hit;
    hit.t = ray.tmax, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
    bool found = false;
    for (uint32_t i = 0; i < scene.num_analytic; ++i) {
        const AnalyticPrim &p = scene.analytic[i];
        if (stats) ++counters->nodes;
        if (!IntersectBox(p.bmin, p.bmax, ray, pre)) continue;
        if (stats) ++counters->prims;
        float t;
        V2 uv = {0.0f, 0.0f};
        if (IntersectAnalytic(p, ray, &t, opacity ? &uv : nullptr)) {
            if (opacity && OpacityRejects(scene, p.inst, uv, rng, kPrimAnalyticBit | i)) continue;
            found = true;
            if (any) return true;
            ray.tmax = t;
            hit.t = t;
            hit.prim = kPrimAnalyticBit | i;
        }
    }
    int cur = scene.num_nodes ? 0 : kSentinel;
    while (cur != kSentinel) {
        if (cur >= 0) {
            float4 n0, n1, nz;
            int child0, child1;
            LoadNode<false>(scene.nodes, nullptr, 0, cur, &n0, &n1, &nz, &child0, &child1);
            if (stats) counters->nodes += 2;
            const float c0lox = fmaf(n0.x, pre.idir.x, -pre.ood.x), c0hix = fmaf(n0.y, pre.idir.x, -pre.ood.x);
            const float c0loy = fmaf(n0.z, pre.idir.y, -pre.ood.y), c0hiy = fmaf(n0.w, pre.idir.y, -pre.ood.y);
            const float c0loz = fmaf(nz.x, pre.idir.z, -pre.ood.z), c0hiz = fmaf(nz.y, pre.idir.z, -pre.ood.z);
            const float c1lox = fmaf(n1.x, pre.idir.x, -pre.ood.x), c1hix = fmaf(n1.y, pre.idir.x, -pre.ood.x);
            const float c1loy = fmaf(n1.z, pre.idir.y, -pre.ood.y), c1hiy = fmaf(n1.w, pre.idir.y, -pre.ood.y);
            const float c1loz = fmaf(nz.z, pre.idir.z, -pre.ood.z), c1hiz = fmaf(nz.w, pre.idir.z, -pre.ood.z);
            const float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), ray.tmin));
            const float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), ray.tmax));
            const float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), ray.tmin));
            const float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), ray.tmax));
            const bool hit0 = c0min <= c0max, hit1 = c1min <= c1max;
            if (!hit0 && !hit1) {
                cur = sp > 0 ? stack[--sp] : kSentinel;
            } else if (hit0 && hit1) {
                const bool swap = c1min < c0min;
                stack[sp++] = swap ? child0 : child1;
                cur = swap ? child1 : child0;
            } else {
                cur = hit0 ? child0 : child1;
            }
        } else {
            const uint32_t leaf = static_cast<uint32_t>(~cur);
            const uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
            const float4 *verts = reinterpret_cast<const float4 *>(scene.tri_verts + first);
            cur = sp > 0 ? stack[--sp] : kSentinel;
            for (uint32_t j = 0; j < count; ++j) {
                const float4 p0 = __ldg(verts + 3 * j), p1 = __ldg(verts + 3 * j + 1), p2 = __ldg(verts + 3 * j + 2);
                if (stats) ++counters->prims;
                float t, u, v;
                bool inside;
                if (IntersectTriangleWoop(ray, pre, p0, p1, p2, &t, &u, &v, &inside)) {
                    if (opacity) {
                        const float *tc = &scene.tri_shade[first + j].uv[0][0];
                        const float w = 1.0f - u - v;
                        const V2 uv = {u * __ldg(tc) + v * __ldg(tc + 2) + w * __ldg(tc + 4), u * __ldg(tc + 1) + v * __ldg(tc + 3) + w * __ldg(tc + 5)};
                        if (OpacityRejects(scene, __float_as_uint(p0.w), uv, rng, __float_as_uint(p1.w))) continue;
                    }
                    found = true;
                    if (any) return true;
                    ray.tmax = t;
                    hit.t = t;
                    hit.prim = (first + j) | (inside ? kPrimInsideBit : 0u);
                    hit.u = u;
                    hit.v = v;
                }
            }
        }
    }
    *hit_out = hit;
    return found;
}

} // namespace b200pt
