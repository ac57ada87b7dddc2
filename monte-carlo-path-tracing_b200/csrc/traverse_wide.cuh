// traverse_wide.cuh — closest-hit / any-hit traversal of the compressed 8-wide BVH (WideNode, device_scene.h).
//
// Same job as traverse.cuh (TLAS::Intersect / IntersectAny src/rtcore/accel/tlas.cpp:13-76, BLAS::Intersect
// blas.cpp:18-77, AABB::Intersect aabb.cpp:29-48) on the default acceleration structure: one 80-byte node fetch decides
// eight children, so a ray takes a third of the dependent fetches of the binary tree, and the traversal stack holds
// (base index, hit mask) groups instead of single nodes (Ylitie, Karras, Laine: "Efficient incoherent ray traversal on
// GPUs through compressed wide BVHs", HPG 2017).  The ray-triangle test is the unchanged Woop test of traverse.cuh on the
// unchanged world-space vertices, so both trees report the same hits.
#pragma once
#include "traverse.cuh"

namespace b200pt {

constexpr int kWideStack = 64;          // = kWideStackEntries (bvh_wide.hpp): the builder rejects deeper trees
constexpr int kWideNodeVec = 5;         // uint4 loads per node
constexpr int kWideTopNodesMax = 2048;  // nodes of the top of the tree that fit the shared-memory copy (160 KB)

// 8388608 + b for byte `j` of `w`, as a float (b sits in the low mantissa bits): one PRMT.
template <int J>
__device__ __forceinline__ float ByteAsBiasedFloat(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + J));
}

// 0xff in every byte of the result whose counterpart in `x` has its top bit set, 0x00 elsewhere.  PTX prmt with the
// "replicate sign" selector bit; __byte_perm() masks that bit off, hence the inline PTX.
__device__ __forceinline__ uint32_t SignExtendBytes(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

#ifndef B200PT_WIDE_X2
#define B200PT_WIDE_X2 0   // 1: packed fp32x2 plane distances (measured variant, profiles/README.md)
#endif

// (b0 - 2^23) * adj + org and (b1 - 2^23) * adj + org with one packed add and one packed fma.
__device__ __forceinline__ void Plane2(float b0, float b1, float adj, float org, float *t0, float *t1) {
    asm("{\n\t.reg .b64 v, m, s, o;\n\t"
        "mov.b64 v, {%2, %3};\n\tmov.b64 m, {%4, %4};\n\tadd.rn.f32x2 v, v, m;\n\t"
        "mov.b64 s, {%5, %5};\n\tmov.b64 o, {%6, %6};\n\tfma.rn.f32x2 v, v, s, o;\n\t"
        "mov.b64 {%0, %1}, v;\n\t}"
        : "=f"(*t0), "=f"(*t1)
        : "f"(b0), "f"(b1), "f"(-8388608.0f), "f"(adj), "f"(org));
}

struct WideRay {            // per-ray constants of the node test
    float idx, idy, idz;    // 1 / d (ray.cpp:21-22 with its 1e-4 substitute)
    uint32_t octinv4;       // r * 0x01010101, r = bits of the positive direction components (slot visiting order s ^ r)
};

__device__ __forceinline__ WideRay MakeWideRay(const RayPre &pre) {
    WideRay w;
    w.idx = pre.idir.x, w.idy = pre.idir.y, w.idz = pre.idir.z;
    const uint32_t r = (pre.idir.x >= 0.0f ? 1u : 0u) | (pre.idir.y >= 0.0f ? 2u : 0u) | (pre.idir.z >= 0.0f ? 4u : 0u);
    w.octinv4 = r * 0x01010101u;
    return w;
}

// Slab test (aabb.cpp:29-48) of the eight quantised child boxes of one node.  Returns the hit mask: bit 24 + (s ^ r)
// for a hit inner child in slot s, bits [offset, offset + count) for the triangles of a hit leaf.
__device__ __forceinline__ uint32_t WideNodeHits(const uint4 &n0, const uint4 &n1, const uint4 &n2, const uint4 &n3, const uint4 &n4,
                                                 const Ray &ray, const WideRay &wr) {
    const uint32_t e_imask = n0.w;
    // grid spacing 2^(e - 127): the exponent byte moved into a float's exponent field
    const float adjx = __uint_as_float((e_imask << 23) & 0x7f800000u) * wr.idx;
    const float adjy = __uint_as_float((e_imask << 15) & 0x7f800000u) * wr.idy;
    const float adjz = __uint_as_float((e_imask << 7) & 0x7f800000u) * wr.idz;
    const float orgx = (__uint_as_float(n0.x) - ray.o.x) * wr.idx;
    const float orgy = (__uint_as_float(n0.y) - ray.o.y) * wr.idy;
    const float orgz = (__uint_as_float(n0.z) - ray.o.z) * wr.idz;
    const bool negx = wr.idx < 0.0f, negy = wr.idy < 0.0f, negz = wr.idz < 0.0f;
    constexpr float kBias = 8388608.0f;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = SignExtendBytes(is_inner4 << 3); // 0xff in the bytes of inner children
        const uint32_t bit_index4 = (meta4 ^ (wr.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        // the plane a ray enters through is the low one for a positive direction component, the high one otherwise
        const uint32_t nearx = negx ? hix : lox, farx = negx ? lox : hix;
        const uint32_t neary = negy ? hiy : loy, fary = negy ? loy : hiy;
        const uint32_t nearz = negz ? hiz : loz, farz = negz ? loz : hiz;
#if B200PT_WIDE_X2
        // two children per instruction: Blackwell's packed fp32 add / fma (FADD2 / FFMA2) on register pairs
        auto pair = [&](auto jc) {
            constexpr int J = decltype(jc)::value;
            float tnx0, tnx1, tny0, tny1, tnz0, tnz1, tfx0, tfx1, tfy0, tfy1, tfz0, tfz1;
            Plane2(ByteAsBiasedFloat<J>(nearx), ByteAsBiasedFloat<J + 1>(nearx), adjx, orgx, &tnx0, &tnx1);
            Plane2(ByteAsBiasedFloat<J>(neary), ByteAsBiasedFloat<J + 1>(neary), adjy, orgy, &tny0, &tny1);
            Plane2(ByteAsBiasedFloat<J>(nearz), ByteAsBiasedFloat<J + 1>(nearz), adjz, orgz, &tnz0, &tnz1);
            Plane2(ByteAsBiasedFloat<J>(farx), ByteAsBiasedFloat<J + 1>(farx), adjx, orgx, &tfx0, &tfx1);
            Plane2(ByteAsBiasedFloat<J>(fary), ByteAsBiasedFloat<J + 1>(fary), adjy, orgy, &tfy0, &tfy1);
            Plane2(ByteAsBiasedFloat<J>(farz), ByteAsBiasedFloat<J + 1>(farz), adjz, orgz, &tfz0, &tfz1);
            const float enter0 = fmaxf(fmaxf(tnx0, tny0), fmaxf(tnz0, ray.tmin)), exit0 = fminf(fminf(tfx0, tfy0), fminf(tfz0, ray.tmax));
            const float enter1 = fmaxf(fmaxf(tnx1, tny1), fmaxf(tnz1, ray.tmin)), exit1 = fminf(fminf(tfx1, tfy1), fminf(tfz1, ray.tmax));
            if (enter0 <= exit0) hitmask |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);
            if (enter1 <= exit1) hitmask |= ((child_bits4 >> (8 * J + 8)) & 0xffu) << ((bit_index4 >> (8 * J + 8)) & 0xffu);
        };
        pair(std::integral_constant<int, 0>{});
        pair(std::integral_constant<int, 2>{});
#else
        auto child = [&](auto jc) {
            constexpr int J = decltype(jc)::value;
            const float tnx = fmaf(ByteAsBiasedFloat<J>(nearx) - kBias, adjx, orgx), tfx = fmaf(ByteAsBiasedFloat<J>(farx) - kBias, adjx, orgx);
            const float tny = fmaf(ByteAsBiasedFloat<J>(neary) - kBias, adjy, orgy), tfy = fmaf(ByteAsBiasedFloat<J>(fary) - kBias, adjy, orgy);
            const float tnz = fmaf(ByteAsBiasedFloat<J>(nearz) - kBias, adjz, orgz), tfz = fmaf(ByteAsBiasedFloat<J>(farz) - kBias, adjz, orgz);
            const float t_enter = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, ray.tmin));
            const float t_exit = fminf(fminf(tfx, tfy), fminf(tfz, ray.tmax));
            if (t_enter <= t_exit) hitmask |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);
        };
        child(std::integral_constant<int, 0>{});
        child(std::integral_constant<int, 1>{});
        child(std::integral_constant<int, 2>{});
        child(std::integral_constant<int, 3>{});
#endif
    }
    return hitmask;
}

// Node fetch.  TOP: nodes [0, num_top) come from the CTA's shared-memory copy (StageTopNodes), through a generic pointer so
// that both sources share one load sequence.
template <bool TOP>
__device__ __forceinline__ void LoadWideNode(const WideNode *__restrict__ nodes, const uint4 *top, uint32_t num_top, uint32_t index,
                                             uint4 *n0, uint4 *n1, uint4 *n2, uint4 *n3, uint4 *n4) {
    if (TOP) {
        const uint4 *src = index < num_top ? top + index * kWideNodeVec : reinterpret_cast<const uint4 *>(nodes + index);
        *n0 = src[0], *n1 = src[1], *n2 = src[2], *n3 = src[3], *n4 = src[4];
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(nodes + index);
        *n0 = __ldg(src), *n1 = __ldg(src + 1), *n2 = __ldg(src + 2), *n3 = __ldg(src + 3), *n4 = __ldg(src + 4);
    }
}

// One candidate triangle of a leaf: the Woop test, the alpha test, the hit record.  Returns true when an occlusion ray is done.
template <bool OPACITY>
__device__ __forceinline__ bool WideTestTriangle(const DeviceScene &scene, uint32_t tri, Ray &ray, const RayPre &pre, bool any, const Rng &rng,
                                                 HitRec &hit, bool &found) {
    const float4 *verts = reinterpret_cast<const float4 *>(scene.tri_verts + tri);
    const float4 p0 = __ldg(verts), p1 = __ldg(verts + 1), p2 = __ldg(verts + 2);
    float t, u, v;
    bool inside;
    if (!IntersectTriangleWoop(ray, pre, p0, p1, p2, &t, &u, &v, &inside)) return false;
    if (OPACITY) { // triangle.cpp:115-118: texcoord = Lerp(texcoords, u, v, w)
        const float *tc = &scene.tri_shade[tri].uv[0][0];
        const float w = 1.0f - u - v;
        const V2 uv = {u * __ldg(tc) + v * __ldg(tc + 2) + w * __ldg(tc + 4), u * __ldg(tc + 1) + v * __ldg(tc + 3) + w * __ldg(tc + 5)};
        // keyed by the triangle's index in the scene description, which both tree layouts share
        if (OpacityRejects(scene, __float_as_uint(p0.w), uv, rng, __float_as_uint(p1.w))) return false;
    }
    found = true;
    if (any) return true;
    ray.tmax = t;
    hit.t = t;
    hit.prim = tri | (inside ? kPrimInsideBit : 0u);
    hit.u = u;
    hit.v = v;
    return false;
}

// Persistent-threads traversal with per-lane ray replacement (the loop of TraversePersistent, traverse.cuh) over the wide tree.
// Every lane owns at most one ray and two groups: `ngroup` = (first inner child, hit mask in the top byte | imask) of the node
// it is in, `tgroup` = (first triangle, triangle hit mask).  One iteration = one node fetch + the triangles it produced.
// `tri_min_lanes`: a lane whose node step produced triangles while fewer than this many lanes of the warp have some puts
// them on its stack and goes on with inner nodes when it has any (triangle postponing of the paper), so that the
// triangle test runs with more lanes later.
template <bool MIXED, bool STATS, bool OPACITY, bool TOP, typename Fetch, typename Finish>
__device__ __forceinline__ void TraversePersistentWide(const DeviceScene &scene, const uint4 *top, uint32_t num_top, uint32_t num_rays,
                                                       uint32_t *work_counter, int refill_threshold, int tri_min_lanes, uint2 key,
                                                       Fetch fetch, Finish finish, TraversalCounters *counters, uint32_t *rays_traced) {
    uint2 stack[kWideStack];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    bool has = false, exhausted = false, found = false, any = false;
    uint32_t index = 0;
    Ray ray;
    ray.o = ray.d = mk3(0.0f);
    ray.tmin = ray.tmax = 0.0f;
    RayPre pre = Precompute(ray);
    WideRay wr = MakeWideRay(pre);
    HitRec hit;
    hit.t = 0.0f, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
    Rng rng(0, 0, 0, key, kRngDomainClosest);

    for (;;) {
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
                wr = MakeWideRay(pre);
                hit.t = ray.tmax, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
                found = false;
                has = true;
                sp = 0;
                tgroup = make_uint2(0u, 0u);
                ngroup = make_uint2(0u, scene.num_wide_nodes ? 0x80000000u : 0u); // "child 0 of nothing, hit": the root
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
                            ngroup.y = 0u;
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
            // ---- one node: closest unvisited hit child of the current group ----
            if (ngroup.y > 0x00ffffffu) {
                const uint32_t hits_imask = ngroup.y;
                const uint32_t bit = 31u - __clz(hits_imask);
                const uint32_t child_base = ngroup.x;
                ngroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup; // siblings still to visit
                const uint32_t slot = (bit - 24u) ^ (wr.octinv4 & 0xffu);
                const uint32_t node_index = child_base + __popc(hits_imask & ~(0xffffffffu << slot));
                uint4 n0, n1, n2, n3, n4;
                LoadWideNode<TOP>(scene.wide_nodes, top, num_top, node_index, &n0, &n1, &n2, &n3, &n4);
                if (STATS) counters[any].nodes += 1;
                const uint32_t hitmask = WideNodeHits(n0, n1, n2, n3, n4, ray, wr);
                ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
                tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
            } else { // a postponed triangle group came off the stack
                tgroup = ngroup;
                ngroup = make_uint2(0u, 0u);
            }
            // ---- triangles of the hit leaves ----
            if (tgroup.y != 0u && ngroup.y > 0x00ffffffu && __popc(__activemask()) < tri_min_lanes) {
                stack[sp++] = tgroup; // too few lanes hold triangles: keep walking inner nodes, test these later
                tgroup.y = 0u;
            }
            while (tgroup.y != 0u) {
                const uint32_t k = 31u - __clz(tgroup.y);
                tgroup.y &= ~(1u << k);
                if (STATS) ++counters[any].prims;
                if (WideTestTriangle<OPACITY>(scene, tgroup.x + k, ray, pre, any, rng, hit, found)) { // occluded
                    tgroup.y = 0u, ngroup.y = 0u, sp = 0;
                }
            }
            // ---- next group ----
            if (ngroup.y <= 0x00ffffffu) {
                if (sp > 0) {
                    ngroup = stack[--sp];
                } else {
                    finish(index, hit, found, any);
                    has = false;
                }
            }
        }
    }
}

// One ray per lane to the end (k_tail, debug ray entry): same node / triangle tests and visiting order as
// TraversePersistentWide without postponing.
__device__ __forceinline__ bool TraverseSingleWide(const DeviceScene &scene, Ray ray, bool any, bool opacity, Rng rng, HitRec *hit_out,
                                                   bool stats, TraversalCounters *counters) {
    uint2 stack[kWideStack];
    int sp = 0;
    const RayPre pre = Precompute(ray);
    const WideRay wr = MakeWideRay(pre);
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
    uint2 ngroup = make_uint2(0u, scene.num_wide_nodes ? 0x80000000u : 0u);
    for (;;) {
        uint2 tgroup = make_uint2(0u, 0u);
        if (ngroup.y > 0x00ffffffu) {
            const uint32_t hits_imask = ngroup.y;
            const uint32_t bit = 31u - __clz(hits_imask);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            const uint32_t slot = (bit - 24u) ^ (wr.octinv4 & 0xffu);
            const uint32_t node_index = child_base + __popc(hits_imask & ~(0xffffffffu << slot));
            uint4 n0, n1, n2, n3, n4;
            LoadWideNode<false>(scene.wide_nodes, nullptr, 0u, node_index, &n0, &n1, &n2, &n3, &n4);
            if (stats) counters->nodes += 1;
            const uint32_t hitmask = WideNodeHits(n0, n1, n2, n3, n4, ray, wr);
            ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
        }
        while (tgroup.y != 0u) {
            const uint32_t k = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << k);
            if (stats) ++counters->prims;
            bool done;
            if (opacity)
                done = WideTestTriangle<true>(scene, tgroup.x + k, ray, pre, any, rng, hit, found);
            else
                done = WideTestTriangle<false>(scene, tgroup.x + k, ray, pre, any, rng, hit, found);
            if (done) return true;
        }
        if (ngroup.y <= 0x00ffffffu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    *hit_out = hit;
    return found;
}

} // namespace b200pt
