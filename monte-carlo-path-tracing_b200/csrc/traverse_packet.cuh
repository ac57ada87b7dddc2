// traverse_packet.cuh — warp packets: 32 rays that start close together and point the same way walk the binary tree TOGETHER.
//
// Same job and same tests as TraversePersistent (traverse.cuh): TLAS::Intersect / IntersectAny (src/rtcore/accel/tlas.cpp:13-76),
// AABB::Intersect (aabb.cpp:29-48), the Woop triangle test (triangle.cpp:23-87).  Where it differs is the control flow.  The camera
// rays of ONE pixel (32 consecutive sample slots: renderer.cpp:66-75 varies only the sub-pixel offset) and the NEE rays that
// neighbouring hits send towards one light visit almost the same nodes, so the warp keeps ONE node pointer and ONE stack:
// every lane tests its own ray against the two child boxes of the warp's node, the warp descends into a child when ANY lane
// hits it (near child by majority vote), and at a leaf every lane tests the triangles against its own [tmin, tmax].  Control
// flow is warp-uniform — no divergence, no per-lane stack in local memory, node fetches are one address per warp — and all 32
// lanes do useful work as long as the rays stay together (per-lane ray replacement runs at 14-22 active lanes per instruction
// on the same rays, profiles/).  The price is the union of the 32 paths instead of each ray's own: only for coherent sets.
//
// Each lane runs exactly the per-ray box and triangle arithmetic of traverse.cuh on its own ray, so it finds the same closest
// hit (a tie between two triangles at the same t may pick the other one: the order of the tests differs, as between trees).
#pragma once
#include "traverse.cuh"

namespace b200pt {

constexpr int kPacketStack = 64; // = kStackSize: the builder rejects deeper trees

// `stack`: kPacketStack ints of shared memory private to this warp.  ANY: occlusion rays (a lane is done at its first hit).
// fetch(index, &ray, &ctr) -> bool builds ray `index`; finish(index, hit, found) consumes the result.
template <bool ANY, bool STATS, bool OPACITY, typename Fetch, typename Finish>
__device__ __forceinline__ void TraversePacket(const DeviceScene &scene, uint32_t begin, uint32_t num_rays, uint32_t *work_counter, uint2 key,
                                               int *stack, Fetch fetch, Finish finish, TraversalCounters *counters, uint32_t *rays_traced) {
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= num_rays) break;
        const uint32_t index = begin + base + lane;
        Ray ray;
        ray.o = ray.d = mk3(0.0f);
        ray.tmin = 0.0f, ray.tmax = -1.0f;
        uint3 ctr = make_uint3(0, 0, 0);
        const bool valid = base + lane < num_rays && fetch(index, &ray, &ctr);
        bool alive = valid, found = false;
        const RayPre pre = Precompute(ray);
        const Rng rng(ctr.x, ctr.y, ctr.z, key, ANY ? kRngDomainShadow : kRngDomainClosest);
        HitRec hit;
        hit.t = ray.tmax, hit.prim = kPrimMiss, hit.u = hit.v = 0.0f;
        if (valid) ++*rays_traced;
        // Analytic primitives (spheres, disks, cylinders) are few: tested linearly up front.
        for (uint32_t i = 0; i < scene.num_analytic; ++i) {
            const AnalyticPrim &p = scene.analytic[i];
            if (!alive) continue;
            if (STATS) ++counters->nodes;
            if (!IntersectBox(p.bmin, p.bmax, ray, pre)) continue;
            if (STATS) ++counters->prims;
            float t;
            V2 uv = {0.0f, 0.0f};
            if (IntersectAnalytic(p, ray, &t, OPACITY ? &uv : nullptr)) {
                if (OPACITY && OpacityRejects(scene, p.inst, uv, rng, kPrimAnalyticBit | i)) continue;
                found = true;
                if (ANY) {
                    alive = false;
                    continue;
                }
                ray.tmax = t;
                hit.t = t;
                hit.prim = kPrimAnalyticBit | i;
            }
        }
        int sp = 0;
        int cur = (scene.num_nodes && __any_sync(0xffffffffu, alive)) ? 0 : kSentinel;
        while (cur != kSentinel) {
            if (cur >= 0) {
                float4 n0, n1, nz;
                int child0, child1;
                LoadNode<false>(scene.nodes, nullptr, 0, cur, &n0, &n1, &nz, &child0, &child1);
                if (STATS && alive) counters->nodes += 2;
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
                const bool hit0 = alive && c0min <= c0max, hit1 = alive && c1min <= c1max;
                const unsigned m0 = __ballot_sync(0xffffffffu, hit0), m1 = __ballot_sync(0xffffffffu, hit1);
                if ((m0 | m1) == 0u) {
                    cur = sp > 0 ? stack[--sp] : kSentinel;
                } else if (m0 != 0u && m1 != 0u) {
                    // near child first, as most of the lanes that hit something see it
                    const unsigned first1 = __ballot_sync(0xffffffffu, hit1 && (!hit0 || c1min < c0min));
                    const bool swap = 2 * __popc(first1) > __popc(m0 | m1);
                    // one writer, fenced on both sides: the slot may still be being read (a pop) by the slower lanes of the warp
                    __syncwarp();
                    if (lane == 0) stack[sp] = swap ? child0 : child1;
                    ++sp;
                    __syncwarp();
                    cur = swap ? child1 : child0;
                } else {
                    cur = m0 != 0u ? child0 : child1;
                }
            } else {
                const uint32_t leaf = static_cast<uint32_t>(~cur);
                const uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
                const float4 *verts = reinterpret_cast<const float4 *>(scene.tri_verts + first);
                for (uint32_t j = 0; j < count; ++j) {
                    const float4 p0 = __ldg(verts + 3 * j), p1 = __ldg(verts + 3 * j + 1), p2 = __ldg(verts + 3 * j + 2);
                    if (!alive) continue;
                    if (STATS) ++counters->prims;
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
                        if (ANY) {
                            alive = false;
                            continue;
                        }
                        ray.tmax = t;
                        hit.t = t;
                        hit.prim = (first + j) | (inside ? kPrimInsideBit : 0u);
                        hit.u = u;
                        hit.v = v;
                    }
                }
                cur = sp > 0 ? stack[--sp] : kSentinel;
                if (ANY && !__any_sync(0xffffffffu, alive)) cur = kSentinel; // every ray of the packet is occluded
            }
        }
        if (valid) finish(index, hit, found);
    }
}

} // namespace b200pt
