// wide_bvh_check.cpp — CPU check of the compressed 8-wide BVH builder (csrc/bvh_wide.cpp) and of the traversal scheme the
// CUDA kernels implement (csrc/traverse_wide.cuh), against brute force.  Compiled and run by tests/test_wide_bvh.py.
//
// What it proves without a GPU: every triangle is referenced exactly once; every quantised child box contains the exact
// box of its subtree; the (base index, hit mask) group walk with octant-ordered slots visits every node a ray can hit, so
// the closest hit equals the brute-force one BIT FOR BIT (same triangle test on both sides).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "bvh_wide.hpp"

using namespace b200pt;

namespace {

struct Tri {
    float p[3][3];
};

uint32_t g_state = 12345u;
float Rand() {
    g_state = g_state * 1664525u + 1013904223u;
    return (g_state >> 8) * (1.0f / 16777216.0f);
}

// Median-split binary tree with leaves of <= 3 triangles (the product uses binned SAH; any binary tree must collapse correctly).
int32_t BuildBinary(std::vector<Bvh2Node> &nodes, const std::vector<Tri> &tris, std::vector<uint32_t> &order, uint32_t begin, uint32_t end) {
    const int32_t id = static_cast<int32_t>(nodes.size());
    nodes.emplace_back();
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t i = begin; i < end; ++i)
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) lo[k] = fminf(lo[k], tris[order[i]].p[v][k]), hi[k] = fmaxf(hi[k], tris[order[i]].p[v][k]);
    memcpy(nodes[id].lo, lo, 12), memcpy(nodes[id].hi, hi, 12);
    const uint32_t n = end - begin;
    if (n <= 1 + g_state % 3) { // leaves of 1..3 triangles
        nodes[id].first = begin, nodes[id].count = n;
        g_state = g_state * 1664525u + 1013904223u;
        return id;
    }
    int axis = 0;
    for (int k = 1; k < 3; ++k)
        if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
    const uint32_t mid = begin + n / 2;
    std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](uint32_t a, uint32_t b) {
        return tris[a].p[0][axis] + tris[a].p[1][axis] + tris[a].p[2][axis] < tris[b].p[0][axis] + tris[b].p[1][axis] + tris[b].p[2][axis];
    });
    const int32_t l = BuildBinary(nodes, tris, order, begin, mid), r = BuildBinary(nodes, tris, order, mid, end);
    nodes[id].left = l, nodes[id].right = r;
    return id;
}

// Moeller-Trumbore (any deterministic test does: both sides of the comparison use this one).
bool HitTri(const Tri &t, const float o[3], const float d[3], float tmax, float *t_out) {
    float e1[3], e2[3], pv[3], tv[3], qv[3];
    for (int k = 0; k < 3; ++k) e1[k] = t.p[1][k] - t.p[0][k], e2[k] = t.p[2][k] - t.p[0][k];
    pv[0] = d[1] * e2[2] - d[2] * e2[1], pv[1] = d[2] * e2[0] - d[0] * e2[2], pv[2] = d[0] * e2[1] - d[1] * e2[0];
    const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    for (int k = 0; k < 3; ++k) tv[k] = o[k] - t.p[0][k];
    const float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    qv[0] = tv[1] * e1[2] - tv[2] * e1[1], qv[1] = tv[2] * e1[0] - tv[0] * e1[2], qv[2] = tv[0] * e1[1] - tv[1] * e1[0];
    const float v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float tt = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
    if (tt < 1e-4f || tt > tmax) return false;
    *t_out = tt;
    return true;
}

// The node test of traverse_wide.cuh (WideNodeHits) in plain C++.
uint32_t NodeHits(const WideNode &n, const float o[3], const float idir[3], float tmin, float tmax, uint32_t r) {
    float adj[3], org[3];
    for (int k = 0; k < 3; ++k) {
        const uint32_t bits = static_cast<uint32_t>(n.e[k]) << 23;
        float cell;
        memcpy(&cell, &bits, 4);
        adj[k] = cell * idir[k];
        org[k] = (n.origin[k] - o[k]) * idir[k];
    }
    const uint8_t *qlo[3] = {n.qlo_x, n.qlo_y, n.qlo_z}, *qhi[3] = {n.qhi_x, n.qhi_y, n.qhi_z};
    uint32_t mask = 0;
    for (int s = 0; s < 8; ++s) {
        const uint32_t meta = n.meta[s];
        if (meta == 0) continue;
        float enter = tmin, exit = tmax;
        for (int k = 0; k < 3; ++k) {
            const bool neg = idir[k] < 0.0f;
            const float tn = fmaf(static_cast<float>(neg ? qhi[k][s] : qlo[k][s]), adj[k], org[k]);
            const float tf = fmaf(static_cast<float>(neg ? qlo[k][s] : qhi[k][s]), adj[k], org[k]);
            enter = fmaxf(enter, tn), exit = fminf(exit, tf);
        }
        if (!(enter <= exit)) continue;
        const bool inner = (meta & 0x18u) == 0x18u; // low five bits >= 24
        const uint32_t bit_index = (inner ? (meta ^ r) : meta) & 0x1fu, bits = meta >> 5;
        mask |= bits << bit_index;
    }
    return mask;
}

int Fail(const char *what) {
    printf("FAIL: %s\n", what);
    return 1;
}

// Exact box of the subtree under a wide node (recursing through the wide tree itself).
void SubtreeBox(const std::vector<WideNode> &wide, const std::vector<Tri> &tris, const std::vector<uint32_t> &order, uint32_t index, float lo[3],
                float hi[3], std::vector<uint32_t> *seen, int *errors) {
    const WideNode &n = wide[index];
    uint32_t inner_rank = 0;
    for (int s = 0; s < 8; ++s) {
        const uint32_t meta = n.meta[s];
        if (meta == 0) continue;
        float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
        if ((meta & 0x18u) == 0x18u && (meta >> 5) == 1u) {
            if (!((n.imask >> s) & 1u) || (meta & 7u) != static_cast<uint32_t>(s)) ++*errors;
            SubtreeBox(wide, tris, order, n.child_base + inner_rank++, clo, chi, seen, errors);
        } else {
            if ((n.imask >> s) & 1u) ++*errors;
            const uint32_t count = __builtin_popcount(meta >> 5), offset = meta & 0x1fu;
            if (count < 1 || count > 3 || offset + count > 24) ++*errors;
            for (uint32_t j = 0; j < count; ++j) {
                const uint32_t tri = order[n.tri_base + offset + j];
                ++(*seen)[tri];
                for (int v = 0; v < 3; ++v)
                    for (int k = 0; k < 3; ++k) clo[k] = fminf(clo[k], tris[tri].p[v][k]), chi[k] = fmaxf(chi[k], tris[tri].p[v][k]);
            }
        }
        const uint8_t *qlo[3] = {n.qlo_x, n.qlo_y, n.qlo_z}, *qhi[3] = {n.qhi_x, n.qhi_y, n.qhi_z};
        for (int k = 0; k < 3; ++k) {
            const double cell = ldexp(1.0, static_cast<int>(n.e[k]) - 127);
            if (static_cast<double>(n.origin[k]) + qlo[k][s] * cell > clo[k] || static_cast<double>(n.origin[k]) + qhi[k][s] * cell < chi[k]) ++*errors;
            lo[k] = fminf(lo[k], clo[k]), hi[k] = fmaxf(hi[k], chi[k]);
        }
    }
}

} // namespace

int main(int argc, char **argv) {
    const uint32_t num_tris = argc > 1 ? static_cast<uint32_t>(atoi(argv[1])) : 20000u;
    const uint32_t num_rays = argc > 2 ? static_cast<uint32_t>(atoi(argv[2])) : 4000u;
    g_state = argc > 3 ? static_cast<uint32_t>(atoi(argv[3])) : 12345u;
    std::vector<Tri> tris(num_tris);
    for (Tri &t : tris) {
        // clustered sizes: mostly small triangles, a few large ones, some axis-aligned (flat boxes)
        const float c[3] = {Rand() * 10.0f - 5.0f, Rand() * 10.0f - 5.0f, Rand() * 10.0f - 5.0f};
        const float size = Rand() < 0.02f ? 3.0f : 0.15f;
        const int flat = Rand() < 0.1f ? static_cast<int>(Rand() * 3.0f) : -1;
        for (int v = 0; v < 3; ++v)
            for (int k = 0; k < 3; ++k) t.p[v][k] = c[k] + (k == flat ? 0.0f : (Rand() - 0.5f) * size);
    }
    std::vector<uint32_t> order(num_tris);
    std::iota(order.begin(), order.end(), 0u);
    std::vector<Bvh2Node> binary;
    binary.reserve(2 * num_tris);
    const int32_t root = num_tris ? BuildBinary(binary, tris, order, 0, num_tris) : -1;

    std::vector<WideNode> wide;
    WideBuildInfo info;
    std::string error;
    if (!BuildWideBvh(binary, root, 64, &order, &wide, &info, &error)) return Fail(error.c_str());
    if (num_tris == 0) {
        printf("OK empty\n");
        return wide.empty() ? 0 : Fail("empty scene produced nodes");
    }

    // ---- structure ----
    std::vector<uint32_t> seen(num_tris, 0);
    int errors = 0;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    SubtreeBox(wide, tris, order, 0, lo, hi, &seen, &errors);
    if (errors) return Fail("node invariants (imask / meta / box containment)");
    for (uint32_t c : seen)
        if (c != 1) return Fail("a triangle is referenced zero or several times");

    // ---- traversal vs brute force ----
    uint64_t nodes_visited = 0, tri_tests = 0, hits = 0;
    for (uint32_t i = 0; i < num_rays; ++i) {
        float o[3], d[3], idir[3];
        for (int k = 0; k < 3; ++k) o[k] = Rand() * 14.0f - 7.0f, d[k] = Rand() * 2.0f - 1.0f;
        if (i % 7 == 0) d[i % 3] = 0.0f; // axis-parallel rays take the reference's 1e-4 substitute (ray.cpp:21-22)
        const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (len == 0.0f) continue;
        for (int k = 0; k < 3; ++k) d[k] /= len, idir[k] = 1.0f / (d[k] != 0.0f ? d[k] : 1e-4f);
        float best = 3.0e38f;
        int best_tri = -1;
        for (uint32_t t = 0; t < num_tris; ++t) {
            float tt;
            if (HitTri(tris[t], o, d, best, &tt)) best = tt, best_tri = static_cast<int>(t);
        }
        // the group walk of TraverseSingleWide
        const uint32_t r = (idir[0] >= 0.0f ? 1u : 0u) | (idir[1] >= 0.0f ? 2u : 0u) | (idir[2] >= 0.0f ? 4u : 0u);
        struct Group {
            uint32_t base, mask;
        };
        Group stack[kWideStackEntries];
        int sp = 0;
        Group ng = {0u, 0x80000000u};
        float tmax = 3.0e38f;
        int found_tri = -1;
        for (;;) {
            Group tg = {0u, 0u};
            if (ng.mask > 0x00ffffffu) {
                const uint32_t hits_imask = ng.mask, bit = 31u - __builtin_clz(hits_imask), child_base = ng.base;
                ng.mask &= ~(1u << bit);
                if (ng.mask > 0x00ffffffu) {
                    if (sp >= static_cast<int>(kWideStackEntries)) return Fail("stack overflow");
                    stack[sp++] = ng;
                }
                const uint32_t slot = (bit - 24u) ^ r;
                const uint32_t index = child_base + __builtin_popcount(hits_imask & ~(0xffffffffu << slot));
                if (index >= wide.size()) return Fail("child index out of range");
                ++nodes_visited;
                const uint32_t mask = NodeHits(wide[index], o, idir, 1e-4f, tmax, r);
                ng = {wide[index].child_base, (mask & 0xff000000u) | wide[index].imask};
                tg = {wide[index].tri_base, mask & 0x00ffffffu};
            }
            while (tg.mask) {
                const uint32_t k = 31u - __builtin_clz(tg.mask);
                tg.mask &= ~(1u << k);
                ++tri_tests;
                float tt;
                // same acceptance rule as the kernels: t <= tmax replaces (ties go to the triangle tested last)
                if (HitTri(tris[order[tg.base + k]], o, d, tmax, &tt)) tmax = tt, found_tri = static_cast<int>(order[tg.base + k]);
            }
            if (ng.mask <= 0x00ffffffu) {
                if (sp == 0) break;
                ng = stack[--sp];
            }
        }
        if ((best_tri < 0) != (found_tri < 0)) return Fail("hit / miss disagrees with brute force");
        if (best_tri >= 0) {
            ++hits;
            if (memcmp(&best, &tmax, 4) != 0) return Fail("closest distance differs from brute force");
        }
    }
    printf("OK tris %u wide_nodes %zu depth %u inner_slots %u leaf_slots %u rays %u hits %llu nodes/ray %.2f tris/ray %.2f\n", num_tris, wide.size(),
           info.depth, info.inner_slots, info.leaf_slots, num_rays, (unsigned long long)hits, double(nodes_visited) / num_rays, double(tri_tests) / num_rays);
    return 0;
}
