// sah_bvh_check.cpp — test program (no GPU): the DEFAULT acceleration structure of the product, built by the product's own host
// builder (BuildHostScene, linked from libb200pt.so: parallel binned SAH + flatten, csrc/scene_build.cpp) for a scene pack, checked
// on the CPU:
//   1. structure: every triangle is in exactly one leaf, leaves hold 1..8 triangles, every inner node is referenced once, child
//      boxes contain everything below them, the tree is shallower than the traversal stack;
//   2. closest hits: a plain two-children-per-node walk of the flattened nodes (the scheme of traverse.cuh, restated here)
//      finds, for random rays, the same distance as brute force over all triangles with the same triangle test;
//   3. the two box covers of the camera-ray visibility pre-pass contain every triangle.
// Usage: sah_bvh_check scene.b200scene num_rays seed   -> prints "OK ..." or the first violation.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "b200pt.h"
#include "host_scene.hpp"

using namespace b200pt;

namespace {

struct Ray {
    float o[3], d[3], idir[3];
};

uint32_t g_rng = 1u;
float Rand() {
    g_rng = g_rng * 1664525u + 1013904223u;
    return (g_rng >> 8) * (1.0f / 16777216.0f);
}

bool HitTri(const TriVerts &t, const Ray &r, float tmax, float *t_out) { // Moeller-Trumbore; only has to be the SAME test on both sides
    const float e1[3] = {t.v1.x - t.v0.x, t.v1.y - t.v0.y, t.v1.z - t.v0.z}, e2[3] = {t.v2.x - t.v0.x, t.v2.y - t.v0.y, t.v2.z - t.v0.z};
    const float pv[3] = {r.d[1] * e2[2] - r.d[2] * e2[1], r.d[2] * e2[0] - r.d[0] * e2[2], r.d[0] * e2[1] - r.d[1] * e2[0]};
    const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (det == 0.0f) return false;
    const float inv = 1.0f / det;
    const float tv[3] = {r.o[0] - t.v0.x, r.o[1] - t.v0.y, r.o[2] - t.v0.z};
    const float u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
    const float v = (r.d[0] * qv[0] + r.d[1] * qv[1] + r.d[2] * qv[2]) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float tt = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
    if (tt < 1e-4f || tt > tmax) return false;
    *t_out = tt;
    return true;
}

struct Box3 {
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    void Grow(const F4 &p) {
        const float v[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; ++k) lo[k] = fminf(lo[k], v[k]), hi[k] = fmaxf(hi[k], v[k]);
    }
    void Grow(const Box3 &b) {
        for (int k = 0; k < 3; ++k) lo[k] = fminf(lo[k], b.lo[k]), hi[k] = fmaxf(hi[k], b.hi[k]);
    }
    bool Inside(const Box3 &outer) const {
        for (int k = 0; k < 3; ++k)
            if (lo[k] < outer.lo[k] || hi[k] > outer.hi[k]) return false;
        return true;
    }
};

Box3 ChildBox(const BvhNode &n, int which) {
    Box3 b;
    if (which == 0)
        b.lo[0] = n.c0xy.x, b.hi[0] = n.c0xy.y, b.lo[1] = n.c0xy.z, b.hi[1] = n.c0xy.w, b.lo[2] = n.cz.x, b.hi[2] = n.cz.y;
    else
        b.lo[0] = n.c1xy.x, b.hi[0] = n.c1xy.y, b.lo[1] = n.c1xy.z, b.hi[1] = n.c1xy.w, b.lo[2] = n.cz.z, b.hi[2] = n.cz.w;
    return b;
}

const char *g_error = nullptr;
std::vector<uint32_t> g_seen_tri, g_seen_node;
uint32_t g_depth = 0;

// bounds of everything below `child`; checks containment on the way up
Box3 Check(const HostScene &hs, int32_t child, uint32_t depth) {
    Box3 box;
    if (child < 0) {
        const uint32_t leaf = static_cast<uint32_t>(~child), first = leaf >> 3, count = (leaf & 7u) + 1u;
        if (first + count > hs.tri_verts.size()) {
            if (hs.tri_verts.empty() || first != 0) g_error = "leaf range outside the triangle array"; // (the padding leaf of a one-leaf tree points at triangle 0)
            return box;
        }
        for (uint32_t j = 0; j < count; ++j) {
            ++g_seen_tri[first + j];
            const TriVerts &t = hs.tri_verts[first + j];
            box.Grow(t.v0), box.Grow(t.v1), box.Grow(t.v2);
        }
        return box;
    }
    if (static_cast<size_t>(child) >= hs.nodes.size()) {
        g_error = "child index outside the node array";
        return box;
    }
    if (++g_seen_node[child] > 1) {
        g_error = "inner node referenced twice";
        return box;
    }
    if (depth > g_depth) g_depth = depth;
    const BvhNode &n = hs.nodes[child];
    for (int which = 0; which < 2; ++which) {
        const Box3 stored = ChildBox(n, which);
        if (stored.hi[0] < stored.lo[0]) continue; // the inverted box of a one-leaf tree's padding child
        const Box3 below = Check(hs, which == 0 ? n.child0 : n.child1, depth + 1);
        if (!below.Inside(stored)) g_error = "a child box does not contain its subtree";
        box.Grow(stored);
    }
    return box;
}

float TraceBinary(const HostScene &hs, const Ray &r) {
    int stack[64], sp = 0, cur = hs.nodes.empty() ? 0x7fffffff : 0;
    float tmax = 3.0e38f;
    while (cur != 0x7fffffff) {
        if (cur >= 0) {
            const BvhNode &n = hs.nodes[cur];
            auto slab = [&](const Box3 &b, float *tn) {
                float a = 1e-4f, e = tmax;
                for (int k = 0; k < 3; ++k) {
                    const float t0 = (b.lo[k] - r.o[k]) * r.idir[k], t1 = (b.hi[k] - r.o[k]) * r.idir[k];
                    a = fmaxf(a, fminf(t0, t1)), e = fminf(e, fmaxf(t0, t1));
                }
                *tn = a;
                return a <= e * 1.0000004f; // one ulp of slack: the box test must never lose a hit the brute force finds
            };
            float t0, t1;
            const bool h0 = slab(ChildBox(n, 0), &t0), h1 = slab(ChildBox(n, 1), &t1);
            if (!h0 && !h1) {
                cur = sp > 0 ? stack[--sp] : 0x7fffffff;
            } else if (h0 && h1) {
                const bool swap = t1 < t0;
                stack[sp++] = swap ? n.child0 : n.child1;
                cur = swap ? n.child1 : n.child0;
            } else {
                cur = h0 ? n.child0 : n.child1;
            }
        } else {
            const uint32_t leaf = static_cast<uint32_t>(~cur), first = leaf >> 3, count = (leaf & 7u) + 1u;
            cur = sp > 0 ? stack[--sp] : 0x7fffffff;
            for (uint32_t j = 0; j < count && first + j < hs.tri_verts.size(); ++j) {
                float tt;
                if (HitTri(hs.tri_verts[first + j], r, tmax, &tt)) tmax = tt;
            }
        }
    }
    return tmax;
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int num_rays = atoi(argv[2]);
    g_rng = static_cast<uint32_t>(atoi(argv[3])) * 2654435761u + 1u;
    b200pt_scene *scene = nullptr;
    if (b200pt_scene_load(argv[1], &scene) != B200PT_OK) {
        printf("FAIL cannot load %s: %s\n", argv[1], b200pt_last_error(nullptr));
        return 1;
    }
    HostScene hs;
    std::string error;
    if (!BuildHostScene(*b200pt_scene_get_desc(scene), 0, false, false, &hs, &error)) {
        printf("FAIL BuildHostScene: %s\n", error.c_str());
        return 1;
    }
    const size_t nt = hs.tri_verts.size();
    g_seen_tri.assign(nt, 0), g_seen_node.assign(hs.nodes.size(), 0);
    if (!hs.nodes.empty()) Check(hs, 0, 1);
    if (g_error == nullptr)
        for (size_t i = 0; i < nt; ++i)
            if (g_seen_tri[i] != 1) {
                g_error = "a triangle is in no leaf or in several";
                break;
            }
    if (g_error == nullptr)
        for (size_t i = 0; i < hs.nodes.size(); ++i)
            if (g_seen_node[i] != 1) {
                g_error = "an inner node is unreachable";
                break;
            }
    if (g_error == nullptr && g_depth >= 64) g_error = "tree deeper than the traversal stack";
    if (g_error != nullptr) {
        printf("FAIL structure: %s\n", g_error);
        return 1;
    }
    // 3. covers of the visibility pre-pass (k_cull_tiles): every triangle lies inside a coarse box, and inside a fine box filed
    //    under that coarse box — otherwise a pixel that sees it could be dropped
    {
        const size_t ncoarse = hs.cull_boxes.size() / 6 - hs.analytic.size(), nfine = hs.fine_cull_boxes.size() / 6;
        if (hs.cull_fine_begin.size() != hs.cull_boxes.size() / 6 + 1 || hs.cull_fine_begin.back() != nfine) {
            printf("FAIL cull covers: fine ranges do not tile the fine boxes\n");
            return 1;
        }
        auto inside = [](const Box3 &b, const float *q) {
            for (int k = 0; k < 3; ++k)
                if (b.lo[k] < q[k] || b.hi[k] > q[k + 3]) return false;
            return true;
        };
        for (size_t i = 0; i < nt; ++i) {
            Box3 tb;
            tb.Grow(hs.tri_verts[i].v0), tb.Grow(hs.tri_verts[i].v1), tb.Grow(hs.tri_verts[i].v2);
            bool covered = false;
            for (size_t cidx = 0; cidx < ncoarse && !covered; ++cidx) {
                if (!inside(tb, &hs.cull_boxes[6 * cidx])) continue;
                for (uint32_t f = hs.cull_fine_begin[cidx]; f < hs.cull_fine_begin[cidx + 1] && !covered; ++f) covered = inside(tb, &hs.fine_cull_boxes[6 * f]);
            }
            if (!covered) {
                printf("FAIL cull covers: triangle %zu is in no (coarse, fine) box pair\n", i);
                return 1;
            }
        }
    }
    // rays from around the scene towards points inside its bounds
    float c[3], ext[3];
    for (int k = 0; k < 3; ++k) c[k] = 0.5f * (hs.scene_bmin[k] + hs.scene_bmax[k]), ext[k] = 0.5f * (hs.scene_bmax[k] - hs.scene_bmin[k]) + 1e-3f;
    int hits = 0;
    for (int i = 0; i < num_rays; ++i) {
        Ray r;
        float len = 0.0f;
        for (int k = 0; k < 3; ++k) {
            r.o[k] = c[k] + (2.0f * Rand() - 1.0f) * ext[k] * ((i & 1) ? 2.5f : 0.9f); // outside and inside the bounds
            const float target = c[k] + (2.0f * Rand() - 1.0f) * ext[k];
            r.d[k] = target - r.o[k];
            len += r.d[k] * r.d[k];
        }
        len = sqrtf(len);
        for (int k = 0; k < 3; ++k) r.d[k] /= len, r.idir[k] = 1.0f / (r.d[k] != 0.0f ? r.d[k] : 1e-4f);
        float brute = 3.0e38f, tt;
        for (size_t j = 0; j < nt; ++j)
            if (HitTri(hs.tri_verts[j], r, brute, &tt)) brute = tt;
        const float walked = TraceBinary(hs, r);
        if (walked != brute) {
            printf("FAIL ray %d: tree walk %.9g, brute force %.9g\n", i, walked, brute);
            return 1;
        }
        hits += brute < 3.0e38f;
    }
    printf("OK %zu triangles, %zu nodes, depth %u, %d rays (%d hits)\n", nt, hs.nodes.size(), g_depth, num_rays, hits);
    b200pt_scene_free(scene);
    return 0;
}
