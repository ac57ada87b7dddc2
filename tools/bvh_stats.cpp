// bvh_stats.cpp — development tool (no GPU needed): builds the product's acceleration structures for a scene pack with the
// product's own builder (BuildHostScene, linked from libb200pt.so) and walks them on the CPU with the traversal schemes
// of traverse.cuh / traverse_wide.cuh, counting node fetches and triangle tests per ray for camera rays and for one
// generation of diffuse bounce rays.  Used to judge builder changes (collapse strategy, leaf sizes) before spending GPU time.
//
//   g++ -std=c++17 -O2 -I include -I monte-carlo-path-tracing_b200/csrc tools/bvh_stats.cpp -L monte-carlo-path-tracing_b200 -lb200pt \
//       -Wl,-rpath,$PWD/monte-carlo-path-tracing_b200 -o /tmp/bvh_stats && /tmp/bvh_stats scenes/dragon.b200scene 256
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "b200pt.h"
#include "bvh_wide.hpp"
#include "host_scene.hpp"

using namespace b200pt;

namespace {

struct Ray {
    float o[3], d[3], idir[3];
};

uint32_t g_rng = 99991u;
float Rand() {
    g_rng = g_rng * 1664525u + 1013904223u;
    return (g_rng >> 8) * (1.0f / 16777216.0f);
}

void Finish(Ray *r) {
    const float len = sqrtf(r->d[0] * r->d[0] + r->d[1] * r->d[1] + r->d[2] * r->d[2]);
    for (int k = 0; k < 3; ++k) r->d[k] /= len, r->idir[k] = 1.0f / (r->d[k] != 0.0f ? r->d[k] : 1e-4f);
}

bool HitTri(const TriVerts &t, const Ray &r, float tmax, float *t_out) {
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

struct Counts {
    double nodes = 0, tris = 0, rays = 0, hits = 0, max_stack = 0;
};

// traverse.cuh: two child boxes per fetch, near child first.
int TraceBinary(const HostScene &hs, const Ray &r, float *t_hit, Counts *c) {
    int stack[64], sp = 0, cur = hs.nodes.empty() ? 0x7fffffff : 0, found = -1;
    float tmax = 3.0e38f;
    while (cur != 0x7fffffff) {
        if (cur >= 0) {
            const BvhNode &n = hs.nodes[cur];
            c->nodes += 1;
            auto slab = [&](float lox, float hix, float loy, float hiy, float loz, float hiz, float *tn) {
                const float tx0 = (lox - r.o[0]) * r.idir[0], tx1 = (hix - r.o[0]) * r.idir[0];
                const float ty0 = (loy - r.o[1]) * r.idir[1], ty1 = (hiy - r.o[1]) * r.idir[1];
                const float tz0 = (loz - r.o[2]) * r.idir[2], tz1 = (hiz - r.o[2]) * r.idir[2];
                const float a = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 1e-4f));
                const float b = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tmax));
                *tn = a;
                return a <= b;
            };
            float t0, t1;
            const bool h0 = slab(n.c0xy.x, n.c0xy.y, n.c0xy.z, n.c0xy.w, n.cz.x, n.cz.y, &t0);
            const bool h1 = slab(n.c1xy.x, n.c1xy.y, n.c1xy.z, n.c1xy.w, n.cz.z, n.cz.w, &t1);
            if (!h0 && !h1) {
                cur = sp > 0 ? stack[--sp] : 0x7fffffff;
            } else if (h0 && h1) {
                const bool swap = t1 < t0;
                stack[sp++] = swap ? n.child0 : n.child1;
                cur = swap ? n.child1 : n.child0;
                if (sp > c->max_stack) c->max_stack = sp;
            } else {
                cur = h0 ? n.child0 : n.child1;
            }
        } else {
            const uint32_t leaf = static_cast<uint32_t>(~cur), first = leaf >> 3, count = (leaf & 7u) + 1u;
            cur = sp > 0 ? stack[--sp] : 0x7fffffff;
            for (uint32_t j = 0; j < count; ++j) {
                c->tris += 1;
                float tt;
                if (HitTri(hs.tri_verts[first + j], r, tmax, &tt)) tmax = tt, found = static_cast<int>(first + j);
            }
        }
    }
    *t_hit = tmax;
    return found;
}

uint32_t NodeHits(const WideNode &n, const Ray &r, float tmax, uint32_t oct) {
    float adj[3], org[3];
    for (int k = 0; k < 3; ++k) {
        const uint32_t bits = static_cast<uint32_t>(n.e[k]) << 23;
        float cell;
        memcpy(&cell, &bits, 4);
        adj[k] = cell * r.idir[k];
        org[k] = (n.origin[k] - r.o[k]) * r.idir[k];
    }
    const uint8_t *qlo[3] = {n.qlo_x, n.qlo_y, n.qlo_z}, *qhi[3] = {n.qhi_x, n.qhi_y, n.qhi_z};
    uint32_t mask = 0;
    for (int s = 0; s < 8; ++s) {
        const uint32_t meta = n.meta[s];
        if (meta == 0) continue;
        float enter = 1e-4f, exit = tmax;
        for (int k = 0; k < 3; ++k) {
            const bool neg = r.idir[k] < 0.0f;
            enter = fmaxf(enter, fmaf(static_cast<float>(neg ? qhi[k][s] : qlo[k][s]), adj[k], org[k]));
            exit = fminf(exit, fmaf(static_cast<float>(neg ? qlo[k][s] : qhi[k][s]), adj[k], org[k]));
        }
        if (!(enter <= exit)) continue;
        const bool inner = (meta & 0x18u) == 0x18u;
        mask |= (meta >> 5) << ((inner ? (meta ^ oct) : meta) & 0x1fu);
    }
    return mask;
}

int TraceWide(const HostScene &hs, const Ray &r, float *t_hit, Counts *c) {
    struct Group {
        uint32_t base, mask;
    };
    Group stack[kWideStackEntries];
    int sp = 0, found = -1;
    const uint32_t oct = (r.idir[0] >= 0.0f ? 1u : 0u) | (r.idir[1] >= 0.0f ? 2u : 0u) | (r.idir[2] >= 0.0f ? 4u : 0u);
    Group ng = {0u, hs.wide_nodes.empty() ? 0u : 0x80000000u};
    float tmax = 3.0e38f;
    for (;;) {
        Group tg = {0u, 0u};
        if (ng.mask > 0x00ffffffu) {
            const uint32_t hits_imask = ng.mask, bit = 31u - __builtin_clz(hits_imask), child_base = ng.base;
            ng.mask &= ~(1u << bit);
            if (ng.mask > 0x00ffffffu) {
                stack[sp++] = ng;
                if (sp > c->max_stack) c->max_stack = sp;
            }
            const uint32_t slot = (bit - 24u) ^ oct;
            const WideNode &n = hs.wide_nodes[child_base + __builtin_popcount(hits_imask & ~(0xffffffffu << slot))];
            c->nodes += 1;
            const uint32_t mask = NodeHits(n, r, tmax, oct);
            ng = {n.child_base, (mask & 0xff000000u) | n.imask};
            tg = {n.tri_base, mask & 0x00ffffffu};
        }
        while (tg.mask) {
            const uint32_t k = 31u - __builtin_clz(tg.mask);
            tg.mask &= ~(1u << k);
            c->tris += 1;
            float tt;
            if (HitTri(hs.tri_verts[tg.base + k], r, tmax, &tt)) tmax = tt, found = static_cast<int>(tg.base + k);
        }
        if (ng.mask <= 0x00ffffffu) {
            if (sp == 0) break;
            ng = stack[--sp];
        }
    }
    *t_hit = tmax;
    return found;
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: bvh_stats <pack> [grid]\n");
        return 2;
    }
    const int grid = argc > 2 ? atoi(argv[2]) : 192;
    b200pt_scene *scene = nullptr;
    if (b200pt_scene_load(argv[1], &scene) != B200PT_OK) {
        fprintf(stderr, "cannot load %s\n", argv[1]);
        return 1;
    }
    const b200pt_scene_desc *desc = b200pt_scene_get_desc(scene);
    for (int layout = 0; layout < 2; ++layout) {
        HostScene hs;
        std::string err;
        if (!BuildHostScene(*desc, 0, false, layout == 1, &hs, &err)) {
            fprintf(stderr, "build failed: %s\n", err.c_str());
            return 1;
        }
        const DCamera cam = MakeCamera(hs.camera, static_cast<uint32_t>(grid), static_cast<uint32_t>(grid));
        Counts primary, bounce;
        g_rng = 99991u;
        for (int j = 0; j < grid; ++j)
            for (int i = 0; i < grid; ++i) {
                Ray r;
                const float x = 2.0f * (i + 0.5f) / grid - 1.0f, y = 1.0f - 2.0f * (j + 0.5f) / grid;
                r.o[0] = cam.eye.x, r.o[1] = cam.eye.y, r.o[2] = cam.eye.z;
                r.d[0] = cam.front.x + x * cam.view_dx.x + y * cam.view_dy.x;
                r.d[1] = cam.front.y + x * cam.view_dx.y + y * cam.view_dy.y;
                r.d[2] = cam.front.z + x * cam.view_dx.z + y * cam.view_dy.z;
                Finish(&r);
                float t;
                const int tri = layout == 0 ? TraceBinary(hs, r, &t, &primary) : TraceWide(hs, r, &t, &primary);
                primary.rays += 1;
                if (tri < 0) continue;
                primary.hits += 1;
                // 4 cosine-ish bounce rays from the hit point, on the side the ray came from
                const TriVerts &tv = hs.tri_verts[tri];
                const float e1[3] = {tv.v1.x - tv.v0.x, tv.v1.y - tv.v0.y, tv.v1.z - tv.v0.z}, e2[3] = {tv.v2.x - tv.v0.x, tv.v2.y - tv.v0.y, tv.v2.z - tv.v0.z};
                float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                const float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                if (nl == 0.0f) continue;
                const float flip = (n[0] * r.d[0] + n[1] * r.d[1] + n[2] * r.d[2]) > 0.0f ? -1.0f : 1.0f;
                for (int k = 0; k < 3; ++k) n[k] *= flip / nl;
                for (int s = 0; s < 4; ++s) {
                    Ray b;
                    float v[3], vl;
                    do {
                        for (int k = 0; k < 3; ++k) v[k] = 2.0f * Rand() - 1.0f;
                        vl = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                    } while (vl > 1.0f || vl < 1e-6f);
                    vl = sqrtf(vl);
                    for (int k = 0; k < 3; ++k) b.o[k] = r.o[k] + t * r.d[k], b.d[k] = n[k] + v[k] / vl; // cosine lobe
                    if (b.d[0] * b.d[0] + b.d[1] * b.d[1] + b.d[2] * b.d[2] < 1e-8f) continue;
                    Finish(&b);
                    float tb;
                    const int hit = layout == 0 ? TraceBinary(hs, b, &tb, &bounce) : TraceWide(hs, b, &tb, &bounce);
                    bounce.rays += 1;
                    bounce.hits += hit >= 0;
                }
            }
        printf("%s: %zu nodes (%zu B each), %zu tris, depth %u | primary: %.0f rays, %.1f%% hit, %.2f node fetches/ray, %.2f tri tests/ray | "
               "bounce: %.0f rays, %.1f%% hit, %.2f node fetches/ray, %.2f tri tests/ray, max stack %.0f\n",
               layout == 0 ? "binary" : "wide  ", layout == 0 ? hs.nodes.size() : hs.wide_nodes.size(), layout == 0 ? sizeof(BvhNode) : sizeof(WideNode),
               hs.tri_verts.size(), hs.wide_depth, primary.rays, 100.0 * primary.hits / primary.rays, primary.nodes / primary.rays,
               primary.tris / primary.rays, bounce.rays, 100.0 * bounce.hits / std::max(1.0, bounce.rays), bounce.nodes / std::max(1.0, bounce.rays),
               bounce.tris / std::max(1.0, bounce.rays), std::max(primary.max_stack, bounce.max_stack));
    }
    b200pt_scene_free(scene);
    return 0;
}
