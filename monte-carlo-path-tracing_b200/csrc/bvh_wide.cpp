// bvh_wide.cpp — binary SAH tree -> compressed 8-wide BVH (WideNode, device_scene.h).
//
// Collapse: a wide node starts from the two children of a binary node and keeps opening the inner child with the
// largest surface area until it has eight children or only leaves.  Children are then dealt to the eight octant slots
// by a greedy assignment that maximises, over (slot, child) pairs, the projection of the child's centre (relative to the
// node's centre) on the slot's diagonal direction: a ray then meets the slots front to back in descending (slot ^ r)
// order without any sorting.  Boxes are quantised outwards on the grid origin + q * 2^e (8 bits per plane).
#include "bvh_wide.hpp"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <deque>

namespace b200pt {

namespace {

float HalfArea(const Bvh2Node &n) {
    const float dx = n.hi[0] - n.lo[0], dy = n.hi[1] - n.lo[1], dz = n.hi[2] - n.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// Smallest biased exponent byte b with extent <= 255 * 2^(b - 127).
uint8_t GridExponent(float extent) {
    if (!(extent > 0.0f)) return 1; // flat axis: every plane quantises to 0
    int exp2 = 0;
    frexp(static_cast<double>(extent) / 255.0, &exp2); // extent/255 = m * 2^exp2, m in [0.5, 1)  =>  extent/255 <= 2^exp2
    int b = exp2 + 127;
    while (b < 254 && static_cast<double>(extent) > 255.0 * ldexp(1.0, b - 127)) ++b;
    return static_cast<uint8_t>(std::min(std::max(b, 1), 254));
}

struct Item {
    int32_t node;        // binary node this wide node is built from
    uint32_t index;      // its place in the output
    uint32_t depth;
};

} // namespace

bool BuildWideBvh(const std::vector<Bvh2Node> &nodes, int32_t root, uint32_t top_target, std::vector<uint32_t> *order,
                  std::vector<WideNode> *out, WideBuildInfo *info, std::string *error) {
    out->clear();
    *info = WideBuildInfo{};
    if (root < 0 || nodes.empty()) return true;
    const std::vector<uint32_t> old_order = *order;
    std::vector<uint32_t> new_order;
    new_order.reserve(old_order.size());
    out->reserve(nodes.size() / 4 + 16);
    out->resize(1);

    std::deque<Item> work;
    work.push_back({root, 0u, 1u});
    bool breadth_first = true;
    while (!work.empty()) {
        if (breadth_first && out->size() >= top_target) {
            breadth_first = false;
            info->top_nodes = static_cast<uint32_t>(out->size()); // places handed out so far went level by level
        }
        Item it;
        if (breadth_first) {
            it = work.front();
            work.pop_front();
        } else {
            it = work.back();
            work.pop_back();
        }
        info->depth = std::max(info->depth, it.depth);

        // ---- collapse ----
        int32_t child[8];
        int n = 0;
        if (nodes[it.node].left < 0) {
            child[n++] = it.node; // the whole tree is one leaf
        } else {
            child[n++] = nodes[it.node].left;
            child[n++] = nodes[it.node].right;
        }
        while (n < 8) {
            int best = -1;
            float best_area = -1.0f;
            for (int i = 0; i < n; ++i)
                if (nodes[child[i]].left >= 0 && HalfArea(nodes[child[i]]) > best_area) best = i, best_area = HalfArea(nodes[child[i]]);
            if (best < 0) break;
            const Bvh2Node &open = nodes[child[best]];
            child[best] = open.left;
            child[n++] = open.right;
        }

        // ---- node box, grid ----
        float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                lo[k] = fminf(lo[k], nodes[child[i]].lo[k]);
                hi[k] = fmaxf(hi[k], nodes[child[i]].hi[k]);
            }
        WideNode w;
        memset(&w, 0, sizeof(w));
        double cell[3];
        for (int k = 0; k < 3; ++k) {
            w.origin[k] = lo[k];
            w.e[k] = GridExponent(hi[k] - lo[k]);
            cell[k] = ldexp(1.0, static_cast<int>(w.e[k]) - 127);
        }

        // ---- octant slots: greedy assignment, largest projection first ----
        int slot_of[8], child_in_slot[8];
        for (int i = 0; i < 8; ++i) slot_of[i] = -1, child_in_slot[i] = -1;
        float cost[8][8];
        for (int i = 0; i < n; ++i) {
            const Bvh2Node &c = nodes[child[i]];
            float rel[3];
            for (int k = 0; k < 3; ++k) rel[k] = 0.5f * (c.lo[k] + c.hi[k]) - 0.5f * (lo[k] + hi[k]);
            for (int s = 0; s < 8; ++s) cost[s][i] = ((s & 1) ? rel[0] : -rel[0]) + ((s & 2) ? rel[1] : -rel[1]) + ((s & 4) ? rel[2] : -rel[2]);
        }
        for (int round = 0; round < n; ++round) {
            int bs = -1, bc = -1;
            for (int s = 0; s < 8; ++s) {
                if (child_in_slot[s] >= 0) continue;
                for (int i = 0; i < n; ++i) {
                    if (slot_of[i] >= 0) continue;
                    if (bs < 0 || cost[s][i] > cost[bs][bc]) bs = s, bc = i;
                }
            }
            slot_of[bc] = bs;
            child_in_slot[bs] = bc;
        }

        // ---- children: inner ones get consecutive places, leaves consecutive triangles, both in slot order ----
        uint32_t num_inner = 0;
        for (int s = 0; s < 8; ++s)
            if (child_in_slot[s] >= 0 && nodes[child[child_in_slot[s]]].left >= 0) ++num_inner;
        w.child_base = static_cast<uint32_t>(out->size());
        w.tri_base = static_cast<uint32_t>(new_order.size());
        out->resize(out->size() + num_inner);
        uint32_t next_inner = w.child_base, tri_offset = 0;
        Item pending[8];
        int num_pending = 0;
        for (int s = 0; s < 8; ++s) {
            // empty slot: meta 0 keeps it out of every hit mask; the inverted box is only a second line of defence
            w.qlo_x[s] = w.qlo_y[s] = w.qlo_z[s] = 255;
            w.qhi_x[s] = w.qhi_y[s] = w.qhi_z[s] = 0;
            if (child_in_slot[s] < 0) continue;
            const Bvh2Node &c = nodes[child[child_in_slot[s]]];
            uint8_t *qlo[3] = {w.qlo_x, w.qlo_y, w.qlo_z}, *qhi[3] = {w.qhi_x, w.qhi_y, w.qhi_z};
            for (int k = 0; k < 3; ++k) {
                const double a = floor((static_cast<double>(c.lo[k]) - static_cast<double>(lo[k])) / cell[k]);
                const double b = ceil((static_cast<double>(c.hi[k]) - static_cast<double>(lo[k])) / cell[k]);
                if (a < 0.0 || b > 255.0 || b < a) {
                    *error = "internal error: child box outside the quantisation grid of its wide BVH node.";
                    return false;
                }
                qlo[k][s] = static_cast<uint8_t>(a);
                qhi[k][s] = static_cast<uint8_t>(b);
            }
            if (c.left >= 0) {
                w.imask |= static_cast<uint8_t>(1u << s);
                w.meta[s] = static_cast<uint8_t>((1u << 5) | (24u + s));
                pending[num_pending++] = {child[child_in_slot[s]], next_inner++, it.depth + 1};
                ++info->inner_slots;
            } else {
                if (c.count == 0 || c.count > kWideMaxLeaf) {
                    *error = "internal error: binary BVH leaf does not fit a wide BVH leaf slot.";
                    return false;
                }
                w.meta[s] = static_cast<uint8_t>((((1u << c.count) - 1u) << 5) | tri_offset);
                for (uint32_t j = 0; j < c.count; ++j) new_order.push_back(old_order[c.first + j]);
                tri_offset += c.count;
                ++info->leaf_slots;
            }
        }
        (*out)[it.index] = w;
        if (breadth_first) {
            for (int i = 0; i < num_pending; ++i) work.push_back(pending[i]);
        } else {
            for (int i = num_pending; i-- > 0;) work.push_back(pending[i]); // depth-first: lowest slot on top
        }
    }
    if (breadth_first) info->top_nodes = static_cast<uint32_t>(out->size());
    if (new_order.size() != old_order.size()) {
        *error = "internal error: wide BVH lost triangles.";
        return false;
    }
    if (2u * info->depth + 2u > kWideStackEntries) {
        *error = "BVH too deep for the traversal stack (" + std::to_string(info->depth) + " wide levels).";
        return false;
    }
    *order = new_order;
    return true;
}

} // namespace b200pt
