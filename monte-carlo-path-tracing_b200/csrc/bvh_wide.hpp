// bvh_wide.hpp — collapse of a binary BVH into the compressed 8-wide layout the traversal kernels walk by default.
//
// Stands where the reference flattens its per-instance LBVHs + TLAS into one node array (src/rtcore/scene.cpp:499-525,
// src/rtcore/accel/bvh_builder.cpp:74-206); the layout itself (WideNode, device_scene.h) has no reference counterpart.
#pragma once
#include <string>
#include <vector>

#include "device_scene.h"

namespace b200pt {

// One node of the binary tree handed to the collapse: leaves have left < 0 and own order[first .. first + count).
struct Bvh2Node {
    float lo[3], hi[3];
    int32_t left = -1, right = -1;
    uint32_t first = 0, count = 0;
};

constexpr uint32_t kWideMaxLeaf = 3;     // triangles per leaf slot (unary count in 3 meta bits)
constexpr uint32_t kWideStackEntries = 64; // traversal stack of the wide kernels (traverse_wide.cuh), in uint2 entries

struct WideBuildInfo {
    uint32_t depth = 0;          // levels of wide nodes
    uint32_t inner_slots = 0, leaf_slots = 0;
    uint32_t top_nodes = 0;      // leading nodes that were placed breadth-first (the top of the tree): any prefix of them is worth staging in shared memory
};

// `order`: on input the triangle permutation the binary tree's leaf ranges index; on output the permutation in which
// WideNode::tri_base + offset addresses triangles (leaf order of the wide tree).  Binary leaves must hold at most
// kWideMaxLeaf triangles.  `top_target`: nodes are emitted breadth-first until this many exist, depth-first below.
bool BuildWideBvh(const std::vector<Bvh2Node> &nodes, int32_t root, uint32_t top_target, std::vector<uint32_t> *order,
                  std::vector<WideNode> *out, WideBuildInfo *info, std::string *error);

} // namespace b200pt
