// bvh_gpu.hpp — GPU LBVH builder (bvh_gpu.cu): a binary radix tree over Morton-sorted primitive boxes as plain arrays.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace b200pt {

struct LbvhResult {
    // n - 1 internal nodes, node 0 is the root.  Links: >= 0 internal node, < 0 primitive at sorted position ~link.
    std::vector<int32_t> left, right;
    std::vector<uint32_t> first, last; // sorted-position range covered by the node (inclusive)
    std::vector<float> boxes;          // 6 per internal node: lo.xyz hi.xyz
    std::vector<uint32_t> order;       // sorted position -> primitive index
    double gpu_ms = 0.0;               // Morton codes + sort + hierarchy + refit (CUDA events)
};

// prim_boxes: 6 floats per primitive (lo.xyz hi.xyz).  Returns false and sets *error on a CUDA failure.
bool BuildLbvhGpu(const float *prim_boxes, uint32_t n, const float scene_lo[3], const float scene_hi[3], LbvhResult *out,
                  std::string *error);

} // namespace b200pt
