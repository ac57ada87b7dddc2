// bvh_gpu.hpp — GPU LBVH builder (bvh_gpu.cu): Morton sort, Karras radix tree, refit, collapse and flatten on the device.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "device_scene.h"

namespace b200pt {

// prim_boxes: 6 floats per primitive (lo.xyz hi.xyz); needs n > max_leaf.  On success `nodes` holds the tree in the traversal
// layout (device_scene.h: BvhNode, depth-first pre-order, node 0 = root, leaves of <= max_leaf primitives), `order` maps
// leaf positions to primitive indices and *gpu_ms is the device time of the whole build (CUDA events).
// Returns false and sets *error on a CUDA failure.
bool BuildLbvhGpuFlat(const float *prim_boxes, uint32_t n, const float scene_lo[3], const float scene_hi[3], uint32_t max_leaf,
                      std::vector<BvhNode> *nodes, std::vector<uint32_t> *order, double *gpu_ms, std::string *error);

} // namespace b200pt
