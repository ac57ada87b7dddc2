// bvh_gpu.cu — LBVH construction on the GPU (the step immediately before the hot path).
//
// Replaces, as an option, the host-side builder for large scenes: the reference builds one LBVH per instance plus a TLAS
// on one CPU thread (src/rtcore/accel/bvh_builder.cpp:50-206: Morton codes, std::sort, top-down FindSplit), ~8 s for the
// Dragon scene; our default host builder (binned SAH, scene_build.cpp) takes 1.4 s on 8 cores.  This builder does the same
// job as the reference's — a linear BVH over Morton-sorted primitives — in a few milliseconds:
//
//   k_lbvh_morton      63-bit Morton code of every primitive's box centre (21 bits per axis, scene box normalised)
//   cub radix sort     (key, primitive) pairs                                   [library code, not on the render path]
//   k_lbvh_hierarchy   Karras 2012: every internal node of the binary radix tree in parallel (direction, range, split)
//   k_lbvh_refit       bottom-up boxes, one thread per leaf, the second arrival at a node continues upwards
//   k_lbvh_flags/index/emit   subtrees of <= max_leaf primitives become leaves (a subtree of a radix tree always covers a
//                      contiguous range of the sorted order); the surviving nodes are numbered in depth-first pre-order
//                      WITHOUT walking the tree top-down — pre-order rank = (surviving nodes whose split lies left of the
//                      node's range, a prefix sum over split positions) + (ancestors entered through their left child) —
//                      and written straight into the 64-byte traversal layout (both child boxes + links per node).
//
// Traversal quality is that of an LBVH (measured: profiles/r01_sweep_gpu_lbvh.log), which is why the SAH builder stays the
// default and this one is selected with B200PT_CREATE_GPU_LBVH / B200PT_BVH_BUILDER=lbvh.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "bvh_gpu.hpp"

namespace b200pt {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint64_t ExpandBits21(uint32_t v) { // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_lbvh_morton(const float *__restrict__ prim_boxes, uint32_t n, float3 lo, float3 inv_extent, uint64_t *keys,
                              uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *b = prim_boxes + 6ull * i;
    const float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]);
    const float scale = 2097152.0f; // 2^21
    const uint32_t x = static_cast<uint32_t>(fminf(fmaxf((cx - lo.x) * inv_extent.x * scale, 0.0f), scale - 1.0f));
    const uint32_t y = static_cast<uint32_t>(fminf(fmaxf((cy - lo.y) * inv_extent.y * scale, 0.0f), scale - 1.0f));
    const uint32_t z = static_cast<uint32_t>(fminf(fmaxf((cz - lo.z) * inv_extent.z * scale, 0.0f), scale - 1.0f));
    keys[i] = (ExpandBits21(x) << 2) | (ExpandBits21(y) << 1) | ExpandBits21(z);
    vals[i] = i;
}

// Length of the common prefix of the keys at sorted positions i and j (-1 outside the array); equal keys are told apart
// by their positions, so the tree is well defined for duplicate Morton codes (Karras 2012, §4).
__device__ __forceinline__ int Delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a != b) return __clzll(static_cast<long long>(a ^ b));
    return 64 + __clz(i ^ j);
}

// Child links: >= 0 internal node, < 0 leaf at sorted position ~link.
__global__ void k_lbvh_hierarchy(const uint64_t *__restrict__ keys, int n, int32_t *left, int32_t *right, uint32_t *first,
                                 uint32_t *last, uint32_t *split, int32_t *parent_of_internal, int32_t *parent_of_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = Delta(keys, n, i, i + 1) - Delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int delta_min = Delta(keys, n, i, i - d);
    int l_max = 2;
    while (Delta(keys, n, i, i + l_max * d) > delta_min) l_max *= 2;
    int l = 0;
    for (int t = l_max / 2; t >= 1; t /= 2)
        if (Delta(keys, n, i, i + (l + t) * d) > delta_min) l += t;
    const int j = i + l * d;
    const int delta_node = Delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) { // ceil(l / 2^k)
        if (Delta(keys, n, i, i + (s + t) * d) > delta_node) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int32_t l_link = lo == gamma ? ~gamma : gamma;
    const int32_t r_link = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    left[i] = l_link, right[i] = r_link;
    first[i] = lo, last[i] = hi;
    split[i] = gamma;
    if (l_link >= 0) parent_of_internal[l_link] = i; else parent_of_leaf[~l_link] = i;
    if (r_link >= 0) parent_of_internal[r_link] = i; else parent_of_leaf[~r_link] = i;
    if (i == 0) parent_of_internal[0] = -1;
}

__global__ void k_lbvh_refit(const float *__restrict__ prim_boxes, const uint32_t *__restrict__ order, int n, const int32_t *left,
                             const int32_t *right, const int32_t *parent_of_internal, const int32_t *parent_of_leaf, uint32_t *visits,
                             float *node_boxes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    auto child_box = [&](int32_t link, float *out) {
        const float *src = link >= 0 ? node_boxes + 6ull * link : prim_boxes + 6ull * order[~link];
        // boxes of internal children were written by another thread: read them through L2
        for (int k = 0; k < 6; ++k) out[k] = link >= 0 ? __ldcg(src + k) : src[k];
    };
    int32_t node = parent_of_leaf[p];
    while (node >= 0) {
        if (atomicAdd(visits + node, 1u) == 0u) return; // the sibling subtree is not finished yet: its thread continues
        __threadfence();
        float a[6], b[6];
        child_box(left[node], a);
        child_box(right[node], b);
        float *dst = node_boxes + 6ull * node;
        for (int k = 0; k < 3; ++k) dst[k] = fminf(a[k], b[k]);
        for (int k = 3; k < 6; ++k) dst[k] = fmaxf(a[k], b[k]);
        __threadfence();
        node = parent_of_internal[node];
    }
}

// flags[split position] = 1 for every node that stays an inner node (more than max_leaf primitives below it).
__global__ void k_lbvh_flags(const uint32_t *first, const uint32_t *last, const uint32_t *split, int n, uint32_t max_leaf, uint32_t *flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    flags[split[i]] = (last[i] - first[i] + 1u > max_leaf) ? 1u : 0u;
}

// Pre-order rank of every surviving node among the surviving nodes (see the header comment).
__global__ void k_lbvh_index(const uint32_t *first, const uint32_t *last, const int32_t *left, const int32_t *parent_of_internal, int n,
                             uint32_t max_leaf, const uint32_t *prefix, int32_t *node_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (last[i] - first[i] + 1u <= max_leaf) {
        node_index[i] = -1;
        return;
    }
    uint32_t entered_left = 0;
    for (int32_t child = i, parent = parent_of_internal[i]; parent >= 0; child = parent, parent = parent_of_internal[parent])
        entered_left += left[parent] == child ? 1u : 0u;
    node_index[i] = static_cast<int32_t>(prefix[first[i]] + entered_left);
}

__global__ void k_lbvh_emit(const float *__restrict__ prim_boxes, const uint32_t *__restrict__ order, const float *__restrict__ node_boxes,
                            const int32_t *left, const int32_t *right, const uint32_t *first, const uint32_t *last,
                            const int32_t *node_index, int n, uint32_t max_leaf, BvhNode *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || node_index[i] < 0) return;
    BvhNode node{};
    auto child = [&](int32_t link, int which) {
        const float *b;
        int32_t out_link;
        if (link < 0) { // a single primitive
            const uint32_t pos = static_cast<uint32_t>(~link);
            b = prim_boxes + 6ull * order[pos];
            out_link = ~static_cast<int32_t>(pos << 3);
        } else {
            b = node_boxes + 6ull * link;
            const uint32_t count = last[link] - first[link] + 1u;
            out_link = count <= max_leaf ? ~static_cast<int32_t>((first[link] << 3) | (count - 1u)) : node_index[link];
        }
        if (which == 0) {
            node.c0xy = {b[0], b[3], b[1], b[4]};
            node.cz.x = b[2], node.cz.y = b[5];
            node.child0 = out_link;
        } else {
            node.c1xy = {b[0], b[3], b[1], b[4]};
            node.cz.z = b[2], node.cz.w = b[5];
            node.child1 = out_link;
        }
    };
    child(left[i], 0);
    child(right[i], 1);
    out[node_index[i]] = node;
}

template <typename T>
struct Dev {
    T *p = nullptr;
    cudaError_t Alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~Dev() { cudaFree(p); }
};

} // namespace

#define LBVH_CHECK(call)                                                                       \
    do {                                                                                       \
        const cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                               \
            *error = std::string("GPU BVH build: ") + cudaGetErrorString(e_) + " (" #call ")"; \
            return false;                                                                      \
        }                                                                                      \
    } while (0)

// The whole build on the device: returns the tree already in traversal layout (depth-first pre-order, node 0 = root).
// Needs n > max_leaf (otherwise the scene is a single leaf and the caller builds it on the host).
bool BuildLbvhGpuFlat(const float *prim_boxes, uint32_t n, const float scene_lo[3], const float scene_hi[3], uint32_t max_leaf,
                      std::vector<BvhNode> *nodes, std::vector<uint32_t> *order, double *gpu_ms, std::string *error) {
    if (n < 2 || n <= max_leaf) {
        *error = "GPU BVH build: too few primitives";
        return false;
    }
    Dev<float> d_boxes, d_node_boxes;
    Dev<uint64_t> d_keys, d_keys_sorted;
    Dev<uint32_t> d_vals, d_order, d_first, d_last, d_split, d_visits, d_flags, d_prefix;
    Dev<int32_t> d_left, d_right, d_parent_internal, d_parent_leaf, d_index;
    Dev<uint8_t> d_temp;
    Dev<BvhNode> d_nodes;
    LBVH_CHECK(d_boxes.Alloc(6ull * n));
    LBVH_CHECK(d_node_boxes.Alloc(6ull * (n - 1)));
    LBVH_CHECK(d_keys.Alloc(n));
    LBVH_CHECK(d_keys_sorted.Alloc(n));
    LBVH_CHECK(d_vals.Alloc(n));
    LBVH_CHECK(d_order.Alloc(n));
    LBVH_CHECK(d_first.Alloc(n - 1));
    LBVH_CHECK(d_last.Alloc(n - 1));
    LBVH_CHECK(d_split.Alloc(n - 1));
    LBVH_CHECK(d_visits.Alloc(n - 1));
    LBVH_CHECK(d_flags.Alloc(n));
    LBVH_CHECK(d_prefix.Alloc(n));
    LBVH_CHECK(d_left.Alloc(n - 1));
    LBVH_CHECK(d_right.Alloc(n - 1));
    LBVH_CHECK(d_parent_internal.Alloc(n - 1));
    LBVH_CHECK(d_parent_leaf.Alloc(n));
    LBVH_CHECK(d_index.Alloc(n - 1));
    LBVH_CHECK(d_nodes.Alloc(n - 1));
    size_t sort_bytes = 0, scan_bytes = 0;
    LBVH_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys.p, d_keys_sorted.p, d_vals.p, d_order.p, static_cast<int>(n), 0, 63));
    LBVH_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flags.p, d_prefix.p, static_cast<int>(n)));
    LBVH_CHECK(d_temp.Alloc(std::max(sort_bytes, scan_bytes)));

    cudaEvent_t ev0, ev1;
    LBVH_CHECK(cudaEventCreate(&ev0));
    LBVH_CHECK(cudaEventCreate(&ev1));
    LBVH_CHECK(cudaMemcpy(d_boxes.p, prim_boxes, 6ull * n * sizeof(float), cudaMemcpyHostToDevice));
    LBVH_CHECK(cudaEventRecord(ev0));
    const float3 lo = make_float3(scene_lo[0], scene_lo[1], scene_lo[2]);
    auto inv = [](float e) { return e > 0.0f ? 1.0f / e : 0.0f; };
    const float3 inv_extent = make_float3(inv(scene_hi[0] - scene_lo[0]), inv(scene_hi[1] - scene_lo[1]), inv(scene_hi[2] - scene_lo[2]));
    const int blocks_n = static_cast<int>((n + kThreads - 1) / kThreads), ni = static_cast<int>(n);
    k_lbvh_morton<<<blocks_n, kThreads>>>(d_boxes.p, n, lo, inv_extent, d_keys.p, d_vals.p);
    LBVH_CHECK(cub::DeviceRadixSort::SortPairs(d_temp.p, sort_bytes, d_keys.p, d_keys_sorted.p, d_vals.p, d_order.p, ni, 0, 63));
    k_lbvh_hierarchy<<<blocks_n, kThreads>>>(d_keys_sorted.p, ni, d_left.p, d_right.p, d_first.p, d_last.p, d_split.p, d_parent_internal.p,
                                            d_parent_leaf.p);
    LBVH_CHECK(cudaMemsetAsync(d_visits.p, 0, (n - 1) * sizeof(uint32_t)));
    k_lbvh_refit<<<blocks_n, kThreads>>>(d_boxes.p, d_order.p, ni, d_left.p, d_right.p, d_parent_internal.p, d_parent_leaf.p, d_visits.p,
                                        d_node_boxes.p);
    LBVH_CHECK(cudaMemsetAsync(d_flags.p, 0, n * sizeof(uint32_t)));
    k_lbvh_flags<<<blocks_n, kThreads>>>(d_first.p, d_last.p, d_split.p, ni, max_leaf, d_flags.p);
    LBVH_CHECK(cub::DeviceScan::ExclusiveSum(d_temp.p, scan_bytes, d_flags.p, d_prefix.p, ni));
    k_lbvh_index<<<blocks_n, kThreads>>>(d_first.p, d_last.p, d_left.p, d_parent_internal.p, ni, max_leaf, d_prefix.p, d_index.p);
    k_lbvh_emit<<<blocks_n, kThreads>>>(d_boxes.p, d_order.p, d_node_boxes.p, d_left.p, d_right.p, d_first.p, d_last.p, d_index.p, ni, max_leaf,
                                       d_nodes.p);
    LBVH_CHECK(cudaEventRecord(ev1));
    LBVH_CHECK(cudaGetLastError());
    uint32_t num_nodes = 0; // flags has n entries (the last one is always 0), so prefix[n - 1] is the total
    LBVH_CHECK(cudaMemcpy(&num_nodes, d_prefix.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    nodes->resize(num_nodes);
    order->resize(n);
    LBVH_CHECK(cudaMemcpy(nodes->data(), d_nodes.p, static_cast<size_t>(num_nodes) * sizeof(BvhNode), cudaMemcpyDeviceToHost));
    LBVH_CHECK(cudaMemcpy(order->data(), d_order.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    float ms = 0.0f;
    LBVH_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    *gpu_ms = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return true;
}

} // namespace b200pt
