// lbvh.h — device-side BVH build (see lbvh.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "bvh.h"

namespace rt1w {

// h_boxes: n x 6 floats (conservative min.xyz, max.xyz of every primitive, host memory).
// Returns the device node array in the traversal layout of bvh.h (caller frees with cudaFree), the leaf order
// (leaf -> input index) on the host and the depth of the deepest leaf.  n must be >= 2.
cudaError_t build_lbvh(const float *h_boxes, size_t n_prims, cudaStream_t stream, BvhNode32 **d_nodes, size_t *n_nodes,
                       std::vector<uint32_t> &prim_order, int *depth);

} // namespace rt1w
