// bvh.h — host binned-SAH builder producing the flat 32-byte-node BVH the extend kernel walks.
// Replaces BVHNode::new (random axis + median split, bvh.rs:54-103) and the nested per-object
// trees (AABox inner BVH aabox.rs:82, nested BVHNode main.rs:665,786): ONE tree over all leaves.
#pragma once

#include <cstdint>
#include <vector>

namespace rt1w {

struct BvhNode32 { // 2 x float4
    float min[3];
    uint32_t left_first; // interior: index of the left child (right = left+1); leaf: first primitive (leaf order)
    float max[3];
    uint32_t count;      // 0 = interior
};
static_assert(sizeof(BvhNode32) == 32, "node must be 32 bytes");

struct BvhBuildResult {
    std::vector<BvhNode32> nodes;    // node 0 = root, node 1 = padding so that sibling pairs are 64-B aligned
    std::vector<uint32_t> prim_order; // leaf order -> input index
    int depth = 0;
    double sah_cost = 0.0;
};

// Conservative f32 bounds of an f64 box: padded and rounded outward so that f32 slab arithmetic never
// culls a primitive the f64 solve would reach.  out = {min.xyz, max.xyz}.
// Extra padding of the traversal's node boxes: 1e-6 of the largest coordinate of the scene.  The FMA-form slab test
// rounds by 2^-24 (|o| + |plane - o|) in space units; covered for ray origins up to ~5 x that coordinate away.
double traversal_pad(const double *bmin, const double *bmax, size_t n);
void conservative_box(const double lo[3], const double hi[3], float out_min[3], float out_max[3], double extra_pad = 0.0);

// bmin/bmax: n x 3 doubles (world-space bounds of the lowered primitives).
void build_sah_bvh(const double *bmin, const double *bmax, size_t n, int max_leaf, BvhBuildResult &out);

} // namespace rt1w
