// bvh8.h — the compressed 8-wide BVH big scenes are traversed through (after Ylitie, Karras, Laine 2017, "Efficient
// Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs"; own layout, one primitive per leaf slot).
//
// Why: closest hit over the binary tree of bvh.h is a chain of dependent 64-byte fetches, ~130 of them per ray on the
// 1 M-sphere stress scene (23 levels), and the wave kernel spends its time waiting for them (ncu: long_scoreboard 5.9
// cycles per issue, issue slots 30 % used).  An 8-wide node is ONE 80-byte fetch (three 32-byte sectors) for eight child
// boxes: a third of the dependent fetches and less than half the bytes per ray.  Replaces bvh.rs:25-50 / aabb.rs:13-32
// exactly like the binary tree does: same leaves, same conservative boxes, same closest hits.
//
// Node (80 bytes, 5 x 16):
//   word 0  origin of the child grid (3 x f32 = the node box minimum), then 4 bytes: the grid's step per axis as a
//           biased f32 exponent (step = 2^(e - 127), the smallest power of two with 255 steps >= the node's extent)
//           and `imask` (bit s set: slot s holds an INTERIOR child)
//   word 1  child_base (node index of the first interior child; the interior children of a node are adjacent, in slot
//           order), prim_base (leaf index of the first leaf slot's primitive; likewise adjacent, in slot order),
//           leaf_mask (bit s set: slot s holds ONE primitive), unused
//   words 2-4  child boxes on the grid, one byte per plane: lo.x[8], lo.y[8], lo.z[8], hi.x[8], hi.y[8], hi.z[8]
//           (lo rounded down, hi rounded up: the decoded box contains the child's conservative f32 box; an empty
//           slot holds lo = 255, hi = 0)
// Slots are assigned so that for a ray of octant r (bit k set: direction component k >= 0) visiting the slots in
// DEcreasing order of (slot XOR r) is approximately front to back (children sorted along the diagonals of the node), so
// a traversal needs no distance sort and keeps ONE stack entry per visited node (the not-yet-visited hit slots).
#pragma once

#include <cstdint>
#include <vector>

#include "bvh.h"

namespace rt1w {

struct Bvh8Node {
    float origin[3];
    uint8_t exp[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t prim_base;
    uint32_t leaf_mask;
    uint32_t unused;
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
};
static_assert(sizeof(Bvh8Node) == 80, "wide node must be 5 x 16 bytes");

struct Bvh8BuildResult {
    std::vector<Bvh8Node> nodes;      // node 0 = root
    std::vector<uint32_t> leaf_remap; // new leaf index -> leaf index of the binary tree it was collapsed from
    int depth = 0;                    // deepest node (root = 0)
    double avg_children = 0.0;        // occupied slots per node
};

// Collapses a binary tree in the layout of bvh.h (single-primitive leaves, node 0 = root, siblings adjacent) into the wide
// tree: a node's children are found by opening, largest surface area first, interior children of the binary subtree
// until eight are in hand.  `nodes` may come from either builder (host SAH, device LBVH).
void collapse_to_bvh8(const BvhNode32 *nodes, size_t n_nodes, Bvh8BuildResult &out);

} // namespace rt1w
