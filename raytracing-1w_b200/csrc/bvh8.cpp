// bvh8.cpp — see bvh8.h.
#include "bvh8.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace rt1w {
namespace {

double half_area(const BvhNode32 &n) {
    const double dx = double(n.max[0]) - n.min[0], dy = double(n.max[1]) - n.min[1], dz = double(n.max[2]) - n.min[2];
    return dx * dy + dy * dz + dz * dx;
}

struct Work {
    uint32_t binary; // interior node of the binary tree (or the root, whatever it is)
    uint32_t wide;   // its place in the wide tree
    int depth;
};

// Which binary nodes become wide nodes: the surface-area-heuristic optimum of Ylitie et al. (section 3.1) by dynamic
// programming over the binary tree.  cost[n][i - 1] = least cost of turning the subtree of n into at most i wide-tree
// children (i = 1: ONE child - the primitive itself for a leaf, else a wide node whose own up-to-8 children come out of
// n's two subtrees); a wide node costs its area (one fetch + eight box tests per visit), a leaf slot its area (one f64
// solve per visit).  split[n][i - 1]: how many of the i come from the left subtree (0: "i - 1 were enough").
struct Plan {
    std::vector<float> cost;    // 7 per node
    std::vector<uint8_t> split; // 8 per node: [0..6] for i = 1..7 (unused for i = 1), [7] for the node's own 8 children
};

void plan_collapse(const BvhNode32 *nodes, size_t n_nodes, Plan &plan) {
    plan.cost.assign(7 * n_nodes, 0.0f), plan.split.assign(8 * n_nodes, 0);
    // post-order without recursion: children before parents
    std::vector<uint32_t> order, stack;
    order.reserve(n_nodes), stack.push_back(0u);
    while (!stack.empty()) {
        const uint32_t n = stack.back();
        stack.pop_back();
        order.push_back(n);
        if (nodes[n].count == 0) stack.push_back(nodes[n].left_first), stack.push_back(nodes[n].left_first + 1);
    }
    for (size_t k = order.size(); k-- > 0;) {
        const uint32_t n = order[k];
        const float area = float(half_area(nodes[n]));
        float *c = &plan.cost[7 * size_t(n)];
        uint8_t *sp = &plan.split[8 * size_t(n)];
        if (nodes[n].count != 0) {
            for (int i = 0; i < 7; ++i) c[i] = area;
            continue;
        }
        const float *cl = &plan.cost[7 * size_t(nodes[n].left_first)], *cr = cl + 7;
        auto distribute = [&](int j, uint8_t &best_k) { // j >= 2 children out of the two subtrees
            float best = INFINITY;
            for (int kk = 1; kk < j; ++kk) {
                if (kk > 7 || j - kk > 7) continue;
                const float v = cl[kk - 1] + cr[j - kk - 1];
                if (v < best) best = v, best_k = uint8_t(kk);
            }
            return best;
        };
        c[0] = distribute(8, sp[7]) + area; // n itself becomes a wide node
        for (int i = 2; i <= 7; ++i) {
            uint8_t kk = 0;
            const float d = distribute(i, kk);
            if (d < c[i - 2]) c[i - 1] = d, sp[i - 1] = kk;
            else c[i - 1] = c[i - 2], sp[i - 1] = 0;
        }
    }
}

// the (at most i) wide-tree children the plan makes of the subtree of n
void collect_children(const BvhNode32 *nodes, const Plan &plan, uint32_t n, int i, uint32_t *kids, int &n_kids) {
    while (nodes[n].count == 0 && i > 1 && plan.split[8 * size_t(n) + size_t(i - 1)] == 0) --i; // "i - 1 were enough"
    if (nodes[n].count != 0 || i == 1) {
        kids[n_kids++] = n;
        return;
    }
    const int k = plan.split[8 * size_t(n) + size_t(i - 1)];
    collect_children(nodes, plan, nodes[n].left_first, k, kids, n_kids);
    collect_children(nodes, plan, nodes[n].left_first + 1, i - k, kids, n_kids);
}

} // namespace

void collapse_to_bvh8(const BvhNode32 *nodes, size_t n_nodes, Bvh8BuildResult &out) {
    out.nodes.clear(), out.leaf_remap.clear(), out.depth = 0, out.avg_children = 0.0;
    if (n_nodes == 0) return;
    Plan plan;
    plan_collapse(nodes, n_nodes, plan);
    out.nodes.reserve(n_nodes / 5 + 16);
    out.leaf_remap.reserve(n_nodes / 2 + 1);
    std::vector<Work> queue;
    queue.reserve(n_nodes / 5 + 16);
    out.nodes.push_back(Bvh8Node{});
    queue.push_back(Work{0u, 0u, 0});
    size_t slots_used = 0;
    for (size_t qi = 0; qi < queue.size(); ++qi) {
        const Work w = queue[qi];
        out.depth = std::max(out.depth, w.depth);
        // ---- the children the plan gives this node
        uint32_t kids[8];
        int n_kids = 0;
        const BvhNode32 &top = nodes[w.binary];
        if (top.count != 0) { // a single-leaf tree
            kids[n_kids++] = w.binary;
        } else {
            const int k = plan.split[8 * size_t(w.binary) + 7];
            collect_children(nodes, plan, top.left_first, k, kids, n_kids);
            collect_children(nodes, plan, top.left_first + 1, 8 - k, kids, n_kids);
        }
        slots_used += size_t(n_kids);
        // ---- the node's box and grid
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = nodes[kids[0]].min[a], hi[a] = nodes[kids[0]].max[a];
            for (int k = 1; k < n_kids; ++k) lo[a] = std::min(lo[a], nodes[kids[k]].min[a]), hi[a] = std::max(hi[a], nodes[kids[k]].max[a]);
        }
        // The grid: step = the smallest power of two with 254 steps >= the extent, origin 1/64 of a step below the box minimum.
        // Child planes are rounded outward with 1/128 of a step to spare (they then lie in [0, 255]): the traversal's
        // byte-to-float trick (kernels.cuh: wide_visit) may misplace a decoded plane by 2^-9 of a step.
        Bvh8Node node;
        std::memset(&node, 0, sizeof(node));
        double step[3], origin[3];
        for (int a = 0; a < 3; ++a) {
            const double need = (double(hi[a]) - double(lo[a])) / 254.0;
            int e = -126;
            if (need > 0.0) std::frexp(need, &e); // need = m 2^e with m in [0.5, 1): 2^e >= need
            e = std::min(std::max(e + 127, 24), 230); // (kept away from the ends of the exponent range: the traversal scales the step by 2^15)
            node.exp[a] = uint8_t(e);
            step[a] = std::ldexp(1.0, e - 127);
            float o = float(double(lo[a]) - step[a] / 64.0);
            if (double(o) > double(lo[a]) - step[a] / 64.0) o = std::nextafter(o, -std::numeric_limits<float>::infinity());
            node.origin[a] = o, origin[a] = double(o);
        }
        // ---- slots: child c goes where the diagonal of its slot (bit k set: positive side of axis k) points at it, greedily
        double centre[3], off[8][3];
        for (int a = 0; a < 3; ++a) centre[a] = 0.5 * (double(lo[a]) + double(hi[a]));
        for (int k = 0; k < n_kids; ++k)
            for (int a = 0; a < 3; ++a) off[k][a] = 0.5 * (double(nodes[kids[k]].min[a]) + double(nodes[kids[k]].max[a])) - centre[a];
        int slot_of[8], kid_in[8];
        for (int k = 0; k < 8; ++k) slot_of[k] = -1, kid_in[k] = -1;
        for (int round = 0; round < n_kids; ++round) {
            int bk = -1, bs = -1;
            double best = -1e300;
            for (int k = 0; k < n_kids; ++k) {
                if (slot_of[k] >= 0) continue;
                for (int s = 0; s < 8; ++s) {
                    if (kid_in[s] >= 0) continue;
                    const double c = ((s & 1) ? off[k][0] : -off[k][0]) + ((s & 2) ? off[k][1] : -off[k][1]) + ((s & 4) ? off[k][2] : -off[k][2]);
                    if (c > best) best = c, bk = k, bs = s;
                }
            }
            slot_of[bk] = bs, kid_in[bs] = bk;
        }
        // ---- emit: interior children adjacent in slot order, leaf primitives adjacent in slot order
        node.child_base = uint32_t(out.nodes.size());
        node.prim_base = uint32_t(out.leaf_remap.size());
        for (int s = 0; s < 8; ++s) {
            for (int a = 0; a < 3; ++a) node.qlo[a][s] = 255, node.qhi[a][s] = 0; // empty
            const int k = kid_in[s];
            if (k < 0) continue;
            const BvhNode32 &c = nodes[kids[k]];
            for (int a = 0; a < 3; ++a) {
                const double ql = std::floor((double(c.min[a]) - origin[a]) / step[a] - 1.0 / 128.0);
                const double qh = std::ceil((double(c.max[a]) - origin[a]) / step[a] + 1.0 / 128.0);
                node.qlo[a][s] = uint8_t(std::min(std::max(ql, 0.0), 255.0));
                node.qhi[a][s] = uint8_t(std::min(std::max(qh, 0.0), 255.0));
            }
            if (c.count == 0) {
                node.imask |= uint8_t(1u << s);
                queue.push_back(Work{kids[k], uint32_t(out.nodes.size()), w.depth + 1});
                out.nodes.push_back(Bvh8Node{});
            } else {
                node.leaf_mask |= 1u << s;
                out.leaf_remap.push_back(c.left_first);
            }
        }
        out.nodes[w.wide] = node;
    }
    out.avg_children = double(slots_used) / double(out.nodes.size());
}

} // namespace rt1w
