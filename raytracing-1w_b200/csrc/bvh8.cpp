// bvh8.cpp — see bvh8.h.
#include "bvh8.h"

#include <algorithm>
#include <cmath>
#include <cstring>

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

} // namespace

void collapse_to_bvh8(const BvhNode32 *nodes, size_t n_nodes, Bvh8BuildResult &out) {
    out.nodes.clear(), out.leaf_remap.clear(), out.depth = 0, out.avg_children = 0.0;
    if (n_nodes == 0) return;
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
        // ---- the children: open the binary subtree, largest surface area first, until eight are in hand
        uint32_t kids[8];
        int n_kids = 0;
        const BvhNode32 &top = nodes[w.binary];
        if (top.count != 0) { // a single-leaf tree
            kids[n_kids++] = w.binary;
        } else {
            kids[n_kids++] = top.left_first, kids[n_kids++] = top.left_first + 1;
            while (n_kids < 8) {
                int pick = -1;
                double best = -1.0;
                for (int k = 0; k < n_kids; ++k)
                    if (nodes[kids[k]].count == 0 && half_area(nodes[kids[k]]) > best) best = half_area(nodes[kids[k]]), pick = k;
                if (pick < 0) break;
                const uint32_t l = nodes[kids[pick]].left_first;
                kids[pick] = l, kids[n_kids++] = l + 1;
            }
        }
        slots_used += size_t(n_kids);
        // ---- the node's box and grid
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = nodes[kids[0]].min[a], hi[a] = nodes[kids[0]].max[a];
            for (int k = 1; k < n_kids; ++k) lo[a] = std::min(lo[a], nodes[kids[k]].min[a]), hi[a] = std::max(hi[a], nodes[kids[k]].max[a]);
        }
        Bvh8Node node;
        std::memset(&node, 0, sizeof(node));
        double step[3];
        for (int a = 0; a < 3; ++a) {
            node.origin[a] = lo[a];
            const double need = (double(hi[a]) - double(lo[a])) / 255.0; // 255 steps must reach the far side
            int e = 0;
            if (need > 0.0) {
                std::frexp(need, &e); // need = m 2^e with m in [0.5, 1): 2^e >= need
            } else {
                e = -126;
            }
            e = std::min(std::max(e + 127, 1), 254);
            node.exp[a] = uint8_t(e);
            step[a] = std::ldexp(1.0, e - 127);
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
                const double ql = std::floor((double(c.min[a]) - double(lo[a])) / step[a]);
                const double qh = std::ceil((double(c.max[a]) - double(lo[a])) / step[a]);
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
