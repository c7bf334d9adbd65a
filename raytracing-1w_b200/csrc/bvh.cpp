// bvh.cpp — binned SAH BVH builder (Wald 2007 style, 16 bins, leaves <= max_leaf primitives).
#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace rt1w {
namespace {

constexpr int kBins = 16;
constexpr double kInf = std::numeric_limits<double>::infinity();

struct Box {
    double lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
    void grow(const double *mn, const double *mx) {
        for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], mn[k]), hi[k] = std::max(hi[k], mx[k]);
    }
    void grow(const Box &b) { grow(b.lo, b.hi); }
    void grow_point(const double *p) { grow(p, p); }
    double area() const {
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct Builder {
    const double *bmin, *bmax;
    int max_leaf;
    std::vector<uint32_t> &order;
    std::vector<BvhNode32> &nodes;
    std::vector<double> centroid;
    int max_depth = 0;
    double cost = 0.0;
    double root_area = 1.0;
    double pad = 0.0; // extra padding of every node box: the rounding of the traversal's FMA-form slab test (kernels.cuh: slab)

    // Conservative f32 bounds: round outward and pad, so that f32 slab arithmetic never culls
    // a primitive the f64 reference would reach.
    void store_box(BvhNode32 &n, const Box &b) const { conservative_box(b.lo, b.hi, n.min, n.max, pad); }

    void make_leaf(uint32_t node, uint32_t first, uint32_t count, const Box &b, int depth) {
        store_box(nodes[node], b);
        nodes[node].left_first = first;
        nodes[node].count = count;
        max_depth = std::max(max_depth, depth);
        cost += b.area() / root_area * count;
    }

    void build(uint32_t node, uint32_t first, uint32_t count, int depth) {
        Box b, cb;
        for (uint32_t i = first; i < first + count; ++i) {
            uint32_t p = order[i];
            b.grow(bmin + 3 * p, bmax + 3 * p);
            cb.grow_point(&centroid[3 * p]);
        }
        if (depth == 0) root_area = std::max(b.area(), 1e-300);
        if (count <= 1) return make_leaf(node, first, count, b, depth);
        if (depth > 64) { // deeper than any traversal stack: stop recursing and report a depth the caller rejects (RT1W_ERR_UNSUPPORTED)
            max_depth = 1000;
            return make_leaf(node, first, 1, b, depth);
        }

        // binned SAH over the three axes
        int best_axis = -1, best_split = -1;
        double best_cost = kInf;
        for (int axis = 0; axis < 3; ++axis) {
            double lo = cb.lo[axis], hi = cb.hi[axis];
            if (!(hi > lo)) continue;
            double scale = kBins / (hi - lo);
            Box bin_box[kBins];
            uint32_t bin_n[kBins] = {0};
            for (uint32_t i = first; i < first + count; ++i) {
                uint32_t p = order[i];
                int bi = std::min(kBins - 1, int((centroid[3 * p + axis] - lo) * scale));
                bin_n[bi]++;
                bin_box[bi].grow(bmin + 3 * p, bmax + 3 * p);
            }
            double right_area[kBins];
            uint32_t right_n[kBins];
            Box acc;
            uint32_t n_acc = 0;
            for (int i = kBins - 1; i > 0; --i) {
                acc.grow(bin_box[i]);
                n_acc += bin_n[i];
                right_area[i] = acc.area(), right_n[i] = n_acc;
            }
            Box lacc;
            uint32_t ln = 0;
            for (int i = 0; i < kBins - 1; ++i) {
                lacc.grow(bin_box[i]);
                ln += bin_n[i];
                if (ln == 0 || right_n[i + 1] == 0) continue;
                double c = lacc.area() * ln + right_area[i + 1] * right_n[i + 1];
                if (c < best_cost) best_cost = c, best_axis = axis, best_split = i;
            }
        }
        const double leaf_cost = b.area() * count;
        const double trav_cost = 0.5 * b.area(); // node visit ~ half a primitive test
        bool split_by_sah = best_axis >= 0 && (best_cost + trav_cost < leaf_cost || count > uint32_t(max_leaf));
        uint32_t mid;
        if (split_by_sah) {
            double lo = cb.lo[best_axis], hi = cb.hi[best_axis];
            double scale = kBins / (hi - lo);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t p) {
                int bi = std::min(kBins - 1, int((centroid[3 * p + best_axis] - lo) * scale));
                return bi <= best_split;
            });
            mid = uint32_t(it - order.begin());
        } else if (count > uint32_t(max_leaf)) { // coincident centroids: split by index
            mid = first + count / 2;
        } else {
            return make_leaf(node, first, count, b, depth);
        }
        if (mid == first || mid == first + count) {
            if (count <= uint32_t(max_leaf)) return make_leaf(node, first, count, b, depth);
            mid = first + count / 2;
        }
        uint32_t left = uint32_t(nodes.size());
        nodes.push_back(BvhNode32{});
        nodes.push_back(BvhNode32{});
        store_box(nodes[node], b);
        nodes[node].left_first = left;
        nodes[node].count = 0;
        cost += 0.5 * b.area() / root_area;
        build(left, first, mid - first, depth + 1);
        build(left + 1, mid, first + count - mid, depth + 1);
    }
};

} // namespace

double traversal_pad(const double *bmin, const double *bmax, size_t n) {
    double reach = 0.0;
    for (size_t i = 0; i < 3 * n; ++i) reach = std::max(reach, std::max(std::fabs(bmin[i]), std::fabs(bmax[i])));
    return 1e-6 * reach;
}

void conservative_box(const double lo[3], const double hi[3], float out_min[3], float out_max[3], double extra_pad) {
    for (int k = 0; k < 3; ++k) {
        double ext = std::max(hi[k] - lo[k], std::max(std::fabs(lo[k]), std::fabs(hi[k])));
        double pad = 1e-5 * ext + 1e-5 + extra_pad;
        float l = float(lo[k] - pad), h = float(hi[k] + pad);
        out_min[k] = std::nextafter(l, -std::numeric_limits<float>::infinity());
        out_max[k] = std::nextafter(h, std::numeric_limits<float>::infinity());
    }
}

void build_sah_bvh(const double *bmin, const double *bmax, size_t n, int max_leaf, BvhBuildResult &out) {
    out.nodes.clear();
    out.prim_order.resize(n);
    for (size_t i = 0; i < n; ++i) out.prim_order[i] = uint32_t(i);
    out.nodes.reserve(2 * n + 2);
    out.nodes.push_back(BvhNode32{}); // root
    out.nodes.push_back(BvhNode32{}); // padding: children pairs start at even indices
    if (max_leaf > 7) max_leaf = 7; // node_ref (kernels.cuh) keeps three bits of the leaf count
    Builder b{bmin, bmax, max_leaf, out.prim_order, out.nodes};
    b.centroid.resize(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) b.centroid[3 * i + k] = 0.5 * (bmin[3 * i + k] + bmax[3 * i + k]);
    b.pad = traversal_pad(bmin, bmax, n);
    b.build(0, 0, uint32_t(n), 0);
    out.depth = b.max_depth;
    out.sah_cost = b.cost;
}

} // namespace rt1w
