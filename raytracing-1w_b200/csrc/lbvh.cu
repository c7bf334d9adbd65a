// lbvh.cu — BVH build on the device for big scenes (replaces BVHNode::new, bvh.rs:54-103,
// where the host SAH builder of bvh.cpp would take seconds: 1 M spheres = 2.2 s on the host, ~10 ms here).
//
// Linear BVH (Lauterbach 2009 / Karras 2012): 63-bit Morton codes of the box centroids, one radix sort
// (cub::DeviceRadixSort, part of the CUDA toolkit), every interior node found independently from the sorted
// codes, boxes fitted bottom-up with one atomic flag per interior node, and the result written straight into the
// traversal layout of bvh.h (32-byte nodes, the two children of a node adjacent in one 64-byte pair).
#include "lbvh.h"

#include <cub/device/device_radix_sort.cuh>

#include <cmath>

namespace rt1w {
namespace {

constexpr uint32_t kLeafFlag = 0x80000000u;

__device__ __forceinline__ uint64_t spread21(uint32_t v) { // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float *__restrict__ boxes, uint32_t n, float3 lo, float3 inv_ext, uint64_t *__restrict__ keys,
                         uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *b = boxes + 6 * size_t(i);
    const float cx = (0.5f * (b[0] + b[3]) - lo.x) * inv_ext.x, cy = (0.5f * (b[1] + b[4]) - lo.y) * inv_ext.y,
                cz = (0.5f * (b[2] + b[5]) - lo.z) * inv_ext.z;
    const float s = 2097151.0f; // 2^21 - 1
    const uint32_t qx = uint32_t(fminf(fmaxf(cx * s, 0.0f), s)), qy = uint32_t(fminf(fmaxf(cy * s, 0.0f), s)),
                   qz = uint32_t(fminf(fmaxf(cz * s, 0.0f), s));
    keys[i] = spread21(qx) << 2 | spread21(qy) << 1 | spread21(qz);
    vals[i] = i;
}

// length of the common prefix of the keys of sorted positions i and j (ties broken by the position itself)
__device__ __forceinline__ int delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(uint32_t(i) ^ uint32_t(j)) : __clzll(a ^ b);
}

// Karras 2012, one thread per interior node i in [0, n - 2]
__global__ void k_topology(const uint64_t *__restrict__ keys, int n, uint32_t *__restrict__ child, uint32_t *__restrict__ iparent,
                           uint32_t *__restrict__ lparent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t left = lo == gamma ? uint32_t(gamma) | kLeafFlag : uint32_t(gamma);
    const uint32_t right = hi == gamma + 1 ? uint32_t(gamma + 1) | kLeafFlag : uint32_t(gamma + 1);
    child[2 * i] = left, child[2 * i + 1] = right;
    if (left & kLeafFlag) lparent[left & ~kLeafFlag] = uint32_t(i);
    else iparent[left] = uint32_t(i);
    if (right & kLeafFlag) lparent[right & ~kLeafFlag] = uint32_t(i);
    else iparent[right] = uint32_t(i);
}

struct Box6 {
    float v[6];
};

// bottom-up: the second thread to reach an interior node fits its box and goes on; also the subtree heights
__global__ void k_fit(const float *__restrict__ boxes, const uint32_t *__restrict__ order, int n, const uint32_t *__restrict__ child,
                      const uint32_t *__restrict__ iparent, const uint32_t *__restrict__ lparent, uint32_t *__restrict__ flags,
                      Box6 *__restrict__ ibox, int *__restrict__ iheight) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint32_t node = lparent[l];
    for (;;) {
        __threadfence();
        if (atomicAdd(&flags[node], 1u) == 0u) return; // the sibling subtree is not ready yet
        __threadfence();
        Box6 b;
        int h = 0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const uint32_t c = child[2 * node + side];
            Box6 cb;
            int ch = 0;
            if (c & kLeafFlag) {
                const float *p = boxes + 6 * size_t(order[c & ~kLeafFlag]);
#pragma unroll
                for (int k = 0; k < 6; ++k) cb.v[k] = p[k];
            } else { // written by another thread before its fence + flag increment: read past L1
                const float *p = reinterpret_cast<const float *>(ibox + c);
#pragma unroll
                for (int k = 0; k < 6; ++k) cb.v[k] = __ldcg(p + k);
                ch = __ldcg(iheight + c);
            }
            if (side == 0) {
                b = cb, h = ch;
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) b.v[k] = fminf(b.v[k], cb.v[k]), b.v[3 + k] = fmaxf(b.v[3 + k], cb.v[3 + k]);
                h = max(h, ch);
            }
        }
        ibox[node] = b, iheight[node] = h + 1;
        if (node == 0u) return;
        node = iparent[node];
    }
}

// traversal layout: node 0 = root, node 1 = padding, the children of interior node i at 2 + 2 i and 3 + 2 i
__global__ void k_emit(const float *__restrict__ boxes, const uint32_t *__restrict__ order, int n, const uint32_t *__restrict__ child,
                       const Box6 *__restrict__ ibox, BvhNode32 *__restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const uint32_t c = child[2 * i + side];
        BvhNode32 o;
        if (c & kLeafFlag) {
            const uint32_t l = c & ~kLeafFlag;
            const float *p = boxes + 6 * size_t(order[l]);
            o.min[0] = p[0], o.min[1] = p[1], o.min[2] = p[2], o.max[0] = p[3], o.max[1] = p[4], o.max[2] = p[5];
            o.left_first = l, o.count = 1;
        } else {
            const Box6 b = ibox[c];
            o.min[0] = b.v[0], o.min[1] = b.v[1], o.min[2] = b.v[2], o.max[0] = b.v[3], o.max[1] = b.v[4], o.max[2] = b.v[5];
            o.left_first = 2u + 2u * c, o.count = 0;
        }
        nodes[2 + 2 * i + side] = o;
    }
    if (i == 0) {
        const Box6 b = ibox[0];
        BvhNode32 o;
        o.min[0] = b.v[0], o.min[1] = b.v[1], o.min[2] = b.v[2], o.max[0] = b.v[3], o.max[1] = b.v[4], o.max[2] = b.v[5];
        o.left_first = 2, o.count = 0;
        nodes[0] = o;
        o.min[0] = o.min[1] = o.min[2] = o.max[0] = o.max[1] = o.max[2] = 0.0f, o.left_first = 0, o.count = 0; // never referenced
        nodes[1] = o;
    }
}

struct DeviceBuffers {
    void *p[12] = {nullptr};
    int n = 0;
    template <class T> cudaError_t alloc(T **out, size_t count) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(out), sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
    ~DeviceBuffers() {
        for (int i = 0; i < n; ++i) cudaFree(p[i]);
    }
};

} // namespace

#define RT1W_TRY(call)                                                                                                                                \
    do {                                                                                                                                              \
        cudaError_t e__ = (call);                                                                                                                     \
        if (e__ != cudaSuccess) return e__;                                                                                                           \
    } while (0)

cudaError_t build_lbvh(const float *h_boxes, size_t n_prims, cudaStream_t stream, BvhNode32 **d_nodes, size_t *n_nodes,
                       std::vector<uint32_t> &prim_order, int *depth) {
    const int n = int(n_prims);
    *d_nodes = nullptr, *n_nodes = 0, *depth = 0;
    cudaGetLastError(); // the launches below are checked with cudaGetLastError: start from a clean slate
    if (n < 2) return cudaErrorInvalidValue;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < n_prims; ++i)
        for (int k = 0; k < 3; ++k) {
            const float c = 0.5f * (h_boxes[6 * i + k] + h_boxes[6 * i + 3 + k]);
            lo[k] = std::fmin(lo[k], c), hi[k] = std::fmax(hi[k], c);
        }
    float3 flo = make_float3(lo[0], lo[1], lo[2]);
    float3 inv = make_float3(hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.0f, hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.0f,
                             hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.0f);

    DeviceBuffers buf;
    float *d_boxes;
    uint64_t *d_keys, *d_keys_sorted;
    uint32_t *d_vals, *d_order, *d_child, *d_iparent, *d_lparent, *d_flags;
    Box6 *d_ibox;
    int *d_iheight;
    RT1W_TRY(buf.alloc(&d_boxes, 6 * n_prims));
    RT1W_TRY(buf.alloc(&d_keys, n_prims));
    RT1W_TRY(buf.alloc(&d_keys_sorted, n_prims));
    RT1W_TRY(buf.alloc(&d_vals, n_prims));
    RT1W_TRY(buf.alloc(&d_order, n_prims));
    RT1W_TRY(buf.alloc(&d_child, 2 * n_prims));
    RT1W_TRY(buf.alloc(&d_iparent, n_prims));
    RT1W_TRY(buf.alloc(&d_lparent, n_prims));
    RT1W_TRY(buf.alloc(&d_flags, n_prims));
    RT1W_TRY(buf.alloc(&d_ibox, n_prims));
    RT1W_TRY(buf.alloc(&d_iheight, n_prims));
    RT1W_TRY(cudaMemcpyAsync(d_boxes, h_boxes, sizeof(float) * 6 * n_prims, cudaMemcpyHostToDevice, stream));
    RT1W_TRY(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * n_prims, stream));

    const int threads = 256, blocks = (n + threads - 1) / threads;
    k_morton<<<blocks, threads, 0, stream>>>(d_boxes, uint32_t(n), flo, inv, d_keys, d_vals);
    size_t temp_bytes = 0;
    RT1W_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys, d_keys_sorted, d_vals, d_order, n, 0, 63, stream));
    void *d_temp = nullptr;
    RT1W_TRY(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_sorted, d_vals, d_order, n, 0, 63, stream);
    if (e == cudaSuccess) {
        k_topology<<<blocks, threads, 0, stream>>>(d_keys_sorted, n, d_child, d_iparent, d_lparent);
        k_fit<<<blocks, threads, 0, stream>>>(d_boxes, d_order, n, d_child, d_iparent, d_lparent, d_flags, d_ibox, d_iheight);
        e = cudaGetLastError();
    }
    BvhNode32 *nodes = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&nodes), sizeof(BvhNode32) * 2 * n_prims);
    if (e == cudaSuccess) {
        k_emit<<<blocks, threads, 0, stream>>>(d_boxes, d_order, n, d_child, d_ibox, nodes);
        e = cudaGetLastError();
    }
    prim_order.resize(n_prims);
    int height = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(prim_order.data(), d_order, sizeof(uint32_t) * n_prims, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&height, d_iheight, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_temp);
    if (e != cudaSuccess) {
        cudaFree(nodes);
        return e;
    }
    *d_nodes = nodes, *n_nodes = 2 * n_prims, *depth = height;
    return cudaSuccess;
}

} // namespace rt1w
