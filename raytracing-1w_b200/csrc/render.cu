// render.cu — the wavefront kernels and the wave loop (sm_100a).
//
// The reference's pixel loop + recursive ray_color (main.rs:51-190, 957-1001) become, per wave:
//
//   generate   refills terminated path slots with new camera paths   (main.rs:968-971, camera.rs:61-73)
//   extend     closest hit over the flat SAH BVH, sorts hits into per-material queues (bvh.rs:25-50 ...)
//   shade_<m>  one kernel per material family: scatter + pdf weighting (material.rs, pdf.rs)
//
// Path state lives in SoA arrays in HBM (render.h: Pool); kernels exchange 4-byte slot indices
// through queues filled with warp-aggregated atomics (__ballot_sync + __popc + __shfl_sync).
// The recursion `emitted + attenuation * f * L / pdf` is unrolled into a running throughput:
// only terminal events (DiffuseLight, miss) carry radiance, so a path adds to its pixel exactly once.
#include "render.h"

#include <cstdio>
#include <vector>

namespace rt1w {

constexpr int kGenThreads = 256;
constexpr int kExtendThreads = 128;
constexpr int kShadeThreads = 128;

// ------------------------------------------------------------------------------------------
// Queue push: one atomic per warp per queue.
// ------------------------------------------------------------------------------------------
RT1W_DEV void warp_push(uint32_t *__restrict__ queue, uint32_t *counter, bool pred, uint32_t value) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// Sorts the lanes of a warp into the queues of the table with ONE atomic instruction: lane q reserves
// queue q's slots for the whole warp (QS_COUNT lanes, QS_COUNT addresses, one round trip), then every
// lane fetches the base of its own destination with a shuffle.  dest < 0: the lane pushes nothing.
// `live` is a compile-time mask of the queue slots this call site can produce.
template <uint32_t LIVE> RT1W_DEV void warp_sort_push(const Pool &pool, int dest, uint32_t value) {
    const int lane = threadIdx.x & 31;
    uint32_t mine = 0, count_for_lane = 0;
#pragma unroll
    for (int q = 0; q < QS_COUNT; ++q) {
        if (!(LIVE & (1u << q))) continue;
        const unsigned m = __ballot_sync(0xffffffffu, dest == q);
        if (dest == q) mine = m;
        if (lane == q) count_for_lane = uint32_t(__popc(m));
    }
    uint32_t base = 0;
    if (count_for_lane) base = atomicAdd(&pool.ctr->n[lane], count_for_lane);
    base = __shfl_sync(0xffffffffu, base, dest < 0 ? 0 : dest);
    if (dest >= 0) pool.q[dest][base + __popc(mine & ((1u << lane) - 1u))] = value;
}

// A path ends: pixel += throughput * radiance.  A NaN product must reach the sum even when the
// radiance is zero: the reference turns a NaN pixel SUM into black (color.rs:14-21), and
// `li / pdf` with pdf == 0 is NaN there whatever li is (main.rs:102).
RT1W_DEV void splat(const RenderArgs &a, uint32_t pixel, f3 thr, f3 radiance) {
    const float c[3] = {thr.x * radiance.x, thr.y * radiance.y, thr.z * radiance.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (c[k] != 0.0f) atomicAdd(a.accum + 3 * size_t(pixel) + k, c[k]);
    }
    if (a.stat) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float x = c[k];
            if (x != x) x = 0.0f;
            x = fminf(x, a.rp.stat_clamp);
            if (x != 0.0f) {
                atomicAdd(a.stat + 6 * size_t(pixel) + k, x);
                atomicAdd(a.stat + 6 * size_t(pixel) + 3 + k, x * x);
            }
        }
    }
}

RT1W_DEV Ray load_ray(const Pool &p, uint32_t slot, RayC &c) {
    const double2 a = p.o_xy[slot];
    const RayB b = p.o_zd[slot];
    c = p.dzm[slot];
    Ray r;
    r.ox = a.x, r.oy = a.y, r.oz = b.oz;
    r.dx = b.dx, r.dy = b.dy, r.dz = c.dz;
    r.time = c.time;
    return r;
}

RT1W_DEV void store_ray(const Pool &p, uint32_t slot, double ox, double oy, double oz, f3 d, float time, uint32_t state, uint32_t pixel) {
    p.o_xy[slot] = make_double2(ox, oy);
    RayB b;
    b.oz = oz, b.dx = d.x, b.dy = d.y;
    p.o_zd[slot] = b;
    RayC c;
    c.dz = d.z, c.time = time, c.state = state, c.pixel = pixel;
    p.dzm[slot] = c;
}

// Philox key/counter of a path: key = (reference pixel seed j*w+i (main.rs:964), seed), counter = (sample, bounce, purpose, block)
RT1W_DEV void path_rng_key(const DRenderParams &rp, uint32_t pixel, uint32_t &k0, uint32_t &k1) {
    const uint32_t row = pixel / uint32_t(rp.width), col = pixel - row * uint32_t(rp.width);
    const uint32_t j = uint32_t(rp.height) - 1u - row;
    k0 = j * uint32_t(rp.width) + col;
    k1 = rp.seed_lo;
}
RT1W_DEV uint32_t purpose_word(const DRenderParams &rp, uint32_t stream) { return stream ^ (rp.seed_hi << 4); }

// ------------------------------------------------------------------------------------------
// generate
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenThreads) k_generate(const __grid_constant__ RenderArgs a, const int parity, const int initial) {
    Counters *ctr = a.pool.ctr;
    const uint32_t n = initial ? a.pool.capacity : ctr->n[QS_FREE + (parity ^ 1)];
    if (blockIdx.x == 0 && threadIdx.x == 0) { // recycle the counters nobody reads during this wave's generate
#pragma unroll
        for (int q = 0; q < Q_COUNT; ++q) ctr->n[QS_MAT + q] = 0;
        ctr->n[QS_FREE + parity] = 0;
        ctr->n[QS_EXTEND + (parity ^ 1)] = 0;
    }
    const uint32_t *free_q = a.pool.q[QS_FREE + (parity ^ 1)];
    const unsigned long long total = a.rp.total_paths;
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        const uint32_t slot = valid ? (initial ? i : free_q[i]) : 0u;
        // claim path indices, one atomic per warp
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        unsigned long long base = 0;
        if (lane == 0 && m) base = atomicAdd(&ctr->next_path, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned long long k = base + __popc(m & ((1u << lane) - 1u));
        const bool live = valid && k < total;
        if (live) {
            const uint32_t sample_rel = uint32_t(k / a.rp.n_pixels);
            const uint32_t pixel = uint32_t(k - (unsigned long long)sample_rel * a.rp.n_pixels);
            const uint32_t row = pixel / uint32_t(a.rp.width), col = pixel - row * uint32_t(a.rp.width);
            const uint32_t j = uint32_t(a.rp.height) - 1u - row; // main.rs:959: rows are emitted top first
            Rng rng;
            path_rng_key(a.rp, pixel, rng.k0, rng.k1);
            rng.c0 = uint32_t(a.rp.sample_begin) + sample_rel, rng.c1 = 0, rng.c2 = purpose_word(a.rp, RNG_CAMERA), rng.block = 0;
            const Philox4 x = rng.next4();
            const double s = (double(col) + double(u01(x.x))) / double(a.rp.width - 1);  // main.rs:968
            const double t = (double(j) + double(u01(x.y))) / double(a.rp.height - 1);   // main.rs:969
            const float time = a.cam.time0 + (a.cam.time1 - a.cam.time0) * u01(x.z);     // camera.rs:71
            double offx = 0.0, offy = 0.0, offz = 0.0;
            if (a.cam.lens_radius != 0.0) { // camera.rs:62-63; the rejection loop of math.rs:30-37
                float px, py;
                for (;;) {
                    const Philox4 y = rng.next4();
                    px = 2.0f * u01(y.x) - 1.0f, py = 2.0f * u01(y.y) - 1.0f;
                    if (px * px + py * py < 1.0f) break;
                    px = 2.0f * u01(y.z) - 1.0f, py = 2.0f * u01(y.w) - 1.0f;
                    if (px * px + py * py < 1.0f) break;
                }
                const double rx = a.cam.lens_radius * double(px), ry = a.cam.lens_radius * double(py);
                offx = a.cam.u[0] * rx + a.cam.v[0] * ry;
                offy = a.cam.u[1] * rx + a.cam.v[1] * ry;
                offz = a.cam.u[2] * rx + a.cam.v[2] * ry;
            }
            // camera.rs:67-70: direction = lower_left_corner + s*horizontal + t*vertical - origin - offset (never normalised)
            const f3 d = mk3(float(a.cam.llc_rel[0] + s * a.cam.horizontal[0] + t * a.cam.vertical[0] - offx),
                             float(a.cam.llc_rel[1] + s * a.cam.horizontal[1] + t * a.cam.vertical[1] - offy),
                             float(a.cam.llc_rel[2] + s * a.cam.horizontal[2] + t * a.cam.vertical[2] - offz));
            store_ray(a.pool, slot, a.cam.origin[0] + offx, a.cam.origin[1] + offy, a.cam.origin[2] + offz, d, time, sample_rel << 8, pixel);
            a.pool.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        }
        warp_push(a.pool.q[QS_EXTEND + parity], &ctr->n[QS_EXTEND + parity], live, slot);
    }
}

// ------------------------------------------------------------------------------------------
// extend
// ------------------------------------------------------------------------------------------
// FLAT: scan the primitive list staged in shared memory (scenes of <= kFlatMax primitives) instead of walking the BVH.
template <bool FLAT>
__global__ void __launch_bounds__(kExtendThreads) k_extend(const __grid_constant__ RenderArgs a, const int parity, const int material_mask) {
    // one shared buffer: the staged primitive list (FLAT) or the per-thread traversal stacks (BVH)
    __shared__ __align__(16) unsigned char s_raw[FLAT ? sizeof(FlatScene) : sizeof(uint2) * kStackSmem * kExtendThreads];
    uint2 *s_stack = reinterpret_cast<uint2 *>(s_raw);
    FlatScene *s_flat = reinterpret_cast<FlatScene *>(s_raw);
    Counters *ctr = a.pool.ctr;
    const uint32_t n = ctr->n[QS_EXTEND + parity];
    if (n == 0) return;
    if (FLAT) {
        flat_stage(a.sc, s_flat[0]);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr->rays, (unsigned long long)n);
    const uint32_t *in_q = a.pool.q[QS_EXTEND + parity];
    const bool has_background = a.rp.background[0] != 0.0f || a.rp.background[1] != 0.0f || a.rp.background[2] != 0.0f;
    const bool has_media = (material_mask & (1 << RT1W_MAT_ISOTROPIC)) != 0;
    const int free_dest = QS_FREE + parity;
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        uint32_t slot = 0;
        int dest = -1;
        if (valid) {
            slot = in_q[i];
            RayC c;
            const Ray r = load_ray(a.pool, slot, c);
            MediumRng mr = {0, 0, 0, 0, 0};
            if (has_media) { // only ConstantMedium candidates draw random numbers inside the traversal (constant_medium.rs:85)
                path_rng_key(a.rp, c.pixel, mr.k0, mr.k1);
                mr.c0 = uint32_t(a.rp.sample_begin) + (c.state >> 8), mr.c1 = c.state & 255u, mr.c2 = purpose_word(a.rp, RNG_MEDIUM);
            }
            double t;
            int leaf;
            const bool hit = FLAT ? closest_hit_flat<false>(a.sc, s_flat[0], r, mr, t, leaf)
                                  : closest_hit<false>(a.sc, r, mr, s_stack + threadIdx.x, kExtendThreads, t, leaf);
            if (hit) {
                const uint32_t meta = FLAT ? s_flat[0].prims[leaf].meta : __ldg(&a.sc.prims[leaf].meta);
                const int mat_type = int((meta >> 8) & 15u);
                if (mat_type == RT1W_MAT_NONE) { // `impl Material for ()`: no emission, no scatter (material.rs:68)
                    dest = free_dest;
                } else {
                    HitRec h;
                    h.t = t, h.leaf = leaf, h.pad = 0;
                    a.pool.hit[slot] = h;
                    dest = mat_type;
                }
            } else {
                dest = free_dest;
            }
            if (dest == free_dest) { // main.rs:113-115 (miss -> background) or a null-material hit (zero radiance)
                const float4 th = a.pool.thr[slot];
                const f3 rad = hit ? mk3(0.0f, 0.0f, 0.0f) : mk3(a.rp.background[0], a.rp.background[1], a.rp.background[2]);
                const bool finite = (fabsf(th.x) + fabsf(th.y) + fabsf(th.z)) < CUDART_INF_F; // false for NaN and inf
                if (has_background || !finite) splat(a, c.pixel, mk3(th.x, th.y, th.z), rad);
            }
        }
        __syncwarp();
        warp_sort_push<(1u << (Q_COUNT + 2)) - 1u>(a.pool, dest, slot);
    }
}

// ------------------------------------------------------------------------------------------
// shade: one instantiation per material family
// ------------------------------------------------------------------------------------------
template <int MAT> __global__ void __launch_bounds__(kShadeThreads) k_shade(const __grid_constant__ RenderArgs a, const int parity, const int perlin_in_smem) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ DLight s_lights[MAT == RT1W_MAT_LAMBERTIAN ? RT1W_MAX_LIGHTS : 1];
    Counters *ctr = a.pool.ctr;
    const uint32_t n = ctr->n[QS_MAT + MAT];
    if (n == 0) return;
    // stage the Perlin tables (perlin.rs:7-12) and the light list in shared memory
    const DPerlin *perlins = a.sc.perlins;
    constexpr bool kTextured = MAT == RT1W_MAT_LAMBERTIAN || MAT == RT1W_MAT_ISOTROPIC || MAT == RT1W_MAT_DIFFUSE_LIGHT;
    if (kTextured && perlin_in_smem) {
        const uint32_t words = uint32_t(a.sc.n_perlins) * uint32_t(sizeof(DPerlin) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.sc.perlins);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_dyn);
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
        perlins = reinterpret_cast<const DPerlin *>(s_dyn);
    }
    if (MAT == RT1W_MAT_LAMBERTIAN) {
        for (int l = threadIdx.x; l < a.sc.n_lights; l += blockDim.x) s_lights[l] = a.sc.lights[l];
    }
    __syncthreads();

    const uint32_t *in_q = a.pool.q[QS_MAT + MAT];
    for (uint32_t i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        uint32_t slot = 0;
        bool go_on = false, ended = false;
        if (valid) {
            slot = in_q[i];
            RayC c;
            const Ray r = load_ray(a.pool, slot, c);
            const HitRec hr = a.pool.hit[slot];
            const float4 th4 = a.pool.thr[slot];
            f3 thr = mk3(th4.x, th4.y, th4.z);
            const HitInfo h = finalize_hit<false>(a.sc, hr.leaf, r, hr.t);
            const DMaterial m = a.sc.materials[h.meta >> 12];
            const uint32_t depth = c.state & 255u;
            if (MAT == RT1W_MAT_DIFFUSE_LIGHT) { // material.rs:168-181: emits on the front face only, never scatters (main.rs:110-112)
                const f3 e = h.front_face ? texture_value(a.sc, perlins, m.texture, h) : mk3(0.0f, 0.0f, 0.0f);
                splat(a, c.pixel, thr, e);
                ended = true;
            } else {
                Rng rng;
                path_rng_key(a.rp, c.pixel, rng.k0, rng.k1);
                rng.c0 = uint32_t(a.rp.sample_begin) + (c.state >> 8), rng.c1 = depth, rng.c2 = purpose_word(a.rp, RNG_SCATTER), rng.block = 0;
                f3 dir;
                float time = r.time; // specular scatters keep ray.time (material.rs:104,157; constant_medium.rs:46)
                if (MAT == RT1W_MAT_LAMBERTIAN) {
                    const f3 att = texture_value(a.sc, perlins, m.texture, h);
                    f3 weight;
                    dir = scatter_lambertian(a.sc, s_lights, h, rng, weight);
                    thr = thr * att * weight;
                    time = float(hr.t); // main.rs:86,145: the scattered ray's time is the hit parameter t
                } else if (MAT == RT1W_MAT_METAL) {
                    dir = scatter_metal(m, r, h, rng);
                    thr = thr * mk3(m.albedo[0], m.albedo[1], m.albedo[2]);
                } else if (MAT == RT1W_MAT_DIELECTRIC) {
                    dir = scatter_dielectric(m, r, h, rng); // attenuation (1,1,1)
                } else {                                    // Isotropic, constant_medium.rs:36-51
                    thr = thr * texture_value(a.sc, perlins, m.texture, h);
                    dir = random_in_unit_sphere(rng);
                }
                if (depth + 1u >= uint32_t(a.rp.max_depth)) { // main.rs:59-61: the next ray_color call returns black
                    const bool finite = (fabsf(thr.x) + fabsf(thr.y) + fabsf(thr.z)) < CUDART_INF_F; // false for NaN and inf
                    if (!finite) splat(a, c.pixel, thr, mk3(0.0f, 0.0f, 0.0f));
                    ended = true;
                } else {
                    store_ray(a.pool, slot, h.px, h.py, h.pz, dir, time, c.state + 1u, c.pixel);
                    a.pool.thr[slot] = make_float4(thr.x, thr.y, thr.z, 0.0f);
                    go_on = true;
                }
            }
        }
        __syncwarp();
        warp_sort_push<((1u << QS_FREE) | (1u << (QS_FREE + 1)) | (1u << QS_EXTEND) | (1u << (QS_EXTEND + 1)))>(
            a.pool, go_on ? QS_EXTEND + (parity ^ 1) : (ended ? QS_FREE + parity : -1), slot);
    }
}

// ------------------------------------------------------------------------------------------
// closest-hit parity kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kExtendThreads) k_trace(const __grid_constant__ SceneView sc, const rt1w_ray *__restrict__ rays, const size_t n,
                                                          const uint32_t seed_lo, const uint32_t seed_hi, int32_t *prim_id, float *t_out,
                                                          float *normal3, uint8_t *front_face, float *uv2) {
    __shared__ uint2 s_stack[kStackSmem * kExtendThreads];
    __shared__ FlatScene s_flat;
    const bool flat = sc.n_prims <= kFlatMax; // same choice as the render path, so parity covers both traversals
    if (flat) {
        flat_stage(sc, s_flat);
        __syncthreads();
    }
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const rt1w_ray in = rays[i];
        Ray r;
        r.ox = in.origin[0], r.oy = in.origin[1], r.oz = in.origin[2];
        r.dx = in.direction[0], r.dy = in.direction[1], r.dz = in.direction[2];
        r.time = in.time;
        MediumRng mr;
        mr.c0 = uint32_t(i), mr.c1 = uint32_t(uint64_t(i) >> 32), mr.c2 = RNG_TRACE_MEDIUM, mr.k0 = seed_lo, mr.k1 = seed_hi;
        double t;
        int leaf;
        const bool hit = flat ? closest_hit_flat<true>(sc, s_flat, r, mr, t, leaf) : closest_hit<true>(sc, r, mr, s_stack + threadIdx.x, kExtendThreads, t, leaf);
        HitInfo h;
        if (hit) h = finalize_hit<true>(sc, leaf, r, t);
        if (prim_id) prim_id[i] = hit ? sc.prim_id[leaf] : -1;
        if (t_out) t_out[i] = hit ? float(t) : CUDART_INF_F;
        if (normal3) {
            normal3[3 * i] = hit ? h.normal.x : 0.0f, normal3[3 * i + 1] = hit ? h.normal.y : 0.0f, normal3[3 * i + 2] = hit ? h.normal.z : 0.0f;
        }
        if (front_face) front_face[i] = hit ? uint8_t(h.front_face) : uint8_t(0);
        if (uv2) uv2[2 * i] = hit ? h.u : 0.0f, uv2[2 * i + 1] = hit ? h.v : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
cudaError_t pool_alloc(Pool &pool, uint32_t capacity) {
    pool_free(pool);
    cudaError_t e;
#define RT1W_TRY(x)                                                                                                                                   \
    if ((e = (x)) != cudaSuccess) {                                                                                                                   \
        pool_free(pool);                                                                                                                              \
        return e;                                                                                                                                     \
    }
    RT1W_TRY(cudaMalloc(&pool.o_xy, sizeof(double2) * size_t(capacity)));
    RT1W_TRY(cudaMalloc(&pool.o_zd, sizeof(RayB) * size_t(capacity)));
    RT1W_TRY(cudaMalloc(&pool.dzm, sizeof(RayC) * size_t(capacity)));
    RT1W_TRY(cudaMalloc(&pool.thr, sizeof(float4) * size_t(capacity)));
    RT1W_TRY(cudaMalloc(&pool.hit, sizeof(HitRec) * size_t(capacity)));
    for (int q = 0; q < QS_COUNT; ++q) RT1W_TRY(cudaMalloc(&pool.q[q], sizeof(uint32_t) * size_t(capacity)));
    RT1W_TRY(cudaMalloc(&pool.ctr, sizeof(Counters)));
#undef RT1W_TRY
    pool.capacity = capacity;
    return cudaSuccess;
}

void pool_free(Pool &pool) {
    cudaFree(pool.o_xy), cudaFree(pool.o_zd), cudaFree(pool.dzm), cudaFree(pool.thr), cudaFree(pool.hit);
    for (int q = 0; q < QS_COUNT; ++q) cudaFree(pool.q[q]);
    cudaFree(pool.ctr);
    pool = Pool();
}

namespace {

template <int MAT> void launch_shade(const RenderArgs &args, int parity, int blocks, size_t smem, int perlin_in_smem, cudaStream_t stream) {
    k_shade<MAT><<<blocks, kShadeThreads, smem, stream>>>(args, parity, perlin_in_smem);
}

} // namespace

cudaError_t render_waves(const RenderArgs &args, int material_mask, Counters *h_ctr, cudaStream_t stream, int sm_count, WaveStats &ws) {
    cudaError_t e = cudaMemsetAsync(args.pool.ctr, 0, sizeof(Counters), stream);
    if (e != cudaSuccess) return e;
    // grids: a fixed multiple of the SM count; every kernel grid-strides over a device-side count
    const int gen_blocks = sm_count * 4, ext_blocks = sm_count * 8, shade_blocks = sm_count * 8;
    size_t perlin_bytes = size_t(args.sc.n_perlins) * sizeof(DPerlin);
    const int perlin_in_smem = perlin_bytes > 0 && perlin_bytes <= 40 * 1024;
    if (!perlin_in_smem) perlin_bytes = 0;
    const int poll_every = 8;
    const bool flat = args.sc.n_prims <= kFlatMax;

    // profiling mode: one event before every launch and one at the end of the chunk
    struct Mark {
        cudaEvent_t ev;
        int slot;
    };
    std::vector<Mark> marks;
    std::vector<cudaEvent_t> spare;
    auto mark = [&](int slot) {
        if (!ws.profile) return;
        cudaEvent_t ev;
        if (!spare.empty()) ev = spare.back(), spare.pop_back();
        else cudaEventCreate(&ev);
        cudaEventRecord(ev, stream);
        marks.push_back(Mark{ev, slot});
    };
    auto drain_marks = [&]() {
        for (size_t i = 0; i + 1 < marks.size(); ++i) {
            if (marks[i].slot < 0) continue;
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, marks[i].ev, marks[i + 1].ev);
            ws.kernel_ms[marks[i].slot] += ms, ws.kernel_launches[marks[i].slot] += 1;
        }
        for (auto &m : marks) spare.push_back(m.ev);
        marks.clear();
    };

    uint64_t wave = 0;
    for (;;) {
        for (int k = 0; k < poll_every; ++k, ++wave) {
            const int parity = int(wave & 1);
            mark(K_GENERATE);
            k_generate<<<gen_blocks, kGenThreads, 0, stream>>>(args, parity, wave == 0 ? 1 : 0);
            mark(K_EXTEND);
            if (flat) k_extend<true><<<ext_blocks, kExtendThreads, 0, stream>>>(args, parity, material_mask);
            else k_extend<false><<<ext_blocks, kExtendThreads, 0, stream>>>(args, parity, material_mask);
            ws.launches += 2;
#define RT1W_SHADE(MAT, SMEM, PSM)                                                                                                                    \
    if (material_mask & (1 << MAT)) {                                                                                                                 \
        mark(K_SHADE0 + MAT);                                                                                                                         \
        launch_shade<MAT>(args, parity, shade_blocks, SMEM, PSM, stream);                                                                             \
        ++ws.launches;                                                                                                                                \
    }
            RT1W_SHADE(RT1W_MAT_LAMBERTIAN, perlin_bytes, perlin_in_smem)
            RT1W_SHADE(RT1W_MAT_METAL, 0, 0)
            RT1W_SHADE(RT1W_MAT_DIELECTRIC, 0, 0)
            RT1W_SHADE(RT1W_MAT_ISOTROPIC, perlin_bytes, perlin_in_smem)
            RT1W_SHADE(RT1W_MAT_DIFFUSE_LIGHT, perlin_bytes, perlin_in_smem)
#undef RT1W_SHADE
        }
        mark(-1);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(h_ctr, args.pool.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        drain_marks();
        // after wave (wave-1): continuing paths sit in q_extend[wave & 1]
        if (h_ctr->next_path >= args.rp.total_paths && h_ctr->n[QS_EXTEND + (wave & 1)] == 0) break;
    }
    for (auto &m : marks) cudaEventDestroy(m.ev);
    for (auto ev : spare) cudaEventDestroy(ev);
    if (e != cudaSuccess) return e;
    ws.waves = wave;
    ws.rays = h_ctr->rays;
    return cudaSuccess;
}

cudaError_t trace_closest_launch(const SceneView &sc, const rt1w_ray *rays, size_t n, uint64_t seed, int32_t *prim_id, float *t,
                                 float *normal3, uint8_t *front_face, float *uv2, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const size_t want = (n + kExtendThreads - 1) / kExtendThreads;
    const int blocks = int(want < 148 * 16 ? want : 148 * 16);
    k_trace<<<blocks, kExtendThreads, 0, stream>>>(sc, rays, n, uint32_t(seed), uint32_t(seed >> 32), prim_id, t, normal3, front_face, uv2);
    return cudaGetLastError();
}

} // namespace rt1w
