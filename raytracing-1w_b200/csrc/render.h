// render.h — host-side interface of the wavefront renderer (render.cu) used by the C ABI (api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rt1w.h"
#include "device_types.h"
#include "kernels.cuh"

namespace rt1w {

// Device-side counters of the wavefront queues (one struct per context, zeroed per render).
// Queue table: one index space for every queue so that a lane can address "its" queue with a
// register index into kernel-parameter (constant-bank) arrays.
enum QueueSlot : int {
    QS_MAT = 0,             // + rt1w_material_type: hits queued for the shade kernels of the current wave
    QS_FREE = Q_COUNT,      // + (wave & 1): path slots that terminated during wave w
    QS_EXTEND = Q_COUNT + 2, // + (wave & 1): rays queued for the extend kernel of wave w
    QS_COUNT = Q_COUNT + 4
};

struct Counters {
    uint32_t n[QS_COUNT];         // entries per queue
    uint32_t pad[3];
    unsigned long long next_path; // next (pixel, sample) pair to start
    unsigned long long rays;      // closest-hit queries so far
};

// Path pool: SoA ray/path state + index queues, all in HBM (DESIGN.md "Data layout").
struct Pool {
    double2 *o_xy = nullptr; // origin.x, origin.y
    RayB *o_zd = nullptr;    // origin.z, direction.x, direction.y
    RayC *dzm = nullptr;     // direction.z, time, state, pixel
    float4 *thr = nullptr;   // throughput rgb (+ unused lane)
    HitRec *hit = nullptr;   // t, leaf
    uint32_t *q[QS_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // slot-index queues (QueueSlot)
    Counters *ctr = nullptr;
    uint32_t capacity = 0;
};

cudaError_t pool_alloc(Pool &pool, uint32_t capacity);
void pool_free(Pool &pool);

struct RenderArgs {
    SceneView sc;
    Pool pool;
    DRenderParams rp;
    DCamera cam;
    float *accum; // width*height*3 radiance sums
    float *stat;  // width*height*6 clamped sum / sum of squares, or nullptr
};

enum KernelSlot : int { K_GENERATE = 0, K_EXTEND = 1, K_SHADE0 = 2, K_COUNT = 2 + Q_COUNT }; // K_SHADE0 + rt1w_material_type

struct WaveStats {
    uint64_t waves = 0, launches = 0, rays = 0;
    bool profile = false;        // RT1W_FLAG_PROFILE: bracket every launch with CUDA events
    double kernel_ms[K_COUNT] = {0};
    uint64_t kernel_launches[K_COUNT] = {0};
};

// Runs the wave loop to completion on `stream` (accum/stat must be zeroed by the caller).
// material_mask: bit m set when some primitive uses rt1w_material_type m.
// h_ctr: pinned host mirror of the counters used for the termination poll.
cudaError_t render_waves(const RenderArgs &args, int material_mask, Counters *h_ctr, cudaStream_t stream, int sm_count, WaveStats &ws);

// Closest-hit parity kernel (rt1w_trace_closest).  All pointers are device pointers; outputs may be null.
cudaError_t trace_closest_launch(const SceneView &sc, const rt1w_ray *rays, size_t n, uint64_t seed, int32_t *prim_id, float *t,
                                 float *normal3, uint8_t *front_face, float *uv2, cudaStream_t stream);

} // namespace rt1w
