// device_types.h — POD layouts shared by the host lowering code and the sm_100a kernels.
//
// HBM layout of a committed scene (all arrays read-only during a render, small enough to
// stay L1/L2 resident for every BASELINE config):
//   nodes   : 2 x float4 per BVH node (32 B).  n0 = (min.xyz, bits leftFirst), n1 = (max.xyz, bits count).
//             count == 0 -> interior, children at leftFirst and leftFirst+1 (a 64-B aligned pair);
//             count  > 0 -> leaf, primitives [leftFirst, leftFirst+count) in LEAF order.
//   prims   : 4 x float4 per primitive (64 B) in leaf order, see DPrim (spheres/rects use the first 32 B).
//   prims64 : 10 doubles per primitive (f64 parameters; read only by the f64 confirm path of
//             the sphere tests and by rt1w_scene_get_prims).
//   prim_id : leaf index -> primitive id (DFS order of the description).
//   frames, materials, textures, perlin tables, image texture objects, lights: small tables.
#pragma once

#include <stdint.h>

namespace rt1w {

enum PrimType : int {
    P_SPHERE = 0,
    P_MOVING_SPHERE = 1,
    P_XY_RECT = 2,
    P_XZ_RECT = 3,
    P_YZ_RECT = 4,
    P_MEDIUM_SPHERE = 5,
    P_MEDIUM_BOX = 6,
};

enum PrimFlags : int { PF_FLIP_FACE = 1 };

// Shading queues of the wavefront (one per reference material family + a terminal queue
// for misses, DiffuseLight and the null material).
enum QueueId : int { Q_LAMBERTIAN = 0, Q_METAL = 1, Q_DIELECTRIC = 2, Q_ISOTROPIC = 3, Q_TERMINAL = 4, Q_COUNT = 5 };

// meta word: type[0:4) | flags[4:8) | material type[8:12) | material id[12:32)
static inline uint32_t pack_meta(int type, int flags, int mat_type, int mat_id) {
    return uint32_t(type) | (uint32_t(flags) << 4) | (uint32_t(mat_type) << 8) | (uint32_t(mat_id) << 12);
}

// Host view of one 64-byte device primitive (uploaded as 4 x float4).  Words 6 and 7 hold the
// meta word and the frame id so that spheres and rects are fully described by the first 32 bytes:
//   float4 #0 = p0..p3 | float4 #1 = p4, p5, meta, frame | float4 #2 = p6..p9 | float4 #3 = p10..p13
//   sphere        : p0..2 = center, p3 = radius
//   rects         : p0,p1 = first in-plane interval, p2,p3 = second, p4 = k
//   moving sphere : p0..2 = center(T0), p3 = radius, p4..6 = center(T1)-center(T0), p7 = T0, p8 = 1/(T1-T0)
//                   with (T0,T1) the bounding_box(time0,time1) arguments in scope (bvh.rs:54-59), so that
//                   the [T0,T1] box of moving_sphere.rs:72-84 is min/max(p0..2, p0..2+p4..6) -/+ radius
//   medium sphere : p0..2 = center, p3 = radius, p4 = -1/density
//   medium box    : p0..2 = box min, p3 = -1/density, p4..6 = box max
struct DPrim {
    float p03[4];
    float p4, p5;
    uint32_t meta;
    int32_t frame; // -1: no wrapper chain
    float p69[4];
    float p1013[4];
};
static_assert(sizeof(DPrim) == 64, "DPrim must be 4 x float4");

// One wrapper chain (Translate / RotateY / FlipFace stack above a leaf).
// local = Ry(p) + b with Ry: x' = c*x - s*z, z' = s*x + c*z (hittable.rs:241-245 composed with :207).
// ops replay the HitRecord::new calls of the wrappers, innermost first (hittable.rs:221-229, :269-277, :290).
#define RT1W_MAX_CHAIN_OPS 8
enum ChainOpKind : int { OP_TRANSLATE = 0, OP_ROTATE_Y = 1, OP_FLIP_FACE = 2 };
struct DChainOp {
    int32_t kind;
    float sin_own, cos_own; // this RotateY's own angle (normal back-rotation)
    float sin_cum, cos_cum; // cumulative angle of the ray direction INSIDE this wrapper
};
struct DFrame {
    float sin_t, cos_t;
    float bx, by, bz;
    int32_t n_ops;
    int32_t pad0, pad1;
    DChainOp ops[RT1W_MAX_CHAIN_OPS];
};

struct DMaterial { // 2 x float4
    float albedo[3];
    float fuzz;
    float ir;
    int32_t type;
    int32_t texture;
    int32_t pad;
};
static_assert(sizeof(DMaterial) == 32, "DMaterial must be 2 x float4");

struct DTexture { // 2 x float4
    float color[3];
    float scale;
    int32_t type;
    int32_t odd, even;
    int32_t table;
};
static_assert(sizeof(DTexture) == 32, "DTexture must be 2 x float4");

struct DPerlin {
    float ranvec[256][4]; // xyz + pad, 16-B aligned rows
    uint8_t perm[3][256];
};

enum LightKind : int { L_XZ_RECT = 0, L_SPHERE = 1, L_OTHER = 2 };
struct DLight { // pdf.rs / hittable.rs:144-154: only XZRect and Sphere implement pdf_value/random
    int32_t kind;
    float p[5]; // XZ: x0,x1,z0,z1,k ; sphere: cx,cy,cz,r
    float pad[2];
};
#define RT1W_MAX_LIGHTS 32

struct DCamera {
    float origin[3];
    float llc_rel[3]; // lower_left_corner - origin, differenced in f64 on the host
    float horizontal[3];
    float vertical[3];
    float u[3], v[3];
    float lens_radius;
    float time0, time1;
};

struct DRenderParams {
    int32_t width, height;
    int32_t sample_begin, n_samples;
    int32_t max_depth;
    uint32_t flags;
    uint32_t seed_lo, seed_hi;
    float background[3];
    float stat_clamp;
    int32_t has_lights, n_lights;
    float inv_wm1, inv_hm1;
    unsigned long long total_paths;
    int32_t pool;
    int32_t has_perlin;
};

// Philox counter "stream" words (counter[2]); counter = {sample, bounce, stream, block}.
enum RngStream : uint32_t { RNG_CAMERA = 0, RNG_SCATTER = 1, RNG_MEDIUM = 2, RNG_TRACE_MEDIUM = 0x4d454449u };

} // namespace rt1w
