/* oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE, never shipped).
 *
 * The oracle is a C++17 / f64 restatement of the reference's per-pixel sample loop
 * (src/main.rs:51-190,957-1001 and every module it calls).  Only tests/, the smoke
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  It consumes the same POD scene description as the product
 * (include/rt1w.h) but shares no code with it.
 */
#ifndef RT1W_ORACLE_H
#define RT1W_ORACLE_H
#include "../include/rt1w.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_scene oracle_scene;

typedef struct oracle_render_stats {
    uint64_t paths;
    uint64_t rays; /* closest-hit queries: calls of world.hit from ray_color (main.rs:62) */
    double seconds;
    int32_t threads;
    int32_t reserved;
} oracle_render_stats;

const char *oracle_last_error(void);

/* Builds the reference object tree: nested BVHNode::new with random axis + median split
 * (bvh.rs:54-103, axis RNG = StdRng::seed_from_u64(bvh_seed)), AABox with its inner BVH
 * (aabox.rs:22-84), wrappers as wrappers. */
oracle_scene *oracle_scene_load(const rt1w_scene_desc *desc, uint64_t bvh_seed);
void oracle_scene_free(oracle_scene *s);
int32_t oracle_scene_num_prims(const oracle_scene *s);
/* BVHNode count and depth of the world tree (bvh.rs:60-101), AABox inner trees excluded. */
void oracle_scene_bvh_shape(const oracle_scene *s, int32_t *n_nodes, int32_t *depth);

/* world.hit(ray, 0.001, inf) for n rays (f32 inputs widened to f64).  Media replay the
 * device's keyed Philox free-flight draw.  ambiguous[i] != 0 on genuine (near-)ties only: two
 * leaves within 1e-9 relative of the winning t (e.g. the Cornell box bottom on the floor — the
 * reference's winner depends on its random BVH order there, bvh.rs:39-46,84), a candidate within
 * 1e-9 of t_min / t_max, or a winner that changes under a 1e-9-relative perturbation of the ray. */
int32_t oracle_trace_closest(const oracle_scene *s, const rt1w_ray *rays, size_t n, uint64_t seed, int32_t *prim_id,
                             double *t, double *normal3, uint8_t *front_face, double *uv2, uint8_t *ambiguous,
                             int32_t threads);

/* The pixel loop (main.rs:957-1001) for samples [sample_begin, sample_end).
 * rgb_sum: width*height*3 doubles, row 0 = top.  stat (nullable): width*height*6 doubles
 * (sum and sum of squares of min(sample, stat_clamp), NaN -> 0).  The per-pixel RNG is
 * StdRng::seed_from_u64(j*width+i) advanced past the first sample_begin samples' worth of
 * draws only when sample_begin == 0 (otherwise seeded with a mixed key). threads<=0: all cores. */
int32_t oracle_render(const oracle_scene *s, const rt1w_camera *camera, const rt1w_render_params *params,
                      int32_t threads, double *rgb_sum, double *stat, oracle_render_stats *stats);

/* ---- known-answer probes ---- */
double oracle_xz_rect_pdf_value(const double rect[5], const double o[3], const double v[3]);
double oracle_sphere_pdf_value(const double center[3], double radius, const double o[3], const double v[3]);
double oracle_sphere_hit_t(const double center[3], double radius, const double o[3], const double d[3], double t_min, double t_max);
double oracle_reflectance(double cosine, double ref_idx);
void oracle_refract(const double uv[3], const double n[3], double etai_over_etat, double out[3]);
void oracle_sphere_uv(const double p[3], double out_uv[2]);
void oracle_onb_from_w(const double n[3], double out_uvw[9]);
void oracle_quantise(const double rgb_sum[3], int32_t spp, int32_t out[3]);
void oracle_camera_ray(const rt1w_camera *cam, double s, double t, double out_origin[3], double out_dir[3]); /* lens_radius 0 */
void oracle_rotate_y_bbox(const double bmin[3], const double bmax[3], double deg, double out_min[3], double out_max[3]);
/* Closest hit of one f64 ray against a described scene: returns prim id (or -1). */
int32_t oracle_hit_one(const oracle_scene *s, const double o[3], const double d[3], double time, double *t, double p[3],
                       double n[3], int32_t *front_face);
double oracle_perlin_noise(const rt1w_perlin *tab, const double p[3]);
double oracle_perlin_turb(const rt1w_perlin *tab, const double p[3], int32_t depth);
void oracle_texture_value(const oracle_scene *s, int32_t texture, double u, double v, const double p[3], double out[3]);
void oracle_bvh_count(int32_t n_objects, uint64_t seed, int32_t *n_nodes, int32_t *depth);
/* RNG probes */
void oracle_chacha_block(const uint32_t state[16], int32_t rounds, uint32_t out[16]);
void oracle_stdrng_u32(uint64_t seed, int32_t n, uint32_t *out);
void oracle_stdrng_from_seed_u32(const uint8_t seed[32], int32_t n, uint32_t *out);
void oracle_stdrng_f64(uint64_t seed, int32_t n, double *out);
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
