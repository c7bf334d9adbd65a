/* host_api.h — C entry points of the host-side mirror (librt1w_host.so).
 *
 * These sit ABOVE the drop-in boundary (include/rt1w.h): they run the C++
 * mirror of the reference's scene functions (main.rs:192-795) and `match` arms
 * (main.rs:815-937) and hand back the POD scene description that
 * rt1w_scene_create consumes.  Python tests and bench.py reach the scene
 * functions through these; C++ callers include scenes.hpp directly.
 */
#ifndef RT1W_HOST_API_H
#define RT1W_HOST_API_H
#include "../../include/rt1w.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt1w_host_scene rt1w_host_scene;

typedef struct rt1w_host_settings { /* the per-arm locals of main (main.rs:798-801,815-937) */
    int32_t image_width, image_height, samples_per_pixel, max_depth;
    double aspect_ratio, aperture, vfov_deg;
    double background[3], look_from[3], look_at[3];
} rt1w_host_settings;

/* which: 0..6 = the match arms, 7 = default arm (final_scene), 8 = stress (C5), 9 = one-weekend variant.
 * earth_rgb8 may be NULL for scenes without the image texture. Returns NULL on error. */
rt1w_host_scene *rt1w_host_scene_build(int32_t which, uint64_t seed, const uint8_t *earth_rgb8, int32_t earth_w,
                                       int32_t earth_h, int32_t stress_spheres);
int32_t rt1w_host_scene_id(const char *name);
const rt1w_scene_desc *rt1w_host_scene_desc(const rt1w_host_scene *s);
void rt1w_host_scene_settings(const rt1w_host_scene *s, rt1w_host_settings *out);
/* Camera::new with this scene's look_from/look_at/vfov/aperture and the given aspect (main.rs:940-951). */
void rt1w_host_scene_camera(const rt1w_host_scene *s, double aspect_ratio, rt1w_camera *out);
void rt1w_host_scene_free(rt1w_host_scene *s);
/* Camera::new (camera.rs:22-59). */
void rt1w_host_camera_new(const double look_from[3], const double look_at[3], const double vup[3], double vfov_deg,
                          double aspect_ratio, double aperture, double focus_dist, double time0, double time1,
                          rt1w_camera *out);
/* The PPM P3 text of main.rs:953,1003-1007 from quantised pixels (row 0 = top). Returns 0 on success. */
int32_t rt1w_host_write_ppm(const char *path, const uint8_t *rgb8, int32_t width, int32_t height);
const char *rt1w_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
