// host_api.cpp — C wrappers around the C++ mirror API (see host_api.h).
#include "host_api.h"

#include <cstdio>
#include <string>

#include "scenes.hpp"

struct rt1w_host_scene {
    std::unique_ptr<rt1w::SceneSetup> setup;
    rt1w_scene_desc desc;
};

static thread_local std::string g_host_error;

extern "C" {

rt1w_host_scene *rt1w_host_scene_build(int32_t which, uint64_t seed, const uint8_t *earth_rgb8, int32_t earth_w,
                                       int32_t earth_h, int32_t stress_spheres) {
    try {
        rt1w::EarthMap map;
        if (earth_rgb8 && earth_w > 0 && earth_h > 0) {
            map.rgb8 = std::make_shared<std::vector<uint8_t>>(earth_rgb8, earth_rgb8 + size_t(earth_w) * earth_h * 3);
            map.width = earth_w, map.height = earth_h;
        }
        auto h = new rt1w_host_scene();
        h->setup = rt1w::select_scene(which, seed, map, stress_spheres > 0 ? stress_spheres : 1000000);
        h->desc = h->setup->builder.desc();
        return h;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return nullptr;
    }
}

int32_t rt1w_host_scene_id(const char *name) { return name ? rt1w::scene_id_from_name(name) : -1; }

const rt1w_scene_desc *rt1w_host_scene_desc(const rt1w_host_scene *s) { return s ? &s->desc : nullptr; }

void rt1w_host_scene_settings(const rt1w_host_scene *s, rt1w_host_settings *out) {
    if (!s || !out) return;
    const rt1w::SceneSetup &u = *s->setup;
    out->image_width = u.image_width, out->image_height = u.image_height();
    out->samples_per_pixel = u.samples_per_pixel, out->max_depth = u.max_depth;
    out->aspect_ratio = u.aspect_ratio, out->aperture = u.aperture, out->vfov_deg = u.vfov.value;
    out->background[0] = u.background.v.x, out->background[1] = u.background.v.y, out->background[2] = u.background.v.z;
    out->look_from[0] = u.look_from.x, out->look_from[1] = u.look_from.y, out->look_from[2] = u.look_from.z;
    out->look_at[0] = u.look_at.x, out->look_at[1] = u.look_at.y, out->look_at[2] = u.look_at.z;
}

void rt1w_host_scene_camera(const rt1w_host_scene *s, double aspect_ratio, rt1w_camera *out) {
    if (!s || !out) return;
    *out = s->setup->camera_for_aspect(aspect_ratio).pod;
}

void rt1w_host_scene_free(rt1w_host_scene *s) { delete s; }

void rt1w_host_camera_new(const double look_from[3], const double look_at[3], const double vup[3], double vfov_deg,
                          double aspect_ratio, double aperture, double focus_dist, double time0, double time1,
                          rt1w_camera *out) {
    using namespace rt1w;
    *out = Camera::new_(point3(look_from[0], look_from[1], look_from[2]), point3(look_at[0], look_at[1], look_at[2]),
                        vec3(vup[0], vup[1], vup[2]), Deg{vfov_deg}, aspect_ratio, aperture, focus_dist, time0, time1)
               .pod;
}

int32_t rt1w_host_write_ppm(const char *path, const uint8_t *rgb8, int32_t width, int32_t height) {
    FILE *f = (path && std::string(path) != "-") ? std::fopen(path, "w") : stdout;
    if (!f) {
        g_host_error = "cannot open output file";
        return 1;
    }
    std::fprintf(f, "P3\n%d %d\n255\n", width, height); // main.rs:953
    const size_t n = size_t(width) * height;
    for (size_t i = 0; i < n; ++i) std::fprintf(f, "%d %d %d\n", rgb8[3 * i], rgb8[3 * i + 1], rgb8[3 * i + 2]); // main.rs:1003-1007
    if (f != stdout) std::fclose(f);
    return 0;
}

const char *rt1w_host_last_error(void) { return g_host_error.c_str(); }
}
