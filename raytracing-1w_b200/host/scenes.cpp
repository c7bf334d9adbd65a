// scenes.cpp — see scenes.hpp.  Values follow main.rs:192-795 / 815-937.
#include "scenes.hpp"

#include <cstring>

namespace rt1w {
namespace {

TextureRef solid(Float r, Float g, Float b) { return std::make_shared<SolidColor>(Color(vec3(r, g, b))); }
MaterialRef lambertian(TextureRef t) { return std::make_shared<Lambertian>(std::move(t)); }
MaterialRef lambertian(Float r, Float g, Float b) { return lambertian(solid(r, g, b)); }
MaterialRef metal(Float r, Float g, Float b, Float fuzz) { return std::make_shared<Metal>(Color(vec3(r, g, b)), fuzz); }
MaterialRef glass(Float ir) { return std::make_shared<Dielectric>(ir); }
MaterialRef emitter(Float r, Float g, Float b) { return std::make_shared<DiffuseLight>(solid(r, g, b)); }

template <class T, class... A> HittableBox boxed(A &&...a) { return std::make_unique<T>(std::forward<A>(a)...); }

TextureRef ground_checker() { // main.rs:193-202, 298-307: odd = (0.9,0.9,0.9), even = (0.2,0.3,0.1)
    return std::make_shared<CheckerTexture>(solid(0.9, 0.9, 0.9), solid(0.2, 0.3, 0.1));
}

TextureRef earth_texture(const EarthMap &map) {
    if (!map.rgb8 || map.width <= 0 || map.height <= 0) throw std::runtime_error("this scene needs the decoded earthmap image");
    return std::make_shared<DynamicImage>(map.rgb8, map.width, map.height);
}

// The five walls + ceiling light shared by cornel_box / cornel_smoke (main.rs:451-494, 579-622).
void push_cornell_room(std::vector<HittableBox> &world, Float lx0, Float lx1, Float lz0, Float lz1, const MaterialRef &light) {
    MaterialRef red = lambertian(0.65, 0.05, 0.05);
    MaterialRef white = lambertian(0.73, 0.73, 0.73);
    MaterialRef green = lambertian(0.12, 0.45, 0.15);
    world.push_back(boxed<YZRect>(0.0, 555.0, 0.0, 555.0, 555.0, green));
    world.push_back(boxed<YZRect>(0.0, 555.0, 0.0, 555.0, 0.0, red));
    world.push_back(boxed<FlipFace>(boxed<XZRect>(lx0, lx1, lz0, lz1, 554.0, light)));
    world.push_back(boxed<XZRect>(0.0, 555.0, 0.0, 555.0, 0.0, white));
    world.push_back(boxed<XZRect>(0.0, 555.0, 0.0, 555.0, 555.0, white));
    world.push_back(boxed<XYRect>(0.0, 555.0, 0.0, 555.0, 555.0, white));
}

HittableBox placed_box(Vec3 extent, const MaterialRef &m, Float degrees, Vec3 offset, SceneRng &rng) { // main.rs:425-435
    HittableBox bx = AABox::new_(point3(0.0, 0.0, 0.0), extent, m, rng);
    HittableBox rot = RotateY::new_(std::move(bx), 0.0, 1.0, Deg{degrees});
    return boxed<Translate>(std::move(rot), offset);
}

} // namespace

std::unique_ptr<BVHNode> random_scene(SceneRng &rng) {
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, -1000.0, 0.0), 1000.0, lambertian(ground_checker())));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            Float choose_mat = rng.gen();
            Float cx = Float(a) + 0.9 * rng.gen();
            Float cz = Float(b) + 0.9 * rng.gen();
            Vec3 center = point3(cx, 0.2, cz);
            if (magnitude(center - point3(4.0, 0.2, 0.0)) <= 0.9) continue;
            if (choose_mat < 0.8) { // diffuse, moving (main.rs:220-237)
                Color c0 = rng.gen_color();
                Color c1 = rng.gen_color();
                Vec3 albedo = mul_element_wise(c0.v, c1.v);
                Vec3 center2 = center + vec3(0.0, rng.gen_range(0.0, 0.5), 0.0);
                world.push_back(boxed<MovingSphere>(center, center2, 0.0, 1.0, 0.2, lambertian(albedo.x, albedo.y, albedo.z)));
            } else if (choose_mat < 0.95) { // metal (main.rs:239-253)
                Float r = rng.gen_range(0.5, 1.0), g = rng.gen_range(0.5, 1.0), bl = rng.gen_range(0.5, 1.0);
                Float fuzz = rng.gen_range(0.5, 1.0);
                world.push_back(boxed<Sphere>(center, 0.2, metal(r, g, bl, fuzz)));
            } else { // glass (main.rs:255-263)
                world.push_back(boxed<Sphere>(center, 0.2, glass(1.5)));
            }
        }
    }
    world.push_back(boxed<Sphere>(point3(0.0, 1.0, 0.0), 1.0, glass(1.5)));
    world.push_back(boxed<Sphere>(point3(-4.0, 1.0, 0.0), 1.0, lambertian(0.4, 0.2, 0.1)));
    world.push_back(boxed<Sphere>(point3(4.0, 1.0, 0.0), 1.0, metal(0.7, 0.6, 0.5, 0.0)));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> random_scene_one_weekend(SceneRng &rng) {
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, -1000.0, 0.0), 1000.0, lambertian(0.5, 0.5, 0.5)));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            Float choose_mat = rng.gen();
            Float cx = Float(a) + 0.9 * rng.gen();
            Float cz = Float(b) + 0.9 * rng.gen();
            Vec3 center = point3(cx, 0.2, cz);
            if (magnitude(center - point3(4.0, 0.2, 0.0)) <= 0.9) continue;
            if (choose_mat < 0.8) {
                Color c0 = rng.gen_color();
                Color c1 = rng.gen_color();
                Vec3 albedo = mul_element_wise(c0.v, c1.v);
                world.push_back(boxed<Sphere>(center, 0.2, lambertian(albedo.x, albedo.y, albedo.z)));
            } else if (choose_mat < 0.95) {
                Float r = rng.gen_range(0.5, 1.0), g = rng.gen_range(0.5, 1.0), bl = rng.gen_range(0.5, 1.0);
                Float fuzz = rng.gen_range(0.0, 0.5);
                world.push_back(boxed<Sphere>(center, 0.2, metal(r, g, bl, fuzz)));
            } else {
                world.push_back(boxed<Sphere>(center, 0.2, glass(1.5)));
            }
        }
    }
    world.push_back(boxed<Sphere>(point3(0.0, 1.0, 0.0), 1.0, glass(1.5)));
    world.push_back(boxed<Sphere>(point3(-4.0, 1.0, 0.0), 1.0, lambertian(0.4, 0.2, 0.1)));
    world.push_back(boxed<Sphere>(point3(4.0, 1.0, 0.0), 1.0, metal(0.7, 0.6, 0.5, 0.0)));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> two_spheres(SceneRng &rng) {
    MaterialRef checker_material = lambertian(ground_checker());
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, -10.0, 0.0), 10.0, checker_material));
    world.push_back(boxed<Sphere>(point3(0.0, 10.0, 0.0), 10.0, checker_material));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> two_perlin_spheres(SceneRng &rng) {
    MaterialRef pertext = lambertian(NoiseTexture256::new_(4.0, rng));
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, -1000.0, 0.0), 1000.0, pertext));
    world.push_back(boxed<Sphere>(point3(0.0, 2.0, 0.0), 2.0, pertext));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> earth(SceneRng &rng, const EarthMap &map) {
    MaterialRef earth_surface = lambertian(earth_texture(map));
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, 0.0, 0.0), 2.0, earth_surface));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> simple_light(SceneRng &rng) {
    MaterialRef pertext = lambertian(NoiseTexture256::new_(4.0, rng));
    MaterialRef difflight = emitter(4.0, 4.0, 4.0);
    std::vector<HittableBox> world;
    world.push_back(boxed<Sphere>(point3(0.0, -1000.0, 0.0), 1000.0, pertext));
    world.push_back(boxed<Sphere>(point3(0.0, 2.0, 0.0), 2.0, pertext));
    world.push_back(boxed<XYRect>(3.0, 5.0, 1.0, 3.0, -2.0, difflight));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> cornel_box(SceneRng &rng) {
    std::vector<HittableBox> world;
    push_cornell_room(world, 213.0, 343.0, 227.0, 332.0, emitter(15.0, 15.0, 15.0));
    MaterialRef aluminum = metal(0.8, 0.85, 0.88, 0.0);
    world.push_back(placed_box(point3(165.0, 330.0, 165.0), aluminum, 15.0, vec3(265.0, 0.0, 295.0), rng));
    // box2 is commented out in the reference (main.rs:437-449,503)
    world.push_back(boxed<Sphere>(point3(190.0, 90.0, 190.0), 90.0, glass(1.5)));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> cornel_smoke(SceneRng &rng) {
    std::vector<HittableBox> world;
    push_cornell_room(world, 113.0, 443.0, 127.0, 432.0, emitter(7.0, 7.0, 7.0));
    MaterialRef white = lambertian(0.73, 0.73, 0.73);
    HittableBox box1 = placed_box(point3(165.0, 330.0, 165.0), white, 15.0, vec3(265.0, 0.0, 295.0), rng);
    HittableBox box2 = placed_box(point3(165.0, 165.0, 165.0), white, -18.0, vec3(130.0, 0.0, 65.0), rng);
    world.push_back(ConstantMedium::new_(std::move(box1), 0.01, solid(0.0, 0.0, 0.0)));
    world.push_back(ConstantMedium::new_(std::move(box2), 0.01, solid(1.0, 1.0, 1.0)));
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> final_scene(SceneRng &rng, const EarthMap &map) {
    MaterialRef ground = lambertian(0.48, 0.83, 0.53);
    const int boxes_per_side = 20;
    std::vector<HittableBox> boxes1;
    for (int i = 0; i < boxes_per_side; ++i) {
        for (int j = 0; j < boxes_per_side; ++j) {
            Float w = 100.0;
            Float x0 = -1000.0 + Float(i) * w;
            Float z0 = -1000.0 + Float(j) * w;
            Float y1 = rng.gen_range(1.0, 101.0);
            boxes1.push_back(AABox::new_(point3(x0, 0.0, z0), point3(x0 + w, y1, z0 + w), ground, rng));
        }
    }
    std::vector<HittableBox> objects;
    objects.push_back(BVHNode::new_(std::move(boxes1), 0.0, 1.0, rng));
    objects.push_back(boxed<FlipFace>(boxed<XZRect>(123.0, 423.0, 147.0, 412.0, 554.0, emitter(7.0, 7.0, 7.0))));

    Vec3 center1 = point3(400.0, 400.0, 200.0);
    Vec3 center2 = center1 + vec3(30.0, 0.0, 0.0);
    objects.push_back(boxed<MovingSphere>(center1, center2, 0.0, 1.0, 50.0, lambertian(0.7, 0.3, 0.1)));
    objects.push_back(boxed<Sphere>(point3(260.0, 150.0, 45.0), 50.0, glass(1.5)));
    objects.push_back(boxed<Sphere>(point3(0.0, 150.0, 145.0), 50.0, metal(0.8, 0.8, 0.9, 1.0)));

    // glass ball with blue participating medium inside (main.rs:720-733)
    objects.push_back(boxed<Sphere>(point3(360.0, 150.0, 145.0), 70.0, glass(1.5)));
    objects.push_back(ConstantMedium::new_(boxed<Sphere>(point3(360.0, 150.0, 145.0), 70.0, glass(1.5)), 0.2, solid(0.2, 0.4, 0.9)));
    // scene-wide thin fog (main.rs:734-745)
    objects.push_back(ConstantMedium::new_(boxed<Sphere>(point3(0.0, 0.0, 0.0), 5000.0, glass(1.5)), 0.0001, solid(1.0, 1.0, 1.0)));

    objects.push_back(boxed<Sphere>(point3(400.0, 200.0, 400.0), 100.0, lambertian(earth_texture(map))));
    objects.push_back(boxed<Sphere>(point3(220.0, 280.0, 300.0), 80.0, lambertian(NoiseTexture256::new_(0.1, rng))));

    std::vector<HittableBox> boxes2;
    MaterialRef white = lambertian(0.73, 0.73, 0.73);
    const int ns = 1000;
    for (int i = 0; i < ns; ++i) {
        Float x = rng.gen_range(0.0, 165.0), y = rng.gen_range(0.0, 165.0), z = rng.gen_range(0.0, 165.0);
        boxes2.push_back(boxed<Sphere>(point3(x, y, z), 10.0, white));
    }
    HittableBox cluster = RotateY::new_(BVHNode::new_(std::move(boxes2), 0.0, 1.0, rng), 0.0, 1.0, Deg{15.0});
    objects.push_back(boxed<Translate>(std::move(cluster), vec3(-100.0, 270.0, 395.0)));
    return BVHNode::new_(std::move(objects), 0.0, 1.0, rng);
}

std::unique_ptr<BVHNode> stress_scene(SceneRng &rng, int n_spheres, std::vector<HittableBox> *lights_out) {
    std::vector<HittableBox> world;
    world.reserve(size_t(n_spheres) + 16);
    MaterialRef glass15 = glass(1.5);
    for (int i = 0; i < n_spheres; ++i) {
        Float x = rng.gen_range(-1000.0, 1000.0), y = rng.gen_range(-1000.0, 1000.0), z = rng.gen_range(-1000.0, 1000.0);
        Float radius = rng.gen_range(0.5, 2.0);
        Float choose_mat = rng.gen();
        MaterialRef m;
        if (choose_mat < 0.8) {
            Color c0 = rng.gen_color();
            Color c1 = rng.gen_color();
            Vec3 a = mul_element_wise(c0.v, c1.v);
            m = lambertian(a.x, a.y, a.z);
        } else if (choose_mat < 0.95) {
            Float r = rng.gen_range(0.5, 1.0), g = rng.gen_range(0.5, 1.0), b = rng.gen_range(0.5, 1.0);
            m = metal(r, g, b, rng.gen_range(0.0, 0.5));
        } else {
            m = glass15;
        }
        world.push_back(boxed<Sphere>(point3(x, y, z), radius, m));
    }
    MaterialRef light = emitter(15.0, 15.0, 15.0);
    MaterialRef null_mat = std::make_shared<NullMaterial>();
    for (int gx = 0; gx < 4; ++gx) {
        for (int gz = 0; gz < 4; ++gz) {
            Float x0 = -1000.0 + 500.0 * gx + 150.0, z0 = -1000.0 + 500.0 * gz + 150.0;
            world.push_back(boxed<FlipFace>(boxed<XZRect>(x0, x0 + 200.0, z0, z0 + 200.0, 1100.0, light)));
            if (lights_out) lights_out->push_back(boxed<XZRect>(x0, x0 + 200.0, z0, z0 + 200.0, 1100.0, null_mat));
        }
    }
    return BVHNode::new_(std::move(world), 0.0, 1.0, rng);
}

int scene_id_from_name(const std::string &name) {
    static const std::pair<const char *, int> names[] = {
        {"random_scene", 0}, {"two_spheres", 1}, {"two_perlin_spheres", 2}, {"earth", 3},       {"simple_light", 4},
        {"cornel_box", 5},   {"cornell_box", 5}, {"cornel_smoke", 6},       {"cornell_smoke", 6}, {"final_scene", 7},
        {"stress", 8},       {"one_weekend", 9}};
    for (auto &p : names)
        if (name == p.first) return p.second;
    return -1;
}

std::unique_ptr<SceneSetup> select_scene(int which, uint64_t seed, const EarthMap &map, int stress_spheres) {
    auto s = std::make_unique<SceneSetup>();
    SceneRng rng = SceneRng::seed_from_u64(seed);
    MaterialRef null_mat = std::make_shared<NullMaterial>(); // main.rs:805
    auto sky_camera = [&] {                                  // arms 0..3 (main.rs:816-855)
        s->background = Color(vec3(0.70, 0.80, 1.00));
        s->look_from = point3(13.0, 2.0, 3.0), s->look_at = point3(0.0, 0.0, 0.0);
        s->vfov = Deg{20.0};
    };
    auto cornell_camera = [&](int width, int spp) { // arms 5, 6 (main.rs:867-915)
        s->aspect_ratio = 1.0, s->image_width = width, s->samples_per_pixel = spp;
        s->background = Color(vec3(0.0, 0.0, 0.0));
        s->look_from = point3(278.0, 278.0, -800.0), s->look_at = point3(278.0, 278.0, 0.0);
        s->vfov = Deg{40.0};
    };
    std::vector<HittableBox> lights;
    bool have_lights = false;
    std::unique_ptr<BVHNode> world;
    switch (which) {
    case 0:
        s->samples_per_pixel = 500;
        world = random_scene(rng);
        sky_camera();
        s->aperture = 0.1;
        break;
    case 1:
        world = two_spheres(rng);
        sky_camera();
        break;
    case 2:
        world = two_perlin_spheres(rng);
        sky_camera();
        break;
    case 3:
        world = earth(rng, map);
        sky_camera();
        break;
    case 4:
        s->samples_per_pixel = 400;
        world = simple_light(rng);
        s->background = Color(vec3(0.0, 0.0, 0.0));
        s->look_from = point3(26.0, 3.0, 6.0), s->look_at = point3(0.0, 2.0, 0.0);
        s->vfov = Deg{20.0};
        break;
    case 5:
        world = cornel_box(rng);
        cornell_camera(600, 100);
        lights.push_back(boxed<XZRect>(213.0, 343.0, 227.0, 332.0, 554.0, null_mat));
        lights.push_back(boxed<Sphere>(point3(190.0, 90.0, 190.0), 90.0, null_mat));
        have_lights = true;
        break;
    case 6:
        world = cornel_smoke(rng);
        cornell_camera(600, 200);
        lights.push_back(boxed<XZRect>(113.0, 443.0, 127.0, 432.0, 554.0, null_mat));
        have_lights = true;
        break;
    case 7: // the `_` arm (main.rs:916-936)
        s->aspect_ratio = 1.0, s->image_width = 800, s->samples_per_pixel = 10000;
        world = final_scene(rng, map);
        lights.push_back(boxed<XZRect>(123.0, 423.0, 147.0, 412.0, 554.0, null_mat));
        have_lights = true;
        s->background = Color(vec3(0.0, 0.0, 0.0));
        s->look_from = point3(478.0, 278.0, -600.0), s->look_at = point3(278.0, 278.0, 0.0);
        s->vfov = Deg{40.0};
        break;
    case 8: // BASELINE.json config 5
        s->aspect_ratio = 16.0 / 9.0, s->image_width = 1920, s->samples_per_pixel = 256;
        world = stress_scene(rng, stress_spheres, &lights);
        have_lights = true;
        s->background = Color(vec3(0.0, 0.0, 0.0));
        s->look_from = point3(0.0, 300.0, -2500.0), s->look_at = point3(0.0, 0.0, 0.0);
        s->vfov = Deg{40.0};
        break;
    case 9: // BASELINE.json config 2 in its "One Weekend" flavour
        s->aspect_ratio = 1.5, s->image_width = 1200, s->samples_per_pixel = 500;
        world = random_scene_one_weekend(rng);
        sky_camera();
        s->aperture = 0.1;
        break;
    default:
        throw std::runtime_error("unknown scene id");
    }
    s->set_world(*world);
    if (have_lights) s->set_lights(lights);
    return s;
}

} // namespace rt1w
