// scenes.hpp — the scene functions and `match` arms of the reference's main.rs,
// written against the C++ mirror API (rt1w.hpp).  Geometry, materials and
// per-arm settings follow main.rs:192-795 and main.rs:815-937 value for value;
// the scene RNG is seeded (the reference uses OS entropy, main.rs:803).
#pragma once

#include "rt1w.hpp"

namespace rt1w {

struct EarthMap { // decoded assets/earthmap.jpg (main.rs:347,748); decoding is host tooling
    std::shared_ptr<std::vector<uint8_t>> rgb8;
    int32_t width = 0, height = 0;
};

std::unique_ptr<BVHNode> random_scene(SceneRng &rng);                              // main.rs:192-295
std::unique_ptr<BVHNode> two_spheres(SceneRng &rng);                               // main.rs:297-323
std::unique_ptr<BVHNode> two_perlin_spheres(SceneRng &rng);                        // main.rs:325-344
std::unique_ptr<BVHNode> earth(SceneRng &rng, const EarthMap &map);                // main.rs:346-358
std::unique_ptr<BVHNode> simple_light(SceneRng &rng);                              // main.rs:360-393
std::unique_ptr<BVHNode> cornel_box(SceneRng &rng);                                // main.rs:395-512
std::unique_ptr<BVHNode> cornel_smoke(SceneRng &rng);                              // main.rs:514-633
std::unique_ptr<BVHNode> final_scene(SceneRng &rng, const EarthMap &map);          // main.rs:635-795
// Not in the reference: BASELINE.json config 5 (SURVEY.md §8d "C5 inputs").
std::unique_ptr<BVHNode> stress_scene(SceneRng &rng, int n_spheres, std::vector<HittableBox> *lights_out);
// "One Weekend" flavour of random_scene (static spheres, grey ground, fuzz U[0,0.5)); source branch not mounted.
std::unique_ptr<BVHNode> random_scene_one_weekend(SceneRng &rng);

// The arms of `match 5 { ... }` (main.rs:815-937): 0..6 and the default arm (7).
// Extra ids: 8 = C5 stress scene, 9 = One-Weekend random_scene variant.
std::unique_ptr<SceneSetup> select_scene(int which, uint64_t seed, const EarthMap &map, int stress_spheres = 1000000);
int scene_id_from_name(const std::string &name);

} // namespace rt1w
