// lower.h — host-side lowering of an rt1w_scene_desc into flat device tables.
// Replaces the reference's object tree (`Box<dyn Hittable>` nesting, main.rs:192-795) by
// a primitive table + wrapper-chain frames; see DESIGN.md "Lowering".
#pragma once

#include <string>
#include <vector>

#include "../../include/rt1w.h"
#include "device_types.h"

namespace rt1w {

struct LoweredImage {
    const uint8_t *rgb8;
    int32_t width, height;
};

struct LoweredScene {
    std::vector<rt1w_flat_prim> prims; // primitive-id (DFS) order
    std::vector<DFrame> frames;
    std::vector<DMaterial> materials;
    std::vector<DTexture> textures;
    std::vector<DPerlin> perlins;
    std::vector<LoweredImage> images;
    std::vector<DLight> lights;
    bool has_lights = false;
    int material_mask = 0;
};

// Returns RT1W_OK or an error status with a message in `err`.
rt1w_status lower_scene(const rt1w_scene_desc *desc, LoweredScene &out, std::string &err);

// Device form of one lowered primitive.
DPrim make_device_prim(const rt1w_flat_prim &fp, const std::vector<DMaterial> &materials);

} // namespace rt1w
