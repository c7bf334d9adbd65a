// rt1w.hpp — C++ mirror of the reference's Rust scene-construction API.
//
// The reference is a Rust binary; no Rust toolchain exists in this image, so the
// host side above the C ABI (include/rt1w.h) is written in C++ with the SAME type
// names, field names and constructor argument order as the Rust sources, so that
// scene code reads like main.rs:192-795.  Every type gains one lowering hook
// (`lower(SceneBuilder&)`) that appends POD rows to an rt1w_scene_desc; nothing
// is evaluated on the host.  Citations are relative to the reference `src/`.
//
//   Rust                                   here
//   Arc<Box<dyn Material>>                 MaterialRef  (std::shared_ptr<Material>)
//   Box<dyn Hittable>                      HittableBox  (std::unique_ptr<Hittable>)
//   Box<dyn Texture> / generic T: Texture  TextureRef   (std::shared_ptr<Texture>)
//   point3/vec3 (cgmath)                   point3()/vec3() -> Vec3
//   Deg(15.0)                              Deg{15.0}
//   impl Rng                               SceneRng (Philox-backed, seeded; the reference seeds from entropy, main.rs:803)
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rt1w.h"

namespace rt1w {

using Float = double; // main.rs:1

struct Vec3 {
    Float x = 0, y = 0, z = 0;
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 vec3(Float x, Float y, Float z) { return Vec3{x, y, z}; }
inline Vec3 point3(Float x, Float y, Float z) { return Vec3{x, y, z}; }
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(Vec3 a, Float s) { return s * a; }
inline Vec3 operator/(Vec3 a, Float s) { return {a.x / s, a.y / s, a.z / s}; }
inline Float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Float magnitude(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { return a / magnitude(a); }
inline Vec3 mul_element_wise(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }

struct Deg {
    Float value;
};

struct Color { // color.rs:8
    Vec3 v;
    Color() = default;
    explicit Color(Vec3 c) : v(c) {}
};

// ---------------------------------------------------------------------------
// Scene RNG: the distributions of rand 0.8 that the scene code uses
// (main.rs:212-245,653,777-779; perlin.rs:21,30-32; color.rs:31-35), driven by
// Philox4x32-10 so that scenes are reproducible from a seed.
// ---------------------------------------------------------------------------
class SceneRng {
  public:
    explicit SceneRng(uint64_t seed) : key_{uint32_t(seed), uint32_t(seed >> 32)} {}
    static SceneRng seed_from_u64(uint64_t seed) { return SceneRng(seed); }

    uint32_t next_u32() {
        if (have_ == 0) {
            uint32_t ctr[4] = {uint32_t(block_), uint32_t(block_ >> 32), 0x5ce9eu, 0};
            rt1w_philox4x32_host(ctr, key_, buf_);
            ++block_;
            have_ = 4;
        }
        return buf_[4 - have_--];
    }
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
    // `rng.gen::<Float>()`: 53-bit uniform in [0,1).
    Float gen() { return Float(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // `rng.gen_range(a..b)` for floats.
    Float gen_range(Float lo, Float hi) {
        for (;;) {
            Float r = lo + (hi - lo) * gen();
            if (r < hi) return r;
        }
    }
    // `rng.gen_range(0..=n)` for integers (unbiased by rejection).
    uint32_t gen_range_inclusive(uint32_t n) {
        uint64_t range = uint64_t(n) + 1;
        uint64_t zone = (uint64_t(1) << 32) - ((uint64_t(1) << 32) % range);
        for (;;) {
            uint64_t v = next_u32();
            if (v < zone) return uint32_t(v % range);
        }
    }
    // `rng.gen::<Color>()` (color.rs:31-35).
    Color gen_color() {
        Float r = gen(), g = gen(), b = gen();
        return Color(vec3(r, g, b));
    }
    // `slice.shuffle(rng)`: Fisher-Yates from the back, as rand's SliceRandom does.
    template <class T> void shuffle(T *data, size_t len) {
        for (size_t i = len; i > 1; --i) {
            size_t j = gen_range_inclusive(uint32_t(i - 1));
            std::swap(data[i - 1], data[j]);
        }
    }

    static void rt1w_philox4x32_host(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
        uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
        uint32_t k0 = key_in[0], k1 = key_in[1];
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = uint64_t(0xD2511F53u) * c0;
            uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
            uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
            uint32_t n1 = uint32_t(p1);
            uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
            uint32_t n3 = uint32_t(p0);
            c0 = n0, c1 = n1, c2 = n2, c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
    }

  private:
    uint32_t key_[2];
    uint64_t block_ = 0;
    uint32_t buf_[4] = {0, 0, 0, 0};
    int have_ = 0;
};

// ---------------------------------------------------------------------------
// SceneBuilder: owns the POD tables an rt1w_scene_desc points into.
// ---------------------------------------------------------------------------
class Material;
class Texture;

class SceneBuilder {
  public:
    std::vector<rt1w_node> nodes;
    std::vector<int32_t> children;
    std::vector<rt1w_material> materials;
    std::vector<rt1w_texture> textures;
    std::vector<rt1w_perlin> perlins;
    std::vector<rt1w_image> images;
    std::vector<std::shared_ptr<std::vector<uint8_t>>> image_storage;
    std::vector<int32_t> lights;
    int32_t world = -1;
    bool has_lights = false;

    int32_t add_node(int32_t type, int32_t material, std::initializer_list<double> p, const std::vector<int32_t> &kids = {}) {
        rt1w_node n{};
        n.type = type;
        n.material = material;
        n.child_begin = int32_t(children.size());
        n.child_count = int32_t(kids.size());
        int i = 0;
        for (double v : p) n.p[i++] = v;
        for (int32_t k : kids) children.push_back(k);
        nodes.push_back(n);
        return int32_t(nodes.size()) - 1;
    }
    int32_t material_id(const std::shared_ptr<Material> &m);
    int32_t texture_id(const std::shared_ptr<Texture> &t);

    rt1w_scene_desc desc() const {
        rt1w_scene_desc d{};
        d.nodes = nodes.data(), d.n_nodes = int32_t(nodes.size());
        d.children = children.data(), d.n_children = int32_t(children.size());
        d.materials = materials.data(), d.n_materials = int32_t(materials.size());
        d.textures = textures.data(), d.n_textures = int32_t(textures.size());
        d.perlins = perlins.data(), d.n_perlins = int32_t(perlins.size());
        d.images = images.data(), d.n_images = int32_t(images.size());
        d.world = world;
        d.has_lights = has_lights ? 1 : 0;
        d.lights = lights.data(), d.n_lights = int32_t(lights.size());
        return d;
    }

  private:
    std::map<const Material *, int32_t> material_ids_;
    std::map<const Texture *, int32_t> texture_ids_;
};

// ---------------------------------------------------------------------------
// Textures (texture.rs, perlin.rs)
// ---------------------------------------------------------------------------
class Texture { // trait Texture, texture.rs:8-10
  public:
    virtual ~Texture() = default;
    virtual int32_t lower(SceneBuilder &b) const = 0;
};
using TextureRef = std::shared_ptr<Texture>;

struct SolidColor : Texture { // texture.rs:12-15
    Color color_value;
    explicit SolidColor(Color c) : color_value(c) {}
    int32_t lower(SceneBuilder &b) const override {
        rt1w_texture t{};
        t.type = RT1W_TEX_SOLID, t.odd = t.even = t.table = -1;
        t.color[0] = color_value.v.x, t.color[1] = color_value.v.y, t.color[2] = color_value.v.z;
        b.textures.push_back(t);
        return int32_t(b.textures.size()) - 1;
    }
};

struct CheckerTexture : Texture { // texture.rs:17-21 (field order: odd, even)
    TextureRef odd, even;
    CheckerTexture(TextureRef odd_, TextureRef even_) : odd(std::move(odd_)), even(std::move(even_)) {}
    int32_t lower(SceneBuilder &b) const override {
        rt1w_texture t{};
        t.type = RT1W_TEX_CHECKER, t.table = -1;
        t.odd = b.texture_id(odd);
        t.even = b.texture_id(even);
        b.textures.push_back(t);
        return int32_t(b.textures.size()) - 1;
    }
};

struct Perlin { // Perlin<256>, perlin.rs:7-43
    static constexpr int POINT_COUNT = 256;
    rt1w_perlin tables;
    static Perlin new_(SceneRng &rng) {
        Perlin p;
        for (int i = 0; i < POINT_COUNT; ++i) { // perlin.rs:27-35: normalize(U[-1,1)^3), no rejection
            Float x = rng.gen_range(-1.0, 1.0), y = rng.gen_range(-1.0, 1.0), z = rng.gen_range(-1.0, 1.0);
            Vec3 v = normalize(vec3(x, y, z));
            p.tables.ranvec[i][0] = v.x, p.tables.ranvec[i][1] = v.y, p.tables.ranvec[i][2] = v.z;
        }
        generate_perm(rng, p.tables.perm_x);
        generate_perm(rng, p.tables.perm_y);
        generate_perm(rng, p.tables.perm_z);
        return p;
    }

  private:
    static void generate_perm(SceneRng &rng, int32_t *perm) { // perlin.rs:15-23
        for (int i = 0; i < POINT_COUNT; ++i) perm[i] = i;
        rng.shuffle(perm, POINT_COUNT);
    }
};

struct NoiseTexture256 : Texture { // texture.rs:23-38
    Perlin perlin;
    Float scale;
    NoiseTexture256(Perlin p, Float s) : perlin(p), scale(s) {}
    static std::shared_ptr<NoiseTexture256> new_(Float scale, SceneRng &rng) {
        return std::make_shared<NoiseTexture256>(Perlin::new_(rng), scale);
    }
    int32_t lower(SceneBuilder &b) const override {
        b.perlins.push_back(perlin.tables);
        rt1w_texture t{};
        t.type = RT1W_TEX_NOISE, t.odd = t.even = -1;
        t.table = int32_t(b.perlins.size()) - 1;
        t.scale = scale;
        b.textures.push_back(t);
        return int32_t(b.textures.size()) - 1;
    }
};

// `image::DynamicImage` as a texture (texture.rs:67-89).  JPEG decoding is host
// tooling (the `image` crate in the reference, PIL here); this type holds RGB8.
struct DynamicImage : Texture {
    std::shared_ptr<std::vector<uint8_t>> rgb8;
    int32_t width = 0, height = 0;
    DynamicImage(std::shared_ptr<std::vector<uint8_t>> px, int32_t w, int32_t h) : rgb8(std::move(px)), width(w), height(h) {}
    int32_t lower(SceneBuilder &b) const override {
        if (!rgb8 || int64_t(rgb8->size()) < int64_t(width) * height * 3) throw std::runtime_error("DynamicImage: pixel buffer too small");
        b.image_storage.push_back(rgb8);
        rt1w_image im{rgb8->data(), width, height};
        b.images.push_back(im);
        rt1w_texture t{};
        t.type = RT1W_TEX_IMAGE, t.odd = t.even = -1;
        t.table = int32_t(b.images.size()) - 1;
        b.textures.push_back(t);
        return int32_t(b.textures.size()) - 1;
    }
};

// ---------------------------------------------------------------------------
// Materials (material.rs, constant_medium.rs:31-51)
// ---------------------------------------------------------------------------
class Material { // trait Material, material.rs:25-50
  public:
    virtual ~Material() = default;
    virtual rt1w_material lower(SceneBuilder &b) const = 0;
};
using MaterialRef = std::shared_ptr<Material>; // Arc<Box<dyn Material>>

struct Lambertian : Material { // material.rs:52-55
    TextureRef albedo;
    explicit Lambertian(TextureRef a) : albedo(std::move(a)) {}
    rt1w_material lower(SceneBuilder &b) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_LAMBERTIAN;
        m.texture = b.texture_id(albedo);
        return m;
    }
};
struct Metal : Material { // material.rs:57-61
    Color albedo;
    Float fuzz;
    Metal(Color a, Float f) : albedo(a), fuzz(f) {}
    rt1w_material lower(SceneBuilder &) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_METAL, m.texture = -1;
        m.albedo[0] = albedo.v.x, m.albedo[1] = albedo.v.y, m.albedo[2] = albedo.v.z;
        m.fuzz = fuzz;
        return m;
    }
};
struct Dielectric : Material { // material.rs:127-130
    Float ir;
    explicit Dielectric(Float i) : ir(i) {}
    rt1w_material lower(SceneBuilder &) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_DIELECTRIC, m.texture = -1, m.ir = ir;
        return m;
    }
};
struct DiffuseLight : Material { // material.rs:63-66
    TextureRef emit;
    explicit DiffuseLight(TextureRef e) : emit(std::move(e)) {}
    rt1w_material lower(SceneBuilder &b) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_DIFFUSE_LIGHT;
        m.texture = b.texture_id(emit);
        return m;
    }
};
struct Isotropic : Material { // constant_medium.rs:31-34
    TextureRef albedo;
    explicit Isotropic(TextureRef a) : albedo(std::move(a)) {}
    rt1w_material lower(SceneBuilder &b) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_ISOTROPIC;
        m.texture = b.texture_id(albedo);
        return m;
    }
};
struct NullMaterial : Material { // `impl Material for ()`, material.rs:68
    rt1w_material lower(SceneBuilder &) const override {
        rt1w_material m{};
        m.type = RT1W_MAT_NONE, m.texture = -1;
        return m;
    }
};

inline int32_t SceneBuilder::material_id(const MaterialRef &m) {
    if (!m) throw std::runtime_error("null material handle");
    auto it = material_ids_.find(m.get());
    if (it != material_ids_.end()) return it->second;
    rt1w_material row = m->lower(*this);
    materials.push_back(row);
    int32_t id = int32_t(materials.size()) - 1;
    material_ids_[m.get()] = id;
    return id;
}
inline int32_t SceneBuilder::texture_id(const TextureRef &t) {
    if (!t) throw std::runtime_error("null texture handle");
    auto it = texture_ids_.find(t.get());
    if (it != texture_ids_.end()) return it->second;
    int32_t id = t->lower(*this);
    texture_ids_[t.get()] = id;
    return id;
}

// ---------------------------------------------------------------------------
// Hittables (hittable.rs, sphere.rs, moving_sphere.rs, aarect.rs, aabox.rs,
// constant_medium.rs, bvh.rs)
// ---------------------------------------------------------------------------
class Hittable { // trait Hittable, hittable.rs:63-72
  public:
    virtual ~Hittable() = default;
    virtual int32_t lower(SceneBuilder &b) const = 0;
};
using HittableBox = std::unique_ptr<Hittable>; // Box<dyn Hittable>

struct Sphere : Hittable { // sphere.rs:16-20
    Vec3 center;
    Float radius;
    MaterialRef material;
    Sphere(Vec3 c, Float r, MaterialRef m) : center(c), radius(r), material(std::move(m)) {}
    int32_t lower(SceneBuilder &b) const override {
        return b.add_node(RT1W_NODE_SPHERE, b.material_id(material), {center.x, center.y, center.z, radius});
    }
};
struct MovingSphere : Hittable { // moving_sphere.rs:13-20
    Vec3 center0, center1;
    Float time0, time1, radius;
    MaterialRef material;
    MovingSphere(Vec3 c0, Vec3 c1, Float t0, Float t1, Float r, MaterialRef m)
        : center0(c0), center1(c1), time0(t0), time1(t1), radius(r), material(std::move(m)) {}
    int32_t lower(SceneBuilder &b) const override {
        return b.add_node(RT1W_NODE_MOVING_SPHERE, b.material_id(material),
                          {center0.x, center0.y, center0.z, center1.x, center1.y, center1.z, time0, time1, radius});
    }
};
struct XYRect : Hittable { // aarect.rs:15-22
    Float x0, x1, y0, y1, k;
    MaterialRef material;
    XYRect(Float x0_, Float x1_, Float y0_, Float y1_, Float k_, MaterialRef m) : x0(x0_), x1(x1_), y0(y0_), y1(y1_), k(k_), material(std::move(m)) {}
    int32_t lower(SceneBuilder &b) const override { return b.add_node(RT1W_NODE_XY_RECT, b.material_id(material), {x0, x1, y0, y1, k}); }
};
struct XZRect : Hittable { // aarect.rs:25-32
    Float x0, x1, z0, z1, k;
    MaterialRef material;
    XZRect(Float x0_, Float x1_, Float z0_, Float z1_, Float k_, MaterialRef m) : x0(x0_), x1(x1_), z0(z0_), z1(z1_), k(k_), material(std::move(m)) {}
    int32_t lower(SceneBuilder &b) const override { return b.add_node(RT1W_NODE_XZ_RECT, b.material_id(material), {x0, x1, z0, z1, k}); }
};
struct YZRect : Hittable { // aarect.rs:35-42
    Float y0, y1, z0, z1, k;
    MaterialRef material;
    YZRect(Float y0_, Float y1_, Float z0_, Float z1_, Float k_, MaterialRef m) : y0(y0_), y1(y1_), z0(z0_), z1(z1_), k(k_), material(std::move(m)) {}
    int32_t lower(SceneBuilder &b) const override { return b.add_node(RT1W_NODE_YZ_RECT, b.material_id(material), {y0, y1, z0, z1, k}); }
};
struct AABox : Hittable { // aabox.rs:16-27 (the rng only feeds the inner BVHNode's axis choice)
    Vec3 box_min, box_max;
    MaterialRef material;
    static std::unique_ptr<AABox> new_(Vec3 p0, Vec3 p1, MaterialRef material, SceneRng &) {
        auto bx = std::make_unique<AABox>();
        bx->box_min = p0, bx->box_max = p1, bx->material = std::move(material);
        return bx;
    }
    int32_t lower(SceneBuilder &b) const override {
        return b.add_node(RT1W_NODE_AABOX, b.material_id(material), {box_min.x, box_min.y, box_min.z, box_max.x, box_max.y, box_max.z});
    }
};
struct Translate : Hittable { // hittable.rs:49-52
    HittableBox hittable;
    Vec3 offset;
    Translate(HittableBox h, Vec3 o) : hittable(std::move(h)), offset(o) {}
    int32_t lower(SceneBuilder &b) const override {
        int32_t c = hittable->lower(b);
        return b.add_node(RT1W_NODE_TRANSLATE, -1, {offset.x, offset.y, offset.z}, {c});
    }
};
struct RotateY : Hittable { // hittable.rs:54-59,158
    HittableBox hittable;
    Float time0, time1;
    Deg angle;
    static std::unique_ptr<RotateY> new_(HittableBox h, Float time0, Float time1, Deg angle) {
        auto r = std::make_unique<RotateY>();
        r->hittable = std::move(h), r->time0 = time0, r->time1 = time1, r->angle = angle;
        return r;
    }
    int32_t lower(SceneBuilder &b) const override {
        int32_t c = hittable->lower(b);
        return b.add_node(RT1W_NODE_ROTATE_Y, -1, {angle.value, time0, time1}, {c});
    }
};
struct FlipFace : Hittable { // hittable.rs:61
    HittableBox inner;
    explicit FlipFace(HittableBox h) : inner(std::move(h)) {}
    int32_t lower(SceneBuilder &b) const override {
        int32_t c = inner->lower(b);
        return b.add_node(RT1W_NODE_FLIP_FACE, -1, {}, {c});
    }
};
struct ConstantMedium : Hittable { // constant_medium.rs:15-28
    HittableBox boundary;
    MaterialRef phase_function;
    Float density;
    static std::unique_ptr<ConstantMedium> new_(HittableBox boundary, Float d, TextureRef texture) {
        auto m = std::make_unique<ConstantMedium>();
        m->boundary = std::move(boundary);
        m->phase_function = std::make_shared<Isotropic>(std::move(texture));
        m->density = d;
        return m;
    }
    int32_t lower(SceneBuilder &b) const override {
        int32_t c = boundary->lower(b);
        return b.add_node(RT1W_NODE_CONSTANT_MEDIUM, b.material_id(phase_function), {density}, {c});
    }
};
struct BVHNode : Hittable { // bvh.rs:15-18,54-59
    std::vector<HittableBox> objects;
    Float time0 = 0, time1 = 1;
    static std::unique_ptr<BVHNode> new_(std::vector<HittableBox> objects, Float time0, Float time1, SceneRng &) {
        if (objects.empty()) throw std::runtime_error("objects mut not be empty"); // bvh.rs:61
        auto n = std::make_unique<BVHNode>();
        n->objects = std::move(objects), n->time0 = time0, n->time1 = time1;
        return n;
    }
    int32_t lower(SceneBuilder &b) const override {
        std::vector<int32_t> kids;
        kids.reserve(objects.size());
        for (auto &o : objects) kids.push_back(o->lower(b));
        return b.add_node(RT1W_NODE_BVH, -1, {time0, time1}, kids);
    }
};

// ---------------------------------------------------------------------------
// Camera (camera.rs:22-59); get_ray runs on the device.
// ---------------------------------------------------------------------------
struct Camera {
    rt1w_camera pod;
    static Camera new_(Vec3 look_from, Vec3 look_at, Vec3 vup, Deg vfov, Float aspect_ratio, Float aperture, Float focus_dist,
                       Float time0, Float time1) {
        const Float pi = 3.14159265358979323846;
        Float theta = vfov.value * pi / 180.0;
        Float h = std::tan(theta / 2.0);
        Float viewport_height = 2.0 * h;
        Float viewport_width = aspect_ratio * viewport_height;
        Vec3 w = normalize(look_from - look_at);
        Vec3 u = normalize(cross(vup, w));
        Vec3 v = cross(w, u);
        Vec3 origin = look_from;
        Vec3 horizontal = focus_dist * viewport_width * u;
        Vec3 vertical = focus_dist * viewport_height * v;
        Vec3 llc = origin - horizontal / 2.0 - vertical / 2.0 - focus_dist * w;
        Camera c;
        auto put = [](double *d, Vec3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
        put(c.pod.origin, origin), put(c.pod.lower_left_corner, llc), put(c.pod.horizontal, horizontal), put(c.pod.vertical, vertical);
        put(c.pod.u, u), put(c.pod.v, v), put(c.pod.w, w);
        c.pod.lens_radius = aperture / 2.0;
        c.pod.time0 = time0, c.pod.time1 = time1;
        return c;
    }
};

// ---------------------------------------------------------------------------
// A lowered, self-contained scene + the per-arm settings of main.rs:815-937.
// ---------------------------------------------------------------------------
struct SceneSetup {
    SceneBuilder builder;
    Color background;
    Vec3 look_from, look_at;
    Deg vfov{40.0};
    Float aperture = 0.0;
    Float aspect_ratio = 16.0 / 9.0; // main.rs:798
    int image_width = 400;           // main.rs:799
    int samples_per_pixel = 100;     // main.rs:800
    int max_depth = 50;              // main.rs:801

    int image_height() const { return int(Float(image_width) / aspect_ratio); } // main.rs:939
    Camera camera() const { return camera_for_aspect(aspect_ratio); }
    Camera camera_for_aspect(Float aspect) const { // main.rs:940-951
        return Camera::new_(look_from, look_at, vec3(0.0, 1.0, 0.0), vfov, aspect, aperture, 10.0, 0.0, 1.0);
    }
    void set_world(const Hittable &world) { builder.world = world.lower(builder); }
    void set_lights(const std::vector<HittableBox> &lights) { // Some(vec![...]), main.rs:873-887
        builder.has_lights = true;
        for (auto &l : lights) builder.lights.push_back(l->lower(builder));
    }
};

} // namespace rt1w
