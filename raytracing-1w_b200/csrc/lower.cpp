// lower.cpp — see lower.h.
#include "lower.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>

namespace rt1w {
namespace {

struct Fail : std::runtime_error {
    rt1w_status status;
    Fail(rt1w_status s, const std::string &m) : std::runtime_error(m), status(s) {}
};

struct Wrapper {
    int node;
    int kind; // ChainOpKind
    double sin_t = 0, cos_t = 1;
    double off[3] = {0, 0, 0};
};

// local = R(angle) p + b, composed outermost-first.
struct Xform {
    double angle = 0; // radians, cumulative
    double s = 0, c = 1;
    double b[3] = {0, 0, 0};
};

struct Walker {
    const rt1w_scene_desc &d;
    LoweredScene &out;
    std::vector<Wrapper> chain;                 // wrappers above the current node, outermost first
    std::map<std::vector<int>, int> frame_ids;  // wrapper-node path -> frame id
    double time0 = 0.0, time1 = 1.0;            // bounding_box(time0,time1) arguments in scope
    int depth = 0;

    const rt1w_node &node(int id) const {
        if (id < 0 || id >= d.n_nodes) throw Fail(RT1W_ERR_INVALID, "node id out of range");
        return d.nodes[id];
    }
    int child(const rt1w_node &n, int i) const {
        if (i >= n.child_count || n.child_begin < 0 || n.child_begin + i >= d.n_children)
            throw Fail(RT1W_ERR_INVALID, "node child index out of range");
        return d.children[n.child_begin + i];
    }
    int material(const rt1w_node &n) const {
        if (n.material < 0 || n.material >= d.n_materials) throw Fail(RT1W_ERR_INVALID, "material id out of range");
        return n.material;
    }

    bool chain_has_transform() const {
        for (auto &w : chain)
            if (w.kind != OP_FLIP_FACE) return true;
        return false;
    }

    // Frame for the current chain; `ops_upto` = number of outermost wrappers whose HitRecord
    // rewrites apply (all of them for surface leaves; only those above the medium node for media).
    int frame_for_chain(size_t ops_upto) {
        if (!chain_has_transform()) return -1;
        std::vector<int> key;
        for (auto &w : chain) key.push_back(w.node);
        key.push_back(-int(ops_upto) - 1);
        auto it = frame_ids.find(key);
        if (it != frame_ids.end()) return it->second;
        Xform x;
        for (size_t i = 0; i < chain.size(); ++i) {
            const Wrapper &w = chain[i];
            if (w.kind == OP_TRANSLATE) {
                for (int k = 0; k < 3; ++k) x.b[k] -= w.off[k];
            } else if (w.kind == OP_ROTATE_Y) {
                double nb0 = w.cos_t * x.b[0] - w.sin_t * x.b[2];
                double nb2 = w.sin_t * x.b[0] + w.cos_t * x.b[2];
                x.b[0] = nb0, x.b[2] = nb2;
                // compose the rotation exactly from the wrapper's own sin/cos (hittable.rs:159-160)
                double ns = x.s * w.cos_t + x.c * w.sin_t;
                double nc = x.c * w.cos_t - x.s * w.sin_t;
                x.s = ns, x.c = nc;
            }
        }
        DFrame f;
        std::memset(&f, 0, sizeof(f));
        f.sin_t = x.s, f.cos_t = x.c;
        f.bx = x.b[0], f.by = x.b[1], f.bz = x.b[2];
        // ops: innermost first among the first `ops_upto` wrappers
        double cs = 0, cc = 1; // cumulative rotation of the direction, outermost -> inner
        std::vector<DChainOp> outer_first;
        for (size_t i = 0; i < ops_upto && i < chain.size(); ++i) {
            const Wrapper &w = chain[i];
            DChainOp op;
            std::memset(&op, 0, sizeof(op));
            op.kind = w.kind;
            op.sin_own = float(w.sin_t), op.cos_own = float(w.cos_t);
            if (w.kind == OP_ROTATE_Y) {
                double ns = cs * w.cos_t + cc * w.sin_t;
                double nc = cc * w.cos_t - cs * w.sin_t;
                cs = ns, cc = nc;
            }
            op.sin_cum = float(cs), op.cos_cum = float(cc); // direction INSIDE this wrapper
            outer_first.push_back(op);
        }
        if (outer_first.size() > RT1W_MAX_CHAIN_OPS) throw Fail(RT1W_ERR_UNSUPPORTED, "wrapper chain deeper than 8");
        f.n_ops = int(outer_first.size());
        for (int i = 0; i < f.n_ops; ++i) f.ops[i] = outer_first[f.n_ops - 1 - i];
        out.frames.push_back(f);
        int id = int(out.frames.size()) - 1;
        frame_ids[key] = id;
        return id;
    }

    // world-space bounds of a local box under the current chain
    void world_bounds(const double lmin[3], const double lmax[3], double wmin[3], double wmax[3]) const {
        if (!chain_has_transform()) { // the common case (a million bare spheres): nothing to compose
            for (int k = 0; k < 3; ++k) wmin[k] = lmin[k], wmax[k] = lmax[k];
            return;
        }
        Xform x;
        for (auto &w : chain) {
            if (w.kind == OP_TRANSLATE) {
                for (int k = 0; k < 3; ++k) x.b[k] -= w.off[k];
            } else if (w.kind == OP_ROTATE_Y) {
                double nb0 = w.cos_t * x.b[0] - w.sin_t * x.b[2];
                double nb2 = w.sin_t * x.b[0] + w.cos_t * x.b[2];
                x.b[0] = nb0, x.b[2] = nb2;
                double ns = x.s * w.cos_t + x.c * w.sin_t;
                double nc = x.c * w.cos_t - x.s * w.sin_t;
                x.s = ns, x.c = nc;
            }
        }
        for (int k = 0; k < 3; ++k) wmin[k] = std::numeric_limits<double>::infinity(), wmax[k] = -wmin[k];
        for (int i = 0; i < 8; ++i) {
            double l[3] = {(i & 1) ? lmax[0] : lmin[0], (i & 2) ? lmax[1] : lmin[1], (i & 4) ? lmax[2] : lmin[2]};
            double q[3] = {l[0] - x.b[0], l[1] - x.b[1], l[2] - x.b[2]};
            // inverse rotation: x = c*x' + s*z', z = -s*x' + c*z'
            double wv[3] = {x.c * q[0] + x.s * q[2], q[1], -x.s * q[0] + x.c * q[2]};
            for (int k = 0; k < 3; ++k) wmin[k] = std::fmin(wmin[k], wv[k]), wmax[k] = std::fmax(wmax[k], wv[k]);
        }
    }

    int flip_parity() const {
        int f = 0;
        for (auto &w : chain)
            if (w.kind == OP_FLIP_FACE) f ^= 1;
        return f;
    }

    void emit(int kind, int node_id, int material_id, const double *p, int np, const double lmin[3], const double lmax[3], int boundary,
              size_t ops_upto) {
        rt1w_flat_prim fp;
        std::memset(&fp, 0, sizeof(fp));
        fp.kind = kind, fp.node = node_id, fp.material = material_id, fp.boundary = boundary;
        fp.frame = frame_for_chain(ops_upto);
        fp.flags = (fp.frame < 0 && flip_parity()) ? PF_FLIP_FACE : 0;
        for (int i = 0; i < np; ++i) fp.p[i] = p[i];
        fp.time0 = time0, fp.time1 = time1;
        world_bounds(lmin, lmax, fp.bbox_min, fp.bbox_max);
        out.prims.push_back(fp);
        out.material_mask |= 1 << out.materials[material_id].type;
    }

    void emit_rect(int kind, int node_id, int mat, double a0, double a1, double b0, double b1, double k) {
        double p[5] = {a0, a1, b0, b1, k};
        double lo[3], hi[3];
        const double pad = 0.0001; // aarect.rs:74-79,112-117,180-185
        if (kind == RT1W_NODE_XY_RECT) lo[0] = a0, hi[0] = a1, lo[1] = b0, hi[1] = b1, lo[2] = k - pad, hi[2] = k + pad;
        else if (kind == RT1W_NODE_XZ_RECT) lo[0] = a0, hi[0] = a1, lo[2] = b0, hi[2] = b1, lo[1] = k - pad, hi[1] = k + pad;
        else lo[1] = a0, hi[1] = a1, lo[2] = b0, hi[2] = b1, lo[0] = k - pad, hi[0] = k + pad;
        emit(kind, node_id, mat, p, 5, lo, hi, 0, chain.size());
    }

    void walk(int id) {
        if (++depth > 256) throw Fail(RT1W_ERR_INVALID, "scene graph deeper than 256 (cycle?)");
        const rt1w_node &n = node(id);
        const double *p = n.p;
        switch (n.type) {
        case RT1W_NODE_SPHERE: {
            double lo[3] = {p[0] - p[3], p[1] - p[3], p[2] - p[3]}, hi[3] = {p[0] + p[3], p[1] + p[3], p[2] + p[3]};
            emit(RT1W_NODE_SPHERE, id, material(n), p, 4, lo, hi, 0, chain.size());
            break;
        }
        case RT1W_NODE_MOVING_SPHERE: { // bbox = union of the boxes at the enclosing (time0,time1), moving_sphere.rs:72-84
            double st0 = p[6], st1 = p[7], r = p[8];
            double lo[3], hi[3];
            for (int k = 0; k < 3; ++k) {
                double ca = p[k] + ((time0 - st0) / (st1 - st0)) * (p[3 + k] - p[k]);
                double cb = p[k] + ((time1 - st0) / (st1 - st0)) * (p[3 + k] - p[k]);
                lo[k] = std::fmin(ca, cb) - r, hi[k] = std::fmax(ca, cb) + r;
            }
            emit(RT1W_NODE_MOVING_SPHERE, id, material(n), p, 9, lo, hi, 0, chain.size());
            break;
        }
        case RT1W_NODE_XY_RECT:
        case RT1W_NODE_XZ_RECT:
        case RT1W_NODE_YZ_RECT: emit_rect(n.type, id, material(n), p[0], p[1], p[2], p[3], p[4]); break;
        case RT1W_NODE_AABOX: { // six sides in the order of aabox.rs:29-76
            int m = material(n);
            emit_rect(RT1W_NODE_XY_RECT, id, m, p[0], p[3], p[1], p[4], p[5]);
            emit_rect(RT1W_NODE_XY_RECT, id, m, p[0], p[3], p[1], p[4], p[2]);
            emit_rect(RT1W_NODE_XZ_RECT, id, m, p[0], p[3], p[2], p[5], p[4]);
            emit_rect(RT1W_NODE_XZ_RECT, id, m, p[0], p[3], p[2], p[5], p[1]);
            emit_rect(RT1W_NODE_YZ_RECT, id, m, p[1], p[4], p[2], p[5], p[3]);
            emit_rect(RT1W_NODE_YZ_RECT, id, m, p[1], p[4], p[2], p[5], p[0]);
            break;
        }
        case RT1W_NODE_TRANSLATE: {
            Wrapper w;
            w.node = id, w.kind = OP_TRANSLATE, w.off[0] = p[0], w.off[1] = p[1], w.off[2] = p[2];
            chain.push_back(w);
            walk(child(n, 0));
            chain.pop_back();
            break;
        }
        case RT1W_NODE_ROTATE_Y: {
            Wrapper w;
            w.node = id, w.kind = OP_ROTATE_Y;
            double radians = p[0] * (3.14159265358979323846264338327950288 / 180.0);
            w.sin_t = std::sin(radians), w.cos_t = std::cos(radians);
            double st0 = time0, st1 = time1;
            time0 = p[1], time1 = p[2]; // RotateY::new(hittable, time0, time1, ..) queries the child's box with these
            chain.push_back(w);
            walk(child(n, 0));
            chain.pop_back();
            time0 = st0, time1 = st1;
            break;
        }
        case RT1W_NODE_FLIP_FACE: {
            Wrapper w;
            w.node = id, w.kind = OP_FLIP_FACE;
            chain.push_back(w);
            walk(child(n, 0));
            chain.pop_back();
            break;
        }
        case RT1W_NODE_BVH: {
            if (n.child_count <= 0) throw Fail(RT1W_ERR_INVALID, "objects mut not be empty (BVHNode::new, bvh.rs:61)");
            double st0 = time0, st1 = time1;
            time0 = p[0], time1 = p[1];
            for (int i = 0; i < n.child_count; ++i) walk(child(n, i));
            time0 = st0, time1 = st1;
            break;
        }
        case RT1W_NODE_CONSTANT_MEDIUM: {
            int m = material(n);
            if (out.materials[m].type != RT1W_MAT_ISOTROPIC && out.materials[m].type != RT1W_MAT_NONE) {
                // ConstantMedium::new always builds an Isotropic (constant_medium.rs:22-28); other phase materials
                // would need a surface record the medium does not have.
                throw Fail(RT1W_ERR_UNSUPPORTED, "ConstantMedium phase function must be Isotropic");
            }
            if (!(p[0] != 0.0)) throw Fail(RT1W_ERR_INVALID, "ConstantMedium density must be non-zero");
            size_t ops_upto = chain.size();
            // descend through wrappers to the boundary leaf
            int cur = child(n, 0);
            size_t pushed = 0;
            for (;;) {
                const rt1w_node &b = node(cur);
                if (b.type == RT1W_NODE_TRANSLATE) {
                    Wrapper w;
                    w.node = cur, w.kind = OP_TRANSLATE, w.off[0] = b.p[0], w.off[1] = b.p[1], w.off[2] = b.p[2];
                    chain.push_back(w), ++pushed;
                    cur = child(b, 0);
                } else if (b.type == RT1W_NODE_ROTATE_Y) {
                    Wrapper w;
                    w.node = cur, w.kind = OP_ROTATE_Y;
                    double radians = b.p[0] * (3.14159265358979323846264338327950288 / 180.0);
                    w.sin_t = std::sin(radians), w.cos_t = std::cos(radians);
                    chain.push_back(w), ++pushed;
                    cur = child(b, 0);
                } else if (b.type == RT1W_NODE_FLIP_FACE) { // only toggles a flag the medium never reads
                    cur = child(b, 0);
                } else if (b.type == RT1W_NODE_BVH && b.child_count == 1) {
                    cur = child(b, 0);
                } else {
                    break;
                }
                if (pushed > 64) throw Fail(RT1W_ERR_INVALID, "boundary wrapper chain too deep");
            }
            const rt1w_node &b = node(cur);
            double nid = -1.0 / p[0]; // neg_inv_density, constant_medium.rs:26
            if (b.type == RT1W_NODE_SPHERE) {
                double q[5] = {b.p[0], b.p[1], b.p[2], b.p[3], nid};
                double lo[3] = {q[0] - q[3], q[1] - q[3], q[2] - q[3]}, hi[3] = {q[0] + q[3], q[1] + q[3], q[2] + q[3]};
                emit(RT1W_NODE_CONSTANT_MEDIUM, id, m, q, 5, lo, hi, RT1W_NODE_SPHERE, ops_upto);
            } else if (b.type == RT1W_NODE_AABOX) {
                double q[7] = {b.p[0], b.p[1], b.p[2], nid, b.p[3], b.p[4], b.p[5]};
                double lo[3] = {b.p[0], b.p[1], b.p[2]}, hi[3] = {b.p[3], b.p[4], b.p[5]};
                emit(RT1W_NODE_CONSTANT_MEDIUM, id, m, q, 7, lo, hi, RT1W_NODE_AABOX, ops_upto);
            } else {
                throw Fail(RT1W_ERR_UNSUPPORTED, "ConstantMedium boundary must be a Sphere or an AABox (optionally under Translate/RotateY)");
            }
            for (size_t i = 0; i < pushed; ++i) chain.pop_back();
            break;
        }
        default: throw Fail(RT1W_ERR_INVALID, "unknown node type");
        }
        --depth;
    }
};

void lower_tables(const rt1w_scene_desc &d, LoweredScene &out) {
    if (d.n_materials < 0 || d.n_textures < 0 || d.n_perlins < 0 || d.n_images < 0) throw Fail(RT1W_ERR_INVALID, "negative table size");
    if (d.n_materials >= (1 << 20)) throw Fail(RT1W_ERR_UNSUPPORTED, "more than 2^20 materials");
    for (int i = 0; i < d.n_perlins; ++i) {
        DPerlin p;
        std::memset(&p, 0, sizeof(p));
        for (int k = 0; k < 256; ++k) {
            for (int c = 0; c < 3; ++c) p.ranvec[k][c] = float(d.perlins[i].ranvec[k][c]);
            const int32_t *perms[3] = {d.perlins[i].perm_x, d.perlins[i].perm_y, d.perlins[i].perm_z};
            for (int a = 0; a < 3; ++a) {
                if (perms[a][k] < 0 || perms[a][k] > 255) throw Fail(RT1W_ERR_INVALID, "perlin permutation entry out of range");
                p.perm[a][k] = uint8_t(perms[a][k]);
            }
        }
        out.perlins.push_back(p);
    }
    for (int i = 0; i < d.n_images; ++i) {
        if (!d.images[i].rgb8 || d.images[i].width <= 0 || d.images[i].height <= 0) throw Fail(RT1W_ERR_INVALID, "empty image");
        out.images.push_back(LoweredImage{d.images[i].rgb8, d.images[i].width, d.images[i].height});
    }
    for (int i = 0; i < d.n_textures; ++i) {
        const rt1w_texture &t = d.textures[i];
        DTexture o;
        std::memset(&o, 0, sizeof(o));
        o.type = t.type, o.odd = t.odd, o.even = t.even, o.table = t.table, o.scale = float(t.scale);
        for (int c = 0; c < 3; ++c) o.color[c] = float(t.color[c]);
        switch (t.type) {
        case RT1W_TEX_SOLID: break;
        case RT1W_TEX_CHECKER:
            if (t.odd < 0 || t.odd >= d.n_textures || t.even < 0 || t.even >= d.n_textures) throw Fail(RT1W_ERR_INVALID, "checker child texture out of range");
            break;
        case RT1W_TEX_NOISE:
        case RT1W_TEX_PERLIN:
            if (t.table < 0 || t.table >= d.n_perlins) throw Fail(RT1W_ERR_INVALID, "perlin table id out of range");
            break;
        case RT1W_TEX_IMAGE:
            if (t.table < 0 || t.table >= d.n_images) throw Fail(RT1W_ERR_INVALID, "image id out of range");
            break;
        default: throw Fail(RT1W_ERR_INVALID, "unknown texture type");
        }
        out.textures.push_back(o);
    }
    // checker nesting must terminate within the device's fixed walk length
    for (int i = 0; i < d.n_textures; ++i) {
        int cur = i, steps = 0;
        while (d.textures[cur].type == RT1W_TEX_CHECKER) {
            cur = d.textures[cur].odd; // both branches are bounded the same way; follow one and cap globally below
            if (++steps > 8) throw Fail(RT1W_ERR_UNSUPPORTED, "checker textures nested deeper than 8");
        }
        cur = i, steps = 0;
        while (d.textures[cur].type == RT1W_TEX_CHECKER) {
            cur = d.textures[cur].even;
            if (++steps > 8) throw Fail(RT1W_ERR_UNSUPPORTED, "checker textures nested deeper than 8");
        }
    }
    for (int i = 0; i < d.n_materials; ++i) {
        const rt1w_material &m = d.materials[i];
        DMaterial o;
        std::memset(&o, 0, sizeof(o));
        o.type = m.type, o.texture = m.texture, o.fuzz = float(m.fuzz), o.ir = float(m.ir);
        for (int c = 0; c < 3; ++c) o.albedo[c] = float(m.albedo[c]);
        if (m.type < 0 || m.type > RT1W_MAT_NONE) throw Fail(RT1W_ERR_INVALID, "unknown material type");
        bool needs_tex = m.type == RT1W_MAT_LAMBERTIAN || m.type == RT1W_MAT_DIFFUSE_LIGHT || m.type == RT1W_MAT_ISOTROPIC;
        if (needs_tex && (m.texture < 0 || m.texture >= d.n_textures)) throw Fail(RT1W_ERR_INVALID, "material texture id out of range");
        out.materials.push_back(o);
    }
}

} // namespace

rt1w_status lower_scene(const rt1w_scene_desc *desc, LoweredScene &out, std::string &err) {
    try {
        if (!desc) throw Fail(RT1W_ERR_INVALID, "null scene description");
        if (desc->n_nodes <= 0 || !desc->nodes) throw Fail(RT1W_ERR_INVALID, "empty scene (objects mut not be empty, bvh.rs:61)");
        lower_tables(*desc, out);
        out.prims.reserve(size_t(desc->n_nodes));
        Walker w{*desc, out};
        w.walk(desc->world);
        if (out.prims.empty()) throw Fail(RT1W_ERR_INVALID, "scene has no primitives");
        out.has_lights = desc->has_lights != 0;
        if (out.has_lights) {
            if (desc->n_lights <= 0) throw Fail(RT1W_ERR_INVALID, "lights = Some(vec![]) would panic in choose().unwrap() (hittable.rs:153)");
            if (desc->n_lights > RT1W_MAX_LIGHTS) throw Fail(RT1W_ERR_UNSUPPORTED, "more than 32 lights");
            for (int i = 0; i < desc->n_lights; ++i) {
                const rt1w_node &n = w.node(desc->lights[i]);
                DLight l;
                std::memset(&l, 0, sizeof(l));
                if (n.type == RT1W_NODE_XZ_RECT) { // aarect.rs:119-147
                    l.kind = L_XZ_RECT;
                    for (int k = 0; k < 5; ++k) l.p[k] = n.p[k];
                } else if (n.type == RT1W_NODE_SPHERE) { // sphere.rs:72-99
                    l.kind = L_SPHERE;
                    for (int k = 0; k < 4; ++k) l.p[k] = n.p[k];
                } else { // trait defaults: pdf 0, direction (1,0,0) (hittable.rs:66-71); wrappers do not forward
                    l.kind = L_OTHER;
                }
                out.lights.push_back(l);
            }
        }
        return RT1W_OK;
    } catch (const Fail &f) {
        err = f.what();
        return f.status;
    } catch (const std::exception &e) {
        err = e.what();
        return RT1W_ERR_INVALID;
    }
}

DPrim make_device_prim(const rt1w_flat_prim &fp, const std::vector<DMaterial> &materials) {
    DPrim d;
    std::memset(&d, 0, sizeof(d));
    int type = 0;
    const double *p = fp.p;
    switch (fp.kind) {
    case RT1W_NODE_SPHERE:
        type = P_SPHERE;
        for (int i = 0; i < 4; ++i) d.p[i] = p[i];
        break;
    case RT1W_NODE_MOVING_SPHERE: // center(time) = center0 + ((time - time0) / (time1 - time0)) * (center1 - center0), moving_sphere.rs:23-26
        type = P_MOVING_SPHERE;
        for (int k = 0; k < 3; ++k) d.p[k] = p[k], d.f[k] = float(p[3 + k] - p[k]);
        d.p[3] = p[8];
        d.f[3] = float(p[6]), d.f[4] = float(1.0 / (p[7] - p[6]));
        break;
    case RT1W_NODE_XY_RECT:
    case RT1W_NODE_XZ_RECT:
    case RT1W_NODE_YZ_RECT:
        type = fp.kind == RT1W_NODE_XY_RECT ? P_XY_RECT : (fp.kind == RT1W_NODE_XZ_RECT ? P_XZ_RECT : P_YZ_RECT);
        for (int i = 0; i < 4; ++i) d.p[i] = p[i];
        d.q[0] = p[4];
        break;
    case RT1W_NODE_AABOX: // six rectangles merged by the commit (api.cu): p = min xyz, max xyz
        type = P_BOX;
        for (int i = 0; i < 3; ++i) d.p[i] = p[i], d.q[i] = p[3 + i];
        break;
    case RT1W_NODE_CONSTANT_MEDIUM:
        if (fp.boundary == RT1W_NODE_SPHERE) { // p = center, radius, -1/density
            type = P_MEDIUM_SPHERE;
            for (int i = 0; i < 4; ++i) d.p[i] = p[i];
            d.q[0] = p[4];
        } else { // p = min xyz, -1/density, max xyz
            type = P_MEDIUM_BOX;
            for (int i = 0; i < 4; ++i) d.p[i] = p[i];
            for (int i = 0; i < 3; ++i) d.q[i] = p[4 + i];
        }
        break;
    default: break;
    }
    d.meta = pack_meta(type, fp.flags, materials[fp.material].type, fp.material);
    d.frame = fp.frame;
    return d;
}

} // namespace rt1w
