// Scene description -> flat device layout.
//
// Replaces the pointer tree the reference traverses (Box<dyn Hittable> of HittableList / BHV / Translate /
// Rotate / ConstantMedium; src/hittable.rs:37-68, src/bhv.rs:85-165, src/transforms.rs, src/volumes.rs):
//   * Translate/Rotate chains are composed into one rigid instance per primitive; spheres are moved to
//     world space outright (a rigid transform of a sphere is a sphere), boxes keep an object-space record.
//   * HittableList and BHV are transparent: all surface primitives go into ONE SAH-built BVH (the reference's
//     random-axis median split, bhv.rs:122-145, only matters for its RNG consumption, which worlds.cpp replays).
//   * ConstantMedium nodes are pulled out into a short media list evaluated after the surface search.
#include "flatten.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>

namespace rtb {
namespace {

const double kPi = 3.14159265358979323846264338327950288;

struct M3 {
    double m[9];
    static M3 identity() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
    M3 operator*(const M3& o) const {
        M3 r{};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += m[3 * i + k] * o.m[3 * k + j];
                r.m[3 * i + j] = s;
            }
        return r;
    }
    M3 transposed() const { return M3{{m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]}}; }
    void apply(const double v[3], double out[3]) const {
        for (int i = 0; i < 3; ++i) out[i] = m[3 * i] * v[0] + m[3 * i + 1] * v[1] + m[3 * i + 2] * v[2];
    }
};

// matrix of Rotate::rotate (object -> world), transforms.rs:116-124
M3 rotation_of(int axis, double degrees) {
    int a1 = axis, a2 = (a1 + 1) % 3, a0 = (a1 + 2) % 3;
    double theta = degrees * kPi / 180.0;
    double s = std::sin(theta), c = std::cos(theta);
    M3 r{};
    r.m[3 * a1 + a1] = 1.0;
    r.m[3 * a0 + a0] = c, r.m[3 * a0 + a2] = s;
    r.m[3 * a2 + a0] = -s, r.m[3 * a2 + a2] = c;
    return r;
}

struct Wrapper {
    bool is_rotate;
    M3 rot;         // rotate: its matrix
    double off[3];  // translate: its offset
};

struct Box3 {
    double lo[3], hi[3];
    void reset() {
        for (int i = 0; i < 3; ++i) lo[i] = std::numeric_limits<double>::infinity(), hi[i] = -lo[i];
    }
    void grow(const double p[3]) {
        for (int i = 0; i < 3; ++i) lo[i] = std::min(lo[i], p[i]), hi[i] = std::max(hi[i], p[i]);
    }
    void grow(const Box3& b) { grow(b.lo), grow(b.hi); }
    double area() const {
        double d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        return 2.0 * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    }
};

float round_down(double x) {
    float f = (float)x;
    return (double)f > x ? std::nextafterf(f, -std::numeric_limits<float>::infinity()) : f;
}
float round_up(double x) {
    float f = (float)x;
    return (double)f < x ? std::nextafterf(f, std::numeric_limits<float>::infinity()) : f;
}

class Flattener {
   public:
    Flattener(const RtSceneDesc* d, FlatScene& out, std::string& err) : d(d), out(out), err(err) {}
    const RtSceneDesc* d;
    FlatScene& out;
    std::string& err;
    std::vector<Box3> prim_box;
    int status = RT_OK;

    bool fail(int code, const std::string& msg) {
        if (status == RT_OK) status = code, err = msg;
        return false;
    }

    std::map<std::vector<double>, int> inst_cache;  // one instance record per distinct wrapper chain

    int instance_of(const std::vector<Wrapper>& chain) {
        if (chain.empty()) return 0;
        std::vector<double> key;
        for (const Wrapper& w : chain) {
            key.push_back(w.is_rotate ? 1.0 : 0.0);
            key.insert(key.end(), w.rot.m, w.rot.m + 9);
            key.insert(key.end(), w.off, w.off + 3);
        }
        auto found = inst_cache.find(key);
        if (found != inst_cache.end()) return found->second;
        M3 R = M3::identity();
        double T[3] = {0, 0, 0};
        for (const Wrapper& w : chain) {
            if (w.is_rotate) {
                R = R * w.rot;
            } else {
                double t[3];
                R.apply(w.off, t);
                for (int i = 0; i < 3; ++i) T[i] += t[i];
            }
        }
        // how the two outermost wrappers re-face-forward the normal (see DInstance in rt_types.h)
        M3 m1 = chain[0].is_rotate ? chain[0].rot.transposed() : M3::identity();
        M3 m2 = M3::identity();
        if (chain.size() >= 2 && chain[1].is_rotate) {
            M3 r1 = chain[0].is_rotate ? chain[0].rot : M3::identity();
            m2 = r1 * chain[1].rot.transposed() * r1.transposed();
        }
        DInstance di{};
        for (int i = 0; i < 9; ++i) di.rot[i] = (float)R.m[i], di.m1[i] = (float)m1.m[i], di.m2[i] = (float)m2.m[i];
        for (int i = 0; i < 3; ++i) di.trans[i] = (float)T[i];
        out.inst.push_back(di);
        inst_R.push_back(R);
        inst_T.push_back({T[0], T[1], T[2]});
        if (out.inst.size() > PRIM_INST_MASK) fail(RT_ERR_UNSUPPORTED, "too many transform instances");
        inst_cache[key] = (int)out.inst.size();
        return (int)out.inst.size();  // index + 1
    }
    std::vector<M3> inst_R;
    struct T3 {
        double v[3];
    };
    std::vector<T3> inst_T;

    bool valid_material(int m) { return m >= 0 && m < d->n_materials; }

    // build the device record of one primitive node under `chain`; false if the node is not a primitive
    bool make_prim(const RtNode& n, const std::vector<Wrapper>& chain, DPrim& p, Box3& box) {
        std::memset(&p, 0, sizeof p);
        p.mat = n.material;
        box.reset();
        if (n.kind == RT_NODE_SPHERE) {
            if (!valid_material(n.material)) return fail(RT_ERR_INVALID, "sphere without a material");
            int inst = instance_of(chain);
            double c[3] = {n.f[0], n.f[1], n.f[2]}, r = n.f[3];
            if (inst) {
                double cw[3];
                inst_R[inst - 1].apply(c, cw);
                for (int i = 0; i < 3; ++i) c[i] = cw[i] + inst_T[inst - 1].v[i];
            }
            p.meta = PRIM_SPHERE | ((uint32_t)inst << PRIM_INST_SHIFT);
            for (int i = 0; i < 3; ++i) p.v[i] = (float)c[i];
            p.v[3] = (float)r;
            if (std::fabs(r) >= RTB_BIG_SPHERE_RADIUS) {
                p.meta |= PRIM_BIG;
                uint32_t idx = (uint32_t)out.big.size();
                std::memcpy(&p.v[4], &idx, 4);
                out.big.push_back(DBigSphere{{c[0], c[1], c[2]}, r});
            }
            double ar = std::fabs(r);
            double a[3] = {c[0] - ar, c[1] - ar, c[2] - ar}, b[3] = {c[0] + ar, c[1] + ar, c[2] + ar};
            box.grow(a), box.grow(b);
            return true;
        }
        if (n.kind == RT_NODE_MOVING_SPHERE) {  // EXTENSION: The Next Week's MovingSphere, time in [0, 1]
            if (!valid_material(n.material)) return fail(RT_ERR_INVALID, "moving sphere without a material");
            int inst = instance_of(chain);
            double c0[3] = {n.f[0], n.f[1], n.f[2]}, c1[3] = {n.f[3], n.f[4], n.f[5]}, r = n.f[6];
            if (inst) {
                double w0[3], w1[3];
                inst_R[inst - 1].apply(c0, w0), inst_R[inst - 1].apply(c1, w1);
                for (int i = 0; i < 3; ++i) c0[i] = w0[i] + inst_T[inst - 1].v[i], c1[i] = w1[i] + inst_T[inst - 1].v[i];
            }
            p.meta = PRIM_SPHERE | PRIM_MOVING | ((uint32_t)inst << PRIM_INST_SHIFT);
            for (int i = 0; i < 3; ++i) p.v[i] = (float)c0[i];
            p.v[3] = (float)r;
            uint32_t idx = (uint32_t)(out.moving.size() / 4);
            std::memcpy(&p.v[4], &idx, 4);
            for (int i = 0; i < 3; ++i) out.moving.push_back((float)(c1[i] - c0[i]));
            out.moving.push_back(0.0f);
            double ar = std::fabs(r) * (1.0 + 1e-6) + 1e-6;  // the f32 centre moves along a rounded displacement
            for (const double* c : {c0, c1}) {
                double a[3] = {c[0] - ar, c[1] - ar, c[2] - ar}, b[3] = {c[0] + ar, c[1] + ar, c[2] + ar};
                box.grow(a), box.grow(b);
            }
            return true;
        }
        double lo[3], hi[3];
        uint32_t rect_axis = 0;
        if (n.kind == RT_NODE_XYRECT || n.kind == RT_NODE_XZRECT || n.kind == RT_NODE_YZRECT) {
            // AARect::new (aarects.rs:31-43) incl. its one-sided normalisation of the second axis
            int a0 = n.kind == RT_NODE_YZRECT ? 1 : 0;
            int a1 = n.kind == RT_NODE_XYRECT ? 1 : 2;
            int ap = 3 - a0 - a1;
            lo[a0] = std::min(n.f[0], n.f[1]), hi[a0] = std::max(n.f[1], n.f[0]);
            lo[a1] = n.f[2], hi[a1] = std::max(n.f[3], n.f[2]);
            lo[ap] = hi[ap] = n.f[4];
            rect_axis = (uint32_t)ap + 1;
        } else if (n.kind == RT_NODE_BLOCK) {
            for (int i = 0; i < 3; ++i) lo[i] = n.f[i], hi[i] = n.f[3 + i];  // ordered: see block_sides for the other case
        } else {
            return false;
        }
        if (!valid_material(n.material)) return fail(RT_ERR_INVALID, "rect/block without a material");
        int inst = instance_of(chain);
        p.meta = PRIM_BOX | ((uint32_t)inst << PRIM_INST_SHIFT) | (rect_axis << PRIM_RECT_SHIFT);
        for (int i = 0; i < 3; ++i) p.v[i] = (float)lo[i], p.v[3 + i] = (float)hi[i];
        for (int corner = 0; corner < 8; ++corner) {
            double q[3] = {corner & 1 ? hi[0] : lo[0], corner & 2 ? hi[1] : lo[1], corner & 4 ? hi[2] : lo[2]};
            if (inst) {
                double w[3];
                inst_R[inst - 1].apply(q, w);
                for (int i = 0; i < 3; ++i) q[i] = w[i] + inst_T[inst - 1].v[i];
            }
            box.grow(q);
        }
        return true;
    }

    // Block::new (shapes.rs:173-186) with corners that are NOT ordered: the reference builds six rects whose bounds go
    // through AARect::new's one-sided normalisation (aarects.rs:31-43: the second axis keeps its first value as minimum),
    // so some sides degenerate to a line.  Such a block is flattened as those six rects, not as one slab primitive.
    static bool block_is_ordered(const RtNode& n) { return n.f[0] <= n.f[3] && n.f[1] <= n.f[4] && n.f[2] <= n.f[5]; }
    static void block_sides(const RtNode& b, RtNode out6[6]) {
        const double* p0 = b.f;
        const double* p1 = b.f + 3;
        const double sides[6][6] = {
            {RT_NODE_XYRECT, p0[0], p1[0], p0[1], p1[1], p1[2]}, {RT_NODE_XYRECT, p0[0], p1[0], p0[1], p1[1], p0[2]},
            {RT_NODE_XZRECT, p0[0], p1[0], p0[2], p1[2], p0[1]}, {RT_NODE_XZRECT, p0[0], p1[0], p0[2], p1[2], p1[1]},
            {RT_NODE_YZRECT, p0[1], p1[1], p0[2], p1[2], p0[0]}, {RT_NODE_YZRECT, p0[1], p1[1], p0[2], p1[2], p1[0]}};
        for (int i = 0; i < 6; ++i) {
            out6[i] = b;
            out6[i].kind = (int32_t)sides[i][0];
            for (int k = 0; k < 5; ++k) out6[i].f[k] = sides[i][1 + k];
        }
    }

    // the surface primitives below `node` (lists, BVHs and transforms are transparent), appended to prims/boxes
    bool collect_prims(int node, std::vector<Wrapper>& chain, int depth, std::vector<DPrim>& prims, std::vector<Box3>* boxes, std::vector<int32_t>* nodes_out) {
        if (status != RT_OK) return false;
        if (node < 0 || node >= d->n_nodes) return fail(RT_ERR_INVALID, "node index out of range");
        if (depth > 64) return fail(RT_ERR_INVALID, "description nests deeper than 64 levels (cycle?)");
        const RtNode& n = d->nodes[node];
        switch (n.kind) {
            case RT_NODE_TRANSLATE: {
                chain.push_back(Wrapper{false, M3::identity(), {n.f[0], n.f[1], n.f[2]}});
                bool ok = collect_prims(n.first_child, chain, depth + 1, prims, boxes, nodes_out);
                chain.pop_back();
                return ok;
            }
            case RT_NODE_ROTATE: {
                if (n.axis < 0 || n.axis > 2) return fail(RT_ERR_INVALID, "rotate axis must be 0, 1 or 2");
                chain.push_back(Wrapper{true, rotation_of(n.axis, n.f[0]), {0, 0, 0}});
                bool ok = collect_prims(n.first_child, chain, depth + 1, prims, boxes, nodes_out);
                chain.pop_back();
                return ok;
            }
            case RT_NODE_BVH:
            case RT_NODE_LIST: {
                if (n.first_child < 0 || n.child_count < 0 || (long long)n.first_child + n.child_count > (long long)d->n_children)
                    return fail(RT_ERR_INVALID, "list/bvh child range out of bounds");
                for (int i = 0; i < n.child_count; ++i)
                    if (!collect_prims(d->children[n.first_child + i], chain, depth + 1, prims, boxes, nodes_out)) return false;
                return true;
            }
            case RT_NODE_MEDIUM: return fail(RT_ERR_UNSUPPORTED, "a constant medium cannot be part of another medium's boundary");
            default: {
                RtNode parts[6];
                int n_parts = 1;
                parts[0] = n;
                if (n.kind == RT_NODE_BLOCK && !block_is_ordered(n)) block_sides(n, parts), n_parts = 6;
                for (int i = 0; i < n_parts; ++i) {
                    DPrim p;
                    Box3 b;
                    if (!make_prim(parts[i], chain, p, b)) {
                        if (status == RT_OK) fail(RT_ERR_INVALID, "unknown node kind");
                        return false;
                    }
                    prims.push_back(p);
                    if (boxes) boxes->push_back(b);
                    if (nodes_out) nodes_out->push_back(node);
                }
                return true;
            }
        }
    }

    void walk(int node, std::vector<Wrapper>& chain, int depth) {
        if (status != RT_OK) return;
        if (node < 0 || node >= d->n_nodes) {
            fail(RT_ERR_INVALID, "node index out of range");
            return;
        }
        if (depth > 64) {
            fail(RT_ERR_INVALID, "description nests deeper than 64 levels (cycle?)");
            return;
        }
        const RtNode& n = d->nodes[node];
        switch (n.kind) {
            case RT_NODE_TRANSLATE: {
                chain.push_back(Wrapper{false, M3::identity(), {n.f[0], n.f[1], n.f[2]}});
                walk(n.first_child, chain, depth + 1);
                chain.pop_back();
                break;
            }
            case RT_NODE_ROTATE: {
                if (n.axis < 0 || n.axis > 2) {
                    fail(RT_ERR_INVALID, "rotate axis must be 0, 1 or 2");
                    return;
                }
                chain.push_back(Wrapper{true, rotation_of(n.axis, n.f[0]), {0, 0, 0}});
                walk(n.first_child, chain, depth + 1);
                chain.pop_back();
                break;
            }
            case RT_NODE_BVH:
            case RT_NODE_LIST: {
                if (n.first_child < 0 || n.child_count < 0 || (long long)n.first_child + n.child_count > (long long)d->n_children) {
                    fail(RT_ERR_INVALID, "list/bvh child range out of bounds");
                    return;
                }
                for (int i = 0; i < n.child_count; ++i) walk(d->children[n.first_child + i], chain, depth + 1);
                break;
            }
            case RT_NODE_MEDIUM: {
                if (!valid_material(n.material) || d->materials[n.material].kind != RT_MAT_ISOTROPIC) {
                    fail(RT_ERR_INVALID, "medium needs an isotropic phase material");
                    return;
                }
                if (!(n.f[0] > 0.0)) {
                    fail(RT_ERR_INVALID, "medium density must be positive");
                    return;
                }
                // ConstantMedium's boundary is any Hittable (volumes.rs:7-11): every surface primitive below it
                std::vector<DPrim> boundary;
                std::vector<Box3> boundary_box;
                if (!collect_prims(n.first_child, chain, depth + 1, boundary, &boundary_box, nullptr)) return;
                if (boundary.empty()) return;  // an empty boundary is never hit: the medium does nothing
                if (boundary.size() > 0xFFFF || out.media.size() >= RTB_MAX_MEDIA) {
                    fail(RT_ERR_UNSUPPORTED, "too many constant media / boundary primitives");
                    return;
                }
                DMedium m{};
                m.boundary = boundary[0];
                m.neg_inv_density = (float)(-1.0 / n.f[0]);
                m.mat = n.material;
                m.first = (int32_t)out.media_prims.size();
                m.count = (int32_t)boundary.size();
                out.media_prims.insert(out.media_prims.end(), boundary.begin(), boundary.end());
                out.media.push_back(m);
                out.media_node.push_back(node);
                media_box.push_back(boundary_box[0]);
                break;
            }
            default: collect_prims(node, chain, depth, out.prims, &prim_box, &out.prim_node);
        }
    }

    // ---- "clear" media: the interior of the boundary holds no surface.  A ray that starts inside such a medium and draws
    // its next free-flight event inside it too cannot hit a surface first (the boundary is one convex primitive, so the
    // whole segment stays inside), hence the persistent kernel advances such paths from event to event without a
    // surface search (rt_persist.cu, chain phase).  Conservative test on the primitives as emitted (before the BVH build
    // reorders them): a surface primitive disqualifies the medium when its world box reaches into the interior of the
    // boundary's world box — for a spherical boundary: into the ball; a sphere is tested exactly against it, so that the
    // boundary itself, added to the world as a glass shell (final_scene, worlds.rs:430-437), does not count.
    std::vector<Box3> media_box;
    void find_clear_media() {
        out.clear_media = 0u;
        for (size_t m = 0; m < out.media.size() && m < 32; ++m) {
            const DMedium& M = out.media[m];
            if (M.count != 1 || out.mats[M.mat].tex >= 0) continue;  // compound boundary / textured phase function: general path
            const DPrim& B = M.boundary;
            const bool ball = (B.meta & PRIM_KIND_MASK) == PRIM_SPHERE && !(B.meta & PRIM_MOVING);
            if (!ball && (B.meta & PRIM_KIND_MASK) != PRIM_BOX) continue;
            if (!ball && ((B.meta >> PRIM_RECT_SHIFT) & PRIM_RECT_MASK)) continue;  // a rect encloses nothing
            const Box3& bb = media_box[m];
            bool clear = true;
            for (size_t i = 0; i < out.prims.size() && clear; ++i) {
                const DPrim& q = out.prims[i];
                const Box3& qb = prim_box[i];
                bool reaches = true;
                for (int k = 0; k < 3; ++k) reaches = reaches && qb.lo[k] < bb.hi[k] && qb.hi[k] > bb.lo[k];
                if (!reaches) continue;
                if (ball) {
                    const double r = std::fabs((double)B.v[3]);
                    double d2 = 0.0;  // squared distance from the centre to q's box
                    for (int k = 0; k < 3; ++k) {
                        const double c = (double)B.v[k], d = c < qb.lo[k] ? qb.lo[k] - c : (c > qb.hi[k] ? c - qb.hi[k] : 0.0);
                        d2 += d * d;
                    }
                    if (d2 >= r * r) continue;
                    const bool q_ball = (q.meta & PRIM_KIND_MASK) == PRIM_SPHERE && !(q.meta & PRIM_MOVING);
                    if (q_ball) {  // sphere against ball, exactly: the surface of q comes within |dist - |r_q|| of the centre
                        double dist2 = 0.0;
                        for (int k = 0; k < 3; ++k) dist2 += ((double)q.v[k] - (double)B.v[k]) * ((double)q.v[k] - (double)B.v[k]);
                        if (std::fabs(std::sqrt(dist2) - std::fabs((double)q.v[3])) >= r) continue;  // (the coincident shell: == r)
                    }
                }
                clear = false;
            }
            if (clear) out.clear_media |= 1u << m;
        }
        if (getenv("RT_NO_CLEAR_MEDIA")) out.clear_media = 0u;
    }

    // ---- F_* feature bits (rt_types.h): lets the persistent pipeline launch a kernel instance without the code this scene
    // cannot reach.  Conservative: textures and materials count when the description holds them, used or not.
    void find_features() {
        uint32_t f = 0u;
        auto prim_bits = [&](const DPrim& q, bool boundary) {
            const bool sphere = (q.meta & PRIM_KIND_MASK) == PRIM_SPHERE;
            const bool inst = ((q.meta >> PRIM_INST_SHIFT) & PRIM_INST_MASK) != 0;
            if (q.meta & PRIM_BIG) f |= F_BIG;
            if (q.meta & PRIM_MOVING) f |= F_MOVING;
            if (boundary) {
                if (!sphere) f |= F_BOXMEDIA;
                return;
            }
            f |= sphere ? F_SPHERE : F_BOX;
            if (inst) f |= F_INSTANCE | (sphere ? 0u : F_INSTBOX);
        };
        for (const DPrim& q : out.prims) prim_bits(q, false);
        for (const DPrim& q : out.media_prims) prim_bits(q, true);
        if (!out.media.empty()) f |= F_MEDIA;
        for (const DTexture& t : out.texs) f |= t.kind == TEX_CHECKER ? F_CHECKER : (t.kind == TEX_NOISE ? F_NOISE : (t.kind == TEX_IMAGE ? F_IMAGE : 0u));
        out.features = f;
    }

    // ---- SAH BVH over prim_box ----
    std::vector<int> order;
    static const int kMedianDepth = 20;
    int kMaxLeaf = 1;  // one primitive per leaf measured fastest (C2 +1 %, C3 +5 %, C4 +2 % over 4); RT_BVH_MAX_LEAF overrides

    Box3 bounds_of(int begin, int end) const {
        Box3 b;
        b.reset();
        for (int i = begin; i < end; ++i) b.grow(prim_box[order[i]]);
        return b;
    }

    void build_node(int idx, int begin, int end, int depth) {
        out.bvh_depth = std::max(out.bvh_depth, depth);
        Box3 b = bounds_of(begin, end);
        for (int i = 0; i < 3; ++i) out.nodes[idx].lo[i] = round_down(b.lo[i]), out.nodes[idx].hi[i] = round_up(b.hi[i]);
        int n = end - begin;
        double best_cost = std::numeric_limits<double>::infinity();
        int best_axis = -1, best_split = -1;
        if (n >= 2) {
            double parent_area = std::max(b.area(), 1e-300);
            std::vector<double> right_area(n);
            for (int axis = 0; axis < 3; ++axis) {
                std::stable_sort(order.begin() + begin, order.begin() + end, [&](int x, int y) {
                    return prim_box[x].lo[axis] + prim_box[x].hi[axis] < prim_box[y].lo[axis] + prim_box[y].hi[axis];
                });
                Box3 acc;
                acc.reset();
                for (int i = n - 1; i >= 1; --i) {
                    acc.grow(prim_box[order[begin + i]]);
                    right_area[i] = acc.area();
                }
                acc.reset();
                for (int i = 1; i < n; ++i) {
                    acc.grow(prim_box[order[begin + i - 1]]);
                    double cost = 1.0 + (acc.area() * i + right_area[i] * (n - i)) / parent_area;
                    if (cost < best_cost) best_cost = cost, best_axis = axis, best_split = i;
                }
            }
        }
        bool leaf = n <= kMaxLeaf && (n < 2 || (double)n <= best_cost);
        if (leaf) {
            out.nodes[idx].a = ~(begin | (n << 24)), out.nodes[idx].b = n;  // ready-made traversal link (n <= kMaxLeaf <= 8)
            return;
        }
        if (depth >= kMedianDepth || best_axis < 0) {
            // A skewed SAH tree (nested or concentric primitives) could outgrow the traversal stack.  From this depth on
            // the subtree is split at the median of the widest centroid axis: at most ceil(log2 n) <= 24 further levels,
            // so the whole tree stays below RTB_BVH_STACK (checked after the build).
            Box3 cb;
            cb.reset();
            for (int i = begin; i < end; ++i) {
                const Box3& pb = prim_box[order[i]];
                double c[3] = {pb.lo[0] + pb.hi[0], pb.lo[1] + pb.hi[1], pb.lo[2] + pb.hi[2]};
                cb.grow(c);
            }
            best_axis = 0;
            for (int a = 1; a < 3; ++a)
                if (cb.hi[a] - cb.lo[a] > cb.hi[best_axis] - cb.lo[best_axis]) best_axis = a;
            best_split = n / 2;
        }
        std::stable_sort(order.begin() + begin, order.begin() + end, [&](int x, int y) {
            return prim_box[x].lo[best_axis] + prim_box[x].hi[best_axis] < prim_box[y].lo[best_axis] + prim_box[y].hi[best_axis];
        });
        int left = (int)out.nodes.size();
        out.nodes.push_back(DNode{}), out.nodes.push_back(DNode{});
        out.nodes[idx].a = left, out.nodes[idx].b = 0;
        build_node(left, begin, begin + best_split, depth + 1);
        build_node(left + 1, begin + best_split, end, depth + 1);
    }

    void build_bvh() {
        if (const char* e = getenv("RT_BVH_MAX_LEAF")) kMaxLeaf = std::max(1, std::min(8, atoi(e)));
        int n = (int)out.prims.size();
        out.nodes.clear();
        out.nodes.push_back(DNode{});
        if (n == 0) {  // empty world: a leaf with no primitives
            out.nodes[0].a = ~0, out.nodes[0].b = 0;
            return;
        }
        order.resize(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        build_node(0, 0, n, 0);
        std::vector<DPrim> p2(n);
        std::vector<int32_t> n2(n);
        for (int i = 0; i < n; ++i) p2[i] = out.prims[order[i]], n2[i] = out.prim_node[order[i]];
        out.prims.swap(p2), out.prim_node.swap(n2);
        if (out.bvh_depth > RTB_BVH_STACK - 2) {
            fail(RT_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
            return;
        }
        collapse4();
    }

    // ---- 4-wide collapse of the binary tree (DNode4, rt_types.h).  Which binary nodes survive as wide nodes is chosen
    // by the dynamic programme of Ylitie, Karras & Laine (HPG 2017, section 4.1) for K = 4: cost[n][j] = the smallest
    // summed surface area of wide nodes with which subtree n can fill at most j child slots of its parent.  Leaves hold
    // exactly one primitive, so their cost is the same in every arrangement and drops out.  Only for scenes whose
    // links fit 16 bits; otherwise nodes4 stays empty and every pipeline walks the binary tree.
    static double area_of(const DNode& n) {
        double d[3] = {(double)n.hi[0] - n.lo[0], (double)n.hi[1] - n.lo[1], (double)n.hi[2] - n.lo[2]};
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    }
    struct Cost4 {
        double c[4];  // c[j], j = 1..3 (c[0] unused)
    };
    std::vector<Cost4> cost4;
    bool wide_overflow = false;
    double distribute4(int n2, int j) const {  // best split of j slots between the two children of binary node n2
        int l = out.nodes[n2].a, r = l + 1;
        double best = std::numeric_limits<double>::infinity();
        for (int k = 1; k < j; ++k) best = std::min(best, cost4[l].c[k] + cost4[r].c[j - k]);
        return best;
    }
    void solve4(int n2) {
        Cost4& me = cost4[n2];
        if (out.nodes[n2].a < 0) {
            me.c[1] = me.c[2] = me.c[3] = 0.0;
            return;
        }
        solve4(out.nodes[n2].a), solve4(out.nodes[n2].a + 1);
        me.c[1] = area_of(out.nodes[n2]) + distribute4(n2, 4);
        me.c[2] = std::min(me.c[1], distribute4(n2, 2));
        me.c[3] = std::min(me.c[2], distribute4(n2, 3));
    }
    void expand4(int n2, int j, std::vector<int>& kids) const {  // the binary nodes that fill <= j slots at cost4[n2].c[j]
        if (out.nodes[n2].a < 0 || j == 1) {
            kids.push_back(n2);
            return;
        }
        if (cost4[n2].c[j] == cost4[n2].c[j - 1]) return expand4(n2, j - 1, kids);
        int l = out.nodes[n2].a, r = l + 1;
        for (int k = 1; k < j; ++k)
            if (cost4[l].c[k] + cost4[r].c[j - k] == cost4[n2].c[j]) {
                expand4(l, k, kids), expand4(r, j - k, kids);
                return;
            }
        kids.push_back(n2);  // not reached (the minimum above is attained by some k)
    }
    int make4(int n2, int depth) {
        int me = (int)out.nodes4.size();
        out.nodes4.push_back(DNode4{});
        out.bvh4_depth = std::max(out.bvh4_depth, depth);
        if (me >= RTB_WIDE_MAX_NODES) wide_overflow = true;
        if (wide_overflow) return 0;
        std::vector<int> kids;
        if (out.nodes[n2].a < 0) {
            kids.push_back(n2);  // the whole world is one leaf
        } else {
            int l = out.nodes[n2].a, r = l + 1;
            double want = cost4[n2].c[1] - area_of(out.nodes[n2]);
            for (int k = 1; k < 4; ++k)
                if (cost4[l].c[k] + cost4[r].c[4 - k] == want) {
                    expand4(l, k, kids), expand4(r, 4 - k, kids);
                    break;
                }
            if (kids.empty()) kids.push_back(l), kids.push_back(r);
        }
        for (int c = 0; c < 4; ++c) {
            uint32_t link = RTB_LINK4_EMPTY;
            float box[6] = {0, 0, 0, 0, 0, 0};
            if (c < (int)kids.size()) {
                const DNode k = out.nodes[kids[c]];
                for (int a = 0; a < 3; ++a) box[a] = k.lo[a], box[3 + a] = k.hi[a];
                if (k.a < 0) {
                    int v = ~k.a;
                    link = RTB_LINK4_LEAF | (uint32_t)(v & 0xFFFFFF);  // exactly one primitive per leaf here
                } else {
                    link = (uint32_t)make4(kids[c], depth + 1);
                }
            }
            // centre and half-extent; the half-extent is rounded up so that [c - h, c + h] contains [lo, hi]
            float ctr[3], half[3];
            for (int a = 0; a < 3; ++a) {
                ctr[a] = (float)(0.5 * ((double)box[a] + (double)box[3 + a]));
                half[a] = round_up(std::max((double)box[3 + a] - (double)ctr[a], (double)ctr[a] - (double)box[a]));
            }
            DNode4& dst = out.nodes4[me];  // looked up after the recursion, which reallocates out.nodes4
            dst.ch[c][0] = ctr[0], dst.ch[c][1] = ctr[1], dst.ch[c][2] = half[0], dst.ch[c][3] = half[1];
            dst.cz[c] = ctr[2], dst.hz[c] = half[2];
            dst.link[c] = link;
        }
        return me;
    }
    void collapse4() {
        out.nodes4.clear();
        out.bvh4_depth = 0;
        int n = (int)out.prims.size();
        if (n == 0 || n > RTB_WIDE_MAX_PRIMS || kMaxLeaf != 1 || getenv("RT_BVH_NO_WIDE")) return;
        cost4.assign(out.nodes.size(), Cost4{});
        solve4(0);
        make4(0, 1);
        // a step pushes at most three entries, so 3 x depth bounds the stack
        if (wide_overflow || 3 * out.bvh4_depth > RTB_WIDE_STACK) out.nodes4.clear(), out.bvh4_depth = 0;
    }

    int checker_depth(int tex, int level) const {  // levels of checker above the deepest leaf; -1 = too deep / cyclic
        const RtTexture& t = d->textures[tex];
        if (t.kind != RT_TEX_CHECKER) return 0;
        if (level >= RTB_CHECKER_DEPTH) return -1;
        int a = checker_depth(t.a, level + 1), b = checker_depth(t.b, level + 1);
        return (a < 0 || b < 0) ? -1 : 1 + std::max(a, b);
    }
    bool leaf_needs_uv(int tex, int level) const {
        const RtTexture& t = d->textures[tex];
        if (t.kind == RT_TEX_IMAGE) return true;
        if (t.kind != RT_TEX_CHECKER || level >= RTB_CHECKER_DEPTH) return false;
        return leaf_needs_uv(t.a, level + 1) || leaf_needs_uv(t.b, level + 1);
    }

    bool tables() {
        for (int i = 0; i < d->n_materials; ++i) {
            const RtMaterial& m = d->materials[i];
            DMaterial dm{};
            dm.kind = m.kind, dm.tex = m.texture, dm.fuzz = (float)m.fuzz, dm.ior = (float)m.ior;
            for (int k = 0; k < 3; ++k) dm.albedo[k] = (float)m.albedo[k];
            bool textured = m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT || m.kind == RT_MAT_ISOTROPIC;
            if (m.kind < RT_MAT_LAMBERTIAN || m.kind > RT_MAT_ISOTROPIC) return fail(RT_ERR_INVALID, "unknown material kind");
            if (textured) {
                if (m.texture < 0 || m.texture >= d->n_textures) return fail(RT_ERR_INVALID, "material texture index out of range");
                const RtTexture& t = d->textures[m.texture];
                if (t.kind == RT_TEX_SOLID) {  // fold constant colours into the material record
                    dm.tex = -1;
                    for (int k = 0; k < 3; ++k) dm.albedo[k] = (float)t.color[k];
                }
            } else {
                dm.tex = -1;
            }
            out.mats.push_back(dm);
        }
        for (int i = 0; i < d->n_textures; ++i) {
            const RtTexture& t = d->textures[i];
            DTexture dt{};
            dt.kind = t.kind, dt.a = t.a, dt.b = t.b, dt.scale = (float)t.scale;
            for (int k = 0; k < 3; ++k) dt.color[k] = (float)t.color[k];
            if (t.kind == RT_TEX_CHECKER) {
                if (t.a < 0 || t.a >= d->n_textures || t.b < 0 || t.b >= d->n_textures) return fail(RT_ERR_INVALID, "checker child out of range");
            } else if (t.kind == RT_TEX_NOISE) {
                if (t.a < 0 || t.a >= d->n_perlins) return fail(RT_ERR_INVALID, "noise texture: perlin table out of range");
            } else if (t.kind == RT_TEX_IMAGE) {
                if (t.a < 0 || t.a >= d->n_images) return fail(RT_ERR_INVALID, "image texture: image out of range");
            } else if (t.kind != RT_TEX_SOLID) {
                return fail(RT_ERR_INVALID, "unknown texture kind");
            }
            out.texs.push_back(dt);
        }
        // Checker<Odd, Even> nests by type in the reference (textures.rs:28-49): the device resolves a chain of checkers
        // in a loop of at most RTB_CHECKER_DEPTH levels.  DTexture.needs_uv: some leaf below is an image (sphere_uv needed).
        for (int i = 0; i < d->n_textures; ++i) {
            int depth = checker_depth(i, 0);
            if (depth < 0) return fail(RT_ERR_UNSUPPORTED, "checker textures nest deeper than 8 levels (or form a cycle)");
            out.texs[i].needs_uv = leaf_needs_uv(i, 0) ? 1 : 0;
        }
        for (int i = 0; i < d->n_perlins; ++i) {
            const RtPerlin& p = d->perlins[i];
            for (int k = 0; k < RTB_PERLIN_POINTS; ++k) {
                for (int c = 0; c < 3; ++c) out.perlin_vec.push_back((float)p.ranvec[k][c]);
                out.perlin_vec.push_back(0.0f);
            }
            const int32_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
            for (const int32_t* perm : perms)
                for (int k = 0; k < RTB_PERLIN_POINTS; ++k) {
                    if (perm[k] < 0 || perm[k] >= RTB_PERLIN_POINTS) return fail(RT_ERR_INVALID, "perlin permutation entry out of range");
                    out.perlin_perm.push_back((unsigned short)perm[k]);
                }
        }
        for (int i = 0; i < d->n_images; ++i) {
            const RtImage& im = d->images[i];
            if (im.width <= 0 || im.height <= 0 || !im.rgb) return fail(RT_ERR_INVALID, "image without pixels");
            FlatScene::Image fi;
            fi.width = im.width, fi.height = im.height;
            fi.rgba.resize((size_t)4 * im.width * im.height);
            for (size_t px = 0; px < (size_t)im.width * im.height; ++px) {
                fi.rgba[4 * px] = im.rgb[3 * px], fi.rgba[4 * px + 1] = im.rgb[3 * px + 1], fi.rgba[4 * px + 2] = im.rgb[3 * px + 2];
                fi.rgba[4 * px + 3] = 255;
            }
            out.images.push_back(std::move(fi));
        }
        out.bg_kind = d->background_kind;
        for (int k = 0; k < 3; ++k) out.bg_top[k] = (float)d->background_top[k], out.bg_bottom[k] = (float)d->background_bottom[k];
        return true;
    }
};

}  // namespace

int flatten_scene(const RtSceneDesc* desc, int32_t root, int build, FlatScene& out, std::string& err) {
    if (!desc || desc->n_nodes <= 0 || !desc->nodes) {
        err = "empty scene description";
        return RT_ERR_INVALID;
    }
    if (desc->n_children > 0 && !desc->children) {
        err = "children array missing";
        return RT_ERR_INVALID;
    }
    Flattener f(desc, out, err);
    if (!f.tables()) return f.status;
    std::vector<Wrapper> chain;
    f.walk(root, chain, 0);
    if (f.status != RT_OK) return f.status;
    f.find_clear_media();
    f.find_features();
    size_t gpu_min = RTB_GPU_BUILD_MIN;
    if (const char* e = getenv("RT_BVH_GPU_MIN")) gpu_min = (size_t)std::max(2, atoi(e));
    const bool device_build = build == BUILD_AUTO && out.prims.size() >= gpu_min;
    if (build != BUILD_NONE && !device_build) f.build_bvh();
    if (f.status != RT_OK) return f.status;
    out.needs_device_build = device_build;
    out.prim_bounds.resize(6 * out.prims.size());
    for (size_t i = 0; i < out.prims.size(); ++i) {  // after build_bvh the primitives are in leaf order: prims[i] <-> prim_box[order[i]]
        const Box3& b = f.prim_box[f.order.empty() ? i : (size_t)f.order[i]];
        for (int k = 0; k < 3; ++k) out.prim_bounds[6 * i + k] = round_down(b.lo[k]), out.prim_bounds[6 * i + 3 + k] = round_up(b.hi[k]);
    }
    return f.status;
}

void make_camera(const RtCamera& in, DCamera& out) {
    // Camera::new, camera.rs:15-38, all in f64 like the reference; only the result is narrowed
    auto sub = [](const double a[3], const double b[3], double r[3]) { for (int i = 0; i < 3; ++i) r[i] = a[i] - b[i]; };
    auto unit = [](double v[3]) {
        double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int i = 0; i < 3; ++i) v[i] /= l;
    };
    auto cross = [](const double a[3], const double b[3], double r[3]) {
        r[0] = a[1] * b[2] - a[2] * b[1], r[1] = a[2] * b[0] - a[0] * b[2], r[2] = a[0] * b[1] - a[1] * b[0];
    };
    double theta = in.vfov_deg * kPi / 180.0;
    double h = std::tan(theta / 2.0);
    double vh = 2.0 * h, vw = in.aspect_ratio * vh;
    double w[3], u[3], v[3];
    sub(in.lookfrom, in.lookat, w);
    unit(w);
    cross(in.vup, w, u);
    unit(u);
    cross(w, u, v);
    for (int i = 0; i < 3; ++i) {
        double hor = in.focus_dist * vw * u[i], ver = in.focus_dist * vh * v[i];
        out.origin[i] = (float)in.lookfrom[i];
        out.horizontal[i] = (float)hor, out.vertical[i] = (float)ver;
        // stored relative to the origin: get_ray only ever uses lower_left_corner - origin (camera.rs:46)
        double llc = in.lookfrom[i] - hor / 2.0 - ver / 2.0 - in.focus_dist * w[i];
        out.lower_left[i] = (float)(llc - in.lookfrom[i]);
        out.u[i] = (float)u[i], out.v[i] = (float)v[i];
    }
    out.lens_radius = (float)(in.aperture / 2.0);
    out.time0 = (float)std::min(std::max(in.time0, 0.0), 1.0), out.time1 = (float)std::min(std::max(in.time1, 0.0), 1.0);
}

}  // namespace rtb
