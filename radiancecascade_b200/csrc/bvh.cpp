// bvh.cpp — binned surface-area-heuristic BVH2 builder (host).  Scenes here are
// <= ~35k triangles, so the build is a one-off millisecond-scale cost at rc_create.
#include "bvh.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <queue>

namespace rc {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const
    {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Prim { Box box; float c[3]; uint32_t id; };

// Bounding box of triangle (a, b, c) clipped to `cell` (Sutherland-Hodgman against the six planes, in double).
// Returns false when the clipped polygon is empty.  The result is grown by a relative 1e-6 of the cell (the
// caller's pad is added on top in write_node), so the union of the fragments' boxes always covers the triangle.
bool clipped_bounds(const float* a, const float* b, const float* c, const Box& cell, Box& out)
{
    double poly[16][3], tmp[16][3];
    int n = 3;
    for (int k = 0; k < 3; k++) { poly[0][k] = a[k]; poly[1][k] = b[k]; poly[2][k] = c[k]; }
    for (int axis = 0; axis < 3 && n; axis++)
        for (int side = 0; side < 2 && n; side++) {
            const double plane = side ? cell.hi[axis] : cell.lo[axis];
            const double sgn = side ? -1.0 : 1.0;   // inside: sgn * (x - plane) >= 0
            int m = 0;
            for (int i = 0; i < n; i++) {
                const double* p = poly[i];
                const double* q = poly[(i + 1) % n];
                const double dp = sgn * (p[axis] - plane), dq = sgn * (q[axis] - plane);
                if (dp >= 0.0) { for (int k = 0; k < 3; k++) tmp[m][k] = p[k]; m++; }
                if ((dp >= 0.0) != (dq >= 0.0)) {
                    const double t = dp / (dp - dq);
                    for (int k = 0; k < 3; k++) tmp[m][k] = p[k] + t * (q[k] - p[k]);
                    tmp[m][axis] = plane;
                    m++;
                }
            }
            n = m > 15 ? 15 : m;
            memcpy(poly, tmp, sizeof(double) * 3 * (size_t)n);
        }
    if (n == 0) return false;
    out.reset();
    for (int i = 0; i < n; i++) {
        float p[3] = {(float)poly[i][0], (float)poly[i][1], (float)poly[i][2]};
        out.grow(p);
    }
    for (int k = 0; k < 3; k++) {
        const float eps = 1e-6f * std::max(std::fabs(cell.hi[k] - cell.lo[k]), std::max(std::fabs(cell.lo[k]), std::fabs(cell.hi[k])));
        out.lo[k] = std::max(out.lo[k] - eps, cell.lo[k] - eps);
        out.hi[k] = std::min(out.hi[k] + eps, cell.hi[k] + eps);
    }
    return true;
}

// Triangle pre-splitting (early split clipping): the largest boxes are halved along their longest axis, each
// half keeping the clipped triangle's bounds, until `budget` extra references exist.  Walls and floors that
// span the whole scene otherwise put a scene-sized box around every node they fall into.  A triangle may then
// sit in several leaves; closest hits are min (t, id) over all tests (S5), so results do not change.
void presplit(std::vector<Prim>& prims, const float* v0, const float* e1, const float* e2, size_t budget, float min_area)
{
    auto cmp = [&](uint32_t x, uint32_t y) { return prims[x].box.area() < prims[y].box.area(); };
    std::priority_queue<uint32_t, std::vector<uint32_t>, decltype(cmp)> heap(cmp);
    for (uint32_t i = 0; i < prims.size(); i++) heap.push(i);
    size_t added = 0;
    while (!heap.empty() && added < budget) {
        const uint32_t i = heap.top();
        heap.pop();
        const Box bx = prims[i].box;
        if (!(bx.area() > min_area)) break;
        int axis = 0;
        for (int k = 1; k < 3; k++) if (bx.hi[k] - bx.lo[k] > bx.hi[axis] - bx.lo[axis]) axis = k;
        const float mid = 0.5f * (bx.lo[axis] + bx.hi[axis]);
        if (!(mid > bx.lo[axis] && mid < bx.hi[axis])) continue;   // cannot be halved any further
        const uint32_t t = prims[i].id;
        const float a[3] = {v0[3 * t], v0[3 * t + 1], v0[3 * t + 2]};
        const float b[3] = {a[0] + e1[3 * t], a[1] + e1[3 * t + 1], a[2] + e1[3 * t + 2]};
        const float c[3] = {a[0] + e2[3 * t], a[1] + e2[3 * t + 1], a[2] + e2[3 * t + 2]};
        Box lc = bx, rc_ = bx, lb, rb;
        lc.hi[axis] = mid;
        rc_.lo[axis] = mid;
        const bool hl = clipped_bounds(a, b, c, lc, lb), hr = clipped_bounds(a, b, c, rc_, rb);
        if (!hl && !hr) continue;            // numerically empty: keep the fragment as it is
        auto set = [&](Prim& p, const Box& nb) {
            p.box = nb;
            for (int k = 0; k < 3; k++) p.c[k] = 0.5f * (nb.lo[k] + nb.hi[k]);
        };
        if (hl && hr) {
            Prim q = prims[i];
            set(prims[i], lb);
            set(q, rb);
            prims.push_back(q);
            added++;
            heap.push(i);
            heap.push((uint32_t)prims.size() - 1);
        } else {
            const Box& nb = hl ? lb : rb;
            const bool shrunk = nb.area() < bx.area();
            set(prims[i], nb);
            if (shrunk) heap.push(i);       // tighter than before: may still be worth splitting
        }
    }
}

struct Builder {
    std::vector<Prim> prims;
    Bvh* out;
    float pad;

    static constexpr int kBins = 16;
    int kLeaf = 4;          // largest leaf (<= 7: the leaf link stores the count in 3 bits)
    float node_cost = 0.f;  // > 0: SAH termination — make a leaf of <= kLeaf triangles when splitting does not pay

    // Returns the child link (inner index or encoded leaf) for prims[first, first+count).
    int32_t build(uint32_t first, uint32_t count, Box& box_out, int depth)
    {
        Box box, cbox;
        box.reset();
        cbox.reset();
        for (uint32_t i = 0; i < count; i++) { box.grow(prims[first + i].box); cbox.grow(prims[first + i].c); }
        box_out = box;
        out->max_depth = std::max(out->max_depth, depth);
        if (count <= (uint32_t)(node_cost > 0.f ? 1 : kLeaf)) return make_leaf(first, count);

        int best_axis = -1, best_split = -1;
        float best_cost = FLT_MAX;
        for (int a = 0; a < 3; a++) {
            float ext = cbox.hi[a] - cbox.lo[a];
            if (!(ext > 0.f)) continue;
            Box bb[kBins];
            int bc[kBins] = {0};
            for (auto& b : bb) b.reset();
            float scale = kBins / ext;
            for (uint32_t i = 0; i < count; i++) {
                const Prim& p = prims[first + i];
                int b = std::min(kBins - 1, std::max(0, (int)((p.c[a] - cbox.lo[a]) * scale)));
                bb[b].grow(p.box);
                bc[b]++;
            }
            float la[kBins], ra[kBins];
            int lc[kBins], rc_[kBins];
            Box acc;
            acc.reset();
            int n = 0;
            for (int b = 0; b < kBins; b++) { acc.grow(bb[b]); n += bc[b]; la[b] = acc.area(); lc[b] = n; }
            acc.reset();
            n = 0;
            for (int b = kBins - 1; b >= 0; b--) { acc.grow(bb[b]); n += bc[b]; ra[b] = acc.area(); rc_[b] = n; }
            for (int b = 0; b < kBins - 1; b++) {
                if (lc[b] == 0 || rc_[b + 1] == 0) continue;
                float cost = la[b] * lc[b] + ra[b + 1] * rc_[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
            }
        }
        if (node_cost > 0.f && count <= (uint32_t)kLeaf && best_axis >= 0) {
            const float area = box.area();
            if (area > 0.f && node_cost + best_cost / area >= (float)count) return make_leaf(first, count);
        }
        uint32_t mid;
        if (best_axis < 0) {
            mid = first + count / 2;  // all centroids coincide: split by order
        } else {
            float ext = cbox.hi[best_axis] - cbox.lo[best_axis];
            float scale = kBins / ext, lo = cbox.lo[best_axis];
            int a = best_axis, s = best_split;
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const Prim& p) {
                int b = std::min(kBins - 1, std::max(0, (int)((p.c[a] - lo) * scale)));
                return b <= s;
            });
            mid = (uint32_t)(it - prims.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        int32_t idx = (int32_t)out->nodes.size();
        out->nodes.emplace_back();
        Box b0, b1;
        int32_t c0 = build(first, mid - first, b0, depth + 1);
        int32_t c1 = build(mid, first + count - mid, b1, depth + 1);
        write_node(idx, b0, b1, c0, c1);
        return idx;
    }

    int32_t make_leaf(uint32_t first, uint32_t count)
    {
        uint32_t at = (uint32_t)out->leaf_tris.size();
        uint32_t n = 0;
        for (uint32_t i = 0; i < count; i++) {   // two fragments of one pre-split triangle may share a leaf
            bool dup = false;
            for (uint32_t j = 0; j < n; j++) dup = dup || out->leaf_tris[at + j] == prims[first + i].id;
            if (!dup) { out->leaf_tris.push_back(prims[first + i].id); n++; }
        }
        return ~(int32_t)((at << 3) | n);
    }

    void write_node(int32_t idx, const Box& a, const Box& b, int32_t c0, int32_t c1)
    {
        float* q = out->nodes[idx].q;
        // an empty child is the point (1e30,1e30,1e30): the sorted-slab test can never reach it
        auto lo = [&](const Box& x, int k) { return x.lo[k] <= x.hi[k] ? x.lo[k] - pad : 1e30f; };
        auto hi = [&](const Box& x, int k) { return x.lo[k] <= x.hi[k] ? x.hi[k] + pad : 1e30f; };
        q[0] = lo(a, 0); q[1] = lo(a, 1); q[2] = lo(a, 2); q[3] = hi(a, 0);
        q[4] = hi(a, 1); q[5] = hi(a, 2); q[6] = lo(b, 0); q[7] = lo(b, 1);
        q[8] = lo(b, 2); q[9] = hi(b, 0); q[10] = hi(b, 1); q[11] = hi(b, 2);
        memcpy(&q[12], &c0, 4);
        memcpy(&q[13], &c1, 4);
        q[14] = q[15] = 0.f;
    }
};

}  // namespace

void build_bvh(const float* v0, const float* e1, const float* e2, const uint8_t* skip, uint32_t n_tris,
               float pad, Bvh& out, int max_leaf, float node_cost, float split_budget)
{
    out = Bvh();
    Builder b;
    b.out = &out;
    b.pad = pad;
    b.kLeaf = max_leaf < 1 ? 1 : (max_leaf > 7 ? 7 : max_leaf);
    b.node_cost = node_cost;
    b.prims.reserve(n_tris);
    for (uint32_t t = 0; t < n_tris; t++) {
        if (skip && skip[t]) continue;
        Prim p;
        p.id = t;
        p.box.reset();
        float a[3] = {v0[3 * t], v0[3 * t + 1], v0[3 * t + 2]};
        float bq[3] = {a[0] + e1[3 * t], a[1] + e1[3 * t + 1], a[2] + e1[3 * t + 2]};
        float cq[3] = {a[0] + e2[3 * t], a[1] + e2[3 * t + 1], a[2] + e2[3 * t + 2]};
        p.box.grow(a);
        p.box.grow(bq);
        p.box.grow(cq);
        bool finite = true;
        for (int k = 0; k < 3; k++) finite = finite && std::isfinite(p.box.lo[k]) && std::isfinite(p.box.hi[k]);
        if (!finite) continue;
        for (int k = 0; k < 3; k++) p.c[k] = 0.5f * (p.box.lo[k] + p.box.hi[k]);
        b.prims.push_back(p);
    }
    if (split_budget > 0.f && !b.prims.empty()) {
        Box scene;
        scene.reset();
        for (const Prim& p : b.prims) scene.grow(p.box);
        presplit(b.prims, v0, e1, e2, (size_t)(split_budget * (float)b.prims.size()), 1e-5f * scene.area());
    }
    // Root is always an inner node (index 0) so traversal starts uniformly.
    out.nodes.emplace_back();
    Box empty;
    empty.reset();
    uint32_t n = (uint32_t)b.prims.size();
    if (n == 0) {
        b.write_node(0, empty, empty, ~0, ~0);
        return;
    }
    if (n <= (uint32_t)b.kLeaf) {
        b.node_cost = 0.f;
        Box bx;
        int32_t c0 = b.build(0, n, bx, 1);
        b.write_node(0, bx, empty, c0, ~0);
        return;
    }
    // build() allocates its own inner node; make node 0 that node by building the halves here
    out.nodes.clear();
    Box bx;
    int32_t root = b.build(0, n, bx, 0);
    (void)root;  // == 0 because the first inner node emplaced is the root
}

}  // namespace rc
