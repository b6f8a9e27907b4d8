// bvh.h — host-side binned-SAH BVH2 build; flattened for the CUDA traversal kernels.
// The acceleration structure is product-internal: closest hits are defined by
// rc_spec.h S5 independently of it (boxes are padded so culling is conservative).
#pragma once
#include <cstdint>
#include <vector>

namespace rc {

// One 64-byte node = both children's boxes + child links, fetched as 4 x float4.
//   q0 = (lo0.x, lo0.y, lo0.z, hi0.x)   q1 = (hi0.y, hi0.z, lo1.x, lo1.y)
//   q2 = (lo1.z, hi1.x, hi1.y, hi1.z)   q3 = (child0, child1, -, -) as int bits
// child >= 0: inner node index; child < 0: leaf, ~child = (first << 3) | count, count <= 4,
// `first` indexing the leaf-ordered triangle arrays.
struct BvhNode { float q[16]; };

struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> leaf_tris;  // leaf order -> global triangle id
    int max_depth = 0;
};

// v0/e1/e2: per-triangle geometry (3 floats each); skip[t] != 0 excludes a triangle (S5).
// split_budget > 0: pre-split the largest triangles' boxes, adding up to split_budget * n_tris references
// (a triangle can then appear in several leaves: leaf_tris.size() >= number of triangles).
void build_bvh(const float* v0, const float* e1, const float* e2, const uint8_t* skip, uint32_t n_tris,
               float pad, Bvh& out, int max_leaf = 4, float node_cost = 0.f, float split_budget = 0.f);

}  // namespace rc
