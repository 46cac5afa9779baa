// Host-side SAH BVH construction for the flattened scene.
//
// Mirrors BVHAccel::new / recursive_build / split_sah / flatten_bvhtree of
// pbrt-rust (src/accelerators/bvh.rs:145-375, 662-693) so that the node array and
// the primitive order handed to the device are the ones the Rust host would hand
// over.  The structure is not a transcription: the tree is grown with an explicit
// work stack (right range first, bvh.rs:275-276, which fixes `ordered` order) and
// is emitted straight into the linear array by a second explicit-stack pass
// (first child at i+1, second child at `offset`).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/pbrt_b200.h"
#include "error.h"

namespace {

struct Box {
    float lo[3], hi[3];
    Box() {
        for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::max(); hi[k] = std::numeric_limits<float>::lowest(); }
    }
    void grow(const Box& b) {
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], b.lo[k]); hi[k] = fmaxf(hi[k], b.hi[k]); }
    }
    void grow(const float p[3]) {
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); }
    }
    float area() const {  // bounds.rs:507-513
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return (dx * dy + dx * dz + dy * dz) * 2.0f;
    }
    int widest() const {  // bounds.rs:343-356
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx > dy && dx > dz) return 0;
        return dy > dz ? 1 : 2;
    }
};

struct Item {  // BVHPrimitiveInfo, bvh.rs:53-68
    uint32_t id;
    Box box;
    float c[3];
};

struct TreeNode {
    Box box;
    int32_t kid[2];   // first (left) / second (right); -1 for leaves
    uint32_t first, count;
    uint8_t axis;
};

constexpr int kBuckets = 12;  // bvh.rs:35

inline int bucket_index(const Box& cb, int dim, const Item& it) {
    // Bounds3::offset (bounds.rs:371-391) then `as usize` (saturating)
    float o = it.c[dim] - cb.lo[dim];
    if (cb.hi[dim] > cb.lo[dim]) o /= cb.hi[dim] - cb.lo[dim];
    float f = (float)kBuckets * o;
    int b = (f != f || f <= 0.0f) ? 0 : (f >= 2147483647.0f ? 2147483647 : (int)f);
    if (b == kBuckets) b = kBuckets - 1;
    return b;
}

// Two-ended partition with the element motion of Iterator::partition_in_place
// (first failing element swapped with the last passing one).
template <typename Pred>
size_t split_range(Item* a, size_t n, Pred keep) {
    size_t i = 0, j = n, kept = 0;
    while (true) {
        while (i < j && keep(a[i])) { ++i; ++kept; }
        if (i >= j) break;
        // a[i] fails; look for the last passing element in (i, j)
        size_t k = j;
        bool found = false;
        while (k > i + 1) {
            --k;
            if (keep(a[k])) { found = true; break; }
        }
        if (!found) break;
        std::swap(a[i], a[k]);
        ++kept; ++i; j = k;
    }
    return kept;
}

struct Builder {
    std::vector<Item> items;
    std::vector<TreeNode> tree;
    std::vector<uint32_t> order;
    size_t max_leaf;
    int method;

    void make_leaf(int ni, size_t b, size_t e, const Box& box) {
        TreeNode& t = tree[ni];
        t.box = box; t.kid[0] = t.kid[1] = -1;
        t.first = (uint32_t)order.size(); t.count = (uint32_t)(e - b);
        for (size_t i = b; i < e; ++i) order.push_back(items[i].id);
    }

    // returns true if the range must become a leaf, else sets mid
    bool sah_split(const Box& box, const Box& cb, int dim, size_t b, size_t e, size_t* mid) {
        size_t n = e - b;
        if (n <= 2) {  // bvh.rs:305-311
            *mid = (b + e) / 2;
            if (b != e - 1 && items[e - 1].c[dim] < items[b].c[dim]) std::swap(items[b], items[e - 1]);
            return false;
        }
        size_t cnt[kBuckets] = {0};
        Box bb[kBuckets];
        for (size_t i = b; i < e; ++i) {
            int k = bucket_index(cb, dim, items[i]);
            cnt[k]++; bb[k].grow(items[i].box);
        }
        // prefix/suffix unions give the same boxes as the reference's O(B^2) loops
        // (fmin/fmax are associative and commutative on non-NaN input).
        Box pre[kBuckets], suf[kBuckets];
        size_t pc[kBuckets], sc[kBuckets];
        Box acc; size_t c = 0;
        for (int k = 0; k < kBuckets; ++k) { acc.grow(bb[k]); c += cnt[k]; pre[k] = acc; pc[k] = c; }
        acc = Box(); c = 0;
        for (int k = kBuckets - 1; k >= 0; --k) { acc.grow(bb[k]); c += cnt[k]; suf[k] = acc; sc[k] = c; }
        float best = 0.0f; int best_k = 0;
        float total = box.area();
        for (int k = 0; k < kBuckets - 1; ++k) {
            float cost = 1.0f + ((float)pc[k] * pre[k].area() + (float)sc[k + 1] * suf[k + 1].area()) / total;
            if (k == 0 || cost < best) { best = cost; best_k = k; }
        }
        if (n > max_leaf || best < (float)n) {
            *mid = b + split_range(&items[b], n, [&](const Item& it) { return bucket_index(cb, dim, it) <= best_k; });
            return false;
        }
        return true;
    }

    int grow() {
        struct Job { int node; size_t b, e; };
        std::vector<Job> todo;
        tree.reserve(2 * items.size());
        order.reserve(items.size());
        tree.push_back(TreeNode());
        todo.push_back({0, 0, items.size()});
        while (!todo.empty()) {
            Job j = todo.back(); todo.pop_back();
            Box box;
            for (size_t i = j.b; i < j.e; ++i) box.grow(items[i].box);
            size_t n = j.e - j.b;
            if (n == 1) { make_leaf(j.node, j.b, j.e, box); continue; }
            Box cb;
            for (size_t i = j.b; i < j.e; ++i) cb.grow(items[i].c);
            int dim = cb.widest();
            if (cb.hi[dim] == cb.lo[dim]) { make_leaf(j.node, j.b, j.e, box); continue; }
            size_t mid = (j.b + j.e) / 2;
            bool equal_counts = (method == PBRT_B200_SPLIT_EQUAL);
            if (method == PBRT_B200_SPLIT_MIDDLE) {  // bvh.rs:285-289
                float pm = (cb.lo[dim] + cb.hi[dim]) / 2.0f;
                mid = j.b + split_range(&items[j.b], n, [&](const Item& it) { return it.c[dim] < pm; });
                if (mid == j.b || mid == j.e) equal_counts = true;
            }
            if (equal_counts) {  // bvh.rs:291-299
                mid = (j.b + j.e) / 2;
                std::nth_element(items.begin() + j.b, items.begin() + mid, items.begin() + j.e,
                                 [dim](const Item& a, const Item& b) { return a.c[dim] < b.c[dim]; });
            } else if (sah_split(box, cb, dim, j.b, j.e, &mid)) {  // also reached by a successful Middle (bvh.rs:253-270)
                make_leaf(j.node, j.b, j.e, box);
                continue;
            }
            int l = (int)tree.size(); tree.push_back(TreeNode());
            int r = (int)tree.size(); tree.push_back(TreeNode());
            TreeNode& t = tree[j.node];
            t.kid[0] = l; t.kid[1] = r; t.axis = (uint8_t)dim; t.count = 0; t.first = 0;
            // the reference finishes the whole right range before the left one
            todo.push_back({l, j.b, mid});
            todo.push_back({r, mid, j.e});
        }
        // interior boxes = union of children (bvh.rs:115-121), bottom-up: kids have larger indices
        for (int i = (int)tree.size() - 1; i >= 0; --i) {
            TreeNode& t = tree[i];
            if (t.kid[0] >= 0) { Box b = tree[t.kid[0]].box; b.grow(tree[t.kid[1]].box); t.box = b; }
        }
        return 0;
    }

    void emit(pbrt_b200_bvh_node* out) {
        // pre-order, first child adjacent (bvh.rs:662-693)
        struct Slot { int node; int parent_out; };
        std::vector<Slot> st;
        st.push_back({0, -1});
        uint32_t next = 0;
        while (!st.empty()) {
            Slot s = st.back(); st.pop_back();
            uint32_t me = next++;
            if (s.parent_out >= 0) out[s.parent_out].offset = me;  // we are a second child
            const TreeNode& t = tree[s.node];
            pbrt_b200_bvh_node n;
            std::memset(&n, 0, sizeof n);
            for (int k = 0; k < 3; ++k) { n.bounds[k] = t.box.lo[k]; n.bounds[3 + k] = t.box.hi[k]; }
            if (t.kid[0] < 0) { n.n_prims = (uint16_t)t.count; n.offset = t.first; }
            else {
                n.axis = t.axis;
                st.push_back({t.kid[1], (int)me});
                st.push_back({t.kid[0], -1});
            }
            out[me] = n;
        }
    }
};

}  // namespace

extern "C" int pbrt_b200_bvh_build(const float* prim_bounds, uint64_t n, int max_prims_in_node, int split_method,
                                   pbrt_b200_bvh_node* nodes_out, uint32_t* ordered_out, uint64_t* n_nodes_out) {
    if (!n_nodes_out || (n && (!prim_bounds || !nodes_out || !ordered_out)))
        return pbrt_b200::fail(PBRT_B200_ERR_INVALID, "pbrt_b200_bvh_build: null argument");
    if (split_method != PBRT_B200_SPLIT_SAH && split_method != PBRT_B200_SPLIT_MIDDLE && split_method != PBRT_B200_SPLIT_EQUAL)
        return pbrt_b200::fail(PBRT_B200_ERR_UNSUPPORTED, "pbrt_b200_bvh_build: split method not supported (sah, middle, equal)");
    if (n > 0xfffffff0ull) return pbrt_b200::fail(PBRT_B200_ERR_INVALID, "pbrt_b200_bvh_build: too many primitives");
    *n_nodes_out = 0;
    if (n == 0) return PBRT_B200_OK;
    Builder b;
    b.max_leaf = (size_t)std::min(255, std::max(0, max_prims_in_node));  // bvh.rs:154
    b.method = split_method;
    b.items.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        Item& it = b.items[i];
        it.id = (uint32_t)i;
        for (int k = 0; k < 3; ++k) {
            it.box.lo[k] = prim_bounds[6 * i + k]; it.box.hi[k] = prim_bounds[6 * i + 3 + k];
            it.c[k] = it.box.lo[k] * 0.5f + it.box.hi[k] * 0.5f;  // bvh.rs:66
        }
    }
    b.grow();
    b.emit(nodes_out);
    std::memcpy(ordered_out, b.order.data(), n * sizeof(uint32_t));
    *n_nodes_out = b.tree.size();
    return PBRT_B200_OK;
}
