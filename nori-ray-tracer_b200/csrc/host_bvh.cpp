// host_bvh.cpp -- host-side SAH BVH construction producing the tree the reference would build
// (src/bvh.cpp:54-382), so that scenes which never pass through the reference's loader (synthetic
// benchmark scenes, host-authored scenes) are traversed with the same node layout and the same
// left-first order as reference-exported ones.
//
// It restates BVHBuildTask::execute (16 centroid bins along the largest axis, sweep from the right,
// cost = 2*T + (1/SA(node)) * (nL*SA(L) + nR*SA(R)), accept if < N; bvh.cpp:100-233) and
// execute_serially (< 32 primitives or no binned split: per axis std::sort by centroid + prefix-area
// sweep, node bbox recomputed when axis == 0, leaf if nothing beats cost N; bvh.cpp:236-305), then the
// compaction of the 2N node array (bvh.cpp:356-381).  The reference's only non-determinism is the
// order in which its parallel partition places 1000-primitive chunks (bvh.cpp:188-213); this builder
// takes the chunks in order (a stable partition), i.e. it reproduces the schedule of a
// single-threaded run.  Subtrees are independent memory regions (left child at i+1, right child at
// i+2*nL), so large subtrees are built on separate threads without changing the result.
// Float arithmetic follows the reference build (no FMA contraction: plain x86-64 SSE2).
#include "nori_gpu.h"
#include "host_layout.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

struct Box {
    float mn[3], mx[3];
    Box() { reset(); }
    void reset() { for (int i = 0; i < 3; ++i) { mn[i] = std::numeric_limits<float>::infinity(); mx[i] = -std::numeric_limits<float>::infinity(); } }
    void expand(const Box &b) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); } }
    void expand(const float *p) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], p[i]); mx[i] = std::max(mx[i], p[i]); } }
    float area() const {                                    // bbox.h:87-100
        float d0 = mx[0] - mn[0], d1 = mx[1] - mn[1], d2 = mx[2] - mn[2];
        float r = 0.0f; r += 1.0f * d1 * d2; r += 1.0f * d0 * d2; r += 1.0f * d0 * d1;
        return 2.0f * r;
    }
    int largestAxis() const {                               // bbox.h:308-317
        float e0 = mx[0] - mn[0], e1 = mx[1] - mn[1], e2 = mx[2] - mn[2];
        if (e0 >= e1 && e0 >= e2) return 0;
        if (e1 >= e0 && e1 >= e2) return 1;
        return 2;
    }
};

struct Builder {
    std::vector<Box> pbox;              // per-primitive bounding boxes (mesh.cpp:172-177, sphere.cpp:38)
    std::vector<float> cen;             // per-primitive centroids, 3 floats (mesh.cpp:179-184, sphere.cpp:40)
    std::vector<nori_gpu_bvh_node> nodes;   // 2N, zero-initialised (bvh.cpp:340-341)
    uint32_t *indices = nullptr;
    std::atomic<int> threadsLeft{0};

    static void setBox(nori_gpu_bvh_node &n, const Box &b) { memcpy(n.bmin, b.mn, 12); memcpy(n.bmax, b.mx, 12); }
    static Box getBox(const nori_gpu_bvh_node &n) { Box b; memcpy(b.mn, n.bmin, 12); memcpy(b.mx, n.bmax, 12); return b; }

    // bvh.cpp:236-305
    void serial(uint32_t node_idx, uint32_t *start, uint32_t *end, uint32_t *temp) {
        nori_gpu_bvh_node &node = nodes[node_idx];
        const uint32_t size = (uint32_t) (end - start);
        float best_cost = (float) 1 * size;
        int64_t best_index = -1, best_axis = -1;
        float *left_areas = (float *) temp;
        Box nodeBox = getBox(node);
        for (int axis = 0; axis < 3; ++axis) {
            std::sort(start, end, [&](uint32_t f1, uint32_t f2) { return cen[3 * (size_t) f1 + axis] < cen[3 * (size_t) f2 + axis]; });
            Box bbox;
            for (uint32_t i = 0; i < size; ++i) { bbox.expand(pbox[start[i]]); left_areas[i] = bbox.area(); }
            if (axis == 0) { nodeBox = bbox; setBox(node, bbox); }
            bbox.reset();
            float tri_factor = 1 / nodeBox.area();
            for (uint32_t i = size - 1; i >= 1; --i) {
                bbox.expand(pbox[start[i]]);
                float left_area = left_areas[i - 1], right_area = bbox.area();
                uint32_t prims_left = i, prims_right = size - i;
                float sah_cost = 2.0f * 1 + tri_factor * (prims_left * left_area + prims_right * right_area);
                if (sah_cost < best_cost) { best_cost = sah_cost; best_index = i; best_axis = axis; }
            }
        }
        if (best_index == -1) {                              // leaf (any size)
            node.data[0] = 1u | (size << 1);
            node.data[1] = (uint32_t) (start - indices);
            return;
        }
        std::sort(start, end, [&](uint32_t f1, uint32_t f2) { return cen[3 * (size_t) f1 + best_axis] < cen[3 * (size_t) f2 + best_axis]; });
        const uint32_t left_count = (uint32_t) best_index;
        const uint32_t left = node_idx + 1, right = node_idx + 2 * left_count;
        node.data[0] = ((uint32_t) best_axis) << 1;          // flag 0 | axis
        node.data[1] = right;
        serial(left, start, start + left_count, temp);
        serial(right, start + left_count, end, temp + left_count);
    }

    // bvh.cpp:100-233
    void build(uint32_t node_idx, uint32_t *start, uint32_t *end, uint32_t *temp) {
        while (true) {
            const uint32_t size = (uint32_t) (end - start);
            nori_gpu_bvh_node &node = nodes[node_idx];
            if (size < 32) { serial(node_idx, start, end, temp); return; }
            const Box nb = getBox(node);
            const int axis = nb.largestAxis();
            const float mn = nb.mn[axis], mx = nb.mx[axis], inv_bin_size = 16 / (mx - mn);
            uint32_t counts[16]; Box bins[16];
            memset(counts, 0, sizeof(counts));
            for (uint32_t i = 0; i < size; ++i) {
                const uint32_t f = start[i];
                const float c = cen[3 * (size_t) f + axis];
                int index = std::min(std::max((int) ((c - mn) * inv_bin_size), 0), 15);
                counts[index]++; bins[index].expand(pbox[f]);
            }
            Box bbox_left[16]; bbox_left[0] = bins[0];
            for (int i = 1; i < 16; ++i) { counts[i] += counts[i - 1]; bbox_left[i] = bbox_left[i - 1]; bbox_left[i].expand(bins[i]); }
            Box bbox_right = bins[15], best_bbox_right;
            int64_t best_index = -1;
            float best_cost = (float) 1 * size;
            const float tri_factor = (float) 1 / nb.area();
            for (int i = 14; i >= 0; --i) {
                uint32_t prims_left = counts[i], prims_right = size - counts[i];
                float sah_cost = 2.0f * 1 + tri_factor * (prims_left * bbox_left[i].area() + prims_right * bbox_right.area());
                if (sah_cost < best_cost) { best_cost = sah_cost; best_index = i; best_bbox_right = bbox_right; }
                bbox_right.expand(bins[i]);
            }
            if (best_index == -1) { serial(node_idx, start, end, temp); return; }
            const uint32_t left_count = counts[best_index];
            const uint32_t left = node_idx + 1, right = node_idx + 2 * left_count;
            setBox(nodes[left], bbox_left[best_index]);
            setBox(nodes[right], best_bbox_right);
            node.data[0] = ((uint32_t) axis) << 1;
            node.data[1] = right;
            uint32_t il = 0, ir = left_count;                // stable partition (chunks taken in order)
            for (uint32_t i = 0; i < size; ++i) {
                const uint32_t f = start[i];
                const int index = (int) ((cen[3 * (size_t) f + axis] - mn) * inv_bin_size);
                if (index <= best_index) temp[il++] = f; else temp[ir++] = f;
            }
            memcpy(start, temp, size * sizeof(uint32_t));
            // right subtree: own thread when it is large and a thread is available; else inline
            uint32_t *rs = start + left_count, *re = end, *rt = temp + left_count;
            std::thread worker;
            bool spawned = false;
            if (size - left_count > 65536 && threadsLeft.fetch_sub(1) > 0) {
                worker = std::thread([this, right, rs, re, rt]() { build(right, rs, re, rt); threadsLeft.fetch_add(1); });
                spawned = true;
            } else if (size - left_count > 65536) threadsLeft.fetch_add(1);
            if (!spawned) build(right, rs, re, rt);
            // left subtree: continue in this frame (the reference recycles the task, bvh.cpp:228-232)
            build(left, start, start + left_count, temp);
            if (spawned) worker.join();
            return;
        }
    }
};

} // namespace

// ---------------------------------------------------------------------------------------------
// traversal layouts for the large-scene kernels (host_layout.h)
// ---------------------------------------------------------------------------------------------
namespace {
struct NodeWords {
    const uint32_t *w;
    bool leaf(uint32_t i) const { return (w[8 * (size_t) i] & 1u) != 0; }
    uint32_t size(uint32_t i) const { return w[8 * (size_t) i] >> 1; }            // leaf: primitive count; inner: split axis
    uint32_t second(uint32_t i) const { return w[8 * (size_t) i + 1]; }           // leaf: first primitive; inner: right child
    bool empty(uint32_t i) const { return leaf(i) && size(i) == 0u; }
    const uint32_t *box(uint32_t i) const { return &w[8 * (size_t) i + 2]; }
    uint32_t leafRef(uint32_t i) const { return 0x80000000u | (size(i) << 25) | second(i); }
    double area(uint32_t i) const {
        const float *b = (const float *) box(i);
        const double dx = (double) b[3] - b[0], dy = (double) b[4] - b[1], dz = (double) b[5] - b[2];
        return dx * dy + dy * dz + dz * dx;
    }
};
bool layoutEncodable(const NodeWords &t, uint32_t n_nodes, uint32_t n_indices) {
    if (n_nodes == 0 || t.leaf(0) || n_indices >= (1u << 25) || n_nodes >= (1u << 29)) return false;
    for (uint32_t i = 0; i < n_nodes; ++i) if (t.leaf(i) && t.size(i) > 63u) return false;
    return true;
}
} // namespace

bool noriBuildPairLayout(const uint32_t *w, uint32_t n_nodes, uint32_t n_indices, std::vector<uint32_t> &out, uint32_t &rootRef) {
    out.clear();
    const NodeWords t{w};
    if (!layoutEncodable(t, n_nodes, n_indices)) return false;
    std::vector<uint32_t> innerIdx(n_nodes, 0);
    uint32_t n = 0;
    for (uint32_t i = 0; i < n_nodes; ++i) if (!t.leaf(i)) innerIdx[i] = n++;
    auto ref = [&](uint32_t c) { return t.leaf(c) ? t.leafRef(c) : ((innerIdx[c] << 2) | (t.size(c) & 3u)); };
    out.assign(16 * (size_t) n, 0u);
    for (uint32_t i = 0; i < n_nodes; ++i) {
        if (t.leaf(i)) continue;
        const uint32_t l = i + 1, r = t.second(i);
        if (l >= n_nodes || r >= n_nodes) { out.clear(); return false; }
        const uint32_t *a = t.box(l), *b = t.box(r);
        uint32_t *o = &out[16 * (size_t) innerIdx[i]];
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = ref(l);
        o[4] = a[3]; o[5] = a[4]; o[6] = a[5]; o[7] = ref(r);
        o[8] = b[0]; o[9] = b[1]; o[10] = b[2];
        o[12] = b[3]; o[13] = b[4]; o[14] = b[5];
    }
    rootRef = t.size(0) & 3u;
    return true;
}

bool noriBuildWideLayout(const uint32_t *w, uint32_t n_nodes, uint32_t n_indices, uint32_t maxStack, std::vector<uint32_t> &out) {
    out.clear();
    const NodeWords t{w};
    if (!layoutEncodable(t, n_nodes, n_indices)) return false;
    std::vector<uint32_t> slots;                                 // 4 per record: node index or 0xffffffff
    // Record numbering: the top of the tree breadth-first (the first NORI_WIDE_TOP_RECORDS records: the records
    // most rays pass through are contiguous, which is what an L2 access-policy window can pin), everything below
    // depth-first (a subtree's records stay together).  The numbering is only a naming: references carry it.
    std::vector<std::pair<uint32_t, uint32_t>> st; st.reserve(256);   // (binary node that roots a record, record depth)
    std::vector<uint32_t> recOf(n_nodes, 0xffffffffu);
    uint32_t n = 0, maxDepth = 0;
    size_t head = 0;                                             // breadth-first phase: st is a queue read at `head`
    bool bfs = true;
    st.push_back({0u, 1u});
    while (bfs ? head < st.size() : !st.empty()) {
        uint32_t i, depth;
        if (bfs) {
            i = st[head].first; depth = st[head].second; ++head;
            if (n + 1 >= NORI_WIDE_TOP_RECORDS) {                // switch: the rest of the queue becomes the depth-first stack
                st.erase(st.begin(), st.begin() + head); std::reverse(st.begin(), st.end());
                bfs = false; head = 0;
            }
        } else { i = st.back().first; depth = st.back().second; st.pop_back(); }
        if (n >= n_nodes || i + 1 >= n_nodes || t.second(i) >= n_nodes) return false;     // not a tree
        recOf[i] = n++; maxDepth = std::max(maxDepth, depth);
        uint32_t sl[4]; int cnt = 0;
        for (uint32_t c : { i + 1, t.second(i) }) if (!t.empty(c)) sl[cnt++] = c;
        while (cnt < 4) {
            int best = -1; double bestA = -1.0;
            for (int k = 0; k < cnt; ++k) if (!t.leaf(sl[k]) && t.area(sl[k]) > bestA) { bestA = t.area(sl[k]); best = k; }
            if (best < 0) break;
            const uint32_t c = sl[best];
            if (c + 1 >= n_nodes || t.second(c) >= n_nodes) return false;
            sl[best] = sl[--cnt];
            for (uint32_t gc : { c + 1, t.second(c) }) if (!t.empty(gc)) sl[cnt++] = gc;
        }
        for (int k = 0; k < 4; ++k) slots.push_back(k < cnt ? sl[k] : 0xffffffffu);
        if (bfs) { for (int k = 0; k < cnt; ++k) if (!t.leaf(sl[k])) st.push_back({sl[k], depth + 1}); }
        else for (int k = cnt - 1; k >= 0; --k) if (!t.leaf(sl[k])) st.push_back({sl[k], depth + 1});
    }
    if (3u * maxDepth > maxStack) return false;
    out.assign(32 * (size_t) n, 0u);
    for (uint32_t r = 0; r < n; ++r)
        for (int k = 0; k < 4; ++k) {
            uint32_t *o = &out[32 * (size_t) r + 8 * k];
            const uint32_t j = slots[4 * (size_t) r + k];
            if (j == 0xffffffffu) { o[3] = 0x80000000u; continue; }
            const uint32_t *b = t.box(j);
            o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = t.leaf(j) ? t.leafRef(j) : recOf[j];
            o[4] = b[3]; o[5] = b[4]; o[6] = b[5];
            // rank of this slot in the reference's depth-first order (node indices ARE that order: a left subtree precedes
            // its sibling): visiting the hit slots by rank walks the leaves exactly as bvh.cpp:430-433 does
            for (int q = 0; q < 4; ++q) { const uint32_t jq = slots[4 * (size_t) r + q]; if (jq != 0xffffffffu && jq < j) ++o[7]; }
        }
    return true;
}

extern "C" {

// Build the reference's SAH BVH over `shapes` (bvh.cpp:329-382).  Outputs: nodes_out (capacity
// 2 * total primitives; compacted, *n_nodes_out entries are valid), indices_out (total primitives),
// shape_offset_out (n_shapes + 1).  threads <= 0: hardware concurrency.  Returns 0 on success.
int nori_gpu_build_bvh(const nori_gpu_shape *shapes, uint32_t n_shapes, nori_gpu_bvh_node *nodes_out,
                       uint32_t *indices_out, uint32_t *shape_offset_out, uint32_t *n_nodes_out, int threads) {
    if (!shapes || !nodes_out || !indices_out || !shape_offset_out || !n_nodes_out) return 1;
    Builder b;
    shape_offset_out[0] = 0;
    for (uint32_t s = 0; s < n_shapes; ++s) shape_offset_out[s + 1] = shape_offset_out[s] + shapes[s].n_triangles;   // bvh.cpp:308-312
    const uint32_t size = shape_offset_out[n_shapes];
    *n_nodes_out = 0;
    if (size == 0) return 0;
    b.pbox.resize(size); b.cen.resize(3 * (size_t) size);
    Box root;
    for (uint32_t s = 0; s < n_shapes; ++s) {
        const nori_gpu_shape &h = shapes[s];
        const uint32_t base = shape_offset_out[s];
        if (h.type == NORI_SHAPE_MESH) {
            if (!h.V || !h.F) return 1;
            for (uint32_t v = 0; v < h.n_vertices; ++v) root.expand(&h.V[3 * (size_t) v]);       // Mesh bbox over all vertices (obj.cpp)
            for (uint32_t t = 0; t < h.n_triangles; ++t) {
                const float *p0 = &h.V[3 * (size_t) h.F[3 * (size_t) t]], *p1 = &h.V[3 * (size_t) h.F[3 * (size_t) t + 1]], *p2 = &h.V[3 * (size_t) h.F[3 * (size_t) t + 2]];
                Box bb; bb.expand(p0); bb.expand(p1); bb.expand(p2);
                b.pbox[base + t] = bb;
                for (int k = 0; k < 3; ++k) b.cen[3 * (size_t) (base + t) + k] = (1.0f / 3.0f) * ((p0[k] + p1[k]) + p2[k]);   // mesh.cpp:179-184
            }
        } else {
            Box bb;
            float lo[3], hi[3];
            for (int k = 0; k < 3; ++k) { lo[k] = h.center[k] - h.radius; hi[k] = h.center[k] + h.radius; }
            bb.expand(lo); bb.expand(hi);                                                         // sphere.cpp:33-34
            b.pbox[base] = bb; root.expand(bb);
            for (int k = 0; k < 3; ++k) b.cen[3 * (size_t) base + k] = h.center[k];
        }
    }
    b.nodes.assign(2 * (size_t) size, nori_gpu_bvh_node{});
    Builder::setBox(b.nodes[0], root);
    for (uint32_t i = 0; i < size; ++i) indices_out[i] = i;
    b.indices = indices_out;
    std::vector<uint32_t> temp(size);
    int nt = threads > 0 ? threads : (int) std::thread::hardware_concurrency();
    b.threadsLeft = std::max(0, nt - 1);
    b.build(0, indices_out, indices_out + size, temp.data());
    // compaction of unused (all-zero) nodes, bvh.cpp:356-381: preorder, left child stays at i+1
    std::vector<uint32_t> remap(b.nodes.size());
    uint32_t used = 0;
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const nori_gpu_bvh_node &n = b.nodes[i];
        remap[i] = used;
        if (n.data[0] != 0 || n.data[1] != 0) ++used;
    }
    uint32_t j = 0;
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        nori_gpu_bvh_node n = b.nodes[i];
        if (n.data[0] == 0 && n.data[1] == 0) continue;
        if (!(n.data[0] & 1u)) n.data[1] = remap[n.data[1]];
        nodes_out[j++] = n;
    }
    *n_nodes_out = used;
    return 0;
}

// Mesh::activate (mesh.cpp:30-38) + DiscretePDF::append / normalize (dpdf.h:56-109):
// cdf_out has n_triangles + 1 entries; *normalization_out = 1 / total area.
int nori_gpu_mesh_area_cdf(const float *V, const uint32_t *F, uint32_t n_triangles, float *cdf_out, float *normalization_out) {
    if (!V || !F || !cdf_out || !normalization_out) return 1;
    cdf_out[0] = 0.0f;
    for (uint32_t t = 0; t < n_triangles; ++t) {
        const float *p0 = &V[3 * (size_t) F[3 * (size_t) t]], *p1 = &V[3 * (size_t) F[3 * (size_t) t + 1]], *p2 = &V[3 * (size_t) F[3 * (size_t) t + 2]];
        float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, c[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        float x = a[1] * c[2] - a[2] * c[1], y = a[2] * c[0] - a[0] * c[2], z = a[0] * c[1] - a[1] * c[0];
        float area = 0.5f * std::sqrt(x * x + (y * y + z * z));                                   // mesh.cpp:75-81
        cdf_out[t + 1] = cdf_out[t] + area;
    }
    float sum = cdf_out[n_triangles];
    if (sum > 0) {
        float norm = 1.0f / sum;
        for (uint32_t i = 1; i <= n_triangles; ++i) cdf_out[i] *= norm;
        cdf_out[n_triangles] = 1.0f;
        *normalization_out = norm;
    } else *normalization_out = 0.0f;
    return 0;
}

// Test hook for the 4-wide layout (host_layout.h).  records_out: 32 words per record, capacity in records
// (n_nodes / 2 + 1 always suffices); *n_records_out = 0 when the layout is not built for this tree.
int nori_gpu_wide_layout(const nori_gpu_bvh_node *nodes, uint32_t n_nodes, uint32_t n_indices, uint32_t *records_out,
                         uint32_t capacity, uint32_t *n_records_out) {
    if (!nodes || !records_out || !n_records_out) return 1;
    std::vector<uint32_t> out;
    *n_records_out = 0;
    if (!noriBuildWideLayout((const uint32_t *) nodes, n_nodes, n_indices, NORI_STACK2_MAX, out)) return 0;
    if (out.size() / 32 > capacity) return 1;
    memcpy(records_out, out.data(), out.size() * sizeof(uint32_t));
    *n_records_out = (uint32_t) (out.size() / 32);
    return 0;
}

} // extern "C"
