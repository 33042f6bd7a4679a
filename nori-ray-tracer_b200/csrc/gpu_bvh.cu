// gpu_bvh.cu -- nori_gpu_build_bvh_device: a BVH builder that runs on the GPU.
//
// The reference builds its tree on the host with a binned SAH sweep (bvh.cpp:54-382; restated in
// host_bvh.cpp, which reproduces the reference's trees node for node).  For scenes of millions of
// primitives that build dominates everything outside the timed region, so this file offers a second
// builder: a linear BVH (30-bit Morton codes of the primitive centroids, one radix sort, the parallel
// radix-tree construction of Karras 2012, bottom-up box fitting) whose OUTPUT IS IN THE REFERENCE'S FORMAT --
// 32-byte nodes (bvh.h:127-164) in depth-first order with the left child right behind its parent, a
// primitive index array and the shape offset table -- so every consumer of a reference tree (the traversal
// kernels, the oracle, nori_gpu_trace) takes it unchanged.  It is an alternative, never a silent
// replacement: which primitive wins an exact tie depends on leaf order (traverse.cuh), and parity with the
// reference is defined on the reference's tree.
//
// Pipeline (all on the context's stream):
//   k_prim_bounds   primitive boxes + centroids (mesh.cpp:172-178, sphere.cpp:39-41), scene box by atomics
//   k_morton        64-bit keys = (30-bit Morton code << 32) | primitive id  (unique => no tie handling)
//   cub radix sort  keys
//   k_radix_tree    one thread per internal node: range, split, children, parents (Karras 2012, sec. 4)
//   k_fit           bottom-up from the primitives: boxes and output-node counts; a subtree of <= leaf_size
//                   primitives collapses into one leaf
//   k_dfs_index     depth-first index of every surviving node = sum over its ancestors of (1 + size of the
//                   left sibling subtree when it is a right child)
//   k_emit          nodes in the reference layout; split axis = the coordinate of the highest differing
//                   Morton bit, so that "left child = smaller centroid along the axis" holds as for the
//                   reference's builder (used by the near-child-first traversal order)
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdint.h>
#include <string>
#include <vector>
#include "nori_gpu.h"

namespace {

struct Box { float mn[3], mx[3]; };

__device__ __forceinline__ uint32_t orderedFloat(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float unorderedFloat(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

struct DevShape { int32_t type; uint32_t first, count; const float *V; const uint32_t *F; float c[3], r; };

__global__ void k_prim_bounds(const DevShape *shapes, uint32_t nShapes, uint32_t n, Box *boxes, float3 *centroids, uint32_t *sceneBox /*6 ordered uints*/) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    if (g < n) {
        uint32_t lo = 0, hi = nShapes;                       // last shape with first <= g (findShape, bvh.h:105-109)
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (shapes[mid].first <= g) lo = mid; else hi = mid; }
        const DevShape &s = shapes[lo];
        const uint32_t idx = g - s.first;
        if (s.type == NORI_SHAPE_MESH) {
            for (int k = 0; k < 3; ++k) {
                const float *p = &s.V[3 * (size_t) s.F[3 * (size_t) idx + k]];
                for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], p[a]); mx[a] = fmaxf(mx[a], p[a]); }
            }
        } else {
            for (int a = 0; a < 3; ++a) { mn[a] = s.c[a] - s.r; mx[a] = s.c[a] + s.r; }
        }
        Box b; for (int a = 0; a < 3; ++a) { b.mn[a] = mn[a]; b.mx[a] = mx[a]; }
        boxes[g] = b;
        centroids[g] = make_float3(0.5f * (mn[0] + mx[0]), 0.5f * (mn[1] + mx[1]), 0.5f * (mn[2] + mx[2]));
    }
    // scene box of the CENTROIDS' extent is what the Morton grid needs; the boxes' extent contains it
    for (int a = 0; a < 3; ++a) {
        float lo = mn[a], hi = mx[a];
        for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&sceneBox[a], orderedFloat(lo)); atomicMax(&sceneBox[3 + a], orderedFloat(hi)); }
    }
}

__device__ __forceinline__ uint32_t expandBits(uint32_t v) {        // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float3 *centroids, uint32_t n, const uint32_t *sceneBox, uint64_t *keys) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    float q[3]; const float c[3] = {centroids[g].x, centroids[g].y, centroids[g].z};
    for (int a = 0; a < 3; ++a) {
        const float lo = unorderedFloat(sceneBox[a]), hi = unorderedFloat(sceneBox[3 + a]);
        const float e = hi - lo;
        q[a] = e > 0.f ? fminf(fmaxf((c[a] - lo) / e * 1024.f, 0.f), 1023.f) : 0.f;
    }
    const uint32_t code = (expandBits((uint32_t) q[0]) << 2) | (expandBits((uint32_t) q[1]) << 1) | expandBits((uint32_t) q[2]);
    keys[g] = ((uint64_t) code << 32) | g;
}

// Radix tree over n sorted unique keys: internal nodes 0..n-2, leaves 0..n-1.  A child reference is
// (index << 1) | isLeaf.
struct Tree {
    uint32_t *left, *right, *parent /* of internal nodes */, *leafParent, *first, *last;
};

__device__ __forceinline__ int delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long) (keys[i] ^ keys[j]));
}

__global__ void k_radix_tree(const uint64_t *keys, int n, Tree t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1) if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int div = 2, step = (l + div - 1) / div; ; div <<= 1, step = (l + div - 1) / div) {
        if (delta(keys, n, i, i + (s + step) * d) > dnode) s += step;
        if (step <= 1) break;
    }
    const int split = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t lref = (split == lo) ? (((uint32_t) split << 1) | 1u) : ((uint32_t) split << 1);
    const uint32_t rref = (split + 1 == hi) ? (((uint32_t) (split + 1) << 1) | 1u) : ((uint32_t) (split + 1) << 1);
    t.left[i] = lref; t.right[i] = rref; t.first[i] = (uint32_t) lo; t.last[i] = (uint32_t) hi;
    if (lref & 1u) t.leafParent[split] = (uint32_t) i; else t.parent[split] = (uint32_t) i;
    if (rref & 1u) t.leafParent[split + 1] = (uint32_t) i; else t.parent[split + 1] = (uint32_t) i;
    if (i == 0) t.parent[0] = 0xffffffffu;
}

// bottom-up: boxes of internal nodes and the number of OUTPUT nodes of every subtree (a subtree with at most
// leafSize primitives becomes one leaf)
__global__ void k_fit(const uint64_t *keys, int n, Tree t, const Box *primBoxes, Box *nodeBoxes, uint32_t *cnt, uint32_t *visits, uint32_t leafSize) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t node = t.leafParent[k];
    while (true) {
        if (atomicAdd(&visits[node], 1u) == 0u) return;           // the second child to arrive continues
        __threadfence();
        Box b; uint32_t c = 1;
        const uint32_t refs[2] = {t.left[node], t.right[node]};
        for (int a = 0; a < 3; ++a) { b.mn[a] = 1e30f; b.mx[a] = -1e30f; }
        for (int ch = 0; ch < 2; ++ch) {
            const uint32_t r = refs[ch];
            const Box cb = (r & 1u) ? primBoxes[(uint32_t) keys[r >> 1]] : nodeBoxes[r >> 1];
            for (int a = 0; a < 3; ++a) { b.mn[a] = fminf(b.mn[a], cb.mn[a]); b.mx[a] = fmaxf(b.mx[a], cb.mx[a]); }
            c += (r & 1u) ? 1u : cnt[r >> 1];
        }
        nodeBoxes[node] = b;
        cnt[node] = (t.last[node] - t.first[node] + 1u <= leafSize) ? 1u : c;
        __threadfence();
        if (t.parent[node] == 0xffffffffu) return;
        node = t.parent[node];
    }
}

// A node survives in the output iff no proper ancestor collapsed into a leaf.  Its depth-first index is the
// sum, over the path from the root, of 1 per step plus the size of the left sibling subtree on right turns.
__device__ bool dfsIndex(const Tree &t, const uint32_t *cnt, uint32_t leafSize, uint32_t ref, uint32_t &out) {
    uint32_t idx = 0;
    uint32_t child = ref;
    uint32_t p = (ref & 1u) ? t.leafParent[ref >> 1] : t.parent[ref >> 1];
    while (p != 0xffffffffu) {
        if (t.last[p] - t.first[p] + 1u <= leafSize) return false;           // an ancestor is an output leaf
        idx += 1u;
        if (t.right[p] == child) { const uint32_t l = t.left[p]; idx += (l & 1u) ? 1u : cnt[l >> 1]; }
        child = p << 1; p = t.parent[p];
    }
    out = idx;
    return true;
}

__global__ void k_emit(const uint64_t *keys, int n, Tree t, const Box *primBoxes, const Box *nodeBoxes, const uint32_t *cnt, uint32_t leafSize,
                       nori_gpu_bvh_node *nodes, uint32_t *indices) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) indices[k] = (uint32_t) keys[k];
    // thread k handles primitive-leaf k and internal node k
    if (k < n) {
        uint32_t di;
        if (dfsIndex(t, cnt, leafSize, ((uint32_t) k << 1) | 1u, di)) {
            nori_gpu_bvh_node nd; const Box b = primBoxes[(uint32_t) keys[k]];
            nd.data[0] = 1u | (1u << 1); nd.data[1] = (uint32_t) k;
            for (int a = 0; a < 3; ++a) { nd.bmin[a] = b.mn[a]; nd.bmax[a] = b.mx[a]; }
            nodes[di] = nd;
        }
    }
    if (k < n - 1) {
        uint32_t di;
        if (dfsIndex(t, cnt, leafSize, (uint32_t) k << 1, di)) {
            nori_gpu_bvh_node nd; const Box b = nodeBoxes[k];
            for (int a = 0; a < 3; ++a) { nd.bmin[a] = b.mn[a]; nd.bmax[a] = b.mx[a]; }
            const uint32_t size = t.last[k] - t.first[k] + 1u;
            if (size <= leafSize) { nd.data[0] = 1u | (size << 1); nd.data[1] = t.first[k]; }
            else {
                const uint32_t l = t.left[k];
                const uint32_t split = (l & 1u) ? (l >> 1) : t.last[l >> 1];            // last key of the left child
                const uint64_t x = keys[split] ^ keys[split + 1];
                const int hb = 63 - __clzll((long long) x);                              // highest differing bit
                const uint32_t axis = hb >= 32 ? (uint32_t) (2 - ((hb - 32) % 3)) : 0u;     // Morton bit 3k+2 = x, 3k+1 = y, 3k = z
                nd.data[0] = axis << 1;
                nd.data[1] = di + 1u + ((l & 1u) ? 1u : cnt[l >> 1]);                       // right child; the left one is di + 1
            }
            nodes[di] = nd;
        }
    }
}

struct Buf {                        // RAII device buffer
    void *p = nullptr;
    ~Buf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T *as() { return (T *) p; }
};

} // namespace

extern "C" int nori_gpu_build_bvh_device(int device, const nori_gpu_shape *shapes, uint32_t n_shapes, nori_gpu_bvh_node *nodes_out,
                                         uint32_t *indices_out, uint32_t *shape_offset_out, uint32_t *n_nodes_out, uint32_t leaf_size,
                                         float *build_ms_out) {
    if (!shapes || !nodes_out || !indices_out || !shape_offset_out || !n_nodes_out) return 1;
    if (leaf_size < 1) leaf_size = 1;
    if (leaf_size > 63) leaf_size = 63;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
#define GK(call) do { if ((call) != cudaSuccess) return 1; } while (0)
    uint32_t total = 0;
    std::vector<DevShape> hs(n_shapes);
    std::vector<Buf> geo(2 * (size_t) n_shapes);
    for (uint32_t i = 0; i < n_shapes; ++i) {
        shape_offset_out[i] = total;
        DevShape &d = hs[i]; d.type = shapes[i].type; d.first = total; d.V = nullptr; d.F = nullptr;
        d.count = shapes[i].type == NORI_SHAPE_MESH ? shapes[i].n_triangles : 1u;
        for (int a = 0; a < 3; ++a) d.c[a] = shapes[i].center[a];
        d.r = shapes[i].radius;
        if (shapes[i].type == NORI_SHAPE_MESH) {
            if (!shapes[i].V || !shapes[i].F) return 1;
            GK(geo[2 * i].alloc(12 * (size_t) shapes[i].n_vertices)); GK(geo[2 * i + 1].alloc(12 * (size_t) shapes[i].n_triangles));
            GK(cudaMemcpy(geo[2 * i].p, shapes[i].V, 12 * (size_t) shapes[i].n_vertices, cudaMemcpyHostToDevice));
            GK(cudaMemcpy(geo[2 * i + 1].p, shapes[i].F, 12 * (size_t) shapes[i].n_triangles, cudaMemcpyHostToDevice));
            d.V = geo[2 * i].as<float>(); d.F = geo[2 * i + 1].as<uint32_t>();
        }
        total += d.count;
    }
    shape_offset_out[n_shapes] = total;
    *n_nodes_out = 0;
    if (build_ms_out) *build_ms_out = 0.f;
    if (total == 0) return 0;
    const int n = (int) total;
    Buf dShapes, boxes, cents, sceneBox, keys, keysSorted, tmp, left, right, parent, leafParent, first, last, nodeBoxes, cnt, visits, dNodes, dIdx;
    GK(dShapes.alloc(sizeof(DevShape) * n_shapes)); GK(cudaMemcpy(dShapes.p, hs.data(), sizeof(DevShape) * n_shapes, cudaMemcpyHostToDevice));
    GK(boxes.alloc(sizeof(Box) * (size_t) n)); GK(cents.alloc(sizeof(float3) * (size_t) n)); GK(sceneBox.alloc(24));
    GK(keys.alloc(8 * (size_t) n)); GK(keysSorted.alloc(8 * (size_t) n));
    GK(left.alloc(4 * (size_t) n)); GK(right.alloc(4 * (size_t) n)); GK(parent.alloc(4 * (size_t) n)); GK(leafParent.alloc(4 * (size_t) n));
    GK(first.alloc(4 * (size_t) n)); GK(last.alloc(4 * (size_t) n)); GK(nodeBoxes.alloc(sizeof(Box) * (size_t) n));
    GK(cnt.alloc(4 * (size_t) n)); GK(visits.alloc(4 * (size_t) n)); GK(dNodes.alloc(sizeof(nori_gpu_bvh_node) * 2 * (size_t) n)); GK(dIdx.alloc(4 * (size_t) n));
    size_t tmpBytes = 0;
    GK(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keys.as<uint64_t>(), keysSorted.as<uint64_t>(), n));
    GK(tmp.alloc(tmpBytes));
    cudaEvent_t e0, e1; GK(cudaEventCreate(&e0)); GK(cudaEventCreate(&e1));
    const uint32_t initBox[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    GK(cudaMemcpy(sceneBox.p, initBox, 24, cudaMemcpyHostToDevice));
    GK(cudaMemset(visits.p, 0, 4 * (size_t) n));
    const int B = 256, G = (n + B - 1) / B;
    cudaEventRecord(e0);
    k_prim_bounds<<<G, B>>>(dShapes.as<DevShape>(), n_shapes, (uint32_t) n, boxes.as<Box>(), cents.as<float3>(), sceneBox.as<uint32_t>());
    k_morton<<<G, B>>>(cents.as<float3>(), (uint32_t) n, sceneBox.as<uint32_t>(), keys.as<uint64_t>());
    GK(cub::DeviceRadixSort::SortKeys(tmp.p, tmpBytes, keys.as<uint64_t>(), keysSorted.as<uint64_t>(), n));
    Tree t{left.as<uint32_t>(), right.as<uint32_t>(), parent.as<uint32_t>(), leafParent.as<uint32_t>(), first.as<uint32_t>(), last.as<uint32_t>()};
    uint32_t nNodes = 1;
    if (n == 1) {                                            // a single primitive: the root is its leaf
        const uint32_t none = 0xffffffffu;
        GK(cudaMemcpy(leafParent.p, &none, 4, cudaMemcpyHostToDevice));
        k_emit<<<1, 32>>>(keysSorted.as<uint64_t>(), n, t, boxes.as<Box>(), nodeBoxes.as<Box>(), cnt.as<uint32_t>(), leaf_size, dNodes.as<nori_gpu_bvh_node>(), dIdx.as<uint32_t>());
    } else {
        k_radix_tree<<<G, B>>>(keysSorted.as<uint64_t>(), n, t);
        k_fit<<<G, B>>>(keysSorted.as<uint64_t>(), n, t, boxes.as<Box>(), nodeBoxes.as<Box>(), cnt.as<uint32_t>(), visits.as<uint32_t>(), leaf_size);
        k_emit<<<G, B>>>(keysSorted.as<uint64_t>(), n, t, boxes.as<Box>(), nodeBoxes.as<Box>(), cnt.as<uint32_t>(), leaf_size, dNodes.as<nori_gpu_bvh_node>(), dIdx.as<uint32_t>());
        GK(cudaMemcpy(&nNodes, cnt.p, 4, cudaMemcpyDeviceToHost));          // output nodes of the root's subtree
    }
    cudaEventRecord(e1);
    GK(cudaEventSynchronize(e1));
    GK(cudaGetLastError());
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (build_ms_out) *build_ms_out = ms;
    GK(cudaMemcpy(nodes_out, dNodes.p, sizeof(nori_gpu_bvh_node) * (size_t) nNodes, cudaMemcpyDeviceToHost));
    GK(cudaMemcpy(indices_out, dIdx.p, 4 * (size_t) n, cudaMemcpyDeviceToHost));
    *n_nodes_out = nNodes;
#undef GK
    return 0;
}
