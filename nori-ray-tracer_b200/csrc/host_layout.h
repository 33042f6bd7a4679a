// host_layout.h -- traversal layouts derived on the host from a reference-format tree (bvh.h:127-164) for the
// large-scene kernels of wave_extend.cu.  Built by nori_gpu_upload_scene; plain words, no CUDA types.
#pragma once
#include <cstdint>
#include <vector>

// deepest per-ray stack of the large-scene kernels' layouts (wave_extend.cu: LaneStack2)
#define NORI_STACK2_MAX 96

// Child-box pairs: one 64-byte record (16 words) per INNER node, in node order:
//   {L.bmin, refL}{L.bmax, refR}{R.bmin, 0}{R.bmax, 0}
//   child reference: bit 31 = leaf; leaf: size in bits 30..25, first primitive in bits 24..0;
//                    inner: record index in bits 30..2, split axis in bits 1..0
// `w` = the nodes as 8 words each (flag|size-or-axis, start-or-right, bmin[3], bmax[3]); the tree must have passed
// the checks of nori_gpu_upload_scene (children in range, a tree).  Returns false (out empty) when the root is a
// leaf, a leaf holds more than 63 primitives or the counts exceed the reference encoding; rootRef = the root's
// own reference.
bool noriBuildPairLayout(const uint32_t *w, uint32_t n_nodes, uint32_t n_indices, std::vector<uint32_t> &out, uint32_t &rootRef);

// records of the 4-wide layout numbered breadth-first from the root before the numbering turns depth-first: 2^19
// records = 64 MB, the part of the tree an L2 access-policy window can hold (nori_gpu.cu: option "l2_window")
#define NORI_WIDE_TOP_RECORDS 524288u

// 4-wide records: one 128-byte record (32 words) per group of merged nodes, record 0 = root, the top of the tree
// breadth-first, the rest depth-first:
//   slot k = words 8k..8k+7 = {bmin, ref}{bmax, rank}; unused slots hold the empty-leaf reference 0x80000000;
//   rank = position of the slot's subtree among the record's slots in the reference's depth-first (leaf) order
//   child reference: bit 31 = leaf (as above); inner: record index
// Starting from a binary inner node's two children, the inner slot with the largest box surface is replaced by its
// own children until four slots are filled; empty leaves (bvh.cpp:437) are dropped.  Same preconditions as the
// pair layout; also returns false when a ray's stack (3 entries per record level) could exceed maxStack.
bool noriBuildWideLayout(const uint32_t *w, uint32_t n_nodes, uint32_t n_indices, uint32_t maxStack, std::vector<uint32_t> &out);
