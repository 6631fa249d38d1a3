/*
 * map_layout.h -- the flattened device image of a BSP map (shared by host and device code).
 *
 * The reference walks AoS file lumps: node -> plane, leaf -> leafbrush -> brush -> texture,
 * brush -> brushside -> plane (reference bsp.cc:127-129,186-190,214-215), using 12 of 36,
 * 8 of 48 and 4 of 72 bytes of the records it touches.  The device image keeps only what
 * the trace reads and removes every second hop:
 *
 *   NodeRec  (32 B)  split plane embedded + both children; nodes in breadth-first order so
 *                    the top of the tree is one contiguous prefix (staged in shared memory).
 *                    child >= 0: node index; child == kEmptyLeaf: leaf without solid brushes;
 *                    otherwise ~child = index of the leaf's first BrushRef.
 *   BrushRef (16 B)  one per (leaf, solid brush): side range + brush index; the leaf's last
 *                    entry carries kLastRef, so no leaf record and no count are needed.
 *                    The content test (textures[brush.texture].content_flags & 1, bsp.cc:193)
 *                    is baked in: non-solid brushes are dropped at flatten time.
 *   SideRec  (16 B)  the side's plane embedded (n.xyz, d), indexed like the brushsides lump.
 *   side_plane (4 B) planes-lump index per side, read once per hit for the `plane` extension.
 *
 * One device allocation holds [nodes | refs | sides | side_plane | qnodes | start grid], each section
 * 128-byte aligned; the whole-map shared-memory variant copies the first three sections verbatim.
 */
#ifndef BSPGPU_MAP_LAYOUT_H
#define BSPGPU_MAP_LAYOUT_H

#include <stdint.h>

namespace bspk {

struct alignas( 16 ) NodeRec {
	float nx, ny, nz, d;
	int32_t child[ 2 ];      /* [0] front (dist >= 0 side), [1] back -- same slot order as the reference */
	int32_t orig_child[ 2 ]; /* the file's own child values (leaf = -(idx+1)), for position_to_leaf */
};

/* QNodeRec (32 B): the record the read-only-path walk fetches -- ONE scattered 256-bit gather per visit, and for
 * axis-aligned split planes (every node plane of a kd-style map; q3map prefers them too) that one gather carries
 * TWO tree levels: the node's plane distance, both children's plane distances and the four grandchildren.
 *   general  words 0-3 plane (n.xyz, d), 4-5 children, 6 unused, 7 = 0
 *   packed   words 0-2 d of the node / of child 0 / of child 1, words 3-6 g0..g3, word 7 = w:
 *              bit 0      = 1 (packed)
 *              bit 1,2    child 0 / child 1 is "expanded": an internal node with an axial plane whose own d is in
 *                         word 1 / 2 and whose children are g0,g1 / g2,g3.  A child that is not expanded keeps its
 *                         ordinary reference in g0 / g2.
 *              bits 3-4, 5-6, 7-8   axis (0 x, 1 y, 2 z) of the node's, child 0's, child 1's plane
 *              bits 9-31  node index of the first expanded child; the second expanded child is the next index
 *                         (children are numbered consecutively in breadth-first order)
 * Only planes whose normal is bitwise (+1,+0,+0) / (+0,+1,+0) / (+0,+0,+1) are packed; the kernel rebuilds exactly
 * that normal and runs the same arithmetic as for a general record, so packing never changes a result.  Every node
 * keeps its own record at its own index (pops and start-grid references land there). */
struct alignas( 32 ) QNodeRec {
	uint32_t w[ 8 ];
};
static const uint32_t kQPacked = 1u, kQExpand0 = 2u, kQExpand1 = 4u;
static const uint32_t kQAxisShift = 3, kQAxis0Shift = 5, kQAxis1Shift = 7, kQBaseShift = 9;

struct alignas( 16 ) BrushRef {
	uint32_t first_side;
	uint32_t num_sides;      /* bit 31 (kLastRef) = last solid brush of this leaf */
	int32_t brush;           /* brushes-lump index */
	uint32_t reserved;
};

struct alignas( 16 ) SideRec {
	float nx, ny, nz, d;
};

static const int32_t kEmptyLeaf = INT32_MIN;
static const uint32_t kLastRef = 0x80000000u;
static const uint32_t kSectionAlign = 128;

/* Start grid: a uniform grid over the world box; cell -> the deepest child reference such that every
 * ancestor's plane has the whole (padded) cell at least `delta` away on one side.  A segment with both
 * endpoints in one cell is a pass-through at all those ancestors in the reference (bsp.cc:226-237: the
 * crossing parameter is provably < 0 or > 1, so only the near child is visited, interval unchanged), so the
 * walk may start at that reference instead of node 0 and still be bit-identical.  See flatten.cc. */
struct StartGridDesc {
	int32_t dim[ 3 ];        /* 0 = no grid */
	float lo[ 3 ], hi[ 3 ];  /* world box; a point is in cell floor((p - lo) * inv) when that lies in [0, dim) (trace_core.h: grid_start_ref) */
	float inv[ 3 ];          /* cells per unit */
};

/* byte offsets of the sections inside the single device image */
struct MapImageLayout {
	uint32_t num_nodes, num_refs, num_sides, tree_depth;
	uint64_t off_nodes, off_refs, off_sides, off_side_plane, off_qnodes, off_grid, total_bytes;
	uint64_t staged_bytes;   /* prefix [nodes|refs|sides] the whole-map variant copies to shared memory */
	StartGridDesc grid;
};

} // namespace bspk

#endif
