/*
 * map_layout.h -- the flattened device image of a BSP map (shared by host and device code).
 *
 * The reference walks AoS file lumps: node -> plane, leaf -> leafbrush -> brush -> texture,
 * brush -> brushside -> plane (reference bsp.cc:127-129,186-190,214-215), using 12 of 36,
 * 8 of 48 and 4 of 72 bytes of the records it touches.  The device image keeps only what
 * the trace reads and removes every second hop:
 *
 *   NodeRec  (16 B)  split-plane distance + both children + `w`.  Nodes are in breadth-first order.
 *                    child >= 0: node index; child == kEmptyLeaf: leaf without solid brushes;
 *                    otherwise ~child = index of the leaf's first BrushRef.
 *                    w = 0/1/2: the plane normal is bitwise +e_x / +e_y / +e_z (every node of a kd-style
 *                    map, most nodes of a q3map tree): the walk then needs no normal at all -- for a ray
 *                    whose six components are finite and whose start components are non-zero,
 *                    dot(start, e_a) and dot(e_a, dir) evaluated the reference's way are bitwise start.a
 *                    and dir.a (trace_core.h: ray_is_plain; other rays take the general arithmetic).
 *                    w >= kGeneralPlane: normal = gplanes[ w - kGeneralPlane ] (one more 16-byte fetch).
 *   PlaneRec (16 B)  (n.xyz, d): the general node planes, and the brush sides.
 *   BrushRef (16 B)  one per (leaf, solid brush): side range + brush index; the leaf's last
 *                    entry carries kLastRef, so no leaf record and no count are needed.
 *                    The content test (textures[brush.texture].content_flags & 1, bsp.cc:193)
 *                    is baked in: non-solid brushes are dropped at flatten time.
 *   sides            re-laid out per solid brush: even start, even count (zero-normal, d = +inf padding
 *                    sides are a no-op for every input), so a leaf step clips TWO sides per fetch.
 *                    Stored twice: `side_pairs` (sides 2k, 2k+1 adjacent: ONE 256-bit gather on the
 *                    read-only path) and `sides_even | sides_odd` (two dense 16-byte arrays: when the map
 *                    is staged in shared memory a dense array spreads eight lanes over eight bank groups,
 *                    the 32-byte-stride pair layout only over four -- ncu round 1: 57 % of the shared-memory
 *                    wavefronts of the staged kernel were bank conflicts).
 *   side_plane (4 B) planes-lump index per side, read once per hit for the `plane` extension.
 *   node_orig  (8 B) the file's own child values per node, read only by batched position_to_leaf.
 *   leaf_cluster (4 B) BSP_Leaf::cluster per leaves-lump index (bsp.h:68), for the PVS lookup.
 *
 * One device allocation; sections 128-byte aligned.  [nodes | gplanes | refs | sides_even | sides_odd] is a
 * contiguous prefix that the shared-memory variant copies verbatim (TMA bulk copy).
 */
#ifndef BSPGPU_MAP_LAYOUT_H
#define BSPGPU_MAP_LAYOUT_H

#include <stdint.h>

namespace bspk {

struct alignas( 16 ) NodeRec {
	float d;
	int32_t child[ 2 ];      /* [0] front (dist >= 0 side), [1] back -- same slot order as the reference */
	uint32_t w;              /* 0,1,2 = axis of a bitwise +unit normal; >= kGeneralPlane = gplanes index + kGeneralPlane */
};
static const uint32_t kGeneralPlane = 4;

struct alignas( 16 ) PlaneRec {
	float nx, ny, nz, d;
};
typedef PlaneRec SideRec;

struct NodeOrig {
	int32_t orig_child[ 2 ]; /* the file's own child values (leaf = -(idx+1)), for position_to_leaf */
};

struct alignas( 16 ) BrushRef {
	uint32_t first_side;     /* even */
	uint32_t num_sides;      /* even; bit 31 (kLastRef) = last solid brush of this leaf */
	int32_t brush;           /* brushes-lump index */
	uint32_t reserved;
};

/* Child slots: what the trace kernels walk.  Every child of every node -- internal node or leaf -- owns one 8-byte slot, and
 * the two children of a node sit side by side in one 16-byte pair, so ONE 16-byte gather fetches both children of the node
 * being tested while its plane test is still running (the plane of a node travels in the slot that points to it, not in a
 * record of its own): the dependent chain of a step is max(gather, plane test), not their sum.
 *   x & 7 = 0, 1, 2   internal node whose plane normal is bitwise +e_x / +e_y / +e_z: a = bits of the plane's d
 *   x & 7 = 4         internal node with any other plane: a = index into gplanes (which carries n and d)
 *                     in both cases x >> 3 = index of the pair that holds the node's children (breadth-first index + 1;
 *                     pair 0 = { root, unused })
 *   x == kSlotLeaf    leaf with solid brushes: a = index of its first BrushRef
 *   x == kSlotEmpty   leaf without solid brushes
 * kSlotDone / kSlotIdle never appear in the image: they are the two other states of a kernel lane. */
struct alignas( 8 ) Slot {
	uint32_t a, x;
};
static const uint32_t kSlotGeneral = 4, kSlotLeaf = 3, kSlotEmpty = 7, kSlotDone = 11, kSlotIdle = 15;

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
	uint32_t num_nodes, num_gplanes, num_refs, num_sides, tree_depth, num_leaves;
	uint64_t off_nodes, off_gplanes, off_refs, off_sides_even, off_sides_odd;   /* the staged prefix */
	uint64_t off_side_pairs, off_side_plane, off_node_orig, off_leaf_cluster, off_grid, total_bytes;
	uint64_t staged_bytes;   /* prefix the whole-map variant copies to shared memory */
	uint64_t off_slots, off_grid_slots;   /* child slots in pairs (2 * (num_nodes + 1) slots); start grid as slots (8 B per cell) */
	Slot root;               /* the slot of node 0: where every walk starts (bsp.cc:263) */
	StartGridDesc grid;
};

} // namespace bspk

#endif
