/*
 * bsp_host.h -- C++ host side of the drop-in: the reference's BSP load / query interface
 * (reference bsp.h:156-214) with the trace running on the GPU behind include/bspgpu.h.
 *
 * Kept from the reference, same names, argument meaning and ownership:
 *     void bsp_init( BSP *, std::string filename );              bsp.h:213 / bsp.cc:58
 *     void bsp_destroy( BSP * );                                  bsp.h:214 / bsp.cc:88
 *     bool BSP::trace_seg( start, end, Intersection & ) const;    bsp.h:210 / bsp.cc:255
 *     BSP_Leaf & BSP::position_to_leaf( pos ) const;              bsp.h:207 / bsp.cc:273
 *   and the public lump members (num_*, textures, planes, nodes, leaves, leaf_brushes,
 *   brushes, brush_sides) that point into one file blob owned by the BSP.
 * Added (north_star): the batched entry
 *     void BSP::trace_segs( const Segment *, size_t n, TraceResult * out ) const;
 *   plus reference-shaped and device-pointer overloads.
 * GPU state hides behind one opaque pointer so sizeof(BSP) grows by a single word
 * (the reference embeds BSP by value in GameState, game.h:95).
 *
 * Errors: the reference asserts (message + backtrace + trap, intrinsics.h:60-74).  Here a
 * failed load or a CUDA failure prints bspgpu_last_error() and aborts; there is no CPU
 * fallback for the trace.
 */
#ifndef BSP_HOST_H
#define BSP_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <string>

#include "vec3.h"
#include "../../../include/bspgpu.h"
#include "../../../include/bspgpu_multi.h"

/* ---- Quake-3 IBSP v46 records as they sit in the file (reference bsp.h:41-96) ---- */
struct BSP_Texture { char name[ 64 ]; int32_t surface_flags; int32_t content_flags; };
struct BSP_Plane { glm::vec3 n; float d; };
struct BSP_Node { uint32_t plane; int32_t children[ 2 ]; int32_t mins[ 3 ]; int32_t maxs[ 3 ]; };
struct BSP_Leaf {
	int32_t cluster, area;
	int32_t mins[ 3 ], maxs[ 3 ];
	uint32_t first_face, num_faces;
	uint32_t first_brush, num_brushes;
};
typedef uint32_t BSP_LeafFace;
typedef uint32_t BSP_LeafBrush;
struct BSP_Brush { uint32_t first_side, num_sides; int32_t texture; };
struct BSP_BrushSide { uint32_t plane; int32_t texture; };
/* render-side records (reference bsp.h:98-127): not read by the trace, kept so that code written against the
 * reference's BSP -- bsp_renderer.cc walks leaf_faces, faces, vertices, mesh_verts and vis -- compiles unchanged */
struct BSP_Vertex { glm::vec3 pos; float uv[ 2 ][ 2 ]; glm::vec3 normal; uint8_t rgba[ 4 ]; };
typedef int32_t BSP_MeshVert;
struct BSP_Face {
	int32_t texture, effect, type;
	int32_t first_vert, num_verts, first_mesh_vert, num_mesh_verts;
	int32_t lightmap, lmpos[ 2 ], lmsize[ 2 ];
	glm::vec3 lmorigin, a, b, normal;
	int32_t c, d;
};
struct BSP_Vis { uint32_t num_clusters, cluster_size; uint8_t pvs[ 1 ]; };   /* the reference declares a flexible array, bsp.h:129-134 */

static_assert( sizeof( BSP_Texture ) == 72 && sizeof( BSP_Plane ) == 16 && sizeof( BSP_Node ) == 36, "IBSP layout" );
static_assert( sizeof( BSP_Leaf ) == 48 && sizeof( BSP_Brush ) == 12 && sizeof( BSP_BrushSide ) == 8, "IBSP layout" );
static_assert( sizeof( BSP_Vertex ) == 44 && sizeof( BSP_Face ) == 104, "IBSP layout" );

/* result of one trace, reference bsp.h:143-147.  `plane` stays null exactly as in the
 * reference (bsp.cc:256,268 never assign it) unless BSP_HOST_FILL_PLANE is defined. */
struct Intersection {
	const BSP_Plane * plane;
	glm::vec3 pos;
	float t;
};

/* one segment = the (start, end) argument pair of trace_seg, 24 bytes */
struct Segment {
	glm::vec3 start;
	glm::vec3 end;
};

/* one batched result, 16 bytes, identical to bspgpu_result (see include/bspgpu.h for the fields) */
typedef bspgpu_result TraceResult;

class BSP {
public:
	char * contents;

	/* same members, same order as the reference (bsp.h:160-193) */
	uint32_t num_textures;      BSP_Texture * textures;
	uint32_t num_planes;        BSP_Plane * planes;
	uint32_t num_nodes;         BSP_Node * nodes;
	uint32_t num_leaves;        BSP_Leaf * leaves;
	uint32_t num_leaf_faces;    BSP_LeafFace * leaf_faces;
	uint32_t num_leaf_brushes;  BSP_LeafBrush * leaf_brushes;
	uint32_t num_brushes;       BSP_Brush * brushes;
	uint32_t num_brush_sides;   BSP_BrushSide * brush_sides;
	uint32_t num_vertices;      BSP_Vertex * vertices;
	uint32_t num_mesh_verts;    BSP_MeshVert * mesh_verts;
	uint32_t num_faces;         BSP_Face * faces;
	BSP_Vis * vis;

	bspgpu_t * gpu;             /* opaque: flattened map + streams on one device (null when the map lives on several) */
	bspgpu_multi_t * multi;     /* opaque: one replica per GPU of the box (bsp_init with device == BSP_ALL_GPUS) */

	/* reference API */
	bool trace_seg( const glm::vec3 & start, const glm::vec3 & end, Intersection & is ) const;
	BSP_Leaf & position_to_leaf( const glm::vec3 & pos ) const;

	/* batched API (host arrays; synchronous) */
	void trace_segs( const Segment * segs, size_t n, TraceResult * out ) const;
	/* same, shaped like n calls of trace_seg: hit[i] = return value, is[i] = Intersection */
	void trace_segs( const Segment * segs, size_t n, Intersection * is, bool * hit ) const;
	/* device-resident variant: 32-byte bspgpu_segment in, TraceResult out, both in HBM, enqueued
	 * on `cuda_stream` (a cudaStream_t, null = legacy default stream) without synchronising */
	void trace_segs_device( const bspgpu_segment * d_segs, size_t n, TraceResult * d_out, void * cuda_stream ) const;
	/* batched position_to_leaf: leaf index per point */
	void positions_to_leaves( const glm::vec3 * pos, size_t n, int32_t * leaf ) const;
	/* which leaves bspr_render would draw from `pos` (bsp_renderer.cc:147-168): visible[ num_leaves ] */
	void visible_leaves( const glm::vec3 & pos, uint8_t * visible ) const;
	/* hits of the last trace_segs call */
	uint64_t last_hit_count() const;
};

void bsp_init( BSP * bsp, std::string filename );
/* device >= 0: that GPU; BSP_CURRENT_GPU: the calling thread's current CUDA device; BSP_ALL_GPUS: the map is replicated on every
 * GPU of the box and trace_segs shards its segments across them (contiguous ranges, results in place) */
enum { BSP_CURRENT_GPU = -1, BSP_ALL_GPUS = -2 };
void bsp_init( BSP * bsp, std::string filename, int device );
void bsp_destroy( BSP * bsp );

#endif
