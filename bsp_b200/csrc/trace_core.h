/*
 * trace_core.h -- the per-segment arithmetic of the trace, written once.
 *
 * Compiled by nvcc into the sm_100a kernels (trace_kernels.cuh) and, for tests only,
 * by g++ into tests/hostsim (a CPU build of the SAME inline functions over the SAME
 * flattened image, compared bit-for-bit with the reference object before any GPU time
 * is spent).  The product never runs this on the CPU.
 *
 * Semantics are the reference's, statement for statement:
 *   descend()  = BSP::trace_seg_tree  (reference bsp.cc:205-253) with the recursion turned
 *                into a loop + an explicit stack of pending far children;
 *   leaf()     = BSP::trace_seg_leaf  (bsp.cc:185-203) + BSP::trace_seg_brush (bsp.cc:120-183).
 * No epsilon anywhere (bsp.cc:47's EPSILON is never read).  Every fp32 operation is a
 * single IEEE-754 rounding in the reference's order: build with -fmad=false -prec-div=true
 * -ftz=false (nvcc) / -ffp-contract=off (g++).  glm::dot(vec3) is (x*x' + y*y') + z*z'.
 *
 * Stack invariant (lets an entry be 8 bytes): the reference's far call gets [t, tmax] where
 * tmax is the interval end of the node that straddled (bsp.cc:250-252).  At any moment the
 * current interval end equals the start (`t`) of the entry on top of the stack, or 1.0f when
 * the stack is empty -- so `tmax` is not stored, it is read back from the entry below on pop.
 * For that to hold every straddle pushes, including far children that are empty leaves; those
 * are discarded at pop time.
 */
#ifndef BSPGPU_TRACE_CORE_H
#define BSPGPU_TRACE_CORE_H

#include "map_layout.h"

#ifdef __CUDACC__
#define BSPK_HD __host__ __device__ __forceinline__
#else
#define BSPK_HD inline
#endif

namespace bspk {

struct Ray {
	float sx, sy, sz;   /* start */
	float dx, dy, dz;   /* dir = end - start (bsp.cc:261) */
};

struct StackEntry {
	int32_t node;
	float t;
};

/* running result = the reference's BSP_Intersection {hit, is.t} (bsp.h:149-154) + extension fields */
struct HitState {
	float t;            /* bis.is.t, starts at 1.0f (bsp.cc:257) */
	int32_t hit;        /* bis.hit */
	int32_t far_side;   /* extension: brushsides index that last raised tfar in the brush that supplied t, or -1 */
	int32_t brush;      /* extension: brushes index that supplied t, or -1 */
	int32_t startsolid; /* extension */
};

BSPK_HD float dot3( float ax, float ay, float az, float bx, float by, float bz ) {
	const float px = ax * bx, py = ay * by, pz = az * bz;
	return ( px + py ) + pz;
}

/* One trace_seg_tree step at an internal node (bsp.cc:213-252).  Returns the near child and
 * narrows tmax; `far`/`both` tell the caller what to push.
 * `u >= 0 && u <= tmax` already excludes denom == 0 (u is then +-inf or NaN), which is the
 * reference's explicit `denom != 0.0f` guard (bsp.cc:226). */
BSPK_HD int32_t node_step( const Ray & r, float nx, float ny, float nz, float d, int32_t c0, int32_t c1,
                           float tmin, float & tmax, int32_t & far, bool & both ) {
	const float denom = dot3( nx, ny, nz, r.dx, r.dy, r.dz );                  /* :220 */
	const float dist = -( dot3( r.sx, r.sy, r.sz, nx, ny, nz ) - d );         /* :221 */
	bool near_child = dist > 0.0f;                                             /* :222 */
	both = false;
	const float u = dist / denom;                                              /* :227 */
	if( denom != 0.0f && u >= 0.0f && u <= tmax ) {                            /* :226,:232 */
		if( u < tmin ) near_child = !near_child;                               /* :235 */
		else { both = true; tmax = u; }                                        /* :238-241 */
	}
	far = near_child ? c0 : c1;                                                /* children[ !near ] */
	return near_child ? c1 : c0;                                               /* children[ near ] :248 */
}

/* trace_seg_tree from `node` down to a leaf reference; pending far sides go on the stack. */
template< class Map >
BSPK_HD int32_t descend( const Map & map, const Ray & r, int32_t node, float tmin, float & tmax, StackEntry * stack, int32_t & sp ) {
	while( node >= 0 ) {
		float nx, ny, nz, d; int32_t c0, c1;
		map.node( node, nx, ny, nz, d, c0, c1 );
		int32_t far; bool both;
		node = node_step( r, nx, ny, nz, d, c0, c1, tmin, tmax, far, both );
		if( both ) {
			stack[ sp ].node = far;
			stack[ sp ].t = tmax;   /* = u: the far interval is [u, old tmax] (bsp.cc:251) */
			sp++;
		}
	}
	return node;
}

/* One side of trace_seg_brush's loop (bsp.cc:128-165).  Returns false when the brush is
 * rejected (both points in front, :135). */
struct BrushClip {
	float tfar, tnear;
	int32_t far_side;
	bool hit, starts_inside;
};

BSPK_HD bool side_step( const Ray & r, float ex, float ey, float ez, float nx, float ny, float nz, float d,
                        float tmin, float tmax, int32_t side_index, BrushClip & b ) {
	const float start_dist = dot3( r.sx, r.sy, r.sz, nx, ny, nz ) - d;        /* :131 */
	const float end_dist = dot3( ex, ey, ez, nx, ny, nz ) - d;                /* :132 */
	if( start_dist > 0.0f && end_dist > 0.0f ) return false;                   /* :135 */
	if( start_dist <= 0.0f && end_dist <= 0.0f ) return true;                  /* :137 */
	if( start_dist >= 0.0f ) b.starts_inside = false;                          /* :139 */
	const float denom = -dot3( nx, ny, nz, r.dx, r.dy, r.dz );                /* :141 */
	const float t = start_dist / denom;                                        /* :142 */
	if( t >= tmin && t <= tmax ) {                                             /* :144 */
		if( start_dist >= 0.0f ) {                                             /* :147 */
			if( t >= b.tfar ) { b.tfar = t; b.hit = true; b.far_side = side_index; }
		}
		else if( t <= b.tnear ) { b.tnear = t; b.hit = true; }                 /* :154 */
	}
	return true;
}

/* The tail of trace_seg_brush (bsp.cc:169-182) folded into trace_seg_leaf's update (bsp.cc:197-200). */
BSPK_HD void brush_finish( const BrushClip & b, int32_t brush_index, HitState & h ) {
	if( b.starts_inside ) h.startsolid = 1;
	if( !b.hit ) return;
	float t; int32_t side;
	if( !b.starts_inside ) {
		if( !( b.tfar <= b.tnear ) ) return;                                   /* :171 */
		t = b.tfar; side = b.far_side;
	}
	else {
		if( !( b.tnear < b.tfar ) ) return;                                    /* :176 */
		t = b.tnear; side = -1;
	}
	if( !h.hit || t < h.t ) { h.far_side = side; h.brush = brush_index; }      /* extension bookkeeping */
	h.hit = 1;                                                                 /* :198 */
	if( t < h.t ) h.t = t;                                                     /* :199 */
}

/* trace_seg_leaf over the leaf's solid brushes; `leaf_ref` is a negative non-empty child value. */
template< class Map >
BSPK_HD void leaf( const Map & map, const Ray & r, int32_t leaf_ref, float tmin, float tmax, HitState & h ) {
	/* end = start + dir * tmax (bsp.cc:121): depends on the leaf interval only, same for every brush */
	const float ex = r.sx + r.dx * tmax, ey = r.sy + r.dy * tmax, ez = r.sz + r.dz * tmax;
	uint32_t ri = ( uint32_t ) ~leaf_ref;
	for( ;; ) {
		uint32_t first, nraw; int32_t brush;
		map.ref( ri, first, nraw, brush );
		const uint32_t ns = nraw & ~kLastRef;
		BrushClip b;
		b.tfar = tmin; b.tnear = tmax; b.far_side = -1; b.hit = false; b.starts_inside = true;  /* :122-125 */
		bool alive = true;
		for( uint32_t s = first; s < first + ns; s++ ) {
			float nx, ny, nz, d;
			map.side( s, nx, ny, nz, d );
			if( !side_step( r, ex, ey, ez, nx, ny, nz, d, tmin, tmax, ( int32_t ) s, b ) ) { alive = false; break; }
		}
		if( alive ) brush_finish( b, brush, h );
		if( nraw & kLastRef ) break;
		ri++;
	}
}

BSPK_HD void hit_init( HitState & h ) {
	h.t = 1.0f; h.hit = 0; h.far_side = -1; h.brush = -1; h.startsolid = 0;
}

/* BSP::trace_seg (bsp.cc:255-271) for one segment, start to finish.  Used by the
 * thread-per-segment kernel and by tests/hostsim; the persistent kernel inlines the same
 * pieces around its warp-level work fetch. */
template< class Map, int kStack >
BSPK_HD void trace_one( const Map & map, const Ray & r, HitState & h ) {
	StackEntry stack[ kStack ];
	int32_t sp = 0;
	int32_t node = 0;
	float tmin = 0.0f, tmax = 1.0f;
	hit_init( h );
	for( ;; ) {
		node = descend( map, r, node, tmin, tmax, stack, sp );
		if( node != kEmptyLeaf ) {
			leaf( map, r, node, tmin, tmax, h );
			if( h.hit ) break;                                                 /* bsp.cc:206: every later call returns at once */
		}
		/* next pending far side; entries that are empty leaves do nothing */
		do {
			if( sp == 0 ) return;
			sp--;
			node = stack[ sp ].node;
			tmin = stack[ sp ].t;
			tmax = sp > 0 ? stack[ sp - 1 ].t : 1.0f;
		} while( node == kEmptyLeaf );
	}
}

/* BSP::position_to_leaf (bsp.cc:273-286) on the flattened nodes; returns the file's leaf index */
template< class Map >
BSPK_HD int32_t point_leaf( const Map & map, float px, float py, float pz ) {
	int32_t node = 0;
	for( ;; ) {
		float nx, ny, nz, d; int32_t c0, c1, o0, o1;
		map.node_full( node, nx, ny, nz, d, c0, c1, o0, o1 );
		const float dist = dot3( px, py, pz, nx, ny, nz ) - d;
		const bool back = dist < 0.0f;
		const int32_t c = back ? c1 : c0;
		if( c < 0 ) return -( ( back ? o1 : o0 ) + 1 );
		node = c;
	}
}

/* wire format of one result == bspgpu_result (include/bspgpu.h) */
struct ResultRec {
	float t;
	int32_t plane;
	int32_t brush;
	uint32_t flags;   /* bit0 hit, bit1 startsolid */
};

BSPK_HD ResultRec make_result( const HitState & h, const uint32_t * side_plane ) {
	ResultRec o;
	o.t = h.t;
	o.plane = ( h.hit && h.far_side >= 0 ) ? ( int32_t ) side_plane[ h.far_side ] : -1;
	o.brush = h.hit ? h.brush : -1;
	o.flags = ( h.hit ? 1u : 0u ) | ( h.startsolid ? 2u : 0u );
	return o;
}

/* Intersection::pos (bsp.cc:265-267): start + t * (end - start) on a hit, the zero-initialised
 * (0,0,0) of `bis = { }` (bsp.cc:256) on a miss */
BSPK_HD void make_pos( const Ray & r, const HitState & h, float & px, float & py, float & pz ) {
	px = 0.0f; py = 0.0f; pz = 0.0f;
	if( h.hit ) {
		px = r.sx + h.t * r.dx; py = r.sy + h.t * r.dy; pz = r.sz + h.t * r.dz;
	}
}

/* plain-pointer accessor over the flattened image (global memory on the device, host memory in hostsim) */
struct ImageMap {
	const NodeRec * nodes;
	const BrushRef * refs;
	const SideRec * sides;
	BSPK_HD void node( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) const {
		const NodeRec & n = nodes[ i ];
		nx = n.nx; ny = n.ny; nz = n.nz; d = n.d; c0 = n.child[ 0 ]; c1 = n.child[ 1 ];
	}
	BSPK_HD void node_full( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1, int32_t & o0, int32_t & o1 ) const {
		const NodeRec & n = nodes[ i ];
		nx = n.nx; ny = n.ny; nz = n.nz; d = n.d; c0 = n.child[ 0 ]; c1 = n.child[ 1 ]; o0 = n.orig_child[ 0 ]; o1 = n.orig_child[ 1 ];
	}
	BSPK_HD void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const BrushRef & r = refs[ i ];
		first = r.first_side; nraw = r.num_sides; brush = r.brush;
	}
	BSPK_HD void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const SideRec & s = sides[ i ];
		nx = s.nx; ny = s.ny; nz = s.nz; d = s.d;
	}
};

} // namespace bspk

#endif
