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
 * Node planes whose normal is bitwise +e_x / +e_y / +e_z carry only their axis (map_layout.h: NodeRec).  Two ways to
 * evaluate them, both exact:
 *   general  rebuild the unit normal and run the same dot products as for any plane (node_plane_exact): bit-identical
 *            to the reference for EVERY input, because it is the reference's arithmetic;
 *   plain    for a ray whose six components are finite and whose start components are non-zero (ray_is_plain),
 *            (s.x*1 + s.y*0) + s.z*0 is bitwise s.x (adding +-0 to a non-zero finite number changes nothing) and
 *            (1*d.x + 0*d.y) + 0*d.z is d.x up to the sign of a zero, which no comparison and no stored value can
 *            see (a zero denominator fails `denom != 0`, bsp.cc:226).  The resumable walk the kernels run uses this
 *            form; rays that are not plain (NaN / infinite / zero start coordinates) are traced start to finish with
 *            the general form instead (trace_one), so the result is the reference's in both cases.
 *
 * Stack invariant of the simple form trace_one (lets an entry be 8 bytes; the resumable form the shipped kernel
 * runs stores its intervals explicitly, see below): the reference's far call gets [t, tmax] where
 * tmax is the interval end of the node that straddled (bsp.cc:250-252).  At any moment the
 * current interval end equals the start (`t`) of the entry on top of the stack, or 1.0f when
 * the stack is empty -- so `tmax` is not stored, it is read back from the entry below on pop.
 * For that to hold every straddle pushes, including far children that are empty leaves; those
 * are discarded at pop time.
 */
#ifndef BSPGPU_TRACE_CORE_H
#define BSPGPU_TRACE_CORE_H

#include <math.h>
#include <string.h>

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

/* plane and children of node i with the normal spelled out: an axis record becomes the exact unit normal it stands for */
template< class Map >
BSPK_HD void node_plane_exact( const Map & map, int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) {
	uint32_t w;
	map.node( i, d, c0, c1, w );
	if( w < kGeneralPlane ) { nx = w == 0u ? 1.0f : 0.0f; ny = w == 1u ? 1.0f : 0.0f; nz = w == 2u ? 1.0f : 0.0f; }
	else map.gplane( w - kGeneralPlane, nx, ny, nz );
}

/* see the header: dot(start, e_a) == start.a and dot(e_a, dir) == dir.a (up to the sign of a zero denominator) bit for bit */
BSPK_HD bool ray_is_plain( const Ray & r ) {
	const float p = ( r.sx * r.sy ) * r.sz;   /* finite and non-zero => every start component is finite and non-zero */
	const float q = ( r.dx * r.dy ) * r.dz;   /* finite => every direction component is finite (inf * 0 = NaN) */
	/* x * 0 is 0 for finite x and NaN for an infinite or NaN x: one multiply-add and one compare test both products
	 * (products that overflow or underflow send a plain ray down the general path: slower, never wrong) */
	return p != 0.0f && ( p * 0.0f + q * 0.0f ) == 0.0f;
}

/* trace_seg_tree from `node` down to a leaf reference; pending far sides go on the stack. */
template< class Map >
BSPK_HD int32_t descend( const Map & map, const Ray & r, int32_t node, float tmin, float & tmax, StackEntry * stack, int32_t & sp ) {
	while( node >= 0 ) {
		float nx, ny, nz, d; int32_t c0, c1;
		node_plane_exact( map, node, nx, ny, nz, d, c0, c1 );
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

/* ------------------------------------------------------------------------------------------
 * Resumable per-lane state machine: the same walk as trace_one, cut into two kinds of steps
 * so that a warp can run "node steps" and "side steps" in separate, bounded phases and pick
 * the phase most of its lanes are waiting for (trace_kernels.cuh).  Both phases can stop after
 * any number of steps and resume later without changing one arithmetic operation or its order.
 *
 * Differences from trace_one that do not touch the arithmetic:
 *  - stack entries carry their interval explicitly {node, tmin, tmax} (16 B, one 128-bit local
 *    load/store), so far children that are empty leaves are never pushed, a straddle whose
 *    near child is an empty leaf continues straight into the far child, and a pop is one load;
 *  - the step bodies are written as selects over predicates (same comparisons as the
 *    reference, evaluated unconditionally) so that lanes of a warp do not diverge inside a step;
 *    rare events (brush end, leaf end, pop, next brush reference) are short separate blocks.
 * ------------------------------------------------------------------------------------------ */

/* Lane state is packed so the step bodies stay short:
 *   cur >= 0                  node mode: index of the node to visit next
 *   kCurDone < cur < 0        leaf mode: ~cur = brush reference being clipped (the child encoding itself);
 *                             s == s_end means "fetch that reference first"
 *   cur == kCurDone / kCurIdle  finished (result not yet written) / no segment                              */
static const int32_t kCurIdle = INT32_MIN;        /* == kEmptyLeaf, which is never a live state */
static const int32_t kCurDone = INT32_MIN + 1;
static const uint32_t kBrushHit = 1u, kBrushInside = 2u, kBrushLastRef = 4u;
static const uint32_t kHitFlag = 1u, kStartSolidFlag = 2u;
static const uint32_t kSlotHit = 8u, kSlotStartSolid = 16u;   /* the same two flags where LaneS keeps them: beside the brush flags */

struct alignas( 16 ) StackEntry16 {
	int32_t node;
	float tmin, tmax;
	int32_t pad;
};

struct Lane {
	Ray r;
	float tmin, tmax;     /* current interval (the tmin/tmax arguments of bsp.cc:205) */
	int32_t cur;
	int32_t sp;           /* stack height */
	/* running result (bis.hit / bis.is.t, bsp.h:149-154, + extensions) */
	float ht;
	uint32_t hflags;      /* kHitFlag | kStartSolidFlag */
	int32_t hside, hbrush;
	/* cursor inside a leaf */
	uint32_t s, s_end;    /* side range still to test */
	int32_t brush;
	uint32_t bflags;      /* kBrushHit | kBrushInside | kBrushLastRef */
	float tfar, tnear;
	int32_t far_side;
};

BSPK_HD bool lane_in_node( const Lane & l ) { return l.cur >= 0; }
BSPK_HD bool lane_in_leaf( const Lane & l ) { return ( uint32_t ) l.cur > ( uint32_t ) kCurDone; }   /* negative, above kCurDone */
BSPK_HD bool lane_done( const Lane & l ) { return l.cur == kCurDone; }
BSPK_HD bool lane_idle( const Lane & l ) { return l.cur == kCurIdle; }

/* Where the pending far sides live is the caller's business: the step functions below are written against
 * stack_put / stack_get.  A plain array (local memory on the device) is the default; the kernels add a variant
 * that keeps the first slots in shared memory (trace_kernels.cuh: SmemStack). */
BSPK_HD void stack_put( StackEntry16 * s, int32_t i, int32_t node, float tmin, float tmax ) {
	/* three words, not four: local memory interleaves threads word by word, so every word stored is another L1 wavefront */
	s[ i ].node = node; s[ i ].tmin = tmin; s[ i ].tmax = tmax;
}
BSPK_HD void stack_get( const StackEntry16 * s, int32_t i, int32_t & node, float & tmin, float & tmax ) {
	const StackEntry16 e = s[ i ];
	node = e.node; tmin = e.tmin; tmax = e.tmax;
}

/* slots [0, D) in one array, the rest in another -- the host-side model of SmemStack (tests/hostsim) */
template< int D >
struct SplitStack {
	StackEntry16 * lo, * hi;
};
template< int D >
BSPK_HD void stack_put( const SplitStack< D > & s, int32_t i, int32_t node, float tmin, float tmax ) {
	if( i < D ) stack_put( s.lo, i, node, tmin, tmax );
	else stack_put( s.hi, i - D, node, tmin, tmax );
}
template< int D >
BSPK_HD void stack_get( const SplitStack< D > & s, int32_t i, int32_t & node, float & tmin, float & tmax ) {
	if( i < D ) stack_get( s.lo, i, node, tmin, tmax );
	else stack_get( s.hi, i - D, node, tmin, tmax );
}

BSPK_HD void lane_start( Lane & l, float sx, float sy, float sz, float ex, float ey, float ez ) {
	l.r.sx = sx; l.r.sy = sy; l.r.sz = sz;
	l.r.dx = ex - sx; l.r.dy = ey - sy; l.r.dz = ez - sz;      /* bsp.cc:261 */
	l.ht = 1.0f; l.hflags = 0; l.hside = -1; l.hbrush = -1;      /* bsp.cc:256-257 */
	l.tmin = 0.0f; l.tmax = 1.0f; l.cur = 0; l.sp = 0;          /* bsp.cc:263 */
	l.s = 0; l.s_end = 0;
	l.brush = -1; l.bflags = 0; l.tfar = 0.0f; l.tnear = 0.0f; l.far_side = -1;
}

struct StartGrid {
	const int32_t * cells;   /* null = disabled */
	StartGridDesc d;
};

/* where the walk of segment (s, e) may start: node 0 (bsp.cc:263), or -- when both endpoints lie in the same
 * start-grid cell -- the child reference stored for that cell (map_layout.h).  Comparisons are written so
 * that NaN / infinite coordinates fall back to node 0. */
BSPK_HD int32_t grid_start_ref( const StartGrid & g, float sx, float sy, float sz, float ex, float ey, float ez ) {
	if( !g.cells ) return 0;
	const StartGridDesc & d = g.d;
	/* cell coordinates of both endpoints, as floats.  Same cell <=> the floors are equal -- false for NaN, and an
	 * infinite coordinate fails the range test -- so non-finite input starts at node 0 without a test of its own.
	 * A point whose coordinate rounds to dim (it lies within an ulp of the upper face) is sent to node 0 as well:
	 * starting at the root is always exact, only starting below it needs the cell's guarantee. */
	const float cx = floorf( ( sx - d.lo[ 0 ] ) * d.inv[ 0 ] ), cy = floorf( ( sy - d.lo[ 1 ] ) * d.inv[ 1 ] ), cz = floorf( ( sz - d.lo[ 2 ] ) * d.inv[ 2 ] );
	const float ox = floorf( ( ex - d.lo[ 0 ] ) * d.inv[ 0 ] ), oy = floorf( ( ey - d.lo[ 1 ] ) * d.inv[ 1 ] ), oz = floorf( ( ez - d.lo[ 2 ] ) * d.inv[ 2 ] );
	const bool same = cx == ox && cy == oy && cz == oz;
	const bool inside = cx >= 0.0f && cx < ( float ) d.dim[ 0 ] && cy >= 0.0f && cy < ( float ) d.dim[ 1 ] && cz >= 0.0f && cz < ( float ) d.dim[ 2 ];
	if( !( same && inside ) ) return 0;
	const int32_t ix = ( int32_t ) cx, iy = ( int32_t ) cy, iz = ( int32_t ) cz;
	return g.cells[ ( uint32_t ) ( ( iz * d.dim[ 1 ] + iy ) * d.dim[ 0 ] + ix ) ];   /* < 2^30 cells (flatten.cc) */
}

/* begin the walk at a start reference from grid_start_ref (0 = the root) */
BSPK_HD void lane_start_at( Lane & l, int32_t ref ) {
	l.cur = ref == kEmptyLeaf ? kCurDone : ref;   /* the whole segment lies in a leaf without solid brushes: a miss, t = 1 */
}

BSPK_HD void lane_hit_state( const Lane & l, HitState & h ) {
	h.t = l.ht; h.hit = ( l.hflags & kHitFlag ) ? 1 : 0; h.startsolid = ( l.hflags & kStartSolidFlag ) ? 1 : 0;
	h.far_side = l.hside; h.brush = l.hbrush;
}

/* one trace_seg_tree step (bsp.cc:213-252).  Written as selects over the reference's comparisons:
 *   near child has work            -> go there over [tmin, u or tmax]; push the far child over [u, tmax] if it has work
 *   near child is an empty leaf    -> go straight to the far child over [u, tmax] if the segment straddles and it has work
 *   neither                        -> next pending far side from the stack (the second recursive call, :250-252), or done */
template< class Stack >
BSPK_HD void lane_node_apply( Lane & l, const Stack & stack, float denom, float dist, int32_t c0, int32_t c1 ) {
	const float u = dist / denom;                                                  /* :227 */
	/* :226,:232.  `denom != 0` is implied: a zero denominator makes u +-inf or NaN, and tmax is finite (1, or an earlier u that
	 * passed this very test), so `u >= 0 && u <= tmax` already fails */
	const bool in_range = u >= 0.0f && u <= l.tmax;
	const bool below = u < l.tmin;                                                 /* :235 */
	const bool both = in_range && !below;                                          /* :238 */
	const bool near_back = ( dist > 0.0f ) != ( in_range && below );               /* :222,:236 */
	const int32_t cn = near_back ? c1 : c0;                                        /* children[ near ] */
	const int32_t cf = near_back ? c0 : c1;                                        /* children[ !near ] */
	const bool near_ok = cn != kEmptyLeaf;
	const bool far_work = both && cf != kEmptyLeaf;
	const bool do_push = near_ok && far_work;
	const bool do_pop = !near_ok && !far_work;
	const bool can_pop = do_pop && l.sp > 0;
	if( do_push ) {
		stack_put( stack, l.sp, cf, u, l.tmax );
		l.sp++;
	}
	/* continue into the near child, else the far child; with neither, the next pending far side (or done) */
	l.cur = near_ok ? cn : cf;
	l.tmax = ( near_ok && both ) ? u : l.tmax;
	l.tmin = near_ok ? l.tmin : u;
	if( do_pop ) l.cur = kCurDone;
	if( can_pop ) {
		l.sp--;
		stack_get( stack, l.sp, l.cur, l.tmin, l.tmax );
	}
}

/* component `w` (0, 1, 2) of the ray's start and direction.  On the device this is spelled as predicated selects: written as
 * `w == 0 ? x : w == 1 ? y : z` the compiler turns it into an INDEXED load and moves the whole ray to local memory (two LDL per
 * node step and the lane state spilled around it: measured -19 % on long rays). */
BSPK_HD void ray_axis( const Ray & r, uint32_t w, float & s, float & d ) {
#ifdef __CUDA_ARCH__
	/* w is 0, 1 or 2: bit 0 <=> y, bit 1 <=> z (ptxas turns the two bit tests into one R2P) */
	asm( "{\n\t.reg .pred p1, p2;\n\t.reg .b32 t1, t2;\n\tand.b32 t1, %2, 1;\n\tand.b32 t2, %2, 2;\n\tsetp.ne.u32 p1, t1, 0;\n\tsetp.ne.u32 p2, t2, 0;\n\t"
	     "selp.f32 %0, %4, %3, p1;\n\tselp.f32 %0, %5, %0, p2;\n\tselp.f32 %1, %7, %6, p1;\n\tselp.f32 %1, %8, %1, p2;\n\t}"
	     : "=&f"( s ), "=&f"( d ) : "r"( w ), "f"( r.sx ), "f"( r.sy ), "f"( r.sz ), "f"( r.dx ), "f"( r.dy ), "f"( r.dz ) );
#else
	s = w == 0u ? r.sx : ( w == 1u ? r.sy : r.sz );
	d = w == 0u ? r.dx : ( w == 1u ? r.dy : r.dz );
#endif
}

/* the lane's ray must be plain (ray_is_plain) */
template< class Map, class Stack >
BSPK_HD void lane_node_step( const Map & map, Lane & l, const Stack & stack ) {
	float d; int32_t c0, c1; uint32_t w;
	map.node( l.cur, d, c0, c1, w );
	float denom, sdot;
	/* Map::kGeneralPlanes == false: the map has no node plane other than +unit normals (every kd-style map), and the general
	 * path is not even compiled in -- left in, the compiler predicates it into every step (+18 instructions of 67) */
	if( !Map::kGeneralPlanes || w < kGeneralPlane ) ray_axis( l.r, w, sdot, denom );   /* = dot( start, e_w ) :221, dot( e_w, dir ) :220 */
	else {
		float nx, ny, nz;
		map.gplane( w - kGeneralPlane, nx, ny, nz );
		denom = dot3( nx, ny, nz, l.r.dx, l.r.dy, l.r.dz );                        /* :220 */
		sdot = dot3( l.r.sx, l.r.sy, l.r.sz, nx, ny, nz );
	}
	lane_node_apply( l, stack, denom, -( sdot - d ), c0, c1 );                     /* dist, :221 */
}

/* up to `max_steps` node steps; stops early when the lane leaves node mode */
template< class Map, class Stack >
BSPK_HD void lane_node_phase( const Map & map, Lane & l, const Stack & stack, uint32_t max_steps ) {
	if( !lane_in_node( l ) ) return;
	for( ; max_steps > 0 && lane_in_node( l ); max_steps-- ) {
		lane_node_step( map, l, stack );
	}
	l.s = 0; l.s_end = 0;                                                          /* if the walk reached a leaf: fetch its first brush reference */
}

/* one turn of trace_seg_brush's side loop (bsp.cc:128-165) for side `index` with plane (n, d); `live` = false leaves the
 * lane untouched.  Returns whether the side rejected the brush (:135). */
template< class LaneT >
BSPK_HD bool lane_side_apply( LaneT & l, float ex, float ey, float ez, int32_t index, bool live, float nx, float ny, float nz, float d ) {
	const float start_dist = dot3( l.r.sx, l.r.sy, l.r.sz, nx, ny, nz ) - d;    /* :131 */
	const float end_dist = dot3( ex, ey, ez, nx, ny, nz ) - d;                  /* :132 */
	const float denom = -dot3( nx, ny, nz, l.r.dx, l.r.dy, l.r.dz );           /* :141 */
	const float t = start_dist / denom;                                         /* :142 */
	const bool rejected = live && start_dist > 0.0f && end_dist > 0.0f;         /* :135 */
	const bool behind = start_dist <= 0.0f && end_dist <= 0.0f;                 /* :137 */
	const bool crossing = live && !rejected && !behind;
	const bool entering = start_dist >= 0.0f;
	const bool in_range = crossing && t >= l.tmin && t <= l.tmax;               /* :144 */
	const bool raise_far = in_range && entering && t >= l.tfar;                 /* :147-148 */
	const bool lower_near = in_range && !entering && t <= l.tnear;              /* :154 */
	if( crossing && entering ) l.bflags &= ~kBrushInside;                       /* :139 */
	if( raise_far ) { l.tfar = t; l.far_side = index; }
	if( lower_near ) l.tnear = t;
	if( raise_far || lower_near ) l.bflags |= kBrushHit;
	return rejected;
}

/* up to `max_steps` iterations of trace_seg_brush's side loop (bsp.cc:128-165), moving through the
 * leaf's brushes (trace_seg_leaf, bsp.cc:188-201).  Flatten guarantees every referenced brush has >= 1 side. */
template< class Map, class Stack >
BSPK_HD void lane_leaf_phase( const Map & map, Lane & l, const Stack & stack, uint32_t max_steps ) {
	/* end = start + dir * tmax (bsp.cc:121); tmax is the leaf interval's and is constant inside a leaf */
	float ex = l.r.sx + l.r.dx * l.tmax, ey = l.r.sy + l.r.dy * l.tmax, ez = l.r.sz + l.r.dz * l.tmax;
	for( ; max_steps > 0 && lane_in_leaf( l ); max_steps-- ) {
		if( l.s == l.s_end ) {
			/* next brush of the leaf: reference ~cur; reset the clip state (bsp.cc:122-125) */
			uint32_t first, nraw;
			map.ref( ( uint32_t ) ~l.cur, first, nraw, l.brush );
			l.s = first; l.s_end = first + ( nraw & ~kLastRef );
			l.bflags = kBrushInside | ( ( nraw & kLastRef ) ? kBrushLastRef : 0u );
			l.tfar = l.tmin; l.tnear = l.tmax; l.far_side = -1;
		}
		/* two sides per step: flatten starts every brush at an even side index with an even count (zero-normal padding
		 * sides change nothing), so sides s and s+1 arrive in one 32-byte fetch.  They are clipped in order; the second
		 * one only if the first did not reject the brush, exactly like two turns of the reference's loop. */
		float n0x, n0y, n0z, d0, n1x, n1y, n1z, d1;
		map.side_pair( l.s, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
		bool rejected = lane_side_apply( l, ex, ey, ez, ( int32_t ) l.s, true, n0x, n0y, n0z, d0 );
		rejected = lane_side_apply( l, ex, ey, ez, ( int32_t ) l.s + 1, !rejected, n1x, n1y, n1z, d1 ) || rejected;
		l.s += 2;
		if( !rejected && l.s != l.s_end ) continue;

		/* ---- brush over (rejected, or side loop complete): bsp.cc:169-182 folded into :197-200 */
		if( !rejected ) {
			const bool inside = ( l.bflags & kBrushInside ) != 0;
			if( inside ) l.hflags |= kStartSolidFlag;
			const bool brush_hit = ( l.bflags & kBrushHit ) && ( inside ? ( l.tnear < l.tfar ) : ( l.tfar <= l.tnear ) );   /* :171,:176 */
			if( brush_hit ) {
				const float bt = inside ? l.tnear : l.tfar;
				if( !( l.hflags & kHitFlag ) || bt < l.ht ) { l.hside = inside ? -1 : l.far_side; l.hbrush = l.brush; }
				l.hflags |= kHitFlag;                                               /* :198 */
				if( bt < l.ht ) l.ht = bt;                                          /* :199 */
			}
		}
		l.s = l.s_end;                                                              /* a rejected brush skips its remaining sides */
		if( !( l.bflags & kBrushLastRef ) ) { l.cur--; continue; }                  /* ~cur + 1: next brush reference of this leaf */

		/* ---- leaf over: a hit ends the whole trace (bsp.cc:206), otherwise the next pending far side */
		if( l.hflags & kHitFlag ) { l.cur = kCurDone; break; }
		if( l.sp == 0 ) { l.cur = kCurDone; break; }
		l.sp--;
		stack_get( stack, l.sp, l.cur, l.tmin, l.tmax );
		l.s = 0; l.s_end = 0;
		ex = l.r.sx + l.r.dx * l.tmax; ey = l.r.sy + l.r.dy * l.tmax; ez = l.r.sz + l.r.dz * l.tmax;
	}
}

/* trace_one written with the resumable phases (budgets make no difference to the result); kSplit > 0 keeps the
 * first kSplit stack slots in a separate array, the way the kernels keep them in shared memory */
template< class Map, int kStack, int kSplit = 0 >
BSPK_HD void trace_one_phased( const Map & map, float sx, float sy, float sz, float ex, float ey, float ez,
                               uint32_t node_budget, uint32_t side_budget, Ray & r_out, HitState & h_out, const StartGrid * grid = nullptr ) {
	StackEntry16 stack[ kStack ];
	StackEntry16 low[ kSplit > 0 ? kSplit : 1 ];
	Lane l;
	lane_start( l, sx, sy, sz, ex, ey, ez );
	if( !ray_is_plain( l.r ) ) {                       /* as the kernels do: such a ray takes the general arithmetic start to finish */
		r_out = l.r;
		trace_one< Map, kStack >( map, l.r, h_out );
		return;
	}
	if( grid ) lane_start_at( l, grid_start_ref( *grid, sx, sy, sz, ex, ey, ez ) );
	while( !lane_done( l ) ) {
		if constexpr( kSplit > 0 ) {
			SplitStack< kSplit > st; st.lo = low; st.hi = stack;
			lane_node_phase( map, l, st, node_budget );
			lane_leaf_phase( map, l, st, side_budget );
		}
		else {
			StackEntry16 * st = stack;
			lane_node_phase( map, l, st, node_budget );
			lane_leaf_phase( map, l, st, side_budget );
		}
	}
	r_out = l.r;
	lane_hit_state( l, h_out );
}

/* ------------------------------------------------------------------------------------------
 * The same resumable walk over CHILD SLOTS (map_layout.h: Slot) -- what the GLOBAL-variant kernel (trace_slots) runs; the
 * SMEM-variant kernel (trace_phased) runs the node-record walk above.
 *
 * A lane carries the slot it is visiting {ca, cx}: plane + where the children are for an internal node, first brush
 * reference for a leaf.  A node step first issues the 16-byte gather of the node's two child slots, then runs the plane
 * test on what it already holds, and only then selects near / far from the pair: the gather of step k + 1 no longer waits
 * for the division of step k.  Pending far sides are pushed as whole slots {x, a, tmin, tmax} (16 bytes), so a pop touches
 * no map record either.  Every fp32 operation and comparison is the one lane_node_apply / lane_leaf_phase above perform.
 * ------------------------------------------------------------------------------------------ */
struct LaneS {
	Ray r;
	float tmin, tmax;
	uint32_t ca, cx;      /* the slot being visited; cx also encodes done / idle (kSlotDone / kSlotIdle) */
	int32_t sp;
	float ht;
	int32_t hside, hbrush;
	uint32_t s, s_end;
	uint32_t bflags;      /* kBrushHit | kBrushInside | kBrushLastRef of the brush being clipped, and the running result's
	                         kSlotHit | kSlotStartSolid (one register for both sets) */
	float tfar, tnear;
	int32_t far_side;
};

BSPK_HD bool slane_in_node( const LaneS & l ) { return ( l.cx & 3u ) != 3u; }
BSPK_HD bool slane_in_leaf( const LaneS & l ) { return l.cx == kSlotLeaf; }
BSPK_HD bool slane_done( const LaneS & l ) { return l.cx == kSlotDone; }
BSPK_HD bool slane_idle( const LaneS & l ) { return l.cx == kSlotIdle; }

BSPK_HD float bits_to_float( uint32_t b ) {
#ifdef __CUDA_ARCH__
	return __uint_as_float( b );
#else
	float f; memcpy( &f, &b, 4 ); return f;
#endif
}

struct alignas( 16 ) SlotEntry {
	uint32_t x, a;
	float tmin, tmax;
};
BSPK_HD void sstack_put( SlotEntry * s, int32_t i, uint32_t x, uint32_t a, float tmin, float tmax ) {
	s[ i ].x = x; s[ i ].a = a; s[ i ].tmin = tmin; s[ i ].tmax = tmax;
}
BSPK_HD void sstack_get( const SlotEntry * s, int32_t i, uint32_t & x, uint32_t & a, float & tmin, float & tmax ) {
	const SlotEntry e = s[ i ];
	x = e.x; a = e.a; tmin = e.tmin; tmax = e.tmax;
}
template< int D >
struct SplitSlotStack {
	SlotEntry * lo, * hi;
};
template< int D >
BSPK_HD void sstack_put( const SplitSlotStack< D > & s, int32_t i, uint32_t x, uint32_t a, float tmin, float tmax ) {
	if( i < D ) sstack_put( s.lo, i, x, a, tmin, tmax );
	else sstack_put( s.hi, i - D, x, a, tmin, tmax );
}
template< int D >
BSPK_HD void sstack_get( const SplitSlotStack< D > & s, int32_t i, uint32_t & x, uint32_t & a, float & tmin, float & tmax ) {
	if( i < D ) sstack_get( s.lo, i, x, a, tmin, tmax );
	else sstack_get( s.hi, i - D, x, a, tmin, tmax );
}

struct SlotGrid {
	const Slot * cells;   /* null = disabled */
	StartGridDesc d;
};

BSPK_HD void slane_start( LaneS & l, float sx, float sy, float sz, float ex, float ey, float ez ) {
	l.r.sx = sx; l.r.sy = sy; l.r.sz = sz;
	l.r.dx = ex - sx; l.r.dy = ey - sy; l.r.dz = ez - sz;      /* bsp.cc:261 */
	l.ht = 1.0f; l.hside = -1; l.hbrush = -1;                    /* bsp.cc:256-257 */
	l.tmin = 0.0f; l.tmax = 1.0f; l.sp = 0;                      /* bsp.cc:263 */
	l.s = 0; l.s_end = 0;
	l.bflags = 0; l.tfar = 0.0f; l.tnear = 0.0f; l.far_side = -1;
}

/* the slot where the walk of segment (s, e) may start: the root's (bsp.cc:263), or -- both endpoints in one start-grid cell --
 * the one stored for that cell (same lookup and same guarantees as grid_start_ref above) */
BSPK_HD Slot grid_start_slot( const SlotGrid & g, const Slot & root, float sx, float sy, float sz, float ex, float ey, float ez ) {
	if( !g.cells ) return root;
	const StartGridDesc & d = g.d;
	const float cx = floorf( ( sx - d.lo[ 0 ] ) * d.inv[ 0 ] ), cy = floorf( ( sy - d.lo[ 1 ] ) * d.inv[ 1 ] ), cz = floorf( ( sz - d.lo[ 2 ] ) * d.inv[ 2 ] );
	const float ox = floorf( ( ex - d.lo[ 0 ] ) * d.inv[ 0 ] ), oy = floorf( ( ey - d.lo[ 1 ] ) * d.inv[ 1 ] ), oz = floorf( ( ez - d.lo[ 2 ] ) * d.inv[ 2 ] );
	const bool same = cx == ox && cy == oy && cz == oz;
	const bool inside = cx >= 0.0f && cx < ( float ) d.dim[ 0 ] && cy >= 0.0f && cy < ( float ) d.dim[ 1 ] && cz >= 0.0f && cz < ( float ) d.dim[ 2 ];
	if( !( same && inside ) ) return root;
	const int32_t ix = ( int32_t ) cx, iy = ( int32_t ) cy, iz = ( int32_t ) cz;
#ifdef __CUDA_ARCH__
	const uint2 v = __ldg( reinterpret_cast< const uint2 * >( g.cells + ( uint32_t ) ( ( iz * d.dim[ 1 ] + iy ) * d.dim[ 0 ] + ix ) ) );
	Slot s; s.a = v.x; s.x = v.y;
	return s;
#else
	return g.cells[ ( uint32_t ) ( ( iz * d.dim[ 1 ] + iy ) * d.dim[ 0 ] + ix ) ];
#endif
}

BSPK_HD void slane_start_at( LaneS & l, const Slot & s ) {
	l.ca = s.a;
	l.cx = s.x == kSlotEmpty ? kSlotDone : s.x;   /* the whole segment lies in a leaf without solid brushes: a miss, t = 1 */
}

BSPK_HD void slane_hit_state( const LaneS & l, HitState & h ) {
	h.t = l.ht; h.hit = ( l.bflags & kSlotHit ) ? 1 : 0; h.startsolid = ( l.bflags & kSlotStartSolid ) ? 1 : 0;
	h.far_side = l.hside; h.brush = l.hbrush;
}

/* component (0, 1, 2) of the ray's start and direction named by the low bits of a node slot's x (see ray_axis) */
BSPK_HD void ray_axis_bits( const Ray & r, uint32_t x, float & s, float & d ) {
#ifdef __CUDA_ARCH__
	asm( "{\n\t.reg .pred p1, p2;\n\t.reg .b32 t1, t2;\n\tand.b32 t1, %2, 1;\n\tand.b32 t2, %2, 2;\n\tsetp.ne.u32 p1, t1, 0;\n\tsetp.ne.u32 p2, t2, 0;\n\t"
	     "selp.f32 %0, %4, %3, p1;\n\tselp.f32 %0, %5, %0, p2;\n\tselp.f32 %1, %7, %6, p1;\n\tselp.f32 %1, %8, %1, p2;\n\t}"
	     : "=&f"( s ), "=&f"( d ) : "r"( x ), "f"( r.sx ), "f"( r.sy ), "f"( r.sz ), "f"( r.dx ), "f"( r.dy ), "f"( r.dz ) );
#else
	s = ( x & 2u ) ? r.sz : ( ( x & 1u ) ? r.sy : r.sx );
	d = ( x & 2u ) ? r.dz : ( ( x & 1u ) ? r.dy : r.dx );
#endif
}

/* one trace_seg_tree step (bsp.cc:213-252) at the internal node whose slot the lane holds; the lane's ray must be plain */
template< class Map, class Stack >
BSPK_HD void slane_node_step( const Map & map, LaneS & l, const Stack & stack ) {
	uint32_t a0, x0, a1, x1;
	map.pair( l.cx >> 3, a0, x0, a1, x1 );                                         /* children[ 0 ], children[ 1 ]: needed last, asked for first */
	float denom, sdot, d;
	if( !Map::kGeneralPlanes || ( l.cx & kSlotGeneral ) == 0u ) {
		ray_axis_bits( l.r, l.cx, sdot, denom );                                   /* = dot( start, e_w ) :221, dot( e_w, dir ) :220 */
		d = bits_to_float( l.ca );
	}
	else {
		float nx, ny, nz;
		map.gplane_d( l.ca, nx, ny, nz, d );
		denom = dot3( nx, ny, nz, l.r.dx, l.r.dy, l.r.dz );                        /* :220 */
		sdot = dot3( l.r.sx, l.r.sy, l.r.sz, nx, ny, nz );
	}
	const float dist = -( sdot - d );                                              /* :221 */
	const float u = dist / denom;                                                  /* :227 */
	/* :226,:232.  The reference's `denom != 0` is implied: a zero denominator makes u +-inf or NaN, and tmax is finite
	 * (1, or an earlier u that passed this very test), so `u >= 0 && u <= tmax` already fails */
	const bool in_range = u >= 0.0f && u <= l.tmax;
	const bool below = u < l.tmin;                                                 /* :235 */
	const bool both = in_range && !below;                                          /* :238 */
	const bool near_back = ( dist > 0.0f ) != ( in_range && below );               /* :222,:236 */
	const uint32_t an = near_back ? a1 : a0, xn = near_back ? x1 : x0;             /* children[ near ] */
	const uint32_t af = near_back ? a0 : a1, xf = near_back ? x0 : x1;             /* children[ !near ] */
	const bool near_ok = xn != kSlotEmpty;
	const bool far_work = both && xf != kSlotEmpty;
	const bool do_push = near_ok && far_work;
	const bool do_pop = !near_ok && !far_work;
	const bool can_pop = do_pop && l.sp > 0;
	/* push and pop stay two separate short blocks: merged into one (one address computation, one reconvergence pair)
	 * the push lanes and the pop lanes still run one after the other, and the kernel measured 4 % slower */
	if( do_push ) {
		sstack_put( stack, l.sp, xf, af, u, l.tmax );
		l.sp++;
	}
	/* continue into the near child, else the far child; with neither, the next pending far side (or done) */
	l.ca = near_ok ? an : af;
	l.cx = near_ok ? xn : xf;
	l.tmax = ( near_ok && both ) ? u : l.tmax;
	l.tmin = near_ok ? l.tmin : u;
	if( do_pop ) l.cx = kSlotDone;
	if( can_pop ) {
		l.sp--;
		sstack_get( stack, l.sp, l.cx, l.ca, l.tmin, l.tmax );
	}
}

template< class Map, class Stack >
BSPK_HD void slane_node_phase( const Map & map, LaneS & l, const Stack & stack, uint32_t max_steps ) {
	if( !slane_in_node( l ) ) return;
	for( ; max_steps > 0 && slane_in_node( l ); max_steps-- ) {
		slane_node_step( map, l, stack );
	}
	l.s = 0; l.s_end = 0;                                                          /* if the walk reached a leaf: fetch its first brush reference */
}

/* lane_leaf_phase over slots: in leaf mode ca is the brush reference being clipped */
template< class Map, class Stack >
BSPK_HD void slane_leaf_phase( const Map & map, LaneS & l, const Stack & stack, uint32_t max_steps ) {
	float ex = l.r.sx + l.r.dx * l.tmax, ey = l.r.sy + l.r.dy * l.tmax, ez = l.r.sz + l.r.dz * l.tmax;   /* bsp.cc:121 */
	for( ; max_steps > 0 && slane_in_leaf( l ); max_steps-- ) {
		if( l.s == l.s_end ) {
			uint32_t first, nraw; int32_t brush;
			map.ref( l.ca, first, nraw, brush );
			l.s = first; l.s_end = first + ( nraw & ~kLastRef );
			l.bflags = ( l.bflags & ( kSlotHit | kSlotStartSolid ) ) | kBrushInside | ( ( nraw & kLastRef ) ? kBrushLastRef : 0u );
			l.tfar = l.tmin; l.tnear = l.tmax; l.far_side = -1;                         /* bsp.cc:122-125 */
		}
		float n0x, n0y, n0z, d0, n1x, n1y, n1z, d1;
		map.side_pair( l.s, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
		bool rejected = lane_side_apply( l, ex, ey, ez, ( int32_t ) l.s, true, n0x, n0y, n0z, d0 );
		rejected = lane_side_apply( l, ex, ey, ez, ( int32_t ) l.s + 1, !rejected, n1x, n1y, n1z, d1 ) || rejected;
		l.s += 2;
		if( !rejected && l.s != l.s_end ) continue;

		/* ---- brush over: bsp.cc:169-182 folded into :197-200 */
		if( !rejected ) {
			const bool inside = ( l.bflags & kBrushInside ) != 0;
			if( inside ) l.bflags |= kSlotStartSolid;
			const bool brush_hit = ( l.bflags & kBrushHit ) && ( inside ? ( l.tnear < l.tfar ) : ( l.tfar <= l.tnear ) );   /* :171,:176 */
			if( brush_hit ) {
				const float bt = inside ? l.tnear : l.tfar;
				if( !( l.bflags & kSlotHit ) || bt < l.ht ) {
					/* the brush index is not carried through the side loop: its reference is read again on the (rare) hit */
					uint32_t first, nraw;
					l.hside = inside ? -1 : l.far_side; map.ref( l.ca, first, nraw, l.hbrush );
				}
				l.bflags |= kSlotHit;                                               /* :198 */
				if( bt < l.ht ) l.ht = bt;                                          /* :199 */
			}
		}
		l.s = l.s_end;
		if( !( l.bflags & kBrushLastRef ) ) { l.ca++; continue; }                   /* next brush reference of this leaf */

		/* ---- leaf over: a hit ends the whole trace (bsp.cc:206), otherwise the next pending far side */
		if( l.bflags & kSlotHit ) { l.cx = kSlotDone; break; }
		if( l.sp == 0 ) { l.cx = kSlotDone; break; }
		l.sp--;
		sstack_get( stack, l.sp, l.cx, l.ca, l.tmin, l.tmax );
		l.s = 0; l.s_end = 0;
		ex = l.r.sx + l.r.dx * l.tmax; ey = l.r.sy + l.r.dy * l.tmax; ez = l.r.sz + l.r.dz * l.tmax;
	}
}

/* trace_one through the slot walk (budgets make no difference to the result); kSplit > 0 keeps the first kSplit stack
 * entries in a separate array, the way the kernels keep them in shared memory */
template< class Map, int kStack, int kSplit = 0 >
BSPK_HD void trace_one_slots( const Map & map, const Slot & root, float sx, float sy, float sz, float ex, float ey, float ez,
                              uint32_t node_budget, uint32_t side_budget, Ray & r_out, HitState & h_out, const SlotGrid * grid = nullptr ) {
	SlotEntry stack[ kStack ];
	SlotEntry low[ kSplit > 0 ? kSplit : 1 ];
	LaneS l;
	slane_start( l, sx, sy, sz, ex, ey, ez );
	if( !ray_is_plain( l.r ) ) {                       /* as the kernels do: such a ray takes the general arithmetic start to finish */
		r_out = l.r;
		trace_one< Map, kStack >( map, l.r, h_out );
		return;
	}
	slane_start_at( l, grid ? grid_start_slot( *grid, root, sx, sy, sz, ex, ey, ez ) : root );
	while( !slane_done( l ) ) {
		if constexpr( kSplit > 0 ) {
			SplitSlotStack< kSplit > st; st.lo = low; st.hi = stack;
			slane_node_phase( map, l, st, node_budget );
			slane_leaf_phase( map, l, st, side_budget );
		}
		else {
			SlotEntry * st = stack;
			slane_node_phase( map, l, st, node_budget );
			slane_leaf_phase( map, l, st, side_budget );
		}
	}
	r_out = l.r;
	slane_hit_state( l, h_out );
}

/* BSP::position_to_leaf (bsp.cc:273-286) on the flattened nodes; returns the file's leaf index */
template< class Map >
BSPK_HD int32_t point_leaf( const Map & map, float px, float py, float pz ) {
	int32_t node = 0;
	for( ;; ) {
		float nx, ny, nz, d; int32_t c0, c1;
		node_plane_exact( map, node, nx, ny, nz, d, c0, c1 );
		const float dist = dot3( px, py, pz, nx, ny, nz ) - d;
		const bool back = dist < 0.0f;
		const int32_t c = back ? c1 : c0;
		if( c < 0 ) return -( map.orig_child( node, back ? 1 : 0 ) + 1 );
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
	static const bool kGeneralPlanes = true;
	const NodeRec * nodes;
	const PlaneRec * gplanes;
	const BrushRef * refs;
	const SideRec * side_pairs;
	const NodeOrig * node_orig;
	const Slot * slots;
	BSPK_HD void pair( uint32_t p, uint32_t & a0, uint32_t & x0, uint32_t & a1, uint32_t & x1 ) const {
		a0 = slots[ 2 * p ].a; x0 = slots[ 2 * p ].x; a1 = slots[ 2 * p + 1 ].a; x1 = slots[ 2 * p + 1 ].x;
	}
	BSPK_HD void gplane_d( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const PlaneRec & p = gplanes[ i ];
		nx = p.nx; ny = p.ny; nz = p.nz; d = p.d;
	}
	BSPK_HD void node( int32_t i, float & d, int32_t & c0, int32_t & c1, uint32_t & w ) const {
		const NodeRec & n = nodes[ i ];
		d = n.d; c0 = n.child[ 0 ]; c1 = n.child[ 1 ]; w = n.w;
	}
	BSPK_HD void gplane( uint32_t i, float & nx, float & ny, float & nz ) const {
		const PlaneRec & p = gplanes[ i ];
		nx = p.nx; ny = p.ny; nz = p.nz;
	}
	BSPK_HD int32_t orig_child( int32_t i, int k ) const { return node_orig[ i ].orig_child[ k ]; }
	BSPK_HD void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const BrushRef & r = refs[ i ];
		first = r.first_side; nraw = r.num_sides; brush = r.brush;
	}
	BSPK_HD void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const SideRec & s = side_pairs[ i ];
		nx = s.nx; ny = s.ny; nz = s.nz; d = s.d;
	}
	BSPK_HD void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		side( i, n0x, n0y, n0z, d0 );
		side( i + 1, n1x, n1y, n1z, d1 );
	}
};

/* hostsim's model of the shared-memory accessor: sides come from the split even / odd arrays */
struct ImageMapSplit : ImageMap {
	const SideRec * sides_even, * sides_odd;
	BSPK_HD void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		const SideRec & a = sides_even[ i >> 1 ], & b = sides_odd[ i >> 1 ];
		n0x = a.nx; n0y = a.ny; n0z = a.nz; d0 = a.d; n1x = b.nx; n1y = b.ny; n1z = b.nz; d1 = b.d;
	}
};

} // namespace bspk

#endif
