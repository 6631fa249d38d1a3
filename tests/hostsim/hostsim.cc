/*
 * tests/hostsim/hostsim.cc -- TEST-ONLY CPU build of the device trace core.
 *
 * Compiles bsp_b200/csrc/trace_core.h (the exact inline functions the sm_100a kernels are
 * made of) and flatten.cc with g++ -ffp-contract=off, so the flattened image and the
 * iterative/stack formulation can be checked bit-for-bit against the reference object on
 * millions of segments in this GPU-less container.  It is not part of libbspgpu.so and the
 * product never calls it.
 */
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../bsp_b200/csrc/flatten.h"
#include "../../bsp_b200/csrc/trace_core.h"

using namespace bspk;

namespace {
struct Sim {
	FlatMap flat;
	ImageMap view;
	ImageMapPacked packed;
	const uint32_t * side_plane;
	StartGrid grid;
};
std::string g_err;
}

/* ImageMap that counts record fetches (for sizing the start grid, not used by the parity tests) */
struct CountingMap {
	static const bool kPackedNodes = false;
	ImageMap m;
	mutable uint64_t nodes = 0, refs = 0, sides = 0;
	void node( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) const { nodes++; m.node( i, nx, ny, nz, d, c0, c1 ); }
	void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const { refs++; m.ref( i, first, nraw, brush ); }
	void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const { sides++; m.side( i, nx, ny, nz, d ); }
	void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		sides++; m.side_pair( i, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

struct CountingMapPacked {
	static const bool kPackedNodes = true;
	ImageMapPacked m;
	mutable uint64_t nodes = 0, refs = 0, sides = 0;
	void qnode( int32_t i, uint32_t & q0, uint32_t & q1, uint32_t & q2, uint32_t & q3, uint32_t & q4, uint32_t & q5, uint32_t & q6, uint32_t & w ) const { nodes++; m.qnode( i, q0, q1, q2, q3, q4, q5, q6, w ); }
	void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const { refs++; m.ref( i, first, nraw, brush ); }
	void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const { sides++; m.side( i, nx, ny, nz, d ); }
	void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		sides++; m.side_pair( i, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

extern "C" {

/* record fetches of the phased walk over n segments, with or without the start grid: out = {nodes, refs, sides, started below the root} */
void hostsim_count( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n, int use_grid, uint64_t * out ) {
	const Sim * s = ( const Sim * ) h;
	/* use_grid: bit 0 = start grid, bit 1 = walk the packed two-level records */
	CountingMap cm; cm.m = s->view;
	CountingMapPacked cp; cp.m = s->packed;
	const bool grid = ( use_grid & 1 ) != 0, packed = ( use_grid & 2 ) != 0;
	uint64_t below = 0;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * a = segs + i * stride;
		const float * b = a + end_off;
		Ray r; HitState hs;
		if( grid && grid_start_ref( s->grid, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ] ) != 0 ) below++;
		if( packed ) trace_one_phased< CountingMapPacked, 64 >( cp, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], 8, 8, r, hs, grid ? &s->grid : nullptr );
		else trace_one_phased< CountingMap, 64 >( cm, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], 8, 8, r, hs, grid ? &s->grid : nullptr );
	}
	out[ 0 ] = cm.nodes + cp.nodes; out[ 1 ] = cm.refs + cp.refs; out[ 2 ] = cm.sides + cp.sides; out[ 3 ] = below;
}

void * hostsim_create( const bspgpu_map_desc * lumps ) {
	Sim * s = new Sim();
	if( flatten_map( *lumps, s->flat, g_err ) != BSPGPU_OK ) { delete s; return nullptr; }
	const unsigned char * img = s->flat.image.data();
	s->view.nodes = ( const NodeRec * ) ( img + s->flat.layout.off_nodes );
	s->view.refs = ( const BrushRef * ) ( img + s->flat.layout.off_refs );
	s->view.sides = ( const SideRec * ) ( img + s->flat.layout.off_sides );
	s->side_plane = ( const uint32_t * ) ( img + s->flat.layout.off_side_plane );
	s->packed.nodes = s->view.nodes; s->packed.refs = s->view.refs; s->packed.sides = s->view.sides;
	s->packed.qnodes = ( const QNodeRec * ) ( img + s->flat.layout.off_qnodes );
	s->grid.d = s->flat.layout.grid;
	s->grid.cells = s->flat.layout.grid.dim[ 0 ] ? ( const int32_t * ) ( img + s->flat.layout.off_grid ) : nullptr;
	return s;
}

const char * hostsim_error() { return g_err.c_str(); }

void hostsim_destroy( void * h ) { delete ( Sim * ) h; }

void hostsim_layout( void * h, uint64_t * out ) {
	const MapImageLayout & l = ( ( Sim * ) h )->flat.layout;
	out[ 0 ] = l.num_nodes; out[ 1 ] = l.num_refs; out[ 2 ] = l.num_sides; out[ 3 ] = l.tree_depth;
	out[ 4 ] = l.staged_bytes; out[ 5 ] = l.total_bytes;
	out[ 6 ] = ( uint64_t ) l.grid.dim[ 0 ] * l.grid.dim[ 1 ] * l.grid.dim[ 2 ];
	/* packed two-level records (QNodeRec) and the children folded into them */
	const QNodeRec * q = ( ( Sim * ) h )->packed.qnodes;
	out[ 7 ] = out[ 8 ] = 0;
	for( uint32_t i = 0; i < l.num_nodes; i++ ) {
		if( q[ i ].w[ 7 ] & kQPacked ) out[ 7 ]++;
		out[ 8 ] += ( ( q[ i ].w[ 7 ] & kQExpand0 ) ? 1 : 0 ) + ( ( q[ i ].w[ 7 ] & kQExpand1 ) ? 1 : 0 );
	}
}

/* segs: `stride` floats per segment, end point `end_off` floats after the start point */
void hostsim_trace( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n, ResultRec * out, float * pos4 ) {
	const Sim * s = ( const Sim * ) h;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * a = segs + i * stride;
		const float * b = a + end_off;
		Ray r;
		r.sx = a[ 0 ]; r.sy = a[ 1 ]; r.sz = a[ 2 ];
		r.dx = b[ 0 ] - a[ 0 ]; r.dy = b[ 1 ] - a[ 1 ]; r.dz = b[ 2 ] - a[ 2 ];
		HitState hs;
		trace_one< ImageMap, 64 >( s->view, r, hs );
		if( out ) out[ i ] = make_result( hs, s->side_plane );
		if( pos4 ) {
			make_pos( r, hs, pos4[ 4 * i + 0 ], pos4[ 4 * i + 1 ], pos4[ 4 * i + 2 ] );
			pos4[ 4 * i + 3 ] = hs.t;
		}
	}
}

/* same walk through the resumable node/leaf phases the persistent kernel runs, with small and
 * varying step budgets so every suspend/resume point is exercised */
void hostsim_trace_phased( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n, ResultRec * out, float * pos4,
                           uint32_t node_budget, uint32_t side_budget, int use_grid ) {
	const Sim * s = ( const Sim * ) h;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * a = segs + i * stride;
		const float * b = a + end_off;
		Ray r; HitState hs;
		const uint32_t nb = node_budget ? node_budget : 1 + ( uint32_t ) ( i % 7 );
		const uint32_t sb = side_budget ? side_budget : 1 + ( uint32_t ) ( ( i / 7 ) % 5 );
		/* use_grid: bit 0 = start grid, bit 1 = walk the packed two-level records (QNodeRec),
		 * bit 2 = first two stack slots in a separate array (the model of the kernels' shared-memory stack) */
		const StartGrid * grid = ( use_grid & 1 ) ? &s->grid : nullptr;
		if( use_grid & 4 ) trace_one_phased< ImageMap, 64, 2 >( s->view, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else if( use_grid & 2 ) trace_one_phased< ImageMapPacked, 64 >( s->packed, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else trace_one_phased< ImageMap, 64 >( s->view, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		if( out ) out[ i ] = make_result( hs, s->side_plane );
		if( pos4 ) {
			make_pos( r, hs, pos4[ 4 * i + 0 ], pos4[ 4 * i + 1 ], pos4[ 4 * i + 2 ] );
			pos4[ 4 * i + 3 ] = hs.t;
		}
	}
}

void hostsim_leaf( void * h, const float * pts, uint64_t stride, uint64_t n, int32_t * leaf ) {
	const Sim * s = ( const Sim * ) h;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * p = pts + i * stride;
		leaf[ i ] = point_leaf( s->view, p[ 0 ], p[ 1 ], p[ 2 ] );
	}
}

} // extern "C"
