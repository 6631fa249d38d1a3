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
	ImageMapSplit split;
	const uint32_t * side_plane;
	StartGrid grid;
	SlotGrid sgrid;
	Slot root;
};
std::string g_err;
}

/* ImageMap that counts record fetches (for sizing the start grid, not used by the parity tests) */
struct CountingMap {
	static const bool kGeneralPlanes = true;
	ImageMap m;
	mutable uint64_t nodes = 0, refs = 0, sides = 0;
	void node( int32_t i, float & d, int32_t & c0, int32_t & c1, uint32_t & w ) const { nodes++; m.node( i, d, c0, c1, w ); }
	void pair( uint32_t p, uint32_t & a0, uint32_t & x0, uint32_t & a1, uint32_t & x1 ) const { nodes++; m.pair( p, a0, x0, a1, x1 ); }
	void gplane_d( uint32_t i, float & nx, float & ny, float & nz, float & d ) const { nodes++; m.gplane_d( i, nx, ny, nz, d ); }
	void gplane( uint32_t i, float & nx, float & ny, float & nz ) const { nodes++; m.gplane( i, nx, ny, nz ); }
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
	CountingMap cm; cm.m = s->view;
	const bool grid = ( use_grid & 1 ) != 0;
	uint64_t below = 0;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * a = segs + i * stride;
		const float * b = a + end_off;
		Ray r; HitState hs;
		if( grid && grid_start_ref( s->grid, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ] ) != 0 ) below++;
		trace_one_slots< CountingMap, 128 >( cm, s->root, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], 8, 8, r, hs, grid ? &s->sgrid : nullptr );
	}
	out[ 0 ] = cm.nodes; out[ 1 ] = cm.refs; out[ 2 ] = cm.sides; out[ 3 ] = below;
}

/* flags: bit 0 = store every node plane as a general plane (no axis shortcut); grid_cells as bspgpu_create_opts */
void * hostsim_create_opts( const bspgpu_map_desc * lumps, int flags, int64_t grid_cells ) {
	Sim * s = new Sim();
	FlattenOptions opt;
	opt.general_planes_only = ( flags & 1 ) != 0;
	opt.grid_cells = grid_cells;
	if( flatten_map( *lumps, opt, s->flat, g_err ) != BSPGPU_OK ) { delete s; return nullptr; }
	const unsigned char * img = s->flat.image.data();
	const MapImageLayout & l = s->flat.layout;
	s->view.nodes = ( const NodeRec * ) ( img + l.off_nodes );
	s->view.gplanes = ( const PlaneRec * ) ( img + l.off_gplanes );
	s->view.refs = ( const BrushRef * ) ( img + l.off_refs );
	s->view.side_pairs = ( const SideRec * ) ( img + l.off_side_pairs );
	s->view.node_orig = ( const NodeOrig * ) ( img + l.off_node_orig );
	s->view.slots = ( const Slot * ) ( img + l.off_slots );
	s->root = l.root;
	s->sgrid.d = l.grid;
	s->sgrid.cells = l.grid.dim[ 0 ] ? ( const Slot * ) ( img + l.off_grid_slots ) : nullptr;
	static_cast< ImageMap & >( s->split ) = s->view;
	s->split.sides_even = ( const SideRec * ) ( img + l.off_sides_even );
	s->split.sides_odd = ( const SideRec * ) ( img + l.off_sides_odd );
	s->side_plane = ( const uint32_t * ) ( img + l.off_side_plane );
	s->grid.d = l.grid;
	s->grid.cells = l.grid.dim[ 0 ] ? ( const int32_t * ) ( img + l.off_grid ) : nullptr;
	return s;
}

void * hostsim_create( const bspgpu_map_desc * lumps ) { return hostsim_create_opts( lumps, 0, -1 ); }

const char * hostsim_error() { return g_err.c_str(); }

void hostsim_destroy( void * h ) { delete ( Sim * ) h; }

void hostsim_layout( void * h, uint64_t * out ) {
	const MapImageLayout & l = ( ( Sim * ) h )->flat.layout;
	out[ 0 ] = l.num_nodes; out[ 1 ] = l.num_refs; out[ 2 ] = l.num_sides; out[ 3 ] = l.tree_depth;
	out[ 4 ] = l.staged_bytes; out[ 5 ] = l.total_bytes;
	out[ 6 ] = ( uint64_t ) l.grid.dim[ 0 ] * l.grid.dim[ 1 ] * l.grid.dim[ 2 ];
	out[ 7 ] = l.num_gplanes;   /* node planes that are not a bitwise +unit normal */
	out[ 8 ] = 0;
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
		trace_one< ImageMap, 128 >( s->view, r, hs );
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
		/* use_grid: bit 0 = start grid, bit 1 = sides from the split even / odd arrays (the model of the shared-memory accessor),
		 * bit 2 = first two stack slots in a separate array (the model of the kernels' shared-memory stack) */
		const StartGrid * grid = ( use_grid & 1 ) ? &s->grid : nullptr;
		if( use_grid & 4 ) trace_one_phased< ImageMap, 128, 2 >( s->view, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else if( use_grid & 2 ) trace_one_phased< ImageMapSplit, 128 >( s->split, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else trace_one_phased< ImageMap, 128 >( s->view, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		if( out ) out[ i ] = make_result( hs, s->side_plane );
		if( pos4 ) {
			make_pos( r, hs, pos4[ 4 * i + 0 ], pos4[ 4 * i + 1 ], pos4[ 4 * i + 2 ] );
			pos4[ 4 * i + 3 ] = hs.t;
		}
	}
}

/* the walk the shipped kernels run: over child slots (trace_core.h: LaneS); flags as hostsim_trace_phased's use_grid */
void hostsim_trace_slots( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n, ResultRec * out, float * pos4,
                          uint32_t node_budget, uint32_t side_budget, int flags ) {
	const Sim * s = ( const Sim * ) h;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * a = segs + i * stride;
		const float * b = a + end_off;
		Ray r; HitState hs;
		const uint32_t nb = node_budget ? node_budget : 1 + ( uint32_t ) ( i % 7 );
		const uint32_t sb = side_budget ? side_budget : 1 + ( uint32_t ) ( ( i / 7 ) % 5 );
		const SlotGrid * grid = ( flags & 1 ) ? &s->sgrid : nullptr;
		if( flags & 4 ) trace_one_slots< ImageMap, 128, 2 >( s->view, s->root, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else if( flags & 2 ) trace_one_slots< ImageMapSplit, 128 >( s->split, s->root, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
		else trace_one_slots< ImageMap, 128 >( s->view, s->root, a[ 0 ], a[ 1 ], a[ 2 ], b[ 0 ], b[ 1 ], b[ 2 ], nb, sb, r, hs, grid );
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
