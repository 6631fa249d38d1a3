/*
 * bsp_host.cc -- see bsp_host.h.  Loads the file the way the reference does (one blob, typed
 * pointers into it, bsp.cc:58-86), hands the seven trace lumps to bspgpu_create, and forwards
 * every query to the C ABI.
 */
#include "bsp_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

[[noreturn]] void die( const char * what ) {
	fprintf( stderr, "bsp: %s: %s\n", what, bspgpu_last_error() );
	abort();
}

void check( int rc, const char * what ) {
	if( rc != BSPGPU_OK ) die( what );
}

struct Dirent { uint32_t off, len; };

template< typename T >
void point_lump( const BSP * bsp, long file_len, int slot, uint32_t & count, T *& ptr, const char * name ) {
	/* reference BSP::load_lump (bsp.cc:92-103): len must be a whole number of records */
	const Dirent * dir = reinterpret_cast< const Dirent * >( bsp->contents + 8 );
	const Dirent & d = dir[ slot ];
	if( d.len % sizeof( T ) != 0 || ( uint64_t ) d.off + d.len > ( uint64_t ) file_len ) {
		fprintf( stderr, "bsp: lump %s is malformed (off %u len %u)\n", name, d.off, d.len );
		abort();
	}
	count = d.len / ( uint32_t ) sizeof( T );
	ptr = reinterpret_cast< T * >( bsp->contents + d.off );
}

} // namespace

void bsp_init( BSP * bsp, std::string filename, int device ) {
	memset( static_cast< void * >( bsp ), 0, sizeof( *bsp ) );
	FILE * f = fopen( filename.c_str(), "rb" );
	if( !f ) { fprintf( stderr, "bsp: cannot open %s\n", filename.c_str() ); abort(); }
	fseek( f, 0, SEEK_END );
	const long len = ftell( f );
	fseek( f, 0, SEEK_SET );
	if( len < 8 + 8 * 17 ) { fprintf( stderr, "bsp: %s is shorter than an IBSP header\n", filename.c_str() ); abort(); }
	bsp->contents = new char[ len ];
	if( fread( bsp->contents, 1, ( size_t ) len, f ) != ( size_t ) len ) { fprintf( stderr, "bsp: short read on %s\n", filename.c_str() ); abort(); }
	fclose( f );

	/* directory slots: 1 textures, 2 planes, 3 nodes, 4 leaves, 6 leafbrushes, 8 brushes, 9 brushsides (bsp.h:7-27) */
	point_lump( bsp, len, 1, bsp->num_textures, bsp->textures, "textures" );
	point_lump( bsp, len, 2, bsp->num_planes, bsp->planes, "planes" );
	point_lump( bsp, len, 3, bsp->num_nodes, bsp->nodes, "nodes" );
	point_lump( bsp, len, 4, bsp->num_leaves, bsp->leaves, "leaves" );
	point_lump( bsp, len, 6, bsp->num_leaf_brushes, bsp->leaf_brushes, "leafbrushes" );
	point_lump( bsp, len, 8, bsp->num_brushes, bsp->brushes, "brushes" );
	point_lump( bsp, len, 9, bsp->num_brush_sides, bsp->brush_sides, "brushsides" );

	bspgpu_map_desc d;
	d.textures = bsp->textures;         d.num_textures = bsp->num_textures;
	d.planes = bsp->planes;             d.num_planes = bsp->num_planes;
	d.nodes = bsp->nodes;               d.num_nodes = bsp->num_nodes;
	d.leaves = bsp->leaves;             d.num_leaves = bsp->num_leaves;
	d.leaf_brushes = bsp->leaf_brushes; d.num_leaf_brushes = bsp->num_leaf_brushes;
	d.brushes = bsp->brushes;           d.num_brushes = bsp->num_brushes;
	d.brush_sides = bsp->brush_sides;   d.num_brush_sides = bsp->num_brush_sides;
	check( bspgpu_create( &bsp->gpu, &d, device ), "bspgpu_create" );
}

void bsp_init( BSP * bsp, std::string filename ) {
	bsp_init( bsp, filename, -1 );
}

void bsp_destroy( BSP * bsp ) {
	bspgpu_destroy( bsp->gpu );
	bsp->gpu = nullptr;
	delete[] bsp->contents;
	bsp->contents = nullptr;
}

void BSP::trace_segs( const Segment * segs, size_t n, TraceResult * out ) const {
	static_assert( sizeof( Segment ) == 24, "Segment is two packed vec3" );
	check( bspgpu_trace( gpu, segs, n, out, BSPGPU_SEGS_PACKED24 ), "bspgpu_trace" );
}

void BSP::trace_segs( const Segment * segs, size_t n, Intersection * is, bool * hit ) const {
	std::vector< TraceResult > res( n );
	std::vector< bspgpu_pos > pos( n );
	check( bspgpu_trace_pos( gpu, segs, n, res.data(), pos.data(), BSPGPU_SEGS_PACKED24 ), "bspgpu_trace_pos" );
	for( size_t i = 0; i < n; i++ ) {
#ifdef BSP_HOST_FILL_PLANE
		is[ i ].plane = res[ i ].plane >= 0 ? &planes[ res[ i ].plane ] : nullptr;
#else
		is[ i ].plane = nullptr;                                   /* as the reference: never assigned */
#endif
		is[ i ].pos = glm::vec3( pos[ i ].pos[ 0 ], pos[ i ].pos[ 1 ], pos[ i ].pos[ 2 ] );
		is[ i ].t = res[ i ].t;
		if( hit ) hit[ i ] = ( res[ i ].flags & BSPGPU_HIT ) != 0;
	}
}

bool BSP::trace_seg( const glm::vec3 & start, const glm::vec3 & end, Intersection & is ) const {
	Segment s;
	s.start = start; s.end = end;
	bool hit = false;
	trace_segs( &s, 1, &is, &hit );
	return hit;
}

void BSP::trace_segs_device( const bspgpu_segment * d_segs, size_t n, TraceResult * d_out, void * cuda_stream ) const {
	check( bspgpu_set_stream( gpu, cuda_stream ), "bspgpu_set_stream" );
	check( bspgpu_trace( gpu, d_segs, n, d_out, BSPGPU_SEGS_DEVICE | BSPGPU_OUT_DEVICE | BSPGPU_ASYNC ), "bspgpu_trace" );
	check( bspgpu_reset_stream( gpu ), "bspgpu_reset_stream" );
}

void BSP::positions_to_leaves( const glm::vec3 * pos, size_t n, int32_t * leaf ) const {
	check( bspgpu_leaf( gpu, reinterpret_cast< const float * >( pos ), 3, n, leaf, 0 ), "bspgpu_leaf" );
}

BSP_Leaf & BSP::position_to_leaf( const glm::vec3 & pos ) const {
	int32_t leaf = 0;
	positions_to_leaves( &pos, 1, &leaf );
	return leaves[ leaf ];
}

uint64_t BSP::last_hit_count() const {
	uint64_t hits = 0;
	check( bspgpu_hit_count( gpu, &hits ), "bspgpu_hit_count" );
	return hits;
}
