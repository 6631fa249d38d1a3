/*
 * bsp_host.cc -- see bsp_host.h.  Loads the file the way the reference does (one blob, typed
 * pointers into it, bsp.cc:58-86), hands the seven trace lumps to bspgpu_create, and forwards
 * every query to the C ABI.
 */
#include "bsp_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
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

	/* the eleven lumps of bsp_init (bsp.cc:74-84; directory slots per bsp.h:7-27) + load_vis (bsp.cc:105-114) */
	point_lump( bsp, len, 1, bsp->num_textures, bsp->textures, "textures" );
	point_lump( bsp, len, 2, bsp->num_planes, bsp->planes, "planes" );
	point_lump( bsp, len, 3, bsp->num_nodes, bsp->nodes, "nodes" );
	point_lump( bsp, len, 4, bsp->num_leaves, bsp->leaves, "leaves" );
	point_lump( bsp, len, 5, bsp->num_leaf_faces, bsp->leaf_faces, "leaffaces" );
	point_lump( bsp, len, 6, bsp->num_leaf_brushes, bsp->leaf_brushes, "leafbrushes" );
	point_lump( bsp, len, 8, bsp->num_brushes, bsp->brushes, "brushes" );
	point_lump( bsp, len, 9, bsp->num_brush_sides, bsp->brush_sides, "brushsides" );
	point_lump( bsp, len, 10, bsp->num_vertices, bsp->vertices, "vertices" );
	point_lump( bsp, len, 11, bsp->num_mesh_verts, bsp->mesh_verts, "meshverts" );
	point_lump( bsp, len, 13, bsp->num_faces, bsp->faces, "faces" );
	const Dirent & vd = reinterpret_cast< const Dirent * >( bsp->contents + 8 )[ 16 ];
	if( ( uint64_t ) vd.off + vd.len > ( uint64_t ) len || vd.len < 8 ) { fprintf( stderr, "bsp: visibility lump is malformed (off %u len %u)\n", vd.off, vd.len ); abort(); }
	bsp->vis = reinterpret_cast< BSP_Vis * >( bsp->contents + vd.off );
	if( vd.len != 8 + ( uint64_t ) bsp->vis->num_clusters * bsp->vis->cluster_size ) { fprintf( stderr, "bsp: visibility lump length does not match its header (bsp.cc:111)\n" ); abort(); }

	bspgpu_map_desc d;
	d.textures = bsp->textures;         d.num_textures = bsp->num_textures;
	d.planes = bsp->planes;             d.num_planes = bsp->num_planes;
	d.nodes = bsp->nodes;               d.num_nodes = bsp->num_nodes;
	d.leaves = bsp->leaves;             d.num_leaves = bsp->num_leaves;
	d.leaf_brushes = bsp->leaf_brushes; d.num_leaf_brushes = bsp->num_leaf_brushes;
	d.brushes = bsp->brushes;           d.num_brushes = bsp->num_brushes;
	d.brush_sides = bsp->brush_sides;   d.num_brush_sides = bsp->num_brush_sides;
	if( device == BSP_ALL_GPUS ) {
		check( bspgpu_multi_create( &bsp->multi, &d, nullptr, 0 ), "bspgpu_multi_create" );
		for( int r = 0; r < bspgpu_multi_device_count( bsp->multi ); r++ )
			check( bspgpu_set_vis( bspgpu_multi_handle( bsp->multi, r ), bsp->vis, vd.len ), "bspgpu_set_vis" );
		return;
	}
	check( bspgpu_create( &bsp->gpu, &d, device ), "bspgpu_create" );
	check( bspgpu_set_vis( bsp->gpu, bsp->vis, vd.len ), "bspgpu_set_vis" );
}

/* point queries and device-pointer calls go to one replica */
static bspgpu_t * one_gpu( const BSP * bsp ) { return bsp->gpu ? bsp->gpu : bspgpu_multi_handle( bsp->multi, 0 ); }

void bsp_init( BSP * bsp, std::string filename ) {
	bsp_init( bsp, filename, -1 );
}

void bsp_destroy( BSP * bsp ) {
	bspgpu_destroy( bsp->gpu );
	bspgpu_multi_destroy( bsp->multi );
	bsp->gpu = nullptr; bsp->multi = nullptr;
	delete[] bsp->contents;
	bsp->contents = nullptr;
}

void BSP::trace_segs( const Segment * segs, size_t n, TraceResult * out ) const {
	static_assert( sizeof( Segment ) == 24, "Segment is two packed vec3" );
	if( multi ) check( bspgpu_multi_trace( multi, segs, n, out, nullptr, BSPGPU_SEGS_PACKED24 ), "bspgpu_multi_trace" );
	else check( bspgpu_trace( gpu, segs, n, out, BSPGPU_SEGS_PACKED24 ), "bspgpu_trace" );
}

void BSP::trace_segs( const Segment * segs, size_t n, Intersection * is, bool * hit ) const {
	/* What the reference returns per segment is `hit` and `is.t`; `is.pos` is a function of those and of the caller's own
	 * segment (bsp.cc:266).  So only 4 bytes per segment come back from the GPU (BSPGPU_OUT_COMPACT4: t, or BSPGPU_T_MISS) --
	 * 8 with BSP_HOST_FILL_PLANE, for the plane index -- and pos is evaluated here, with the reference's own expression and one
	 * rounding per operation (this file is built with -ffp-contract=off): bit-identical to the 16-byte position records.
	 * A batch of one (trace_seg) or a handful: no heap traffic. */
#ifdef BSP_HOST_FILL_PLANE
	typedef bspgpu_result8 Wire;
	const unsigned layout = BSPGPU_OUT_COMPACT8;
#else
	typedef float Wire;
	const unsigned layout = BSPGPU_OUT_COMPACT4;
#endif
	Wire small[ 16 ];
	std::vector< Wire > big;
	Wire * w = small;
	if( n > 16 ) { big.resize( n ); w = big.data(); }
	if( multi ) check( bspgpu_multi_trace( multi, segs, n, reinterpret_cast< TraceResult * >( w ), nullptr, BSPGPU_SEGS_PACKED24 | layout ), "bspgpu_multi_trace" );
	else check( bspgpu_trace( gpu, segs, n, reinterpret_cast< TraceResult * >( w ), BSPGPU_SEGS_PACKED24 | layout ), "bspgpu_trace" );
	/* unpacking is a streaming loop over 24 + 4 bytes in, 24 + 1 bytes out per segment: large batches split it over a few threads */
	auto unpack = [ & ]( size_t lo, size_t hi ) {
	for( size_t i = lo; i < hi; i++ ) {
#ifdef BSP_HOST_FILL_PLANE
		const bool h = ( w[ i ].bits & BSPGPU_HIT ) != 0;
		const float t = w[ i ].t;
		const int32_t plane = BSPGPU_RESULT8_PLANE( w[ i ].bits );
		is[ i ].plane = plane >= 0 ? &planes[ plane ] : nullptr;
#else
		const bool h = w[ i ] != BSPGPU_T_MISS;
		const float t = h ? w[ i ] : 1.0f;                         /* bis.is.t = 1.0f (bsp.cc:257) survives a miss */
		is[ i ].plane = nullptr;                                   /* as the reference: never assigned */
#endif
		const glm::vec3 & a = segs[ i ].start, & b = segs[ i ].end;
		/* is.pos = start + is.t * ( end - start ) on a hit (bsp.cc:266), the zero-initialised (0,0,0) of `bis = { }` on a miss */
		is[ i ].pos = h ? glm::vec3( a.x + t * ( b.x - a.x ), a.y + t * ( b.y - a.y ), a.z + t * ( b.z - a.z ) ) : glm::vec3( 0.0f, 0.0f, 0.0f );
		is[ i ].t = t;
		if( hit ) hit[ i ] = h;
	}
	};
	const unsigned hw = std::thread::hardware_concurrency();
	const size_t workers = n >= ( size_t ) 1 << 20 ? ( hw >= 16 ? 8 : ( hw >= 4 ? hw / 2 : 1 ) ) : 1;
	if( workers <= 1 ) { unpack( 0, n ); return; }
	std::vector< std::thread > pool;
	const size_t per = ( n + workers - 1 ) / workers;
	for( size_t k = 0; k < workers; k++ ) {
		const size_t lo = k * per, hi = lo + per < n ? lo + per : n;
		if( lo < hi ) pool.emplace_back( unpack, lo, hi );
	}
	for( std::thread & t : pool ) t.join();
}

bool BSP::trace_seg( const glm::vec3 & start, const glm::vec3 & end, Intersection & is ) const {
	Segment s;
	s.start = start; s.end = end;
	bool hit = false;
	trace_segs( &s, 1, &is, &hit );
	return hit;
}

void BSP::trace_segs_device( const bspgpu_segment * d_segs, size_t n, TraceResult * d_out, void * cuda_stream ) const {
	/* one call, its own stream argument: nothing of the handle's state changes, two threads cannot interleave a set / reset */
	check( bspgpu_trace_stream( one_gpu( this ), d_segs, n, d_out, nullptr, BSPGPU_SEGS_DEVICE | BSPGPU_OUT_DEVICE | BSPGPU_ASYNC, cuda_stream ), "bspgpu_trace_stream" );
}

void BSP::positions_to_leaves( const glm::vec3 * pos, size_t n, int32_t * leaf ) const {
	check( bspgpu_leaf( one_gpu( this ), reinterpret_cast< const float * >( pos ), 3, n, leaf, 0 ), "bspgpu_leaf" );
}

BSP_Leaf & BSP::position_to_leaf( const glm::vec3 & pos ) const {
	int32_t leaf = 0;
	positions_to_leaves( &pos, 1, &leaf );
	return leaves[ leaf ];
}

void BSP::visible_leaves( const glm::vec3 & pos, uint8_t * visible ) const {
	const float p[ 3 ] = { pos.x, pos.y, pos.z };
	check( bspgpu_visible_leaves( one_gpu( this ), p, visible, nullptr ), "bspgpu_visible_leaves" );
}

uint64_t BSP::last_hit_count() const {
	uint64_t hits = 0;
	if( multi ) check( bspgpu_multi_hit_count( multi, &hits ), "bspgpu_multi_hit_count" );   /* NCCL all-reduce of the per-GPU counters */
	else check( bspgpu_hit_count( gpu, &hits ), "bspgpu_hit_count" );
	return hits;
}
