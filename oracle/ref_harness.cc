/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded
 * by or called from the product (bsp_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * A C-ABI driver around the UNMODIFIED reference translation units, compiled
 * where they lie (/root/reference/bsp.cc, memory_arena.cc, work_queue.cc) by
 * oracle/Makefile into oracle/_ref/libbspref.so:
 *   - bspref_trace      loops BSP::trace_seg (bsp.cc:255-271) over a segment array,
 *   - bspref_trace_mt   chunks the same loop into <=255 jobs on the reference's own
 *                       work_queue (work_queue.cc:50,77,91), caller thread helping,
 *   - bspref_leaf       loops BSP::position_to_leaf (bsp.cc:273-286).
 * The five link stubs at the bottom replace the GL debug drawing the reference
 * calls from inside the trace (bsp.cc:159,162,168,242): unmodified, it tessellates
 * a sphere per plane crossing into a static 512000-triangle buffer and traps when
 * that fills (immediate.cc:18).  The spheres are write-only, results are unaffected.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <chrono>
#include <string>

#include <GL/glew.h>
#include <glm/glm.hpp>

#include "intrinsics.h"
#include "bsp.h"
#include "bsp_renderer.h"
#include "immediate.h"
#include "memory_arena.h"
#include "work_queue.h"

namespace {

struct SegView {
	const float * base;
	uint64_t stride;   // floats between consecutive segments
	uint64_t end_off;  // floats from a segment's start point to its end point
};

struct Outputs {
	uint8_t * hit;       // may be null
	float * t;           // may be null
	float * pos;         // 3 floats per segment, may be null
	uint8_t * plane_set; // 1 if Intersection::plane != nullptr, may be null
};

inline void run_range( const BSP * bsp, const SegView & sv, const Outputs & out, uint64_t first, uint64_t count ) {
	for( uint64_t i = first; i < first + count; i++ ) {
		const float * s = sv.base + i * sv.stride;
		const float * e = s + sv.end_off;
		const glm::vec3 start( s[ 0 ], s[ 1 ], s[ 2 ] );
		const glm::vec3 end( e[ 0 ], e[ 1 ], e[ 2 ] );
		Intersection is;
		const bool hit = bsp->trace_seg( start, end, is );
		if( out.hit ) out.hit[ i ] = hit ? 1 : 0;
		if( out.t ) out.t[ i ] = is.t;
		if( out.pos ) {
			out.pos[ 3 * i + 0 ] = is.pos.x;
			out.pos[ 3 * i + 1 ] = is.pos.y;
			out.pos[ 3 * i + 2 ] = is.pos.z;
		}
		if( out.plane_set ) out.plane_set[ i ] = is.plane != nullptr ? 1 : 0;
	}
}

struct ChunkJob {
	const BSP * bsp;
	SegView sv;
	Outputs out;
	uint64_t first;
	uint64_t count;
};

WORK_QUEUE_CALLBACK( chunk_job ) {
	( void ) arena;
	const ChunkJob * job = ( const ChunkJob * ) data;
	run_range( job->bsp, job->sv, job->out, job->first, job->count );
}

/* The reference's workers never exit (work_queue.cc:41-45), so the queue is a
 * process-lifetime singleton, initialised by the first threaded call. */
WorkQueue g_queue;
MemoryArena g_arena;
uint32_t g_workers = 0;
bool g_queue_ready = false;

double now_s() {
	return std::chrono::duration< double >( std::chrono::steady_clock::now().time_since_epoch() ).count();
}

} // namespace

extern "C" {

void * bspref_load( const char * path ) {
	BSP * bsp = new BSP();
	/* bsp_init printf()s one line per lump (bsp.cc:102,113); keep callers' stdout clean */
	fflush( stdout );
	const int saved = dup( STDOUT_FILENO );
	const int devnull = open( "/dev/null", O_WRONLY );
	if( saved >= 0 && devnull >= 0 ) dup2( devnull, STDOUT_FILENO );
	bsp_init( bsp, std::string( path ) );
	fflush( stdout );
	if( saved >= 0 ) { dup2( saved, STDOUT_FILENO ); close( saved ); }
	if( devnull >= 0 ) close( devnull );
	return bsp;
}

void bspref_free( void * h ) {
	BSP * bsp = ( BSP * ) h;
	if( !bsp ) return;
	delete[] bsp->contents; /* bsp_destroy uses scalar delete on new[] (bsp.cc:89) -- not repeated here */
	delete bsp;
}

/* out[0..6] = textures, planes, nodes, leaves, leaf_brushes, brushes, brush_sides */
void bspref_counts( void * h, uint32_t * out ) {
	const BSP * bsp = ( const BSP * ) h;
	out[ 0 ] = bsp->num_textures; out[ 1 ] = bsp->num_planes; out[ 2 ] = bsp->num_nodes;
	out[ 3 ] = bsp->num_leaves; out[ 4 ] = bsp->num_leaf_brushes; out[ 5 ] = bsp->num_brushes;
	out[ 6 ] = bsp->num_brush_sides;
}

/* sizeof() of the reference's on-disk structs, for the generator's layout test */
void bspref_sizes( uint32_t * out ) {
	out[ 0 ] = sizeof( BSP_Header ); out[ 1 ] = sizeof( BSP_Texture ); out[ 2 ] = sizeof( BSP_Plane );
	out[ 3 ] = sizeof( BSP_Node ); out[ 4 ] = sizeof( BSP_Leaf ); out[ 5 ] = sizeof( BSP_Brush );
	out[ 6 ] = sizeof( BSP_BrushSide ); out[ 7 ] = sizeof( BSP_Vertex ); out[ 8 ] = sizeof( BSP_Face );
	out[ 9 ] = sizeof( Intersection );
}

/* single-threaded loop; returns seconds spent inside the loop */
double bspref_trace( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n,
                     uint8_t * hit, float * t, float * pos, uint8_t * plane_set ) {
	const SegView sv = { segs, stride, end_off };
	const Outputs out = { hit, t, pos, plane_set };
	const double t0 = now_s();
	run_range( ( const BSP * ) h, sv, out, 0, n );
	return now_s() - t0;
}

/* work_queue-threaded loop: nthreads-1 reference workers + the calling thread.
 * returns seconds spent between the first enqueue and the end of exhaust, or <0 on misuse */
double bspref_trace_mt( void * h, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n,
                        uint8_t * hit, float * t, float * pos, uint8_t * plane_set, uint32_t nthreads ) {
	if( nthreads < 1 ) return -1.0;
	if( !g_queue_ready ) {
		g_workers = nthreads - 1;
		const size_t bytes = megabytes( ( size_t ) nthreads + 3 );
		memarena_init( &g_arena, ( u8 * ) malloc( bytes ), bytes );
		workqueue_init( &g_queue, &g_arena, g_workers );
		g_queue_ready = true;
	}
	else if( g_workers != nthreads - 1 ) {
		return -2.0; /* the reference queue cannot be resized or torn down */
	}

	/* work_queue.cc:78 asserts jobs_queued < 256, so work is handed out in rounds of <= 255 jobs, each
	 * followed by workqueue_exhaust (the caller helps, work_queue.cc:91-98).  A round holds a whole
	 * multiple of the thread count so every thread gets the same number of ~16k-segment jobs. */
	static ChunkJob jobs[ 255 ];
	if( n == 0 ) return 0.0;
	const uint64_t per_round = nthreads <= 255 ? ( 255 / nthreads ) * ( uint64_t ) nthreads : 255;
	uint64_t chunk = ( n + per_round - 1 ) / per_round;
	if( chunk > 16384 ) chunk = 16384;
	if( chunk < 1 ) chunk = 1;

	const SegView sv = { segs, stride, end_off };
	const Outputs out = { hit, t, pos, plane_set };

	const double t0 = now_s();
	uint64_t queued = 0;
	for( uint64_t first = 0; first < n; first += chunk ) {
		const uint64_t count = ( first + chunk <= n ) ? chunk : n - first;
		jobs[ queued ] = { ( const BSP * ) h, sv, out, first, count };
		workqueue_enqueue( &g_queue, chunk_job, &jobs[ queued ] );
		queued++;
		if( queued == per_round ) { workqueue_exhaust( &g_queue ); queued = 0; }
	}
	if( queued ) workqueue_exhaust( &g_queue );
	return now_s() - t0;
}

/* BSP::position_to_leaf (bsp.cc:273-286) -> leaf index */
void bspref_leaf( void * h, const float * pts, uint64_t stride, uint64_t n, int32_t * leaf ) {
	const BSP * bsp = ( const BSP * ) h;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * p = pts + i * stride;
		const BSP_Leaf & l = bsp->position_to_leaf( glm::vec3( p[ 0 ], p[ 1 ], p[ 2 ] ) );
		leaf[ i ] = ( int32_t ) ( &l - bsp->leaves );
	}
}

} // extern "C"

/* ---- link stubs for the GL debug/viewer code bsp.cc references ---- */
void immediate_init( ImmediateContext * const, ImmediateTriangle * const, const size_t ) { }
void immediate_sphere( ImmediateContext * const, const glm::vec3, const float, const glm::vec4, const u32 ) { }
void immediate_render( const ImmediateContext * const, const GLint, const GLint, const bool, const GLint, const GLint ) { }
void bspr_init( BSPRenderer * const, MemoryArena * const, const BSP * const, const GLint, const GLint ) { }
void bspr_render( const BSPRenderer * const, const glm::vec3 & ) { }
