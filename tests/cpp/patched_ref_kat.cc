/*
 * tests/cpp/patched_ref_kat.cc -- TEST DRIVER for the reference's OWN BSP class after INTEGRATION.md's patch
 * (tools/apply_integration.py) has been applied to copies of /root/reference/bsp.h and bsp.cc.
 *
 * Everything else -- the class declaration with all of its members (bsp.h:156-211), bsp_init, load_lump, load_vis,
 * position_to_leaf, the plugin glue -- is the reference's unmodified text; bsp_renderer.cc is compiled next to it
 * unmodified to show that code written against the reference's BSP still builds.  Runs the KAT-0 / KAT-1 vectors
 * (captured from the unmodified reference) through bsp_init + BSP::trace_seg + BSP::trace_segs.
 * usage: patched_ref_kat kat0.bsp kat1.bsp      (prints OK, or the mismatches)
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include <GL/glew.h>
#include <glm/glm.hpp>

#include "intrinsics.h"
#include "bsp.h"
#include "bsp_renderer.h"
#include "immediate.h"
#include "memory_arena.h"

struct Kat { float s[ 3 ], e[ 3 ]; int hit; uint32_t t_bits; };

static const Kat kat0[] = {
	{ { 64, 0, 0 }, { -64, 0, 0 }, 1, 0x3EC00000u }, { { -64, 1, 2 }, { 64, 3, 4 }, 1, 0x3EC00000u },
	{ { 64, 0, 40 }, { -64, 0, 40 }, 0, 0x3F800000u }, { { 1, 0, 0 }, { 64, 0, 0 }, 0, 0x3F800000u },
	{ { 1, 0, 0 }, { -64, 0, 0 }, 0, 0x3F800000u }, { { 64, 0, 0 }, { 4, 0, 0 }, 1, 0x3F4CCCCDu },
	{ { 16, 0, 0 }, { -64, 0, 0 }, 0, 0x3F800000u }, { { 64, 0, 0 }, { 16, 0, 0 }, 1, 0x3F800000u },
	{ { 50, 40, 30 }, { -50, -40, -30 }, 1, 0x3EAE147Bu }, { { 0, 64, 0 }, { 0, -64, 0 }, 1, 0x3EC00000u },
};
static const Kat kat1[] = {
	{ { 64, 0, 0 }, { -64, 0, 0 }, 1, 0x3F000000u }, { { -64, 0, 0 }, { 64, 0, 0 }, 1, 0x3EC00000u },
	{ { 180, 0, 0 }, { 60, 0, 0 }, 1, 0x3ED55555u }, { { 180, 12, 0 }, { 60, 12, 0 }, 1, 0x3F000000u },
	{ { 190, 0, 0 }, { 260, 0, 0 }, 1, 0x3F124925u }, { { 190, 0, 0 }, { 225, 0, 0 }, 0, 0x3F800000u },
	{ { -10, 45, 0 }, { 207.5f, 210, 0 }, 1, 0x3F22E8BAu }, { { 200, 157, 3 }, { 100, 157, -5 }, 1, 0x3F11EB85u },
	{ { 100, 158, 0 }, { 160, 156, 1 }, 1, 0x3EF564F6u }, { { 110, 40, 0 }, { 110, -40, 0 }, 1, 0x3E99999Au },
	{ { 10, 0, 0 }, { -64, 0, 0 }, 0, 0x3F800000u }, { { 140, 36, 0 }, { 100, -4, 0 }, 1, 0x3F000000u },
};

static uint32_t bits( float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }

static int run( const char * path, const Kat * kat, size_t n, const char * name ) {
	BSP bsp;
	bsp_init( &bsp, path );
	/* every member of the reference's class is there and loaded (bsp.h:160-193) */
	printf( "%s: %u nodes %u leaves %u leaf_faces %u vertices %u mesh_verts %u faces, vis %u clusters\n", name, bsp.num_nodes, bsp.num_leaves,
		bsp.num_leaf_faces, bsp.num_vertices, bsp.num_mesh_verts, bsp.num_faces, bsp.vis->num_clusters );
	int bad = 0;
	std::vector< Segment > segs( n );
	for( size_t i = 0; i < n; i++ ) {
		segs[ i ].start = glm::vec3( kat[ i ].s[ 0 ], kat[ i ].s[ 1 ], kat[ i ].s[ 2 ] );
		segs[ i ].end = glm::vec3( kat[ i ].e[ 0 ], kat[ i ].e[ 1 ], kat[ i ].e[ 2 ] );
		Intersection is;
		const bool hit = bsp.trace_seg( segs[ i ].start, segs[ i ].end, is );   /* the reference's call shape, bsp.cc:365 */
		if( hit != ( kat[ i ].hit != 0 ) || bits( is.t ) != kat[ i ].t_bits || is.plane != NULL ) {
			printf( "%s trace_seg %zu: hit %d t %08x (want %d %08x)\n", name, i, ( int ) hit, bits( is.t ), kat[ i ].hit, kat[ i ].t_bits );
			bad++;
		}
	}
	std::vector< TraceResult > res( n );
	bsp.trace_segs( segs.data(), n, res.data() );
	for( size_t i = 0; i < n; i++ ) {
		if( ( int ) ( res[ i ].flags & BSPGPU_HIT ) != kat[ i ].hit || bits( res[ i ].t ) != kat[ i ].t_bits ) {
			printf( "%s trace_segs %zu: flags %u t %08x (want %d %08x)\n", name, i, res[ i ].flags, bits( res[ i ].t ), kat[ i ].hit, kat[ i ].t_bits );
			bad++;
		}
	}
	/* the reference's own CPU position_to_leaf still works on the same object */
	const BSP_Leaf & leaf = bsp.position_to_leaf( glm::vec3( -5, 0, 0 ) );
	if( &leaf < bsp.leaves || &leaf >= bsp.leaves + bsp.num_leaves ) { printf( "%s: position_to_leaf out of range\n", name ); bad++; }
	bsp_destroy( &bsp );
	return bad;
}

int main( int argc, char ** argv ) {
	if( argc < 3 ) { fprintf( stderr, "usage: %s kat0.bsp kat1.bsp\n", argv[ 0 ] ); return 2; }
	int bad = run( argv[ 1 ], kat0, sizeof( kat0 ) / sizeof( kat0[ 0 ] ), "kat0" );
	bad += run( argv[ 2 ], kat1, sizeof( kat1 ) / sizeof( kat1[ 0 ] ), "kat1" );
	printf( bad ? "FAILED (%d)\n" : "OK\n", bad );
	return bad ? 1 : 0;
}

/* link stubs: the GL debug drawing the reference's (now dead) CPU trace helpers and plugin glue call */
void immediate_init( ImmediateContext * const, ImmediateTriangle * const, const size_t ) { }
void immediate_sphere( ImmediateContext * const, const glm::vec3, const float, const glm::vec4, const u32 ) { }
void immediate_render( const ImmediateContext * const, const GLint, const GLint, const bool, const GLint, const GLint ) { }
