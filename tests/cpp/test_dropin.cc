/*
 * tests/cpp/test_dropin.cc -- the reference-facing C++ interface, used the way the reference's
 * own call site uses it (game_frame, reference bsp.cc:363-370): bsp_init, trace_seg per
 * segment, plus the new batched trace_segs.  Checks the KAT-1 vectors captured from the
 * unmodified reference.  usage: test_dropin <kat1.bsp>
 */
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../../bsp_b200/csrc/host/bsp_host.h"

struct Kat { float s[ 3 ], e[ 3 ]; int hit; uint32_t t_bits; };

static const Kat kat1[] = {
	{ { 64, 0, 0 }, { -64, 0, 0 }, 1, 0x3F000000u }, { { -64, 0, 0 }, { 64, 0, 0 }, 1, 0x3EC00000u },
	{ { 180, 0, 0 }, { 60, 0, 0 }, 1, 0x3ED55555u }, { { 180, 12, 0 }, { 60, 12, 0 }, 1, 0x3F000000u },
	{ { 190, 0, 0 }, { 260, 0, 0 }, 1, 0x3F124925u }, { { 190, 0, 0 }, { 225, 0, 0 }, 0, 0x3F800000u },
	{ { -10, 45, 0 }, { 207.5f, 210, 0 }, 1, 0x3F22E8BAu }, { { 200, 157, 3 }, { 100, 157, -5 }, 1, 0x3F11EB85u },
	{ { 100, 158, 0 }, { 160, 156, 1 }, 1, 0x3EF564F6u }, { { 110, 40, 0 }, { 110, -40, 0 }, 1, 0x3E99999Au },
	{ { 10, 0, 0 }, { -64, 0, 0 }, 0, 0x3F800000u }, { { 140, 36, 0 }, { 100, -4, 0 }, 1, 0x3F000000u },
};

static uint32_t bits( float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }

int main( int argc, char ** argv ) {
	if( argc < 2 ) { fprintf( stderr, "usage: %s kat1.bsp [all]\n", argv[ 0 ] ); return 2; }
	BSP bsp;
	/* "all": the map replicated on every GPU of the box, trace_segs sharded across them (bspgpu_multi_*, NCCL hit-count reduction) */
	if( argc > 2 && !strcmp( argv[ 2 ], "all" ) ) bsp_init( &bsp, argv[ 1 ], BSP_ALL_GPUS );
	else bsp_init( &bsp, argv[ 1 ] );
	int bad = 0;
	const size_t n = sizeof( kat1 ) / sizeof( kat1[ 0 ] );

	std::vector< Segment > segs( n );
	for( size_t i = 0; i < n; i++ ) {
		segs[ i ].start = glm::vec3( kat1[ i ].s[ 0 ], kat1[ i ].s[ 1 ], kat1[ i ].s[ 2 ] );
		segs[ i ].end = glm::vec3( kat1[ i ].e[ 0 ], kat1[ i ].e[ 1 ], kat1[ i ].e[ 2 ] );
		/* the reference call shape: one segment, one Intersection */
		Intersection is;
		const bool hit = bsp.trace_seg( segs[ i ].start, segs[ i ].end, is );
		if( hit != ( kat1[ i ].hit != 0 ) || bits( is.t ) != kat1[ i ].t_bits || is.plane != nullptr ) {
			printf( "trace_seg %zu: hit %d t %08x (want %d %08x)\n", i, ( int ) hit, bits( is.t ), kat1[ i ].hit, kat1[ i ].t_bits );
			bad++;
		}
		if( !hit && ( is.pos.x != 0.0f || is.pos.y != 0.0f || is.pos.z != 0.0f ) ) { printf( "trace_seg %zu: pos not zero on a miss\n", i ); bad++; }
	}

	std::vector< TraceResult > res( n );
	bsp.trace_segs( segs.data(), n, res.data() );
	uint64_t hits = 0;
	for( size_t i = 0; i < n; i++ ) {
		hits += kat1[ i ].hit;
		if( ( int ) ( res[ i ].flags & BSPGPU_HIT ) != kat1[ i ].hit || bits( res[ i ].t ) != kat1[ i ].t_bits ) {
			printf( "trace_segs %zu: flags %u t %08x (want %d %08x)\n", i, res[ i ].flags, bits( res[ i ].t ), kat1[ i ].hit, kat1[ i ].t_bits );
			bad++;
		}
	}
	if( bsp.last_hit_count() != hits ) { printf( "hit count %llu, want %llu\n", ( unsigned long long ) bsp.last_hit_count(), ( unsigned long long ) hits ); bad++; }
	if( res[ 0 ].plane != -1 || !( res[ 10 ].flags & BSPGPU_STARTSOLID ) ) { printf( "extension fields off\n" ); bad++; }

	/* the Intersection-shaped batch: only 4 bytes per segment come back from the GPU and `pos` is evaluated on the host with the
	 * reference's expression (bsp_host.cc) -- it must carry the very bits of the device's 16-byte position records.  The KAT
	 * segments plus 1.2 million pseudo-random ones around the brushes (beyond the 16-segment stack buffers, and enough for the threaded unpacking). */
	{
		std::vector< Segment > many( segs );
		uint64_t x = 0x9e3779b97f4a7c15ull;
		auto rnd = [ &x ]( float lo, float hi ) { x = x * 6364136223846793005ull + 1442695040888963407ull; return lo + ( hi - lo ) * ( float ) ( ( x >> 40 ) * ( 1.0 / 16777216.0 ) ); };
		for( int i = 0; i < 1200000; i++ ) {
			Segment sg;
			sg.start = glm::vec3( rnd( -80, 280 ), rnd( -40, 220 ), rnd( -30, 30 ) );
			sg.end = glm::vec3( rnd( -80, 280 ), rnd( -40, 220 ), rnd( -30, 30 ) );
			many.push_back( sg );
		}
		const size_t m = many.size();
		std::vector< Intersection > is( m );
		std::vector< uint8_t > hit8( m );
		bsp.trace_segs( many.data(), m, is.data(), reinterpret_cast< bool * >( hit8.data() ) );
		std::vector< TraceResult > r16( m );
		std::vector< bspgpu_pos > p16( m );
		const int rc = bsp.gpu ? bspgpu_trace_pos( bsp.gpu, many.data(), m, r16.data(), p16.data(), BSPGPU_SEGS_PACKED24 )
		                       : bspgpu_multi_trace( bsp.multi, many.data(), m, r16.data(), p16.data(), BSPGPU_SEGS_PACKED24 );
		if( rc != BSPGPU_OK ) { printf( "bspgpu_trace_pos: %s\n", bspgpu_last_error() ); bad++; }
		size_t nhit = 0, diff = 0;
		for( size_t i = 0; i < m; i++ ) {
			const bool h = ( r16[ i ].flags & BSPGPU_HIT ) != 0;
			nhit += h;
			if( ( hit8[ i ] != 0 ) != h || bits( is[ i ].t ) != bits( r16[ i ].t ) || bits( is[ i ].pos.x ) != bits( p16[ i ].pos[ 0 ] ) ||
			    bits( is[ i ].pos.y ) != bits( p16[ i ].pos[ 1 ] ) || bits( is[ i ].pos.z ) != bits( p16[ i ].pos[ 2 ] ) || is[ i ].plane != nullptr ) {
				if( diff++ < 5 ) printf( "Intersection %zu differs from the 16-byte records: hit %d/%d t %08x/%08x\n", i, ( int ) hit8[ i ], ( int ) h, bits( is[ i ].t ), bits( r16[ i ].t ) );
			}
		}
		if( diff ) { printf( "%zu of %zu Intersections differ\n", diff, m ); bad++; }
		if( nhit < m / 50 || nhit > m - m / 50 ) { printf( "implausible hit count %zu of %zu\n", nhit, m ); bad++; }
	}

	/* position_to_leaf (bsp.cc:273-286): x<0 -> leaf 0; x>0,y>50 -> leaf 1; x>0,y<50 -> leaf 2 */
	if( &bsp.position_to_leaf( glm::vec3( -5, 0, 0 ) ) != &bsp.leaves[ 0 ] ) { printf( "leaf 0\n" ); bad++; }
	if( &bsp.position_to_leaf( glm::vec3( 5, 60, 0 ) ) != &bsp.leaves[ 1 ] ) { printf( "leaf 1\n" ); bad++; }
	if( &bsp.position_to_leaf( glm::vec3( 5, 40, 0 ) ) != &bsp.leaves[ 2 ] ) { printf( "leaf 2\n" ); bad++; }

	bsp_destroy( &bsp );
	printf( bad ? "FAILED %d\n" : "OK\n", bad );
	return bad ? 1 : 0;
}
