/*
 * tests/cpp/depth_frame.cc -- a game_frame-shaped C++ consumer of the batched trace.
 *
 * The reference issues ONE debug segment per frame from game_frame (bsp.cc:357-370: press `t`, trace from the camera
 * along its forward vector).  This program does what a caller with a GPU behind trace_segs would do instead: per frame,
 * every pixel of a pinhole camera gets its own segment eye -> eye + far * ray (the C4 workload of BASELINE.json), the
 * whole frame goes through ONE BSP::trace_segs call, and the hit fractions become a depth image -- written as PNG with
 * the reference's own vendored stb_image_write.h (included from $BSP_REF_DIR at build time, never copied; without it
 * the image is a binary PGM).
 *
 *   depth_frame map.bsp W H out.{png,pgm} ex ey ez fx fy fz rx ry rz ux uy uz [tan_half far frames]
 *
 * prints one line: hits, an FNV-1a hash of the frame's t bits (compared with tools/render_depth.py by the tests), and
 * the wall time of the trace_segs call per frame.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>

#include "../../bsp_b200/csrc/host/bsp_host.h"

#ifdef BSP_HAVE_STB
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image_write.h"
#endif

int main( int argc, char ** argv ) {
	if( argc < 17 ) { fprintf( stderr, "usage: %s map.bsp W H out ex ey ez fx fy fz rx ry rz ux uy uz [tan_half far frames]\n", argv[ 0 ] ); return 2; }
	const int width = atoi( argv[ 2 ] ), height = atoi( argv[ 3 ] );
	float view[ 12 ];
	for( int i = 0; i < 12; i++ ) view[ i ] = strtof( argv[ 5 + i ], nullptr );
	const float tan_half = argc > 17 ? strtof( argv[ 17 ], nullptr ) : 1.0f;
	const float far_dist = argc > 18 ? strtof( argv[ 18 ], nullptr ) : 32768.0f;
	const int frames = argc > 19 ? atoi( argv[ 19 ] ) : 1;
	if( width < 1 || height < 1 || frames < 1 ) return 2;

	BSP bsp;
	bsp_init( &bsp, argv[ 1 ] );

	/* the frame's segments: the same arithmetic, in the same order, as tools/seggen.py camera() and the device generator */
	const size_t n = ( size_t ) width * height;
	std::vector< Segment > segs( n );
	const float inv2w = ( float ) ( 2.0 / ( double ) width ), inv2h = ( float ) ( 2.0 / ( double ) height );
	const float tanx = tan_half, tany = tan_half * ( float ) height / ( float ) width;
	for( int py = 0; py < height; py++ ) for( int px = 0; px < width; px++ ) {
		const float sx = ( ( ( float ) px + 0.5f ) * inv2w - 1.0f ) * tanx;
		const float sy = ( 1.0f - ( ( float ) py + 0.5f ) * inv2h ) * tany;
		float e[ 3 ];
		for( int a = 0; a < 3; a++ ) {
			const float ray = ( view[ 3 + a ] + view[ 6 + a ] * sx ) + view[ 9 + a ] * sy;
			e[ a ] = view[ a ] + ray * far_dist;
		}
		Segment & s = segs[ ( size_t ) py * width + px ];
		s.start = glm::vec3( view[ 0 ], view[ 1 ], view[ 2 ] );
		s.end = glm::vec3( e[ 0 ], e[ 1 ], e[ 2 ] );
	}

	std::vector< TraceResult > res( n );
	double best_ms = 1e30;
	for( int f = 0; f < frames; f++ ) {       /* one call per frame, like game_frame */
		const auto t0 = std::chrono::steady_clock::now();
		bsp.trace_segs( segs.data(), n, res.data() );
		const double ms = std::chrono::duration< double, std::milli >( std::chrono::steady_clock::now() - t0 ).count();
		if( ms < best_ms ) best_ms = ms;
	}

	uint64_t fnv = 1469598103934665603ull, hits = 0;
	std::vector< unsigned char > img( n );
	for( size_t i = 0; i < n; i++ ) {
		uint32_t tb; memcpy( &tb, &res[ i ].t, 4 );
		for( int k = 0; k < 4; k++ ) { fnv ^= ( tb >> ( 8 * k ) ) & 0xffu; fnv *= 1099511628211ull; }
		const bool hit = ( res[ i ].flags & BSPGPU_HIT ) != 0;
		hits += hit;
		float t = res[ i ].t < 0.0f ? 0.0f : ( res[ i ].t > 1.0f ? 1.0f : res[ i ].t );
		img[ i ] = hit ? ( unsigned char ) ( 255.0f * ( 1.0f - __builtin_sqrtf( t ) ) ) : 0;
	}
	const std::string out = argv[ 4 ];
	bool wrote_png = false;
#ifdef BSP_HAVE_STB
	if( out.size() > 4 && out.substr( out.size() - 4 ) == ".png" ) wrote_png = stbi_write_png( out.c_str(), width, height, 1, img.data(), width ) != 0;
#endif
	if( !wrote_png ) {
		FILE * f = fopen( out.c_str(), "wb" );
		if( !f ) { fprintf( stderr, "cannot write %s\n", out.c_str() ); return 1; }
		fprintf( f, "P5\n%d %d\n255\n", width, height );
		fwrite( img.data(), 1, n, f );
		fclose( f );
	}
	printf( "frame %dx%d hits=%llu fnv=%016llx trace_segs_ms=%.3f format=%s hit_count=%llu\n", width, height, ( unsigned long long ) hits,
		( unsigned long long ) fnv, best_ms, wrote_png ? "png" : "pgm", ( unsigned long long ) bsp.last_hit_count() );
	bsp_destroy( &bsp );
	return 0;
}
