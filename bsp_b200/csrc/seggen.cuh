/*
 * seggen.cuh -- K0: counter-based synthetic segment generator, bit-identical to
 * tools/seggen.py (same integer hash, same fp32 operations in the same order; build with
 * -fmad=false -prec-div=true -prec-sqrt=true).  Lets the 1e9-segment configurations be
 * created in HBM without a 32 GB host->device copy, each rank generating only its shard
 * from (seed, global index).
 */
#ifndef BSPGPU_SEGGEN_CUH
#define BSPGPU_SEGGEN_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace bspk {

struct GenParams {
	uint32_t kind;                 /* 0 uniform, 1 long, 2 camera */
	uint64_t key;                  /* fin(seed + GOLDEN) */
	float lo[ 3 ], ext[ 3 ];       /* ext = hi - lo in fp32 */
	float len_lo, len_ext;
	const float * views;           /* device, 12 floats per view */
	uint32_t width, height;
	float inv2w, inv2h, tanx, tany, far_dist;
	uint64_t first, count;
	float4 * out;
};

static const uint64_t kGolden = 0x9E3779B97F4A7C15ull;
static const uint32_t kSlots = 32;

__host__ __device__ __forceinline__ uint64_t sm64_fin( uint64_t z ) {
	z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
	z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
	return z ^ ( z >> 31 );
}

__device__ __forceinline__ float unit_float( uint64_t key, uint64_t idx, uint32_t slot ) {
	const uint64_t u = sm64_fin( key + ( idx * kSlots + slot + 1 ) * kGolden );
	return ( float ) ( uint32_t ) ( u >> 40 ) * 5.9604644775390625e-8f;   /* 2^-24 */
}

__global__ void generate_kernel( const GenParams g ) {
	for( uint64_t k = blockIdx.x * ( uint64_t ) blockDim.x + threadIdx.x; k < g.count; k += ( uint64_t ) gridDim.x * blockDim.x ) {
		const uint64_t idx = g.first + k;
		float s[ 3 ], e[ 3 ];
		if( g.kind == 2 ) {
			const uint64_t per = ( uint64_t ) g.width * g.height;
			const uint64_t v = idx / per, pix = idx % per;
			const float px = ( float ) ( uint32_t ) ( pix % g.width ), py = ( float ) ( uint32_t ) ( pix / g.width );
			const float sx = ( ( px + 0.5f ) * g.inv2w - 1.0f ) * g.tanx;
			const float sy = ( 1.0f - ( py + 0.5f ) * g.inv2h ) * g.tany;
			const float * view = g.views + v * 12;
			for( int a = 0; a < 3; a++ ) {
				s[ a ] = view[ a ];
				const float ray = ( view[ 3 + a ] + view[ 6 + a ] * sx ) + view[ 9 + a ] * sy;
				e[ a ] = s[ a ] + ray * g.far_dist;
			}
		}
		else {
			for( int a = 0; a < 3; a++ ) s[ a ] = g.lo[ a ] + g.ext[ a ] * unit_float( g.key, idx, a );
			if( g.kind == 0 ) {
				for( int a = 0; a < 3; a++ ) e[ a ] = g.lo[ a ] + g.ext[ a ] * unit_float( g.key, idx, 3 + a );
			}
			else {
				float vx = 0.0f, vy = 0.0f, vz = 0.0f, q = 0.0f;
				bool done = false;
				for( int j = 0; j < 8; j++ ) {
					const float cx = 2.0f * unit_float( g.key, idx, 3 + 3 * j ) - 1.0f;
					const float cy = 2.0f * unit_float( g.key, idx, 4 + 3 * j ) - 1.0f;
					const float cz = 2.0f * unit_float( g.key, idx, 5 + 3 * j ) - 1.0f;
					const float cq = ( cx * cx + cy * cy ) + cz * cz;
					if( !done ) {
						vx = cx; vy = cy; vz = cz; q = cq;
						done = ( cq <= 1.0f ) && ( cq >= 0.0009765625f );
					}
				}
				if( q < 0.0009765625f ) { vx = 1.0f; vy = 0.0f; vz = 0.0f; q = 1.0f; }
				const float sq = sqrtf( q );
				const float ln = g.len_lo + g.len_ext * unit_float( g.key, idx, 27 );
				e[ 0 ] = s[ 0 ] + ( vx / sq ) * ln;
				e[ 1 ] = s[ 1 ] + ( vy / sq ) * ln;
				e[ 2 ] = s[ 2 ] + ( vz / sq ) * ln;
			}
		}
		g.out[ 2 * k ] = make_float4( s[ 0 ], s[ 1 ], s[ 2 ], 0.0f );
		g.out[ 2 * k + 1 ] = make_float4( e[ 0 ], e[ 1 ], e[ 2 ], 0.0f );
	}
}

} // namespace bspk

#endif
