/*
 * bin_kernels.cuh -- N4: optional coherence pre-pass.  A counting sort of one device-resident batch of
 * segments by the (coarsened) start-grid cell of their start point: histogram -> exclusive scan -> scatter of segment
 * indices.  The trace kernel then walks the permutation (TraceParams::perm), so neighbouring lanes start in the same
 * part of the tree; every result still lands at its segment's own index.  The order inside a bin depends on atomic
 * timing -- it only decides which lane traces which segment, never a result.
 */
#ifndef BSPGPU_BIN_KERNELS_CUH
#define BSPGPU_BIN_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "map_layout.h"

namespace bspk {

struct BinParams {
	StartGridDesc grid;
	uint32_t shift[ 3 ];      /* start-grid cell coordinate >> shift = bin coordinate */
	uint32_t bdim[ 3 ];
	uint32_t num_bins;        /* <= 65536 */
	const float4 * segs;
	const float2 * segs24;
	uint32_t n;
	uint32_t * bins;          /* histogram, then exclusive offsets, then running cursors */
	uint32_t * perm;
};

/* bin of segment i: the start point's cell; anything outside the world box (or NaN) goes to bin 0 */
__device__ __forceinline__ uint32_t bin_of( const BinParams & b, uint32_t i ) {
	float sx, sy, sz;
	if( b.segs24 ) {
		const float2 a = b.segs24[ ( size_t ) i * 3 ], c = b.segs24[ ( size_t ) i * 3 + 1 ];
		sx = a.x; sy = a.y; sz = c.x;
	}
	else {
		const float4 a = b.segs[ ( size_t ) i * 2 ];
		sx = a.x; sy = a.y; sz = a.z;
	}
	const float cx = floorf( ( sx - b.grid.lo[ 0 ] ) * b.grid.inv[ 0 ] ), cy = floorf( ( sy - b.grid.lo[ 1 ] ) * b.grid.inv[ 1 ] ), cz = floorf( ( sz - b.grid.lo[ 2 ] ) * b.grid.inv[ 2 ] );
	const bool inside = cx >= 0.0f && cx < ( float ) b.grid.dim[ 0 ] && cy >= 0.0f && cy < ( float ) b.grid.dim[ 1 ] && cz >= 0.0f && cz < ( float ) b.grid.dim[ 2 ];
	if( !inside ) return 0;
	const uint32_t ix = ( uint32_t ) cx >> b.shift[ 0 ], iy = ( uint32_t ) cy >> b.shift[ 1 ], iz = ( uint32_t ) cz >> b.shift[ 2 ];
	return ( iz * b.bdim[ 1 ] + iy ) * b.bdim[ 0 ] + ix;
}

__global__ void bin_count_kernel( const BinParams b ) {
	for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x ) atomicAdd( &b.bins[ bin_of( b, i ) ], 1u );
}

/* exclusive scan of <= 65536 counters by one 1024-thread CTA: 64 counters per thread, warp shuffles across threads */
__global__ void bin_scan_kernel( const BinParams b ) {
	__shared__ uint32_t warp_sums[ 32 ];
	const uint32_t per = ( b.num_bins + 1023u ) / 1024u;
	const uint32_t lo = threadIdx.x * per, hi = min( lo + per, b.num_bins );
	uint32_t sum = 0;
	for( uint32_t k = lo; k < hi; k++ ) sum += b.bins[ k ];
	uint32_t incl = sum;
	for( int o = 1; o < 32; o <<= 1 ) { const uint32_t v = __shfl_up_sync( 0xffffffffu, incl, o ); if( ( threadIdx.x & 31u ) >= ( unsigned ) o ) incl += v; }
	if( ( threadIdx.x & 31u ) == 31u ) warp_sums[ threadIdx.x >> 5 ] = incl;
	__syncthreads();
	if( threadIdx.x < 32 ) {
		uint32_t w = warp_sums[ threadIdx.x ], wi = w;
		for( int o = 1; o < 32; o <<= 1 ) { const uint32_t v = __shfl_up_sync( 0xffffffffu, wi, o ); if( threadIdx.x >= ( unsigned ) o ) wi += v; }
		warp_sums[ threadIdx.x ] = wi - w;
	}
	__syncthreads();
	uint32_t run = warp_sums[ threadIdx.x >> 5 ] + incl - sum;
	for( uint32_t k = lo; k < hi; k++ ) { const uint32_t c = b.bins[ k ]; b.bins[ k ] = run; run += c; }
}

__global__ void bin_scatter_kernel( const BinParams b ) {
	for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x ) b.perm[ atomicAdd( &b.bins[ bin_of( b, i ) ], 1u ) ] = i;
}

} // namespace bspk

#endif
