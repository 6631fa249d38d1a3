/*
 * tools/microbench/microbench.cu -- measures the roofline denominators MEASURED_PEAKS.json does
 * not carry ("L2 and shared-memory peaks must be measured on the box"),
 * plus the one that actually binds the trace: how many SCATTERED 32-byte record gathers per
 * second the L1/L2 path sustains when every lane of a warp wants a different record (what a node
 * step is).  Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
 *
 *   hbm_copy          float4 copy over a buffer >> L2                       (cross-check of MEASURED_PEAKS)
 *   l2_read           float4 reads of an L2-resident buffer (32 MB), coalesced
 *   smem_read         conflict-free LDS.128 from a 64 KB tile per CTA
 *   gather_l2_indep   per lane: independent random 32-byte loads (LDG.256) from a `map_bytes` buffer
 *   gather_l2_chase   per lane: dependent chain (next index = f(loaded word)), like the tree walk
 *   gather_smem_chase same chain over a tile in shared memory (LDS.128 + LDS.64)
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#define CK( x ) do { cudaError_t e = ( x ); if( e != cudaSuccess ) { fprintf( stderr, "%s: %s\n", #x, cudaGetErrorString( e ) ); exit( 1 ); } } while( 0 )

__global__ void copy_kernel( const float4 * __restrict__ a, float4 * __restrict__ b, size_t n ) {
	for( size_t i = blockIdx.x * ( size_t ) blockDim.x + threadIdx.x; i < n; i += ( size_t ) gridDim.x * blockDim.x ) b[ i ] = a[ i ];
}

__global__ void read_kernel( const float4 * __restrict__ a, size_t n, int reps, float * sink ) {
	float acc = 0.0f;
	for( int r = 0; r < reps; r++ )
		for( size_t i = blockIdx.x * ( size_t ) blockDim.x + threadIdx.x; i < n; i += ( size_t ) gridDim.x * blockDim.x ) {
			const float4 v = a[ i ];
			acc += v.x + v.y + v.z + v.w;
		}
	if( acc == 12345.678f ) *sink = acc;
}

__global__ void smem_kernel( int iters, float * sink ) {
	extern __shared__ float4 tile[];
	const int n = 4096;   /* 64 KB */
	for( int i = threadIdx.x; i < n; i += blockDim.x ) tile[ i ] = make_float4( i, 1, 2, 3 );
	__syncthreads();
	float acc = 0.0f;
	int idx = threadIdx.x;
	for( int it = 0; it < iters; it++ ) {
		const float4 v = tile[ idx ];
		acc += v.x + v.y + v.z + v.w;
		idx = ( idx + blockDim.x ) & ( n - 1 );
	}
	if( acc == 12345.678f ) *sink = acc;
}

struct alignas( 32 ) Rec { uint32_t w[ 8 ]; };

__device__ __forceinline__ void ldg256( const Rec * p, uint32_t ( &w )[ 8 ] ) {
	asm volatile( "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"( w[ 0 ] ), "=r"( w[ 1 ] ), "=r"( w[ 2 ] ), "=r"( w[ 3 ] ), "=r"( w[ 4 ] ), "=r"( w[ 5 ] ), "=r"( w[ 6 ] ), "=r"( w[ 7 ] ) : "l"( p ) );
}

__device__ __forceinline__ uint32_t mix( uint32_t x ) {
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}

/* independent gathers: index sequence does not depend on loaded data (4 loads in flight per lane) */
__global__ void gather_indep_kernel( const Rec * recs, uint32_t mask, int iters, uint32_t * sink ) {
	uint32_t s = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ), acc = 0;
	for( int it = 0; it < iters; it += 4 ) {
		uint32_t w0[ 8 ], w1[ 8 ], w2[ 8 ], w3[ 8 ];
		const uint32_t i0 = mix( s ), i1 = mix( s + 1 ), i2 = mix( s + 2 ), i3 = mix( s + 3 );
		s += 4;
		ldg256( recs + ( i0 & mask ), w0 ); ldg256( recs + ( i1 & mask ), w1 );
		ldg256( recs + ( i2 & mask ), w2 ); ldg256( recs + ( i3 & mask ), w3 );
		acc += w0[ 0 ] + w1[ 1 ] + w2[ 2 ] + w3[ 3 ];
	}
	if( acc == 0x12345678u ) *sink = acc;
}

/* dependent chain: the next record index comes out of the loaded record (one load in flight per lane) */
__global__ void gather_chase_kernel( const Rec * recs, uint32_t mask, int iters, uint32_t * sink ) {
	uint32_t idx = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ) & mask, acc = 0;
	for( int it = 0; it < iters; it++ ) {
		uint32_t w[ 8 ];
		ldg256( recs + idx, w );
		acc += w[ 1 ];
		idx = ( w[ 0 ] + acc ) & mask;
	}
	if( acc == 0x12345678u ) *sink = acc;
}

__global__ void gather_smem_chase_kernel( const Rec * recs, uint32_t nrec, int iters, uint32_t * sink ) {
	extern __shared__ Rec stile[];
	for( uint32_t i = threadIdx.x; i < nrec; i += blockDim.x ) stile[ i ] = recs[ i ];
	__syncthreads();
	const uint32_t mask = nrec - 1;
	uint32_t idx = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ) & mask, acc = 0;
	for( int it = 0; it < iters; it++ ) {
		const uint4 a = *reinterpret_cast< const uint4 * >( &stile[ idx ].w[ 0 ] );
		const uint2 b = *reinterpret_cast< const uint2 * >( &stile[ idx ].w[ 4 ] );
		acc += a.y + b.x;
		idx = ( a.x + acc ) & mask;
	}
	if( acc == 0x12345678u ) *sink = acc;
}

template< typename F >
static float time_ms( F launch, int reps = 5 ) {
	cudaEvent_t e0, e1;
	CK( cudaEventCreate( &e0 ) ); CK( cudaEventCreate( &e1 ) );
	launch(); launch();
	CK( cudaDeviceSynchronize() );
	float best = 1e30f;
	for( int r = 0; r < reps; r++ ) {
		CK( cudaEventRecord( e0 ) );
		launch();
		CK( cudaEventRecord( e1 ) );
		CK( cudaEventSynchronize( e1 ) );
		float ms; CK( cudaEventElapsedTime( &ms, e0, e1 ) );
		if( ms < best ) best = ms;
	}
	CK( cudaGetLastError() );
	return best;
}

int main( int argc, char ** argv ) {
	size_t map_bytes = argc > 1 ? ( size_t ) atoll( argv[ 1 ] ) : ( size_t ) 2 << 20;   /* gather buffer, rounded to a power of two */
	cudaDeviceProp prop; CK( cudaGetDeviceProperties( &prop, 0 ) );
	const int sms = prop.multiProcessorCount;
	float * fsink; uint32_t * usink;
	CK( cudaMalloc( &fsink, 4 ) ); CK( cudaMalloc( &usink, 4 ) );
	printf( "{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms );

	{   /* HBM copy, 2 x 2 GiB */
		const size_t n = ( size_t ) 2 << 30 >> 4;
		float4 * a, * b; CK( cudaMalloc( &a, n * 16 ) ); CK( cudaMalloc( &b, n * 16 ) );
		CK( cudaMemset( a, 1, n * 16 ) );
		const float ms = time_ms( [&] { copy_kernel<<< sms * 16, 512 >>>( a, b, n ); } );
		printf( ", \"hbm_copy_gbs\": %.1f", 2.0 * n * 16 / ms / 1e6 );
		CK( cudaFree( a ) ); CK( cudaFree( b ) );
	}
	{   /* L2-resident coalesced reads: 32 MB, 20 passes */
		const size_t n = ( size_t ) 32 << 20 >> 4;
		float4 * a; CK( cudaMalloc( &a, n * 16 ) ); CK( cudaMemset( a, 1, n * 16 ) );
		const int reps = 20;
		const float ms = time_ms( [&] { read_kernel<<< sms * 8, 512 >>>( a, n, reps, fsink ); } );
		printf( ", \"l2_read_gbs\": %.1f", ( double ) reps * n * 16 / ms / 1e6 );
		CK( cudaFree( a ) );
	}
	{   /* shared memory, conflict-free LDS.128 */
		const int iters = 20000, threads = 1024;
		CK( cudaFuncSetAttribute( smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 ) );
		const float ms = time_ms( [&] { smem_kernel<<< sms, threads, 65536 >>>( iters, fsink ); } );
		printf( ", \"smem_read_gbs\": %.1f", ( double ) sms * threads * iters * 16.0 / ms / 1e6 );
	}
	{   /* scattered 32-byte gathers over a map-sized buffer */
		uint32_t nrec = 1; while( ( size_t ) nrec * 2 * sizeof( Rec ) <= map_bytes ) nrec *= 2;
		std::vector< Rec > h( nrec );
		uint32_t s = 12345;
		for( uint32_t i = 0; i < nrec; i++ ) for( int k = 0; k < 8; k++ ) { s = s * 1664525u + 1013904223u; h[ i ].w[ k ] = s; }
		Rec * d; CK( cudaMalloc( &d, ( size_t ) nrec * sizeof( Rec ) ) );
		CK( cudaMemcpy( d, h.data(), ( size_t ) nrec * sizeof( Rec ), cudaMemcpyHostToDevice ) );
		printf( ", \"gather_buffer_bytes\": %zu", ( size_t ) nrec * sizeof( Rec ) );
		const int iters = 2048;
		for( int tpb : { 256 } ) {
			for( int ctas : { 4, 5, 8 } ) {
				const double total = ( double ) sms * ctas * tpb * iters;
				float ms = time_ms( [&] { gather_indep_kernel<<< sms * ctas, tpb >>>( d, nrec - 1, iters, usink ); } );
				printf( ", \"gather_indep_%dx%d_per_s\": %.4g", ctas, tpb, total / ms * 1e3 );
				ms = time_ms( [&] { gather_chase_kernel<<< sms * ctas, tpb >>>( d, nrec - 1, iters, usink ); } );
				printf( ", \"gather_chase_%dx%d_per_s\": %.4g", ctas, tpb, total / ms * 1e3 );
			}
		}
		/* shared-memory chain over a 128 KB tile (4096 records), 1024 threads per SM */
		const uint32_t tile_rec = 4096;
		CK( cudaFuncSetAttribute( gather_smem_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) ( tile_rec * sizeof( Rec ) ) ) );
		const float ms = time_ms( [&] { gather_smem_chase_kernel<<< sms, 1024, tile_rec * sizeof( Rec ) >>>( d, tile_rec, iters, usink ); } );
		printf( ", \"gather_smem_chase_1x1024_per_s\": %.4g", ( double ) sms * 1024 * iters / ms * 1e3 );
		CK( cudaFree( d ) );
	}
	printf( "}\n" );
	return 0;
}
