/*
 * tools/microbench/microbench.cu -- measures the roofline denominators MEASURED_PEAKS.json does
 * not carry ("L2 and shared-memory peaks must be measured on the box").
 * Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
 *
 *   hbm_copy            float4 copy over a buffer >> L2                     (cross-check of MEASURED_PEAKS)
 *   l2_read_cg_<MB>     float4 reads with ld.global.cg (L1 BYPASSED: every byte comes over the L2 -> SM
 *                       crossbar) of an L2-resident buffer of 16..96 MB, many passes.  Round 1 read a 32 MB
 *                       buffer with plain loads at identical addresses per SM -- that measured L1, not L2.
 *   l2_read_ca_32       the round-1 measurement, kept for comparison (plain ld.global, L1 allocating)
 *   smem_read           conflict-free LDS.128 from a 64 KB tile per CTA (the pipe's 128 B/clk ceiling)
 *   smem_rec32_*        the trace's own shared-memory pattern: every lane reads a DIFFERENT random 32-byte
 *                       record as LDS.128 + LDS.64 (round-1 NodeRec layout: 16-byte plane + 8 bytes of children,
 *                       32-byte stride), independent addresses, 4 in flight per lane
 *   smem_rec16_*        same for dense 16-byte records (one LDS.128 per lane): the round-2 node / side layout
 *   gather_*            scattered 32-byte (LDG.256) / 16-byte (LDG.128) record gathers through the read-only
 *                       path over buffers that sit in L1 (64 KB), L1+L2 (2 MB) and L2 (32 MB)
 * Bandwidths of the record patterns count the bytes the lane asked for (32 / 24 / 16 per record).
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

__global__ void read_ca_kernel( const float4 * __restrict__ a, size_t n, int reps, float * sink ) {
	float acc = 0.0f;
	for( int r = 0; r < reps; r++ )
		for( size_t i = blockIdx.x * ( size_t ) blockDim.x + threadIdx.x; i < n; i += ( size_t ) gridDim.x * blockDim.x ) {
			const float4 v = a[ i ];
			acc += v.x + v.y + v.z + v.w;
		}
	if( acc == 12345.678f ) *sink = acc;
}

__device__ __forceinline__ float4 ld_cg( const float4 * p ) {
	float4 v;
	asm volatile( "ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( v.x ), "=f"( v.y ), "=f"( v.z ), "=f"( v.w ) : "l"( p ) );
	return v;
}

/* L1-bypassing reads; the block -> address mapping rotates every pass so that no SM re-reads what it read last */
__global__ void read_cg_kernel( const float4 * __restrict__ a, size_t n, int reps, float * sink ) {
	float acc = 0.0f;
	const size_t stride = ( size_t ) gridDim.x * blockDim.x;
	for( int r = 0; r < reps; r++ ) {
		const size_t rot = ( ( size_t ) r * 7919u * blockDim.x ) % n;
		size_t i = blockIdx.x * ( size_t ) blockDim.x + threadIdx.x;
		for( ; i + 3 * stride < n; i += 4 * stride ) {   /* four loads in flight per thread */
			size_t j0 = i + rot, j1 = j0 + stride, j2 = j1 + stride, j3 = j2 + stride;
			if( j0 >= n ) j0 -= n;
			if( j1 >= n ) j1 -= n;
			if( j2 >= n ) j2 -= n;
			if( j3 >= n ) j3 -= n;
			const float4 v0 = ld_cg( a + j0 ), v1 = ld_cg( a + j1 ), v2 = ld_cg( a + j2 ), v3 = ld_cg( a + j3 );
			acc += ( ( v0.x + v0.y ) + ( v0.z + v0.w ) ) + ( ( v1.x + v1.y ) + ( v1.z + v1.w ) ) + ( ( v2.x + v2.y ) + ( v2.z + v2.w ) ) + ( ( v3.x + v3.y ) + ( v3.z + v3.w ) );
		}
		for( ; i < n; i += stride ) {
			size_t j = i + rot;
			if( j >= n ) j -= n;
			const float4 v = ld_cg( a + j );
			acc += ( v.x + v.y ) + ( v.z + v.w );
		}
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
__device__ __forceinline__ uint4 ldg128( const void * p ) {
	uint4 v;
	asm volatile( "ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"( v.x ), "=r"( v.y ), "=r"( v.z ), "=r"( v.w ) : "l"( p ) );
	return v;
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
		for( int k = 0; k < 8; k++ ) acc += ( w0[ k ] ^ w1[ k ] ) + ( w2[ k ] ^ w3[ k ] );
	}
	if( acc == 0x12345678u ) *sink = acc;
}

/* same with 16-byte records (LDG.128): twice the records per line */
__global__ void gather16_indep_kernel( const uint4 * recs, uint32_t mask, int iters, uint32_t * sink ) {
	uint32_t s = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ), acc = 0;
	for( int it = 0; it < iters; it += 4 ) {
		const uint32_t i0 = mix( s ), i1 = mix( s + 1 ), i2 = mix( s + 2 ), i3 = mix( s + 3 );
		s += 4;
		const uint4 a = ldg128( recs + ( i0 & mask ) ), b = ldg128( recs + ( i1 & mask ) ), c = ldg128( recs + ( i2 & mask ) ), d = ldg128( recs + ( i3 & mask ) );
		acc += ( a.x ^ a.y ^ a.z ^ a.w ) + ( b.x ^ b.y ^ b.z ^ b.w ) + ( c.x ^ c.y ^ c.z ^ c.w ) + ( d.x ^ d.y ^ d.z ^ d.w );
	}
	if( acc == 0x12345678u ) *sink = acc;
}

/* dependent chain: the next record index comes out of the loaded record (one load in flight per lane) */
__global__ void gather_chase_kernel( const Rec * recs, uint32_t mask, int iters, uint32_t * sink ) {
	uint32_t idx = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ) & mask, acc = 0;
	for( int it = 0; it < iters; it++ ) {
		uint32_t w[ 8 ];
		ldg256( recs + idx, w );
		acc += ( w[ 1 ] ^ w[ 2 ] ^ w[ 3 ] ) + ( w[ 4 ] ^ w[ 5 ] ^ w[ 6 ] ^ w[ 7 ] );
		idx = ( w[ 0 ] + acc ) & mask;
	}
	if( acc == 0x12345678u ) *sink = acc;
}

/* the trace's shared-memory record pattern, independent addresses.  kBytes = 32: LDS.128 at rec*32 + LDS.64 at rec*32+16
 * (round-1 NodeRec);  kBytes = 16: one LDS.128 at rec*16 (dense records) */
template< int kBytes >
__global__ void smem_rec_kernel( const uint4 * src, uint32_t tile_bytes, int iters, uint32_t * sink ) {
	extern __shared__ uint4 stile4[];
	for( uint32_t i = threadIdx.x; i < tile_bytes / 16; i += blockDim.x ) stile4[ i ] = src[ i ];
	__syncthreads();
	const uint32_t base = ( uint32_t ) __cvta_generic_to_shared( stile4 );
	const uint32_t mask = tile_bytes / kBytes - 1;
	uint32_t s = mix( blockIdx.x * blockDim.x + threadIdx.x + 1 ), acc = 0;
	for( int it = 0; it < iters; it += 4 ) {
		uint32_t a[ 4 ];
#pragma unroll
		for( int k = 0; k < 4; k++ ) a[ k ] = base + ( mix( s + k ) & mask ) * kBytes;
		s += 4;
#pragma unroll
		for( int k = 0; k < 4; k++ ) {
			uint4 v;
			asm volatile( "ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"( v.x ), "=r"( v.y ), "=r"( v.z ), "=r"( v.w ) : "r"( a[ k ] ) );
			acc += ( v.x ^ v.y ) + ( v.z ^ v.w );
			if( kBytes == 32 ) {
				uint2 c;
				asm volatile( "ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"( c.x ), "=r"( c.y ) : "r"( a[ k ] + 16 ) );
				acc += c.x ^ c.y;
			}
		}
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

int main( int, char ** ) {
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
	{   /* L2: L1-bypassing reads over resident sets of 16..96 MB */
		float4 * a; CK( cudaMalloc( &a, ( size_t ) 96 << 20 ) ); CK( cudaMemset( a, 1, ( size_t ) 96 << 20 ) );
		for( int mb : { 16, 32, 64, 96 } ) {
			const size_t n = ( size_t ) mb << 20 >> 4;
			const int reps = 1536 / mb;
			float best = 1e30f;
			for( int ctas : { 4, 8 } ) for( int tpb : { 256, 512 } ) {
				const float ms = time_ms( [&] { read_cg_kernel<<< sms * ctas, tpb >>>( a, n, reps, fsink ); } );
				if( ms < best ) best = ms;
			}
			printf( ", \"l2_read_cg_%dmb_gbs\": %.1f", mb, ( double ) reps * n * 16 / best / 1e6 );
		}
		{
			const size_t n = ( size_t ) 32 << 20 >> 4;
			const int reps = 20;
			const float ms = time_ms( [&] { read_ca_kernel<<< sms * 8, 512 >>>( a, n, reps, fsink ); } );
			printf( ", \"l2_read_ca_32mb_gbs\": %.1f", ( double ) reps * n * 16 / ms / 1e6 );
		}
		CK( cudaFree( a ) );
	}
	{   /* shared memory, conflict-free LDS.128 */
		const int iters = 20000, threads = 1024;
		CK( cudaFuncSetAttribute( smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 ) );
		const float ms = time_ms( [&] { smem_kernel<<< sms, threads, 65536 >>>( iters, fsink ); } );
		printf( ", \"smem_read_gbs\": %.1f", ( double ) sms * threads * iters * 16.0 / ms / 1e6 );
	}
	{   /* the trace's shared-memory record patterns over a 128 KB tile, 1024 threads per SM */
		const uint32_t tile = 128 * 1024;
		std::vector< uint32_t > h( tile / 4 );
		uint32_t s = 777;
		for( auto & v : h ) { s = s * 1664525u + 1013904223u; v = s; }
		uint4 * d; CK( cudaMalloc( &d, tile ) ); CK( cudaMemcpy( d, h.data(), tile, cudaMemcpyHostToDevice ) );
		const int iters = 4096;
		CK( cudaFuncSetAttribute( smem_rec_kernel< 32 >, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) tile ) );
		CK( cudaFuncSetAttribute( smem_rec_kernel< 16 >, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) tile ) );
		float ms = time_ms( [&] { smem_rec_kernel< 32 ><<< sms, 1024, tile >>>( d, tile, iters, usink ); } );
		double rate = ( double ) sms * 1024 * iters / ms * 1e3;
		printf( ", \"smem_rec32_per_s\": %.4g, \"smem_rec32_gbs_24B\": %.1f, \"smem_rec32_gbs_32B\": %.1f", rate, rate * 24 / 1e9, rate * 32 / 1e9 );
		ms = time_ms( [&] { smem_rec_kernel< 16 ><<< sms, 1024, tile >>>( d, tile, iters, usink ); } );
		rate = ( double ) sms * 1024 * iters / ms * 1e3;
		printf( ", \"smem_rec16_per_s\": %.4g, \"smem_rec16_gbs\": %.1f", rate, rate * 16 / 1e9 );
		CK( cudaFree( d ) );
	}
	for( size_t map_bytes : { ( size_t ) 64 << 10, ( size_t ) 2 << 20, ( size_t ) 32 << 20 } ) {   /* scattered record gathers, read-only path */
		uint32_t nrec = 1; while( ( size_t ) nrec * 2 * sizeof( Rec ) <= map_bytes ) nrec *= 2;
		std::vector< Rec > h( nrec );
		uint32_t s = 12345;
		for( uint32_t i = 0; i < nrec; i++ ) for( int k = 0; k < 8; k++ ) { s = s * 1664525u + 1013904223u; h[ i ].w[ k ] = s; }
		Rec * d; CK( cudaMalloc( &d, ( size_t ) nrec * sizeof( Rec ) ) );
		CK( cudaMemcpy( d, h.data(), ( size_t ) nrec * sizeof( Rec ), cudaMemcpyHostToDevice ) );
		const char * tag = map_bytes < ( 1u << 20 ) ? "64kb" : ( map_bytes < ( 16u << 20 ) ? "2mb" : "32mb" );
		const int iters = 2048, tpb = 256, ctas = 5;
		const double total = ( double ) sms * ctas * tpb * iters;
		float ms = time_ms( [&] { gather_indep_kernel<<< sms * ctas, tpb >>>( d, nrec - 1, iters, usink ); } );
		printf( ", \"gather32_indep_%s_per_s\": %.4g, \"gather32_indep_%s_gbs\": %.1f", tag, total / ms * 1e3, tag, total / ms * 1e3 * 32 / 1e9 );
		ms = time_ms( [&] { gather_chase_kernel<<< sms * ctas, tpb >>>( d, nrec - 1, iters, usink ); } );
		printf( ", \"gather32_chase_%s_per_s\": %.4g", tag, total / ms * 1e3 );
		ms = time_ms( [&] { gather16_indep_kernel<<< sms * ctas, tpb >>>( ( const uint4 * ) d, nrec * 2 - 1, iters, usink ); } );
		printf( ", \"gather16_indep_%s_per_s\": %.4g", tag, total / ms * 1e3 );
		CK( cudaFree( d ) );
	}
	printf( "}\n" );
	return 0;
}
