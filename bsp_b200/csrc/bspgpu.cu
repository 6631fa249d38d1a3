/*
 * bspgpu.cu -- implementation of the C ABI in include/bspgpu.h (libbspgpu.so).
 *
 * Owns all CUDA state behind the opaque handle: the flattened map image in HBM, the work /
 * hit counters, streams, and (for host-pointer calls) a 3-deep ring of device staging
 * buffers so that the H2D copy of chunk c+1, the trace kernel of chunk c and the D2H copy of
 * chunk c-1 overlap on three streams.  There is no CPU fallback anywhere in this file: if
 * the device or the kernels are unavailable the call fails with BSPGPU_E_CUDA.
 */
#include <cuda_runtime.h>

#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bspgpu.h"
#include "flatten.h"
#include "seggen.cuh"
#include "trace_kernels.cuh"

using namespace bspk;

static_assert( sizeof( bspgpu_result ) == sizeof( ResultRec ) && sizeof( ResultRec ) == 16, "result wire format" );
static_assert( sizeof( bspgpu_segment ) == 32 && sizeof( bspgpu_pos ) == 16, "segment wire format" );
static_assert( sizeof( QNodeRec ) == 32, "image records" );
static_assert( sizeof( NodeRec ) == 32 && sizeof( BrushRef ) == 16 && sizeof( SideRec ) == 16, "image records" );

namespace {

thread_local std::string g_err;

int fail( int code, const std::string & msg ) { g_err = msg; return code; }

#define CK( call ) do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { \
	return fail( BSPGPU_E_CUDA, std::string( #call ) + ": " + cudaGetErrorString( e_ ) ); } } while( 0 )

/* device scratch that is released on every exit path */
struct DeviceScratch {
	void * p = nullptr;
	~DeviceScratch() { if( p ) cudaFree( p ); }
};

const int kBufs = 3;
const int kCounterSlots = 64;               /* one work counter per in-flight launch */
const uint64_t kMaxLaunchSegs = 1ull << 30; /* in-kernel segment indices are 32-bit */

typedef void ( *TraceKernel )( const TraceParams );

template< int M, int S >
TraceKernel pick_threads( int threads, int sched ) {
	if constexpr( M != kGlobalPacked ) {   /* the experimental schedulers walk NodeRecs only */
	if( sched == 1 ) {
		if( threads <= 256 ) return trace_persistent< M, S, 256 >;
		if( threads <= 512 ) return trace_persistent< M, S, 512 >;
		return trace_persistent< M, S, 1024 >;
	}
	if( sched == 3 ) {   /* two segments per lane, one parked in shared memory */
		if( threads <= 256 ) return trace_dual< M, S, 256 >;
		if( threads <= 512 ) return trace_dual< M, S, 512 >;
		return trace_dual< M, S, 1024 >;
	}
	}
	if( sched == 2 ) {   /* experiment: 1536 threads per SM (<= 40 registers) */
		if( threads <= 256 ) return trace_phased< M, S, 256, 6 >;
		if( threads <= 512 ) return trace_phased< M, S, 512, 3 >;
		return trace_phased< M, S, 768, 2 >;
	}
	if( threads <= 256 ) return trace_phased< M, S, 256, M == kGlobalPacked ? 5 : 1 >;   /* 5 CTAs of 256 threads per SM: <= 51 registers */
	if( threads <= 512 ) return trace_phased< M, S, 512, 1 >;
	return trace_phased< M, S, 1024, 1 >;
}
template< int M >
TraceKernel pick_stack( int depth, int threads, int sched ) {
	return depth <= 32 ? pick_threads< M, 32 >( threads, sched ) : pick_threads< M, 64 >( threads, sched );
}
TraceKernel pick_simple( int depth, int threads ) {
	if( depth <= 32 ) return threads <= 256 ? trace_simple< kGlobal, 32, 256 > : trace_simple< kGlobal, 32, 1024 >;
	return threads <= 256 ? trace_simple< kGlobal, 64, 256 > : trace_simple< kGlobal, 64, 1024 >;
}
/* phase-scheduled kernel, read-only map path, 256-thread CTAs, first `slots` stack entries in shared memory */
template< int S >
TraceKernel pick_smem_stack( int slots ) {
	switch( slots ) {
		case 2: return trace_phased< kGlobal, S, 256, 5, 2 >;
		case 4: return trace_phased< kGlobal, S, 256, 5, 4 >;
		case 6: return trace_phased< kGlobal, S, 256, 5, 6 >;
		default: return trace_phased< kGlobal, S, 256, 5, 8 >;
	}
}
/* same with the whole map staged and one 1024-thread CTA per SM */
template< int S >
TraceKernel pick_smem_stack_staged( int slots ) {
	switch( slots ) {
		case 2: return trace_phased< kSmemAll, S, 1024, 1, 2 >;
		case 4: return trace_phased< kSmemAll, S, 1024, 1, 4 >;
		default: return trace_phased< kSmemAll, S, 1024, 1, 6 >;
	}
}
TraceKernel pick_kernel( int mode, int depth, int threads, int sched, int smem_stack ) {
	if( sched == 4 ) return pick_simple( depth, threads );
	if( smem_stack > 0 && mode == kSmemAll ) return depth <= 32 ? pick_smem_stack_staged< 32 >( smem_stack ) : pick_smem_stack_staged< 64 >( smem_stack );
	if( smem_stack > 0 ) return depth <= 32 ? pick_smem_stack< 32 >( smem_stack ) : pick_smem_stack< 64 >( smem_stack );
	switch( mode ) {
		case kSmemAll: return pick_stack< kSmemAll >( depth, threads, sched );
		case kSmemTop: return pick_stack< kSmemTop >( depth, threads, sched );
		case kGlobalPacked: return pick_stack< kGlobalPacked >( depth, threads, sched );
		default: return pick_stack< kGlobal >( depth, threads, sched );
	}
}

} // namespace

struct bspgpu {
	std::mutex lock;              /* calls on one handle are serialised (host state + the handle's stream) */
	int device = 0;
	int sm_count = 0;
	int max_smem_optin = 0;
	cudaStream_t own_stream = nullptr, user_stream = nullptr, copy_in = nullptr, copy_out = nullptr;
	unsigned char * d_image = nullptr;
	MapImageLayout lay;
	unsigned long long * d_counters = nullptr;   /* [0] hits, [1 .. kCounterSlots] work counters */
	uint32_t launch_seq = 0;

	/* options */
	int64_t opt_variant = BSPGPU_VARIANT_AUTO, opt_threads = 0, opt_ctas = 0, opt_top_bytes = 96 * 1024;
	int64_t opt_chunk = 2ll << 20, opt_refill = 8, opt_max_steps = 0, opt_block_segs = 256;   /* 0 = per-variant default */
	int64_t opt_sched = 0, opt_side_steps = 3, opt_leaf_threshold = 16;
	int64_t opt_smem_stack = -1;   /* GLOBAL variant, 256-thread CTAs: traversal-stack slots per thread kept in shared memory (0, 2, 4, 6, 8; -1 auto = 6) */
	int64_t opt_packed_nodes = 0;  /* 1: two-level node records on the read-only path (measured slower, profiles/README.md; kept as a tested option) */
	int64_t opt_start_grid = -1;   /* -1 auto: on unless the whole map sits in shared memory (there a node step is cheaper than the grid's global load) */
	uint64_t opt_max_launch = kMaxLaunchSegs;   /* segments per kernel launch (lowered by tests to exercise the multi-launch path) */

	/* launch configuration, recomputed only after an option changed */
	bool cfg_valid = false;
	TraceKernel cfg_kernel = nullptr;
	int cfg_variant = 0, cfg_mode = 0, cfg_threads = 0, cfg_grid = 0;
	uint32_t cfg_smem = 0, cfg_staged = 0, cfg_top_nodes = 0;

	/* what the last trace did */
	uint32_t last_variant = 0, last_threads = 0, last_grid = 0, last_smem = 0, last_launches = 0;
	cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
	bool timing_valid = false;

	/* host-pointer path */
	void * d_in[ kBufs ] = { }, * d_out[ kBufs ] = { }, * d_pos[ kBufs ] = { };
	size_t buf_segs = 0;
	cudaEvent_t ev_in[ kBufs ] = { }, ev_k[ kBufs ] = { }, ev_out[ kBufs ] = { };

	bool use_user_stream = false;
	cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }
};

namespace {

int enqueue_trace( bspgpu * h, const void * d_segs, bool packed24, uint64_t n, bspgpu_result * d_out, bspgpu_pos * d_pos, cudaStream_t st ) {
	if( !h->cfg_valid ) {
		int variant = ( int ) h->opt_variant;
		const uint64_t smem_cap = ( uint64_t ) h->max_smem_optin > 1024 ? ( uint64_t ) h->max_smem_optin - 1024 : 0;
		/* AUTO (measured on B200, profiles/): the whole map in shared memory when it fits; otherwise the read-only
		 * global path -- for a few-MB map L1 already keeps the top of the tree resident and staging it buys nothing */
		if( variant == BSPGPU_VARIANT_AUTO ) variant = h->lay.staged_bytes <= smem_cap ? BSPGPU_VARIANT_SMEM : BSPGPU_VARIANT_GLOBAL;
		if( h->opt_sched == 4 ) variant = BSPGPU_VARIANT_GLOBAL;   /* the one-thread-per-segment kernel stages nothing */
		int mode; uint32_t smem = 0, top_nodes = 0;
		if( variant == BSPGPU_VARIANT_SMEM ) {
			if( h->lay.staged_bytes > smem_cap ) return fail( BSPGPU_E_INVALID, "SMEM variant: flattened map does not fit in shared memory" );
			mode = kSmemAll; smem = ( uint32_t ) h->lay.staged_bytes;
		}
		else if( variant == BSPGPU_VARIANT_TOPTREE ) {
			mode = kSmemTop;
			uint64_t bytes = ( uint64_t ) h->opt_top_bytes;
			if( bytes > smem_cap ) bytes = smem_cap;
			const uint64_t all = ( uint64_t ) h->lay.num_nodes * sizeof( NodeRec );
			if( bytes > all ) bytes = all;
			top_nodes = ( uint32_t ) ( bytes / sizeof( NodeRec ) );
			smem = top_nodes * ( uint32_t ) sizeof( NodeRec );
		}
		else mode = ( h->opt_packed_nodes > 0 && ( h->opt_sched == 0 || h->opt_sched == 2 ) ) ? kGlobalPacked : kGlobal;

		int threads = ( int ) h->opt_threads;
		if( threads <= 0 ) threads = ( mode == kSmemAll ) ? 1024 : 256;   /* measured: 1 x 1024 with the map in smem, 5 x 256 otherwise */
		threads = ( threads + 31 ) / 32 * 32;
		if( threads > 1024 ) threads = 1024;
		if( h->opt_sched == 2 && threads > 768 ) threads = 768;

		const uint32_t staged = smem;
		if( h->opt_sched == 3 ) {
			smem += 96u * ( uint32_t ) threads;   /* parking slots of trace_dual */
			if( smem > ( uint32_t ) h->max_smem_optin ) return fail( BSPGPU_E_INVALID, "two-segment scheduler: staged map + 96 B/thread of parking slots exceed shared memory" );
		}
		int smem_stack = 0;
		const int64_t want_stack = h->opt_smem_stack < 0 ? 6 : h->opt_smem_stack;   /* measured: 4 -> +10 %, 6 -> +17 %, 8 -> +10 % on city10k long rays */
		if( mode == kGlobal && h->opt_sched == 0 && threads <= 256 && want_stack > 0 ) {
			smem_stack = want_stack <= 2 ? 2 : ( want_stack <= 4 ? 4 : ( want_stack <= 6 ? 6 : 8 ) );
			smem += ( uint32_t ) smem_stack * 256u * 16u;
		}
		else if( mode == kSmemAll && h->opt_sched == 0 && threads > 512 && h->opt_smem_stack > 0 ) {
			/* behind the staged map, as many slots as still fit (explicit request only: measured in profiles/README.md) */
			smem = ( smem + 127u ) & ~127u;
			smem_stack = h->opt_smem_stack <= 2 ? 2 : ( h->opt_smem_stack <= 4 ? 4 : 6 );
			while( smem_stack > 0 && ( uint64_t ) smem + ( uint64_t ) smem_stack * 1024u * 16u > smem_cap ) smem_stack -= 2;   /* smem_cap leaves room for static shared memory */
			smem += ( uint32_t ) smem_stack * 1024u * 16u;
		}
		TraceKernel kernel = pick_kernel( mode, ( int ) h->lay.tree_depth, threads, ( int ) h->opt_sched, smem_stack );
		CK( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) smem ) );
		int ctas = ( int ) h->opt_ctas;
		int occ = 0;
		CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, kernel, threads, smem ) );
		if( occ < 1 ) return fail( BSPGPU_E_CUDA, "kernel does not fit on an SM with the requested block size / shared memory" );
		if( ctas <= 0 || ctas > occ ) ctas = occ;
		h->cfg_kernel = kernel; h->cfg_variant = variant; h->cfg_mode = mode; h->cfg_threads = threads;
		h->cfg_grid = h->sm_count * ctas * ( h->opt_sched == 4 ? 4 : 1 ); h->cfg_smem = smem; h->cfg_staged = staged; h->cfg_top_nodes = top_nodes;
		h->cfg_valid = true;
	}
	TraceKernel kernel = h->cfg_kernel;
	const int variant = h->cfg_variant, mode = h->cfg_mode, threads = h->cfg_threads, grid = h->cfg_grid;
	const uint32_t smem = h->cfg_smem, top_nodes = h->cfg_top_nodes;

	TraceParams p;
	p.image = h->d_image;
	p.off_refs = h->lay.off_refs; p.off_sides = h->lay.off_sides; p.off_side_plane = h->lay.off_side_plane; p.off_qnodes = h->lay.off_qnodes;
	p.staged_bytes = h->cfg_staged; p.top_nodes = top_nodes;
	p.block_segs = ( uint32_t ) ( ( h->opt_block_segs + 31 ) / 32 * 32 );
	p.refill = ( uint32_t ) ( h->opt_refill < 1 ? 1 : ( h->opt_refill > 32 ? 32 : h->opt_refill ) );
	p.max_steps = ( uint32_t ) ( h->opt_max_steps < 1 ? ( mode == kSmemAll ? 16 : 8 ) : h->opt_max_steps );
	p.side_steps = ( uint32_t ) ( h->opt_side_steps < 1 ? 1 : h->opt_side_steps );
	p.leaf_threshold = ( uint32_t ) ( h->opt_leaf_threshold < 1 ? 1 : ( h->opt_leaf_threshold > 32 ? 32 : h->opt_leaf_threshold ) );
	p.hits = h->d_counters;
	p.grid.d = h->lay.grid;
	const bool use_grid = h->opt_start_grid < 0 ? mode != kSmemAll : h->opt_start_grid != 0;
	p.grid.cells = ( use_grid && h->opt_sched != 3 && h->lay.grid.dim[ 0 ] ) ? reinterpret_cast< const int32_t * >( h->d_image + h->lay.off_grid ) : nullptr;

	h->last_variant = ( uint32_t ) variant; h->last_threads = ( uint32_t ) threads; h->last_grid = ( uint32_t ) grid; h->last_smem = smem;

	for( uint64_t done = 0; done < n; done += h->opt_max_launch ) {
		const uint64_t cnt = n - done < h->opt_max_launch ? n - done : h->opt_max_launch;
		unsigned long long * next = h->d_counters + 1 + ( h->launch_seq++ % kCounterSlots );
		CK( cudaMemsetAsync( next, 0, sizeof( unsigned long long ), st ) );
		p.segs = packed24 ? nullptr : reinterpret_cast< const float4 * >( d_segs ) + done * 2;
		p.segs24 = packed24 ? reinterpret_cast< const float2 * >( d_segs ) + done * 3 : nullptr;
		p.out = d_out ? reinterpret_cast< ResultRec * >( d_out ) + done : nullptr;
		p.pos = d_pos ? reinterpret_cast< float4 * >( d_pos ) + done : nullptr;
		p.n = ( uint32_t ) cnt;
		p.next = next;
		kernel<<< grid, threads, smem, st >>>( p );
		CK( cudaGetLastError() );
		h->last_launches++;
	}
	return BSPGPU_OK;
}

/* device staging buffers of the host-pointer pipeline: kBufs chunks of each kind that is actually needed */
int ensure_staging( bspgpu * h, size_t segs, bool want_in, bool want_out, bool want_pos ) {
	if( h->buf_segs < segs ) {
		for( int b = 0; b < kBufs; b++ ) {
			if( h->d_in[ b ] ) cudaFree( h->d_in[ b ] );
			if( h->d_out[ b ] ) cudaFree( h->d_out[ b ] );
			if( h->d_pos[ b ] ) cudaFree( h->d_pos[ b ] );
			h->d_in[ b ] = h->d_out[ b ] = h->d_pos[ b ] = nullptr;
		}
		h->buf_segs = segs;
	}
	for( int b = 0; b < kBufs; b++ ) {
		if( want_in && !h->d_in[ b ] ) CK( cudaMalloc( &h->d_in[ b ], h->buf_segs * sizeof( bspgpu_segment ) ) );
		if( want_out && !h->d_out[ b ] ) CK( cudaMalloc( &h->d_out[ b ], h->buf_segs * sizeof( bspgpu_result ) ) );
		if( want_pos && !h->d_pos[ b ] ) CK( cudaMalloc( &h->d_pos[ b ], h->buf_segs * sizeof( bspgpu_pos ) ) );
	}
	return BSPGPU_OK;
}

} // namespace

extern "C" {

const char * bspgpu_last_error( void ) { return g_err.c_str(); }
uint32_t bspgpu_abi_version( void ) { return BSPGPU_ABI_VERSION; }

void bspgpu_shard_range( uint64_t n, uint32_t rank, uint32_t world, uint64_t * first, uint64_t * count ) {
	if( world == 0 ) world = 1;
	const uint64_t per = ( n + world - 1 ) / world;
	uint64_t f = per * rank;
	if( f > n ) f = n;
	uint64_t c = n - f < per ? n - f : per;
	if( first ) *first = f;
	if( count ) *count = c;
}

int bspgpu_create( bspgpu_t ** out, const bspgpu_map_desc * lumps, int device ) {
	if( !out || !lumps ) return fail( BSPGPU_E_INVALID, "bspgpu_create: null argument" );
	*out = nullptr;
	FlatMap flat;
	std::string err;
	const int rc = flatten_map( *lumps, flat, err );
	if( rc != BSPGPU_OK ) return fail( rc, "bspgpu_create: " + err );
	if( flat.layout.tree_depth > 64 ) return fail( BSPGPU_E_MAP, "bspgpu_create: tree deeper than 64 levels" );

	int count = 0;
	if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 ) {
		cudaGetLastError();
		return fail( BSPGPU_E_CUDA, "bspgpu_create: no CUDA device (this library has no CPU fallback)" );
	}
	if( device < 0 ) CK( cudaGetDevice( &device ) );
	if( device >= count ) return fail( BSPGPU_E_INVALID, "bspgpu_create: device index out of range" );
	CK( cudaSetDevice( device ) );
	cudaDeviceProp prop;
	CK( cudaGetDeviceProperties( &prop, device ) );
	if( prop.major != 10 ) return fail( BSPGPU_E_CUDA, "bspgpu_create: kernels are built for sm_100a only, device is sm_" + std::to_string( prop.major * 10 + prop.minor ) );

	bspgpu * h = new bspgpu();
	h->device = device;
	h->sm_count = prop.multiProcessorCount;
	h->max_smem_optin = ( int ) prop.sharedMemPerBlockOptin;
	h->lay = flat.layout;
	cudaError_t e = cudaSuccess;
	do {
		if( ( e = cudaStreamCreateWithFlags( &h->own_stream, cudaStreamNonBlocking ) ) != cudaSuccess ) break;
		if( ( e = cudaStreamCreateWithFlags( &h->copy_in, cudaStreamNonBlocking ) ) != cudaSuccess ) break;
		if( ( e = cudaStreamCreateWithFlags( &h->copy_out, cudaStreamNonBlocking ) ) != cudaSuccess ) break;
		if( ( e = cudaEventCreate( &h->ev_start ) ) != cudaSuccess ) break;
		if( ( e = cudaEventCreate( &h->ev_stop ) ) != cudaSuccess ) break;
		for( int b = 0; b < kBufs && e == cudaSuccess; b++ ) {
			if( ( e = cudaEventCreateWithFlags( &h->ev_in[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
			if( ( e = cudaEventCreateWithFlags( &h->ev_k[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
			if( ( e = cudaEventCreateWithFlags( &h->ev_out[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
		}
		if( e != cudaSuccess ) break;
		if( ( e = cudaMalloc( &h->d_image, flat.layout.total_bytes ) ) != cudaSuccess ) break;
		if( ( e = cudaMemcpy( h->d_image, flat.image.data(), flat.layout.total_bytes, cudaMemcpyHostToDevice ) ) != cudaSuccess ) break;
		if( ( e = cudaMalloc( &h->d_counters, ( 1 + kCounterSlots ) * sizeof( unsigned long long ) ) ) != cudaSuccess ) break;
		if( ( e = cudaMemset( h->d_counters, 0, ( 1 + kCounterSlots ) * sizeof( unsigned long long ) ) ) != cudaSuccess ) break;
	} while( 0 );
	if( e != cudaSuccess ) {
		const std::string msg = std::string( "bspgpu_create: " ) + cudaGetErrorString( e );
		bspgpu_destroy( h );
		return fail( BSPGPU_E_CUDA, msg );
	}
	*out = h;
	return BSPGPU_OK;
}

int bspgpu_create_from_file( bspgpu_t ** out, const char * path, int device ) {
	if( !out || !path ) return fail( BSPGPU_E_INVALID, "bspgpu_create_from_file: null argument" );
	std::vector< unsigned char > blob;
	bspgpu_map_desc lumps;
	memset( &lumps, 0, sizeof( lumps ) );
	std::string err;
	const int rc = read_ibsp_file( path, blob, lumps, err );
	if( rc != BSPGPU_OK ) return fail( rc, "bspgpu_create_from_file: " + err );
	return bspgpu_create( out, &lumps, device );
}

void bspgpu_destroy( bspgpu_t * h ) {
	if( !h ) return;
	cudaSetDevice( h->device );
	/* nothing of this handle may still be in flight when its buffers go away (BSPGPU_ASYNC work on a caller's stream included) */
	if( h->own_stream ) cudaStreamSynchronize( h->own_stream );
	if( h->use_user_stream ) cudaStreamSynchronize( h->user_stream );
	if( h->copy_in ) cudaStreamSynchronize( h->copy_in );
	if( h->copy_out ) cudaStreamSynchronize( h->copy_out );
	for( int b = 0; b < kBufs; b++ ) {
		if( h->d_in[ b ] ) cudaFree( h->d_in[ b ] );
		if( h->d_out[ b ] ) cudaFree( h->d_out[ b ] );
		if( h->d_pos[ b ] ) cudaFree( h->d_pos[ b ] );
		if( h->ev_in[ b ] ) cudaEventDestroy( h->ev_in[ b ] );
		if( h->ev_k[ b ] ) cudaEventDestroy( h->ev_k[ b ] );
		if( h->ev_out[ b ] ) cudaEventDestroy( h->ev_out[ b ] );
	}
	if( h->d_image ) cudaFree( h->d_image );
	if( h->d_counters ) cudaFree( h->d_counters );
	if( h->ev_start ) cudaEventDestroy( h->ev_start );
	if( h->ev_stop ) cudaEventDestroy( h->ev_stop );
	if( h->own_stream ) cudaStreamDestroy( h->own_stream );
	if( h->copy_in ) cudaStreamDestroy( h->copy_in );
	if( h->copy_out ) cudaStreamDestroy( h->copy_out );
	delete h;
}

int bspgpu_trace_pos( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( n && ( !segs || ( !out && !pos ) ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null segs, or both out and pos null" );
	const bool segs_dev = ( flags & BSPGPU_SEGS_DEVICE ) != 0, out_dev = ( flags & BSPGPU_OUT_DEVICE ) != 0;
	const bool packed = ( flags & BSPGPU_SEGS_PACKED24 ) != 0;
	if( ( flags & BSPGPU_ASYNC ) && !( segs_dev && out_dev ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: BSPGPU_ASYNC needs device pointers" );
	/* the kernels read segments with 128-bit (packed: 64-bit) loads and write 16-byte records */
	if( segs_dev && ( ( uintptr_t ) segs % ( packed ? 8 : 16 ) ) != 0 ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: device segments must be 16-byte aligned (8-byte for the packed layout)" );
	if( out_dev && ( ( ( uintptr_t ) out % 16 ) != 0 || ( ( uintptr_t ) pos % 16 ) != 0 ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: device results must be 16-byte aligned" );
	CK( cudaSetDevice( h->device ) );
	cudaStream_t st = h->stream();
	h->last_launches = 0;
	h->timing_valid = false;
	CK( cudaMemsetAsync( h->d_counters, 0, sizeof( unsigned long long ), st ) );
	if( n == 0 ) return BSPGPU_OK;
	const size_t seg_bytes = packed ? 24 : sizeof( bspgpu_segment );

	if( segs_dev && out_dev ) {
		CK( cudaEventRecord( h->ev_start, st ) );
		const int rc = enqueue_trace( h, segs, packed, n, out, pos, st );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaEventRecord( h->ev_stop, st ) );
		h->timing_valid = true;
		if( !( flags & BSPGPU_ASYNC ) ) CK( cudaStreamSynchronize( st ) );
		return BSPGPU_OK;
	}

	/* ---- at least one side lives in host memory: chunked 3-stream pipeline */
	uint64_t chunk = ( uint64_t ) ( h->opt_chunk < 1024 ? 1024 : h->opt_chunk );
	if( chunk > n ) chunk = n;
	if( chunk > kMaxLaunchSegs ) chunk = kMaxLaunchSegs;
	{
		const int rc = ensure_staging( h, ( size_t ) chunk, !segs_dev, out && !out_dev, pos && !out_dev );
		if( rc != BSPGPU_OK ) return rc;
	}
	/* order the side streams after whatever the compute stream already holds */
	CK( cudaEventRecord( h->ev_start, st ) );
	CK( cudaStreamWaitEvent( h->copy_in, h->ev_start, 0 ) );
	uint64_t c = 0;
	for( uint64_t done = 0; done < n; done += chunk, c++ ) {
		const uint64_t cnt = n - done < chunk ? n - done : chunk;
		const int b = ( int ) ( c % kBufs );
		const void * d_segs;
		if( segs_dev ) d_segs = ( const char * ) segs + done * seg_bytes;
		else {
			if( c >= ( uint64_t ) kBufs ) CK( cudaStreamWaitEvent( h->copy_in, h->ev_k[ b ], 0 ) );   /* kernel that last read d_in[b] */
			CK( cudaMemcpyAsync( h->d_in[ b ], ( const char * ) segs + done * seg_bytes, cnt * seg_bytes, cudaMemcpyHostToDevice, h->copy_in ) );
			CK( cudaEventRecord( h->ev_in[ b ], h->copy_in ) );
			CK( cudaStreamWaitEvent( st, h->ev_in[ b ], 0 ) );
			d_segs = h->d_in[ b ];
		}
		bspgpu_result * k_out = out ? ( out_dev ? out + done : ( bspgpu_result * ) h->d_out[ b ] ) : nullptr;
		bspgpu_pos * k_pos = pos ? ( out_dev ? pos + done : ( bspgpu_pos * ) h->d_pos[ b ] ) : nullptr;
		if( !out_dev && c >= ( uint64_t ) kBufs ) CK( cudaStreamWaitEvent( st, h->ev_out[ b ], 0 ) );   /* D2H that last read d_out[b] */
		const int rc = enqueue_trace( h, d_segs, packed, cnt, k_out, k_pos, st );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaEventRecord( h->ev_k[ b ], st ) );
		if( !out_dev ) {
			CK( cudaStreamWaitEvent( h->copy_out, h->ev_k[ b ], 0 ) );
			if( out ) CK( cudaMemcpyAsync( out + done, h->d_out[ b ], cnt * sizeof( bspgpu_result ), cudaMemcpyDeviceToHost, h->copy_out ) );
			if( pos ) CK( cudaMemcpyAsync( pos + done, h->d_pos[ b ], cnt * sizeof( bspgpu_pos ), cudaMemcpyDeviceToHost, h->copy_out ) );
			CK( cudaEventRecord( h->ev_out[ b ], h->copy_out ) );
		}
	}
	CK( cudaEventRecord( h->ev_stop, st ) );
	h->timing_valid = true;
	CK( cudaStreamSynchronize( st ) );
	CK( cudaStreamSynchronize( h->copy_in ) );
	CK( cudaStreamSynchronize( h->copy_out ) );
	return BSPGPU_OK;
}

int bspgpu_trace( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, unsigned flags ) {
	if( n && !out ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null out" );
	return bspgpu_trace_pos( h, segs, n, out, nullptr, flags );
}

int bspgpu_hit_count( bspgpu_t * h, uint64_t * hits ) {
	if( !h || !hits ) return fail( BSPGPU_E_INVALID, "bspgpu_hit_count: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	CK( cudaSetDevice( h->device ) );
	unsigned long long v = 0;
	CK( cudaMemcpyAsync( &v, h->d_counters, sizeof( v ), cudaMemcpyDeviceToHost, h->stream() ) );
	CK( cudaStreamSynchronize( h->stream() ) );
	*hits = v;
	return BSPGPU_OK;
}

int bspgpu_hit_count_device( bspgpu_t * h, void * dev_u64 ) {
	if( !h || !dev_u64 ) return fail( BSPGPU_E_INVALID, "bspgpu_hit_count_device: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	CK( cudaSetDevice( h->device ) );
	CK( cudaMemcpyAsync( dev_u64, h->d_counters, sizeof( unsigned long long ), cudaMemcpyDeviceToDevice, h->stream() ) );
	return BSPGPU_OK;
}

int bspgpu_leaf( bspgpu_t * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, unsigned flags ) {
	if( !h || ( n && ( !pts || !leaf ) ) || stride < 3 ) return fail( BSPGPU_E_INVALID, "bspgpu_leaf: bad argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( n == 0 ) return BSPGPU_OK;
	CK( cudaSetDevice( h->device ) );
	cudaStream_t st = h->stream();
	const bool pts_dev = ( flags & BSPGPU_SEGS_DEVICE ) != 0, out_dev = ( flags & BSPGPU_OUT_DEVICE ) != 0;
	DeviceScratch s_pts, s_leaf;
	if( !pts_dev ) {
		CK( cudaMalloc( &s_pts.p, n * stride * sizeof( float ) ) );
		CK( cudaMemcpyAsync( s_pts.p, pts, n * stride * sizeof( float ), cudaMemcpyHostToDevice, st ) );
	}
	if( !out_dev ) CK( cudaMalloc( &s_leaf.p, n * sizeof( int32_t ) ) );
	const float * k_pts = pts_dev ? pts : ( const float * ) s_pts.p;
	int32_t * k_leaf = out_dev ? leaf : ( int32_t * ) s_leaf.p;
	const int threads = 256;
	uint64_t blocks = ( n + threads - 1 ) / threads;
	if( blocks > ( uint64_t ) h->sm_count * 16 ) blocks = ( uint64_t ) h->sm_count * 16;
	leaf_kernel<<< ( unsigned ) blocks, threads, 0, st >>>( h->d_image, k_pts, stride, n, k_leaf );
	CK( cudaGetLastError() );
	if( !out_dev ) CK( cudaMemcpyAsync( leaf, k_leaf, n * sizeof( int32_t ), cudaMemcpyDeviceToHost, st ) );
	/* scratch is freed when this function returns, so wait unless the caller owns every buffer */
	if( !( ( flags & BSPGPU_ASYNC ) && pts_dev && out_dev ) ) CK( cudaStreamSynchronize( st ) );
	return BSPGPU_OK;
}

int bspgpu_generate( bspgpu_t * h, const bspgpu_gen_desc * g, uint64_t first, uint64_t count, void * dev_segs ) {
	if( !h || !g || ( count && !dev_segs ) ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( g->kind > 2 ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: unknown kind" );
	if( count == 0 ) return BSPGPU_OK;
	CK( cudaSetDevice( h->device ) );
	cudaStream_t st = h->stream();
	GenParams p;
	memset( &p, 0, sizeof( p ) );
	p.kind = g->kind;
	p.key = sm64_fin( g->seed + kGolden );
	for( int a = 0; a < 3; a++ ) { p.lo[ a ] = g->lo[ a ]; p.ext[ a ] = g->hi[ a ] - g->lo[ a ]; }
	p.len_lo = g->len_lo; p.len_ext = g->len_hi - g->len_lo;
	p.first = first; p.count = count; p.out = reinterpret_cast< float4 * >( dev_segs );
	DeviceScratch s_views;
	float * d_views = nullptr;
	if( g->kind == 2 ) {
		if( !g->views || !g->num_views || !g->width || !g->height ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: camera kind needs views and a grid size" );
		if( first + count > ( uint64_t ) g->num_views * g->width * g->height ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: camera range exceeds views*width*height" );
		CK( cudaMalloc( &s_views.p, ( size_t ) g->num_views * 12 * sizeof( float ) ) );
		d_views = ( float * ) s_views.p;
		CK( cudaMemcpyAsync( d_views, g->views, ( size_t ) g->num_views * 12 * sizeof( float ), cudaMemcpyHostToDevice, st ) );
		p.views = d_views; p.width = g->width; p.height = g->height;
		p.inv2w = ( float ) ( 2.0 / ( double ) g->width ); p.inv2h = ( float ) ( 2.0 / ( double ) g->height );
		p.tanx = g->tan_half; p.tany = g->tan_half * ( float ) g->height / ( float ) g->width;
		p.far_dist = g->far_dist;
	}
	const int threads = 256;
	uint64_t blocks = ( count + threads - 1 ) / threads;
	if( blocks > ( uint64_t ) h->sm_count * 32 ) blocks = ( uint64_t ) h->sm_count * 32;
	generate_kernel<<< ( unsigned ) blocks, threads, 0, st >>>( p );
	CK( cudaGetLastError() );
	CK( cudaStreamSynchronize( st ) );
	return BSPGPU_OK;
}

int bspgpu_set_option( bspgpu_t * h, int option, int64_t value ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_set_option: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	h->cfg_valid = false;
	switch( option ) {
		case BSPGPU_OPT_VARIANT: if( value < 0 || value > 3 ) return fail( BSPGPU_E_INVALID, "bad variant" ); h->opt_variant = value; break;
		case BSPGPU_OPT_BLOCK_THREADS: h->opt_threads = value; break;
		case BSPGPU_OPT_CTAS_PER_SM: h->opt_ctas = value; break;
		case BSPGPU_OPT_TOPTREE_BYTES: h->opt_top_bytes = value < 0 ? 0 : value; break;
		case BSPGPU_OPT_CHUNK_SEGMENTS: h->opt_chunk = value; break;
		case BSPGPU_OPT_REFILL: h->opt_refill = value; break;
		case BSPGPU_OPT_MAX_STEPS: h->opt_max_steps = value; break;
		case BSPGPU_OPT_BLOCK_SEGS: h->opt_block_segs = value; break;
		case BSPGPU_OPT_SCHEDULER: h->opt_sched = value; break;
		case BSPGPU_OPT_SIDE_STEPS: h->opt_side_steps = value; break;
		case BSPGPU_OPT_LEAF_THRESHOLD: h->opt_leaf_threshold = value; break;
		case BSPGPU_OPT_SMEM_STACK: h->opt_smem_stack = value < 0 ? -1 : value; break;
		case BSPGPU_OPT_PACKED_NODES: h->opt_packed_nodes = value > 0 ? 1 : 0; break;
		case BSPGPU_OPT_START_GRID: h->opt_start_grid = value < 0 ? -1 : ( value ? 1 : 0 ); break;
		case BSPGPU_OPT_MAX_LAUNCH_SEGMENTS:
			if( value < 32 || ( uint64_t ) value > kMaxLaunchSegs ) return fail( BSPGPU_E_INVALID, "max launch segments must be in [32, 2^30]" );
			h->opt_max_launch = ( uint64_t ) value; break;
		default: return fail( BSPGPU_E_INVALID, "bspgpu_set_option: unknown option" );
	}
	return BSPGPU_OK;
}

int bspgpu_set_stream( bspgpu_t * h, void * cuda_stream ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_set_stream: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	h->user_stream = ( cudaStream_t ) cuda_stream;
	h->use_user_stream = true;
	return BSPGPU_OK;
}

int bspgpu_reset_stream( bspgpu_t * h ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_reset_stream: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	h->user_stream = nullptr;
	h->use_user_stream = false;
	return BSPGPU_OK;
}

int bspgpu_sync( bspgpu_t * h ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_sync: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	CK( cudaSetDevice( h->device ) );
	CK( cudaStreamSynchronize( h->stream() ) );
	return BSPGPU_OK;
}

int bspgpu_get_info( bspgpu_t * h, bspgpu_info * info ) {
	if( !h || !info ) return fail( BSPGPU_E_INVALID, "bspgpu_get_info: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	memset( info, 0, sizeof( *info ) );
	info->abi_version = BSPGPU_ABI_VERSION;
	info->device = h->device; info->sm_count = ( uint32_t ) h->sm_count;
	info->variant = h->last_variant; info->block_threads = h->last_threads; info->grid_ctas = h->last_grid;
	info->smem_bytes = h->last_smem; info->tree_depth = h->lay.tree_depth;
	info->num_nodes = h->lay.num_nodes; info->num_brush_refs = h->lay.num_refs; info->num_sides = h->lay.num_sides;
	info->map_bytes = h->lay.total_bytes; info->last_launches = h->last_launches;
	info->staged_bytes = ( uint32_t ) h->lay.staged_bytes;
	info->start_grid_cells = ( uint32_t ) ( ( uint64_t ) h->lay.grid.dim[ 0 ] * h->lay.grid.dim[ 1 ] * h->lay.grid.dim[ 2 ] );
	if( h->timing_valid ) {
		CK( cudaSetDevice( h->device ) );
		CK( cudaEventSynchronize( h->ev_stop ) );
		float ms = 0.0f;
		CK( cudaEventElapsedTime( &ms, h->ev_start, h->ev_stop ) );
		info->last_kernel_ms = ms;
	}
	return BSPGPU_OK;
}

void * bspgpu_host_alloc( size_t bytes ) {
	void * p = nullptr;
	if( cudaMallocHost( &p, bytes ) != cudaSuccess ) { g_err = "bspgpu_host_alloc: cudaMallocHost failed"; cudaGetLastError(); return nullptr; }
	return p;
}

void bspgpu_host_free( void * p ) { if( p ) cudaFreeHost( p ); }

} // extern "C"
