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
#ifdef BSPGPU_EXPERIMENTS
#include "pool_kernels.cuh"
#endif
#include "bin_kernels.cuh"
#include "bspgpu_internal.h"

using namespace bspk;

static_assert( sizeof( bspgpu_result ) == sizeof( ResultRec ) && sizeof( ResultRec ) == 16, "result wire format" );
static_assert( sizeof( bspgpu_result8 ) == 8, "compact result wire format" );
static_assert( sizeof( bspgpu_segment ) == 32 && sizeof( bspgpu_pos ) == 16, "segment wire format" );
static_assert( sizeof( NodeRec ) == 16 && sizeof( BrushRef ) == 16 && sizeof( SideRec ) == 16 && sizeof( NodeOrig ) == 8, "image records" );

std::string & bspgpu_err() {
	thread_local std::string err;
	return err;
}

CopyPool::CopyPool( int n ) {
	for( int i = 0; i < n; i++ ) {
		workers.emplace_back( [ this ] {
			for( ;; ) {
				Task t;
				{
					std::unique_lock< std::mutex > lk( m );
					cv_work.wait( lk, [ this ] { return stop || !q.empty(); } );
					if( q.empty() ) return;                   /* stop */
					t = q.front(); q.pop_front();
				}
				memcpy( t.dst, t.src, t.bytes );
				{
					std::lock_guard< std::mutex > lk( m );
					if( --pending == 0 ) cv_done.notify_all();
				}
			}
		} );
	}
}
CopyPool::~CopyPool() {
	{ std::lock_guard< std::mutex > lk( m ); stop = true; }
	cv_work.notify_all();
	for( std::thread & w : workers ) w.join();
}
void CopyPool::copy( void * dst, const void * src, size_t bytes ) {
	if( bytes == 0 ) return;
	const size_t kSlice = 1u << 20;                           /* >= 1 MB per slice, at most one slice per worker and copy */
	size_t slices = ( bytes + kSlice - 1 ) / kSlice;
	if( slices > workers.size() ) slices = workers.size();
	const size_t per = ( ( bytes + slices - 1 ) / slices + 63 ) & ~( size_t ) 63;
	{
		std::lock_guard< std::mutex > lk( m );
		for( size_t off = 0; off < bytes; off += per ) {
			q.push_back( Task{ ( char * ) dst + off, ( const char * ) src + off, bytes - off < per ? bytes - off : per } );
			pending++;
		}
	}
	cv_work.notify_all();
}
void CopyPool::wait() {
	std::unique_lock< std::mutex > lk( m );
	cv_done.wait( lk, [ this ] { return pending == 0; } );
}

int bspgpu_order_after_user_work( bspgpu * h, cudaStream_t st ) {
	if( h->user_pending ) { CK( cudaStreamWaitEvent( st, h->ev_user, 0 ) ); h->user_pending = false; }
	return BSPGPU_OK;
}

namespace {

/* page-locks a caller's host range for the duration of a call (BSPGPU_OPT_PAGEABLE) */
struct HostPin {
	void * p = nullptr;
	bool pin( const void * ptr, size_t bytes, bool read_only ) {
		if( cudaHostRegister( const_cast< void * >( ptr ), bytes, read_only ? cudaHostRegisterReadOnly : cudaHostRegisterDefault ) == cudaSuccess ) {
			p = const_cast< void * >( ptr );
			return true;
		}
		cudaGetLastError();
		return false;
	}
	~HostPin() { if( p ) cudaHostUnregister( p ); }
};

bool is_pageable_host( const void * ptr ) {
	cudaPointerAttributes a;
	if( cudaPointerGetAttributes( &a, ptr ) != cudaSuccess ) { cudaGetLastError(); return true; }
	return a.type == cudaMemoryTypeUnregistered;
}

const uint64_t kTinyCall = kTinyRays;       /* host-pointer calls up to this many segments are ONE kernel over the pinned buffer (trace_tiny_kernel) */
const uint64_t kSmallCall = 4096;           /* host-pointer calls up to this many segments go through one pinned bounce buffer on one stream */
const uint64_t kPinThresholdBytes = 8ull << 20;


/* the kernels the product ships: phase-scheduled lanes; map staged (1 CTA of up to 1024 threads per SM) or read through
 * L1/L2 (5 x 256 threads per SM); the first `slots` traversal-stack entries of every thread in shared memory.  The two
 * shapes that are defaults get every slot count, the other block sizes 0 or 6 (keeps the number of instantiations down). */
bool full_slot_range( int mode, int threads ) { return ( mode == kGlobal && threads == 256 ) || ( mode == kSmemAll && threads == 1024 ); }
int effective_slots( int mode, int threads, int64_t want ) {
	if( want <= 0 ) return 0;
	if( !full_slot_range( mode, threads ) ) return 6;
	const int s = want <= 2 ? 2 : ( want <= 4 ? 4 : ( want <= 6 ? 6 : 8 ) );
	return threads > 256 && s > 6 ? 6 : s;
}
/* GLOBAL variant: lanes walk child slots (trace_slots: the children of a node are gathered while its plane test runs);
 * SMEM variant: lanes walk node records (trace_phased: with the map in shared memory a gather is 29 cycles and there is
 * nothing to overlap -- measured +-0).  Experiments builds can put the node-record walk on the GLOBAL variant as well
 * (BSPGPU_OPT_SCHEDULER = 1), for A/B runs. */
template< int M, int S, int T, int B, int K, bool G >
TraceKernel kernel_of( bool node_walk ) {
	( void ) node_walk;
	if constexpr( M == kGlobal ) {
#ifdef BSPGPU_EXPERIMENTS
		if( node_walk ) return trace_phased< M, S, T, B, K, G >;
#endif
		return trace_slots< M, S, T, B, K, G >;
	}
	else return trace_phased< M, S, T, B, K, G >;
}
template< int M, int S, int T, int B, bool G >
TraceKernel pick_slots( int slots, bool node_walk ) {
	if constexpr( ( M == kGlobal && T == 256 ) || ( M == kSmemAll && T == 1024 ) ) {
		switch( slots ) {
			case 0: return kernel_of< M, S, T, B, 0, G >( node_walk );
			case 2: return kernel_of< M, S, T, B, 2, G >( node_walk );
			case 4: return kernel_of< M, S, T, B, 4, G >( node_walk );
			case 6: return kernel_of< M, S, T, B, 6, G >( node_walk );
			default:
				if constexpr( T == 256 ) return kernel_of< M, S, T, B, 8, G >( node_walk );
				else return kernel_of< M, S, T, B, 6, G >( node_walk );
		}
	}
	else return slots ? kernel_of< M, S, T, B, 6, G >( node_walk ) : kernel_of< M, S, T, B, 0, G >( node_walk );
}
template< int M, int S, bool G >
TraceKernel pick_threads( int threads, int slots, bool node_walk ) {
	if( threads <= 256 ) return pick_slots< M, S, 256, M == kGlobal ? 5 : 1, G >( slots, node_walk );
	if( threads <= 512 ) return pick_slots< M, S, 512, 1, G >( slots, node_walk );
	return pick_slots< M, S, 1024, 1, G >( slots, node_walk );
}
/* depth <= 32: 32-entry local overflow stack; anything deeper (up to the 128 levels bspgpu_create accepts): 128 entries */
template< int M >
TraceKernel pick_stack( int depth, int threads, int slots, bool general, bool node_walk ) {
	if( depth <= 32 ) return general ? pick_threads< M, 32, true >( threads, slots, node_walk ) : pick_threads< M, 32, false >( threads, slots, node_walk );
	return general ? pick_threads< M, 128, true >( threads, slots, node_walk ) : pick_threads< M, 128, false >( threads, slots, node_walk );
}
/* general: the map has node planes that are not axis records */
TraceKernel pick_kernel( int mode, int depth, int threads, int sched, int slots, bool general ) {
#ifdef BSPGPU_EXPERIMENTS
	if( sched == 4 ) {
		if( depth <= 32 ) return threads <= 256 ? trace_simple< kGlobal, 32, 256 > : trace_simple< kGlobal, 32, 1024 >;
		return threads <= 256 ? trace_simple< kGlobal, 128, 256 > : trace_simple< kGlobal, 128, 1024 >;
	}
	if( mode == kSmemTop ) return pick_stack< kSmemTop >( depth, threads, slots, general, true );
#endif
	return mode == kSmemAll ? pick_stack< kSmemAll >( depth, threads, slots, general, false ) : pick_stack< kGlobal >( depth, threads, slots, general, sched == 1 );
}

#ifdef BSPGPU_EXPERIMENTS
/* BSPGPU_OPT_SCHEDULER = 5: CTA-level ray queues (pool_kernels.cuh; measured slower than the shipped kernel, profiles/README.md).  Shapes: threads per CTA, CTAs per SM, pool slots, stack slots in shared memory. */
struct PoolShape { int threads, rays, stack_slots; uint32_t pool_bytes; TraceKernel kernel[ 2 ]; };
template< int M, int T, int B, int R, int S >
PoolShape pool_shape() { return PoolShape{ T, R, S, PoolLayout< R, S >::kBytes, { trace_pool< M, T, B, R, S, false >, trace_pool< M, T, B, R, S, true > } }; }
const PoolShape * pick_pool( int mode, uint64_t staged, uint64_t smem_cap, int64_t want_threads ) {
	static const PoolShape global_shapes[] = { pool_shape< kGlobal, 256, 4, 320, 4 >(), pool_shape< kGlobal, 256, 3, 384, 6 >() };
	static const PoolShape smem_shapes[] = { pool_shape< kSmemAll, 768, 1, 960, 4 >(), pool_shape< kSmemAll, 512, 1, 640, 4 >() };
	if( mode == kGlobal ) return &global_shapes[ want_threads == 384 ? 1 : 0 ];
	for( const PoolShape & sh : smem_shapes )
		if( ( ( staged + 127u ) & ~( uint64_t ) 127u ) + sh.pool_bytes <= smem_cap && ( want_threads <= 0 || want_threads >= sh.threads ) ) return &sh;
	return nullptr;
}
#endif

} // namespace

namespace {

int ensure_scratch( bspgpu * h, size_t bytes ) {
	if( h->scratch_bytes >= bytes ) return BSPGPU_OK;
	if( h->d_scratch ) { CK( cudaStreamSynchronize( h->stream() ) ); CK( cudaStreamSynchronize( h->own_stream ) ); cudaFree( h->d_scratch ); h->d_scratch = nullptr; h->scratch_bytes = 0; }
	bytes = ( bytes + ( 1u << 20 ) - 1 ) & ~( size_t ) ( ( 1u << 20 ) - 1 );
	CK( cudaMalloc( &h->d_scratch, bytes ) );
	h->scratch_bytes = bytes;
	return BSPGPU_OK;
}

int configure( bspgpu * h ) {
	int variant = ( int ) h->opt_variant;
	const uint64_t smem_cap = ( uint64_t ) h->max_smem_optin > 1024 ? ( uint64_t ) h->max_smem_optin - 1024 : 0;
	/* AUTO (measured on B200, profiles/): the whole map in shared memory when it fits; otherwise the read-only
	 * global path -- for a few-MB map L1 already keeps the top of the tree resident and staging it buys nothing */
	if( variant == BSPGPU_VARIANT_AUTO ) variant = h->lay.staged_bytes <= smem_cap ? BSPGPU_VARIANT_SMEM : BSPGPU_VARIANT_GLOBAL;
#ifdef BSPGPU_EXPERIMENTS
	if( h->opt_sched == 4 ) variant = BSPGPU_VARIANT_GLOBAL;   /* the one-thread-per-segment kernel stages nothing */
#else
	if( h->opt_sched != 0 ) return fail( BSPGPU_E_INVALID, "BSPGPU_OPT_SCHEDULER: only the phase-scheduled kernel (0) is in this build (others need -DBSPGPU_EXPERIMENTS)" );
	if( variant == BSPGPU_VARIANT_TOPTREE ) return fail( BSPGPU_E_INVALID, "BSPGPU_VARIANT_TOPTREE needs a -DBSPGPU_EXPERIMENTS build" );
#endif
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
	else mode = kGlobal;

	int threads = ( int ) h->opt_threads;
	if( threads <= 0 ) threads = ( mode == kSmemAll ) ? 1024 : 256;   /* measured: 1 x 1024 with the map in smem, 5 x 256 otherwise */
	threads = threads <= 256 ? 256 : ( threads <= 512 ? 512 : 1024 );  /* the instantiated CTA shapes; smaller requests run as 256 */
	const int launch_threads = h->opt_threads > 0 && h->opt_threads < 256 ? ( int ) ( ( h->opt_threads + 31 ) / 32 * 32 ) : threads;

	const uint32_t staged = smem;
#ifdef BSPGPU_EXPERIMENTS
	if( h->opt_sched == 5 ) {
		if( mode != kGlobal && mode != kSmemAll ) return fail( BSPGPU_E_INVALID, "the queue scheduler runs the SMEM and GLOBAL variants" );
		if( h->lay.tree_depth > kPoolOverflowDepth ) return fail( BSPGPU_E_INVALID, "the queue scheduler holds traversal stacks of up to 32 entries" );
		const PoolShape * sh = pick_pool( mode, staged, smem_cap, h->opt_threads );
		if( !sh ) return fail( BSPGPU_E_INVALID, "the queue scheduler's ray pool does not fit in shared memory beside the staged map" );
		TraceKernel kernel = sh->kernel[ h->lay.num_gplanes > 0 ? 1 : 0 ];
		smem = ( ( staged + 127u ) & ~127u ) + sh->pool_bytes;
		CK( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) smem ) );
		int occ = 0;
		CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, kernel, sh->threads, smem ) );
		if( occ < 1 ) return fail( BSPGPU_E_CUDA, "queue-scheduled kernel does not fit on an SM" );
		int ctas = ( int ) h->opt_ctas;
		if( ctas <= 0 || ctas > occ ) ctas = occ;
		const size_t ovf = ( size_t ) h->sm_count * ctas * sh->rays * kPoolScratchPerRay * sizeof( float4 );
		if( h->pool_ovf_bytes < ovf ) {
			if( h->d_pool_ovf ) { CK( cudaDeviceSynchronize() ); cudaFree( h->d_pool_ovf ); h->d_pool_ovf = nullptr; h->pool_ovf_bytes = 0; }
			CK( cudaMalloc( &h->d_pool_ovf, ovf ) );
			h->pool_ovf_bytes = ovf;
		}
		h->cfg_kernel = kernel; h->cfg_variant = variant; h->cfg_mode = mode; h->cfg_threads = sh->threads; h->cfg_slots = sh->stack_slots; h->cfg_pool_rays = sh->rays;
		h->cfg_grid = h->sm_count * ctas; h->cfg_smem = smem; h->cfg_staged = staged; h->cfg_top_nodes = 0;
		h->cfg_valid = true;
		return BSPGPU_OK;
	}
#endif
	h->cfg_pool_rays = 0;
	int slots = 0;
	int64_t want = h->opt_smem_stack;
	if( want < 0 ) want = mode == kSmemAll ? 0 : 6;   /* measured: 4 -> +10 %, 6 -> +17 %, 8 -> +10 % on city10k long rays; +-0 behind a staged map */
	if( want > 0 && ( h->opt_sched == 0 || h->opt_sched == 1 ) ) {
		slots = effective_slots( mode, threads, want );
		smem = ( smem + 127u ) & ~127u;
		while( slots > 0 && ( uint64_t ) smem + ( uint64_t ) slots * ( uint32_t ) threads * 16u > smem_cap )   /* smem_cap leaves room for static shared memory */
			slots = full_slot_range( mode, threads ) ? slots - 2 : 0;
		smem += ( uint32_t ) slots * ( uint32_t ) threads * 16u;
	}
	TraceKernel kernel = pick_kernel( mode, ( int ) h->lay.tree_depth, threads, ( int ) h->opt_sched, slots, h->lay.num_gplanes > 0 );
	CK( cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ( int ) smem ) );
	int ctas = ( int ) h->opt_ctas;
	int occ = 0;
	CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, kernel, launch_threads, smem ) );
	if( occ < 1 ) return fail( BSPGPU_E_CUDA, "kernel does not fit on an SM with the requested block size / shared memory" );
	if( ctas <= 0 || ctas > occ ) ctas = occ;
	h->cfg_kernel = kernel; h->cfg_variant = variant; h->cfg_mode = mode; h->cfg_threads = launch_threads; h->cfg_slots = slots;
	h->cfg_grid = h->sm_count * ctas * ( h->opt_sched == 4 ? 4 : 1 ); h->cfg_smem = smem; h->cfg_staged = staged; h->cfg_top_nodes = top_nodes;
	h->cfg_valid = true;
	return BSPGPU_OK;
}

/* the part of the kernel parameters that describes the map image */
void map_params( bspgpu * h, TraceParams & p ) {
	memset( &p, 0, sizeof( p ) );
	p.image = h->d_image;
	p.off_gplanes = h->lay.off_gplanes; p.off_refs = h->lay.off_refs; p.off_sides_even = h->lay.off_sides_even; p.off_sides_odd = h->lay.off_sides_odd;
	p.off_side_pairs = h->lay.off_side_pairs; p.off_side_plane = h->lay.off_side_plane;
	p.hits = h->d_counters;
	p.grid.d = h->lay.grid;
	p.root = h->lay.root;
	p.g_slot_pairs = reinterpret_cast< const uint4 * >( h->d_image + h->lay.off_slots );
	p.g_gplanes = reinterpret_cast< const PlaneRec * >( h->d_image + h->lay.off_gplanes );
	p.g_refs = reinterpret_cast< const BrushRef * >( h->d_image + h->lay.off_refs );
	p.g_side_pairs = reinterpret_cast< const SideRec * >( h->d_image + h->lay.off_side_pairs );
	p.g_side_plane = reinterpret_cast< const uint32_t * >( h->d_image + h->lay.off_side_plane );
}

/* results go to d_out as 16-byte records (compact 0), 8-byte records (1: BSPGPU_OUT_COMPACT8) or one float each (2: BSPGPU_OUT_COMPACT4) */
int enqueue_trace( bspgpu * h, const void * d_segs, bool packed24, uint64_t n, void * d_out, int compact, bspgpu_pos * d_pos, const uint32_t * d_perm, cudaStream_t st ) {
	if( !h->cfg_valid ) { const int rc = configure( h ); if( rc != BSPGPU_OK ) return rc; }
	TraceKernel kernel = h->cfg_kernel;
	const int mode = h->cfg_mode, threads = h->cfg_threads;
	const uint32_t smem = h->cfg_smem;

	TraceParams p;
	map_params( h, p );
	p.staged_bytes = h->cfg_staged; p.top_nodes = h->cfg_top_nodes;
	p.refill = ( uint32_t ) h->opt_refill;
	p.max_steps = ( uint32_t ) ( h->opt_max_steps < 1 ? ( mode == kSmemAll ? 16 : 8 ) : h->opt_max_steps );
	p.max_steps_long = ( uint32_t ) ( h->opt_max_steps < 1 ? 10 : h->opt_max_steps );   /* GLOBAL variant, warps whose segments start at the root */
	p.side_steps = ( uint32_t ) h->opt_side_steps;
	p.leaf_threshold = ( uint32_t ) h->opt_leaf_threshold;
	p.exotic_list = h->d_exotic;
	p.pool_overflow = reinterpret_cast< float4 * >( h->d_pool_ovf );
	const bool use_grid = h->opt_start_grid < 0 ? mode != kSmemAll : h->opt_start_grid != 0;
	p.grid.cells = ( use_grid && h->lay.grid.dim[ 0 ] ) ? reinterpret_cast< const int32_t * >( h->d_image + h->lay.off_grid ) : nullptr;
	p.grid_slots = p.grid.cells ? reinterpret_cast< const Slot * >( h->d_image + h->lay.off_grid_slots ) : nullptr;

	h->last_variant = ( uint32_t ) h->cfg_variant; h->last_threads = ( uint32_t ) threads; h->last_smem = smem; h->last_slots = ( uint32_t ) h->cfg_slots;

	const uint32_t warps_per_cta = ( uint32_t ) threads / 32u;
	for( uint64_t done = 0; done < n; done += h->opt_max_launch ) {
		const uint64_t cnt = n - done < h->opt_max_launch ? n - done : h->opt_max_launch;
		/* small launches: warps claim 32 segments at a time and only as many CTAs start as there is work for --
		 * a batch of one (BSP::trace_seg) is one CTA, not the whole persistent grid */
		uint32_t block_segs = ( uint32_t ) h->opt_block_segs;
		if( cnt < ( uint64_t ) h->cfg_grid * warps_per_cta * block_segs ) block_segs = 32;
		uint64_t grid = ( cnt + ( uint64_t ) block_segs * warps_per_cta - 1 ) / ( ( uint64_t ) block_segs * warps_per_cta );
		if( grid > ( uint64_t ) h->cfg_grid ) grid = ( uint64_t ) h->cfg_grid;
		if( h->opt_sched == 4 ) grid = ( uint64_t ) h->cfg_grid;
		unsigned long long * next = h->d_counters + 2 + 2 * ( h->launch_seq++ % kCounterSlots );   /* {work counter, exotic-ray count} */
		CK( cudaMemsetAsync( next, 0, 2 * sizeof( unsigned long long ), st ) );
		p.block_segs = block_segs;
		p.perm = d_perm ? d_perm + done : nullptr;
		/* with a permutation every slot names an absolute segment index, so the arrays are not advanced */
		const uint64_t base = d_perm ? 0 : done;
		p.segs = packed24 ? nullptr : reinterpret_cast< const float4 * >( d_segs ) + base * 2;
		p.segs24 = packed24 ? reinterpret_cast< const float2 * >( d_segs ) + base * 3 : nullptr;
		p.out = ( d_out && compact == 0 ) ? reinterpret_cast< ResultRec * >( d_out ) + base : nullptr;
		p.out8 = ( d_out && compact == 1 ) ? reinterpret_cast< uint2 * >( d_out ) + base : nullptr;
		p.out4 = ( d_out && compact == 2 ) ? reinterpret_cast< float * >( d_out ) + base : nullptr;
		p.pos = d_pos ? reinterpret_cast< float4 * >( d_pos ) + base : nullptr;
		p.n = ( uint32_t ) cnt;
		p.next = next;
		kernel<<< ( unsigned ) grid, threads, smem, st >>>( p );
		/* rays the axis shortcut does not cover (zero start coordinate, non-finite component): counted by the kernel above, traced
		 * here; returns immediately when there were none */
		uint64_t xgrid = ( cnt + 255 ) / 256;
		if( xgrid > ( uint64_t ) h->sm_count * 2 ) xgrid = ( uint64_t ) h->sm_count * 2;   /* grid-stride; almost always there is nothing to do */
		trace_exotic_kernel<<< ( unsigned ) xgrid, 256, 0, st >>>( p );
		CK( cudaGetLastError() );
		h->last_grid = ( uint32_t ) grid;
		h->last_launches += 2;
	}
	return BSPGPU_OK;
}

/* N4: counting sort of the segments of one device-resident batch by the start-grid cell of their start point.
 * Leaves the permutation in the handle's scratch; *d_perm = null when the map has no grid. */
int enqueue_binning( bspgpu * h, const void * d_segs, bool packed24, uint64_t n, const uint32_t ** d_perm, cudaStream_t st ) {
	*d_perm = nullptr;
	if( !h->lay.grid.dim[ 0 ] || n < 2 || n > kMaxLaunchSegs ) return BSPGPU_OK;
	BinParams b;
	memset( &b, 0, sizeof( b ) );
	b.grid = h->lay.grid;
	for( int a = 0; a < 3; a++ ) { b.shift[ a ] = 0; b.bdim[ a ] = ( uint32_t ) h->lay.grid.dim[ a ]; }
	/* coarsen the start grid until at most 2^16 bins remain (the scan is one CTA) */
	while( ( uint64_t ) b.bdim[ 0 ] * b.bdim[ 1 ] * b.bdim[ 2 ] > 65536 ) {
		int a = 0;
		if( b.bdim[ 1 ] > b.bdim[ a ] ) a = 1;
		if( b.bdim[ 2 ] > b.bdim[ a ] ) a = 2;
		b.shift[ a ]++; b.bdim[ a ] = ( b.bdim[ a ] + 1 ) / 2;
	}
	b.num_bins = b.bdim[ 0 ] * b.bdim[ 1 ] * b.bdim[ 2 ];
	const size_t bins_bytes = ( ( size_t ) b.num_bins * 4 + 255 ) & ~( size_t ) 255;
	{ const int rc = ensure_scratch( h, bins_bytes + n * 4 ); if( rc != BSPGPU_OK ) return rc; }
	b.bins = reinterpret_cast< uint32_t * >( h->d_scratch );
	b.perm = reinterpret_cast< uint32_t * >( ( char * ) h->d_scratch + bins_bytes );
	b.segs = packed24 ? nullptr : reinterpret_cast< const float4 * >( d_segs );
	b.segs24 = packed24 ? reinterpret_cast< const float2 * >( d_segs ) : nullptr;
	b.n = ( uint32_t ) n;
	CK( cudaEventRecord( h->ev_pre0, st ) );
	CK( cudaMemsetAsync( b.bins, 0, ( size_t ) b.num_bins * 4, st ) );
	const int threads = 256;
	uint64_t blocks = ( n + threads - 1 ) / threads;
	if( blocks > ( uint64_t ) h->sm_count * 16 ) blocks = ( uint64_t ) h->sm_count * 16;
	bin_count_kernel<<< ( unsigned ) blocks, threads, 0, st >>>( b );
	bin_scan_kernel<<< 1, 1024, 0, st >>>( b );
	bin_scatter_kernel<<< ( unsigned ) blocks, threads, 0, st >>>( b );
	CK( cudaGetLastError() );
	CK( cudaEventRecord( h->ev_pre1, st ) );
	h->prepass_valid = true;
	*d_perm = b.perm;
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

/* Pageable caller arrays (what an existing caller of the reference has: std::vector, new[]) cannot be the source or target of an
 * asynchronous copy -- the driver stages them through one small pinned buffer of its own, synchronously -- and page-locking them
 * for one call costs more than the call (cudaHostRegister touches every page).  So the library keeps a pinned bounce ring, one
 * slot per device staging buffer, and a few host threads that copy chunk c+1 from the caller's array into the ring and chunk c-2
 * out of it while the GPU works on the chunks in between. */
int ensure_ring( bspgpu * h, size_t segs ) {
	if( h->ring_segs >= segs ) return BSPGPU_OK;
	for( int b = 0; b < kBufs; b++ ) {
		if( h->h_ring_in[ b ] ) cudaFreeHost( h->h_ring_in[ b ] );
		if( h->h_ring_out[ b ] ) cudaFreeHost( h->h_ring_out[ b ] );
		if( h->h_ring_pos[ b ] ) cudaFreeHost( h->h_ring_pos[ b ] );
		h->h_ring_in[ b ] = h->h_ring_out[ b ] = h->h_ring_pos[ b ] = nullptr;
	}
	h->ring_segs = 0;
	for( int b = 0; b < kBufs; b++ ) {
		CK( cudaMallocHost( &h->h_ring_in[ b ], segs * sizeof( bspgpu_segment ) ) );
		CK( cudaMallocHost( &h->h_ring_out[ b ], segs * sizeof( bspgpu_result ) ) );
		CK( cudaMallocHost( &h->h_ring_pos[ b ], segs * sizeof( bspgpu_pos ) ) );
	}
	h->ring_segs = segs;
	if( !h->pool ) {
		unsigned hw = std::thread::hardware_concurrency();
		int n = hw >= 16 ? 8 : ( hw >= 4 ? ( int ) hw / 2 : 1 );
		h->pool = new CopyPool( n );
	}
	return BSPGPU_OK;
}

int trace_bounced( bspgpu * h, const void * segs, uint64_t n, void * out, bspgpu_pos * pos, bool packed, int compact, cudaStream_t st ) {
	const size_t seg_bytes = packed ? 24 : sizeof( bspgpu_segment );
	const size_t out_bytes = compact == 2 ? sizeof( float ) : ( compact == 1 ? sizeof( bspgpu_result8 ) : sizeof( bspgpu_result ) );
	uint64_t chunk = ( uint64_t ) h->opt_chunk;
	if( chunk > n ) chunk = n;
	{ const int rc = ensure_staging( h, ( size_t ) ( chunk < kSmallCall ? kSmallCall : chunk ), true, out != nullptr, pos != nullptr ); if( rc != BSPGPU_OK ) return rc; }
	{ const int rc = ensure_ring( h, ( size_t ) chunk ); if( rc != BSPGPU_OK ) return rc; }
	const uint64_t chunks = ( n + chunk - 1 ) / chunk;
	CK( cudaEventRecord( h->ev_start, st ) );
	CK( cudaStreamWaitEvent( h->copy_in, h->ev_start, 0 ) );
	/* iteration c: host copies chunk c into the ring and chunk c-2 out of it (in parallel), then queues chunk c on the GPU */
	for( uint64_t c = 0; c < chunks + 2; c++ ) {
		const int b = ( int ) ( c % kBufs );
		const uint64_t done = c * chunk, cnt = c < chunks ? ( n - done < chunk ? n - done : chunk ) : 0;
		if( c < chunks ) {
			if( c >= ( uint64_t ) kBufs ) CK( cudaEventSynchronize( h->ev_in[ b ] ) );                /* the H2D copy that last read this ring slot */
			h->pool->copy( h->h_ring_in[ b ], ( const char * ) segs + done * seg_bytes, cnt * seg_bytes );
		}
		if( c >= 2 ) {
			const uint64_t d = c - 2, ddone = d * chunk, dcnt = n - ddone < chunk ? n - ddone : chunk;
			const int db = ( int ) ( d % kBufs );
			CK( cudaEventSynchronize( h->ev_out[ db ] ) );                                             /* chunk d's results are in the ring */
			if( out ) h->pool->copy( ( char * ) out + ddone * out_bytes, h->h_ring_out[ db ], dcnt * out_bytes );
			if( pos ) h->pool->copy( pos + ddone, h->h_ring_pos[ db ], dcnt * sizeof( bspgpu_pos ) );
		}
		h->pool->wait();
		if( c < chunks ) {
			if( c >= ( uint64_t ) kBufs ) CK( cudaStreamWaitEvent( h->copy_in, h->ev_k[ b ], 0 ) );    /* kernel that last read d_in[b] */
			CK( cudaMemcpyAsync( h->d_in[ b ], h->h_ring_in[ b ], cnt * seg_bytes, cudaMemcpyHostToDevice, h->copy_in ) );
			CK( cudaEventRecord( h->ev_in[ b ], h->copy_in ) );
			CK( cudaStreamWaitEvent( st, h->ev_in[ b ], 0 ) );
			/* d_out[b] / the ring slot were last used by chunk c-3, whose results left the ring in iteration c-1 */
			const int rc = enqueue_trace( h, h->d_in[ b ], packed, cnt, out ? h->d_out[ b ] : nullptr, compact, pos ? ( bspgpu_pos * ) h->d_pos[ b ] : nullptr, nullptr, st );
			if( rc != BSPGPU_OK ) return rc;
			CK( cudaEventRecord( h->ev_k[ b ], st ) );
			CK( cudaStreamWaitEvent( h->copy_out, h->ev_k[ b ], 0 ) );
			if( out ) CK( cudaMemcpyAsync( h->h_ring_out[ b ], h->d_out[ b ], cnt * out_bytes, cudaMemcpyDeviceToHost, h->copy_out ) );
			if( pos ) CK( cudaMemcpyAsync( h->h_ring_pos[ b ], h->d_pos[ b ], cnt * sizeof( bspgpu_pos ), cudaMemcpyDeviceToHost, h->copy_out ) );
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

/* the hot path behind bspgpu_trace / _pos / _stream; `st` is the stream this call runs on */
int trace_impl( bspgpu * h, const void * segs, uint64_t n, void * out, bspgpu_pos * pos, unsigned flags, cudaStream_t st, bool foreign_stream ) {
	if( n && ( !segs || ( !out && !pos ) ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null segs, or both out and pos null" );
	const bool segs_dev = ( flags & BSPGPU_SEGS_DEVICE ) != 0, out_dev = ( flags & BSPGPU_OUT_DEVICE ) != 0;
	const bool packed = ( flags & BSPGPU_SEGS_PACKED24 ) != 0;
	if( ( flags & BSPGPU_OUT_COMPACT8 ) && ( flags & BSPGPU_OUT_COMPACT4 ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: BSPGPU_OUT_COMPACT8 and BSPGPU_OUT_COMPACT4 exclude each other" );
	const int compact = ( flags & BSPGPU_OUT_COMPACT4 ) ? 2 : ( ( flags & BSPGPU_OUT_COMPACT8 ) ? 1 : 0 );
	if( ( flags & BSPGPU_ASYNC ) && !( segs_dev && out_dev ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: BSPGPU_ASYNC needs device pointers" );
	/* the kernels read segments with 128-bit (packed: 64-bit) loads and write 16-byte (compact: 8-byte) records */
	if( segs_dev && ( ( uintptr_t ) segs % ( packed ? 8 : 16 ) ) != 0 ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: device segments must be 16-byte aligned (8-byte for the packed layout)" );
	if( out_dev && ( ( ( uintptr_t ) out % ( compact == 2 ? 4 : ( compact == 1 ? 8 : 16 ) ) ) != 0 || ( ( uintptr_t ) pos % 16 ) != 0 ) ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: device results must be 16-byte aligned (8-byte for compact records, 4-byte for BSPGPU_OUT_COMPACT4)" );
	DeviceGuard guard( h->device );
	/* work of an earlier call that went to a caller's stream shares this handle's counters and buffers: wait for it */
	if( h->user_pending && !( foreign_stream && st == h->pending_stream ) ) { CK( cudaStreamWaitEvent( st, h->ev_user, 0 ) ); h->user_pending = false; }
	h->last_launches = 0;
	h->timing_valid = false; h->prepass_valid = false;

	/* ---- a handful of host segments (a batch of one = BSP::trace_seg, the reference's only call shape): ONE kernel that reads
	 *      the segments from and writes the results to the pinned buffer over the link, stores the hit count, and nothing else
	 *      on the stream (trace_kernels.cuh: trace_tiny_kernel) */
	if( n > 0 && n <= kTinyCall && !segs_dev && !out_dev && h->opt_tiny ) {
		/* options that cannot be honoured (a variant that does not fit, a scheduler this build lacks) fail here as on every other path */
		if( !h->cfg_valid ) { const int rc = configure( h ); if( rc != BSPGPU_OK ) return rc; }
		unsigned char * hs = h->h_small, * ho = hs + kSmallCall * 32, * hp = ho + kSmallCall * 16;
		memcpy( hs, segs, n * ( packed ? 24 : sizeof( bspgpu_segment ) ) );
		TraceParams p;
		map_params( h, p );
		p.segs = packed ? nullptr : reinterpret_cast< const float4 * >( hs );
		p.segs24 = packed ? reinterpret_cast< const float2 * >( hs ) : nullptr;
		p.out = ( out && compact == 0 ) ? reinterpret_cast< ResultRec * >( ho ) : nullptr;
		p.out8 = ( out && compact == 1 ) ? reinterpret_cast< uint2 * >( ho ) : nullptr;
		p.out4 = ( out && compact == 2 ) ? reinterpret_cast< float * >( ho ) : nullptr;
		p.pos = pos ? reinterpret_cast< float4 * >( hp ) : nullptr;
		p.n = ( uint32_t ) n;
		trace_tiny_kernel<<< 1, ( unsigned ) n * 32u, 0, st >>>( p );
		CK( cudaGetLastError() );
		CK( cudaStreamSynchronize( st ) );
		const size_t ob = compact == 2 ? sizeof( float ) : ( compact == 1 ? sizeof( bspgpu_result8 ) : sizeof( bspgpu_result ) );
		if( out ) memcpy( out, ho, n * ob );
		if( pos ) memcpy( pos, hp, n * sizeof( bspgpu_pos ) );
		h->last_variant = BSPGPU_VARIANT_GLOBAL; h->last_threads = ( uint32_t ) n * 32u; h->last_smem = 0; h->last_slots = 0; h->last_grid = 1; h->last_launches = 1;
		return BSPGPU_OK;
	}

	CK( cudaMemsetAsync( h->d_counters, 0, sizeof( unsigned long long ), st ) );
	if( n == 0 ) return BSPGPU_OK;
	const size_t seg_bytes = packed ? 24 : sizeof( bspgpu_segment );
	const size_t out_bytes = compact == 2 ? sizeof( float ) : ( compact == 1 ? sizeof( bspgpu_result8 ) : sizeof( bspgpu_result ) );

	if( segs_dev && out_dev ) {
		const uint32_t * d_perm = nullptr;
		if( h->opt_bin > 0 ) { const int rc = enqueue_binning( h, segs, packed, n, &d_perm, st ); if( rc != BSPGPU_OK ) return rc; }
		CK( cudaEventRecord( h->ev_start, st ) );
		const int rc = enqueue_trace( h, segs, packed, n, out, compact, pos, d_perm, st );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaEventRecord( h->ev_stop, st ) );
		h->timing_valid = true;
		if( foreign_stream ) { CK( cudaEventRecord( h->ev_user, st ) ); h->user_pending = true; h->pending_stream = st; }
		if( !( flags & BSPGPU_ASYNC ) ) CK( cudaStreamSynchronize( st ) );
		return BSPGPU_OK;
	}

	/* ---- small host-pointer calls (a batch of one = BSP::trace_seg): one pinned bounce buffer, one stream, one synchronise */
	if( n <= kSmallCall && !segs_dev && !out_dev ) {
		unsigned char * hs = h->h_small, * ho = hs + kSmallCall * 32, * hp = ho + kSmallCall * 16;
		{ const int rc = ensure_staging( h, ( size_t ) kSmallCall, true, out != nullptr, pos != nullptr ); if( rc != BSPGPU_OK ) return rc; }
		memcpy( hs, segs, n * seg_bytes );
		CK( cudaMemcpyAsync( h->d_in[ 0 ], hs, n * seg_bytes, cudaMemcpyHostToDevice, st ) );
		CK( cudaEventRecord( h->ev_start, st ) );
		const int rc = enqueue_trace( h, h->d_in[ 0 ], packed, n, out ? h->d_out[ 0 ] : nullptr, compact, pos ? ( bspgpu_pos * ) h->d_pos[ 0 ] : nullptr, nullptr, st );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaEventRecord( h->ev_stop, st ) );
		h->timing_valid = true;
		if( out ) CK( cudaMemcpyAsync( ho, h->d_out[ 0 ], n * out_bytes, cudaMemcpyDeviceToHost, st ) );
		if( pos ) CK( cudaMemcpyAsync( hp, h->d_pos[ 0 ], n * sizeof( bspgpu_pos ), cudaMemcpyDeviceToHost, st ) );
		CK( cudaStreamSynchronize( st ) );
		if( out ) memcpy( out, ho, n * out_bytes );
		if( pos ) memcpy( pos, hp, n * sizeof( bspgpu_pos ) );
		return BSPGPU_OK;
	}

	/* ---- at least one side lives in host memory: chunked 3-stream pipeline.  Pageable caller memory would make every
	 *      cudaMemcpyAsync a staged, host-synchronous copy and serialise the pipeline: page-lock it for the call. */
	if( h->opt_pageable == 2 && !segs_dev && !out_dev && n * seg_bytes >= kPinThresholdBytes
	    && is_pageable_host( segs ) && ( !out || is_pageable_host( out ) ) && ( !pos || is_pageable_host( pos ) ) )
		return trace_bounced( h, segs, n, out, pos, packed, compact, st );
	HostPin pin_in, pin_out, pin_pos;
	if( h->opt_pageable > 0 ) {
		if( !segs_dev && n * seg_bytes >= kPinThresholdBytes && is_pageable_host( segs ) ) pin_in.pin( segs, n * seg_bytes, true );
		if( !out_dev && out && n * out_bytes >= kPinThresholdBytes && is_pageable_host( out ) ) pin_out.pin( out, n * out_bytes, false );
		if( !out_dev && pos && n * sizeof( bspgpu_pos ) >= kPinThresholdBytes && is_pageable_host( pos ) ) pin_pos.pin( pos, n * sizeof( bspgpu_pos ), false );
	}
	uint64_t chunk = ( uint64_t ) h->opt_chunk;
	if( chunk > n ) chunk = n;
	if( chunk > kMaxLaunchSegs ) chunk = kMaxLaunchSegs;
	{
		const int rc = ensure_staging( h, ( size_t ) ( chunk < kSmallCall ? kSmallCall : chunk ), !segs_dev, out && !out_dev, pos && !out_dev );
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
		void * k_out = out ? ( out_dev ? ( void * ) ( ( char * ) out + done * out_bytes ) : h->d_out[ b ] ) : nullptr;
		bspgpu_pos * k_pos = pos ? ( out_dev ? pos + done : ( bspgpu_pos * ) h->d_pos[ b ] ) : nullptr;
		if( !out_dev && c >= ( uint64_t ) kBufs ) CK( cudaStreamWaitEvent( st, h->ev_out[ b ], 0 ) );   /* D2H that last read d_out[b] */
		const int rc = enqueue_trace( h, d_segs, packed, cnt, k_out, compact, k_pos, nullptr, st );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaEventRecord( h->ev_k[ b ], st ) );
		if( !out_dev ) {
			CK( cudaStreamWaitEvent( h->copy_out, h->ev_k[ b ], 0 ) );
			if( out ) CK( cudaMemcpyAsync( ( char * ) out + done * out_bytes, h->d_out[ b ], cnt * out_bytes, cudaMemcpyDeviceToHost, h->copy_out ) );
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

int upload_vis( bspgpu * h, const void * vis_lump, uint64_t len ) {
	if( !vis_lump || len < 8 ) return fail( BSPGPU_E_MAP, "visibility lump shorter than its 8-byte header (the reference asserts, bsp.cc:111)" );
	uint32_t hdr[ 2 ];
	memcpy( hdr, vis_lump, 8 );
	if( ( uint64_t ) hdr[ 0 ] * hdr[ 1 ] + 8 > len ) return fail( BSPGPU_E_MAP, "visibility lump shorter than 8 + num_clusters * cluster_size (bsp.cc:111)" );
	DeviceGuard guard( h->device );
	if( h->d_vis ) { CK( cudaStreamSynchronize( h->stream() ) ); cudaFree( h->d_vis ); h->d_vis = nullptr; }
	h->num_clusters = hdr[ 0 ]; h->cluster_size = hdr[ 1 ];
	const size_t bytes = ( size_t ) hdr[ 0 ] * hdr[ 1 ];
	if( bytes ) {
		CK( cudaMalloc( &h->d_vis, bytes ) );
		CK( cudaMemcpy( h->d_vis, ( const char * ) vis_lump + 8, bytes, cudaMemcpyHostToDevice ) );
	}
	return BSPGPU_OK;
}

/* batched position_to_leaf (+ cluster, + PVS bit) through the handle's grow-only scratch */
int leaf_impl( bspgpu * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, int32_t * cluster, int32_t from_cluster, uint8_t * visible, unsigned flags ) {
	if( n == 0 ) return BSPGPU_OK;
	if( visible && from_cluster >= 0 && ( !h->d_vis || ( uint32_t ) from_cluster >= h->num_clusters ) )
		return fail( BSPGPU_E_INVALID, "bspgpu_leaf_clusters: from_cluster out of range or no visibility lump uploaded (bspgpu_set_vis)" );
	DeviceGuard guard( h->device );
	cudaStream_t st = h->stream();
	if( h->user_pending ) { CK( cudaStreamWaitEvent( st, h->ev_user, 0 ) ); h->user_pending = false; }
	const bool pts_dev = ( flags & BSPGPU_SEGS_DEVICE ) != 0, out_dev = ( flags & BSPGPU_OUT_DEVICE ) != 0;
	const size_t pts_bytes = ( ( size_t ) n * stride * sizeof( float ) + 255 ) & ~( size_t ) 255, i_bytes = ( ( size_t ) n * 4 + 255 ) & ~( size_t ) 255, v_bytes = ( ( size_t ) n + 255 ) & ~( size_t ) 255;
	size_t need = 0;
	if( !pts_dev ) need += pts_bytes;
	if( !out_dev ) need += ( leaf ? i_bytes : 0 ) + ( cluster ? i_bytes : 0 ) + ( visible ? v_bytes : 0 );
	{ const int rc = ensure_scratch( h, need ); if( rc != BSPGPU_OK ) return rc; }
	char * sp = ( char * ) h->d_scratch;
	LeafParams p;
	memset( &p, 0, sizeof( p ) );
	p.image = h->d_image; p.off_gplanes = h->lay.off_gplanes; p.off_node_orig = h->lay.off_node_orig; p.off_leaf_cluster = h->lay.off_leaf_cluster;
	p.stride = stride; p.n = n;
	p.vis_rows = h->d_vis; p.num_clusters = h->num_clusters; p.cluster_size = h->cluster_size; p.from_cluster = from_cluster;
	if( pts_dev ) p.pts = pts;
	else {
		p.pts = ( const float * ) sp; sp += pts_bytes;
		CK( cudaMemcpyAsync( ( void * ) p.pts, pts, ( size_t ) n * stride * sizeof( float ), cudaMemcpyHostToDevice, st ) );
	}
	if( out_dev ) { p.leaf_out = leaf; p.cluster_out = cluster; p.visible_out = visible; }
	else {
		if( leaf ) { p.leaf_out = ( int32_t * ) sp; sp += i_bytes; }
		if( cluster ) { p.cluster_out = ( int32_t * ) sp; sp += i_bytes; }
		if( visible ) { p.visible_out = ( unsigned char * ) sp; sp += v_bytes; }
	}
	const int threads = 256;
	uint64_t blocks = ( n + threads - 1 ) / threads;
	if( blocks > ( uint64_t ) h->sm_count * 16 ) blocks = ( uint64_t ) h->sm_count * 16;
	leaf_kernel<<< ( unsigned ) blocks, threads, 0, st >>>( p );
	CK( cudaGetLastError() );
	if( !out_dev ) {
		if( leaf ) CK( cudaMemcpyAsync( leaf, p.leaf_out, ( size_t ) n * 4, cudaMemcpyDeviceToHost, st ) );
		if( cluster ) CK( cudaMemcpyAsync( cluster, p.cluster_out, ( size_t ) n * 4, cudaMemcpyDeviceToHost, st ) );
		if( visible ) CK( cudaMemcpyAsync( visible, p.visible_out, ( size_t ) n, cudaMemcpyDeviceToHost, st ) );
	}
	/* the scratch is reused by the next call, so wait unless the caller owns every buffer */
	if( !( ( flags & BSPGPU_ASYNC ) && pts_dev && out_dev ) ) CK( cudaStreamSynchronize( st ) );
	return BSPGPU_OK;
}

} // namespace

extern "C" {

const char * bspgpu_last_error( void ) { return bspgpu_err().c_str(); }
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

int bspgpu_create_opts( bspgpu_t ** out, const bspgpu_map_desc * lumps, int device, const bspgpu_create_options * opts ) {
	if( !out || !lumps ) return fail( BSPGPU_E_INVALID, "bspgpu_create: null argument" );
	*out = nullptr;
	FlatMap flat;
	std::string err;
	FlattenOptions fo;
	if( opts ) {
		if( opts->start_grid_cells < -1 || opts->start_grid_cells > 64000000 ) return fail( BSPGPU_E_INVALID, "bspgpu_create_opts: start_grid_cells out of range" );
		fo.grid_cells = opts->start_grid_cells == 0 ? -1 : ( opts->start_grid_cells < 0 ? 0 : opts->start_grid_cells );
		fo.general_planes_only = opts->general_planes_only != 0;
	}
	const int rc = flatten_map( *lumps, fo, flat, err );
	if( rc != BSPGPU_OK ) return fail( rc, "bspgpu_create: " + err );
	if( flat.layout.tree_depth > 128 ) return fail( BSPGPU_E_MAP, "bspgpu_create: tree deeper than 128 levels (the traversal stack holds one entry per pending far side; 128 is the largest instantiated)" );

	int count = 0;
	if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 ) {
		cudaGetLastError();
		return fail( BSPGPU_E_CUDA, "bspgpu_create: no CUDA device (this library has no CPU fallback)" );
	}
	if( device < 0 ) CK( cudaGetDevice( &device ) );
	if( device >= count ) return fail( BSPGPU_E_INVALID, "bspgpu_create: device index out of range" );
	DeviceGuard guard( device );
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
		if( ( e = cudaEventCreate( &h->ev_pre0 ) ) != cudaSuccess ) break;
		if( ( e = cudaEventCreate( &h->ev_pre1 ) ) != cudaSuccess ) break;
		if( ( e = cudaEventCreateWithFlags( &h->ev_user, cudaEventDisableTiming ) ) != cudaSuccess ) break;
		for( int b = 0; b < kBufs && e == cudaSuccess; b++ ) {
			if( ( e = cudaEventCreateWithFlags( &h->ev_in[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
			if( ( e = cudaEventCreateWithFlags( &h->ev_k[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
			if( ( e = cudaEventCreateWithFlags( &h->ev_out[ b ], cudaEventDisableTiming ) ) != cudaSuccess ) break;
		}
		if( e != cudaSuccess ) break;
		if( ( e = cudaMalloc( &h->d_image, flat.layout.total_bytes ) ) != cudaSuccess ) break;
		if( ( e = cudaMemcpy( h->d_image, flat.image.data(), flat.layout.total_bytes, cudaMemcpyHostToDevice ) ) != cudaSuccess ) break;
		if( ( e = cudaMalloc( &h->d_counters, ( 2 + 2 * kCounterSlots ) * sizeof( unsigned long long ) ) ) != cudaSuccess ) break;
		if( ( e = cudaMemset( h->d_counters, 0, ( 2 + 2 * kCounterSlots ) * sizeof( unsigned long long ) ) ) != cudaSuccess ) break;
		if( ( e = cudaMalloc( &h->d_exotic, kExoticCap * sizeof( uint32_t ) ) ) != cudaSuccess ) break;
		if( ( e = cudaMallocHost( ( void ** ) &h->h_small, kSmallCall * 64 ) ) != cudaSuccess ) break;
	} while( 0 );
	if( e != cudaSuccess ) {
		const std::string msg = std::string( "bspgpu_create: " ) + cudaGetErrorString( e );
		bspgpu_destroy( h );
		return fail( BSPGPU_E_CUDA, msg );
	}
	*out = h;
	return BSPGPU_OK;
}

int bspgpu_create( bspgpu_t ** out, const bspgpu_map_desc * lumps, int device ) {
	return bspgpu_create_opts( out, lumps, device, nullptr );
}

int bspgpu_create_from_file( bspgpu_t ** out, const char * path, int device ) {
	if( !out || !path ) return fail( BSPGPU_E_INVALID, "bspgpu_create_from_file: null argument" );
	std::vector< unsigned char > blob;
	bspgpu_map_desc lumps;
	memset( &lumps, 0, sizeof( lumps ) );
	std::string err;
	const void * vis = nullptr; uint64_t vis_len = 0;
	int rc = read_ibsp_file( path, blob, lumps, &vis, &vis_len, err );
	if( rc != BSPGPU_OK ) return fail( rc, "bspgpu_create_from_file: " + err );
	rc = bspgpu_create( out, &lumps, device );
	if( rc != BSPGPU_OK ) return rc;
	/* the reference asserts on a malformed vis lump (bsp.cc:111); the trace does not need it, so a map without one still loads */
	if( vis_len >= 8 ) {
		uint32_t hdr[ 2 ];
		memcpy( hdr, vis, 8 );
		if( ( uint64_t ) hdr[ 0 ] * hdr[ 1 ] + 8 <= vis_len ) {
			rc = upload_vis( *out, vis, vis_len );
			if( rc != BSPGPU_OK ) { bspgpu_destroy( *out ); *out = nullptr; return rc; }
		}
	}
	return BSPGPU_OK;
}

void bspgpu_destroy( bspgpu_t * h ) {
	if( !h ) return;
	DeviceGuard guard( h->device );
	/* nothing of this handle may still be in flight when its buffers go away (BSPGPU_ASYNC work on a caller's stream included) */
	if( h->own_stream ) cudaStreamSynchronize( h->own_stream );
	if( h->use_user_stream ) cudaStreamSynchronize( h->user_stream );
	if( h->user_pending && h->ev_user ) cudaEventSynchronize( h->ev_user );
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
	bspgpu_dist_release( h );
	if( h->d_image ) cudaFree( h->d_image );
	if( h->d_counters ) cudaFree( h->d_counters );
	if( h->d_exotic ) cudaFree( h->d_exotic );
	if( h->d_vis ) cudaFree( h->d_vis );
	if( h->d_scratch ) cudaFree( h->d_scratch );
	if( h->d_pool_ovf ) cudaFree( h->d_pool_ovf );
	if( h->h_small ) cudaFreeHost( h->h_small );
	for( int b = 0; b < kBufs; b++ ) { if( h->h_ring_in[ b ] ) cudaFreeHost( h->h_ring_in[ b ] ); if( h->h_ring_out[ b ] ) cudaFreeHost( h->h_ring_out[ b ] ); if( h->h_ring_pos[ b ] ) cudaFreeHost( h->h_ring_pos[ b ] ); }
	delete h->pool;
	for( cudaEvent_t ev : { h->ev_start, h->ev_stop, h->ev_pre0, h->ev_pre1, h->ev_user } ) if( ev ) cudaEventDestroy( ev );
	if( h->own_stream ) cudaStreamDestroy( h->own_stream );
	if( h->copy_in ) cudaStreamDestroy( h->copy_in );
	if( h->copy_out ) cudaStreamDestroy( h->copy_out );
	delete h;
}

int bspgpu_trace_pos( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	return trace_impl( h, segs, n, out, pos, flags, h->stream(), h->use_user_stream );
}

int bspgpu_trace_stream( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags, void * cuda_stream ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	return trace_impl( h, segs, n, out, pos, flags, ( cudaStream_t ) cuda_stream, true );
}

int bspgpu_trace( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, unsigned flags ) {
	if( n && !out ) return fail( BSPGPU_E_INVALID, "bspgpu_trace: null out" );
	return bspgpu_trace_pos( h, segs, n, out, nullptr, flags );
}

int bspgpu_hit_count( bspgpu_t * h, uint64_t * hits ) {
	if( !h || !hits ) return fail( BSPGPU_E_INVALID, "bspgpu_hit_count: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	if( h->user_pending ) { CK( cudaStreamWaitEvent( st, h->ev_user, 0 ) ); h->user_pending = false; }
	unsigned long long v = 0;
	CK( cudaMemcpyAsync( &v, h->d_counters, sizeof( v ), cudaMemcpyDeviceToHost, st ) );
	CK( cudaStreamSynchronize( st ) );
	*hits = v;
	return BSPGPU_OK;
}

int bspgpu_hit_count_device( bspgpu_t * h, void * dev_u64 ) {
	if( !h || !dev_u64 ) return fail( BSPGPU_E_INVALID, "bspgpu_hit_count_device: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	if( h->user_pending ) { CK( cudaStreamWaitEvent( st, h->ev_user, 0 ) ); h->user_pending = false; }
	CK( cudaMemcpyAsync( dev_u64, h->d_counters, sizeof( unsigned long long ), cudaMemcpyDeviceToDevice, st ) );
	return BSPGPU_OK;
}

int bspgpu_leaf( bspgpu_t * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, unsigned flags ) {
	if( !h || ( n && ( !pts || !leaf ) ) || stride < 3 ) return fail( BSPGPU_E_INVALID, "bspgpu_leaf: bad argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	return leaf_impl( h, pts, stride, n, leaf, nullptr, -1, nullptr, flags );
}

int bspgpu_leaf_clusters( bspgpu_t * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, int32_t * cluster, int32_t from_cluster, uint8_t * visible, unsigned flags ) {
	if( !h || ( n && ( !pts || ( !leaf && !cluster && !visible ) ) ) || stride < 3 ) return fail( BSPGPU_E_INVALID, "bspgpu_leaf_clusters: bad argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	return leaf_impl( h, pts, stride, n, leaf, cluster, from_cluster, visible, flags );
}

int bspgpu_set_vis( bspgpu_t * h, const void * vis_lump, uint64_t len ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_set_vis: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	return upload_vis( h, vis_lump, len );
}

int bspgpu_visible_leaves( bspgpu_t * h, const float pos[ 3 ], uint8_t * visible, int32_t * viewer_cluster ) {
	if( !h || !pos || !visible ) return fail( BSPGPU_E_INVALID, "bspgpu_visible_leaves: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	/* bsp_renderer.cc:148: cluster of the viewer's leaf */
	int32_t cluster = -1;
	int rc = leaf_impl( h, pos, 3, 1, nullptr, &cluster, -1, nullptr, 0 );
	if( rc != BSPGPU_OK ) return rc;
	if( viewer_cluster ) *viewer_cluster = cluster;
	const uint32_t nl = h->lay.num_leaves;
	if( nl == 0 ) return BSPGPU_OK;
	if( cluster < 0 ) { memset( visible, 1, nl ); return BSPGPU_OK; }                 /* :150-156 */
	if( !h->d_vis || ( uint32_t ) cluster >= h->num_clusters ) return fail( BSPGPU_E_INVALID, "bspgpu_visible_leaves: the viewer's cluster has no row in the visibility lump (bspgpu_set_vis)" );
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	rc = ensure_scratch( h, nl );
	if( rc != BSPGPU_OK ) return rc;
	leaf_vis_kernel<<< ( nl + 255 ) / 256, 256, 0, st >>>( reinterpret_cast< const int32_t * >( h->d_image + h->lay.off_leaf_cluster ), nl,
		h->d_vis, h->num_clusters, h->cluster_size, cluster, ( unsigned char * ) h->d_scratch );
	CK( cudaGetLastError() );
	CK( cudaMemcpyAsync( visible, h->d_scratch, nl, cudaMemcpyDeviceToHost, st ) );
	CK( cudaStreamSynchronize( st ) );
	return BSPGPU_OK;
}

int bspgpu_generate( bspgpu_t * h, const bspgpu_gen_desc * g, uint64_t first, uint64_t count, void * dev_segs ) {
	if( !h || !g || ( count && !dev_segs ) ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: null argument" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( g->kind > 2 ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: unknown kind" );
	if( count == 0 ) return BSPGPU_OK;
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	GenParams p;
	memset( &p, 0, sizeof( p ) );
	p.kind = g->kind;
	p.key = sm64_fin( g->seed + kGolden );
	for( int a = 0; a < 3; a++ ) { p.lo[ a ] = g->lo[ a ]; p.ext[ a ] = g->hi[ a ] - g->lo[ a ]; }
	p.len_lo = g->len_lo; p.len_ext = g->len_hi - g->len_lo;
	p.first = first; p.count = count; p.out = reinterpret_cast< float4 * >( dev_segs );
	if( g->kind == 2 ) {
		if( !g->views || !g->num_views || !g->width || !g->height ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: camera kind needs views and a grid size" );
		if( first + count > ( uint64_t ) g->num_views * g->width * g->height ) return fail( BSPGPU_E_INVALID, "bspgpu_generate: camera range exceeds views*width*height" );
		const int rc = ensure_scratch( h, ( size_t ) g->num_views * 12 * sizeof( float ) );
		if( rc != BSPGPU_OK ) return rc;
		CK( cudaMemcpyAsync( h->d_scratch, g->views, ( size_t ) g->num_views * 12 * sizeof( float ), cudaMemcpyHostToDevice, st ) );
		p.views = ( const float * ) h->d_scratch; p.width = g->width; p.height = g->height;
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
	switch( option ) {
		case BSPGPU_OPT_VARIANT: if( value < 0 || value > 3 ) return fail( BSPGPU_E_INVALID, "variant must be 0..3" ); h->opt_variant = value; break;
		case BSPGPU_OPT_BLOCK_THREADS: if( value < 0 || value > 1024 ) return fail( BSPGPU_E_INVALID, "block threads must be 0 (default) or 32..1024" ); h->opt_threads = value > 0 && value < 32 ? 32 : value; break;
		case BSPGPU_OPT_CTAS_PER_SM: if( value < 0 || value > 32 ) return fail( BSPGPU_E_INVALID, "CTAs per SM must be 0 (as many as fit) .. 32" ); h->opt_ctas = value; break;
		case BSPGPU_OPT_TOPTREE_BYTES: if( value < 0 ) return fail( BSPGPU_E_INVALID, "toptree bytes must be >= 0" ); h->opt_top_bytes = value; break;
		case BSPGPU_OPT_CHUNK_SEGMENTS: if( value < 1024 || ( uint64_t ) value > kMaxLaunchSegs ) return fail( BSPGPU_E_INVALID, "chunk segments must be in [1024, 2^30]" ); h->opt_chunk = value; break;
		case BSPGPU_OPT_REFILL: if( value < 1 || value > 32 ) return fail( BSPGPU_E_INVALID, "refill must be 1..32" ); h->opt_refill = value; break;
		case BSPGPU_OPT_MAX_STEPS: if( value < 0 || value > 1000000 ) return fail( BSPGPU_E_INVALID, "max steps must be 0 (default) .. 1000000" ); h->opt_max_steps = value; break;
		case BSPGPU_OPT_BLOCK_SEGS:
			/* a claim of zero segments would let a warp spin on the work counter for ever */
			if( value != 0 && ( value < 32 || value > 65536 ) ) return fail( BSPGPU_E_INVALID, "block segments must be 0 (default) or 32..65536" );
			h->opt_block_segs = value == 0 ? 256 : ( value + 31 ) / 32 * 32; break;
		case BSPGPU_OPT_SCHEDULER: {
#ifdef BSPGPU_EXPERIMENTS
			const bool ab_walk = value == 1;
#else
			const bool ab_walk = false;
#endif
			if( value != 0 && !ab_walk && value != 4 && value != 5 ) return fail( BSPGPU_E_INVALID, "scheduler must be 0 (or 1 / 4 / 5 in an experiments build)" );
		} h->opt_sched = value; break;
		case BSPGPU_OPT_SIDE_STEPS: if( value < 1 || value > 1000000 ) return fail( BSPGPU_E_INVALID, "side steps must be 1..1000000" ); h->opt_side_steps = value; break;
		case BSPGPU_OPT_LEAF_THRESHOLD: if( value < 1 || value > 32 ) return fail( BSPGPU_E_INVALID, "leaf threshold must be 1..32" ); h->opt_leaf_threshold = value; break;
		case BSPGPU_OPT_SMEM_STACK: if( value < -1 || value > 8 ) return fail( BSPGPU_E_INVALID, "smem stack slots must be -1 (default), 0, 2, 4, 6 or 8" ); h->opt_smem_stack = value; break;
		case BSPGPU_OPT_START_GRID: h->opt_start_grid = value < 0 ? -1 : ( value ? 1 : 0 ); break;
		case BSPGPU_OPT_BIN_SEGMENTS: h->opt_bin = value > 0 ? 1 : 0; break;
		case BSPGPU_OPT_TINY_CALLS: h->opt_tiny = value ? 1 : 0; break;
		case BSPGPU_OPT_PAGEABLE: if( value < 0 || value > 2 ) return fail( BSPGPU_E_INVALID, "pageable mode must be 0, 1 or 2" ); h->opt_pageable = value; break;
		case BSPGPU_OPT_MAX_LAUNCH_SEGMENTS:
			if( value < 32 || ( uint64_t ) value > kMaxLaunchSegs ) return fail( BSPGPU_E_INVALID, "max launch segments must be in [32, 2^30]" );
			h->opt_max_launch = ( uint64_t ) value; break;
		default: return fail( BSPGPU_E_INVALID, "bspgpu_set_option: unknown option" );
	}
	h->cfg_valid = false;
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
	if( h->use_user_stream ) {
		/* what is still in flight on the caller's stream shares this handle's counters: the own stream waits for it */
		DeviceGuard dg( h->device );
		CK( cudaEventRecord( h->ev_user, h->user_stream ) );
		h->user_pending = true; h->pending_stream = h->user_stream;
	}
	h->user_stream = nullptr;
	h->use_user_stream = false;
	return BSPGPU_OK;
}

int bspgpu_sync( bspgpu_t * h ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_sync: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	DeviceGuard dg( h->device );
	if( h->user_pending ) { CK( cudaEventSynchronize( h->ev_user ) ); h->user_pending = false; }
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
	info->num_general_planes = h->lay.num_gplanes; info->num_leaves = h->lay.num_leaves; info->num_clusters = h->num_clusters;
	info->smem_stack_slots = h->last_slots;
	if( h->timing_valid ) {
		DeviceGuard dg( h->device );
		CK( cudaEventSynchronize( h->ev_stop ) );
		float ms = 0.0f;
		CK( cudaEventElapsedTime( &ms, h->ev_start, h->ev_stop ) );
		info->last_kernel_ms = ms;
		if( h->prepass_valid ) {
			CK( cudaEventElapsedTime( &ms, h->ev_pre0, h->ev_pre1 ) );
			info->last_prepass_ms = ms;
		}
	}
	return BSPGPU_OK;
}

void * bspgpu_host_alloc( size_t bytes ) {
	void * p = nullptr;
	if( cudaMallocHost( &p, bytes ) != cudaSuccess ) { fail( BSPGPU_E_CUDA, "bspgpu_host_alloc: cudaMallocHost failed" ); cudaGetLastError(); return nullptr; }
	return p;
}

void bspgpu_host_free( void * p ) { if( p ) cudaFreeHost( p ); }

} // extern "C"
