/*
 * trace_kernels.cuh -- sm_100a kernels of the batched segment trace.
 *
 * All three trace kernels are persistent: each warp claims a private block of consecutive
 * segments with one global atomicAdd (warp-level work fetch) and re-fills lanes whose segment
 * has finished from that block as soon as `refill` lanes are idle (__ballot_sync / __popc ranks
 * = the ray compaction), so a warp keeps ~32 live segments instead of waiting for its slowest
 * one.  The per-lane arithmetic is trace_core.h (the reference's statements, one rounding each).
 *
 *   trace_phased      THE SHIPPED KERNEL.  Lanes are resumable node/leaf state machines; each warp
 *                     iteration runs the one bounded phase most of its lanes wait for.  A node step is one
 *                     32-byte gather + one plane test, a leaf step one 32-byte gather + two brush sides;
 *                     with the map in global memory the first traversal-stack slots of every thread live
 *                     in shared memory (SmemStack below).
 *   trace_persistent  the first version (rounds: every lane descends to a leaf, then every lane
 *                     clips its leaf).  Kept selectable (BSPGPU_OPT_SCHEDULER=1) as a cross-check.
 *   trace_dual        experiment (BSPGPU_OPT_SCHEDULER=3): two segments per lane, one parked in
 *                     shared memory.  Higher SIMT utilisation, but loses to L1 capacity on rooms8.
 *
 * Map staging (template `kMode`):
 *   kSmemAll : the whole flattened map [nodes|refs|sides] is copied into shared memory with
 *              cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) completing on an mbarrier   (C1/C2)
 *   kSmemTop : only the first `top_nodes` breadth-first nodes are staged; the rest of the map is
 *              read through the read-only path (L1 -> L2)
 *   kGlobal  : nothing staged; a node is one 256-bit read-only load                      (C3-C5)
 *   kGlobalPacked : kGlobal over two-level node records (QNodeRec, map_layout.h); a measured-slower option
 * K3  hit count: per-lane counters, one shuffle reduction + one atomicAdd per warp.
 * N3  leaf_kernel: batched BSP::position_to_leaf (reference bsp.cc:273-286).
 *
 * No tensor cores: the walk is a dependent chain of 16/32-byte gathers and a handful of fp32
 * ops per node/side, not a contraction.
 */
#ifndef BSPGPU_TRACE_KERNELS_CUH
#define BSPGPU_TRACE_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "trace_core.h"

namespace bspk {

enum Mode { kSmemAll = 0, kSmemTop = 1, kGlobal = 2, kGlobalPacked = 3 };

struct TraceParams {
	const unsigned char * image;      /* flattened map in global memory */
	uint64_t off_refs, off_sides, off_side_plane, off_qnodes;
	uint32_t staged_bytes;            /* bytes copied to shared memory by this mode */
	uint32_t top_nodes;               /* kSmemTop: nodes [0, top_nodes) live in shared memory */
	const float4 * segs;              /* 2 float4 per segment ... */
	const float2 * segs24;            /* ... or 3 float2 (6 packed floats) per segment */
	ResultRec * out;                  /* may be null */
	float4 * pos;                     /* may be null */
	uint32_t n;
	uint32_t block_segs;              /* segments a warp claims per global atomic */
	uint32_t refill;                  /* idle lanes that trigger a refill (1..32) */
	uint32_t max_steps;               /* node steps per lane per node phase (bounds divergence) */
	uint32_t side_steps;              /* side steps per lane per leaf phase */
	uint32_t leaf_threshold;          /* lanes waiting in a leaf that make the warp run a leaf phase */
	StartGrid grid;                   /* cells == null: every walk starts at node 0 */
	unsigned long long * next;        /* work counter, zeroed before launch */
	unsigned long long * hits;        /* accumulated hit count */
};

/* ---------------------------------------------------------------- PTX helpers (TMA bulk copy + mbarrier) */

__device__ __forceinline__ uint32_t smem_u32( const void * p ) {
	return ( uint32_t ) __cvta_generic_to_shared( p );
}
__device__ __forceinline__ void mbar_init( uint64_t * bar, uint32_t count ) {
	asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"( smem_u32( bar ) ), "r"( count ) );
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
}
__device__ __forceinline__ void mbar_expect_tx( uint64_t * bar, uint32_t bytes ) {
	asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ bool mbar_try_wait( uint64_t * bar, uint32_t phase ) {
	uint32_t ok;
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		"selp.u32 %0, 1, 0, p;\n"
		"}\n" : "=r"( ok ) : "r"( smem_u32( bar ) ), "r"( phase ) : "memory" );
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait( uint64_t * bar, uint32_t phase ) {
	while( !mbar_try_wait( bar, phase ) ) { }
}
/* 1-D TMA bulk copy global -> shared, completion counted on the mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s( void * dst, const void * src, uint32_t bytes, uint64_t * bar ) {
	asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"( smem_u32( dst ) ), "l"( src ), "r"( bytes ), "r"( smem_u32( bar ) ) : "memory" );
}

/* streaming loads/stores for the segment and result streams: read once / written once, keep them
 * out of L1 and first in line for L2 eviction so the map stays cached */
__device__ __forceinline__ float4 ld_stream4( const float4 * p ) {
	float4 v;
	asm volatile( "ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( v.x ), "=f"( v.y ), "=f"( v.z ), "=f"( v.w ) : "l"( p ) );
	return v;
}
__device__ __forceinline__ float2 ld_stream2( const float2 * p ) {
	float2 v;
	asm volatile( "ld.global.cs.v2.f32 {%0,%1}, [%2];" : "=f"( v.x ), "=f"( v.y ) : "l"( p ) );
	return v;
}
__device__ __forceinline__ void st_stream4( void * p, float a, float b, float c, float d ) {
	asm volatile( "st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"( p ), "f"( a ), "f"( b ), "f"( c ), "f"( d ) : "memory" );
}

/* ---------------------------------------------------------------- map accessors */

/* explicit ld.shared through 32-bit shared-window addresses: the generic->shared conversion of an
 * `extern __shared__` pointer costs four uniform-datapath instructions at every load site */
__device__ __forceinline__ float4 lds128( uint32_t addr ) {
	float4 v;
	asm volatile( "ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( v.x ), "=f"( v.y ), "=f"( v.z ), "=f"( v.w ) : "r"( addr ) );
	return v;
}
__device__ __forceinline__ int2 lds64i( uint32_t addr ) {
	int2 v;
	asm volatile( "ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"( v.x ), "=r"( v.y ) : "r"( addr ) );
	return v;
}
__device__ __forceinline__ uint4 lds128u( uint32_t addr ) {
	uint4 v;
	asm volatile( "ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"( v.x ), "=r"( v.y ), "=r"( v.z ), "=r"( v.w ) : "r"( addr ) );
	return v;
}

/* keep a computed shared-window address in a register: without this the compiler re-derives it from
 * the CTA's shared window (S2UR SR_CgaCtaId + 3 uniform ops) at every load site inside the hot loops */
__device__ __forceinline__ uint32_t opaque_u32( uint32_t v ) {
	asm volatile( "" : "+r"( v ) );
	return v;
}

struct SmemAllMap {
	static const bool kPackedNodes = false;
	uint32_t nodes, refs, sides;   /* shared-window byte addresses of the three sections */
	__device__ __forceinline__ void node( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) const {
		const uint32_t a = nodes + ( ( uint32_t ) i << 5 );
		const float4 p = lds128( a );
		const int2 c = lds64i( a + 16 );
		nx = p.x; ny = p.y; nz = p.z; d = p.w; c0 = c.x; c1 = c.y;
	}
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const uint4 r = lds128u( refs + ( i << 4 ) );
		first = r.x; nraw = r.y; brush = ( int32_t ) r.z;
	}
	__device__ __forceinline__ void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const float4 p = lds128( sides + ( i << 4 ) );
		nx = p.x; ny = p.y; nz = p.z; d = p.w;
	}
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		const uint32_t a = sides + ( i << 4 );
		const float4 p = lds128( a ), q = lds128( a + 16 );
		n0x = p.x; n0y = p.y; n0z = p.z; d0 = p.w; n1x = q.x; n1y = q.y; n1z = q.z; d1 = q.w;
	}
};

/* one 256-bit read-only load (sm_100: LDG.E.ENL2.256.CONSTANT) fetches a whole 32-byte NodeRec --
 * plane and children in ONE L1 request instead of two (the city10k profile showed the L1 data pipe
 * at 75 % with two scattered requests per node step) */
__device__ __forceinline__ void ldg256( const void * p, float & a, float & b, float & c, float & d, int32_t & e, int32_t & f, int32_t & g, int32_t & h ) {
	asm volatile( "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"( a ), "=f"( b ), "=f"( c ), "=f"( d ), "=r"( e ), "=r"( f ), "=r"( g ), "=r"( h ) : "l"( p ) );
}

/* two consecutive SideRecs (i even: 32-byte aligned) in one 256-bit read-only load */
__device__ __forceinline__ void ldg_side_pair( const SideRec * p, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) {
	asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"( n0x ), "=f"( n0y ), "=f"( n0z ), "=f"( d0 ), "=f"( n1x ), "=f"( n1y ), "=f"( n1z ), "=f"( d1 ) : "l"( p ) );
}

struct GlobalMap {
	static const bool kPackedNodes = false;
	const NodeRec * nodes; const BrushRef * refs; const SideRec * sides;   /* global memory, read-only path */
	__device__ __forceinline__ void node( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) const {
		int32_t o0, o1;
		ldg256( &nodes[ i ], nx, ny, nz, d, c0, c1, o0, o1 );
	}
	__device__ __forceinline__ void node_full( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1, int32_t & o0, int32_t & o1 ) const {
		ldg256( &nodes[ i ], nx, ny, nz, d, c0, c1, o0, o1 );
	}
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const uint4 r = __ldg( reinterpret_cast< const uint4 * >( &refs[ i ] ) );
		first = r.x; nraw = r.y; brush = ( int32_t ) r.z;
	}
	__device__ __forceinline__ void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const float4 p = __ldg( reinterpret_cast< const float4 * >( &sides[ i ] ) );
		nx = p.x; ny = p.y; nz = p.z; d = p.w;
	}
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		ldg_side_pair( &sides[ i ], n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

/* kGlobalPacked: the walk fetches two-level QNodeRecs (map_layout.h) -- still one 256-bit gather per fetch, about half
 * as many fetches on maps with axial node planes.  Brush references and sides as in GlobalMap. */
struct GlobalPackedMap {
	static const bool kPackedNodes = true;
	const QNodeRec * qnodes; const BrushRef * refs; const SideRec * sides;
	__device__ __forceinline__ void qnode( int32_t i, uint32_t & q0, uint32_t & q1, uint32_t & q2, uint32_t & q3, uint32_t & q4, uint32_t & q5, uint32_t & q6, uint32_t & w ) const {
		asm volatile( "ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			: "=r"( q0 ), "=r"( q1 ), "=r"( q2 ), "=r"( q3 ), "=r"( q4 ), "=r"( q5 ), "=r"( q6 ), "=r"( w ) : "l"( &qnodes[ i ] ) );
	}
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const uint4 r = __ldg( reinterpret_cast< const uint4 * >( &refs[ i ] ) );
		first = r.x; nraw = r.y; brush = ( int32_t ) r.z;
	}
	__device__ __forceinline__ void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const float4 p = __ldg( reinterpret_cast< const float4 * >( &sides[ i ] ) );
		nx = p.x; ny = p.y; nz = p.z; d = p.w;
	}
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		ldg_side_pair( &sides[ i ], n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

struct SmemTopMap {
	static const bool kPackedNodes = false;
	uint32_t top;          /* shared-window byte address of nodes [0, top_nodes) */
	uint32_t top_nodes;
	GlobalMap g;
	__device__ __forceinline__ void node( int32_t i, float & nx, float & ny, float & nz, float & d, int32_t & c0, int32_t & c1 ) const {
		if( ( uint32_t ) i < top_nodes ) {
			const uint32_t a = top + ( ( uint32_t ) i << 5 );
			const float4 p = lds128( a );
			const int2 c = lds64i( a + 16 );
			nx = p.x; ny = p.y; nz = p.z; d = p.w; c0 = c.x; c1 = c.y;
		}
		else g.node( i, nx, ny, nz, d, c0, c1 );
	}
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const { g.ref( i, first, nraw, brush ); }
	__device__ __forceinline__ void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const { g.side( i, nx, ny, nz, d ); }
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		g.side_pair( i, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

template< int kMode > struct MapOf;
template<> struct MapOf< kSmemAll > { typedef SmemAllMap type; };
template<> struct MapOf< kSmemTop > { typedef SmemTopMap type; };
template<> struct MapOf< kGlobal > { typedef GlobalMap type; };
template<> struct MapOf< kGlobalPacked > { typedef GlobalPackedMap type; };

/* Stage what this mode keeps in shared memory -- one thread issues TMA bulk copies (cp.async.bulk, 32 KB each),
 * everyone waits on the mbarrier's transaction count -- and bind the map accessor to shared / global memory. */
template< int kMode >
__device__ __forceinline__ void stage_and_bind( const TraceParams & p, unsigned char * smem, uint64_t * stage_bar, typename MapOf< kMode >::type & map ) {
	if( ( kMode == kSmemAll || kMode == kSmemTop ) && p.staged_bytes > 0 ) {
		if( threadIdx.x == 0 ) {
			mbar_init( stage_bar, 1 );
			mbar_fence_init();
		}
		__syncthreads();
		if( threadIdx.x == 0 ) {
			mbar_expect_tx( stage_bar, p.staged_bytes );
			const uint32_t kChunk = 32768;
			for( uint32_t off = 0; off < p.staged_bytes; off += kChunk ) {
				const uint32_t sz = min( kChunk, p.staged_bytes - off );
				bulk_g2s( smem + off, p.image + off, sz, stage_bar );
			}
		}
		mbar_wait( stage_bar, 0 );
	}

	const NodeRec * g_nodes = reinterpret_cast< const NodeRec * >( p.image );
	const BrushRef * g_refs = reinterpret_cast< const BrushRef * >( p.image + p.off_refs );
	const SideRec * g_sides = reinterpret_cast< const SideRec * >( p.image + p.off_sides );
	if constexpr( kMode == kSmemAll ) {
		map.nodes = opaque_u32( smem_u32( smem ) );
		map.refs = opaque_u32( map.nodes + ( uint32_t ) p.off_refs );
		map.sides = opaque_u32( map.nodes + ( uint32_t ) p.off_sides );
	}
	else if constexpr( kMode == kSmemTop ) {
		map.top = opaque_u32( smem_u32( smem ) );
		map.top_nodes = p.top_nodes;
		map.g.nodes = g_nodes; map.g.refs = g_refs; map.g.sides = g_sides;
	}
	else if constexpr( kMode == kGlobalPacked ) {
		map.qnodes = reinterpret_cast< const QNodeRec * >( p.image + p.off_qnodes ); map.refs = g_refs; map.sides = g_sides;
	}
	else {
		map.nodes = g_nodes; map.refs = g_refs; map.sides = g_sides;
	}
}

/* ---------------------------------------------------------------- K1/K2 */

template< int kMode, int kStack, int kMaxThreads >
__global__ void __launch_bounds__( kMaxThreads ) trace_persistent( const TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;

	/* ---- stage the map (or its top) into shared memory: one thread issues TMA bulk copies,
	 *      everyone waits on the mbarrier's transaction count */
	typename MapOf< kMode >::type map;
	stage_and_bind< kMode >( p, smem, &stage_bar, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );

	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = ( 1u << lane ) - 1u;
	const unsigned kFull = 0xffffffffu;

	/* warp-private block of claimed segments [blk_next, blk_end): warp-uniform registers */
	uint32_t blk_next = 0, blk_end = 0;
	bool more_work = true;

	/* per-lane segment state */
	bool active = false;
	uint32_t seg = 0;
	Ray r; r.sx = r.sy = r.sz = r.dx = r.dy = r.dz = 0.0f;
	HitState h; hit_init( h );
	StackEntry stack[ kStack ];
	int32_t sp = 0, node = 0;
	float tmin = 0.0f, tmax = 1.0f;
	uint32_t my_hits = 0;

	for( ;; ) {
		/* ---- warp-level work fetch + compaction: idle lanes take consecutive new segments */
		const unsigned idle = __ballot_sync( kFull, !active );
		const unsigned n_idle = __popc( idle );
		if( more_work && n_idle >= p.refill ) {
			if( blk_next == blk_end ) {
				unsigned long long b = 0;
				if( lane == 0 ) b = atomicAdd( p.next, ( unsigned long long ) p.block_segs );
				b = __shfl_sync( kFull, b, 0 );
				if( b >= ( unsigned long long ) p.n ) more_work = false;
				else {
					blk_next = ( uint32_t ) b;
					blk_end = ( uint32_t ) min( b + ( unsigned long long ) p.block_segs, ( unsigned long long ) p.n );
				}
			}
			if( more_work ) {
				const uint32_t avail = blk_end - blk_next;
				const uint32_t rank = __popc( idle & lt_mask );
				if( !active && rank < avail ) {
					seg = blk_next + rank;
					float ex, ey, ez;
					if( p.segs24 ) {
						const float2 * s = p.segs24 + ( size_t ) seg * 3;
						const float2 a = ld_stream2( s ), b2 = ld_stream2( s + 1 ), c = ld_stream2( s + 2 );
						r.sx = a.x; r.sy = a.y; r.sz = b2.x; ex = b2.y; ey = c.x; ez = c.y;
					}
					else {
						const float4 a = ld_stream4( p.segs + ( size_t ) seg * 2 );
						const float4 b4 = ld_stream4( p.segs + ( size_t ) seg * 2 + 1 );
						r.sx = a.x; r.sy = a.y; r.sz = a.z; ex = b4.x; ey = b4.y; ez = b4.z;
					}
					r.dx = ex - r.sx; r.dy = ey - r.sy; r.dz = ez - r.sz;      /* bsp.cc:261 */
					hit_init( h );                                              /* bsp.cc:256-257 */
					sp = 0; node = 0; tmin = 0.0f; tmax = 1.0f;                 /* bsp.cc:263 */
					active = true;
				}
				blk_next += min( n_idle, avail );
			}
		}
		if( __ballot_sync( kFull, active ) == 0u ) {
			if( !more_work ) break;
			continue;
		}

		/* ---- walk: at most max_steps node steps, then the leaf if one was reached */
		if( active ) {
			uint32_t steps = p.max_steps;
			while( node >= 0 && steps-- ) {
				float nx, ny, nz, d; int32_t c0, c1;
				map.node( node, nx, ny, nz, d, c0, c1 );
				int32_t far; bool both;
				node = node_step( r, nx, ny, nz, d, c0, c1, tmin, tmax, far, both );
				if( both ) { stack[ sp ].node = far; stack[ sp ].t = tmax; sp++; }
			}
		}
		if( active && node < 0 ) {
			if( node != kEmptyLeaf ) leaf( map, r, node, tmin, tmax, h );
			bool done = h.hit != 0;                                             /* bsp.cc:206 */
			if( !done ) {
				do {
					if( sp == 0 ) { done = true; break; }
					sp--;
					node = stack[ sp ].node;
					tmin = stack[ sp ].t;
					tmax = sp > 0 ? stack[ sp - 1 ].t : 1.0f;
				} while( node == kEmptyLeaf );
			}
			if( done ) {
				const ResultRec o = make_result( h, side_plane );
				if( p.out ) st_stream4( p.out + seg, o.t, __int_as_float( o.plane ), __int_as_float( o.brush ), __uint_as_float( o.flags ) );
				if( p.pos ) {
					float px, py, pz;
					make_pos( r, h, px, py, pz );                               /* bsp.cc:265-267 */
					st_stream4( p.pos + seg, px, py, pz, h.t );
				}
				my_hits += ( uint32_t ) h.hit;
				active = false;
			}
		}
	}

	/* ---- K3: hit count, one atomic per warp */
	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( kFull, my_hits, o );
	if( lane == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

/* read segment `sidx`, begin its walk: at node 0 (bsp.cc:263) or at the start-grid reference of its cell */
__device__ __forceinline__ void lane_load_segment( const TraceParams & p, uint32_t sidx, Lane & l ) {
	float sx, sy, sz, ex, ey, ez;
	if( p.segs24 ) {
		const float2 * sp24 = p.segs24 + ( size_t ) sidx * 3;
		const float2 a = ld_stream2( sp24 ), b2 = ld_stream2( sp24 + 1 ), c = ld_stream2( sp24 + 2 );
		sx = a.x; sy = a.y; sz = b2.x; ex = b2.y; ey = c.x; ez = c.y;
	}
	else {
		const float4 a = ld_stream4( p.segs + ( size_t ) sidx * 2 );
		const float4 b4 = ld_stream4( p.segs + ( size_t ) sidx * 2 + 1 );
		sx = a.x; sy = a.y; sz = a.z; ex = b4.x; ey = b4.y; ez = b4.z;
	}
	/* the start reference first, the lane set-up once after it (one copy of the initialisation instead of one per grid outcome) */
	const int32_t ref = grid_start_ref( p.grid, sx, sy, sz, ex, ey, ez );   /* 0 when there is no grid */
	lane_start( l, sx, sy, sz, ex, ey, ez );
	lane_start_at( l, ref );
}

/* Traversal stack with its first kSlots entries in shared memory.
 *
 * Local memory interleaves threads word by word, so a 16-byte entry of one lane is four words in four different
 * 128-byte lines: a push by k lanes at k different heights costs 4k L1 wavefronts (ncu on city10k long rays: stack
 * traffic was 48 % of all L1 sectors, and the L1 data pipe the busiest unit at 82 %).  In shared memory, laid out
 * [slot][thread] with 16 bytes per thread, the bank depends on the thread only: any mix of heights is conflict-free
 * and a warp-wide push or pop is at most four wavefronts, outside the L1 tag path.  Measured heights:
 * 87 % of the pushes of long rays land in slots 0-3, 99 % in 0-5; deeper entries overflow to local memory. */
template< int kSlots, int kStride > struct SmemStack;
template< int kSlots, int kThreads > struct StackOf { typedef SmemStack< kSlots, kThreads * 16 > type; };
template< int kThreads > struct StackOf< 0, kThreads > { typedef StackEntry16 * type; };

template< int kSlots, int kStride >
struct SmemStack {
	uint32_t base;              /* shared-window byte address of this thread's slot 0; slot i is kStride bytes further */
	StackEntry16 * overflow;    /* local memory, indexed by the absolute slot number */
};
/* (A branch-free form -- predicated STS/LDS plus a predicated local overflow -- measured 6-12 % slower on every
 * workload: as branches, a warp in which no lane pushes or pops skips the block.) */
template< int kSlots, int kStride >
__device__ __forceinline__ void stack_put( const SmemStack< kSlots, kStride > & s, int32_t i, int32_t node, float tmin, float tmax ) {
	/* both stores name the same four registers, so one register quad serves either destination; the fourth word is padding */
	uint32_t pad;
	asm( "" : "=r"( pad ) );
	asm volatile( "{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %0, %7;\n\t@q st.shared.v4.b32 [%1], {%3,%4,%5,%6};\n\t@!q st.local.v4.b32 [%2], {%3,%4,%5,%6};\n\t}"
		:: "r"( i ), "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ), "l"( __cvta_generic_to_local( s.overflow + i ) ),
		   "r"( node ), "f"( tmin ), "f"( tmax ), "r"( pad ), "n"( kSlots ) : "memory" );
}
template< int kSlots, int kStride >
__device__ __forceinline__ void stack_get( const SmemStack< kSlots, kStride > & s, int32_t i, int32_t & node, float & tmin, float & tmax ) {
	if( i < kSlots ) {
		int32_t pad;
		asm volatile( "ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"( node ), "=f"( tmin ), "=f"( tmax ), "=r"( pad )
			: "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ) : "memory" );
	}
	else stack_get( s.overflow, i, node, tmin, tmax );
}

/* ---------------------------------------------------------------- K1/K2, phase-scheduled
 *
 * ncu on the round-based kernel above showed it issue-bound (88 % issue-active on rooms8) at
 * only ~9 of 32 threads per instruction: lanes that reached an empty leaf idled through the
 * whole brush loop of the few lanes that had brushes, and vice versa.  Here every lane is a
 * small state machine (trace_core.h: Lane) that is either stepping nodes or stepping brush
 * sides, empty leaves are popped on the spot, and each warp iteration runs ONE bounded phase:
 * a leaf phase when at least `leaf_threshold` lanes wait in a leaf (or nobody can step nodes),
 * otherwise a node phase.  Lanes of the other kind just keep their registers.  Finished lanes
 * are re-filled as before. */
template< int kMode, int kStack, int kMaxThreads, int kMinBlocks, int kSmemSlots = 0 >
__global__ void __launch_bounds__( kMaxThreads, kMinBlocks ) trace_phased( const TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;

	typename MapOf< kMode >::type map;
	stage_and_bind< kMode >( p, smem, &stage_bar, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );

	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = ( 1u << lane ) - 1u;
	const unsigned kFull = 0xffffffffu;

	uint32_t blk_next = 0, blk_end = 0;   /* warp-private block of claimed segments (warp-uniform) */
	bool more_work = true;

	Lane l;
	lane_start( l, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f );
	l.cur = kCurIdle;
	l.brush = -1; l.bflags = 0; l.tfar = l.tnear = 0.0f; l.far_side = -1;
	StackEntry16 stack_local[ kStack ];
	/* kSmemSlots > 0: the first slots live in shared memory, behind whatever this mode staged there */
	typename StackOf< kSmemSlots, kMaxThreads >::type stack;
	if constexpr( kSmemSlots > 0 ) {
		stack.base = opaque_u32( smem_u32( smem ) + ( ( p.staged_bytes + 127u ) & ~127u ) + threadIdx.x * 16u );
		stack.overflow = stack_local;
	}
	else stack = stack_local;
	uint32_t seg = 0, my_hits = 0;

	for( ;; ) {
		/* ---- retire finished segments (also those the start grid resolved at load time: a segment inside an empty leaf) */
		if( lane_done( l ) ) {
			HitState h;
			lane_hit_state( l, h );
			const ResultRec o = make_result( h, side_plane );
			if( p.out ) st_stream4( p.out + seg, o.t, __int_as_float( o.plane ), __int_as_float( o.brush ), __uint_as_float( o.flags ) );
			if( p.pos ) {
				float px, py, pz;
				make_pos( l.r, h, px, py, pz );                                 /* bsp.cc:265-267 */
				st_stream4( p.pos + seg, px, py, pz, h.t );
			}
			my_hits += ( uint32_t ) h.hit;
			l.cur = kCurIdle;
		}

		/* ---- warp-level work fetch + compaction */
		const unsigned idle = __ballot_sync( kFull, lane_idle( l ) );
		const unsigned n_idle = __popc( idle );
		if( more_work && n_idle >= p.refill ) {
			if( blk_next == blk_end ) {
				unsigned long long b = 0;
				if( lane == 0 ) b = atomicAdd( p.next, ( unsigned long long ) p.block_segs );
				b = __shfl_sync( kFull, b, 0 );
				if( b >= ( unsigned long long ) p.n ) more_work = false;
				else {
					blk_next = ( uint32_t ) b;
					blk_end = ( uint32_t ) min( b + ( unsigned long long ) p.block_segs, ( unsigned long long ) p.n );
				}
			}
			if( more_work ) {
				const uint32_t avail = blk_end - blk_next;
				const uint32_t rank = __popc( idle & lt_mask );
				if( lane_idle( l ) && rank < avail ) {
					seg = blk_next + rank;
					lane_load_segment( p, seg, l );
				}
				blk_next += min( n_idle, avail );
			}
		}

		/* ---- pick the phase most lanes are waiting for */
		const unsigned in_node = __ballot_sync( kFull, lane_in_node( l ) );
		const unsigned in_leaf = __ballot_sync( kFull, lane_in_leaf( l ) );
		if( ( in_node | in_leaf ) == 0u ) {
			if( __ballot_sync( kFull, lane_done( l ) ) != 0u ) continue;       /* loaded and already finished: retire them first */
			if( !more_work ) break;
			continue;
		}
		if( in_node == 0u || ( uint32_t ) __popc( in_leaf ) >= p.leaf_threshold ) lane_leaf_phase( map, l, stack, p.side_steps );
		else lane_node_phase( map, l, stack, p.max_steps );

	}

	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( kFull, my_hits, o );
	if( lane == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

/* ---------------------------------------------------------------- one thread per segment, no scheduling
 *
 * BSPGPU_OPT_SCHEDULER=4.  Grid-stride loop, each thread walks its segment start to finish.  No work
 * fetch, no phases, no compaction: for workloads whose walks are a handful of steps (short moves that
 * begin at a start-grid reference) the persistent kernel's bookkeeping costs more than divergence does. */
template< int kMode, int kStack, int kMaxThreads >
__global__ void __launch_bounds__( kMaxThreads ) trace_simple( const TraceParams p ) {
	static_assert( kMode == kGlobal, "the simple kernel reads the map through the read-only path" );
	GlobalMap map;
	map.nodes = reinterpret_cast< const NodeRec * >( p.image );
	map.refs = reinterpret_cast< const BrushRef * >( p.image + p.off_refs );
	map.sides = reinterpret_cast< const SideRec * >( p.image + p.off_sides );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );
	StackEntry16 stack_local[ kStack ];
	StackEntry16 * const stack = stack_local;
	uint32_t my_hits = 0;
	for( uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x; seg < p.n; seg += gridDim.x * blockDim.x ) {
		Lane l;
		lane_load_segment( p, seg, l );
		while( !lane_done( l ) ) {
			lane_node_phase( map, l, stack, 0xffffffffu );
			lane_leaf_phase( map, l, stack, 0xffffffffu );
		}
		HitState h;
		lane_hit_state( l, h );
		const ResultRec o = make_result( h, side_plane );
		if( p.out ) st_stream4( p.out + seg, o.t, __int_as_float( o.plane ), __int_as_float( o.brush ), __uint_as_float( o.flags ) );
		if( p.pos ) {
			float px, py, pz;
			make_pos( l.r, h, px, py, pz );
			st_stream4( p.pos + seg, px, py, pz, h.t );
		}
		my_hits += ( uint32_t ) h.hit;
	}
	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( 0xffffffffu, my_hits, o );
	if( ( threadIdx.x & 31u ) == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

/* ---------------------------------------------------------------- K1 with two segments per lane
 *
 * Two-phase scheduling of independent lanes tops out near 55 % SIMT utilisation: a lane is in one
 * mode at a time and waits whenever its warp runs the other phase.  Here every lane owns TWO
 * segments: one live in registers, one parked in shared memory (6 x 16 bytes per thread, laid
 * out [word][thread] so a warp's accesses are conflict-free), each with its own traversal stack
 * in local memory.  Before a phase, a lane whose live segment does not match the phase but whose
 * parked one does swaps them (six LDS.128/STS.128 pairs through four temporaries).  A lane can
 * then follow the warp's phase with probability 1-(1-p)^2 instead of p.  The arithmetic is
 * untouched: the same lane_node_phase / lane_leaf_phase run on whichever segment is live.
 * Costs 96 bytes of shared memory per thread on top of whatever is staged. */

struct ParkSlot {
	uint32_t base;       /* shared-window byte address of this thread's word 0 */
	uint32_t stride;     /* bytes between consecutive words = 16 * blockDim.x */
};

__device__ __forceinline__ void sts128( uint32_t addr, float a, float b, float c, float d ) {
	asm volatile( "st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"( addr ), "f"( a ), "f"( b ), "f"( c ), "f"( d ) : "memory" );
}
__device__ __forceinline__ float4 lds128v( uint32_t addr ) {
	float4 v;
	asm volatile( "ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( v.x ), "=f"( v.y ), "=f"( v.z ), "=f"( v.w ) : "r"( addr ) : "memory" );
	return v;
}

#define BSPK_SWAP4( k, a, b, c, d ) { \
	const float4 t_ = lds128v( slot.base + ( k ) * slot.stride ); \
	sts128( slot.base + ( k ) * slot.stride, a, b, c, d ); \
	a = t_.x; b = t_.y; c = t_.z; d = t_.w; }

/* exchange the live segment (registers) with the parked one (shared memory) */
__device__ __forceinline__ void lane_swap( const ParkSlot & slot, Lane & l, uint32_t & seg ) {
	float f_cur = __int_as_float( l.cur ), f_sp = __int_as_float( l.sp ), f_hflags = __uint_as_float( l.hflags );
	float f_hside = __int_as_float( l.hside ), f_hbrush = __int_as_float( l.hbrush ), f_s = __uint_as_float( l.s );
	float f_send = __uint_as_float( l.s_end ), f_brush = __int_as_float( l.brush ), f_bflags = __uint_as_float( l.bflags );
	float f_fside = __int_as_float( l.far_side ), f_seg = __uint_as_float( seg ), f_pad0 = 0.0f, f_pad1 = 0.0f;
	BSPK_SWAP4( 0, l.r.sx, l.r.sy, l.r.sz, l.r.dx )
	BSPK_SWAP4( 1, l.r.dy, l.r.dz, l.tmin, l.tmax )
	BSPK_SWAP4( 2, f_cur, f_sp, l.ht, f_hflags )
	BSPK_SWAP4( 3, f_hside, f_hbrush, f_s, f_send )
	BSPK_SWAP4( 4, f_brush, f_bflags, l.tfar, l.tnear )
	BSPK_SWAP4( 5, f_fside, f_seg, f_pad0, f_pad1 )
	l.cur = __float_as_int( f_cur ); l.sp = __float_as_int( f_sp ); l.hflags = __float_as_uint( f_hflags );
	l.hside = __float_as_int( f_hside ); l.hbrush = __float_as_int( f_hbrush ); l.s = __float_as_uint( f_s );
	l.s_end = __float_as_uint( f_send ); l.brush = __float_as_int( f_brush ); l.bflags = __float_as_uint( f_bflags );
	l.far_side = __float_as_int( f_fside ); seg = __float_as_uint( f_seg );
}

template< int kMode, int kStack, int kMaxThreads >
__global__ void __launch_bounds__( kMaxThreads, 1 ) trace_dual( const TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;

	typename MapOf< kMode >::type map;
	stage_and_bind< kMode >( p, smem, &stage_bar, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );

	/* parking area right behind the staged bytes (sections are 128-byte aligned) */
	ParkSlot slot;
	slot.stride = 16u * blockDim.x;
	slot.base = opaque_u32( smem_u32( smem ) + p.staged_bytes + 16u * threadIdx.x );

	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = ( 1u << lane ) - 1u;
	const unsigned kFull = 0xffffffffu;

	uint32_t blk_next = 0, blk_end = 0;
	bool more_work = true;

	Lane l;
	lane_start( l, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f );
	l.cur = kCurIdle;
	l.brush = -1; l.bflags = 0; l.tfar = l.tnear = 0.0f; l.far_side = -1;
	uint32_t seg = 0, my_hits = 0;
	/* the parked slot starts out idle: write one idle state (cur == kCurIdle) into it.
	 * Invariant from here on: the slot always holds a complete state and its `cur` word == pcur. */
	lane_swap( slot, l, seg );
	l.cur = kCurIdle;
	int32_t pcur = kCurIdle;
	uint32_t which = 0;                   /* which local stack belongs to the live segment */
	StackEntry16 stacks[ 2 ][ kStack ];

	for( ;; ) {
		/* ---- work fetch: every empty slot (live first, then parked) takes a new segment */
		const bool live_idle = l.cur == kCurIdle, park_idle = pcur == kCurIdle;
		const unsigned m_live = __ballot_sync( kFull, live_idle );
		const unsigned m_park = __ballot_sync( kFull, park_idle );
		const uint32_t n_live = __popc( m_live ), n_want = n_live + __popc( m_park );
		if( more_work && n_want >= p.refill ) {
			if( blk_next == blk_end ) {
				unsigned long long b = 0;
				if( lane == 0 ) b = atomicAdd( p.next, ( unsigned long long ) p.block_segs );
				b = __shfl_sync( kFull, b, 0 );
				if( b >= ( unsigned long long ) p.n ) more_work = false;
				else {
					blk_next = ( uint32_t ) b;
					blk_end = ( uint32_t ) min( b + ( unsigned long long ) p.block_segs, ( unsigned long long ) p.n );
				}
			}
			if( more_work ) {
				const uint32_t avail = blk_end - blk_next;
				const uint32_t rank_live = __popc( m_live & lt_mask );
				const uint32_t rank_park = n_live + __popc( m_park & lt_mask );
				if( live_idle && rank_live < avail ) {
					seg = blk_next + rank_live;
					lane_load_segment( p, seg, l );
				}
				if( park_idle && rank_park < avail ) {
					/* move whatever is live (old or just started) to the parked slot, start the new one live */
					const int32_t old_cur = l.cur;
					lane_swap( slot, l, seg );
					pcur = old_cur; which ^= 1u;
					seg = blk_next + rank_park;
					lane_load_segment( p, seg, l );
				}
				blk_next += min( n_want, avail );
			}
		}

		/* ---- pick the phase most lanes can follow with one of their two segments */
		const bool live_node = lane_in_node( l ), live_leaf = lane_in_leaf( l );
		const bool park_node = pcur >= 0, park_leaf = ( uint32_t ) pcur > ( uint32_t ) kCurDone;
		const unsigned can_node = __ballot_sync( kFull, live_node || park_node );
		const unsigned can_leaf = __ballot_sync( kFull, live_leaf || park_leaf );
		if( ( can_node | can_leaf ) == 0u ) {
			if( !more_work ) break;
			continue;
		}
		const bool leaf_phase = can_node == 0u || ( uint32_t ) __popc( can_leaf ) >= p.leaf_threshold;
		if( leaf_phase ? ( !live_leaf && park_leaf ) : ( !live_node && park_node ) ) {
			const int32_t old_cur = l.cur;
			lane_swap( slot, l, seg );
			pcur = old_cur; which ^= 1u;
		}
		StackEntry16 * stack = stacks[ which ];
		if( leaf_phase ) lane_leaf_phase( map, l, stack, p.side_steps );
		else lane_node_phase( map, l, stack, p.max_steps );

		/* ---- retire finished segments */
		if( lane_done( l ) ) {
			HitState h;
			lane_hit_state( l, h );
			const ResultRec o = make_result( h, side_plane );
			if( p.out ) st_stream4( p.out + seg, o.t, __int_as_float( o.plane ), __int_as_float( o.brush ), __uint_as_float( o.flags ) );
			if( p.pos ) {
				float px, py, pz;
				make_pos( l.r, h, px, py, pz );                                 /* bsp.cc:265-267 */
				st_stream4( p.pos + seg, px, py, pz, h.t );
			}
			my_hits += ( uint32_t ) h.hit;
			l.cur = kCurIdle;
		}
	}

	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( kFull, my_hits, o );
	if( lane == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

/* ---------------------------------------------------------------- N3: batched position_to_leaf */

__global__ void leaf_kernel( const unsigned char * image, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf_out ) {
	GlobalMap map;
	map.nodes = reinterpret_cast< const NodeRec * >( image );
	map.refs = nullptr; map.sides = nullptr;
	for( uint64_t i = blockIdx.x * ( uint64_t ) blockDim.x + threadIdx.x; i < n; i += ( uint64_t ) gridDim.x * blockDim.x ) {
		const float * q = pts + i * stride;
		leaf_out[ i ] = point_leaf( map, q[ 0 ], q[ 1 ], q[ 2 ] );
	}
}

} // namespace bspk

#endif
