/*
 * trace_kernels.cuh -- sm_100a kernels of the batched segment trace.
 *
 * Both trace kernels are persistent: each warp claims a private block of consecutive segments with one global atomicAdd
 * (warp-level work fetch) and re-fills lanes whose segment has finished from that block as soon as `refill` lanes
 * are idle (__ballot_sync / __popc ranks = the ray compaction), so a warp keeps ~32 live segments instead of waiting
 * for its slowest one.  Lanes are resumable node/leaf state machines (trace_core.h); each warp iteration runs the one
 * bounded phase most of its lanes wait for.  A leaf step is one 32-byte gather + two brush sides.
 *
 *   trace_slots  (GLOBAL variant: the map is read through L1/L2)  lanes walk CHILD SLOTS (map_layout.h: Slot): a node
 *                step first asks for the 16-byte pair holding both children of its node, then runs the plane test on the
 *                plane it already carries, then selects near / far -- the gather overlaps the IEEE division instead of
 *                waiting for it (+4 % long rays, +9 % short moves on B200).  The first traversal-stack slots of every
 *                thread live in shared memory (SmemStack below); the warp's scheduling policy follows the kind of
 *                segments it drew last (short moves / long rays).
 *   trace_phased (SMEM variant: the map is staged in shared memory; also the walk experiments builds can A/B)  lanes
 *                walk node records: one 16-byte LDS per step (+ one for a node plane that is not axis-aligned).
 *
 * Map staging (template `kMode`):
 *   kSmemAll : the flattened map [nodes|gplanes|refs|sides_even|sides_odd] is copied into shared memory with
 *              cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) completing on an mbarrier; every record is then one
 *              dense 16-byte LDS.128                                                        (C1/C2)
 *   kGlobal  : nothing staged; records come through the read-only path (L1 -> L2)          (C3-C5)
 *   kSmemTop : (BSPGPU_EXPERIMENTS builds only) the first `top_nodes` breadth-first nodes staged, rest read-only path
 * K3  hit count: per-lane counters, one shuffle reduction + one atomicAdd per warp.
 * N3  leaf_kernel: batched BSP::position_to_leaf (reference bsp.cc:273-286) + cluster / PVS lookup
 *     (reference bsp_renderer.cc:147-168).
 * N4  bin_* kernels: optional counting sort of the segments by start-grid cell (a coherence pre-pass).
 *
 * Rays with a zero start coordinate or a non-finite component do not take the axis shortcut (trace_core.h:
 * ray_is_plain); the lane that draws one only counts it, and trace_exotic_kernel -- launched behind every trace
 * kernel, a no-op when the count is zero -- traces such rays with the general arithmetic.
 *
 * No tensor cores: the walk is a dependent chain of 16/32-byte gathers and a handful of fp32
 * ops per node/side, not a contraction.
 *
 * Round-1 experiments that are no longer in the tree (measured slower, profiles/README.md; last present at commit
 * 440db56): trace_persistent (round-based scheduler), trace_dual (two segments per lane), QNodeRec two-level records.
 */
#ifndef BSPGPU_TRACE_KERNELS_CUH
#define BSPGPU_TRACE_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "trace_core.h"

namespace bspk {

enum Mode { kSmemAll = 0, kSmemTop = 1, kGlobal = 2 };
static const uint32_t kExoticCap = 1u << 16;

struct TraceParams {
	const unsigned char * image;      /* flattened map in global memory */
	uint64_t off_gplanes, off_refs, off_sides_even, off_sides_odd, off_side_pairs, off_side_plane;
	uint32_t staged_bytes;            /* bytes copied to shared memory by this mode */
	uint32_t top_nodes;               /* kSmemTop: nodes [0, top_nodes) live in shared memory */
	const float4 * segs;              /* 2 float4 per segment ... */
	const float2 * segs24;            /* ... or 3 float2 (6 packed floats) per segment */
	const uint32_t * perm;            /* null, or: slot i of this launch traces segment perm[i] (segment binning, N4) */
	ResultRec * out;                  /* may be null */
	float4 * pos;                     /* may be null */
	uint2 * out8;                     /* may be null: compact results {t, flags | (plane+1) << 2} */
	float * out4;                     /* may be null: t on a hit, 2.0f on a miss (BSPGPU_OUT_COMPACT4) */
	uint32_t n;
	uint32_t block_segs;              /* segments a warp claims per global atomic */
	uint32_t refill;                  /* idle lanes that trigger a refill (1..32) */
	uint32_t max_steps;               /* node steps per lane per node phase (bounds divergence) */
	uint32_t side_steps;              /* side steps per lane per leaf phase */
	uint32_t leaf_threshold;          /* lanes waiting in a leaf that make the warp run a leaf phase */
	StartGrid grid;                   /* cells == null: every walk starts at node 0 */
	/* sections as absolute pointers (g_slot_pairs: the child slots in pairs, map_layout.h: Slot): under register pressure the
	 * compiler re-derives a base from the constant bank inside the hot loops, and image + offset is five instructions where a
	 * pointer parameter is two */
	const uint4 * g_slot_pairs; const PlaneRec * g_gplanes; const BrushRef * g_refs; const SideRec * g_side_pairs; const uint32_t * g_side_plane;
	Slot root;                        /* slot of node 0 */
	const Slot * grid_slots;          /* start grid as slots, or null */
	uint32_t max_steps_long;          /* node steps per phase while the warp's segments start at the root (long rays); max_steps otherwise */
	unsigned long long * next;        /* work counter, zeroed before launch; next[1] counts the segments left to trace_exotic_kernel */
	uint32_t * exotic_list;           /* the first kExoticCap of those segments' indices (beyond that the follow-up kernel rescans the launch) */
	unsigned long long * hits;        /* accumulated hit count */
	float4 * pool_overflow;           /* trace_pool: global scratch for traversal-stack entries beyond the shared-memory slots */
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
	static const bool kGeneralPlanes = true;
	uint32_t nodes, gplanes, refs, sides_even, sides_odd;   /* shared-window byte addresses of the sections */
	__device__ __forceinline__ void node( int32_t i, float & d, int32_t & c0, int32_t & c1, uint32_t & w ) const {
		const uint4 r = lds128u( nodes + ( ( uint32_t ) i << 4 ) );
		d = __uint_as_float( r.x ); c0 = ( int32_t ) r.y; c1 = ( int32_t ) r.z; w = r.w;
	}
	__device__ __forceinline__ void gplane( uint32_t i, float & nx, float & ny, float & nz ) const {
		const float4 p = lds128( gplanes + ( i << 4 ) );
		nx = p.x; ny = p.y; nz = p.z;
	}
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const uint4 r = lds128u( refs + ( i << 4 ) );
		first = r.x; nraw = r.y; brush = ( int32_t ) r.z;
	}
	/* sides i (even) and i+1 from the two dense arrays: eight lanes spread over eight bank groups in each */
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		const float4 p = lds128( sides_even + ( i << 3 ) ), q = lds128( sides_odd + ( i << 3 ) );
		n0x = p.x; n0y = p.y; n0z = p.z; d0 = p.w; n1x = q.x; n1y = q.y; n1z = q.z; d1 = q.w;
	}
};

/* two consecutive SideRecs (i even: 32-byte aligned) in one 256-bit read-only load (sm_100: LDG.E.ENL2.256.CONSTANT) */
__device__ __forceinline__ void ldg_side_pair( const SideRec * p, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) {
	asm volatile( "ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"( n0x ), "=f"( n0y ), "=f"( n0z ), "=f"( d0 ), "=f"( n1x ), "=f"( n1y ), "=f"( n1z ), "=f"( d1 ) : "l"( p ) );
}

struct GlobalMap {
	static const bool kGeneralPlanes = true;
	const NodeRec * nodes; const PlaneRec * gplanes; const BrushRef * refs; const SideRec * side_pairs;   /* global memory, read-only path */
	const NodeOrig * node_orig;
	const uint4 * slot_pairs;
	/* both child slots of a node in one 16-byte read-only gather */
	__device__ __forceinline__ void pair( uint32_t p, uint32_t & a0, uint32_t & x0, uint32_t & a1, uint32_t & x1 ) const {
		const uint4 r = __ldg( slot_pairs + p );
		a0 = r.x; x0 = r.y; a1 = r.z; x1 = r.w;
	}
	__device__ __forceinline__ void gplane_d( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const float4 p = __ldg( reinterpret_cast< const float4 * >( &gplanes[ i ] ) );
		nx = p.x; ny = p.y; nz = p.z; d = p.w;
	}
	__device__ __forceinline__ void node( int32_t i, float & d, int32_t & c0, int32_t & c1, uint32_t & w ) const {
		const uint4 r = __ldg( reinterpret_cast< const uint4 * >( &nodes[ i ] ) );
		d = __uint_as_float( r.x ); c0 = ( int32_t ) r.y; c1 = ( int32_t ) r.z; w = r.w;
	}
	__device__ __forceinline__ void gplane( uint32_t i, float & nx, float & ny, float & nz ) const {
		const float4 p = __ldg( reinterpret_cast< const float4 * >( &gplanes[ i ] ) );
		nx = p.x; ny = p.y; nz = p.z;
	}
	__device__ __forceinline__ int32_t orig_child( int32_t i, int k ) const { return __ldg( &node_orig[ i ].orig_child[ k ] ); }
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const {
		const uint4 r = __ldg( reinterpret_cast< const uint4 * >( &refs[ i ] ) );
		first = r.x; nraw = r.y; brush = ( int32_t ) r.z;
	}
	__device__ __forceinline__ void side( uint32_t i, float & nx, float & ny, float & nz, float & d ) const {
		const float4 p = __ldg( reinterpret_cast< const float4 * >( &side_pairs[ i ] ) );
		nx = p.x; ny = p.y; nz = p.z; d = p.w;
	}
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		ldg_side_pair( &side_pairs[ i ], n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};

__device__ __forceinline__ void bind_global( const TraceParams & p, GlobalMap & g ) {
	g.nodes = reinterpret_cast< const NodeRec * >( p.image );
	g.gplanes = p.g_gplanes;
	g.refs = p.g_refs;
	g.side_pairs = p.g_side_pairs;
	g.node_orig = nullptr;
	g.slot_pairs = p.g_slot_pairs;
}

#ifdef BSPGPU_EXPERIMENTS
struct SmemTopMap {
	static const bool kGeneralPlanes = true;
	uint32_t top;          /* shared-window byte address of nodes [0, top_nodes) */
	uint32_t top_nodes;
	GlobalMap g;
	__device__ __forceinline__ void node( int32_t i, float & d, int32_t & c0, int32_t & c1, uint32_t & w ) const {
		if( ( uint32_t ) i < top_nodes ) {
			const uint4 r = lds128u( top + ( ( uint32_t ) i << 4 ) );
			d = __uint_as_float( r.x ); c0 = ( int32_t ) r.y; c1 = ( int32_t ) r.z; w = r.w;
		}
		else g.node( i, d, c0, c1, w );
	}
	__device__ __forceinline__ void gplane( uint32_t i, float & nx, float & ny, float & nz ) const { g.gplane( i, nx, ny, nz ); }
	__device__ __forceinline__ void ref( uint32_t i, uint32_t & first, uint32_t & nraw, int32_t & brush ) const { g.ref( i, first, nraw, brush ); }
	__device__ __forceinline__ void side_pair( uint32_t i, float & n0x, float & n0y, float & n0z, float & d0, float & n1x, float & n1y, float & n1z, float & d1 ) const {
		g.side_pair( i, n0x, n0y, n0z, d0, n1x, n1y, n1z, d1 );
	}
};
#endif

/* kGeneral = false: the accessor of a map all of whose node planes are axis records (trace_core.h: lane_node_step) */
template< class Base, bool kGeneral > struct PlaneKinds : Base { static const bool kGeneralPlanes = kGeneral; };
template< int kMode, bool kGeneral > struct MapOf;
template< bool kGeneral > struct MapOf< kSmemAll, kGeneral > { typedef PlaneKinds< SmemAllMap, kGeneral > type; };
template< bool kGeneral > struct MapOf< kGlobal, kGeneral > { typedef PlaneKinds< GlobalMap, kGeneral > type; };
#ifdef BSPGPU_EXPERIMENTS
template< bool kGeneral > struct MapOf< kSmemTop, kGeneral > { typedef PlaneKinds< SmemTopMap, kGeneral > type; };
#endif

/* Stage what this mode keeps in shared memory -- one thread issues TMA bulk copies (cp.async.bulk, 32 KB each),
 * everyone waits on the mbarrier's transaction count -- and bind the map accessor to shared / global memory. */
template< int kMode, class Map >
__device__ __forceinline__ void stage_and_bind( const TraceParams & p, unsigned char * smem, uint64_t * stage_bar, Map & map ) {
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

	if constexpr( kMode == kSmemAll ) {
		map.nodes = opaque_u32( smem_u32( smem ) );
		map.gplanes = opaque_u32( map.nodes + ( uint32_t ) p.off_gplanes );
		map.refs = opaque_u32( map.nodes + ( uint32_t ) p.off_refs );
		map.sides_even = opaque_u32( map.nodes + ( uint32_t ) p.off_sides_even );
		map.sides_odd = opaque_u32( map.nodes + ( uint32_t ) p.off_sides_odd );
	}
#ifdef BSPGPU_EXPERIMENTS
	else if constexpr( kMode == kSmemTop ) {
		map.top = opaque_u32( smem_u32( smem ) );
		map.top_nodes = p.top_nodes;
		bind_global( p, map.g );
	}
#endif
	else bind_global( p, map );
}

/* the endpoints of segment `sidx` */
__device__ __forceinline__ void load_endpoints( const TraceParams & p, uint32_t sidx, float & sx, float & sy, float & sz, float & ex, float & ey, float & ez ) {
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
}

/* write the outputs of one finished segment (bsp.cc:265-270): 16-byte record, and / or position, and / or compact record */
__device__ __forceinline__ void store_result( const TraceParams & p, uint32_t seg, const Ray & r, const HitState & h, const uint32_t * side_plane ) {
	const ResultRec o = make_result( h, side_plane );
	if( p.out ) st_stream4( p.out + seg, o.t, __int_as_float( o.plane ), __int_as_float( o.brush ), __uint_as_float( o.flags ) );
	if( p.out8 ) {
		const uint32_t packed = o.flags | ( ( uint32_t ) ( o.plane + 1 ) << 2 );
		asm volatile( "st.global.cs.v2.b32 [%0], {%1,%2};" :: "l"( p.out8 + seg ), "r"( __float_as_uint( o.t ) ), "r"( packed ) : "memory" );
	}
	if( p.out4 ) asm volatile( "st.global.cs.f32 [%0], %1;" :: "l"( p.out4 + seg ), "f"( h.hit ? h.t : 2.0f ) : "memory" );
	if( p.pos ) {
		float px, py, pz;
		make_pos( r, h, px, py, pz );                                   /* bsp.cc:265-267 */
		st_stream4( p.pos + seg, px, py, pz, h.t );
	}
}

/* read segment `sidx` and begin its walk: at node 0 (bsp.cc:263) or at the start-grid reference of its cell.
 * A ray that is not plain (trace_core.h: ray_is_plain -- a zero start coordinate, an infinite or NaN component) cannot take
 * the axis shortcut: it is only counted here (next[1]) and left to trace_exotic_kernel, which runs right behind this kernel
 * and traces such rays start to finish with the general arithmetic.  Returns false for them; the lane stays idle. */
__device__ __forceinline__ bool lane_load_segment( const TraceParams & p, uint32_t sidx, Lane & l ) {
	float sx, sy, sz, ex, ey, ez;
	load_endpoints( p, sidx, sx, sy, sz, ex, ey, ez );
	/* the start reference first, the lane set-up once after it (one copy of the initialisation instead of one per grid outcome) */
	const int32_t ref = grid_start_ref( p.grid, sx, sy, sz, ex, ey, ez );   /* 0 when there is no grid */
	lane_start( l, sx, sy, sz, ex, ey, ez );
	lane_start_at( l, ref );
	if( !ray_is_plain( l.r ) ) {
		const unsigned long long k = atomicAdd( p.next + 1, 1ull );
		if( k < kExoticCap ) p.exotic_list[ k ] = sidx;
		l.cur = kCurIdle;
		return false;
	}
	return true;
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

/* the slot walk's entries {x, a, tmin, tmax}: all four words carry data */
template< int kSlots, int kThreads > struct SlotStackOf { typedef SmemStack< kSlots, kThreads * 16 > type; };
template< int kThreads > struct SlotStackOf< 0, kThreads > { typedef SlotEntry * type; };
template< int kSlots, int kStride >
__device__ __forceinline__ void sstack_put( const SmemStack< kSlots, kStride > & s, int32_t i, uint32_t x, uint32_t a, float tmin, float tmax ) {
	asm volatile( "{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %0, %7;\n\t@q st.shared.v4.b32 [%1], {%3,%4,%5,%6};\n\t@!q st.local.v4.b32 [%2], {%3,%4,%5,%6};\n\t}"
		:: "r"( i ), "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ), "l"( __cvta_generic_to_local( reinterpret_cast< SlotEntry * >( s.overflow ) + i ) ),
		   "r"( x ), "r"( a ), "f"( tmin ), "f"( tmax ), "n"( kSlots ) : "memory" );
}
template< int kSlots, int kStride >
__device__ __forceinline__ void sstack_get( const SmemStack< kSlots, kStride > & s, int32_t i, uint32_t & x, uint32_t & a, float & tmin, float & tmax ) {
	if( i < kSlots ) {
		asm volatile( "ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"( x ), "=r"( a ), "=f"( tmin ), "=f"( tmax )
			: "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ) : "memory" );
	}
	else sstack_get( reinterpret_cast< const SlotEntry * >( s.overflow ), i, x, a, tmin, tmax );
}

/* read segment `sidx` and begin its slot walk (see lane_load_segment) */
__device__ __forceinline__ bool slane_load_segment( const TraceParams & p, uint32_t sidx, LaneS & l ) {
	float sx, sy, sz, ex, ey, ez;
	load_endpoints( p, sidx, sx, sy, sz, ex, ey, ez );
	SlotGrid g; g.cells = p.grid_slots; g.d = p.grid.d;
	const Slot start = grid_start_slot( g, p.root, sx, sy, sz, ex, ey, ez );
	slane_start( l, sx, sy, sz, ex, ey, ez );
	slane_start_at( l, start );
	if( !ray_is_plain( l.r ) ) {
		const unsigned long long k = atomicAdd( p.next + 1, 1ull );
		if( k < kExoticCap ) p.exotic_list[ k ] = sidx;
		l.cx = kSlotIdle;
		return false;
	}
	return true;
}

/* ---------------------------------------------------------------- K1/K2 over child slots: the same scheduling as trace_phased
 * below, lanes walk Slot pairs (trace_core.h: LaneS) */
template< int kMode, int kStack, int kMaxThreads, int kMinBlocks, int kSmemSlots, bool kGeneral >
__global__ void __launch_bounds__( kMaxThreads, kMinBlocks ) trace_slots( const __grid_constant__ TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;
	/* the warp's claimed block of segments {next, end}: warp-uniform, touched only at a refill -- kept out of the register file */
	__shared__ uint2 warp_blk[ kMaxThreads / 32 ];

	typename MapOf< kMode, kGeneral >::type map;
	stage_and_bind< kMode >( p, smem, &stage_bar, map );

	const unsigned kFull = 0xffffffffu;
	const uint32_t warp_slot = opaque_u32( smem_u32( &warp_blk[ threadIdx.x >> 5 ] ) );
	asm volatile( "st.shared.v2.u32 [%0], {%1,%1};" :: "r"( warp_slot ), "r"( 0u ) : "memory" );
	uint32_t refill_at = p.refill;          /* idle lanes that trigger a refill; 33 once the launch has no segments left */
	bool long_mode = true;                  /* scheduling follows the kind of segments the warp drew last (see the refill) */

	LaneS l;
	slane_start( l, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f );
	l.ca = 0; l.cx = kSlotIdle;
	SlotEntry stack_local[ kStack ];
	typename SlotStackOf< kSmemSlots, kMaxThreads >::type stack;
	if constexpr( kSmemSlots > 0 ) {
		stack.base = opaque_u32( smem_u32( smem ) + ( ( p.staged_bytes + 127u ) & ~127u ) + threadIdx.x * 16u );
		stack.overflow = reinterpret_cast< StackEntry16 * >( stack_local );   /* SmemStack is shared with trace_phased; both entry types are 16 bytes */
	}
	else stack = stack_local;
	uint32_t seg = 0, my_hits = 0;

	for( ;; ) {
		/* ---- retire finished segments (also those the start grid resolved at load time: a segment inside an empty leaf) */
		if( slane_done( l ) ) {
			HitState h;
			slane_hit_state( l, h );
			store_result( p, seg, l.r, h, p.g_side_plane );
			my_hits += ( uint32_t ) h.hit;
			l.cx = kSlotIdle;
		}

		/* ---- warp-level work fetch + compaction */
		const unsigned idle = __ballot_sync( kFull, slane_idle( l ) );
		const unsigned n_idle = __popc( idle );
		if( n_idle >= refill_at ) {
			uint32_t blk_next, blk_end;
			asm volatile( "ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"( blk_next ), "=r"( blk_end ) : "r"( warp_slot ) : "memory" );
			if( blk_next == blk_end ) {
				unsigned long long b = 0;
				if( ( threadIdx.x & 31u ) == 0 ) b = atomicAdd( p.next, ( unsigned long long ) p.block_segs );
				b = __shfl_sync( kFull, b, 0 );
				if( b >= ( unsigned long long ) p.n ) refill_at = 33u;
				else {
					blk_next = ( uint32_t ) b;
					blk_end = ( uint32_t ) min( b + ( unsigned long long ) p.block_segs, ( unsigned long long ) p.n );
				}
			}
			if( refill_at <= 32u ) {
				const uint32_t avail = blk_end - blk_next;
				const uint32_t rank = __popc( idle & ( ( 1u << ( threadIdx.x & 31u ) ) - 1u ) );
				const bool take = slane_idle( l ) && rank < avail;
				if( take ) {
					seg = blk_next + rank;
					if( p.perm ) seg = __ldg( p.perm + seg );                       /* binned order (N4): results still land at the segment's own index */
					slane_load_segment( p, seg, l );
				}
				/* what kind of batch is this?  Segments the start grid placed below the root are short moves: a few node steps
				 * each, so long node phases run half empty and waiting leaf work should go first; segments that start at the
				 * root are long rays, where the opposite holds (measured: profiles/README.md).  The warp follows the majority
				 * of the segments it has just drawn; results do not depend on it. */
				const unsigned at_root = __ballot_sync( kFull, take && l.cx == p.root.x && l.ca == p.root.a );
				long_mode = 4u * __popc( at_root ) >= 3u * min( n_idle, avail );
				blk_next += min( n_idle, avail );
				/* every lane stores the same pair, so each reads back its own store: no warp barrier needed */
				asm volatile( "st.shared.v2.u32 [%0], {%1,%2};" :: "r"( warp_slot ), "r"( blk_next ), "r"( blk_end ) : "memory" );
			}
		}

		/* ---- pick the phase most lanes are waiting for */
		const unsigned in_node = __ballot_sync( kFull, slane_in_node( l ) );
		const unsigned in_leaf = __ballot_sync( kFull, slane_in_leaf( l ) );
		if( ( in_node | in_leaf ) == 0u ) {
			if( __ballot_sync( kFull, slane_done( l ) ) != 0u ) continue;       /* loaded and already finished: retire them first */
			if( refill_at > 32u ) break;
			continue;
		}
		const uint32_t n_leaf = __popc( in_leaf );
		const bool leaf_first = long_mode ? n_leaf >= p.leaf_threshold : 2u * n_leaf >= ( uint32_t ) __popc( in_node );
		if( in_node == 0u || leaf_first ) slane_leaf_phase( map, l, stack, p.side_steps );
		else slane_node_phase( map, l, stack, long_mode ? p.max_steps_long : p.max_steps );
	}

	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( kFull, my_hits, o );
	if( ( threadIdx.x & 31u ) == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
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
template< int kMode, int kStack, int kMaxThreads, int kMinBlocks, int kSmemSlots, bool kGeneral >
__global__ void __launch_bounds__( kMaxThreads, kMinBlocks ) trace_phased( const __grid_constant__ TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;

	typename MapOf< kMode, kGeneral >::type map;
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
			store_result( p, seg, l.r, h, side_plane );
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
					if( p.perm ) seg = __ldg( p.perm + seg );                       /* binned order (N4): results still land at the segment's own index */
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

#ifdef BSPGPU_EXPERIMENTS
/* ---------------------------------------------------------------- one thread per segment, no scheduling
 *
 * BSPGPU_OPT_SCHEDULER=4 (experiments build).  Grid-stride loop, each thread walks its segment start to finish.  No
 * work fetch, no phases, no compaction: the naive baseline the shipped kernel is measured against. */
template< int kMode, int kStack, int kMaxThreads >
__global__ void __launch_bounds__( kMaxThreads ) trace_simple( const __grid_constant__ TraceParams p ) {
	static_assert( kMode == kGlobal, "the simple kernel reads the map through the read-only path" );
	GlobalMap map;
	bind_global( p, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );
	StackEntry16 stack_local[ kStack ];
	StackEntry16 * const stack = stack_local;
	uint32_t my_hits = 0;
	for( uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x; seg < p.n; seg += gridDim.x * blockDim.x ) {
		Lane l;
		if( !lane_load_segment( p, seg, l ) ) continue;
		while( !lane_done( l ) ) {
			lane_node_phase( map, l, stack, 0xffffffffu );
			lane_leaf_phase( map, l, stack, 0xffffffffu );
		}
		HitState h;
		lane_hit_state( l, h );
		store_result( p, seg, l.r, h, side_plane );
		my_hits += ( uint32_t ) h.hit;
	}
	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( 0xffffffffu, my_hits, o );
	if( ( threadIdx.x & 31u ) == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}
#endif

/* ---------------------------------------------------------------- the rays the axis shortcut does not cover
 *
 * Launched right behind every trace kernel.  When that kernel counted no such ray (the normal case) it returns at once;
 * otherwise it re-reads the launch's segments, and every ray that is not plain is traced start to finish, from the root,
 * with the reference's own arithmetic at every node (node_plane_exact: an axis record becomes the unit normal it stands for).
 * One thread per segment, no scheduling: these rays are rare outside of tests. */
__global__ void __launch_bounds__( 256 ) trace_exotic_kernel( const __grid_constant__ TraceParams p ) {
	/* one load per CTA; L1 starts out invalid at a launch boundary, so the plain load sees the count of the kernel before */
	__shared__ unsigned long long todo_s;
	if( threadIdx.x == 0 ) todo_s = p.next[ 1 ];
	__syncthreads();
	const unsigned long long todo = todo_s;
	if( todo == 0ull ) return;
	GlobalMap map;
	bind_global( p, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );
	/* few such rays (the usual case: a coordinate that is exactly 0 turns up about once per 2^24 random floats): their indices
	 * were listed.  More than the list holds: re-read the launch's segments and pick them out again. */
	const bool listed = todo <= ( unsigned long long ) kExoticCap;
	const uint32_t count = listed ? ( uint32_t ) todo : p.n;
	uint32_t my_hits = 0;
	for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x ) {
		const uint32_t seg = listed ? p.exotic_list[ i ] : ( p.perm ? __ldg( p.perm + i ) : i );
		float sx, sy, sz, ex, ey, ez;
		load_endpoints( p, seg, sx, sy, sz, ex, ey, ez );
		Ray r; r.sx = sx; r.sy = sy; r.sz = sz; r.dx = ex - sx; r.dy = ey - sy; r.dz = ez - sz;      /* bsp.cc:261 */
		if( ray_is_plain( r ) ) continue;
		HitState h;
		trace_one< GlobalMap, 128 >( map, r, h );
		store_result( p, seg, r, h, side_plane );
		my_hits += ( uint32_t ) h.hit;
	}
	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( 0xffffffffu, my_hits, o );
	if( ( threadIdx.x & 31u ) == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

/* ---------------------------------------------------------------- a handful of segments, one launch, no copies
 *
 * BSP::trace_seg -- the reference's only call shape (bsp.cc:365: one segment per frame) -- is a batch of ONE.  For up to
 * kTinyRays host segments the library launches this kernel alone: one CTA, one WARP per segment (lane 0 walks, so no two
 * rays share a SIMT instruction stream), segments read from and results written to the handle's pinned buffer directly over
 * the host link (no memcpy, no memset, no second kernel), the hit count STORED by the CTA.  Every ray is traced start to
 * finish with the reference's own arithmetic (trace_one, the general form): nothing to schedule, no shortcut to qualify for.
 * Measured from Python: 46 -> 27 us for a batch of one. */
static const uint32_t kTinyRays = 16;
__global__ void __launch_bounds__( kTinyRays * 32 ) trace_tiny_kernel( const __grid_constant__ TraceParams p ) {
	__shared__ unsigned int total;
	if( threadIdx.x == 0 ) total = 0;
	__syncthreads();
	const uint32_t ray = threadIdx.x >> 5;
	if( ( threadIdx.x & 31u ) == 0 && ray < p.n ) {
		GlobalMap map;
		bind_global( p, map );
		float sx, sy, sz, ex, ey, ez;
		load_endpoints( p, ray, sx, sy, sz, ex, ey, ez );
		Ray r; r.sx = sx; r.sy = sy; r.sz = sz; r.dx = ex - sx; r.dy = ey - sy; r.dz = ez - sz;      /* bsp.cc:261 */
		HitState h;
		trace_one< GlobalMap, 128 >( map, r, h );
		store_result( p, ray, r, h, p.g_side_plane );
		if( h.hit ) atomicAdd( &total, 1u );
	}
	__syncthreads();
	if( threadIdx.x == 0 ) *p.hits = ( unsigned long long ) total;
}

/* ---------------------------------------------------------------- N3: batched position_to_leaf */

struct LeafParams {
	const unsigned char * image;
	uint64_t off_gplanes, off_node_orig, off_leaf_cluster;
	const float * pts; uint32_t stride; uint64_t n;
	int32_t * leaf_out;       /* may be null */
	int32_t * cluster_out;    /* may be null: BSP_Leaf::cluster of the point's leaf (bsp.h:68) */
	/* PVS test (reference bsp_renderer.cc:147-168): visible_out[i] = bit `cluster` of row `from_cluster` of the vis lump */
	const unsigned char * vis_rows; uint32_t num_clusters, cluster_size;
	int32_t from_cluster;
	unsigned char * visible_out;   /* may be null */
};

__global__ void leaf_kernel( const LeafParams p ) {
	GlobalMap map;
	map.nodes = reinterpret_cast< const NodeRec * >( p.image );
	map.gplanes = reinterpret_cast< const PlaneRec * >( p.image + p.off_gplanes );
	map.node_orig = reinterpret_cast< const NodeOrig * >( p.image + p.off_node_orig );
	map.refs = nullptr; map.side_pairs = nullptr; map.slot_pairs = nullptr;
	const int32_t * leaf_cluster = reinterpret_cast< const int32_t * >( p.image + p.off_leaf_cluster );
	for( uint64_t i = blockIdx.x * ( uint64_t ) blockDim.x + threadIdx.x; i < p.n; i += ( uint64_t ) gridDim.x * blockDim.x ) {
		const float * q = p.pts + i * p.stride;
		const int32_t leaf = point_leaf( map, q[ 0 ], q[ 1 ], q[ 2 ] );
		if( p.leaf_out ) p.leaf_out[ i ] = leaf;
		if( p.cluster_out || p.visible_out ) {
			const int32_t cluster = __ldg( leaf_cluster + leaf );
			if( p.cluster_out ) p.cluster_out[ i ] = cluster;
			if( p.visible_out ) {
				/* bsp_renderer.cc:150-165: no PVS row for the viewer (cluster < 0) => everything is drawn; a leaf outside every
				 * cluster is never drawn; else bit (cluster % 8) of byte cluster / 8 of the viewer's row */
				unsigned char vis = 1;
				if( p.from_cluster >= 0 ) {
					vis = 0;
					if( cluster >= 0 && ( uint32_t ) cluster < p.num_clusters && ( uint32_t ) ( cluster >> 3 ) < p.cluster_size )
						vis = ( __ldg( p.vis_rows + ( size_t ) p.from_cluster * p.cluster_size + ( cluster >> 3 ) ) >> ( cluster & 7 ) ) & 1;
				}
				p.visible_out[ i ] = vis;
			}
		}
	}
}

/* the loop of bspr_render (bsp_renderer.cc:158-167) over every leaf of the leaves lump, seen from cluster `from` (>= 0) */
__global__ void leaf_vis_kernel( const int32_t * leaf_cluster, uint32_t num_leaves, const unsigned char * vis_rows, uint32_t num_clusters, uint32_t cluster_size,
                                 int32_t from, unsigned char * visible ) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if( i >= num_leaves ) return;
	const int32_t c = leaf_cluster[ i ];
	unsigned char vis = 0;
	if( c >= 0 && ( uint32_t ) c < num_clusters && ( uint32_t ) ( c >> 3 ) < cluster_size )
		vis = ( vis_rows[ ( size_t ) from * cluster_size + ( c >> 3 ) ] >> ( c & 7 ) ) & 1;
	visible[ i ] = vis;
}

} // namespace bspk

#endif
