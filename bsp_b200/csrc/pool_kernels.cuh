/*
 * pool_kernels.cuh -- trace_pool: the phase-scheduled trace with CTA-level ray queues (BSPGPU_OPT_SCHEDULER = 5).
 *
 * trace_phased (trace_kernels.cuh) keeps every ray in the lane that loaded it, so a warp's 32 rays are split between
 * the two kinds of work (tree descent, brush clipping) and each phase runs with the lanes of one kind only: ncu shows
 * 17-18 of 32 lanes per instruction while the issue slots are 75-83 % busy.  Here the rays that wait for the OTHER kind
 * of work do not sit in a lane.  Every ray in flight owns a slot in a shared-memory pool (its 48-byte travelling state
 * -- ray, interval, position in the tree, stack height -- plus the first slots of its traversal stack); three CTA-wide
 * queues of slot numbers (free slots, rays waiting for node steps, rays waiting for a leaf) connect the warps:
 *
 *   node role   lanes step nodes.  A ray that reaches a leaf with solid brushes is written to its slot and queued for
 *               the leaf role; the lane is re-filled from the node queue (rays coming back from a leaf), then with new
 *               segments.  When 32 rays wait for a leaf, the warp parks its node rays in the node queue and takes them.
 *   leaf role   32 lanes clip brush sides.  A hit retires the ray (bsp.cc:206); a miss pops the ray's stack and queues
 *               it for the node role.  When no lane has a leaf left the warp goes back to the node role.
 *
 * The per-lane arithmetic is trace_core.h's, unchanged (lane_node_phase / lane_leaf_phase): only where a ray waits differs.
 */
#ifndef BSPGPU_POOL_KERNELS_CUH
#define BSPGPU_POOL_KERNELS_CUH

#include "trace_kernels.cuh"

namespace bspk {

static const uint32_t kPoolOverflowDepth = 32;   /* stack entries beyond the shared-memory slots: global scratch, indexed by the absolute slot */
static const uint32_t kPoolScratchPerRay = kPoolOverflowDepth + 3;   /* + the brush cursor of a ray parked in the middle of a leaf (3 x 16 bytes) */
static const uint32_t kParkedFlag = 0x80000000u; /* in the travelling hflags word: the cursor chunks are valid */
static const uint32_t kPoolSpinLimit = 1u << 22; /* watchdogs: a protocol error ends the kernel with wrong results instead of hanging the GPU */
static const uint32_t kPoolIterLimit = 1u << 28;

/* traversal stack of the ray in pool slot `id`: the first kS entries in shared memory ([entry][slot], 16 bytes each),
 * deeper ones in global scratch */
template< int kS, int kStride >
struct PoolStack {
	uint32_t base;              /* shared-window byte address of entry 0 of this ray */
	float4 * overflow;          /* global memory, indexed by the absolute entry number */
};
template< int kS, int kStride >
__device__ __forceinline__ void stack_put( const PoolStack< kS, kStride > & s, int32_t i, int32_t node, float tmin, float tmax ) {
	/* both stores name the same registers (see SmemStack) */
	uint32_t pad;
	asm( "" : "=r"( pad ) );
	asm volatile( "{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %0, %7;\n\t@q st.shared.v4.b32 [%1], {%3,%4,%5,%6};\n\t@!q st.global.v4.b32 [%2], {%3,%4,%5,%6};\n\t}"
		:: "r"( i ), "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ), "l"( s.overflow + i ),
		   "r"( node ), "f"( tmin ), "f"( tmax ), "r"( pad ), "n"( kS ) : "memory" );
}
template< int kS, int kStride >
__device__ __forceinline__ void stack_get( const PoolStack< kS, kStride > & s, int32_t i, int32_t & node, float & tmin, float & tmax ) {
	if( i < kS ) {
		int32_t pad;
		asm volatile( "ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"( node ), "=f"( tmin ), "=f"( tmax ), "=r"( pad )
			: "r"( s.base + ( uint32_t ) i * ( uint32_t ) kStride ) : "memory" );
	}
	else {
		int32_t pad;
		asm volatile( "ld.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"( node ), "=f"( tmin ), "=f"( tmax ), "=r"( pad ) : "l"( s.overflow + i ) : "memory" );
	}
}

/* multi-producer multi-consumer ring of slot numbers in shared memory.  An entry is (slot + 1); 0 = empty.  `avail` counts
 * published entries; positions are handed out by head / tail, and a consumer whose position has not been written yet
 * (a slower producer reserved it earlier) spins on the entry itself.  At most kR slot numbers exist, the ring holds
 * more, so a producer can only meet a non-empty entry that a consumer is about to take. */
struct PoolQueue {
	int * avail;
	unsigned * head, * tail;
	unsigned * e;
	unsigned mask;
};

__device__ __forceinline__ int warp_read( const int * p, unsigned lane ) {
	int v = 0;
	if( lane == 0 ) v = *( volatile const int * ) p;
	return __shfl_sync( 0xffffffffu, v, 0 );
}

/* every lane of the warp calls; lanes with `pred` append `id` */
__device__ __forceinline__ void q_push( const PoolQueue & q, bool pred, int id, unsigned lane, unsigned lt_mask ) {
	const unsigned m = __ballot_sync( 0xffffffffu, pred );
	if( m == 0u ) return;
	const unsigned k = __popc( m );
	unsigned pos = 0;
	if( lane == 0 ) pos = atomicAdd( q.tail, k );
	pos = __shfl_sync( 0xffffffffu, pos, 0 );
	__threadfence_block();                                   /* the ray's state is in its slot before its number is in the queue */
	if( pred ) {
		const unsigned i = ( pos + __popc( m & lt_mask ) ) & q.mask;
		uint32_t spin = 0;
		while( atomicCAS( q.e + i, 0u, ( unsigned ) id + 1u ) != 0u && ++spin < kPoolSpinLimit ) { }
	}
	__threadfence_block();
	__syncwarp();
	if( lane == 0 ) atomicAdd( q.avail, ( int ) k );
}

/* every lane calls; takes up to `want` entries (all or nothing when `exact`), hands them to the lanes of `want_m` in lane order.
 * Returns the number taken; a lane that got one has id >= 0 afterwards. */
__device__ __forceinline__ unsigned q_pop( const PoolQueue & q, unsigned want, bool exact, unsigned want_m, int & id, unsigned lane, unsigned lt_mask ) {
	unsigned k = 0, pos = 0;
	if( lane == 0 ) {
		const int a = *( volatile int * ) q.avail;
		if( a > 0 ) {
			k = min( want, ( unsigned ) a );
			if( exact && k < want ) k = 0;
			if( k ) {
				const int old = atomicSub( q.avail, ( int ) k );
				if( old < ( int ) k ) { atomicAdd( q.avail, ( int ) k ); k = 0; }   /* another warp was faster */
			}
			if( k ) pos = atomicAdd( q.head, k );
		}
	}
	k = __shfl_sync( 0xffffffffu, k, 0 );
	if( k == 0u ) return 0u;
	pos = __shfl_sync( 0xffffffffu, pos, 0 );
	const unsigned rank = __popc( want_m & lt_mask );
	if( ( ( want_m >> lane ) & 1u ) && rank < k ) {
		const unsigned i = ( pos + rank ) & q.mask;
		unsigned v = 0; uint32_t spin = 0;
		while( ( v = atomicExch( q.e + i, 0u ) ) == 0u && ++spin < kPoolSpinLimit ) { }
		id = ( int ) v - 1;
	}
	__threadfence_block();
	return k;
}

/* the travelling state of a ray: what a lane needs to resume it in either role.  Three 16-byte chunks, [chunk][slot].
 * The running hit (ht, hside, hbrush) never travels: a hit ends the ray in the leaf phase that found it (bsp.cc:206);
 * only the start-solid bit (hflags) outlives a leaf. */
template< int kR >
__device__ __forceinline__ void pool_save( uint32_t pool0, int id, const Lane & l, uint32_t seg, uint32_t flags ) {
	const uint32_t a = pool0 + ( uint32_t ) id * 16u;
	asm volatile( "st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"( a ), "f"( l.r.sx ), "f"( l.r.sy ), "f"( l.r.sz ), "f"( l.tmin ) : "memory" );
	asm volatile( "st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"( a + ( uint32_t ) kR * 16u ), "f"( l.r.dx ), "f"( l.r.dy ), "f"( l.r.dz ), "f"( l.tmax ) : "memory" );
	asm volatile( "st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"( a + ( uint32_t ) kR * 32u ), "r"( l.cur ), "r"( l.sp ), "r"( seg ), "r"( l.hflags | flags ) : "memory" );
}
/* a ray parked in the middle of a leaf also keeps its brush cursor and running hit (global scratch: rare) */
__device__ __forceinline__ void pool_save_cursor( float4 * c, const Lane & l ) {
	c[ 0 ] = make_float4( __uint_as_float( l.s ), __uint_as_float( l.s_end ), __int_as_float( l.brush ), __uint_as_float( l.bflags ) );
	c[ 1 ] = make_float4( l.tfar, l.tnear, __int_as_float( l.far_side ), l.ht );
	c[ 2 ] = make_float4( __int_as_float( l.hside ), __int_as_float( l.hbrush ), 0.0f, 0.0f );
}
template< int kR >
__device__ __forceinline__ void pool_load( uint32_t pool0, int id, Lane & l, uint32_t & seg, const float4 * cursor ) {
	const uint32_t a = pool0 + ( uint32_t ) id * 16u;
	const float4 c0 = lds128( a ), c1 = lds128( a + ( uint32_t ) kR * 16u );
	const uint4 c2 = lds128u( a + ( uint32_t ) kR * 32u );
	l.r.sx = c0.x; l.r.sy = c0.y; l.r.sz = c0.z; l.tmin = c0.w;
	l.r.dx = c1.x; l.r.dy = c1.y; l.r.dz = c1.z; l.tmax = c1.w;
	l.cur = ( int32_t ) c2.x; l.sp = ( int32_t ) c2.y; seg = c2.z; l.hflags = c2.w & ~kParkedFlag;
	l.ht = 1.0f; l.hside = -1; l.hbrush = -1;                      /* no hit so far, or the ray would have ended (bsp.cc:206,257) */
	l.s = 0; l.s_end = 0;                                          /* in a leaf: fetch its first brush reference */
	l.brush = -1; l.bflags = 0; l.tfar = 0.0f; l.tnear = 0.0f; l.far_side = -1;
	if( c2.w & kParkedFlag ) {
		const float4 k0 = cursor[ 0 ], k1 = cursor[ 1 ], k2 = cursor[ 2 ];
		l.s = __float_as_uint( k0.x ); l.s_end = __float_as_uint( k0.y ); l.brush = __float_as_int( k0.z ); l.bflags = __float_as_uint( k0.w );
		l.tfar = k1.x; l.tnear = k1.y; l.far_side = __float_as_int( k1.z ); l.ht = k1.w;
		l.hside = __float_as_int( k2.x ); l.hbrush = __float_as_int( k2.y );
	}
}

/* shared memory of one CTA behind the staged map: [state 3 x kR x 16][stack kS x kR x 16][3 rings x kCap x 4][control] */
template< int kR, int kS > struct PoolLayout {
	static const uint32_t kCap = kR <= 256 ? 512u : ( kR <= 512 ? 1024u : 2048u );      /* power of two > kR */
	static const uint32_t kStateBytes = 3u * kR * 16u, kStackBytes = ( uint32_t ) kS * kR * 16u, kRingBytes = kCap * 4u;
	static const uint32_t kBytes = kStateBytes + kStackBytes + 3u * kRingBytes + 64u;
};

template< int kMode, int kThreads, int kMinBlocks, int kR, int kS, bool kGeneral >
__global__ void __launch_bounds__( kThreads, kMinBlocks ) trace_pool( const __grid_constant__ TraceParams p ) {
	extern __shared__ __align__( 128 ) unsigned char smem[];
	__shared__ __align__( 8 ) uint64_t stage_bar;
	typedef PoolLayout< kR, kS > L;
	static_assert( kR >= kThreads, "every thread starts with a slot of its own" );

	typename MapOf< kMode, kGeneral >::type map;
	stage_and_bind< kMode >( p, smem, &stage_bar, map );
	const uint32_t * side_plane = reinterpret_cast< const uint32_t * >( p.image + p.off_side_plane );

	const uint32_t pool_off = ( p.staged_bytes + 127u ) & ~127u;
	const uint32_t pool0 = opaque_u32( smem_u32( smem ) + pool_off );
	const uint32_t stack0 = opaque_u32( pool0 + L::kStateBytes );
	unsigned * const rings = reinterpret_cast< unsigned * >( smem + pool_off + L::kStateBytes + L::kStackBytes );
	int * const ctrl = reinterpret_cast< int * >( rings + 3u * L::kCap );
	PoolQueue freeq, nodeq, leafq;
	freeq.e = rings; nodeq.e = rings + L::kCap; leafq.e = rings + 2u * L::kCap;
	freeq.mask = nodeq.mask = leafq.mask = L::kCap - 1u;
	freeq.avail = ctrl + 0; freeq.head = reinterpret_cast< unsigned * >( ctrl + 1 ); freeq.tail = reinterpret_cast< unsigned * >( ctrl + 2 );
	nodeq.avail = ctrl + 4; nodeq.head = reinterpret_cast< unsigned * >( ctrl + 5 ); nodeq.tail = reinterpret_cast< unsigned * >( ctrl + 6 );
	leafq.avail = ctrl + 8; leafq.head = reinterpret_cast< unsigned * >( ctrl + 9 ); leafq.tail = reinterpret_cast< unsigned * >( ctrl + 10 );
	int * const in_flight = ctrl + 12;

	/* slots [0, blockDim) start in the hands of the threads, the rest in the free ring */
	for( uint32_t i = threadIdx.x; i < 3u * L::kCap; i += blockDim.x ) {
		const uint32_t spare = ( uint32_t ) kR - blockDim.x;
		rings[ i ] = i < spare ? blockDim.x + i + 1u : 0u;
	}
	if( threadIdx.x < 16 ) ctrl[ threadIdx.x ] = 0;
	__syncthreads();
	if( threadIdx.x == 0 ) { *freeq.avail = ( int ) ( ( uint32_t ) kR - blockDim.x ); *freeq.tail = ( uint32_t ) kR - blockDim.x; }
	__syncthreads();

	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = ( 1u << lane ) - 1u;
	const unsigned kFull = 0xffffffffu;

	uint32_t blk_next = 0, blk_end = 0;
	bool more_work = true, leaf_role = false, drain = false;   /* drain: this warp clips because nothing else is left for it -- it never parks */
	Lane l;
	lane_start( l, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f );
	l.cur = kCurIdle;
	int id = ( int ) threadIdx.x;
	uint32_t seg = 0, my_hits = 0;
	float4 * const scratch_cta = p.pool_overflow + ( size_t ) blockIdx.x * kR * kPoolScratchPerRay;
	PoolStack< kS, kR * 16 > stack;
#define POOL_SCRATCH( id_ ) ( scratch_cta + ( size_t ) ( uint32_t ) ( id_ ) * kPoolScratchPerRay )
#define POOL_LOAD( id_ ) pool_load< kR >( pool0, id_, l, seg, POOL_SCRATCH( id_ ) + kPoolOverflowDepth )

	for( uint32_t iter = 0; iter < kPoolIterLimit; iter++ ) {
		if( leaf_role ) {
			unsigned live = __ballot_sync( kFull, lane_in_leaf( l ) );
			if( 32u - __popc( live ) >= p.refill || live == 0u ) {
				/* ---- rays that ended here are written out and their slots go back to the pool; rays that popped a node go back to the tree */
				const bool done = lane_done( l );
				if( done ) {
					HitState h;
					lane_hit_state( l, h );
					store_result( p, seg, l.r, h, side_plane );
					my_hits += ( uint32_t ) h.hit;
					l.cur = kCurIdle;
				}
				const unsigned done_m = __ballot_sync( kFull, done );
				if( done_m != 0u && lane == 0 ) atomicSub( in_flight, ( int ) __popc( done_m ) );
				q_push( freeq, done, id, lane, lt_mask );
				if( done ) id = -1;
				const bool to_node = lane_in_node( l );
				if( __ballot_sync( kFull, to_node ) != 0u ) {
					if( to_node ) pool_save< kR >( pool0, id, l, seg, 0u );
					q_push( nodeq, to_node, id, lane, lt_mask );
					if( to_node ) { id = -1; l.cur = kCurIdle; }
				}
				/* ---- take what waits for a leaf */
				if( warp_read( leafq.avail, lane ) > 0 ) {
					const int before = id;
					q_pop( leafq, 32u - __popc( live ), false, ~live, id, lane, lt_mask );
					if( id != before ) POOL_LOAD( id );
					live = __ballot_sync( kFull, lane_in_leaf( l ) );
				}
				/* ---- a few long leaves left and nothing waiting: park them (with their brush cursor) and go back to the tree */
				if( live == 0u ) { leaf_role = false; drain = false; continue; }
				if( !drain && __popc( live ) < p.leaf_threshold ) {
					const bool park = lane_in_leaf( l );
					if( park ) {
						pool_save< kR >( pool0, id, l, seg, kParkedFlag );
						pool_save_cursor( POOL_SCRATCH( id ) + kPoolOverflowDepth, l );
					}
					q_push( leafq, park, id, lane, lt_mask );
					if( park ) { id = -1; l.cur = kCurIdle; }
					leaf_role = false;
					continue;
				}
			}
			stack.base = stack0 + ( uint32_t ) id * 16u;
			stack.overflow = POOL_SCRATCH( id );
			lane_leaf_phase( map, l, stack, p.side_steps );
			continue;
		}

		/* ---- node role */
		unsigned in_node = __ballot_sync( kFull, lane_in_node( l ) );
		if( 32u - __popc( in_node ) >= p.refill || in_node == 0u ) {
			const bool done = lane_done( l );
			if( done ) {
				HitState h;
				lane_hit_state( l, h );
				store_result( p, seg, l.r, h, side_plane );
				my_hits += ( uint32_t ) h.hit;
				l.cur = kCurIdle;                                    /* the slot stays with the lane for the next segment */
			}
			const unsigned done_m = __ballot_sync( kFull, done );
			if( done_m != 0u && lane == 0 ) atomicSub( in_flight, ( int ) __popc( done_m ) );

			/* rays that reached a leaf with solid brushes wait for the leaf role in the queue, not in this lane */
			const bool to_leaf = lane_in_leaf( l );
			if( __ballot_sync( kFull, to_leaf ) != 0u ) {
				if( to_leaf ) pool_save< kR >( pool0, id, l, seg, 0u );
				q_push( leafq, to_leaf, id, lane, lt_mask );
				if( to_leaf ) { id = -1; l.cur = kCurIdle; }
			}

			/* a full warp of rays waits for a leaf: park the node rays and take them */
			if( warp_read( leafq.avail, lane ) >= 32 ) {
				int nid = -1;
				if( q_pop( leafq, 32u, true, kFull, nid, lane, lt_mask ) == 32u ) {
					const bool nd = lane_in_node( l );
					if( nd ) pool_save< kR >( pool0, id, l, seg, 0u );
					q_push( nodeq, nd, id, lane, lt_mask );
					q_push( freeq, !nd && id >= 0, id, lane, lt_mask );
					id = nid;
					POOL_LOAD( id );
					leaf_role = true;
					continue;
				}
			}

			/* refill idle lanes: rays coming back from a leaf first, then new segments */
			const unsigned want_m = __ballot_sync( kFull, lane_idle( l ) && id < 0 );
			if( want_m != 0u && warp_read( nodeq.avail, lane ) > 0 ) {
				const int before = id;
				q_pop( nodeq, __popc( want_m ), false, want_m, id, lane, lt_mask );
				if( id != before ) POOL_LOAD( id );
			}
			if( more_work ) {
				const unsigned need_m = __ballot_sync( kFull, lane_idle( l ) && id < 0 );
				if( need_m != 0u && warp_read( freeq.avail, lane ) > 0 ) q_pop( freeq, __popc( need_m ), false, need_m, id, lane, lt_mask );
				const bool can = lane_idle( l ) && id >= 0;
				const unsigned can_m = __ballot_sync( kFull, can );
				if( can_m != 0u ) {
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
						const uint32_t rank = __popc( can_m & lt_mask );
						bool loaded = false;
						if( can && rank < avail ) {
							seg = blk_next + rank;
							if( p.perm ) seg = __ldg( p.perm + seg );
							loaded = lane_load_segment( p, seg, l );      /* false: a ray the follow-up kernel traces; the lane stays idle */
						}
						blk_next += min( ( uint32_t ) __popc( can_m ), avail );
						const unsigned loaded_m = __ballot_sync( kFull, loaded );
						if( loaded_m != 0u && lane == 0 ) atomicAdd( in_flight, ( int ) __popc( loaded_m ) );
					}
				}
			}
			if( !more_work ) {
				/* no segment left for this warp: its spare slots may be what another warp is waiting for */
				const bool spare = lane_idle( l ) && id >= 0;
				q_push( freeq, spare, id, lane, lt_mask );
				if( spare ) id = -1;
			}

			in_node = __ballot_sync( kFull, lane_in_node( l ) );
			if( in_node == 0u ) {
				if( __ballot_sync( kFull, lane_done( l ) || lane_in_leaf( l ) ) != 0u ) continue;   /* loaded and already finished, or already at a leaf */
				if( warp_read( leafq.avail, lane ) > 0 ) {
					/* nothing to descend: clip what waits, however few */
					q_push( freeq, id >= 0, id, lane, lt_mask );
					id = -1;
					if( q_pop( leafq, 32u, false, kFull, id, lane, lt_mask ) != 0u ) {
						if( id >= 0 ) POOL_LOAD( id );
						leaf_role = true; drain = true;
					}
					continue;
				}
				if( !more_work && warp_read( in_flight, lane ) == 0 ) break;
				__nanosleep( 200 );
				continue;
			}
		}
		stack.base = stack0 + ( uint32_t ) id * 16u;
		stack.overflow = POOL_SCRATCH( id );
		lane_node_phase( map, l, stack, p.max_steps );
	}
#undef POOL_SCRATCH
#undef POOL_LOAD

	for( int o = 16; o > 0; o >>= 1 ) my_hits += __shfl_xor_sync( kFull, my_hits, o );
	if( lane == 0 && my_hits ) atomicAdd( p.hits, ( unsigned long long ) my_hits );
}

} // namespace bspk

#endif
