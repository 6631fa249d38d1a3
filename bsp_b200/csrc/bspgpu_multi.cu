/*
 * bspgpu_multi.cu -- the multi-GPU entry points of include/bspgpu_multi.h.
 *
 * The trace shards embarrassingly: contiguous segment ranges per GPU, map replicated, results in
 * place.  NCCL carries only what north_star gives it: the hit-count all-reduce (8 bytes) and the gather of
 * device-resident result shards.  NCCL is bound at run time (dlopen "libnccl.so.2" -- in a process that already
 * loaded one, e.g. torch's, that same library is used), so libbspgpu.so has no link-time dependency on it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <stdint.h>
#include <string.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bspgpu_multi.h"
#include "flatten.h"
#include "bspgpu_internal.h"

namespace {

/* ---- the slice of the NCCL API this file uses, declared here so that building needs no nccl.h (values are ABI-stable) */
typedef struct ncclComm * ncclComm_t;
typedef struct { char internal[ BSPGPU_NCCL_ID_BYTES ]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclUint8 = 1, kNcclUint64 = 5, kNcclSum = 0 };

struct Nccl {
	void * lib = nullptr;
	const char * ( *GetErrorString )( int ) = nullptr;
	int ( *GetUniqueId )( ncclUniqueId * ) = nullptr;
	int ( *CommInitRank )( ncclComm_t *, int, ncclUniqueId, int ) = nullptr;
	int ( *CommInitAll )( ncclComm_t *, int, const int * ) = nullptr;
	int ( *CommDestroy )( ncclComm_t ) = nullptr;
	int ( *AllReduce )( const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
	int ( *Send )( const void *, size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
	int ( *Recv )( void *, size_t, int, int, ncclComm_t, cudaStream_t ) = nullptr;
	int ( *GroupStart )() = nullptr;
	int ( *GroupEnd )() = nullptr;
};

Nccl g_nccl;
std::mutex g_nccl_lock;

template< typename F >
bool bind( void * lib, const char * name, F & fn ) {
	fn = reinterpret_cast< F >( dlsym( lib, name ) );
	return fn != nullptr;
}

int load_nccl() {
	std::lock_guard< std::mutex > guard( g_nccl_lock );
	if( g_nccl.lib ) return BSPGPU_OK;
	void * lib = dlopen( "libnccl.so.2", RTLD_NOW | RTLD_LOCAL );
	if( !lib ) lib = dlopen( "libnccl.so", RTLD_NOW | RTLD_LOCAL );
	if( !lib ) return fail( BSPGPU_E_CUDA, std::string( "NCCL is needed for the multi-GPU entry points and could not be loaded: " ) + dlerror() );
	Nccl n;
	n.lib = lib;
	const bool ok = bind( lib, "ncclGetErrorString", n.GetErrorString ) && bind( lib, "ncclGetUniqueId", n.GetUniqueId ) &&
		bind( lib, "ncclCommInitRank", n.CommInitRank ) && bind( lib, "ncclCommInitAll", n.CommInitAll ) && bind( lib, "ncclCommDestroy", n.CommDestroy ) &&
		bind( lib, "ncclAllReduce", n.AllReduce ) && bind( lib, "ncclSend", n.Send ) && bind( lib, "ncclRecv", n.Recv ) &&
		bind( lib, "ncclGroupStart", n.GroupStart ) && bind( lib, "ncclGroupEnd", n.GroupEnd );
	if( !ok ) { dlclose( lib ); return fail( BSPGPU_E_CUDA, "libnccl.so.2 lacks a symbol this library needs" ); }
	g_nccl = n;
	return BSPGPU_OK;
}

#define NK( call ) do { const int r_ = ( call ); if( r_ != kNcclSuccess ) { \
	return fail( BSPGPU_E_CUDA, std::string( #call ) + ": " + g_nccl.GetErrorString( r_ ) ); } } while( 0 )

} // namespace

struct bspgpu_multi {
	std::mutex lock;
	std::vector< int > devices;
	std::vector< bspgpu * > handles;
	std::vector< ncclComm_t > comms;
	std::vector< unsigned long long * > d_hits;          /* one u64 per rank: the all-reduce operand */
	/* resident shards (bspgpu_multi_generate) */
	std::vector< void * > d_segs, d_out;
	std::vector< uint64_t > shard_first, shard_count;    /* relative to res_first */
	uint64_t res_total = 0;
	void * d_gathered = nullptr;
	int gathered_rank = -1;
	uint64_t last_hits = 0;
	bool hits_valid = false;
};

namespace {

int world( const bspgpu_multi * m ) { return ( int ) m->handles.size(); }

/* sum the per-GPU hit counters of the last trace: ncclAllReduce over the library's communicators */
int allreduce_hits( bspgpu_multi * m ) {
	const int w = world( m );
	for( int r = 0; r < w; r++ ) {
		const int rc = bspgpu_hit_count_device( m->handles[ r ], m->d_hits[ r ] );
		if( rc != BSPGPU_OK ) return rc;
	}
	NK( g_nccl.GroupStart() );
	for( int r = 0; r < w; r++ ) {
		DeviceGuard guard( m->devices[ r ] );
		NK( g_nccl.AllReduce( m->d_hits[ r ], m->d_hits[ r ], 1, kNcclUint64, kNcclSum, m->comms[ r ], m->handles[ r ]->stream() ) );
	}
	NK( g_nccl.GroupEnd() );
	unsigned long long total = 0;
	{
		DeviceGuard guard( m->devices[ 0 ] );
		CK( cudaMemcpyAsync( &total, m->d_hits[ 0 ], sizeof( total ), cudaMemcpyDeviceToHost, m->handles[ 0 ]->stream() ) );
		CK( cudaStreamSynchronize( m->handles[ 0 ]->stream() ) );
	}
	for( int r = 1; r < w; r++ ) {
		DeviceGuard guard( m->devices[ r ] );
		CK( cudaStreamSynchronize( m->handles[ r ]->stream() ) );
	}
	m->last_hits = total;
	m->hits_valid = true;
	return BSPGPU_OK;
}

void free_resident( bspgpu_multi * m ) {
	for( size_t r = 0; r < m->d_segs.size(); r++ ) {
		DeviceGuard guard( m->devices[ r ] );
		if( m->d_segs[ r ] ) cudaFree( m->d_segs[ r ] );
		if( m->d_out[ r ] ) cudaFree( m->d_out[ r ] );
		m->d_segs[ r ] = m->d_out[ r ] = nullptr;
	}
	if( m->d_gathered ) {
		DeviceGuard guard( m->devices[ m->gathered_rank ] );
		cudaFree( m->d_gathered );
		m->d_gathered = nullptr; m->gathered_rank = -1;
	}
	m->res_total = 0;
}

} // namespace

extern "C" {

int bspgpu_multi_create( bspgpu_multi_t ** out, const bspgpu_map_desc * lumps, const int * devices, int ndev ) {
	if( !out || !lumps || ndev < 0 || ( ndev > 0 && !devices ) ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_create: bad argument" );
	*out = nullptr;
	int count = 0;
	if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 ) {
		cudaGetLastError();
		return fail( BSPGPU_E_CUDA, "bspgpu_multi_create: no CUDA device (this library has no CPU fallback)" );
	}
	bspgpu_multi * m = new bspgpu_multi();
	if( ndev == 0 ) for( int d = 0; d < count; d++ ) m->devices.push_back( d );
	else m->devices.assign( devices, devices + ndev );
	const int w = ( int ) m->devices.size();
	int rc = load_nccl();
	for( int r = 0; rc == BSPGPU_OK && r < w; r++ ) {
		bspgpu_t * h = nullptr;
		rc = bspgpu_create( &h, lumps, m->devices[ r ] );
		if( rc == BSPGPU_OK ) m->handles.push_back( h );
	}
	if( rc == BSPGPU_OK ) {
		m->comms.assign( w, nullptr );
		const int nr = g_nccl.CommInitAll( m->comms.data(), w, m->devices.data() );
		if( nr != kNcclSuccess ) { m->comms.clear(); rc = fail( BSPGPU_E_CUDA, std::string( "ncclCommInitAll: " ) + g_nccl.GetErrorString( nr ) ); }
	}
	for( int r = 0; rc == BSPGPU_OK && r < w; r++ ) {
		DeviceGuard guard( m->devices[ r ] );
		unsigned long long * p = nullptr;
		if( cudaMalloc( &p, sizeof( unsigned long long ) ) != cudaSuccess ) { cudaGetLastError(); rc = fail( BSPGPU_E_CUDA, "bspgpu_multi_create: cudaMalloc failed" ); }
		else m->d_hits.push_back( p );
	}
	if( rc != BSPGPU_OK ) {
		const std::string keep = bspgpu_err();
		bspgpu_multi_destroy( m );
		return fail( rc, keep );
	}
	m->d_segs.assign( w, nullptr ); m->d_out.assign( w, nullptr ); m->shard_first.assign( w, 0 ); m->shard_count.assign( w, 0 );
	*out = m;
	return BSPGPU_OK;
}

int bspgpu_multi_create_from_file( bspgpu_multi_t ** out, const char * path, const int * devices, int ndev ) {
	if( !out || !path ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_create_from_file: null argument" );
	std::vector< unsigned char > blob;
	bspgpu_map_desc lumps;
	memset( &lumps, 0, sizeof( lumps ) );
	std::string err;
	const void * vis = nullptr; uint64_t vis_len = 0;
	int rc = bspk::read_ibsp_file( path, blob, lumps, &vis, &vis_len, err );
	if( rc != BSPGPU_OK ) return fail( rc, "bspgpu_multi_create_from_file: " + err );
	rc = bspgpu_multi_create( out, &lumps, devices, ndev );
	if( rc != BSPGPU_OK ) return rc;
	if( vis_len >= 8 ) {
		uint32_t hdr[ 2 ];
		memcpy( hdr, vis, 8 );
		if( ( uint64_t ) hdr[ 0 ] * hdr[ 1 ] + 8 <= vis_len )
			for( bspgpu * h : ( *out )->handles ) if( ( rc = bspgpu_set_vis( h, vis, vis_len ) ) != BSPGPU_OK ) { bspgpu_multi_destroy( *out ); *out = nullptr; return rc; }
	}
	return BSPGPU_OK;
}

void bspgpu_multi_destroy( bspgpu_multi_t * m ) {
	if( !m ) return;
	free_resident( m );
	for( size_t r = 0; r < m->d_hits.size(); r++ ) { DeviceGuard guard( m->devices[ r ] ); cudaFree( m->d_hits[ r ] ); }
	for( size_t r = 0; r < m->comms.size(); r++ ) if( m->comms[ r ] ) g_nccl.CommDestroy( m->comms[ r ] );
	for( bspgpu * h : m->handles ) bspgpu_destroy( h );
	delete m;
}

int bspgpu_multi_device_count( bspgpu_multi_t * m ) { return m ? world( m ) : 0; }

bspgpu_t * bspgpu_multi_handle( bspgpu_multi_t * m, int rank ) {
	if( !m || rank < 0 || rank >= world( m ) ) { fail( BSPGPU_E_INVALID, "bspgpu_multi_handle: rank out of range" ); return nullptr; }
	return m->handles[ rank ];
}

int bspgpu_multi_trace( bspgpu_multi_t * m, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags ) {
	if( !m ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_trace: null handle" );
	if( flags & ( BSPGPU_SEGS_DEVICE | BSPGPU_OUT_DEVICE | BSPGPU_ASYNC ) ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_trace takes host arrays (device-resident shards: bspgpu_multi_generate / _trace_resident)" );
	if( n && ( !segs || ( !out && !pos ) ) ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_trace: null segs, or both out and pos null" );
	std::lock_guard< std::mutex > guard( m->lock );
	const int w = world( m );
	const size_t seg_bytes = ( flags & BSPGPU_SEGS_PACKED24 ) ? 24 : sizeof( bspgpu_segment );
	const size_t out_bytes = ( flags & BSPGPU_OUT_COMPACT4 ) ? sizeof( float ) : ( ( flags & BSPGPU_OUT_COMPACT8 ) ? sizeof( bspgpu_result8 ) : sizeof( bspgpu_result ) );
	/* a batch too small to be worth a thread per GPU runs on rank 0; the other ranks trace an empty shard so that their
	 * hit counters read zero */
	const uint32_t active = n < 16384ull * ( uint64_t ) w ? 1u : ( uint32_t ) w;
	std::vector< int > rcs( w, BSPGPU_OK );
	std::vector< std::string > errs( w );
	auto work = [&]( int r ) {
		uint64_t first = 0, count = 0;
		if( ( uint32_t ) r < active ) bspgpu_shard_range( n, ( uint32_t ) r, active, &first, &count );
		rcs[ r ] = bspgpu_trace_pos( m->handles[ r ], ( const char * ) segs + first * seg_bytes, count,
			out ? ( bspgpu_result * ) ( ( char * ) out + first * out_bytes ) : nullptr, pos ? pos + first : nullptr, flags );
		if( rcs[ r ] != BSPGPU_OK ) errs[ r ] = bspgpu_last_error();
	};
	std::vector< std::thread > threads;
	for( int r = 1; r < w; r++ ) threads.emplace_back( work, r );
	work( 0 );
	for( std::thread & t : threads ) t.join();
	for( int r = 0; r < w; r++ ) if( rcs[ r ] != BSPGPU_OK ) return fail( rcs[ r ], "bspgpu_multi_trace (GPU " + std::to_string( m->devices[ r ] ) + "): " + errs[ r ] );
	return allreduce_hits( m );
}

int bspgpu_multi_hit_count( bspgpu_multi_t * m, uint64_t * hits ) {
	if( !m || !hits ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_hit_count: null argument" );
	std::lock_guard< std::mutex > guard( m->lock );
	if( !m->hits_valid ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_hit_count: no multi-GPU trace has completed on this handle" );
	*hits = m->last_hits;
	return BSPGPU_OK;
}

int bspgpu_multi_generate( bspgpu_multi_t * m, const bspgpu_gen_desc * g, uint64_t first, uint64_t n_total ) {
	if( !m || !g ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_generate: null argument" );
	std::lock_guard< std::mutex > guard( m->lock );
	free_resident( m );
	const int w = world( m );
	for( int r = 0; r < w; r++ ) {
		uint64_t f = 0, c = 0;
		bspgpu_shard_range( n_total, ( uint32_t ) r, ( uint32_t ) w, &f, &c );
		m->shard_first[ r ] = f; m->shard_count[ r ] = c;
		if( c == 0 ) continue;
		DeviceGuard dg( m->devices[ r ] );
		CK( cudaMalloc( &m->d_segs[ r ], c * sizeof( bspgpu_segment ) ) );
		CK( cudaMalloc( &m->d_out[ r ], c * sizeof( bspgpu_result ) ) );
		const int rc = bspgpu_generate( m->handles[ r ], g, first + f, c, m->d_segs[ r ] );
		if( rc != BSPGPU_OK ) return rc;
	}
	m->res_total = n_total;
	return BSPGPU_OK;
}

int bspgpu_multi_trace_resident( bspgpu_multi_t * m ) {
	if( !m ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_trace_resident: null handle" );
	std::lock_guard< std::mutex > guard( m->lock );
	if( !m->res_total ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_trace_resident: no resident segments (bspgpu_multi_generate)" );
	const int w = world( m );
	for( int r = 0; r < w; r++ ) {   /* asynchronous enqueues: all GPUs run at once */
		const int rc = bspgpu_trace( m->handles[ r ], m->d_segs[ r ], m->shard_count[ r ], ( bspgpu_result * ) m->d_out[ r ],
			BSPGPU_SEGS_DEVICE | BSPGPU_OUT_DEVICE | BSPGPU_ASYNC );
		if( rc != BSPGPU_OK && m->shard_count[ r ] ) return rc;
	}
	return allreduce_hits( m );
}

int bspgpu_multi_gather( bspgpu_multi_t * m, int dst_rank, const bspgpu_result ** dev_results ) {
	if( !m || !dev_results ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_gather: null argument" );
	std::lock_guard< std::mutex > guard( m->lock );
	const int w = world( m );
	if( dst_rank < 0 || dst_rank >= w ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_gather: rank out of range" );
	if( !m->res_total ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_gather: no resident results (bspgpu_multi_generate / _trace_resident)" );
	if( m->d_gathered && m->gathered_rank != dst_rank ) { DeviceGuard dg( m->devices[ m->gathered_rank ] ); cudaFree( m->d_gathered ); m->d_gathered = nullptr; }
	if( !m->d_gathered ) {
		DeviceGuard dg( m->devices[ dst_rank ] );
		CK( cudaMalloc( &m->d_gathered, m->res_total * sizeof( bspgpu_result ) ) );
		m->gathered_rank = dst_rank;
	}
	cudaStream_t st_dst = m->handles[ dst_rank ]->stream();
	NK( g_nccl.GroupStart() );
	for( int r = 0; r < w; r++ ) {
		const size_t bytes = m->shard_count[ r ] * sizeof( bspgpu_result );
		if( !bytes ) continue;
		char * dst = ( char * ) m->d_gathered + m->shard_first[ r ] * sizeof( bspgpu_result );
		if( r == dst_rank ) {
			DeviceGuard dg( m->devices[ r ] );
			CK( cudaMemcpyAsync( dst, m->d_out[ r ], bytes, cudaMemcpyDeviceToDevice, st_dst ) );
			continue;
		}
		{ DeviceGuard dg( m->devices[ r ] ); NK( g_nccl.Send( m->d_out[ r ], bytes, kNcclUint8, dst_rank, m->comms[ r ], m->handles[ r ]->stream() ) ); }
		{ DeviceGuard dg( m->devices[ dst_rank ] ); NK( g_nccl.Recv( dst, bytes, kNcclUint8, r, m->comms[ dst_rank ], st_dst ) ); }
	}
	NK( g_nccl.GroupEnd() );
	for( int r = 0; r < w; r++ ) { DeviceGuard dg( m->devices[ r ] ); CK( cudaStreamSynchronize( m->handles[ r ]->stream() ) ); }
	*dev_results = ( const bspgpu_result * ) m->d_gathered;
	return BSPGPU_OK;
}

int bspgpu_multi_read_results( bspgpu_multi_t * m, uint64_t first, uint64_t count, bspgpu_result * host_out ) {
	if( !m || ( count && !host_out ) ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_read_results: null argument" );
	std::lock_guard< std::mutex > guard( m->lock );
	if( first + count > m->res_total ) return fail( BSPGPU_E_INVALID, "bspgpu_multi_read_results: range exceeds the resident results" );
	if( m->d_gathered ) {
		DeviceGuard dg( m->devices[ m->gathered_rank ] );
		CK( cudaMemcpy( host_out, ( const char * ) m->d_gathered + first * sizeof( bspgpu_result ), count * sizeof( bspgpu_result ), cudaMemcpyDeviceToHost ) );
		return BSPGPU_OK;
	}
	for( int r = 0; r < world( m ); r++ ) {   /* not gathered: read each owning shard */
		const uint64_t lo = first > m->shard_first[ r ] ? first : m->shard_first[ r ];
		const uint64_t hi = first + count < m->shard_first[ r ] + m->shard_count[ r ] ? first + count : m->shard_first[ r ] + m->shard_count[ r ];
		if( lo >= hi ) continue;
		DeviceGuard dg( m->devices[ r ] );
		CK( cudaStreamSynchronize( m->handles[ r ]->stream() ) );
		CK( cudaMemcpy( host_out + ( lo - first ), ( const char * ) m->d_out[ r ] + ( lo - m->shard_first[ r ] ) * sizeof( bspgpu_result ),
			( hi - lo ) * sizeof( bspgpu_result ), cudaMemcpyDeviceToHost ) );
	}
	return BSPGPU_OK;
}

/* ---------------------------------------------------------------- one process per GPU */

int bspgpu_dist_unique_id( void * id128 ) {
	if( !id128 ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_unique_id: null argument" );
	const int rc = load_nccl();
	if( rc != BSPGPU_OK ) return rc;
	ncclUniqueId id;
	NK( g_nccl.GetUniqueId( &id ) );
	memcpy( id128, &id, BSPGPU_NCCL_ID_BYTES );
	return BSPGPU_OK;
}

int bspgpu_dist_init( bspgpu_t * h, const void * id128, int rank, int world_size ) {
	if( !h || !id128 || world_size < 1 || rank < 0 || rank >= world_size ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_init: bad argument" );
	const int rc = load_nccl();
	if( rc != BSPGPU_OK ) return rc;
	std::lock_guard< std::mutex > guard( h->lock );
	if( h->dist_comm ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_init: the handle already has a communicator (bspgpu_dist_shutdown first)" );
	DeviceGuard dg( h->device );
	ncclUniqueId id;
	memcpy( &id, id128, BSPGPU_NCCL_ID_BYTES );
	ncclComm_t comm = nullptr;
	NK( g_nccl.CommInitRank( &comm, world_size, id, rank ) );
	if( cudaMalloc( &h->d_dist, sizeof( unsigned long long ) ) != cudaSuccess ) { cudaGetLastError(); g_nccl.CommDestroy( comm ); return fail( BSPGPU_E_CUDA, "bspgpu_dist_init: cudaMalloc failed" ); }
	h->dist_comm = comm; h->dist_rank = rank; h->dist_world = world_size;
	return BSPGPU_OK;
}

int bspgpu_dist_hit_count( bspgpu_t * h, uint64_t * hits, void * dev_u64 ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_hit_count: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( !h->dist_comm ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_hit_count: bspgpu_dist_init has not been called on this handle" );
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	{ const int rc = bspgpu_order_after_user_work( h, st ); if( rc != BSPGPU_OK ) return rc; }
	CK( cudaMemcpyAsync( h->d_dist, h->d_counters, sizeof( unsigned long long ), cudaMemcpyDeviceToDevice, st ) );
	NK( g_nccl.AllReduce( h->d_dist, h->d_dist, 1, kNcclUint64, kNcclSum, ( ncclComm_t ) h->dist_comm, st ) );
	if( dev_u64 ) CK( cudaMemcpyAsync( dev_u64, h->d_dist, sizeof( unsigned long long ), cudaMemcpyDeviceToDevice, st ) );
	if( hits ) {
		unsigned long long v = 0;
		CK( cudaMemcpyAsync( &v, h->d_dist, sizeof( v ), cudaMemcpyDeviceToHost, st ) );
		CK( cudaStreamSynchronize( st ) );
		*hits = v;
	}
	return BSPGPU_OK;
}

int bspgpu_dist_gather( bspgpu_t * h, const void * dev_src, uint64_t bytes, void * dev_dst, int dst ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_gather: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	if( !h->dist_comm ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_gather: bspgpu_dist_init has not been called on this handle" );
	if( dst < 0 || dst >= h->dist_world || ( bytes && ( !dev_src || ( h->dist_rank == dst && !dev_dst ) ) ) ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_gather: bad argument" );
	if( !bytes ) return BSPGPU_OK;
	DeviceGuard dg( h->device );
	cudaStream_t st = h->stream();
	{ const int rc = bspgpu_order_after_user_work( h, st ); if( rc != BSPGPU_OK ) return rc; }
	ncclComm_t comm = ( ncclComm_t ) h->dist_comm;
	NK( g_nccl.GroupStart() );
	if( h->dist_rank == dst ) {
		for( int r = 0; r < h->dist_world; r++ ) {
			char * to = ( char * ) dev_dst + ( uint64_t ) r * bytes;
			if( r == dst ) CK( cudaMemcpyAsync( to, dev_src, bytes, cudaMemcpyDeviceToDevice, st ) );
			else NK( g_nccl.Recv( to, bytes, kNcclUint8, r, comm, st ) );
		}
	}
	else NK( g_nccl.Send( dev_src, bytes, kNcclUint8, dst, comm, st ) );
	NK( g_nccl.GroupEnd() );
	return BSPGPU_OK;
}

int bspgpu_dist_shutdown( bspgpu_t * h ) {
	if( !h ) return fail( BSPGPU_E_INVALID, "bspgpu_dist_shutdown: null handle" );
	std::lock_guard< std::mutex > guard( h->lock );
	bspgpu_dist_release( h );
	return BSPGPU_OK;
}

} // extern "C"

void bspgpu_dist_release( bspgpu * h ) {
	if( h->dist_comm ) {
		DeviceGuard dg( h->device );
		cudaStreamSynchronize( h->stream() );
		g_nccl.CommDestroy( ( ncclComm_t ) h->dist_comm );
		h->dist_comm = nullptr;
	}
	if( h->d_dist ) { DeviceGuard dg( h->device ); cudaFree( h->d_dist ); h->d_dist = nullptr; }
}
