/*
 * bspgpu_internal.h -- what the translation units of libbspgpu.so share (bspgpu.cu: single-GPU C ABI; bspgpu_multi.cu:
 * multi-GPU entry points + NCCL).  Not installed, not part of the ABI.
 */
#ifndef BSPGPU_INTERNAL_H
#define BSPGPU_INTERNAL_H

#include <cuda_runtime.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bspgpu.h"
#include "map_layout.h"

namespace bspk { struct TraceParams; }

std::string & bspgpu_err();                               /* the calling thread's last error message */
inline int fail( int code, const std::string & msg ) { bspgpu_err() = msg; return code; }

#define CK( call ) do { cudaError_t e_ = ( call ); if( e_ != cudaSuccess ) { \
	return fail( BSPGPU_E_CUDA, std::string( #call ) + ": " + cudaGetErrorString( e_ ) ); } } while( 0 )

/* every entry point leaves the calling thread's current device as it found it */
struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard( int device ) {
		if( cudaGetDevice( &prev ) != cudaSuccess ) { prev = -1; cudaGetLastError(); }
		if( prev != device ) cudaSetDevice( device );
		else prev = -1;
	}
	~DeviceGuard() { if( prev >= 0 ) cudaSetDevice( prev ); }
};

/* a few host threads that copy between a caller's pageable arrays and the handle's pinned bounce ring (bspgpu.cu: trace_bounced) */
struct CopyPool {
	struct Task { char * dst; const char * src; size_t bytes; };
	std::vector< std::thread > workers;
	std::mutex m;
	std::condition_variable cv_work, cv_done;
	std::deque< Task > q;
	size_t pending = 0;
	bool stop = false;
	explicit CopyPool( int n );
	~CopyPool();
	void copy( void * dst, const void * src, size_t bytes );   /* asynchronous: split into slices for the workers */
	void wait();                                                  /* until every submitted slice has been copied */
};

const int kBufs = 3;
const int kCounterSlots = 64;               /* one work counter per in-flight launch */
const uint64_t kMaxLaunchSegs = 1ull << 30; /* in-kernel segment indices are 32-bit */

typedef void ( *TraceKernel )( const bspk::TraceParams );

struct bspgpu {
	std::mutex lock;              /* calls on one handle are serialised (host state + the handle's stream) */
	int device = 0;
	int sm_count = 0;
	int max_smem_optin = 0;
	cudaStream_t own_stream = nullptr, user_stream = nullptr, copy_in = nullptr, copy_out = nullptr;
	unsigned char * d_image = nullptr;
	bspk::MapImageLayout lay;
	unsigned long long * d_counters = nullptr;   /* [0] hits, then per in-flight launch a pair {work counter, exotic-ray count} */
	uint32_t * d_exotic = nullptr;               /* indices of the rays the follow-up kernel traces (trace_kernels.cuh: kExoticCap entries) */
	uint32_t launch_seq = 0;

	/* visibility lump (bspgpu_set_vis) */
	unsigned char * d_vis = nullptr;
	uint32_t num_clusters = 0, cluster_size = 0;

	/* options */
	int64_t opt_variant = BSPGPU_VARIANT_AUTO, opt_threads = 0, opt_ctas = 0, opt_top_bytes = 16 * 1024;
	int64_t opt_chunk = 2ll << 20, opt_refill = 8, opt_max_steps = 0, opt_block_segs = 256;
	int64_t opt_sched = 0, opt_side_steps = 3, opt_leaf_threshold = 16;
	int64_t opt_smem_stack = -1, opt_start_grid = -1, opt_bin = 0, opt_pageable = 2, opt_tiny = 1;
	uint64_t opt_max_launch = kMaxLaunchSegs;   /* segments per kernel launch (lowered by tests to exercise the multi-launch path) */

	/* launch configuration, recomputed only after an option changed */
	bool cfg_valid = false;
	TraceKernel cfg_kernel = nullptr;
	int cfg_variant = 0, cfg_mode = 0, cfg_threads = 0, cfg_grid = 0, cfg_slots = 0, cfg_pool_rays = 0;
	uint32_t cfg_smem = 0, cfg_staged = 0, cfg_top_nodes = 0;

	/* what the last trace did */
	uint32_t last_variant = 0, last_threads = 0, last_grid = 0, last_smem = 0, last_launches = 0, last_slots = 0;
	cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_pre0 = nullptr, ev_pre1 = nullptr, ev_user = nullptr;
	bool timing_valid = false, prepass_valid = false, user_pending = false;
	cudaStream_t pending_stream = nullptr;       /* the caller's stream that work of this handle may still be running on */

	/* host-pointer path */
	void * d_in[ kBufs ] = { }, * d_out[ kBufs ] = { }, * d_pos[ kBufs ] = { };
	size_t buf_segs = 0;
	cudaEvent_t ev_in[ kBufs ] = { }, ev_k[ kBufs ] = { }, ev_out[ kBufs ] = { };
	unsigned char * h_small = nullptr;           /* pinned bounce buffer of the small-call path: kSmallCall x (32 + 16 + 16) bytes */
	/* pageable callers (BSPGPU_OPT_PAGEABLE = 2): pinned bounce ring, one slot per device staging buffer, filled / drained by host threads */
	void * h_ring_in[ kBufs ] = { }, * h_ring_out[ kBufs ] = { }, * h_ring_pos[ kBufs ] = { };
	size_t ring_segs = 0;
	CopyPool * pool = nullptr;

	/* grow-only device scratch (leaf queries, segment binning) */
	void * d_scratch = nullptr;
	size_t scratch_bytes = 0;
	void * d_pool_ovf = nullptr;                 /* trace_pool: deep traversal-stack entries (grow-only) */
	size_t pool_ovf_bytes = 0;

	bool use_user_stream = false;
	cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }

	/* NCCL communicator of the rank/world entry points (bspgpu_dist_*, bspgpu_multi.cu); null until bspgpu_dist_init */
	void * dist_comm = nullptr;
	int dist_rank = 0, dist_world = 1;
	unsigned long long * d_dist = nullptr;       /* all-reduce staging: one u64 */
};

/* bspgpu.cu: waits (on `st`) for work of this handle that an earlier call left on a caller's stream */
int bspgpu_order_after_user_work( bspgpu * h, cudaStream_t st );
/* bspgpu_multi.cu: releases dist_comm / d_dist (called by bspgpu_destroy) */
void bspgpu_dist_release( bspgpu * h );


#endif
