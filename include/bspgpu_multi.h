/*
 * bspgpu_multi.h -- multi-GPU entry points of libbspgpu.so (SURVEY.md 8b "multi-GPU variant takes an NCCL
 * communicator or rank/world and does shard+gather internally", 8e; north_star: "the work partitions across the
 * 8 GPUs of one box by sharding segments, with the map replicated on every GPU.  NCCL over NVLink is used only
 * for the small result gather and hit-count reduction").
 *
 * The path shards embarrassingly (every BSP::trace_seg call is independent, reference bsp.cc:255-271): rank r of
 * `world` traces the contiguous range bspgpu_shard_range( n, r, world ), results land at the same indices, the map is
 * replicated.  There is no data-path collective; NCCL (loaded at run time with dlopen("libnccl.so.2"): single-GPU
 * users do not need it installed) carries the hit-count all-reduce and the optional gather of device-resident
 * result shards.  Two shapes:
 *
 *   bspgpu_multi_*  ONE process drives several GPUs (what BSP::trace_segs uses after bsp_init( .., BSP_ALL_GPUS )):
 *                   one handle + one host thread per GPU, communicators from ncclCommInitAll.
 *   bspgpu_dist_*   one process PER GPU (torchrun, MPI ...): the launcher passes rank / world and ships the 128-byte
 *                   NCCL unique id from rank 0 to the others over whatever channel it has; the communicator then
 *                   belongs to the handle.
 *
 * Errors as in bspgpu.h: 0 or a negative BSPGPU_E_* code + bspgpu_last_error().
 */
#ifndef BSPGPU_MULTI_H
#define BSPGPU_MULTI_H

#include "bspgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bspgpu_multi bspgpu_multi_t;

/* ---- one process, several GPUs ---- */
/* devices == NULL / ndev == 0: every visible device.  The map is flattened once and uploaded to each. */
int bspgpu_multi_create( bspgpu_multi_t ** out, const bspgpu_map_desc * lumps, const int * devices, int ndev );
int bspgpu_multi_create_from_file( bspgpu_multi_t ** out, const char * path, const int * devices, int ndev );
void bspgpu_multi_destroy( bspgpu_multi_t * m );
int bspgpu_multi_device_count( bspgpu_multi_t * m );
/* the single-GPU handle of rank r (options, info, leaf queries); owned by `m` */
bspgpu_t * bspgpu_multi_handle( bspgpu_multi_t * m, int rank );
/* n x BSP::trace_seg over HOST arrays, sharded across the GPUs; out / pos as in bspgpu_trace_pos (either may be NULL);
 * flags: BSPGPU_SEGS_PACKED24, BSPGPU_OUT_COMPACT8 / BSPGPU_OUT_COMPACT4.  Synchronous.  The hit counts of the shards are summed with
 * ncclAllReduce (every GPU ends up with the total; bspgpu_multi_hit_count reads it). */
int bspgpu_multi_trace( bspgpu_multi_t * m, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags );
int bspgpu_multi_hit_count( bspgpu_multi_t * m, uint64_t * hits );
/* device-resident variant: every GPU generates ITS shard of segments [first, first + n_total) of a synthetic workload
 * (bspgpu_generate) into buffers the library owns, and traces it there; no host copies at all */
int bspgpu_multi_generate( bspgpu_multi_t * m, const bspgpu_gen_desc * g, uint64_t first, uint64_t n_total );
int bspgpu_multi_trace_resident( bspgpu_multi_t * m );
/* gather the result shards of the last resident trace onto GPU `dst_rank` in segment order (ncclSend / ncclRecv over
 * NVLink); *dev_results then points at n_total 16-byte records on that device, valid until the next generate / destroy */
int bspgpu_multi_gather( bspgpu_multi_t * m, int dst_rank, const bspgpu_result ** dev_results );
/* copy results [first, first+count) of the gathered array (or, before a gather, of the owning shards) to host memory */
int bspgpu_multi_read_results( bspgpu_multi_t * m, uint64_t first, uint64_t count, bspgpu_result * host_out );

/* ---- one process per GPU ---- */
#define BSPGPU_NCCL_ID_BYTES 128
int bspgpu_dist_unique_id( void * id128 );                                       /* rank 0; ship the bytes to the other ranks */
int bspgpu_dist_init( bspgpu_t * h, const void * id128, int rank, int world );   /* collective: every rank calls it */
/* all-reduced hit count of the last trace on every rank's handle (collective, on the handle's stream): the total goes to
 * *hits (host, may be NULL; synchronises) and / or dev_u64 (device, may be NULL; asynchronous) */
int bspgpu_dist_hit_count( bspgpu_t * h, uint64_t * hits, void * dev_u64 );
/* gather `bytes` device bytes from every rank into dev_dst on rank `dst` (world x bytes, rank order = segment order for
 * equal shards); dev_dst is ignored on the other ranks.  Collective, asynchronous on the handle's stream. */
int bspgpu_dist_gather( bspgpu_t * h, const void * dev_src, uint64_t bytes, void * dev_dst, int dst );
int bspgpu_dist_shutdown( bspgpu_t * h );

#ifdef __cplusplus
}
#endif
#endif /* BSPGPU_MULTI_H */
