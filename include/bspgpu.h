/*
 * bspgpu.h -- C ABI of the B200 batched BSP segment tracer (libbspgpu.so).
 *
 * This is the drop-in boundary for ONE path of mikejsavage/bsp: the per-call CPU
 * segment trace  BSP::trace_seg -> trace_seg_tree -> trace_seg_leaf -> trace_seg_brush
 * (reference bsp.cc:120-271, declared bsp.h:203-210).  The reference has no FFI of
 * its own (C++ only); each entry point below names the reference interface it
 * replaces or extends.  Plain pointers and sizes only -- no CUDA, torch or C++
 * types cross this boundary.
 *
 * Threading: calls on one handle are serialised (a per-handle mutex + the handle's stream) and,
 * unless BSPGPU_ASYNC is passed, synchronous on return.  Different handles are independent.
 * Every call leaves the calling thread's current CUDA device as it found it.
 * Streams: host-pointer calls run on streams the handle owns.  DEVICE-pointer calls must be ordered after
 * whatever produced their inputs: pass the producing stream to bspgpu_trace_stream (per call), or make it the
 * handle's stream with bspgpu_set_stream; without either the call runs on the handle's own non-blocking
 * stream, which is NOT ordered after the caller's streams -- synchronise first in that case.
 * Errors: every int-returning call gives 0 on success or a negative BSPGPU_E_* code;
 * bspgpu_last_error() then returns a thread-local message.  There is no CPU fallback:
 * without a usable sm_100 device every compute call fails with BSPGPU_E_CUDA.
 */
#ifndef BSPGPU_H
#define BSPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSPGPU_ABI_VERSION 2

enum {
	BSPGPU_OK = 0,
	BSPGPU_E_INVALID = -1,   /* bad argument */
	BSPGPU_E_MAP = -2,       /* map failed validation (index out of range, cycle, too deep ...) */
	BSPGPU_E_CUDA = -3,      /* CUDA runtime error / no device */
	BSPGPU_E_IO = -4,        /* file could not be read */
	BSPGPU_E_NOMEM = -5
};

typedef struct bspgpu bspgpu_t;

/* The seven lumps the trace reads, exactly as bsp_init leaves them in the BSP object
 * (reference bsp.cc:74-81 / bsp.h:160-182): pointers into the file blob + counts.
 * Element layouts are the on-disk Q3 IBSP v46 structs (bsp.h:41-96):
 * texture 72 B, plane 16 B, node 36 B, leaf 48 B, leafbrush 4 B, brush 12 B, brushside 8 B. */
typedef struct bspgpu_map_desc {
	const void * textures;     uint32_t num_textures;
	const void * planes;       uint32_t num_planes;
	const void * nodes;        uint32_t num_nodes;
	const void * leaves;       uint32_t num_leaves;
	const void * leaf_brushes; uint32_t num_leaf_brushes;
	const void * brushes;      uint32_t num_brushes;
	const void * brush_sides;  uint32_t num_brush_sides;
} bspgpu_map_desc;

/* One segment = the (start, end) argument pair of BSP::trace_seg (bsp.h:210), padded to
 * two float4 so a warp loads it with coalesced 16-byte accesses.  The pads are ignored. */
typedef struct bspgpu_segment {
	float start[ 3 ]; float pad0;
	float end[ 3 ];   float pad1;
} bspgpu_segment;

/* One result = the outputs of BSP::trace_seg (bool return + Intersection, bsp.h:143-147).
 *   flags bit0 (BSPGPU_HIT)  == the reference's return value                 [reference parity]
 *   t                        == Intersection::t (1.0f on a miss)             [reference parity]
 *   plane, brush, BSPGPU_STARTSOLID are EXTENSIONS with no reference ground truth (the
 *   reference never assigns Intersection::plane and has no startsolid output); they are
 *   defined by oracle/restate.c:
 *     plane = planes-lump index of the side that last raised tfar in the brush that supplied
 *             t (first brush in leaf order on exact ties), -1 on a miss or when the hit was
 *             reported at the interval start without any side raising tfar;
 *     brush = brushes-lump index of that brush, -1 on a miss;
 *     STARTSOLID = some tested solid brush had the segment start on/behind all of its sides.
 * Intersection::pos is start + t*(end-start) on a hit, (0,0,0) on a miss: written to the
 * optional pos array of bspgpu_trace_pos. */
typedef struct bspgpu_result {
	float t;
	int32_t plane;
	int32_t brush;
	uint32_t flags;
} bspgpu_result;

#define BSPGPU_HIT        1u
#define BSPGPU_STARTSOLID 2u

/* compact result: everything the reference defines (hit, t) plus the extensions that fit in 8 bytes.
 *   bits = flags (bit0 BSPGPU_HIT, bit1 BSPGPU_STARTSOLID) | (plane + 1) << 2      (plane as in bspgpu_result, 0 = none) */
typedef struct bspgpu_result8 {
	float t;
	uint32_t bits;
} bspgpu_result8;
#define BSPGPU_RESULT8_PLANE( bits ) ( ( int32_t ) ( ( bits ) >> 2 ) - 1 )

/* {pos.xyz, t}: Intersection::pos and ::t as one float4 */
typedef struct bspgpu_pos { float pos[ 3 ]; float t; } bspgpu_pos;

/* bspgpu_trace flags */
#define BSPGPU_SEGS_DEVICE   0x01u  /* segs is a device pointer (already resident in HBM), 16-byte aligned (8 if packed) */
#define BSPGPU_OUT_DEVICE    0x02u  /* out / pos are device pointers, 16-byte aligned */
#define BSPGPU_SEGS_PACKED24 0x04u  /* segs are 6 floats each (two glm::vec3, the reference's own argument layout) */
#define BSPGPU_ASYNC         0x08u  /* device pointers only: enqueue on the handle's stream and return */
#define BSPGPU_OUT_COMPACT8  0x10u  /* `out` points to 8-byte bspgpu_result8 records instead of 16-byte bspgpu_result (no `brush`):
                                       halves the device->host bytes of the host-pointer path */
#define BSPGPU_OUT_COMPACT4  0x20u  /* `out` points to one float per segment: Intersection::t on a hit, BSPGPU_T_MISS on a miss -- exactly
                                       what the reference returns (`hit`, `is.t`, bsp.cc:265-270; a hit has 0 <= t <= 1, never 2), in a
                                       quarter of the bytes; `is.pos` is start + t * (end - start) (bsp.cc:266) from the caller's own
                                       segment.  No extensions.  Device pointers must be 4-byte aligned. */
#define BSPGPU_T_MISS 2.0f

/* kernel variant override for bspgpu_set_option(BSPGPU_OPT_VARIANT): default AUTO */
enum {
	BSPGPU_VARIANT_AUTO = 0,
	BSPGPU_VARIANT_SMEM = 1,     /* whole flattened map staged in shared memory (fails if it does not fit) */
	BSPGPU_VARIANT_TOPTREE = 2,  /* experiments build only (-DBSPGPU_EXPERIMENTS): top of the tree in shared memory, rest read-only through L1/L2 */
	BSPGPU_VARIANT_GLOBAL = 3    /* nothing staged */
};
/* Options.  None is needed for correct or fast results: the defaults are the measured best on B200.  Values out of
 * range are rejected with BSPGPU_E_INVALID (never clamped silently into something that could hang a kernel). */
enum {
	BSPGPU_OPT_VARIANT = 1,
	BSPGPU_OPT_BLOCK_THREADS = 2,   /* threads per CTA: 0 (default per variant) or 32..1024 (rounded up to a multiple of 32) */
	BSPGPU_OPT_CTAS_PER_SM = 3,     /* 0 = as many as fit */
	BSPGPU_OPT_TOPTREE_BYTES = 4,   /* shared-memory budget of the TOPTREE variant (experiments build) */
	BSPGPU_OPT_CHUNK_SEGMENTS = 5,  /* host-pointer path: segments per pipelined H2D/kernel/D2H chunk (>= 1024) */
	BSPGPU_OPT_REFILL = 6,          /* persistent-warp refill threshold (idle lanes before a warp fetches work), 1..32 */
	BSPGPU_OPT_MAX_STEPS = 7,       /* node steps a lane takes per node phase; 0 = per-variant default */
	BSPGPU_OPT_BLOCK_SEGS = 8,      /* consecutive segments a warp claims per global atomic: 0 = default (256), else 32..65536
	                                   (rounded up to a multiple of 32) */
	BSPGPU_OPT_SCHEDULER = 9,       /* 0 = phase-scheduled lanes (the shipped kernel).  1 = the same with the node-record walk of the SMEM variant
	                                   on the GLOBAL variant too (experiments build only, for A/B runs).  4 = one thread per segment, no scheduling
	                                   (experiments build only; the naive baseline).  5 = rays wait in CTA-wide shared-memory queues and
	                                   warps specialise on node or leaf work (experiments build only: bit-identical, measured 30 % slower) */
	BSPGPU_OPT_SIDE_STEPS = 10,     /* side steps a lane takes per leaf phase; one step clips two brush sides (default 3) */
	BSPGPU_OPT_LEAF_THRESHOLD = 11, /* lanes waiting in a leaf that make a warp run a leaf phase (1..32) */
	BSPGPU_OPT_MAX_LAUNCH_SEGMENTS = 12, /* segments per kernel launch, default and maximum 2^30 (in-kernel indices are 32-bit) */
	BSPGPU_OPT_START_GRID = 13,     /* segments whose endpoints share a start-grid cell begin below the root (bit-identical, see map_layout.h):
	                                   1 on, 0 off (always node 0), -1 (default) on unless the map is staged in shared memory */
	BSPGPU_OPT_SMEM_STACK = 15,     /* keep the first N (0, 2, 4, 6, 8) traversal-stack entries of every thread in shared memory instead
	                                   of local memory (deeper entries overflow to local memory); -1 (default) = 6 for the GLOBAL
	                                   variant, 0 for the SMEM variant.  Bit-identical; +17 % on long rays (profiles/README.md) */
	BSPGPU_OPT_BIN_SEGMENTS = 16,   /* 1: device-pointer traces first counting-sort the segments by the start-grid cell of their start
	                                   point (a coherence pre-pass) and trace in that order; results still land at each
	                                   segment's own index.  0 (default) off: measured in profiles/README.md */
	BSPGPU_OPT_TINY_CALLS = 18,     /* host-pointer calls of up to 16 segments (a batch of one = BSP::trace_seg) run as ONE kernel that reads the
	                                   segments from and writes the results to the handle's pinned buffer over the link -- no copies, no
	                                   second kernel: 1 (default) on, 0 = the general small-call path (one bounce buffer, one stream) */
	BSPGPU_OPT_PAGEABLE = 17        /* host-pointer calls with pageable (not page-locked) caller memory:
	                                   2 (default) = bounce through the handle's pinned ring: host threads copy chunk c+1 in and
	                                       chunk c-2 out while the GPU works on the chunks in between,
	                                   1 = page-lock the caller's arrays for the duration of a large call (cudaHostRegister),
	                                   0 = hand them to cudaMemcpyAsync as they are (staged by the driver, serialises the pipeline) */
};

/* build-time choices of the flattened image (bspgpu_create_opts); zero-initialise for the defaults */
typedef struct bspgpu_create_options {
	int64_t start_grid_cells;    /* 0 = default (~2^20 cells); -1 = no start grid; else the target cell count (<= 6.4e7) */
	int32_t general_planes_only; /* test knob: 1 = store every node plane with its full normal (no axis records) */
	int32_t reserved;
} bspgpu_create_options;

typedef struct bspgpu_info {
	uint32_t abi_version;
	int32_t device;
	uint32_t sm_count;
	uint32_t variant;            /* variant the last trace ran */
	uint32_t block_threads, grid_ctas;
	uint32_t smem_bytes;         /* dynamic shared memory per CTA of the last trace */
	uint32_t tree_depth;         /* deepest node chain of the flattened tree */
	uint32_t num_nodes, num_brush_refs, num_sides;  /* flattened (reachable, solid-only) counts */
	uint64_t map_bytes;          /* flattened device image size */
	float last_kernel_ms;        /* CUDA-event time of the last trace's kernel(s) on the handle's stream */
	uint32_t last_launches;      /* kernels launched by the last call */
	uint32_t start_grid_cells;   /* 0 when the map carries no world bounds (root node mins/maxs) */
	uint32_t staged_bytes;       /* size of the staged prefix: what the SMEM variant needs in shared memory */
	uint32_t num_general_planes; /* node planes that are not a bitwise +unit normal (the others are stored as their axis) */
	uint32_t num_leaves, num_clusters;   /* leaves-lump count; clusters of the vis lump (0 = none uploaded) */
	uint32_t smem_stack_slots;   /* traversal-stack slots per thread kept in shared memory by the last trace */
	float last_prepass_ms;       /* CUDA-event time of the last trace's segment-binning pre-pass (0 when off) */
} bspgpu_info;

/* ---- lifetime (replaces nothing: the reference keeps the map in host memory; this COPIES and
 *      flattens it into device buffers, the caller's lumps may be freed afterwards) ---- */
int bspgpu_create( bspgpu_t ** out, const bspgpu_map_desc * lumps, int device );
int bspgpu_create_opts( bspgpu_t ** out, const bspgpu_map_desc * lumps, int device, const bspgpu_create_options * opts );
/* convenience: read an IBSP file the way bsp_init does (bsp.cc:58-86; 17- or 18-slot header) */
int bspgpu_create_from_file( bspgpu_t ** out, const char * path, int device );
void bspgpu_destroy( bspgpu_t * h );

/* ---- the hot path: n x BSP::trace_seg (bsp.cc:255-271) ---- */
int bspgpu_trace( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, unsigned flags );
/* same, also writing Intersection::pos (bsp.cc:265-267); out or pos may be NULL (not both) */
int bspgpu_trace_pos( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags );
/* same as bspgpu_trace_pos, enqueued on the caller's cudaStream_t for THIS call only (NULL = the legacy default stream):
 * ordered after the work already in that stream, e.g. the kernel that produced device-resident segments.  The handle's
 * own stream waits for the call before it runs anything else, so hit counts and later calls stay ordered. */
int bspgpu_trace_stream( bspgpu_t * h, const void * segs, uint64_t n, bspgpu_result * out, bspgpu_pos * pos, unsigned flags, void * cuda_stream );
/* number of hits of the last completed trace on this handle (device-side reduction) */
int bspgpu_hit_count( bspgpu_t * h, uint64_t * hits );
/* same counter copied (8 bytes, async on the handle's stream) into caller-owned DEVICE memory, so a
 * multi-GPU launcher can all-reduce it (NCCL) without a host round trip */
int bspgpu_hit_count_device( bspgpu_t * h, void * dev_u64 );

/* ---- batched BSP::position_to_leaf (bsp.cc:273-286): pts = 3 floats each (stride in floats), leaf = leaves-lump index ---- */
int bspgpu_leaf( bspgpu_t * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, unsigned flags );
/* ---- PVS (reference bsp.h:129-134 BSP_Vis, bsp.cc:105-114 load_vis, bsp_renderer.cc:147-168 bspr_render) ----
 * upload the visibility lump {u32 num_clusters, u32 cluster_size, u8 pvs[num_clusters * cluster_size]}; the library copies it.
 * bspgpu_create_from_file does this itself.  A lump shorter than its own header says is rejected (the reference asserts, bsp.cc:111). */
int bspgpu_set_vis( bspgpu_t * h, const void * vis_lump, uint64_t len );
/* per point: leaves-lump index (as bspgpu_leaf), BSP_Leaf::cluster of that leaf, and whether that cluster is potentially
 * visible from cluster `from_cluster` -- the test of bsp_renderer.cc:161-164: bit (c % 8) of pvs[from * cluster_size + c / 8];
 * from_cluster < 0 => everything visible (bsp_renderer.cc:150-156); a point in a leaf without a cluster (c < 0, where the
 * reference's shift is undefined) => not visible.  Any of leaf / cluster / visible may be NULL. */
int bspgpu_leaf_clusters( bspgpu_t * h, const float * pts, uint32_t stride, uint64_t n, int32_t * leaf, int32_t * cluster,
                          int32_t from_cluster, uint8_t * visible, unsigned flags );
/* the loop of bspr_render (bsp_renderer.cc:147-168) as one call: visible[i] for every leaf i of the leaves lump, seen from `pos` */
int bspgpu_visible_leaves( bspgpu_t * h, const float pos[ 3 ], uint8_t * visible /* num_leaves, host */, int32_t * viewer_cluster /* may be NULL */ );

/* ---- synthetic workload generator (tools/seggen.py, bit-identical): writes `count` device segments ---- */
typedef struct bspgpu_gen_desc {
	uint32_t kind;            /* 0 uniform, 1 long (isotropic direction, length range), 2 camera grid */
	uint64_t seed;
	float lo[ 3 ], hi[ 3 ];   /* box for start points (and end points when kind==0) */
	float len_lo, len_hi;     /* kind 1 */
	const float * views;      /* kind 2: host array, 12 floats per view {eye, fwd, right, up} */
	uint32_t num_views, width, height;
	float tan_half, far_dist;
} bspgpu_gen_desc;
int bspgpu_generate( bspgpu_t * h, const bspgpu_gen_desc * g, uint64_t first, uint64_t count, void * dev_segs );

/* ---- plumbing ---- */
int bspgpu_set_option( bspgpu_t * h, int option, int64_t value );
/* run on a caller-owned cudaStream_t from now on (NULL is the legacy default stream, as in CUDA);
 * bspgpu_reset_stream goes back to the handle's own non-blocking stream (the initial state) */
int bspgpu_set_stream( bspgpu_t * h, void * cuda_stream );
int bspgpu_reset_stream( bspgpu_t * h );
int bspgpu_get_info( bspgpu_t * h, bspgpu_info * info );
int bspgpu_sync( bspgpu_t * h );
void * bspgpu_host_alloc( size_t bytes );   /* pinned host memory for the host-pointer path */
void bspgpu_host_free( void * p );
/* contiguous shard of rank r of `world` over n segments: [first, first+count) */
void bspgpu_shard_range( uint64_t n, uint32_t rank, uint32_t world, uint64_t * first, uint64_t * count );
const char * bspgpu_last_error( void );
uint32_t bspgpu_abi_version( void );

#ifdef __cplusplus
}
#endif
#endif /* BSPGPU_H */
