/*
 * oracle/restate.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's
 * BSP::trace_seg path).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED -- checked bit-for-bit against the unmodified reference
 * object (oracle/_ref/libbspref.so) and against the KAT-0 / KAT-1 vectors the
 * survey captured from it (tests/test_oracle.py).  The one unpinned assumption is
 * the evaluation order of glm::dot, see oracle/shim/glm/glm.hpp.
 */
#ifndef BSP_ORACLE_RESTATE_H
#define BSP_ORACLE_RESTATE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* raw IBSP lumps, exactly as they sit in the file (reference bsp.h:41-96) */
typedef struct {
	const void * textures;     uint32_t num_textures;      /* 72 B each */
	const void * planes;       uint32_t num_planes;        /* 16 B each */
	const void * nodes;        uint32_t num_nodes;         /* 36 B each */
	const void * leaves;       uint32_t num_leaves;        /* 48 B each */
	const void * leaf_brushes; uint32_t num_leaf_brushes;  /*  4 B each */
	const void * brushes;      uint32_t num_brushes;       /* 12 B each */
	const void * brush_sides;  uint32_t num_brush_sides;   /*  8 B each */
} orc_map;

/* {hit,t} (+pos) are reference outputs; plane/brush/startsolid are the extension
 * fields of SURVEY.md 8a R5x, DEFINED here (the reference leaves plane null). */
typedef struct {
	float t;
	int32_t plane;   /* planes-lump index of the side that last raised tfar in the brush that supplied t; -1 if none */
	int32_t brush;   /* brushes-lump index of that brush; -1 on miss */
	uint32_t flags;  /* bit0 hit, bit1 startsolid */
} orc_result;

/* visit counters = the per-trace means of BASELINE.md section 5's bytes formula */
typedef struct {
	uint64_t traces;
	uint64_t nodes;       /* trace_seg_tree calls on an internal node */
	uint64_t leaves;      /* trace_seg_leaf calls */
	uint64_t leaf_brushes;/* leaf-brush loop iterations */
	uint64_t solid;       /* brushes passing the content test (trace_seg_brush calls) */
	uint64_t sides;       /* side-loop iterations executed */
	uint64_t crossings;   /* side iterations that reach the divide */
	uint64_t hits;
	uint64_t max_depth;   /* deepest recursion seen */
} orc_counters;

/* segment i: start = segs[i*stride .. +2], end = segs[i*stride + end_off .. +2] */
void orc_trace( const orc_map * map, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n,
                orc_result * out, float * pos3, orc_counters * counters );

/* BSP::position_to_leaf */
void orc_leaf( const orc_map * map, const float * pts, uint64_t stride, uint64_t n, int32_t * leaf );

#ifdef __cplusplus
}
#endif
#endif
