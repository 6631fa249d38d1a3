/*
 * oracle/restate.c -- TEST INFRASTRUCTURE ONLY.  See restate.h for who may use it.
 *
 * A plain-C restatement of the reference's segment trace, reading the raw IBSP
 * lumps, with visit counters and the R5x extension outputs.  It keeps the
 * reference's recursion, operation order and comparisons; no epsilon anywhere
 * (the reference's EPSILON, bsp.cc:47, is never read).  All arithmetic is IEEE
 * binary32, compiled with -ffp-contract=off.
 */
#include "restate.h"

#include <string.h>

typedef struct { float x, y, z; } v3;

/* on-disk layouts, reference bsp.h:41-96 (Q3 IBSP v46) */
typedef struct { char name[ 64 ]; int32_t surface_flags; int32_t content_flags; } disk_texture;
typedef struct { float n[ 3 ]; float d; } disk_plane;
typedef struct { uint32_t plane; int32_t children[ 2 ]; int32_t mins[ 3 ]; int32_t maxs[ 3 ]; } disk_node;
typedef struct { int32_t cluster, area; int32_t mins[ 3 ], maxs[ 3 ]; uint32_t first_face, num_faces, first_brush, num_brushes; } disk_leaf;
typedef struct { uint32_t first_side, num_sides; int32_t texture; } disk_brush;
typedef struct { uint32_t plane; int32_t texture; } disk_side;

typedef char check_tex[ sizeof( disk_texture ) == 72 ? 1 : -1 ];
typedef char check_node[ sizeof( disk_node ) == 36 ? 1 : -1 ];
typedef char check_leaf[ sizeof( disk_leaf ) == 48 ? 1 : -1 ];

typedef struct {
	const disk_texture * textures;
	const disk_plane * planes;
	const disk_node * nodes;
	const disk_leaf * leaves;
	const uint32_t * leaf_brushes;
	const disk_brush * brushes;
	const disk_side * sides;
} lumps;

typedef struct {
	int hit;          /* bis.hit */
	float t;          /* bis.is.t */
	int32_t plane;    /* extension */
	int32_t brush;    /* extension */
	int startsolid;   /* extension */
	orc_counters * c;
	uint64_t depth;
} trace_state;

/* glm::dot for vec3: tmp = a*b; (tmp.x + tmp.y) + tmp.z */
static inline float dot3( v3 a, v3 b ) {
	const float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
	return ( px + py ) + pz;
}

/* reference bsp.cc:54-56 point_plane_distance */
static inline float plane_dist( v3 p, const disk_plane * pl ) {
	const v3 n = { pl->n[ 0 ], pl->n[ 1 ], pl->n[ 2 ] };
	return dot3( p, n ) - pl->d;
}

/* reference bsp.cc:120-183 BSP::trace_seg_brush.  Extension outputs: *tfar_plane is the
 * planes-lump index of the side that last raised tfar (-1 if none did); *inside is set
 * when the side loop ran to completion with starts_inside still true. */
static int clip_brush( const lumps * L, const disk_brush * brush, v3 start, v3 dir, float tmin, float tmax,
                       float * tout, int32_t * tfar_plane, int * inside, orc_counters * c ) {
	const v3 end = { start.x + dir.x * tmax, start.y + dir.y * tmax, start.z + dir.z * tmax }; /* :121 */
	float tfar = tmin;
	float tnear = tmax;
	int hit = 0;
	int starts_inside = 1;
	int32_t far_plane = -1;

	for( uint32_t i = 0; i < brush->num_sides; i++ ) {                 /* :127 */
		const disk_side * side = &L->sides[ i + brush->first_side ];
		const disk_plane * pl = &L->planes[ side->plane ];
		c->sides++;

		const float start_dist = plane_dist( start, pl );              /* :131 */
		const float end_dist = plane_dist( end, pl );                  /* :132 */

		if( start_dist > 0.0f && end_dist > 0.0f ) return 0;           /* :135 */
		if( start_dist <= 0.0f && end_dist <= 0.0f ) continue;         /* :137 */

		if( start_dist >= 0.0f ) starts_inside = 0;                    /* :139 */

		const v3 n = { pl->n[ 0 ], pl->n[ 1 ], pl->n[ 2 ] };
		const float denom = -dot3( n, dir );                           /* :141 */
		const float t = start_dist / denom;                            /* :142 */
		c->crossings++;

		if( t >= tmin && t <= tmax ) {                                 /* :144 */
			if( start_dist >= 0.0f ) {                                 /* :147 */
				if( t >= tfar ) { tfar = t; hit = 1; far_plane = ( int32_t ) side->plane; }
			}
			else if( t <= tnear ) { tnear = t; hit = 1; }              /* :154 */
		}
	}

	if( starts_inside ) *inside = 1;

	if( hit ) {                                                        /* :169 */
		if( !starts_inside ) {
			if( tfar <= tnear ) { *tout = tfar; *tfar_plane = far_plane; return 1; }
		}
		else if( tnear < tfar ) { *tout = tnear; *tfar_plane = -1; return 1; } /* :176, unreachable in practice */
	}
	return 0;
}

/* reference bsp.cc:185-203 BSP::trace_seg_leaf */
static void visit_leaf( const lumps * L, uint32_t leaf_idx, v3 start, v3 dir, float tmin, float tmax, trace_state * st ) {
	const disk_leaf * leaf = &L->leaves[ leaf_idx ];
	st->c->leaves++;

	for( uint32_t i = 0; i < leaf->num_brushes; i++ ) {
		const uint32_t bidx = L->leaf_brushes[ i + leaf->first_brush ];
		const disk_brush * brush = &L->brushes[ bidx ];
		const disk_texture * tex = &L->textures[ brush->texture ];
		st->c->leaf_brushes++;

		if( tex->content_flags & 1 ) {                                 /* :193 */
			float t;
			int32_t pl = -1;
			int inside = 0;
			st->c->solid++;
			const int hit = clip_brush( L, brush, start, dir, tmin, tmax, &t, &pl, &inside, st->c );
			if( inside ) st->startsolid = 1;
			if( hit ) {
				/* extension: the brush that "supplies" t is the first one to hit, then any strictly nearer one */
				if( !st->hit || t < st->t ) { st->plane = pl; st->brush = ( int32_t ) bidx; }
				st->hit = 1;                                           /* :198 */
				if( t < st->t ) st->t = t;                             /* :199 */
			}
		}
	}
}

/* reference bsp.cc:205-253 BSP::trace_seg_tree */
static void visit_tree( const lumps * L, int32_t node_idx, v3 start, v3 dir, float tmin, float tmax, trace_state * st ) {
	if( st->hit ) return;                                              /* :206 */

	if( node_idx < 0 ) {                                               /* :208 */
		visit_leaf( L, ( uint32_t ) -( node_idx + 1 ), start, dir, tmin, tmax, st );
		return;
	}

	const disk_node * node = &L->nodes[ node_idx ];
	const disk_plane * pl = &L->planes[ node->plane ];
	st->c->nodes++;
	st->depth++;
	if( st->depth > st->c->max_depth ) st->c->max_depth = st->depth;

	const v3 n = { pl->n[ 0 ], pl->n[ 1 ], pl->n[ 2 ] };
	const float denom = dot3( n, dir );                                /* :220 */
	const float dist = -plane_dist( start, pl );                       /* :221 */
	int near_child = dist > 0.0f;
	int check_both = 0;
	float t = tmax;

	if( denom != 0.0f ) {                                              /* :226 */
		const float u = dist / denom;
		if( u >= 0 && u <= tmax ) {                                    /* :232 */
			if( u < tmin ) near_child = !near_child;                   /* :235 */
			else { check_both = 1; t = u; }
		}
	}

	visit_tree( L, node->children[ near_child ], start, dir, tmin, t, st );            /* :248 */
	if( check_both ) visit_tree( L, node->children[ !near_child ], start, dir, t, tmax, st ); /* :250 */
	st->depth--;
}

void orc_trace( const orc_map * map, const float * segs, uint64_t stride, uint64_t end_off, uint64_t n,
                orc_result * out, float * pos3, orc_counters * counters ) {
	const lumps L = { map->textures, map->planes, map->nodes, map->leaves, map->leaf_brushes, map->brushes, map->brush_sides };
	orc_counters local;
	memset( &local, 0, sizeof( local ) );
	orc_counters * c = counters ? counters : &local;

	for( uint64_t i = 0; i < n; i++ ) {
		const float * s = segs + i * stride;
		const float * e = s + end_off;
		const v3 start = { s[ 0 ], s[ 1 ], s[ 2 ] };
		const v3 end = { e[ 0 ], e[ 1 ], e[ 2 ] };

		/* reference bsp.cc:255-271 BSP::trace_seg */
		trace_state st;
		st.hit = 0; st.t = 1.0f; st.plane = -1; st.brush = -1; st.startsolid = 0; st.c = c; st.depth = 0;
		const v3 dir = { end.x - start.x, end.y - start.y, end.z - start.z };  /* :261 */
		visit_tree( &L, 0, start, dir, 0.0f, 1.0f, &st );

		c->traces++;
		if( st.hit ) c->hits++;
		if( out ) {
			out[ i ].t = st.t;
			out[ i ].plane = st.hit ? st.plane : -1;
			out[ i ].brush = st.hit ? st.brush : -1;
			out[ i ].flags = ( st.hit ? 1u : 0u ) | ( st.startsolid ? 2u : 0u );
		}
		if( pos3 ) {
			v3 p = { 0.0f, 0.0f, 0.0f };                               /* bis = { } :256 */
			if( st.hit ) {                                             /* :265-267 */
				const v3 d2 = { end.x - start.x, end.y - start.y, end.z - start.z };
				p.x = start.x + st.t * d2.x; p.y = start.y + st.t * d2.y; p.z = start.z + st.t * d2.z;
			}
			pos3[ 3 * i + 0 ] = p.x; pos3[ 3 * i + 1 ] = p.y; pos3[ 3 * i + 2 ] = p.z;
		}
	}
}

/* reference bsp.cc:273-286 BSP::position_to_leaf */
void orc_leaf( const orc_map * map, const float * pts, uint64_t stride, uint64_t n, int32_t * leaf ) {
	const disk_node * nodes = map->nodes;
	const disk_plane * planes = map->planes;
	for( uint64_t i = 0; i < n; i++ ) {
		const float * p = pts + i * stride;
		const v3 pos = { p[ 0 ], p[ 1 ], p[ 2 ] };
		int32_t node_idx = 0;
		do {
			const disk_node * node = &nodes[ node_idx ];
			const float dist = plane_dist( pos, &planes[ node->plane ] );
			node_idx = node->children[ dist < 0 ];
		} while( node_idx >= 0 );
		leaf[ i ] = -( node_idx + 1 );
	}
}
