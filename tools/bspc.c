/*
 * tools/bspc.c -- the surface-area-heuristic kd split search of tools/mapgen.py:compile_bsp, in C.
 *
 * Same algorithm, same double-precision operations in the same order as the numpy version
 * (which stays in mapgen.py as the reference implementation and fallback; tests compare the
 * two byte for byte), ~100x faster: mega200k compiles in seconds instead of minutes.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (tools/mapgen.py does it on first use).
 *
 * Output is the tree in creation order (explicit stack, front child compiled first):
 *   nodes : axis, split position, two child references (>=0 node, <0 = -(leaf+1)), cell bounds
 *   leaves: first/count into leaf_ids, cell bounds
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { int32_t axis; double c; int32_t child[ 2 ]; double lo[ 3 ], hi[ 3 ]; } bspc_node;
typedef struct { int64_t first, count; double lo[ 3 ], hi[ 3 ]; } bspc_leaf;

typedef struct {
	bspc_node * nodes; int64_t num_nodes, cap_nodes;
	bspc_leaf * leaves; int64_t num_leaves, cap_leaves;
	int32_t * leaf_ids; int64_t num_ids, cap_ids;
	int32_t max_depth_seen; int64_t forced_leaves;
} bspc_tree;

typedef struct { int32_t * ids; int64_t n; double lo[ 3 ], hi[ 3 ]; int32_t depth, parent, slot; } work_item;

static int cmp_double( const void * a, const void * b ) {
	const double x = *( const double * ) a, y = *( const double * ) b;
	return ( x > y ) - ( x < y );
}
static int cmp_int( const void * a, const void * b ) {
	const int32_t x = *( const int32_t * ) a, y = *( const int32_t * ) b;
	return ( x > y ) - ( x < y );
}

static void * grow( void * p, int64_t * cap, int64_t need, size_t elem ) {
	if( need <= *cap ) return p;
	int64_t nc = *cap ? *cap * 2 : 1024;
	while( nc < need ) nc *= 2;
	*cap = nc;
	return realloc( p, ( size_t ) nc * elem );
}

static int32_t make_leaf( bspc_tree * t, const int32_t * ids, int64_t n, const double * lo, const double * hi, int32_t depth ) {
	t->leaves = grow( t->leaves, &t->cap_leaves, t->num_leaves + 1, sizeof( bspc_leaf ) );
	t->leaf_ids = grow( t->leaf_ids, &t->cap_ids, t->num_ids + n + 1, sizeof( int32_t ) );
	bspc_leaf * l = &t->leaves[ t->num_leaves ];
	l->first = t->num_ids; l->count = n;
	memcpy( l->lo, lo, sizeof( l->lo ) ); memcpy( l->hi, hi, sizeof( l->hi ) );
	memcpy( t->leaf_ids + t->num_ids, ids, ( size_t ) n * sizeof( int32_t ) );
	qsort( t->leaf_ids + t->num_ids, ( size_t ) n, sizeof( int32_t ), cmp_int );
	t->num_ids += n;
	if( depth > t->max_depth_seen ) t->max_depth_seen = depth;
	t->num_leaves++;
	return ( int32_t ) -t->num_leaves;   /* -(leaf_index + 1) */
}

/* returns the tree; free with bspc_free */
bspc_tree * bspc_build( int64_t nb, const double * bmin, const double * bmax, const double * world_lo, const double * world_hi,
                        int32_t leaf_max, double cost_traverse, double cost_brush, int32_t max_depth, double empty_bonus ) {
	bspc_tree * t = calloc( 1, sizeof( bspc_tree ) );
	int64_t cap_work = 0, num_work = 0;
	work_item * work = NULL;

	/* root: brushes overlapping the world box with positive volume */
	int32_t * root = malloc( ( size_t ) ( nb + 1 ) * sizeof( int32_t ) );
	int64_t nroot = 0;
	for( int64_t i = 0; i < nb; i++ ) {
		int ok = 1;
		for( int a = 0; a < 3; a++ ) ok &= bmax[ 3 * i + a ] > world_lo[ a ] && bmin[ 3 * i + a ] < world_hi[ a ];
		if( ok ) root[ nroot++ ] = ( int32_t ) i;
	}
	work = grow( work, &cap_work, 1, sizeof( work_item ) );
	work[ 0 ].ids = root; work[ 0 ].n = nroot; work[ 0 ].depth = 0; work[ 0 ].parent = -1; work[ 0 ].slot = 0;
	memcpy( work[ 0 ].lo, world_lo, 3 * sizeof( double ) ); memcpy( work[ 0 ].hi, world_hi, 3 * sizeof( double ) );
	num_work = 1;

	int64_t cap_tmp = 0;
	double * mn = NULL, * mx = NULL, * cand = NULL;

	while( num_work ) {
		work_item w = work[ --num_work ];
		const int64_t n = w.n;
		int32_t ref;
		if( n == 0 || w.depth >= max_depth ) {
			if( n ) t->forced_leaves++;
			ref = make_leaf( t, w.ids, n, w.lo, w.hi, w.depth );
		}
		else {
			if( 2 * n + 2 > cap_tmp ) {
				cap_tmp = 2 * n + 2;
				mn = realloc( mn, ( size_t ) cap_tmp * sizeof( double ) );
				mx = realloc( mx, ( size_t ) cap_tmp * sizeof( double ) );
				cand = realloc( cand, ( size_t ) cap_tmp * sizeof( double ) );
			}
			const double ext[ 3 ] = { w.hi[ 0 ] - w.lo[ 0 ], w.hi[ 1 ] - w.lo[ 1 ], w.hi[ 2 ] - w.lo[ 2 ] };
			const double sa = 2.0 * ( ext[ 0 ] * ext[ 1 ] + ext[ 1 ] * ext[ 2 ] + ext[ 0 ] * ext[ 2 ] );
			double best_cost = INFINITY, best_c = 0.0; int best_axis = -1; int64_t best_nb = 0, best_nf = 0;
			for( int a = 0; a < 3; a++ ) {
				for( int64_t i = 0; i < n; i++ ) {
					double lo_v = bmin[ 3 * ( int64_t ) w.ids[ i ] + a ], hi_v = bmax[ 3 * ( int64_t ) w.ids[ i ] + a ];
					if( lo_v < w.lo[ a ] ) lo_v = w.lo[ a ]; if( lo_v > w.hi[ a ] ) lo_v = w.hi[ a ];
					if( hi_v < w.lo[ a ] ) hi_v = w.lo[ a ]; if( hi_v > w.hi[ a ] ) hi_v = w.hi[ a ];
					mn[ i ] = lo_v; mx[ i ] = hi_v;
				}
				qsort( mn, ( size_t ) n, sizeof( double ), cmp_double );
				qsort( mx, ( size_t ) n, sizeof( double ), cmp_double );
				/* sorted unique union of mn and mx, strictly inside the cell */
				int64_t nc = 0, i = 0, j = 0;
				while( i < n || j < n ) {
					double v;
					if( j >= n || ( i < n && mn[ i ] <= mx[ j ] ) ) v = mn[ i++ ]; else v = mx[ j++ ];
					if( v > w.lo[ a ] && v < w.hi[ a ] && ( nc == 0 || cand[ nc - 1 ] != v ) ) cand[ nc++ ] = v;
				}
				if( nc == 0 ) continue;
				const double o1 = ext[ ( a + 1 ) % 3 ], o2 = ext[ ( a + 2 ) % 3 ];
				int64_t ib = 0, jf = 0;   /* ib = #mn < c ; jf = #mx <= c */
				double axis_cost = INFINITY, axis_c = 0.0; int64_t axis_nb = 0, axis_nf = 0;
				for( int64_t k = 0; k < nc; k++ ) {
					const double c = cand[ k ];
					while( ib < n && mn[ ib ] < c ) ib++;
					while( jf < n && mx[ jf ] <= c ) jf++;
					const int64_t n_back = ib, n_front = n - jf;
					const double sa_back = 2.0 * ( o1 * o2 + ( c - w.lo[ a ] ) * ( o1 + o2 ) );
					const double sa_front = 2.0 * ( o1 * o2 + ( w.hi[ a ] - c ) * ( o1 + o2 ) );
					double cost = cost_traverse + cost_brush * ( sa_back * ( double ) n_back + sa_front * ( double ) n_front ) / sa;
					if( n_back == 0 || n_front == 0 ) cost = cost * empty_bonus;
					if( cost < axis_cost ) { axis_cost = cost; axis_c = c; axis_nb = n_back; axis_nf = n_front; }
				}
				if( axis_cost < best_cost ) { best_cost = axis_cost; best_axis = a; best_c = axis_c; best_nb = axis_nb; best_nf = axis_nf; }
			}
			const double leaf_cost = cost_brush * ( double ) n;
			const int progress = best_axis >= 0 && ( best_nb > best_nf ? best_nb : best_nf ) < n;
			if( best_axis < 0 || ( best_cost >= leaf_cost && n <= leaf_max ) || ( n > leaf_max && !progress && best_cost >= leaf_cost ) ) {
				ref = make_leaf( t, w.ids, n, w.lo, w.hi, w.depth );
			}
			else {
				t->nodes = grow( t->nodes, &t->cap_nodes, t->num_nodes + 1, sizeof( bspc_node ) );
				ref = ( int32_t ) t->num_nodes++;
				bspc_node * nd = &t->nodes[ ref ];
				nd->axis = best_axis; nd->c = best_c; nd->child[ 0 ] = nd->child[ 1 ] = 0;
				memcpy( nd->lo, w.lo, sizeof( nd->lo ) ); memcpy( nd->hi, w.hi, sizeof( nd->hi ) );
				int32_t * back = malloc( ( size_t ) ( n + 1 ) * sizeof( int32_t ) ), * front = malloc( ( size_t ) ( n + 1 ) * sizeof( int32_t ) );
				int64_t nbk = 0, nfr = 0;
				for( int64_t i = 0; i < n; i++ ) {
					const int64_t b = w.ids[ i ];
					if( bmin[ 3 * b + best_axis ] < best_c ) back[ nbk++ ] = ( int32_t ) b;
					if( bmax[ 3 * b + best_axis ] > best_c ) front[ nfr++ ] = ( int32_t ) b;
				}
				work = grow( work, &cap_work, num_work + 2, sizeof( work_item ) );
				/* back first so that the front child is compiled next (depth-first, front first) */
				work_item * wb = &work[ num_work++ ];
				wb->ids = back; wb->n = nbk; wb->depth = w.depth + 1; wb->parent = ref; wb->slot = 1;
				memcpy( wb->lo, w.lo, sizeof( wb->lo ) ); memcpy( wb->hi, w.hi, sizeof( wb->hi ) ); wb->hi[ best_axis ] = best_c;
				work_item * wf = &work[ num_work++ ];
				wf->ids = front; wf->n = nfr; wf->depth = w.depth + 1; wf->parent = ref; wf->slot = 0;
				memcpy( wf->lo, w.lo, sizeof( wf->lo ) ); memcpy( wf->hi, w.hi, sizeof( wf->hi ) ); wf->lo[ best_axis ] = best_c;
			}
		}
		if( w.parent >= 0 ) t->nodes[ w.parent ].child[ w.slot ] = ref;
		free( w.ids );
	}
	free( work ); free( mn ); free( mx ); free( cand );
	return t;
}

void bspc_counts( const bspc_tree * t, int64_t * out ) {
	out[ 0 ] = t->num_nodes; out[ 1 ] = t->num_leaves; out[ 2 ] = t->num_ids; out[ 3 ] = t->max_depth_seen; out[ 4 ] = t->forced_leaves;
}

/* copies out: node_i32[n][3] = axis, child0, child1 ; node_f64[n][7] = c, lo[3], hi[3] ;
 * leaf_i64[n][2] = first, count ; leaf_f64[n][6] = lo[3], hi[3] ; ids[num_ids] */
void bspc_export( const bspc_tree * t, int32_t * node_i32, double * node_f64, int64_t * leaf_i64, double * leaf_f64, int32_t * ids ) {
	for( int64_t i = 0; i < t->num_nodes; i++ ) {
		const bspc_node * n = &t->nodes[ i ];
		node_i32[ 3 * i ] = n->axis; node_i32[ 3 * i + 1 ] = n->child[ 0 ]; node_i32[ 3 * i + 2 ] = n->child[ 1 ];
		node_f64[ 7 * i ] = n->c;
		memcpy( node_f64 + 7 * i + 1, n->lo, 3 * sizeof( double ) ); memcpy( node_f64 + 7 * i + 4, n->hi, 3 * sizeof( double ) );
	}
	for( int64_t i = 0; i < t->num_leaves; i++ ) {
		const bspc_leaf * l = &t->leaves[ i ];
		leaf_i64[ 2 * i ] = l->first; leaf_i64[ 2 * i + 1 ] = l->count;
		memcpy( leaf_f64 + 6 * i, l->lo, 3 * sizeof( double ) ); memcpy( leaf_f64 + 6 * i + 3, l->hi, 3 * sizeof( double ) );
	}
	memcpy( ids, t->leaf_ids, ( size_t ) t->num_ids * sizeof( int32_t ) );
}

void bspc_free( bspc_tree * t ) {
	if( !t ) return;
	free( t->nodes ); free( t->leaves ); free( t->leaf_ids ); free( t );
}
