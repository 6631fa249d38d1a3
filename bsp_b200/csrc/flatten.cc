/*
 * flatten.cc -- raw IBSP lumps -> validated, flattened device image (see map_layout.h).
 *
 * Follows what the reference reads on the trace path: node.plane / node.children
 * (bsp.cc:214-215,248-251), leaf.first_brush / num_brushes (bsp.cc:188), leaf_brushes[]
 * (bsp.cc:189), brush.texture -> texture.content_flags & 1 (bsp.cc:190-193),
 * brush.first_side / num_sides -> side.plane -> plane (bsp.cc:127-129).
 * The reference performs no bounds checks on any of these; here every index is checked
 * once so the kernels do not have to.
 */
#include "flatten.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>

namespace bspk {

namespace {

/* on-disk Q3 IBSP v46 records (reference bsp.h:41-96) */
#pragma pack( push, 1 )
struct DiskTexture { char name[ 64 ]; int32_t surface_flags; int32_t content_flags; };
struct DiskPlane { float n[ 3 ]; float d; };
struct DiskNode { uint32_t plane; int32_t children[ 2 ]; int32_t mins[ 3 ]; int32_t maxs[ 3 ]; };
struct DiskLeaf { int32_t cluster, area; int32_t mins[ 3 ], maxs[ 3 ]; uint32_t first_face, num_faces, first_brush, num_brushes; };
struct DiskBrush { uint32_t first_side, num_sides; int32_t texture; };
struct DiskSide { uint32_t plane; int32_t texture; };
#pragma pack( pop )

static_assert( sizeof( DiskTexture ) == 72, "texture" );
static_assert( sizeof( DiskPlane ) == 16, "plane" );
static_assert( sizeof( DiskNode ) == 36, "node" );
static_assert( sizeof( DiskLeaf ) == 48, "leaf" );
static_assert( sizeof( DiskBrush ) == 12, "brush" );
static_assert( sizeof( DiskSide ) == 8, "side" );

const int kSolidBit = 1; /* the reference's "magic number", bsp.cc:192-193 */

uint64_t align_up( uint64_t v, uint64_t a ) { return ( v + a - 1 ) / a * a; }

std::string fmt( const char * f, long long a = 0, long long b = 0, long long c = 0 ) {
	char buf[ 256 ];
	snprintf( buf, sizeof( buf ), f, a, b, c );
	return buf;
}

} // namespace

int flatten_map( const bspgpu_map_desc & L, const FlattenOptions & opt, FlatMap & out, std::string & err ) {
	if( L.num_nodes == 0 || !L.nodes ) { err = "map has no nodes (the trace starts at node 0, bsp.cc:263)"; return BSPGPU_E_MAP; }
	if( ( L.num_planes && !L.planes ) || ( L.num_leaves && !L.leaves ) || ( L.num_brushes && !L.brushes ) ||
	    ( L.num_brush_sides && !L.brush_sides ) || ( L.num_leaf_brushes && !L.leaf_brushes ) || ( L.num_textures && !L.textures ) ) {
		err = "null lump pointer with non-zero count"; return BSPGPU_E_INVALID;
	}
	const DiskTexture * textures = ( const DiskTexture * ) L.textures;
	const DiskPlane * planes = ( const DiskPlane * ) L.planes;
	const DiskNode * nodes = ( const DiskNode * ) L.nodes;
	const DiskLeaf * leaves = ( const DiskLeaf * ) L.leaves;
	const uint32_t * leaf_brushes = ( const uint32_t * ) L.leaf_brushes;
	const DiskBrush * brushes = ( const DiskBrush * ) L.brushes;
	const DiskSide * sides = ( const DiskSide * ) L.brush_sides;

	/* ---- brushes: side ranges, side planes, texture index (checked for every brush a leaf can name) */
	std::vector< uint8_t > brush_solid( L.num_brushes, 0 );
	for( uint32_t b = 0; b < L.num_brushes; b++ ) {
		const DiskBrush & br = brushes[ b ];
		if( br.texture < 0 || ( uint32_t ) br.texture >= L.num_textures ) { err = fmt( "brush %lld: texture %lld out of range", b, br.texture ); return BSPGPU_E_MAP; }
		if( ( uint64_t ) br.first_side + br.num_sides > L.num_brush_sides ) { err = fmt( "brush %lld: sides [%lld,+%lld) out of range", b, br.first_side, br.num_sides ); return BSPGPU_E_MAP; }
		if( br.num_sides >= kLastRef ) { err = fmt( "brush %lld: too many sides", b ); return BSPGPU_E_MAP; }
		brush_solid[ b ] = ( textures[ br.texture ].content_flags & kSolidBit ) ? 1 : 0;
	}
	for( uint32_t s = 0; s < L.num_brush_sides; s++ ) {
		if( sides[ s ].plane >= L.num_planes ) { err = fmt( "brushside %lld: plane %lld out of range", s, sides[ s ].plane ); return BSPGPU_E_MAP; }
	}

	/* ---- walk the tree from node 0 breadth-first: validates children, rejects nodes reached twice
	 *      (cycles would hang the reference too), assigns BFS indices, measures depth */
	std::vector< int32_t > bfs_of( L.num_nodes, -1 );
	std::vector< uint32_t > order;      /* BFS position -> file node index */
	std::vector< uint32_t > depth_of;   /* per BFS position */
	order.reserve( L.num_nodes ); depth_of.reserve( L.num_nodes );
	order.push_back( 0 ); depth_of.push_back( 1 ); bfs_of[ 0 ] = 0;
	uint32_t tree_depth = 1;
	for( size_t head = 0; head < order.size(); head++ ) {
		const uint32_t ni = order[ head ];
		const DiskNode & nd = nodes[ ni ];
		if( nd.plane >= L.num_planes ) { err = fmt( "node %lld: plane %lld out of range", ni, nd.plane ); return BSPGPU_E_MAP; }
		for( int k = 0; k < 2; k++ ) {
			const int32_t c = nd.children[ k ];
			if( c >= 0 ) {
				if( ( uint32_t ) c >= L.num_nodes ) { err = fmt( "node %lld: child %lld out of range", ni, c ); return BSPGPU_E_MAP; }
				if( bfs_of[ c ] != -1 ) { err = fmt( "node %lld reached twice (cycle or shared subtree) via node %lld", c, ni ); return BSPGPU_E_MAP; }
				bfs_of[ c ] = ( int32_t ) order.size();
				order.push_back( ( uint32_t ) c );
				depth_of.push_back( depth_of[ head ] + 1 );
				if( depth_of[ head ] + 1 > tree_depth ) tree_depth = depth_of[ head ] + 1;
			}
			else {
				const int64_t leaf = -( ( int64_t ) c + 1 );
				if( leaf >= ( int64_t ) L.num_leaves ) { err = fmt( "node %lld: leaf %lld out of range", ni, leaf ); return BSPGPU_E_MAP; }
				const DiskLeaf & lf = leaves[ leaf ];
				if( ( uint64_t ) lf.first_brush + lf.num_brushes > L.num_leaf_brushes ) { err = fmt( "leaf %lld: leafbrushes [%lld,+%lld) out of range", leaf, lf.first_brush, lf.num_brushes ); return BSPGPU_E_MAP; }
				for( uint32_t i = 0; i < lf.num_brushes; i++ ) {
					if( leaf_brushes[ lf.first_brush + i ] >= L.num_brushes ) { err = fmt( "leaf %lld: brush %lld out of range", leaf, leaf_brushes[ lf.first_brush + i ] ); return BSPGPU_E_MAP; }
				}
			}
		}
	}

	/* ---- sides, re-laid out per solid brush so that every brush starts at an even index and has an even count:
	 *      the kernels fetch and clip TWO sides per step (one 32-byte gather).  Padding sides have a zero normal and
	 *      d = +inf: dot(p, 0) is 0 or NaN, so start_dist and end_dist are -inf or NaN for ANY input -- never > 0 (no
	 *      reject, bsp.cc:135), never >= 0 (not entering: `starts_inside` kept, :139), and a crossing parameter of
	 *      +inf or NaN is out of every range (:144).  (d = 0 is not enough: a finite start with an infinite end gives
	 *      start_dist = 0, end_dist = NaN, i.e. a crossing that "enters".)  A solid brush WITHOUT sides (the reference's
	 *      loop, :127, does not run: never hits, `starts_inside` stays true) becomes one such pair. */
	std::vector< uint32_t > new_first( L.num_brushes, 0 ), new_count( L.num_brushes, 0 );
	std::vector< SideRec > osides;
	std::vector< uint32_t > oside_plane;
	for( uint32_t b = 0; b < L.num_brushes; b++ ) {
		if( !brush_solid[ b ] ) continue;
		new_first[ b ] = ( uint32_t ) osides.size();
		for( uint32_t k = 0; k < brushes[ b ].num_sides; k++ ) {
			const uint32_t pi = sides[ brushes[ b ].first_side + k ].plane;
			SideRec r; r.nx = planes[ pi ].n[ 0 ]; r.ny = planes[ pi ].n[ 1 ]; r.nz = planes[ pi ].n[ 2 ]; r.d = planes[ pi ].d;
			osides.push_back( r ); oside_plane.push_back( pi );
		}
		while( osides.size() == new_first[ b ] || ( osides.size() & 1 ) ) {
			SideRec z; z.nx = z.ny = z.nz = 0.0f; z.d = INFINITY;
			osides.push_back( z ); oside_plane.push_back( 0 );
		}
		new_count[ b ] = ( uint32_t ) osides.size() - new_first[ b ];
		if( osides.size() >= 0x7ffffff0u ) { err = "too many brush sides"; return BSPGPU_E_MAP; }
	}

	/* ---- emit.  A node whose plane normal is bitwise +e_x / +e_y / +e_z keeps only the axis (map_layout.h: NodeRec);
	 *      every other plane goes to the general-plane section. */
	auto axis_of = [&]( const DiskPlane & pl ) -> int {
		uint32_t b[ 3 ];
		memcpy( b, pl.n, 12 );
		const uint32_t one = 0x3f800000u;
		if( b[ 0 ] == one && b[ 1 ] == 0 && b[ 2 ] == 0 ) return 0;
		if( b[ 0 ] == 0 && b[ 1 ] == one && b[ 2 ] == 0 ) return 1;
		if( b[ 0 ] == 0 && b[ 1 ] == 0 && b[ 2 ] == one ) return 2;
		return -1;
	};
	const uint32_t num_nodes = ( uint32_t ) order.size();
	std::vector< NodeRec > onodes( num_nodes );
	std::vector< NodeOrig > oorig( num_nodes );
	std::vector< PlaneRec > ogplanes;
	std::vector< PlaneRec > full_planes( num_nodes );   /* every node's plane, for the start grid below */
	std::vector< BrushRef > orefs;
	for( uint32_t i = 0; i < num_nodes; i++ ) {
		const DiskNode & nd = nodes[ order[ i ] ];
		const DiskPlane & pl = planes[ nd.plane ];
		NodeRec & o = onodes[ i ];
		PlaneRec full; full.nx = pl.n[ 0 ]; full.ny = pl.n[ 1 ]; full.nz = pl.n[ 2 ]; full.d = pl.d;
		full_planes[ i ] = full;
		o.d = pl.d;
		const int ax = opt.general_planes_only ? -1 : axis_of( pl );
		if( ax >= 0 ) o.w = ( uint32_t ) ax;
		else {
			if( ogplanes.size() >= 0xfffffff0u - kGeneralPlane ) { err = "too many general node planes"; return BSPGPU_E_MAP; }
			o.w = kGeneralPlane + ( uint32_t ) ogplanes.size();
			ogplanes.push_back( full );
		}
		for( int k = 0; k < 2; k++ ) {
			const int32_t c = nd.children[ k ];
			oorig[ i ].orig_child[ k ] = c;
			if( c >= 0 ) { o.child[ k ] = bfs_of[ c ]; continue; }
			const DiskLeaf & lf = leaves[ -( ( int64_t ) c + 1 ) ];
			const size_t first_ref = orefs.size();
			for( uint32_t j = 0; j < lf.num_brushes; j++ ) {
				const uint32_t b = leaf_brushes[ lf.first_brush + j ];
				if( !brush_solid[ b ] ) continue;
				BrushRef r;
				r.first_side = new_first[ b ]; r.num_sides = new_count[ b ];   /* even start, even count (see above) */
				r.brush = ( int32_t ) b; r.reserved = 0;
				orefs.push_back( r );
			}
			if( orefs.size() == first_ref ) { o.child[ k ] = kEmptyLeaf; continue; }
			if( orefs.size() >= 0x7ffffff0u ) { err = "too many leaf brush references"; return BSPGPU_E_MAP; }
			orefs.back().num_sides |= kLastRef;
			o.child[ k ] = ~( int32_t ) first_ref;
		}
	}

	MapImageLayout & lay = out.layout;
	const uint32_t num_sides_out = ( uint32_t ) osides.size();   /* even */
	const uint64_t half_bytes = ( uint64_t ) ( num_sides_out / 2 ) * sizeof( SideRec );
	lay.num_nodes = num_nodes; lay.num_gplanes = ( uint32_t ) ogplanes.size(); lay.num_refs = ( uint32_t ) orefs.size(); lay.num_sides = num_sides_out;
	lay.tree_depth = tree_depth; lay.num_leaves = L.num_leaves;
	lay.off_nodes = 0;
	lay.off_gplanes = align_up( lay.off_nodes + ( uint64_t ) num_nodes * sizeof( NodeRec ), kSectionAlign );
	lay.off_refs = align_up( lay.off_gplanes + ( uint64_t ) ogplanes.size() * sizeof( PlaneRec ), kSectionAlign );
	lay.off_sides_even = align_up( lay.off_refs + ( uint64_t ) orefs.size() * sizeof( BrushRef ), kSectionAlign );
	lay.off_sides_odd = align_up( lay.off_sides_even + half_bytes, kSectionAlign );
	lay.staged_bytes = align_up( lay.off_sides_odd + half_bytes, kSectionAlign );
	lay.off_side_pairs = lay.staged_bytes;
	lay.off_side_plane = align_up( lay.off_side_pairs + ( uint64_t ) num_sides_out * sizeof( SideRec ), kSectionAlign );
	lay.off_node_orig = align_up( lay.off_side_plane + ( uint64_t ) num_sides_out * sizeof( uint32_t ), kSectionAlign );
	lay.off_leaf_cluster = align_up( lay.off_node_orig + ( uint64_t ) num_nodes * sizeof( NodeOrig ), kSectionAlign );
	lay.off_grid = align_up( lay.off_leaf_cluster + ( uint64_t ) L.num_leaves * sizeof( int32_t ), kSectionAlign );

	/* ---- start grid (map_layout.h).  World box = the root node's bounds (reference bsp.h:63-64 mins/maxs, unused
	 *      by the reference's trace); no grid when a map does not carry them. */
	std::vector< int32_t > grid_cells;
	memset( &lay.grid, 0, sizeof( lay.grid ) );
	{
		const DiskNode & root = nodes[ 0 ];
		double ext[ 3 ], wlo[ 3 ];
		bool ok = true;
		double scale = 1.0;
		for( int a = 0; a < 3; a++ ) {
			wlo[ a ] = root.mins[ a ]; ext[ a ] = ( double ) root.maxs[ a ] - root.mins[ a ];
			ok = ok && ext[ a ] >= 1.0 && ext[ a ] < 1e7;
			if( fabs( ( double ) root.mins[ a ] ) > scale ) scale = fabs( ( double ) root.mins[ a ] );
			if( fabs( ( double ) root.maxs[ a ] ) > scale ) scale = fabs( ( double ) root.maxs[ a ] );
		}
		ok = ok && opt.grid_cells != 0;
		for( uint32_t i = 0; ok && i < num_nodes; i++ ) {
			const PlaneRec & n = full_planes[ i ];
			ok = ok && std::isfinite( n.nx ) && std::isfinite( n.ny ) && std::isfinite( n.nz ) && std::isfinite( n.d );
			const double nlen = fabs( ( double ) n.nx ) + fabs( ( double ) n.ny ) + fabs( ( double ) n.nz );
			ok = ok && nlen < 4.0;                       /* the margins below assume roughly unit normals */
			if( fabs( ( double ) n.d ) > scale ) scale = fabs( ( double ) n.d );
		}
		if( ok ) {
			/* ~cubic cells, about 2^20 of them (4 MB), at most 256 along an axis */
			double target_cells = 1048576.0;
			if( opt.grid_cells > 0 ) target_cells = ( double ) ( opt.grid_cells > 64000000 ? 64000000 : opt.grid_cells );   /* bspgpu_create_opts */
			double cell = cbrt( ext[ 0 ] * ext[ 1 ] * ext[ 2 ] / target_cells );
			int32_t dim[ 3 ];
			for( int pass = 0; pass < 2; pass++ ) {
				for( int a = 0; a < 3; a++ ) {
					dim[ a ] = ( int32_t ) ceil( ext[ a ] / cell );
					if( dim[ a ] < 1 ) dim[ a ] = 1;
					if( dim[ a ] > 1024 ) dim[ a ] = 1024;
				}
				if( ( double ) dim[ 0 ] * dim[ 1 ] * dim[ 2 ] <= 1.15 * target_cells ) break;
				cell *= 1.3;
			}
			/* margins: fp32 evaluation of dot(p,n)-d is off by < 2^-20*scale, the crossing parameter by a relative
			 * 2^-22; delta leaves a 20x reserve.  pad covers the rounding of the kernel's cell index. */
			const double delta = 0.25 + scale / 32768.0;
			const double pad = 0.25 + scale / 65536.0;
			grid_cells.resize( ( size_t ) dim[ 0 ] * dim[ 1 ] * dim[ 2 ] );
			for( int32_t iz = 0; iz < dim[ 2 ]; iz++ ) for( int32_t iy = 0; iy < dim[ 1 ]; iy++ ) for( int32_t ix = 0; ix < dim[ 0 ]; ix++ ) {
				const int32_t ic[ 3 ] = { ix, iy, iz };
				double blo[ 3 ], bhi[ 3 ];
				for( int a = 0; a < 3; a++ ) {
					const double cs = ext[ a ] / dim[ a ];
					blo[ a ] = wlo[ a ] + ic[ a ] * cs - pad; bhi[ a ] = wlo[ a ] + ( ic[ a ] + 1 ) * cs + pad;
				}
				int32_t ref = 0;
				for( ;; ) {
					const PlaneRec & n = full_planes[ ref ];
					const NodeRec & nrec = onodes[ ref ];
					const double nn[ 3 ] = { n.nx, n.ny, n.nz };
					double dmin = -( double ) n.d, dmax = -( double ) n.d;
					for( int a = 0; a < 3; a++ ) {
						dmin += nn[ a ] >= 0.0 ? nn[ a ] * blo[ a ] : nn[ a ] * bhi[ a ];
						dmax += nn[ a ] >= 0.0 ? nn[ a ] * bhi[ a ] : nn[ a ] * blo[ a ];
					}
					int32_t child;
					if( dmin >= delta ) child = nrec.child[ 0 ];       /* everything in front: children[ dist > 0 ] = children[ 0 ] */
					else if( dmax <= -delta ) child = nrec.child[ 1 ];
					else break;
					ref = child;
					if( child < 0 ) break;                          /* a leaf reference or kEmptyLeaf */
				}
				grid_cells[ ( ( size_t ) iz * dim[ 1 ] + iy ) * dim[ 0 ] + ix ] = ref;
			}
			for( int a = 0; a < 3; a++ ) {
				lay.grid.dim[ a ] = dim[ a ];
				lay.grid.lo[ a ] = ( float ) wlo[ a ];
				lay.grid.hi[ a ] = ( float ) ( wlo[ a ] + ext[ a ] );
				lay.grid.inv[ a ] = ( float ) ( dim[ a ] / ext[ a ] );
			}
		}
	}
	/* ---- child slots (map_layout.h: Slot): pair i + 1 = the two children of breadth-first node i, pair 0 = { root, unused } */
	if( num_nodes >= ( 1u << 29 ) - 2u ) { err = "too many nodes"; return BSPGPU_E_MAP; }
	auto node_slot = [&]( uint32_t i ) -> Slot {
		const NodeRec & n = onodes[ i ];
		Slot s;
		if( n.w < kGeneralPlane ) { memcpy( &s.a, &n.d, 4 ); s.x = ( ( i + 1 ) << 3 ) | n.w; }
		else { s.a = n.w - kGeneralPlane; s.x = ( ( i + 1 ) << 3 ) | kSlotGeneral; }
		return s;
	};
	auto child_slot = [&]( int32_t c ) -> Slot {
		if( c >= 0 ) return node_slot( ( uint32_t ) c );
		Slot s;
		if( c == kEmptyLeaf ) { s.a = 0xffffffffu; s.x = kSlotEmpty; }
		else { s.a = ( uint32_t ) ~c; s.x = kSlotLeaf; }
		return s;
	};
	std::vector< Slot > oslots( 2 * ( ( size_t ) num_nodes + 1 ) );
	oslots[ 0 ] = node_slot( 0 );
	oslots[ 1 ].a = 0xffffffffu; oslots[ 1 ].x = kSlotEmpty;
	for( uint32_t i = 0; i < num_nodes; i++ ) for( int k = 0; k < 2; k++ ) oslots[ 2 * ( ( size_t ) i + 1 ) + k ] = child_slot( onodes[ i ].child[ k ] );
	lay.root = oslots[ 0 ];
	std::vector< Slot > grid_slots( grid_cells.size() );
	for( size_t i = 0; i < grid_cells.size(); i++ ) grid_slots[ i ] = grid_cells[ i ] == 0 ? oslots[ 0 ] : child_slot( grid_cells[ i ] );
	lay.off_slots = align_up( lay.off_grid + ( uint64_t ) grid_cells.size() * sizeof( int32_t ), kSectionAlign );
	lay.off_grid_slots = align_up( lay.off_slots + ( uint64_t ) oslots.size() * sizeof( Slot ), kSectionAlign );
	lay.total_bytes = align_up( lay.off_grid_slots + ( uint64_t ) grid_slots.size() * sizeof( Slot ), kSectionAlign );

	out.image.assign( lay.total_bytes, 0 );
	unsigned char * img = out.image.data();
	memcpy( img + lay.off_nodes, onodes.data(), onodes.size() * sizeof( NodeRec ) );
	if( !ogplanes.empty() ) memcpy( img + lay.off_gplanes, ogplanes.data(), ogplanes.size() * sizeof( PlaneRec ) );
	if( !orefs.empty() ) memcpy( img + lay.off_refs, orefs.data(), orefs.size() * sizeof( BrushRef ) );
	if( !osides.empty() ) {
		memcpy( img + lay.off_side_pairs, osides.data(), osides.size() * sizeof( SideRec ) );
		memcpy( img + lay.off_side_plane, oside_plane.data(), oside_plane.size() * sizeof( uint32_t ) );
		SideRec * even = ( SideRec * ) ( img + lay.off_sides_even ), * odd = ( SideRec * ) ( img + lay.off_sides_odd );
		for( size_t k = 0; k + 1 < osides.size(); k += 2 ) { even[ k / 2 ] = osides[ k ]; odd[ k / 2 ] = osides[ k + 1 ]; }
	}
	memcpy( img + lay.off_node_orig, oorig.data(), oorig.size() * sizeof( NodeOrig ) );
	for( uint32_t i = 0; i < L.num_leaves; i++ ) memcpy( img + lay.off_leaf_cluster + ( uint64_t ) i * 4, &leaves[ i ].cluster, 4 );
	if( !grid_cells.empty() ) memcpy( img + lay.off_grid, grid_cells.data(), grid_cells.size() * sizeof( int32_t ) );
	memcpy( img + lay.off_slots, oslots.data(), oslots.size() * sizeof( Slot ) );
	if( !grid_slots.empty() ) memcpy( img + lay.off_grid_slots, grid_slots.data(), grid_slots.size() * sizeof( Slot ) );
	return BSPGPU_OK;
}

int read_ibsp_file( const char * path, std::vector< unsigned char > & blob, bspgpu_map_desc & lumps, const void ** vis, uint64_t * vis_len, std::string & err ) {
	FILE * f = fopen( path, "rb" );
	if( !f ) { err = std::string( "cannot open " ) + path; return BSPGPU_E_IO; }
	fseek( f, 0, SEEK_END );
	const long len = ftell( f );
	fseek( f, 0, SEEK_SET );
	if( len < 8 + 8 * 17 ) { fclose( f ); err = "file shorter than an IBSP header"; return BSPGPU_E_IO; }
	blob.resize( ( size_t ) len );
	const size_t got = fread( blob.data(), 1, ( size_t ) len, f );
	fclose( f );
	if( got != ( size_t ) len ) { err = "short read"; return BSPGPU_E_IO; }

	/* header: 4-byte magic, u32 version, then {off,len} per lump (reference bsp.h:29-39).  Like the
	 * reference (bsp.cc:58-86) magic and version are not checked; lump slots 1,2,3,4,6,8,9 are used. */
	struct Dirent { uint32_t off, len; };
	const Dirent * dir = ( const Dirent * ) ( blob.data() + 8 );
	struct Want { int slot; size_t elem; const void ** ptr; uint32_t * count; const char * name; };
	const Want wants[] = {
		{ 1, sizeof( DiskTexture ), &lumps.textures, &lumps.num_textures, "textures" },
		{ 2, sizeof( DiskPlane ), &lumps.planes, &lumps.num_planes, "planes" },
		{ 3, sizeof( DiskNode ), &lumps.nodes, &lumps.num_nodes, "nodes" },
		{ 4, sizeof( DiskLeaf ), &lumps.leaves, &lumps.num_leaves, "leaves" },
		{ 6, sizeof( uint32_t ), &lumps.leaf_brushes, &lumps.num_leaf_brushes, "leafbrushes" },
		{ 8, sizeof( DiskBrush ), &lumps.brushes, &lumps.num_brushes, "brushes" },
		{ 9, sizeof( DiskSide ), &lumps.brush_sides, &lumps.num_brush_sides, "brushsides" },
	};
	for( const Want & w : wants ) {
		const Dirent & d = dir[ w.slot ];
		if( d.len % w.elem != 0 ) { err = std::string( w.name ) + ": lump length not a multiple of the record size (bsp.cc:97)"; return BSPGPU_E_MAP; }
		if( ( uint64_t ) d.off + d.len > ( uint64_t ) len ) { err = std::string( w.name ) + ": lump runs past end of file"; return BSPGPU_E_MAP; }
		*w.ptr = blob.data() + d.off;
		*w.count = ( uint32_t ) ( d.len / w.elem );
	}
	if( vis ) {
		const Dirent & d = dir[ 16 ];   /* LUMP_VISIBILITY (bsp.h:24) */
		if( ( uint64_t ) d.off + d.len > ( uint64_t ) len ) { err = "visibility: lump runs past end of file"; return BSPGPU_E_MAP; }
		*vis = blob.data() + d.off;
		*vis_len = d.len;
	}
	return BSPGPU_OK;
}

} // namespace bspk
