#!/usr/bin/env python
"""Apply INTEGRATION.md's patch to the reference's bsp.h / bsp.cc MECHANICALLY.

    python tools/apply_integration.py <reference dir> <out dir>

Writes patched copies of bsp.h and bsp.cc into <out dir> and symlinks every other reference file next to them (the
reference's headers include "bsp.h" with quotes, which resolves next to the including file: the tree must be whole).
This is the "tools/integration.patch" of the round-1 review in script form: a unified diff would carry the
reference's own lines as context / deletions, and no reference source may be copied into this repository --
so the edits are expressed as anchors (declaration names) + the text to insert or to put in place of a body.

What it does:
  bsp.h   + #include "bspgpu.h", struct Segment, typedef TraceResult (next to Intersection)
          + members `bspgpu_t * gpu;` and `void trace_segs( ... ) const;` in class BSP (all reference members stay)
  bsp.cc  bsp_init      ... + bspgpu_create from the lump pointers, + bspgpu_set_vis
          bsp_destroy   + bspgpu_destroy
          BSP::trace_seg body -> bspgpu_trace_pos on a batch of one
          + BSP::trace_segs
  (trace_seg_brush / _leaf / _tree stay in the file as dead code: the GPU kernels replace them.)
"""
from __future__ import annotations

import os
import re
import sys

H_INCLUDE = '#include "bspgpu.h"\n'
H_TYPES = '''
/* ---- bspgpu integration: batched trace ---- */
struct Segment { glm::vec3 start; glm::vec3 end; };   /* 24 bytes = trace_seg's argument pair */
typedef bspgpu_result TraceResult;                    /* { float t; int32 plane; int32 brush; uint32 flags; } */
'''
H_MEMBERS = '''	bspgpu_t * gpu;   /* opaque device state (libbspgpu.so) */
	void trace_segs( const Segment * segs, size_t n, TraceResult * out ) const;

'''
CC_CREATE = '''
	{
		/* bspgpu integration: flatten + upload the seven trace lumps (the library copies them) and the vis lump */
		bspgpu_map_desc d = {
			bsp->textures, bsp->num_textures, bsp->planes, bsp->num_planes, bsp->nodes, bsp->num_nodes,
			bsp->leaves, bsp->num_leaves, bsp->leaf_brushes, bsp->num_leaf_brushes,
			bsp->brushes, bsp->num_brushes, bsp->brush_sides, bsp->num_brush_sides };
		const int ok = bspgpu_create( &bsp->gpu, &d, -1 );
		if( ok != BSPGPU_OK ) fprintf( stderr, "bsp: %s\\n", bspgpu_last_error() );
		assert( ok == BSPGPU_OK );
		const int vis_ok = bspgpu_set_vis( bsp->gpu, bsp->vis, 2 * sizeof( u32 ) + bsp->vis->num_clusters * bsp->vis->cluster_size );
		assert( vis_ok == BSPGPU_OK );
	}
'''
CC_DESTROY = '	bspgpu_destroy( bsp->gpu );\n'
CC_TRACE_BODY = '''
	/* bspgpu integration: a batch of one through the GPU trace */
	const Segment s = { start, end };
	TraceResult r; bspgpu_pos p;
	const int ok = bspgpu_trace_pos( gpu, &s, 1, &r, &p, BSPGPU_SEGS_PACKED24 );
	assert( ok == BSPGPU_OK );
	is.plane = NULL;   /* the reference never assigns it */
	is.pos = glm::vec3( p.pos[ 0 ], p.pos[ 1 ], p.pos[ 2 ] );
	is.t = r.t;
	return ( r.flags & BSPGPU_HIT ) != 0;
'''
CC_TRACE_SEGS = '''
void BSP::trace_segs( const Segment * segs, size_t n, TraceResult * out ) const {
	const int ok = bspgpu_trace( gpu, segs, n, out, BSPGPU_SEGS_PACKED24 );   /* host arrays, synchronous */
	assert( ok == BSPGPU_OK );
}
'''


def _body_span(text: str, open_brace: int):
    """[open_brace+1, index of the matching close brace)"""
    depth = 0
    for i in range(open_brace, len(text)):
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                return open_brace + 1, i
    raise ValueError("unbalanced braces")


def _function(text: str, pattern: str):
    m = re.search(pattern, text)
    if not m:
        raise ValueError(f"anchor not found: {pattern}")
    brace = text.index("{", m.end() - 1)
    return _body_span(text, brace)


def patch_header(h: str) -> str:
    m = re.search(r'#include "intrinsics.h"\n', h)
    h = h[: m.end()] + H_INCLUDE + h[m.end():]
    m = re.search(r"struct BSP_Intersection \{", h)
    _, close = _body_span(h, h.index("{", m.start()))
    end = h.index(";", close) + 1
    h = h[:end] + "\n" + H_TYPES + h[end:]
    m = re.search(r"\n\ttemplate< typename T > void load_lump", h)
    h = h[: m.start() + 1] + H_MEMBERS + h[m.start() + 1:]
    return h


def patch_source(c: str) -> str:
    # bsp_init: after the last loader call
    lo, hi = _function(c, r"void bsp_init\( BSP \* bsp, std::string filename \) \{")
    c = c[:hi] + CC_CREATE + c[hi:]
    lo, hi = _function(c, r"void bsp_destroy\( BSP \* bsp \) \{")
    c = c[:lo] + "\n" + CC_DESTROY + c[lo:]
    lo, hi = _function(c, r"bool BSP::trace_seg\( const glm::vec3 & start, const glm::vec3 & end, Intersection & is \) const \{")
    c = c[:lo] + CC_TRACE_BODY + c[hi:]
    lo, hi = _function(c, r"bool BSP::trace_seg\( const glm::vec3 & start, const glm::vec3 & end, Intersection & is \) const \{")
    c = c[: hi + 1] + "\n" + CC_TRACE_SEGS + c[hi + 1:]
    return c


def main(ref: str, out: str) -> None:
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(ref, "bsp.h")) as f:
        h = patch_header(f.read())
    with open(os.path.join(ref, "bsp.cc")) as f:
        c = patch_source(f.read())
    with open(os.path.join(out, "bsp.h"), "w") as f:
        f.write(h)
    with open(os.path.join(out, "bsp.cc"), "w") as f:
        f.write(c)
    for name in os.listdir(ref):
        src, dst = os.path.join(ref, name), os.path.join(out, name)
        if name not in ("bsp.h", "bsp.cc") and os.path.isfile(src) and not os.path.lexists(dst):
            os.symlink(src, dst)


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    main(sys.argv[1], sys.argv[2])
