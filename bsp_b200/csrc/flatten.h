/* flatten.h -- host-side validation + flattening of raw IBSP lumps into the device image. */
#ifndef BSPGPU_FLATTEN_H
#define BSPGPU_FLATTEN_H

#include <string>
#include <vector>

#include "../../include/bspgpu.h"
#include "map_layout.h"

namespace bspk {

struct FlatMap {
	MapImageLayout layout;
	std::vector< unsigned char > image;   /* layout.total_bytes */
};

/* build-time choices of the image (bspgpu_create_opts; defaults = what bspgpu_create uses) */
struct FlattenOptions {
	int64_t grid_cells = -1;            /* start-grid size: -1 default (~2^20 cells), 0 no grid, else the target cell count */
	bool general_planes_only = false;   /* test knob: store every node plane as a general plane (no axis shortcut) */
};

/* Validates every index the trace will follow (so kernels are check-free) and builds the
 * image.  Returns BSPGPU_OK or BSPGPU_E_MAP / BSPGPU_E_INVALID with `err` set. */
int flatten_map( const bspgpu_map_desc & lumps, const FlattenOptions & opt, FlatMap & out, std::string & err );

/* Reads an IBSP file (17- or 18-slot header) into `blob` and points `lumps` into it; `vis` / `vis_len` (may be null)
 * receive the visibility lump (slot 16, reference bsp.h:129-134, bsp.cc:105-114). */
int read_ibsp_file( const char * path, std::vector< unsigned char > & blob, bspgpu_map_desc & lumps, const void ** vis, uint64_t * vis_len, std::string & err );

} // namespace bspk

#endif
