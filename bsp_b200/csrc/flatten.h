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

/* Validates every index the trace will follow (so kernels are check-free) and builds the
 * image.  Returns BSPGPU_OK or BSPGPU_E_MAP / BSPGPU_E_INVALID with `err` set. */
int flatten_map( const bspgpu_map_desc & lumps, FlatMap & out, std::string & err );

/* Reads an IBSP file (17- or 18-slot header) into `blob` and points `lumps` into it. */
int read_ibsp_file( const char * path, std::vector< unsigned char > & blob, bspgpu_map_desc & lumps, std::string & err );

} // namespace bspk

#endif
