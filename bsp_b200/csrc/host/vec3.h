/* vec3.h -- the reference's API is written in terms of glm::vec3 (reference bsp.h:48,145,210).
 * glm is a system header there and is not vendored; use it when present, otherwise provide the
 * three-float POD the interface needs under the same name so call sites compile unchanged. */
#ifndef BSP_HOST_VEC3_H
#define BSP_HOST_VEC3_H

#if defined( __has_include )
#if __has_include( <glm/glm.hpp> )
#include <glm/glm.hpp>
#define BSP_HOST_HAVE_GLM 1
#endif
#endif

#ifndef BSP_HOST_HAVE_GLM
namespace glm {
struct vec3 {
	float x, y, z;
	vec3() : x( 0.0f ), y( 0.0f ), z( 0.0f ) { }
	vec3( float x_, float y_, float z_ ) : x( x_ ), y( y_ ), z( z_ ) { }
};
}
#endif

#endif
