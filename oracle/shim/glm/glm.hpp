/*
 * oracle/shim/glm/glm.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Minimal stand-in for the un-vendored, un-pinned <glm/glm.hpp> that the
 * reference includes (bsp.cc:6, game.h:73, immediate.h:4 ...).  Only what the
 * reference translation units we compile actually name is provided.
 *
 * Arithmetic contract (the only part that matters for parity):
 *   - every vec3 operator is component-wise IEEE-754 binary32;
 *   - dot(vec3,vec3) = (a.x*b.x + a.y*b.y) + a.z*b.z, i.e. upstream glm's
 *     compute_dot<vec<3,...>> ("tmp = a*b; tmp.x + tmp.y + tmp.z").
 * glm is absent from /root/reference and from this image, so this order is an
 * ASSUMPTION that cannot be re-verified offline ("parity unpinned w.r.t. glm",
 * BASELINE.md section 4).  Call sites on the traced path: bsp.cc:55,121,141,
 * 220,261,266.
 */
#ifndef ORACLE_SHIM_GLM_HPP
#define ORACLE_SHIM_GLM_HPP

namespace glm {

struct vec2 { float x, y; vec2() : x( 0 ), y( 0 ) { } vec2( float a, float b ) : x( a ), y( b ) { } };
struct ivec2 { int x, y; ivec2() : x( 0 ), y( 0 ) { } ivec2( int a, int b ) : x( a ), y( b ) { } };

struct vec3 {
	float x, y, z;
	vec3() : x( 0 ), y( 0 ), z( 0 ) { }
	explicit vec3( float s ) : x( s ), y( s ), z( s ) { }
	vec3( float a, float b, float c ) : x( a ), y( b ), z( c ) { }
	template< typename A, typename B, typename C >
	vec3( A a, B b, C c ) : x( float( a ) ), y( float( b ) ), z( float( c ) ) { }
	vec3 & operator+=( const vec3 & o ) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3 & operator-=( const vec3 & o ) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	vec3 operator-() const { return vec3( -x, -y, -z ); }
};

struct vec4 {
	float x, y, z, w;
	vec4() : x( 0 ), y( 0 ), z( 0 ), w( 0 ) { }
	template< typename A, typename B, typename C, typename D >
	vec4( A a, B b, C c, D d ) : x( float( a ) ), y( float( b ) ), z( float( c ) ), w( float( d ) ) { }
	vec4( const vec3 & v, float d ) : x( v.x ), y( v.y ), z( v.z ), w( d ) { }
};

struct mat4 { float m[ 16 ]; mat4() { for( int i = 0; i < 16; i++ ) m[ i ] = ( i % 5 == 0 ) ? 1.0f : 0.0f; } };

inline vec3 operator+( const vec3 & a, const vec3 & b ) { return vec3( a.x + b.x, a.y + b.y, a.z + b.z ); }
inline vec3 operator-( const vec3 & a, const vec3 & b ) { return vec3( a.x - b.x, a.y - b.y, a.z - b.z ); }
inline vec3 operator*( const vec3 & a, const vec3 & b ) { return vec3( a.x * b.x, a.y * b.y, a.z * b.z ); }
inline vec3 operator*( const vec3 & a, float s ) { return vec3( a.x * s, a.y * s, a.z * s ); }
inline vec3 operator*( float s, const vec3 & a ) { return vec3( s * a.x, s * a.y, s * a.z ); }
inline vec3 operator/( const vec3 & a, float s ) { return vec3( a.x / s, a.y / s, a.z / s ); }

inline float dot( const vec3 & a, const vec3 & b ) {
	const vec3 tmp = a * b;
	return ( tmp.x + tmp.y ) + tmp.z;
}

/* Everything below is only named by GL glue we never execute. */
inline float radians( float deg ) { return deg * 0.01745329251994329576923690768489f; }
inline vec3 radians( const vec3 & d ) { return vec3( radians( d.x ), radians( d.y ), radians( d.z ) ); }
inline mat4 perspective( float, float, float, float ) { return mat4(); }
inline mat4 rotate( const mat4 & m, float, const vec3 & ) { return m; }
inline mat4 translate( const mat4 & m, const vec3 & ) { return m; }
inline const float * value_ptr( const mat4 & m ) { return m.m; }

} // namespace glm

#endif
