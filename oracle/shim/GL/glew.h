/*
 * oracle/shim/GL/glew.h -- TEST INFRASTRUCTURE ONLY.
 * No-op OpenGL surface so that the reference's bsp.cc (which carries a GL
 * viewer next to the trace code) compiles unmodified.  None of these run on
 * the traced path.
 */
#ifndef ORACLE_SHIM_GLEW_H
#define ORACLE_SHIM_GLEW_H

#include <stddef.h>

typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef char GLchar;
typedef unsigned char GLboolean;
typedef float GLfloat;
typedef unsigned int GLbitfield;
typedef void GLvoid;

typedef void ( *PFNGLGETSHADERIVPROC )( GLuint, GLenum, GLint * );
typedef void ( *PFNGLGETSHADERINFOLOGPROC )( GLuint, GLsizei, GLsizei *, GLchar * );

#define GL_FALSE 0
#define GL_TRUE 1
#define GL_NO_ERROR 0
#define GL_INVALID_ENUM 0x0500
#define GL_INVALID_VALUE 0x0501
#define GL_INVALID_OPERATION 0x0502
#define GL_OUT_OF_MEMORY 0x0505
#define GL_INVALID_FRAMEBUFFER_OPERATION 0x0506
#define GL_DEPTH_BUFFER_BIT 0x00000100
#define GL_COLOR_BUFFER_BIT 0x00004000
#define GL_FRAGMENT_SHADER 0x8B30
#define GL_VERTEX_SHADER 0x8B31
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82
#define GL_INFO_LOG_LENGTH 0x8B84

inline GLenum glGetError() { return GL_NO_ERROR; }
inline void glGetShaderiv( GLuint, GLenum, GLint * p ) { if( p ) *p = GL_TRUE; }
inline void glGetProgramiv( GLuint, GLenum, GLint * p ) { if( p ) *p = GL_TRUE; }
inline void glGetShaderInfoLog( GLuint, GLsizei, GLsizei *, GLchar * ) { }
inline void glGetProgramInfoLog( GLuint, GLsizei, GLsizei *, GLchar * ) { }
inline void glDeleteShader( GLuint ) { }
inline void glDeleteProgram( GLuint ) { }
inline GLuint glCreateShader( GLenum ) { return 0; }
inline GLuint glCreateProgram() { return 0; }
inline void glShaderSource( GLuint, GLsizei, const GLchar * const *, const GLint * ) { }
inline void glCompileShader( GLuint ) { }
inline void glAttachShader( GLuint, GLuint ) { }
inline void glBindFragDataLocation( GLuint, GLuint, const GLchar * ) { }
inline void glLinkProgram( GLuint ) { }
inline GLint glGetAttribLocation( GLuint, const GLchar * ) { return 0; }
inline GLint glGetUniformLocation( GLuint, const GLchar * ) { return 0; }
inline void glClear( GLbitfield ) { }
inline void glUseProgram( GLuint ) { }
inline void glUniformMatrix4fv( GLint, GLsizei, GLboolean, const GLfloat * ) { }

/* the surface of the reference's bsp_renderer.cc (tests/test_integration_patch.py compiles it, unmodified, against the
 * patched bsp.h and links it: it must still build and bspr_init must still walk every lump it reads) */
typedef ptrdiff_t GLsizeiptr;
#define GL_TRIANGLES 0x0004
#define GL_UNSIGNED_INT 0x1405
#define GL_FLOAT 0x1406
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_STATIC_DRAW 0x88E4
inline void glGenVertexArrays( GLsizei n, GLuint * a ) { for( GLsizei i = 0; i < n; i++ ) a[ i ] = ( GLuint ) i + 1; }
inline void glGenBuffers( GLsizei n, GLuint * a ) { for( GLsizei i = 0; i < n; i++ ) a[ i ] = ( GLuint ) i + 1; }
inline void glBindVertexArray( GLuint ) { }
inline void glBindBuffer( GLenum, GLuint ) { }
inline void glBufferData( GLenum, GLsizeiptr, const void *, GLenum ) { }
inline void glEnableVertexAttribArray( GLuint ) { }
inline void glVertexAttribPointer( GLuint, GLint, GLenum, GLboolean, GLsizei, const void * ) { }
inline void glDrawElements( GLenum, GLsizei, GLenum, const void * ) { }
inline void glDeleteBuffers( GLsizei, const GLuint * ) { }
inline void glDeleteVertexArrays( GLsizei, const GLuint * ) { }

#endif
