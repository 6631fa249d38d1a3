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

#endif
