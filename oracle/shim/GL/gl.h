/* oracle shim: GL/gl.h -- empty, GL/glew.h carries the typedefs */
