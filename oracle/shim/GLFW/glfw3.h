/* oracle shim: GLFW/glfw3.h -- gl.h only names the window type (gl.h:6). */
#ifndef ORACLE_SHIM_GLFW3_H
#define ORACLE_SHIM_GLFW3_H
struct GLFWwindow;
#endif
