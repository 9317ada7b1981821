/* Stand-in for <GL/glew.h> (included by the reference's particlebot.cpp:1): buffer objects without OpenGL.
 * TEST INFRASTRUCTURE (oracle/_ref build) — lets the reference's host class compile and run headless: a buffer
 * object is a plain device allocation (oracle/gl_stub/gl_headless.cpp), mapping it for writing hands out a host
 * staging copy, and the CUDA-GL interop calls of particlebot_cuda.cu:69-93 are redirected (oracle/Makefile:
 * -DcudaGraphics...=prs_glstub_...) to functions that return that allocation. */
#ifndef PRS_GLEW_STUB_H
#define PRS_GLEW_STUB_H
#include <stddef.h>
#include <GL/gl.h>
typedef ptrdiff_t GLsizeiptr;
typedef ptrdiff_t GLintptr;
typedef int GLsizei;
typedef unsigned char GLboolean;
typedef void GLvoid;
#define GL_ARRAY_BUFFER 0x8892
#define GL_WRITE_ONLY 0x88B9
#define GL_DYNAMIC_DRAW 0x88E8
#ifdef __cplusplus
extern "C" {
#endif
void glGenBuffers(GLsizei n, GLuint *buffers);
void glDeleteBuffers(GLsizei n, const GLuint *buffers);
void glBindBuffer(GLenum target, GLuint buffer);
void glBufferData(GLenum target, GLsizeiptr size, const void *data, GLenum usage);
void glBufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void *data);
void *glMapBuffer(GLenum target, GLenum access);
GLboolean glUnmapBuffer(GLenum target);
#ifdef __cplusplus
}
#endif
#endif
