/* Minimal stand-in for <GL/gl.h>: only the types, error enums and glGetError that the
 * reference's include/helper_cuda_gl.h:120-170 mentions.  Lets particlebot_cuda.cu compile
 * headless (no OpenGL on the build or GPU boxes).  TEST INFRASTRUCTURE (oracle/_ref build). */
#ifndef PRS_GL_STUB_H
#define PRS_GL_STUB_H
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
#define GL_NO_ERROR 0
#define GL_INVALID_ENUM 0x0500
#define GL_INVALID_VALUE 0x0501
#define GL_INVALID_OPERATION 0x0502
#define GL_STACK_OVERFLOW 0x0503
#define GL_STACK_UNDERFLOW 0x0504
#define GL_OUT_OF_MEMORY 0x0505
static inline GLenum glGetError(void) { return GL_NO_ERROR; }
#endif
