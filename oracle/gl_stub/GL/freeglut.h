/* Stand-in for <GL/freeglut.h> (included by the reference's particlebot_cuda.cu:5). */
#include <GL/gl.h>
