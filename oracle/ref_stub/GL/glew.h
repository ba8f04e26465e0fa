/* Test infrastructure only: 4-typedef stand-in for the generated GLEW header so that the
 * reference's hot-path translation units compile without OpenGL.
 * Nothing here is product code. */
#ifndef PFREF_STUB_GLEW_H
#define PFREF_STUB_GLEW_H
typedef float        GLfloat;
typedef unsigned int GLuint;
typedef int          GLint;
typedef unsigned int GLenum;
#endif
