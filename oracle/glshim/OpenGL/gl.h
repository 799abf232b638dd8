/* TEST INFRASTRUCTURE ONLY (oracle build).
 *
 * No-op stand-in for <OpenGL/gl.h> so that the reference's framebuffer.{h,cpp}
 * (which create one GL texture per tile, framebuffer.cpp:223-276) compile and run
 * headless on Linux.  Nothing here draws anything; the reference's own tile
 * buffers (Tile::m_bgra) are read back by oracle/ref_driver.cpp instead.
 */
#ifndef RTM_ORACLE_GL_SHIM_H
#define RTM_ORACLE_GL_SHIM_H

typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int          GLint;
typedef int          GLsizei;
typedef float        GLfloat;

enum
{
    GL_TEXTURE_2D = 1, GL_TEXTURE_MIN_FILTER, GL_TEXTURE_MAG_FILTER, GL_TEXTURE_WRAP_S,
    GL_TEXTURE_WRAP_T, GL_LINEAR, GL_CLAMP_TO_EDGE, GL_RGBA8, GL_BGRA, GL_UNSIGNED_BYTE,
    GL_QUADS
};

static inline void glGenTextures(GLsizei n, GLuint *t) { for (GLsizei i = 0; i < n; i++) t[i] = 0; }
static inline void glDeleteTextures(GLsizei, const GLuint *) { }
static inline void glBindTexture(GLenum, GLuint) { }
static inline void glTexParameteri(GLenum, GLenum, GLint) { }
static inline void glTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum,
                                const void *) { }
static inline void glEnable(GLenum) { }
static inline void glDisable(GLenum) { }
static inline void glColor3f(GLfloat, GLfloat, GLfloat) { }
static inline void glBegin(GLenum) { }
static inline void glEnd() { }
static inline void glTexCoord2f(GLfloat, GLfloat) { }
static inline void glVertex2f(GLfloat, GLfloat) { }

#endif
