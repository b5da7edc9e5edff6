// Headless no-op OpenGL surface so GLChunk.cpp links without a GL context (oracle/_ref build only).
#pragma once
#include <cstddef>
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef unsigned char GLubyte;
typedef int GLsizei; typedef unsigned char GLboolean; typedef ptrdiff_t GLsizeiptr; typedef ptrdiff_t GLintptr;
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_STATIC_DRAW 0x88E4
#define GL_DYNAMIC_DRAW 0x88E8
#define GL_FLOAT 0x1406
#define GL_FALSE 0
#define GL_TRUE 1
static inline void glGenBuffers(GLsizei, GLuint* p) { *p = 0; }
static inline void glGenVertexArrays(GLsizei, GLuint* p) { *p = 0; }
static inline void glBindVertexArray(GLuint) {}
static inline void glDeleteVertexArrays(GLsizei, const GLuint*) {}
static inline void glDeleteBuffers(GLsizei, const GLuint*) {}
static inline void glBindBuffer(GLenum, GLuint) {}
static inline void glBufferData(GLenum, GLsizeiptr, const void*, GLenum) {}
static inline void glBufferSubData(GLenum, GLintptr, GLsizeiptr, const void*) {}
static inline void glVertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}
static inline void glEnableVertexAttribArray(GLuint) {}
