#ifndef RTM_HOST_BMP_WRITER_H
#define RTM_HOST_BMP_WRITER_H

#include "types.h"

// 32 bpp uncompressed BMP, bottom-up (row 0 of `bitmap` is the bottom row), byte-identical to
// what the reference writes (bmp_writer.cpp:7-56).  Returns false if the file cannot be written.
bool WriteBitmap(const char *filename, uint width, uint height, const uint32 *bitmap);

#endif
