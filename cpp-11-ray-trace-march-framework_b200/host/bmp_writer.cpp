#include "bmp_writer.h"

#include <cstdio>
#include <cstring>

namespace
{
void put16(unsigned char *p, uint16 v) { p[0] = uchar(v); p[1] = uchar(v >> 8); }
void put32(unsigned char *p, uint32 v) { p[0] = uchar(v); p[1] = uchar(v >> 8); p[2] = uchar(v >> 16); p[3] = uchar(v >> 24); }
}

bool WriteBitmap(const char *filename, uint width, uint height, const uint32 *bitmap)
{
    // BITMAPFILEHEADER (14 bytes) + BITMAPINFOHEADER (40 bytes), little endian
    unsigned char hdr[54];
    std::memset(hdr, 0, sizeof(hdr));
    hdr[0] = 'B';
    hdr[1] = 'M';
    put32(hdr + 2, uint32(sizeof(hdr)) + width * height * 4); // file size
    put32(hdr + 10, uint32(sizeof(hdr)));                     // offset of the pixel data
    put32(hdr + 14, 40);                                      // info header size
    put32(hdr + 18, width);
    put32(hdr + 22, height);                                  // positive: bottom-up
    put16(hdr + 26, 1);                                       // planes
    put16(hdr + 28, 32);                                      // bits per pixel
    std::FILE *f = std::fopen(filename, "wb");
    if (!f)
        return false;
    bool ok = std::fwrite(hdr, sizeof(hdr), 1, f) == 1;
    ok = ok && std::fwrite(bitmap, 4, size_t(width) * height, f) == size_t(width) * height;
    ok = (std::fclose(f) == 0) && ok;
    return ok;
}
