// Fixed-width aliases used across the host API (same spellings as the reference's types.h so
// that code written against it compiles unchanged).
#ifndef RTM_HOST_TYPES_H
#define RTM_HOST_TYPES_H

#include <cstdint>

typedef std::int8_t   int8;
typedef std::int16_t  int16;
typedef std::int32_t  int32;
typedef std::int64_t  int64;
typedef unsigned int  uint;
typedef unsigned char uchar;
typedef std::uint8_t  uint8;
typedef std::uint16_t uint16;
typedef std::uint32_t uint32;
typedef std::uint64_t uint64;

static_assert(sizeof(uint) == 4 && sizeof(float) == 4, "ILP32/LP64 with 32-bit int expected");

#endif
