// printf-style logging with the reference's line format (reference trace.cpp:11-26): thread id,
// seconds since start, message.  Goes to stderr here so that tools can keep stdout for results;
// set RTM_QUIET=1 to silence it.
#ifndef RTM_HOST_TRACE_H
#define RTM_HOST_TRACE_H

#include <string>

#include "types.h"

void Trace(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
std::string PrintBytesHumanReadable(uint64 bytes);
double TimerGetTick(); // seconds since the first call

#endif
