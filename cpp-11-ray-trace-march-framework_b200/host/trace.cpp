#include "trace.h"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <thread>

double TimerGetTick()
{
    static const auto start = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
}

void Trace(const char *fmt, ...)
{
    static const bool quiet = std::getenv("RTM_QUIET") != nullptr;
    const double tick = TimerGetTick();
    if (quiet)
        return;
    std::ostringstream tid;
    tid << std::this_thread::get_id();
    char msg[2048];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    std::fprintf(stderr, "%-14s @ %6.2fs - %s\n", tid.str().c_str(), tick, msg);
}

std::string PrintBytesHumanReadable(uint64 bytes)
{
    static const char *unit[] = { "B", "KB", "MB", "GB", "TB" };
    double v = double(bytes);
    int u = 0;
    while (v >= 1024.0 && u < 4)
    {
        v /= 1024.0;
        u++;
    }
    char buf[64];
    std::snprintf(buf, sizeof(buf), u == 0 ? "%.0f%s" : "%.2f%s", v, unit[u]);
    return buf;
}
