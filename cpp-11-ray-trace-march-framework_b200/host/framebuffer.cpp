#include "framebuffer.h"

#include <cstring>

#include "bmp_writer.h"
#include "trace.h"

Framebuffer::Framebuffer() : m_threads_stop(false)
{
    Trace("Initializing framebuffer: %i x %i tiles, one GPU launcher thread", m_tiles_x, m_tiles_y);
}

void Framebuffer::Tile::SetPosition(uint x0, uint y0, uint x1, uint y1)
{
    m_x0 = x0; m_y0 = y0; m_x1 = x1; m_y1 = y1;
    m_bgra.assign(size_t(GetWidth()) * GetHeight(), 0u);
}

void Framebuffer::Tile::Clear() { std::fill(m_bgra.begin(), m_bgra.end(), 0u); }

void Framebuffer::RenderTiles(Tile * const *tiles, uint count)
{
    for (uint i = 0; i < count && !m_threads_stop; i++)
        RenderTile(*tiles[i]);
}

void Framebuffer::LauncherThread()
{
    // Hold every tile's mutex for the duration of the frame: readers (SaveToBMP) use try_lock
    // and see tiles in flight as black, like the reference (framebuffer.cpp:72,203)
    std::vector<Tile *> list;
    for (Tile& t : m_tiles)
    {
        t.GetMutex().lock();
        list.push_back(&t);
    }
    if (!m_threads_stop)
        RenderTiles(list.data(), uint(list.size()));
    for (Tile *t : list)
        t->GetMutex().unlock();
    if (!m_threads_stop)
    {
        m_last_render_seconds = TimerGetTick() - m_render_start_time;
        Trace("Finished rendering after %.4fs", m_last_render_seconds);
    }
}

void Framebuffer::CreateWorkerThreads()
{
    m_render_start_time = TimerGetTick();
    m_launcher = std::thread(&Framebuffer::LauncherThread, this);
}

void Framebuffer::KillAllWorkerThreads()
{
    if (!m_launcher.joinable())
        return;
    m_threads_stop = true;
    OnCancel();
    m_launcher.join();
    m_threads_stop = false;
}

void Framebuffer::WaitRendering()
{
    if (m_launcher.joinable())
        m_launcher.join();
}

// Tile rectangles: floor(size / count) each, the last column / row takes the remainder
// (reference framebuffer.cpp:94-122).  No-op when the size is unchanged.
void Framebuffer::Resize(uint width, uint height)
{
    if (width == m_width && height == m_height)
        return;
    KillAllWorkerThreads();
    WaitRendering();
    m_width = width;
    m_height = height;
    const uint tw = width / m_tiles_x, th = height / m_tiles_y;
    for (uint ty = 0; ty < m_tiles_y; ty++)
        for (uint tx = 0; tx < m_tiles_x; tx++)
            m_tiles[tx + ty * m_tiles_x].SetPosition(tx * tw, ty * th,
                                                     tx == m_tiles_x - 1 ? width : (tx + 1) * tw,
                                                     ty == m_tiles_y - 1 ? height : (ty + 1) * th);
    CreateWorkerThreads();
}

void Framebuffer::StartRendering()
{
    WaitRendering(); // the reference asserts no threads are alive (framebuffer.cpp:18)
    for (Tile& t : m_tiles)
        t.Clear();
    CreateWorkerThreads();
}

void Framebuffer::Draw(uint, uint, uint, uint) { }

void Framebuffer::CopyToBitmap(uint32 *bgra)
{
    WaitRendering();
    for (Tile& t : m_tiles)
    {
        uint x0, y0, x1, y1;
        t.GetPosition(x0, y0, x1, y1);
        const uint32 *src = t.GetBuffer();
        for (uint y = 0; y < t.GetHeight(); y++)
            std::memcpy(bgra + x0 + size_t(y0 + y) * m_width, src + size_t(y) * t.GetWidth(), size_t(t.GetWidth()) * 4);
    }
}

// Tiles still being rendered are left black (reference framebuffer.cpp:195-221)
void Framebuffer::SaveToBMP(const char *filename)
{
    std::vector<uint32> bitmap(size_t(m_width) * m_height, 0u);
    for (Tile& t : m_tiles)
    {
        if (!t.GetMutex().try_lock())
            continue;
        uint x0, y0, x1, y1;
        t.GetPosition(x0, y0, x1, y1);
        const uint32 *src = t.GetBuffer();
        for (uint y = 0; y < t.GetHeight(); y++)
            std::memcpy(&bitmap[x0 + size_t(y0 + y) * m_width], src + size_t(y) * t.GetWidth(), size_t(t.GetWidth()) * 4);
        t.GetMutex().unlock();
    }
    if (WriteBitmap(filename, m_width, m_height, bitmap.data()))
        Trace("Saved screenshot to '%s'", filename);
    else
        Trace("Could not write screenshot '%s'", filename);
}
