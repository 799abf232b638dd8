#include "framebuffer.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <exception>

#include "../../include/cuda_trace.h"
#include "bmp_writer.h"
#include "trace.h"

Framebuffer::Framebuffer() : m_threads_stop(false), m_rendering(false)
{
    Trace("Initializing framebuffer: %i x %i tiles, one GPU launcher thread", m_tiles_x, m_tiles_y);
}

Framebuffer::~Framebuffer()
{
    // (a derived class has called KillAllWorkerThreads() in its own destructor: the launcher is idle)
    {
        std::lock_guard<std::mutex> lock(m_state_mtx);
        m_quit = true;
    }
    m_state_cv.notify_all();
    if (m_launcher.joinable())
        m_launcher.join();
    if (m_arena_pinned)
        cuda_trace_host_free(m_arena);
    else
        std::free(m_arena);
}

void Framebuffer::Tile::SetPosition(uint x0, uint y0, uint x1, uint y1, uint32 *storage)
{
    m_x0 = x0; m_y0 = y0; m_x1 = x1; m_y1 = y1;
    m_bgra = storage ? storage : m_own;
}

void Framebuffer::Tile::Clear() { std::memset(m_bgra, 0, size_t(GetWidth()) * GetHeight() * sizeof(uint32)); }

void Framebuffer::RenderTiles(Tile * const *tiles, uint count)
{
    for (uint i = 0; i < count && !m_threads_stop; i++)
    {
        RenderTile(*tiles[i]);
        TileFinished(*tiles[i]);
    }
}

// Launcher thread only: the tile's pixels are complete -- let readers at it (framebuffer.cpp:72-77)
void Framebuffer::TileFinished(Tile& tile)
{
    if (tile.m_locked)
    {
        tile.m_locked = false;
        tile.GetMutex().unlock();
    }
}

void Framebuffer::RenderFrame()
{
    // Hold every tile's mutex until its pixels are in: readers (SaveToBMP) use try_lock and see tiles in
    // flight as black, like the reference (framebuffer.cpp:72,203)
    std::vector<Tile *> list;
    for (Tile& t : m_tiles)
    {
        t.GetMutex().lock();
        t.m_locked = true;
        list.push_back(&t);
    }
    m_failed = false;
    m_error.clear();
    {
        // StartRendering() / Resize() return once every tile is held: the thread that asked for the frame never
        // sees a tile of the previous one afterwards (the reference clears them on that thread, framebuffer.cpp:124-134)
        std::lock_guard<std::mutex> lock(m_state_mtx);
        m_tiles_held = true;
    }
    m_state_cv.notify_all();
    try
    {
        if (!m_threads_stop)
            RenderTiles(list.data(), uint(list.size()));
    }
    catch (const std::exception& e)
    {
        m_failed = true;
        m_error = e.what();
    }
    catch (...)
    {
        m_failed = true;
        m_error = "unknown exception in RenderTiles";
    }
    // whatever was not finished (cancelled or failed frame) is black, as after the reference's StartRendering()
    for (Tile *t : list)
        if (t->m_locked)
        {
            t->Clear();
            TileFinished(*t);
        }
    if (m_failed)
        Trace("Rendering failed: %s", m_error.c_str());
    else if (!m_threads_stop)
    {
        m_last_render_seconds = TimerGetTick() - m_render_start_time;
        Trace("Finished rendering after %.4fs", m_last_render_seconds);
    }
}

void Framebuffer::LauncherThread()
{
    std::unique_lock<std::mutex> lock(m_state_mtx);
    for (;;)
    {
        m_state_cv.wait(lock, [this] { return m_frame_requested || m_quit; });
        if (m_quit)
            return;
        m_frame_requested = false;
        lock.unlock();
        RenderFrame();
        lock.lock();
        m_rendering = false;
        m_state_cv.notify_all();
    }
}

void Framebuffer::CreateWorkerThreads()
{
    m_render_start_time = TimerGetTick();
    {
        std::lock_guard<std::mutex> lock(m_state_mtx);
        m_frame_requested = true;
        m_rendering = true;
        m_tiles_held = false;
    }
    if (!m_launcher.joinable())
        m_launcher = std::thread(&Framebuffer::LauncherThread, this);
    m_state_cv.notify_all();
    std::unique_lock<std::mutex> lock(m_state_mtx);
    m_state_cv.wait(lock, [this] { return m_tiles_held || !m_rendering; });
}

void Framebuffer::KillAllWorkerThreads()
{
    if (!m_rendering)
        return;
    m_threads_stop = true;
    {
        // the request is repeated until the frame ends: the launcher may not have reached the device call yet
        std::unique_lock<std::mutex> lock(m_state_mtx);
        while (m_rendering)
        {
            lock.unlock();
            OnCancel();
            lock.lock();
            m_state_cv.wait_for(lock, std::chrono::milliseconds(1), [this] { return !m_rendering; });
        }
    }
    m_threads_stop = false;
}

bool Framebuffer::WaitRendering()
{
    std::unique_lock<std::mutex> lock(m_state_mtx);
    m_state_cv.wait(lock, [this] { return !m_rendering; });
    return !m_failed;
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
    // one page-locked allocation for all tiles (the device copies straight into it); plain memory if there
    // is no CUDA device to lock it for -- rendering itself then fails loudly in Grid / Renderer
    if (m_arena_pinned)
        cuda_trace_host_free(m_arena);
    else
        std::free(m_arena);
    const size_t bytes = size_t(width) * height * sizeof(uint32);
    m_arena = static_cast<uint32 *>(cuda_trace_host_alloc(bytes));
    m_arena_pinned = m_arena != nullptr;
    if (!m_arena)
        m_arena = static_cast<uint32 *>(std::malloc(bytes ? bytes : 1));
    std::memset(m_arena, 0, bytes);
    const uint tw = width / m_tiles_x, th = height / m_tiles_y;
    size_t offset = 0;
    for (uint ty = 0; ty < m_tiles_y; ty++)
        for (uint tx = 0; tx < m_tiles_x; tx++)
        {
            Tile& t = m_tiles[tx + ty * m_tiles_x];
            t.SetPosition(tx * tw, ty * th, tx == m_tiles_x - 1 ? width : (tx + 1) * tw,
                          ty == m_tiles_y - 1 ? height : (ty + 1) * th, m_arena + offset);
            offset += size_t(t.GetWidth()) * t.GetHeight();
        }
    CreateWorkerThreads();
}

// The reference clears every tile here (framebuffer.cpp:124-134).  Same picture without touching 33 MB per
// frame: a tile stays locked -- readers skip it, it shows black -- until its new pixels are in, and tiles a
// cancelled frame never reached are cleared when that frame ends (RenderFrame).
void Framebuffer::StartRendering()
{
    WaitRendering(); // the reference asserts no threads are alive (framebuffer.cpp:18)
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

uint Framebuffer::CountFinishedTiles()
{
    uint n = 0;
    for (Tile& t : m_tiles)
        if (t.GetMutex().try_lock())
        {
            n++;
            t.GetMutex().unlock();
        }
    return n;
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
