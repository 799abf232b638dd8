// Tiled frame buffer with the reference's public + protected surface (framebuffer.h:16-102):
// Resize / StartRendering / StopRendering / SaveToBMP / Draw, the nested Tile with its
// "minimal interface needed by implementors of RenderTile()", the pure virtual RenderTile(Tile&)
// plugin hook, KillAllWorkerThreads / WorkerThreadsRunning and the m_threads_stop flag.
//
// What changed underneath: the reference starts hardware_concurrency() worker threads that pop
// tiles from a shuffled queue (framebuffer.cpp:16-92).  Here ONE persistent launcher thread hands the
// whole tile list to the virtual RenderTiles(), whose default just loops RenderTile() -- so any
// RenderTile() implementor keeps working -- and which Renderer overrides with a single batched
// GPU launch (cuda_trace_tiles_into).  The launcher holds every tile's mutex when a frame starts and
// releases each one when that tile's pixels have arrived (TileFinished), the point where the
// reference's worker leaves its lock_guard (framebuffer.cpp:72-77): SaveToBMP / Draw see a frame in
// progress, tile by tile.  Tile pixels live in one page-locked allocation so that the device can
// copy straight into them.  OpenGL is gone: Draw() is a no-op, SaveToBMP() is headless.
#ifndef RTM_HOST_FRAMEBUFFER_H
#define RTM_HOST_FRAMEBUFFER_H

#include <array>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "types.h"

class Framebuffer
{
public:
    Framebuffer();
    virtual ~Framebuffer();

    void Resize(uint width, uint height);                 // re-tile and start rendering (async)
    void Draw(uint x, uint y, uint width, uint height);   // no display on a headless GPU box
    void SaveToBMP(const char *filename);
    void StopRendering() { KillAllWorkerThreads(); }
    void StartRendering();                                // clear tiles and render (async)

    // --- additions ---------------------------------------------------------------------------
    bool WaitRendering();                                 // block until the frame is done; false if it failed
    const std::string& GetLastError() const { return m_error; } // what went wrong in the last frame ("" = nothing)
    double GetLastRenderSeconds() const { return m_last_render_seconds; }
    uint GetWidth() const { return m_width; }
    uint GetHeight() const { return m_height; }
    void CopyToBitmap(uint32 *bgra);                      // width*height, row 0 = y 0, waits first
    uint CountFinishedTiles();                            // tiles a reader could lock right now (all of them between frames)

protected:
    uint m_width  = 1;
    uint m_height = 1;

    std::atomic<bool> m_threads_stop; // cooperative cancel, polled by RenderTile implementations

    class Tile
    {
    public:
        Tile() { SetPosition(0, 0, 1, 1); }

        // Minimal interface needed by implementors of RenderTile()
        void GetPosition(uint& x0, uint& y0, uint& x1, uint& y1) const { x0 = m_x0; y0 = m_y0; x1 = m_x1; y1 = m_y1; }
        uint GetWidth()  const { return m_x1 - m_x0; }
        uint GetHeight() const { return m_y1 - m_y0; }
        uint32 * GetBuffer()   { return m_bgra; }

    protected:
        friend class Framebuffer;
        void SetPosition(uint x0, uint y0, uint x1, uint y1, uint32 *storage = nullptr);
        void Clear();
        std::mutex& GetMutex() { return m_mtx; }

    private:
        std::mutex m_mtx;            // held while the tile is being rendered
        uint32 *m_bgra = nullptr;    // row-major within the tile; a slice of the frame buffer's page-locked arena
        uint32  m_own[1] = { 0 };    // (the 1 x 1 tile of a frame buffer that was never resized)
        uint m_x0, m_y0, m_x1, m_y1;
        bool m_locked = false;       // launcher thread only: it holds m_mtx for the frame in flight
    };

    // Override to provide actual rendering functionality (one tile)
    virtual void RenderTile(Tile& tile) = 0;

    // Batch hook: render `count` tiles.  The default calls RenderTile() for each of them.  An implementation
    // calls TileFinished() for every tile whose pixels are complete, as early as it can; tiles it leaves
    // unfinished (cancelled frame) come out black.  It runs on the launcher thread and may throw: the frame
    // then fails (WaitRendering() returns false, GetLastError() says why), the application lives on.
    virtual void RenderTiles(Tile * const *tiles, uint count);
    void TileFinished(Tile& tile);

    // Derived classes call this in their destructor (RenderTile may touch their state)
    void KillAllWorkerThreads();
    bool WorkerThreadsRunning() const { return m_rendering; }

    // Called by KillAllWorkerThreads() after raising m_threads_stop, before joining: lets a
    // derived class interrupt work it has in flight (Renderer: cuda_trace_cancel)
    virtual void OnCancel() { }

    static const uint m_tiles_x = 12; // the reference's fixed layout (framebuffer.h:87-88)
    static const uint m_tiles_y = 9;
    std::array<Tile, m_tiles_x * m_tiles_y> m_tiles;

private:
    std::thread m_launcher;              // persistent; sleeps between frames
    std::mutex m_state_mtx;
    std::condition_variable m_state_cv;
    bool m_frame_requested = false, m_tiles_held = false, m_quit = false;
    std::atomic<bool> m_rendering;       // a frame is requested or in flight
    bool m_failed = false;
    std::string m_error;
    uint32 *m_arena = nullptr;           // page-locked pixels of all tiles, tile after tile
    bool m_arena_pinned = false;
    double m_render_start_time = 0.0;
    double m_last_render_seconds = 0.0;

    void LauncherThread();
    void RenderFrame();
    void CreateWorkerThreads();
};

#endif
