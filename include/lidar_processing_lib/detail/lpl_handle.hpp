// Shared plumbing of the drop-in adaptors: an owning wrapper around lpl_ctx and the translation of
// C-ABI status codes into the exceptions the reference library throws
// (static_unordered_map.hpp:79-82 std::overflow_error, :43-57 std::invalid_argument, others
// std::runtime_error). Header-only: a caller links liblpl_b200.so and nothing else.
#ifndef LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP
#define LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../lpl_b200.h"

namespace lidar_processing_lib
{
namespace detail
{
[[noreturn]] inline void raise(int code, const lpl_ctx* ctx, const char* what)
{
    const std::string msg = std::string(what) + ": " + (ctx != nullptr ? lpl_last_error(ctx) : "no context");
    switch (code)
    {
    case LPL_ERR_INVALID_ARGUMENT:
        throw std::invalid_argument(msg);
    case LPL_ERR_CAPACITY:
        throw std::overflow_error(msg);
    case LPL_ERR_NO_DEVICE:
        throw std::runtime_error(std::string(what) + ": no CUDA device (lpl_b200 has no CPU fallback)");
    default:
        throw std::runtime_error(msg);
    }
}

inline void check(int code, const lpl_ctx* ctx, const char* what)
{
    if (code != LPL_OK)
    {
        raise(code, ctx, what);
    }
}

// One context = one CUDA stream + device scratch for single frames of up to max_points points (~0.75 KB of device
// memory per point of capacity). The reference node owns one Segmenter, one Clusterer, one Polygonizer (and
// possibly a NoiseRemover) and calls them in turn from one thread (processor.cpp:882-893), so the adaptor objects
// of a thread SHARE one context: the first one creates it, it grows when any of them needs more points, and the
// last one to go destroys it. Every adaptor pushes its own configuration before its call (host-side, cheap), so
// sharing is invisible to the caller. Like the reference objects, the adaptors are stateful and not thread-safe;
// objects used from different threads get different contexts.
struct SharedContext
{
    lpl_ctx* ctx = nullptr;
    std::uint32_t max_points = 0;
    std::int32_t height = 0;
    std::int32_t width = 0;
    int users = 0;
    unsigned long long generation = 0; // bumps whenever the context is (re)created

    ~SharedContext()
    {
        if (ctx != nullptr)
        {
            lpl_destroy(ctx);
        }
    }
};

inline SharedContext& shared_context()
{
    thread_local SharedContext s;
    return s;
}

// What the thread's last Clusterer::cluster call saw and returned. The node follows it with one
// Polygonizer::convexHull call per cluster (processor.cpp:627-663); on a GPU that is ~260 host <-> device round
// trips per frame. The Polygonizer therefore looks here first: when the points it is given ARE the gather of the
// next cluster of this cloud (checked coordinate by coordinate), all hulls of the frame are computed in ONE device
// pass (lpl_cluster_hulls) and this and the following calls are answered from that result.
struct ClusterCache
{
    bool valid = false;
    bool hulls_ready = false;
    std::uint32_t num_clusters = 0;
    std::uint32_t next_label = 0;
    std::vector<float> xyz;               // 3 floats per clustered point
    std::vector<std::int32_t> labels;     // as returned by Clusterer::cluster
    std::vector<std::uint32_t> start;     // [K + 1] first member of every cluster in `members`
    std::vector<std::uint32_t> members;   // point indices grouped by cluster, cloud order inside a cluster
    std::vector<std::uint32_t> rank;      // position of a point inside its cluster
    std::vector<std::uint32_t> hull_off;  // [K + 1]
    std::vector<std::int32_t> hull_idx;   // cloud index per hull vertex
    std::vector<float> hull_xy, zminmax;

    void invalidate() noexcept
    {
        valid = false;
        hulls_ready = false;
    }

    void index_members()
    {
        const std::size_t n = labels.size();
        start.assign(static_cast<std::size_t>(num_clusters) + 1, 0U);
        for (std::size_t i = 0; i < n; ++i)
        {
            if (labels[i] >= 0)
            {
                start[static_cast<std::size_t>(labels[i]) + 1] += 1U;
            }
        }
        for (std::uint32_t k = 0; k < num_clusters; ++k)
        {
            start[k + 1] += start[k];
        }
        members.assign(start[num_clusters], 0U);
        rank.assign(n, 0U);
        std::vector<std::uint32_t> fill(start.begin(), start.end() - 1);
        for (std::size_t i = 0; i < n; ++i)
        {
            if (labels[i] >= 0)
            {
                const std::uint32_t pos = fill[static_cast<std::size_t>(labels[i])]++;
                members[pos] = static_cast<std::uint32_t>(i);
                rank[i] = pos - start[static_cast<std::size_t>(labels[i])];
            }
        }
    }
};

inline ClusterCache& cluster_cache()
{
    thread_local ClusterCache c;
    return c;
}

class Handle
{
  public:
    Handle() = default;
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    ~Handle() { reset(); }

    void reset()
    {
        if (attached_)
        {
            SharedContext& s = shared_context();
            attached_ = false;
            if (--s.users <= 0 && s.ctx != nullptr)
            {
                lpl_destroy(s.ctx);
                s = SharedContext{};
            }
        }
    }

    // the thread's context, (re)created when the capacity has to grow or the image size changes; image size 0 =
    // "whatever the context has" (only the Segmenter cares)
    lpl_ctx* ensure(std::uint32_t max_points, std::int32_t image_height = 0, std::int32_t image_width = 0, int device = 0)
    {
        SharedContext& s = shared_context();
        if (!attached_)
        {
            attached_ = true;
            s.users += 1;
        }
        const std::int32_t h = image_height > 0 ? image_height : (s.height > 0 ? s.height : 64);
        const std::int32_t w = image_width > 0 ? image_width : (s.width > 0 ? s.width : 2048);
        if (s.ctx == nullptr || max_points > s.max_points || h != s.height || w != s.width)
        {
            if (s.ctx != nullptr)
            {
                lpl_destroy(s.ctx);
                s.ctx = nullptr;
            }
            const std::uint32_t cap = max_points > s.max_points ? max_points : s.max_points;
            const int rc = lpl_create(&s.ctx, device, cap, 1U, h, w);
            if (rc != LPL_OK)
            {
                s.ctx = nullptr;
                s.max_points = 0;
                raise(rc, nullptr, "lpl_create");
            }
            s.max_points = cap;
            s.height = h;
            s.width = w;
            s.generation += 1;
        }
        return s.ctx;
    }

    lpl_ctx* get() const noexcept { return shared_context().ctx; }
    std::uint32_t capacity() const noexcept { return shared_context().max_points; }

  private:
    bool attached_ = false;
};
} // namespace detail
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP
