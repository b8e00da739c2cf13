// Drop-in NoiseRemover (DROR) over the B200 C ABI.
// Replaces lidar_processing_lib/include/lidar_processing_lib/noise_remover.hpp:35-88 and
// src/noise_remover.cpp:38-68: same enum, configuration struct, method names and ownership
// (the caller owns both vectors; `labels` is resized to the input size).
// Semantics: a point is NOISE iff fewer than min_neighbours points (itself included) lie within
// r = max(radius_multiplier * range_xy, min_search_radius), evaluated with the reference's float
// expressions. The reference's KD-tree query leaves a stale traversal stack behind its early exit,
// which makes its own answer depend on the tree shape (hazard H1); this is the
// well-defined neighbour count.
#ifndef LIDAR_PROCESSING_LIB__NOISE_REMOVER_HPP
#define LIDAR_PROCESSING_LIB__NOISE_REMOVER_HPP

#include <array>
#include <cstdint>
#include <vector>

#include "detail/lpl_handle.hpp"

namespace lidar_processing_lib
{
enum class NoiseRemoverLabel : std::uint8_t
{
    VALID = 0,
    NOISE = 1
};

struct NoiseRemoverConfiguration
{
    float radius_multiplier_m_per_m = 0.02F; // search radius per metre of xy range
    float min_search_radius_m = 0.1F;        // lower bound of the search radius
    std::uint32_t min_neighbours = 4U;       // neighbours (self included) a VALID point needs
};

class NoiseRemover final
{
  public:
    using PointT = std::array<float, 3>; // == KDTree<float, 3>::PointT of the reference
    struct NeighbourT                     // == KDTree<float, 3>::Neighbour (kdtree.hpp:50-54)
    {
        std::uint32_t index;
        float distance;
    };

    NoiseRemover() = default;

    void filter(const std::vector<PointT>& points, std::vector<NoiseRemoverLabel>& labels)
    {
        static_assert(sizeof(NoiseRemoverLabel) == 1 && sizeof(PointT) == 12, "layout the C ABI relies on");
        labels.assign(points.size(), NoiseRemoverLabel::VALID);
        if (points.empty())
        {
            return;
        }
        const auto n = static_cast<std::uint32_t>(points.size());
        lpl_ctx* ctx = handle_.ensure(n > reserve_ ? n : reserve_);
        push_config(ctx);
        detail::check(lpl_dror_filter(ctx, points.data(), sizeof(PointT), n,
                                      reinterpret_cast<std::uint8_t*>(labels.data())),
                      ctx, "NoiseRemover::filter");
    }

    void reserve(std::uint32_t max_pts) { reserve_ = max_pts; }

    void config(const NoiseRemoverConfiguration& config) { config_ = config; }
    const NoiseRemoverConfiguration& config() const noexcept { return config_; }

  private:
    void push_config(lpl_ctx* ctx)
    {
        const lpl_dror_cfg c{config_.radius_multiplier_m_per_m, config_.min_search_radius_m, config_.min_neighbours};
        detail::check(lpl_dror_config(ctx, &c), ctx, "NoiseRemover::config");
    }

    NoiseRemoverConfiguration config_{};
    std::uint32_t reserve_ = 131072U;
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__NOISE_REMOVER_HPP
