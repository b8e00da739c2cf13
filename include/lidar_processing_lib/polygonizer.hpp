// Drop-in Polygonizer::convexHull over the B200 C ABI.
// Replaces lidar_processing_lib/include/lidar_processing_lib/polygonizer.hpp:41-243 /
// src/polygonizer.cpp:33-91 for the caller in src/processor/src/processor.cpp:663. The batched
// entry point the GPU is built for is lpl_cluster_hulls (all clusters of a frame in one call, see
// hulls()); convexHull() keeps the reference's one-polygon-per-call signature.
// The oriented-bounding-box members are declared so that the node still compiles, but they are
// outside the hot path (dead code in the node: processor.cpp:676 `perform_polygon_simplification =
// false`) and throw.
#ifndef LIDAR_PROCESSING_LIB__POLYGONIZER_HPP
#define LIDAR_PROCESSING_LIB__POLYGONIZER_HPP

#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "detail/lpl_handle.hpp"

namespace lidar_processing_lib
{
enum class Orientation : std::uint8_t
{
    ANTICLOCKWISE = 0,
    CLOCKWISE = 1,
    COLINEAR = 2
};

enum class PolygonContour : std::uint8_t
{
    OPEN = 0,
    ENCLOSED = 1
};

struct PolygonizerConfiguration
{
    Orientation orientation = Orientation::ANTICLOCKWISE;
    PolygonContour contour = PolygonContour::OPEN;

    std::uint32_t max_points = 100'000;
};

struct PointXY
{
    double x;
    double y;
};

struct PointXYZ
{
    double x;
    double y;
    double z;
};

struct BoundingBox
{
    std::array<PointXY, 4> corners;
    float area;
    float angle_rad;
    bool is_valid;
};

struct AntipodalPair final
{
    std::int32_t index_1;
    std::int32_t index_2;
};

class Polygonizer final
{
  public:
    Polygonizer() = default;

    /// Indices of the convex hull vertices, counter-clockwise from the lexicographically smallest
    /// point, collinear points dropped; fewer than three points are returned as they are.
    /// Coordinates must be float-representable (they are when they come from a PCL cloud, as in
    /// processor.cpp:645-646); otherwise std::invalid_argument.
    template <typename PointT>
    void convexHull(const std::vector<PointT>& points, std::vector<std::int32_t>& indices)
    {
        static_assert(sizeof(PointT) >= 2 * sizeof(double), "PointXY / PointXYZ of doubles");
        indices.clear();
        if (points.empty())
        {
            return;
        }
        const auto n = static_cast<std::uint32_t>(points.size());
        lpl_ctx* ctx = handle_.ensure(n > config_.max_points ? n : config_.max_points);
        indices.resize(n);
        std::uint32_t count = 0;
        detail::check(lpl_convex_hull(ctx, points.data(), sizeof(PointT), n, indices.data(), &count), ctx,
                      "Polygonizer::convexHull");
        indices.resize(count);
    }

    void findAntipodalPairsOfConvexHull(const std::vector<PointXY>&, std::vector<AntipodalPair>&)
    {
        throw std::logic_error("Polygonizer::findAntipodalPairsOfConvexHull is outside the lpl_b200 hot path");
    }
    BoundingBox boundingBoxRotatingCalipers(const std::vector<PointXY>&)
    {
        throw std::logic_error("Polygonizer::boundingBoxRotatingCalipers is outside the lpl_b200 hot path");
    }
    BoundingBox boundingBoxPrincipalComponentAnalysis(const std::vector<PointXY>&)
    {
        throw std::logic_error("Polygonizer::boundingBoxPrincipalComponentAnalysis is outside the lpl_b200 hot path");
    }
    template <typename PointT>
    void concaveHull(const std::vector<PointT>&, std::vector<std::int32_t>&)
    {
        // empty in the reference as well (polygonizer.hpp:233-237)
    }

    void config(const PolygonizerConfiguration& config) { config_ = config; }
    const PolygonizerConfiguration& config() const noexcept { return config_; }

  private:
    PolygonizerConfiguration config_{};
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__POLYGONIZER_HPP
