// Drop-in Polygonizer over the B200 C ABI.
// Replaces lidar_processing_lib/include/lidar_processing_lib/polygonizer.hpp:41-243 /
// src/polygonizer.cpp:33-362 for the callers in src/processor/src/processor.cpp:663,704. The batched
// entry points the GPU is built for are lpl_cluster_hulls (all clusters of a frame in one call) and
// lpl_bounding_boxes (all hulls in one call, see boundingBoxes()); convexHull() and the
// boundingBox*() members keep the reference's one-polygon-per-call signatures.
// findAntipodalPairsOfConvexHull is a few dozen sequential steps on a hull of ~7 vertices: it runs
// inside the box kernel on the device and, as a public member, here on the host (same arithmetic).
#ifndef LIDAR_PROCESSING_LIB__POLYGONIZER_HPP
#define LIDAR_PROCESSING_LIB__POLYGONIZER_HPP

#include <array>
#include <cmath>
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

// free helpers of the reference header (polygonizer.hpp:176-231), used by the node's shape matching
template <typename PointT>
inline double crossProduct(const PointT& p1, const PointT& p2) noexcept
{
    return (p1.x * p2.y) - (p2.x * p1.y);
}

/// Polygon area by the shoelace formula.
template <typename PointT>
inline double polygonArea(const std::vector<PointT>& points) noexcept
{
    double area = 0.0;
    if (const auto n = static_cast<std::int32_t>(points.size()); n > 2)
    {
        for (std::int32_t i = 0; i < n - 1; ++i)
        {
            area += crossProduct(points[i], points[i + 1]);
        }
        area += crossProduct(points[n - 1], points[0]);
    }
    return std::fabs(area) * 0.5;
}

template <typename PointT>
inline double distanceSquared(const PointT& p1, const PointT& p2) noexcept
{
    const double dx = p1.x - p2.x;
    const double dy = p1.y - p2.y;
    return dx * dx + dy * dy;
}

template <typename PointT>
inline double distance(const PointT& p1, const PointT& p2) noexcept
{
    return std::sqrt(distanceSquared(p1, p2));
}

/// Area of a rectangle with 4 ordered vertices.
template <typename PointT>
inline double areaOfRectangle(const PointT& p1, const PointT& p2, const PointT& p3, [[maybe_unused]] const PointT& p4) noexcept
{
    return std::sqrt(distanceSquared(p1, p2) * distanceSquared(p2, p3));
}

/// Area of a triangle with 3 ordered vertices.
template <typename PointT>
inline double areaOfTriangle(const PointT& p1, const PointT& p2, const PointT& p3) noexcept
{
    return std::fabs((p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y)) * 0.5);
}

class Polygonizer final
{
  public:
    Polygonizer() = default;

    /// Indices of the convex hull vertices, counter-clockwise from the lexicographically smallest
    /// point, collinear points dropped; fewer than three points are returned as they are.
    /// Coordinates that are float-representable (they are when they come from a PCL cloud, as in
    /// processor.cpp:645-646) take the batched fast path; any other doubles are sorted and swept in fp64 on the device.
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
        if (from_cluster_cache(points, indices))
        {
            return;
        }
        lpl_ctx* ctx = handle_.ensure(n > config_.max_points ? n : config_.max_points);
        indices.resize(n);
        std::uint32_t count = 0;
        detail::check(lpl_convex_hull(ctx, points.data(), sizeof(PointT), n, indices.data(), &count), ctx,
                      "Polygonizer::convexHull");
        indices.resize(count);
    }

    /// Batched form of the node's per-label loop (processor.cpp:627-663: for every label, gather the cluster's
    /// points in cloud order, z extent, convexHull): ONE device pass for all clusters of a frame instead of one
    /// host <-> device round trip per cluster. `cloud` is any container of points with float x, y, z whose first
    /// 12 bytes are x, y, z (every PCL point type), `labels[i]` in [-1, num_clusters).
    /// hull k = hull_points[hull_offsets[k] .. hull_offsets[k + 1]) with hull_indices pointing into `cloud`
    /// (the reference's per-call indices are positions inside the gathered cluster; these are cloud positions);
    /// z_min_max[k] = {z_min, z_max} as the node computes them.
    template <typename CloudPointT>
    void convexHulls(const std::vector<CloudPointT>& cloud, const std::vector<std::int32_t>& labels,
                     std::uint32_t num_clusters, std::vector<std::uint32_t>& hull_offsets,
                     std::vector<std::int32_t>& hull_indices, std::vector<PointXY>& hull_points,
                     std::vector<std::array<double, 2>>& z_min_max)
    {
        static_assert(sizeof(CloudPointT) >= 3 * sizeof(float), "points start with float x, y, z");
        if (labels.size() != cloud.size())
        {
            throw std::invalid_argument("Polygonizer::convexHulls: one label per point");
        }
        const auto n = static_cast<std::uint32_t>(cloud.size());
        hull_offsets.assign(static_cast<std::size_t>(num_clusters) + 1, 0U);
        hull_indices.assign(n, 0);
        hull_points.clear();
        z_min_max.assign(num_clusters, {0.0, 0.0});
        if (n == 0 || num_clusters == 0)
        {
            hull_indices.clear();
            return;
        }
        lpl_ctx* ctx = handle_.ensure(n > config_.max_points ? n : config_.max_points);
        xy_scratch_.resize(static_cast<std::size_t>(n) * 2);
        z_scratch_.resize(static_cast<std::size_t>(num_clusters) * 2);
        detail::check(lpl_cluster_hulls(ctx, cloud.data(), sizeof(CloudPointT), labels.data(), n, num_clusters,
                                        hull_offsets.data(), hull_indices.data(), xy_scratch_.data(), z_scratch_.data()),
                      ctx, "Polygonizer::convexHulls");
        const std::uint32_t total = hull_offsets[num_clusters];
        hull_indices.resize(total);
        hull_points.resize(total);
        for (std::uint32_t v = 0; v < total; ++v)
        {
            hull_points[v] = {static_cast<double>(xy_scratch_[2 * v]), static_cast<double>(xy_scratch_[2 * v + 1])};
        }
        for (std::uint32_t k = 0; k < num_clusters; ++k)
        {
            z_min_max[k] = {static_cast<double>(z_scratch_[2 * k]), static_cast<double>(z_scratch_[2 * k + 1])};
        }
    }

    /// Shamos' antipodal pairs of a convex polygon (src/polygonizer.cpp:93-163), host-side.
    void findAntipodalPairsOfConvexHull(const std::vector<PointXY>& h, std::vector<AntipodalPair>& pairs)
    {
        pairs.clear();
        const auto n = static_cast<std::int32_t>(h.size());
        if (n < 2)
        {
            return;
        }
        const auto area = [](const PointXY& p1, const PointXY& p2, const PointXY& p3) {
            return std::fabs((p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y)) * 0.5);
        };
        const auto nx = [n](std::int32_t k) { return (k + 1 == n) ? 0 : (k + 1); };
        const std::int32_t i0 = n - 1;
        std::int32_t i = 0, j = 1;
        while (area(h[i], h[nx(i)], h[nx(j)]) > area(h[i], h[nx(i)], h[j]))
        {
            j = nx(j);
        }
        const std::int32_t j0 = j;
        while (i != j0)
        {
            i = nx(i);
            pairs.push_back({i, j});
            while (area(h[i], h[nx(i)], h[nx(j)]) > area(h[i], h[nx(i)], h[j]))
            {
                j = nx(j);
                if (i == j0 && j == i0)
                {
                    return;
                }
                pairs.push_back({i, j});
            }
            if (area(h[j], h[nx(i)], h[nx(j)]) == area(h[i], h[nx(i)], h[j]))
            {
                pairs.push_back((i == j0 && j == i0) ? AntipodalPair{nx(i), j} : AntipodalPair{i, nx(j)});
            }
        }
    }

    /// Minimum-area oriented box by rotating calipers (src/polygonizer.cpp:165-278).
    BoundingBox boundingBoxRotatingCalipers(const std::vector<PointXY>& convex_hull_points)
    {
        return box(convex_hull_points, LPL_BOX_ROTATING_CALIPERS, "Polygonizer::boundingBoxRotatingCalipers");
    }

    /// Principal-axes box (src/polygonizer.cpp:280-362).
    BoundingBox boundingBoxPrincipalComponentAnalysis(const std::vector<PointXY>& convex_hull_points)
    {
        return box(convex_hull_points, LPL_BOX_PCA, "Polygonizer::boundingBoxPrincipalComponentAnalysis");
    }

    /// Batched form: hull k is hull_points[offsets[k] .. offsets[k + 1]).
    void boundingBoxes(const std::vector<PointXY>& hull_points, const std::vector<std::uint32_t>& offsets,
                       std::vector<BoundingBox>& boxes, int method = LPL_BOX_ROTATING_CALIPERS)
    {
        boxes.clear();
        if (offsets.size() < 2)
        {
            return;
        }
        const auto k = static_cast<std::uint32_t>(offsets.size() - 1);
        const auto n = static_cast<std::uint32_t>(hull_points.size());
        lpl_ctx* ctx = handle_.ensure(n > config_.max_points ? n : config_.max_points);
        std::vector<lpl_bbox> raw(k);
        detail::check(lpl_bounding_boxes(ctx, hull_points.data(), sizeof(PointXY), offsets.data(), k, method, raw.data()),
                      ctx, "Polygonizer::boundingBoxes");
        boxes.resize(k);
        for (std::uint32_t b = 0; b < k; ++b)
        {
            boxes[b] = convert(raw[b]);
        }
    }

    template <typename PointT>
    void concaveHull(const std::vector<PointT>&, std::vector<std::int32_t>&)
    {
        // empty in the reference as well (polygonizer.hpp:233-237)
    }

    void config(const PolygonizerConfiguration& config) { config_ = config; }
    const PolygonizerConfiguration& config() const noexcept { return config_; }

  private:
    // The node calls convexHull once per cluster right after Clusterer::cluster, labels ascending
    // (processor.cpp:627-663). If `points` is exactly the gather of the next cluster of the thread's last
    // clustered cloud, answer from ONE batched device pass over all clusters of that cloud. Everything is verified
    // against the coordinates given, so a caller that does something else simply takes the per-call path.
    template <typename PointT>
    bool from_cluster_cache(const std::vector<PointT>& points, std::vector<std::int32_t>& indices)
    {
        detail::ClusterCache& c = detail::cluster_cache();
        if (!c.valid || c.num_clusters == 0)
        {
            return false;
        }
        const auto n = static_cast<std::uint32_t>(points.size());
        std::uint32_t l = c.next_label < c.num_clusters ? c.next_label : 0U;
        if (c.start[l + 1] - c.start[l] != n)
        {
            return false;
        }
        const std::uint32_t* mem = c.members.data() + c.start[l];
        for (std::uint32_t j = 0; j < n; ++j)
        {
            const float* q = c.xyz.data() + static_cast<std::size_t>(mem[j]) * 3;
            if (points[j].x != static_cast<double>(q[0]) || points[j].y != static_cast<double>(q[1]))
            {
                return false;
            }
        }
        if (!c.hulls_ready)
        {
            const auto m = static_cast<std::uint32_t>(c.labels.size());
            lpl_ctx* ctx = handle_.ensure(m > config_.max_points ? m : config_.max_points);
            c.hull_off.assign(static_cast<std::size_t>(c.num_clusters) + 1, 0U);
            c.hull_idx.assign(m, 0);
            c.hull_xy.assign(static_cast<std::size_t>(m) * 2, 0.F);
            c.zminmax.assign(static_cast<std::size_t>(c.num_clusters) * 2, 0.F);
            detail::check(lpl_cluster_hulls(ctx, c.xyz.data(), 3 * sizeof(float), c.labels.data(), m, c.num_clusters,
                                            c.hull_off.data(), c.hull_idx.data(), c.hull_xy.data(), c.zminmax.data()),
                          ctx, "Polygonizer::convexHull (batched over the clustered cloud)");
            c.hulls_ready = true;
        }
        const std::uint32_t a = c.hull_off[l], b = c.hull_off[l + 1];
        indices.resize(b - a);
        for (std::uint32_t v = a; v < b; ++v)
        {
            indices[v - a] = static_cast<std::int32_t>(c.rank[static_cast<std::size_t>(c.hull_idx[v])]);
        }
        c.next_label = l + 1;
        return true;
    }

    static BoundingBox convert(const lpl_bbox& r)
    {
        BoundingBox b{};
        for (int k = 0; k < 4; ++k)
        {
            b.corners[k] = {r.corners[k][0], r.corners[k][1]};
        }
        b.area = r.area;
        b.angle_rad = r.angle_rad;
        b.is_valid = r.is_valid != 0;
        return b;
    }

    BoundingBox box(const std::vector<PointXY>& hull, int method, const char* what)
    {
        BoundingBox invalid{};
        invalid.is_valid = false;
        if (hull.size() < 3)
        {
            return invalid; // src/polygonizer.cpp:172-175, 285-289
        }
        const auto n = static_cast<std::uint32_t>(hull.size());
        lpl_ctx* ctx = handle_.ensure(n > config_.max_points ? n : config_.max_points);
        const std::uint32_t offsets[2] = {0u, n};
        lpl_bbox raw{};
        detail::check(lpl_bounding_boxes(ctx, hull.data(), sizeof(PointXY), offsets, 1, method, &raw), ctx, what);
        return convert(raw);
    }

    PolygonizerConfiguration config_{};
    std::vector<float> xy_scratch_, z_scratch_;
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__POLYGONIZER_HPP
