// Drop-in Clusterer (curved-voxel clustering) over the B200 C ABI.
// Replaces lidar_processing_lib/include/lidar_processing_lib/clusterer.hpp:52-188 and
// src/clusterer.cpp:43-239 for the caller in src/processor/src/processor.cpp:599: labels is
// assigned to the cloud size; clusters are numbered by their first point in cloud order after the
// small ones (< min_cluster_size) became INVALID_LABEL.
#ifndef LIDAR_PROCESSING_LIB__CLUSTERER_HPP
#define LIDAR_PROCESSING_LIB__CLUSTERER_HPP

#include <cmath>
#include <cstdint>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "detail/lpl_handle.hpp"

namespace lidar_processing_lib
{
using ClusterLabel = std::int32_t;

// spherical image of a point (clusterer.hpp:54-59 of the reference); the GPU path keeps these as device planes,
// the type stays for source compatibility
struct SphericalPoint
{
    float range_m;
    float azimuth_rad;   // Convention: 0 -> 2 * pi
    float elevation_rad; // Convention: 0 -> pi
};

struct ClustererConfiguration
{
    float voxel_grid_range_resolution_m = 0.4F;
    float voxel_grid_azimuth_resolution_deg = 1.0F;
    float voxel_grid_elevation_resolution_deg = 1.5F;

    std::uint32_t min_cluster_size = 3;
};

class Clusterer
{
  public:
    static constexpr float TWO_M_PIf = static_cast<float>(2.0 * M_PI);
    static constexpr float DEG_TO_RAD = static_cast<float>(M_PI / 180.0);
    static constexpr std::int32_t INVALID_LABEL = -1;

    Clusterer() = default;

    void config(const ClustererConfiguration& config) { config_ = config; }
    const ClustererConfiguration& config() const noexcept { return config_; }

    template <typename PointT>
    void cluster(const pcl::PointCloud<PointT>& cloud, std::vector<ClusterLabel>& labels)
    {
        labels.assign(cloud.points.size(), INVALID_LABEL);
        detail::cluster_cache().invalidate();
        if (cloud.points.empty())
        {
            return; // clusterer.cpp:62-65
        }
        const auto n = static_cast<std::uint32_t>(cloud.points.size());
        lpl_ctx* ctx = handle_.ensure(n > 200'000U ? n : 200'000U);
        const lpl_cluster_cfg c{config_.voxel_grid_range_resolution_m, config_.voxel_grid_azimuth_resolution_deg,
                                config_.voxel_grid_elevation_resolution_deg, config_.min_cluster_size};
        detail::check(lpl_cluster_config(ctx, &c), ctx, "Clusterer::config");
        std::uint32_t num_clusters = 0;
        detail::check(lpl_cluster(ctx, cloud.points.data(), sizeof(PointT), n, labels.data(), &num_clusters), ctx,
                      "Clusterer::cluster");
        num_clusters_ = num_clusters;
        // remembered for the Polygonizer calls that follow (detail::ClusterCache)
        detail::ClusterCache& cache = detail::cluster_cache();
        cache.xyz.resize(static_cast<std::size_t>(n) * 3);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            cache.xyz[3 * i] = cloud.points[i].x;
            cache.xyz[3 * i + 1] = cloud.points[i].y;
            cache.xyz[3 * i + 2] = cloud.points[i].z;
        }
        cache.labels = labels;
        cache.num_clusters = num_clusters;
        cache.next_label = 0;
        cache.hulls_ready = false;
        cache.valid = true;
        cache.index_members();
    }

    // extension: number of clusters found by the last call (max label + 1)
    std::uint32_t lastClusterCount() const noexcept { return num_clusters_; }

  private:
    ClustererConfiguration config_{};
    std::uint32_t num_clusters_ = 0;
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__CLUSTERER_HPP
