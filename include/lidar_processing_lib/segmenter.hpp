// Drop-in Segmenter (JCP ground segmentation on the RECM with near-field RANSAC) over the B200 C ABI.
// Replaces lidar_processing_lib/include/lidar_processing_lib/segmenter.hpp:69-271 and
// src/segmenter.cpp:38-669 for the caller in src/processor/src/processor.cpp:552-556: same Label
// enum, SegmenterConfiguration fields, config()/image()/segment<PointT>() signatures, ownership
// (labels is assigned to the cloud size; image() stays valid until the next segment()).
#ifndef LIDAR_PROCESSING_LIB__SEGMENTER_HPP
#define LIDAR_PROCESSING_LIB__SEGMENTER_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <ostream>
#include <type_traits>
#include <vector>

#include <opencv2/opencv.hpp>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "detail/lpl_handle.hpp"
#include "point_types.hpp"

namespace lidar_processing_lib
{
enum class Label : std::uint32_t
{
    UNKNOWN = 0,
    GROUND,
    OBSTACLE
};

struct SegmenterConfiguration
{
    // sensor (defaults: Velodyne HDL-64E)
    float elevation_up_deg = 2.0F;
    float elevation_down_deg = -24.8F;
    std::int32_t image_width = 2048;
    std::int32_t image_height = 64;

    // algorithm
    bool assume_unorganized_cloud = false;
    float grid_radial_spacing_m = 2.0F;
    float grid_slice_resolution_deg = 1.0F;
    float ground_height_threshold_m = 0.2F;
    float road_maximum_slope_m_per_m = 0.2F;
    float min_distance_m = 2.0F;
    float max_distance_m = 100.0F;
    float sensor_height_m = 1.73F;
    float kernel_threshold_distance_m = 1.0F;
    float amplification_factor = 5.0F;
    float z_min_m = -3.0F;
    float z_max_m = 4.0F;

    // accepted for source compatibility; the GPU path never opens a window
    bool display_recm_with_low_confidence_points = false;
};

inline std::ostream& operator<<(std::ostream& os, const SegmenterConfiguration& c)
{
    return os << "grid_radial_spacing_m: " << c.grid_radial_spacing_m << "\n"
              << "grid_slice_resolution_deg: " << c.grid_slice_resolution_deg << "\n"
              << "ground_height_threshold_m: " << c.ground_height_threshold_m << "\n"
              << "road_maximum_slope_m_per_m: " << c.road_maximum_slope_m_per_m << "\n"
              << "min_distance_m: " << c.min_distance_m << "\n"
              << "max_distance_m: " << c.max_distance_m << "\n"
              << "sensor_height_m: " << c.sensor_height_m << "\n"
              << "kernel_threshold_distance_m: " << c.kernel_threshold_distance_m << "\n"
              << "amplification_factor: " << c.amplification_factor << "\n"
              << "z_min_m: " << c.z_min_m << "\n"
              << "z_max_m: " << c.z_max_m;
}

namespace detail
{
template <typename PointT, typename = void>
struct RingOffset
{
    static constexpr std::int32_t value = -1; // ring-less point type: height index from the elevation angle
};
template <typename PointT>
struct RingOffset<PointT, std::void_t<decltype(std::declval<PointT>().ring)>>
{
    static constexpr std::int32_t value = static_cast<std::int32_t>(offsetof(PointT, ring));
};
} // namespace detail

// One binned point of the polar grid (segmenter.hpp:76-85 of the reference). The GPU path keeps the grid as
// device planes; the type stays for source compatibility.
struct SegmenterPoint
{
    float x;
    float y;
    float z;
    Label label;
    std::uint16_t height_index;
    std::uint16_t width_index;
    std::uint32_t cloud_index;
};

class Segmenter
{
  public:
    // public constants of the reference class (segmenter.hpp:131-151)
    static constexpr float DEG_TO_RAD = static_cast<float>(M_PI / 180.0);
    static constexpr float TWO_M_PIf = static_cast<float>(2.0 * M_PI);
    static constexpr std::int32_t INVALID_INDEX = -1;
    static constexpr float INVALID_Z = std::numeric_limits<float>::max();
    static constexpr float INVALID_DEPTH_M = std::numeric_limits<float>::max();
    static constexpr std::uint32_t MAX_CLOUD_SIZE = 200'000U;
    inline static const cv::Vec3b CV_OBSTACLE{0, 0, 255};
    inline static const cv::Vec3b CV_GROUND{0, 255, 0};
    inline static const cv::Vec3b CV_INTERSECTION_OR_UNKNOWN{0, 255, 255};
    inline static const cv::Vec3b CV_INTERSECTION{255, 0, 0};
    inline static const cv::Vec3b CV_UNKNOWN{255, 255, 255};

    Segmenter() { config(SegmenterConfiguration{}); }
    ~Segmenter() = default;

    void config(const SegmenterConfiguration& config)
    {
        config_ = config;
        image_.create(config_.image_height, config_.image_width, CV_8UC3);
        image_.setTo(cv::Scalar(0, 0, 0));
    }

    inline const SegmenterConfiguration& config() const noexcept { return config_; }

    inline const cv::Mat& image() const noexcept { return image_; }

    template <typename PointT>
    void segment(const pcl::PointCloud<PointT>& cloud, std::vector<Label>& labels)
    {
        static_assert(sizeof(Label) == sizeof(std::uint32_t), "Label must stay a 32-bit enum");
        labels.assign(cloud.points.size(), Label::UNKNOWN);
        const auto n = static_cast<std::uint32_t>(cloud.points.size());
        lpl_ctx* ctx = handle_.ensure(n > MAX_CLOUD_SIZE ? n : MAX_CLOUD_SIZE, config_.image_height, config_.image_width);
        {
            // the context is shared with the thread's other adaptor objects: the configuration goes with every call
            lpl_segmenter_cfg c{};
            c.elevation_up_deg = config_.elevation_up_deg;
            c.elevation_down_deg = config_.elevation_down_deg;
            c.image_width = config_.image_width;
            c.image_height = config_.image_height;
            c.assume_unorganized_cloud = config_.assume_unorganized_cloud ? 1 : 0;
            c.grid_radial_spacing_m = config_.grid_radial_spacing_m;
            c.grid_slice_resolution_deg = config_.grid_slice_resolution_deg;
            c.ground_height_threshold_m = config_.ground_height_threshold_m;
            c.road_maximum_slope_m_per_m = config_.road_maximum_slope_m_per_m;
            c.min_distance_m = config_.min_distance_m;
            c.max_distance_m = config_.max_distance_m;
            c.sensor_height_m = config_.sensor_height_m;
            c.kernel_threshold_distance_m = config_.kernel_threshold_distance_m;
            c.amplification_factor = config_.amplification_factor;
            c.z_min_m = config_.z_min_m;
            c.z_max_m = config_.z_max_m;
            detail::check(lpl_segmenter_config(ctx, &c), ctx, "Segmenter::config");
        }
        // an empty cloud is a valid no-op that still clears the image (segmenter.cpp:73-85,116-119)
        detail::check(lpl_segment(ctx, cloud.points.data(), sizeof(PointT), detail::RingOffset<PointT>::value, n,
                                  reinterpret_cast<std::uint32_t*>(labels.data()), image_.template ptr<std::uint8_t>(0)),
                      ctx, "Segmenter::segment");
    }

  private:
    SegmenterConfiguration config_{};
    cv::Mat image_;
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__SEGMENTER_HPP
