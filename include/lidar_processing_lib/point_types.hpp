// Ring-carrying PCL point types of the drop-in surface (reference:
// lidar_processing_lib/include/lidar_processing_lib/point_types.hpp:10-31). Both are 16-byte
// aligned PCL 4D points followed by the fields below; the GPU path reads x, y, z from the first
// 12 bytes of a record and the ring at its byte offset.
#ifndef LIDAR_PROCESSING_LIB__POINT_TYPES_HPP
#define LIDAR_PROCESSING_LIB__POINT_TYPES_HPP

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdint>

namespace pcl
{
struct EIGEN_ALIGN16 PointXYZR
{
    PCL_ADD_POINT4D;
    std::uint16_t ring;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

struct EIGEN_ALIGN16 PointXYZIR
{
    PCL_ADD_POINT4D;
    float intensity;
    std::uint16_t ring;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
} // namespace pcl

POINT_CLOUD_REGISTER_POINT_STRUCT(pcl::PointXYZR,
                                  (float, x, x)(float, y, y)(float, z, z)(std::uint16_t, ring, ring))
POINT_CLOUD_REGISTER_POINT_STRUCT(pcl::PointXYZIR, (float, x, x)(float, y, y)(float, z, z)(float, intensity, intensity)(
                                                       std::uint16_t, ring, ring))

#endif // LIDAR_PROCESSING_LIB__POINT_TYPES_HPP
