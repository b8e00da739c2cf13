// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <pcl/point_cloud.h>.
// See oracle/shim/pcl/point_types.h for why this exists.
#ifndef ORACLE_SHIM_PCL_POINT_CLOUD_H
#define ORACLE_SHIM_PCL_POINT_CLOUD_H

#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace pcl
{
template <typename PointT>
class PointCloud
{
  public:
    std::vector<PointT> points;
    std::uint32_t width = 0;
    std::uint32_t height = 1;
    bool is_dense = true;

    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear()
    {
        points.clear();
        width = 0;
        height = 1;
    }
    void reserve(std::size_t n) { points.reserve(n); }
    void resize(std::size_t n)
    {
        points.resize(n);
        width = static_cast<std::uint32_t>(n);
        height = 1;
    }
    void push_back(const PointT& p)
    {
        points.push_back(p);
        width = static_cast<std::uint32_t>(points.size());
        height = 1;
    }
    template <typename... Args>
    PointT& emplace_back(Args&&... args)
    {
        points.emplace_back(std::forward<Args>(args)...);
        width = static_cast<std::uint32_t>(points.size());
        height = 1;
        return points.back();
    }
    PointT& operator[](std::size_t i) { return points[i]; }
    const PointT& operator[](std::size_t i) const { return points[i]; }
    auto begin() { return points.begin(); }
    auto end() { return points.end(); }
    auto begin() const { return points.begin(); }
    auto end() const { return points.end(); }
};
} // namespace pcl

#endif
