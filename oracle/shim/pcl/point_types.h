// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <pcl/point_types.h>.
//
// PCL is not installed in this image. The reference sources under
// /root/reference/lidar_processing_lib/src are compiled UNMODIFIED against this
// shim by oracle/Makefile to produce oracle/_ref/libref_oracle.so. Only the few
// names those sources touch are provided; layouts follow PCL's documented point
// layouts (xyz + pad in the first 16 bytes, 16-byte alignment).
#ifndef ORACLE_SHIM_PCL_POINT_TYPES_H
#define ORACLE_SHIM_PCL_POINT_TYPES_H

#include <cstddef>
#include <cstdint>

#define EIGEN_ALIGN16 alignas(16)
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define PCL_ADD_POINT4D                                                                            \
    union {                                                                                        \
        float data[4];                                                                             \
        struct                                                                                     \
        {                                                                                          \
            float x;                                                                               \
            float y;                                                                               \
            float z;                                                                               \
        };                                                                                         \
    };
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)

namespace pcl
{
struct alignas(16) PointXYZ
{
    PCL_ADD_POINT4D
    PointXYZ() : data{0.F, 0.F, 0.F, 1.F} {}
    PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.F} {}
};

struct alignas(16) PointXYZI
{
    PCL_ADD_POINT4D
    union {
        struct
        {
            float intensity;
        };
        float data_c[4];
    };
    PointXYZI() : data{0.F, 0.F, 0.F, 1.F}, data_c{0.F, 0.F, 0.F, 0.F} {}
};

struct alignas(16) PointXYZRGB
{
    PCL_ADD_POINT4D
    union {
        struct
        {
            std::uint8_t b;
            std::uint8_t g;
            std::uint8_t r;
            std::uint8_t a;
        };
        float rgb;
        std::uint32_t rgba;
    };
    float pad_[3];
    PointXYZRGB() : data{0.F, 0.F, 0.F, 1.F}, rgba{0xff000000U}, pad_{0.F, 0.F, 0.F} {}
    PointXYZRGB(float x_, float y_, float z_, std::uint8_t r_, std::uint8_t g_, std::uint8_t b_)
        : data{x_, y_, z_, 1.F}, pad_{0.F, 0.F, 0.F}
    {
        b = b_;
        g = g_;
        r = r_;
        a = 255;
    }
};
} // namespace pcl

// The reference's segmenter.hpp uses Eigen::Vector<float, 24> (it gets Eigen through PCL).
namespace Eigen
{
template <typename T, int N>
class Vector
{
  public:
    void fill(T v)
    {
        for (int i = 0; i < N; ++i)
        {
            v_[i] = v;
        }
    }
    T& operator[](std::size_t i) { return v_[i]; }
    const T& operator[](std::size_t i) const { return v_[i]; }
    int size() const { return N; }
    Vector& noalias() { return *this; }
    // Eigen's `vector / scalar` is an element-wise IEEE division (scalar_quotient_op).
    friend Vector operator/(const Vector& a, T s)
    {
        Vector r;
        for (int i = 0; i < N; ++i)
        {
            r.v_[i] = a.v_[i] / s;
        }
        return r;
    }

  private:
    T v_[N];
};
} // namespace Eigen

#endif
