// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <opencv2/opencv.hpp>.
//
// OpenCV's C++ headers are not installed in this image. The reference segmenter
// (src/segmenter.cpp:486-489) only needs an 8-bit image container, split/merge and a
// 5x5 rectangular dilation. cv::dilate with the default border (BORDER_CONSTANT,
// morphologyDefaultBorderValue) ignores out-of-image taps, i.e. it is the in-bounds
// window maximum; tests/test_oracle_cpu.py checks this shim against cv2.dilate.
#ifndef ORACLE_SHIM_OPENCV_HPP
#define ORACLE_SHIM_OPENCV_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16

namespace cv
{
enum MorphShapes
{
    MORPH_RECT = 0
};

struct Size
{
    int width = 0;
    int height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar
{
    double val[4];
    Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
};

struct Vec3b
{
    std::uint8_t val[3];
    Vec3b() : val{0, 0, 0} {}
    Vec3b(std::uint8_t a, std::uint8_t b, std::uint8_t c) : val{a, b, c} {}
    bool operator==(const Vec3b& o) const
    {
        return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2];
    }
    bool operator!=(const Vec3b& o) const { return !(*this == o); }
    std::uint8_t& operator[](int i) { return val[i]; }
    const std::uint8_t& operator[](int i) const { return val[i]; }
};

class Mat
{
  public:
    int rows = 0;
    int cols = 0;

    Mat() = default;

    static Mat zeros(int r, int c, int type)
    {
        Mat m;
        m.create(r, c, type);
        std::fill(m.buf_.begin(), m.buf_.end(), 0);
        return m;
    }

    void create(int r, int c, int type)
    {
        rows = r;
        cols = c;
        channels_ = (type >> 3) + 1;
        buf_.assign(static_cast<std::size_t>(r) * c * channels_, 0);
    }

    int depth() const { return CV_8U; }
    int channels() const { return channels_; }
    int type() const { return (channels_ - 1) << 3; }
    bool empty() const { return buf_.empty(); }

    Mat& setTo(const Scalar& s)
    {
        for (std::size_t i = 0; i < buf_.size(); ++i)
        {
            buf_[i] = static_cast<std::uint8_t>(s.val[i % channels_]);
        }
        return *this;
    }

    template <typename T>
    T& at(int r, int c)
    {
        return *reinterpret_cast<T*>(&buf_[(static_cast<std::size_t>(r) * cols + c) * sizeof(T)]);
    }
    template <typename T>
    const T& at(int r, int c) const
    {
        return *reinterpret_cast<const T*>(
            &buf_[(static_cast<std::size_t>(r) * cols + c) * sizeof(T)]);
    }

    std::uint8_t* data() { return buf_.data(); }
    const std::uint8_t* data() const { return buf_.data(); }
    // cv::Mat::ptr<T>(row): the accessor the drop-in adaptors use (same spelling as real OpenCV)
    template <typename T = std::uint8_t>
    T* ptr(int r = 0)
    {
        return reinterpret_cast<T*>(buf_.data() + static_cast<std::size_t>(r) * cols * channels_);
    }
    template <typename T = std::uint8_t>
    const T* ptr(int r = 0) const
    {
        return reinterpret_cast<const T*>(buf_.data() + static_cast<std::size_t>(r) * cols * channels_);
    }

  private:
    int channels_ = 1;
    std::vector<std::uint8_t> buf_;
};

inline Mat getStructuringElement(int /*shape*/, Size ksize)
{
    Mat k;
    k.create(ksize.height, ksize.width, CV_8UC1);
    k.setTo(Scalar(1));
    return k;
}

inline void split(const Mat& src, std::vector<Mat>& dst)
{
    const int ch = src.channels();
    dst.resize(ch);
    for (int c = 0; c < ch; ++c)
    {
        if (dst[c].rows != src.rows || dst[c].cols != src.cols || dst[c].channels() != 1)
        {
            dst[c].create(src.rows, src.cols, CV_8UC1);
        }
    }
    const std::size_t n = static_cast<std::size_t>(src.rows) * src.cols;
    for (std::size_t i = 0; i < n; ++i)
    {
        for (int c = 0; c < ch; ++c)
        {
            dst[c].data()[i] = src.data()[i * ch + c];
        }
    }
}

inline void merge(const std::vector<Mat>& src, Mat& dst)
{
    const int ch = static_cast<int>(src.size());
    if (dst.rows != src[0].rows || dst.cols != src[0].cols || dst.channels() != ch)
    {
        dst.create(src[0].rows, src[0].cols, (ch - 1) << 3);
    }
    const std::size_t n = static_cast<std::size_t>(dst.rows) * dst.cols;
    for (std::size_t i = 0; i < n; ++i)
    {
        for (int c = 0; c < ch; ++c)
        {
            dst.data()[i * ch + c] = src[c].data()[i];
        }
    }
}

// Grey-scale dilation = window maximum over in-image taps, anchor at the kernel centre.
inline void dilate(const Mat& src, Mat& dst, const Mat& kernel)
{
    const int kh = kernel.rows;
    const int kw = kernel.cols;
    const int ay = kh / 2;
    const int ax = kw / 2;
    const int rows = src.rows;
    const int cols = src.cols;
    std::vector<std::uint8_t> in(src.data(), src.data() + static_cast<std::size_t>(rows) * cols);
    if (dst.rows != rows || dst.cols != cols || dst.channels() != 1)
    {
        dst.create(rows, cols, CV_8UC1);
    }
    // separable for a full rectangle: horizontal then vertical maximum
    std::vector<std::uint8_t> tmp(in.size());
    for (int r = 0; r < rows; ++r)
    {
        for (int c = 0; c < cols; ++c)
        {
            std::uint8_t m = 0;
            const int c0 = std::max(0, c - ax);
            const int c1 = std::min(cols - 1, c - ax + kw - 1);
            for (int cc = c0; cc <= c1; ++cc)
            {
                m = std::max(m, in[static_cast<std::size_t>(r) * cols + cc]);
            }
            tmp[static_cast<std::size_t>(r) * cols + c] = m;
        }
    }
    for (int r = 0; r < rows; ++r)
    {
        const int r0 = std::max(0, r - ay);
        const int r1 = std::min(rows - 1, r - ay + kh - 1);
        for (int c = 0; c < cols; ++c)
        {
            std::uint8_t m = 0;
            for (int rr = r0; rr <= r1; ++rr)
            {
                m = std::max(m, tmp[static_cast<std::size_t>(rr) * cols + c]);
            }
            dst.data()[static_cast<std::size_t>(r) * cols + c] = m;
        }
    }
}

inline void flip(const Mat& src, Mat& dst, int /*flip_code*/) { dst = src; }
inline void imshow(const std::string& /*name*/, const Mat& /*m*/) {}
inline int waitKey(int /*delay*/) { return -1; }
} // namespace cv

#endif
