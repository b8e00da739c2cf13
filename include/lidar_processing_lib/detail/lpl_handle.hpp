// Shared plumbing of the drop-in adaptors: an owning wrapper around lpl_ctx and the translation of
// C-ABI status codes into the exceptions the reference library throws
// (static_unordered_map.hpp:79-82 std::overflow_error, :43-57 std::invalid_argument, others
// std::runtime_error). Header-only: a caller links liblpl_b200.so and nothing else.
#ifndef LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP
#define LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP

#include <cstdint>
#include <stdexcept>
#include <string>

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

// One context = one CUDA stream + device scratch for single frames of up to max_points points.
// Like the reference objects it is stateful and not thread-safe.
class Handle
{
  public:
    Handle() = default;
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    ~Handle() { reset(); }

    void reset()
    {
        if (ctx_ != nullptr)
        {
            lpl_destroy(ctx_);
            ctx_ = nullptr;
        }
    }

    // (re)creates the context when the capacity or the image size changes
    lpl_ctx* ensure(std::uint32_t max_points, std::int32_t image_height = 64, std::int32_t image_width = 2048,
                    int device = 0)
    {
        if (ctx_ == nullptr || max_points > max_points_ || image_height != height_ || image_width != width_)
        {
            reset();
            const std::uint32_t cap = max_points > max_points_ ? max_points : max_points_;
            const int rc = lpl_create(&ctx_, device, cap, 1U, image_height, image_width);
            if (rc != LPL_OK)
            {
                ctx_ = nullptr;
                raise(rc, nullptr, "lpl_create");
            }
            max_points_ = cap;
            height_ = image_height;
            width_ = image_width;
        }
        return ctx_;
    }

    lpl_ctx* get() const noexcept { return ctx_; }
    std::uint32_t capacity() const noexcept { return max_points_; }

  private:
    lpl_ctx* ctx_ = nullptr;
    std::uint32_t max_points_ = 0;
    std::int32_t height_ = 0;
    std::int32_t width_ = 0;
};
} // namespace detail
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__DETAIL__LPL_HANDLE_HPP
