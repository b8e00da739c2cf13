// Drop-in KDTree over the B200 C ABI (general nearest-neighbour queries).
// Replaces lidar_processing_lib/include/lidar_processing_lib/kdtree.hpp:40-400 of the reference: same class
// template, nested Neighbour / Compare types, rebuild / k_nearest / radius_search / radius_search_k_nearest /
// dist_sqr signatures. The searches run on the device (exact, exhaustive; lpl_knn_*); only KDTree<float, 3> - the
// one instantiation the reference library uses (noise_remover.hpp:59-60) - is provided.
// Differences a caller can observe: equal distances are ordered by point index (the reference: by its tree
// traversal); radius_search without `sort` returns ascending point indices; radius_search_k_nearest returns the k
// NEAREST points within the radius (the reference: the first k its traversal meets). Batched forms
// (k_nearest_batch / radius_search_batch) answer many targets in one device pass - one call per target costs a
// host <-> device round trip.
#ifndef LIDAR_PROCESSING_LIB__KDTREE_HPP
#define LIDAR_PROCESSING_LIB__KDTREE_HPP

#include <algorithm>
#include <array>
#include <cstdint>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "detail/lpl_handle.hpp"

namespace lidar_processing_lib
{
template <typename T, std::uint8_t Dim>
using Point = std::array<T, Dim>;

template <typename T, std::uint8_t Dim>
class KDTree final
{
    static_assert(std::is_same<T, float>::value && Dim == 3, "the device search is built for KDTree<float, 3>");

  public:
    using PointT = Point<T, Dim>;
    using KDTreeT = KDTree<T, Dim>;

    struct Neighbour final
    {
        std::uint32_t index;
        T distance;
    };

    struct Compare final
    {
        inline bool operator()(const Neighbour& a, const Neighbour& b) const noexcept { return a.distance < b.distance; }
    };

    KDTree& operator=(const KDTreeT& other) = delete;
    KDTree(const KDTreeT& other) = delete;
    KDTree& operator=(KDTreeT&& other) noexcept = default;
    KDTree(KDTreeT&& other) noexcept = default;

    KDTree(bool sort = false) : sort_(sort) {}

    void reserve(std::uint32_t num_pts = 200'000U) { reserve_ = num_pts; }

    void rebuild(const std::vector<PointT>& points)
    {
        points_ = points; // the reference copies the points into its nodes as well (kdtree.hpp:167-177)
        token_ = 0;
    }

    void k_nearest(const PointT& target, std::uint32_t num_neigh, std::vector<Neighbour>& neigh)
    {
        neigh.clear();
        if (num_neigh == 0 || points_.empty())
        {
            return;
        }
        std::vector<std::vector<Neighbour>> all;
        k_nearest_batch({target}, num_neigh, all);
        neigh = std::move(all[0]);
    }

    void radius_search(const PointT& target, T proximity_sqr, std::vector<Neighbour>& neigh)
    {
        neigh.clear();
        if (points_.empty())
        {
            return;
        }
        std::vector<std::vector<Neighbour>> all;
        radius_search_batch({target}, {proximity_sqr}, all);
        neigh = std::move(all[0]);
    }

    void radius_search_k_nearest(const PointT& target, T proximity_sqr, std::uint32_t num_neigh, std::vector<Neighbour>& neigh)
    {
        neigh.clear();
        if (points_.empty() || num_neigh == 0)
        {
            return;
        }
        const std::uint32_t k = clamp_k(num_neigh);
        std::vector<std::uint32_t> idx(k), cnt(1);
        std::vector<float> dist(k);
        lpl_ctx* ctx = resident();
        detail::check(lpl_knn_k_nearest(ctx, target.data(), sizeof(PointT), 1, k, &proximity_sqr, idx.data(), dist.data(), cnt.data()),
                      ctx, "KDTree::radius_search_k_nearest");
        for (std::uint32_t j = 0; j < cnt[0]; ++j)
        {
            neigh.push_back({idx[j], dist[j]});
        }
    }

    /// Batched k_nearest: neigh[q] = the num_neigh nearest points of targets[q], ascending distance.
    void k_nearest_batch(const std::vector<PointT>& targets, std::uint32_t num_neigh, std::vector<std::vector<Neighbour>>& neigh)
    {
        neigh.assign(targets.size(), {});
        if (targets.empty() || num_neigh == 0 || points_.empty())
        {
            return;
        }
        const std::uint32_t k = clamp_k(num_neigh);
        const auto m = static_cast<std::uint32_t>(targets.size());
        std::vector<std::uint32_t> idx(static_cast<std::size_t>(m) * k), cnt(m);
        std::vector<float> dist(static_cast<std::size_t>(m) * k);
        lpl_ctx* ctx = resident();
        detail::check(lpl_knn_k_nearest(ctx, targets.data(), sizeof(PointT), m, k, nullptr, idx.data(), dist.data(), cnt.data()), ctx,
                      "KDTree::k_nearest");
        for (std::uint32_t q = 0; q < m; ++q)
        {
            neigh[q].reserve(cnt[q]);
            for (std::uint32_t j = 0; j < cnt[q]; ++j)
            {
                neigh[q].push_back({idx[static_cast<std::size_t>(q) * k + j], dist[static_cast<std::size_t>(q) * k + j]});
            }
        }
    }

    /// Batched radius_search: neigh[q] = every point with dist_sqr <= proximity_sqr[q].
    void radius_search_batch(const std::vector<PointT>& targets, const std::vector<T>& proximity_sqr,
                             std::vector<std::vector<Neighbour>>& neigh)
    {
        if (proximity_sqr.size() != targets.size())
        {
            throw std::invalid_argument("KDTree::radius_search_batch: one radius per target");
        }
        neigh.assign(targets.size(), {});
        if (targets.empty() || points_.empty())
        {
            return;
        }
        const auto m = static_cast<std::uint32_t>(targets.size());
        lpl_ctx* ctx = resident();
        std::uint32_t cap = 64;
        std::vector<std::uint32_t> idx, cnt(m);
        std::vector<float> dist;
        for (;;)
        {
            idx.resize(static_cast<std::size_t>(m) * cap);
            dist.resize(static_cast<std::size_t>(m) * cap);
            detail::check(lpl_knn_radius_search(ctx, targets.data(), sizeof(PointT), m, proximity_sqr.data(), cap, idx.data(),
                                                dist.data(), cnt.data()),
                          ctx, "KDTree::radius_search");
            const std::uint32_t most = *std::max_element(cnt.begin(), cnt.end());
            if (most <= cap)
            {
                break;
            }
            cap = most; // a second pass with room for the fullest neighbourhood
        }
        for (std::uint32_t q = 0; q < m; ++q)
        {
            neigh[q].reserve(cnt[q]);
            for (std::uint32_t j = 0; j < cnt[q]; ++j)
            {
                neigh[q].push_back({idx[static_cast<std::size_t>(q) * cap + j], dist[static_cast<std::size_t>(q) * cap + j]});
            }
            if (sort_)
            {
                std::stable_sort(neigh[q].begin(), neigh[q].end(), Compare{});
            }
        }
    }

    constexpr T dist_sqr(const PointT& a, const PointT& b) noexcept
    {
        return (a[0] - b[0]) * (a[0] - b[0]) + ((a[1] - b[1]) * (a[1] - b[1]) + ((a[2] - b[2]) * (a[2] - b[2]) + 0));
    }

  private:
    std::uint32_t clamp_k(std::uint32_t k) const
    {
        const auto n = static_cast<std::uint32_t>(points_.size());
        k = k < n ? k : n;
        if (k > 128U)
        {
            throw std::overflow_error("KDTree::k_nearest: at most 128 neighbours per query on the device");
        }
        return k;
    }

    // the thread's context with THIS tree's points resident (another object or another adaptor call of the thread
    // may have replaced them since)
    lpl_ctx* resident()
    {
        const auto n = static_cast<std::uint32_t>(points_.size());
        lpl_ctx* ctx = handle_.ensure(n > reserve_ ? n : reserve_);
        if (ctx != built_on_ || token_ == 0 || lpl_knn_token(ctx) != token_)
        {
            detail::check(lpl_knn_build(ctx, points_.data(), sizeof(PointT), n), ctx, "KDTree::rebuild");
            token_ = lpl_knn_token(ctx);
            built_on_ = ctx;
        }
        return ctx;
    }

    bool sort_;
    std::uint32_t reserve_ = 200'000U;
    std::vector<PointT> points_;
    unsigned long long token_ = 0;
    lpl_ctx* built_on_ = nullptr;
    detail::Handle handle_;
};
} // namespace lidar_processing_lib

#endif // LIDAR_PROCESSING_LIB__KDTREE_HPP
