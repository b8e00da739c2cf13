// General nearest-neighbour queries: the public KDTree<float, 3> API of the
// reference library - k_nearest, radius_search, radius_search_k_nearest
// (lidar_processing_lib/include/lidar_processing_lib/kdtree.hpp:216-400) - as BATCHED exact searches on the device.
//
// The node itself never calls these (the DROR stage has its own grid search, ring_dror.cu); they are offered for
// source compatibility through include/lidar_processing_lib/kdtree.hpp. A query set is answered by an exhaustive
// tiled scan: the point set streams through shared memory in tiles, every thread owns one query and keeps its
// current k best in a sorted list (insertions become rare after the first tiles). Exact by construction, with the
// reference's distance expression (a0-b0)^2 + ((a1-b1)^2 + ((a2-b2)^2 + 0)), a = target, b = tree point
// (kdtree.hpp:131-143), no FMA. O(n m) work - 15 G pair tests per second-scale for 120k x 120k - so it is a
// utility, not a hot-path stage.
// Order of results: ascending (distance, point index). The reference returns k_nearest in ascending distance and
// leaves ties / radius_search order to its tree traversal; (distance, index) is one valid such order.
#include "common.cuh"

namespace lpl
{
constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;

__device__ __forceinline__ float knn_dist(const float4& t, const float4& p)
{
    const float d0 = t.x - p.x, d1 = t.y - p.y, d2 = t.z - p.z;
    return d0 * d0 + (d1 * d1 + (d2 * d2 + 0.0f));
}

// k nearest of every query, optionally restricted to dist <= radius_sqr[q]. best_d / best_i: [m][k], sorted
// ascending by (distance, index); count[q] = neighbours found (min(k, points within the radius)).
template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
    k_knn(const float4* __restrict__ pts, std::uint32_t n, const float4* __restrict__ queries, std::uint32_t m, std::uint32_t k,
          const float* __restrict__ radius_sqr, float radius_all, float* __restrict__ best_d, std::uint32_t* __restrict__ best_i,
          std::uint32_t* __restrict__ count)
{
    __shared__ float4 tile[kKnnTile];
    const std::uint32_t q = blockIdx.x * kKnnThreads + threadIdx.x;
    const bool live = q < m;
    const float4 t = live ? queries[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float r2 = radius_sqr != nullptr ? (live ? radius_sqr[q] : 0.f) : radius_all;
    float bd[KMAX];
    std::uint32_t bi[KMAX];
    std::uint32_t have = 0;
    for (std::uint32_t base = 0; base < n; base += kKnnTile)
    {
        __syncthreads();
        for (std::uint32_t j = threadIdx.x; j < kKnnTile; j += kKnnThreads)
        {
            tile[j] = base + j < n ? pts[base + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const std::uint32_t cnt = min(static_cast<std::uint32_t>(kKnnTile), n - base);
        if (!live)
        {
            continue;
        }
        for (std::uint32_t j = 0; j < cnt; ++j)
        {
            const float d = knn_dist(t, tile[j]);
            // points arrive in ascending index order: a tie with the current worst keeps the earlier point
            if (!(d <= r2) || (have == k && !(d < bd[k - 1])))
            {
                continue;
            }
            std::uint32_t pos = have < k ? have : k - 1;
            while (pos > 0 && d < bd[pos - 1])
            {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = base + j;
            have = have < k ? have + 1 : have;
        }
    }
    if (live)
    {
        for (std::uint32_t j = 0; j < have; ++j)
        {
            best_d[static_cast<std::size_t>(q) * k + j] = bd[j];
            best_i[static_cast<std::size_t>(q) * k + j] = bi[j];
        }
        count[q] = have;
    }
}

// every point within the radius of every query: count[q] = how many there are; the first `cap` of them (ascending
// point index) go to out_i / out_d [m][cap]
__global__ void __launch_bounds__(kKnnThreads)
    k_radius(const float4* __restrict__ pts, std::uint32_t n, const float4* __restrict__ queries, std::uint32_t m,
             const float* __restrict__ radius_sqr, float radius_all, std::uint32_t cap, float* __restrict__ out_d,
             std::uint32_t* __restrict__ out_i, std::uint32_t* __restrict__ count)
{
    __shared__ float4 tile[kKnnTile];
    const std::uint32_t q = blockIdx.x * kKnnThreads + threadIdx.x;
    const bool live = q < m;
    const float4 t = live ? queries[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float r2 = radius_sqr != nullptr ? (live ? radius_sqr[q] : 0.f) : radius_all;
    std::uint32_t have = 0;
    for (std::uint32_t base = 0; base < n; base += kKnnTile)
    {
        __syncthreads();
        for (std::uint32_t j = threadIdx.x; j < kKnnTile; j += kKnnThreads)
        {
            tile[j] = base + j < n ? pts[base + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const std::uint32_t cnt = min(static_cast<std::uint32_t>(kKnnTile), n - base);
        if (!live)
        {
            continue;
        }
        for (std::uint32_t j = 0; j < cnt; ++j)
        {
            const float d = knn_dist(t, tile[j]);
            if (d <= r2)
            {
                if (have < cap)
                {
                    out_d[static_cast<std::size_t>(q) * cap + have] = d;
                    out_i[static_cast<std::size_t>(q) * cap + have] = base + j;
                }
                ++have;
            }
        }
    }
    if (live)
    {
        count[q] = have;
    }
}

int launch_knn(Ctx* c, const float4* pts, std::uint32_t n, const float4* queries, std::uint32_t m, std::uint32_t k,
               const float* radius_sqr, float radius_all, float* best_d, std::uint32_t* best_i, std::uint32_t* count)
{
    const dim3 grid((m + kKnnThreads - 1) / kKnnThreads);
    if (k <= 8)
    {
        k_knn<8><<<grid, kKnnThreads, 0, c->stream>>>(pts, n, queries, m, k, radius_sqr, radius_all, best_d, best_i, count);
    }
    else if (k <= 32)
    {
        k_knn<32><<<grid, kKnnThreads, 0, c->stream>>>(pts, n, queries, m, k, radius_sqr, radius_all, best_d, best_i, count);
    }
    else if (k <= 128)
    {
        k_knn<128><<<grid, kKnnThreads, 0, c->stream>>>(pts, n, queries, m, k, radius_sqr, radius_all, best_d, best_i, count);
    }
    else
    {
        return -1;
    }
    mark(c, "knn");
    return 0;
}

void launch_radius(Ctx* c, const float4* pts, std::uint32_t n, const float4* queries, std::uint32_t m, const float* radius_sqr,
                   float radius_all, std::uint32_t cap, float* out_d, std::uint32_t* out_i, std::uint32_t* count)
{
    k_radius<<<dim3((m + kKnnThreads - 1) / kKnnThreads), kKnnThreads, 0, c->stream>>>(pts, n, queries, m, radius_sqr, radius_all,
                                                                                       cap, out_d, out_i, count);
    mark(c, "radius_search");
}
} // namespace lpl
