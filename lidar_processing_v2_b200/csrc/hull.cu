// Stage 4: per-cluster gather and Andrew monotone-chain convex hulls, batched over frames.
//
// Reference:
//   cluster gather  src/processor/src/processor.cpp:627-658 (O(K*M) rescan per label; here one
//                   counting-sort pass over the obstacle cloud) incl. z_min / z_max
//   convexHull      lidar_processing_lib/src/polygonizer.cpp:33-91 on PointXY{double x, y}
//
// Every cluster is sorted by (x, y) on order-preserving 64-bit keys and then swept by the same
// sequential lower/upper chain as the reference, with the same fp64 orientation predicate
// evaluated without FMA contraction, so the vertex list (coordinates) is identical; only the
// *index* reported for exactly duplicated (x, y) points may differ, as it does between
// std::sort implementations. Clusters of up to kHullSmem points are sorted and swept in shared
// memory by one CTA; larger ones use the same network on their global-memory segment.
#include "common.cuh"

namespace lpl
{
constexpr int kHullSmem = 2048;
constexpr int kHullThreads = 128;
constexpr int kHullBlocks = 592; // 4 CTAs per SM x 148 SMs, grid-stride over clusters

__device__ __forceinline__ std::uint32_t ord_f32(float v)
{
    v = v + 0.0f; // -0.0 -> +0.0 (the reference comparator treats them as equal)
    const std::uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float unord_f32(std::uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256) k_hull_scatter(Dev d)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_o[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::int32_t l = d.clabel[o + i];
    if (l < 0)
    {
        return;
    }
    // counting down returns ccount to zero; the segment size stays available from cstart
    const std::uint32_t k = atomicSub(&d.ccount[o + l], 1u) - 1u;
    const std::uint32_t pos = d.cstart[static_cast<std::size_t>(f) * (d.cap + 1) + l] + k;
    const float4 p = d.pts_o[o + i];
    d.hsk[o + pos] = (static_cast<unsigned long long>(ord_f32(p.x)) << 32) | ord_f32(p.y);
    d.hsi[o + pos] = i;
}

// CTA-wide ascending sort of (key, value) pairs for arbitrary n. Mirror-first bitonic network:
// every exchange moves the larger key to the higher index, so the virtual +inf padding beyond
// n never moves and no physical padding is needed.
__device__ __forceinline__ void block_sort_pairs(unsigned long long* keys, std::uint32_t* vals, std::uint32_t n)
{
    for (std::uint32_t k = 2; (k >> 1) < n; k <<= 1)
    {
        for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
        {
            const std::uint32_t u = t ^ (k - 1);
            if (u > t && u < n)
            {
                const unsigned long long a = keys[t], b = keys[u];
                if (b < a)
                {
                    keys[t] = b;
                    keys[u] = a;
                    const std::uint32_t va = vals[t];
                    vals[t] = vals[u];
                    vals[u] = va;
                }
            }
        }
        __syncthreads();
        for (std::uint32_t j = k >> 2; j > 0; j >>= 1)
        {
            for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
            {
                const std::uint32_t u = t ^ j;
                if (u > t && u < n)
                {
                    const unsigned long long a = keys[t], b = keys[u];
                    if (b < a)
                    {
                        keys[t] = b;
                        keys[u] = a;
                        const std::uint32_t va = vals[t];
                        vals[t] = vals[u];
                        vals[u] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// polygonizer.cpp:45-48: true when p3 is not strictly left of p1 -> p2 (pop p2)
__device__ __forceinline__ bool not_left(unsigned long long k1, unsigned long long k2, unsigned long long k3)
{
    const double x1 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k1 >> 32)));
    const double y1 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k1)));
    const double x2 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k2 >> 32)));
    const double y2 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k2)));
    const double x3 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k3 >> 32)));
    const double y3 = static_cast<double>(unord_f32(static_cast<std::uint32_t>(k3)));
    return (x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1) <= 0.0;
}

// sequential monotone chain over sorted keys; st needs n + 1 entries; returns the vertex count
__device__ std::uint32_t monotone_chain(const unsigned long long* keys, std::uint32_t n, std::uint32_t* st)
{
    std::int32_t k = 0;
    for (std::int32_t i = 0; i < static_cast<std::int32_t>(n); ++i)
    {
        const unsigned long long ki = keys[i];
        while (k > 1 && not_left(keys[st[k - 2]], keys[st[k - 1]], ki))
        {
            --k;
        }
        st[k++] = static_cast<std::uint32_t>(i);
    }
    for (std::int32_t i = static_cast<std::int32_t>(n) - 2, t = k + 1; i >= 0; --i)
    {
        const unsigned long long ki = keys[i];
        while (k >= t && not_left(keys[st[k - 2]], keys[st[k - 1]], ki))
        {
            --k;
        }
        st[k++] = static_cast<std::uint32_t>(i);
    }
    return static_cast<std::uint32_t>(k - 1);
}

__global__ void __launch_bounds__(kHullThreads) k_hull(Dev d)
{
    __shared__ unsigned long long s_keys[kHullSmem];
    __shared__ std::uint32_t s_vals[kHullSmem];
    __shared__ std::uint32_t s_stack[kHullSmem + 1];
    __shared__ float s_red[2][kHullThreads / 32];
    __shared__ std::uint32_t s_cnt;
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    for (std::uint32_t c = blockIdx.x; c < K; c += gridDim.x)
    {
        const std::uint32_t seg = cstart[c];
        const std::uint32_t n = cstart[c + 1] - seg;
        unsigned long long* gk = d.hsk + o + seg;
        std::uint32_t* gv = d.hsi + o + seg;
        std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + seg + c; // n + 1 entries
        // z extent of the cluster
        float zmn = 3.402823466e+38f, zmx = -3.402823466e+38f;
        for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
        {
            const float z = d.pts_o[o + gv[t]].z;
            zmn = fminf(zmn, z);
            zmx = fmaxf(zmx, z);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1)
        {
            zmn = fminf(zmn, __shfl_xor_sync(0xffffffffu, zmn, s));
            zmx = fmaxf(zmx, __shfl_xor_sync(0xffffffffu, zmx, s));
        }
        if (lane_id() == 0)
        {
            s_red[0][threadIdx.x >> 5] = zmn;
            s_red[1][threadIdx.x >> 5] = zmx;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            for (int w = 1; w < kHullThreads / 32; ++w)
            {
                zmn = fminf(zmn, s_red[0][w]);
                zmx = fmaxf(zmx, s_red[1][w]);
            }
            d.zminmax[o + c] = make_float2(zmn, zmx);
        }
        if (n < 3)
        {
            // identity order = obstacle-cloud order (polygonizer.cpp:36-41)
            if (threadIdx.x == 0)
            {
                std::uint32_t a = (n > 0) ? gv[0] : 0u, b = (n > 1) ? gv[1] : 0u;
                if (n == 2 && b < a)
                {
                    const std::uint32_t t = a;
                    a = b;
                    b = t;
                }
                if (n > 0)
                {
                    gst[0] = a;
                }
                if (n > 1)
                {
                    gst[1] = b;
                }
                d.hcnt[o + c] = n;
            }
            __syncthreads();
            continue;
        }
        if (n <= kHullSmem)
        {
            for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
            {
                s_keys[t] = gk[t];
                s_vals[t] = gv[t];
            }
            __syncthreads();
            block_sort_pairs(s_keys, s_vals, n);
            if (threadIdx.x == 0)
            {
                s_cnt = monotone_chain(s_keys, n, s_stack);
            }
            __syncthreads();
            const std::uint32_t hc = s_cnt;
            for (std::uint32_t t = threadIdx.x; t < hc; t += blockDim.x)
            {
                gst[t] = s_vals[s_stack[t]];
            }
            if (threadIdx.x == 0)
            {
                d.hcnt[o + c] = hc;
            }
        }
        else
        {
            __syncthreads();
            block_sort_pairs(gk, gv, n);
            if (threadIdx.x == 0)
            {
                const std::uint32_t hc = monotone_chain(gk, n, gst);
                s_cnt = hc;
                d.hcnt[o + c] = hc;
            }
            __syncthreads();
            const std::uint32_t hc = s_cnt;
            for (std::uint32_t t = threadIdx.x; t < hc; t += blockDim.x)
            {
                gst[t] = gv[gst[t]]; // sorted position -> obstacle-cloud index
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kHullThreads) k_hull_gather(Dev d)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    for (std::uint32_t c = blockIdx.x; c < K; c += gridDim.x)
    {
        const std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + cstart[c] + c;
        const std::uint32_t off = hoff[c], hc = hoff[c + 1] - off;
        for (std::uint32_t t = threadIdx.x; t < hc; t += blockDim.x)
        {
            const std::uint32_t idx = gst[t];
            const float4 p = d.pts_o[o + idx];
            d.hull_idx[o + off + t] = idx;
            d.hull_xy[o + off + t] = make_float2(p.x, p.y);
        }
    }
}

void launch_hulls(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    k_excl_scan<<<nf, 1024, 0, s>>>(d.ccount, d.cap, d.cstart, d.cap + 1, d.cap, d.n_clusters, nullptr);
    mark(c, "hull_seg_scan");
    k_hull_scatter<<<dim3((d.cap + 255) / 256, nf), 256, 0, s>>>(d);
    mark(c, "hull_scatter");
    k_hull<<<dim3(kHullBlocks, nf), kHullThreads, 0, s>>>(d);
    mark(c, "hull");
    k_excl_scan<<<nf, 1024, 0, s>>>(d.hcnt, d.cap, d.hull_off, d.cap + 1, d.cap, d.n_clusters, d.n_hull);
    mark(c, "hull_off_scan");
    k_hull_gather<<<dim3(kHullBlocks, nf), kHullThreads, 0, s>>>(d);
    mark(c, "hull_gather");
}
} // namespace lpl
