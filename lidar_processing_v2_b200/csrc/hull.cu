// Stage 4: per-cluster gather and Andrew monotone-chain convex hulls, batched over frames.
//
// Reference:
//   cluster gather  src/processor/src/processor.cpp:627-658 (O(K*M) rescan per label) incl.
//                   z_min / z_max
//   convexHull      lidar_processing_lib/src/polygonizer.cpp:33-91 on PointXY{double x, y}
//
// One frame-wide merge sort on the key (cluster label, x, y, point index) replaces both the
// per-label gather and the per-cluster std::sort: after it every cluster is a contiguous,
// (x, y)-sorted segment (k_hull_tilesort: 2048-element bitonic tiles in shared memory;
// k_hull_merge: merge-path passes, each output tile merged in shared memory).
// The segments are then cut into chunks and one warp per chunk thins and sweeps it in shared memory
// (k_hull_chunks, k_hull_join - see "Shared-memory hull pass" below) with the reference's own lower / upper
// chain and its fp64 orientation predicate evaluated without FMA contraction. The vertex list
// (coordinates) equals the reference's; only the *index* reported for exactly duplicated (x, y)
// points may differ, as it does between std::sort implementations.
#include "common.cuh"

namespace lpl
{
// element = (label, x bits, y bits, obstacle-cloud index) with -0.0 folded into +0.0 (the
// reference comparator treats them as equal); the index makes the order total
__device__ __forceinline__ bool elem_less(const uint4& a, const uint4& b)
{
    if (a.x != b.x)
    {
        return a.x < b.x;
    }
    const float ax = __uint_as_float(a.y), bx = __uint_as_float(b.y);
    if (ax != bx)
    {
        return ax < bx;
    }
    const float ay = __uint_as_float(a.z), by = __uint_as_float(b.z);
    if (ay != by)
    {
        return ay < by;
    }
    return a.w < b.w;
}

// ------------------------------------------------------------------------------------------
// polygon filter (Akl-Toussaint): a point strictly inside the polygon spanned by up to kExtDirs
// extreme points of its cluster cannot be a hull vertex, so it never enters the sort. The
// extreme points are actual points of the cluster (found while labelling, cluster.cu), the test
// is the reference's own fp64 orientation predicate, and a polygon that is not convex in
// counter-clockwise order (possible only through float rounding of x + y / x - y) disables the
// filter for that cluster.
// ------------------------------------------------------------------------------------------
struct P2
{
    double x, y;
};

// polygonizer.cpp:45-48: true when p3 is not strictly left of p1 -> p2 (pop p2)
__device__ __forceinline__ bool not_left(const P2& p1, const P2& p2, const P2& p3)
{
    return (p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x) <= 0.0;
}

constexpr std::uint32_t kOctaMinPoints = 16; // smaller clusters skip the filter

__global__ void __launch_bounds__(128) k_hull_octagon(Dev d)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    for (std::uint32_t c = blockIdx.x * 128u + threadIdx.x; c < K; c += gridDim.x * 128u)
    {
        float2 v[kExtDirs];
        bool ok = d.ccount[o + c] >= kOctaMinPoints;
        // a slot still in its initial state (possible when only a subset of the points contributes,
        // cluster.cu) leaves the cluster without a filter polygon
        ok = ok && d.ext[(o + c) * kExtDirs] != 0ULL;
        if (ok)
        {
#pragma unroll
            for (int k = 0; k < kExtDirs; ++k)
            {
                const std::uint32_t idx = static_cast<std::uint32_t>(d.ext[(o + c) * kExtDirs + k]);
                const float4 p = d.pts_o[o + idx];
                v[k] = make_float2(p.x + 0.0f, p.y + 0.0f);
            }
            // convex and counter-clockwise (repeated vertices allowed), and not degenerate
            bool any_turn = false;
#pragma unroll
            for (int k = 0; k < kExtDirs; ++k)
            {
                const P2 a = {static_cast<double>(v[k].x), static_cast<double>(v[k].y)};
                const P2 b = {static_cast<double>(v[(k + 1) % kExtDirs].x), static_cast<double>(v[(k + 1) % kExtDirs].y)};
#pragma unroll
                for (int j = 2; j < kExtDirs; ++j)
                {
                    const P2 q = {static_cast<double>(v[(k + j) % kExtDirs].x), static_cast<double>(v[(k + j) % kExtDirs].y)};
                    const double cr = (b.x - a.x) * (q.y - a.y) - (b.y - a.y) * (q.x - a.x);
                    ok = ok && !(cr < 0.0); // every other vertex on or left of every edge
                    any_turn = any_turn || cr > 0.0;
                }
            }
            ok = ok && any_turn;
        }
#pragma unroll
        for (int k = 0; k < kExtDirs; ++k)
        {
            d.octa[(o + c) * kExtDirs + k] = ok ? v[k] : make_float2(__int_as_float(0x7fc00000), 0.f);
        }
        d.hseg_cnt[o + c] = 0;
    }
}

struct HullKeepPred
{
    Dev d;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        const std::size_t o = static_cast<std::size_t>(f) * d.cap;
        const std::int32_t l = d.clabel[o + i];
        if (l < 0)
        {
            return false;
        }
        const float2* v = d.octa + (o + l) * kExtDirs;
        const float2 v0 = v[0];
        if (v0.x != v0.x)
        {
            return true; // no usable polygon for this cluster
        }
        const float4 pt = d.pts_o[o + i];
        const P2 q = {static_cast<double>(pt.x), static_cast<double>(pt.y)};
        P2 a = {static_cast<double>(v0.x), static_cast<double>(v0.y)};
        bool inside = true;
#pragma unroll
        for (int k = 1; k <= kExtDirs; ++k)
        {
            const float2 vk = v[k % kExtDirs];
            const P2 b = {static_cast<double>(vk.x), static_cast<double>(vk.y)};
            const bool degenerate = (a.x == b.x) && (a.y == b.y);
            inside = inside && (degenerate || !not_left(a, b, q)); // strictly left of every proper edge
            a = b;
        }
        return !inside;
    }
};

struct HullKeepEmit
{
    Dev d;
    __device__ void operator()(std::uint32_t f, std::uint32_t i, std::uint32_t pos) const
    {
        const std::size_t o = static_cast<std::size_t>(f) * d.cap;
        const float4 p = d.pts_o[o + i];
        const std::uint32_t l = static_cast<std::uint32_t>(d.clabel[o + i]);
        d.hsB[o + pos] = make_uint4(l, __float_as_uint(p.x + 0.0f), __float_as_uint(p.y + 0.0f), i);
        atomicAdd(&d.hseg_cnt[o + l], 1u);
    }
};

// ------------------------------------------------------------------------------------------
// tile sort: 2048 elements per CTA, mirror-first bitonic network for arbitrary n (every exchange
// moves the larger key to the higher index, so the virtual +inf padding beyond n never moves)
// ------------------------------------------------------------------------------------------
// kThreads: 256 for a batch that fills the GPU with tiles; 1024 for a launch of a few frames, whose handful of tiles are
// pure latency (four compare-exchanges per thread and stage become one: single frame 0.073 -> 0.039 ms)
template <int kThreads>
__global__ void __launch_bounds__(kThreads) k_hull_tilesort(Dev d)
{
    __shared__ uint4 s[kTile];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_h[f];
    const std::uint32_t base = blockIdx.x * kTile;
    if (base >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t m = min(static_cast<std::uint32_t>(kTile), n - base);
    for (std::uint32_t t = threadIdx.x; t < m; t += kThreads)
    {
        s[t] = d.hsB[o + base + t]; // (label, x, y, index) of the points that survived the octagon filter
    }
    __syncthreads();
    // every thread owns compare-exchange PAIRS (q-th pair of a stage: insert a zero bit at the stride
    // position), so no lane idles on the upper element of a pair
    for (std::uint32_t k = 2; (k >> 1) < m; k <<= 1)
    {
        const std::uint32_t hk = k >> 1;
        for (std::uint32_t q = threadIdx.x; q < kTile / 2; q += kThreads)
        {
            const std::uint32_t t = ((q & ~(hk - 1u)) << 1) | (q & (hk - 1u));
            const std::uint32_t u = t ^ (k - 1u);
            if (t >= m)
            {
                break; // t grows with q
            }
            if (u < m)
            {
                const uint4 a = s[t], b = s[u];
                if (elem_less(b, a))
                {
                    s[t] = b;
                    s[u] = a;
                }
            }
        }
        __syncthreads();
        for (std::uint32_t j = k >> 2; j > 0; j >>= 1)
        {
            for (std::uint32_t q = threadIdx.x; q < kTile / 2; q += kThreads)
            {
                const std::uint32_t t = ((q & ~(j - 1u)) << 1) | (q & (j - 1u));
                const std::uint32_t u = t | j;
                if (t >= m)
                {
                    break;
                }
                if (u < m)
                {
                    const uint4 a = s[t], b = s[u];
                    if (elem_less(b, a))
                    {
                        s[t] = b;
                        s[u] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (std::uint32_t t = threadIdx.x; t < m; t += kThreads)
    {
        d.hsA[o + base + t] = s[t];
    }
}

// number of elements taken from A among the first `diag` outputs of merge(A, B)
template <class GetA, class GetB>
__device__ __forceinline__ std::uint32_t merge_path(GetA A, std::uint32_t la, GetB B, std::uint32_t lb, std::uint32_t diag)
{
    std::uint32_t lo = diag > lb ? diag - lb : 0u;
    std::uint32_t hi = min(diag, la);
    while (lo < hi)
    {
        const std::uint32_t mid = (lo + hi) >> 1;
        if (elem_less(A(mid), B(diag - 1u - mid)))
        {
            lo = mid + 1u;
        }
        else
        {
            hi = mid;
        }
    }
    return lo;
}

// the same split found by a whole warp: 32 probe positions per round (the predicate is true on a prefix
// of the range), so a run of 131,072 elements takes four dependent rounds of loads instead of seventeen
template <class GetA, class GetB>
__device__ __forceinline__ std::uint32_t merge_path_warp(GetA A, std::uint32_t la, GetB B, std::uint32_t lb, std::uint32_t diag)
{
    std::uint32_t lo = diag > lb ? diag - lb : 0u;
    std::uint32_t hi = min(diag, la);
    const std::uint32_t lane = lane_id();
    while (lo < hi) // warp-uniform
    {
        const unsigned long long span = hi - lo;
        const std::uint32_t m = lo + static_cast<std::uint32_t>((span * lane) >> 5);
        const bool before = elem_less(A(m), B(diag - 1u - m));
        const std::uint32_t cnt = __popc(__ballot_sync(0xffffffffu, before));
        const std::uint32_t new_hi = cnt < 32u ? lo + static_cast<std::uint32_t>((span * cnt) >> 5) : hi;
        const std::uint32_t new_lo = cnt > 0u ? lo + static_cast<std::uint32_t>((span * (cnt - 1u)) >> 5) + 1u : lo;
        lo = new_lo;
        hi = new_hi;
    }
    return lo;
}

constexpr std::uint32_t kMergeCtas = 64;

// merge pass p: runs of (kTile << p) elements, pairwise, one output tile per CTA and trip
__global__ void __launch_bounds__(kTileThreads) k_hull_merge(Dev d, std::uint32_t pass)
{
    __shared__ uint4 s[kTile];
    __shared__ std::uint32_t s_split[2];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_h[f];
    // at most kMergeCtas CTAs per frame stride over the output tiles: a 2 M-point capacity means 977 tiles and ten
    // passes, nearly all of them empty
    for (std::uint32_t tile = blockIdx.x;; tile += gridDim.x)
    {
    const std::uint32_t out0 = tile * kTile;
    if (out0 >= n || sort_passes(n) <= pass)
    {
        return; // past the end, or this frame was fully sorted by an earlier pass
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const uint4* src = ((pass & 1u) ? d.hsB : d.hsA) + o;
    uint4* dst = ((pass & 1u) ? d.hsA : d.hsB) + o;
    const std::uint32_t run = static_cast<std::uint32_t>(kTile) << pass;
    const std::uint32_t pair0 = (out0 / (2u * run)) * (2u * run);
    const std::uint32_t a0 = pair0, a1 = min(n, a0 + run);
    const std::uint32_t b0 = a1, b1 = min(n, b0 + run);
    const std::uint32_t la = a1 - a0, lb = b1 - b0;
    const std::uint32_t d0 = out0 - pair0;
    const std::uint32_t d1 = min(d0 + static_cast<std::uint32_t>(kTile), la + lb);
    const std::uint32_t cnt = d1 - d0;
    if (lb == 0)
    {
        for (std::uint32_t t = threadIdx.x; t < cnt; t += kTileThreads)
        {
            dst[out0 + t] = src[out0 + t];
        }
        continue;
    }
    if (threadIdx.x < 64u)
    {
        // warp 0 finds the split of the tile's first output, warp 1 of its last
        const std::uint32_t dg = threadIdx.x < 32u ? d0 : d1;
        const std::uint32_t split = merge_path_warp([&](std::uint32_t i) { return src[a0 + i]; }, la,
                                                    [&](std::uint32_t i) { return src[b0 + i]; }, lb, dg);
        if (lane_id() == 0)
        {
            s_split[threadIdx.x >> 5] = split;
        }
    }
    __syncthreads();
    const std::uint32_t ai0 = s_split[0], ai1 = s_split[1];
    const std::uint32_t bi0 = d0 - ai0, bi1 = d1 - ai1;
    const std::uint32_t na = ai1 - ai0, nb = bi1 - bi0; // na + nb == cnt <= kTile
    for (std::uint32_t t = threadIdx.x; t < na; t += kTileThreads)
    {
        s[t] = src[a0 + ai0 + t];
    }
    for (std::uint32_t t = threadIdx.x; t < nb; t += kTileThreads)
    {
        s[na + t] = src[b0 + bi0 + t];
    }
    __syncthreads();
    // each thread merges kItems consecutive outputs
    const std::uint32_t dg = min(threadIdx.x * static_cast<std::uint32_t>(kItems), cnt);
    std::uint32_t ia = merge_path([&](std::uint32_t i) { return s[i]; }, na,
                                  [&](std::uint32_t i) { return s[na + i]; }, nb, dg);
    std::uint32_t ib = dg - ia;
    uint4 out[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k)
    {
        const bool has_a = ia < na, has_b = ib < nb;
        uint4 va = make_uint4(0, 0, 0, 0), vb = va;
        if (has_a)
        {
            va = s[ia];
        }
        if (has_b)
        {
            vb = s[na + ib];
        }
        const bool take_a = has_a && (!has_b || elem_less(va, vb));
        out[k] = take_a ? va : vb;
        ia += take_a ? 1u : 0u;
        ib += take_a ? 0u : 1u;
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k)
    {
        if (dg + k < cnt)
        {
            dst[out0 + dg + k] = out[k];
        }
    }
    __syncthreads(); // s / s_split are rewritten by the next trip
    }
}

// ------------------------------------------------------------------------------------------
// hull chains
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ P2 elem_pt(const uint4& e)
{
    P2 p;
    p.x = static_cast<double>(__uint_as_float(e.y));
    p.y = static_cast<double>(__uint_as_float(e.z));
    return p;
}

// the reference's sweep (polygonizer.cpp:67-90) over m >= 1 sorted points; st needs m + 1 entries.
// The two topmost stack points stay in registers; only a pop reads the stack and a point again.
template <class Get, class Stack>
__device__ __forceinline__ std::uint32_t monotone_chain(Get P, std::uint32_t m, Stack* st)
{
    std::int32_t k = 0;
    P2 s2 = {0.0, 0.0}, s1 = {0.0, 0.0};
    for (std::int32_t i = 0; i < static_cast<std::int32_t>(m); ++i)
    {
        const P2 pi = P(i);
        while (k > 1 && not_left(s2, s1, pi))
        {
            --k;
            s1 = s2;
            if (k > 1)
            {
                s2 = P(st[k - 2]);
            }
        }
        st[k++] = static_cast<Stack>(i);
        s2 = s1;
        s1 = pi;
    }
    for (std::int32_t i = static_cast<std::int32_t>(m) - 2, t = k + 1; i >= 0; --i)
    {
        const P2 pi = P(i);
        while (k >= t && not_left(s2, s1, pi))
        {
            --k;
            s1 = s2;
            if (k > 1)
            {
                s2 = P(st[k - 2]);
            }
        }
        st[k++] = static_cast<Stack>(i);
        s2 = s1;
        s1 = pi;
    }
    return static_cast<std::uint32_t>(k - 1);
}

__global__ void __launch_bounds__(128) k_hull_gather(Dev d)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (std::uint32_t c = blockIdx.x * 4u + warp; c < K; c += gridDim.x * 4u)
    {
        if (lane == 0)
        {
            // z extent of the cluster (processor.cpp:648-655), reduced while labelling (cluster.cu)
            float zlo = unord_f32(d.zmin_u[o + c]), zhi = unord_f32(d.zmax_u[o + c]);
            const std::uint32_t zz = d.zzero[o + c];
            if (zz != 0xffffffffu && (zz & 1u) != 0u)
            {
                // a zero extent carries the sign of the cluster's first zero-height point (see accumulate_cluster_stats)
                zlo = zlo == 0.0f ? -0.0f : zlo;
                zhi = zhi == 0.0f ? -0.0f : zhi;
            }
            d.zminmax[o + c] = make_float2(zlo, zhi);
        }
        const std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + cstart[c] + c;
        const std::uint32_t off = hoff[c], hc = hoff[c + 1] - off;
        for (std::uint32_t t = lane; t < hc; t += 32)
        {
            const std::uint32_t idx = gst[t];
            const float4 p = d.pts_o[o + idx];
            d.hull_idx[o + off + t] = idx;
            d.hull_xy[o + off + t] = make_float2(p.x, p.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Shared-memory hull pass (the default path of launch_hulls).
//
// After the frame-wide sort every cluster is a contiguous, (x, y)-sorted segment. The segment is cut into chunks of
// at most kChunk points and ONE WARP turns a chunk into its hull entirely in shared memory: thinning by per-lane
// monotone chains over contiguous slices (a point that is on neither chain of its slice cannot be a hull vertex),
// then the reference's own sweep (polygonizer.cpp:67-90) over what is left, by one lane, with its operands a
// shared-memory access away - the global-memory round trips of the chains were what the old thinning passes spent
// their time on. A cluster of several chunks - the walls and hedges of 5-20k points that used to keep one warp busy
// while the rest of the GPU had finished - is thinned by as many warps in parallel (hull(A u B) = hull(hull(A) u
// hull(B))), and a second pass joins the chunks' survivors, which are still in sorted order.
// The orientation predicate is the reference's fp64 expression throughout, so the vertex set is the reference's.
// ------------------------------------------------------------------------------------------
#ifndef LPL_HULL_CHUNK
#define LPL_HULL_CHUNK 256 // measured per 154-frame batch (chunk pass): 1024 -> 0.34 ms, 512 -> 0.22, 256 -> 0.15, 128 -> 0.11 (join pass grows)
#endif
#ifndef LPL_HULL_JOIN
#define LPL_HULL_JOIN 1024 // buffer of the join pass: a cluster's chunk survivors should fit in one go
#endif
constexpr std::uint32_t kChunk = LPL_HULL_CHUNK;      // points one warp of the chunk pass holds in shared memory
constexpr std::uint32_t kJoin = LPL_HULL_JOIN;        // points one warp of the join pass holds
constexpr std::uint32_t kChunkWarps = 4;              // warps per CTA
#ifndef LPL_HULL_SWEEP
#define LPL_HULL_SWEEP 64 // survivors one lane sweeps without further thinning
#endif
#ifndef LPL_HULL_CHUNK_CTAS
#define LPL_HULL_CHUNK_CTAS 12 // CTAs per frame of the chunk pass (154-frame batch: about one resident wave)
#endif
constexpr std::uint32_t kSweepBelow = LPL_HULL_SWEEP;             // survivors one lane sweeps without further thinning
static_assert(kChunk % 32u == 0 && kChunk >= 64u && kChunk <= 2048u && kJoin % 32u == 0 && kJoin >= kChunk && kJoin <= 2048u, "buffer sizes");

template <std::uint32_t C>
struct WarpBufT
{
    static constexpr std::uint32_t kCap = C;
    static constexpr std::uint32_t kSliceLen = C / 32u; // points per lane in a thinning pass (= chain stack depth)
    float x[C];
    float y[C];
    std::uint32_t id[C];
    std::uint16_t st[2u * C + 2u]; // lane chain stacks (lower | upper); the sweep's m + 1 vertex stack + m vertex flags
};

template <class WarpBuf>
__device__ __forceinline__ P2 buf_pt(const WarpBuf& b, std::uint32_t i)
{
    P2 p;
    p.x = static_cast<double>(b.x[i]);
    p.y = static_cast<double>(b.y[i]);
    return p;
}

// cnt elements from global memory into the buffer at position `at`: four independent 16-byte loads per lane in flight
// before the first store (a load - store loop would pay one memory round trip per 32 elements)
template <class WarpBuf>
__device__ __forceinline__ void buf_load(WarpBuf& buf, std::uint32_t at, const uint4* __restrict__ src, std::uint32_t cnt)
{
    const std::uint32_t lane = lane_id();
    for (std::uint32_t t0 = 0; t0 < cnt; t0 += 128u)
    {
        uint4 e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const std::uint32_t t = t0 + k * 32u + lane;
            e[k] = t < cnt ? src[t] : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const std::uint32_t t = t0 + k * 32u + lane;
            if (t < cnt)
            {
                buf.x[at + t] = __uint_as_float(e[k].y);
                buf.y[at + t] = __uint_as_float(e[k].z);
                buf.id[at + t] = e[k].w;
            }
        }
    }
    __syncwarp();
}

// One thinning pass over the m sorted entries: lane l sweeps slice [a, b) with a lower (left to right) and an
// upper (right to left) chain; the survivors - the union of both chains - are compacted to the front in sorted
// order. Returns how many are left.
template <class WarpBuf>
__device__ std::uint32_t buf_thin(WarpBuf& buf, std::uint32_t m)
{
    const std::uint32_t lane = lane_id();
    constexpr std::uint32_t kSliceCap = WarpBuf::kSliceLen;
    // slices of ~sqrt(2 m) points balance this pass against what it leaves for the next one (a slice of s sorted points
    // keeps a handful); never more than the stack depth kSlice, so large inputs use all 32 lanes
    std::uint32_t per = min(8u, kSliceCap);
    while (per * per < 2u * m && per < kSliceCap)
    {
        ++per;
    }
    per = max(per, (m + 31u) / 32u);
    const std::uint32_t a = min(m, lane * per), b = min(m, a + per);
    std::uint16_t* L = buf.st + lane;                // entry k at L[k * 32]
    constexpr std::uint32_t kSlice = WarpBuf::kSliceLen;
    std::uint16_t* U = buf.st + 32u * kSlice + lane;
    std::uint32_t kl = 0, ku = 0;
    for (std::uint32_t i = a; i < b; ++i)
    {
        const P2 p = buf_pt(buf, i);
        while (kl >= 2 && not_left(buf_pt(buf, L[(kl - 2) * 32u]), buf_pt(buf, L[(kl - 1) * 32u]), p))
        {
            --kl;
        }
        L[kl * 32u] = static_cast<std::uint16_t>(i);
        ++kl;
    }
    for (std::uint32_t i = b; i > a; --i)
    {
        const P2 p = buf_pt(buf, i - 1u);
        while (ku >= 2 && not_left(buf_pt(buf, U[(ku - 2) * 32u]), buf_pt(buf, U[(ku - 1) * 32u]), p))
        {
            --ku;
        }
        U[ku * 32u] = static_cast<std::uint16_t>(i - 1u);
        ++ku;
    }
    // survivors of the slice, ascending position: merge of L (ascending) and U read from its top (ascending)
    float kx[kSlice], ky[kSlice];
    std::uint32_t kid[kSlice];
    std::uint32_t cnt = 0;
    {
        std::uint32_t i = 0, j = ku;
        while (i < kl || j > 0)
        {
            const std::uint32_t pl = i < kl ? L[i * 32u] : 0xffffffffu;
            const std::uint32_t pu = j > 0 ? U[(j - 1u) * 32u] : 0xffffffffu;
            i += (pl <= pu) ? 1u : 0u;
            j -= (pu <= pl) ? 1u : 0u;
            const std::uint32_t p = min(pl, pu);
            kx[cnt] = buf.x[p];
            ky[cnt] = buf.y[p];
            kid[cnt] = buf.id[p];
            ++cnt;
        }
    }
    const std::uint32_t incl = warp_incl_scan(cnt);
    const std::uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp(); // every lane has read its survivors
    const std::uint32_t w = incl - cnt;
    for (std::uint32_t k = 0; k < cnt; ++k)
    {
        buf.x[w + k] = kx[k];
        buf.y[w + k] = ky[k];
        buf.id[w + k] = kid[k];
    }
    __syncwarp();
    return total;
}

// thin the m sorted entries until few enough are left for one lane's sweep (or thinning stops paying: points in
// convex position); returns the survivor count, the survivors still sorted at the front of the buffer
template <class WarpBuf>
__device__ std::uint32_t buf_reduce(WarpBuf& buf, std::uint32_t m)
{
    bool stalled = false;
    while (m > kSweepBelow && !stalled)
    {
        const std::uint32_t m2 = buf_thin(buf, m);
        stalled = m2 * 4u > m * 3u;
        m = m2;
    }
    return m;
}

// exact reduction: the m sorted entries are replaced by their hull vertices (still sorted). Used by the join pass when
// thinning alone does not make room - its input is what earlier passes already thinned, so slices drop little.
template <class WarpBuf>
__device__ std::uint32_t buf_hull_only(WarpBuf& buf, std::uint32_t m)
{
    if (m < 3u)
    {
        return m; // nothing to drop (and the sweep is only defined from three points on)
    }
    const std::uint32_t lane = lane_id();
    std::uint16_t* flag = buf.st + WarpBuf::kCap + 1u; // behind the sweep's m + 1 stack entries
    for (std::uint32_t t = lane; t < m; t += 32u)
    {
        flag[t] = 0;
    }
    __syncwarp();
    if (lane == 0)
    {
        const std::uint32_t hc = monotone_chain([&](std::uint32_t i) { return buf_pt(buf, i); }, m, buf.st);
        for (std::uint32_t t = 0; t < hc; ++t)
        {
            flag[buf.st[t]] = 1;
        }
    }
    __syncwarp();
    std::uint32_t w = 0;
    for (std::uint32_t base = 0; base < m; base += 32u)
    {
        const std::uint32_t t = base + lane;
        const bool keep = t < m && flag[t] != 0;
        float x = 0.f, y = 0.f;
        std::uint32_t id = 0;
        if (keep)
        {
            x = buf.x[t];
            y = buf.y[t];
            id = buf.id[t];
        }
        const std::uint32_t mask = __ballot_sync(0xffffffffu, keep);
        __syncwarp(); // all reads of this batch before its writes (destinations never lie ahead of their sources)
        if (keep)
        {
            const std::uint32_t dst = w + __popc(mask & ((1u << lane) - 1u));
            buf.x[dst] = x;
            buf.y[dst] = y;
            buf.id[dst] = id;
        }
        w += __popc(mask);
        __syncwarp();
    }
    return w;
}

// the reference's sweep over the m sorted survivors by lane 0; vertex ids (obstacle-cloud indices) to gst, count returned
template <class WarpBuf>
__device__ std::uint32_t buf_sweep(WarpBuf& buf, std::uint32_t m, std::uint32_t* __restrict__ gst)
{
    std::uint32_t hc = 0;
    if (lane_id() == 0)
    {
        hc = monotone_chain([&](std::uint32_t i) { return buf_pt(buf, i); }, m, buf.st);
    }
    hc = __shfl_sync(0xffffffffu, hc, 0);
    __syncwarp();
    for (std::uint32_t t = lane_id(); t < hc; t += 32u)
    {
        gst[t] = buf.id[buf.st[t]];
    }
    __syncwarp();
    return hc;
}

// fewer than three points: identity order = obstacle-cloud order (polygonizer.cpp:36-41)
__device__ void hull_trivial(const uint4* __restrict__ seg, std::uint32_t m, std::uint32_t* __restrict__ gst, std::uint32_t* hcnt)
{
    if (lane_id() == 0)
    {
        std::uint32_t a = m > 0 ? seg[0].w : 0u, b = m > 1 ? seg[1].w : 0u;
        if (m == 2 && b < a)
        {
            const std::uint32_t t = a;
            a = b;
            b = t;
        }
        if (m > 0)
        {
            gst[0] = a;
        }
        if (m > 1)
        {
            gst[1] = b;
        }
        *hcnt = m;
    }
}

// chunks per cluster and their running sum (one CTA per frame)
__global__ void __launch_bounds__(1024) k_hull_plan(Dev d)
{
    __shared__ std::uint32_t sh[33];
    __shared__ std::uint32_t s_multi;
    if (threadIdx.x == 0)
    {
        s_multi = 0;
    }
    __syncthreads();
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::uint32_t* off = d.hwk_off + static_cast<std::size_t>(f) * (d.cap + 1);
    std::uint32_t carry = 0;
    for (std::uint32_t base = 0; base < K; base += 1024u)
    {
        const std::uint32_t c = base + threadIdx.x;
        const std::uint32_t m = c < K ? d.hseg_cnt[o + c] : 0u;
        const std::uint32_t chunks = (m + kChunk - 1u) / kChunk;
        std::uint32_t total;
        const std::uint32_t ex = block_excl_scan(chunks, sh, &total);
        if (c < K)
        {
            off[c] = carry + ex;
            d.hcnt[o + c] = 0; // clusters without a kept point keep an empty hull
            if (chunks > 1u)
            {
                d.hfin[o + atomicAdd(&s_multi, 1u)] = c; // work list of the join pass (order does not matter)
            }
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        off[K] = carry;
        d.n_work[f] = carry;
        d.n_multi[f] = s_multi;
        d.hull_next[f] = 0; // hand-out counter of k_hull_chunks
    }
}

// pass 1: one warp per chunk, chunks handed out dynamically (their cost differs by orders of magnitude); the
// cluster of a chunk is found by bisection over the chunk offsets, staged in shared memory when they fit
constexpr std::uint32_t kOffCache = 2048;

__global__ void __launch_bounds__(kChunkWarps * 32) k_hull_chunks(Dev d)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ std::uint32_t s_off[kOffCache];
    using WarpBuf = WarpBufT<kChunk>;
    WarpBuf& buf = reinterpret_cast<WarpBuf*>(s_raw)[threadIdx.x >> 5];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::uint32_t W = d.n_work[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* goff = d.hwk_off + static_cast<std::size_t>(f) * (d.cap + 1);
    const bool cached = K + 1u <= kOffCache;
    if (cached)
    {
        for (std::uint32_t t = threadIdx.x; t <= K; t += blockDim.x)
        {
            s_off[t] = goff[t];
        }
    }
    __syncthreads();
    const std::uint32_t* off = cached ? s_off : goff;
    const bool in_b = (sort_passes(d.n_h[f]) & 1u) != 0u;
    const uint4* sorted = (in_b ? d.hsB : d.hsA) + o;
    uint4* other = (in_b ? d.hsA : d.hsB) + o;
    const std::uint32_t lane = lane_id();
    while (true)
    {
        std::uint32_t w = 0;
        if (lane == 0)
        {
            w = atomicAdd(&d.hull_next[f], 1u);
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= W)
        {
            break;
        }
        // cluster of chunk w: the last c with off[c] <= w (clusters without chunks repeat the offset)
        std::uint32_t lo = 0, hi = K;
        while (hi - lo > 1u)
        {
            const std::uint32_t mid = (lo + hi) >> 1;
            if (off[mid] <= w)
            {
                lo = mid;
            }
            else
            {
                hi = mid;
            }
        }
        const std::uint32_t c = lo;
        const std::uint32_t j = w - off[c];
        const std::uint32_t mc = d.hseg_cnt[o + c];
        const std::uint32_t first = j * kChunk;
        const std::uint32_t m = min(kChunk, mc - first);
        const uint4* seg = sorted + cstart[c] + first;
        std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + cstart[c] + c; // segment length + 1 entries
        const bool single = mc <= kChunk;
        if (single && m < 3u)
        {
            hull_trivial(seg, m, gst, &d.hcnt[o + c]);
            continue;
        }
        buf_load(buf, 0u, seg, m);
        const std::uint32_t thinned = buf_reduce(buf, m);
        if (single)
        {
            const std::uint32_t hc = buf_sweep(buf, thinned, gst);
            if (lane == 0)
            {
                d.hcnt[o + c] = hc;
            }
        }
        else
        {
            // this chunk's own hull vertices (an exact reduction: the join pass gathers a few dozen points per chunk
            // instead of the up to kSweepBelow the thinning stops at), for the join pass
            const std::uint32_t left = buf_hull_only(buf, thinned);
            uint4* out = other + cstart[c] + first;
            for (std::uint32_t t = lane; t < left; t += 32u)
            {
                out[t] = make_uint4(c, __float_as_uint(buf.x[t]), __float_as_uint(buf.y[t]), buf.id[t]);
            }
            if (lane == 0)
            {
                d.hck_cnt[static_cast<std::size_t>(f) * 2 * d.cap + w] = left;
            }
        }
        __syncwarp();
    }
}

// pass 2: clusters of several chunks. The chunks' survivors (consecutive x ranges, so their concatenation is sorted)
// are gathered into the warp's buffer, thinned whenever the next chunk would not fit, and swept at the end. A cluster
// whose survivors cannot be brought under one buffer (more than ~kChunk points in convex position) is swept from
// its sorted segment in global memory: slow, always correct.
constexpr std::uint32_t kJoinWarps = 2; // warps per CTA of the join pass

__global__ void __launch_bounds__(kJoinWarps * 32) k_hull_join(Dev d)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    using WarpBuf = WarpBufT<kJoin>;
    WarpBuf& buf = reinterpret_cast<WarpBuf*>(s_raw)[threadIdx.x >> 5];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* off = d.hwk_off + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* ck = d.hck_cnt + static_cast<std::size_t>(f) * 2 * d.cap;
    const bool in_b = (sort_passes(d.n_h[f]) & 1u) != 0u;
    const uint4* sorted = (in_b ? d.hsB : d.hsA) + o;
    const uint4* other = (in_b ? d.hsA : d.hsB) + o;
    const std::uint32_t lane = lane_id();
    const std::uint32_t warps = gridDim.x * kJoinWarps;
    const std::uint32_t nm = d.n_multi[f];
    (void)K;
    for (std::uint32_t q = blockIdx.x * kJoinWarps + (threadIdx.x >> 5); q < nm; q += warps)
    {
        const std::uint32_t c = d.hfin[o + q];
        const std::uint32_t w0 = off[c], nchunks = off[c + 1] - w0;
        std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + cstart[c] + c;
        std::uint32_t fill = 0;
        bool overflow = false;
        // rounds over up to 32 chunks: lane j holds the survivor count of chunk j0 + j; as many leading chunks as fit
        // the buffer are gathered in one flat loop (every lane loading, four loads in flight each); when none fits,
        // the buffer is thinned first
        std::uint32_t j0 = 0;
        while (j0 < nchunks && !overflow)
        {
            const std::uint32_t cntj = j0 + lane < nchunks ? ck[w0 + j0 + lane] : 0u;
            const std::uint32_t incl = warp_incl_scan(cntj);
            std::uint32_t fit = __popc(__ballot_sync(0xffffffffu, j0 + lane < nchunks && fill + incl <= kJoin));
            if (fit == 0u)
            {
                fill = buf_reduce(buf, fill);
                fit = __popc(__ballot_sync(0xffffffffu, j0 + lane < nchunks && fill + incl <= kJoin));
                if (fit == 0u)
                {
                    fill = buf_hull_only(buf, fill);
                    fit = __popc(__ballot_sync(0xffffffffu, j0 + lane < nchunks && fill + incl <= kJoin));
                }
                if (fit == 0u)
                {
                    overflow = true; // even thinned, buffer + next chunk do not fit: points in convex position
                    break;
                }
            }
            const std::uint32_t total = __shfl_sync(0xffffffffu, incl, fit - 1u);
            const uint4* base = other + cstart[c] + j0 * kChunk;
            for (std::uint32_t e0 = 0; e0 < total; e0 += 128u)
            {
                uint4 v[4];
                std::uint32_t pos[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const std::uint32_t e = e0 + k * 32u + lane;
                    // chunk of flat element e: the first lane whose inclusive count exceeds e
                    std::uint32_t jj = 0;
#pragma unroll
                    for (std::uint32_t step = 16u; step > 0u; step >>= 1)
                    {
                        const std::uint32_t probe = __shfl_sync(0xffffffffu, incl, jj + step - 1u);
                        jj += (probe <= e) ? step : 0u;
                    }
                    jj = min(jj, 31u);
                    const std::uint32_t before = __shfl_sync(0xffffffffu, incl - cntj, jj);
                    pos[k] = e;
                    v[k] = e < total ? base[static_cast<std::size_t>(jj) * kChunk + (e - before)] : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    if (pos[k] < total)
                    {
                        buf.x[fill + pos[k]] = __uint_as_float(v[k].y);
                        buf.y[fill + pos[k]] = __uint_as_float(v[k].z);
                        buf.id[fill + pos[k]] = v[k].w;
                    }
                }
            }
            fill += total;
            j0 += fit;
            __syncwarp();
        }
        std::uint32_t hc = 0;
        if (!overflow)
        {
            hc = buf_sweep(buf, buf_reduce(buf, fill), gst);
        }
        else
        {
            const uint4* seg = sorted + cstart[c];
            const std::uint32_t m = d.hseg_cnt[o + c];
            if (lane == 0) // the vertex stack lives in the cluster's own output slots (segment length + 1 entries)
            {
                hc = monotone_chain([&](std::uint32_t i) { return elem_pt(seg[i]); }, m, gst);
            }
            hc = __shfl_sync(0xffffffffu, hc, 0);
            __syncwarp();
            for (std::uint32_t t0 = 0; t0 < hc; t0 += 32u)
            {
                const std::uint32_t t = t0 + lane;
                std::uint32_t v = 0;
                if (t < hc)
                {
                    v = seg[gst[t]].w;
                }
                __syncwarp();
                if (t < hc)
                {
                    gst[t] = v;
                }
            }
            __syncwarp();
        }
        if (lane == 0)
        {
            d.hcnt[o + c] = hc;
        }
        __syncwarp();
    }
}

// frame-wide merge sort of the n_h[f] elements in hsB by (x = label, y, z as floats, w = index); the result is in
// hsB when sort_passes(n_h[f]) is odd, else in hsA
void launch_hull_sort(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    if (nf <= 8)
    {
        k_hull_tilesort<1024><<<dim3(d.tiles, nf), 1024, 0, s>>>(d);
    }
    else
    {
        k_hull_tilesort<kTileThreads><<<dim3(d.tiles, nf), kTileThreads, 0, s>>>(d);
    }
    mark(c, "hull_tilesort");
    std::uint32_t passes = 0;
    while ((1u << passes) < d.tiles)
    {
        ++passes;
    }
    for (std::uint32_t p = 0; p < passes; ++p)
    {
        // frames whose obstacle cloud is already one sorted run leave at once
        k_hull_merge<<<dim3(std::min<std::uint32_t>(d.tiles, kMergeCtas), nf), kTileThreads, 0, s>>>(d, p);
        mark(c, "hull_merge");
    }
}

// ------------------------------------------------------------------------------------------
// Polygonizer::convexHull on coordinates that are NOT float-representable (arbitrary doubles; nothing on the node's
// path produces them - its points come from PCL floats - but the reference's signature accepts them): one CTA sorts
// the point indices by (x, y, index) on the doubles themselves (bitonic network in global memory) and one thread
// runs the reference's sweep. A fallback for generality, not a fast path.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool f64_less(const double2* xy, std::uint32_t a, std::uint32_t b)
{
    const double2 p = xy[a], q = xy[b];
    if (p.x != q.x)
    {
        return p.x < q.x;
    }
    if (p.y != q.y)
    {
        return p.y < q.y;
    }
    return a < b;
}

__global__ void __launch_bounds__(1024)
    k_hull_f64(const double2* __restrict__ xy, std::uint32_t n, std::uint32_t* __restrict__ order, std::uint32_t* __restrict__ st,
               std::uint32_t* __restrict__ out_idx, std::uint32_t* __restrict__ out_cnt)
{
    for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
    {
        order[t] = t;
    }
    __syncthreads();
    for (std::uint32_t k = 2; (k >> 1) < n; k <<= 1)
    {
        for (std::uint32_t t = threadIdx.x; t < n; t += blockDim.x)
        {
            const std::uint32_t u = t ^ (k - 1u);
            if (u > t && u < n)
            {
                const std::uint32_t a = order[t], b = order[u];
                if (f64_less(xy, b, a))
                {
                    order[t] = b;
                    order[u] = a;
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
                    const std::uint32_t a = order[t], b = order[u];
                    if (f64_less(xy, b, a))
                    {
                        order[t] = b;
                        order[u] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0)
    {
        std::uint32_t hc = n;
        if (n < 3u)
        {
            for (std::uint32_t t = 0; t < n; ++t)
            {
                out_idx[t] = t; // polygonizer.cpp:36-41: fewer than three points are returned as they are
            }
        }
        else
        {
            hc = monotone_chain(
                [&](std::uint32_t i) {
                    const double2 v = xy[order[i]];
                    P2 p;
                    p.x = v.x;
                    p.y = v.y;
                    return p;
                },
                n, st);
            for (std::uint32_t t = 0; t < hc; ++t)
            {
                out_idx[t] = order[st[t]];
            }
        }
        *out_cnt = hc;
    }
}

void launch_hull_f64(Ctx* c, const double2* xy, std::uint32_t n, std::uint32_t* order, std::uint32_t* st, std::uint32_t* out_idx,
                     std::uint32_t* out_cnt)
{
    k_hull_f64<<<1, 1024, 0, c->stream>>>(xy, n, order, st, out_idx, out_cnt);
    mark(c, "hull_f64");
}

void launch_hulls(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    k_hull_octagon<<<dim3(4, nf), 128, 0, s>>>(d);
    mark(c, "hull_octagon");
    // d.lab (RECM labels of the segmenter) is free by now: it holds the recorded verdicts
    launch_compact_recorded(c, "hull_keep", nf, d.tiles, d.n_o, d.tile_cnt, d.n_h, d.lab, HullKeepPred{d}, HullKeepEmit{d});
    k_excl_scan<<<nf, 1024, 0, s>>>(d.hseg_cnt, d.cap, d.cstart, d.cap + 1, d.cap, d.n_clusters, nullptr, d.f0);
    mark(c, "hull_seg_scan");
    launch_hull_sort(c, nf);
    k_hull_plan<<<nf, 1024, 0, s>>>(d);
    mark(c, "hull_plan");
    constexpr std::size_t smem = sizeof(WarpBufT<kChunk>) * kChunkWarps, smem_join = sizeof(WarpBufT<kJoin>) * kJoinWarps;
    cudaFuncSetAttribute(k_hull_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(k_hull_join, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_join));
    k_hull_chunks<<<dim3(per_frame_ctas(LPL_HULL_CHUNK_CTAS, nf, 256), nf), kChunkWarps * 32, smem, s>>>(d);
    mark(c, "hull_chunks");
    k_hull_join<<<dim3(per_frame_ctas(4, nf, 64), nf), kJoinWarps * 32, smem_join, s>>>(d);
    mark(c, "hull_join");
    k_excl_scan<<<nf, 1024, 0, s>>>(d.hcnt, d.cap, d.hull_off, d.cap + 1, d.cap, d.n_clusters, d.n_hull, d.f0);
    mark(c, "hull_off_scan");
    k_hull_gather<<<dim3(64, nf), 128, 0, s>>>(d);
    mark(c, "hull_gather");
}
} // namespace lpl
