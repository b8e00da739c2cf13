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
// One warp then builds each hull (k_hull_chain): clusters above 128 points are first thinned
// by per-lane monotone chains over contiguous chunks (a point that is not on the lower or upper
// hull of its chunk cannot be on the cluster's hull), repeatedly, and the survivors are swept by
// the reference's own lower/upper chain with the same fp64 orientation predicate evaluated
// without FMA contraction. The vertex list (coordinates) equals the reference's; only the *index*
// reported for exactly duplicated (x, y) points may differ, as it does between std::sort
// implementations.
#include "common.cuh"

namespace lpl
{
constexpr int kHullThreads = 32;   // one warp per CTA: the hardware balances clusters of very different size
#ifndef LPL_HULL_CTAS
#define LPL_HULL_CTAS 24
#endif
#ifndef LPL_HULL_BIG
#define LPL_HULL_BIG 1024
#endif
#ifndef LPL_HULL_LANEDIV
#define LPL_HULL_LANEDIV 2
#endif
#ifndef LPL_HULL_FILTER_ABOVE
#define LPL_HULL_FILTER_ABOVE 48
#endif
#ifndef LPL_HULL_FINAL_MAX
#define LPL_HULL_FINAL_MAX 256
#endif
constexpr int kHullCtasPerFrame = LPL_HULL_CTAS;  // single-warp CTAs per frame (about one resident wave for a 154-frame batch);
                                       // clusters are handed out dynamically, their sizes differ by orders of magnitude
constexpr std::uint32_t kChainSmem = 512; // survivors swept from shared memory
constexpr std::uint32_t kFilterAbove = LPL_HULL_FILTER_ABOVE;  // clusters above this are thinned by all lanes first
constexpr std::uint32_t kLaneStack = 16;  // per-lane chain stack entries kept in shared memory

// per-lane stack of (position, x, y): the first kLaneStack entries live in shared memory - a pop
// is on the critical path of the sweep and must not cost a trip to L2 for the popped point's
// coordinates - deeper ones spill to the lane's slice of a global scratch array (positions only).
// Entries of the lanes are interleaved (entry k of lane l at k * stride + l): no bank conflicts
// when the lanes work at the same depth.
struct LaneStack
{
    std::uint32_t* sm;    // shared-memory positions, already offset by the lane
    float2* smp;          // shared-memory coordinates, same layout
    std::uint32_t stride; // lanes in the group
    std::uint32_t* gl;    // this lane's global slice
    __device__ __forceinline__ std::uint32_t get(std::uint32_t k) const { return k < kLaneStack ? sm[k * stride] : gl[k]; }
    __device__ __forceinline__ void set(std::uint32_t k, std::uint32_t v, float2 xy) const
    {
        if (k < kLaneStack)
        {
            sm[k * stride] = v;
            smp[k * stride] = xy;
        }
        else
        {
            gl[k] = v;
        }
    }
};

// plain (coherent) load: the thinning passes re-read what earlier passes of the same kernel wrote
__device__ __forceinline__ uint4 ldg4(const uint4* p) { return *p; }

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
    const std::uint32_t f = blockIdx.y;
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
__global__ void __launch_bounds__(kTileThreads) k_hull_tilesort(Dev d)
{
    __shared__ uint4 s[kTile];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_h[f];
    const std::uint32_t base = blockIdx.x * kTile;
    if (base >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t m = min(static_cast<std::uint32_t>(kTile), n - base);
    for (std::uint32_t t = threadIdx.x; t < m; t += kTileThreads)
    {
        s[t] = d.hsB[o + base + t]; // (label, x, y, index) of the points that survived the octagon filter
    }
    __syncthreads();
    // every thread owns compare-exchange PAIRS (q-th pair of a stage: insert a zero bit at the stride
    // position), so no lane idles on the upper element of a pair
    for (std::uint32_t k = 2; (k >> 1) < m; k <<= 1)
    {
        const std::uint32_t hk = k >> 1;
        for (std::uint32_t q = threadIdx.x; q < kTile / 2; q += kTileThreads)
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
            for (std::uint32_t q = threadIdx.x; q < kTile / 2; q += kTileThreads)
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
    for (std::uint32_t t = threadIdx.x; t < m; t += kTileThreads)
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

// merge pass p: runs of (kTile << p) elements, pairwise, one output tile per CTA
__global__ void __launch_bounds__(kTileThreads) k_hull_merge(Dev d, std::uint32_t pass)
{
    __shared__ uint4 s[kTile];
    __shared__ std::uint32_t s_split[2];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_h[f];
    const std::uint32_t out0 = blockIdx.x * kTile;
    if (out0 >= n || sort_passes(n) <= pass)
    {
        return; // this frame was fully sorted by an earlier pass
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
        return;
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
}

// ------------------------------------------------------------------------------------------
// hull chains
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 elem_xy(const uint4& e)
{
    return make_float2(__uint_as_float(e.y), __uint_as_float(e.z));
}

__device__ __forceinline__ P2 stack_pt(const LaneStack& S, std::uint32_t k, const uint4* __restrict__ src)
{
    P2 p;
    if (k < kLaneStack)
    {
        const float2 v = S.smp[k * S.stride];
        p.x = static_cast<double>(v.x);
        p.y = static_cast<double>(v.y);
    }
    else
    {
        const uint4 e = src[S.gl[k]];
        p.x = static_cast<double>(__uint_as_float(e.y));
        p.y = static_cast<double>(__uint_as_float(e.z));
    }
    return p;
}

__device__ __forceinline__ P2 elem_pt(const uint4& e)
{
    P2 p;
    p.x = static_cast<double>(__uint_as_float(e.y));
    p.y = static_cast<double>(__uint_as_float(e.z));
    return p;
}

// One lane's share of a thinning pass: the lower chain (left to right) and the upper chain (right to
// left, as the reference walks it) over the contiguous chunk src[a..b). A point that is on neither
// chain of its chunk cannot be on the cluster's hull.
// Both chains advance in ONE warp-uniform loop as two independent state machines: per iteration a
// lane evaluates one orientation predicate per chain and either pops its stack or pushes the
// current point and moves on. Compared with a loop over points with an inner pop loop this keeps
// the lanes of a warp converged (a step costs one predicate, not the longest pop run among the
// lanes) and gives every lane two independent fp64 dependency chains to overlap. The next two
// elements of each direction are in flight while the current one is worked on.
// Must be called by all lanes of the warp (lanes with an empty chunk just vote).
__device__ __forceinline__ void lane_chains(const uint4* __restrict__ src, std::uint32_t a, std::uint32_t b,
                                            const LaneStack& L, const LaneStack& U, std::uint32_t& kl,
                                            std::uint32_t& ku)
{
    kl = 0;
    ku = 0;
    bool act_l = b > a, act_u = b > a;
    std::uint32_t il = a, iu = b - 1u; // current point of each chain (valid while active)
    P2 s2l = {0.0, 0.0}, s1l = s2l, s2u = s2l, s1u = s2l;
    uint4 el = make_uint4(0, 0, 0, 0), el1 = el, el2 = el, eu = el, eu1 = el, eu2 = el;
    if (act_l)
    {
        el = ldg4(src + a);
        el1 = ldg4(src + min(a + 1u, b - 1u));
        el2 = ldg4(src + min(a + 2u, b - 1u));
        eu = ldg4(src + (b - 1u));
        eu1 = ldg4(src + max(b - 1u, a + 1u) - 1u);
        eu2 = ldg4(src + max(b - 1u, a + 2u) - 2u);
    }
    while (__any_sync(0xffffffffu, act_l || act_u))
    {
        if (act_l)
        {
            const P2 p = elem_pt(el);
            if (kl >= 2 && not_left(s2l, s1l, p))
            {
                --kl;
                s1l = s2l;
                if (kl >= 2)
                {
                    s2l = stack_pt(L, kl - 2, src);
                }
            }
            else
            {
                L.set(kl, il, elem_xy(el));
                ++kl;
                s2l = s1l;
                s1l = p;
                ++il;
                if (il < b)
                {
                    el = el1;
                    el1 = el2;
                    el2 = ldg4(src + min(il + 2u, b - 1u));
                }
                else
                {
                    act_l = false;
                }
            }
        }
        if (act_u)
        {
            const P2 p = elem_pt(eu);
            if (ku >= 2 && not_left(s2u, s1u, p))
            {
                --ku;
                s1u = s2u;
                if (ku >= 2)
                {
                    s2u = stack_pt(U, ku - 2, src);
                }
            }
            else
            {
                U.set(ku, iu, elem_xy(eu));
                ++ku;
                s2u = s1u;
                s1u = p;
                if (iu > a)
                {
                    --iu;
                    eu = eu1;
                    eu1 = eu2;
                    eu2 = ldg4(src + max(iu, a + 2u) - 2u);
                }
                else
                {
                    act_u = false;
                }
            }
        }
    }
}

// survivors of one lane = union of its ascending lower list and its descending upper list
__device__ __forceinline__ std::uint32_t lane_survivors(const LaneStack& L, const LaneStack& U, std::uint32_t kl,
                                                        std::uint32_t ku)
{
    std::uint32_t cnt = 0;
    std::uint32_t i = 0, j = ku;
    while (i < kl || j > 0)
    {
        const std::uint32_t pl = i < kl ? L.get(i) : 0xffffffffu;
        const std::uint32_t pu = j > 0 ? U.get(j - 1) : 0xffffffffu;
        i += (pl <= pu) ? 1u : 0u;
        j -= (pu <= pl) ? 1u : 0u;
        ++cnt;
    }
    return cnt;
}

__device__ __forceinline__ void lane_emit(const uint4* __restrict__ src, uint4* __restrict__ dst, std::uint32_t w,
                                          const LaneStack& L, const LaneStack& U, std::uint32_t kl, std::uint32_t ku)
{
    std::uint32_t i = 0, j = ku;
    while (i < kl || j > 0)
    {
        const std::uint32_t pl = i < kl ? L.get(i) : 0xffffffffu;
        const std::uint32_t pu = j > 0 ? U.get(j - 1) : 0xffffffffu;
        i += (pl <= pu) ? 1u : 0u;
        j -= (pu <= pl) ? 1u : 0u;
        dst[w++] = ldg4(src + min(pl, pu));
    }
}

// Warp-wide thinning pass: lane l sweeps its contiguous chunk of src[0..m); the survivor lists, in
// sorted order, are compacted into dst. stL / stU: global spill space for the per-lane stacks
// (m entries each); smL / smU / spL / spU: 32 * kLaneStack shared-memory entries each. Returns the
// survivor count.
__device__ std::uint32_t hull_filter(const uint4* __restrict__ src, std::uint32_t m, uint4* __restrict__ dst,
                                     std::uint32_t* __restrict__ stL, std::uint32_t* __restrict__ stU,
                                     std::uint32_t* smL, std::uint32_t* smU, float2* spL, float2* spU)
{
    const std::uint32_t lane = lane_id();
    // balance the two sequential phases: a lane sweeps m / lanes points now and every lane leaves
    // roughly eight survivors for the next (sequential or thinner) sweep -> lanes ~ sqrt(m / 8)
    std::uint32_t lanes = 2u;
    while (lanes < 32u && lanes * lanes * LPL_HULL_LANEDIV < m)
    {
        ++lanes;
    }
    const std::uint32_t chunk = (m + lanes - 1u) / lanes;
    const std::uint32_t a = min(m, lane * chunk), b = min(m, a + chunk);
    const LaneStack L{smL + lane, spL + lane, 32u, stL + a};
    const LaneStack U{smU + lane, spU + lane, 32u, stU + a};
    std::uint32_t kl, ku;
    lane_chains(src, a, b, L, U, kl, ku);
    const std::uint32_t cnt = lane_survivors(L, U, kl, ku);
    const std::uint32_t incl = warp_incl_scan(cnt);
    const std::uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    lane_emit(src, dst, incl - cnt, L, U, kl, ku);
    __syncwarp();
    return total;
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

// hfin[c]: what k_hull_thin leaves for k_hull_final
constexpr std::uint32_t kFinDone = 0x80000000u;   // hull already written by k_hull_thin
constexpr std::uint32_t kFinOther = 0x40000000u;  // survivors live in the second sort buffer
constexpr std::uint32_t kFinalMax = LPL_HULL_FINAL_MAX;          // survivors swept by one thread (local-memory stack)

constexpr std::uint32_t kFinPre = 0x20000000u;    // k_hull_thin_big already thinned the cluster once (into the other buffer)
constexpr std::uint32_t kBigAbove = LPL_HULL_BIG;         // clusters above this get a CTA-wide first thinning pass
constexpr int kBigThreads = 256;
constexpr int kBigCtasPerFrame = 8;

// Pass 0: the handful of very large clusters of a frame (walls, vegetation: up to ~20k points) would
// leave one warp sweeping 600-point chunks per lane while everything else has finished; a whole
// CTA thins them once first (256 chunks), the warp passes of k_hull_thin take over from there.
// Also resets hfin for every cluster of the frame.
__global__ void __launch_bounds__(kBigThreads) k_hull_thin_big(Dev d)
{
    extern __shared__ __align__(16) unsigned char s_big[]; // 2 x (positions + coordinates) x kBigThreads x kLaneStack
    float2* s_xy = reinterpret_cast<float2*>(s_big);
    std::uint32_t* s_pos = reinterpret_cast<std::uint32_t*>(s_xy + 2 * kBigThreads * kLaneStack);
    __shared__ std::uint32_t s_scan[33];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const bool in_b = (sort_passes(d.n_h[f]) & 1u) != 0u;
    const uint4* sorted = (in_b ? d.hsB : d.hsA) + o;
    uint4* other = (in_b ? d.hsA : d.hsB) + o;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        d.hull_next[f] = 0; // hand-out counter of k_hull_thin
    }
    for (std::uint32_t c = blockIdx.x; c < K; c += gridDim.x)
    {
        const std::uint32_t seg = cstart[c];
        const std::uint32_t m = cstart[c + 1] - seg;
        if (m <= kBigAbove)
        {
            if (threadIdx.x == 0)
            {
                d.hfin[o + c] = 0;
            }
            continue;
        }
        const uint4* src = sorted + seg;
        uint4* dst = other + seg;
        const std::uint32_t chunk = (m + kBigThreads - 1u) / kBigThreads;
        const std::uint32_t a = min(m, threadIdx.x * chunk), b = min(m, a + chunk);
        const LaneStack L{s_pos + threadIdx.x, s_xy + threadIdx.x, kBigThreads, d.hstL + o + seg + a};
        const LaneStack U{s_pos + kBigThreads * kLaneStack + threadIdx.x, s_xy + kBigThreads * kLaneStack + threadIdx.x,
                          kBigThreads, d.hstU + o + seg + a};
        std::uint32_t kl, ku;
        lane_chains(src, a, b, L, U, kl, ku);
        const std::uint32_t cnt = lane_survivors(L, U, kl, ku);
        std::uint32_t total;
        const std::uint32_t w = block_excl_scan(cnt, s_scan, &total);
        lane_emit(src, dst, w, L, U, kl, ku);
        if (threadIdx.x == 0)
        {
            d.hfin[o + c] = total | kFinPre;
        }
        __syncthreads(); // the stacks / s_scan are reused by the next cluster
    }
}

// Pass 1, one warp per cluster above kFilterAbove points: thinning passes ping-pong between the two
// sort buffers (the segment is private to the warp) until few enough points survive for a single
// thread, or thinning stops paying (points in convex position), in which case the warp sweeps the
// survivors itself.
__global__ void __launch_bounds__(kHullThreads) k_hull_thin(Dev d)
{
    __shared__ float2 s_key[kChainSmem];
    __shared__ std::uint16_t s_st[kChainSmem + 2];
    __shared__ std::uint32_t s_lane[2][32 * kLaneStack];
    __shared__ float2 s_lxy[2][32 * kLaneStack];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::uint32_t lane = lane_id();
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const bool in_b = (sort_passes(d.n_h[f]) & 1u) != 0u;
    uint4* sorted = (in_b ? d.hsB : d.hsA) + o;
    uint4* other = (in_b ? d.hsA : d.hsB) + o;
    while (true)
    {
        std::uint32_t c = 0;
        if (lane == 0)
        {
            c = atomicAdd(&d.hull_next[f], 1u);
        }
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= K)
        {
            break;
        }
        const std::uint32_t seg = cstart[c];
        const std::uint32_t n = cstart[c + 1] - seg;
#ifdef LPL_HULL_TRACE
        // diagnostics (build with LPL_NVCC_EXTRA=-DLPL_HULL_TRACE): per cluster, the survivors and the
        // cycles of every thinning pass and of what follows; printed for clusters above 50k cycles
        struct Trace
        {
            long long t0, tp[6];
            std::uint32_t f, c, n, mp[6], np;
            __device__ void pass(std::uint32_t m)
            {
                if (np < 6)
                {
                    tp[np] = clock64();
                    mp[np] = m;
                    ++np;
                }
            }
            __device__ ~Trace()
            {
                const long long t1 = clock64();
                if (t1 - t0 > 50000 && (threadIdx.x & 31u) == 0)
                {
                    printf("hull_thin f %u c %u n %u total %lld passes %u:", f, c, n, t1 - t0, np);
                    long long prev = t0;
                    for (std::uint32_t k = 0; k < np; ++k)
                    {
                        printf(" [m %u cyc %lld]", mp[k], tp[k] - prev);
                        prev = tp[k];
                    }
                    printf(" tail %lld\n", t1 - prev);
                }
            }
        } trace{clock64(), {}, f, c, n, {}, 0u};
#endif
        if (n <= kFilterAbove)
        {
            if (lane == 0)
            {
                d.hfin[o + c] = n;
            }
            continue;
        }
        std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + seg + c; // n + 1 entries
        uint4* cur = sorted + seg;
        uint4* nxt = other + seg;
        std::uint32_t m = n;
        bool in_other = false;
        const std::uint32_t pre = d.hfin[o + c];
        if (pre & kFinPre)
        {
            // k_hull_thin_big left its survivors in the other buffer
            m = pre & 0x1fffffffu;
            cur = other + seg;
            nxt = sorted + seg;
            in_other = true;
        }
        bool stalled = false;
        while (m > kFilterAbove && !stalled)
        {
            const std::uint32_t m2 = hull_filter(cur, m, nxt, d.hstL + o + seg, d.hstU + o + seg, s_lane[0], s_lane[1], s_lxy[0], s_lxy[1]);
            uint4* t = cur;
            cur = nxt;
            nxt = t;
            in_other = !in_other;
            stalled = m2 * 4u > m * 3u; // convex-position input: thinning does not pay
            m = m2;
#ifdef LPL_HULL_TRACE
            trace.pass(m);
#endif
        }
        if (m <= kFinalMax)
        {
            if (lane == 0)
            {
                d.hfin[o + c] = m | (in_other ? kFinOther : 0u);
            }
            continue;
        }
        std::uint32_t hc = 0;
        if (m <= kChainSmem)
        {
            for (std::uint32_t t = lane; t < m; t += 32)
            {
                const uint4 e = cur[t];
                s_key[t] = make_float2(__uint_as_float(e.y), __uint_as_float(e.z));
            }
            __syncwarp();
            if (lane == 0)
            {
                const float2* key = s_key;
                hc = monotone_chain(
                    [&](std::uint32_t i) {
                        const float2 v = key[i];
                        P2 p;
                        p.x = static_cast<double>(v.x);
                        p.y = static_cast<double>(v.y);
                        return p;
                    },
                    m, s_st);
            }
            hc = __shfl_sync(0xffffffffu, hc, 0);
            __syncwarp();
            for (std::uint32_t t = lane; t < hc; t += 32)
            {
                gst[t] = cur[s_st[t]].w;
            }
            __syncwarp();
        }
        else
        {
            // rare: more than kChainSmem points in (near) convex position; sweep in global memory
            std::uint32_t* st = (m == n) ? gst : (d.hstL + o + seg); // m + 1 entries
            if (lane == 0)
            {
                hc = monotone_chain([&](std::uint32_t i) { return elem_pt(cur[i]); }, m, st);
            }
            hc = __shfl_sync(0xffffffffu, hc, 0);
            __syncwarp();
            // positions -> obstacle-cloud indices (in place when st == gst)
            for (std::uint32_t t0 = 0; t0 < hc; t0 += 32)
            {
                const std::uint32_t t = t0 + lane;
                std::uint32_t v = 0;
                if (t < hc)
                {
                    v = cur[st[t]].w;
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
            d.hfin[o + c] = kFinDone;
        }
    }
}

// Pass 2, one thread per cluster: z extent, the trivial cases, and the reference's sweep over the
// (at most kFinalMax) surviving points with the stack in local memory. 32 clusters per warp keep
// the sequential sweeps from wasting 31 of 32 lanes.
__global__ void __launch_bounds__(64) k_hull_final(Dev d)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const bool in_b = (sort_passes(d.n_h[f]) & 1u) != 0u;
    for (std::uint32_t c = blockIdx.x * 64u + threadIdx.x; c < K; c += gridDim.x * 64u)
    {
    const std::uint32_t seg = cstart[c];
    const std::uint32_t n = cstart[c + 1] - seg;
    // z extent of the cluster (processor.cpp:648-655), reduced while labelling (cluster.cu)
    {
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
    const std::uint32_t fin = d.hfin[o + c];
    if (fin & kFinDone)
    {
        continue;
    }
    const bool use_b = in_b != ((fin & kFinOther) != 0u);
    const uint4* cur = (use_b ? d.hsB : d.hsA) + o + seg;
    std::uint32_t* gst = d.hstack + static_cast<std::size_t>(f) * 2 * d.cap + seg + c; // n + 1 entries
    if (n < 3)
    {
        // identity order = obstacle-cloud order (polygonizer.cpp:36-41)
        std::uint32_t a = (n > 0) ? cur[0].w : 0u, b = (n > 1) ? cur[1].w : 0u;
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
        continue;
    }
    const std::uint32_t m = fin & 0x3fffffffu;
    std::uint16_t st[kFinalMax + 2];
    const std::uint32_t hc = monotone_chain([&](std::uint32_t i) { return elem_pt(cur[i]); }, m, st);
    for (std::uint32_t t = 0; t < hc; ++t)
    {
        gst[t] = cur[st[t]].w;
    }
    d.hcnt[o + c] = hc;
    }
}

__global__ void __launch_bounds__(128) k_hull_gather(Dev d)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* cstart = d.cstart + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (std::uint32_t c = blockIdx.x * 4u + warp; c < K; c += gridDim.x * 4u)
    {
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

// frame-wide merge sort of the n_h[f] elements in hsB by (x = label, y, z as floats, w = index); the result is in
// hsB when sort_passes(n_h[f]) is odd, else in hsA
void launch_hull_sort(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    k_hull_tilesort<<<dim3(d.tiles, nf), kTileThreads, 0, s>>>(d);
    mark(c, "hull_tilesort");
    std::uint32_t passes = 0;
    while ((1u << passes) < d.tiles)
    {
        ++passes;
    }
    for (std::uint32_t p = 0; p < passes; ++p)
    {
        // frames whose obstacle cloud is already one sorted run leave at once
        k_hull_merge<<<dim3(d.tiles, nf), kTileThreads, 0, s>>>(d, p);
        mark(c, "hull_merge");
    }
}

void launch_hulls(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    k_hull_octagon<<<dim3(4, nf), 128, 0, s>>>(d);
    mark(c, "hull_octagon");
    // d.lab (RECM labels of the segmenter) is free by now: it holds the recorded verdicts
    launch_compact_recorded(c, "hull_keep", nf, d.tiles, d.n_o, d.tile_cnt, d.n_h, d.lab, HullKeepPred{d}, HullKeepEmit{d});
    k_excl_scan<<<nf, 1024, 0, s>>>(d.hseg_cnt, d.cap, d.cstart, d.cap + 1, d.cap, d.n_clusters, nullptr);
    mark(c, "hull_seg_scan");
    launch_hull_sort(c, nf);
    constexpr std::size_t big_smem = static_cast<std::size_t>(2) * kBigThreads * kLaneStack * (sizeof(float2) + sizeof(std::uint32_t));
    cudaFuncSetAttribute(k_hull_thin_big, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(big_smem));
    k_hull_thin_big<<<dim3(per_frame_ctas(kBigCtasPerFrame, nf, 64), nf), kBigThreads, big_smem, s>>>(d);
    mark(c, "hull_thin_big");
    k_hull_thin<<<dim3(per_frame_ctas(kHullCtasPerFrame, nf, 1024), nf), kHullThreads, 0, s>>>(d);
    mark(c, "hull_thin");
    k_hull_final<<<dim3(per_frame_ctas(16, nf, 256), nf), 64, 0, s>>>(d);
    mark(c, "hull_final");
    k_excl_scan<<<nf, 1024, 0, s>>>(d.hcnt, d.cap, d.hull_off, d.cap + 1, d.cap, d.n_clusters, d.n_hull);
    mark(c, "hull_off_scan");
    k_hull_gather<<<dim3(64, nf), 128, 0, s>>>(d);
    mark(c, "hull_gather");
}
} // namespace lpl
