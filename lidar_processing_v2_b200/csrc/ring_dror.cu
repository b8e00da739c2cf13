// Stage 0 (ring partition) and stage 1 (DROR outlier filter).
//
// Reference behaviour:
//   ring partition  src/dataloader/src/dataloader.cpp:68-137 (Dataloader::addRingInfo)
//   DROR            lidar_processing_lib/src/noise_remover.cpp:38-68 (NoiseRemover::filter) with
//                   the neighbour predicate of kdtree.hpp:131-149,363-371; "exact" semantics
//                   (count of points with dist_sqr <= r_sqr, self included, compared with
//                   min_neighbours) — the reference's KD-tree early exit leaves a stale traversal
//                   stack whose effect depends on nth_element's tree shape (hazard H1).
//
// One scan-line pass over float4 points settles ~95 % of an organised sweep and the ring wrap flags (16 B/pt in,
// 4 B out); the rest are searched exactly in a two-level grid built from the points near them, the queries taken in
// grid order (k_dror_query).
#include "common.cuh"

namespace lpl
{
// ------------------------------------------------------------------------------------------
// Ring partition: ring(i) = max(0, 63 - #{j <= i : quadrant(j) == FIRST && quadrant(j-1) == FOURTH}).
// The sequential counter of the reference becomes a two-pass tile scan.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int quadrant_of(float x, float y)
{
    // More than ~1e-3 rad away from every axis the quadrant follows from the signs alone: neither the
    // rounding of atan2f nor of the "+ 2 pi" / the float thresholds can move the azimuth across a
    // boundary there. Only points next to an axis (and zeros / NaNs) take the exact glibc path.
    const float ax = fabsf(x), ay = fabsf(y);
    if (ay > ax * 9.765625e-4f && ax > ay * 9.765625e-4f)
    {
        return y > 0.f ? (x > 0.f ? 0 : 1) : (x > 0.f ? 3 : 2);
    }
    // dataloader.cpp:96-114; M_PI_2f, M_PIf and 1.5F * M_PIf as float constants
    float az = atan2f_glibc(y, x);
    az = (az < 0.f) ? (az + 2.0f * 3.14159265358979323846f) : az;
    if (az < 1.57079632679489661923f)
    {
        return 0;
    }
    if (az < 3.14159265358979323846f)
    {
        return 1;
    }
    if (az < 1.5f * 3.14159265358979323846f)
    {
        return 2;
    }
    return 3;
}

struct RingWrapPred
{
    const float4* pts;
    std::uint32_t cap;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        if (i == 0)
        {
            return false; // previous_quadrant starts as FIRST (dataloader.cpp:86)
        }
        const float4* p = pts + static_cast<std::size_t>(f) * cap;
        const float4 a = p[i];
        if (quadrant_of(a.x, a.y) != 0)
        {
            return false;
        }
        const float4 b = p[i - 1];
        return quadrant_of(b.x, b.y) == 3;
    }
};

__global__ void __launch_bounds__(kTileThreads)
    k_ring_write(RecordedPred pred, const std::uint32_t* __restrict__ n_arr,
                 const std::uint32_t* __restrict__ tile_cnt, std::uint32_t tiles_per_frame, std::uint32_t cnt_per_tile,
                 std::uint16_t* __restrict__ ring, std::uint32_t cap, std::uint32_t f0)
{
    __shared__ std::uint32_t sh[kItems * (kTileThreads / 32) + 1];
    __shared__ std::uint32_t sh2[33];
    const std::uint32_t f = blockIdx.y + f0;
    const std::uint32_t n = n_arr[f];
    const std::uint32_t base = blockIdx.x * kTile;
    if (base >= n)
    {
        return;
    }
    // wrap counts of everything ahead of this tile: `cnt_per_tile` counters per 2048-point tile (1 from the
    // stand-alone counting pass, 64 per-warp counters from the fused DROR scan-line pass)
    std::uint32_t before = 0;
    for (std::uint32_t t = threadIdx.x; t < blockIdx.x * cnt_per_tile; t += kTileThreads)
    {
        before += tile_cnt[static_cast<std::size_t>(f) * tiles_per_frame * cnt_per_tile + t];
    }
    before = block_sum(before, sh2);
    bool flag[kItems];
    std::uint32_t rank[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j)
    {
        const std::uint32_t i = base + j * kTileThreads + threadIdx.x;
        flag[j] = (i < n) && pred(f, i);
    }
    std::uint32_t total;
    tile_ranks(flag, rank, &total, sh);
#pragma unroll
    for (int j = 0; j < kItems; ++j)
    {
        const std::uint32_t i = base + j * kTileThreads + threadIdx.x;
        if (i < n)
        {
            const std::uint32_t wraps = before + rank[j] + (flag[j] ? 1u : 0u); // inclusive
            ring[static_cast<std::size_t>(f) * cap + i] =
                static_cast<std::uint16_t>(wraps >= 63u ? 0u : 63u - wraps);
        }
    }
}

void launch_ring(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    const dim3 grid(d.tiles, nf);
    // the wrap flags are evaluated once (first pass) and recorded in d.lab, which the segmenter only
    // writes later in the chain
    const RingWrapPred pred{d.pts_in, d.cap};
    k_compact_count<<<grid, kTileThreads, 0, c->stream>>>(RecordingPred<RingWrapPred>{pred, d.lab, d.cap}, d.n_in, 0u,
                                                          d.tile_cnt, d.tiles, d.f0);
    mark(c, "ring_count");
    k_ring_write<<<grid, kTileThreads, 0, c->stream>>>(RecordedPred{d.lab, d.cap}, d.n_in, d.tile_cnt, d.tiles, 1u, d.ring,
                                                      d.cap, d.f0);
    mark(c, "ring_write");
}

// ------------------------------------------------------------------------------------------
// DROR
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float dror_radius_sqr(float x, float y, const DrorParams& p)
{
    // noise_remover.cpp:57-59: float range^2 widened to double, scaled, narrowed, clamped
    const float range_sqr_f = (x * x) + (y * y);
    const float dyn = static_cast<float>(p.scaling * static_cast<double>(range_sqr_f));
    return fmaxf(dyn, p.min_r_sqr);
}

__device__ __forceinline__ bool dror_within(const float4& target, const float4& node, float r_sqr)
{
    // kdtree.hpp:131-143: (a0-b0)^2 + ((a1-b1)^2 + ((a2-b2)^2 + 0)), a = target, b = node
    const float d0 = target.x - node.x;
    const float d1 = target.y - node.y;
    const float d2 = target.z - node.z;
    const float dist = d0 * d0 + (d1 * d1 + (d2 * d2 + 0.0f));
    return dist <= r_sqr;
}

#ifndef LPL_DROR_HALO
#define LPL_DROR_HALO 5
#endif
constexpr int kNearHalo = LPL_DROR_HALO; // scan-line neighbours examined on each side

// Pass A: LiDAR clouds arrive in firing order, so a point's nearest neighbours are almost always
// its predecessors / successors on the same ring. Counting those first settles the vast majority
// of points (VALID as soon as min_neighbours are found) with one coalesced tile load; the
// remaining points go to the exhaustive grid search. Any input order gives the same result.
template <bool kWithRing>
__global__ void __launch_bounds__(256)
    k_dror_near(Dev d, DrorParams prm)
{
    __shared__ float4 sh[256 + 2 * kNearHalo];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t base = blockIdx.x * 256u;
    if (base >= n)
    {
        return;
    }
    const float4* pts = d.pts_in + static_cast<std::size_t>(f) * d.cap;
    for (int t = threadIdx.x; t < 256 + 2 * kNearHalo; t += 256)
    {
        const long long gi = static_cast<long long>(base) + t - kNearHalo;
        // beyond the ends of the cloud: NaN, within() of nothing - the neighbour loop needs no bounds tests
        float4 v = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        if (gi >= 0 && gi < static_cast<long long>(n))
        {
            v = pts[gi];
        }
        sh[t] = v;
    }
    __syncthreads();
    const std::uint32_t i = base + threadIdx.x;
    bool unresolved = false;
    bool wrap = false;
    if (i < n)
    {
        const float4 p = sh[threadIdx.x + kNearHalo];
        if (kWithRing)
        {
            // ring partition, pass 1 (RingWrapPred) on the tile that is in shared memory anyway: the wrap flag of
            // point i needs point i - 1, which the halo holds
            if (i != 0 && quadrant_of(p.x, p.y) == 0)
            {
                const float4 b = sh[threadIdx.x + kNearHalo - 1];
                wrap = quadrant_of(b.x, b.y) == 3;
            }
            d.lab[static_cast<std::size_t>(f) * d.cap + i] = wrap ? 1 : 0;
        }
        const float r_sqr = dror_radius_sqr(p.x, p.y, prm);
        std::uint32_t cnt = dror_within(p, p, r_sqr) ? 1u : 0u; // self (dist 0 unless NaN)
        bool done = false;
#pragma unroll
        for (int o = 1; o <= kNearHalo; ++o)
        {
            if (!done)
            {
                cnt += dror_within(p, sh[threadIdx.x + kNearHalo - o], r_sqr) ? 1u : 0u;
                cnt += dror_within(p, sh[threadIdx.x + kNearHalo + o], r_sqr) ? 1u : 0u;
                // on a continuous surface the two nearest scan neighbours on each side already settle the point
                done = cnt >= prm.min_neighbours;
            }
        }
        unresolved = cnt < prm.min_neighbours;
        // VALID, or 2 = "query of the grid search" (k_dror_grid_scatter tags the point's grid copy with it and
        // k_dror_query writes the final verdict over it)
        d.noise[static_cast<std::size_t>(f) * d.cap + i] = unresolved ? 2 : 0;
    }
    // CTA-aggregated append: one atomic per CTA on the frame's counter (per warp, an unorganised 2 M-point cloud - all of
    // whose points are unresolved - queued 62k atomics on one address per frame: that, not the distance tests, was the
    // whole cost of this pass there)
    {
        __shared__ std::uint32_t s_wcnt[8], s_base;
        const std::uint32_t m = __ballot_sync(0xffffffffu, unresolved);
        const std::uint32_t w = threadIdx.x >> 5;
        if (lane_id() == 0)
        {
            s_wcnt[w] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            std::uint32_t tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
            {
                const std::uint32_t c = s_wcnt[k];
                s_wcnt[k] = tot;
                tot += c;
            }
            s_base = tot != 0u ? atomicAdd(&d.n_unres[f], tot) : 0u;
        }
        __syncthreads();
        if (unresolved)
        {
            d.unres[static_cast<std::size_t>(f) * d.cap + s_base + s_wcnt[w] + __popc(m & ((1u << lane_id()) - 1u))] = i;
        }
    }
    if (kWithRing)
    {
        // wraps of this warp's 32 points (no block reduction: its barriers were the kernel's top stall);
        // k_ring_write sums the counters ahead of its tile
        const std::uint32_t wm = __ballot_sync(0xffffffffu, wrap);
        if (lane_id() == 0)
        {
            d.wrap_cnt[static_cast<std::size_t>(f) * (d.cap / 32u) + (i >> 5)] = __popc(wm);
        }
    }
}

// Two grid levels share one cell index space: level 0 = 1 m cells over +-128 m (cells
// [0, kDrorLevelCells)), level 1 = 0.125 m cells over the inner +-16 m (cells [kDrorLevelCells,
// 2 * kDrorLevelCells)). A point lives in exactly one level (the fine one iff it lies inside the
// inner square), a query scans its search box in both. The search radius grows with range
// (0.1 m below 5 m, 0.02 * range beyond), so near the sensor - where a 1 m cell holds hundreds of
// points - the fine level cuts the candidates per query several times.
#ifndef LPL_DROR_INNER
#define LPL_DROR_INNER 16.0f
#endif
constexpr float kDrorInner = LPL_DROR_INNER;
#ifndef LPL_DROR_FINE
#define LPL_DROR_FINE 8.0f
#endif
constexpr float kDrorFineScale = LPL_DROR_FINE;
static_assert(2.0f * kDrorInner * kDrorFineScale == static_cast<float>(kDrorGrid), "the fine level spans exactly kDrorGrid cells");

__device__ __forceinline__ int dror_cell_coord(float v)
{
    int c = static_cast<int>(floorf(v)) + kDrorGrid / 2;
    return min(max(c, 0), kDrorGrid - 1);
}

__device__ __forceinline__ bool dror_is_inner(float x, float y)
{
    return x >= -kDrorInner && x < kDrorInner && y >= -kDrorInner && y < kDrorInner;
}

__device__ __forceinline__ int dror_point_cell(const float4& p)
{
    if (dror_is_inner(p.x, p.y))
    {
        return kDrorLevelCells + dror_cell_coord(p.y * kDrorFineScale) * kDrorGrid + dror_cell_coord(p.x * kDrorFineScale);
    }
    return dror_cell_coord(p.y) * kDrorGrid + dror_cell_coord(p.x);
}

// cells a query's search disc can touch on one level (conservative cover of the float predicate)
struct DrorBox
{
    int x0, x1, y0, y1; // inclusive cell coordinates; empty when y1 < y0
    int base;           // first cell index of the level
};

__device__ __forceinline__ DrorBox dror_box(const float4& p, float r_sqr, int level)
{
    const float rc = sqrtf(r_sqr) * 1.001f + 1e-4f;
    const float xl = p.x - rc, xh = p.x + rc, yl = p.y - rc, yh = p.y + rc;
    DrorBox b;
    if (level == 0)
    {
        b.base = 0;
        b.x0 = dror_cell_coord(xl);
        b.x1 = dror_cell_coord(xh);
        b.y0 = dror_cell_coord(yl);
        b.y1 = dror_cell_coord(yh);
        if (xl >= -kDrorInner && xh < kDrorInner && yl >= -kDrorInner && yh < kDrorInner)
        {
            b.y1 = b.y0 - 1; // box inside the inner square: every candidate lives on the fine level
        }
    }
    else
    {
        b.base = kDrorLevelCells;
        b.x0 = dror_cell_coord(xl * kDrorFineScale);
        b.x1 = dror_cell_coord(xh * kDrorFineScale);
        b.y0 = dror_cell_coord(yl * kDrorFineScale);
        b.y1 = dror_cell_coord(yh * kDrorFineScale);
        if (!(xh >= -kDrorInner && xl < kDrorInner && yh >= -kDrorInner && yl < kDrorInner))
        {
            b.y1 = b.y0 - 1; // box misses the inner square
        }
    }
    return b;
}

// Only ~4 % of the points reach the exhaustive search, and only the points near them can be
// their neighbours: every unresolved query marks the cells of its search boxes in a per-frame
// bitmap (16 KB), and the grid is then built from the points of marked cells alone.
__global__ void __launch_bounds__(256) k_dror_mark(Dev d, DrorParams prm)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nu = d.n_unres[f];
    std::uint32_t* mask = d.grid_mask + static_cast<std::size_t>(f) * (kDrorCells / 32);
    if (nu >= d.n_in[f] / 4u)
    {
        // an unorganised cloud leaves (nearly) every point to the grid search: their boxes cover every occupied
        // cell anyway, and millions of atomicOr on 16 KB of bitmap would cost more than the whole search
        for (std::uint32_t w = blockIdx.x * 256u + threadIdx.x; w < static_cast<std::uint32_t>(kDrorCells / 32); w += gridDim.x * 256u)
        {
            mask[w] = nu != 0u ? 0xffffffffu : 0u;
        }
        return;
    }
    for (std::uint32_t u = blockIdx.x * 256u + threadIdx.x; u < nu; u += gridDim.x * 256u)
    {
        const std::uint32_t i = d.unres[static_cast<std::size_t>(f) * d.cap + u];
        const float4 p = d.pts_in[static_cast<std::size_t>(f) * d.cap + i];
        const float r_sqr = dror_radius_sqr(p.x, p.y, prm);
        for (int level = 0; level < 2; ++level)
        {
            const DrorBox b = dror_box(p, r_sqr, level);
            for (int cy = b.y0; cy <= b.y1; ++cy)
            {
                // the cells x0..x1 of a row are consecutive bits: one OR per touched word
                const std::uint32_t first = static_cast<std::uint32_t>(b.base + cy * kDrorGrid + b.x0);
                const std::uint32_t last = static_cast<std::uint32_t>(b.base + cy * kDrorGrid + b.x1);
                for (std::uint32_t w = first >> 5; w <= (last >> 5); ++w)
                {
                    const std::uint32_t lo = (w == (first >> 5)) ? (first & 31u) : 0u;
                    const std::uint32_t hi = (w == (last >> 5)) ? (last & 31u) : 31u;
                    const std::uint32_t bits = (0xffffffffu >> (31u - hi)) & (0xffffffffu << lo);
                    if ((__ldcg(mask + w) & bits) != bits)
                    {
                        atomicOr(mask + w, bits);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ bool dror_marked(const std::uint32_t* mask, int cell)
{
    return (mask[static_cast<std::uint32_t>(cell) >> 5] >> (static_cast<std::uint32_t>(cell) & 31u)) & 1u;
}

__global__ void __launch_bounds__(256) k_dror_grid_count(Dev d)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    if (blockIdx.x * 256u >= n || d.n_unres[f] == 0)
    {
        return;
    }
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    int cell = -1;
    if (i < n)
    {
        const float4 p = d.pts_in[static_cast<std::size_t>(f) * d.cap + i];
        const int c = dror_point_cell(p);
        if (dror_marked(d.grid_mask + static_cast<std::size_t>(f) * (kDrorCells / 32), c))
        {
            cell = c;
        }
    }
    if (i < n)
    {
        d.cell[static_cast<std::size_t>(f) * d.cap + i] = cell; // recorded for k_dror_grid_scatter (the segmenter's plane is free here)
    }
    const std::uint32_t peers = __match_any_sync(0xffffffffu, cell);
    if (cell >= 0 && static_cast<int>(lane_id()) == __ffs(peers) - 1)
    {
        atomicAdd(&d.grid_cnt[static_cast<std::size_t>(f) * kDrorCells + cell], static_cast<std::uint32_t>(__popc(peers)));
    }
}

__global__ void __launch_bounds__(256) k_dror_grid_scatter(Dev d)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    if (blockIdx.x * 256u >= n || d.n_unres[f] == 0)
    {
        return;
    }
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    int cell = -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n)
    {
        cell = d.cell[static_cast<std::size_t>(f) * d.cap + i]; // recorded by k_dror_grid_count (-1: cell not marked)
        if (cell >= 0)
        {
            p = d.pts_in[static_cast<std::size_t>(f) * d.cap + i];
            // w: the point's index, bit 31 set when the point is itself a query (k_dror_query walks the grid copy)
            const std::uint32_t q = d.noise[static_cast<std::size_t>(f) * d.cap + i] == 2 ? 0x80000000u : 0u;
            p.w = __uint_as_float(i | q);
        }
    }
    const std::uint32_t peers = __match_any_sync(0xffffffffu, cell);
    if (cell >= 0)
    {
        const int leader = __ffs(peers) - 1;
        std::uint32_t k = 0;
        if (static_cast<int>(lane_id()) == leader)
        {
            // counting down returns grid_cnt to zero for the next batch
            k = atomicSub(&d.grid_cnt[static_cast<std::size_t>(f) * kDrorCells + cell],
                          static_cast<std::uint32_t>(__popc(peers)));
        }
        k = __shfl_sync(peers, k, leader) - 1u - __popc(peers & ((1u << lane_id()) - 1u));
        const std::uint32_t pos = d.grid_start[static_cast<std::size_t>(f) * (kDrorCells + 1) + cell] + k;
        d.grid_pts[static_cast<std::size_t>(f) * d.cap + pos] = p;
    }
}

// Pass B: exhaustive count for the unresolved points over every grid cell the search disc can
// touch. Cells of one grid row are contiguous in grid_pts, so each row is one range. Rows are
// short far from the sensor (where most unresolved points live), so a query is served by a group
// of 8 lanes: four queries per warp in flight, 32 points per step, early exit at min_neighbours.
//
// The queries are taken in GRID order, not in list order: every unresolved point lies in a marked cell (its own
// search box covers it), so its grid copy exists and carries the query flag. A warp reads 32 consecutive grid
// points (one coalesced load: coordinates and index in one record, where the list order paid two dependent loads),
// ballots the flagged ones and serves them four at a time. Consecutive grid points share a cell, so the queries a
// warp - and its CTA - has in flight scan the same rows: the row bounds and the candidates are L1 hits instead of
// one DRAM sector per query and row (the unorganised 2 M-point cloud, where every point is a query: 1.82 -> 0.x ms).
constexpr int kDrorQueryWarps = 8;
#ifndef LPL_DROR_GROUP
#define LPL_DROR_GROUP 8
#endif
constexpr int kDrorGroup = LPL_DROR_GROUP; // lanes per query
#ifndef LPL_DROR_CTAS
#define LPL_DROR_CTAS 48 // measured per 154-frame batch: 12 -> 0.294 ms, 24 -> 0.277, 48 -> 0.266
#endif
constexpr int kDrorQueryCtas = LPL_DROR_CTAS; // per frame; warps stride over the chunks of 32 grid points
#ifndef LPL_DROR_UNROLL
#define LPL_DROR_UNROLL 4
#endif
constexpr int kDrorUnroll = LPL_DROR_UNROLL; // points per lane and step
#ifndef LPL_DROR_DENSE
#define LPL_DROR_DENSE 8
#endif
constexpr int kDrorDense = LPL_DROR_DENSE; // queries among 32 consecutive grid points from which the chunk-mate pass pays

#ifndef LPL_DROR_MINB
#define LPL_DROR_MINB 6 // measured per 154-frame batch: 1 (64 registers, 50 % occupancy) -> 0.288 ms, 6 (40 registers) -> 0.266, 8 -> 0.273
#endif
__global__ void __launch_bounds__(kDrorQueryWarps * 32, LPL_DROR_MINB) k_dror_query(Dev d, DrorParams prm)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    if (d.n_unres[f] == 0u)
    {
        return;
    }
    const std::uint32_t lane = lane_id();
    const std::uint32_t gl = lane & (kDrorGroup - 1);                 // lane inside the group
    const std::uint32_t grp = lane / kDrorGroup;                      // group inside the warp
    const std::uint32_t gmask = ((1u << kDrorGroup) - 1u) << (lane & ~(kDrorGroup - 1u));
    constexpr std::uint32_t kGroups = 32 / kDrorGroup;
    const std::uint32_t* start = d.grid_start + static_cast<std::size_t>(f) * (kDrorCells + 1);
    const float4* gp = d.grid_pts + static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t total = start[kDrorCells];
    for (std::uint32_t c0 = (blockIdx.x * kDrorQueryWarps + (threadIdx.x >> 5)) * 32u; c0 < total;
         c0 += gridDim.x * kDrorQueryWarps * 32u)
    {
        float4 mine = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f); // beyond the end: NaN, within() of nothing
        if (c0 + lane < total)
        {
            mine = gp[c0 + lane];
        }
        const bool is_query = (__float_as_uint(mine.w) >> 31) != 0u;
        std::uint32_t pending = __ballot_sync(0xffffffffu, is_query);
        if (__popc(pending) >= kDrorDense)
        {
            // Dense chunk (an unorganised cloud: every point is a query): the other 31 grid points of the chunk lie in
            // the same or the next cells, and every point is a legitimate candidate (the distance test is exact, the
            // search box only prunes) - so each lane first counts among its chunk mates, nearest in grid order first,
            // entirely out of registers. Most queries of a dense region have their min_neighbours here; the verdict is
            // final for them (the count can only grow), the others take the exhaustive search below from zero.
            const float my_r = dror_radius_sqr(mine.x, mine.y, prm);
            std::uint32_t cnt = dror_within(mine, mine, my_r) ? 1u : 0u;
            bool open = is_query && cnt < prm.min_neighbours;
#pragma unroll 1
            for (int j0 = 1; j0 < 32 && __any_sync(0xffffffffu, open); j0 += 4)
            {
#pragma unroll
                for (int j = j0; j < j0 + 4 && j < 32; ++j)
                {
                    const int sft = ((j & 1) != 0) ? (j + 1) / 2 : -(j / 2);
                    const int src = (static_cast<int>(lane) + sft) & 31;
                    float4 o;
                    o.x = __shfl_sync(0xffffffffu, mine.x, src);
                    o.y = __shfl_sync(0xffffffffu, mine.y, src);
                    o.z = __shfl_sync(0xffffffffu, mine.z, src);
                    cnt += dror_within(mine, o, my_r) ? 1u : 0u;
                }
                open = open && cnt < prm.min_neighbours;
            }
            if (is_query && !open)
            {
                d.noise[static_cast<std::size_t>(f) * d.cap + (__float_as_uint(mine.w) & 0x7fffffffu)] = 0;
            }
            pending = __ballot_sync(0xffffffffu, open);
        }
        while (pending != 0u)
        {
            // group g takes the g-th pending query
            std::uint32_t pm = pending;
#pragma unroll
            for (std::uint32_t j = 0; j + 1 < kGroups; ++j)
            {
                pm = j < grp ? (pm & (pm - 1u)) : pm;
            }
            const bool active = pm != 0u;
            const int src = active ? __ffs(pm) - 1 : 0;
            float4 p;
            p.x = __shfl_sync(0xffffffffu, mine.x, src);
            p.y = __shfl_sync(0xffffffffu, mine.y, src);
            p.z = __shfl_sync(0xffffffffu, mine.z, src);
            const std::uint32_t i = __float_as_uint(__shfl_sync(0xffffffffu, mine.w, src)) & 0x7fffffffu;
#pragma unroll
            for (std::uint32_t j = 0; j < kGroups; ++j)
            {
                pending &= pending - 1u;
            }
            if (!active)
            {
                continue; // (whole groups only: the group-wide shuffles below stay converged)
            }
            const float r_sqr = dror_radius_sqr(p.x, p.y, prm);
            std::uint32_t cnt = 0;
            for (int level = 1; level >= 0 && cnt < prm.min_neighbours; --level)
            {
                const DrorBox bx = dror_box(p, r_sqr, level);
                for (int cy0 = bx.y0; cy0 <= bx.y1 && cnt < prm.min_neighbours; cy0 += kDrorGroup)
                {
                    // lane gl fetches the bounds of row cy0 + gl
                    const int cy = cy0 + static_cast<int>(gl);
                    std::uint32_t ra = 0, rb = 0;
                    if (cy <= bx.y1)
                    {
                        ra = start[bx.base + cy * kDrorGrid + bx.x0];
                        rb = start[bx.base + cy * kDrorGrid + bx.x1 + 1];
                    }
                    const int rows = min(kDrorGroup, bx.y1 - cy0 + 1);
                    // rows are visited outwards from the query's own row: a query that does have its
                    // min_neighbours finds them in the nearest cells and leaves early
                    const float yq = (level == 1) ? p.y * kDrorFineScale : p.y;
                    const int own = min(max(dror_cell_coord(yq) - cy0, 0), rows - 1);
                    for (int t = 0; t < 2 * rows && cnt < prm.min_neighbours; ++t)
                    {
                        const int r = own + (((t & 1) != 0) ? (t + 1) / 2 : -(t / 2));
                        if (r < 0 || r >= rows)
                        {
                            continue;
                        }
                        const std::uint32_t a = __shfl_sync(gmask, ra, r, kDrorGroup);
                        const std::uint32_t b = __shfl_sync(gmask, rb, r, kDrorGroup);
                        // kDrorUnroll independent loads per lane and step: a dense row (hundreds of points in
                        // a near-range cell) is latency bound, a sparse one only ever issues the first load
                        for (std::uint32_t k0 = a; k0 < b && cnt < prm.min_neighbours; k0 += kDrorGroup * kDrorUnroll)
                        {
                            float4 q[kDrorUnroll];
#pragma unroll
                            for (int j = 0; j < kDrorUnroll; ++j)
                            {
                                const std::uint32_t k = k0 + j * kDrorGroup + gl;
                                q[j] = k < b ? gp[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            std::uint32_t hits = 0;
#pragma unroll
                            for (int j = 0; j < kDrorUnroll; ++j)
                            {
                                const std::uint32_t k = k0 + j * kDrorGroup + gl;
                                hits += (k < b && dror_within(p, q[j], r_sqr)) ? 1u : 0u;
                            }
#pragma unroll
                            for (int sft = kDrorGroup / 2; sft > 0; sft >>= 1)
                            {
                                hits += __shfl_xor_sync(gmask, hits, sft, kDrorGroup);
                            }
                            cnt += hits;
                        }
                    }
                }
            }
            if (gl == 0)
            {
                d.noise[static_cast<std::size_t>(f) * d.cap + i] = cnt < prm.min_neighbours ? 1 : 0;
            }
        }
    }
}

__global__ void k_excl_scan(const std::uint32_t* __restrict__ in, std::uint32_t in_stride,
                            std::uint32_t* __restrict__ out, std::uint32_t out_stride,
                            std::uint32_t len, const std::uint32_t* __restrict__ len_arr,
                            std::uint32_t* __restrict__ total_out, std::uint32_t f0)
{
    __shared__ std::uint32_t sh[33];
    const std::uint32_t f = blockIdx.x + f0;
    if (len_arr != nullptr)
    {
        len = min(len, len_arr[f]);
    }
    const std::uint32_t* src = in + static_cast<std::size_t>(f) * in_stride;
    std::uint32_t* dst = out + static_cast<std::size_t>(f) * out_stride;
    // tiles of 4 * blockDim.x counters, four consecutive counters per thread (one 16-byte load when
    // the row is aligned), block scan per tile, running carry across tiles
    std::uint32_t total = 0;
    const bool vec = (reinterpret_cast<std::uintptr_t>(src) & 15u) == 0;
    for (std::uint32_t base = 0; base < len; base += 4u * blockDim.x)
    {
        const std::uint32_t k = base + 4u * threadIdx.x;
        std::uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (vec && k + 3u < len)
        {
            const uint4 q = *reinterpret_cast<const uint4*>(src + k);
            v[0] = q.x;
            v[1] = q.y;
            v[2] = q.z;
            v[3] = q.w;
        }
        else
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                v[j] = (k + j < len) ? src[k + j] : 0u;
            }
        }
        std::uint32_t tile_total;
        std::uint32_t run = total + block_excl_scan(v[0] + v[1] + v[2] + v[3], sh, &tile_total);
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (k + j < len)
            {
                dst[k + j] = run;
            }
            run += v[j];
        }
        total += tile_total;
    }
    if (threadIdx.x == 0)
    {
        dst[len] = total;
        if (total_out != nullptr)
        {
            total_out[f] = total;
        }
    }
}

// Exclusive scan of the grid counters of one frame, restricted to the words of the relevance bitmap
// that matter: a cell outside every search box holds no point, and its start offset is never read.
// A 32-cell word is active when it has a marked cell or follows a word whose last cell is marked (a
// query reads start[last cell of its row range + 1]). One CTA per frame: compact the active words
// (4096 bitmap words, one block scan), sum the 32 counters of each (thread per word, eight 16-byte
// loads in flight), scan the sums, then write the starts (warp per word, coalesced). ~0.2 MB of
// traffic per frame instead of the 1 MB of a full scan over the 131,072 cells.
#ifndef LPL_DROR_SPARSE_SCAN
#define LPL_DROR_SPARSE_SCAN 1
#endif
constexpr int kDrorWords = kDrorCells / 32;
static_assert(kDrorWords == 4096, "k_dror_grid_scan stages one bitmap word index per thread x 4");

__global__ void __launch_bounds__(1024) k_dror_grid_scan(Dev d)
{
    __shared__ std::uint32_t sh[33];
    __shared__ std::uint16_t s_list[kDrorWords];
    __shared__ std::uint32_t s_off[kDrorWords];
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::uint32_t* mask = d.grid_mask + static_cast<std::size_t>(f) * kDrorWords;
    const std::uint32_t* cnt = d.grid_cnt + static_cast<std::size_t>(f) * kDrorCells;
    std::uint32_t* start = d.grid_start + static_cast<std::size_t>(f) * (kDrorCells + 1);
    // active words, compacted in order
    const std::uint32_t w0 = 4u * threadIdx.x;
    const uint4 mq = *reinterpret_cast<const uint4*>(mask + w0);
    const std::uint32_t before = threadIdx.x == 0 ? 0u : mask[w0 - 1u];
    const std::uint32_t m[5] = {before, mq.x, mq.y, mq.z, mq.w};
    bool act[4];
    std::uint32_t na = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        act[j] = m[j + 1] != 0u || (m[j] >> 31) != 0u;
        na += act[j] ? 1u : 0u;
    }
    std::uint32_t nact;
    std::uint32_t pos = block_excl_scan(na, sh, &nact);
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        if (act[j])
        {
            s_list[pos++] = static_cast<std::uint16_t>(w0 + j);
        }
    }
    __syncthreads();
    // counters per active word
    for (std::uint32_t i = threadIdx.x; i < nact; i += blockDim.x)
    {
        const uint4* row = reinterpret_cast<const uint4*>(cnt + 32u * s_list[i]);
        std::uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const uint4 q = row[j];
            sum += q.x + q.y + q.z + q.w;
        }
        s_off[i] = sum;
    }
    __syncthreads();
    // exclusive scan over the active words' sums (at most 4096: four per thread)
    std::uint32_t v[4], tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        v[j] = (w0 + j < nact) ? s_off[w0 + j] : 0u;
        tsum += v[j];
    }
    std::uint32_t total;
    std::uint32_t run = block_excl_scan(tsum, sh, &total);
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        if (w0 + j < nact)
        {
            s_off[w0 + j] = run;
        }
        run += v[j];
    }
    __syncthreads();
    // starts of the cells of the active words
    const std::uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (std::uint32_t i0 = warp; i0 < nact; i0 += 4u * nwarps)
    {
        std::uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const std::uint32_t i = i0 + j * nwarps;
            c[j] = i < nact ? cnt[32u * s_list[i] + lane] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const std::uint32_t i = i0 + j * nwarps;
            if (i < nact)
            {
                start[32u * s_list[i] + lane] = s_off[i] + warp_incl_scan(c[j]) - c[j];
            }
        }
    }
    if (threadIdx.x == 0)
    {
        start[kDrorCells] = total;
    }
}

// with_ring: the ring partition (stage 0) rides along - its wrap flags are evaluated by the scan-line pass, which has
// every point and its predecessor in shared memory, and k_ring_write turns them into ring indices; the cloud is
// then read once for both stages instead of twice.
void launch_dror(Ctx* c, std::uint32_t nf, bool with_ring)
{
    Dev& d = c->d;
    if (!c->counters_cleared)
    {
        cudaMemsetAsync(at_frame(d.n_unres, 1, d.f0), 0, sizeof(std::uint32_t) * nf, c->stream);
    }
    const dim3 grid((d.cap + 255) / 256, nf);
    if (with_ring)
    {
        k_dror_near<true><<<grid, 256, 0, c->stream>>>(d, c->dror);
        mark(c, "front");
        k_ring_write<<<dim3(d.tiles, nf), kTileThreads, 0, c->stream>>>(RecordedPred{d.lab, d.cap}, d.n_in, d.wrap_cnt, d.tiles,
                                                                         static_cast<std::uint32_t>(kTile / 32), d.ring, d.cap, d.f0);
        mark(c, "ring_write");
    }
    else
    {
        k_dror_near<false><<<grid, 256, 0, c->stream>>>(d, c->dror);
        mark(c, "dror_near");
    }
    cudaMemsetAsync(at_frame(d.grid_mask, kDrorCells / 32, d.f0), 0, sizeof(std::uint32_t) * (kDrorCells / 32) * nf, c->stream);
    k_dror_mark<<<dim3(per_frame_ctas(8, nf, 256), nf), 256, 0, c->stream>>>(d, c->dror);
    mark(c, "dror_mark");
    k_dror_grid_count<<<grid, 256, 0, c->stream>>>(d);
    mark(c, "dror_grid_count");
#if LPL_DROR_SPARSE_SCAN
    k_dror_grid_scan<<<nf, 1024, 0, c->stream>>>(d);
#else
    k_excl_scan<<<nf, 1024, 0, c->stream>>>(d.grid_cnt, kDrorCells, d.grid_start, kDrorCells + 1, kDrorCells,
                                            nullptr, nullptr, d.f0);
#endif
    mark(c, "dror_grid_scan");
    k_dror_grid_scatter<<<grid, 256, 0, c->stream>>>(d);
    mark(c, "dror_grid_scatter");
    k_dror_query<<<dim3(per_frame_ctas(kDrorQueryCtas, nf, 2048), nf), kDrorQueryWarps * 32, 0, c->stream>>>(d, c->dror);
    mark(c, "dror_query");
}

} // namespace lpl
