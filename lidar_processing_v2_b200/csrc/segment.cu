// Stage 2: JCP ground segmentation on the RECM (+ near-field RANSAC), batched over frames.
//
// Reference: lidar_processing_lib/src/segmenter.cpp
//   constructPolarGrid :103-204   -> k_seg_bin, k_excl_scan, k_seg_scatter, k_seg_cell
//   RECM               :206-283   -> k_seg_cell (robust per-cell minimum), k_seg_elev; the obstacle test itself is
//                                    evaluated for the pixel winners only (k_seg_px)
//   RANSAC             :321-479   -> candidates (k_seg_label), k_ransac_draw, k_ransac_plane, k_ransac_count,
//                                    k_ransac_best; the plane is applied to the pixel winners (k_seg_px)
//   image scatter      :291-318   -> 64-bit atomicMin keys (issued by k_seg_scatter), k_seg_px (winner decode + its label);
//                                    the range image is a plane of point indices
//   JCP                :481-638   -> k_seg_dilate_tma (5x5 stencil on range-image tiles staged by TMA),
//                                    queue compaction, k_jcp_pre, k_jcp_rows
//   populateLabels     :640-669   -> k_seg_labels_out
//
// Ordered semantics on an unordered machine: the reference iterates polar cells in index order
// and the points of a cell in cloud order. Only two results depend on that order and both are
// recovered without sorting the cells: (i) the <= 120 RANSAC draws address candidates by their
// rank in (cell, cloud) order - a prefix over per-cell candidate counts plus a radix select over
// the point indices of one cell resolves a rank (k_ransac_draw / k_ransac_plane); (ii) the range image keeps the
// first point among equal depths - equal depth^2 implies the same radial bin, so the 64-bit key
// (depth^2, azimuth slice, point index) under atomicMin picks the reference's winner (k_seg_scatter / k_seg_px).
// Everything else (heights per cell, inlier counts, labels) is order independent.
//
// JCP is a Gauss-Seidel sweep in raster order. A queued pixel only depends on queued pixels
// that precede it in raster order and lie within the kernel distance: weights and mask sources are
// computed for all queued pixels in parallel (k_jcp_pre); then one CTA per frame sweeps the image
// row by row, one thread per queued pixel, the in-row recurrence solved by a scan over finite maps
// (k_jcp_rows).
// The reference leaves out-of-image kernel slots untouched, so border pixels inherit those
// slots from the previously popped pixel (hazard H2); k_jcp_pre reproduces that by
// locating the most recent earlier queued pixel for which the slot was inside the image.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace lpl
{
__constant__ int c_off_h[24] = {-2, -2, -2, -2, -2, -1, -1, -1, -1, -1, 0, 0,
                                0,  0,  1,  1,  1,  1,  1,  2,  2,  2,  2, 2};
__constant__ int c_off_w[24] = {-2, -1, 0, 1, 2, -2, -1, 0, 1, 2, -2, -1,
                                1,  2,  -2, -1, 0, 1, 2, -2, -1, 0, 1, 2};

// ------------------------------------------------------------------------------------------
// polar binning (segmenter.cpp:124-203)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_seg_bin(Dev d, SegParams sp)
{
    // The segmenter works on the input cloud in place: points DROR flagged as NOISE simply are not
    // binned. Everything order dependent downstream only uses the relative order of the points,
    // which a stable compaction of the VALID points would preserve - so the compaction (and the
    // index indirection it needs) is skipped; n_v only counts the valid points.
    __shared__ std::uint32_t s_valid;
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    if (threadIdx.x == 0)
    {
        s_valid = 0;
    }
    __syncthreads();
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::int32_t cell = -1;
    std::uint32_t px = 0;
    // the three loads of a point (noise flag, coordinates, ring) are independent: issue them together instead of
    // one round trip after the other (the kernel is bound by exactly that latency chain, not by bandwidth)
    std::uint8_t nz = 1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    std::uint16_t ring_i = 0;
    if (i < n)
    {
        nz = d.noise[o + i];
        p = d.pts_in[o + i];
        ring_i = sp.use_ring ? d.ring[o + i] : std::uint16_t(0);
    }
    const bool valid = i < n && nz == 0;
    if (valid)
    {
        if (!(p.z < sp.z_lo || p.z > sp.z_hi))
        {
            const float dist = sqrtf(p.x * p.x + p.y * p.y);
            const std::int32_t radial = static_cast<std::int32_t>(dist / sp.radial_spacing);
            if (!(dist < sp.min_dist || dist > sp.max_dist || radial >= sp.rings))
            {
                float az = atan2_approx(p.y, p.x);
                az = (az < 0.f) ? (az + 6.28318530717958647692f) : az;
                const std::int32_t az_idx = min(static_cast<std::int32_t>(az / sp.slice_res), sp.slices - 1);
                std::int32_t hgt;
                bool ok = true;
                if (sp.use_ring)
                {
                    hgt = static_cast<std::int32_t>(ring_i);
                    ok = hgt < sp.H;
                }
                else
                {
                    const float el = atanf_glibc(p.z / dist);
                    hgt = static_cast<std::int32_t>((el - sp.el_down) / sp.rad_per_px);
                    ok = !(hgt < 0 || hgt >= sp.H);
                }
                if (ok)
                {
                    // static_cast<uint16_t>((W - 1) * az / TWO_M_PIf): float -> int32 -> low 16 bits
                    const std::int32_t wi = static_cast<std::int32_t>(sp.wscale * az / 6.28318530717958647692f);
                    const std::uint32_t wid = static_cast<std::uint32_t>(wi) & 0xffffu;
                    cell = az_idx * sp.rings + radial;
                    px = static_cast<std::uint32_t>(hgt) * sp.W + wid;
                }
            }
        }
    }
    const std::uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, valid));
    if (lane_id() == 0 && nvalid != 0)
    {
        atomicAdd(&s_valid, nvalid);
    }
    // consecutive points of a scan line mostly share a cell: one atomic per (warp, cell)
    const std::uint32_t peers = __match_any_sync(0xffffffffu, cell);
    std::uint32_t slot = 0;
    if (cell >= 0)
    {
        const int leader = __ffs(peers) - 1;
        std::uint32_t base = 0;
        if (static_cast<int>(lane_id()) == leader)
        {
            base = atomicAdd(&d.cell_cnt[static_cast<std::size_t>(f) * sp.ncell + cell], __popc(peers));
        }
        base = __shfl_sync(peers, base, leader);
        slot = base + __popc(peers & ((1u << lane_id()) - 1u));
    }
    if (i < n)
    {
        d.cell[o + i] = cell;
        d.px[o + i] = px;
        d.slot[o + i] = slot;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_valid != 0)
    {
        atomicAdd(&d.n_v[f], s_valid);
    }
}

// cell-major copy of (point index, z): the only per-cell consumers are the height statistics of
// k_seg_cell and the rank selection of the RANSAC draws, neither of which needs cloud order
__global__ void __launch_bounds__(256) k_seg_scatter(Dev d, SegParams sp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    // independent loads first (cell, slot, z), then the one dependent look-up (the cell's start)
    const std::int32_t cell = d.cell[o + i];
    const std::uint32_t slot = d.slot[o + i];
    const std::uint32_t px = d.px[o + i];
    const float4 p = d.pts_in[o + i];
    if (cell >= 0)
    {
        const std::uint32_t s = d.cell_start[static_cast<std::size_t>(f) * (sp.ncell + 1) + cell];
        d.zo[o + s + slot] = make_uint2(i, __float_as_uint(p.z));
        // range-image scatter (segmenter.cpp:291-318), see k_seg_px: the winner of a pixel does not depend on the
        // labels, so the key goes out with this pass and no later pass re-reads the cloud for it
        const float d2 = (p.x * p.x) + (p.y * p.y);
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(d2)) << 33) |
                                       (static_cast<unsigned long long>(cell / sp.rings) << sp.idx_bits) |
                                       static_cast<unsigned long long>(i);
        atomicMin(&d.key[static_cast<std::size_t>(f) * sp.npx + px], key);
    }
}

// ------------------------------------------------------------------------------------------
// per-cell work: restore cloud order, robust minimum (segmenter.cpp:240-260)
// ------------------------------------------------------------------------------------------
constexpr int kCellSmem = 256; // heights per warp staged in shared memory for the gap scan

// in-place ascending sort of buf[0..n) by one warp (bitonic network for arbitrary n: the first
// step of every merge mirrors, so all exchanges move the larger value to the higher index and
// the virtual +inf padding beyond n is never touched)
template <typename T>
__device__ __forceinline__ void warp_sort(T* buf, std::uint32_t n)
{
    for (std::uint32_t k = 2; (k >> 1) < n; k <<= 1)
    {
        for (std::uint32_t t = lane_id(); t < n; t += 32)
        {
            const std::uint32_t u = t ^ (k - 1);
            if (u > t && u < n)
            {
                const T a = buf[t], b = buf[u];
                if (b < a)
                {
                    buf[t] = b;
                    buf[u] = a;
                }
            }
        }
        __syncwarp();
        for (std::uint32_t j = k >> 2; j > 0; j >>= 1)
        {
            for (std::uint32_t t = lane_id(); t < n; t += 32)
            {
                const std::uint32_t u = t ^ j;
                if (u > t && u < n)
                {
                    const T a = buf[t], b = buf[u];
                    if (b < a)
                    {
                        buf[t] = b;
                        buf[u] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// Bitonic sort of 32 * E values held E per lane (value index = e * 32 + lane): exchanges across
// lanes are shuffles, exchanges across the registers of a lane are plain moves; no shared memory.
template <int E, typename T>
__device__ __forceinline__ void warp_sort_regs(T (&v)[E])
{
    const std::uint32_t lane = lane_id();
#pragma unroll
    for (std::uint32_t k = 2; k <= 32u * E; k <<= 1)
    {
#pragma unroll
        for (std::uint32_t j = k >> 1; j > 0; j >>= 1)
        {
            if (j >= 32u)
            {
                const std::uint32_t je = j >> 5;
#pragma unroll
                for (std::uint32_t e = 0; e < static_cast<std::uint32_t>(E); ++e)
                {
                    const std::uint32_t pe = e ^ je;
                    if (pe > e)
                    {
                        const bool asc = (((e << 5) | lane) & k) == 0;
                        const T a = v[e], b = v[pe];
                        v[e] = asc ? min(a, b) : max(a, b);
                        v[pe] = asc ? max(a, b) : min(a, b);
                    }
                }
            }
            else
            {
#pragma unroll
                for (std::uint32_t e = 0; e < static_cast<std::uint32_t>(E); ++e)
                {
                    const T other = __shfl_xor_sync(0xffffffffu, v[e], j);
                    const bool asc = (((e << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    const bool take_min = lower == asc;
                    v[e] = take_min ? min(v[e], other) : max(v[e], other);
                }
            }
        }
    }
}

#ifndef LPL_CELL_WARPS
#define LPL_CELL_WARPS 8
#endif
constexpr int kCellWarps = LPL_CELL_WARPS; // warps per CTA
#ifndef LPL_CELL_PER_WARP
#define LPL_CELL_PER_WARP 16 // measured per 154-frame batch with the flat-cell shortcut: 4 -> 0.236 ms, 8 -> 0.206, 16 -> 0.197, 32 -> 0.230
#endif
constexpr int kCellsPerWarp = LPL_CELL_PER_WARP; // cells per warp, interleaved across the CTA's warps (most cells are empty)

// largest i in [1, n/2] with zs[i] - zs[i-1] > 0.5 on the ascending heights zs[0..n), scanned from
// the top by whole warps (segmenter.cpp:249-259); returns zs[i] or zs[0] when there is no such gap
template <class Load>
__device__ __forceinline__ float gap_scan(Load zs, std::uint32_t n)
{
    const std::uint32_t lane = lane_id();
    float zmin = zs(0);
    for (std::uint32_t hi = n / 2; hi >= 1;)
    {
        const bool valid = hi > lane;
        const std::uint32_t i = hi - lane;
        const bool hit = valid && (zs(i) - zs(i - 1) > 0.5f);
        const std::uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m != 0)
        {
            zmin = zs(hi - (__ffs(m) - 1));
            break;
        }
        if (hi <= 32)
        {
            break;
        }
        hi -= 32;
    }
    return zmin;
}

// heights are sorted as order-preserving integer keys (a total order: -0.0 before +0.0), one
// integer min / max per compare-exchange
__device__ __forceinline__ std::uint32_t zkey(std::uint32_t bits)
{
    return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}

__device__ __forceinline__ float zval(std::uint32_t key)
{
    return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

// shuffle-only bitonic network over the first KMAX lanes (ascending), one key per lane
template <int KMAX>
__device__ __forceinline__ std::uint32_t sort_lanes(std::uint32_t zk, std::uint32_t lane)
{
#pragma unroll
    for (std::uint32_t k = 2; k <= static_cast<std::uint32_t>(KMAX); k <<= 1)
    {
#pragma unroll
        for (std::uint32_t j = k >> 1; j > 0; j >>= 1)
        {
            const std::uint32_t other = __shfl_xor_sync(0xffffffffu, zk, j);
            const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
            zk = take_min ? min(zk, other) : max(zk, other);
        }
    }
    return zk;
}

// E values per lane sorted in registers, then the gap scan over shared memory
template <int E>
__device__ __forceinline__ float cell_zmin_regs(const uint2* zo, std::uint32_t n, float* zb)
{
    const std::uint32_t lane = lane_id();
    std::uint32_t z[E];
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        const std::uint32_t t = e * 32 + lane;
        z[e] = t < n ? zkey(zo[t].y) : 0xffffffffu;
    }
    // no gap can open below the median when the lower half of the heights spans <= 0.5 m (see cell_is_flat)
    {
        std::uint32_t kmin = z[0];
#pragma unroll
        for (int e = 1; e < E; ++e)
        {
            kmin = min(kmin, z[e]);
        }
        const float zlo = zval(__reduce_min_sync(0xffffffffu, kmin));
        std::uint32_t within = 0;
#pragma unroll
        for (int e = 0; e < E; ++e)
        {
            within += (static_cast<std::uint32_t>(e) * 32u + lane < n && zval(z[e]) - zlo <= 0.5f) ? 1u : 0u;
        }
        if (__reduce_add_sync(0xffffffffu, within) >= n / 2u + 1u)
        {
            return zlo;
        }
    }
    warp_sort_regs<E>(z);
#pragma unroll
    for (int e = 0; e < E; ++e)
    {
        zb[e * 32 + lane] = zval(z[e]);
    }
    __syncwarp();
    const float r = gap_scan([&](std::uint32_t i) { return zb[i]; }, n);
    __syncwarp();
    return r;
}

__device__ __forceinline__ void seg_cell_one(const Dev& d, const SegParams& sp, std::uint32_t f, std::uint32_t cell,
                                             std::uint32_t a, std::uint32_t n, float* zb)
{
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const uint2* zo = d.zo + o + a;
    float zmin;
    if (n <= 32)
    {
        // the common case: one height per lane, shuffle-only bitonic network
        const std::uint32_t lane = lane_id();
        std::uint32_t zk = lane < n ? zkey(zo[lane].y) : 0xffffffffu;
        // Flat cell (the common case: ground): the reference looks for the highest gap > 0.5 m among the sorted
        // heights zs[0 .. n/2]. Rounding is monotonic, so when n/2 + 1 heights satisfy z - zs[0] <= 0.5f (in float,
        // as the reference subtracts) every adjacent difference in that range is <= 0.5f as well: the answer is the
        // plain minimum and the sort is skipped. (A NaN or infinite height fails the comparison and takes the sort.)
        const float zlo = zval(__reduce_min_sync(0xffffffffu, zk));
        const std::uint32_t flat = __ballot_sync(0xffffffffu, lane < n && zval(zk) - zlo <= 0.5f);
        if (static_cast<std::uint32_t>(__popc(flat)) >= n / 2u + 1u)
        {
            zmin = zlo;
        }
        else
        {
        // most cells far from the sensor hold a handful of points: the network only runs up to the
        // next power of two of n (lanes beyond hold the +inf padding and sort among themselves)
        if (n <= 4)
        {
            zk = sort_lanes<4>(zk, lane);
        }
        else if (n <= 8)
        {
            zk = sort_lanes<8>(zk, lane);
        }
        else if (n <= 16)
        {
            zk = sort_lanes<16>(zk, lane);
        }
        else
        {
            zk = sort_lanes<32>(zk, lane);
        }
        const float z = zval(zk);
        const float prev = __shfl_up_sync(0xffffffffu, z, 1);
        const bool hit = lane >= 1 && lane <= n / 2 && (z - prev > 0.5f);
        const std::uint32_t m = __ballot_sync(0xffffffffu, hit);
        const int src_lane = m != 0 ? 31 - __clz(m) : 0;
        zmin = __shfl_sync(0xffffffffu, z, src_lane);
        }
    }
    else if (n <= 64)
    {
        zmin = cell_zmin_regs<2>(zo, n, zb);
    }
    else if (n <= 128)
    {
        zmin = cell_zmin_regs<4>(zo, n, zb);
    }
    else if (n <= kCellSmem)
    {
        zmin = cell_zmin_regs<8>(zo, n, zb);
    }
    else
    {
        // oversized cell (rare): the same bitonic network over global scratch; a warp's own
        // stores are visible to its lanes after __syncwarp
        std::uint32_t kmin = 0xffffffffu;
        for (std::uint32_t t = lane_id(); t < n; t += 32)
        {
            kmin = min(kmin, zkey(zo[t].y));
        }
        const float zlo = zval(__reduce_min_sync(0xffffffffu, kmin));
        std::uint32_t within = 0;
        for (std::uint32_t t = lane_id(); t < n; t += 32)
        {
            within += (__uint_as_float(zo[t].y) - zlo <= 0.5f) ? 1u : 0u;
        }
        if (__reduce_add_sync(0xffffffffu, within) >= n / 2u + 1u)
        {
            zmin = zlo; // flat cell, see above
        }
        else
        {
            float* zt = d.zsort + o + a;
            for (std::uint32_t t = lane_id(); t < n; t += 32)
            {
                zt[t] = __uint_as_float(zo[t].y);
            }
            __syncwarp();
            warp_sort(zt, n);
            zmin = gap_scan([&](std::uint32_t i) { return zt[i]; }, n);
        }
    }
    if (lane_id() == 0)
    {
        d.cell_zmin[static_cast<std::size_t>(f) * sp.ncell + cell] = zmin;
    }
}

__global__ void __launch_bounds__(kCellWarps * 32) k_seg_cell(Dev d, SegParams sp)
{
    __shared__ float sh[kCellWarps][kCellSmem];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t warp = threadIdx.x >> 5;
    // consecutive cells are radial neighbours of one slice (dense near the sensor, empty far out):
    // interleaving them over the warps keeps the warps of a CTA equally loaded. Lane j fetches the
    // bounds of the warp's j-th cell up front, so empty cells cost no dependent load.
    const std::uint32_t first = blockIdx.x * (kCellWarps * kCellsPerWarp) + warp;
    const std::uint32_t* cs = d.cell_start + static_cast<std::size_t>(f) * (sp.ncell + 1);
    std::uint32_t my_a = 0, my_n = 0;
    {
        const std::uint32_t cell = first + lane_id() * kCellWarps;
        if (lane_id() < kCellsPerWarp && cell < static_cast<std::uint32_t>(sp.ncell))
        {
            my_a = cs[cell];
            my_n = cs[cell + 1] - my_a;
        }
    }
#pragma unroll 1
    for (std::uint32_t j = 0; j < kCellsPerWarp; ++j)
    {
        const std::uint32_t a = __shfl_sync(0xffffffffu, my_a, j);
        const std::uint32_t n = __shfl_sync(0xffffffffu, my_n, j);
        if (n != 0)
        {
            seg_cell_one(d, sp, f, first + j * kCellWarps, a, n, sh[warp]);
        }
    }
}

// radial recurrence, one thread per azimuth slice (segmenter.cpp:215-267). The repeated
// `+ delta` is not re-associable bit-exactly, so it stays sequential (50 steps).
__global__ void __launch_bounds__(128) k_seg_elev(Dev d, SegParams sp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t s = blockIdx.x * 128u + threadIdx.x;
    if (s >= static_cast<std::uint32_t>(sp.slices))
    {
        return;
    }
    const std::uint32_t* cnt = d.cell_cnt + static_cast<std::size_t>(f) * sp.ncell;
    const float* zmin = d.cell_zmin + static_cast<std::size_t>(f) * sp.ncell;
    float* elev = d.elev + static_cast<std::size_t>(f) * sp.ncell;
    float prev = sp.e0;
    elev[s * sp.rings] = prev;
    for (int r = 1; r < sp.rings; ++r)
    {
        const int c = s * sp.rings + r;
        float e = prev + sp.delta;
        if (cnt[c] != 0)
        {
            e = fminf(zmin[c], e);
        }
        elev[c] = e;
        prev = e;
    }
}

// RANSAC candidates (segmenter.cpp:328-351), per point of the segmented cloud: counted per cell and copied
// densely. The obstacle classification itself (:271-283) is only ever observed for the pixel winners of the range
// image, so it is evaluated there (k_seg_px) and no per-point label plane exists.
__global__ void __launch_bounds__(256) k_seg_label(Dev d, SegParams sp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_in[f];
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::int32_t ccell = -1; // cell of a candidate, -1 otherwise
    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n)
    {
        const std::int32_t c = d.cell[o + i];
        pt = d.pts_in[o + i]; // in flight together with the cell index (only the elevation look-up depends on it)
        if (c >= 0 && (c % sp.rings) < kRansacBins)
        {
            const float z = pt.z;
            const float e = d.elev[static_cast<std::size_t>(f) * sp.ncell + c];
            if (fabsf(e - z) < sp.thr2)
            {
                ccell = c;
            }
        }
    }
    const std::uint32_t peers = __match_any_sync(0xffffffffu, ccell);
    if (ccell >= 0 && static_cast<int>(lane_id()) == __ffs(peers) - 1)
    {
        atomicAdd(&d.ccnt[static_cast<std::size_t>(f) * sp.ncell + ccell], __popc(peers));
    }
    // dense (unordered) copy of the candidates for the inlier count, one atomic per CTA
    __shared__ std::uint32_t s_cnt, s_base;
    if (threadIdx.x == 0)
    {
        s_cnt = 0;
    }
    __syncthreads();
    const std::uint32_t cm = __ballot_sync(0xffffffffu, ccell >= 0);
    std::uint32_t woff = 0;
    if (lane_id() == 0 && cm != 0)
    {
        woff = atomicAdd(&s_cnt, static_cast<std::uint32_t>(__popc(cm)));
    }
    woff = __shfl_sync(0xffffffffu, woff, 0);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt != 0)
    {
        s_base = atomicAdd(&d.n_cpts[f], s_cnt);
    }
    __syncthreads();
    if (ccell >= 0)
    {
        d.cpts[o + s_base + woff + __popc(cm & ((1u << lane_id()) - 1u))] = pt;
    }
}

// ------------------------------------------------------------------------------------------
// near-field RANSAC (segmenter.cpp:321-479)
// ------------------------------------------------------------------------------------------
// std::mt19937{42} raw outputs are frame independent (pre-generated on the host); libstdc++'s
// uniform_int_distribution<uint32_t>{0, n-1} maps them with Lemire's multiply-shift + rejection.
template <class RawAt>
__device__ __forceinline__ std::uint32_t mt_draw(RawAt raw_at, std::uint32_t& pos, std::uint32_t n, bool& exhausted)
{
    auto next = [&]() -> std::uint32_t {
        if (pos >= kMtRaws)
        {
            exhausted = true;
            return 0u;
        }
        return raw_at(pos++);
    };
    unsigned long long product = static_cast<unsigned long long>(next()) * n;
    std::uint32_t low = static_cast<std::uint32_t>(product);
    if (low < n)
    {
        const std::uint32_t threshold = (0u - n) % n;
        while (low < threshold && !exhausted)
        {
            product = static_cast<unsigned long long>(next()) * n;
            low = static_cast<std::uint32_t>(product);
        }
    }
    return static_cast<std::uint32_t>(product >> 32);
}

constexpr int kDrawThreads = 256;
constexpr int kSelCap = 1024;    // candidate indices of one cell staged per warp
constexpr int kRawStage = 1024;  // generator outputs staged in shared memory

// The reference numbers its candidates in (slice, bin, cloud) order and draws 60 index pairs.
// Only those <= 120 candidates are ever looked at by position, so instead of materialising the
// ordered list the draws are resolved by rank: an exclusive prefix of the per-cell candidate
// counts (cells in index order, bins 0..3 only) locates the cell of rank t (k_ransac_draw), and a
// radix select over the point indices of that cell's candidates finds the (t - prefix)-th in
// cloud order (k_ransac_plane, one warp per draw).
__device__ __forceinline__ std::uint32_t ransac_cell_of(const SegParams& sp, std::uint32_t j)
{
    return (j / sp.nb) * sp.rings + (j % sp.nb);
}

__global__ void __launch_bounds__(kDrawThreads) k_ransac_draw(Dev d, SegParams sp)
{
    __shared__ std::uint32_t sh[33];
    __shared__ std::uint32_t raws[kRawStage];
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::uint32_t J = static_cast<std::uint32_t>(sp.slices) * sp.nb; // candidate cells in rank order
    std::uint32_t* ccnt = d.ccnt + static_cast<std::size_t>(f) * sp.ncell;
    for (std::uint32_t t = threadIdx.x; t < kRawStage; t += kDrawThreads)
    {
        raws[t] = d.mt_raw[t];
    }
    // in-place exclusive prefix over the J candidate cells, a contiguous chunk per thread
    const std::uint32_t chunk = (J + kDrawThreads - 1) / kDrawThreads;
    const std::uint32_t j0 = min(threadIdx.x * chunk, J), j1 = min(j0 + chunk, J);
    std::uint32_t part = 0;
    for (std::uint32_t j = j0; j < j1; ++j)
    {
        part += ccnt[ransac_cell_of(sp, j)];
    }
    std::uint32_t total;
    std::uint32_t run = block_excl_scan(part, sh, &total);
    for (std::uint32_t j = j0; j < j1; ++j)
    {
        const std::uint32_t c = ransac_cell_of(sp, j);
        const std::uint32_t v = ccnt[c];
        ccnt[c] = run;
        run += v;
    }
    const std::uint32_t nc = total;
    if (threadIdx.x < kRansacIters)
    {
        d.inliers[f * kRansacIters + threadIdx.x] = 0;
    }
    // a single candidate makes the reference spin forever (segmenter.cpp:382-386); RANSAC is
    // skipped for nc < 2 (documented deviation).
    if (threadIdx.x == 0)
    {
        d.n_cand[f] = nc;
        if (nc >= 2)
        {
            std::uint32_t pos = 0;
            bool exhausted = false;
            auto raw_at = [&](std::uint32_t k) -> std::uint32_t { return k < kRawStage ? raws[k] : d.mt_raw[k]; };
            std::uint32_t* pairs = d.pairs + static_cast<std::size_t>(f) * kRansacIters * 2;
            for (int it = 0; it < kRansacIters; ++it)
            {
                const std::uint32_t i2 = mt_draw(raw_at, pos, nc, exhausted);
                std::uint32_t i3 = mt_draw(raw_at, pos, nc, exhausted);
                while (i3 == i2 && !exhausted)
                {
                    i3 = mt_draw(raw_at, pos, nc, exhausted);
                }
                pairs[2 * it] = i2;
                pairs[2 * it + 1] = i3;
            }
            if (exhausted)
            {
                atomicOr(&d.status[f], ST_RNG_EXHAUSTED);
            }
        }
    }
}

// one CTA per (draw pair, frame): warp w resolves draw w of the pair, then the plane is formed
__global__ void __launch_bounds__(64) k_ransac_plane(Dev d, SegParams sp)
{
    __shared__ std::uint32_t sel[2][kSelCap];
    __shared__ std::uint32_t pidx[2];
    const std::uint32_t it = blockIdx.x, f = blockIdx.y + d.f0;
    const std::uint32_t nc = d.n_cand[f];
    const bool runit = nc >= 2;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    if (runit)
    {
        const std::uint32_t J = static_cast<std::uint32_t>(sp.slices) * sp.nb;
        const std::uint32_t* ccnt = d.ccnt + static_cast<std::size_t>(f) * sp.ncell;
        const std::uint32_t* cs = d.cell_start + static_cast<std::size_t>(f) * (sp.ncell + 1);
        const std::uint32_t warp = threadIdx.x >> 5, lane = lane_id();
        std::uint32_t* mine = sel[warp];
        const std::uint32_t t = d.pairs[(static_cast<std::size_t>(f) * kRansacIters + it) * 2 + warp];
        // last candidate cell j with prefix[j] <= t (cells without candidates repeat the prefix)
        std::uint32_t lo = 0, hi = J;
        while (hi - lo > 1)
        {
            const std::uint32_t mid = (lo + hi) >> 1;
            if (ccnt[ransac_cell_of(sp, mid)] <= t)
            {
                lo = mid;
            }
            else
            {
                hi = mid;
            }
        }
        const std::uint32_t c = ransac_cell_of(sp, lo);
        std::uint32_t r = t - ccnt[c]; // rank inside the cell, cloud order
        const std::uint32_t a = cs[c], n = cs[c + 1] - a;
        const uint2* zo = d.zo + o + a;
        const float elev_c = d.elev[static_cast<std::size_t>(f) * sp.ncell + c];
        // stage the candidates' point indices
        std::uint32_t m = 0;
        for (std::uint32_t base = 0; base < n; base += 32)
        {
            const std::uint32_t u = base + lane;
            std::uint32_t idx = 0;
            bool is = false;
            if (u < n)
            {
                const uint2 rec = zo[u];
                idx = rec.x;
                is = fabsf(elev_c - __uint_as_float(rec.y)) < sp.thr2; // the candidate test of k_seg_label
            }
            const std::uint32_t b = __ballot_sync(0xffffffffu, is);
            const std::uint32_t w = m + __popc(b & ((1u << lane) - 1u));
            if (is && w < kSelCap)
            {
                mine[w] = idx;
            }
            m += __popc(b);
        }
        __syncwarp();
        std::uint32_t prefix = 0;
        if (m <= kSelCap)
        {
            for (int bit = sp.idx_bits - 1; bit >= 0; --bit)
            {
                std::uint32_t cnt0 = 0;
                for (std::uint32_t u = lane; u < m; u += 32)
                {
                    const std::uint32_t v = mine[u];
                    cnt0 += ((v >> (bit + 1)) == (prefix >> (bit + 1)) && ((v >> bit) & 1u) == 0u) ? 1u : 0u;
                }
                cnt0 = warp_sum(cnt0);
                if (r >= cnt0)
                {
                    r -= cnt0;
                    prefix |= 1u << bit;
                }
            }
        }
        else
        {
            // more candidates in one cell than the staging holds: select straight from memory
            for (int bit = sp.idx_bits - 1; bit >= 0; --bit)
            {
                std::uint32_t cnt0 = 0;
                for (std::uint32_t u = lane; u < n; u += 32)
                {
                    const uint2 rec = zo[u];
                    const std::uint32_t v = rec.x;
                    if (fabsf(elev_c - __uint_as_float(rec.y)) < sp.thr2 && (v >> (bit + 1)) == (prefix >> (bit + 1)) &&
                        ((v >> bit) & 1u) == 0u)
                    {
                        cnt0 += 1;
                    }
                }
                cnt0 = warp_sum(cnt0);
                if (r >= cnt0)
                {
                    r -= cnt0;
                    prefix |= 1u << bit;
                }
            }
        }
        if (lane == 0)
        {
            pidx[warp] = prefix;
        }
    }
    __syncthreads();
    if (threadIdx.x != 0)
    {
        return;
    }
    float4 plane = make_float4(0.f, 0.f, __int_as_float(0x7fc00000), 0.f); // skipped
    if (runit)
    {
        const float4 p2 = d.pts_in[o + pidx[0]];
        const float4 p3 = d.pts_in[o + pidx[1]];
        const float p1x = 0.0f, p1y = 0.0f, p1z = sp.p1z;
        float nx = ((p2.y - p1y) * (p3.z - p1z)) - ((p2.z - p1z) * (p3.y - p1y));
        float ny = ((p2.z - p1z) * (p3.x - p1x)) - ((p2.x - p1x) * (p3.z - p1z));
        float nz = ((p2.x - p1x) * (p3.y - p1y)) - ((p2.y - p1y) * (p3.x - p1x));
        const float den = (nx * nx) + (ny * ny) + (nz * nz);
        if (!(den < 1.0e-5f))
        {
            const float norm = 1.0f / sqrtf(den);
            nz *= norm;
            if (!(fabsf(nz) < sp.cos_max))
            {
                nx *= norm;
                ny *= norm;
                const float pd = (nx * p1x) + (ny * p1y) + (nz * p1z);
                plane = make_float4(nx, ny, nz, pd);
            }
        }
    }
    d.planes[f * kRansacIters + it] = plane;
}

#ifndef LPL_RANSAC_PER
#define LPL_RANSAC_PER 16 // candidates per thread; measured per 154-frame batch: 4 -> 0.176 ms, 8 -> 0.166, 16 -> 0.147, 32 -> 0.195
#endif
constexpr int kRansacPer = LPL_RANSAC_PER; // candidates per thread: one plane fetch serves all of them

__global__ void __launch_bounds__(256) k_ransac_count(Dev d, SegParams sp)
{
    __shared__ float4 pl[kRansacIters];
    __shared__ std::uint32_t cnt[kRansacIters];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nc = d.n_cpts[f]; // = n_cand, the dense copy is unordered
    const std::uint32_t base = blockIdx.x * (256u * kRansacPer);
    if (nc < 2 || base >= nc)
    {
        return;
    }
    if (threadIdx.x < kRansacIters)
    {
        pl[threadIdx.x] = d.planes[f * kRansacIters + threadIdx.x];
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    float4 p[kRansacPer];
    bool live[kRansacPer];
#pragma unroll
    for (int j = 0; j < kRansacPer; ++j)
    {
        const std::uint32_t k = base + j * 256u + threadIdx.x;
        live[j] = k < nc;
        p[j] = live[j] ? d.cpts[static_cast<std::size_t>(f) * d.cap + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // lane j keeps the counts of planes j and j + 32 of its warp in registers
    std::uint32_t c0 = 0, c1 = 0;
    const std::uint32_t lane = lane_id();
#pragma unroll 4
    for (int it = 0; it < kRansacIters; ++it)
    {
        const float4 q = pl[it];
        if (q.z != q.z)
        {
            continue; // skipped draw (uniform branch)
        }
        // per-thread count over its candidates, then one warp reduction per plane (redux.sync) instead of a ballot +
        // popc per candidate
        std::uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < kRansacPer; ++j)
        {
            const float od = fabsf((q.x * p[j].x) + (q.y * p[j].y) + (q.z * p[j].z) - q.w);
            mine += (live[j] && od < sp.thr) ? 1u : 0u;
        }
        const std::uint32_t hits = __reduce_add_sync(0xffffffffu, mine);
        if (static_cast<int>(lane) == (it & 31))
        {
            if (it < 32)
            {
                c0 += hits;
            }
            else
            {
                c1 += hits;
            }
        }
    }
    if (c0 != 0)
    {
        atomicAdd(&cnt[lane], c0);
    }
    if (c1 != 0 && lane + 32 < kRansacIters)
    {
        atomicAdd(&cnt[lane + 32], c1);
    }
    __syncthreads();
    if (threadIdx.x < kRansacIters && cnt[threadIdx.x] != 0)
    {
        atomicAdd(&d.inliers[f * kRansacIters + threadIdx.x], cnt[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------
// plane application + range-image scatter (segmenter.cpp:434-477, 291-318)
// The reference walks cells in index order and the points of a cell in cloud order and replaces a
// pixel only on a strictly smaller depth^2, so among equal depths the first in (cell, cloud) order
// wins. Equal depth^2 means equal radial bin, hence cell order = azimuth-slice order: the key
// (depth^2 bits, slice, point index) under atomicMin reproduces the winner without any sorted list.
// ------------------------------------------------------------------------------------------
// best plane of a frame (segmenter.cpp:434-453): the first iteration with the strictly largest inlier
// count, flipped so that c >= 0. One warp per frame, once - not once per CTA of k_seg_px.
__global__ void __launch_bounds__(32) k_ransac_best(Dev d)
{
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::uint32_t lane = lane_id();
    unsigned long long key = 0; // (count << 8) | (255 - iteration): maximum = largest count, earliest iteration
    if (d.n_cand[f] >= 2)
    {
        for (std::uint32_t it = lane; it < static_cast<std::uint32_t>(kRansacIters); it += 32)
        {
            const std::uint32_t c = d.inliers[f * kRansacIters + it];
            const unsigned long long k = (static_cast<unsigned long long>(c) << 8) | (255u - it);
            key = (c != 0 && k > key) ? k : key;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
    {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, s);
        key = other > key ? other : key;
    }
    if (lane == 0)
    {
        float4 pl = make_float4(0.f, 0.f, 1.f, 0.f);
        const std::uint32_t best = static_cast<std::uint32_t>(key >> 8);
        if (best != 0)
        {
            pl = d.planes[f * kRansacIters + (255u - static_cast<std::uint32_t>(key & 0xffu))];
            if (pl.z < 0.f)
            {
                pl = make_float4(-pl.x, -pl.y, -pl.z, -pl.w);
            }
        }
        d.best_plane[f] = pl;
        d.best_cnt[f] = best; // 0 = no plane was accepted
    }
}

__global__ void __launch_bounds__(256) k_seg_px(Dev d, SegParams sp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t p = blockIdx.x * 256u + threadIdx.x;
    if (p >= static_cast<std::uint32_t>(sp.npx))
    {
        return;
    }
    const std::size_t po = static_cast<std::size_t>(f) * sp.npx;
    const unsigned long long key = d.key[po + p];
    std::uint8_t c = PX_EMPTY;
    std::int32_t widx = -1;
    if (key != ~0ULL)
    {
        const std::size_t o = static_cast<std::size_t>(f) * d.cap;
        const std::uint32_t i = static_cast<std::uint32_t>(key & ((1ULL << sp.idx_bits) - 1ULL));
        const float4 q = d.pts_in[o + i];
        // the winner's polar cell without a gather: the azimuth slice rides in the key, the radial bin follows from the
        // coordinates exactly as in k_seg_bin
        const std::int32_t slice = static_cast<std::int32_t>((key & ((1ULL << 33) - 1ULL)) >> sp.idx_bits);
        const std::int32_t cell = slice * sp.rings + static_cast<std::int32_t>(sqrtf(q.x * q.x + q.y * q.y) / sp.radial_spacing);
        widx = static_cast<std::int32_t>(i);
        // obstacle classification against the cell's elevation (segmenter.cpp:271-283), then the RANSAC plane over the
        // near-field bins (:455-477)
        const float e = d.elev[static_cast<std::size_t>(f) * sp.ncell + cell];
        c = (q.z >= e + sp.thr) ? PX_OBSTACLE : PX_GROUND;
        if (d.best_cnt[f] != 0 && (cell % sp.rings) < kRansacBins)
        {
            const float4 pl = d.best_plane[f];
            const float sd = (pl.x * q.x) + (pl.y * q.y) + (pl.z * q.z) - pl.w;
            if (sd < sp.thr)
            {
                c = PX_GROUND;
            }
        }
    }
    d.pxidx[po + p] = widx; // the range image proper: the winner's index (k_jcp_pre gathers coordinates through it, k_seg_labels_out labels through it)
    d.code[po + p] = c;
}

// ------------------------------------------------------------------------------------------
// 5x5 in-image dilation of the obstacle mask + queue flags (segmenter.cpp:486-514).
// One CTA per 16 x 128 pixel tile; the tile and its 2-pixel halo are staged in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int kDilTh = 16, kDilTw = 128;
constexpr int kDilBoxW = 160; // tile + 16 columns on both sides: TMA wants the innermost box coordinate (in bytes)
                              // and extent to be multiples of 16, the stencil only needs 2 of those 16 columns

// vertical 5-tap maximum over the horizontal maxima + the queue / dilated flags (segmenter.cpp:491-514)
template <int kPitch>
__device__ __forceinline__ void dilate_finish(const Dev& d, const SegParams& sp, std::uint32_t f, int h0, int w0,
                                              const std::uint8_t (*hmax)[kPitch])
{
    std::uint8_t* code = d.code + static_cast<std::size_t>(f) * sp.npx;
    for (int t = threadIdx.x; t < kDilTh * kDilTw; t += 256)
    {
        const int r = t / kDilTw, c = t % kDilTw;
        const int h = h0 + r, w = w0 + c;
        if (h >= sp.H || w >= sp.W)
        {
            continue;
        }
        const std::uint8_t dil = hmax[r][c] | hmax[r + 1][c] | hmax[r + 2][c] | hmax[r + 3][c] | hmax[r + 4][c];
        if (dil)
        {
            const std::uint8_t cur = code[h * sp.W + w];
            if (cur == PX_GROUND)
            {
                code[h * sp.W + w] = PX_QUEUED;
            }
            else if (cur == PX_EMPTY)
            {
                code[h * sp.W + w] = PX_EMPTY | PX_DILATED;
            }
        }
    }
}

// Tile + halo staged by TMA: one elected thread issues a 3-D tiled bulk copy (x = column, y = row,
// z = frame) of the (16 + 4) x 160-byte box whose corner lies two rows up and 16 columns left of the tile;
// the tensor map's out-of-bounds fill supplies the zeros beyond the image borders, which is exactly
// the reference's in-bounds 5x5 maximum (cv::dilate over the image rectangle). Neighbouring tiles
// rewrite GROUND -> QUEUED / EMPTY -> DILATED in the same plane while this box is read, but never an
// OBSTACLE pixel, and only (code & 0xf) == OBSTACLE is looked at.
__global__ void __launch_bounds__(256) k_seg_dilate_tma(Dev d, SegParams sp, const __grid_constant__ CUtensorMap code_map)
{
    __shared__ __align__(128) std::uint8_t raw[kDilTh + 4][kDilBoxW];
    __shared__ std::uint8_t hmax[kDilTh + 4][kDilTw + 4];
    __shared__ __align__(8) unsigned long long mbar;
    const std::uint32_t f = blockIdx.z + d.f0;
    const int h0 = blockIdx.y * kDilTh, w0 = blockIdx.x * kDilTw;
    const std::uint32_t mbar_s = static_cast<std::uint32_t>(__cvta_generic_to_shared(&mbar));
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        constexpr std::uint32_t kBytes = (kDilTh + 4) * kDilBoxW;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(kBytes) : "memory");
        const std::uint32_t dst = static_cast<std::uint32_t>(__cvta_generic_to_shared(&raw[0][0]));
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<unsigned long long>(&code_map)), "r"(mbar_s), "r"(w0 - 16), "r"(h0 - 2),
            "r"(static_cast<int>(f))
            : "memory");
    }
    // every thread waits for the transaction bytes (phase 0)
    {
        std::uint32_t done = 0;
        while (done == 0)
        {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(mbar_s)
                : "memory");
        }
    }
    // separable maximum: horizontal 5 taps on (kDilTh + 4) rows, then vertical 5 taps
    for (int t = threadIdx.x; t < (kDilTh + 4) * kDilTw; t += 256)
    {
        const int r = t / kDilTw, c = t % kDilTw;
        std::uint8_t m = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k)
        {
            m |= ((raw[r][c + 14 + k] & 0xf) == PX_OBSTACLE) ? 1 : 0; // column w0 + c - 2 + k
        }
        hmax[r][c] = m;
    }
    __syncthreads();
    dilate_finish(d, sp, f, h0, w0, hmax);
}

// same stencil with the tile staged by plain loads: image widths that are not a multiple of 16 bytes
// cannot be described by a tensor map (global strides must be)
__global__ void __launch_bounds__(256) k_seg_dilate(Dev d, SegParams sp)
{
    __shared__ std::uint8_t tile[kDilTh + 4][kDilTw + 4 + 4];
    __shared__ std::uint8_t hmax[kDilTh + 4][kDilTw + 4];
    const std::uint32_t f = blockIdx.z + d.f0;
    const int h0 = blockIdx.y * kDilTh, w0 = blockIdx.x * kDilTw;
    const std::uint8_t* code = d.code + static_cast<std::size_t>(f) * sp.npx;
    for (int t = threadIdx.x; t < (kDilTh + 4) * (kDilTw + 4); t += 256)
    {
        const int r = t / (kDilTw + 4), c = t % (kDilTw + 4);
        const int h = h0 + r - 2, w = w0 + c - 2;
        std::uint8_t v = 0;
        if (h >= 0 && h < sp.H && w >= 0 && w < sp.W)
        {
            v = ((code[h * sp.W + w] & 0xf) == PX_OBSTACLE) ? 1 : 0;
        }
        tile[r][c] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (kDilTh + 4) * kDilTw; t += 256)
    {
        const int r = t / kDilTw, c = t % kDilTw;
        hmax[r][c] = tile[r][c] | tile[r][c + 1] | tile[r][c + 2] | tile[r][c + 3] | tile[r][c + 4];
    }
    __syncthreads();
    dilate_finish(d, sp, f, h0, w0, hmax);
}

// Tensor map of the pixel-code planes [B][H][W] (uint8) with a (kDilTh + 4) x kDilBoxW box. The driver
// entry point is looked up at run time (no link-time dependency on libcuda).
bool make_code_map(Ctx* c)
{
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess)
    {
        return false;
    }
    const SegParams& sp = c->seg;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(sp.W), static_cast<cuuint64_t>(sp.H), static_cast<cuuint64_t>(c->d.B)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(sp.W), static_cast<cuuint64_t>(sp.W) * sp.H};
    const cuuint32_t box[3] = {kDilBoxW, kDilTh + 4, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = reinterpret_cast<Encode>(fn)(&c->code_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c->d.code, dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct QueuePred
{
    const std::uint8_t* code;
    std::uint32_t npx;
    __device__ bool operator()(std::uint32_t f, std::uint32_t p) const
    {
        return code[static_cast<std::size_t>(f) * npx + p] == PX_QUEUED;
    }
};

struct QueueEmit
{
    std::uint32_t* queue;
    std::uint32_t* status;
    std::uint32_t qcap;
    __device__ void operator()(std::uint32_t f, std::uint32_t p, std::uint32_t pos) const
    {
        if (pos < qcap)
        {
            queue[static_cast<std::size_t>(f) * qcap + pos] = p;
        }
        else
        {
            atomicOr(&status[f], ST_QUEUE_OVERFLOW);
        }
    }
};

// ------------------------------------------------------------------------------------------
// JCP pre-pass: weights and mask sources for every queued pixel (segmenter.cpp:543-607)
// mask source per slot (2 bits): 0 unknown, 1 ground, 2 obstacle, 3 = final state of a queued
// pixel that precedes this one in raster order (slots 0..11 only).
// mk bit 48: the pixel can be decided (|sum| > FLT_EPSILON); bits 49..63: 1 + row of stale_ref.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void jcp_slot(const Dev& d, const SegParams& sp, std::size_t po, std::size_t o, const float4& core,
                                         int hh, int ww, int i, float& wgt, std::uint32_t& msk)
{
    const std::uint32_t np = static_cast<std::uint32_t>(hh * sp.W + ww);
    // the range image holds point indices; the coordinates come from the cloud itself (a per-pixel copy of them cost
    // 323 MB of writes per batch for the ~7 % of the pixels whose neighbourhoods are ever read)
    const std::int32_t qi = d.pxidx[po + np];
    wgt = 0.f;
    msk = 0;
    if (qi < 0)
    {
        return;
    }
    const float4 q = d.pts_in[o + static_cast<std::uint32_t>(qi)];
    const float dx = core.x - q.x;
    const float dy = core.y - q.y;
    const float dz = core.z - q.z;
    const float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > sp.kthr_sqr)
    {
        return;
    }
    wgt = expf_glibc(-sp.amp * sqrtf(d2));
    const std::uint8_t c = d.code[po + np] & 0xf;
    if (c == PX_GROUND)
    {
        msk = 1;
    }
    else if (c == PX_OBSTACLE)
    {
        msk = 2;
    }
    else if (c == PX_QUEUED && i < 12)
    {
        msk = 3;
    }
}

#ifndef LPL_JCP_PRE_UNROLL
#define LPL_JCP_PRE_UNROLL 4
#endif
constexpr int kJcpPreUnroll = LPL_JCP_PRE_UNROLL; // neighbour slots of a queued pixel evaluated per loop trip
__global__ void __launch_bounds__(128) k_jcp_pre(Dev d, SegParams sp)
{
    __shared__ float s_w[128][25]; // raw weights of the CTA's 128 queued pixels (+1 pad: no bank conflicts)
    __shared__ float s_inv[128];   // the divisor (sum) or 0 when the pixel cannot be decided
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nq = min(d.n_queue[f], d.qcap);
    const std::size_t po = static_cast<std::size_t>(f) * sp.npx;
    const std::uint32_t* queue = d.queue + static_cast<std::size_t>(f) * d.qcap;
    // the grid covers a typical queue (~11k entries per HDL-64E frame) in one trip; longer queues loop
    for (std::uint32_t blk = blockIdx.x; blk * 128u < nq; blk += gridDim.x)
    {
    const std::uint32_t k = min(blk * 128u + threadIdx.x, nq - 1u); // tail threads redo the last pixel
    const bool live = blk * 128u + threadIdx.x < nq;
    const std::uint32_t p = queue[k];
    const int h = static_cast<int>(p / sp.W), w = static_cast<int>(p % sp.W);
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const float4 core = d.pts_in[o + static_cast<std::uint32_t>(d.pxidx[po + p])]; // a queued pixel has a winner
    float* wn = d.wn + static_cast<std::size_t>(f) * 24 * d.qcap;
    unsigned long long mk = 0;
    std::uint32_t brow = 0; // 1 + stale_ref row once allocated
    float sum = 0.f;
#pragma unroll kJcpPreUnroll
    for (int i = 0; i < 24; ++i)
    {
        const int hh = h + c_off_h[i], ww = w + c_off_w[i];
        float wgt = 0.f;
        std::uint32_t msk = 0;
        if (hh >= 0 && hh < sp.H && ww >= 0 && ww < sp.W)
        {
            jcp_slot(d, sp, po, o, core, hh, ww, i, wgt, msk);
            if (wgt != 0.f)
            {
                sum += wgt;
            }
        }
        else if (sp.jcp_emulate_stale)
        {
            // most recent earlier queued pixel for which slot i lies inside the image
            for (std::uint32_t kk = k; kk-- > 0;)
            {
                const std::uint32_t pp = queue[kk];
                const int h2 = static_cast<int>(pp / sp.W), w2 = static_cast<int>(pp % sp.W);
                const int hh2 = h2 + c_off_h[i], ww2 = w2 + c_off_w[i];
                if (hh2 >= 0 && hh2 < sp.H && ww2 >= 0 && ww2 < sp.W)
                {
                    jcp_slot(d, sp, po, o, d.pts_in[o + static_cast<std::uint32_t>(d.pxidx[po + pp])], hh2, ww2, i, wgt, msk);
                    if (msk == 3)
                    {
                        if (brow == 0 && live)
                        {
                            const std::uint32_t r = atomicAdd(&d.n_border[f], 1u);
                            if (r < d.nborder_cap)
                            {
                                brow = r + 1;
                            }
                            else
                            {
                                atomicOr(&d.status[f], ST_BORDER_OVERFLOW);
                            }
                        }
                        if (brow != 0)
                        {
                            d.stale_ref[(static_cast<std::size_t>(f) * d.nborder_cap + (brow - 1)) * 12 + i] =
                                static_cast<std::uint32_t>(hh2 * sp.W + ww2);
                        }
                        else
                        {
                            msk = 0;
                        }
                    }
                    break;
                }
                if (hh2 < 0)
                {
                    break; // every earlier queued pixel sits in the same or a lower row
                }
            }
        }
        s_w[threadIdx.x][i] = wgt; // raw weight, normalised below
        mk |= static_cast<unsigned long long>(msk) << (2 * i);
    }
    const bool decidable = fabsf(sum) > FLT_EPSILON;
    s_inv[threadIdx.x] = decidable ? sum : 0.f;
    mk |= static_cast<unsigned long long>(decidable ? 1 : 0) << 48;
    mk |= static_cast<unsigned long long>(brow) << 49;
    if (live)
    {
        d.mk[static_cast<std::size_t>(f) * d.qcap + k] = mk;
    }
    __syncthreads();
    // weight_matrix = unnormalized / sum (segmenter.cpp:611); the CTA's rows are one contiguous
    // block of the entry-major array -> coalesced stores
    const std::uint32_t rows = min(128u, nq - blk * 128u);
    float* out = wn + static_cast<std::size_t>(blk) * 128u * 24u;
    for (std::uint32_t t = threadIdx.x; t < rows * 24u; t += 128u)
    {
        const std::uint32_t e = t / 24u, i = t % 24u;
        const float dv = s_inv[e];
        out[t] = dv != 0.f ? s_w[e][i] / dv : 0.f;
    }
    __syncthreads(); // s_w / s_inv are rewritten by the next trip
    }
}

// 2-bit state plane of a frame in shared memory: 0 unknown / empty / undecided, 1 ground, 2 obstacle, 3 queued and not yet swept
__device__ __forceinline__ std::uint32_t plane_get(const volatile std::uint32_t* plane, std::uint32_t p)
{
    return (plane[p >> 4] >> ((p & 15u) * 2u)) & 3u;
}


// ------------------------------------------------------------------------------------------
// Row-synchronous JCP sweep, one THREAD per queued pixel.
// Within an image row a queued pixel depends only on its two left neighbours (kernel slots 10 / 11):
// every other dynamic slot - in-image slots of the two rows above and the inherited (stale) slots of
// border pixels, which always refer to rows above - is final once the rows are swept top to bottom.
//   vote:  a thread reads the rows-above slots of its pixel from the state plane and evaluates the vote
//     for every outcome of the two left neighbours at once. The ground sum depends only on [left-2 is
//     ground], [left-1 is ground] and the obstacle sum on [.. is obstacle]: four variants each, every
//     one the reference's exact 24-term chain in slot order (x + 0.0f == x), folded into a 9-bit table
//     (state of left-2, state of left-1) -> outcome.
//   chain: the raster recurrence along the row is a composition of finite maps. With the state
//     X = (outcome of the previous-but-one entry, outcome of the previous entry) in {0,1,2}^2, entry q
//     is a map X -> X' (9 x 4 bits); a warp-level inclusive scan composes the maps of 32 consecutive
//     entries, the warps' totals are chained through shared memory, and every thread reads its outcome
//     off its prefix map. No sequential walk, no polling.
// A row is processed in chunks of kJcpRowsThreads queue entries: one CTA barrier per chunk plus one at
// the end of the row; the queue records of the next chunk are loaded into registers before the
// current chunk is voted on. Reference: the queue loop of Segmenter::JCP (segmenter.cpp:535-637).
// ------------------------------------------------------------------------------------------
#ifndef LPL_JCP_ROWS_THREADS
#define LPL_JCP_ROWS_THREADS 256 // measured per 154-frame batch: 128 -> 0.43 ms, 256 -> 0.33, 384 -> 0.37, 512 -> 0.42
#endif
constexpr int kJcpRowsThreads = LPL_JCP_ROWS_THREADS;
#ifndef LPL_JCP_ROWS_MINB
#define LPL_JCP_ROWS_MINB 2 // CTAs per SM: a 154-frame batch must not spill into a second wave on 148 SMs
#endif
// A map on the nine states X = 3 * (outcome of entry q-2) + (outcome of entry q-1) is nine bytes: byte x of (r0, r1, r2)
// is the image of state x. Composition is a table look-up of nine indices, which is what PRMT does: the indices of the
// earlier map, packed four bits each, are the selector of a byte permute over the later map's bytes 0..7; index 8 sets
// the selector's "replicate the sign" bit, which yields 0x00 from the table (all entries < 0x80) and 0xff from a
// register of 0x80 bytes - the mask that lets the ninth entry in. ~16 instructions per composition (the 4-bit-packed
// form of round 1 shifted and masked its way through nine entries: ~60, a third of the kernel's instructions).
struct JcpMap
{
    std::uint32_t r0, r1, r2;
};

__device__ __forceinline__ std::uint32_t prmt(std::uint32_t a, std::uint32_t b, std::uint32_t sel)
{
    std::uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

__device__ __forceinline__ JcpMap jcp_identity()
{
    return JcpMap{0x03020100u, 0x07060504u, 0x00000008u};
}

// selector form of a map: n0 = the images of the states 0..7, four bits each; n1 = the image of state 8
__device__ __forceinline__ void jcp_pack(const JcpMap& m, std::uint32_t& n0, std::uint32_t& n1)
{
    const std::uint32_t t0 = m.r0 | (m.r0 >> 4), t1 = m.r1 | (m.r1 >> 4);
    n0 = prmt(t0, t1, 0x6420u);
    n1 = m.r2;
}

// (later o earlier)[x] = later[earlier[x]]; the earlier map comes in selector form (e0, e1)
__device__ __forceinline__ JcpMap jcp_compose(const JcpMap& later, std::uint32_t e0, std::uint32_t e1)
{
    const std::uint32_t l8 = prmt(later.r2, 0u, 0x0000u); // later[8] in every byte
    const std::uint32_t k80 = 0x80808080u;
    JcpMap r;
    r.r0 = prmt(later.r0, later.r1, e0) | (prmt(k80, k80, e0) & l8);
    r.r1 = prmt(later.r0, later.r1, e0 >> 16) | (prmt(k80, k80, e0 >> 16) & l8);
    r.r2 = (prmt(later.r0, later.r1, e1) | (prmt(k80, k80, e1) & l8)) & 0xffu;
    return r;
}

__device__ __forceinline__ std::uint32_t jcp_apply(const JcpMap& m, std::uint32_t x)
{
    return (x == 8u ? m.r2 : prmt(m.r0, m.r1, x)) & 0xffu;
}

struct JcpEntry
{
    unsigned long long m;
    std::uint32_t p, p_before; // pixel, pixel of the previous queue entry (0xffffffff: none in this row)
    float4 w[6];
    bool valid;
};

__global__ void __launch_bounds__(kJcpRowsThreads, LPL_JCP_ROWS_MINB) k_jcp_rows(Dev d, SegParams sp)
{
    extern __shared__ std::uint32_t plane[]; // npx / 16 words, then row_start[H + 1]
    __shared__ JcpMap s_tot[2][kJcpRowsThreads / 32];
    __shared__ std::uint32_t s_carry[2];
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::size_t po = static_cast<std::size_t>(f) * sp.npx;
    std::uint8_t* code = d.code + po;
    const std::uint32_t nwords = (sp.npx + 15) / 16;
    const std::uint32_t W = static_cast<std::uint32_t>(sp.W), H = static_cast<std::uint32_t>(sp.H);
    std::uint32_t* row_start = plane + nwords;
    for (std::uint32_t wi = threadIdx.x; wi < nwords; wi += blockDim.x)
    {
        std::uint32_t v = 0;
        const uint4 raw = *reinterpret_cast<const uint4*>(code + static_cast<std::size_t>(wi) * 16);
        const std::uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int b = 0; b < 16; ++b)
        {
            const std::uint32_t c = (r[b >> 2] >> ((b & 3) * 8)) & 0xfu;
            v |= (c & 3u) << (2 * b); // EMPTY 0, GROUND 1, OBSTACLE 2, QUEUED 3
        }
        plane[wi] = v;
    }
    const std::uint32_t nq = min(d.n_queue[f], d.qcap);
    const std::uint32_t* queue = d.queue + static_cast<std::size_t>(f) * d.qcap;
    const unsigned long long* mkv = d.mk + static_cast<std::size_t>(f) * d.qcap;
    const float* wn = d.wn + static_cast<std::size_t>(f) * 24 * d.qcap;
    const std::uint32_t* sref = d.stale_ref + static_cast<std::size_t>(f) * d.nborder_cap * 12;
    // first queue entry of every row (the queue is in raster order)
    for (std::uint32_t h = threadIdx.x; h <= H; h += blockDim.x)
    {
        const std::uint32_t target = h * W;
        std::uint32_t lo = 0, hi = nq;
        while (lo < hi)
        {
            const std::uint32_t mid = (lo + hi) >> 1;
            if (queue[mid] < target)
            {
                lo = mid + 1;
            }
            else
            {
                hi = mid;
            }
        }
        row_start[h] = lo;
    }
    __syncthreads();
    const std::uint32_t T = blockDim.x, lane = lane_id(), warp = threadIdx.x >> 5;
    // chunks of T consecutive queue entries of one row, in raster order
    auto skip_empty = [&](std::uint32_t& hh, std::uint32_t& bb) {
        while (hh < H && row_start[hh] + bb >= row_start[hh + 1])
        {
            ++hh;
            bb = 0;
        }
    };
    auto load = [&](std::uint32_t hh, std::uint32_t bb) {
        JcpEntry en;
        en.valid = false;
        en.m = 0;
        en.p = 0;
        en.p_before = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 6; ++i)
        {
            en.w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (hh < H)
        {
            const std::uint32_t ks = row_start[hh];
            const std::uint32_t e = ks + bb + threadIdx.x;
            if (e < row_start[hh + 1])
            {
                en.valid = true;
                en.m = mkv[e];
                en.p = queue[e];
                if (lane == 0 && e > ks)
                {
                    en.p_before = queue[e - 1];
                }
                const float4* wrow = reinterpret_cast<const float4*>(wn + static_cast<std::size_t>(e) * 24);
#pragma unroll
                for (int i = 0; i < 6; ++i)
                {
                    en.w[i] = wrow[i];
                }
            }
        }
        return en;
    };
    bool bad = false;
    std::uint32_t h = 0, base = 0, parity = 0, chunks = 0;
    skip_empty(h, base);
    JcpEntry cur = load(h, base);
    while (h < H)
    {
        std::uint32_t nh = h, nb = base + T;
        skip_empty(nh, nb);
        const JcpEntry nxt = load(nh, nb); // in flight while this chunk is voted on
        // ---- vote tables
        const int wv = static_cast<int>(cur.p - h * W);
        std::uint32_t entry = 0;
        if (cur.valid && ((cur.m >> 48) & 1ULL))
        {
            const unsigned long long m = cur.m;
            float wt[24];
#pragma unroll
            for (int i = 0; i < 6; ++i)
            {
                wt[4 * i] = cur.w[i].x;
                wt[4 * i + 1] = cur.w[i].y;
                wt[4 * i + 2] = cur.w[i].z;
                wt[4 * i + 3] = cur.w[i].w;
            }
            bool spec10 = false, spec11 = false;
            float g = 0.f, o = 0.f; // sums over slots 0..9
            float g10 = 0.f, o10 = 0.f, g11 = 0.f, o11 = 0.f;
#pragma unroll
            for (int s = 0; s < 12; ++s)
            {
                const int dh = s < 5 ? -2 : (s < 10 ? -1 : 0);
                const int dw = s < 5 ? s - 2 : (s < 10 ? s - 7 : s - 12);
                std::uint32_t mi = static_cast<std::uint32_t>(m >> (2 * s)) & 3u;
                if (mi == 3u)
                {
                    const int hh = static_cast<int>(h) + dh, ww = wv + dw;
                    if (hh >= 0 && ww >= 0 && ww < sp.W)
                    {
                        if (dh == 0)
                        {
                            // left neighbour in this row: resolved by the chain scan
                            spec10 = spec10 || s == 10;
                            spec11 = spec11 || s == 11;
                            mi = 0u;
                        }
                        else
                        {
                            mi = plane_get(plane, static_cast<std::uint32_t>(hh * sp.W + ww));
                            bad = bad || mi == 3u;
                        }
                    }
                    else
                    {
                        // inherited (stale) slot of a border pixel: explicit reference into an earlier row
                        const std::uint32_t brow = static_cast<std::uint32_t>(m >> 49);
                        mi = plane_get(plane, sref[static_cast<std::size_t>(brow - 1) * 12 + s]);
                        bad = bad || mi == 3u;
                    }
                }
                const float cg = mi == 1u ? wt[s] : 0.f, co = mi == 2u ? wt[s] : 0.f;
                if (s < 10)
                {
                    g += cg;
                    o += co;
                }
                else if (s == 10)
                {
                    g10 = cg;
                    o10 = co;
                }
                else
                {
                    g11 = cg;
                    o11 = co;
                }
            }
            // variants [a][b]: a = left-2 is of the class, b = left-1 is of the class
            float G[4], O[4];
#pragma unroll
            for (int v = 0; v < 4; ++v)
            {
                const bool a = (v & 2) != 0, b = (v & 1) != 0;
                G[v] = (g + (spec10 ? (a ? wt[10] : 0.f) : g10)) + (spec11 ? (b ? wt[11] : 0.f) : g11);
                O[v] = (o + (spec10 ? (a ? wt[10] : 0.f) : o10)) + (spec11 ? (b ? wt[11] : 0.f) : o11);
            }
#pragma unroll
            for (int s = 12; s < 24; ++s)
            {
                const std::uint32_t mi = static_cast<std::uint32_t>(m >> (2 * s)) & 3u;
                const float cg = mi == 1u ? wt[s] : 0.f, co = mi == 2u ? wt[s] : 0.f;
#pragma unroll
                for (int v = 0; v < 4; ++v)
                {
                    G[v] += cg;
                    O[v] += co;
                }
            }
            std::uint32_t table = 0;
#pragma unroll
            for (int c = 0; c < 9; ++c)
            {
                const int s10 = c / 3, s11 = c % 3;
                const int vg = (s10 == 1 ? 2 : 0) | (s11 == 1 ? 1 : 0);
                const int vo = (s10 == 2 ? 2 : 0) | (s11 == 2 ? 1 : 0);
                table |= (O[vo] > G[vg]) ? (1u << c) : 0u;
            }
            entry = table | (1u << 9) | (spec10 ? 1u << 10 : 0u) | (spec11 ? 1u << 11 : 0u);
        }
        // ---- the entry as a map on X = (outcome of entry q-2, outcome of entry q-1)
        std::uint32_t p_before = __shfl_up_sync(0xffffffffu, cur.p, 1);
        if (lane == 0)
        {
            p_before = cur.p_before;
        }
        JcpMap map = jcp_identity();
        if (cur.valid)
        {
            // slot 10 (pixel w - 2) is the previous entry when that is not the pixel w - 1
            const bool prev_is_w2 = p_before + 2u == cur.p;
            std::uint32_t img[9];
#pragma unroll
            for (int x = 0; x < 9; ++x)
            {
                const std::uint32_t a = x / 3, b = x % 3;
                std::uint32_t out = 0;
                if (entry & 0x200u)
                {
                    const std::uint32_t s11 = (entry & 0x800u) ? b : 0u;
                    const std::uint32_t s10 = (entry & 0x400u) ? (prev_is_w2 ? b : a) : 0u;
                    out = ((entry >> (s10 * 3u + s11)) & 1u) ? 2u : 1u;
                }
                img[x] = b * 3u + out;
            }
            map.r0 = img[0] | (img[1] << 8) | (img[2] << 16) | (img[3] << 24);
            map.r1 = img[4] | (img[5] << 8) | (img[6] << 16) | (img[7] << 24);
            map.r2 = img[8];
        }
        // ---- inclusive scan of the maps over the warp, totals chained across the warps
#pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            std::uint32_t n0, n1;
            jcp_pack(map, n0, n1);
            const std::uint32_t e0 = __shfl_up_sync(0xffffffffu, n0, off);
            const std::uint32_t e1 = __shfl_up_sync(0xffffffffu, n1, off);
            if (lane >= static_cast<std::uint32_t>(off))
            {
                map = jcp_compose(map, e0, e1);
            }
        }
        if (lane == 31u)
        {
            s_tot[parity][warp] = map;
        }
        __syncthreads();
        std::uint32_t x = (base == 0u) ? 0u : s_carry[parity ^ 1u]; // state at the chunk start
        for (std::uint32_t w2 = 0; w2 < warp; ++w2)
        {
            x = jcp_apply(s_tot[parity][w2], x);
        }
        x = jcp_apply(map, x); // state after this entry: (outcome before, own outcome)
        if (threadIdx.x == T - 1u)
        {
            s_carry[parity] = x;
        }
        if (cur.valid)
        {
            const std::uint32_t out = x % 3u;
            const std::uint32_t px = cur.p;
            atomicAnd(&plane[px >> 4], ~((3u ^ out) << ((px & 15u) * 2u)));
            code[px] = (out == 0u) ? PX_UNDECIDED : static_cast<std::uint8_t>(out);
        }
        if (nh != h)
        {
            __syncthreads(); // the plane is final for the next row
        }
        // (within a row the next chunk's votes only read rows above, and s_tot / s_carry alternate
        // between chunks: the first barrier of the next chunk orders their reuse)
        parity ^= 1u;
        ++chunks;
        cur = nxt;
        h = nh;
        base = nb;
    }
    if (bad)
    {
        atomicOr(&d.status[f], ST_JCP_STALL); // a rows-above dependency was not final: internal error
    }
    if (threadIdx.x == 0)
    {
        d.jcp_rounds[f] = chunks;
    }
}


// populateLabels (segmenter.cpp:640-669): only pixel winners receive a label; optional BGR image
__global__ void __launch_bounds__(256) k_seg_labels_out(Dev d, SegParams sp, int want_image)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t p = blockIdx.x * 256u + threadIdx.x;
    if (p >= static_cast<std::uint32_t>(sp.npx))
    {
        return;
    }
    const std::size_t po = static_cast<std::size_t>(f) * sp.npx;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint8_t cfull = d.code[po + p];
    const int idx = d.pxidx[po + p]; // issued together with the code byte
    const std::uint8_t c = cfull & 0xf;
    if (c == PX_GROUND || c == PX_OBSTACLE)
    {
        if (idx >= 0)
        {
            d.labels_out[o + idx] = c;
        }
    }
    if (want_image)
    {
        std::uint8_t b = 0, g = 0, r = 0;
        if (c == PX_GROUND)
        {
            g = 255;
        }
        else if (c == PX_OBSTACLE)
        {
            r = 255;
        }
        else if (c == PX_QUEUED || c == PX_UNDECIDED)
        {
            b = 255;
        }
        else if (cfull & PX_DILATED)
        {
            r = 255;
        }
        std::uint8_t* px = d.bgr + (po + p) * 3;
        px[0] = b;
        px[1] = g;
        px[2] = r;
    }
}

// ------------------------------------------------------------------------------------------
void launch_segment(Ctx* c, std::uint32_t nf, bool want_image)
{
    Dev& d = c->d;
    const SegParams& sp = c->seg;
    cudaStream_t s = c->stream;
    // per-batch resets
    cudaMemsetAsync(at_frame(d.cell_cnt, sp.ncell, d.f0), 0, sizeof(std::uint32_t) * sp.ncell * nf, s);
    cudaMemsetAsync(at_frame(d.ccnt, sp.ncell, d.f0), 0, sizeof(std::uint32_t) * sp.ncell * nf, s);
    cudaMemsetAsync(at_frame(d.key, sp.npx, d.f0), 0xff, sizeof(unsigned long long) * sp.npx * nf, s);
    cudaMemsetAsync(at_frame(d.labels_out, d.cap, d.f0), 0, static_cast<std::size_t>(d.cap) * nf, s);
    if (!c->counters_cleared)
    {
        cudaMemsetAsync(at_frame(d.n_cpts, 1, d.f0), 0, sizeof(std::uint32_t) * nf, s);
        cudaMemsetAsync(at_frame(d.n_v, 1, d.f0), 0, sizeof(std::uint32_t) * nf, s);
        cudaMemsetAsync(at_frame(d.n_border, 1, d.f0), 0, sizeof(std::uint32_t) * nf, s);
    }

    const dim3 gpts((d.cap + 255) / 256, nf);
    k_seg_bin<<<gpts, 256, 0, s>>>(d, sp);
    mark(c, "seg_bin");
    k_excl_scan<<<nf, 1024, 0, s>>>(d.cell_cnt, sp.ncell, d.cell_start, sp.ncell + 1,
                                    static_cast<std::uint32_t>(sp.ncell), nullptr, d.n_binned, d.f0);
    mark(c, "seg_cell_scan");
    k_seg_scatter<<<gpts, 256, 0, s>>>(d, sp);
    mark(c, "seg_scatter");
    k_seg_cell<<<dim3((sp.ncell + kCellWarps * kCellsPerWarp - 1) / (kCellWarps * kCellsPerWarp), nf), kCellWarps * 32, 0, s>>>(d, sp);
    mark(c, "seg_cell");
    k_seg_elev<<<dim3((sp.slices + 127) / 128, nf), 128, 0, s>>>(d, sp);
    mark(c, "seg_elev");
    k_seg_label<<<gpts, 256, 0, s>>>(d, sp);
    mark(c, "seg_label");
    k_ransac_draw<<<nf, kDrawThreads, 0, s>>>(d, sp);
    mark(c, "ransac_draw");
    k_ransac_plane<<<dim3(kRansacIters, nf), 64, 0, s>>>(d, sp);
    mark(c, "ransac_plane");
    k_ransac_count<<<dim3((d.cap + 256 * kRansacPer - 1) / (256 * kRansacPer), nf), 256, 0, s>>>(d, sp);
    mark(c, "ransac_count");
    k_ransac_best<<<nf, 32, 0, s>>>(d);
    mark(c, "ransac_best");
    k_seg_px<<<dim3((sp.npx + 255) / 256, nf), 256, 0, s>>>(d, sp);
    mark(c, "seg_px");
    const dim3 gdil((sp.W + kDilTw - 1) / kDilTw, (sp.H + kDilTh - 1) / kDilTh, nf);
    if (sp.W % 16 == 0)
    {
        if (!c->have_code_map)
        {
            c->have_code_map = make_code_map(c);
            if (!c->have_code_map)
            {
                std::snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled failed for the %d x %d pixel-code planes", sp.H, sp.W);
                c->launch_failed = true;
                return;
            }
        }
        k_seg_dilate_tma<<<gdil, 256, 0, s>>>(d, sp, c->code_map);
    }
    else
    {
        k_seg_dilate<<<gdil, 256, 0, s>>>(d, sp);
    }
    mark(c, "seg_dilate");
    launch_compact(c, "jcp_queue", nf, d.ptiles, nullptr, static_cast<std::uint32_t>(sp.npx), d.tile_cnt, d.n_queue,
                   QueuePred{d.code, static_cast<std::uint32_t>(sp.npx)},
                   QueueEmit{d.queue, d.status, d.qcap});
    k_jcp_pre<<<dim3(std::min<std::uint32_t>((d.qcap + 127) / 128, per_frame_ctas(128, nf, 1024)), nf), 128, 0, s>>>(d, sp);
    mark(c, "jcp_pre");
    // state plane (32 KB for 64 x 2048, 64 KB for 128-beam images): above the 48 KB default for the
    // larger images, opt in (up to 227 KB per CTA on sm_100a)
    const std::size_t plane_bytes = static_cast<std::size_t>((sp.npx + 15) / 16) * 4;
    const std::size_t rows_bytes = plane_bytes + sizeof(std::uint32_t) * (sp.H + 1);
    if (cudaFuncSetAttribute(k_jcp_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rows_bytes)) != cudaSuccess)
    {
        cudaGetLastError();
        std::snprintf(c->err, sizeof(c->err), "the %d x %d range image needs %zu bytes of shared memory per CTA for the JCP sweep", sp.H,
                      sp.W, rows_bytes);
        c->launch_failed = true;
        return;
    }
    k_jcp_rows<<<nf, kJcpRowsThreads, rows_bytes, s>>>(d, sp);
    mark(c, "jcp_rows");
    k_seg_labels_out<<<dim3((sp.npx + 255) / 256, nf), 256, 0, s>>>(d, sp, want_image ? 1 : 0);
    mark(c, "seg_labels_out");
}
} // namespace lpl
