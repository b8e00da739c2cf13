// Shared declarations for the sm_100a kernels of the LiDAR hot path.
//
// Layout in HBM: every per-point array is frame-major with a fixed stride of `cap` elements
// (ctx->cap, a multiple of kTile), so a batch of B frames is one launch with blockIdx.y = frame
// and per-frame point counts living in device memory (no host sync between stages).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "libm_exact.cuh"

namespace lpl
{
constexpr int kTile = 2048;         // items per compaction / scan tile
constexpr int kTileThreads = 256;   // threads per tile block (8 items per thread)
constexpr int kItems = kTile / kTileThreads;
constexpr int kDrorGrid = 256;      // DROR grids: kDrorGrid x kDrorGrid cells, two levels (1 m over +-128 m, 0.125 m over +-16 m)
constexpr int kDrorLevelCells = kDrorGrid * kDrorGrid;
constexpr int kDrorCells = 2 * kDrorLevelCells;
constexpr int kRansacIters = 60;    // segmenter.cpp:324
constexpr int kRansacBins = 4;      // segmenter.cpp:323
constexpr int kMtRaws = 8192;       // pre-generated std::mt19937{42} outputs
constexpr int kRawRecord = 32;      // bytes reserved per point for raw records (pcl::PointXYZIR = 32)

// pixel codes of the range image
enum : std::uint8_t
{
    PX_EMPTY = 0,
    PX_GROUND = 1,
    PX_OBSTACLE = 2,
    PX_QUEUED = 3,    // CV_INTERSECTION: ground pixel under the dilated obstacle mask
    PX_UNDECIDED = 4, // queued pixel whose neighbourhood carried no weight
    PX_DILATED = 0x10 // flag: red channel set by the 5x5 dilation (only visible on empty pixels)
};

struct SegParams
{
    // image / grid geometry
    int H, W, npx;
    int rings, slices, ncell;
    int use_ring; // 1: height index = ring field, 0: from elevation angle
    // derived float constants (computed on the host with the reference's float expressions)
    float radial_spacing, min_dist, max_dist;
    float slice_res, el_down, rad_per_px;
    float z_lo, z_hi;
    float thr, thr2, delta, e0;
    float cos_max, p1z;
    float kthr_sqr, amp;
    float wscale; // (W - 1) as float
    int jcp_emulate_stale; // 1 = bit-compatible with the reference's stale out-of-image slots
    int idx_bits;          // range-image key: bits of the point index (ceil log2 of the frame capacity)
    int nb;                // radial bins that feed RANSAC: min(rings, kRansacBins)
};

struct DrorParams
{
    double scaling;   // pow((double)radius_multiplier, 2.0)
    float min_r_sqr;  // min_search_radius^2
    std::uint32_t min_neighbours;
};

struct ClusterParams
{
    float range_res, az_res, el_res;
    std::uint32_t min_cluster_size;
};

// oriented bounding box of one hull (binary layout = lpl_bbox of the C ABI)
struct ObbBox
{
    double c[8]; // four corners (x, y)
    float area;
    float angle;
    std::int32_t valid;
    std::int32_t pad;
};

// packed result download (ingest.cu: launch_pack_results; C ABI: lpl_pipeline_download_packed). Plane order =
// the LPL_PLANE_* bits of include/lpl_b200.h.
constexpr int kPackPlanes = 10;
struct PackHeader
{
    unsigned long long offset[kPackPlanes]; // byte offset of every selected plane inside the payload, ~0 = not selected
    unsigned long long total;               // payload bytes
    std::uint32_t fits;                     // 0: the payload would not fit the staging area (nothing was packed)
    std::uint32_t pad;
};

// bytes per element of plane p: labels_u8, noise, ring, obstacle_index, cluster_labels, hull_offsets, hull_indices,
// hull_xy, zminmax, boxes
__host__ __device__ inline std::uint32_t pack_elem(int p)
{
    return p < 2 ? 1u : p == 2 ? 2u : p < 7 ? 4u : p < 9 ? 8u : 80u;
}

// which per-frame count sizes plane p: 0 = n, 1 = n_o, 2 = K + 1, 3 = K, 4 = Hv
__host__ __device__ inline int pack_count_of(int p)
{
    return p < 3 ? 0 : p < 5 ? 1 : p == 5 ? 2 : p < 8 ? 4 : 3;
}

// staging layout: [PackHeader][counts: 5 x nf u32][pad to 8][frame_off: 5 x nf u64][pad to 16][payload]
inline std::size_t pack_frame_off_start(std::uint32_t nf)
{
    return (sizeof(PackHeader) + static_cast<std::size_t>(5) * nf * sizeof(std::uint32_t) + 7u) & ~static_cast<std::size_t>(7);
}

inline std::size_t pack_payload_start(std::uint32_t nf)
{
    return (pack_frame_off_start(nf) + static_cast<std::size_t>(5) * nf * sizeof(unsigned long long) + 15u) & ~static_cast<std::size_t>(15);
}

// All device buffers of a context. Pointers are to the start of frame 0; frame f lives at
// ptr + f * stride (stride noted per field).
constexpr int kEdgePitch = 16; // words per voxel row of Dev::edges (13 forward neighbours + padding): 64 bytes, as a row of Dev::octa

struct Dev
{
    std::uint32_t cap;    // points per frame (multiple of kTile)
    std::uint32_t tiles;  // cap / kTile
    std::uint32_t B;      // frames per batch
    std::uint32_t qcap;   // JCP queue capacity per frame
    std::uint32_t hcap;   // voxel hash slots per frame (power of two)
    std::uint32_t ptiles; // pixel tiles per frame = ceil(npx / kTile)
    std::uint32_t f0;     // first frame of this launch: a kernel's frame is blockIdx + f0 (sub-batches of one run on
                          // concurrent streams, see capi.cu: enqueue_stages)

    // ---- input cloud
    float4* pts_in;           // [B][cap]  x,y,z,(unused)
    std::uint32_t* n_in;      // [B]
    std::uint16_t* ring;      // [B][cap]
    std::uint32_t* wrap_cnt;  // [B][cap / 32] ring wraps per warp of 32 points (fused ring + DROR scan-line pass)
    // ---- DROR
    std::uint8_t* noise;      // [B][cap]  0 valid, 1 noise
    std::uint32_t* grid_cnt;  // [B][kDrorCells]  (self-cleaning)
    std::uint32_t* grid_start;// [B][kDrorCells+1]
    std::uint32_t* grid_mask; // [B][kDrorCells/32] cells inside the search box of an unresolved query
    float4* grid_pts;         // [B][cap]  points in cell order
    std::uint32_t* unres;     // [B][cap]  unresolved point indices after the scan-line pass
    std::uint32_t* n_unres;   // [B]
    // ---- segmentation works on the input cloud in place (noise-masked); n_v counts the valid points
    std::uint32_t* n_v;       // [B]
    // ---- segmentation scratch
    std::int32_t* cell;       // [B][cap]  polar cell or -1
    std::uint32_t* px;        // [B][cap]  pixel index
    std::uint32_t* slot;      // [B][cap]  arrival slot inside the cell
    std::uint32_t* cell_cnt;  // [B][ncell]
    std::uint32_t* cell_start;// [B][ncell+1]
    std::uint32_t* n_binned;  // [B]       points that fell into the polar grid
    uint2* zo;                // [B][cap]  cell-major (unordered inside a cell): (point index, z bits)
    float* zsort;             // [B][cap]  scratch for oversized cells
    std::uint32_t* ccnt;      // [B][ncell] RANSAC candidates per cell, then their exclusive prefix in (slice, bin) order
    float* cell_zmin;         // [B][ncell]
    float* elev;              // [B][ncell]
    std::uint8_t* lab;        // [B][cap]  recorded verdicts of the two-pass compactions (ring wrap flags, cluster representatives, hull filter)
    std::uint32_t* n_cand;    // [B]
    float4* cpts;             // [B][cap]  dense unordered copy of the candidate points (inlier count)
    std::uint32_t* n_cpts;    // [B]
    std::uint32_t* pairs;     // [B][kRansacIters][2] drawn candidate ranks
    float4* planes;           // [B][kRansacIters] (nx, ny, nz, d); nz = NaN marks a skipped draw
    std::uint32_t* inliers;   // [B][kRansacIters]
    float4* best_plane;       // [B]  (a, b, c, d); w component of [B + f] unused
    std::uint32_t* best_cnt;  // [B]
    unsigned long long* key;  // [B][npx]  (depth_sqr bits << 33) | (azimuth slice << idx_bits) | point index
    std::int32_t* pxidx;      // [B][npx]  the range image: index of the pixel's winner in the input cloud (-1 = none)
    std::uint8_t* code;       // [B][npx]  PX_* before / after JCP
    std::uint32_t* queue;     // [B][qcap] queued pixels in raster order
    std::uint32_t* n_queue;   // [B]
    float* wn;                // [B][qcap][24] normalised weights, entry-major (96 B per queued pixel)
    unsigned long long* mk;   // [B][qcap] 2-bit mask source per slot + flags
    std::uint32_t* stale_ref; // [B][nborder][12] explicit pixel refs for inherited slots
    std::uint32_t* n_border;  // [B]
    std::uint32_t nborder_cap;
    std::uint32_t* jcp_rounds;// [B] chunks of queue entries swept by k_jcp_rows (diagnostic)
    std::uint8_t* labels_out; // [B][cap]  Label (0/1/2) per *input* point
    std::uint8_t* bgr;        // [B][npx*3]
    // ---- obstacle cloud / clustering
    float4* pts_o;            // [B][cap]
    std::uint32_t* idx_o;     // [B][cap]  index in the input cloud
    std::uint32_t* n_o;       // [B]
    float4* sph;              // [B][cap]  range, azimuth, elevation
    std::uint32_t* sph_max;   // [B][4]    float bits of max range / azimuth / elevation
    std::int32_t* hkey;       // [B][hcap] voxel flat index or -1
    std::uint32_t* hparent;   // [B][hcap] union-find parent by voxel id (global path only)
    std::uint32_t* hmin;      // [B][hcap] min point index (per voxel, then per root)
    std::uint32_t* hcount;    // [B][hcap] points per root
    std::int32_t* hlabel;     // [B][hcap] final label per root
    std::uint32_t* hroot;     // [B][hcap] root slot per occupied voxel
    std::uint32_t* hvid;      // [B][hcap] position of the slot in the frame's voxel list
    std::uint32_t* edges;     // [B][cap][16] voxel ids of the 13 occupied forward neighbours (or ~0), rows padded to kEdgePitch
    std::uint32_t* vslot;     // [B][cap]  voxel slot per point
    std::uint32_t* vlist;     // [B][cap]  slots of the occupied voxels (unordered)
    std::uint32_t* n_vox;     // [B]
    std::int32_t* clabel;     // [B][cap]  cluster label per obstacle point
    std::uint32_t* n_clusters;// [B]
    // ---- hulls
    std::uint32_t* ccount;    // [B][cap]  points per cluster (indexed by label)
    std::uint32_t* cstart;    // [B][cap+1]
    uint4* hsA;               // [B][cap]  sort elements (label, ord x, ord y, index), ping
    uint4* hsB;               // [B][cap]  pong
    std::uint32_t* hstack;    // [B][2*cap] per-cluster hull vertices (at segment offset + cluster id)
    std::uint32_t* hfin;      // [B][cap]  clusters of several chunks: work list of the join pass
    std::uint32_t* hcnt;      // [B][cap]  hull vertex count per cluster
    std::uint32_t* hull_off;  // [B][cap+1]
    std::uint32_t* hull_idx;  // [B][cap]  obstacle-cloud index per hull vertex
    float2* hull_xy;          // [B][cap]
    float2* zminmax;          // [B][cap]  per cluster
    std::uint32_t* n_hull;    // [B]       hull vertices of the frame
    std::uint32_t* zmin_u;    // [B][cap]  per cluster: order-preserving bits of min z
    std::uint32_t* zmax_u;    // [B][cap]  per cluster: order-preserving bits of max z
    std::uint32_t* zzero;     // [B][cap]  per cluster: (first point in cloud order with z == +-0) << 1 | its sign bit, ~0 = none
    unsigned long long* ext;  // [B][cap][kExtDirs] per cluster: extreme points (ordered value bits << 32 | point index) for
                              //             min x, min x+y, min y, max x-y, max x, max x+y, max y, min x-y (CCW order)
    float2* octa;             // [B][cap][kExtDirs] per cluster: the polygon of those points, or NaN when unusable
    std::uint32_t* hseg_cnt;  // [B][cap]  per cluster: points that survive the octagon filter
    std::uint32_t* n_h;       // [B]       survivors per frame (input size of the hull sort)
    std::uint32_t* hull_next; // [B]       next chunk to hand out in k_hull_chunks
    std::uint32_t* hwk_off;   // [B][cap+1] per cluster: first chunk (work item) of the smem hull pass; [K] = chunks of the frame
    std::uint32_t* hck_cnt;   // [B][2*cap] per chunk: points that survived its thinning (multi-chunk clusters)
    std::uint32_t* n_work;    // [B]       chunks of the frame
    std::uint32_t* n_multi;   // [B]       clusters of several chunks (listed in hfin) for the join pass
    ObbBox* boxes;            // [B][cap]  oriented bounding box per cluster (LPL_STAGE_BOXES)
    unsigned char* raw;       // [B][cap * kRawRecord] raw PointCloud2 records before the device unpack
    unsigned char* raw_desc;  // [B] record layouts (Cloud2Desc)
    // ---- generic
    std::uint32_t* tile_cnt;  // [B][max(tiles, ptiles)]
    std::uint32_t* status;    // [B]  error bits raised by kernels
    const std::uint32_t* mt_raw; // [kMtRaws]
    // the per-frame counters every run starts from zero (status, n_v, n_o, n_clusters, n_hull, n_unres, n_cpts, n_border,
    // n_vox, sph_max) lie next to each other in the slab: lpl_pipeline_run clears them with ONE memset
    unsigned char* ctr_begin;
    std::size_t ctr_bytes;
};

enum : std::uint32_t
{
    ST_QUEUE_OVERFLOW = 1u,  // more queued JCP pixels than qcap
    ST_RNG_EXHAUSTED = 2u,   // RANSAC needed more than kMtRaws generator outputs
    ST_HASH_FULL = 4u,       // voxel hash table full
    ST_BORDER_OVERFLOW = 8u, // more queued border pixels than reserved
    ST_JCP_STALL = 16u,      // JCP sweep met a rows-above dependency that was not final (internal error)
};

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ std::uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ std::uint32_t warp_sum(std::uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

__device__ __forceinline__ std::uint32_t warp_incl_scan(std::uint32_t v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const std::uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= static_cast<std::uint32_t>(o))
        {
            v += t;
        }
    }
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `sh` needs 32 words.
__device__ __forceinline__ std::uint32_t block_sum(std::uint32_t v, std::uint32_t* sh)
{
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0)
    {
        sh[threadIdx.x >> 5] = v;
    }
    __syncthreads();
    const std::uint32_t nw = (blockDim.x + 31) >> 5;
    std::uint32_t r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0;
    if (threadIdx.x < 32)
    {
        r = warp_sum(r);
        if (threadIdx.x == 0)
        {
            sh[0] = r;
        }
    }
    __syncthreads();
    return sh[0];
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum. `sh` needs 33 words.
__device__ __forceinline__ std::uint32_t block_excl_scan(std::uint32_t v, std::uint32_t* sh,
                                                         std::uint32_t* total)
{
    const std::uint32_t incl = warp_incl_scan(v);
    __syncthreads();
    if (lane_id() == 31)
    {
        sh[threadIdx.x >> 5] = incl;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        const std::uint32_t nw = (blockDim.x + 31) >> 5;
        const std::uint32_t w = (threadIdx.x < nw) ? sh[threadIdx.x] : 0;
        const std::uint32_t ws = warp_incl_scan(w);
        sh[threadIdx.x] = ws - w;
        if (threadIdx.x == 31)
        {
            sh[32] = ws;
        }
    }
    __syncthreads();
    *total = sh[32];
    return sh[threadIdx.x >> 5] + incl - v;
}

// Stable ranks for a tile processed as kItems striped rounds of kTileThreads threads:
// item (j, t) = tile_base + j * kTileThreads + t. Returns per-item exclusive rank among the
// flagged items of the tile, in item order. `sh` needs kItems * 8 + 1 words.
__device__ __forceinline__ void tile_ranks(const bool (&flag)[kItems], std::uint32_t (&rank)[kItems],
                                           std::uint32_t* total, std::uint32_t* sh)
{
    constexpr int kWarps = kTileThreads / 32;
    const std::uint32_t w = threadIdx.x >> 5;
    std::uint32_t below[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j)
    {
        const std::uint32_t b = __ballot_sync(0xffffffffu, flag[j]);
        below[j] = __popc(b & ((1u << lane_id()) - 1u));
        if (lane_id() == 0)
        {
            sh[j * kWarps + w] = __popc(b);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        // kItems * kWarps = 64 counters -> two per lane
        const std::uint32_t a = sh[2 * threadIdx.x];
        const std::uint32_t b = sh[2 * threadIdx.x + 1];
        const std::uint32_t s = warp_incl_scan(a + b);
        sh[2 * threadIdx.x] = s - a - b;
        sh[2 * threadIdx.x + 1] = s - b;
        if (threadIdx.x == 31)
        {
            sh[kItems * kWarps] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kItems; ++j)
    {
        rank[j] = sh[j * kWarps + w] + below[j];
    }
    *total = sh[kItems * kWarps];
    __syncthreads();
}

static_assert(kItems * (kTileThreads / 32) == 64, "tile_ranks assumes 64 (round, warp) counters");

// merge passes the frame-wide sort of hull.cu needs for n elements (tiles of kTile): ceil(log2(tiles))
__host__ __device__ __forceinline__ std::uint32_t sort_passes(std::uint32_t n)
{
    const std::uint32_t nt = (n + kTile - 1) / kTile;
    std::uint32_t p = 0;
    while ((1u << p) < nt)
    {
        ++p;
    }
    return p;
}

// order-preserving float <-> uint32 map (-0.0 folded into +0.0)
__device__ __forceinline__ std::uint32_t ord_f32(float v)
{
    v = v + 0.0f;
    const std::uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float unord_f32(std::uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Per-cluster statistics gathered while labels are written: the z extent
// (src/processor/src/processor.cpp:648-655) and the extreme points in kExtDirs directions (hull.cu
// builds an inscribed polygon from them and drops every point strictly inside it before the hull
// sort). Slot k holds the point that maximises the dot product with direction k (counter-clockwise
// from 180 degrees) as (ordered value bits << 32 | point index); 0 = no point yet.
// Scan-ordered clouds are monotone along a ring, so per-point atomics would hammer one address
// per cluster: the lanes of a warp that share a label are reduced first (redux.sync over the
// match group), and the group leader only issues an atomic when an L2 read says it improves.
// Must be called by all 32 lanes; label < 0 = no contribution.
#ifndef LPL_EXT_DIRS
#define LPL_EXT_DIRS 8 // measured on KITTI: 16 directions halve the hull-sort input (20 % -> 10 % of the obstacle points, hull stage -0.1 ms) but double k_clu_labels (+0.28 ms)
#endif
constexpr int kExtDirs = LPL_EXT_DIRS;
static_assert(kExtDirs == 8 || kExtDirs == 16, "8 (octagon) or 16 directions");

__device__ __forceinline__ float ext_dot(int k, float x, float y)
{
    // integer direction vectors: the products are exact, one rounding in the sum
    constexpr int dx16[16] = {-1, -2, -1, -1, 0, 1, 1, 2, 1, 2, 1, 1, 0, -1, -1, -2};
    constexpr int dy16[16] = {0, -1, -1, -2, -1, -2, -1, -1, 0, 1, 1, 2, 1, 2, 1, 1};
    constexpr int dx8[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    constexpr int dy8[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    const float dx = static_cast<float>(kExtDirs == 16 ? dx16[k & 15] : dx8[k & 7]);
    const float dy = static_cast<float>(kExtDirs == 16 ? dy16[k & 15] : dy8[k & 7]);
    return dx * x + dy * y;
}

__device__ __forceinline__ void ext_init(unsigned long long* e)
{
#pragma unroll
    for (int k = 0; k < kExtDirs; ++k)
    {
        e[k] = 0ULL;
    }
}

// `with_extremes` (warp uniform): whether this warp also contributes to the extreme points. The octagon
// only has to be spanned by points of the cluster, so a subset of the points (every kExtWarpStride-th
// warp) gives a valid - for large clusters practically identical - filter at a fraction of the cost;
// the z extent always takes every point.
#ifndef LPL_EXT_WARP_STRIDE
#define LPL_EXT_WARP_STRIDE 1 // measured on KITTI: 2 / 4 / 8 let 25 / 32 / 41 % (instead of 20 %) of the obstacle points into the hull sort and cost more there than they save here
#endif
constexpr std::uint32_t kExtWarpStride = LPL_EXT_WARP_STRIDE;

__device__ __forceinline__ void accumulate_cluster_stats(unsigned long long* ext, std::uint32_t* zmin_u,
                                                         std::uint32_t* zmax_u, std::uint32_t* zzero, std::int32_t label,
                                                         float x, float y, float z, std::uint32_t idx, bool with_extremes)
{
    const std::uint32_t peers = __match_any_sync(0xffffffffu, label);
    if (label < 0)
    {
        return;
    }
    // The reference keeps the FIRST point in cloud order among equal extremes (strict compares,
    // processor.cpp:648-655), which is only observable for z == +-0: the ordered keys below fold -0.0 into
    // +0.0, so the sign of a zero extent is taken from the first zero-height point of the cluster (rare).
    if (z == 0.0f)
    {
        atomicMin(&zzero[label], (idx << 1) | (__float_as_uint(z) >> 31));
    }
    const bool lead = static_cast<int>(lane_id()) == __ffs(peers) - 1;
    // the leader fetches the cluster's running values up front, all loads in flight together: their
    // L2 latency overlaps the reductions below (loads placed behind the first atomic could not be
    // hoisted by the compiler and would cost one round trip each)
    unsigned long long* e = ext + static_cast<std::size_t>(label) * kExtDirs;
    ulonglong2 cur2[kExtDirs / 2];
    std::uint32_t cur_zmin = 0, cur_zmax = 0;
    if (lead)
    {
        if (with_extremes)
        {
            const ulonglong2* e2 = reinterpret_cast<const ulonglong2*>(e);
#pragma unroll
            for (int k = 0; k < kExtDirs / 2; ++k)
            {
                cur2[k] = __ldcg(e2 + k);
            }
        }
        cur_zmin = __ldcg(zmin_u + label);
        cur_zmax = __ldcg(zmax_u + label);
    }
    const std::uint32_t zk = ord_f32(z);
    const std::uint32_t zlo = __reduce_min_sync(peers, zk);
    const std::uint32_t zhi = __reduce_max_sync(peers, zk);
    if (lead)
    {
        if (zlo < cur_zmin)
        {
            atomicMin(&zmin_u[label], zlo);
        }
        if (zhi > cur_zmax)
        {
            atomicMax(&zmax_u[label], zhi);
        }
    }
    if (!with_extremes)
    {
        return;
    }
#pragma unroll
    for (int k = 0; k < kExtDirs; ++k)
    {
        const std::uint32_t vk = ord_f32(ext_dot(k, x, y));
        const std::uint32_t b = __reduce_max_sync(peers, vk);
        const std::uint32_t bi = __reduce_max_sync(peers, vk == b ? idx : 0u);
        if (lead)
        {
            const unsigned long long key = (static_cast<unsigned long long>(b) << 32) | bi;
            const unsigned long long cur = (k & 1) ? cur2[k >> 1].y : cur2[k >> 1].x;
            if (key > cur)
            {
                atomicMax(e + k, key);
            }
        }
    }
}

// ---------------------------------------------------------------- generic tile compaction
// Stable, order-preserving selection of the items i in [0, n_f) of every frame f for which
// pred(f, i) holds. Pass 1 counts per tile, pass 2 adds the counts of the preceding tiles and
// scatters: emit(f, i, pos). n_out[f] (nullable) receives the number selected.
template <class Pred>
__global__ void __launch_bounds__(kTileThreads)
    k_compact_count(Pred pred, const std::uint32_t* __restrict__ n_arr, std::uint32_t n_const,
                    std::uint32_t* __restrict__ tile_cnt, std::uint32_t tiles_per_frame, std::uint32_t f0)
{
    __shared__ std::uint32_t sh[33];
    const std::uint32_t f = blockIdx.y + f0;
    const std::uint32_t n = n_arr ? n_arr[f] : n_const;
    const std::uint32_t base = blockIdx.x * kTile;
    std::uint32_t c = 0;
    if (base < n)
    {
#pragma unroll
        for (int j = 0; j < kItems; ++j)
        {
            const std::uint32_t i = base + j * kTileThreads + threadIdx.x;
            c += (i < n && pred(f, i)) ? 1u : 0u;
        }
    }
    const std::uint32_t s = block_sum(c, sh);
    if (threadIdx.x == 0)
    {
        tile_cnt[f * tiles_per_frame + blockIdx.x] = s;
    }
}

template <class Pred, class Emit>
__global__ void __launch_bounds__(kTileThreads)
    k_compact_scatter(Pred pred, Emit emit, const std::uint32_t* __restrict__ n_arr,
                      std::uint32_t n_const, const std::uint32_t* __restrict__ tile_cnt,
                      std::uint32_t tiles_per_frame, std::uint32_t* __restrict__ n_out, std::uint32_t f0)
{
    __shared__ std::uint32_t sh[kItems * (kTileThreads / 32) + 1];
    __shared__ std::uint32_t sh2[33];
    const std::uint32_t f = blockIdx.y + f0;
    const std::uint32_t n = n_arr ? n_arr[f] : n_const;
    const std::uint32_t base = blockIdx.x * kTile;
    const std::uint32_t* tc = tile_cnt + f * tiles_per_frame;
    // sum of the preceding tiles (and, in tile 0, of all tiles for n_out)
    std::uint32_t before = 0, all = 0;
    for (std::uint32_t t = threadIdx.x; t < tiles_per_frame; t += kTileThreads)
    {
        const std::uint32_t v = tc[t];
        all += v;
        before += (t < blockIdx.x) ? v : 0u;
    }
    before = block_sum(before, sh2);
    if (blockIdx.x == 0 && n_out != nullptr)
    {
        all = block_sum(all, sh2);
        if (threadIdx.x == 0)
        {
            n_out[f] = all;
        }
    }
    if (base >= n)
    {
        return;
    }
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
        if (flag[j])
        {
            emit(f, base + j * kTileThreads + threadIdx.x, before + rank[j]);
        }
    }
}

// Exclusive scan of `len` counters per frame by one block of 1024 threads:
// out[f][0..len] (len + 1 entries, out[len] = total).
// `len_arr` (nullable) gives a per-frame length <= len; strides are in elements; total_out
// (nullable) receives the per-frame total.
__global__ void k_excl_scan(const std::uint32_t* __restrict__ in, std::uint32_t in_stride,
                            std::uint32_t* __restrict__ out, std::uint32_t out_stride,
                            std::uint32_t len, const std::uint32_t* __restrict__ len_arr,
                            std::uint32_t* __restrict__ total_out, std::uint32_t f0 = 0);

// start of frame f0 in a frame-major array of `stride` elements per frame (memsets of a sub-batch)
template <typename T>
inline T* at_frame(T* p, std::size_t stride, std::uint32_t f0)
{
    return p + stride * f0;
}

// host-side launchers implemented per stage
struct Ctx;
void launch_ring(Ctx* c, std::uint32_t nf);
void launch_dror(Ctx* c, std::uint32_t nf, bool with_ring);
void launch_segment(Ctx* c, std::uint32_t nf, bool want_image);
void launch_take_obstacles(Ctx* c, std::uint32_t nf);
void launch_cluster(Ctx* c, std::uint32_t nf);
void launch_hulls(Ctx* c, std::uint32_t nf);
void launch_boxes(Ctx* c, std::uint32_t nf, int method);
void launch_spread_packed(Ctx* c, std::uint32_t nf, const void* packed, const std::uint32_t* start, bool xyz12);
void launch_hull_sort(Ctx* c, std::uint32_t nf);
void launch_split_clouds(Ctx* c, std::uint32_t nf, unsigned char* out, std::size_t frame_records, const std::uint8_t* colours,
                         std::uint32_t colour_stride, std::uint32_t* cnt3, std::uint32_t* totals);
void launch_marker_lines(Ctx* c, std::uint32_t nf, std::uint32_t* mcount, std::uint32_t* moff, std::uint32_t* mtotal, double* out,
                         std::size_t frame_stride);
int launch_knn(Ctx* c, const float4* pts, std::uint32_t n, const float4* queries, std::uint32_t m, std::uint32_t k,
               const float* radius_sqr, float radius_all, float* best_d, std::uint32_t* best_i, std::uint32_t* count);
void launch_radius(Ctx* c, const float4* pts, std::uint32_t n, const float4* queries, std::uint32_t m, const float* radius_sqr,
                   float radius_all, std::uint32_t cap, float* out_d, std::uint32_t* out_i, std::uint32_t* count);
void launch_vehicle_match(Ctx* c, const double2* xy, const std::uint32_t* off, std::uint32_t K, const double2* zmm,
                          const std::uint32_t* sizes, const ObbBox* boxes, std::int32_t* cls, double* area_out);
void launch_hull_f64(Ctx* c, const double2* xy, std::uint32_t n, std::uint32_t* order, std::uint32_t* st, std::uint32_t* out_idx,
                     std::uint32_t* out_cnt);
void launch_pack_results(Ctx* c, std::uint32_t nf, std::uint32_t planes, unsigned char* staging, std::size_t staging_bytes);
void launch_unpack_cloud2(Ctx* c, std::uint32_t nf, const unsigned char* raw, std::size_t raw_stride, const void* desc);
void launch_boxes_hulls(Ctx* c, const double2* xy, const std::uint32_t* off, std::uint32_t K, int method, ObbBox* out);

struct Ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    Dev d{};
    SegParams seg{};
    DrorParams dror{};
    ClusterParams clu{};
    void* slab = nullptr;       // one device allocation backing every Dev pointer
    std::size_t slab_bytes = 0;
    char err[512] = {0};
    // pinned host staging for the C-ABI calls that take host pointers
    void* h_stage = nullptr;
    std::size_t h_stage_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long launches = 0; // kernels launched since the last reset
    bool hash_clean = false;         // the voxel hash planes are in their cleared state (cluster.cu)
    alignas(64) CUtensorMap code_map{}; // TMA descriptor of the pixel-code planes (segment.cu: k_seg_dilate_tma)
    bool have_code_map = false;
    bool launch_failed = false;      // a launcher could not set up its kernel (message in err)
    bool counters_cleared = false;   // inside lpl_pipeline_run: the launchers' own counter memsets are already done
    // per-kernel CUDA-event profile of the last lpl_pipeline_run (lpl_profile_*)
    static constexpr int kProfMax = 96;
    bool prof_on = false;
    int prof_n = 0;                       // marks recorded in the current run
    cudaEvent_t prof_ev[kProfMax + 1] = {};
    const char* prof_name[kProfMax] = {};
};

// Called after every kernel launch: counts it and, when profiling, drops an event behind it so
// the time between consecutive marks is that kernel's duration on the context stream
// (preceding memsets are attributed to the kernel that follows them).
// CTAs per frame of the kernels that stride a fixed number of CTAs over a frame's work list. The
// defaults are tuned on 154-frame batches; a small batch (single-frame latency mode, a few very large
// clouds) gets proportionally more CTAs per frame so that the whole GPU stays occupied.
inline unsigned per_frame_ctas(unsigned tuned, unsigned nf, unsigned cap)
{
    const unsigned total = tuned * 154u;
    const unsigned scaled = (total + nf - 1u) / (nf == 0u ? 1u : nf);
    return scaled < tuned ? tuned : (scaled > cap ? cap : scaled);
}

inline void mark(Ctx* c, const char* name)
{
    c->launches += 1;
    if (c->prof_on && c->prof_n < Ctx::kProfMax)
    {
        c->prof_name[c->prof_n] = name;
        cudaEventRecord(c->prof_ev[c->prof_n + 1], c->stream);
        c->prof_n += 1;
    }
}


// a predicate that is expensive to evaluate (fp64 polygon tests, dependent gathers) is evaluated once:
// the counting pass stores its verdicts as bytes, the scatter pass reads them back
template <class Pred>
struct RecordingPred
{
    Pred pred;
    std::uint8_t* flags;
    std::uint32_t cap;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        const bool r = pred(f, i);
        flags[static_cast<std::size_t>(f) * cap + i] = r ? 1 : 0;
        return r;
    }
};

struct RecordedPred
{
    const std::uint8_t* flags;
    std::uint32_t cap;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        return flags[static_cast<std::size_t>(f) * cap + i] != 0;
    }
};

template <class Pred, class Emit>
inline void launch_compact_recorded(Ctx* c, const char* name, std::uint32_t B, std::uint32_t tiles_per_frame,
                                    const std::uint32_t* n_arr, std::uint32_t* tile_cnt, std::uint32_t* n_out,
                                    std::uint8_t* flags, Pred pred, Emit emit)
{
    const dim3 grid(tiles_per_frame, B);
    k_compact_count<<<grid, kTileThreads, 0, c->stream>>>(RecordingPred<Pred>{pred, flags, c->d.cap}, n_arr, 0u, tile_cnt,
                                                          tiles_per_frame, c->d.f0);
    mark(c, name);
    k_compact_scatter<<<grid, kTileThreads, 0, c->stream>>>(RecordedPred{flags, c->d.cap}, emit, n_arr, 0u, tile_cnt,
                                                            tiles_per_frame, n_out, c->d.f0);
    mark(c, name);
}

template <class Pred, class Emit>
inline void launch_compact(Ctx* c, const char* name, std::uint32_t B, std::uint32_t tiles_per_frame,
                           const std::uint32_t* n_arr, std::uint32_t n_const,
                           std::uint32_t* tile_cnt, std::uint32_t* n_out, Pred pred, Emit emit)
{
    const dim3 grid(tiles_per_frame, B);
    k_compact_count<<<grid, kTileThreads, 0, c->stream>>>(pred, n_arr, n_const, tile_cnt, tiles_per_frame, c->d.f0);
    mark(c, name);
    k_compact_scatter<<<grid, kTileThreads, 0, c->stream>>>(pred, emit, n_arr, n_const, tile_cnt,
                                                            tiles_per_frame, n_out, c->d.f0);
    mark(c, name);
}

#define LPL_CUDA_OK(call)                                                                          \
    do                                                                                             \
    {                                                                                              \
        const cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                                     \
        {                                                                                          \
            std::snprintf(c->err, sizeof(c->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                          cudaGetErrorString(e_));                                                 \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

} // namespace lpl
