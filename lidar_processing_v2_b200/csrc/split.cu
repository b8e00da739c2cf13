// Processor glue on the device: what the reference node does on the host between
// and after the library calls.
//
//   label split      src/processor/src/processor.cpp:562-579  three clouds of pcl::PointXYZRGB in cloud order:
//                    GROUND (124, 252, 0), OBSTACLE (200, 0, 0), everything else (255, 255, 0)
//   clustered cloud  src/processor/src/processor.cpp:627-647  for every label ascending, the cluster's points in
//                    obstacle-cloud order, one colour per cluster (the node draws r, g, b = std::rand() % 256)
//   marker lines     src/processor/src/processor.cpp:206-343  convertPolygonPointsToMarker: LINE_LIST vertices of the
//                    bottom ring, the top ring and the vertical edges of every hull with >= 3 vertices
//
// A record is the 32-byte pcl::PointXYZRGB: x, y, z, 1.0f | b, g, r, a = 255 | 12 bytes of padding (zeroed here).
// All outputs are stable compactions / segmented gathers: the same order the node's host loops produce.
#include "common.cuh"

namespace lpl
{
struct RgbRecord
{
    float x, y, z, w;
    std::uint32_t bgra;
    std::uint32_t pad[3];
};
static_assert(sizeof(RgbRecord) == 32, "pcl::PointXYZRGB is 32 bytes");

__device__ __forceinline__ std::uint32_t pack_bgra(std::uint32_t r, std::uint32_t g, std::uint32_t b)
{
    return b | (g << 8) | (r << 16) | (255u << 24);
}

__device__ __forceinline__ void store_record(RgbRecord* out, const float4& p, std::uint32_t bgra)
{
    // two 16-byte stores per record
    reinterpret_cast<float4*>(out)[0] = make_float4(p.x, p.y, p.z, 1.0f);
    reinterpret_cast<uint4*>(out)[1] = make_uint4(bgra, 0u, 0u, 0u);
}

// per 2048-point tile: how many GROUND / OBSTACLE / other points
__global__ void __launch_bounds__(kTileThreads) k_split_count(Dev d, std::uint32_t* __restrict__ cnt3)
{
    __shared__ std::uint32_t sh[33];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t base = blockIdx.x * kTile;
    std::uint32_t c[3] = {0u, 0u, 0u};
    if (base < n)
    {
        const std::uint8_t* lab = d.labels_out + static_cast<std::size_t>(f) * d.cap;
#pragma unroll
        for (int j = 0; j < kItems; ++j)
        {
            const std::uint32_t i = base + j * kTileThreads + threadIdx.x;
            if (i < n)
            {
                const std::uint8_t l = lab[i];
                c[l == PX_GROUND ? 0 : (l == PX_OBSTACLE ? 1 : 2)] += 1u;
            }
        }
    }
    for (int k = 0; k < 3; ++k)
    {
        const std::uint32_t s = block_sum(c[k], sh);
        if (threadIdx.x == 0)
        {
            cnt3[(static_cast<std::size_t>(f) * d.tiles + blockIdx.x) * 3 + k] = s;
        }
        __syncthreads();
    }
}

// stable scatter of every point into the cloud of its label; out3[k] = start of cloud k of frame 0, frames
// `frame_stride` records apart; totals[f][k] receives the cloud sizes
__global__ void __launch_bounds__(kTileThreads)
    k_split_scatter(Dev d, const std::uint32_t* __restrict__ cnt3, RgbRecord* out_g, RgbRecord* out_o, RgbRecord* out_u,
                    std::size_t frame_stride, std::uint32_t* __restrict__ totals)
{
    __shared__ std::uint32_t sh[kItems * (kTileThreads / 32) + 1];
    __shared__ std::uint32_t sh2[33];
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t base = blockIdx.x * kTile;
    const std::uint32_t* tc = cnt3 + static_cast<std::size_t>(f) * d.tiles * 3;
    std::uint32_t before[3];
    for (int k = 0; k < 3; ++k)
    {
        std::uint32_t b = 0, all = 0;
        for (std::uint32_t t = threadIdx.x; t < d.tiles; t += kTileThreads)
        {
            const std::uint32_t v = tc[t * 3 + k];
            all += v;
            b += t < blockIdx.x ? v : 0u;
        }
        before[k] = block_sum(b, sh2);
        if (blockIdx.x == 0)
        {
            all = block_sum(all, sh2);
            if (threadIdx.x == 0)
            {
                totals[f * 4 + k] = all;
            }
        }
    }
    if (base >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::uint8_t cls[kItems];
    float4 p[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j)
    {
        const std::uint32_t i = base + j * kTileThreads + threadIdx.x;
        cls[j] = 3;
        if (i < n)
        {
            const std::uint8_t l = d.labels_out[o + i];
            cls[j] = l == PX_GROUND ? 0 : (l == PX_OBSTACLE ? 1 : 2);
            p[j] = d.pts_in[o + i];
        }
    }
    RgbRecord* outs[3] = {out_g + f * frame_stride, out_o + f * frame_stride, out_u + f * frame_stride};
    const std::uint32_t colour[3] = {pack_bgra(124, 252, 0), pack_bgra(200, 0, 0), pack_bgra(255, 255, 0)}; // processor.cpp:568-578
    for (int k = 0; k < 3; ++k)
    {
        bool flag[kItems];
        std::uint32_t rank[kItems];
#pragma unroll
        for (int j = 0; j < kItems; ++j)
        {
            flag[j] = cls[j] == k;
        }
        std::uint32_t total;
        tile_ranks(flag, rank, &total, sh);
#pragma unroll
        for (int j = 0; j < kItems; ++j)
        {
            if (flag[j])
            {
                store_record(outs[k] + before[k] + rank[j], p[j], colour[k]);
            }
        }
    }
}

// sort elements (cluster label, 0, 0, obstacle index) of the clustered points of a frame, for the frame-wide merge
// sort of hull.cu: afterwards position p holds the p-th point of the clustered cloud
struct ClusteredPred
{
    const std::int32_t* clabel;
    std::uint32_t cap;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const { return clabel[static_cast<std::size_t>(f) * cap + i] >= 0; }
};

struct ClusteredEmit
{
    const std::int32_t* clabel;
    uint4* hsB;
    std::uint32_t cap;
    __device__ void operator()(std::uint32_t f, std::uint32_t i, std::uint32_t pos) const
    {
        const std::size_t o = static_cast<std::size_t>(f) * cap;
        hsB[o + pos] = make_uint4(static_cast<std::uint32_t>(clabel[o + i]), 0u, 0u, i);
    }
};

__global__ void __launch_bounds__(256)
    k_split_clustered(Dev d, const std::uint8_t* __restrict__ colours, std::uint32_t colour_stride, RgbRecord* out,
                      std::size_t frame_stride, std::uint32_t* __restrict__ totals)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_h[f];
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        totals[f * 4 + 3] = n;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const bool in_b = (sort_passes(n) & 1u) != 0u;
    const uint4* sorted = (in_b ? d.hsB : d.hsA) + o;
    for (std::uint32_t p = blockIdx.x * 256u + threadIdx.x; p < n; p += gridDim.x * 256u)
    {
        const uint4 e = sorted[p];
        const std::uint8_t* c = colours + static_cast<std::size_t>(f) * colour_stride + static_cast<std::size_t>(e.x) * 3u;
        store_record(out + f * frame_stride + p, d.pts_o[o + e.w], pack_bgra(c[0], c[1], c[2]));
    }
}

// LINE_LIST vertices of the hulls with >= 3 vertices (processor.cpp:254-343): per hull of n vertices 6 n points of
// three doubles: bottom ring (2 n), top ring (2 n), vertical edges (2 n). marker_off[k] = first vertex of hull k.
__global__ void __launch_bounds__(128) k_marker_count(Dev d, std::uint32_t* __restrict__ mcount)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    for (std::uint32_t c = blockIdx.x * 128u + threadIdx.x; c < K; c += gridDim.x * 128u)
    {
        const std::uint32_t n = hoff[c + 1] - hoff[c];
        mcount[static_cast<std::size_t>(f) * d.cap + c] = n >= 3u ? 6u * n : 0u;
    }
}

__global__ void __launch_bounds__(128)
    k_marker_lines(Dev d, const std::uint32_t* __restrict__ moff, double* __restrict__ out, std::size_t frame_stride)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t K = d.n_clusters[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t* mo = moff + static_cast<std::size_t>(f) * (d.cap + 1);
    const std::uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (std::uint32_t c = blockIdx.x * 4u + warp; c < K; c += gridDim.x * 4u)
    {
        const std::uint32_t a = hoff[c], n = hoff[c + 1] - a;
        if (n < 3u)
        {
            continue;
        }
        const float2 zz = d.zminmax[o + c];
        const double z_min = static_cast<double>(zz.x), z_max = static_cast<double>(zz.y);
        double* dst = out + (f * frame_stride + mo[c]) * 3;
        for (std::uint32_t i = lane; i < n; i += 32)
        {
            // segment i of a ring runs from vertex i - 1 to vertex i, the closing segment (last -> first) comes last
            const float2 cur = d.hull_xy[o + a + i];
            const float2 nxt = d.hull_xy[o + a + (i + 1u == n ? 0u : i + 1u)];
            const double cx = static_cast<double>(cur.x), cy = static_cast<double>(cur.y);
            const double nx = static_cast<double>(nxt.x), ny = static_cast<double>(nxt.y);
            // segment (cur -> nxt) is the (i + 1)-th of the ring when i + 1 < n, the closing one otherwise: both at
            // slot i of the "i = 1 .. n - 1, then closing" order the node emits (slot s holds vertices s, s + 1)
            const std::uint32_t slot = i; // segment starting at vertex i
            double* b = dst + static_cast<std::size_t>(2u * slot) * 3;
            b[0] = cx; b[1] = cy; b[2] = z_min; b[3] = nx; b[4] = ny; b[5] = z_min;
            double* t = dst + static_cast<std::size_t>(2u * n + 2u * slot) * 3;
            t[0] = cx; t[1] = cy; t[2] = z_max; t[3] = nx; t[4] = ny; t[5] = z_max;
            double* v = dst + static_cast<std::size_t>(4u * n + 2u * i) * 3;
            v[0] = cx; v[1] = cy; v[2] = z_min; v[3] = cx; v[4] = cy; v[5] = z_max;
        }
    }
}

void launch_split_clouds(Ctx* c, std::uint32_t nf, unsigned char* out, std::size_t frame_records, const std::uint8_t* colours,
                         std::uint32_t colour_stride, std::uint32_t* cnt3, std::uint32_t* totals)
{
    Dev& d = c->d;
    auto* rec = reinterpret_cast<RgbRecord*>(out);
    // layout of `out`: four planes of nf x frame_records records: ground, obstacle, unsegmented, clustered
    const std::size_t plane = static_cast<std::size_t>(nf) * frame_records;
    const dim3 grid(d.tiles, nf);
    k_split_count<<<grid, kTileThreads, 0, c->stream>>>(d, cnt3);
    mark(c, "split_count");
    k_split_scatter<<<grid, kTileThreads, 0, c->stream>>>(d, cnt3, rec, rec + plane, rec + 2 * plane, frame_records, totals);
    mark(c, "split_scatter");
    launch_compact(c, "split_clustered_keys", nf, d.tiles, d.n_o, 0u, d.tile_cnt, d.n_h, ClusteredPred{d.clabel, d.cap},
                   ClusteredEmit{d.clabel, d.hsB, d.cap});
    launch_hull_sort(c, nf);
    k_split_clustered<<<dim3(per_frame_ctas(32, nf, 512), nf), 256, 0, c->stream>>>(d, colours, colour_stride, rec + 3 * plane,
                                                                                  frame_records, totals);
    mark(c, "split_clustered");
}

void launch_marker_lines(Ctx* c, std::uint32_t nf, std::uint32_t* mcount, std::uint32_t* moff, std::uint32_t* mtotal, double* out,
                         std::size_t frame_stride)
{
    Dev& d = c->d;
    k_marker_count<<<dim3(8, nf), 128, 0, c->stream>>>(d, mcount);
    mark(c, "marker_count");
    k_excl_scan<<<nf, 1024, 0, c->stream>>>(mcount, d.cap, moff, d.cap + 1, d.cap, d.n_clusters, mtotal);
    mark(c, "marker_scan");
    k_marker_lines<<<dim3(64, nf), 128, 0, c->stream>>>(d, moff, out, frame_stride);
    mark(c, "marker_lines");
}
} // namespace lpl
