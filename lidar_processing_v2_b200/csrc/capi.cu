// C ABI (include/lpl_b200.h): context, configuration, staging and stage orchestration.
// No torch types, no CPU fallback: every entry point fails with LPL_ERR_NO_DEVICE / LPL_ERR_CUDA
// when the GPU is not usable.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <random>
#include <vector>

#include "../../include/lpl_b200.h"
#include "common.cuh"

using namespace lpl;

struct lpl_ctx
{
    Ctx c;
    lpl_segmenter_cfg seg_cfg;
    lpl_dror_cfg dror_cfg;
    lpl_cluster_cfg clu_cfg;
    int ncell_cap = 0;
    int want_image = 0;
    int jcp_mode = LPL_JCP_AS_REFERENCE;
    int have_ring = 0; // ring plane of the current batch is meaningful
    bool ran_hulls = false; // the last lpl_pipeline_run included LPL_STAGE_HULLS (hull_off is this batch's)
    // CUDA graphs of lpl_pipeline_run, keyed by (frames, stages, image / ring flags); cfg_epoch moves with every
    // configuration call because kernel arguments (SegParams, ...) are baked into a captured graph
    struct RunGraph
    {
        unsigned long long key;
        unsigned long long epoch;
        cudaGraphExec_t exec;
        unsigned long long launches;
        int use_ring;
    };
    std::vector<RunGraph> graphs;
    // device scratch of lpl_pipeline_split_clouds (allocated on first use) and the replica of the C library's
    // rand() stream the node colours its clusters with (processor.cpp:629-631)
    void* split_dev = nullptr;
    std::size_t split_bytes = 0;
    // lpl_knn_*: the searched point set lives in frame 0 of the input plane; per-chunk query / result scratch
    std::uint32_t knn_n = 0;
    bool knn_built = false;          // false again as soon as anything else writes the input plane
    unsigned long long knn_token = 0; // identifies the resident point set (lpl_knn_token)
    void* knn_dev = nullptr;
    std::size_t knn_bytes = 0;
    std::int32_t rand_r[34] = {0};
    int rand_i = -1; // -1: not seeded yet
    unsigned long long cfg_epoch = 0;
    bool use_graphs = true;
    // sub-batches of one run on concurrent streams (enqueue_stages): streams and join events, created on first use
    struct Aux
    {
        cudaStream_t stream;
        cudaEvent_t done;
    };
    std::vector<Aux> aux;
    cudaEvent_t fork_ev = nullptr;
    std::uint32_t split_parts = 3;
    std::vector<std::uint32_t> h_status;
};

namespace
{
constexpr std::size_t kAlign = 256;

struct Carver
{
    std::size_t off = 0;
    char* base = nullptr;
    template <typename T>
    void take(T*& p, std::size_t count)
    {
        off = (off + kAlign - 1) / kAlign * kAlign;
        if (base != nullptr)
        {
            p = reinterpret_cast<T*>(base + off);
        }
        off += count * sizeof(T);
    }
};

void carve(Dev& d, Carver& cv, int npx, int ncell_cap, std::uint32_t** mt_raw)
{
    const std::size_t B = d.B, cap = d.cap, q = d.qcap, h = d.hcap;
    cv.take(d.pts_in, B * cap);
    cv.take(d.n_in, B);
    cv.take(d.ring, B * cap);
    cv.take(d.wrap_cnt, B * (cap / 32));
    cv.take(d.noise, B * cap);
    cv.take(d.grid_cnt, B * kDrorCells);
    cv.take(d.grid_start, B * (kDrorCells + 1));
    cv.take(d.grid_mask, B * (kDrorCells / 32));
    cv.take(d.grid_pts, B * cap);
    cv.take(d.unres, B * cap);       // = slot = vslot = hfin (see below)
    cv.take(d.cell, B * cap);
    cv.take(d.px, B * cap);          // = vlist
    cv.take(d.cell_cnt, B * ncell_cap);
    cv.take(d.cell_start, B * (ncell_cap + 1));
    cv.take(d.n_binned, B);
    cv.take(d.zo, B * cap);          // = hstack
    cv.take(d.zsort, B * cap);       // = hseg_cnt
    cv.take(d.ccnt, B * ncell_cap);
    cv.take(d.cell_zmin, B * ncell_cap);
    cv.take(d.elev, B * ncell_cap);
    cv.take(d.lab, B * cap);
    cv.take(d.n_cand, B);
    cv.take(d.pairs, B * kRansacIters * 2);
    cv.take(d.planes, B * kRansacIters);
    cv.take(d.inliers, B * kRansacIters);
    cv.take(d.best_plane, B);
    cv.take(d.best_cnt, B);
    cv.take(d.key, B * npx);
    cv.take(d.pxidx, B * npx);
    cv.take(d.code, B * npx);
    cv.take(d.queue, B * q);
    cv.take(d.n_queue, B);
    cv.take(d.wn, B * 24 * q);
    cv.take(d.mk, B * q);
    cv.take(d.stale_ref, B * d.nborder_cap * 12);
    cv.take(d.jcp_rounds, B);
    cv.take(d.labels_out, B * cap);
    cv.take(d.bgr, B * npx * 3);
    cv.take(d.pts_o, B * cap);
    cv.take(d.idx_o, B * cap);
    cv.take(d.hkey, B * h);
    cv.take(d.hparent, B * h);
    cv.take(d.hmin, B * h);
    cv.take(d.hcount, B * h);
    cv.take(d.hlabel, B * h);
    cv.take(d.hroot, B * h);
    cv.take(d.hvid, B * h);
    cv.take(d.edges, B * cap * kEdgePitch); // = octa
    cv.take(d.clabel, B * cap);
    cv.take(d.ccount, B * cap);
    cv.take(d.cstart, B * (cap + 1));
    cv.take(d.hsB, B * cap);
    cv.take(d.hcnt, B * cap);
    cv.take(d.hull_off, B * (cap + 1));
    cv.take(d.hull_idx, B * cap);
    cv.take(d.hull_xy, B * cap);
    cv.take(d.zminmax, B * cap);
    cv.take(d.zmin_u, B * cap);
    cv.take(d.zmax_u, B * cap);
    cv.take(d.zzero, B * cap);
    cv.take(d.ext, B * cap * kExtDirs);
    cv.take(d.n_h, B);
    cv.take(d.hull_next, B);
    cv.take(d.hwk_off, B * (cap + 1));
    cv.take(d.hck_cnt, B * 2 * cap);
    cv.take(d.n_work, B);
    cv.take(d.n_multi, B);
    cv.take(d.boxes, B * cap);
    cv.take(d.raw_desc, B * 32);
    const std::size_t tl = std::max<std::size_t>(d.tiles, d.ptiles);
    cv.take(d.tile_cnt, B * tl);
    // counter block (one memset per run): every array below is zero at the start of a run
    cv.take(d.status, B);
    const std::size_t ctr_first = cv.off - B * sizeof(std::uint32_t);
    cv.take(d.n_v, B);
    cv.take(d.n_o, B);
    cv.take(d.n_clusters, B);
    cv.take(d.n_hull, B);
    cv.take(d.n_unres, B);
    cv.take(d.n_cpts, B);
    cv.take(d.n_border, B);
    cv.take(d.n_vox, B);
    cv.take(d.sph_max, B * 4);
    d.ctr_begin = cv.base != nullptr ? reinterpret_cast<unsigned char*>(cv.base) + ctr_first : nullptr;
    d.ctr_bytes = cv.off - ctr_first;
    cv.take(*mt_raw, kMtRaws);
    // Stage-local scratch shares storage. A frame's chain runs its stages in order (S1 DROR, S2 segmentation, S3
    // clustering, S4 hulls), and every pair below has the same element size AND the same per-frame stride, so frame f
    // of one plane is frame f of the other: safe for sub-batches of one run that are at different stages, and for the
    // stage entry points, which use these planes as scratch only.
    //   16 B / point: grid_pts (S1) = cpts (S2) = sph (S3) = hsA (S4)
    //    4 B / point: unres (S1) = slot (S2) = vslot (S3) = hfin (S4);  px (S2) = vlist (S3);  zsort (S2) = hseg_cnt (S4)
    //    8 B / point: zo (S2) = hstack (S4)
    //   64 B / point: edges (S3, rows of 13 padded to 16) = octa (S4) = raw (the 32 B / point staging of the packed
    //                 uploads and downloads and of raw PointCloud2 records: consumed before / filled after a run)
    d.cpts = d.grid_pts;
    d.sph = d.grid_pts;
    d.hsA = reinterpret_cast<uint4*>(d.grid_pts);
    d.slot = d.unres;
    d.vslot = d.unres;
    d.hfin = d.unres;
    d.vlist = d.px;
    d.hstack = reinterpret_cast<std::uint32_t*>(d.zo);
    d.hseg_cnt = reinterpret_cast<std::uint32_t*>(d.zsort);
    d.octa = reinterpret_cast<float2*>(d.edges);
    d.raw = reinterpret_cast<unsigned char*>(d.edges);
    static_assert(kRawRecord <= kEdgePitch * sizeof(std::uint32_t), "the raw-record staging fits the edge plane");
    static_assert(kEdgePitch * sizeof(std::uint32_t) == kExtDirs * sizeof(float2) || kExtDirs != 8, "edges and octa share a plane");
}

void drop_graphs(lpl_ctx* ctx)
{
    for (auto& g : ctx->graphs)
    {
        if (g.exec != nullptr)
        {
            cudaGraphExecDestroy(g.exec);
        }
    }
    ctx->graphs.clear();
}

int fail(lpl_ctx* ctx, int code, const char* msg)
{
    if (ctx != nullptr)
    {
        std::snprintf(ctx->c.err, sizeof(ctx->c.err), "%s", msg);
    }
    return code;
}

#define LPL_TRY(call)                                                                              \
    do                                                                                             \
    {                                                                                              \
        const cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                                     \
        {                                                                                          \
            std::snprintf(ctx->c.err, sizeof(ctx->c.err), "%s:%d %s: %s", __FILE__, __LINE__,      \
                          #call, cudaGetErrorString(e_));                                          \
            return LPL_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int ensure_stage(lpl_ctx* ctx, std::size_t bytes)
{
    Ctx& c = ctx->c;
    if (c.h_stage_bytes >= bytes)
    {
        return 0;
    }
    if (c.h_stage != nullptr)
    {
        LPL_TRY(cudaStreamSynchronize(c.stream));
        cudaFreeHost(c.h_stage);
        c.h_stage = nullptr;
        c.h_stage_bytes = 0;
    }
    bytes = (bytes + (1u << 20)) & ~((std::size_t(1) << 20) - 1);
    LPL_TRY(cudaMallocHost(&c.h_stage, bytes));
    c.h_stage_bytes = bytes;
    return 0;
}

// derived constants, in the reference's own float expressions (segmenter.cpp:42-46,106-109,
// 121-122,209-211,220-221,359-360,532-533)
int apply_seg_cfg(lpl_ctx* ctx)
{
    const lpl_segmenter_cfg& g = ctx->seg_cfg;
    SegParams& s = ctx->c.seg;
    const float kD2R = static_cast<float>(M_PI / 180.0);
    const float kTwoPi = static_cast<float>(2.0 * M_PI);
    if (g.image_height != s.H || g.image_width != s.W)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT,
                    "image_height/image_width differ from the size the context was created for");
    }
    if (!(g.grid_radial_spacing_m > 0.f) || !(g.grid_slice_resolution_deg > 0.f))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "grid spacing / slice resolution must be positive");
    }
    const float slice_res = g.grid_slice_resolution_deg * kD2R;
    const int rings = static_cast<std::int32_t>(g.max_distance_m / g.grid_radial_spacing_m);
    const int slices = static_cast<std::int32_t>(kTwoPi / slice_res);
    if (rings <= 0 || slices <= 0 || static_cast<long long>(rings) * slices > ctx->ncell_cap)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "polar grid has no cells or more cells than reserved (262144)");
    }
    // range-image keys carry (depth^2 bits : 31, azimuth slice, point index) in 64 bits
    int idx_bits = 11, slice_bits = 0;
    while ((1u << idx_bits) < ctx->c.d.cap)
    {
        ++idx_bits;
    }
    while ((1 << slice_bits) < slices)
    {
        ++slice_bits;
    }
    if (idx_bits + slice_bits > 33)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "azimuth slices x points per frame exceed the 33-bit tie-break field of the range image");
    }
    s.idx_bits = idx_bits;
    s.nb = std::min(rings, kRansacBins);
    s.rings = rings;
    s.slices = slices;
    s.ncell = rings * slices;
    s.radial_spacing = g.grid_radial_spacing_m;
    s.min_dist = g.min_distance_m;
    s.max_dist = g.max_distance_m;
    s.slice_res = slice_res;
    const float el_up = g.elevation_up_deg * kD2R;
    const float el_down = g.elevation_down_deg * kD2R;
    const float vfov = el_up - el_down;
    s.el_down = el_down;
    s.rad_per_px = vfov / g.image_height;
    s.z_lo = -g.sensor_height_m + g.z_min_m;
    s.z_hi = -g.sensor_height_m + g.z_max_m;
    s.thr = g.ground_height_threshold_m;
    s.thr2 = 2.0F * g.ground_height_threshold_m;
    s.delta = std::min(g.grid_radial_spacing_m * std::tan(g.road_maximum_slope_m_per_m),
                       g.ground_height_threshold_m - std::numeric_limits<float>::epsilon());
    s.e0 = -g.sensor_height_m + g.ground_height_threshold_m;
    s.cos_max = std::cos(std::tan(g.road_maximum_slope_m_per_m));
    s.p1z = -g.sensor_height_m;
    s.kthr_sqr = g.kernel_threshold_distance_m * g.kernel_threshold_distance_m;
    s.amp = g.amplification_factor;
    s.wscale = static_cast<float>(g.image_width - 1);
    s.jcp_emulate_stale = (ctx->jcp_mode == LPL_JCP_AS_REFERENCE) ? 1 : 0;
    return 0;
}

void apply_dror_cfg(lpl_ctx* ctx)
{
    // noise_remover.cpp:46-49
    ctx->c.dror.scaling = std::pow(static_cast<double>(ctx->dror_cfg.radius_multiplier_m_per_m), 2.0);
    ctx->c.dror.min_r_sqr = ctx->dror_cfg.min_search_radius_m * ctx->dror_cfg.min_search_radius_m;
    ctx->c.dror.min_neighbours = ctx->dror_cfg.min_neighbours;
}

void apply_cluster_cfg(lpl_ctx* ctx)
{
    // clusterer.hpp:79-86
    const float kD2R = static_cast<float>(M_PI / 180.0);
    ctx->c.clu.range_res = ctx->clu_cfg.voxel_grid_range_resolution_m;
    ctx->c.clu.az_res = ctx->clu_cfg.voxel_grid_azimuth_resolution_deg * kD2R;
    ctx->c.clu.el_res = ctx->clu_cfg.voxel_grid_elevation_resolution_deg * kD2R;
    ctx->c.clu.min_cluster_size = ctx->clu_cfg.min_cluster_size;
}

// pack n strided records (first `take` bytes of each) into pinned staging at `dst`
void pack(void* dst, std::size_t dst_pitch, const void* src, std::size_t stride, std::size_t take, std::uint32_t n)
{
    auto* o = static_cast<char*>(dst);
    const auto* s = static_cast<const char*>(src);
    if (stride == take && dst_pitch == take)
    {
        std::memcpy(o, s, static_cast<std::size_t>(n) * take);
        return;
    }
    for (std::uint32_t i = 0; i < n; ++i)
    {
        std::memcpy(o + i * dst_pitch, s + i * stride, take);
    }
}

// Syncs the stream and reads the per-frame status words. Returns LPL_ERR_CAPACITY (message names the first
// flagged frame) when any frame raised a capacity bit; ctx->h_status keeps all of them (lpl_pipeline_status).
int check_status(lpl_ctx* ctx, std::uint32_t nf)
{
    Ctx& c = ctx->c;
    ctx->h_status.resize(nf);
    LPL_TRY(cudaMemcpyAsync(ctx->h_status.data(), c.d.status, sizeof(std::uint32_t) * nf, cudaMemcpyDeviceToHost,
                            c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        const std::uint32_t s = ctx->h_status[f];
        if (s != 0)
        {
            c.hash_clean = false; // an aborted batch may leave voxel hash slots behind: clear all next time
            std::snprintf(c.err, sizeof(c.err),
                          "frame %u exceeded a reserved capacity:%s%s%s%s", f,
                          (s & ST_QUEUE_OVERFLOW) ? " JCP queue" : "", (s & ST_RNG_EXHAUSTED) ? " RANSAC RNG table" : "",
                          (s & ST_HASH_FULL) ? " voxel hash" : "",
                          (s & (ST_BORDER_OVERFLOW | ST_JCP_STALL)) ? " JCP border rows / sweep stall" : "");
            return LPL_ERR_CAPACITY;
        }
    }
    return 0;
}

__global__ void k_label_count(Dev d, std::uint32_t K)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_o[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::int32_t l = -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < n)
    {
        l = d.clabel[o + i];
        if (l >= 0 && static_cast<std::uint32_t>(l) < K)
        {
            atomicAdd(&d.ccount[o + l], 1u);
            const float4 p = d.pts_o[o + i];
            x = p.x;
            y = p.y;
            z = p.z;
        }
        else
        {
            l = -1;
            d.clabel[o + i] = -1;
        }
    }
    accumulate_cluster_stats(d.ext + o * kExtDirs, d.zmin_u + o, d.zmax_u + o, d.zzero + o, l, x, y, z, i, true);
}

__global__ void k_ext_init(Dev d, std::uint32_t K)
{
    const std::uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c < K)
    {
        ext_init(d.ext + static_cast<std::size_t>(c) * kExtDirs);
    }
}
} // namespace

extern "C"
{
const char* lpl_version(void) { return "lpl_b200 0.1 (sm_100a)"; }

void lpl_segmenter_default_cfg(lpl_segmenter_cfg* g)
{
    // segmenter.hpp:87-112
    g->elevation_up_deg = 2.0F;
    g->elevation_down_deg = -24.8F;
    g->image_width = 2048;
    g->image_height = 64;
    g->assume_unorganized_cloud = 0;
    g->grid_radial_spacing_m = 2.0F;
    g->grid_slice_resolution_deg = 1.0F;
    g->ground_height_threshold_m = 0.2F;
    g->road_maximum_slope_m_per_m = 0.2F;
    g->min_distance_m = 2.0F;
    g->max_distance_m = 100.0F;
    g->sensor_height_m = 1.73F;
    g->kernel_threshold_distance_m = 1.0F;
    g->amplification_factor = 5.0F;
    g->z_min_m = -3.0F;
    g->z_max_m = 4.0F;
}

void lpl_dror_default_cfg(lpl_dror_cfg* g)
{
    // noise_remover.hpp:41-54
    g->radius_multiplier_m_per_m = 0.02F;
    g->min_search_radius_m = 0.1F;
    g->min_neighbours = 4U;
}

void lpl_cluster_default_cfg(lpl_cluster_cfg* g)
{
    // clusterer.hpp:61-68
    g->voxel_grid_range_resolution_m = 0.4F;
    g->voxel_grid_azimuth_resolution_deg = 1.0F;
    g->voxel_grid_elevation_resolution_deg = 1.5F;
    g->min_cluster_size = 3;
}

int lpl_create(lpl_ctx** out, int device, std::uint32_t max_points, std::uint32_t max_frames,
               std::int32_t image_height, std::int32_t image_width)
{
    if (out == nullptr || max_points == 0 || max_frames == 0)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    {
        return LPL_ERR_NO_DEVICE;
    }
    lpl_ctx* ctx = new (std::nothrow) lpl_ctx();
    if (ctx == nullptr)
    {
        return LPL_ERR_CUDA;
    }
    Ctx& c = ctx->c;
    c.device = device;
    const int H = image_height > 0 ? image_height : 64;
    const int W = image_width > 0 ? image_width : 2048;
    if ((static_cast<long long>(H) * W) % 16 != 0)
    {
        delete ctx;
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Dev& d = c.d;
    d.cap = (max_points + kTile - 1) / kTile * kTile;
    d.tiles = d.cap / kTile;
    d.B = max_frames;
    const int npx = H * W;
    d.ptiles = (npx + kTile - 1) / kTile;
    // the reference reserves H * W queue entries (segmenter.cpp: index_queue_.reserve): wall scenes queue well
    // over half of the image
    d.qcap = static_cast<std::uint32_t>(npx);
    // k_jcp_rows keeps the 2-bit state plane of one frame (npx / 4 bytes) + H + 1 row starts in shared memory
    if (static_cast<std::size_t>(npx) / 4 + sizeof(std::uint32_t) * (H + 2) + 4096 > 227u * 1024u)
    {
        delete ctx;
        return LPL_ERR_CAPACITY;
    }
    std::uint32_t h = 1024;
    while (h < 2u * d.cap)
    {
        h <<= 1;
    }
    d.hcap = h;
    d.nborder_cap = static_cast<std::uint32_t>(4 * W + 4 * H);
    ctx->ncell_cap = 262144;
    c.seg.H = H;
    c.seg.W = W;
    c.seg.npx = npx;
    auto bail = [&](int code) {
        lpl_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess)
    {
        return bail(LPL_ERR_CUDA);
    }
    // experiment hook (tools/overlap_probe.py): LPL_STREAM_PRIORITY = 0 (default) .. -5 (highest on B200)
    int prio = 0;
    if (const char* e = std::getenv("LPL_STREAM_PRIORITY"))
    {
        prio = std::atoi(e);
    }
    if (cudaStreamCreateWithPriority(&c.stream, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreate(&c.ev0) != cudaSuccess || cudaEventCreate(&c.ev1) != cudaSuccess)
    {
        return bail(LPL_ERR_CUDA);
    }
    std::uint32_t* mt_raw = nullptr;
    Carver measure;
    carve(d, measure, npx, ctx->ncell_cap, &mt_raw);
    c.slab_bytes = measure.off + kAlign;
    if (cudaMalloc(&c.slab, c.slab_bytes) != cudaSuccess)
    {
        cudaGetLastError();
        return bail(LPL_ERR_CAPACITY);
    }
    if (cudaMemsetAsync(c.slab, 0, c.slab_bytes, c.stream) != cudaSuccess)
    {
        return bail(LPL_ERR_CUDA);
    }
    Carver real;
    real.base = static_cast<char*>(c.slab);
    carve(d, real, npx, ctx->ncell_cap, &mt_raw);
    d.mt_raw = mt_raw;
    // std::mt19937{42} raw stream (segmenter.cpp:369); frame independent
    {
        std::vector<std::uint32_t> raws(kMtRaws);
        std::mt19937 gen{42};
        for (auto& r : raws)
        {
            r = static_cast<std::uint32_t>(gen());
        }
        if (cudaMemcpyAsync(mt_raw, raws.data(), sizeof(std::uint32_t) * kMtRaws, cudaMemcpyHostToDevice,
                            c.stream) != cudaSuccess ||
            cudaStreamSynchronize(c.stream) != cudaSuccess)
        {
            return bail(LPL_ERR_CUDA);
        }
    }
    ctx->use_graphs = std::getenv("LPL_NO_GRAPH") == nullptr;
    if (const char* sp = std::getenv("LPL_SPLIT"))
    {
        const int v = std::atoi(sp);
        ctx->split_parts = v >= 1 && v <= 8 ? static_cast<std::uint32_t>(v) : 3u;
    }
    lpl_segmenter_default_cfg(&ctx->seg_cfg);
    ctx->seg_cfg.image_height = H;
    ctx->seg_cfg.image_width = W;
    lpl_dror_default_cfg(&ctx->dror_cfg);
    lpl_cluster_default_cfg(&ctx->clu_cfg);
    if (apply_seg_cfg(ctx) != 0)
    {
        return bail(LPL_ERR_INVALID_ARGUMENT);
    }
    apply_dror_cfg(ctx);
    apply_cluster_cfg(ctx);
    *out = ctx;
    return LPL_OK;
}

void lpl_destroy(lpl_ctx* ctx)
{
    if (ctx == nullptr)
    {
        return;
    }
    Ctx& c = ctx->c;
    cudaSetDevice(c.device);
    if (c.stream != nullptr)
    {
        cudaStreamSynchronize(c.stream);
    }
    drop_graphs(ctx);
    if (ctx->split_dev != nullptr)
    {
        cudaFree(ctx->split_dev);
    }
    if (ctx->knn_dev != nullptr)
    {
        cudaFree(ctx->knn_dev);
    }
    if (c.slab != nullptr)
    {
        cudaFree(c.slab);
    }
    if (c.h_stage != nullptr)
    {
        cudaFreeHost(c.h_stage);
    }
    if (c.ev0 != nullptr)
    {
        cudaEventDestroy(c.ev0);
    }
    if (c.ev1 != nullptr)
    {
        cudaEventDestroy(c.ev1);
    }
    for (cudaEvent_t e : c.prof_ev)
    {
        if (e != nullptr)
        {
            cudaEventDestroy(e);
        }
    }
    for (auto& a : ctx->aux)
    {
        cudaEventDestroy(a.done);
        cudaStreamDestroy(a.stream);
    }
    if (ctx->fork_ev != nullptr)
    {
        cudaEventDestroy(ctx->fork_ev);
    }
    if (c.stream != nullptr)
    {
        cudaStreamDestroy(c.stream);
    }
    delete ctx;
}

const char* lpl_last_error(const lpl_ctx* ctx) { return ctx != nullptr ? ctx->c.err : "null context"; }

size_t lpl_device_bytes(const lpl_ctx* ctx) { return ctx != nullptr ? ctx->c.slab_bytes : 0; }

int lpl_segmenter_config(lpl_ctx* ctx, const lpl_segmenter_cfg* cfg)
{
    if (ctx != nullptr)
    {
        ctx->cfg_epoch += 1; // captured graphs carry the old parameters
    }
    if (ctx == nullptr || cfg == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    const lpl_segmenter_cfg old = ctx->seg_cfg;
    ctx->seg_cfg = *cfg;
    const int rc = apply_seg_cfg(ctx);
    if (rc != 0)
    {
        ctx->seg_cfg = old;
        apply_seg_cfg(ctx);
    }
    return rc;
}

int lpl_dror_config(lpl_ctx* ctx, const lpl_dror_cfg* cfg)
{
    if (ctx != nullptr)
    {
        ctx->cfg_epoch += 1; // captured graphs carry the old parameters
    }
    if (ctx == nullptr || cfg == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    ctx->dror_cfg = *cfg;
    apply_dror_cfg(ctx);
    return LPL_OK;
}

int lpl_cluster_config(lpl_ctx* ctx, const lpl_cluster_cfg* cfg)
{
    if (ctx != nullptr)
    {
        ctx->cfg_epoch += 1; // captured graphs carry the old parameters
    }
    if (ctx == nullptr || cfg == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    if (!(cfg->voxel_grid_range_resolution_m > 0.f) || !(cfg->voxel_grid_azimuth_resolution_deg > 0.f) ||
        !(cfg->voxel_grid_elevation_resolution_deg > 0.f))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "voxel resolutions must be positive");
    }
    ctx->clu_cfg = *cfg;
    apply_cluster_cfg(ctx);
    return LPL_OK;
}

int lpl_set_jcp_mode(lpl_ctx* ctx, int mode)
{
    if (ctx != nullptr)
    {
        ctx->cfg_epoch += 1; // captured graphs carry the old parameters
    }
    if (ctx == nullptr || (mode != LPL_JCP_AS_REFERENCE && mode != LPL_JCP_CLEAN))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    ctx->jcp_mode = mode;
    ctx->c.seg.jcp_emulate_stale = (mode == LPL_JCP_AS_REFERENCE) ? 1 : 0;
    return LPL_OK;
}

// ------------------------------------------------------------------------------------------
// batched pipeline
// ------------------------------------------------------------------------------------------
static int upload_impl(lpl_ctx* ctx, const lpl_frame* frames, std::uint32_t nf, bool from_device)
{
    if (ctx != nullptr)
    {
        ctx->knn_built = false;
    }
    if (ctx == nullptr || frames == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad frame batch (null, empty or larger than max_frames)");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    if (ensure_stage(ctx, sizeof(std::uint32_t) * d.B) != 0)
    {
        return LPL_ERR_CUDA;
    }
    // the counts staging may still be in flight from the previous batch
    LPL_TRY(cudaStreamSynchronize(c.stream));
    auto* h_n = static_cast<std::uint32_t*>(c.h_stage);
    bool any_ring = false;
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        if (frames[f].n > d.cap)
        {
            return fail(ctx, LPL_ERR_CAPACITY, "frame has more points than the context was created for");
        }
        if (frames[f].n != 0 && frames[f].xyzw == nullptr)
        {
            return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "frame without points pointer");
        }
        h_n[f] = frames[f].n;
        any_ring = any_ring || frames[f].ring != nullptr;
    }
    const cudaMemcpyKind kind = from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    LPL_TRY(cudaMemcpyAsync(d.n_in, h_n, sizeof(std::uint32_t) * nf, cudaMemcpyHostToDevice, c.stream));
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        if (frames[f].n == 0)
        {
            continue;
        }
        LPL_TRY(cudaMemcpyAsync(d.pts_in + static_cast<std::size_t>(f) * d.cap, frames[f].xyzw,
                                sizeof(float4) * frames[f].n, kind, c.stream));
        if (frames[f].ring != nullptr)
        {
            LPL_TRY(cudaMemcpyAsync(d.ring + static_cast<std::size_t>(f) * d.cap, frames[f].ring,
                                    sizeof(std::uint16_t) * frames[f].n, kind, c.stream));
        }
        else if (any_ring)
        {
            LPL_TRY(cudaMemsetAsync(d.ring + static_cast<std::size_t>(f) * d.cap, 0,
                                    sizeof(std::uint16_t) * frames[f].n, c.stream));
        }
    }
    ctx->have_ring = any_ring ? 1 : 0;
    return LPL_OK;
}

int lpl_pipeline_upload(lpl_ctx* ctx, const lpl_frame* frames, std::uint32_t nf)
{
    return upload_impl(ctx, frames, nf, false);
}

int lpl_pipeline_upload_device(lpl_ctx* ctx, const lpl_frame* frames, std::uint32_t nf)
{
    return upload_impl(ctx, frames, nf, true);
}

static int upload_packed_impl(lpl_ctx* ctx, const float* pts, const std::uint32_t* counts, std::uint32_t nf,
                              std::uint32_t bytes_per_point)
{
    if (ctx != nullptr)
    {
        ctx->knn_built = false;
    }
    if (ctx == nullptr || counts == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad frame batch (null, empty or larger than max_frames)");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    if (ensure_stage(ctx, sizeof(std::uint32_t) * 2 * d.B) != 0)
    {
        return LPL_ERR_CUDA;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream)); // the staging may still be in flight from the previous batch
    auto* h_n = static_cast<std::uint32_t*>(c.h_stage);
    std::uint32_t* h_start = h_n + d.B;
    unsigned long long total = 0;
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        if (counts[f] > d.cap)
        {
            return fail(ctx, LPL_ERR_CAPACITY, "frame has more points than the context was created for");
        }
        h_n[f] = counts[f];
        h_start[f] = static_cast<std::uint32_t>(total);
        total += counts[f];
    }
    // the raw-record area (32 bytes per point of capacity) doubles as the packed staging: >= 2 x the need
    if (total != 0 && pts == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "null points");
    }
    std::uint32_t* d_start = reinterpret_cast<std::uint32_t*>(d.raw_desc);
    static_assert(sizeof(std::uint32_t) <= 32, "one start offset fits a descriptor slot");
    LPL_TRY(cudaMemcpyAsync(d.n_in, h_n, sizeof(std::uint32_t) * nf, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(d_start, h_start, sizeof(std::uint32_t) * nf, cudaMemcpyHostToDevice, c.stream));
    if (total != 0)
    {
        LPL_TRY(cudaMemcpyAsync(d.raw, pts, static_cast<std::size_t>(total) * bytes_per_point, cudaMemcpyHostToDevice,
                                c.stream));
        launch_spread_packed(&c, nf, d.raw, d_start, bytes_per_point == 12u);
    }
    ctx->have_ring = 0;
    return LPL_OK;
}

int lpl_pipeline_upload_packed(lpl_ctx* ctx, const float* xyzw, const std::uint32_t* counts, std::uint32_t nf)
{
    return upload_packed_impl(ctx, xyzw, counts, nf, 16u);
}

int lpl_pipeline_upload_packed_xyz(lpl_ctx* ctx, const float* xyz, const std::uint32_t* counts, std::uint32_t nf)
{
    return upload_packed_impl(ctx, xyz, counts, nf, 12u);
}

int lpl_pipeline_upload_cloud2(lpl_ctx* ctx, const lpl_cloud2_frame* frames, std::uint32_t nf)
{
    if (ctx != nullptr)
    {
        ctx->knn_built = false;
    }
    if (ctx == nullptr || frames == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad frame batch (null, empty or larger than max_frames)");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    struct Desc
    {
        std::uint32_t width, height, point_step, row_step;
        std::int32_t x_off, y_off, z_off, ring_off;
    };
    static_assert(sizeof(Desc) == 32, "record layout descriptor is 32 bytes");
    if (ensure_stage(ctx, (sizeof(std::uint32_t) + sizeof(Desc)) * d.B) != 0)
    {
        return LPL_ERR_CUDA;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream)); // the staging may still be in flight from the previous batch
    auto* h_n = static_cast<std::uint32_t*>(c.h_stage);
    auto* h_d = reinterpret_cast<Desc*>(h_n + d.B);
    const std::size_t raw_stride = static_cast<std::size_t>(d.cap) * kRawRecord;
    bool any_ring = false;
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        const lpl_cloud2_frame& fr = frames[f];
        const unsigned long long n = static_cast<unsigned long long>(fr.width) * fr.height;
        if (n > d.cap)
        {
            return fail(ctx, LPL_ERR_CAPACITY, "frame has more points than the context was created for");
        }
        const std::size_t bytes = static_cast<std::size_t>(fr.height) * fr.row_step;
        const std::int32_t hi = std::max(std::max(fr.x_offset, fr.y_offset), fr.z_offset) + 4;
        if (n != 0 && (fr.data == nullptr || fr.point_step == 0 || fr.row_step < static_cast<std::uint64_t>(fr.width) * fr.point_step ||
                       fr.x_offset < 0 || fr.y_offset < 0 || fr.z_offset < 0 || static_cast<std::uint32_t>(hi) > fr.point_step ||
                       (fr.ring_offset >= 0 && static_cast<std::uint32_t>(fr.ring_offset) + 2u > fr.point_step)))
        {
            return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "inconsistent PointCloud2 layout (steps / field offsets)");
        }
        if (bytes > raw_stride)
        {
            return fail(ctx, LPL_ERR_CAPACITY, "raw records larger than 32 bytes per point of capacity");
        }
        h_n[f] = static_cast<std::uint32_t>(n);
        h_d[f] = Desc{fr.width, fr.height, fr.point_step, fr.row_step, fr.x_offset, fr.y_offset, fr.z_offset,
                      fr.ring_offset >= 0 ? fr.ring_offset : -1};
        any_ring = any_ring || fr.ring_offset >= 0;
    }
    LPL_TRY(cudaMemcpyAsync(d.n_in, h_n, sizeof(std::uint32_t) * nf, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(d.raw_desc, h_d, sizeof(Desc) * nf, cudaMemcpyHostToDevice, c.stream));
    if (any_ring)
    {
        // frames without a ring field read as ring 0, as in the xyzw upload
        LPL_TRY(cudaMemsetAsync(d.ring, 0, sizeof(std::uint16_t) * static_cast<std::size_t>(d.cap) * nf, c.stream));
    }
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        const std::size_t bytes = static_cast<std::size_t>(frames[f].height) * frames[f].row_step;
        if (bytes != 0)
        {
            LPL_TRY(cudaMemcpyAsync(d.raw + f * raw_stride, frames[f].data, bytes, cudaMemcpyHostToDevice, c.stream));
        }
    }
    launch_unpack_cloud2(&c, nf, d.raw, raw_stride, d.raw_desc);
    ctx->have_ring = any_ring ? 1 : 0;
    return LPL_OK;
}

// streams and events of the sub-batches (created outside any stream capture: lpl_pipeline_run calls this first)
static bool ensure_split_resources(lpl_ctx* ctx, std::uint32_t parts)
{
    while (ctx->aux.size() + 1 < parts)
    {
        lpl_ctx::Aux a{};
        if (cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming) != cudaSuccess)
        {
            cudaGetLastError();
            if (a.stream != nullptr)
            {
                cudaStreamDestroy(a.stream);
            }
            return false;
        }
        try
        {
            ctx->aux.push_back(a);
        }
        catch (...)
        {
            cudaEventDestroy(a.done);
            cudaStreamDestroy(a.stream);
            return false;
        }
    }
    if (ctx->fork_ev == nullptr && cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return true;
}

// one pass of the selected stages over the frames [c.d.f0, c.d.f0 + nf) on c.stream (c: the context itself or the
// view of one sub-batch, see enqueue_stages)
static int enqueue_chain(lpl_ctx* ctx, Ctx& c, std::uint32_t nf, std::uint32_t stages)
{
    Dev& d = c.d;
    const std::uint32_t f0 = d.f0;
    if ((stages & LPL_STAGE_DROR) == 0)
    {
        LPL_TRY(cudaMemsetAsync(at_frame(d.noise, d.cap, f0), 0, static_cast<std::size_t>(d.cap) * nf, c.stream));
    }
    const bool ring_stage = (stages & LPL_STAGE_RING) != 0;
    const bool fused_front = ring_stage && (stages & LPL_STAGE_DROR) != 0; // one read of the cloud for both stages
    if (ring_stage && !fused_front)
    {
        launch_ring(&c, nf);
    }
    else if (!ctx->have_ring)
    {
        LPL_TRY(cudaMemsetAsync(at_frame(d.ring, d.cap, f0), 0, sizeof(std::uint16_t) * static_cast<std::size_t>(d.cap) * nf, c.stream));
    }
    if (stages & LPL_STAGE_DROR)
    {
        launch_dror(&c, nf, fused_front);
    }
    if ((stages & LPL_STAGE_SEGMENT) == 0 && (stages & (LPL_STAGE_CLUSTER | LPL_STAGE_HULLS)) != 0)
    {
        // no labels in this run: the obstacle cloud is empty rather than a previous batch's
        LPL_TRY(cudaMemsetAsync(at_frame(d.labels_out, d.cap, f0), 0, static_cast<std::size_t>(d.cap) * nf, c.stream));
    }
    if (stages & LPL_STAGE_SEGMENT)
    {
        launch_segment(&c, nf, ctx->want_image != 0);
        if (c.launch_failed)
        {
            c.launch_failed = false;
            return LPL_ERR_CUDA; // message already in err
        }
    }
    if (stages & LPL_STAGE_CLUSTER)
    {
        launch_take_obstacles(&c, nf);
        launch_cluster(&c, nf);
    }
    if (stages & LPL_STAGE_HULLS)
    {
        launch_hulls(&c, nf);
    }
    if ((stages & LPL_STAGE_BOXES) && (stages & LPL_STAGE_HULLS))
    {
        launch_boxes(&c, nf, LPL_BOX_ROTATING_CALIPERS);
    }
    return LPL_OK;
}

// Every kernel and memset of one pass of the selected stages (also what a graph captures).
//
// Sub-batches. Frames are independent, and the chain alternates between kernels that fill the GPU (the per-point
// passes) and kernels that cannot (one CTA per frame: the JCP row sweep, the shared-memory union-find, the RECM
// recurrence, the scans). A batch of >= kSplitMinFrames frames is therefore cut into `split_parts` contiguous
// sub-batches whose chains run on concurrent streams (fork / join by events; inside a captured graph these become
// parallel branches): while one sub-batch sits in a latency-bound kernel the per-point kernels of another use the
// SMs it leaves idle. Every kernel takes its frame as blockIdx + Dev::f0, so a sub-batch is the same launch with a
// smaller grid and a frame offset; all per-frame state is disjoint. Not while the per-kernel profile is on (its
// events time one stream).
constexpr std::uint32_t kSplitMinFrames = 4;

static int enqueue_stages(lpl_ctx* ctx, std::uint32_t nf, std::uint32_t stages)
{
    Ctx& c = ctx->c;
    Dev& d = c.d;
    if (c.prof_on)
    {
        c.prof_n = 0;
        LPL_TRY(cudaEventRecord(c.prof_ev[0], c.stream));
    }
    // status words and every per-frame counter in one memset (counts of stages that are not selected read as zero)
    LPL_TRY(cudaMemsetAsync(d.ctr_begin, 0, d.ctr_bytes, c.stream));
    struct Cleared
    {
        Ctx& c;
        explicit Cleared(Ctx& cc) : c(cc) { c.counters_cleared = true; }
        ~Cleared() { c.counters_cleared = false; }
    } cleared(c);
    const bool ring_stage = (stages & LPL_STAGE_RING) != 0;
    c.seg.use_ring = ((ring_stage || ctx->have_ring) && ctx->seg_cfg.assume_unorganized_cloud == 0) ? 1 : 0;
    ctx->ran_hulls = (stages & LPL_STAGE_HULLS) != 0;
    std::uint32_t parts = ctx->split_parts;
    if (c.prof_on || !c.hash_clean || nf < kSplitMinFrames || parts > nf)
    {
        parts = 1; // (the first run of a context clears whole hash planes: one stream)
    }
    if (parts > 1 && !ensure_split_resources(ctx, parts))
    {
        parts = 1;
    }
    if (parts <= 1)
    {
        d.f0 = 0;
        const int rc = enqueue_chain(ctx, c, nf, stages);
        if (rc != 0)
        {
            return rc;
        }
        LPL_TRY(cudaGetLastError());
        return LPL_OK;
    }
    LPL_TRY(cudaEventRecord(ctx->fork_ev, c.stream));
    int rc = LPL_OK;
    for (std::uint32_t p = 0; p < parts; ++p)
    {
        const std::uint32_t a = static_cast<std::uint32_t>(static_cast<unsigned long long>(nf) * p / parts);
        const std::uint32_t b = static_cast<std::uint32_t>(static_cast<unsigned long long>(nf) * (p + 1) / parts);
        Ctx sub = c; // a view: same buffers and parameters, its own stream and frame offset
        sub.d.f0 = a;
        sub.launches = 0;
        if (p != 0)
        {
            sub.stream = ctx->aux[p - 1].stream;
            if (cudaStreamWaitEvent(sub.stream, ctx->fork_ev, 0) != cudaSuccess)
            {
                rc = LPL_ERR_CUDA;
                break;
            }
        }
        const int r = rc == LPL_OK ? enqueue_chain(ctx, sub, b - a, stages) : LPL_OK;
        c.launches += sub.launches;
        if (sub.have_code_map && !c.have_code_map)
        {
            c.code_map = sub.code_map;
            c.have_code_map = true;
        }
        if (r != LPL_OK)
        {
            std::memcpy(c.err, sub.err, sizeof(c.err));
            rc = r;
        }
        if (p != 0)
        {
            // joined even after an error: a capture in progress must end with every forked stream back in
            if (cudaEventRecord(ctx->aux[p - 1].done, sub.stream) != cudaSuccess ||
                cudaStreamWaitEvent(c.stream, ctx->aux[p - 1].done, 0) != cudaSuccess)
            {
                rc = rc != LPL_OK ? rc : LPL_ERR_CUDA;
            }
        }
    }
    if (rc != LPL_OK)
    {
        cudaGetLastError();
        return rc == LPL_ERR_CUDA && c.err[0] == 0 ? fail(ctx, LPL_ERR_CUDA, "sub-batch fork / join failed") : rc;
    }
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}


// The chain is a fixed sequence of ~55 kernels and ~15 memsets whose arguments only depend on (frame count,
// stages, configuration): from the second run with the same key on, the sequence is replayed as one CUDA graph -
// one launch instead of ~70, no host work between the kernels (what the single-frame latency is made of).
// Not while the per-kernel profile is on (its events sit between the kernels), nor with LPL_NO_GRAPH set.
int lpl_pipeline_run(lpl_ctx* ctx, std::uint32_t nf, std::uint32_t stages)
{
    if (ctx == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad frame count");
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaSetDevice(c.device));
    if (ctx->split_parts > 1 && nf >= kSplitMinFrames)
    {
        ensure_split_resources(ctx, ctx->split_parts); // not inside a capture; enqueue_stages falls back to one stream without them
    }
    const bool ring_stage = (stages & LPL_STAGE_RING) != 0;
    const unsigned long long key = (static_cast<unsigned long long>(nf) << 32) | (static_cast<unsigned long long>(stages & 0xffffu) << 8) |
                                   (ctx->want_image ? 1u : 0u) | (ctx->have_ring ? 2u : 0u) | (ring_stage ? 4u : 0u);
    if (c.prof_on || !ctx->use_graphs || !c.hash_clean)
    {
        // (the first run of a context also clears the voxel hash planes: not something to replay)
        return enqueue_stages(ctx, nf, stages);
    }
    for (auto& g : ctx->graphs)
    {
        if (g.key == key && g.epoch == ctx->cfg_epoch)
        {
            ctx->ran_hulls = (stages & LPL_STAGE_HULLS) != 0;
            c.seg.use_ring = g.use_ring;
            LPL_TRY(cudaGraphLaunch(g.exec, c.stream));
            c.launches += g.launches;
            return LPL_OK;
        }
    }
    if (!ctx->graphs.empty() && ctx->graphs.front().epoch != ctx->cfg_epoch)
    {
        drop_graphs(ctx); // configuration changed: kernel arguments baked into the graphs are stale
    }
    if (ctx->graphs.size() >= 8)
    {
        if (ctx->graphs.front().exec != nullptr)
        {
            cudaGraphExecDestroy(ctx->graphs.front().exec);
        }
        ctx->graphs.erase(ctx->graphs.begin());
    }
    const unsigned long long before = c.launches;
    if (cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    {
        cudaGetLastError();
        return enqueue_stages(ctx, nf, stages);
    }
    const int rc = enqueue_stages(ctx, nf, stages);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr)
    {
        cudaGetLastError();
        if (graph != nullptr)
        {
            cudaGraphDestroy(graph);
        }
        c.launches = before;
        return rc != 0 ? rc : enqueue_stages(ctx, nf, stages);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || exec == nullptr)
    {
        cudaGetLastError();
        c.launches = before;
        return enqueue_stages(ctx, nf, stages);
    }
    try
    {
        ctx->graphs.push_back({key, ctx->cfg_epoch, exec, c.launches - before, c.seg.use_ring});
    }
    catch (...)
    {
        cudaGraphExecDestroy(exec);
        c.launches = before;
        return enqueue_stages(ctx, nf, stages);
    }
    LPL_TRY(cudaGraphLaunch(exec, c.stream));
    return LPL_OK;
}

int lpl_pipeline_use_graph(lpl_ctx* ctx, int enable)
{
    if (ctx == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    ctx->use_graphs = enable != 0 && std::getenv("LPL_NO_GRAPH") == nullptr;
    return LPL_OK;
}

int lpl_pipeline_use_split(lpl_ctx* ctx, std::uint32_t parts)
{
    if (ctx == nullptr || parts == 0 || parts > 8)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    if (parts != ctx->split_parts)
    {
        ctx->split_parts = parts;
        ctx->cfg_epoch += 1; // captured graphs hold the old branch structure
    }
    return LPL_OK;
}

int lpl_pipeline_sync(lpl_ctx* ctx, std::uint32_t nf)
{
    if (ctx == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    return check_status(ctx, nf);
}

int lpl_pipeline_status(lpl_ctx* ctx, std::uint32_t nf, std::uint32_t* status_out)
{
    if (ctx == nullptr || status_out == nullptr || nf == 0 || nf > ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    const int rc = check_status(ctx, nf);
    if (rc != 0 && rc != LPL_ERR_CAPACITY)
    {
        return rc;
    }
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        status_out[f] = ctx->h_status[f];
    }
    return LPL_OK;
}

int lpl_pipeline_want_image(lpl_ctx* ctx, int enable)
{
    if (ctx == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    ctx->want_image = enable ? 1 : 0;
    return LPL_OK;
}

int lpl_pipeline_counts(lpl_ctx* ctx, std::uint32_t f, lpl_frame_result* r)
{
    if (ctx == nullptr || r == nullptr || f >= ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    std::uint32_t v[5] = {0, 0, 0, 0, 0};
    LPL_TRY(cudaMemcpyAsync(&v[0], d.n_in + f, 4, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaMemcpyAsync(&v[1], d.n_v + f, 4, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaMemcpyAsync(&v[2], d.n_o + f, 4, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaMemcpyAsync(&v[3], d.n_clusters + f, 4, cudaMemcpyDeviceToHost, c.stream));
    // n_hull is zeroed by every run and only written by the hull stage: a run without LPL_STAGE_HULLS reports
    // no vertices instead of the previous batch's offsets
    LPL_TRY(cudaMemcpyAsync(&v[4], d.n_hull + f, 4, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    r->n = v[0];
    r->num_valid = v[1];
    r->num_obstacles = v[2];
    r->num_clusters = v[3];
    r->num_hull_vertices = v[4];
    return LPL_OK;
}

int lpl_pipeline_download(lpl_ctx* ctx, std::uint32_t f, lpl_frame_result* r)
{
    const int rc = lpl_pipeline_counts(ctx, f, r);
    if (rc != 0)
    {
        return rc;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
    if (r->noise != nullptr && r->n != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->noise, d.noise + o, r->n, k, c.stream));
    }
    if (r->ring != nullptr && r->n != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->ring, d.ring + o, sizeof(std::uint16_t) * r->n, k, c.stream));
    }
    if (r->labels != nullptr && r->n != 0)
    {
        // device plane is u8; the reference's Label is uint32_t (segmenter.hpp:69-74)
        auto* raw = reinterpret_cast<std::uint8_t*>(r->labels) + static_cast<std::size_t>(r->n) * 3;
        LPL_TRY(cudaMemcpyAsync(raw, d.labels_out + o, r->n, k, c.stream));
        LPL_TRY(cudaStreamSynchronize(c.stream));
        for (std::uint32_t i = 0; i < r->n; ++i)
        {
            r->labels[i] = raw[i]; // forward in-place widening: byte i sits at offset 3n + i >= 4i
        }
    }
    if (r->obstacle_index != nullptr && r->num_obstacles != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->obstacle_index, d.idx_o + o, sizeof(std::uint32_t) * r->num_obstacles, k, c.stream));
    }
    if (r->cluster_labels != nullptr && r->num_obstacles != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->cluster_labels, d.clabel + o, sizeof(std::int32_t) * r->num_obstacles, k, c.stream));
    }
    if (r->hull_offsets != nullptr)
    {
        if (ctx->ran_hulls)
        {
            LPL_TRY(cudaMemcpyAsync(r->hull_offsets, d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1),
                                    sizeof(std::uint32_t) * (r->num_clusters + 1), k, c.stream));
        }
        else
        {
            std::memset(r->hull_offsets, 0, sizeof(std::uint32_t) * (r->num_clusters + 1)); // no hull stage in this run
        }
    }
    if (r->hull_indices != nullptr && r->num_hull_vertices != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->hull_indices, d.hull_idx + o, sizeof(std::uint32_t) * r->num_hull_vertices, k, c.stream));
    }
    if (r->hull_xy != nullptr && r->num_hull_vertices != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->hull_xy, d.hull_xy + o, sizeof(float2) * r->num_hull_vertices, k, c.stream));
    }
    if (r->zminmax != nullptr && r->num_clusters != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->zminmax, d.zminmax + o, sizeof(float2) * r->num_clusters, k, c.stream));
    }
    if (r->boxes != nullptr && r->num_clusters != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->boxes, d.boxes + o, sizeof(ObbBox) * r->num_clusters, k, c.stream));
    }
    if (r->bgr != nullptr)
    {
        LPL_TRY(cudaMemcpyAsync(r->bgr, d.bgr + static_cast<std::size_t>(f) * c.seg.npx * 3,
                                static_cast<std::size_t>(c.seg.npx) * 3, k, c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    return LPL_OK;
}

// ------------------------------------------------------------------------------------------
// single-frame entry points (the reference's per-object calls)
// ------------------------------------------------------------------------------------------
static int stage_points(lpl_ctx* ctx, const void* points, std::size_t stride, std::uint32_t n, float4* dst,
                        std::uint32_t* n_dst, std::int32_t ring_offset)
{
    Ctx& c = ctx->c;
    Dev& d = c.d;
    if (n > d.cap)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "more points than the context was created for");
    }
    if (n != 0 && (points == nullptr || stride < 12))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "null points or stride < 12 bytes");
    }
    LPL_TRY(cudaSetDevice(c.device));
    if (dst == d.pts_in)
    {
        ctx->knn_built = false; // the searched point set of lpl_knn_* lives there
    }
    const std::size_t need = 64 + static_cast<std::size_t>(n) * 16 + static_cast<std::size_t>(n) * 2;
    if (ensure_stage(ctx, need) != 0)
    {
        return LPL_ERR_CUDA;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    auto* base = static_cast<char*>(c.h_stage);
    *reinterpret_cast<std::uint32_t*>(base) = n;
    LPL_TRY(cudaMemcpyAsync(n_dst, base, 4, cudaMemcpyHostToDevice, c.stream));
    if (n == 0)
    {
        return LPL_OK;
    }
    char* hp = base + 64;
    if (stride < 16)
    {
        std::memset(hp, 0, static_cast<std::size_t>(n) * 16);
    }
    pack(hp, 16, points, stride, stride < 16 ? 12 : 16, n);
    LPL_TRY(cudaMemcpyAsync(dst, hp, static_cast<std::size_t>(n) * 16, cudaMemcpyHostToDevice, c.stream));
    if (ring_offset >= 0)
    {
        char* hr = hp + static_cast<std::size_t>(n) * 16;
        pack(hr, 2, static_cast<const char*>(points) + ring_offset, stride, 2, n);
        LPL_TRY(cudaMemcpyAsync(d.ring, hr, static_cast<std::size_t>(n) * 2, cudaMemcpyHostToDevice, c.stream));
    }
    return LPL_OK;
}

int lpl_ring_partition(lpl_ctx* ctx, const void* points, std::size_t stride, std::uint32_t n, std::uint16_t* ring_out)
{
    if (ctx == nullptr || (n != 0 && ring_out == nullptr))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    int rc = stage_points(ctx, points, stride, n, c.d.pts_in, c.d.n_in, -1);
    if (rc != 0 || n == 0)
    {
        return rc;
    }
    launch_ring(&c, 1);
    LPL_TRY(cudaMemcpyAsync(ring_out, c.d.ring, sizeof(std::uint16_t) * n, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

int lpl_dror_filter(lpl_ctx* ctx, const void* points, std::size_t stride, std::uint32_t n, std::uint8_t* labels_out)
{
    if (ctx == nullptr || (n != 0 && labels_out == nullptr))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    int rc = stage_points(ctx, points, stride, n, c.d.pts_in, c.d.n_in, -1);
    if (rc != 0 || n == 0)
    {
        return rc; // empty cloud: valid no-op (kdtree.hpp:167-170)
    }
    launch_dror(&c, 1, false);
    LPL_TRY(cudaMemcpyAsync(labels_out, c.d.noise, n, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

int lpl_segment(lpl_ctx* ctx, const void* points, std::size_t stride, std::int32_t ring_offset, std::uint32_t n,
                std::uint32_t* labels_out, std::uint8_t* bgr_image_out)
{
    if (ctx == nullptr || (n != 0 && labels_out == nullptr))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    if (ring_offset >= 0 && static_cast<std::size_t>(ring_offset) + 2 > stride)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "ring_offset outside the point record");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    int rc = stage_points(ctx, points, stride, n, d.pts_in, d.n_in, ring_offset);
    if (rc != 0)
    {
        return rc;
    }
    LPL_TRY(cudaMemsetAsync(d.status, 0, sizeof(std::uint32_t), c.stream));
    if (ring_offset < 0)
    {
        LPL_TRY(cudaMemsetAsync(d.ring, 0, sizeof(std::uint16_t) * d.cap, c.stream));
    }
    c.seg.use_ring = (ring_offset >= 0 && ctx->seg_cfg.assume_unorganized_cloud == 0) ? 1 : 0;
    LPL_TRY(cudaMemsetAsync(d.noise, 0, d.cap, c.stream)); // every point enters the segmenter
    launch_segment(&c, 1, bgr_image_out != nullptr);
    if (c.launch_failed)
    {
        c.launch_failed = false;
        return LPL_ERR_CUDA;
    }
    std::uint8_t* raw = reinterpret_cast<std::uint8_t*>(labels_out) + static_cast<std::size_t>(n) * 3;
    if (n != 0)
    {
        LPL_TRY(cudaMemcpyAsync(raw, d.labels_out, n, cudaMemcpyDeviceToHost, c.stream));
    }
    if (bgr_image_out != nullptr)
    {
        LPL_TRY(cudaMemcpyAsync(bgr_image_out, d.bgr, static_cast<std::size_t>(c.seg.npx) * 3, cudaMemcpyDeviceToHost,
                                c.stream));
    }
    rc = check_status(ctx, 1);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        labels_out[i] = raw[i]; // u8 device plane -> uint32_t Label, widened in place front to back
    }
    return rc;
}

int lpl_cluster(lpl_ctx* ctx, const void* points, std::size_t stride, std::uint32_t n, std::int32_t* labels_out,
                std::uint32_t* num_clusters_out)
{
    if (ctx == nullptr || (n != 0 && labels_out == nullptr))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    int rc = stage_points(ctx, points, stride, n, d.pts_o, d.n_o, -1);
    if (rc != 0)
    {
        return rc;
    }
    if (num_clusters_out != nullptr)
    {
        *num_clusters_out = 0;
    }
    if (n == 0)
    {
        return LPL_OK; // clusterer.cpp:62-65
    }
    LPL_TRY(cudaMemsetAsync(d.status, 0, sizeof(std::uint32_t), c.stream));
    launch_cluster(&c, 1);
    LPL_TRY(cudaMemcpyAsync(labels_out, d.clabel, sizeof(std::int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
    std::uint32_t k = 0;
    LPL_TRY(cudaMemcpyAsync(&k, d.n_clusters, 4, cudaMemcpyDeviceToHost, c.stream));
    rc = check_status(ctx, 1);
    if (num_clusters_out != nullptr)
    {
        *num_clusters_out = k;
    }
    return rc;
}

int lpl_cluster_hulls(lpl_ctx* ctx, const void* points, std::size_t stride, const std::int32_t* labels, std::uint32_t n,
                      std::uint32_t num_clusters, std::uint32_t* hull_offsets, std::int32_t* hull_indices, float* hull_xy,
                      float* zminmax)
{
    if (ctx == nullptr || hull_offsets == nullptr || (n != 0 && labels == nullptr))
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    if (num_clusters > d.cap)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "more clusters than point capacity");
    }
    int rc = stage_points(ctx, points, stride, n, d.pts_o, d.n_o, -1);
    if (rc != 0)
    {
        return rc;
    }
    if (n == 0 || num_clusters == 0)
    {
        for (std::uint32_t k = 0; k <= num_clusters; ++k)
        {
            hull_offsets[k] = 0;
        }
        return LPL_OK;
    }
    LPL_TRY(cudaMemcpyAsync(d.clabel, labels, sizeof(std::int32_t) * n, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(d.n_clusters, &num_clusters, 4, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemsetAsync(d.ccount, 0, sizeof(std::uint32_t) * num_clusters, c.stream));
    LPL_TRY(cudaMemsetAsync(d.zmin_u, 0xff, sizeof(std::uint32_t) * num_clusters, c.stream));
    LPL_TRY(cudaMemsetAsync(d.zmax_u, 0, sizeof(std::uint32_t) * num_clusters, c.stream));
    LPL_TRY(cudaMemsetAsync(d.zzero, 0xff, sizeof(std::uint32_t) * num_clusters, c.stream));
    k_ext_init<<<(num_clusters + 255) / 256, 256, 0, c.stream>>>(d, num_clusters);
    mark(&c, "ext_init");
    k_label_count<<<dim3((d.cap + 255) / 256, 1), 256, 0, c.stream>>>(d, num_clusters);
    mark(&c, "label_count");
    launch_hulls(&c, 1);
    ctx->ran_hulls = true;
    LPL_TRY(cudaMemcpyAsync(hull_offsets, d.hull_off, sizeof(std::uint32_t) * (num_clusters + 1), cudaMemcpyDeviceToHost,
                            c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    const std::uint32_t tot = hull_offsets[num_clusters];
    if (tot > n)
    {
        return fail(ctx, LPL_ERR_CUDA, "internal error: more hull vertices than points");
    }
    if (hull_indices != nullptr && tot != 0)
    {
        LPL_TRY(cudaMemcpyAsync(hull_indices, d.hull_idx, sizeof(std::uint32_t) * tot, cudaMemcpyDeviceToHost, c.stream));
    }
    if (hull_xy != nullptr && tot != 0)
    {
        LPL_TRY(cudaMemcpyAsync(hull_xy, d.hull_xy, sizeof(float2) * tot, cudaMemcpyDeviceToHost, c.stream));
    }
    if (zminmax != nullptr)
    {
        LPL_TRY(cudaMemcpyAsync(zminmax, d.zminmax, sizeof(float2) * num_clusters, cudaMemcpyDeviceToHost, c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

static_assert(sizeof(lpl_bbox) == sizeof(ObbBox), "lpl_bbox and the device box record share one layout");

int lpl_bounding_boxes(lpl_ctx* ctx, const void* xy, std::size_t stride, const std::uint32_t* offsets,
                       std::uint32_t num_hulls, int method, lpl_bbox* boxes_out)
{
    if (ctx == nullptr || (num_hulls != 0 && (offsets == nullptr || boxes_out == nullptr)) || stride < 16 ||
        (method != LPL_BOX_ROTATING_CALIPERS && method != LPL_BOX_PCA))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad bounding-box request");
    }
    if (num_hulls == 0)
    {
        return LPL_OK;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    const std::size_t total = offsets[num_hulls];
    for (std::uint32_t k = 0; k < num_hulls; ++k)
    {
        if (offsets[k] > offsets[k + 1])
        {
            return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "hull offsets must be non-decreasing");
        }
    }
    const std::size_t room = static_cast<std::size_t>(d.B) * d.cap;
    if (total > room || num_hulls >= room)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "more hull vertices / hulls than the context was created for");
    }
    if (total != 0 && xy == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "null hull points");
    }
    LPL_TRY(cudaSetDevice(c.device));
    const std::size_t need = total * 16 + (static_cast<std::size_t>(num_hulls) + 1) * 4;
    if (ensure_stage(ctx, need) != 0)
    {
        return LPL_ERR_CUDA;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    auto* base = static_cast<char*>(c.h_stage);
    pack(base, 16, xy, stride, 16, static_cast<std::uint32_t>(total));
    std::memcpy(base + total * 16, offsets, (static_cast<std::size_t>(num_hulls) + 1) * 4);
    // device scratch that is idle outside a pipeline run: the hull sort buffer and the segment starts
    auto* dxy = reinterpret_cast<double2*>(d.hsA);
    std::uint32_t* doff = d.cstart;
    if (total != 0)
    {
        LPL_TRY(cudaMemcpyAsync(dxy, base, total * 16, cudaMemcpyHostToDevice, c.stream));
    }
    LPL_TRY(cudaMemcpyAsync(doff, base + total * 16, (static_cast<std::size_t>(num_hulls) + 1) * 4, cudaMemcpyHostToDevice,
                            c.stream));
    launch_boxes_hulls(&c, dxy, doff, num_hulls, method, d.boxes);
    LPL_TRY(cudaMemcpyAsync(boxes_out, d.boxes, sizeof(ObbBox) * num_hulls, cudaMemcpyDeviceToHost, c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

int lpl_vehicle_match(lpl_ctx* ctx, const void* hull_xy, std::size_t stride, const std::uint32_t* offsets, std::uint32_t num_hulls,
                      const double* z_min_max, const std::uint32_t* cluster_sizes, const lpl_bbox* boxes, std::int32_t* class_out,
                      double* polygon_area_out)
{
    if (ctx == nullptr || stride < 16 ||
        (num_hulls != 0 && (offsets == nullptr || z_min_max == nullptr || cluster_sizes == nullptr || boxes == nullptr || class_out == nullptr)))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad vehicle match request");
    }
    if (num_hulls == 0)
    {
        return LPL_OK;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    const std::size_t total = offsets[num_hulls];
    for (std::uint32_t k = 0; k < num_hulls; ++k)
    {
        if (offsets[k] > offsets[k + 1])
        {
            return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "hull offsets must be non-decreasing");
        }
    }
    const std::size_t room = static_cast<std::size_t>(d.B) * d.cap;
    if (total > room || static_cast<std::size_t>(num_hulls) * 3 >= room)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "more hull vertices / hulls than the context was created for");
    }
    if (total != 0 && hull_xy == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "null hull points");
    }
    LPL_TRY(cudaSetDevice(c.device));
    const std::size_t K = num_hulls;
    const std::size_t need = total * 16 + (K + 1) * 4;
    if (ensure_stage(ctx, need) != 0)
    {
        return LPL_ERR_CUDA;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    auto* hbase = static_cast<char*>(c.h_stage);
    pack(hbase, 16, hull_xy, stride, 16, static_cast<std::uint32_t>(total));
    std::memcpy(hbase + total * 16, offsets, (K + 1) * 4);
    // device scratch that is idle outside a pipeline run: the hull sort buffers, the segment starts, the box plane
    auto* dxy = reinterpret_cast<double2*>(d.hsA);
    std::uint32_t* doff = d.cstart;
    auto* dz = reinterpret_cast<double2*>(d.hsB);
    auto* darea = reinterpret_cast<double*>(dz + K);
    auto* dsize = reinterpret_cast<std::uint32_t*>(darea + K);
    auto* dcls = reinterpret_cast<std::int32_t*>(dsize + K);
    if (total != 0)
    {
        LPL_TRY(cudaMemcpyAsync(dxy, hbase, total * 16, cudaMemcpyHostToDevice, c.stream));
    }
    LPL_TRY(cudaMemcpyAsync(doff, hbase + total * 16, (K + 1) * 4, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(dz, z_min_max, K * 16, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(dsize, cluster_sizes, K * 4, cudaMemcpyHostToDevice, c.stream));
    LPL_TRY(cudaMemcpyAsync(d.boxes, boxes, K * sizeof(ObbBox), cudaMemcpyHostToDevice, c.stream));
    launch_vehicle_match(&c, dxy, doff, num_hulls, dz, dsize, d.boxes, dcls, darea);
    LPL_TRY(cudaMemcpyAsync(class_out, dcls, K * 4, cudaMemcpyDeviceToHost, c.stream));
    if (polygon_area_out != nullptr)
    {
        LPL_TRY(cudaMemcpyAsync(polygon_area_out, darea, K * 8, cudaMemcpyDeviceToHost, c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

int lpl_convex_hull(lpl_ctx* ctx, const void* xy, std::size_t stride, std::uint32_t n, std::int32_t* indices_out,
                    std::uint32_t* count_out)
{
    if (ctx == nullptr || count_out == nullptr || (n != 0 && (xy == nullptr || indices_out == nullptr)) || stride < 16)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    *count_out = 0;
    if (n == 0)
    {
        return LPL_OK;
    }
    // The fast device path sorts on float keys: every PCL-derived PointXY is float-representable
    // (processor.cpp:645-646). Anything else takes the fp64 path below - never a silent rounding.
    std::vector<float> pts;
    std::vector<std::int32_t> lab;
    try
    {
        pts.assign(static_cast<std::size_t>(n) * 4, 0.f);
        lab.assign(n, 0);
    }
    catch (...)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "convexHull: host staging allocation failed"); // nothing unwinds through the C ABI
    }
    bool float_exact = true;
    for (std::uint32_t i = 0; i < n && float_exact; ++i)
    {
        double v[2];
        std::memcpy(v, static_cast<const char*>(xy) + i * stride, 16);
        const float fx = static_cast<float>(v[0]), fy = static_cast<float>(v[1]);
        float_exact = static_cast<double>(fx) == v[0] && static_cast<double>(fy) == v[1];
        pts[4 * i] = fx;
        pts[4 * i + 1] = fy;
    }
    if (!float_exact)
    {
        Ctx& c = ctx->c;
        Dev& d = c.d;
        if (n > d.cap)
        {
            return fail(ctx, LPL_ERR_CAPACITY, "more points than the context was created for");
        }
        LPL_TRY(cudaSetDevice(c.device));
        if (ensure_stage(ctx, static_cast<std::size_t>(n) * 16) != 0)
        {
            return LPL_ERR_CUDA;
        }
        LPL_TRY(cudaStreamSynchronize(c.stream));
        pack(c.h_stage, 16, xy, stride, 16, n);
        auto* dxy = reinterpret_cast<double2*>(d.hsA); // idle outside a pipeline run
        LPL_TRY(cudaMemcpyAsync(dxy, c.h_stage, static_cast<std::size_t>(n) * 16, cudaMemcpyHostToDevice, c.stream));
        launch_hull_f64(&c, dxy, n, d.hstack, d.hull_off, d.hull_idx, d.n_hull); // order: n entries, sweep stack: n + 1
        LPL_TRY(cudaMemcpyAsync(count_out, d.n_hull, 4, cudaMemcpyDeviceToHost, c.stream));
        LPL_TRY(cudaStreamSynchronize(c.stream));
        if (*count_out > n)
        {
            *count_out = 0;
            return fail(ctx, LPL_ERR_CUDA, "internal error: more hull vertices than points");
        }
        if (*count_out != 0)
        {
            LPL_TRY(cudaMemcpyAsync(indices_out, d.hull_idx, sizeof(std::uint32_t) * *count_out, cudaMemcpyDeviceToHost, c.stream));
            LPL_TRY(cudaStreamSynchronize(c.stream));
        }
        LPL_TRY(cudaGetLastError());
        return LPL_OK;
    }
    std::uint32_t off[2] = {0, 0};
    const int rc = lpl_cluster_hulls(ctx, pts.data(), 16, lab.data(), n, 1, off, indices_out, nullptr, nullptr);
    if (rc == 0)
    {
        *count_out = off[1];
    }
    return rc;
}

int lpl_pipeline_download_batch(lpl_ctx* ctx, std::uint32_t nf, lpl_batch_result* r)
{
    if (ctx == nullptr || r == nullptr || nf == 0 || nf > ctx->c.d.B || r->counts == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad batch download request");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
    // phase 1: per-frame counts (sizes the plane copies) + status
    std::uint32_t* cn = r->counts;
    LPL_TRY(cudaMemcpyAsync(cn + 0 * nf, d.n_in, 4 * nf, k, c.stream));
    LPL_TRY(cudaMemcpyAsync(cn + 1 * nf, d.n_v, 4 * nf, k, c.stream));
    LPL_TRY(cudaMemcpyAsync(cn + 2 * nf, d.n_o, 4 * nf, k, c.stream));
    LPL_TRY(cudaMemcpyAsync(cn + 3 * nf, d.n_clusters, 4 * nf, k, c.stream));
    LPL_TRY(cudaMemcpyAsync(cn + 4 * nf, d.n_hull, 4 * nf, k, c.stream));
    // a frame that ran out of a reserved capacity does not hold back the others: every plane is still delivered
    // and the call returns LPL_ERR_CAPACITY afterwards (lpl_pipeline_status tells which frames to discard)
    const int rc_status = check_status(ctx, nf);
    if (rc_status != 0 && rc_status != LPL_ERR_CAPACITY)
    {
        return rc_status;
    }
    std::uint32_t mx[5] = {0, 0, 0, 0, 0};
    for (int a = 0; a < 5; ++a)
    {
        for (std::uint32_t f = 0; f < nf; ++f)
        {
            mx[a] = std::max(mx[a], cn[a * nf + f]);
        }
    }
    const std::size_t hs = r->stride; // host plane stride in elements
    if (hs < mx[0] || hs < static_cast<std::size_t>(mx[3]) + 1)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "host result stride smaller than the largest frame");
    }
    auto plane = [&](void* dst, const void* src, std::size_t elem, std::size_t dev_stride, std::uint32_t width) -> cudaError_t {
        if (dst == nullptr || width == 0)
        {
            return cudaSuccess;
        }
        return cudaMemcpy2DAsync(dst, hs * elem, src, dev_stride * elem, static_cast<std::size_t>(width) * elem, nf, k,
                                 c.stream);
    };
    LPL_TRY(plane(r->labels_u8, d.labels_out, 1, d.cap, mx[0]));
    LPL_TRY(plane(r->noise, d.noise, 1, d.cap, mx[0]));
    LPL_TRY(plane(r->ring, d.ring, 2, d.cap, mx[0]));
    LPL_TRY(plane(r->obstacle_index, d.idx_o, 4, d.cap, mx[2]));
    LPL_TRY(plane(r->cluster_labels, d.clabel, 4, d.cap, mx[2]));
    LPL_TRY(plane(r->hull_offsets, d.hull_off, 4, d.cap + 1, mx[3] + 1));
    LPL_TRY(plane(r->hull_indices, d.hull_idx, 4, d.cap, mx[4]));
    LPL_TRY(plane(r->hull_xy, d.hull_xy, 8, d.cap, mx[4]));
    LPL_TRY(plane(r->zminmax, d.zminmax, 8, d.cap, mx[3]));
    LPL_TRY(plane(r->boxes, d.boxes, sizeof(ObbBox), d.cap, mx[3]));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    return rc_status;
}

static_assert(LPL_PLANE_COUNT == kPackPlanes, "plane bits of the C ABI = planes the pack kernels know");

int lpl_pipeline_download_packed(lpl_ctx* ctx, std::uint32_t nf, lpl_packed_result* r)
{
    if (ctx == nullptr || r == nullptr || nf == 0 || nf > ctx->c.d.B || r->counts == nullptr ||
        (r->planes >> LPL_PLANE_COUNT) != 0u)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad packed download request");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
    // device staging = the raw-record area (32 bytes per point of capacity), idle once a batch is uploaded
    unsigned char* staging = d.raw;
    const std::size_t staging_bytes = static_cast<std::size_t>(d.B) * d.cap * kRawRecord;
    const std::size_t head = pack_payload_start(nf);
    if (staging_bytes <= head)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "context too small for a packed download");
    }
    launch_pack_results(&c, nf, r->planes, staging, staging_bytes);
    // phase 1: the layout header with the per-frame counts behind it (one small transfer) and the status words
    std::uint32_t* cn = r->counts;
    const std::size_t hc_bytes = sizeof(PackHeader) + static_cast<std::size_t>(5) * nf * sizeof(std::uint32_t);
    if (ensure_stage(ctx, sizeof(std::uint32_t) * 2 * d.B + hc_bytes) != 0)
    {
        return LPL_ERR_CUDA;
    }
    // (the upload staging at the front of h_stage is consumed by then: this copy is behind it in the stream)
    auto* h_hdr = reinterpret_cast<PackHeader*>(static_cast<char*>(c.h_stage) + sizeof(std::uint32_t) * 2 * d.B);
    LPL_TRY(cudaMemcpyAsync(h_hdr, staging, hc_bytes, k, c.stream));
    const int rc_status = check_status(ctx, nf); // flagged frames do not hold back the others (see download_batch)
    if (rc_status != 0 && rc_status != LPL_ERR_CAPACITY)
    {
        return rc_status;
    }
    std::memcpy(cn, reinterpret_cast<const char*>(h_hdr) + sizeof(PackHeader), static_cast<std::size_t>(5) * nf * sizeof(std::uint32_t));
    for (int p = 0; p < LPL_PLANE_COUNT; ++p)
    {
        r->offset[p] = (r->planes & (1u << p)) ? static_cast<std::size_t>(h_hdr->offset[p]) : static_cast<std::size_t>(-1);
    }
    r->bytes_used = static_cast<std::size_t>(h_hdr->total);
    if (h_hdr->fits == 0u)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "packed results exceed the device staging area (32 bytes per point of capacity)");
    }
    if (r->bytes_used > r->buffer_bytes || (r->bytes_used != 0 && r->buffer == nullptr))
    {
        return fail(ctx, LPL_ERR_CAPACITY, "host buffer smaller than the packed results");
    }
    // phase 2: the payload, one transfer
    if (r->bytes_used != 0)
    {
        LPL_TRY(cudaMemcpyAsync(r->buffer, staging + head, r->bytes_used, k, c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    return rc_status;
}

// ------------------------------------------------------------------------------------------
// glibc rand() (TYPE_3 additive feedback generator, r[i] = r[i-3] + r[i-31], top 31 bits): the node colours
// cluster k of every frame with three consecutive std::rand() % 256 draws of the process-wide, never seeded
// stream (processor.cpp:629-631). A context replays that stream from srand(1), the state a fresh process has.
// ------------------------------------------------------------------------------------------
static void glibc_srand(std::int32_t* r, int* pos, std::uint32_t seed)
{
    r[0] = static_cast<std::int32_t>(seed == 0u ? 1u : seed);
    for (int i = 1; i < 31; ++i)
    {
        const long long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long long w = 16807 * lo - 2836 * hi;
        if (w < 0)
        {
            w += 2147483647;
        }
        r[i] = static_cast<std::int32_t>(w);
    }
    for (int i = 31; i < 34; ++i)
    {
        r[i] = r[i - 31];
    }
    *pos = 0;
    // the state is a ring of 34 words holding the last 34 outputs-before-shift; 310 draws are discarded
    for (int k = 0; k < 310; ++k)
    {
        const int i = *pos;
        const std::uint32_t v = static_cast<std::uint32_t>(r[(i + 3) % 34]) + static_cast<std::uint32_t>(r[(i + 31) % 34]);
        r[i % 34] = static_cast<std::int32_t>(v);
        *pos = (i + 1) % 34;
    }
}

static std::int32_t glibc_rand(std::int32_t* r, int* pos)
{
    const int i = *pos;
    // ring index i holds o[t - 34]; o[t] = o[t - 31] + o[t - 3]
    const std::uint32_t v = static_cast<std::uint32_t>(r[(i + 3) % 34]) + static_cast<std::uint32_t>(r[(i + 31) % 34]);
    r[i] = static_cast<std::int32_t>(v);
    *pos = (i + 1) % 34;
    return static_cast<std::int32_t>(v >> 1);
}

void lpl_glibc_rand_stream(std::uint32_t seed, std::uint32_t count, std::int32_t* out)
{
    std::int32_t r[34];
    int pos = 0;
    glibc_srand(r, &pos, seed);
    for (std::uint32_t k = 0; k < count && out != nullptr; ++k)
    {
        out[k] = glibc_rand(r, &pos);
    }
}

int lpl_pipeline_split_clouds(lpl_ctx* ctx, std::uint32_t nf, lpl_split_result* r)
{
    if (ctx == nullptr || r == nullptr || nf == 0 || nf > ctx->c.d.B || r->counts == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad cloud split request");
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    LPL_TRY(cudaSetDevice(c.device));
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
    const std::size_t cap = d.cap, B = d.B;
    // device scratch: 4 cloud planes of B x cap records, marker vertices (cap / 2 per frame), colours (3 B per cluster),
    // counters: tile counts [B][tiles][3], totals [B][4], marker counts / offsets / totals
    const std::size_t rec_bytes = 4 * B * cap * 32;
    const std::size_t mk_verts = cap / 2;
    const std::size_t mk_bytes = B * mk_verts * 24;
    const std::size_t col_bytes = (B * cap * 3 + 255) / 256 * 256;
    const std::size_t cnt_bytes = (B * d.tiles * 3 + B * 4 + B * cap + B * (cap + 1) + B) * 4 + 1024;
    const std::size_t need = rec_bytes + mk_bytes + col_bytes + cnt_bytes;
    if (ctx->split_bytes < need)
    {
        LPL_TRY(cudaStreamSynchronize(c.stream));
        if (ctx->split_dev != nullptr)
        {
            cudaFree(ctx->split_dev);
            ctx->split_dev = nullptr;
            ctx->split_bytes = 0;
        }
        if (cudaMalloc(&ctx->split_dev, need) != cudaSuccess)
        {
            cudaGetLastError();
            return fail(ctx, LPL_ERR_CAPACITY, "device allocation for the cloud split failed");
        }
        ctx->split_bytes = need;
    }
    auto* base = static_cast<unsigned char*>(ctx->split_dev);
    unsigned char* d_rec = base;
    auto* d_mk = reinterpret_cast<double*>(base + rec_bytes);
    auto* d_col = base + rec_bytes + mk_bytes;
    auto* d_cnt3 = reinterpret_cast<std::uint32_t*>(base + rec_bytes + mk_bytes + col_bytes);
    std::uint32_t* d_tot = d_cnt3 + B * d.tiles * 3;
    std::uint32_t* d_mcount = d_tot + B * 4;
    std::uint32_t* d_moff = d_mcount + B * cap;
    std::uint32_t* d_mtot = d_moff + B * (cap + 1);
    // colours: from the caller, or the node's own rand() stream (needs the cluster counts of the batch)
    std::vector<std::uint32_t> K;
    std::vector<std::uint8_t> col;
    try
    {
        K.resize(nf);
        LPL_TRY(cudaMemcpyAsync(K.data(), d.n_clusters, 4 * nf, k, c.stream));
        const int rc_status = check_status(ctx, nf);
        if (rc_status != 0)
        {
            return rc_status;
        }
        col.assign(static_cast<std::size_t>(nf) * cap * 3, 0);
        for (std::uint32_t f = 0; f < nf; ++f)
        {
            if (K[f] > cap)
            {
                return fail(ctx, LPL_ERR_CUDA, "internal error: more clusters than points");
            }
            for (std::uint32_t q = 0; q < K[f]; ++q)
            {
                std::uint8_t* o = col.data() + (static_cast<std::size_t>(f) * cap + q) * 3;
                if (r->cluster_colors != nullptr)
                {
                    const std::uint8_t* src = r->cluster_colors + (static_cast<std::size_t>(f) * r->colors_stride + q) * 3;
                    if (q >= r->colors_stride)
                    {
                        return fail(ctx, LPL_ERR_CAPACITY, "fewer colours than clusters");
                    }
                    o[0] = src[0];
                    o[1] = src[1];
                    o[2] = src[2];
                }
                else
                {
                    if (ctx->rand_i < 0)
                    {
                        glibc_srand(ctx->rand_r, &ctx->rand_i, 1u);
                    }
                    for (int ch = 0; ch < 3; ++ch)
                    {
                        o[ch] = static_cast<std::uint8_t>(glibc_rand(ctx->rand_r, &ctx->rand_i) % 256);
                    }
                }
            }
        }
    }
    catch (...)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "host allocation for the cluster colours failed");
    }
    // one strided H2D copy: K[f] colours per frame (pageable source: the call returns after the bytes are staged)
    std::uint32_t kmax = 0;
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        kmax = std::max(kmax, K[f]);
    }
    if (kmax != 0)
    {
        LPL_TRY(cudaMemcpy2DAsync(d_col, cap * 3, col.data(), cap * 3, static_cast<std::size_t>(kmax) * 3, nf, cudaMemcpyHostToDevice,
                                  c.stream));
    }
    launch_split_clouds(&c, nf, d_rec, cap, d_col, static_cast<std::uint32_t>(cap * 3), d_cnt3, d_tot);
    const bool want_markers = r->marker_points != nullptr;
    if (want_markers)
    {
        if (!ctx->ran_hulls)
        {
            return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "marker lines need a batch that ran LPL_STAGE_HULLS");
        }
        launch_marker_lines(&c, nf, d_mcount, d_moff, d_mtot, d_mk, mk_verts);
    }
    else
    {
        LPL_TRY(cudaMemsetAsync(d_mtot, 0, 4 * nf, c.stream));
    }
    // counts: [5][nf] = ground, obstacle, unsegmented, clustered, marker vertices
    std::uint32_t* cn = r->counts;
    for (int q = 0; q < 4; ++q)
    {
        LPL_TRY(cudaMemcpy2DAsync(cn + static_cast<std::size_t>(q) * nf, 4, d_tot + q, 16, 4, nf, k, c.stream));
    }
    LPL_TRY(cudaMemcpyAsync(cn + 4 * static_cast<std::size_t>(nf), d_mtot, 4 * nf, k, c.stream));
    LPL_TRY(cudaStreamSynchronize(c.stream));
    std::uint32_t mx[5] = {0, 0, 0, 0, 0};
    for (int q = 0; q < 5; ++q)
    {
        for (std::uint32_t f = 0; f < nf; ++f)
        {
            mx[q] = std::max(mx[q], cn[static_cast<std::size_t>(q) * nf + f]);
        }
    }
    if (r->stride < std::max(std::max(mx[0], mx[1]), std::max(mx[2], mx[3])))
    {
        return fail(ctx, LPL_ERR_CAPACITY, "host cloud stride smaller than the largest cloud");
    }
    if (want_markers && (mx[4] > mk_verts || mx[4] > r->marker_stride))
    {
        return fail(ctx, LPL_ERR_CAPACITY, "more marker vertices than reserved (half the point capacity) or than the host stride");
    }
    void* host[4] = {r->ground, r->obstacle, r->unsegmented, r->clustered};
    const std::size_t plane = B * cap * 32; // device plane q starts at q * nf * cap records (launch_split_clouds)
    (void)plane;
    for (int q = 0; q < 4; ++q)
    {
        if (host[q] != nullptr && mx[q] != 0)
        {
            LPL_TRY(cudaMemcpy2DAsync(host[q], r->stride * 32, d_rec + static_cast<std::size_t>(q) * nf * cap * 32, cap * 32,
                                      static_cast<std::size_t>(mx[q]) * 32, nf, k, c.stream));
        }
    }
    if (want_markers && mx[4] != 0)
    {
        LPL_TRY(cudaMemcpy2DAsync(r->marker_points, r->marker_stride * 24, d_mk, mk_verts * 24, static_cast<std::size_t>(mx[4]) * 24, nf, k,
                                  c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

// ------------------------------------------------------------------------------------------
// general nearest-neighbour queries (KDTree<float, 3> of the reference, kdtree.hpp:216-400), batched
// ------------------------------------------------------------------------------------------
int lpl_knn_build(lpl_ctx* ctx, const void* points, std::size_t stride, std::uint32_t n)
{
    if (ctx == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    const int rc = stage_points(ctx, points, stride, n, c.d.pts_in, c.d.n_in, -1);
    if (rc != 0)
    {
        return rc;
    }
    LPL_TRY(cudaStreamSynchronize(c.stream)); // the staging buffer is reused by the queries
    ctx->knn_n = n;
    ctx->knn_built = true;
    ctx->knn_token += 1;
    return LPL_OK;
}

unsigned long long lpl_knn_token(const lpl_ctx* ctx)
{
    return (ctx != nullptr && ctx->knn_built) ? ctx->knn_token : 0ULL;
}

// mode 0: k nearest (within radius when one is given); mode 1: everything within the radius (first `k` by index)
static int knn_queries(lpl_ctx* ctx, int mode, const void* queries, std::size_t stride, std::uint32_t m, std::uint32_t k,
                       const float* radius_sqr, float radius_all, std::uint32_t* idx_out, float* dist_out, std::uint32_t* count_out)
{
    if (ctx == nullptr || count_out == nullptr || (m != 0 && queries == nullptr) || stride < 12 || k == 0 ||
        (m != 0 && (idx_out == nullptr || dist_out == nullptr)))
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "bad neighbour query");
    }
    if (!ctx->knn_built)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "lpl_knn_build has not been called on this context");
    }
    if (mode == 0 && k > 128)
    {
        return fail(ctx, LPL_ERR_CAPACITY, "k_nearest supports up to 128 neighbours per query");
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaSetDevice(c.device));
    const std::uint32_t chunk = 16384;
    const std::size_t per_q = 16 + 4 + 4 + static_cast<std::size_t>(k) * 8;
    const std::size_t need = per_q * chunk + 1024;
    if (ctx->knn_bytes < need)
    {
        LPL_TRY(cudaStreamSynchronize(c.stream));
        if (ctx->knn_dev != nullptr)
        {
            cudaFree(ctx->knn_dev);
            ctx->knn_dev = nullptr;
            ctx->knn_bytes = 0;
        }
        if (cudaMalloc(&ctx->knn_dev, need) != cudaSuccess)
        {
            cudaGetLastError();
            return fail(ctx, LPL_ERR_CAPACITY, "device allocation for the neighbour queries failed");
        }
        ctx->knn_bytes = need;
    }
    auto* base = static_cast<unsigned char*>(ctx->knn_dev);
    auto* d_q = reinterpret_cast<float4*>(base);
    auto* d_r = reinterpret_cast<float*>(base + static_cast<std::size_t>(chunk) * 16);
    auto* d_cnt = reinterpret_cast<std::uint32_t*>(base + static_cast<std::size_t>(chunk) * 20);
    auto* d_d = reinterpret_cast<float*>(base + static_cast<std::size_t>(chunk) * 24);
    auto* d_i = reinterpret_cast<std::uint32_t*>(base + static_cast<std::size_t>(chunk) * 24 + static_cast<std::size_t>(chunk) * k * 4);
    if (ensure_stage(ctx, static_cast<std::size_t>(chunk) * 16) != 0)
    {
        return LPL_ERR_CUDA;
    }
    for (std::uint32_t q0 = 0; q0 < m; q0 += chunk)
    {
        const std::uint32_t mc = std::min(chunk, m - q0);
        LPL_TRY(cudaStreamSynchronize(c.stream)); // staging free again
        auto* hp = static_cast<char*>(c.h_stage);
        if (stride < 16)
        {
            std::memset(hp, 0, static_cast<std::size_t>(mc) * 16);
        }
        pack(hp, 16, static_cast<const char*>(queries) + static_cast<std::size_t>(q0) * stride, stride, stride < 16 ? 12 : 16, mc);
        LPL_TRY(cudaMemcpyAsync(d_q, hp, static_cast<std::size_t>(mc) * 16, cudaMemcpyHostToDevice, c.stream));
        if (radius_sqr != nullptr)
        {
            LPL_TRY(cudaMemcpyAsync(d_r, radius_sqr + q0, static_cast<std::size_t>(mc) * 4, cudaMemcpyHostToDevice, c.stream));
        }
        if (mode == 0)
        {
            if (launch_knn(&c, c.d.pts_in, ctx->knn_n, d_q, mc, k, radius_sqr != nullptr ? d_r : nullptr, radius_all, d_d, d_i, d_cnt) != 0)
            {
                return fail(ctx, LPL_ERR_CAPACITY, "unsupported k");
            }
        }
        else
        {
            launch_radius(&c, c.d.pts_in, ctx->knn_n, d_q, mc, radius_sqr != nullptr ? d_r : nullptr, radius_all, k, d_d, d_i, d_cnt);
        }
        LPL_TRY(cudaMemcpyAsync(count_out + q0, d_cnt, static_cast<std::size_t>(mc) * 4, cudaMemcpyDeviceToHost, c.stream));
        LPL_TRY(cudaMemcpyAsync(dist_out + static_cast<std::size_t>(q0) * k, d_d, static_cast<std::size_t>(mc) * k * 4,
                                cudaMemcpyDeviceToHost, c.stream));
        LPL_TRY(cudaMemcpyAsync(idx_out + static_cast<std::size_t>(q0) * k, d_i, static_cast<std::size_t>(mc) * k * 4,
                                cudaMemcpyDeviceToHost, c.stream));
    }
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaGetLastError());
    return LPL_OK;
}

int lpl_knn_k_nearest(lpl_ctx* ctx, const void* queries, std::size_t stride, std::uint32_t m, std::uint32_t k, const float* radius_sqr,
                      std::uint32_t* idx_out, float* dist_out, std::uint32_t* count_out)
{
    return knn_queries(ctx, 0, queries, stride, m, k, radius_sqr, std::numeric_limits<float>::infinity(), idx_out, dist_out, count_out);
}

int lpl_knn_radius_search(lpl_ctx* ctx, const void* queries, std::size_t stride, std::uint32_t m, const float* radius_sqr,
                          std::uint32_t max_per_query, std::uint32_t* idx_out, float* dist_out, std::uint32_t* count_out)
{
    if (radius_sqr == nullptr)
    {
        return fail(ctx, LPL_ERR_INVALID_ARGUMENT, "radius_search needs a squared radius per query");
    }
    return knn_queries(ctx, 1, queries, stride, m, max_per_query, radius_sqr, 0.f, idx_out, dist_out, count_out);
}

int lpl_host_alloc(void** out, std::size_t bytes)
{
    if (out == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    return cudaMallocHost(out, bytes) == cudaSuccess ? LPL_OK : LPL_ERR_CUDA;
}

void lpl_host_free(void* p)
{
    if (p != nullptr)
    {
        cudaFreeHost(p);
    }
}

int lpl_profile_enable(lpl_ctx* ctx, int enable)
{
    if (ctx == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaSetDevice(c.device));
    if (enable && c.prof_ev[0] == nullptr)
    {
        for (cudaEvent_t& e : c.prof_ev)
        {
            LPL_TRY(cudaEventCreate(&e));
        }
    }
    c.prof_on = enable != 0;
    c.prof_n = 0;
    return LPL_OK;
}

int lpl_profile_read(lpl_ctx* ctx, std::uint32_t max_entries, const char** names_out, float* ms_out,
                     std::uint32_t* count_out)
{
    if (ctx == nullptr || count_out == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    *count_out = 0;
    if (!c.prof_on || c.prof_n == 0)
    {
        return LPL_OK;
    }
    LPL_TRY(cudaEventSynchronize(c.prof_ev[c.prof_n]));
    const std::uint32_t m = std::min<std::uint32_t>(max_entries, static_cast<std::uint32_t>(c.prof_n));
    for (std::uint32_t i = 0; i < m; ++i)
    {
        float ms = 0.f;
        LPL_TRY(cudaEventElapsedTime(&ms, c.prof_ev[i], c.prof_ev[i + 1]));
        if (names_out != nullptr)
        {
            names_out[i] = c.prof_name[i];
        }
        if (ms_out != nullptr)
        {
            ms_out[i] = ms;
        }
    }
    *count_out = m;
    return LPL_OK;
}

// ------------------------------------------------------------------------------------------
int lpl_timer_start(lpl_ctx* ctx)
{
    if (ctx == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    LPL_TRY(cudaEventRecord(ctx->c.ev0, ctx->c.stream));
    return LPL_OK;
}

int lpl_timer_stop_ms(lpl_ctx* ctx, float* ms_out)
{
    if (ctx == nullptr || ms_out == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    LPL_TRY(cudaEventRecord(ctx->c.ev1, ctx->c.stream));
    LPL_TRY(cudaEventSynchronize(ctx->c.ev1));
    LPL_TRY(cudaEventElapsedTime(ms_out, ctx->c.ev0, ctx->c.ev1));
    return LPL_OK;
}

std::uint64_t lpl_launch_count(lpl_ctx* ctx, int reset)
{
    if (ctx == nullptr)
    {
        return 0;
    }
    const std::uint64_t v = ctx->c.launches;
    if (reset)
    {
        ctx->c.launches = 0;
    }
    return v;
}

int lpl_debug_segment(lpl_ctx* ctx, std::uint32_t f, float* elevation, float* plane, std::uint32_t* best_inliers,
                      std::uint32_t* counters)
{
    if (ctx == nullptr || f >= ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    Dev& d = c.d;
    const SegParams& s = c.seg;
    LPL_TRY(cudaStreamSynchronize(c.stream));
    if (elevation != nullptr)
    {
        LPL_TRY(cudaMemcpy(elevation, d.elev + static_cast<std::size_t>(f) * s.ncell, sizeof(float) * s.ncell,
                           cudaMemcpyDeviceToHost));
    }
    if (plane != nullptr)
    {
        LPL_TRY(cudaMemcpy(plane, d.best_plane + f, sizeof(float4), cudaMemcpyDeviceToHost));
    }
    if (best_inliers != nullptr)
    {
        LPL_TRY(cudaMemcpy(best_inliers, d.best_cnt + f, 4, cudaMemcpyDeviceToHost));
    }
    if (counters != nullptr)
    {
        LPL_TRY(cudaMemcpy(&counters[0], d.n_binned + f, 4, cudaMemcpyDeviceToHost));
        LPL_TRY(cudaMemcpy(&counters[1], d.n_cand + f, 4, cudaMemcpyDeviceToHost));
        LPL_TRY(cudaMemcpy(&counters[2], d.n_queue + f, 4, cudaMemcpyDeviceToHost));
        LPL_TRY(cudaMemcpy(&counters[3], d.jcp_rounds + f, 4, cudaMemcpyDeviceToHost));
        LPL_TRY(cudaMemcpy(&counters[4], d.n_border + f, 4, cudaMemcpyDeviceToHost));
        counters[5] = static_cast<std::uint32_t>(s.slices);
        counters[6] = static_cast<std::uint32_t>(s.rings);
        LPL_TRY(cudaMemcpy(&counters[7], d.status + f, 4, cudaMemcpyDeviceToHost));
    }
    return LPL_OK;
}

int lpl_debug_dror(lpl_ctx* ctx, std::uint32_t f, std::uint32_t* n_unresolved)
{
    if (ctx == nullptr || n_unresolved == nullptr || f >= ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaMemcpy(n_unresolved, c.d.n_unres + f, 4, cudaMemcpyDeviceToHost));
    return LPL_OK;
}

int lpl_debug_cluster(lpl_ctx* ctx, std::uint32_t f, std::int32_t* dims)
{
    if (ctx == nullptr || dims == nullptr || f >= ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaStreamSynchronize(c.stream));
    float mx[4];
    LPL_TRY(cudaMemcpy(mx, c.d.sph_max + f * 4, 16, cudaMemcpyDeviceToHost));
    dims[0] = static_cast<std::int32_t>(std::ceil(mx[0] / c.clu.range_res) + 1);
    dims[1] = static_cast<std::int32_t>(std::ceil(mx[1] / c.clu.az_res) + 1);
    dims[2] = static_cast<std::int32_t>(std::ceil(mx[2] / c.clu.el_res) + 1);
    return LPL_OK;
}

int lpl_debug_hulls(lpl_ctx* ctx, std::uint32_t f, std::uint32_t* counters)
{
    if (ctx == nullptr || counters == nullptr || f >= ctx->c.d.B)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    Ctx& c = ctx->c;
    LPL_TRY(cudaStreamSynchronize(c.stream));
    LPL_TRY(cudaMemcpy(&counters[0], c.d.n_h + f, 4, cudaMemcpyDeviceToHost));
    LPL_TRY(cudaMemcpy(&counters[1], c.d.n_vox + f, 4, cudaMemcpyDeviceToHost));
    return LPL_OK;
}

void* lpl_stream(lpl_ctx* ctx) { return ctx != nullptr ? static_cast<void*>(ctx->c.stream) : nullptr; }
} // extern "C"
