// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement ("port") of the reference hot path.
//
// This file restates, in plain C++17 with no third-party dependency, what the reference
// library computes for the per-frame perception path. It is the checker for the CUDA path
// (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is never linked into or
// called from the product library. It is pinned against the UNMODIFIED reference
// (oracle/_ref/libref_oracle.so, built from /root/reference by oracle/Makefile) on the 154
// KITTI frames by tests/test_oracle_cpu.py and tools/make_golden.py; the reference's own test
// suite holds no vectors for this path.
//
// Reference (paths relative to /root/reference):
//   ring partition   src/dataloader/src/dataloader.cpp:68-137          (Dataloader::addRingInfo)
//   DROR             lidar_processing_lib/src/noise_remover.cpp:38-68, include/.../kdtree.hpp:131-149,339-397
//   segmentation     lidar_processing_lib/src/segmenter.cpp:38-71,103-204,206-319,321-479,481-638,640-669
//   atan2Approx      lidar_processing_lib/include/lidar_processing_lib/common.hpp:33-62
//   clustering       lidar_processing_lib/src/clusterer.cpp:55-100,102-120,122-193,195-239
//   cluster gather   src/processor/src/processor.cpp:627-658
//   convex hull      lidar_processing_lib/src/polygonizer.cpp:33-91
//
// The decomposition deliberately differs from the reference's (flat sorted cell lists instead
// of vector-of-vectors, min-key range image, union-find instead of BFS, data-flow JCP) because
// it doubles as the executable specification of the CUDA kernels' algorithms.
#include <algorithm>
#include <array>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace
{
constexpr float kDegToRad = static_cast<float>(M_PI / 180.0); // segmenter.hpp:135
constexpr float kTwoPi = static_cast<float>(2.0 * M_PI);      // segmenter.hpp:136
constexpr float kPi = 3.14159265358979323846f;                // M_PIf
constexpr float kPi2 = 1.57079632679489661923f;               // M_PI_2f

// common.hpp:33-62 — plain float ops in the written order.
inline float atan2_approx(float y, float x)
{
    const float ax = std::fabs(x);
    const float ay = std::fabs(y);
    const float mx = std::max(ay, ax);
    const float mn = std::min(ay, ax);
    const float a = mn / mx;
    const float s = a * a;
    const float c = s * a;
    const float q = s * s;
    float r = 0.024840285F * q + 0.18681418F;
    const float t = -0.094097948F * q - 0.33213072F;
    r = r * s + t;
    r = r * c + a;
    if (ay > ax)
    {
        r = 1.57079637F - r;
    }
    if (x < 0)
    {
        r = 3.14159274F - r;
    }
    if (y < 0)
    {
        r = -r;
    }
    return r;
}

// ---- std::mt19937 + libstdc++ uniform_int_distribution<uint32_t> (segmenter.cpp:369-371) ----
struct Mt19937
{
    std::uint32_t mt[624];
    int idx;
    explicit Mt19937(std::uint32_t seed)
    {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i)
        {
            mt[i] = 1812433253U * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<std::uint32_t>(i);
        }
        idx = 624;
    }
    std::uint32_t next()
    {
        if (idx >= 624)
        {
            for (int i = 0; i < 624; ++i)
            {
                const std::uint32_t y = (mt[i] & 0x80000000U) | (mt[(i + 1) % 624] & 0x7fffffffU);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
            }
            idx = 0;
        }
        std::uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= y >> 18;
        return y;
    }
};

// libstdc++ (GCC >= 11) maps a 32-bit engine to [0, n) with Lemire's multiply-shift + rejection.
inline std::uint32_t uniform_below(Mt19937& g, std::uint32_t n)
{
    std::uint64_t product = static_cast<std::uint64_t>(g.next()) * n;
    std::uint32_t low = static_cast<std::uint32_t>(product);
    if (low < n)
    {
        const std::uint32_t threshold = (0U - n) % n;
        while (low < threshold)
        {
            product = static_cast<std::uint64_t>(g.next()) * n;
            low = static_cast<std::uint32_t>(product);
        }
    }
    return static_cast<std::uint32_t>(product >> 32);
}
} // namespace

extern "C"
{
// =================================================================== ring partition
// dataloader.cpp:94-134. ring starts at 63 and decrements (floor 0) at every FOURTH->FIRST
// quadrant transition in point order.
void port_ring_partition(const float* xyz, std::int32_t stride_f, std::uint32_t n, std::uint16_t* ring)
{
    int prev_q = 0;
    std::uint16_t ring_index = 63;
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        const float y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        float az = std::atan2(y, x);
        az = (az < 0) ? (az + 2.0F * kPi) : az;
        int q;
        if (az < kPi2)
        {
            q = 0;
        }
        else if (az < kPi)
        {
            q = 1;
        }
        else if (az < 1.5F * kPi)
        {
            q = 2;
        }
        else
        {
            q = 3;
        }
        if (q == 0 && prev_q == 3 && ring_index > 0U)
        {
            --ring_index;
        }
        prev_q = q;
        ring[i] = ring_index;
    }
}

// =================================================================== DROR
// noise_remover.cpp:38-68 with the KD-tree replaced by an exhaustive count over a uniform
// grid ("exact" semantics: NOISE iff fewer than min_neighbours points, self included, satisfy
// dist_sqr <= r_sqr; see SURVEY.md hazard H1 for why the as-is reference differs).
void port_dror(const float* xyz,
               std::int32_t stride_f,
               std::uint32_t n,
               float radius_multiplier,
               float min_radius,
               std::uint32_t min_neighbours,
               std::uint8_t* labels)
{
    if (n == 0)
    {
        return;
    }
    const double scaling = std::pow(static_cast<double>(radius_multiplier), 2.0);
    const float min_r_sqr = min_radius * min_radius;
    const float cell = 1.0f;
    const int G = 512; // +-256 m, clamped
    auto cell_of = [&](float v) {
        int c = static_cast<int>(std::floor(v / cell)) + G / 2;
        return std::min(std::max(c, 0), G - 1);
    };
    std::vector<std::uint32_t> start(static_cast<std::size_t>(G) * G + 1, 0);
    std::vector<std::uint32_t> cid(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        const float y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        cid[i] = static_cast<std::uint32_t>(cell_of(y)) * G + cell_of(x);
        ++start[cid[i] + 1];
    }
    for (std::size_t c = 0; c < static_cast<std::size_t>(G) * G; ++c)
    {
        start[c + 1] += start[c];
    }
    std::vector<std::uint32_t> fill(start.begin(), start.end() - 1);
    std::vector<std::uint32_t> order(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        order[fill[cid[i]]++] = i;
    }
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float* p = xyz + static_cast<std::size_t>(i) * stride_f;
        const double range_sqr = (p[0] * p[0]) + (p[1] * p[1]); // float expression widened
        const float r_sqr = std::max(static_cast<float>(scaling * range_sqr), min_r_sqr);
        const float r = std::sqrt(r_sqr) * 1.0001f + 1e-6f; // conservative cell cover
        const int x0 = cell_of(p[0] - r), x1 = cell_of(p[0] + r);
        const int y0 = cell_of(p[1] - r), y1 = cell_of(p[1] + r);
        std::uint32_t count = 0;
        for (int cy = y0; cy <= y1 && count < min_neighbours; ++cy)
        {
            const std::uint32_t a = start[static_cast<std::size_t>(cy) * G + x0];
            const std::uint32_t b = start[static_cast<std::size_t>(cy) * G + x1 + 1];
            for (std::uint32_t k = a; k < b; ++k)
            {
                const float* q = xyz + static_cast<std::size_t>(order[k]) * stride_f;
                // kdtree.hpp:131-143: (a0-b0)^2 + ((a1-b1)^2 + ((a2-b2)^2 + 0)), a = target
                const float d0 = p[0] - q[0];
                const float d1 = p[1] - q[1];
                const float d2 = p[2] - q[2];
                const float dist = d0 * d0 + (d1 * d1 + (d2 * d2 + 0.0f));
                if (dist <= r_sqr)
                {
                    if (++count >= min_neighbours)
                    {
                        break;
                    }
                }
            }
        }
        labels[i] = (count < min_neighbours) ? 1 : 0;
    }
}

// =================================================================== segmentation
struct port_seg_cfg
{
    float elevation_up_deg;
    float elevation_down_deg;
    std::int32_t image_width;
    std::int32_t image_height;
    std::int32_t assume_unorganized_cloud;
    float grid_radial_spacing_m;
    float grid_slice_resolution_deg;
    float ground_height_threshold_m;
    float road_maximum_slope_m_per_m;
    float min_distance_m;
    float max_distance_m;
    float sensor_height_m;
    float kernel_threshold_distance_m;
    float amplification_factor;
    float z_min_m;
    float z_max_m;
};

struct port_seg_debug
{
    float* elevation;           // [slices*rings] or null
    std::int32_t* cloud_map;    // [H*W] or null
    std::uint8_t* pre_jcp_code; // [H*W] or null: 0 empty, 1 ground, 2 obstacle, 3 queued
    float plane[4];
    std::uint32_t best_inliers;
    std::uint32_t n_candidates;
    std::uint32_t n_binned;
    std::uint32_t n_queued;
    std::uint32_t n_undecided;
    std::uint32_t rounds;      // data-flow rounds (jcp_mode 2/3)
    std::uint32_t max_cell;    // largest polar cell
    std::uint32_t n_nonempty_cells;
    std::int32_t slices;
    std::int32_t rings;
};

enum : std::uint8_t
{
    PX_EMPTY = 0,
    PX_GROUND = 1,
    PX_OBSTACLE = 2,
    PX_QUEUED = 3
};

static const int kJcpOff[24][2] = { // segmenter.cpp:527-530 (height, width)
    {-2, -2}, {-2, -1}, {-2, 0}, {-2, 1}, {-2, 2}, {-1, -2}, {-1, -1}, {-1, 0},
    {-1, 1},  {-1, 2},  {0, -2}, {0, -1}, {0, 1},  {0, 2},   {1, -2},  {1, -1},
    {1, 0},   {1, 1},   {1, 2},  {2, -2}, {2, -1}, {2, 0},   {2, 1},   {2, 2}};

// jcp_mode: 0 = as-is, sequential raster emulation (stale out-of-image slots, hazard H2)
//           1 = clean, sequential (out-of-image slots contribute nothing)
//           2 = as-is, data-flow formulation (what the CUDA kernels implement)
//           3 = clean, data-flow formulation
int port_segment(const port_seg_cfg* cfg,
                 const float* xyz,
                 std::int32_t stride_f,
                 const std::uint16_t* ring,
                 std::uint32_t n,
                 std::int32_t jcp_mode,
                 std::uint32_t* labels,
                 std::uint8_t* bgr,
                 port_seg_debug* dbg)
{
    const int W = cfg->image_width;
    const int H = cfg->image_height;
    // ---- derived constants (segmenter.cpp:42-46,106-109,121-122,209-211,359-360,532-533)
    const float slice_res = cfg->grid_slice_resolution_deg * kDegToRad;
    const int rings = static_cast<std::int32_t>(cfg->max_distance_m / cfg->grid_radial_spacing_m);
    const int slices = static_cast<std::int32_t>(kTwoPi / slice_res);
    const float el_up = cfg->elevation_up_deg * kDegToRad;
    const float el_down = cfg->elevation_down_deg * kDegToRad;
    const float vfov = el_up - el_down;
    const float rad_per_px = vfov / H;
    const float z_lo = -cfg->sensor_height_m + cfg->z_min_m;
    const float z_hi = -cfg->sensor_height_m + cfg->z_max_m;
    const float thr = cfg->ground_height_threshold_m;
    const float delta = std::min(cfg->grid_radial_spacing_m * std::tan(cfg->road_maximum_slope_m_per_m),
                                 thr - std::numeric_limits<float>::epsilon());
    const float e0 = -cfg->sensor_height_m + thr;
    const float cos_max = std::cos(std::tan(cfg->road_maximum_slope_m_per_m));
    const float kthr_sqr = cfg->kernel_threshold_distance_m * cfg->kernel_threshold_distance_m;
    const bool use_ring = (ring != nullptr) && (cfg->assume_unorganized_cloud == 0);
    const int ncell = rings * slices;

    std::fill(labels, labels + n, 0U);
    if (dbg != nullptr)
    {
        dbg->slices = slices;
        dbg->rings = rings;
        dbg->plane[0] = 0.f;
        dbg->plane[1] = 0.f;
        dbg->plane[2] = 1.f;
        dbg->plane[3] = 0.f;
        dbg->best_inliers = dbg->n_candidates = dbg->n_binned = dbg->n_queued = 0;
        dbg->n_undecided = dbg->rounds = dbg->max_cell = dbg->n_nonempty_cells = 0;
    }

    // ---- a4: per-point binning (segmenter.cpp:124-203)
    std::vector<std::int32_t> cell_of(n, -1);
    std::vector<std::uint32_t> px_of(n, 0);
    std::vector<std::uint32_t> cell_start(static_cast<std::size_t>(ncell) + 1, 0);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        const float y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        const float z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
        if (z < z_lo || z > z_hi)
        {
            continue;
        }
        const float dist = std::sqrt(x * x + y * y);
        const int radial = static_cast<std::int32_t>(dist / cfg->grid_radial_spacing_m);
        if (dist < cfg->min_distance_m || dist > cfg->max_distance_m || radial >= rings)
        {
            continue;
        }
        float az = atan2_approx(y, x);
        az = (az < 0) ? (az + kTwoPi) : az;
        const int az_idx = std::min(static_cast<std::int32_t>(az / slice_res), slices - 1);
        std::int32_t hgt;
        if (use_ring)
        {
            hgt = ring[i];
            if (hgt >= H)
            {
                continue;
            }
        }
        else
        {
            const float el = std::atan(z / dist);
            hgt = static_cast<std::int32_t>((el - el_down) / rad_per_px);
            if (hgt < 0 || hgt >= H)
            {
                continue;
            }
        }
        const std::uint16_t wid = static_cast<std::uint16_t>((W - 1) * az / kTwoPi);
        cell_of[i] = az_idx * rings + radial;
        px_of[i] = static_cast<std::uint32_t>(hgt) * W + wid;
        ++cell_start[cell_of[i] + 1];
    }
    for (int c = 0; c < ncell; ++c)
    {
        cell_start[c + 1] += cell_start[c];
    }
    const std::uint32_t nb = cell_start[ncell];
    // stable counting sort: position in `order` == (cell, cloud index) iteration order
    std::vector<std::uint32_t> order(nb);
    {
        std::vector<std::uint32_t> fill(cell_start.begin(), cell_start.end() - 1);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            if (cell_of[i] >= 0)
            {
                order[fill[cell_of[i]]++] = i;
            }
        }
    }
    auto X = [&](std::uint32_t i) { return xyz[static_cast<std::size_t>(i) * stride_f + 0]; };
    auto Y = [&](std::uint32_t i) { return xyz[static_cast<std::size_t>(i) * stride_f + 1]; };
    auto Z = [&](std::uint32_t i) { return xyz[static_cast<std::size_t>(i) * stride_f + 2]; };

    // ---- a5: per-cell robust minimum + sequential radial recurrence (segmenter.cpp:215-267)
    std::vector<float> cell_zmin(ncell, 0.f);
    std::vector<float> zs;
    std::uint32_t max_cell = 0, nonempty = 0;
    for (int c = 0; c < ncell; ++c)
    {
        const std::uint32_t a = cell_start[c], b = cell_start[c + 1];
        if (a == b)
        {
            continue;
        }
        ++nonempty;
        max_cell = std::max(max_cell, b - a);
        zs.clear();
        for (std::uint32_t k = a; k < b; ++k)
        {
            zs.push_back(Z(order[k]));
        }
        std::sort(zs.begin(), zs.end());
        float zmin = zs[0];
        for (std::int32_t i = static_cast<std::int32_t>(zs.size() / 2); i >= 1; --i)
        {
            if (zs[i] - zs[i - 1] > 0.5F)
            {
                zmin = zs[i];
                break;
            }
        }
        cell_zmin[c] = zmin;
    }
    std::vector<float> elev(ncell);
    for (int s = 0; s < slices; ++s)
    {
        float prev = e0;
        elev[s * rings] = e0;
        for (int r = 1; r < rings; ++r)
        {
            const int c = s * rings + r;
            float e = prev + delta;
            if (cell_start[c] != cell_start[c + 1])
            {
                e = std::min(cell_zmin[c], prev + delta);
            }
            elev[c] = e;
            prev = e;
        }
    }
    // obstacle classification (segmenter.cpp:271-283); lab is indexed by sorted position
    std::vector<std::uint8_t> lab(nb, PX_GROUND);
    for (int c = 0; c < ncell; ++c)
    {
        const float lim = elev[c] + thr;
        for (std::uint32_t k = cell_start[c]; k < cell_start[c + 1]; ++k)
        {
            if (Z(order[k]) >= lim)
            {
                lab[k] = PX_OBSTACLE;
            }
        }
    }

    // ---- a6: near-field RANSAC (segmenter.cpp:321-479)
    {
        const int kBins = 4;
        std::vector<std::uint32_t> cand; // sorted positions, (slice, bin, cloud) order
        for (int s = 0; s < slices; ++s)
        {
            for (int r = 0; r < kBins && r < rings; ++r)
            {
                const int c = s * rings + r;
                for (std::uint32_t k = cell_start[c]; k < cell_start[c + 1]; ++k)
                {
                    if (std::fabs(elev[c] - Z(order[k])) < 2.0F * thr)
                    {
                        cand.push_back(k);
                    }
                }
            }
        }
        const std::uint32_t nc = static_cast<std::uint32_t>(cand.size());
        float pa = 0.f, pb = 0.f, pc = 1.f, pd = 0.f;
        std::uint32_t best = 0;
        // nc == 1 makes the reference spin forever in `while (p3_index == p2_index)`
        // (segmenter.cpp:382-386); every implementation here skips RANSAC for nc < 2.
        if (nc >= 2)
        {
            const float p1x = 0.0f, p1y = 0.0f, p1z = -cfg->sensor_height_m;
            Mt19937 gen(42);
            for (int it = 0; it < 60; ++it)
            {
                const std::uint32_t i2 = uniform_below(gen, nc);
                std::uint32_t i3 = uniform_below(gen, nc);
                while (i3 == i2)
                {
                    i3 = uniform_below(gen, nc);
                }
                const std::uint32_t q2 = order[cand[i2]], q3 = order[cand[i3]];
                const float p2x = X(q2), p2y = Y(q2), p2z = Z(q2);
                const float p3x = X(q3), p3y = Y(q3), p3z = Z(q3);
                float nx = ((p2y - p1y) * (p3z - p1z)) - ((p2z - p1z) * (p3y - p1y));
                float ny = ((p2z - p1z) * (p3x - p1x)) - ((p2x - p1x) * (p3z - p1z));
                float nz = ((p2x - p1x) * (p3y - p1y)) - ((p2y - p1y) * (p3x - p1x));
                const float den = (nx * nx) + (ny * ny) + (nz * nz);
                if (den < 1.0e-5F)
                {
                    continue;
                }
                const float norm = 1.0F / std::sqrt(den);
                nz *= norm;
                if (std::fabs(nz) < cos_max)
                {
                    continue;
                }
                nx *= norm;
                ny *= norm;
                const float d = (nx * p1x) + (ny * p1y) + (nz * p1z);
                std::uint32_t inl = 0;
                for (std::uint32_t k : cand)
                {
                    const std::uint32_t q = order[k];
                    const float od = std::fabs((nx * X(q)) + (ny * Y(q)) + (nz * Z(q)) - d);
                    if (od < thr)
                    {
                        ++inl;
                    }
                }
                if (inl > best)
                {
                    best = inl;
                    pa = nx;
                    pb = ny;
                    pc = nz;
                    pd = d;
                }
            }
            if (best > 0)
            {
                if (pc < 0)
                {
                    pa = -pa;
                    pb = -pb;
                    pc = -pc;
                    pd = -pd;
                }
                for (int s = 0; s < slices; ++s)
                {
                    for (int r = 0; r < kBins && r < rings; ++r)
                    {
                        const int c = s * rings + r;
                        for (std::uint32_t k = cell_start[c]; k < cell_start[c + 1]; ++k)
                        {
                            const std::uint32_t q = order[k];
                            const float sd = (pa * X(q)) + (pb * Y(q)) + (pc * Z(q)) - pd;
                            if (sd < thr)
                            {
                                lab[k] = PX_GROUND;
                            }
                        }
                    }
                }
            }
        }
        if (dbg != nullptr)
        {
            dbg->plane[0] = pa;
            dbg->plane[1] = pb;
            dbg->plane[2] = pc;
            dbg->plane[3] = pd;
            dbg->best_inliers = best;
            dbg->n_candidates = nc;
        }
    }

    // ---- a7: range image = per-pixel minimum of (depth_sqr bits, sorted position)
    //      (segmenter.cpp:291-318: strict `<`, first in (cell, cloud) order wins ties)
    const std::size_t npx = static_cast<std::size_t>(H) * W;
    std::vector<std::uint64_t> key(npx, ~0ULL);
    for (std::uint32_t k = 0; k < nb; ++k)
    {
        const std::uint32_t q = order[k];
        const float d2 = (X(q) * X(q)) + (Y(q) * Y(q));
        std::uint32_t bits;
        std::memcpy(&bits, &d2, 4);
        const std::uint64_t kk = (static_cast<std::uint64_t>(bits) << 32) | k;
        key[px_of[q]] = std::min(key[px_of[q]], kk);
    }
    std::vector<std::int32_t> cmap(npx, -1);
    std::vector<std::uint8_t> code(npx, PX_EMPTY);
    for (std::size_t p = 0; p < npx; ++p)
    {
        if (key[p] != ~0ULL)
        {
            const std::uint32_t k = static_cast<std::uint32_t>(key[p] & 0xffffffffULL);
            cmap[p] = static_cast<std::int32_t>(order[k]);
            code[p] = lab[k];
        }
    }

    // ---- a8 part 1: 5x5 in-bounds dilation of the obstacle (red) channel + queue flags
    //      (segmenter.cpp:486-514)
    std::vector<std::uint8_t> dil(npx, 0);
    for (int h = 0; h < H; ++h)
    {
        for (int w = 0; w < W; ++w)
        {
            std::uint8_t m = 0;
            for (int dh = -2; dh <= 2 && !m; ++dh)
            {
                const int hh = h + dh;
                if (hh < 0 || hh >= H)
                {
                    continue;
                }
                for (int dw = -2; dw <= 2; ++dw)
                {
                    const int ww = w + dw;
                    if (ww >= 0 && ww < W && code[static_cast<std::size_t>(hh) * W + ww] == PX_OBSTACLE)
                    {
                        m = 1;
                        break;
                    }
                }
            }
            dil[static_cast<std::size_t>(h) * W + w] = m;
        }
    }
    std::vector<std::uint32_t> queue; // raster order
    for (std::size_t p = 0; p < npx; ++p)
    {
        if (code[p] == PX_GROUND && dil[p])
        {
            code[p] = PX_QUEUED; // (0,255,255) with a mapped point -> CV_INTERSECTION
            queue.push_back(static_cast<std::uint32_t>(p));
        }
    }
    if (dbg != nullptr)
    {
        dbg->n_binned = nb;
        dbg->n_queued = static_cast<std::uint32_t>(queue.size());
        dbg->max_cell = max_cell;
        dbg->n_nonempty_cells = nonempty;
        if (dbg->elevation != nullptr)
        {
            std::copy(elev.begin(), elev.end(), dbg->elevation);
        }
        if (dbg->cloud_map != nullptr)
        {
            std::copy(cmap.begin(), cmap.end(), dbg->cloud_map);
        }
        if (dbg->pre_jcp_code != nullptr)
        {
            std::copy(code.begin(), code.end(), dbg->pre_jcp_code);
        }
    }

    // ---- a8 part 2: jump-convolution relaxation (segmenter.cpp:523-637)
    // final[p]: 1 ground, 2 obstacle, 3 still queued (pending), 4 undecided (stays CV_INTERSECTION)
    std::vector<std::uint8_t> state(code);
    const float amp = cfg->amplification_factor;
    auto slot_static = [&](int h, int w, int i, std::uint32_t core, float& wgt, std::uint8_t& msk, std::uint32_t& dyn) {
        // weight / mask of in-image slot i seen from queued pixel (h, w); dyn = pixel id whose
        // *final* state supplies the mask (only slots 0..11 precede the pixel in raster order)
        const int hh = h + kJcpOff[i][0], ww = w + kJcpOff[i][1];
        const std::size_t np = static_cast<std::size_t>(hh) * W + ww;
        dyn = 0xffffffffU;
        const std::int32_t ni = cmap[np];
        if (ni < 0)
        {
            wgt = 0.f;
            msk = 0;
            return;
        }
        const float dx = X(core) - X(ni);
        const float dy = Y(core) - Y(ni);
        const float dz = Z(core) - Z(ni);
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > kthr_sqr)
        {
            wgt = 0.f;
            msk = 0;
            return;
        }
        wgt = std::exp(-amp * std::sqrt(d2));
        const std::uint8_t c = code[np];
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
            dyn = static_cast<std::uint32_t>(np);
        }
        else
        {
            msk = 0;
        }
    };
    auto in_image = [&](int h, int w, int i) {
        const int hh = h + kJcpOff[i][0], ww = w + kJcpOff[i][1];
        return hh >= 0 && hh < H && ww >= 0 && ww < W;
    };
    std::uint32_t undecided = 0, rounds = 0;
    if (jcp_mode == 0 || jcp_mode == 1)
    {
        float wv[24];
        std::uint8_t mv[24];
        for (int i = 0; i < 24; ++i)
        {
            wv[i] = 0.f;
            mv[i] = 0;
        }
        for (std::uint32_t p : queue)
        {
            const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
            const std::uint32_t core = static_cast<std::uint32_t>(cmap[p]);
            float sum = 0.f;
            for (int i = 0; i < 24; ++i)
            {
                if (!in_image(h, w, i))
                {
                    if (jcp_mode == 1)
                    {
                        wv[i] = 0.f;
                        mv[i] = 0;
                    }
                    continue; // as-is: slot keeps whatever the previous pixel left (H2)
                }
                std::uint32_t dyn;
                slot_static(h, w, i, core, wv[i], mv[i], dyn);
                if (wv[i] != 0.f)
                {
                    sum += wv[i];
                }
                // mask from the *current* image: already-relaxed pixels show their decision,
                // pending / undecided ones read as unknown
                if (wv[i] != 0.f)
                {
                    const std::size_t np = static_cast<std::size_t>(h + kJcpOff[i][0]) * W + (w + kJcpOff[i][1]);
                    const std::uint8_t s = state[np];
                    mv[i] = (s == 1) ? 1 : (s == 2) ? 2 : 0;
                }
            }
            if (std::fabs(sum) > std::numeric_limits<float>::epsilon())
            {
                float wo = 0.f, wg = 0.f;
                for (int i = 0; i < 24; ++i)
                {
                    const float wn = wv[i] / sum;
                    if (mv[i] == 1)
                    {
                        wg += wn;
                    }
                    else if (mv[i] == 2)
                    {
                        wo += wn;
                    }
                }
                state[p] = (wo > wg) ? 2 : 1;
            }
            else
            {
                state[p] = 4;
                ++undecided;
            }
        }
    }
    else
    {
        // Data-flow formulation. Pre-pass: per queued pixel, 24 (weight, mask-source) slots.
        const bool emulate_stale = (jcp_mode == 2);
        const std::size_t nq = queue.size();
        std::vector<std::uint32_t> qidx(npx, 0xffffffffU);
        for (std::size_t k = 0; k < nq; ++k)
        {
            qidx[queue[k]] = static_cast<std::uint32_t>(k);
        }
        std::vector<float> wn(nq * 24);
        std::vector<std::uint8_t> mk(nq * 24);
        std::vector<std::uint32_t> dy(nq * 24);
        std::vector<std::uint8_t> decidable(nq);
        for (std::size_t k = 0; k < nq; ++k)
        {
            const std::uint32_t p = queue[k];
            const int h = static_cast<int>(p / W), w = static_cast<int>(p % W);
            const std::uint32_t core = static_cast<std::uint32_t>(cmap[p]);
            float sum = 0.f;
            float wv[24];
            for (int i = 0; i < 24; ++i)
            {
                std::uint8_t m = 0;
                std::uint32_t d = 0xffffffffU;
                wv[i] = 0.f;
                if (in_image(h, w, i))
                {
                    slot_static(h, w, i, core, wv[i], m, d);
                    if (wv[i] != 0.f)
                    {
                        sum += wv[i];
                    }
                }
                else if (emulate_stale)
                {
                    // most recent earlier queued pixel for which slot i lies inside the image
                    for (std::size_t kk = k; kk-- > 0;)
                    {
                        const std::uint32_t pp = queue[kk];
                        const int h2 = static_cast<int>(pp / W), w2 = static_cast<int>(pp % W);
                        if (in_image(h2, w2, i))
                        {
                            slot_static(h2, w2, i, static_cast<std::uint32_t>(cmap[pp]), wv[i], m, d);
                            break;
                        }
                        if (h2 + kJcpOff[i][0] < 0)
                        {
                            break; // every earlier pixel is in the same or a lower row
                        }
                    }
                }
                mk[k * 24 + i] = m;
                dy[k * 24 + i] = d;
            }
            decidable[k] = std::fabs(sum) > std::numeric_limits<float>::epsilon();
            for (int i = 0; i < 24; ++i)
            {
                wn[k * 24 + i] = decidable[k] ? wv[i] / sum : 0.f;
            }
        }
        std::vector<std::uint32_t> pending(nq), next;
        std::iota(pending.begin(), pending.end(), 0U);
        while (!pending.empty())
        {
            ++rounds;
            next.clear();
            std::vector<std::pair<std::uint32_t, std::uint8_t>> fired;
            for (std::uint32_t k : pending)
            {
                bool ready = true;
                for (int i = 0; i < 24 && ready; ++i)
                {
                    if (mk[k * 24 + i] == 3 && state[dy[k * 24 + i]] == PX_QUEUED)
                    {
                        ready = false;
                    }
                }
                if (!ready)
                {
                    next.push_back(k);
                    continue;
                }
                std::uint8_t out = 4;
                if (decidable[k])
                {
                    float wo = 0.f, wg = 0.f;
                    for (int i = 0; i < 24; ++i)
                    {
                        std::uint8_t m = mk[k * 24 + i];
                        if (m == 3)
                        {
                            const std::uint8_t s = state[dy[k * 24 + i]];
                            m = (s == 1) ? 1 : (s == 2) ? 2 : 0;
                        }
                        if (m == 1)
                        {
                            wg += wn[k * 24 + i];
                        }
                        else if (m == 2)
                        {
                            wo += wn[k * 24 + i];
                        }
                    }
                    out = (wo > wg) ? 2 : 1;
                }
                fired.emplace_back(queue[k], out);
            }
            for (auto& f : fired) // Jacobi within a round (any interleaving gives the same result)
            {
                state[f.first] = f.second;
                if (f.second == 4)
                {
                    ++undecided;
                }
            }
            pending.swap(next);
        }
    }
    if (dbg != nullptr)
    {
        dbg->n_undecided = undecided;
        dbg->rounds = rounds;
    }

    // ---- a9 / a10: labels of pixel winners + BGR image (segmenter.cpp:640-669, segmenter.hpp:144-148)
    for (std::size_t p = 0; p < npx; ++p)
    {
        if (cmap[p] >= 0 && (state[p] == 1 || state[p] == 2))
        {
            labels[cmap[p]] = state[p];
        }
    }
    if (bgr != nullptr)
    {
        for (std::size_t p = 0; p < npx; ++p)
        {
            std::uint8_t b = 0, g = 0, r = 0;
            switch (state[p])
            {
            case 1: g = 255; r = 0; break;
            case 2: r = 255; break;
            case 3:
            case 4: b = 255; break;
            default: r = dil[p] ? 255 : 0; break;
            }
            // a ground pixel that was not queued keeps red = dilated value = 0 by construction;
            // an obstacle pixel has red 255 already
            bgr[p * 3 + 0] = b;
            bgr[p * 3 + 1] = g;
            bgr[p * 3 + 2] = r;
        }
    }
    return 0;
}

// =================================================================== clustering
// clusterer.cpp:55-239. Connected components of occupied curved voxels under 26-connectivity
// with the reference's literal azimuth wrap (hazard H3); ids ranked by the component's first
// point; components with fewer than min_cluster_size points become -1 and the rest are
// renumbered densely in increasing order.
int port_cluster(const float* xyz,
                 std::int32_t stride_f,
                 std::uint32_t n,
                 float range_res_m,
                 float az_res_deg,
                 float el_res_deg,
                 std::uint32_t min_cluster_size,
                 std::int32_t* labels,
                 std::int32_t* grid_dims,
                 std::uint32_t* n_voxels_out)
{
    if (n == 0)
    {
        return 0;
    }
    const float az_res = az_res_deg * kDegToRad;
    const float el_res = el_res_deg * kDegToRad;
    std::vector<float> rg(n), az(n), el(n);
    float max_r = std::numeric_limits<float>::lowest();
    float max_a = max_r, max_e = max_r;
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        const float y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        const float z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
        float a = std::atan2(y, x);
        a = (a < 0) ? (a + kTwoPi) : a;
        const float dxy2 = x * x + y * y;
        const float dxy = std::sqrt(dxy2);
        const float r = std::sqrt(dxy2 + z * z);
        const float e = std::atan(z / dxy) + kPi2;
        rg[i] = r;
        az[i] = a;
        el[i] = e;
        max_r = std::max(max_r, r);
        max_a = std::max(max_a, a);
        max_e = std::max(max_e, e);
    }
    const std::int32_t nr = static_cast<std::int32_t>(std::ceil(max_r / range_res_m) + 1);
    const std::int32_t na = static_cast<std::int32_t>(std::ceil(max_a / az_res) + 1);
    const std::int32_t ne = static_cast<std::int32_t>(std::ceil(max_e / el_res) + 1);
    if (grid_dims != nullptr)
    {
        grid_dims[0] = nr;
        grid_dims[1] = na;
        grid_dims[2] = ne;
    }
    // occupied voxels
    std::unordered_map<std::int32_t, std::uint32_t> vox; // flat index -> voxel id
    vox.reserve(n);
    std::vector<std::uint32_t> vid(n);
    std::vector<std::array<std::int32_t, 3>> vkey;
    std::vector<std::uint32_t> vmin; // first (minimum) point index per voxel
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const std::int32_t ri = static_cast<std::int32_t>(rg[i] / range_res_m);
        const std::int32_t ai = static_cast<std::int32_t>(az[i] / az_res);
        const std::int32_t ei = static_cast<std::int32_t>(el[i] / el_res);
        const std::int32_t flat = nr * (na * ei + ai) + ri;
        auto it = vox.find(flat);
        if (it == vox.end())
        {
            it = vox.emplace(flat, static_cast<std::uint32_t>(vkey.size())).first;
            vkey.push_back({ri, ai, ei});
            vmin.push_back(i);
        }
        vid[i] = it->second;
    }
    const std::uint32_t nv = static_cast<std::uint32_t>(vkey.size());
    if (n_voxels_out != nullptr)
    {
        *n_voxels_out = nv;
    }
    // union-find over voxels
    std::vector<std::uint32_t> parent(nv);
    std::iota(parent.begin(), parent.end(), 0U);
    auto find = [&](std::uint32_t v) {
        while (parent[v] != v)
        {
            parent[v] = parent[parent[v]];
            v = parent[v];
        }
        return v;
    };
    for (std::uint32_t v = 0; v < nv; ++v)
    {
        const auto [ri, ai, ei] = vkey[v];
        for (int dr = -1; dr <= 1; ++dr)
        {
            for (int da = -1; da <= 1; ++da)
            {
                for (int de = -1; de <= 1; ++de)
                {
                    if (dr == 0 && da == 0 && de == 0)
                    {
                        continue;
                    }
                    const int r2 = ri + dr, e2 = ei + de;
                    if (r2 < 0 || r2 >= nr || e2 < 0 || e2 >= ne)
                    {
                        continue;
                    }
                    int a2 = ai + da;
                    if (a2 < 0)
                    {
                        a2 += na;
                    }
                    else if (a2 >= na)
                    {
                        a2 -= na;
                    }
                    const auto it = vox.find(nr * (na * e2 + a2) + r2);
                    if (it != vox.end())
                    {
                        const std::uint32_t a = find(v), b = find(it->second);
                        if (a != b)
                        {
                            parent[std::max(a, b)] = std::min(a, b);
                        }
                    }
                }
            }
        }
    }
    // component id = rank of its minimum point index; then small-cluster removal
    std::vector<std::uint32_t> root_min(nv, 0xffffffffU), root_cnt(nv, 0);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const std::uint32_t r = find(vid[i]);
        root_min[r] = std::min(root_min[r], i);
        ++root_cnt[r];
    }
    std::vector<std::int32_t> new_label(nv, -1);
    std::int32_t next = 0;
    for (std::uint32_t i = 0; i < n; ++i) // increasing first-point order == reference label order
    {
        const std::uint32_t r = find(vid[i]);
        if (root_min[r] == i && root_cnt[r] >= min_cluster_size)
        {
            new_label[r] = next++;
        }
    }
    for (std::uint32_t i = 0; i < n; ++i)
    {
        labels[i] = new_label[find(vid[i])];
    }
    return next;
}

// =================================================================== oriented bounding boxes
// Restated from polygonizer.cpp:93-163 (Shamos antipodal pairs), :165-278 (rotating calipers) and
// :280-362 (PCA box; the SVD is the 2 x 2 two-sided Jacobi iteration of Eigen 3.4 restated -
// Eigen itself is absent from this image, so the PCA box is "parity unpinned").
namespace
{
struct PXY
{
    double x, y;
};

inline double tri_area(const PXY& p1, const PXY& p2, const PXY& p3)
{
    // polygonizer.hpp:226-231
    return std::fabs((p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y)) * 0.5);
}

inline float atan2_approx_f(float y, float x)
{
    // common.hpp:33-62
    const float ax = std::fabs(x), ay = std::fabs(y);
    const float mx = std::max(ay, ax), mn = std::min(ay, ax);
    const float a = mn / mx;
    const float s = a * a, c = s * a, q = s * s;
    float r = 0.024840285F * q + 0.18681418F;
    const float t = -0.094097948F * q - 0.33213072F;
    r = r * s + t;
    r = r * c + a;
    if (ay > ax)
    {
        r = 1.57079637F - r;
    }
    if (x < 0)
    {
        r = 3.14159274F - r;
    }
    if (y < 0)
    {
        r = -r;
    }
    return r;
}

void antipodal_pairs(const PXY* h, std::int32_t n, std::vector<std::pair<std::int32_t, std::int32_t>>& out)
{
    out.clear();
    if (n < 2)
    {
        return;
    }
    const std::int32_t i0 = n - 1;
    std::int32_t i = 0, j = 1;
    auto nx = [n](std::int32_t k) { return (k + 1 == n) ? 0 : (k + 1); };
    while (tri_area(h[i], h[nx(i)], h[nx(j)]) > tri_area(h[i], h[nx(i)], h[j]))
    {
        j = nx(j);
    }
    const std::int32_t j0 = j;
    while (i != j0)
    {
        i = nx(i);
        out.emplace_back(i, j);
        while (tri_area(h[i], h[nx(i)], h[nx(j)]) > tri_area(h[i], h[nx(i)], h[j]))
        {
            j = nx(j);
            if (!(i == j0 && j == i0))
            {
                out.emplace_back(i, j);
            }
            else
            {
                return;
            }
        }
        if (tri_area(h[j], h[nx(i)], h[nx(j)]) == tri_area(h[i], h[nx(i)], h[j]))
        {
            if (!(i == j0 && j == i0))
            {
                out.emplace_back(i, nx(j));
            }
            else
            {
                out.emplace_back(nx(i), j);
            }
        }
    }
}
} // namespace

// pairs_out needs 2 * (3 * n + 4) entries; returns the number of pairs
std::int32_t port_antipodal_pairs(const double* hull_xy, std::int32_t n, std::int32_t* pairs_out)
{
    std::vector<std::pair<std::int32_t, std::int32_t>> pr;
    antipodal_pairs(reinterpret_cast<const PXY*>(hull_xy), n, pr);
    for (std::size_t k = 0; k < pr.size(); ++k)
    {
        pairs_out[2 * k] = pr[k].first;
        pairs_out[2 * k + 1] = pr[k].second;
    }
    return static_cast<std::int32_t>(pr.size());
}

// out[11] = 4 corners (x, y), area (float), angle_rad (float), is_valid; method 0 = rotating
// calipers, 1 = PCA. An invalid box leaves the rest zero.
void port_bounding_box(const double* hull_xy, std::int32_t n, std::int32_t method, double* out)
{
    std::fill(out, out + 11, 0.0);
    if (n < 3)
    {
        return;
    }
    const PXY* h = reinterpret_cast<const PXY*>(hull_xy);
    if (method == 0)
    {
        std::vector<std::pair<std::int32_t, std::int32_t>> pr;
        antipodal_pairs(h, n, pr);
        PXY c = {0.0, 0.0};
        for (std::int32_t k = 0; k < n; ++k)
        {
            c.x += h[k].x;
            c.y += h[k].y;
        }
        c.x /= n;
        c.y /= n;
        float best = static_cast<float>(std::numeric_limits<double>::max()); // BoundingBox::area is a float
        bool valid = false;
        for (const auto& [pi, pj] : pr)
        {
            const std::int32_t two[2] = {pi, pj};
            for (const std::int32_t index : two)
            {
                for (const std::int32_t offset : {-1, 1})
                {
                    std::int32_t nb = index + offset;
                    nb = nb < 0 ? nb + n : (nb >= n ? nb - n : nb);
                    const double ex = h[nb].x - h[index].x, ey = h[nb].y - h[index].y;
                    const double len = std::sqrt(ex * ex + ey * ey);
                    if (len < 1.0e-6)
                    {
                        continue;
                    }
                    const double ux = ex / len, uy = ey / len;
                    double min_x = std::numeric_limits<double>::max(), max_x = std::numeric_limits<double>::lowest();
                    double min_y = min_x, max_y = max_x;
                    for (std::int32_t k = 0; k < n; ++k)
                    {
                        const double tx = h[k].x - c.x, ty = h[k].y - c.y;
                        const double rx = tx * ux + ty * uy;
                        const double ry = -tx * uy + ty * ux;
                        min_x = std::min(min_x, rx);
                        max_x = std::max(max_x, rx);
                        min_y = std::min(min_y, ry);
                        max_y = std::max(max_y, ry);
                    }
                    const double area = (max_x - min_x) * (max_y - min_y);
                    if (area < best)
                    {
                        best = static_cast<float>(area);
                        out[0] = min_x * ux - min_y * uy + c.x;
                        out[1] = min_x * uy + min_y * ux + c.y;
                        out[2] = max_x * ux - min_y * uy + c.x;
                        out[3] = max_x * uy + min_y * ux + c.y;
                        out[4] = max_x * ux - max_y * uy + c.x;
                        out[5] = max_x * uy + max_y * ux + c.y;
                        out[6] = min_x * ux - max_y * uy + c.x;
                        out[7] = min_x * uy + max_y * ux + c.y;
                        out[9] = static_cast<double>(atan2_approx_f(static_cast<float>(uy), static_cast<float>(ux)));
                        valid = true;
                    }
                }
            }
        }
        out[8] = valid ? static_cast<double>(best) : 0.0;
        out[10] = valid ? 1.0 : 0.0;
        if (!valid)
        {
            std::fill(out, out + 10, 0.0);
        }
        return;
    }
    // PCA (polygonizer.cpp:280-362)
    double mx = 0.0, my = 0.0;
    for (std::int32_t k = 0; k < n; ++k)
    {
        mx += h[k].x;
        my += h[k].y;
    }
    mx /= n;
    my /= n;
    double cxx = 0.0, cxy = 0.0, cyy = 0.0;
    for (std::int32_t k = 0; k < n; ++k)
    {
        const double dx = h[k].x - mx, dy = h[k].y - my;
        cxx += dx * dx;
        cxy += dx * dy;
        cyy += dy * dy;
    }
    const double dn = static_cast<double>(static_cast<std::uint32_t>(n) - 1u);
    double a00 = cxx / dn, a01 = cxy / dn, a10 = cxy / dn, a11 = cyy / dn;
    // JacobiSVD (right singular vectors), 2 x 2
    double v[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    {
        const double precision = 2.0 * std::numeric_limits<double>::epsilon();
        const double tiny = std::numeric_limits<double>::min();
        double scale = std::max(std::max(std::fabs(a00), std::fabs(a01)), std::max(std::fabs(a10), std::fabs(a11)));
        if (!std::isfinite(scale))
        {
            return; // svd_.info() != Success
        }
        if (scale == 0.0)
        {
            scale = 1.0;
        }
        double w[2][2] = {{a00 / scale, a01 / scale}, {a10 / scale, a11 / scale}};
        double max_diag = std::max(std::fabs(w[0][0]), std::fabs(w[1][1]));
        bool finished = false;
        const int p = 1, q = 0;
        while (!finished)
        {
            finished = true;
            const double threshold = std::max(tiny, precision * max_diag);
            if (std::fabs(w[p][q]) > threshold || std::fabs(w[q][p]) > threshold)
            {
                finished = false;
                double m00 = w[p][p], m01 = w[p][q], m10 = w[q][p], m11 = w[q][q];
                double r1c = 1.0, r1s = 0.0;
                const double t = m00 + m11, d = m10 - m01;
                if (std::fabs(d) >= tiny)
                {
                    const double u = t / d;
                    const double tmp = std::sqrt(1.0 + u * u);
                    r1s = 1.0 / tmp;
                    r1c = u / tmp;
                }
                {
                    const double b0 = r1c * m00 + r1s * m10, b1 = r1c * m01 + r1s * m11;
                    const double c0 = -r1s * m00 + r1c * m10, c1 = -r1s * m01 + r1c * m11;
                    m00 = b0;
                    m01 = b1;
                    m10 = c0;
                    m11 = c1;
                }
                double jc = 1.0, js = 0.0;
                {
                    const double deno = 2.0 * std::fabs(m01);
                    if (!(deno < tiny))
                    {
                        const double tau = (m00 - m11) / deno;
                        const double ww = std::sqrt(tau * tau + 1.0);
                        const double tt = tau > 0.0 ? 1.0 / (tau + ww) : 1.0 / (tau - ww);
                        const double sign_t = tt > 0.0 ? 1.0 : -1.0;
                        const double nn = 1.0 / std::sqrt(tt * tt + 1.0);
                        js = -sign_t * (m01 / std::fabs(m01)) * std::fabs(tt) * nn;
                        jc = nn;
                    }
                }
                const double tc = jc, ts = -js;
                const double lc = r1c * tc - r1s * ts;
                const double ls = r1c * ts + r1s * tc;
                {
                    const double b0 = lc * w[p][0] + ls * w[q][0], b1 = lc * w[p][1] + ls * w[q][1];
                    const double c0 = -ls * w[p][0] + lc * w[q][0], c1 = -ls * w[p][1] + lc * w[q][1];
                    w[p][0] = b0;
                    w[p][1] = b1;
                    w[q][0] = c0;
                    w[q][1] = c1;
                }
                for (int i = 0; i < 2; ++i)
                {
                    const double xp = w[i][p], xq = w[i][q];
                    w[i][p] = jc * xp - js * xq;
                    w[i][q] = js * xp + jc * xq;
                    const double vp = v[i][p], vq = v[i][q];
                    v[i][p] = jc * vp - js * vq;
                    v[i][q] = js * vp + jc * vq;
                }
                max_diag = std::max(max_diag, std::max(std::fabs(w[p][p]), std::fabs(w[q][q])));
            }
        }
        if (std::fabs(w[1][1]) > std::fabs(w[0][0]))
        {
            std::swap(v[0][0], v[0][1]);
            std::swap(v[1][0], v[1][1]);
        }
    }
    double min_x = std::numeric_limits<double>::max(), max_x = std::numeric_limits<double>::lowest();
    double min_y = min_x, max_y = max_x;
    for (std::int32_t k = 0; k < n; ++k)
    {
        const double dx = h[k].x - mx, dy = h[k].y - my;
        const double rx = dx * v[0][0] + dy * v[1][0];
        const double ry = dx * v[0][1] + dy * v[1][1];
        min_x = std::min(min_x, rx);
        max_x = std::max(max_x, rx);
        min_y = std::min(min_y, ry);
        max_y = std::max(max_y, ry);
    }
    const double cx[4] = {min_x, max_x, max_x, min_x}, cy[4] = {min_y, min_y, max_y, max_y};
    for (int k = 0; k < 4; ++k)
    {
        out[2 * k] = (cx[k] * v[0][0] + cy[k] * v[0][1]) + mx;
        out[2 * k + 1] = (cx[k] * v[1][0] + cy[k] * v[1][1]) + my;
    }
    out[8] = static_cast<double>(static_cast<float>((max_x - min_x) * (max_y - min_y)));
    out[9] = static_cast<double>(static_cast<float>(std::atan2(v[1][0], v[0][0])));
    out[10] = 1.0;
}

// =================================================================== convex hull
// polygonizer.cpp:33-91 on PointXY{double x, y}. Returns the vertex count; idx receives local
// indices, counter-clockwise from the lexicographic minimum, collinear points dropped.
std::int32_t port_convex_hull(const double* xy, std::int32_t n, std::int32_t* idx)
{
    if (n < 3)
    {
        for (int i = 0; i < n; ++i)
        {
            idx[i] = i;
        }
        return n;
    }
    std::vector<std::int32_t> s(n);
    std::iota(s.begin(), s.end(), 0);
    std::sort(s.begin(), s.end(), [xy](std::int32_t i, std::int32_t j) {
        return xy[2 * i] < xy[2 * j] ||
               (std::fabs(xy[2 * i] - xy[2 * j]) < std::numeric_limits<double>::epsilon() &&
                xy[2 * i + 1] < xy[2 * j + 1]);
    });
    auto not_left = [xy](std::int32_t a, std::int32_t b, std::int32_t c) {
        return (xy[2 * b] - xy[2 * a]) * (xy[2 * c + 1] - xy[2 * a + 1]) -
                   (xy[2 * b + 1] - xy[2 * a + 1]) * (xy[2 * c] - xy[2 * a]) <=
               0;
    };
    std::vector<std::int32_t> st(2 * static_cast<std::size_t>(n));
    std::int32_t k = 0;
    for (std::int32_t i = 0; i < n; ++i)
    {
        while (k > 1 && not_left(st[k - 2], st[k - 1], s[i]))
        {
            --k;
        }
        st[k++] = s[i];
    }
    for (std::int32_t i = n - 2, t = k + 1; i >= 0; --i)
    {
        while (k >= t && not_left(st[k - 2], st[k - 1], s[i]))
        {
            --k;
        }
        st[k++] = s[i];
    }
    std::copy(st.begin(), st.begin() + (k - 1), idx);
    return k - 1;
}

// processor.cpp:627-658 + polygonizer: per-cluster gather (obstacle-cloud order) and hull.
// hull_offsets: [K+1]; hull_xy: vertex coordinates as doubles (2 per vertex), capacity 2*n;
// hull_index: index into the input cloud per vertex; zminmax: [K][2].
std::int32_t port_cluster_hulls(const float* xyz,
                                std::int32_t stride_f,
                                std::uint32_t n,
                                const std::int32_t* labels,
                                std::int32_t num_clusters,
                                std::uint32_t* hull_offsets,
                                double* hull_xy,
                                std::int32_t* hull_index,
                                double* zminmax)
{
    std::vector<std::vector<std::uint32_t>> members(num_clusters);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        if (labels[i] >= 0 && labels[i] < num_clusters)
        {
            members[labels[i]].push_back(i);
        }
    }
    std::uint32_t off = 0;
    std::vector<double> pts;
    std::vector<std::int32_t> idx;
    for (std::int32_t c = 0; c < num_clusters; ++c)
    {
        hull_offsets[c] = off;
        const auto& m = members[c];
        pts.resize(2 * m.size());
        idx.resize(std::max<std::size_t>(m.size(), 1));
        double zmin = std::numeric_limits<double>::max();
        double zmax = std::numeric_limits<double>::lowest();
        for (std::size_t j = 0; j < m.size(); ++j)
        {
            const float* p = xyz + static_cast<std::size_t>(m[j]) * stride_f;
            pts[2 * j] = p[0];
            pts[2 * j + 1] = p[1];
            zmin = std::min(zmin, static_cast<double>(p[2]));
            zmax = std::max(zmax, static_cast<double>(p[2]));
        }
        const std::int32_t k = port_convex_hull(pts.data(), static_cast<std::int32_t>(m.size()), idx.data());
        for (std::int32_t j = 0; j < k; ++j)
        {
            hull_xy[2 * (off + j)] = pts[2 * idx[j]];
            hull_xy[2 * (off + j) + 1] = pts[2 * idx[j] + 1];
            hull_index[off + j] = static_cast<std::int32_t>(m[idx[j]]);
        }
        off += static_cast<std::uint32_t>(k);
        if (zminmax != nullptr)
        {
            zminmax[2 * c] = zmin;
            zminmax[2 * c + 1] = zmax;
        }
    }
    hull_offsets[num_clusters] = off;
    return static_cast<std::int32_t>(off);
}

// RNG self-check hook: first `count` indices drawn exactly as segmenter.cpp:369-386 would for
// a candidate list of size n (used by tests against std::mt19937 + uniform_int_distribution).
void port_rng_draws(std::uint32_t n, std::uint32_t count, std::uint32_t* out)
{
    Mt19937 gen(42);
    for (std::uint32_t i = 0; i < count; ++i)
    {
        out[i] = uniform_below(gen, n);
    }
}

void port_std_rng_draws(std::uint32_t n, std::uint32_t count, std::uint32_t* out);
} // extern "C"

#include <random>
extern "C" void port_std_rng_draws(std::uint32_t n, std::uint32_t count, std::uint32_t* out)
{
    std::mt19937 gen{42};
    std::uniform_int_distribution<std::uint32_t> dist{0, n - 1U};
    for (std::uint32_t i = 0; i < count; ++i)
    {
        out[i] = dist(gen);
    }
}
