// TEST INFRASTRUCTURE ONLY (oracle/): C wrapper around the UNMODIFIED reference library.
//
// oracle/Makefile compiles, in place and without edits,
//   /root/reference/lidar_processing_lib/src/{segmenter,clusterer,noise_remover,polygonizer}.cpp
// against the stand-in headers in oracle/shim/ (PCL, OpenCV and Eigen are not in this image)
// and links them with this file into oracle/_ref/libref_oracle.so.  Nothing here is shipped
// or used by the product path; tests and bench.py's cpu_baseline leg load it through ctypes
// (oracle/oracle.py) as the checker / the reported CPU baseline.
//
// Reference entry points wrapped (file:line under /root/reference/lidar_processing_lib):
//   Segmenter::config / segment / image     include/lidar_processing_lib/segmenter.hpp:153-169
//   Clusterer::config / cluster             include/lidar_processing_lib/clusterer.hpp:77-94
//   NoiseRemover::config / filter           include/lidar_processing_lib/noise_remover.hpp:56-80
//   KDTree::rebuild / radius_search         include/lidar_processing_lib/kdtree.hpp:161-214,283-337
//   Polygonizer::convexHull / findAntipodalPairsOfConvexHull / boundingBoxRotatingCalipers /
//   boundingBoxPrincipalComponentAnalysis   include/lidar_processing_lib/polygonizer.hpp:105-123
//   (the PCA box runs on oracle/shim/Eigen's JacobiSVD restatement, not on Eigen: unpinned)
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <iostream>
#include <limits>
#include <memory>
#include <optional>
#include <queue>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// Expose private scratch members (elevation map, depth image, voxel grid extents) so that
// stage-wise parity tests can compare intermediates. Access specifiers do not change layout
// under the Itanium ABI, so the unmodified reference translation units stay compatible.
#define private public
#include "clusterer.hpp"
#include "noise_remover.hpp"
#include "polygonizer.hpp"
#include "segmenter.hpp"
#undef private

namespace lpl = lidar_processing_lib;

namespace
{
struct CerrSilencer
{
    std::ostringstream sink; // declared first: it must be alive before the swap below
    std::streambuf* old;
    CerrSilencer() : sink(), old(std::cerr.rdbuf(sink.rdbuf())) {}
    ~CerrSilencer() { std::cerr.rdbuf(old); }
};

thread_local std::string g_last_error;
} // namespace

extern "C"
{
struct ref_seg_cfg
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

const char* ref_last_error() { return g_last_error.c_str(); }

// ------------------------------------------------------------------ Segmenter
void* ref_segmenter_create()
{
    CerrSilencer s;
    return new lpl::Segmenter();
}

void ref_segmenter_destroy(void* h) { delete static_cast<lpl::Segmenter*>(h); }

void ref_segmenter_config(void* h, const ref_seg_cfg* c)
{
    CerrSilencer s;
    lpl::SegmenterConfiguration cfg;
    cfg.elevation_up_deg = c->elevation_up_deg;
    cfg.elevation_down_deg = c->elevation_down_deg;
    cfg.image_width = c->image_width;
    cfg.image_height = c->image_height;
    cfg.assume_unorganized_cloud = c->assume_unorganized_cloud != 0;
    cfg.grid_radial_spacing_m = c->grid_radial_spacing_m;
    cfg.grid_slice_resolution_deg = c->grid_slice_resolution_deg;
    cfg.ground_height_threshold_m = c->ground_height_threshold_m;
    cfg.road_maximum_slope_m_per_m = c->road_maximum_slope_m_per_m;
    cfg.min_distance_m = c->min_distance_m;
    cfg.max_distance_m = c->max_distance_m;
    cfg.sensor_height_m = c->sensor_height_m;
    cfg.kernel_threshold_distance_m = c->kernel_threshold_distance_m;
    cfg.amplification_factor = c->amplification_factor;
    cfg.z_min_m = c->z_min_m;
    cfg.z_max_m = c->z_max_m;
    cfg.display_recm_with_low_confidence_points = false;
    static_cast<lpl::Segmenter*>(h)->config(cfg);
}

void ref_segmenter_default_config(ref_seg_cfg* c)
{
    const lpl::SegmenterConfiguration d;
    c->elevation_up_deg = d.elevation_up_deg;
    c->elevation_down_deg = d.elevation_down_deg;
    c->image_width = d.image_width;
    c->image_height = d.image_height;
    c->assume_unorganized_cloud = d.assume_unorganized_cloud ? 1 : 0;
    c->grid_radial_spacing_m = d.grid_radial_spacing_m;
    c->grid_slice_resolution_deg = d.grid_slice_resolution_deg;
    c->ground_height_threshold_m = d.ground_height_threshold_m;
    c->road_maximum_slope_m_per_m = d.road_maximum_slope_m_per_m;
    c->min_distance_m = d.min_distance_m;
    c->max_distance_m = d.max_distance_m;
    c->sensor_height_m = d.sensor_height_m;
    c->kernel_threshold_distance_m = d.kernel_threshold_distance_m;
    c->amplification_factor = d.amplification_factor;
    c->z_min_m = d.z_min_m;
    c->z_max_m = d.z_max_m;
}

// xyz: n points, `stride_f` floats apart. ring == nullptr -> pcl::PointXYZ (ring-less path),
// else pcl::PointXYZIR. labels: n x uint32 (Label). bgr (nullable): H*W*3 bytes.
int ref_segment(void* h,
                const float* xyz,
                std::int32_t stride_f,
                const std::uint16_t* ring,
                std::uint32_t n,
                std::uint32_t* labels,
                std::uint8_t* bgr)
{
    auto* seg = static_cast<lpl::Segmenter*>(h);
    try
    {
        std::vector<lpl::Label> out;
        if (ring != nullptr)
        {
            pcl::PointCloud<pcl::PointXYZIR> cloud;
            cloud.points.resize(n);
            for (std::uint32_t i = 0; i < n; ++i)
            {
                auto& p = cloud.points[i];
                p.x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
                p.y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
                p.z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
                p.intensity = 0.F;
                p.ring = ring[i];
            }
            seg->segment(cloud, out);
        }
        else
        {
            pcl::PointCloud<pcl::PointXYZ> cloud;
            cloud.points.resize(n);
            for (std::uint32_t i = 0; i < n; ++i)
            {
                auto& p = cloud.points[i];
                p.x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
                p.y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
                p.z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
            }
            seg->segment(cloud, out);
        }
        for (std::uint32_t i = 0; i < n; ++i)
        {
            labels[i] = static_cast<std::uint32_t>(out[i]);
        }
        if (bgr != nullptr)
        {
            const cv::Mat& img = seg->image();
            std::memcpy(bgr, img.data(), static_cast<std::size_t>(img.rows) * img.cols * 3);
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1;
    }
    return 0;
}

// Timed variant used by the CPU baseline: the cloud is converted once, `reps` calls are timed.
double ref_segment_timed(void* h,
                         const float* xyz,
                         std::int32_t stride_f,
                         const std::uint16_t* ring,
                         std::uint32_t n,
                         std::int32_t reps)
{
    auto* seg = static_cast<lpl::Segmenter*>(h);
    pcl::PointCloud<pcl::PointXYZIR> cloud;
    cloud.points.resize(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        auto& p = cloud.points[i];
        p.x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        p.y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        p.z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
        p.intensity = 0.F;
        p.ring = ring[i];
    }
    std::vector<lpl::Label> out;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r)
    {
        seg->segment(cloud, out);
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Scratch state left behind by the last segment() call (private members of the reference).
void ref_segment_grid_dims(void* h, std::int32_t* slices, std::int32_t* rings)
{
    auto* seg = static_cast<lpl::Segmenter*>(h);
    *slices = seg->grid_number_of_azimuth_slices_;
    *rings = seg->grid_number_of_radial_rings_;
}

void ref_segment_intermediates(void* h,
                               float* elevation_map,
                               std::int32_t* cloud_map,
                               float* depth,
                               std::uint32_t* ransac_candidates)
{
    auto* seg = static_cast<lpl::Segmenter*>(h);
    if (elevation_map != nullptr)
    {
        std::copy(seg->elevation_map_.begin(), seg->elevation_map_.end(), elevation_map);
    }
    if (cloud_map != nullptr)
    {
        std::copy(
            seg->cloud_mapping_indices_.begin(), seg->cloud_mapping_indices_.end(), cloud_map);
    }
    if (depth != nullptr)
    {
        std::copy(seg->depth_image_.begin(), seg->depth_image_.end(), depth);
    }
    if (ransac_candidates != nullptr)
    {
        *ransac_candidates = static_cast<std::uint32_t>(seg->ransac_points_.size());
    }
}

// ------------------------------------------------------------------ Clusterer
void* ref_clusterer_create() { return new lpl::Clusterer(); }
void ref_clusterer_destroy(void* h) { delete static_cast<lpl::Clusterer*>(h); }

void ref_clusterer_config(void* h, float range_m, float az_deg, float el_deg, std::uint32_t min_sz)
{
    lpl::ClustererConfiguration c;
    c.voxel_grid_range_resolution_m = range_m;
    c.voxel_grid_azimuth_resolution_deg = az_deg;
    c.voxel_grid_elevation_resolution_deg = el_deg;
    c.min_cluster_size = min_sz;
    static_cast<lpl::Clusterer*>(h)->config(c);
}

// Lifts the reference's fixed scratch reservation (clusterer.cpp:36, 200k voxels) for the
// large synthetic configurations; uses only the reference containers' own reserve().
void ref_clusterer_reserve(void* h, std::uint32_t buckets, std::uint32_t elements)
{
    auto* c = static_cast<lpl::Clusterer*>(h);
    c->voxel_labels_.reserve(buckets, elements);
    c->visited_voxels_.reserve(elements + elements / 2);
    c->voxel_queue_.reserve(elements);
}

int ref_cluster(void* h,
                const float* xyz,
                std::int32_t stride_f,
                std::uint32_t n,
                std::int32_t* labels,
                std::int32_t* grid_dims /* nullable: num_range, num_azimuth, num_elevation */)
{
    auto* c = static_cast<lpl::Clusterer*>(h);
    try
    {
        pcl::PointCloud<pcl::PointXYZ> cloud;
        cloud.points.resize(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            auto& p = cloud.points[i];
            p.x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
            p.y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
            p.z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
        }
        std::vector<lpl::ClusterLabel> out;
        c->cluster(cloud, out);
        std::copy(out.begin(), out.end(), labels);
        if (grid_dims != nullptr)
        {
            grid_dims[0] = c->num_range_;
            grid_dims[1] = c->num_azimuth_;
            grid_dims[2] = c->num_elevation_;
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1;
    }
    return 0;
}

double ref_cluster_timed(
    void* h, const float* xyz, std::int32_t stride_f, std::uint32_t n, std::int32_t reps)
{
    auto* c = static_cast<lpl::Clusterer*>(h);
    pcl::PointCloud<pcl::PointXYZ> cloud;
    cloud.points.resize(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        auto& p = cloud.points[i];
        p.x = xyz[static_cast<std::size_t>(i) * stride_f + 0];
        p.y = xyz[static_cast<std::size_t>(i) * stride_f + 1];
        p.z = xyz[static_cast<std::size_t>(i) * stride_f + 2];
    }
    std::vector<lpl::ClusterLabel> out;
    const auto t0 = std::chrono::steady_clock::now();
    try
    {
        for (int r = 0; r < reps; ++r)
        {
            c->cluster(cloud, out);
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1.0;
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ------------------------------------------------------------------ NoiseRemover (DROR)
// mode 0 ("as_is"): NoiseRemover::filter unmodified. Its early-exit leaves entries on the
//   KD-tree traversal stack (kdtree.hpp:367-370), so later queries re-visit nodes; the result
//   depends on nth_element's tree shape and on query order (SURVEY.md hazard H1).
// mode 1 ("exact"): same KDTree and the same radius / threshold arithmetic
//   (noise_remover.cpp:44-66), but every query runs KDTree::radius_search, whose traversal
//   always drains the stack, and compares the full neighbour count with min_neighbours.
int ref_dror(const float* xyz,
             std::int32_t stride_f,
             std::uint32_t n,
             float radius_multiplier,
             float min_radius,
             std::uint32_t min_neighbours,
             std::int32_t mode,
             std::uint8_t* labels)
{
    try
    {
        std::vector<lpl::NoiseRemover::PointT> pts(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            pts[i] = {xyz[static_cast<std::size_t>(i) * stride_f + 0],
                      xyz[static_cast<std::size_t>(i) * stride_f + 1],
                      xyz[static_cast<std::size_t>(i) * stride_f + 2]};
        }
        lpl::NoiseRemoverConfiguration cfg;
        cfg.radius_multiplier_m_per_m = radius_multiplier;
        cfg.min_search_radius_m = min_radius;
        cfg.min_neighbours = min_neighbours;
        if (mode == 0)
        {
            lpl::NoiseRemover nr;
            nr.reserve(n + 16);
            nr.config(cfg);
            std::vector<lpl::NoiseRemoverLabel> out;
            nr.filter(pts, out);
            for (std::uint32_t i = 0; i < n; ++i)
            {
                labels[i] = static_cast<std::uint8_t>(out[i]);
            }
        }
        else
        {
            lpl::KDTree<float, 3> tree(false);
            tree.reserve(n + 16);
            std::vector<lpl::KDTree<float, 3>::Neighbour> neigh;
            neigh.reserve(n + 16);
            if (n > 0)
            {
                tree.rebuild(pts);
            }
            const double scaling_factor = std::pow(static_cast<double>(radius_multiplier), 2.0);
            const float min_r_sqr = min_radius * min_radius;
            for (std::uint32_t i = 0; i < n; ++i)
            {
                const auto& p = pts[i];
                const double range_sqr = (p[0] * p[0]) + (p[1] * p[1]);
                const float r_sqr =
                    std::max(static_cast<float>(scaling_factor * range_sqr), min_r_sqr);
                tree.radius_search(p, r_sqr, neigh);
                labels[i] = (neigh.size() < min_neighbours) ? 1 : 0;
            }
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1;
    }
    return 0;
}

double ref_dror_timed(const float* xyz, std::int32_t stride_f, std::uint32_t n, std::int32_t reps)
{
    std::vector<lpl::NoiseRemover::PointT> pts(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        pts[i] = {xyz[static_cast<std::size_t>(i) * stride_f + 0],
                  xyz[static_cast<std::size_t>(i) * stride_f + 1],
                  xyz[static_cast<std::size_t>(i) * stride_f + 2]};
    }
    lpl::NoiseRemover nr;
    nr.reserve(n + 16);
    std::vector<lpl::NoiseRemoverLabel> out;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r)
    {
        nr.filter(pts, out);
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- KDTree<float, 3>::radius_search (kdtree.hpp:283-337), query by query. idx_out / dist_out: [m][k] = the first k
// of the distance-sorted result; count_out[q] = all neighbours found.
// (KDTree::k_nearest cannot be wrapped: instantiating it fails to compile - kdtree.hpp:246,251 call
// PriorityQueue::push_back, which priority_queue.hpp does not have; the template is never instantiated by the
// reference. The k-nearest checker is therefore a restatement: tests/test_gpu_parity.py, brute force.)
int ref_kdtree_query(const float* xyz, std::uint32_t n, const float* queries, std::uint32_t m, std::uint32_t k,
                     const float* radius_sqr, std::int32_t mode, std::uint32_t* idx_out, float* dist_out, std::uint32_t* count_out)
{
    try
    {
        using Tree = lpl::KDTree<float, 3>;
        std::vector<Tree::PointT> pts(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            pts[i] = {xyz[3 * static_cast<std::size_t>(i)], xyz[3 * static_cast<std::size_t>(i) + 1], xyz[3 * static_cast<std::size_t>(i) + 2]};
        }
        Tree tree(true); // radius results sorted by distance
        tree.reserve(n + 16);
        if (n > 0)
        {
            tree.rebuild(pts);
        }
        std::vector<Tree::Neighbour> neigh;
        for (std::uint32_t q = 0; q < m; ++q)
        {
            const Tree::PointT t = {queries[3 * static_cast<std::size_t>(q)], queries[3 * static_cast<std::size_t>(q) + 1],
                                    queries[3 * static_cast<std::size_t>(q) + 2]};
            if (mode == 0)
            {
                g_last_error = "KDTree::k_nearest does not compile in the reference (PriorityQueue::push_back)";
                return -2;
            }
            tree.radius_search(t, radius_sqr[q], neigh);
            count_out[q] = static_cast<std::uint32_t>(neigh.size());
            for (std::uint32_t j = 0; j < k && j < neigh.size(); ++j)
            {
                idx_out[static_cast<std::size_t>(q) * k + j] = neigh[j].index;
                dist_out[static_cast<std::size_t>(q) * k + j] = neigh[j].distance;
            }
        }
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return -1;
    }
    return 0;
}

// shim self-check: 5x5 dilation of a single-channel image (compared with cv2.dilate in tests)
void ref_shim_dilate5x5(const std::uint8_t* src, std::int32_t rows, std::int32_t cols, std::uint8_t* dst)
{
    cv::Mat m;
    m.create(rows, cols, CV_8UC1);
    std::memcpy(m.data(), src, static_cast<std::size_t>(rows) * cols);
    cv::Mat k = cv::getStructuringElement(cv::MORPH_RECT, cv::Size(5, 5));
    cv::dilate(m, m, k);
    std::memcpy(dst, m.data(), static_cast<std::size_t>(rows) * cols);
}
// ---- Polygonizer (src/polygonizer.cpp:33-362) ------------------------------------------------
// xy: n interleaved (x, y) doubles. idx_out needs n entries; returns the vertex count.
std::int32_t ref_convex_hull(const double* xy, std::int32_t n, std::int32_t* idx_out)
{
    static thread_local lpl::Polygonizer poly;
    std::vector<lpl::PointXY> pts(static_cast<std::size_t>(n));
    for (std::int32_t i = 0; i < n; ++i)
    {
        pts[i] = {xy[2 * i], xy[2 * i + 1]};
    }
    std::vector<std::int32_t> idx;
    poly.convexHull(pts, idx);
    std::copy(idx.begin(), idx.end(), idx_out);
    return static_cast<std::int32_t>(idx.size());
}

// pairs_out needs 2 * (3 * n + 4) entries; returns the number of pairs.
std::int32_t ref_antipodal_pairs(const double* hull_xy, std::int32_t n, std::int32_t* pairs_out)
{
    static thread_local lpl::Polygonizer poly;
    std::vector<lpl::PointXY> pts(static_cast<std::size_t>(n));
    for (std::int32_t i = 0; i < n; ++i)
    {
        pts[i] = {hull_xy[2 * i], hull_xy[2 * i + 1]};
    }
    std::vector<lpl::AntipodalPair> pairs;
    poly.findAntipodalPairsOfConvexHull(pts, pairs);
    for (std::size_t k = 0; k < pairs.size(); ++k)
    {
        pairs_out[2 * k] = pairs[k].index_1;
        pairs_out[2 * k + 1] = pairs[k].index_2;
    }
    return static_cast<std::int32_t>(pairs.size());
}

// method 0 = boundingBoxRotatingCalipers, 1 = boundingBoxPrincipalComponentAnalysis.
// out[11] = 4 corners (x, y), area, angle_rad, is_valid. An invalid box leaves the rest zero.
void ref_bounding_box(const double* hull_xy, std::int32_t n, std::int32_t method, double* out)
{
    static thread_local lpl::Polygonizer poly;
    std::vector<lpl::PointXY> pts(static_cast<std::size_t>(n));
    for (std::int32_t i = 0; i < n; ++i)
    {
        pts[i] = {hull_xy[2 * i], hull_xy[2 * i + 1]};
    }
    const lpl::BoundingBox b = method == 0 ? poly.boundingBoxRotatingCalipers(pts)
                                           : poly.boundingBoxPrincipalComponentAnalysis(pts);
    std::fill(out, out + 11, 0.0);
    out[10] = b.is_valid ? 1.0 : 0.0;
    if (b.is_valid)
    {
        for (int k = 0; k < 4; ++k)
        {
            out[2 * k] = b.corners[k].x;
            out[2 * k + 1] = b.corners[k].y;
        }
        out[8] = static_cast<double>(b.area);
        out[9] = static_cast<double>(b.angle_rad);
    }
}
} // extern "C"
