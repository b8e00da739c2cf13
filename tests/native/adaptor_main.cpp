// Exercises the drop-in C++ surface (include/lidar_processing_lib/*.hpp) the way the reference's
// processor node does (src/processor/src/processor.cpp:552-663): Segmenter::segment on a
// PointXYZIR cloud, label split, Clusterer::cluster on the obstacle cloud, per-label gather and
// Polygonizer::convexHull, plus NoiseRemover::filter on the raw points.
// Built by tests/test_adaptors.py against the PCL / OpenCV shims under oracle/shim (test
// infrastructure) and linked with liblpl_b200.so.
//
// usage: adaptor_main <in.bin> <out.bin>
//        adaptor_main --time <in.bin> <reps>     per-call latency of the same sequence (JSON on stdout): the node's
//                                                per-cluster convexHull loop and the batched Polygonizer::convexHulls
//   in : u32 n, n x (float x, y, z, w), n x u16 ring
//   out: u32 n, n x u8 noise, n x u32 label, u32 m, m x i32 cluster label, u32 K, K x u32 hull size,
//        sum(hull size) x (double x, y), K x (8 double corners, float area, float yaw, u32 valid) from
//        Polygonizer::boundingBoxRotatingCalipers on every hull (processor.cpp:704 passes its points
//        the same way; dead code in the node)
// exit code 3: no CUDA device (std::runtime_error from the adaptors), 4: any other exception
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include <lidar_processing_lib/clusterer.hpp>
#include <lidar_processing_lib/kdtree.hpp>
#include <lidar_processing_lib/noise_remover.hpp>
#include <lidar_processing_lib/polygonizer.hpp>
#include <lidar_processing_lib/segmenter.hpp>

namespace lpl = lidar_processing_lib;

namespace
{
bool read_frame(const char* path, std::vector<float>& xyzw, std::vector<std::uint16_t>& ring)
{
    std::FILE* fi = std::fopen(path, "rb");
    if (fi == nullptr)
    {
        return false;
    }
    std::uint32_t n = 0;
    bool ok = std::fread(&n, 4, 1, fi) == 1;
    if (ok)
    {
        xyzw.resize(static_cast<std::size_t>(n) * 4);
        ring.resize(n);
        ok = n == 0 || (std::fread(xyzw.data(), 16, n, fi) == n && std::fread(ring.data(), 2, n, fi) == n);
    }
    std::fclose(fi);
    return ok;
}

double median(std::vector<double> v)
{
    std::sort(v.begin(), v.end());
    return v.empty() ? 0.0 : v[v.size() / 2];
}

// the node's sequence (processor.cpp:552-663), timed call by call
int time_sequence(const char* path, int reps)
{
    using clk = std::chrono::steady_clock;
    const auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::vector<float> xyzw;
    std::vector<std::uint16_t> ring;
    if (!read_frame(path, xyzw, ring))
    {
        return 2;
    }
    const auto n = static_cast<std::uint32_t>(ring.size());
    std::vector<lpl::NoiseRemover::PointT> raw(n);
    pcl::PointCloud<pcl::PointXYZIR> cloud;
    cloud.points.resize(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        raw[i] = {xyzw[4 * i], xyzw[4 * i + 1], xyzw[4 * i + 2]};
        auto& p = cloud.points[i];
        p.x = xyzw[4 * i];
        p.y = xyzw[4 * i + 1];
        p.z = xyzw[4 * i + 2];
        p.intensity = 0.5F;
        p.ring = ring[i];
    }
    lpl::NoiseRemover noise_remover;
    noise_remover.reserve(200'000U);
    lpl::Segmenter segmenter;
    lpl::Clusterer clusterer;
    lpl::ClustererConfiguration ccfg;
    ccfg.voxel_grid_elevation_resolution_deg = 3.0F;
    clusterer.config(ccfg);
    lpl::Polygonizer polygonizer;
    std::vector<lpl::NoiseRemoverLabel> noise;
    std::vector<lpl::Label> labels;
    std::vector<lpl::ClusterLabel> clabels;
    pcl::PointCloud<pcl::PointXYZRGB> obstacles;
    std::vector<double> t_dror, t_seg, t_split, t_clu, t_loop, t_batched;
    std::size_t hv_loop = 0, hv_batched = 0;
    std::int32_t max_label = -1;
    for (int r = 0; r < reps + 2; ++r)
    {
        const auto t0 = clk::now();
        noise_remover.filter(raw, noise);
        const auto t1 = clk::now();
        segmenter.segment(cloud, labels);
        const auto t2 = clk::now();
        obstacles.points.clear();
        for (std::uint32_t i = 0; i < n; ++i)
        {
            if (labels[i] == lpl::Label::OBSTACLE)
            {
                pcl::PointXYZRGB q;
                q.x = cloud.points[i].x;
                q.y = cloud.points[i].y;
                q.z = cloud.points[i].z;
                obstacles.points.push_back(q);
            }
        }
        const auto t3 = clk::now();
        clusterer.cluster(obstacles, clabels);
        const auto t4 = clk::now();
        max_label = -1;
        for (const auto l : clabels)
        {
            max_label = l > max_label ? l : max_label;
        }
        // (a) the node's loop: O(K * M) gather + one convexHull call per cluster
        std::vector<lpl::PointXY> pts;
        std::vector<std::int32_t> idx;
        hv_loop = 0;
        for (std::int32_t l = 0; l <= max_label; ++l)
        {
            pts.clear();
            for (std::size_t i = 0; i < clabels.size(); ++i)
            {
                if (clabels[i] == l)
                {
                    pts.push_back({static_cast<double>(obstacles.points[i].x), static_cast<double>(obstacles.points[i].y)});
                }
            }
            polygonizer.convexHull(pts, idx);
            hv_loop += idx.size();
        }
        const auto t5 = clk::now();
        // (b) the batched extension: one call for all clusters (gather + z extent + hulls on the device)
        std::vector<std::uint32_t> off;
        std::vector<std::int32_t> hidx;
        std::vector<lpl::PointXY> hpts;
        std::vector<std::array<double, 2>> zmm;
        polygonizer.convexHulls(obstacles.points, clabels, static_cast<std::uint32_t>(max_label + 1), off, hidx, hpts, zmm);
        hv_batched = hpts.size();
        const auto t6 = clk::now();
        if (r >= 2)
        {
            t_dror.push_back(ms(t0, t1));
            t_seg.push_back(ms(t1, t2));
            t_split.push_back(ms(t2, t3));
            t_clu.push_back(ms(t3, t4));
            t_loop.push_back(ms(t4, t5));
            t_batched.push_back(ms(t5, t6));
        }
    }
    std::printf("{\"n\": %u, \"clusters\": %d, \"hull_vertices_loop\": %zu, \"hull_vertices_batched\": %zu, \"reps\": %d, "
                "\"ms\": {\"noise_filter\": %.3f, \"segment\": %.3f, \"label_split_host\": %.3f, \"cluster\": %.3f, "
                "\"hulls_per_cluster_calls\": %.3f, \"hulls_batched_call\": %.3f}}\n",
                n, max_label + 1, hv_loop, hv_batched, reps, median(t_dror), median(t_seg), median(t_split), median(t_clu),
                median(t_loop), median(t_batched));
    return 0;
}
} // namespace

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        return 2;
    }
    try
    {
        if (std::strcmp(argv[1], "--time") == 0)
        {
            return argc >= 4 ? time_sequence(argv[2], std::atoi(argv[3])) : 2;
        }
        std::FILE* fi = std::fopen(argv[1], "rb");
        if (fi == nullptr)
        {
            return 2;
        }
        std::uint32_t n = 0;
        if (std::fread(&n, 4, 1, fi) != 1)
        {
            return 2;
        }
        std::vector<float> xyzw(static_cast<std::size_t>(n) * 4);
        std::vector<std::uint16_t> ring(n);
        if (n != 0 && (std::fread(xyzw.data(), 16, n, fi) != n || std::fread(ring.data(), 2, n, fi) != n))
        {
            return 2;
        }
        std::fclose(fi);

        // ---- NoiseRemover (noise_remover.hpp:56-80)
        std::vector<lpl::NoiseRemover::PointT> raw(n);
        pcl::PointCloud<pcl::PointXYZIR> cloud;
        cloud.points.resize(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            raw[i] = {xyzw[4 * i], xyzw[4 * i + 1], xyzw[4 * i + 2]};
            auto& p = cloud.points[i];
            p.x = xyzw[4 * i];
            p.y = xyzw[4 * i + 1];
            p.z = xyzw[4 * i + 2];
            p.intensity = 0.5F;
            p.ring = ring[i];
        }
        lpl::NoiseRemover noise_remover;
        noise_remover.reserve(200'000U);
        std::vector<lpl::NoiseRemoverLabel> noise;
        noise_remover.filter(raw, noise);

        // ---- Segmenter (processor.cpp:552-556)
        lpl::Segmenter segmenter;
        lpl::SegmenterConfiguration scfg; // node values == struct defaults (processor.param.yaml:10-30)
        segmenter.config(scfg);
        std::vector<lpl::Label> labels;
        segmenter.segment(cloud, labels);
        const cv::Mat& image = segmenter.image();
        if (image.rows != scfg.image_height || image.cols != scfg.image_width)
        {
            return 5;
        }

        // ---- label split (processor.cpp:562-579), cluster (:599)
        pcl::PointCloud<pcl::PointXYZRGB> obstacles;
        for (std::uint32_t i = 0; i < n; ++i)
        {
            if (labels[i] == lpl::Label::OBSTACLE)
            {
                pcl::PointXYZRGB q;
                q.x = cloud.points[i].x;
                q.y = cloud.points[i].y;
                q.z = cloud.points[i].z;
                obstacles.points.push_back(q);
            }
        }
        lpl::Clusterer clusterer;
        lpl::ClustererConfiguration ccfg;
        ccfg.voxel_grid_elevation_resolution_deg = 3.0F; // processor.param.yaml:31-35
        clusterer.config(ccfg);
        std::vector<lpl::ClusterLabel> clabels;
        clusterer.cluster(obstacles, clabels);

        // ---- per-label gather + convexHull (processor.cpp:627-663)
        std::int32_t max_label = -1;
        for (const auto l : clabels)
        {
            max_label = l > max_label ? l : max_label;
        }
        lpl::Polygonizer polygonizer;
        std::vector<std::uint32_t> hull_sizes;
        std::vector<double> hull_xy;
        std::vector<lpl::BoundingBox> boxes;
        std::vector<lpl::PointXY> hull_pts;
        std::vector<lpl::PointXY> pts;
        std::vector<std::int32_t> idx;
        for (std::int32_t l = 0; l <= max_label; ++l)
        {
            pts.clear();
            for (std::size_t i = 0; i < clabels.size(); ++i)
            {
                if (clabels[i] == l)
                {
                    pts.push_back({static_cast<double>(obstacles.points[i].x), static_cast<double>(obstacles.points[i].y)});
                }
            }
            polygonizer.convexHull(pts, idx);
            hull_sizes.push_back(static_cast<std::uint32_t>(idx.size()));
            hull_pts.clear();
            for (const auto j : idx)
            {
                hull_xy.push_back(pts[j].x);
                hull_xy.push_back(pts[j].y);
                hull_pts.push_back(pts[j]);
            }
            boxes.push_back(polygonizer.boundingBoxRotatingCalipers(hull_pts));
        }

        std::FILE* fo = std::fopen(argv[2], "wb");
        if (fo == nullptr)
        {
            return 2;
        }
        const std::uint32_t m = static_cast<std::uint32_t>(clabels.size());
        const std::uint32_t K = static_cast<std::uint32_t>(hull_sizes.size());
        std::fwrite(&n, 4, 1, fo);
        std::fwrite(noise.data(), 1, n, fo);
        std::fwrite(labels.data(), 4, n, fo);
        std::fwrite(&m, 4, 1, fo);
        std::fwrite(clabels.data(), 4, m, fo);
        std::fwrite(&K, 4, 1, fo);
        std::fwrite(hull_sizes.data(), 4, K, fo);
        std::fwrite(hull_xy.data(), 8, hull_xy.size(), fo);
        for (const auto& b : boxes)
        {
            double c[8];
            for (int k = 0; k < 4; ++k)
            {
                c[2 * k] = b.is_valid ? b.corners[k].x : 0.0;
                c[2 * k + 1] = b.is_valid ? b.corners[k].y : 0.0;
            }
            const float area = b.is_valid ? b.area : 0.0F, yaw = b.is_valid ? b.angle_rad : 0.0F;
            const std::uint32_t valid = b.is_valid ? 1U : 0U;
            std::fwrite(c, 8, 8, fo);
            std::fwrite(&area, 4, 1, fo);
            std::fwrite(&yaw, 4, 1, fo);
            std::fwrite(&valid, 4, 1, fo);
        }
        // ---- KDTree<float, 3> (kdtree.hpp:66-77) on the first 5000 points: 3 nearest of 64 targets, their
        // neighbourhoods within 0.5 m (sorted by distance) and the 2 nearest within that radius. Called AFTER the other
        // adaptors used the thread's shared context: the tree re-uploads its points when they were displaced.
        {
            const std::uint32_t nt = n < 5000U ? n : 5000U, nq = nt < 64U ? nt : 64U;
            std::vector<lpl::KDTree<float, 3>::PointT> tp(raw.begin(), raw.begin() + nt);
            lpl::KDTree<float, 3> tree(true);
            tree.rebuild(tp);
            std::vector<lpl::KDTree<float, 3>::Neighbour> ne;
            std::fwrite(&nq, 4, 1, fo);
            for (std::uint32_t q = 0; q < nq; ++q)
            {
                tree.k_nearest(tp[q * 7U % nt], 3, ne);
                std::uint32_t cnt = static_cast<std::uint32_t>(ne.size());
                std::fwrite(&cnt, 4, 1, fo);
                std::fwrite(ne.data(), sizeof(ne[0]), ne.size(), fo);
                if (q < 4)
                {
                    noise_remover.filter(raw, noise); // displaces the tree's points on the device
                }
                tree.radius_search(tp[q * 7U % nt], 0.25F, ne);
                cnt = static_cast<std::uint32_t>(ne.size());
                std::fwrite(&cnt, 4, 1, fo);
                std::fwrite(ne.data(), sizeof(ne[0]), ne.size(), fo);
                tree.radius_search_k_nearest(tp[q * 7U % nt], 0.25F, 2, ne);
                cnt = static_cast<std::uint32_t>(ne.size());
                std::fwrite(&cnt, 4, 1, fo);
                std::fwrite(ne.data(), sizeof(ne[0]), ne.size(), fo);
            }
        }
        std::fclose(fo);
        std::printf("ok n=%u obstacles=%u clusters=%u hull_vertices=%zu\n", n, m, K, hull_xy.size() / 2);
        return 0;
    }
    catch (const std::runtime_error& e)
    {
        std::cerr << "runtime_error: " << e.what() << "\n";
        return std::strstr(e.what(), "no CUDA device") != nullptr ? 3 : 4;
    }
    catch (const std::exception& e)
    {
        std::cerr << "exception: " << e.what() << "\n";
        return 4;
    }
}
