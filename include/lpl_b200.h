/* lpl_b200.h — C ABI of the B200-native LiDAR perception hot path.
 *
 * Drop-in boundary for the reference's `lidar_processing_lib` (a C++ class API, see
 * the headers under include/lidar_processing_lib/ in this repo for the header-only C++ adaptors that keep the
 * reference's class / enum / struct names). Every entry point takes plain pointers and sizes;
 * host pointers unless a name ends in `_device`. All functions return 0 on success or a negative
 * lpl_status; lpl_last_error() gives the text. A context owns one CUDA stream, all device scratch
 * and pinned staging; like the reference objects it is stateful and not thread-safe (one context
 * per calling thread / GPU).
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   lpl_ring_partition      Dataloader::addRingInfo            src/dataloader/src/dataloader.cpp:68-137
 *   lpl_dror_config/filter  NoiseRemover::config / filter      lidar_processing_lib/include/lidar_processing_lib/noise_remover.hpp:56-80
 *   lpl_segmenter_config    Segmenter::config                  lidar_processing_lib/include/lidar_processing_lib/segmenter.hpp:156
 *   lpl_segment             Segmenter::segment / image         .../segmenter.hpp:163-169
 *   lpl_cluster_config      Clusterer::config                  .../clusterer.hpp:79-86
 *   lpl_cluster             Clusterer::cluster                 .../clusterer.hpp:93-94
 *   lpl_convex_hull         Polygonizer::convexHull            .../polygonizer.hpp:105-106
 *   lpl_cluster_hulls       per-label gather + convexHull      src/processor/src/processor.cpp:627-663
 *   lpl_bounding_boxes      Polygonizer::boundingBoxRotatingCalipers / boundingBoxPrincipalComponentAnalysis
 *                           (+ findAntipodalPairsOfConvexHull)  .../polygonizer.hpp:108-123, src/polygonizer.cpp:93-362
 *   lpl_pipeline_*          Processor::run (segment -> split -> cluster -> hulls), batched
 *                                                              src/processor/src/processor.cpp:552-663
 *   lpl_pipeline_upload_cloud2  convert<PointT>(PointCloud2)   src/processor/src/processor.cpp:42-179
 *   lpl_pipeline_upload_packed_xyz  the std::array<float,3> cloud of NoiseRemover::filter   .../noise_remover.hpp:68
 *   lpl_pcd_read            pcl::io::loadPCDFile<PointXYZI>    src/dataloader/src/dataloader.cpp:165
 *   lpl_vehicle_match       vehicle shape matching             src/processor/src/processor.cpp:680-757, processor.hpp:60-192
 *   lpl_knn_build / k_nearest / radius_search   KDTree<float,3>::rebuild / k_nearest / radius_search(_k_nearest)   .../kdtree.hpp:66-77,216-400
 *   lpl_pipeline_split_clouds  label split, clustered cloud, marker lines   src/processor/src/processor.cpp:562-579,627-647,206-343
 */
#ifndef LPL_B200_H
#define LPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpl_ctx lpl_ctx;

typedef enum lpl_status
{
    LPL_OK = 0,
    LPL_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the C++ adaptors            */
    LPL_ERR_CUDA = -2,             /* std::runtime_error                                   */
    LPL_ERR_CAPACITY = -3,         /* std::overflow_error (more points / voxels than reserved) */
    LPL_ERR_NO_DEVICE = -4         /* no CUDA device: there is no CPU fallback             */
} lpl_status;

/* POD mirror of lidar_processing_lib::SegmenterConfiguration (segmenter.hpp:87-112). */
typedef struct lpl_segmenter_cfg
{
    float elevation_up_deg;
    float elevation_down_deg;
    int32_t image_width;
    int32_t image_height;
    int32_t assume_unorganized_cloud;
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
} lpl_segmenter_cfg;

/* POD mirror of NoiseRemoverConfiguration (noise_remover.hpp:41-54). */
typedef struct lpl_dror_cfg
{
    float radius_multiplier_m_per_m;
    float min_search_radius_m;
    uint32_t min_neighbours;
} lpl_dror_cfg;

/* POD mirror of ClustererConfiguration (clusterer.hpp:61-68). */
typedef struct lpl_cluster_cfg
{
    float voxel_grid_range_resolution_m;
    float voxel_grid_azimuth_resolution_deg;
    float voxel_grid_elevation_resolution_deg;
    uint32_t min_cluster_size;
} lpl_cluster_cfg;

/* JCP border behaviour (hazard H2). */
enum
{
    LPL_JCP_AS_REFERENCE = 0, /* reproduce the reference's stale out-of-image kernel slots (default) */
    LPL_JCP_CLEAN = 1         /* out-of-image slots contribute nothing */
};

/* ---- lifetime ------------------------------------------------------------------------- */
/* max_points: capacity per frame (rounded up to a multiple of 2048); max_frames: frames per batch.
 * image_height/width fix the range-image size the scratch is allocated for (64 x 2048 when 0). */
int lpl_create(lpl_ctx** out, int device, uint32_t max_points, uint32_t max_frames,
               int32_t image_height, int32_t image_width);
void lpl_destroy(lpl_ctx* ctx);
const char* lpl_last_error(const lpl_ctx* ctx);
/* Device memory the context holds (one slab: every plane of every frame of the batch), in bytes. About 0.59 KB per
 * point of capacity plus ~0.1 MB per frame of range-image, polar-grid and DROR-grid planes; stage-local scratch
 * planes of equal element size share storage. */
size_t lpl_device_bytes(const lpl_ctx* ctx);
const char* lpl_version(void);

/* ---- configuration (defaults = the reference's struct defaults) ------------------------ */
void lpl_segmenter_default_cfg(lpl_segmenter_cfg* cfg);
void lpl_dror_default_cfg(lpl_dror_cfg* cfg);
void lpl_cluster_default_cfg(lpl_cluster_cfg* cfg);
int lpl_segmenter_config(lpl_ctx* ctx, const lpl_segmenter_cfg* cfg);
int lpl_dror_config(lpl_ctx* ctx, const lpl_dror_cfg* cfg);
int lpl_cluster_config(lpl_ctx* ctx, const lpl_cluster_cfg* cfg);
int lpl_set_jcp_mode(lpl_ctx* ctx, int mode);

/* ---- single-frame, host-pointer entry points (one per reference entry point) ------------ */
/* points: n records `stride` bytes apart whose first 12 bytes are float x, y, z (every PCL point
 * type and std::array<float,3>). */
int lpl_ring_partition(lpl_ctx* ctx, const void* points, size_t stride, uint32_t n, uint16_t* ring_out);

/* labels_out[n]: 0 = VALID, 1 = NOISE (NoiseRemoverLabel). */
int lpl_dror_filter(lpl_ctx* ctx, const void* points, size_t stride, uint32_t n, uint8_t* labels_out);

/* ring_offset: byte offset of the uint16 ring field inside a record, or -1 for ring-less point
 * types (height index from the elevation angle). labels_out[n]: 0 UNKNOWN, 1 GROUND, 2 OBSTACLE
 * (Label). bgr_image_out: nullable, image_height * image_width * 3 bytes (Segmenter::image()). */
int lpl_segment(lpl_ctx* ctx, const void* points, size_t stride, int32_t ring_offset, uint32_t n,
                uint32_t* labels_out, uint8_t* bgr_image_out);

/* labels_out[n]: cluster id or -1 (ClusterLabel); num_clusters_out nullable. */
int lpl_cluster(lpl_ctx* ctx, const void* points, size_t stride, uint32_t n, int32_t* labels_out,
                uint32_t* num_clusters_out);

/* Polygonizer::convexHull on n (x, y) doubles `stride` bytes apart; indices_out needs n entries. */
int lpl_convex_hull(lpl_ctx* ctx, const void* xy, size_t stride, uint32_t n, int32_t* indices_out,
                    uint32_t* count_out);

/* Gather every cluster (labels 0..num_clusters-1, obstacle-cloud order) and build its hull.
 * hull_offsets[num_clusters + 1]; hull_indices / hull_xy sized for n vertices (indices into the
 * input cloud; xy as float pairs); zminmax[num_clusters][2] nullable. */
int lpl_cluster_hulls(lpl_ctx* ctx, const void* points, size_t stride, const int32_t* labels, uint32_t n,
                      uint32_t num_clusters, uint32_t* hull_offsets, int32_t* hull_indices,
                      float* hull_xy, float* zminmax);

/* BoundingBox (polygonizer.hpp:75-81) with the same field meaning; is_valid = 0 leaves the rest zero. */
typedef struct lpl_bbox
{
    double corners[4][2];
    float area;
    float angle_rad;
    int32_t is_valid;
    int32_t reserved;
} lpl_bbox;

enum
{
    LPL_BOX_ROTATING_CALIPERS = 0, /* Polygonizer::boundingBoxRotatingCalipers (polygonizer.cpp:165-278)           */
    LPL_BOX_PCA = 1                /* Polygonizer::boundingBoxPrincipalComponentAnalysis (polygonizer.cpp:280-362) */
};

/* Oriented bounding boxes of num_hulls convex polygons in one call. xy: (x, y) doubles `stride`
 * bytes apart, all hulls back to back, vertices in the order convexHull returns them;
 * offsets[num_hulls + 1] delimits the hulls; boxes_out[num_hulls]. */
int lpl_bounding_boxes(lpl_ctx* ctx, const void* xy, size_t stride, const uint32_t* offsets, uint32_t num_hulls,
                       int method, lpl_bbox* boxes_out);

/* Vehicle shape matching of the node (src/processor/src/processor.cpp:680-757, tables of processor.hpp:60-192; kept
 * behind `perform_polygon_simplification = false` there): per cluster, from its hull (vertices as (x, y) doubles
 * `stride` bytes apart, hulls back to back, offsets[num_hulls + 1]), its z extent z_min_max[k][2], its point count
 * and its oriented box (lpl_bounding_boxes of whatever points the caller boxes - the node passes all cluster
 * points, processor.cpp:704). class_out[k] = 0 compact, 1 sedan, 2 SUV, 3 truck, 4 minivan, or -1 (the polygon is
 * kept); polygon_area_out[k] (nullable) = lidar_processing_lib::polygonArea of the hull (polygonizer.hpp:185-198). */
int lpl_vehicle_match(lpl_ctx* ctx, const void* hull_xy, size_t stride, const uint32_t* offsets, uint32_t num_hulls,
                      const double* z_min_max, const uint32_t* cluster_sizes, const lpl_bbox* boxes, int32_t* class_out,
                      double* polygon_area_out);

/* ---- batched, chained pipeline ----------------------------------------------------------- */
enum
{
    LPL_STAGE_RING = 1,     /* ring partition from point order (else: ring supplied or ring-less) */
    LPL_STAGE_DROR = 2,     /* DROR filter; only VALID points enter segmentation */
    LPL_STAGE_SEGMENT = 4,
    LPL_STAGE_CLUSTER = 8,  /* on OBSTACLE points, cloud order */
    LPL_STAGE_HULLS = 16,
    LPL_STAGE_ALL = 31,     /* the reference node's chain (processor.cpp:552-663) */
    LPL_STAGE_BOXES = 32    /* rotating-calipers box per cluster hull (disabled in the node: processor.cpp:676) */
};

/* One frame of input: n points of 4 floats (x, y, z, unused), contiguous. */
typedef struct lpl_frame
{
    const float* xyzw;    /* host (or device for the *_device call) pointer, 16 B per point */
    uint32_t n;
    const uint16_t* ring; /* nullable; ignored when LPL_STAGE_RING is set */
} lpl_frame;

/* Upload a batch into the context's device buffers (async on the context stream). */
int lpl_pipeline_upload(lpl_ctx* ctx, const lpl_frame* frames, uint32_t num_frames);
/* Same, from device-resident frames (device-to-device copies). */
int lpl_pipeline_upload_device(lpl_ctx* ctx, const lpl_frame* frames, uint32_t num_frames);
/* Same as lpl_pipeline_upload for a batch whose frames lie back to back in ONE host buffer (pinned for
 * an asynchronous copy): counts[num_frames] points per frame, 16 bytes per point. One DMA transfer
 * for the whole batch instead of one per frame; a kernel spreads the frames into place. No ring planes
 * (run with LPL_STAGE_RING, or ring-less). */
int lpl_pipeline_upload_packed(lpl_ctx* ctx, const float* xyzw, const uint32_t* counts, uint32_t num_frames);
/* Same with 12 bytes per point: x, y, z floats back to back - the std::vector<std::array<float, 3>> cloud that
 * NoiseRemover::filter takes (noise_remover.hpp:68), the first library call of the chained pipeline. A quarter
 * fewer bytes cross PCIe than with the 16-byte PCL layout. */
int lpl_pipeline_upload_packed_xyz(lpl_ctx* ctx, const float* xyz, const uint32_t* counts, uint32_t num_frames);
/* One frame as a sensor_msgs/PointCloud2 payload (what Processor::convert<PointT> reads,
 * src/processor/src/processor.cpp:42-179): height * width records, point_step bytes apart inside a
 * row, rows row_step bytes apart; x / y / z are float32 at the given byte offsets, ring (uint16) at
 * ring_offset or -1 when the point type has none. Up to 32 bytes per point of capacity. The raw
 * bytes are uploaded as they are and unpacked by a kernel (no host-side conversion loop). */
typedef struct lpl_cloud2_frame
{
    const void* data;
    uint32_t width, height;
    uint32_t point_step, row_step;
    int32_t x_offset, y_offset, z_offset;
    int32_t ring_offset;
} lpl_cloud2_frame;
int lpl_pipeline_upload_cloud2(lpl_ctx* ctx, const lpl_cloud2_frame* frames, uint32_t num_frames);
/* Enqueue the selected stages for the uploaded batch (async). */
int lpl_pipeline_run(lpl_ctx* ctx, uint32_t num_frames, uint32_t stages);
/* From the second run with the same (frame count, stages, configuration) on, lpl_pipeline_run replays the chain
 * as one CUDA graph (default: enabled). One stream working alone gains ~2 % throughput and ~10 % single-frame
 * latency; several contexts rotating on one GPU (stream.py: FramePipeline) interleave better kernel by kernel
 * and switch it off. */
int lpl_pipeline_use_graph(lpl_ctx* ctx, int enable);
/* Sub-batches: lpl_pipeline_run cuts a batch of >= 4 frames into `parts` (1..8) contiguous sub-batches whose chains
 * run on concurrent streams (parallel branches of the captured graph), so that the one-CTA-per-frame kernels of one
 * sub-batch (JCP row sweep, union-find, scans) overlap the per-point kernels of another. Frames are independent
 * (segmenter.cpp:73-85, clusterer.cpp:104-106), the results do not change. Default: environment LPL_SPLIT, else 3
 * (measured on the 154-frame batch: 1 -> 4.48 ms, 2 -> 4.19, 3 -> 4.17, 4 -> 4.19 when introduced, 2 -> 4.07, 3 -> 4.04,
 * 4 -> 4.08 with the final kernels; starting the sub-batches one stage apart instead of together: 4.36 and worse). */
int lpl_pipeline_use_split(lpl_ctx* ctx, uint32_t parts);
/* Wait for the stream; fails if any kernel raised a capacity flag. */
int lpl_pipeline_sync(lpl_ctx* ctx, uint32_t num_frames);
/* Wait for the stream and report the per-frame capacity flags of the last run: status_out[num_frames], 0 = the
 * frame is good; bit 0 JCP queue, 1 RANSAC RNG table, 2 voxel hash, 3 / 4 JCP border rows / sweep. The batch
 * downloads deliver every frame's planes and return LPL_ERR_CAPACITY when any frame is flagged: only the
 * flagged frames' results are to be discarded. */
int lpl_pipeline_status(lpl_ctx* ctx, uint32_t num_frames, uint32_t* status_out);

/* Per-frame results of the last batch (host buffers, any pointer may be NULL to skip).
 * Copies are synchronous with respect to the context stream. */
typedef struct lpl_frame_result
{
    uint8_t* noise;           /* [n] DROR labels                                   */
    uint16_t* ring;           /* [n]                                               */
    uint32_t* labels;         /* [n] segmentation Label per input point            */
    uint32_t* obstacle_index; /* [num_obstacles] input index of each obstacle point */
    int32_t* cluster_labels;  /* [num_obstacles]                                   */
    uint32_t* hull_offsets;   /* [num_clusters + 1]                                */
    uint32_t* hull_indices;   /* [num_hull_vertices] index into the obstacle cloud */
    float* hull_xy;           /* [num_hull_vertices][2]                            */
    float* zminmax;           /* [num_clusters][2]                                 */
    lpl_bbox* boxes;          /* [num_clusters] (only if the batch ran LPL_STAGE_BOXES) */
    uint8_t* bgr;             /* [H*W*3] (only if the batch ran with lpl_pipeline_want_image) */
    /* counts, filled by lpl_pipeline_counts / lpl_pipeline_download */
    uint32_t n, num_valid, num_obstacles, num_clusters, num_hull_vertices;
} lpl_frame_result;

int lpl_pipeline_want_image(lpl_ctx* ctx, int enable);
int lpl_pipeline_counts(lpl_ctx* ctx, uint32_t frame, lpl_frame_result* res);
int lpl_pipeline_download(lpl_ctx* ctx, uint32_t frame, lpl_frame_result* res);

/* Results of the whole batch in one call: host planes are frame-major, frame f of a plane starts
 * `stride` elements after frame f - 1 (stride >= the largest frame and >= max clusters + 1).
 * Only the occupied width of every plane crosses PCIe (one strided copy per plane). Any plane
 * pointer may be NULL; `counts` is required: [5][num_frames] = n, num_valid, num_obstacles,
 * num_clusters, num_hull_vertices. Pinned host memory (lpl_host_alloc) makes the copies async
 * with respect to other contexts' work. */
typedef struct lpl_batch_result
{
    uint32_t* counts;         /* [5][num_frames]                                   */
    size_t stride;            /* elements per frame in every host plane            */
    uint8_t* labels_u8;       /* Label per input point as a byte (0 / 1 / 2)       */
    uint8_t* noise;
    uint16_t* ring;
    uint32_t* obstacle_index;
    int32_t* cluster_labels;
    uint32_t* hull_offsets;
    uint32_t* hull_indices;
    float* hull_xy;           /* 2 floats per element                              */
    float* zminmax;           /* 2 floats per element                              */
    lpl_bbox* boxes;          /* one box per element (LPL_STAGE_BOXES)             */
} lpl_batch_result;
int lpl_pipeline_download_batch(lpl_ctx* ctx, uint32_t num_frames, lpl_batch_result* res);

/* Results of the whole batch as ONE device-to-host transfer of exactly the occupied bytes: a kernel first packs
 * the selected planes back to back on the device (frames of a plane follow each other without padding, planes
 * start on 16-byte boundaries in the order of the LPL_PLANE_* bits). Frame f of a plane starts at
 * offset[plane] + (sum of that plane's count over the frames before f) * element size, where the count is n for
 * LABELS_U8 / NOISE / RING, num_obstacles for OBSTACLE_INDEX / CLUSTER_LABELS, num_clusters + 1 for HULL_OFFSETS,
 * num_hull_vertices for HULL_INDICES / HULL_XY and num_clusters for ZMINMAX / BOXES - all in `counts`. */
enum
{
    LPL_PLANE_LABELS_U8 = 1u << 0,      /* uint8  per input point: Label 0 / 1 / 2          */
    LPL_PLANE_NOISE = 1u << 1,          /* uint8  per input point: NoiseRemoverLabel        */
    LPL_PLANE_RING = 1u << 2,           /* uint16 per input point                           */
    LPL_PLANE_OBSTACLE_INDEX = 1u << 3, /* uint32 per obstacle point (= positions of Label 2, ascending) */
    LPL_PLANE_CLUSTER_LABELS = 1u << 4, /* int32  per obstacle point                        */
    LPL_PLANE_HULL_OFFSETS = 1u << 5,   /* uint32, num_clusters + 1 per frame               */
    LPL_PLANE_HULL_INDICES = 1u << 6,   /* uint32 per hull vertex                           */
    LPL_PLANE_HULL_XY = 1u << 7,        /* 2 floats per hull vertex                         */
    LPL_PLANE_ZMINMAX = 1u << 8,        /* 2 floats per cluster                             */
    LPL_PLANE_BOXES = 1u << 9,          /* lpl_bbox per cluster (LPL_STAGE_BOXES)           */
    LPL_PLANE_COUNT = 10
};
typedef struct lpl_packed_result
{
    uint32_t* counts;     /* [5][num_frames] = n, num_valid, num_obstacles, num_clusters, num_hull_vertices (required) */
    void* buffer;         /* host buffer for the packed planes (pinned: lpl_host_alloc)                         */
    size_t buffer_bytes;  /* its capacity                                                                       */
    uint32_t planes;      /* LPL_PLANE_* bits to pack                                                           */
    size_t offset[LPL_PLANE_COUNT]; /* out: byte offset of every selected plane in `buffer` ((size_t)-1 otherwise) */
    size_t bytes_used;    /* out: bytes transferred                                                             */
} lpl_packed_result;
int lpl_pipeline_download_packed(lpl_ctx* ctx, uint32_t num_frames, lpl_packed_result* res);

/* Processor glue on the device (src/processor/src/processor.cpp): what the node does on the host around the
 * library calls, for the whole last batch in one call.
 *   ground / obstacle / unsegmented   the label split of :562-579 - every input point, in cloud order, as a 32-byte
 *       pcl::PointXYZRGB record (x, y, z, 1.0f | b, g, r, a = 255 | 12 zero bytes) with the node's colours
 *       (124, 252, 0) / (200, 0, 0) / (255, 255, 0); NOISE points of a DROR run carry label UNKNOWN
 *   clustered   the cloud of :627-647 - for every cluster label ascending, the cluster's points in obstacle-cloud
 *       order, one colour per cluster: cluster_colors[frame][k] = {r, g, b}, or (NULL) three std::rand() % 256 draws
 *       per cluster from the C library stream of a never-seeded process, continued from call to call, as the node does
 *   marker_points   the LINE_LIST vertices of convertPolygonPointsToMarker (:254-343) for every hull with >= 3
 *       vertices: 6 * n points of 3 doubles per hull (bottom ring at z_min, top ring at z_max, vertical edges),
 *       hulls in label order; needs a batch that ran LPL_STAGE_HULLS; at most half the point capacity per frame
 * Host planes are frame-major: frame f of a cloud plane starts `stride` records after frame f - 1 (marker vertices:
 * `marker_stride`). counts: [5][num_frames] = ground, obstacle, unsegmented, clustered points, marker vertices.
 * Any plane pointer may be NULL. Uses the hull stage's sort buffers: call it after the results were downloaded. */
typedef struct lpl_split_result
{
    uint32_t* counts;
    size_t stride;
    void* ground;
    void* obstacle;
    void* unsegmented;
    void* clustered;
    size_t marker_stride;
    double* marker_points;
    const uint8_t* cluster_colors; /* nullable; [num_frames][colors_stride][3] */
    size_t colors_stride;
} lpl_split_result;
int lpl_pipeline_split_clouds(lpl_ctx* ctx, uint32_t num_frames, lpl_split_result* res);
/* First `count` outputs of the C library's rand() after srand(seed) (glibc TYPE_3 generator), host-only: the stream
 * lpl_pipeline_split_clouds colours clusters with. */
void lpl_glibc_rand_stream(uint32_t seed, uint32_t count, int32_t* out);

/* General nearest-neighbour queries: the public KDTree<float, 3> API of the reference
 * (lidar_processing_lib/include/lidar_processing_lib/kdtree.hpp:66-77: rebuild, k_nearest, radius_search,
 * radius_search_k_nearest), BATCHED over queries and exact (an exhaustive tiled scan on the device; the node's hot
 * path does not use it - DROR has its own grid search). Distances are the reference's squared float distances
 * (kdtree.hpp:131-143); results are ordered by ascending (distance, point index) for k_nearest and by ascending point
 * index for radius_search (the reference leaves ties / radius order to its tree traversal).
 *   lpl_knn_build        KDTree::rebuild: uploads the searched set (stride >= 12 bytes, float x, y, z first)
 *   lpl_knn_k_nearest    KDTree::k_nearest for m queries, k <= 128; radius_sqr nullable [m]: with it, only points with
 *                        dist^2 <= radius_sqr[q] count (the k nearest of KDTree::radius_search_k_nearest's candidates).
 *                        idx_out / dist_out [m][k], count_out [m] = neighbours written per query
 *   lpl_knn_radius_search KDTree::radius_search: count_out[q] = points within radius_sqr[q] (may exceed
 *                        max_per_query); the first max_per_query of them by point index are written */
int lpl_knn_build(lpl_ctx* ctx, const void* points, size_t stride, uint32_t n);
/* Identifies the point set resident on the device: changes with every lpl_knn_build, 0 once another call of the
 * context (a segment / filter / pipeline upload) has overwritten it - the caller then builds again. */
unsigned long long lpl_knn_token(const lpl_ctx* ctx);
int lpl_knn_k_nearest(lpl_ctx* ctx, const void* queries, size_t stride, uint32_t m, uint32_t k, const float* radius_sqr,
                      uint32_t* idx_out, float* dist_out, uint32_t* count_out);
int lpl_knn_radius_search(lpl_ctx* ctx, const void* queries, size_t stride, uint32_t m, const float* radius_sqr,
                          uint32_t max_per_query, uint32_t* idx_out, float* dist_out, uint32_t* count_out);

/* PCD v0.7 reader for the reference's data set (FIELDS x y z [intensity], float32, DATA binary or
 * ascii): fills xyzi_out[n][4] (intensity 0 when absent) and *n_out; xyzi_out == NULL only queries
 * the point count. Host-only, needs no context. */
int lpl_pcd_read(const char* path, float* xyzi_out, uint32_t capacity, uint32_t* n_out);

/* Pinned host memory for frame / result buffers (cudaMallocHost / cudaFreeHost). */
int lpl_host_alloc(void** out, size_t bytes);
void lpl_host_free(void* p);

/* ---- measurement / debugging ------------------------------------------------------------- */
/* CUDA-event bracket on the context stream. */
int lpl_timer_start(lpl_ctx* ctx);
int lpl_timer_stop_ms(lpl_ctx* ctx, float* ms_out);
/* Per-kernel CUDA-event profile of the last lpl_pipeline_run on the context stream: entry i is
 * the time between the events dropped behind kernel i - 1 and kernel i (names are static strings;
 * a kernel launched twice appears twice). */
int lpl_profile_enable(lpl_ctx* ctx, int enable);
int lpl_profile_read(lpl_ctx* ctx, uint32_t max_entries, const char** names_out, float* ms_out,
                     uint32_t* count_out);
/* Kernels launched by this context since the last call with reset != 0. */
uint64_t lpl_launch_count(lpl_ctx* ctx, int reset);
/* Intermediates of the last segmentation of `frame` (all nullable): elevation[slices*rings],
 * plane[4] + best inlier count, counters[8] = {binned, candidates, queued, jcp_rounds (chunks of queue
 * entries the row-synchronous JCP sweep processed), border_rows, slices, rings, status}. */
int lpl_debug_segment(lpl_ctx* ctx, uint32_t frame, float* elevation, float* plane,
                      uint32_t* best_inliers, uint32_t* counters);
/* Points the DROR scan-line pass left to the exhaustive grid search in the last run. */
int lpl_debug_dror(lpl_ctx* ctx, uint32_t frame, uint32_t* n_unresolved);
/* Intermediates of the last clustering: dims[3] = num_range, num_azimuth, num_elevation. */
int lpl_debug_cluster(lpl_ctx* ctx, uint32_t frame, int32_t* dims);
/* Hull / cluster stage counters of the last run: counters[2] = {points that survived the octagon
 * filter and entered the hull sort, occupied voxels}. */
int lpl_debug_hulls(lpl_ctx* ctx, uint32_t frame, uint32_t* counters);
/* cudaStream_t of the context (as void*), for callers that order their own work after ours. */
void* lpl_stream(lpl_ctx* ctx);

#ifdef __cplusplus
}
#endif

#endif /* LPL_B200_H */
