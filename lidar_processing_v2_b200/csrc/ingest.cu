// Ingest: the two ways a frame reaches the hot path in the reference.
//
//   lpl_pcd_read                 pcl::io::loadPCDFile<pcl::PointXYZI>  src/dataloader/src/dataloader.cpp:165
//                                (PCD v0.7, float32 FIELDS x y z [intensity ...], DATA binary or ascii;
//                                exactly POINTS * record bytes are consumed - the KITTI files carry a few
//                                KB of padding behind the payload)
//   lpl_pipeline_upload_cloud2   Processor::convert<PointT>            src/processor/src/processor.cpp:42-179
//                                (sensor_msgs/PointCloud2 layout: height x width records, point_step /
//                                row_step, float32 x / y / z and an optional uint16 ring at byte offsets).
//                                The raw message bytes cross PCIe once and are unpacked on the device
//                                into the float4 / ring planes every stage reads (the reference converts
//                                with a scalar host loop, one point at a time).
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lpl_b200.h"
#include "common.cuh"

namespace lpl
{
struct Cloud2Desc
{
    std::uint32_t width, height, point_step, row_step;
    std::int32_t x_off, y_off, z_off, ring_off;
};

__device__ __forceinline__ std::uint32_t load_u32_bytes(const unsigned char* p)
{
    return static_cast<std::uint32_t>(p[0]) | (static_cast<std::uint32_t>(p[1]) << 8) |
           (static_cast<std::uint32_t>(p[2]) << 16) | (static_cast<std::uint32_t>(p[3]) << 24);
}

// one thread per point: AoS records (any stride) -> float4 plane (+ ring plane)
__global__ void __launch_bounds__(256)
    k_unpack_cloud2(Dev d, const unsigned char* __restrict__ raw, std::size_t raw_stride, const Cloud2Desc* __restrict__ desc)
{
    const std::uint32_t f = blockIdx.y;
    const Cloud2Desc c = desc[f];
    const std::uint32_t n = c.width * c.height;
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n)
    {
        return;
    }
    const std::uint32_t row = i / c.width, col = i - row * c.width;
    const unsigned char* rec = raw + static_cast<std::size_t>(f) * raw_stride + static_cast<std::size_t>(row) * c.row_step +
                               static_cast<std::size_t>(col) * c.point_step;
    float x, y, z;
    if (((c.point_step | c.row_step | static_cast<std::uint32_t>(c.x_off) | static_cast<std::uint32_t>(c.y_off) |
          static_cast<std::uint32_t>(c.z_off)) & 3u) == 0u)
    {
        x = *reinterpret_cast<const float*>(rec + c.x_off);
        y = *reinterpret_cast<const float*>(rec + c.y_off);
        z = *reinterpret_cast<const float*>(rec + c.z_off);
    }
    else
    {
        x = __uint_as_float(load_u32_bytes(rec + c.x_off));
        y = __uint_as_float(load_u32_bytes(rec + c.y_off));
        z = __uint_as_float(load_u32_bytes(rec + c.z_off));
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap + i;
    d.pts_in[o] = make_float4(x, y, z, 0.f);
    if (c.ring_off >= 0)
    {
        const unsigned char* r = rec + c.ring_off;
        d.ring[o] = static_cast<std::uint16_t>(r[0] | (r[1] << 8));
    }
}

// packed upload: the frames of a batch lie back to back in one buffer (one DMA transfer); this kernel
// spreads them into the frame-major planes. start[f] = first point of frame f in the packed buffer.
__global__ void __launch_bounds__(256)
    k_spread_packed(Dev d, const float4* __restrict__ packed, const std::uint32_t* __restrict__ start)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n)
    {
        d.pts_in[static_cast<std::size_t>(f) * d.cap + i] = packed[static_cast<std::size_t>(start[f]) + i];
    }
}

// same for 12-byte records (x, y, z floats back to back: the std::array<float, 3> cloud NoiseRemover::filter takes,
// noise_remover.hpp:68). Consecutive threads read consecutive records, so every fetched sector is fully used.
__global__ void __launch_bounds__(256)
    k_spread_packed_xyz(Dev d, const float* __restrict__ packed, const std::uint32_t* __restrict__ start)
{
    const std::uint32_t f = blockIdx.y;
    const std::uint32_t n = d.n_in[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n)
    {
        const float* p = packed + (static_cast<std::size_t>(start[f]) + i) * 3u;
        d.pts_in[static_cast<std::size_t>(f) * d.cap + i] = make_float4(p[0], p[1], p[2], 0.f);
    }
}

void launch_spread_packed(Ctx* c, std::uint32_t nf, const void* packed, const std::uint32_t* start, bool xyz12)
{
    const dim3 grid((c->d.cap + 255) / 256, nf);
    if (xyz12)
    {
        k_spread_packed_xyz<<<grid, 256, 0, c->stream>>>(c->d, static_cast<const float*>(packed), start);
    }
    else
    {
        k_spread_packed<<<grid, 256, 0, c->stream>>>(c->d, static_cast<const float4*>(packed), start);
    }
    mark(c, "spread_packed");
}

// ------------------------------------------------------------------------------------------
// packed result download: the occupied part of every selected result plane of a batch is copied
// back to back into one device staging area, so the batch's results cross PCIe as ONE transfer
// of exactly the bytes that carry information (the frame-major planes have a fixed stride of
// `cap` elements; a strided copy moves the widest frame's width for every frame).
// ------------------------------------------------------------------------------------------
// header (PackHeader) written by k_pack_layout at the start of the staging area
__global__ void __launch_bounds__(1024) k_pack_layout(Dev d, std::uint32_t nf, std::uint32_t planes, PackHeader* hdr,
                                                      std::uint32_t* counts, unsigned long long* frame_off, unsigned long long capacity)
{
    // the per-frame counts, gathered behind the header so that one small transfer brings header + counts to the host
    for (std::uint32_t f = threadIdx.x; f < nf; f += blockDim.x)
    {
        counts[0 * nf + f] = d.n_in[f];
        counts[1 * nf + f] = d.n_v[f];
        counts[2 * nf + f] = d.n_o[f];
        counts[3 * nf + f] = d.n_clusters[f];
        counts[4 * nf + f] = d.n_hull[f];
    }
    // five running sums over the frames: n, n_o, K + 1, K, Hv  (one block; nf is a few thousand at most)
    __shared__ unsigned long long tot[5];
    __shared__ std::uint32_t sh[33];
    for (int q = 0; q < 5; ++q)
    {
        std::uint32_t carry = 0;
        for (std::uint32_t base = 0; base < nf; base += blockDim.x)
        {
            const std::uint32_t f = base + threadIdx.x;
            std::uint32_t v = 0;
            if (f < nf)
            {
                v = q == 0 ? d.n_in[f] : q == 1 ? d.n_o[f] : q == 2 ? d.n_clusters[f] + 1u : q == 3 ? d.n_clusters[f] : d.n_hull[f];
            }
            std::uint32_t t;
            const std::uint32_t ex = block_excl_scan(v, sh, &t);
            if (f < nf)
            {
                frame_off[static_cast<std::size_t>(q) * nf + f] = carry + ex;
            }
            carry += t;
            __syncthreads();
        }
        if (threadIdx.x == 0)
        {
            tot[q] = carry;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long off = 0;
        for (int p = 0; p < kPackPlanes; ++p)
        {
            hdr->offset[p] = ~0ULL;
            if (planes & (1u << p))
            {
                hdr->offset[p] = off;
                off += tot[pack_count_of(p)] * pack_elem(p);
                off = (off + 15ULL) & ~15ULL;
            }
        }
        hdr->total = off;
        hdr->fits = off <= capacity ? 1u : 0u;
    }
}

// element e of frame f of a plane -> dst[offset + (frame_off[f] + e) * elem]; units of 4 bytes when the element is a
// multiple of 4 bytes, half-words / bytes otherwise
template <int kUnit>
__device__ __forceinline__ void pack_plane(const unsigned char* __restrict__ src, std::size_t src_frame_bytes, std::uint32_t elem,
                                           const std::uint32_t* __restrict__ count, std::uint32_t count_add,
                                           const unsigned long long* __restrict__ frame_off, const PackHeader* __restrict__ hdr,
                                           int plane, unsigned char* __restrict__ staging)
{
    const std::uint32_t f = blockIdx.y;
    const std::size_t bytes = static_cast<std::size_t>(count[f] + count_add) * elem;
    const unsigned char* s = src + static_cast<std::size_t>(f) * src_frame_bytes;
    unsigned char* o = staging + hdr->offset[plane] + frame_off[f] * elem;
    // the source frame starts 16-byte aligned and is padded to a multiple of 2048 elements, so it is read in
    // whole words (the last one may run past the count, still inside the frame's plane); the destination of a
    // byte / half-word plane starts wherever the previous frame ended and is written in its own unit
    const std::size_t words = (bytes + 3u) / 4u;
    for (std::size_t u = static_cast<std::size_t>(blockIdx.x) * 256u + threadIdx.x; u < words; u += static_cast<std::size_t>(gridDim.x) * 256u)
    {
        const std::uint32_t w = reinterpret_cast<const std::uint32_t*>(s)[u];
        if (kUnit == 4)
        {
            reinterpret_cast<std::uint32_t*>(o)[u] = w;
        }
        else if (kUnit == 2)
        {
            reinterpret_cast<std::uint16_t*>(o)[2u * u] = static_cast<std::uint16_t>(w);
            if (4u * u + 2u < bytes)
            {
                reinterpret_cast<std::uint16_t*>(o)[2u * u + 1u] = static_cast<std::uint16_t>(w >> 16);
            }
        }
        else
        {
#pragma unroll
            for (std::uint32_t k = 0; k < 4u; ++k)
            {
                if (4u * u + k < bytes)
                {
                    o[4u * u + k] = static_cast<unsigned char>(w >> (8u * k));
                }
            }
        }
    }
}

// all selected planes in ONE launch: blockIdx.z = plane (a launch per plane cost more than the copies of a small batch)
struct PackPlanes
{
    const unsigned char* src[kPackPlanes];
    std::size_t frame_bytes[kPackPlanes];
    const std::uint32_t* cnt[kPackPlanes];
    std::uint32_t add[kPackPlanes];
    std::uint32_t selected; // plane bits
};

__global__ void __launch_bounds__(256)
    k_pack_planes(PackPlanes pp, std::uint32_t nf, const unsigned long long* __restrict__ frame_off,
                  const PackHeader* __restrict__ hdr, unsigned char* __restrict__ staging)
{
    const int p = static_cast<int>(blockIdx.z);
    if ((pp.selected & (1u << p)) == 0u || hdr->fits == 0u)
    {
        return;
    }
    const std::uint32_t elem = pack_elem(p);
    const unsigned long long* fo = frame_off + static_cast<std::size_t>(pack_count_of(p)) * nf;
    if (elem % 4u == 0u)
    {
        pack_plane<4>(pp.src[p], pp.frame_bytes[p], elem, pp.cnt[p], pp.add[p], fo, hdr, p, staging);
    }
    else if (elem == 2u)
    {
        pack_plane<2>(pp.src[p], pp.frame_bytes[p], elem, pp.cnt[p], pp.add[p], fo, hdr, p, staging);
    }
    else
    {
        pack_plane<1>(pp.src[p], pp.frame_bytes[p], elem, pp.cnt[p], pp.add[p], fo, hdr, p, staging);
    }
}

void launch_pack_results(Ctx* c, std::uint32_t nf, std::uint32_t planes, unsigned char* staging, std::size_t staging_bytes)
{
    Dev& d = c->d;
    // staging: [PackHeader][counts: 5 x nf u32][frame_off: 5 x nf u64][payload ...] (common.cuh: pack_payload_start)
    PackHeader* hdr = reinterpret_cast<PackHeader*>(staging);
    std::uint32_t* counts = reinterpret_cast<std::uint32_t*>(staging + sizeof(PackHeader));
    unsigned long long* frame_off = reinterpret_cast<unsigned long long*>(staging + pack_frame_off_start(nf));
    const std::size_t head = pack_payload_start(nf);
    unsigned char* payload = staging + head;
    k_pack_layout<<<1, 1024, 0, c->stream>>>(d, nf, planes, hdr, counts, frame_off, staging_bytes > head ? staging_bytes - head : 0);
    mark(c, "pack_layout");
    const std::size_t cap = d.cap;
    struct Src
    {
        const void* p;
        std::size_t frame_elems;
        const std::uint32_t* cnt;
        std::uint32_t add;
    };
    const Src src[kPackPlanes] = {
        {d.labels_out, cap, d.n_in, 0},  {d.noise, cap, d.n_in, 0},        {d.ring, cap, d.n_in, 0},
        {d.idx_o, cap, d.n_o, 0},        {d.clabel, cap, d.n_o, 0},        {d.hull_off, cap + 1, d.n_clusters, 1},
        {d.hull_idx, cap, d.n_hull, 0},  {d.hull_xy, cap, d.n_hull, 0},    {d.zminmax, cap, d.n_clusters, 0},
        {d.boxes, cap, d.n_clusters, 0},
    };
    PackPlanes pp{};
    pp.selected = planes;
    for (int p = 0; p < kPackPlanes; ++p)
    {
        pp.src[p] = static_cast<const unsigned char*>(src[p].p);
        pp.frame_bytes[p] = src[p].frame_elems * pack_elem(p);
        pp.cnt[p] = src[p].cnt;
        pp.add[p] = src[p].add;
    }
    // CTAs per frame: enough to cover a typical frame's point planes in a few trips (the small planes finish at once)
    k_pack_planes<<<dim3(per_frame_ctas(48, nf, 128), nf, kPackPlanes), 256, 0, c->stream>>>(pp, nf, frame_off, hdr, payload);
    mark(c, "pack_planes");
}

void launch_unpack_cloud2(Ctx* c, std::uint32_t nf, const unsigned char* raw, std::size_t raw_stride, const void* desc)
{
    k_unpack_cloud2<<<dim3((c->d.cap + 255) / 256, nf), 256, 0, c->stream>>>(c->d, raw, raw_stride,
                                                                             static_cast<const Cloud2Desc*>(desc));
    mark(c, "unpack_cloud2");
}
} // namespace lpl

// ------------------------------------------------------------------------------------------
// PCD reader (host)
// ------------------------------------------------------------------------------------------
namespace
{
struct PcdHeader
{
    std::vector<std::string> fields;
    std::vector<int> size, count;
    std::vector<char> type;
    unsigned long long points = 0, width = 0, height = 1;
    std::string data;
    bool have_points = false;
};

std::vector<std::string> split_ws(const std::string& s)
{
    std::vector<std::string> out;
    std::size_t i = 0;
    while (i < s.size())
    {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r'))
        {
            ++i;
        }
        std::size_t j = i;
        while (j < s.size() && s[j] != ' ' && s[j] != '\t' && s[j] != '\r')
        {
            ++j;
        }
        if (j > i)
        {
            out.push_back(s.substr(i, j - i));
        }
        i = j;
    }
    return out;
}

bool read_line(std::FILE* fp, std::string& line)
{
    line.clear();
    int ch;
    while ((ch = std::fgetc(fp)) != EOF)
    {
        if (ch == '\n')
        {
            return true;
        }
        if (line.size() >= (std::size_t(1) << 20))
        {
            return false; // no header or ascii record line is this long: refuse rather than grow without bound
        }
        line.push_back(static_cast<char>(ch));
    }
    return !line.empty();
}
} // namespace

namespace
{
constexpr int kPcdMaxRecord = 4096; // bytes per point record accepted from a header
constexpr int kPcdMaxFields = 256;

int pcd_read_impl(const char* path, float* xyzi_out, uint32_t capacity, uint32_t* n_out)
{
    if (path == nullptr || n_out == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    *n_out = 0;
    std::FILE* fp = std::fopen(path, "rb");
    if (fp == nullptr)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    struct Closer
    {
        std::FILE* f;
        ~Closer() { std::fclose(f); }
    } closer{fp}; // also closes when an allocation throws
    PcdHeader h;
    std::string line;
    while (read_line(fp, line))
    {
        const std::vector<std::string> t = split_ws(line);
        if (t.empty() || t[0][0] == '#')
        {
            continue;
        }
        if (t[0] == "FIELDS" || t[0] == "COLUMNS")
        {
            h.fields.assign(t.begin() + 1, t.end());
        }
        else if (t[0] == "SIZE")
        {
            for (std::size_t k = 1; k < t.size(); ++k)
            {
                h.size.push_back(std::atoi(t[k].c_str()));
            }
        }
        else if (t[0] == "TYPE")
        {
            for (std::size_t k = 1; k < t.size(); ++k)
            {
                h.type.push_back(t[k][0]);
            }
        }
        else if (t[0] == "COUNT")
        {
            for (std::size_t k = 1; k < t.size(); ++k)
            {
                h.count.push_back(std::atoi(t[k].c_str()));
            }
        }
        else if (t[0] == "WIDTH" && t.size() > 1)
        {
            h.width = std::strtoull(t[1].c_str(), nullptr, 10);
        }
        else if (t[0] == "HEIGHT" && t.size() > 1)
        {
            h.height = std::strtoull(t[1].c_str(), nullptr, 10);
        }
        else if (t[0] == "POINTS" && t.size() > 1)
        {
            h.points = std::strtoull(t[1].c_str(), nullptr, 10);
            h.have_points = true;
        }
        else if (t[0] == "DATA" && t.size() > 1)
        {
            h.data = t[1];
            break; // the payload follows this line
        }
    }
    const std::size_t nf = h.fields.size();
    if (h.count.empty())
    {
        h.count.assign(nf, 1);
    }
    if (nf == 0 || nf > static_cast<std::size_t>(kPcdMaxFields) || h.size.size() != nf || h.type.size() != nf ||
        h.count.size() != nf || h.data.empty())
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    // every field - not only the ones that are read - must have a sane size and count: the record
    // length and the byte offsets / columns of x, y, z, intensity are sums over all of them
    for (std::size_t k = 0; k < nf; ++k)
    {
        if (h.size[k] <= 0 || h.size[k] > 8 || h.count[k] <= 0 || h.count[k] > kPcdMaxRecord)
        {
            return LPL_ERR_INVALID_ARGUMENT;
        }
    }
    if (!h.have_points)
    {
        if (h.height != 0 && h.width > 0xffffffffULL / h.height)
        {
            return LPL_ERR_CAPACITY;
        }
        h.points = h.width * h.height;
    }
    if (h.points > 0xffffffffULL)
    {
        return LPL_ERR_CAPACITY; // n_out is 32 bits wide
    }
    // byte offset (binary) / column (ascii) of x, y, z, intensity
    int off[4] = {-1, -1, -1, -1}, colidx[4] = {-1, -1, -1, -1};
    int rec = 0, cols = 0;
    const char* want[4] = {"x", "y", "z", "intensity"};
    for (std::size_t k = 0; k < nf; ++k)
    {
        for (int w = 0; w < 4; ++w)
        {
            if (h.fields[k] == want[w])
            {
                if (h.size[k] != 4 || h.type[k] != 'F' || h.count[k] != 1)
                {
                    return LPL_ERR_INVALID_ARGUMENT; // only float32 coordinates (what the node consumes)
                }
                off[w] = rec;
                colidx[w] = cols;
            }
        }
        rec += h.size[k] * h.count[k];
        cols += h.count[k];
        if (rec > kPcdMaxRecord)
        {
            return LPL_ERR_INVALID_ARGUMENT;
        }
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0 || rec <= 0)
    {
        return LPL_ERR_INVALID_ARGUMENT;
    }
    for (int w = 0; w < 4; ++w)
    {
        if (off[w] >= 0 && (off[w] + 4 > rec || colidx[w] >= cols))
        {
            return LPL_ERR_INVALID_ARGUMENT;
        }
    }
    *n_out = static_cast<uint32_t>(h.points);
    if (xyzi_out == nullptr)
    {
        return LPL_OK; // size query
    }
    if (h.points > capacity)
    {
        return LPL_ERR_CAPACITY;
    }
    int rc = LPL_OK;
    if (h.data == "binary")
    {
        const bool packed = (rec == 16 && off[0] == 0 && off[1] == 4 && off[2] == 8 && off[3] == 12);
        if (packed)
        {
            if (std::fread(xyzi_out, 16, h.points, fp) != h.points)
            {
                rc = LPL_ERR_INVALID_ARGUMENT;
            }
        }
        else
        {
            std::vector<unsigned char> buf(static_cast<std::size_t>(rec) * 4096);
            unsigned long long done = 0;
            while (done < h.points && rc == LPL_OK)
            {
                const std::size_t take = static_cast<std::size_t>(std::min<unsigned long long>(4096, h.points - done));
                if (std::fread(buf.data(), static_cast<std::size_t>(rec), take, fp) != take)
                {
                    rc = LPL_ERR_INVALID_ARGUMENT;
                    break;
                }
                for (std::size_t i = 0; i < take; ++i)
                {
                    float* o = xyzi_out + (done + i) * 4;
                    for (int w = 0; w < 4; ++w)
                    {
                        o[w] = 0.f;
                        if (off[w] >= 0)
                        {
                            std::memcpy(&o[w], buf.data() + i * rec + off[w], 4);
                        }
                    }
                }
                done += take;
            }
        }
    }
    else if (h.data == "ascii")
    {
        for (unsigned long long i = 0; i < h.points && rc == LPL_OK; ++i)
        {
            if (!read_line(fp, line))
            {
                rc = LPL_ERR_INVALID_ARGUMENT;
                break;
            }
            const std::vector<std::string> t = split_ws(line);
            if (static_cast<int>(t.size()) < cols)
            {
                rc = LPL_ERR_INVALID_ARGUMENT;
                break;
            }
            float* o = xyzi_out + i * 4;
            for (int w = 0; w < 4; ++w)
            {
                o[w] = colidx[w] >= 0 ? std::strtof(t[colidx[w]].c_str(), nullptr) : 0.f;
            }
        }
    }
    else
    {
        rc = LPL_ERR_INVALID_ARGUMENT; // binary_compressed is not produced by the reference's data set
    }
    if (rc != LPL_OK)
    {
        *n_out = 0;
    }
    return rc;
}
} // namespace

extern "C" int lpl_pcd_read(const char* path, float* xyzi_out, uint32_t capacity, uint32_t* n_out)
{
    // nothing may unwind through the C ABI (std::string / std::vector growth can throw)
    try
    {
        return pcd_read_impl(path, xyzi_out, capacity, n_out);
    }
    catch (...)
    {
        if (n_out != nullptr)
        {
            *n_out = 0;
        }
        return LPL_ERR_INVALID_ARGUMENT;
    }
}
