// Stage 3: curved-voxel clustering of the obstacle cloud, batched over frames.
//
// Reference: lidar_processing_lib/src/clusterer.cpp
//   cartesianToSpherical :55-100   -> k_clu_sph (glibc-exact atan2f / atanf, frame-wide maxima)
//   buildHashTable       :102-120  -> k_clu_insert (device open-addressing hash, key = flat index)
//   clusterImpl          :122-193  -> k_clu_edges + k_clu_union_sm (k_clu_union / k_clu_flatten for
//                                     oversized frames): the BFS over 26-connected occupied voxels
//                                     computes connected components; here they come from a
//                                     lock-free union-find over the voxel list, held in shared
//                                     memory per frame
//   removeSmallClusters  :195-239  -> stable compaction of the component representatives
//
// Label order: the reference opens a new cluster at the first point (in cloud order) whose
// voxel is unlabelled, so cluster ids are ranked by the component's minimum point index, and
// the small-cluster pass renumbers the survivors in that same order. A stable compaction over
// "point i is the minimum point of a surviving component" yields exactly those ids.
//
// The azimuth wrap is the reference's literal one (index -1 -> num_azimuth - 1, num_azimuth -> 0
// with num_azimuth = ceil(max_az / res) + 1), see DESIGN.md hazard H3.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace lpl
{
// k_clu_sph and k_clu_edges stride a fixed number of CTAs over a frame's points / voxels (kCluCtas tiles of 256 cover a
// typical HDL-64E frame's obstacle points in one trip): a grid sized to the frame CAPACITY launched 60 - 90 % of its CTAs
// only to read the count and leave (edges 0.134 -> 0.106 ms). k_clu_insert and k_clu_labels keep the capacity grid:
// with the loop, the frames that need a second trip became the tail of the launch (labels 0.226 -> 0.262 ms).
constexpr std::uint32_t kCluCtas = 224;

__global__ void __launch_bounds__(256) k_clu_sph(Dev d)
{
    __shared__ std::uint32_t smax[3];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_o[f];
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    if (threadIdx.x < 3)
    {
        smax[threadIdx.x] = 0;
    }
    __syncthreads();
    std::uint32_t br = 0, ba = 0, be = 0;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    for (std::uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u)
    {
        const float4 p = d.pts_o[o + i];
        float az = atan2f_glibc(p.y, p.x);
        az = (az < 0.f) ? (az + 6.28318530717958647692f) : az;
        const float dxy2 = p.x * p.x + p.y * p.y;
        const float dxy = sqrtf(dxy2);
        const float range = sqrtf(dxy2 + p.z * p.z);
        const float el = atanf_glibc(p.z / dxy) + 1.57079632679489661923f;
        d.sph[o + i] = make_float4(range, az, el, 0.f);
        // non-negative floats order like their bit patterns; "+ 0.0f" folds -0.0 into +0.0
        br = max(br, __float_as_uint(range + 0.0f));
        ba = max(ba, __float_as_uint(az + 0.0f));
        be = max(be, __float_as_uint(el + 0.0f));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
    {
        br = max(br, __shfl_xor_sync(0xffffffffu, br, s));
        ba = max(ba, __shfl_xor_sync(0xffffffffu, ba, s));
        be = max(be, __shfl_xor_sync(0xffffffffu, be, s));
    }
    if (lane_id() == 0)
    {
        atomicMax(&smax[0], br);
        atomicMax(&smax[1], ba);
        atomicMax(&smax[2], be);
    }
    __syncthreads();
    if (threadIdx.x < 3)
    {
        atomicMax(&d.sph_max[f * 4 + threadIdx.x], smax[threadIdx.x]);
    }
}

struct VoxelDims
{
    std::int32_t nr, na, ne;
};

__device__ __forceinline__ VoxelDims voxel_dims(const Dev& d, const ClusterParams& cp, std::uint32_t f)
{
    // clusterer.cpp:92-99
    const float mr = __uint_as_float(d.sph_max[f * 4 + 0]);
    const float ma = __uint_as_float(d.sph_max[f * 4 + 1]);
    const float me = __uint_as_float(d.sph_max[f * 4 + 2]);
    VoxelDims v;
    v.nr = static_cast<std::int32_t>(ceilf(mr / cp.range_res) + 1.f);
    v.na = static_cast<std::int32_t>(ceilf(ma / cp.az_res) + 1.f);
    v.ne = static_cast<std::int32_t>(ceilf(me / cp.el_res) + 1.f);
    return v;
}

// Block-coherent open addressing: the table is split into buckets of 8 slots (one 32-byte
// sector); a voxel's bucket comes from hashing (flat >> 3) and its place inside the bucket is
// (flat & 7), so the range neighbours r - 1, r, r + 1 of a voxel (consecutive flat indices)
// almost always share a sector. A taken slot sends the probe to the same place of the next bucket.
// Only the first pow2 >= 2 * n_o slots of the frame's table are used (min 1024), which keeps the
// random look-ups of a frame inside a few hundred KB of L2.
__device__ __forceinline__ std::uint32_t voxel_slots(const Dev& d, std::uint32_t f)
{
    const std::uint32_t n = d.n_o[f];
    std::uint32_t s = 1024u;
    if (n > 512u)
    {
        s = 1u << (32 - __clz(2u * n - 1u));
    }
    return min(s, d.hcap);
}

// probe t of the sequence that starts at `home`: first the same place of every bucket, then (only
// when all of those are taken, e.g. every key congruent modulo 8) plain linear probing
__device__ __forceinline__ std::uint32_t voxel_probe(std::uint32_t home, std::uint32_t t, std::uint32_t slots)
{
    const std::uint32_t nb = slots / 8u;
    return (t < nb ? home + 8u * t : home + (t - nb) + 1u) & (slots - 1u);
}

__device__ __forceinline__ std::uint32_t voxel_home(std::int32_t key, std::uint32_t slots)
{
    std::uint32_t h = (static_cast<std::uint32_t>(key) >> 3) * 0x9E3779B1u;
    h ^= h >> 15;
    return ((h & (slots / 8u - 1u)) << 3) | (static_cast<std::uint32_t>(key) & 7u);
}

__global__ void __launch_bounds__(256) k_clu_insert(Dev d, ClusterParams cp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_o[f];
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::int32_t flat = -1;
    if (i < n)
    {
        const VoxelDims vd = voxel_dims(d, cp, f);
        const float4 s = d.sph[o + i];
        const std::int32_t ri = static_cast<std::int32_t>(s.x / cp.range_res);
        const std::int32_t ai = static_cast<std::int32_t>(s.y / cp.az_res);
        const std::int32_t ei = static_cast<std::int32_t>(s.z / cp.el_res);
        flat = vd.nr * (vd.na * ei + ai) + ri; // clusterer.hpp:136-142
    }
    // neighbours along a scan line mostly share a voxel: the lowest lane of every group of equal
    // keys inserts for the group (it also carries the group's smallest point index)
    const std::uint32_t peers = __match_any_sync(0xffffffffu, flat);
    const int leader = __ffs(peers) - 1;
    __shared__ std::uint32_t s_new, s_base;
    if (threadIdx.x == 0)
    {
        s_new = 0;
    }
    __syncthreads();
    std::uint32_t slot = 0;
    bool created = false;
    const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
    if (i < n && static_cast<int>(lane_id()) == leader)
    {
        const std::uint32_t slots = voxel_slots(d, f);
        std::int32_t* keys = d.hkey + ho;
        const std::uint32_t home = voxel_home(flat, slots);
        slot = 0xffffffffu;
        for (std::uint32_t t = 0; t < slots / 8u + slots; ++t)
        {
            const std::uint32_t h = voxel_probe(home, t, slots);
            const std::int32_t prev = atomicCAS(&keys[h], -1, flat);
            if (prev == -1 || prev == flat)
            {
                slot = h;
                created = prev == -1;
                break;
            }
        }
        if (slot == 0xffffffffu)
        {
            atomicOr(&d.status[f], ST_HASH_FULL);
            slot = 0;
        }
        atomicMin(&d.hmin[ho + slot], i);
        atomicAdd(&d.hcount[ho + slot], static_cast<std::uint32_t>(__popc(peers)));
    }
    // the voxels this CTA created are appended to the frame's voxel list with one atomic per CTA
    // (a per-voxel atomic on the frame's counter serialises ~17k updates on one address; measured:
    // per voxel 0.55 ms, per warp 0.53 ms, per CTA 0.19 ms for the 154-frame batch)
    const std::uint32_t cm = __ballot_sync(0xffffffffu, created);
    std::uint32_t woff = 0;
    if (lane_id() == 0 && cm != 0)
    {
        woff = atomicAdd(&s_new, static_cast<std::uint32_t>(__popc(cm)));
    }
    woff = __shfl_sync(0xffffffffu, woff, 0);
    __syncthreads();
    if (threadIdx.x == 0 && s_new != 0)
    {
        s_base = atomicAdd(&d.n_vox[f], s_new);
    }
    __syncthreads();
    if (created)
    {
        const std::uint32_t vid = s_base + woff + __popc(cm & ((1u << lane_id()) - 1u));
        d.vlist[o + vid] = slot;
        d.hvid[ho + slot] = vid;
    }
    slot = __shfl_sync(0xffffffffu, slot, leader);
    if (i < n)
    {
        d.vslot[o + i] = slot;
    }
}

constexpr std::uint32_t kUfSmemVoxels = 24576; // occupied voxels per frame the shared-memory union-find holds (96 KB)
#ifndef LPL_UF_THREADS
#define LPL_UF_THREADS 1024 // measured: 512 -> 0.25 ms, 256 -> 0.42 ms, 1024 -> 0.19 ms per 154-frame batch
#endif
constexpr int kUfThreads = LPL_UF_THREADS;
constexpr int kFwd = 13; // forward half of the 26-neighbourhood (rows of kEdgePitch = 16 words: the plane doubles as hull.cu's octagon plane)

// Occupied voxels are numbered by their position in the frame's voxel list ("voxel id").
// k_clu_edges (one thread per voxel, whole GPU): the 26-neighbourhood is symmetric (also across
// the literal azimuth wrap and the clipped range / elevation borders), so each voxel only looks
// at the 13 "forward" offsets (de, da, dr) > (0, 0, 0); all first probes are issued before any is
// consumed, and the ids of the neighbours found go to a fixed 13-entry row per voxel.
__global__ void __launch_bounds__(256) k_clu_edges(Dev d, ClusterParams cp)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nvox = d.n_vox[f];
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
    for (std::uint32_t v = blockIdx.x * 256u + threadIdx.x; v < nvox; v += gridDim.x * 256u)
    {
    const std::uint32_t slot = d.vlist[o + v];
    const VoxelDims vd = voxel_dims(d, cp, f);
    const std::int32_t* keys = d.hkey + ho;
    const std::uint32_t slots = voxel_slots(d, f);
    const std::int32_t flat = keys[slot];
    const std::int32_t ri = flat % vd.nr;
    const std::int32_t t = flat / vd.nr;
    const std::int32_t ai = t % vd.na;
    const std::int32_t ei = t / vd.na;
    std::int32_t key2[kFwd];
    std::uint32_t slot2[kFwd];
    std::int32_t found[kFwd];
    int j = 0;
#pragma unroll
    for (int de = 0; de <= 1; ++de)
    {
#pragma unroll
        for (int da = -1; da <= 1; ++da)
        {
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr)
            {
                if (de == 0 && (da < 0 || (da == 0 && dr <= 0)))
                {
                    continue;
                }
                const std::int32_t e2 = ei + de;
                std::int32_t a2 = ai + da;
                a2 = a2 < 0 ? a2 + vd.na : (a2 >= vd.na ? a2 - vd.na : a2);
                const std::int32_t r2 = ri + dr;
                const bool ok = e2 < vd.ne && r2 >= 0 && r2 < vd.nr;
                key2[j] = ok ? vd.nr * (vd.na * e2 + a2) + r2 : -1;
                slot2[j] = voxel_home(key2[j], slots);
                ++j;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kFwd; ++q)
    {
        found[q] = key2[q] >= 0 ? keys[slot2[q]] : -1;
    }
    std::uint32_t* row = d.edges + (o + v) * kEdgePitch;
#pragma unroll
    for (int q = kFwd; q < kEdgePitch; ++q)
    {
        row[q] = 0xffffffffu; // padding reads as "no neighbour"
    }
#pragma unroll
    for (int q = 0; q < kFwd; ++q)
    {
        std::uint32_t nb = 0xffffffffu;
        if (key2[q] >= 0)
        {
            std::uint32_t h = slot2[q];
            std::int32_t kk = found[q];
            for (std::uint32_t tt = 1; tt <= slots / 8u + slots; ++tt)
            {
                if (kk == key2[q])
                {
                    nb = d.hvid[ho + h];
                    break;
                }
                if (kk == -1)
                {
                    break;
                }
                h = voxel_probe(slot2[q], tt, slots);
                kk = keys[h];
            }
        }
        row[q] = nb;
    }
    d.hparent[ho + v] = v; // forest of the global path, indexed by voxel id
    }
}

// parent[] is updated with atomics while other threads walk it. The walk halves the path as it
// goes (every visited node is re-pointed to its grandparent): parents only ever move to
// ancestors, so racing walkers and hooks stay correct. Voxel ids follow the scan order, so
// hooking by id would build path-like trees; a multiplicative hash of the id gives a scattered
// (but fixed) priority instead. The global variant reads through L2 (ld.cg) so that a stale L1
// line can never make a failed CAS retry forever.
__device__ __forceinline__ bool uf_before(std::uint32_t a, std::uint32_t b)
{
    return a * 0x9E3779B1u < b * 0x9E3779B1u;
}

__device__ __forceinline__ std::uint32_t uf_find(std::uint32_t* parent, std::uint32_t x)
{
    std::uint32_t p = __ldcg(parent + x);
    while (p != x)
    {
        const std::uint32_t g = __ldcg(parent + p);
        if (g != p)
        {
            parent[x] = g;
        }
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ std::uint32_t uf_find_sm(volatile std::uint32_t* parent, std::uint32_t x)
{
    std::uint32_t p = parent[x];
    while (p != x)
    {
        const std::uint32_t g = parent[p];
        if (g != p)
        {
            parent[x] = g;
        }
        x = p;
        p = g;
    }
    return x;
}

// component statistics at the root slot (a non-root voxel's own count / minimum are final after
// k_clu_insert)
__device__ __forceinline__ void uf_publish(const Dev& d, std::size_t o, std::size_t ho, std::uint32_t v, std::uint32_t r)
{
    const std::uint32_t slot = d.vlist[o + v];
    const std::uint32_t root = d.vlist[o + r];
    d.hroot[ho + slot] = root;
    d.hlabel[ho + slot] = -1;
    if (r != v)
    {
        atomicMin(&d.hmin[ho + root], d.hmin[ho + slot]);
        atomicAdd(&d.hcount[ho + root], d.hcount[ho + slot]);
    }
}

// Shared-memory union-find: a frame's forest (typically 10-17k voxels) lives in one CTA's shared
// memory, so the pointer chasing of find() costs shared-memory instead of L2 latency; the CTA
// streams the frame's edge rows (coalesced) and hooks lock-free with shared-memory CAS.
__global__ void __launch_bounds__(kUfThreads) k_clu_union_sm(Dev d, std::uint32_t sm_max)
{
    extern __shared__ std::uint32_t par[];
    const std::uint32_t f = blockIdx.x + d.f0;
    const std::uint32_t nv = d.n_vox[f];
    if (nv == 0 || nv > sm_max)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
    for (std::uint32_t v = threadIdx.x; v < nv; v += kUfThreads)
    {
        par[v] = v;
    }
    __syncthreads();
    const std::uint32_t* edges = d.edges + o * kEdgePitch;
    // the next edge word is in flight while the current one is united (a single frame has one CTA walking ~220k edge
    // words: without this every trip waits for its own load)
    std::uint32_t u_next = threadIdx.x < nv * kEdgePitch ? edges[threadIdx.x] : 0xffffffffu;
    for (std::uint32_t e = threadIdx.x; e < nv * kEdgePitch; e += kUfThreads)
    {
        const std::uint32_t u = u_next;
        u_next = e + kUfThreads < nv * kEdgePitch ? edges[e + kUfThreads] : 0xffffffffu;
        if (u == 0xffffffffu)
        {
            continue;
        }
        std::uint32_t a = e / kEdgePitch, b = u;
        while (true)
        {
            a = uf_find_sm(par, a);
            b = uf_find_sm(par, b);
            if (a == b)
            {
                break;
            }
            if (uf_before(a, b))
            {
                const std::uint32_t sw = a;
                a = b;
                b = sw;
            }
            if (atomicCAS(&par[a], a, b) == a)
            {
                break;
            }
        }
    }
    __syncthreads();
    for (std::uint32_t v = threadIdx.x; v < nv; v += kUfThreads)
    {
        uf_publish(d, o, ho, v, uf_find_sm(par, v));
    }
}

// Global-memory path for frames with more occupied voxels than the shared-memory forest holds
// (e.g. the 2M-point clouds): one thread per voxel walks its edge row.
__global__ void __launch_bounds__(256) k_clu_union(Dev d, std::uint32_t sm_max)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nvox = d.n_vox[f];
    if (nvox <= sm_max)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    std::uint32_t* parent = d.hparent + static_cast<std::size_t>(f) * d.hcap;
    for (std::uint32_t v = blockIdx.x * 256u + threadIdx.x; v < nvox; v += gridDim.x * 256u)
    {
    const std::uint32_t* row = d.edges + (o + v) * kEdgePitch;
    for (int q = 0; q < kFwd; ++q)
    {
        const std::uint32_t u = row[q];
        if (u == 0xffffffffu)
        {
            continue;
        }
        std::uint32_t a = v, b = u;
        while (true)
        {
            a = uf_find(parent, a);
            b = uf_find(parent, b);
            if (a == b)
            {
                break;
            }
            if (uf_before(a, b))
            {
                const std::uint32_t sw = a;
                a = b;
                b = sw;
            }
            if (atomicCAS(&parent[a], a, b) == a)
            {
                break;
            }
        }
    }
    }
}

__global__ void __launch_bounds__(256) k_clu_flatten(Dev d, std::uint32_t sm_max)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t nvox = d.n_vox[f];
    if (nvox <= sm_max)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
    for (std::uint32_t v = blockIdx.x * 256u + threadIdx.x; v < nvox; v += gridDim.x * 256u)
    {
        uf_publish(d, o, ho, v, uf_find(d.hparent + ho, v));
    }
}

struct ClusterRepPred
{
    Dev d;
    std::uint32_t min_size;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        // the minimum point of a component is also the minimum point of its own voxel (for a root
        // voxel hmin already holds the component's minimum): one look-up rejects most points
        const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
        const std::uint32_t slot = d.vslot[static_cast<std::size_t>(f) * d.cap + i];
        if (d.hmin[ho + slot] != i)
        {
            return false;
        }
        const std::uint32_t root = d.hroot[ho + slot];
        return d.hmin[ho + root] == i && d.hcount[ho + root] >= min_size;
    }
};

struct ClusterRepEmit
{
    Dev d;
    __device__ void operator()(std::uint32_t f, std::uint32_t i, std::uint32_t pos) const
    {
        const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
        const std::uint32_t root = d.hroot[ho + d.vslot[static_cast<std::size_t>(f) * d.cap + i]];
        d.hlabel[ho + root] = static_cast<std::int32_t>(pos);
        const std::size_t o = static_cast<std::size_t>(f) * d.cap;
        d.ccount[o + pos] = d.hcount[ho + root];
        d.zmin_u[o + pos] = 0xffffffffu; // z extent accumulators of the new label
        d.zmax_u[o + pos] = 0u;
        d.zzero[o + pos] = 0xffffffffu;
        ext_init(d.ext + (o + pos) * kExtDirs);
    }
};

__global__ void __launch_bounds__(256) k_clu_labels(Dev d)
{
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t n = d.n_o[f];
    const std::uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (blockIdx.x * 256u >= n)
    {
        return;
    }
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::size_t ho = static_cast<std::size_t>(f) * d.hcap;
    // The hash planes (key / minimum point / count) are self-cleaning: once the cluster ids are ranked (k_clu_rank is
    // their last reader), the slots of the frame's occupied voxels (~17k of 262k) are reset - instead of three
    // full-table memsets (484 MB per 154-frame batch) ahead of every run. There are at most as many voxels as points,
    // so the threads of this launch cover them; the labels below only read hroot / hlabel.
    if (i < d.n_vox[f])
    {
        const std::uint32_t slot = d.vlist[o + i];
        d.hkey[ho + slot] = -1;
        d.hmin[ho + slot] = 0xffffffffu;
        d.hcount[ho + slot] = 0u;
    }
    std::int32_t l = -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < n)
    {
        const float4 p = d.pts_o[o + i]; // in flight while the three dependent look-ups below resolve
        l = d.hlabel[ho + d.hroot[ho + d.vslot[o + i]]];
        d.clabel[o + i] = l;
        x = p.x;
        y = p.y;
        z = p.z;
    }
    // every kExtWarpStride-th warp of the frame contributes to the extreme points
    const bool with_extremes = (((blockIdx.x * 256u + threadIdx.x) >> 5) % kExtWarpStride) == 0u;
    accumulate_cluster_stats(d.ext + o * kExtDirs, d.zmin_u + o, d.zmax_u + o, d.zzero + o, l, x, y, z, i, with_extremes);
}

// hand-over from segmentation: stable compaction of OBSTACLE points in cloud order
// (src/processor/src/processor.cpp:562-579)
struct ObstaclePred
{
    const std::uint8_t* labels; // Label per input point
    std::uint32_t cap;
    __device__ bool operator()(std::uint32_t f, std::uint32_t i) const
    {
        return labels[static_cast<std::size_t>(f) * cap + i] == PX_OBSTACLE;
    }
};

struct ObstacleEmit
{
    const float4* pts_in;
    float4* pts_o;
    std::uint32_t* idx_o;
    std::uint32_t cap;
    __device__ void operator()(std::uint32_t f, std::uint32_t i, std::uint32_t pos) const
    {
        const std::size_t o = static_cast<std::size_t>(f) * cap;
        pts_o[o + pos] = pts_in[o + i];
        idx_o[o + pos] = i;
    }
};

void launch_take_obstacles(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    launch_compact(c, "take_obstacles", nf, d.tiles, d.n_in, 0u, d.tile_cnt, d.n_o, ObstaclePred{d.labels_out, d.cap},
                   ObstacleEmit{d.pts_in, d.pts_o, d.idx_o, d.cap});
}

void launch_cluster(Ctx* c, std::uint32_t nf)
{
    Dev& d = c->d;
    cudaStream_t s = c->stream;
    if (!c->counters_cleared)
    {
        cudaMemsetAsync(at_frame(d.sph_max, 4, d.f0), 0, sizeof(std::uint32_t) * 4 * nf, s);
        cudaMemsetAsync(at_frame(d.n_vox, 1, d.f0), 0, sizeof(std::uint32_t) * nf, s);
    }
    if (!c->hash_clean)
    {
        // first use of this context: all frames' tables; afterwards k_clu_labels resets the used slots
        cudaMemsetAsync(d.hkey, 0xff, sizeof(std::int32_t) * static_cast<std::size_t>(d.hcap) * d.B, s);
        cudaMemsetAsync(d.hmin, 0xff, sizeof(std::uint32_t) * static_cast<std::size_t>(d.hcap) * d.B, s);
        cudaMemsetAsync(d.hcount, 0, sizeof(std::uint32_t) * static_cast<std::size_t>(d.hcap) * d.B, s);
        c->hash_clean = true;
    }
    const dim3 g((d.cap + 255) / 256, nf);
    k_clu_sph<<<dim3(std::min<std::uint32_t>((d.cap + 255) / 256, per_frame_ctas(kCluCtas, nf, 4096)), nf), 256, 0, s>>>(d);
    mark(c, "clu_sph");
    k_clu_insert<<<g, 256, 0, s>>>(d, c->clu);
    mark(c, "clu_insert");
    k_clu_edges<<<dim3(std::min<std::uint32_t>((d.cap + 255) / 256, per_frame_ctas(72, nf, 4096)), nf), 256, 0, s>>>(d, c->clu);
    mark(c, "clu_edges");
    cudaFuncSetAttribute(k_clu_union_sm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(kUfSmemVoxels * sizeof(std::uint32_t)));
    // One CTA per frame is the right shape for a batch that fills the GPU with frames. A launch of a few frames leaves
    // the other SMs idle for the whole (latency-bound) walk, so it takes the global-memory form, which spreads a frame's
    // voxels over the whole GPU: a single frame 0.105 -> 0.06 ms (chain 0.81 -> 0.76 ms), five 2M-point clouds 2.63 -> 2.44 ms; at 16
    // frames both forms take the same time.
    static const std::uint32_t few = std::getenv("LPL_UF_GLOBAL_MAX_FRAMES") != nullptr
                                         ? static_cast<std::uint32_t>(std::atoi(std::getenv("LPL_UF_GLOBAL_MAX_FRAMES")))
                                         : 8u;
    const std::uint32_t sm_max = nf <= few ? 0u : kUfSmemVoxels;
    k_clu_union_sm<<<nf, kUfThreads, kUfSmemVoxels * sizeof(std::uint32_t), s>>>(d, sm_max);
    mark(c, "clu_union_sm");
    // frames with more occupied voxels than the shared-memory forest holds take the global path
    // about one resident wave of CTAs in total: the kernels stride over a frame's voxels
    const dim3 gbig(std::max(1u, std::min<std::uint32_t>((d.cap + 255) / 256, (148u * 8u + nf - 1u) / nf)), nf);
    k_clu_union<<<gbig, 256, 0, s>>>(d, sm_max);
    mark(c, "clu_union");
    k_clu_flatten<<<gbig, 256, 0, s>>>(d, sm_max);
    mark(c, "clu_flatten");
    // d.lab (RECM labels of the segmenter) is free by now: it holds the recorded verdicts
    launch_compact_recorded(c, "clu_rank", nf, d.tiles, d.n_o, d.tile_cnt, d.n_clusters, d.lab,
                            ClusterRepPred{d, c->clu.min_cluster_size}, ClusterRepEmit{d});
    k_clu_labels<<<g, 256, 0, s>>>(d);
    mark(c, "clu_labels");
}
} // namespace lpl
