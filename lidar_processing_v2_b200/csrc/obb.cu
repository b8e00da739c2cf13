// Oriented bounding boxes of convex hulls, batched: one warp per hull.
//
// Reference: lidar_processing_lib/src/polygonizer.cpp
//   findAntipodalPairsOfConvexHull :93-163  (Shamos)            -> shamos_next (lane 0, resumable)
//   boundingBoxRotatingCalipers    :165-278                     -> obb_calipers
//   boundingBoxPrincipalComponentAnalysis :280-362              -> obb_pca (Eigen's 2 x 2 two-sided
//                                    Jacobi SVD restated; Eigen is absent here: parity unpinned)
//
// Rotating calipers in the reference: for every antipodal pair (i, j), for index in {i, j}, for
// offset in {-1, +1}: the edge hull[index] -> hull[index + offset] gives a direction, every hull
// point is rotated into it, and the axis-aligned extent there is a candidate box; the first
// strictly smaller area wins, where the running minimum is kept in a *float* (BoundingBox::area).
// The candidate sequence is sequential by construction (Shamos walks two pointers around the
// hull), its evaluation is not: lane 0 emits pairs 32 at a time, the 128 candidates of a chunk are
// evaluated by all lanes (each an O(n) sweep in fp64 without FMA contraction, as compiled in the
// reference), and lane 0 replays the float-rounded `area < min` selection in candidate order.
#include <algorithm>

#include "common.cuh"

namespace lpl
{
namespace
{
struct P2d
{
    double x, y;
};

struct HullF2
{
    const float2* p;
    __device__ __forceinline__ P2d operator()(int i) const
    {
        const float2 v = p[i];
        return P2d{static_cast<double>(v.x), static_cast<double>(v.y)};
    }
};

struct HullD2
{
    const double2* p;
    __device__ __forceinline__ P2d operator()(int i) const
    {
        const double2 v = p[i];
        return P2d{v.x, v.y};
    }
};

// polygonizer.hpp:226-231
__device__ __forceinline__ double tri_area(const P2d& p1, const P2d& p2, const P2d& p3)
{
    return fabs((p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y)) * 0.5);
}

__device__ __forceinline__ int nxt(int k, int n) { return (k + 1 == n) ? 0 : (k + 1); }

// polygonizer.cpp:93-163 as a resumable generator: every call yields the next pair
struct Shamos
{
    int n, i, j, j0, i0, state; // state 0: outer loop head, 1: inner while, 2: parallel-edge check, 3: done
};

template <class Pts>
__device__ void shamos_init(Shamos& s, const Pts& h, int n)
{
    s.n = n;
    s.i0 = n - 1;
    s.i = 0;
    s.j = 1;
    s.state = 0;
    while (tri_area(h(s.i), h(nxt(s.i, n)), h(nxt(s.j, n))) > tri_area(h(s.i), h(nxt(s.i, n)), h(s.j)))
    {
        s.j = nxt(s.j, n);
    }
    s.j0 = s.j;
}

template <class Pts>
__device__ bool shamos_next(Shamos& s, const Pts& h, int& pi, int& pj)
{
    const int n = s.n;
    while (true)
    {
        if (s.state == 0)
        {
            if (s.i == s.j0)
            {
                s.state = 3;
                return false;
            }
            s.i = nxt(s.i, n);
            pi = s.i;
            pj = s.j;
            s.state = 1;
            return true;
        }
        if (s.state == 1)
        {
            if (tri_area(h(s.i), h(nxt(s.i, n)), h(nxt(s.j, n))) > tri_area(h(s.i), h(nxt(s.i, n)), h(s.j)))
            {
                s.j = nxt(s.j, n);
                if (!(s.i == s.j0 && s.j == s.i0))
                {
                    pi = s.i;
                    pj = s.j;
                    return true;
                }
                s.state = 3;
                return false;
            }
            s.state = 2;
        }
        if (s.state == 2)
        {
            s.state = 0;
            if (tri_area(h(s.j), h(nxt(s.i, n)), h(nxt(s.j, n))) == tri_area(h(s.i), h(nxt(s.i, n)), h(s.j)))
            {
                if (!(s.i == s.j0 && s.j == s.i0))
                {
                    pi = s.i;
                    pj = nxt(s.j, n);
                }
                else
                {
                    pi = nxt(s.i, n);
                    pj = s.j;
                }
                return true;
            }
        }
        if (s.state == 3)
        {
            return false;
        }
    }
}

struct Extent
{
    double ux, uy, min_x, max_x, min_y, max_y;
    bool ok;
};

// one candidate direction (polygonizer.cpp:215-246)
template <class Pts>
__device__ Extent edge_extent(const Pts& h, int n, int index, int nb, const P2d& c)
{
    Extent e;
    const P2d p0 = h(index), p1 = h(nb);
    const double ex = p1.x - p0.x, ey = p1.y - p0.y;
    const double len = sqrt(ex * ex + ey * ey);
    e.ok = !(len < 1.0e-6);
    e.ux = ex / len;
    e.uy = ey / len;
    e.min_x = 1.7976931348623157e308;
    e.max_x = -1.7976931348623157e308;
    e.min_y = 1.7976931348623157e308;
    e.max_y = -1.7976931348623157e308;
    if (e.ok)
    {
        for (int k = 0; k < n; ++k)
        {
            const P2d pt = h(k);
            const double tx = pt.x - c.x, ty = pt.y - c.y;
            const double rx = tx * e.ux + ty * e.uy;
            const double ry = -tx * e.uy + ty * e.ux;
            e.min_x = rx < e.min_x ? rx : e.min_x; // std::min(min_x, rx)
            e.max_x = e.max_x < rx ? rx : e.max_x; // std::max(max_x, rx)
            e.min_y = ry < e.min_y ? ry : e.min_y;
            e.max_y = e.max_y < ry ? ry : e.max_y;
        }
    }
    return e;
}

constexpr int kObbWarps = 4;

template <class Pts>
__device__ void obb_calipers(const Pts& h, int n, ObbBox* out, int2* s_pairs, double* s_area)
{
    const std::uint32_t lane = lane_id();
    ObbBox box;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        box.c[k] = 0.0;
    }
    box.area = 0.f;
    box.angle = 0.f;
    box.valid = 0;
    box.pad = 0;
    if (n < 3)
    {
        if (lane == 0)
        {
            *out = box;
        }
        return;
    }
    // centroid: the reference's sequential sum (the order matters in fp64)
    P2d c = {0.0, 0.0};
    Shamos sh;
    if (lane == 0)
    {
        for (int k = 0; k < n; ++k)
        {
            const P2d p = h(k);
            c.x += p.x;
            c.y += p.y;
        }
        c.x /= n;
        c.y /= n;
        shamos_init(sh, h, n);
    }
    c.x = __shfl_sync(0xffffffffu, c.x, 0);
    c.y = __shfl_sync(0xffffffffu, c.y, 0);
    float best = __double2float_rn(1.7976931348623157e308); // +inf, as the float member receives DBL_MAX
    int best_index = -1, best_nb = -1;
    while (true)
    {
        int cnt = 0;
        if (lane == 0)
        {
            int pi, pj;
            while (cnt < 32 && shamos_next(sh, h, pi, pj))
            {
                s_pairs[cnt++] = make_int2(pi, pj);
            }
        }
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        if (cnt == 0)
        {
            break;
        }
        __syncwarp();
        for (int q = lane; q < 4 * cnt; q += 32)
        {
            const int2 pr = s_pairs[q >> 2];
            const int index = (q & 2) ? pr.y : pr.x;
            int nb = index + ((q & 1) ? 1 : -1);
            nb = nb < 0 ? nb + n : (nb >= n ? nb - n : nb);
            const Extent e = edge_extent(h, n, index, nb, c);
            // a skipped direction can never satisfy `area < min`
            s_area[q] = e.ok ? (e.max_x - e.min_x) * (e.max_y - e.min_y) : __longlong_as_double(0x7ff0000000000000LL);
        }
        __syncwarp();
        if (lane == 0)
        {
            for (int q = 0; q < 4 * cnt; ++q)
            {
                const double a = s_area[q];
                if (a < static_cast<double>(best))
                {
                    best = __double2float_rn(a);
                    const int2 pr = s_pairs[q >> 2];
                    best_index = (q & 2) ? pr.y : pr.x;
                    int nb = best_index + ((q & 1) ? 1 : -1);
                    best_nb = nb < 0 ? nb + n : (nb >= n ? nb - n : nb);
                }
            }
        }
        __syncwarp();
        if (cnt < 32)
        {
            break;
        }
    }
    if (lane == 0)
    {
        if (best_index >= 0)
        {
            const Extent e = edge_extent(h, n, best_index, best_nb, c);
            box.c[0] = e.min_x * e.ux - e.min_y * e.uy + c.x;
            box.c[1] = e.min_x * e.uy + e.min_y * e.ux + c.y;
            box.c[2] = e.max_x * e.ux - e.min_y * e.uy + c.x;
            box.c[3] = e.max_x * e.uy + e.min_y * e.ux + c.y;
            box.c[4] = e.max_x * e.ux - e.max_y * e.uy + c.x;
            box.c[5] = e.max_x * e.uy + e.max_y * e.ux + c.y;
            box.c[6] = e.min_x * e.ux - e.max_y * e.uy + c.x;
            box.c[7] = e.min_x * e.uy + e.max_y * e.ux + c.y;
            box.area = best;
            box.angle = atan2_approx(__double2float_rn(e.uy), __double2float_rn(e.ux));
            box.valid = 1;
        }
        *out = box;
    }
}

// polygonizer.cpp:280-362; n is small (a hull), one lane does it all
template <class Pts>
__device__ void obb_pca(const Pts& h, int n, ObbBox* out)
{
    ObbBox box;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        box.c[k] = 0.0;
    }
    box.area = 0.f;
    box.angle = 0.f;
    box.valid = 0;
    box.pad = 0;
    if (n < 3)
    {
        *out = box;
        return;
    }
    double mx = 0.0, my = 0.0;
    for (int k = 0; k < n; ++k)
    {
        const P2d p = h(k);
        mx += p.x;
        my += p.y;
    }
    mx /= n;
    my /= n;
    double cxx = 0.0, cxy = 0.0, cyy = 0.0;
    for (int k = 0; k < n; ++k)
    {
        const P2d p = h(k);
        const double dx = p.x - mx, dy = p.y - my;
        cxx += dx * dx;
        cxy += dx * dy;
        cyy += dy * dy;
    }
    const double dn = static_cast<double>(static_cast<std::uint32_t>(n) - 1u);
    const double a00 = cxx / dn, a01 = cxy / dn, a10 = cxy / dn, a11 = cyy / dn;
    double v[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    {
        // Eigen 3.4 JacobiSVD on a 2 x 2 matrix (JacobiSVD.h real_2x2_jacobi_svd, Jacobi.h makeJacobi)
        const double precision = 2.0 * 2.220446049250313e-16;
        const double tiny = 2.2250738585072014e-308;
        double scale = fmax(fmax(fabs(a00), fabs(a01)), fmax(fabs(a10), fabs(a11)));
        if (!isfinite(scale))
        {
            *out = box; // svd_.info() != Success
            return;
        }
        if (scale == 0.0)
        {
            scale = 1.0;
        }
        double w[2][2] = {{a00 / scale, a01 / scale}, {a10 / scale, a11 / scale}};
        double max_diag = fmax(fabs(w[0][0]), fabs(w[1][1]));
        bool finished = false;
        const int p = 1, q = 0;
        while (!finished)
        {
            finished = true;
            const double threshold = fmax(tiny, precision * max_diag);
            if (fabs(w[p][q]) > threshold || fabs(w[q][p]) > threshold)
            {
                finished = false;
                double m00 = w[p][p], m01 = w[p][q], m10 = w[q][p], m11 = w[q][q];
                double r1c = 1.0, r1s = 0.0;
                const double t = m00 + m11, d = m10 - m01;
                if (fabs(d) >= tiny)
                {
                    const double u = t / d;
                    const double tmp = sqrt(1.0 + u * u);
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
                    const double deno = 2.0 * fabs(m01);
                    if (!(deno < tiny))
                    {
                        const double tau = (m00 - m11) / deno;
                        const double ww = sqrt(tau * tau + 1.0);
                        const double tt = tau > 0.0 ? 1.0 / (tau + ww) : 1.0 / (tau - ww);
                        const double sign_t = tt > 0.0 ? 1.0 : -1.0;
                        const double nn = 1.0 / sqrt(tt * tt + 1.0);
                        js = -sign_t * (m01 / fabs(m01)) * fabs(tt) * nn;
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
#pragma unroll
                for (int i = 0; i < 2; ++i)
                {
                    const double xp = w[i][p], xq = w[i][q];
                    w[i][p] = jc * xp - js * xq;
                    w[i][q] = js * xp + jc * xq;
                    const double vp = v[i][p], vq = v[i][q];
                    v[i][p] = jc * vp - js * vq;
                    v[i][q] = js * vp + jc * vq;
                }
                max_diag = fmax(max_diag, fmax(fabs(w[p][p]), fabs(w[q][q])));
            }
        }
        if (fabs(w[1][1]) > fabs(w[0][0]))
        {
            double t0 = v[0][0];
            v[0][0] = v[0][1];
            v[0][1] = t0;
            t0 = v[1][0];
            v[1][0] = v[1][1];
            v[1][1] = t0;
        }
    }
    double min_x = 1.7976931348623157e308, max_x = -1.7976931348623157e308;
    double min_y = min_x, max_y = max_x;
    for (int k = 0; k < n; ++k)
    {
        const P2d pt = h(k);
        const double dx = pt.x - mx, dy = pt.y - my;
        const double rx = dx * v[0][0] + dy * v[1][0];
        const double ry = dx * v[0][1] + dy * v[1][1];
        min_x = fmin(min_x, rx);
        max_x = fmax(max_x, rx);
        min_y = fmin(min_y, ry);
        max_y = fmax(max_y, ry);
    }
    const double cx[4] = {min_x, max_x, max_x, min_x}, cy[4] = {min_y, min_y, max_y, max_y};
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        box.c[2 * k] = (cx[k] * v[0][0] + cy[k] * v[0][1]) + mx;
        box.c[2 * k + 1] = (cx[k] * v[1][0] + cy[k] * v[1][1]) + my;
    }
    box.area = __double2float_rn((max_x - min_x) * (max_y - min_y));
    box.angle = __double2float_rn(atan2(v[1][0], v[0][0]));
    box.valid = 1;
    *out = box;
}

template <class Pts>
__device__ void obb_one(const Pts& h, int n, int method, ObbBox* out, int2* s_pairs, double* s_area)
{
    if (method == 0)
    {
        obb_calipers(h, n, out, s_pairs, s_area);
    }
    else if (lane_id() == 0)
    {
        obb_pca(h, n, out);
    }
}
} // namespace

// pipeline: boxes of every cluster hull of every frame (hull_xy / hull_off from launch_hulls)
__global__ void __launch_bounds__(kObbWarps * 32) k_obb_frames(Dev d, int method)
{
    __shared__ int2 s_pairs[kObbWarps][32];
    __shared__ double s_area[kObbWarps][128];
    const std::uint32_t f = blockIdx.y + d.f0;
    const std::uint32_t K = d.n_clusters[f];
    const std::uint32_t warp = threadIdx.x >> 5;
    const std::size_t o = static_cast<std::size_t>(f) * d.cap;
    const std::uint32_t* hoff = d.hull_off + static_cast<std::size_t>(f) * (d.cap + 1);
    for (std::uint32_t c = blockIdx.x * kObbWarps + warp; c < K; c += gridDim.x * kObbWarps)
    {
        const std::uint32_t a = hoff[c], n = hoff[c + 1] - a;
        obb_one(HullF2{d.hull_xy + o + a}, static_cast<int>(n), method, d.boxes + o + c, s_pairs[warp], s_area[warp]);
        __syncwarp();
    }
}

// standalone entry point: K hulls of doubles with explicit offsets (device buffers)
__global__ void __launch_bounds__(kObbWarps * 32)
    k_obb_hulls(const double2* __restrict__ xy, const std::uint32_t* __restrict__ off, std::uint32_t K, int method,
                ObbBox* __restrict__ out)
{
    __shared__ int2 s_pairs[kObbWarps][32];
    __shared__ double s_area[kObbWarps][128];
    const std::uint32_t warp = threadIdx.x >> 5;
    for (std::uint32_t c = blockIdx.x * kObbWarps + warp; c < K; c += gridDim.x * kObbWarps)
    {
        const std::uint32_t a = off[c], n = off[c + 1] - a;
        obb_one(HullD2{xy + a}, static_cast<int>(n), method, out + c, s_pairs[warp], s_area[warp]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Vehicle shape matching (src/processor/src/processor.cpp:680-757 with the tables of processor.hpp:60-192; the node
// keeps it behind `perform_polygon_simplification = false`). Per cluster, in the node's order of tests: a hull of
// >= 3 vertices, more than 150 points, a height inside the range any vehicle class allows, the polygon volume
// (shoelace area x height) inside the overall bounds, a valid box whose area holds the hull area to more than 0.4,
// then the first of the five classes (compact, sedan, SUV, truck, minivan; dimensions widened by the tolerances)
// whose length / width / height / volume / area windows all hold. All in double, in the node's expression order.
// ------------------------------------------------------------------------------------------
struct VehicleTables
{
    double dim[5][6];    // min / max length, width, height (adjusted)
    double bound[5][4];  // min / max volume, min / max area
    double min_height, max_height, min_volume, max_volume, min_iou;
    std::uint32_t min_points;
};

__global__ void __launch_bounds__(128)
    k_vehicle_match(const double2* __restrict__ xy, const std::uint32_t* __restrict__ off, std::uint32_t K,
                    const double2* __restrict__ zmm, const std::uint32_t* __restrict__ sizes, const ObbBox* __restrict__ boxes,
                    VehicleTables t, std::int32_t* __restrict__ cls, double* __restrict__ area_out)
{
    const std::uint32_t c = blockIdx.x * 128u + threadIdx.x;
    if (c >= K)
    {
        return;
    }
    const std::uint32_t a = off[c], n = off[c + 1] - a;
    // polygonArea (polygonizer.hpp:185-198): shoelace, terms added in vertex order, the closing term last
    double area = 0.0;
    if (n > 2)
    {
        for (std::uint32_t i = 0; i + 1 < n; ++i)
        {
            area += (xy[a + i].x * xy[a + i + 1].y) - (xy[a + i + 1].x * xy[a + i].y);
        }
        area += (xy[a + n - 1].x * xy[a].y) - (xy[a].x * xy[a + n - 1].y);
    }
    area = fabs(area) * 0.5;
    area_out[c] = area;
    std::int32_t found = -1;
    const double height = zmm[c].y - zmm[c].x;
    if (n >= 3 && sizes[c] > t.min_points && height > t.min_height && height < t.max_height)
    {
        const double volume = area * height;
        const ObbBox b = boxes[c];
        if (volume > t.min_volume && volume < t.max_volume && b.valid != 0)
        {
            const double iou = area / static_cast<double>(b.area);
            if (iou > t.min_iou)
            {
                const double dx1 = b.c[0] - b.c[2], dy1 = b.c[1] - b.c[3];
                const double dx2 = b.c[2] - b.c[4], dy2 = b.c[3] - b.c[5];
                const double e1 = sqrt(dx1 * dx1 + dy1 * dy1), e2 = sqrt(dx2 * dx2 + dy2 * dy2);
                const double len = fmax(e1, e2), wid = fmin(e1, e2);
                for (int v = 0; v < 5 && found < 0; ++v)
                {
                    if (len > t.dim[v][0] && len < t.dim[v][1] && wid > t.dim[v][2] && wid < t.dim[v][3] && height > t.dim[v][4] &&
                        height < t.dim[v][5] && volume > t.bound[v][0] && volume < t.bound[v][1] && area > t.bound[v][2] &&
                        area < t.bound[v][3])
                    {
                        found = v;
                    }
                }
            }
        }
    }
    cls[c] = found;
}

void launch_vehicle_match(Ctx* c, const double2* xy, const std::uint32_t* off, std::uint32_t K, const double2* zmm,
                          const std::uint32_t* sizes, const ObbBox* boxes, std::int32_t* cls, double* area_out)
{
    // processor.hpp:60-192, evaluated in double exactly as its constexpr lambdas do
    const double base[5][6] = {{4.3, 4.6, 1.6, 1.9, 1.4, 1.5}, {4.6, 5.0, 1.6, 1.9, 1.4, 1.5}, {4.6, 5.2, 1.7, 2.1, 1.7, 1.8},
                               {5.2, 5.8, 1.9, 2.2, 1.8, 2.0}, {4.9, 5.2, 1.7, 2.1, 1.7, 1.8}};
    const double tol[3] = {0.8, 0.5, 0.5};
    VehicleTables t{};
    t.min_height = 1e300;
    t.max_height = -1e300;
    t.min_volume = 1e300;
    t.max_volume = -1e300;
    for (int v = 0; v < 5; ++v)
    {
        for (int q = 0; q < 3; ++q)
        {
            t.dim[v][2 * q] = base[v][2 * q] - tol[q];
            t.dim[v][2 * q + 1] = base[v][2 * q + 1] + tol[q];
        }
        t.bound[v][0] = t.dim[v][0] * t.dim[v][2] * t.dim[v][4];
        t.bound[v][1] = t.dim[v][1] * t.dim[v][3] * t.dim[v][5];
        t.bound[v][2] = t.dim[v][0] * t.dim[v][2];
        t.bound[v][3] = t.dim[v][1] * t.dim[v][3];
        t.min_height = std::min(t.min_height, t.dim[v][4]);
        t.max_height = std::max(t.max_height, t.dim[v][5]);
        t.min_volume = std::min(t.min_volume, t.bound[v][0]);
        t.max_volume = std::max(t.max_volume, t.bound[v][1]);
    }
    t.min_iou = 0.4;
    t.min_points = 150;
    k_vehicle_match<<<(K + 127) / 128, 128, 0, c->stream>>>(xy, off, K, zmm, sizes, boxes, t, cls, area_out);
    mark(c, "vehicle_match");
}

void launch_boxes(Ctx* c, std::uint32_t nf, int method)
{
    k_obb_frames<<<dim3(16, nf), kObbWarps * 32, 0, c->stream>>>(c->d, method);
    mark(c, "obb_frames");
}

void launch_boxes_hulls(Ctx* c, const double2* xy, const std::uint32_t* off, std::uint32_t K, int method, ObbBox* out)
{
    const std::uint32_t grid = std::min<std::uint32_t>((K + kObbWarps - 1) / kObbWarps, 148u * 8u);
    k_obb_hulls<<<std::max(grid, 1u), kObbWarps * 32, 0, c->stream>>>(xy, off, K, method, out);
    mark(c, "obb_hulls");
}
} // namespace lpl
