"""Frame sources for tests and bench.py (test / measurement infrastructure, not product code).

* KITTI pack: the reference's 154 HDL-64E frames (`/root/reference/data/*.pcd`) are exactly
  millimetre-quantised (x == float32(k / 1000) for an integer k), so `tools/pack_kitti.py`
  re-encodes them as delta-coded int32 millimetres in one compressed `data/kitti154.npz`
  (git-ignored; it travels to the GPU box like the built .so files). Two frames are committed as
  fixtures under tests/golden/.
* Synthetic frames: seeded ray-cast scenes of the shapes BASELINE.json names (HDL-64E 64 x 2048,
  128 x 2048, unorganised 2 M-point clouds), used when the pack is absent and for the scaling runs.
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACK_PATH = os.path.join(ROOT, "data", "kitti154.npz")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def encode_xyz_mm(xyz: np.ndarray):
    """(n, 3) float32 -> (int32 millimetre deltas, flat indices of the -0.0 entries)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    k = np.round(xyz.astype(np.float64) * 1000.0).astype(np.int64)
    back = (k.astype(np.float64) / 1000.0).astype(np.float32)
    if not np.array_equal(back, xyz):
        raise ValueError("cloud is not exactly millimetre-quantised")
    negzero = np.flatnonzero(xyz.reshape(-1).view(np.uint32) == 0x80000000).astype(np.uint32)
    return np.diff(k, axis=0, prepend=0).astype(np.int32), negzero


def decode_xyz_mm(delta: np.ndarray, negzero=None) -> np.ndarray:
    """(n, 3) int32 deltas (+ the -0.0 positions) -> (n, 4) float32 x, y, z, 0, bit-identical to the
    PCD floats (KITTI frames do contain negative zeros)."""
    k = np.cumsum(delta.astype(np.int64), axis=0)
    xyz = (k.astype(np.float64) / 1000.0).astype(np.float32)
    if negzero is not None and len(negzero):
        xyz.reshape(-1).view(np.uint32)[np.asarray(negzero, np.int64)] = 0x80000000
    out = np.zeros((k.shape[0], 4), np.float32)
    out[:, :3] = xyz
    return out


def have_pack() -> bool:
    return os.path.exists(PACK_PATH)


def load_pack(limit: int | None = None) -> list[np.ndarray]:
    z = np.load(PACK_PATH)
    names = sorted(k for k in z.files if k.startswith("f"))
    if limit is not None:
        names = names[:limit]
    # pack members are stored column-major (3, n) for compression
    return [decode_xyz_mm(np.ascontiguousarray(z[k].T), z["z" + k[1:]]) for k in names]


def load_golden(name: str) -> dict:
    """tests/golden/<name>.npz -> dict with 'pts' (n, 4) float32 and the expected outputs."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {k: z[k] for k in z.files}
    out["pts"] = decode_xyz_mm(out.pop("delta"), out.pop("negzero", None))
    return out


# ------------------------------------------------------------------------------------------
# synthetic scenes
# ------------------------------------------------------------------------------------------
def synth_scan(seed: int, beams: int = 64, cols: int = 2048, n_boxes: int = 60, n_poles: int = 40,
               n_walls: int = 2, dropout: float = 0.06, el_up_deg: float = 2.0,
               el_down_deg: float = -24.8, max_range: float = 100.0, extent: float = 80.0):
    """Ray-cast LiDAR sweep of a ground plane with boxes, poles and walls.

    Returns (pts (n, 4) float32 in firing order: top beam first, azimuth ascending; ring (n,) uint16
    = beam index, 0 = lowest). Coordinates are rounded to 1 mm like KITTI.
    """
    rng = np.random.default_rng(seed)
    el = np.deg2rad(el_down_deg + (np.arange(beams) + 0.5) * (el_up_deg - el_down_deg) / beams)
    az = 2.0 * np.pi * (np.arange(cols) + 0.5) / cols
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    dx = ce * np.cos(az)[None, :]
    dy = ce * np.sin(az)[None, :]
    dz = np.broadcast_to(se, dx.shape).copy()
    slope = 0.002
    # ground: z = -1.73 + slope * x
    den = dz - slope * dx
    t = np.where(den < -1e-6, -1.73 / np.where(den < -1e-6, den, -1.0), np.inf)

    def col_span(cx, cy, half):
        a = np.arctan2(cy, cx) % (2 * np.pi)
        r = max(np.hypot(cx, cy) - half, 0.5)
        da = np.arcsin(min(1.0, half * 1.5 / r)) + 2 * np.pi / cols
        c0 = int(np.floor((a - da) / (2 * np.pi) * cols))
        c1 = int(np.ceil((a + da) / (2 * np.pi) * cols))
        return np.arange(c0, c1 + 1) % cols

    def hit_box(cx, cy, sx, sy, h):
        z0 = -1.73 + slope * cx
        lo = np.array([cx - sx / 2, cy - sy / 2, z0])
        hi = np.array([cx + sx / 2, cy + sy / 2, z0 + h])
        cs = np.unique(col_span(cx, cy, 0.5 * np.hypot(sx, sy)))
        d = np.stack([dx[:, cs], dy[:, cs], dz[:, cs]], -1)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = lo / d
            t1 = hi / d
        tn = np.minimum(t0, t1).max(-1)
        tf = np.maximum(t0, t1).min(-1)
        ok = (tf >= tn) & (tn > 0.5)
        sub = t[:, cs]
        t[:, cs] = np.where(ok & (tn < sub), tn, sub)

    def hit_pole(cx, cy, r, h):
        z0 = -1.73 + slope * cx
        cs = np.unique(col_span(cx, cy, r))
        ux, uy = dx[:, cs], dy[:, cs]
        a = ux * ux + uy * uy
        b = -2 * (ux * cx + uy * cy)
        c = cx * cx + cy * cy - r * r
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            tn = (-b - np.sqrt(disc)) / (2 * a)
        z = tn * dz[:, cs]
        ok = (disc > 0) & (tn > 0.5) & (z >= z0) & (z <= z0 + h)
        sub = t[:, cs]
        t[:, cs] = np.where(ok & (tn < sub), tn, sub)

    def place():
        while True:
            cx, cy = rng.uniform(-extent, extent, 2)
            if np.hypot(cx, cy) > 4.0:
                return cx, cy

    for _ in range(n_boxes):
        cx, cy = place()
        hit_box(cx, cy, rng.uniform(0.5, 5.0), rng.uniform(0.5, 5.0), rng.uniform(0.5, 3.0))
    for _ in range(n_poles):
        cx, cy = place()
        hit_pole(cx, cy, rng.uniform(0.1, 0.3), rng.uniform(2.0, 6.0))
    for _ in range(n_walls):
        cx, cy = place()
        if rng.random() < 0.5:
            hit_box(cx, cy, rng.uniform(15, 40), 0.3, rng.uniform(2.0, 4.0))
        else:
            hit_box(cx, cy, 0.3, rng.uniform(15, 40), rng.uniform(2.0, 4.0))
    t = t + rng.normal(0.0, 0.02, t.shape)
    keep = np.isfinite(t) & (t < max_range) & (t > 1.0) & (rng.random(t.shape) >= dropout)
    order_b = np.arange(beams)[::-1]  # top beam first (KITTI firing order, dataloader.cpp:87)
    pts, ring = [], []
    for b in order_b:
        m = keep[b]
        tt = t[b, m]
        p = np.stack([dx[b, m] * tt, dy[b, m] * tt, dz[b, m] * tt], -1)
        pts.append(p)
        ring.append(np.full(p.shape[0], b, np.uint16))
    p = np.concatenate(pts)
    p = (np.round(p * 1000.0) / 1000.0).astype(np.float32)
    out = np.zeros((p.shape[0], 4), np.float32)
    out[:, :3] = p
    return out, np.concatenate(ring)


def synth_ring_walls(seed: int, beams: int = 64, cols: int = 2048,
                     radii=(4.5, 6.5, 9.0, 13.0, 19.0, 28.0), height: float = 0.6):
    """Concentric low walls around the sensor on flat ground: obstacle / ground boundaries that run
    along whole image rows, so the JCP queue holds complete rows of consecutive pixels (the longest
    dependency chains the sweep can meet). Returns (pts (n, 4) float32 in firing order, ring uint16)."""
    rng = np.random.default_rng(seed)
    el = np.deg2rad(-24.8 + (np.arange(beams) + 0.5) * 26.8 / beams)[::-1]
    az = 2.0 * np.pi * (np.arange(cols) + 0.5) / cols
    e, a = np.meshgrid(el, az, indexing="ij")
    t = np.where(e < 0, -1.73 / np.sin(np.minimum(e, -1e-3)), 200.0)
    for r in radii:
        tk = r / np.cos(e)
        zk = tk * np.sin(e)
        t = np.where((zk >= -1.73) & (zk <= -1.73 + height) & (tk < t), tk, t)
    keep = (t < 100.0).ravel()
    zn = rng.normal(0, 0.02, a.shape)
    xyz = (t * np.cos(e) * np.cos(a), t * np.cos(e) * np.sin(a), t * np.sin(e) + zn)
    pts = np.zeros((beams * cols, 4), np.float32)
    for i, v in enumerate(xyz):
        pts[:, i] = np.round(v.ravel() * 1000.0) / 1000.0
    ring = np.repeat(np.arange(beams - 1, -1, -1), cols).astype(np.uint16)
    return np.ascontiguousarray(pts[keep]), np.ascontiguousarray(ring[keep])


def synth_unorganized(seed: int, n: int = 2_000_000, n_blobs: int = 5000, r_max: float = 120.0):
    """Unorganised cloud: 60 % ground disc (density ~ 1/r), 35 % small Gaussian blobs, 5 % uniform
    background (DROR targets). BASELINE.json config 5."""
    rng = np.random.default_rng(seed)
    ng = int(0.60 * n)
    nb = int(0.35 * n)
    nn = n - ng - nb
    r = rng.uniform(0.5, r_max, ng)
    a = rng.uniform(0, 2 * np.pi, ng)
    g = np.stack([r * np.cos(a), r * np.sin(a), -1.73 + rng.normal(0, 0.02, ng)], -1)
    cr = r_max * np.sqrt(rng.uniform(0.002, 1.0, n_blobs))
    ca = rng.uniform(0, 2 * np.pi, n_blobs)
    cz = rng.uniform(-1.5, 1.0, n_blobs)
    sig = rng.uniform(0.15, 0.5, n_blobs)
    which = rng.integers(0, n_blobs, nb)
    b = np.stack([cr[which] * np.cos(ca[which]), cr[which] * np.sin(ca[which]), cz[which]], -1)
    b = b + rng.normal(0, 1, (nb, 3)) * sig[which, None]
    u = np.stack([rng.uniform(-r_max, r_max, nn), rng.uniform(-r_max, r_max, nn), rng.uniform(-2.5, 3.0, nn)], -1)
    p = np.concatenate([g, b, u])
    p = p[rng.permutation(p.shape[0])]
    out = np.zeros((p.shape[0], 4), np.float32)
    out[:, :3] = (np.round(p * 1000.0) / 1000.0).astype(np.float32)
    return out
