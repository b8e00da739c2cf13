"""Golden digests of the CHAINED pipeline with DROR for all 154 KITTI frames, from the UNMODIFIED reference
(oracle/_ref) run in this container -> tests/golden/kitti154_chain_dror.json.

Chain: ring partition on all N points -> DROR on all N (exact = stack-drained semantics of the
reference's own KDTree::radius_search, hazard H1) -> stable compaction of the VALID points (ring kept) ->
Segmenter::segment -> stable compaction of the OBSTACLE points -> Clusterer::cluster -> per-label gather ->
convexHull. Digests are sha1 over the little-endian bytes of the arrays in INPUT index space (labels of NOISE
points are 0), so a test can hash what lpl_pipeline_download returns. The as-is DROR result (the reference as
built) is recorded next to it as the per-frame count of points whose final label differs.
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import PortOracle, RefOracle, read_pcd_xyzi  # noqa: E402
from tools.frames import GOLDEN_DIR  # noqa: E402


def sha(a, dtype) -> str:
    return hashlib.sha1(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


def chain(ref, port, pts, ring, noise):
    keep = np.flatnonzero(noise == 0)
    lv = ref.segment(np.ascontiguousarray(pts[keep]), np.ascontiguousarray(ring[keep]))
    labels = np.zeros(pts.shape[0], np.uint8)
    labels[keep] = lv
    obs = np.ascontiguousarray(pts[keep[lv == 2]])
    cl = ref.cluster(obs)
    off, hxy, hidx, zmm = port.cluster_hulls(obs, cl)  # == the reference's convexHull (tests/golden/kitti_polygonizer.npz)
    return labels, cl, off, hxy, zmm


def work(path):
    idx = int(os.path.basename(path)[:-4])
    ref, port = RefOracle(), PortOracle()
    pts = read_pcd_xyzi(path)
    ring = port.ring_partition(pts)
    noise = ref.dror(pts, mode="exact")
    labels, cl, off, hxy, zmm = chain(ref, port, pts, ring, noise)
    noise_as_is = ref.dror(pts, mode="as_is")
    labels_as_is = chain(ref, port, pts, ring, noise_as_is)[0]
    return dict(frame=idx, n=int(pts.shape[0]), noise=int(noise.sum()), noise_sha1=sha(noise, np.uint8),
                labels_sha1=sha(labels, np.uint8), obstacles=int((labels == 2).sum()),
                clusters=int(cl.max() + 1) if cl.size else 0, cluster_sha1=sha(cl, np.int32),
                hull_vertices=int(hxy.shape[0]), hull_offsets_sha1=sha(off, np.uint32), hull_xy_sha1=sha(hxy, np.float32),
                zminmax_sha1=sha(zmm, np.float32),
                as_is_noise=int(noise_as_is.sum()), as_is_noise_delta=int((noise_as_is != noise).sum()),
                as_is_label_delta=int((labels_as_is != labels).sum()))


def main():
    files = sorted(glob.glob("/root/reference/data/*.pcd"))
    with Pool(8) as p:
        res = p.map(work, files)
    with open(os.path.join(GOLDEN_DIR, "kitti154_chain_dror.json"), "w") as f:
        json.dump(dict(source="oracle/_ref (unmodified reference sources), DROR exact, via tools/make_golden_chain.py",
                       frames=res), f, indent=0)
    tot = lambda k: sum(r[k] for r in res)  # noqa: E731
    print(len(res), "frames; noise exact/as-is per frame", tot("noise") / len(res), tot("as_is_noise") / len(res),
          "label delta per frame", tot("as_is_label_delta") / len(res), "clusters/frame", tot("clusters") / len(res))


if __name__ == "__main__":
    main()
