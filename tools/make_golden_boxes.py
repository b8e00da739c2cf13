"""Generate tests/golden/kitti_polygonizer.npz from the UNMODIFIED reference polygonizer.cpp
(oracle/_ref, compiled against oracle/shim/Eigen): for the clusters of the two golden KITTI frames,
the reference's own convexHull vertices (per-label gather as in processor.cpp:627-658), its
antipodal pairs and its rotating-calipers / PCA boxes. The hulls stored by tools/make_golden.py
came from the restated hull; this file pins that restatement to the reference as well.
The PCA boxes depend on the shim's JacobiSVD restatement (Eigen is absent): unpinned, kept for
regression only."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import RefOracle  # noqa: E402
from tools.frames import GOLDEN_DIR, load_golden  # noqa: E402


def main():
    ref = RefOracle()
    out = {}
    for name in ("kitti_f000", "kitti_f100"):
        g = load_golden(name)
        obs = g["pts"][g["labels"] == 2]
        cl = g["cluster_labels"].astype(np.int32)
        K = int(cl.max()) + 1
        off, hxy, pairs_n, boxes = [0], [], [], []
        for k in range(K):
            xy = obs[cl == k][:, :2].astype(np.float64)
            idx = ref.convex_hull(xy)
            h = xy[idx]
            hxy.append(h)
            off.append(off[-1] + len(h))
            pairs_n.append(len(ref.antipodal_pairs(h)))
            boxes.append(np.concatenate([ref.bounding_box(h, 0), ref.bounding_box(h, 1)]))
        out[name + "_hull_offsets"] = np.asarray(off, np.uint32)
        out[name + "_hull_xy"] = np.concatenate(hxy).astype(np.float32)
        out[name + "_pairs"] = np.asarray(pairs_n, np.uint16)
        out[name + "_boxes"] = np.asarray(boxes, np.float64)  # [K][22]: calipers[11], pca[11]
    np.savez_compressed(os.path.join(GOLDEN_DIR, "kitti_polygonizer.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
