"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref) run in this container.

  tests/golden/kitti_f000.npz, kitti_f100.npz : two KITTI frames (delta-coded mm, see
      tools/frames.py) with the reference's per-point outputs for every stage
  tests/golden/kitti154_summary.json          : per-frame counts + FNV hashes for all 154 frames

The reference's own test-suite holds no vectors for this path, so these
fixtures - outputs of the reference code itself on its own data - are the pin.
Chained-pipeline convention: ring partition on all points -> segment all points
(the node never wires DROR in) -> OBSTACLE points in cloud order -> cluster -> per-label hulls.
DROR is recorded stage-wise on the raw cloud (exact semantics + the as-is count).
"""
from __future__ import annotations

import glob
import json
import os
import sys
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import PortOracle, RefOracle, label_hash, read_pcd_xyzi  # noqa: E402
from tools.frames import GOLDEN_DIR, encode_xyz_mm  # noqa: E402

FULL = (0, 100)


def hash_i32(a) -> str:
    return label_hash(np.asarray(a).astype(np.int64).astype(np.uint32))


def work(path):
    idx = int(os.path.basename(path)[:-4])
    ref, port = RefOracle(), PortOracle()
    pts = read_pcd_xyzi(path)
    ring = port.ring_partition(pts)  # restated (dataloader.cpp:68-137 is a private ROS-node method)
    labels, img = ref.segment(pts, ring, want_image=True)
    inter = ref.segment_intermediates()
    obs = pts[labels == 2]
    clabels, dims = ref.cluster(obs, want_dims=True)
    off, hxy, hidx, zmm = port.cluster_hulls(obs, clabels)  # restated polygonizer.cpp:33-91
    dror_exact = ref.dror(pts, mode="exact")
    dror_as_is = ref.dror(pts, mode="as_is")
    labels_noring = ref.segment(pts, None)
    summary = dict(
        frame=idx, n=int(pts.shape[0]),
        ring_min=int(ring.min()), ring_max=int(ring.max()), ring_hash=label_hash(ring),
        ground=int((labels == 1).sum()), obstacle=int((labels == 2).sum()), unknown=int((labels == 0).sum()),
        label_hash=label_hash(labels), image_hash=label_hash(img.reshape(-1)[::7]),
        noring_label_hash=label_hash(labels_noring),
        ransac_candidates=int(inter["ransac_candidates"]),
        clusters=int(clabels.max() + 1) if clabels.size else 0, unclustered=int((clabels < 0).sum()),
        cluster_hash=hash_i32(clabels), voxel_dims=[int(v) for v in dims],
        hull_vertices=int(hxy.shape[0]), hull_offsets_hash=label_hash(off),
        hull_xy_hash=label_hash(hxy.astype(np.float32).view(np.uint32).reshape(-1)),
        dror_noise_exact=int(dror_exact.sum()), dror_noise_as_is=int(dror_as_is.sum()),
        dror_hash=label_hash(dror_exact),
    )
    if idx in FULL:
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, f"kitti_f{idx:03d}.npz"),
            delta=encode_xyz_mm(pts[:, :3])[0], negzero=encode_xyz_mm(pts[:, :3])[1], ring=ring.astype(np.uint8), labels=labels.astype(np.uint8),
            labels_noring=labels_noring.astype(np.uint8),
            image=np.packbits(img.reshape(-1) > 0), elevation=inter["elevation"],
            cluster_labels=clabels.astype(np.int16), voxel_dims=dims,
            hull_offsets=off, hull_xy=hxy.astype(np.float32), zminmax=zmm.astype(np.float32),
            dror_exact=np.packbits(dror_exact), dror_as_is=np.packbits(dror_as_is))
    return summary


def main():
    files = sorted(glob.glob("/root/reference/data/*.pcd"))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with Pool(8) as p:
        res = p.map(work, files)
    with open(os.path.join(GOLDEN_DIR, "kitti154_summary.json"), "w") as f:
        json.dump(dict(source="oracle/_ref (unmodified reference sources) via tools/make_golden.py",
                       frames=res), f, indent=0)
    tot = lambda k: sum(r[k] for r in res)  # noqa: E731
    print(len(res), "frames; clusters/frame", tot("clusters") / len(res), "hull vertices/frame",
          tot("hull_vertices") / len(res), "obstacle/frame", tot("obstacle") / len(res))


if __name__ == "__main__":
    main()
