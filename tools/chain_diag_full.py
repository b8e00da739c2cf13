"""Diagnostics: the 154-frame chained batch (packed xyz upload, packed download) against the chain golden."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lidar_processing_v2_b200 as lpl  # noqa: E402
import parity  # noqa: E402
from oracle.oracle import NODE_CLUSTER_CFG, PortOracle  # noqa: E402
from tools import frames as F  # noqa: E402


def sha(a, dt):
    return hashlib.sha1(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


gold = json.load(open(os.path.join(F.GOLDEN_DIR, "kitti154_chain_dror.json")))["frames"]
frames = F.load_pack()
port = PortOracle()
c = lpl.Context(0, max_points=max(f.shape[0] for f in frames), max_frames=len(frames))
c.cluster_config(**NODE_CLUSTER_CFG)
mode = sys.argv[1] if len(sys.argv) > 1 else "packed"
for rep_ in range(2):
    if mode == "packed":
        xyz = np.ascontiguousarray(np.concatenate(frames)[:, :3])
        nf = c.upload_packed_xyz(xyz, [f.shape[0] for f in frames])
    else:
        nf = c.upload(frames)
    c.run(nf, lpl.STAGE_ALL)
    names = ("labels_u8", "noise", "cluster_labels", "hull_offsets", "hull_xy", "zminmax")
    bufs = lpl.PackedBuffers(nf, nf * 131072 * 8, want=names)
    counts = c.download_packed(nf, bufs)
    key = dict(labels_u8=("labels_sha1", np.uint8), noise=("noise_sha1", np.uint8), cluster_labels=("cluster_sha1", np.int32),
               hull_offsets=("hull_offsets_sha1", np.uint32), hull_xy=("hull_xy_sha1", np.float32), zminmax=("zminmax_sha1", np.float32))
    bad = {}
    for f, g in enumerate(gold):
        for nm in names:
            if sha(bufs.frame(nm, f), key[nm][1]) != g[key[nm][0]]:
                bad.setdefault(f, []).append(nm)
    print("pass", rep_, "bad frames (packed download vs golden):", bad)
    for f in list(bad)[:6]:
        got = c.download(f)
        exp = parity.oracle_chain(port, frames[f], dror=True)
        print("  frame", f, "per-frame download vs port:", parity.chain_report(got, exp))
        for nm in bad[f]:
            a = bufs.frame(nm, f)
            b = got["labels"].astype(np.uint8) if nm == "labels_u8" else got[nm]
            print("    packed vs per-frame download", nm, "equal" if np.array_equal(a, b) else f"DIFFER {a.shape} {np.asarray(b).shape}")
        for name in ("noise", "labels", "cluster_labels"):
            a, b = np.asarray(got[name]), np.asarray(exp[name])
            if a.shape == b.shape and (a != b).any():
                w = np.flatnonzero(a != b)
                print("    ", name, len(w), "differ at", w[:8], "gpu", a[w[:8]], "exp", b[w[:8]])
    bufs.close()
