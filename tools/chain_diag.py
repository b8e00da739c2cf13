"""Diagnostics: chained pipeline (with DROR) on selected KITTI frames, GPU vs port, plane by plane."""
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

sel = [int(v) for v in sys.argv[1:]] or [28, 50, 75, 77, 140]
fr = F.load_pack()
port = PortOracle()
c = lpl.Context(0, max_points=131072, max_frames=len(sel))
c.cluster_config(**NODE_CLUSTER_CFG)
frames = [fr[i] for i in sel]
nf = c.upload(frames)
c.run(nf, lpl.STAGE_ALL)
c.sync(nf)
for k, i in enumerate(sel):
    got = c.download(k)
    exp = parity.oracle_chain(port, frames[k], dror=True)
    rep = parity.chain_report(got, exp)
    print("frame", i, rep)
    for name in ("noise", "labels"):
        a, b = np.asarray(got[name]), np.asarray(exp[name])
        if a.shape == b.shape and (a != b).any():
            w = np.flatnonzero(a != b)
            print("  ", name, "differs at", w[:10], "gpu", a[w[:10]], "exp", b[w[:10]])
            for j in w[:4]:
                p = frames[k][j]
                print("     pt", j, p, "range_xy", float(np.hypot(p[0], p[1])), "ring", got["ring"][j], "noise g/e", got["noise"][j], exp["noise"][j])
    dbg = c.debug_segment(k)
    print("   gpu plane", dbg["plane"], dbg["best_inliers"], "cand", dbg["n_candidates"], "queued", dbg["n_queued"], "status", dbg["status"])
    keep = np.flatnonzero(exp["noise"] == 0)
    l, img, d = port.segment(np.ascontiguousarray(frames[k][keep]), exp["ring"][keep], want_image=True, want_debug=True)
    print("   port plane", d["plane"], d["best_inliers"], "cand", d["n_candidates"], "queued", d["n_queued"])
    print("   elev diff", int((d["elevation"] != dbg["elevation"]).sum()))
