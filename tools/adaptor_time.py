"""Latency of the drop-in C++ classes through the node's call sequence (processor.cpp:552-663) on one KITTI frame,
next to the reference CPU library timed call by call on the same host (oracle/_ref).
usage (on a GPU box): python tools/adaptor_time.py [frame index] [reps]"""
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_processing_v2_b200 as lpl  # noqa: E402
from oracle.oracle import NODE_CLUSTER_CFG, PortOracle, RefOracle, have_ref  # noqa: E402
from tools import frames as F  # noqa: E402


def main():
    fi = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    pts = F.load_pack(limit=fi + 1)[fi] if F.have_pack() else F.load_golden("kitti_f000")["pts"]
    port = PortOracle()
    ring = port.ring_partition(pts)
    lpl.load_library()
    pkg = os.path.dirname(lpl.SO_PATH)
    exe = os.path.join(ROOT, "tests", "native", "adaptor_main")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "oracle", "shim"), os.path.join(ROOT, "tests", "native", "adaptor_main.cpp"), "-o", exe,
           "-L", pkg, "-llpl_b200", "-Wl,-rpath," + pkg]
    subprocess.run(cmd, check=True)
    fin = "/tmp/adaptor_in.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<I", pts.shape[0]))
        f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        f.write(np.ascontiguousarray(ring, np.uint16).tobytes())
    out = subprocess.run([exe, "--time", fin, str(reps)], capture_output=True, text=True)
    if out.returncode != 0:
        print(out.stderr)
        raise SystemExit(out.returncode)
    gpu = json.loads(out.stdout.strip().splitlines()[-1])
    cpu = None
    if have_ref():
        ref = RefOracle()
        ref.cluster_config(**NODE_CLUSTER_CFG)
        labels = ref.segment(pts, ring)
        obs = np.ascontiguousarray(pts[labels == 2])
        cl = ref.cluster(obs)
        t0 = time.perf_counter()
        for _ in range(3):
            port.cluster_hulls(obs, cl)   # gather + convexHull per label, as the node does
        t_h = (time.perf_counter() - t0) / 3 * 1e3
        cpu = {"noise_filter": ref.dror_timed(pts, 1) * 1e3, "segment": ref.segment_timed(pts, ring, 3) * 1e3 / 3,
               "cluster": ref.cluster_timed(obs, 3) * 1e3 / 3, "hulls_gather_and_convexHull": t_h,
               "what": "reference library (oracle/_ref, unmodified sources) call by call, one thread; the *_timed helpers return seconds for the given repetitions"}
    print(json.dumps({"frame": fi, "adaptors_on_gpu": gpu, "reference_cpu_ms": cpu}))


if __name__ == "__main__":
    main()
