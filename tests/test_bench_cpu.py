"""Host-side checks of the benchmark harness that need no GPU: the workloads BASELINE.json names are
generated deterministically with the shapes the bench line reports, and the reference arm states why it
cannot run the one workload the reference library cannot hold."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from tools import frames as F  # noqa: E402


def test_synthetic_workload_shapes():
    fr, name, _, opts = bench.load_frames(3, "synth64")
    assert name == "synth64" and len(fr) == 3 and opts["image_height"] == 64 and opts["rings"] is None
    assert all(f.dtype == np.float32 and f.shape[1] == 4 and 100_000 < f.shape[0] < 131_072 for f in fr)
    fr, name, _, opts = bench.load_frames(2, "synth128")
    assert name == "synth128" and opts["image_height"] == 128 and opts["stages"] == "ring_field"
    assert len(opts["rings"]) == 2 and opts["rings"][0].shape[0] == fr[0].shape[0]
    assert int(opts["rings"][0].max()) == 127 and 230_000 < fr[0].shape[0] < 262_144
    # same seed, same frame
    again, _, _, _ = bench.load_frames(2, "synth128")
    assert np.array_equal(again[1], fr[1])


def test_unorganised_workload_and_reference_arm_note():
    fr, name, _, opts = bench.load_frames(1, "cloud2m")
    assert name == "cloud2m" and fr[0].shape == (2_000_000, 4) and opts["cpu"] is False and opts["stages"] == "ringless"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cloud2m",
                          "--frames", "1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line


def test_ring_walls_scene_is_organised():
    pts, ring = F.synth_ring_walls(7)
    assert pts.shape[0] == ring.shape[0] and pts.shape[1] == 4 and ring.dtype == np.uint16
    assert np.all(np.diff(ring.astype(np.int32)) <= 0) and 50 <= ring.max() <= 63  # firing order: top beam first (beams above the horizon return nothing)
    r = np.hypot(pts[:, 0], pts[:, 1])
    assert r.min() > 3.0 and r.max() < 100.0
