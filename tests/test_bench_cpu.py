"""Host-side checks of the benchmark harness that need no GPU: the workloads BASELINE.json names are
generated deterministically with the shapes the bench line reports, and the reference arm prints the
contract's JSON line for the same batch (same `config`) our arm runs."""
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


def test_unorganised_workload_and_reference_arm_line():
    fr, name, _, opts = bench.load_frames(1, "cloud2m")
    assert name == "cloud2m" and fr[0].shape == (2_000_000, 4) and opts["stages"] == "ringless"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "synth64",
                          "--frames", "3", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    fr3, _, _, o3 = bench.load_frames(3, "synth64")
    # the same `config` object our arm prints for this batch
    assert line["config"] == bench.step_config("synth64", 3, sum(f.shape[0] for f in fr3), bench.stages_text(o3), 1)


def test_stage_and_kernel_byte_models():
    s = dict(N=1000, V=990, NB=950, M=400, K=5, HV=40, Q=90, C=300, U=50, NH=100, F=1, PX=131072, CELLS=18000)
    assert bench.stage_bytes(s)["S2 segment"] == 22 * 1000 and bench.stage_bytes(s)["S3 cluster"] == 20 * 400
    assert bench.stage_of("jcp_rows") == "S2 segment" and bench.stage_of("take_obstacles") == "S3 cluster"
    assert bench.stage_of("hull_thin") == "S4 hulls" and bench.stage_of("dror_query") == "S1 dror"
    # the JCP pre-pass is charged with the distinct pixels its neighbourhoods touch, not 25 gathers per pixel
    assert bench.algorithmic_bytes("jcp_pre", s) == 4 * 90 * 21 + 90 * 104  # index + point + code per distinct pixel


def test_ring_walls_scene_is_organised():
    pts, ring = F.synth_ring_walls(7)
    assert pts.shape[0] == ring.shape[0] and pts.shape[1] == 4 and ring.dtype == np.uint16
    assert np.all(np.diff(ring.astype(np.int32)) <= 0) and 50 <= ring.max() <= 63  # firing order: top beam first (beams above the horizon return nothing)
    r = np.hypot(pts[:, 0], pts[:, 1])
    assert r.min() > 3.0 and r.max() < 100.0
