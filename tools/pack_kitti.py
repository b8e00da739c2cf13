"""Re-encode /root/reference/data/*.pcd into data/kitti154.npz (see tools/frames.py).

Run in the build container (where /root/reference exists); the output is git-ignored but travels
to the GPU box with the repo snapshot. Intensity is dropped: no stage of the hot path reads it.
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import read_pcd_xyzi  # noqa: E402
from tools.frames import PACK_PATH, decode_xyz_mm, encode_xyz_mm  # noqa: E402


def main(src="/root/reference/data"):
    files = sorted(glob.glob(os.path.join(src, "*.pcd")))
    if not files:
        print("no PCD files under", src)
        return 1
    arrays = {}
    total = 0
    for f in files:
        p = read_pcd_xyzi(f)
        d, nz = encode_xyz_mm(p[:, :3])
        assert np.array_equal(decode_xyz_mm(d, nz)[:, :3].view(np.uint32), p[:, :3].view(np.uint32))
        arrays["f" + os.path.basename(f)[:-4]] = d
        arrays["z" + os.path.basename(f)[:-4]] = nz
        total += p.shape[0]
    os.makedirs(os.path.dirname(PACK_PATH), exist_ok=True)
    # LZMA members (np.load reads them transparently): ~30 % smaller than deflate, and the pack is
    # re-sent to the GPU box with every gpurun call
    import io
    import zipfile

    with zipfile.ZipFile(PACK_PATH, "w", compression=zipfile.ZIP_LZMA) as zf:
        for k, v in arrays.items():
            buf = io.BytesIO()
            np.save(buf, np.ascontiguousarray(v.T) if v.ndim == 2 else v)
            zf.writestr(k + ".npy", buf.getvalue())
    print(f"{len(files)} frames, {total} points -> {PACK_PATH} ({os.path.getsize(PACK_PATH) / 1e6:.1f} MB)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
