"""TEST INFRASTRUCTURE - pins oracle.pixels_to_rays against the reference's own ray generation and writes
tests/golden/raygen.npz.  Runs only where /root/reference exists (the build container):

    python -m oracle.make_raygen_golden

The reference functions executed (unmodified, imported through oracle/ref_shim.py): camera_utils.pixel_coordinates,
camera_utils.pixels_to_rays (internal/camera_utils.py:L368-370, L448-557) and the cam_dirs / near / far / float32
cast of Dataset._make_ray_batch (internal/datasets.py:L414-476, restated inline because the Dataset class needs data
on disk)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, ucnerf_oracle as O  # noqa: E402


def cameras():
    """Three synthetic cameras shaped like the Waymo loader's (float32 intrinsics inverted in float32, float32
    poses, OpenGL convention; datasets.py:L672,L855-857) + one float64 camera like cast_pinhole_rays builds."""
    rng = np.random.default_rng(7)
    out = []
    for i, (w, h) in enumerate([(96, 64), (160, 120), (64, 48), (80, 60)]):
        f = 2000.0 * w / 1920.0 * (1 + 0.05 * i)
        K = np.array([[f, 0, w * 0.5 + 0.3 * i], [0, f * 1.01, h * 0.5 - 0.2 * i], [0, 0, 1]])
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        a, b, c, d = q
        R = np.array([[1 - 2 * (c * c + d * d), 2 * (b * c - d * a), 2 * (b * d + c * a)],
                      [2 * (b * c + d * a), 1 - 2 * (b * b + d * d), 2 * (c * d - b * a)],
                      [2 * (b * d - c * a), 2 * (c * d + b * a), 1 - 2 * (b * b + c * c)]])
        pose = np.concatenate([R, rng.uniform(-0.3, 0.3, (3, 1))], 1)
        if i < 3:
            K = K.astype(np.float32)
            pose = pose.astype(np.float32)
        out.append((np.linalg.inv(K), pose, w, h, 0.0 + 0.05 * i, 8.0 - i))
    return out


def main():
    assert ref_shim.available(), "needs /root/reference"
    ref_shim._install_stubs()
    ref_shim._install_grid_backend()
    sys.path.insert(0, ref_shim.REF_ROOT)
    from internal import camera_utils as cu
    gold = {}
    for i, (P, pose, w, h, near, far) in enumerate(cameras()):
        px, py = cu.pixel_coordinates(w, h)
        origins, directions, viewdirs, radii, imageplane = cu.pixels_to_rays(px, py, P, pose)
        bs = lambda v: np.broadcast_to(v, px.shape)[..., None]
        ref = dict(origins=origins, directions=directions, viewdirs=viewdirs, radii=radii, imageplane=imageplane,
                   near=bs(near), far=bs(far), cam_dirs=np.broadcast_to(-pose[:3, 2], directions.shape))
        ref = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in ref.items()}  # datasets.py:L476 .float()
        mine = O.pixels_to_rays(px, py, P, pose, near, far)
        for k, v in ref.items():
            assert mine[k].shape == v.shape, (k, mine[k].shape, v.shape)
            assert np.array_equal(mine[k], v), (i, k, np.abs(mine[k] - v).max())
        gold[f"cam{i}_pixtocam"] = np.asarray(P, np.float64)
        gold[f"cam{i}_camtoworld"] = np.asarray(pose, np.float64)
        gold[f"cam{i}_whnf"] = np.array([w, h, near, far], np.float64)
        for k in ("directions", "viewdirs", "radii", "imageplane", "origins", "cam_dirs"):
            gold[f"cam{i}_{k}"] = ref[k]
        print(f"camera {i}: {w}x{h} oracle == reference bit-for-bit on {len(ref)} keys")
    path = os.path.join(ROOT, "tests", "golden", "raygen.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
