"""Time eks_triangulate_mean at BASELINE config-3 shape (3 calibrated cameras x 6 keypoints x 10 seeds x 5e5 frames) and
the host mirror (cv2) on a 1/50 sample of the same data.  Usage: python scripts/triangulate_bench.py [T]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from eks_b200 import ops  # noqa: E402
from eks_b200.marker_array import MarkerArray  # noqa: E402
from eks_b200.multicam_smoother import CameraGroup, make_projection_from_camgroup, triangulate_3d_models  # noqa: E402
from oracle import oracle  # noqa: E402  (only to synthesise the projections)

T = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000
M, K = 10, 6
cg = CameraGroup.load(os.path.join(ROOT, 'tests', 'golden', 'fly_calibration.toml'))
cams = np.asarray(make_projection_from_camgroup(cg)[0].cams, dtype=np.float64)
V = cams.shape[0]
rng = np.random.default_rng(0)
X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((K, T, 3)) * 1e-3, axis=1)
uv = np.stack([oracle.project(cams, X[k]) for k in range(K)])                      # (K,T,2V)
raw = np.empty((M, V, T, K, 3), dtype=np.float32)
for m in range(M):
    noisy = uv + rng.standard_normal(uv.shape) * 0.5
    raw[m, :, :, :, :2] = noisy.reshape(K, T, V, 2).transpose(2, 1, 0, 3)
    raw[m, :, :, :, 2] = 0.9
d_raw = torch.as_tensor(raw).cuda()
d_cams = torch.as_tensor(cams)
for _ in range(2):
    out = ops.triangulate_mean(d_raw, d_cams)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = ops.triangulate_mean(d_raw, d_cams)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
Ts = T // 50
t0 = time.perf_counter()
ref = triangulate_3d_models(MarkerArray(raw[:, :, :Ts].astype(np.float64), data_fields=['x', 'y', 'likelihood'],
                                        dtype=np.float64), cg).mean(axis=0)
host_s = time.perf_counter() - t0
err = np.abs(out[:, :Ts].cpu().numpy() - ref).max()
print(json.dumps({'shape': f'{M} seeds x {V} cameras x {K} keypoints x {T} frames', 'device_ms': ms,
                  'points_per_s': M * K * T / (ms * 1e-3), 'host_mirror_s_for_1_50': host_s,
                  'host_points_per_s': M * K * Ts / host_s, 'max_abs_diff_vs_host': float(err)}))
