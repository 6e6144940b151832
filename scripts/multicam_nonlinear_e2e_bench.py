"""Wall-clock of the public calibrated multi-camera entry point at BASELINE config-3 shape (3 calibrated cameras x 6
keypoints x 10 seeds x T frames): host arrays in, DataFrames out.  Usage: python scripts/multicam_nonlinear_e2e_bench.py [T]"""
import json
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from eks_b200.marker_array import MarkerArray  # noqa: E402
from eks_b200.multicam_smoother import CameraGroup, ensemble_kalman_smoother_multicam, make_projection_from_camgroup  # noqa: E402
from oracle import oracle  # noqa: E402  (only to synthesise the projections)

T = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000
M, K = 10, 6
cg = CameraGroup.load(os.path.join(ROOT, 'tests', 'golden', 'fly_calibration.toml'))
cams = np.asarray(make_projection_from_camgroup(cg)[0].cams, dtype=np.float64)
V = cams.shape[0]
rng = np.random.default_rng(0)
X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((K, T, 3)) * 1e-3, axis=1)
uv = np.stack([oracle.project(cams, X[k]) for k in range(K)])
raw = np.empty((M, V, T, K, 3), dtype=np.float32)
for m in range(M):
    noisy = uv + rng.standard_normal(uv.shape) * 0.5
    raw[m, :, :, :, :2] = noisy.reshape(K, T, V, 2).transpose(2, 1, 0, 3)
    raw[m, :, :, :, 2] = 0.9
kps, names = [f'kp{k}' for k in range(K)], [c.name for c in cg.cameras]
ensemble_kalman_smoother_multicam(MarkerArray(raw[:, :, :2000].copy(), data_fields=['x', 'y', 'likelihood']), kps, names,
                                  camgroup=cg)      # warm-up
logging.basicConfig(level=logging.WARNING)
torch.cuda.synchronize()
t0 = time.perf_counter()
dfs, s, df3 = ensemble_kalman_smoother_multicam(MarkerArray(raw, data_fields=['x', 'y', 'likelihood']), kps, names, camgroup=cg)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({'shape': f'{M} seeds x {V} cameras x {K} keypoints x {T} frames', 'sec': dt, 'kf_per_s': K * T / dt,
                  's': [float(x) for x in s]}))
