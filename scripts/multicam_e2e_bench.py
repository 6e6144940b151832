"""Wall-clock profile of the public multi-camera entry point at BASELINE config-2 shape
(2 cameras x 4 keypoints x 10 seeds x T frames, linear PCA latent).  Usage: python scripts/multicam_e2e_bench.py [T]"""
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eks_b200  # noqa: E402
from eks_b200.marker_array import MarkerArray  # noqa: E402
from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam  # noqa: E402

T = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
M, V, K = 10, 2, 4
rng = np.random.default_rng(0)
lat = np.cumsum(rng.normal(0, 0.3, (T, K, 3)), axis=0)
W = rng.standard_normal((K, 2 * V, 3))
truth = np.einsum('tkl,kol->tko', lat, W) + rng.uniform(50, 300, (1, K, 2 * V))     # (T,K,2V)
raw = np.empty((M, V, T, K, 3), dtype=np.float32)
for m in range(M):
    noise = rng.standard_normal((T, K, 2 * V)).astype(np.float32) * 0.5
    raw[m, :, :, :, :2] = (truth + noise).reshape(T, K, V, 2).transpose(2, 0, 1, 3)
    raw[m, :, :, :, 2] = rng.uniform(0.8, 1.0, (V, T, K))
ma = MarkerArray(raw, data_fields=['x', 'y', 'likelihood'])
logging.basicConfig(level=logging.WARNING)
kps, cams = [f'kp{k}' for k in range(K)], [f'cam{v}' for v in range(V)]
ensemble_kalman_smoother_multicam(MarkerArray(raw[:, :, :3000].copy(), data_fields=['x', 'y', 'likelihood']), kps, cams,
                                  quantile_keep_pca=50.0)     # warm-up (library load, allocator)
logging.getLogger('eks_b200').setLevel(logging.DEBUG)
logging.getLogger().setLevel(logging.DEBUG)
torch.cuda.synchronize()
t0 = time.perf_counter()
dfs, s, df3d = ensemble_kalman_smoother_multicam(ma, kps, cams, quantile_keep_pca=50.0)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print({'T': T, 'sec': dt, 'kf_per_s': K * T / dt, 's': [float(x) for x in s]})
