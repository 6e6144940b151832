"""Time the generic (any-model) kernels on multicam-shaped synthetic data: linear D=3/O=4 (config 3 shape)
and pinhole D=3/O=6 (config 4 shape).  Usage: python scripts/generic_bench.py [T]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import eks_b200  # noqa: E402
from eks_b200.core import PinholeProjection  # noqa: E402
from test_oracle import fly_cams  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
import logging
logging.basicConfig(level=logging.DEBUG if os.environ.get('EKS_DEBUG') else logging.WARNING)
rng = np.random.default_rng(0)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


# config-3 shape: 4 keypoints, 2 cameras, linear 3-D latent with a full Q
K, V = 4, 2
W = np.linalg.qr(rng.standard_normal((2 * V, 3)))[0]
lat = np.cumsum(rng.normal(0, 0.3, (K, T, 3)), axis=1)
ev = rng.uniform(0.1, 0.6, (T, K, 2 * V))
ys = lat @ W.T + rng.standard_normal((K, T, 2 * V)) * np.sqrt(np.swapaxes(ev, 0, 1))
ys -= ys.mean(axis=1, keepdims=True)
Q = np.array([[1.0, 0.2, 0.05], [0.2, 0.7, 0.1], [0.05, 0.1, 0.5]])
args = (ys, np.zeros((K, 3)), np.tile(np.diag([5.0, 4.0, 3.0]), (K, 1, 1)), np.tile(np.eye(3), (K, 1, 1)),
        np.tile(W, (K, 1, 1)), np.tile(Q, (K, 1, 1)), ev)
eks_b200.run_kalman_smoother(*[a[:, :2000] if a.ndim == 3 and a.shape[1] == T else a for a in args[:-1]], ev[:2000])
(s, ms, Vs), dt = timed(lambda: eks_b200.run_kalman_smoother(*args))
print(json.dumps({'case': 'linear D3 O4', 'K': K, 'T': T, 'sec': dt, 'kf_per_s': K * T / dt, 's': list(s)}))

# config-4 shape: 6 keypoints, 3 calibrated cameras
K = 6
cams = fly_cams()
X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((K, T, 3)) * 1e-3, axis=1)
from oracle import oracle  # noqa: E402  (only to synthesise projections for this script)
ys = np.stack([oracle.project(cams, X[k]) for k in range(K)]) + rng.standard_normal((K, T, 6)) * 0.5
ev = rng.uniform(0.1, 0.5, (T, K, 6))
args = (ys, X[:, 0, :] + 0.01, np.tile(np.eye(3) * 1e-2, (K, 1, 1)), np.tile(np.eye(3), (K, 1, 1)), None,
        np.tile(np.eye(3) * 1e-6, (K, 1, 1)), ev)
h = PinholeProjection(cams)
(s, ms, Vs), dt = timed(lambda: eks_b200.run_kalman_smoother(*args, h_fn=h))
print(json.dumps({'case': 'pinhole D3 O6', 'K': K, 'T': T, 'sec': dt, 'kf_per_s': K * T / dt, 's': list(s)}))
