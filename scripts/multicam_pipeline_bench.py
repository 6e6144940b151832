"""Device-resident multi-camera pipeline at BASELINE config-2 shape (2 cameras x 4 keypoints x 10 seeds x 1e6
frames, linear PCA latent): stage times from CUDA events on the launching stream, inputs resident in HBM.
Usage: python scripts/multicam_pipeline_bench.py [T] [steps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eks_b200 import ops  # noqa: E402
from eks_b200.pipeline import multicam_smooth_sessions  # noqa: E402

T = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
M, V, K = 10, 2, 4
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
lat = torch.cumsum(torch.randn((T, K, 3), generator=g, device=dev) * 0.3, dim=0)
W = torch.randn((K, 2 * V, 3), generator=g, device=dev)
truth = torch.einsum('tkl,kol->tko', lat, W) + 50 + 250 * torch.rand((1, K, 2 * V), generator=g, device=dev)
raw = torch.empty((1, M, V, T, K, 3), device=dev)
for m in range(M):
    noisy = truth + 0.5 * torch.randn((T, K, 2 * V), generator=g, device=dev)
    raw[0, m, :, :, :, :2] = noisy.view(T, K, V, 2).permute(2, 0, 1, 3)
    raw[0, m, :, :, :, 2] = 0.8 + 0.2 * torch.rand((V, T, K), generator=g, device=dev)
del lat, truth, noisy
out = torch.empty((1, K, V, 9, T), device=dev)
for _ in range(3):
    res = multicam_smooth_sessions(raw, out=out)
torch.cuda.synchronize()
timers, n0 = {}, ops.LAUNCH_COUNT
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    res = multicam_smooth_sessions(raw, out=out, timers=timers)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
stages = {k: sum(a.elapsed_time(b) for a, b in v) / steps for k, v in timers.items()}
print(json.dumps({'workload': f'c2-shape multicam linear: {V} cameras x {K} keypoints x {M} seeds x {T} frames',
                  'ms_per_step': ms, 'kf_per_s': K * T / (ms * 1e-3), 'stages_ms': stages,
                  'iters': res.iters.cpu().tolist(), 's': res.s_finals.cpu().tolist(),
                  'gpu_launches_per_step': (ops.LAUNCH_COUNT - n0) // steps, 'dtype': 'f32'}))
