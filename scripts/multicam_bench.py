"""Device-level timing of the generic (multi-camera) models at BASELINE config 2 / config 3 sizes:
  linear : 4 sequences x T frames, D=3 latent, O=4 (2 cameras), constant-R loss        (config 2: T = 1e6)
  pinhole: 6 sequences x T frames, D=3, O=6 (3 calibrated cameras)                      (config 3: T = 5e5)
Inputs are generated on the host once and are resident in HBM; stages are timed with CUDA events.
Usage: python scripts/multicam_bench.py [linear|pinhole|both] [T] [dtype]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from eks_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'both'
T_arg = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == 'f64') else torch.float32
dev = torch.device('cuda')
rng = np.random.default_rng(0)


def ev_time(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def planes(arr):  # (K,T,O) -> [K][O][T] device planes + view
    K, T, O = arr.shape
    t = torch.as_tensor(np.ascontiguousarray(arr.transpose(0, 2, 1))).to(dev).to(dtype).contiguous()
    return ops.PlaneView(t, O * T, [o * T for o in range(O)])


def run(case, model, ys, ev, T):
    K, _, O = ys.shape
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev).to(dtype).contiguous()
    yv, vv = planes(ys), planes(np.swapaxes(ev, 0, 1))
    Rc = ops.const_R_median(vv, K, T)
    _, s_log0 = ops.initial_guess(vv, K, T)
    res = {'case': case, 'K': K, 'T': T, 'dtype': str(dtype).split('.')[-1]}
    for rep in range(2):   # second repetition is the warm one
        opt, t_opt = ev_time(lambda: ops.optimize_s(model, yv, T, Rc, s_log0))
    iters = opt['iters'].cpu().numpy()
    s = torch.exp(opt['s_log'])
    ops.nll_grad(model, yv, T, Rc, s)
    (_, t_nll) = ev_time(lambda: ops.nll_grad(model, yv, T, Rc, s), reps=5)
    ops.filter_smooth(model, yv, vv, T, s)      # warm-up: first-touch allocations of the 0.4 GB of outputs / scratch
    (_, t_sm) = ev_time(lambda: ops.filter_smooth(model, yv, vv, T, s), reps=3)
    res.update(opt_ms=t_opt, iters=iters.tolist(), ms_per_eval=t_opt / max(1, iters.max()), nll_grad_ms=t_nll,
               smooth_ms=t_sm, kf_per_s=K * T / ((t_opt + t_sm) * 1e-3), s=s.cpu().numpy().round(5).tolist())
    print(json.dumps(res), flush=True)


if which in ('linear', 'both'):
    K, V, T = 4, 2, T_arg or 1_000_000
    W = np.linalg.qr(rng.standard_normal((2 * V, 3)))[0]
    lat = np.cumsum(rng.normal(0, 0.3, (K, T, 3)), axis=1)
    ev = rng.uniform(0.1, 0.6, (T, K, 2 * V))
    ys = lat @ W.T + rng.standard_normal((K, T, 2 * V)) * np.sqrt(np.swapaxes(ev, 0, 1))
    ys -= ys.mean(axis=1, keepdims=True)
    Q = np.array([[1.0, 0.2, 0.05], [0.2, 0.7, 0.1], [0.05, 0.1, 0.5]])
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev).to(dtype).contiguous()
    model = ops.Model(f(np.zeros((K, 3))), f(np.tile(np.diag([5.0, 4.0, 3.0]), (K, 1, 1))), f(np.tile(np.eye(3), (K, 1, 1))),
                      f(np.tile(Q, (K, 1, 1))), f(np.tile(W, (K, 1, 1))))
    run('linear D3 O4', model, ys, ev, T)

if which in ('pinhole', 'both'):
    from test_oracle import fly_cams
    from oracle import oracle  # only to synthesise the projections for this script
    K, T = 6, T_arg or 500_000
    cams = fly_cams()
    X = np.array([-1.75, -0.30, 3.5]) + np.cumsum(rng.standard_normal((K, T, 3)) * 1e-3, axis=1)
    ys = np.stack([oracle.project(cams, X[k]) for k in range(K)]) + rng.standard_normal((K, T, 6)) * 0.5
    ev = rng.uniform(0.1, 0.5, (T, K, 6))
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev).to(dtype).contiguous()
    model = ops.Model(f(X[:, 0, :] + 0.01), f(np.tile(np.eye(3) * 1e-2, (K, 1, 1))), f(np.tile(np.eye(3), (K, 1, 1))),
                      f(np.tile(np.eye(3) * 1e-6, (K, 1, 1))), None, f(cams))
    run('pinhole D3 O6', model, ys, ev, T)
