#!/usr/bin/env python
"""bench.py -- throughput of the full s-optimised EKS hot path on B200 (BASELINE.json metric: keypoint-frames/s
smoothed incl. s-optimisation; % of the HBM roofline).

One "step" = one pass of the whole path (ensemble statistics -> initial guess / median R -> Adam on log s with the
reference's stop rule -> filter + RTS smoother -> output columns) over one batch of synthetic sessions per GPU.

  --workload c5 (default)  BASELINE config 5's per-GPU shard: 8 sessions x (10 seeds x 20 keypoints x 10^6 frames),
                           singlecam; `--gpus 8` = the whole 64-session job (sessions shard over the ranks with no
                           collective on the data path: `scaling: weak`).  Config 5 does not fit ONE GPU resident
                           (153.6 GB in + 46 GB out), so N = 1 runs 8 of its 64 sessions.
  --workload c2            singlecam 5 seeds x 17 keypoints x 10^5 frames
  --workload c3            multicam linear (PCA latent), 2 cameras x 4 keypoints x 10 seeds x 10^6 frames
  --workload c4            calibrated pinhole EKF (fly rig), 3 cameras x 6 keypoints x 5 seeds x 5 10^5 frames

Output: ONE JSON line (rank 0).
  value        kf/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          kf/s with HOST (pinned) inputs and outputs through the batched device pipeline, copies in the timed region
  e2e_public   kf/s through the reference-facing call (ensemble_kalman_smoother_singlecam / _multicam: MarkerArray of
               host arrays in, pandas DataFrames out), one session per call
  roofline     the dominant kernel of the step (largest device time) against the measured HBM peak; `kernels` lists
               every stage the same way; `pipeline_one_touch` the whole step against SURVEY 8(d)'s one-touch bytes
  cpu_baseline the CPU oracle (oracle/, a port of the reference path) on a bounded sample on this box's host cores,
               plus `shared_sequence`: Adam evaluation counts of the GPU and of the fp32 / fp64 oracle on the SAME data
`--impl reference` times that CPU implementation only, on the same `config` (the reference's own JAX path cannot be
installed here: jax / dynamax / optax are absent and there is no network).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, seeds M, cameras V, keypoints K, frames T, default sessions per GPU
    'c5': dict(kind='singlecam', M=10, V=1, K=20, T=1_000_000, S=8),
    'c2': dict(kind='singlecam', M=5, V=1, K=17, T=100_000, S=1),
    'c3': dict(kind='multicam', M=10, V=2, K=4, T=1_000_000, S=1),
    'c4': dict(kind='pinhole', M=5, V=3, K=6, T=500_000, S=1),
}
METRIC = 'keypoint-frames/sec smoothed incl. s-optimisation'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c5', choices=list(WORKLOADS))
    ap.add_argument('--sessions', type=int, default=None, help='sessions per GPU per step')
    ap.add_argument('--frames', type=int, default=None)
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--opt-mode', default='lag', choices=['lag', 'stream'], help='singlecam optimiser (pipeline.py)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-public', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-frames', type=int, default=None, help='frames per sequence of the CPU sample')
    ap.add_argument('--cpu-budget', type=float, default=150.0, help='seconds the reference arm may spend in total')
    return ap.parse_args()


def workload_of(args):
    w = dict(WORKLOADS[args.workload])
    if args.sessions:
        w['S'] = args.sessions
    if args.frames:
        w['T'] = args.frames
    return w


def config_of(args, w):
    """The SAME dict for both arms (--impl b200 / reference): names the workload, nothing measured."""
    kind = {'singlecam': 'singlecam, per-keypoint Adam s-optimisation (reference stop rule)',
            'multicam': 'multicam linear PCA latent (D=3), per-keypoint Adam s-optimisation',
            'pinhole': 'multicam calibrated pinhole EKF (fly rig, D=3), per-keypoint Adam s-optimisation'}[w['kind']]
    return {'workload': f"{args.workload}: {w['S']} sessions/GPU x {w['M']} seeds x {w['V']} cameras x {w['K']} keypoints x "
                        f"{w['T']} frames, {kind}",
            'sessions_per_gpu': w['S'], 'seeds': w['M'], 'cameras': w['V'], 'keypoints': w['K'], 'frames': w['T'],
            'l2': ('inputs larger than L2 (resident raw tensor %.2f GB per GPU)' if
                   w['S'] * w['M'] * w['V'] * w['T'] * w['K'] * 3 * (4 if args.dtype == 'f32' else 8) >= (1 << 30) else
                   'L2 flushed between timed steps by an untimed 512 MB write (resident raw tensor %.2f GB per GPU)') % (
                w['S'] * w['M'] * w['V'] * w['T'] * w['K'] * 3 * (4 if args.dtype == 'f32' else 8) / 1e9)}


# ----------------------------------------------------------------------------- synthetic data
def synth_session_device(torch, M, K, T, seed, device, dtype):
    """(M,1,T,K,3) on device: random-walk latent (sigma .3 px/frame), seeds = truth + N(0, .5^2), x10 noise
    and low likelihood on a 2% Bernoulli occlusion mask shared by the seeds."""
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    truth = torch.cumsum(torch.randn((T, K, 2), generator=g, device=device, dtype=torch.float32) * 0.3, dim=0)
    truth += torch.rand((1, K, 2), generator=g, device=device) * 250.0 + 50.0
    occ = torch.rand((T, K), generator=g, device=device) < 0.02
    sigma = torch.where(occ, 5.0, 0.5)[None, :, :, None]
    raw = torch.empty((M, 1, T, K, 3), device=device, dtype=dtype)
    for m in range(M):
        raw[m, 0, :, :, 0:2] = (truth + torch.randn((T, K, 2), generator=g, device=device) * sigma[0]).to(dtype)
        u = torch.rand((T, K), generator=g, device=device)
        raw[m, 0, :, :, 2] = torch.where(occ, 0.05 + 0.45 * u, 0.9 + 0.1 * u).to(dtype)
    return raw


def synth_session_host(M, K, T, seed):
    rng = np.random.default_rng(1234 + seed)
    truth = np.cumsum(rng.normal(0, 0.3, size=(T, K, 2)), axis=0) + rng.uniform(50, 300, size=(1, K, 2))
    occ = rng.random((T, K)) < 0.02
    sigma = np.where(occ, 5.0, 0.5)[None, :, :, None]
    pred = truth[None] + rng.normal(size=(M, T, K, 2)) * sigma
    lik = np.where(occ[None], rng.uniform(0.05, 0.5, size=(M, T, K)), rng.uniform(0.9, 1.0, size=(M, T, K)))
    return np.concatenate([pred, lik[..., None]], axis=-1)[:, None].astype(np.float32)


def synth_multicam_host(M, V, K, T, seed):
    """Config-3 family: 3-D random-walk latent -> 2V pixel coordinates through a random orthonormal
    W (2V x 3) + per-camera offsets of 100-300 px; seeds = truth + N(0, .5^2), x8 noise and low likelihood on a 2 %
    occlusion mask shared by seeds and cameras.  Returns (M,V,T,K,3) float32."""
    rng = np.random.default_rng(4321 + seed)
    raw = np.empty((M, V, T, K, 3), dtype=np.float32)
    for k in range(K):
        lat = np.cumsum(rng.normal(0, 0.3, size=(T, 3)), axis=0)
        W = np.linalg.qr(rng.standard_normal((2 * V, 3)))[0]                       # orthonormal columns
        truth = lat @ W.T * 1.5 + rng.uniform(100, 300, size=(1, 2 * V))              # (T, 2V)
        occ = rng.random(T) < 0.02
        sigma = np.where(occ, 4.0, 0.5)[:, None]
        for m in range(M):
            noisy = truth + rng.standard_normal((T, 2 * V)) * sigma
            raw[m, :, :, k, 0:2] = noisy.reshape(T, V, 2).transpose(1, 0, 2)
            u = rng.random((V, T))
            raw[m, :, :, k, 2] = np.where(occ[None], 0.05 + 0.45 * u, 0.9 + 0.1 * u)
    return raw


def fly_cameras():
    """(V, 29) packed cameras of the bundled fly rig (tests/golden/fly_calibration.toml = the reference's
    data/fly/calibration.toml) + the product's CameraGroup."""
    from eks_b200.multicam_smoother import CameraGroup, make_projection_from_camgroup
    cg = CameraGroup.load(os.path.join(ROOT, 'tests', 'golden', 'fly_calibration.toml'))
    return np.asarray(make_projection_from_camgroup(cg)[0].cams, dtype=np.float64), cg


def project_pinhole_np(cams, X):
    """The reference's pinhole model (eks/multicam_smoother.py:806-859) in NumPy, for data synthesis only:
    X (N,3) world -> (N, 2V) pixels.  cams: (V,29) = R(9) t(3) fx fy cx cy skew k1 k2 p1 p2 k3 k4 k5 k6 s1..s4."""
    out = np.empty((X.shape[0], 2 * cams.shape[0]))
    for v, c in enumerate(cams):
        Rm, t = c[0:9].reshape(3, 3), c[9:12]
        fx, fy, cx, cy, skew = c[12:17]
        k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = c[17:29]
        Xc = X @ Rm.T + t
        x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
        r2 = x * x + y * y
        radial = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3 + k4 * r2 ** 4 + k5 * r2 ** 5 + k6 * r2 ** 6
        xd = x * radial + 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + s1 * r2 + s2 * r2 * r2
        yd = y * radial + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + s3 * r2 + s4 * r2 * r2
        out[:, 2 * v] = fx * xd + skew * yd + cx
        out[:, 2 * v + 1] = fy * yd + cy
    return out


def synth_fly_host(M, K, T, seed, cams):
    """Config-4 family: 3-D random walk (step sigma 1e-3 world units, reflected into a +-0.1 box around the
    rig centre (-1.75, -0.30, 3.50)) projected through the calibrated cameras, + N(0, .5^2) px noise per seed, x8 on a
    2 % occlusion mask.  Returns (M,V,T,K,3) float32."""
    rng = np.random.default_rng(8765 + seed)
    V = cams.shape[0]
    raw = np.empty((M, V, T, K, 3), dtype=np.float32)
    centre = np.array([-1.75, -0.30, 3.50])
    for k in range(K):
        w = np.cumsum(rng.normal(0, 1e-3, size=(T, 3)), axis=0) + rng.uniform(-0.05, 0.05, size=(1, 3))
        w = np.abs((w + 0.1) % 0.4 - 0.2) - 0.1                                       # reflect into [-0.1, 0.1]
        uv = project_pinhole_np(cams, centre + w)                                     # (T, 2V)
        occ = rng.random(T) < 0.02
        sigma = np.where(occ, 4.0, 0.5)[:, None]
        for m in range(M):
            noisy = uv + rng.standard_normal((T, 2 * V)) * sigma
            raw[m, :, :, k, 0:2] = noisy.reshape(T, V, 2).transpose(1, 0, 2)
            u = rng.random((V, T))
            raw[m, :, :, k, 2] = np.where(occ[None], 0.05 + 0.45 * u, 0.9 + 0.1 * u)
    return raw


# ----------------------------------------------------------------------------- CPU oracle leg
def synth_host(w, T, seed):
    """one session of workload w with T frames: (M,V,T,K,3) float32"""
    if w['kind'] == 'singlecam':
        return synth_session_host(w['M'], w['K'], T, seed)
    if w['kind'] == 'multicam':
        return synth_multicam_host(w['M'], w['V'], w['K'], T, seed)
    return synth_fly_host(w['M'], w['K'], T, seed, fly_cameras()[0])


def oracle_run(w, raw, dtype):
    from oracle import oracle
    if w['kind'] == 'singlecam':
        return oracle.singlecam(raw, dtype=dtype)
    if w['kind'] == 'multicam':
        return oracle.multicam(raw.astype(np.float64), dtype=dtype, quantile_keep_pca=50.0)
    return oracle.multicam(raw.astype(np.float64), dtype=dtype,
                           camgroup=os.path.join(ROOT, 'tests', 'golden', 'fly_calibration.toml'))


def cpu_leg(w, T_sample, steps, warmup, raw=None):
    """Time the CPU oracle (oracle/liboracle.so, OpenMP over sequences; fp32 like the reference's production path) on
    ONE session of T_sample frames."""
    if raw is None:
        raw = synth_host(w, T_sample, seed=0)
    cores = os.cpu_count() or 1
    os.environ['OMP_NUM_THREADS'] = str(cores)   # torchrun exports OMP_NUM_THREADS=1: use every host core
    times, iters = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = oracle_run(w, raw, np.float32)
        dt = time.perf_counter() - t0
        iters = r['info']['iters']
        if i >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return dict(value=w['K'] * T_sample / sec, unit='keypoint-frames/s', cores=cores, kind='port',
                sample=f"1 session x {w['M']} seeds x {w['V']} cameras x {w['K']} keypoints x {T_sample} frames, fp32, "
                       f"oracle/liboracle.so (C++/OpenMP restatement of the reference path; mean Adam iterations "
                       f"{float(np.mean(iters)):.1f}); the path is O(frames), kf/s does not depend on the frame count",
                sec_per_step=sec, iters=[int(x) for x in iters])


def cpu_rate_probe(w):
    """kf/s of the oracle on a small probe (sizes the bounded sample)."""
    T = min(w['T'], 20_000 if w['kind'] == 'singlecam' else 5_000)
    return cpu_leg(w, T, 1, 0)['value']


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = workload_of(args)
    if args.cpu_frames:
        Tc = min(args.cpu_frames, w['T'])
    else:   # the whole --steps/--warmup run must end within a few minutes: bound the frames per step
        rate = cpu_rate_probe(w)
        n_runs = max(1, args.steps) + min(args.warmup, 1)
        Tc = int(min(w['T'], max(10_000, rate * args.cpu_budget / (n_runs * w['K']))))
        Tc = (Tc // 1000) * 1000
    res = cpu_leg(w, Tc, max(1, args.steps), min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': 'keypoint-frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['sec_per_step'] * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_of(args, w),
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': 'keypoint-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'the reference JAX/dynamax path is not installable here; this arm times oracle/, the CPU restatement of '
                'the same algorithm, on all host cores; each step = one session of the workload truncated to '
                f'{Tc} frames (of {w["T"]}) so that the run ends within minutes',
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                   'samples': len(sm)}
        return out


# ----------------------------------------------------------------------------- B200 arm
def stage_bytes_per_kf(w, wbytes, stage, n_eval_mean, opt_mode):
    """ALGORITHMIC bytes per keypoint-frame of one stage (SURVEY 8d "per-kernel" figures)."""
    M, V = w['M'], w['V']
    obs = 2 * V
    D = 2 if w['kind'] == 'singlecam' else 3
    if stage == 'ensemble':
        return wbytes * (3 * M * V + 5 * V)                 # raw x, y, likelihood in; 5 planes per camera out
    if stage == 'const_R_median':
        return wbytes * obs                                 # one read of the variance planes
    if stage == 'optimize_s':
        if (w['kind'] == 'singlecam' and opt_mode == 'lag') or w['kind'] == 'multicam':
            return wbytes * obs                             # the observations are read ONCE (lag statistics)
        return wbytes * obs * n_eval_mean                   # one read of the observations per evaluation
    if stage == 'filter_smooth':
        if w['kind'] == 'singlecam':
            return wbytes * 4 * obs                         # y, var in; x, posterior var out (fused epilogue)
        return wbytes * (2 * obs + D + D * D)               # y, var in; smoothed mean + covariance out
    if stage == 'reproject':
        return wbytes * (D + D * D + obs + 2 * obs)         # latent moments + var in; x, y, 2 posterior vars per camera
    if stage == 'triangulate':
        return wbytes * 3 * M * V + 24
    if stage in ('center', 'pca', 'latent_init'):
        return wbytes * (obs + 1)
    return 0.0


def run_b200(args):
    import torch
    import torch.distributed as dist
    from eks_b200 import ops
    from eks_b200.pipeline import multicam_smooth_sessions, singlecam_smooth_sessions

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    # one rank per GPU: run on the CPUs of the GPU's NUMA node so that the pinned host buffers of the e2e leg are local
    from eks_b200.parallel import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if world > 1 else {'node': None, 'why': 'single rank: not bound'}
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        gathered = [None] * world
        dist.all_gather_object(gathered, numa)
        numa = gathered                      # every rank's binding, in rank order
    w = workload_of(args)
    kind, M, V, K, T, S = w['kind'], w['M'], w['V'], w['K'], w['T'], w['S']
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    wb = 4 if args.dtype == 'f32' else 8
    cams = fly_cameras()[0] if kind == 'pinhole' else None

    # resident inputs: S sessions on this GPU (distinct seeds per rank and session)
    raw = torch.empty((S, M, V, T, K, 3), device=dev, dtype=dtype)
    for s_ in range(S):
        if kind == 'singlecam':
            raw[s_] = synth_session_device(torch, M, K, T, seed=rank * 1000 + s_, device=dev, dtype=dtype)
        else:
            raw[s_] = torch.as_tensor(synth_host(w, T, seed=rank * 1000 + s_)).to(dev, dtype)
    out_shape = (S, K, 9, T) if kind == 'singlecam' else (S, K, V, 9, T)
    out = torch.empty(out_shape, device=dev, dtype=dtype)
    torch.cuda.synchronize()

    def step(raw_, out_, timers=None):
        if kind == 'singlecam':
            return singlecam_smooth_sessions(raw_, dtype=dtype, out=out_, timers=timers, opt_mode=args.opt_mode)
        return multicam_smooth_sessions(raw_, dtype=dtype, out=out_, timers=timers, quantile_keep_pca=50.0, cams=cams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timers = {}
    res = None
    for _ in range(args.warmup):
        res = step(raw, out)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ops.LAUNCH_COUNT = 0
    # inputs larger than L2 need no flush; small workloads (c2) get an untimed 512 MB write between the timed steps
    resident = raw.numel() * wb
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if resident < (1 << 30) else None
    evs = []
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record()
        res = step(raw, out)
        e1_.record()
        evs.append((e0_, e1_))
    barrier()
    launches = ops.LAUNCH_COUNT
    # per-stage device times: a SEPARATE serial pass after the timed region (per-stage CUDA events need the stages of
    # all sessions in ONE stream; the timed steps above run session groups on several streams so that stages overlap,
    # which is why the stage times sum to more than ms_per_step)
    step(raw, out, {})      # untimed: the serial pass has its own workspace sizes (first use = cudaMalloc inside a stage)
    for _ in range(2):
        if flush is not None:
            flush.fill_(1)
        step(raw, out, timers)
    barrier()
    ms_total = float(sum(a_.elapsed_time(b_) for a_, b_ in evs))
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    kf_step = S * K * T                      # per GPU
    value = world * kf_step / (ms_step * 1e-3)
    iters = res.iters.double()
    n_eval_mean = float(iters.mean().item())
    n_eval_max = int(iters.max().item())
    # per-stage device times (CUDA events recorded on the launching stream inside the timed region)
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}

    # ---- e2e, batched device pipeline: host (pinned) inputs and outputs, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        n_pin = min(S, 2)
        sess_in, sess_out = (1, M, V, T, K, 3), (1,) + out_shape[1:]
        h_in = [torch.empty(sess_in, dtype=dtype).pin_memory() for _ in range(n_pin)]
        for i in range(n_pin):
            h_in[i].copy_(raw[i:i + 1])
        h_out = [torch.empty(sess_out, dtype=dtype).pin_memory() for _ in range(n_pin)]
        torch.cuda.synchronize()
        s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        d_in = [torch.empty(sess_in, device=dev, dtype=dtype) for _ in range(2)]
        d_out = [torch.empty(sess_out, device=dev, dtype=dtype) for _ in range(2)]

        def e2e_step():
            ev_in, ev_cmp, ev_out = [None, None], [None, None], [None, None]
            for i in range(S):
                b = i & 1
                with torch.cuda.stream(s_in):
                    if ev_cmp[b] is not None:
                        s_in.wait_event(ev_cmp[b])          # buffer free once its compute finished
                    d_in[b].copy_(h_in[i % n_pin], non_blocking=True)
                    ev_in[b] = torch.cuda.Event()
                    ev_in[b].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[b])
                    if ev_out[b] is not None:
                        s_cmp.wait_event(ev_out[b])
                    step(d_in[b], d_out[b])
                    ev_cmp[b] = torch.cuda.Event()
                    ev_cmp[b].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b])
                    h_out[i % n_pin].copy_(d_out[b], non_blocking=True)
                    ev_out[b] = torch.cuda.Event()
                    ev_out[b].record(s_out)
            torch.cuda.current_stream().wait_stream(s_out)
            torch.cuda.current_stream().wait_stream(s_cmp)

        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(1, min(args.steps, 2))
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = float(te.item()) / n_e2e
        h2d = int(S * M * V * T * K * 3 * wb)
        d2h = int(np.prod(out_shape) * wb)
        e2e = {'value': world * kf_step / (ms_e2e * 1e-3), 'unit': 'keypoint-frames/s',
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
               'h2d_gbs_per_rank': (h2d + d2h) / (ms_e2e * 1e-3) / 1e9,
               'how': 'per session: pinned H2D -> eks_b200.pipeline.*_smooth_sessions -> pinned D2H, double-buffered on '
                      'three streams; h2d_gbs_per_rank = (H2D + D2H bytes) / time on this rank'}
        del h_in, h_out, d_in, d_out

    # ---- e2e through the reference-facing call: MarkerArray of host arrays in, pandas DataFrames out
    e2e_public = None
    if not args.no_public:
        from eks_b200.marker_array import MarkerArray
        host = raw[0].float().cpu().numpy() if dtype == torch.float32 else raw[0].cpu().numpy()
        kps = [f'kp{k}' for k in range(K)]
        if kind == 'singlecam':
            from eks_b200.singlecam_smoother import ensemble_kalman_smoother_singlecam

            def public():
                return ensemble_kalman_smoother_singlecam(MarkerArray(host, data_fields=['x', 'y', 'likelihood']), kps)
        else:
            from eks_b200.multicam_smoother import ensemble_kalman_smoother_multicam
            cg = fly_cameras()[1] if kind == 'pinhole' else None
            names = [c.name for c in cg.cameras] if cg else [f'cam{v}' for v in range(V)]

            def public():
                return ensemble_kalman_smoother_multicam(MarkerArray(host, data_fields=['x', 'y', 'likelihood']), kps, names,
                                                         quantile_keep_pca=50.0, camgroup=cg)
        public()                                        # warm-up (allocator, pinned staging)
        barrier()
        n_pub = 2
        t0 = time.perf_counter()
        for _ in range(n_pub):
            r_pub = public()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / n_pub
        tp = torch.tensor([sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        sec = float(tp.item())
        out_bytes = int(sum(df.to_numpy().nbytes for df in (r_pub[0] if isinstance(r_pub[0], list) else [r_pub[0]])))
        e2e_public = {'value': world * K * T / sec, 'unit': 'keypoint-frames/s', 'sec_per_call': sec,
                      'h2d_bytes_per_call': int(host.nbytes), 'd2h_bytes_per_call': out_bytes,
                      'how': 'wall clock of one call of the reference-facing entry point on host arrays (MarkerArray in, '
                             'DataFrames out), one session per call, every rank concurrently (max over ranks)'}
        del host
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    else:
        peak, peak_src = 6650.0, 'fallback 6.65 TB/s (of fallback)'
    # ---- every stage against the HBM roofline: algorithmic bytes of the stage / its CUDA-event time
    kernel_names = {'ensemble': 'ensemble_staged_kernel',
                    'const_R_median': 'med_sample_kernel + med_count_kernel + med_final_kernel',
                    'optimize_s': ('lag_stats_kernel + diag_lag_opt_kernel' if kind == 'singlecam' and args.opt_mode == 'lag'
                                   else 'diag_nll_kernel' if kind == 'singlecam' else
                                   'ml_signal_kernel + mlag_stats_kernel + lin_lag_opt_kernel' if kind == 'multicam' else
                                   'gen_nll_runs_kernel + gen_runs_reduce_kernel'),
                    'filter_smooth': 'diag_smooth_fused_kernel' if kind == 'singlecam' else
                                     'gen_filter_runs_kernel + gen_rts_runs_kernel',
                    'reproject': 'reproject_kernel', 'triangulate': 'triangulate_mean_kernel'}
    tj = {}
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):   # dram bytes per algorithmic byte from the committed ncu --set full captures
        tj = json.load(open(tpath))
    kernels = {}
    for st, ms in stage_ms.items():
        bpk = stage_bytes_per_kf(w, wb, st, n_eval_mean, args.opt_mode)
        if bpk <= 0 or ms <= 0:
            continue
        alg = bpk * kf_step
        ach = alg / (ms * 1e-3) / 1e9
        tje = tj.get(kernel_names.get(st, st), {})
        ratio = (tje['dram_bytes'] / tje['algorithmic_bytes']) if tje.get('algorithmic_bytes') else None
        kernels[st] = {'kernel': kernel_names.get(st, st), 'ms': ms, 'algorithmic_bytes': alg, 'achieved': ach,
                       'frac': ach / peak, 'share_of_step': ms / sum(stage_ms.values()),
                       'traffic': (alg * ratio) if ratio else None}
    dom = max(kernels, key=lambda k: kernels[k]['ms']) if kernels else None
    roofline = None
    if dom:
        kd = kernels[dom]
        roofline = {'bound': 'hbm', 'kernel': kd['kernel'], 'stage': dom, 'achieved': kd['achieved'], 'peak': peak,
                    'unit': 'GB/s', 'frac': kd['frac'], 'traffic': kd['traffic'], 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': kd['algorithmic_bytes'], 'launch_ms': kd['ms'],
                    'share_of_step': kd['share_of_step'],
                    'how': 'dominant stage of the step = largest CUDA-event time on the launching stream, measured in a '
                           'serial pass (one stream, per-stage events) right after the timed region; achieved = '
                           'algorithmic bytes of that stage / that time; share_of_step = stage time / sum of stage times'}
    D3 = 0 if kind == 'singlecam' else 3
    b_alg = wb * (3 * M * V + 9 * V + 2 * D3)
    pipeline_frac = (kf_step * b_alg / (ms_step * 1e-3) / 1e9) / peak

    cpu = shared = None
    if not args.no_cpu and world == 1:   # the CPU baseline is a rank-0, N = 1 leg (torchrun pins OMP to one thread)
        Tc = min(args.cpu_frames or (100_000 if kind == 'singlecam' else 20_000), T)
        sample = synth_host(w, Tc, seed=0)
        c = cpu_leg(w, Tc, 1, 0, raw=sample)
        cpu = {k: c[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        # Adam evaluation counts on the SAME sequences: GPU (this arm's precision), fp32 oracle, fp64 oracle
        g = step(torch.as_tensor(sample).to(dev, dtype)[None], None)
        torch.cuda.synchronize()
        kk = min(K, 4)
        r64 = oracle_run(w, sample[:, :, :, :kk], np.float64)
        shared = {'frames': Tc, 'keypoints_compared': kk,
                  'n_eval_gpu': [int(x) for x in g.iters[0].cpu().numpy()[:kk]],
                  'n_eval_oracle_f64': [int(x) for x in r64['info']['iters']],
                  'n_eval_oracle_f32': c['iters'][:kk],
                  's_gpu': [float(x) for x in g.s_finals[0].cpu().numpy()[:kk]],
                  's_oracle_f64': [float(x) for x in r64['s_finals']]}

    cfg = config_of(args, w)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'keypoint-frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': cfg,
        'e2e': e2e, 'e2e_public': e2e_public, 'gpu_launches': launches, 'clocks': clocks, 'numa': numa,
        'roofline': roofline,
        'kernels': kernels,
        'pipeline_one_touch': {'bytes_per_kf': b_alg, 'frac_of_hbm_peak': pipeline_frac,
                               'achieved_gbs': kf_step * b_alg / (ms_step * 1e-3) / 1e9},
        'n_eval': {'mean': n_eval_mean, 'max': n_eval_max, 'optimiser': args.opt_mode if kind == 'singlecam' else 'runs'},
        'stage_ms': stage_ms, 'stage_ms_serial_sum': float(sum(stage_ms.values())),
        'stage_ms_note': 'stages timed in a serial pass after the timed region; the timed steps overlap the stages of '
                         'independent session groups on separate streams (pipeline._auto_groups)',
        'cpu_baseline': cpu, 'shared_sequence': shared,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
