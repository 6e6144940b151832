#!/usr/bin/env python
"""bench.py -- throughput of the full s-optimised single-camera EKS hot path on B200.

One "step" = one pass of the whole path (ensemble statistics -> s-optimisation -> filter + RTS smoother
-> nine output columns) over one batch of `--sessions` synthetic sessions per GPU.  The default workload
is BASELINE.json config 5's per-GPU shard: sessions of 10 seeds x 20 keypoints x 1M frames; with
`--gpus 8 --sessions 8` the job is exactly config 5 (64 sessions sharded over 8 B200, no collective on
the data path: `scaling: weak`).  `--workload c2` runs config 2 (5 seeds x 17 keypoints x 100k frames).

Output: ONE JSON line (rank 0).  `value` = keypoint-frames/s with inputs resident in HBM; `e2e` = the
same metric through the public API with HOST (pinned) inputs and outputs, copies inside the timed
region; `roofline` = dominant kernel vs measured HBM peak; `cpu_baseline` = the CPU oracle (a port of
the reference path, oracle/) timed on a bounded sample on this box's host cores.
`--impl reference` times that CPU implementation only (the reference's own JAX path cannot be installed
in this environment: jax/dynamax/optax are absent and there is no network).
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
    # name: (seeds M, keypoints K, frames T, default sessions per GPU)
    'c5': (10, 20, 1_000_000, 8),
    'c2': (5, 17, 100_000, 1),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c5', choices=list(WORKLOADS))
    ap.add_argument('--sessions', type=int, default=None, help='sessions per GPU per step')
    ap.add_argument('--frames', type=int, default=None)
    ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-frames', type=int, default=20_000, help='frames per sequence of the CPU sample')
    return ap.parse_args()


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
def cpu_leg(M, K, T_sample, steps, warmup):
    """Time the CPU oracle (oracle/liboracle.so, OpenMP over sequences) on one session of T_sample frames."""
    from oracle import oracle
    raw = synth_session_host(M, K, T_sample, seed=0)
    cores = os.cpu_count() or 1
    os.environ['OMP_NUM_THREADS'] = str(cores)   # torchrun exports OMP_NUM_THREADS=1: use every host core
    times, iters = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = oracle.singlecam(raw, dtype=np.float32)
        dt = time.perf_counter() - t0
        iters = r['info']['iters']
        if i >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return dict(value=K * T_sample / sec, unit='keypoint-frames/s', cores=cores, kind='port',
                sample=f'1 session x {M} seeds x {K} keypoints x {T_sample} frames, fp32, oracle/liboracle.so '
                       f'(C++/OpenMP restatement of the reference path; mean Adam iterations '
                       f'{float(np.mean(iters)):.1f})',
                sec_per_step=sec)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    M, K, T, S = WORKLOADS[args.workload]
    Tc = min(args.cpu_frames, T)
    res = cpu_leg(M, K, Tc, max(1, args.steps), min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': 'keypoint-frames/sec smoothed incl. s-optimisation', 'value': res['value'],
        'unit': 'keypoint-frames/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': res['sec_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: sessions of {M} seeds x {K} keypoints x {T} frames, '
                               f'singlecam, per-keypoint s-optimisation (CPU sample: {Tc} frames/sequence, '
                               'O(T) path => kf/s is frame-count independent)'},
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': 'keypoint-frames/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'the reference JAX/dynamax path is not installable here; this arm times oracle/, the CPU '
                'restatement of the same algorithm, on all host cores',
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
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    from eks_b200 import ops
    from eks_b200.pipeline import singlecam_smooth_sessions

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    M, K, T, S = WORKLOADS[args.workload]
    if args.sessions:
        S = args.sessions
    if args.frames:
        T = args.frames
    dtype = torch.float32 if args.dtype == 'f32' else torch.float64
    w = 4 if args.dtype == 'f32' else 8

    # resident inputs: S sessions on this GPU (distinct seeds per rank and session)
    raw = torch.empty((S, M, 1, T, K, 3), device=dev, dtype=dtype)
    for s_ in range(S):
        raw[s_] = synth_session_device(torch, M, K, T, seed=rank * 1000 + s_, device=dev, dtype=dtype)
    out = torch.empty((S, K, 9, T), device=dev, dtype=dtype)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timers = {}
    res = None
    for _ in range(args.warmup):
        res = singlecam_smooth_sessions(raw, dtype=dtype, out=out)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ops.LAUNCH_COUNT = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = singlecam_smooth_sessions(raw, dtype=dtype, out=out, timers=timers)
    ev1.record()
    barrier()
    launches = ops.LAUNCH_COUNT
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    clocks = sampler.stop() if sampler else None
    kf_step = S * K * T                      # per GPU
    value = world * kf_step / (ms_step * 1e-3)
    iters = res.iters.double()
    n_eval_mean = float(iters.mean().item())
    n_eval_max = int(iters.max().item())

    # per-stage device times (CUDA events recorded on the launching stream inside the timed region)
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}

    # ---- e2e through the public API: host (pinned) inputs and outputs, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        n_pin = min(S, 2)
        h_in = [torch.empty((1, M, 1, T, K, 3), dtype=dtype).pin_memory() for _ in range(n_pin)]
        for i in range(n_pin):
            h_in[i].copy_(raw[i:i + 1])
        h_out = [torch.empty((1, K, 9, T), dtype=dtype).pin_memory() for _ in range(n_pin)]
        torch.cuda.synchronize()
        s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        d_in = [torch.empty((1, M, 1, T, K, 3), device=dev, dtype=dtype) for _ in range(2)]
        d_out = [torch.empty((1, K, 9, T), device=dev, dtype=dtype) for _ in range(2)]

        def e2e_step():
            ev_in = [None, None]
            ev_cmp = [None, None]
            ev_out = [None, None]
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
                    singlecam_smooth_sessions(d_in[b], dtype=dtype, out=d_out[b])
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
        e2e = {'value': world * kf_step / (ms_e2e * 1e-3), 'unit': 'keypoint-frames/s',
               'h2d_bytes_per_step': int(S * M * T * K * 3 * w), 'd2h_bytes_per_step': int(S * K * 9 * T * w),
               'ms_per_step': ms_e2e,
               'how': 'per session: pinned H2D -> eks_b200.pipeline.singlecam_smooth_sessions -> pinned D2H, '
                      'double-buffered on three streams'}
        del h_in, h_out, d_in, d_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the persistent Adam / NLL kernel)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (of measured)'
    else:
        peak, peak_src = 6650.0, 'fallback 6.65 TB/s (of fallback)'
    opt_ms = stage_ms.get('optimize_s')
    # dominant kernel: diag_nll_kernel, launched once per Adam evaluation.  Algorithmic bytes of one launch:
    # every evaluation reads the observation planes of each still-active sequence once (SURVEY 8d:
    # w * obs bytes per keypoint-frame per evaluation).  Launch duration = CUDA-event time of the optimiser
    # stage / number of launches that had work (the interleaved one-warp Adam kernels, ~2% of the stage per
    # the ncu launch list, are included => the fraction is slightly pessimistic).
    n_launch = max(1, n_eval_max)
    alg_bytes = float(iters.sum().item()) * T * 2 * w / n_launch
    roofline = None
    if opt_ms:
        launch_ms = opt_ms / n_launch
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):   # dram bytes per algorithmic byte from the committed ncu --set full capture
            tj = json.load(open(tpath)).get('diag_nll_kernel')
            if tj:
                traffic = alg_bytes * tj['dram_bytes'] / tj['algorithmic_bytes']
        roofline = {'bound': 'hbm', 'kernel': 'diag_nll_kernel', 'achieved': achieved, 'peak': peak,
                    'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': alg_bytes, 'launches': n_launch, 'launch_ms': launch_ms,
                    'share_of_step': opt_ms / ms_step}
    b_alg = w * (3 * M + 9)
    pipeline_frac = (kf_step * b_alg / (ms_step * 1e-3) / 1e9) / peak

    cpu = None
    if not args.no_cpu and world == 1:   # the CPU baseline is a rank-0, N = 1 leg (torchrun pins OMP to one thread)
        cpu = cpu_leg(M, K, min(args.cpu_frames, T), 1, 0)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

    line = {
        'metric': 'keypoint-frames/sec smoothed incl. s-optimisation', 'value': value, 'unit': 'keypoint-frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: {S} sessions/GPU x {M} seeds x {K} keypoints x {T} frames, '
                               'singlecam, per-keypoint Adam s-optimisation (reference stop rule)',
                   'sessions_per_gpu': S, 'seeds': M, 'keypoints': K, 'frames': T,
                   'l2': 'inputs larger than L2 (resident raw tensor %.1f GB per GPU)' % (raw.numel() * w / 1e9),
                   'n_eval_mean': n_eval_mean, 'n_eval_max': n_eval_max},
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline,
        'pipeline_one_touch': {'bytes_per_kf': b_alg, 'frac_of_hbm_peak': pipeline_frac,
                               'achieved_gbs': kf_step * b_alg / (ms_step * 1e-3) / 1e9},
        'stage_ms': stage_ms, 'cpu_baseline': cpu,
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
