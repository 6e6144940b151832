"""GPU parity at the BASELINE.json shapes (VERDICT r1 "next round" item 1): the computation bench.py times, compared
with the CPU oracle at 10^5 - 10^6 frames, in float64 AND in float32 (tests/parity.py explains the float32 protocol).

  c2  singlecam, 5 seeds x 17 keypoints x 100 000 frames (full config)
  c5  singlecam, 10 seeds x 2 keypoints x 1 000 000 frames (a 2-keypoint sample of one session; full frame count)
  c3  multicam linear, 2 cameras x 4 keypoints x 10 seeds x 200 000 frames (the oracle's dense D=3/O=4 recursion costs
      ~0.3 ms per frame-iteration on a host core; 10^6 frames would take 10 minutes)
  c4  calibrated pinhole EKF on the fly rig, 3 cameras x 2 keypoints x 5 seeds x 100 000 frames

Data = bench.py's own host generators (same distribution as the timed run).  float64: identical Adam iteration
counts, s within 1e-5, every output column within 1e-5.  float32: fp32_stop_protocol against the float64 oracle trace
(loss within a few float32 ulps at every iterate, coinciding trajectories, disputed stop = rounding knife edge),
outputs within 1e-3 of the float64 oracle evaluated at the product's own s; |ds|/s against both oracles is printed.
"""
import os

import numpy as np
import pytest
import torch

import bench
from conftest import GOLDEN
from parity import check_columns, fp32_stop_protocol

pytestmark = pytest.mark.gpu
RTOL64, RTOL32 = 1e-5, 1e-3
TRACE = 300
_cache = {}


def _oracle_singlecam(tag, raw, dtype):
    from oracle import oracle
    key = (tag, np.dtype(dtype).name)
    if key not in _cache:
        _cache[key] = oracle.singlecam(raw.astype(np.float64) if dtype == np.float64 else raw, dtype=dtype,
                                       trace_cap=TRACE)
    return _cache[key]


def _oracle_multicam(tag, raw, dtype, **kw):
    from oracle import oracle
    key = (tag, np.dtype(dtype).name)
    if key not in _cache:
        _cache[key] = oracle.multicam(raw.astype(np.float64), dtype=dtype, trace_cap=TRACE, **kw)
    return _cache[key]


def _raw(tag):
    if tag not in _cache:
        if tag == 'c2':
            _cache[tag] = bench.synth_session_host(5, 17, 100_000, seed=2)
        elif tag == 'c5':
            _cache[tag] = bench.synth_session_host(10, 2, 1_000_000, seed=5)
        elif tag == 'c3':
            _cache[tag] = bench.synth_multicam_host(10, 2, 4, 200_000, seed=3)
        elif tag == 'c4':
            _cache[tag] = bench.synth_fly_host(5, 2, 100_000, seed=4, cams=bench.fly_cameras()[0])
    return _cache[tag]


def _run_singlecam(raw, dtype, **kw):
    from eks_b200.pipeline import singlecam_smooth_sessions
    t = torch.as_tensor(raw).cuda().to(dtype)
    res = singlecam_smooth_sessions(t[None], dtype=dtype, trace_cap=TRACE, **kw)
    torch.cuda.synchronize()
    out = res.out[0].permute(2, 0, 1).double().cpu().numpy()               # (T,K,9)
    trace = None
    if res.iters is not None:
        trace = singlecam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    return out, res, trace


def _run_multicam(raw, dtype, **kw):
    from eks_b200.pipeline import multicam_smooth_sessions
    t = torch.as_tensor(raw).cuda().to(dtype)
    res = multicam_smooth_sessions(t[None], dtype=dtype, trace_cap=TRACE, **kw)
    torch.cuda.synchronize()
    out = res.out[0].permute(1, 3, 0, 2).double().cpu().numpy()            # (V,T,K,9)
    trace = None
    if res.iters is not None:
        trace = multicam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    return out, res, trace


def _report_ds(label, s_gpu, ref64, ref32):
    d64 = np.abs(s_gpu - ref64['s_finals']) / ref64['s_finals']
    d32 = np.abs(s_gpu - ref32['s_finals']) / ref32['s_finals']
    o = np.abs(ref32['s_finals'] - ref64['s_finals']) / ref64['s_finals']
    print(f'[parity fp32] {label}: |ds|/s vs fp64 oracle max {d64.max():.3e} mean {d64.mean():.3e}; vs fp32 oracle max '
          f'{d32.max():.3e}; fp32 oracle vs fp64 oracle max {o.max():.3e}; iterations gpu/fp64/fp32 oracle')


# ----------------------------------------------------------------------------------------------- singlecam c2, c5
@pytest.mark.parametrize('tag', ['c2', 'c5'])
def test_singlecam_fp64(tag):
    raw = _raw(tag)
    ref = _oracle_singlecam(tag, raw, np.float64)
    out, res, _ = _run_singlecam(raw.astype(np.float64), torch.float64)
    it = res.iters[0].cpu().numpy()
    assert list(it) == list(ref['info']['iters']), f'{tag}: iterations {list(it)} vs oracle {list(ref["info"]["iters"])}'
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    check_columns(out, ref['out'], RTOL64, f'{tag} fp64')


@pytest.mark.parametrize('tag', ['c2', 'c5'])
def test_singlecam_fp32(tag):
    from oracle import oracle
    raw = _raw(tag)
    ref64 = _oracle_singlecam(tag, raw, np.float64)
    ref32 = _oracle_singlecam(tag, raw, np.float32)
    out, res, trace = _run_singlecam(raw, torch.float32)
    it = res.iters[0].cpu().numpy()
    K = raw.shape[3]
    for k in range(K):
        fp32_stop_protocol(f'{tag} kp{k}', trace[k], it[k], ref64['info']['trace'][k], ref64['info']['iters'][k])
    s_gpu = res.s_finals[0].cpu().numpy()
    _report_ds(tag, s_gpu, ref64, ref32)
    print('   ', list(it), list(ref64['info']['iters']), list(ref32['info']['iters']))
    at_s = oracle.singlecam(raw.astype(np.float64), smooth_param=list(s_gpu), dtype=np.float64)
    check_columns(out, at_s['out'], RTOL32, f'{tag} fp32 at the product s')


# ----------------------------------------------------------------------------------------------- multicam linear c3
def test_multicam_linear_fp64():
    raw = _raw('c3')
    ref = _oracle_multicam('c3', raw, np.float64, quantile_keep_pca=50.0)
    out, res, _ = _run_multicam(raw.astype(np.float64), torch.float64, quantile_keep_pca=50.0)
    it = res.iters[0].cpu().numpy()
    assert list(it) == list(ref['info']['iters']), f'c3: iterations {list(it)} vs oracle {list(ref["info"]["iters"])}'
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    check_columns(out, ref['cam_out'], RTOL64, 'c3 fp64')


def test_multicam_linear_fp32():
    from oracle import oracle
    raw = _raw('c3')
    ref64 = _oracle_multicam('c3', raw, np.float64, quantile_keep_pca=50.0)
    ref32 = _oracle_multicam('c3', raw, np.float32, quantile_keep_pca=50.0)
    out, res, trace = _run_multicam(raw, torch.float32, quantile_keep_pca=50.0)
    it = res.iters[0].cpu().numpy()
    for k in range(raw.shape[3]):
        fp32_stop_protocol(f'c3 kp{k}', trace[k], it[k], ref64['info']['trace'][k], ref64['info']['iters'][k])
    s_gpu = res.s_finals[0].cpu().numpy()
    _report_ds('c3', s_gpu, ref64, ref32)
    print('   ', list(it), list(ref64['info']['iters']), list(ref32['info']['iters']))
    at_s = oracle.multicam(raw.astype(np.float64), quantile_keep_pca=50.0, dtype=np.float64, smooth_param=list(s_gpu))
    check_columns(out, at_s['cam_out'], RTOL32, 'c3 fp32 at the product s')


# ----------------------------------------------------------------------------------------------- calibrated EKF c4
def test_multicam_pinhole_fp64():
    raw = _raw('c4')
    cams = bench.fly_cameras()[0]
    ref = _oracle_multicam('c4', raw, np.float64, camgroup=os.path.join(GOLDEN, 'fly_calibration.toml'))
    out, res, _ = _run_multicam(raw.astype(np.float64), torch.float64, cams=cams)
    it = res.iters[0].cpu().numpy()
    assert list(it) == list(ref['info']['iters']), f'c4: iterations {list(it)} vs oracle {list(ref["info"]["iters"])}'
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=RTOL64)
    check_columns(out, ref['cam_out'], RTOL64, 'c4 fp64')


def test_multicam_pinhole_fp32():
    from oracle import oracle
    raw = _raw('c4')
    cams = bench.fly_cameras()[0]
    cal = os.path.join(GOLDEN, 'fly_calibration.toml')
    ref64 = _oracle_multicam('c4', raw, np.float64, camgroup=cal)
    ref32 = _oracle_multicam('c4', raw, np.float32, camgroup=cal)
    out, res, trace = _run_multicam(raw, torch.float32, cams=cams)
    it = res.iters[0].cpu().numpy()
    for k in range(raw.shape[3]):
        fp32_stop_protocol(f'c4 kp{k}', trace[k], it[k], ref64['info']['trace'][k], ref64['info']['iters'][k])
    s_gpu = res.s_finals[0].cpu().numpy()
    _report_ds('c4', s_gpu, ref64, ref32)
    print('   ', list(it), list(ref64['info']['iters']), list(ref32['info']['iters']))
    at_s = oracle.multicam(raw.astype(np.float64), camgroup=cal, dtype=np.float64, smooth_param=list(s_gpu))
    check_columns(out, at_s['cam_out'], RTOL32, 'c4 fp32 at the product s')
