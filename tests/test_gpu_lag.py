"""GPU tests of the lag-statistics optimiser (eks_b200/csrc/diag_lag.cu): the closed-form NLL from the lagged products of
the increments against (a) the streaming evaluation of the same loss (opt_mode='stream', one exact time-parallel pass
per Adam iteration) and (b) the CPU oracle, including the cases where the closed form must NOT be used (slow forgetting,
A != 1, short sequences, shared-s blocks) and the evaluation streams inside the persistent kernel."""
import numpy as np
import pytest
import torch

from conftest import synth_singlecam
from parity import check_columns

pytestmark = pytest.mark.gpu
TRACE = 300


def _run(raw, dtype, **kw):
    from eks_b200.pipeline import singlecam_smooth_sessions
    t = torch.as_tensor(raw).cuda().to(dtype)
    res = singlecam_smooth_sessions(t[None], dtype=dtype, trace_cap=TRACE, **kw)
    torch.cuda.synchronize()
    trace = singlecam_smooth_sessions.last_opt['trace'].double().cpu().numpy()
    return res, trace


def _smooth_walk(M, K, T, seed, step, noise):
    """random walk with small steps under large observation noise: s / r << 1, i.e. a slowly forgetting filter"""
    rng = np.random.default_rng(seed)
    truth = np.cumsum(rng.normal(0, step, size=(T, K, 2)), axis=0) + rng.uniform(50, 300, size=(1, K, 2))
    pred = truth[None] + rng.normal(size=(M, T, K, 2)) * noise
    lik = rng.uniform(0.9, 1.0, size=(M, T, K))
    return np.ascontiguousarray(np.concatenate([pred, lik[..., None]], axis=-1)[:, None].astype(np.float32).astype(np.float64))


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_lag_equals_stream_on_bench_like_data(dtype):
    """same data family as bench.py (fast forgetting): every evaluation takes the closed form; the whole Adam trace
    must coincide with the streaming evaluation's."""
    raw = synth_singlecam(M=6, K=3, T=60_000, seed=7)
    r_lag, t_lag = _run(raw, dtype, opt_mode='lag')
    r_str, t_str = _run(raw, dtype, opt_mode='stream')
    it_l, it_s = r_lag.iters[0].cpu().numpy(), r_str.iters[0].cpu().numpy()
    if dtype == torch.float64:
        assert list(it_l) == list(it_s)
        for k in range(3):
            n = it_l[k]
            np.testing.assert_allclose(t_lag[k, :n, 1], t_str[k, :n, 1], rtol=1e-11)      # loss at every iterate
            np.testing.assert_allclose(t_lag[k, :n, 2], t_str[k, :n, 2], rtol=1e-6, atol=1e-9 * np.abs(t_str[k, :n, 2]).max())
        np.testing.assert_allclose(r_lag.s_finals.cpu().numpy(), r_str.s_finals.cpu().numpy(), rtol=1e-8)
    else:
        # float32 mode: the lag path evaluates the float64 loss of the float32 data, the streaming path float32
        # arithmetic with float64 sums: compare the common iterates
        for k in range(3):
            n = min(it_l[k], it_s[k])
            np.testing.assert_allclose(t_lag[k, :n, 0], t_str[k, :n, 0], atol=1e-4)
            np.testing.assert_allclose(t_lag[k, :n, 1], t_str[k, :n, 1], rtol=3e-6)


@pytest.mark.parametrize('step,noise,label', [(0.3, 0.5, 'fast'), (0.05, 0.6, 'mixed'), (0.01, 0.5, 'slow')])
def test_forgetting_regimes_match_oracle_fp64(step, noise, label):
    """fast: closed form throughout; mixed: the excursion to small s streams inside the persistent kernel, the rest is
    closed form; slow: alpha ~ 0.98, every evaluation streams.  All must reproduce the oracle's iteration counts."""
    from oracle import oracle
    raw = _smooth_walk(M=4, K=2, T=12_000, seed=11, step=step, noise=noise)
    ref = oracle.singlecam(raw, dtype=np.float64)
    res, _ = _run(raw, torch.float64)
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters']), label
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=1e-5)
    out = res.out[0].permute(2, 0, 1).double().cpu().numpy()
    check_columns(out, ref['out'], 1e-5, label)


def test_blocks_long_sequences_match_oracle():
    """shared-s blocks (eks/core.py:403-559) through the closed form: member losses summed in order."""
    from oracle import oracle
    raw = synth_singlecam(M=4, K=5, T=9000, seed=23)
    blocks = [[0, 3], [1, 2, 4]]
    ref = oracle.singlecam(raw, dtype=np.float64, blocks=blocks)
    res, _ = _run(raw, torch.float64, blocks=blocks)
    s = res.s_finals[0].cpu().numpy()
    assert s[0] == s[3] and s[1] == s[2] == s[4]
    assert list(res.iters[0].cpu().numpy()) == [int(ref['info']['iters'][j]) for j in (0, 1, 1, 0, 1)]
    np.testing.assert_allclose(s, ref['s_finals'], rtol=1e-5)


def test_spans_and_offsets_match_oracle():
    """one s_frames span: the statistics start at t_begin + T0 of the cropped sequence"""
    from oracle import oracle
    raw = synth_singlecam(M=4, K=2, T=15_000, seed=31)
    ref = oracle.singlecam(raw, dtype=np.float64, s_frames=[(1203, 14001)])
    res, _ = _run(raw, torch.float64, spans=[(1203, 14001)])
    assert list(res.iters[0].cpu().numpy()) == list(ref['info']['iters'])
    np.testing.assert_allclose(res.s_finals[0].cpu().numpy(), ref['s_finals'], rtol=1e-5)


def test_diagonal_dynamics_other_than_identity_stream():
    """run_kalman_smoother with A = 0.97 I, C = 1.1 I (decoupled but not the unit-DC-gain model): the closed form does
    not apply, the persistent kernel streams every evaluation; result = oracle."""
    import eks_b200
    from oracle import oracle
    rng = np.random.default_rng(5)
    K, T = 2, 6000
    lat = np.cumsum(rng.normal(0, 0.3, (K, T, 2)), axis=1) * 0.2
    ev = rng.uniform(0.05, 0.6, (T, K, 2))
    ys = 1.1 * lat + rng.standard_normal((K, T, 2)) * np.sqrt(np.swapaxes(ev, 0, 1))
    ys -= ys.mean(axis=1, keepdims=True)
    eye = np.tile(np.eye(2), (K, 1, 1))
    S0s = np.stack([np.diag(ys[k].var(axis=0)) for k in range(K)])
    eks_b200.set_precision('float64')
    try:
        s, ms, Vs = eks_b200.run_kalman_smoother(ys, np.zeros((K, 2)), S0s, 0.97 * eye, 1.1 * eye, eye, ev)
        s_o, ms_o, Vs_o, info = oracle.run_kalman_smoother(ys, np.zeros((K, 2)), S0s, 0.97 * eye, 1.1 * eye, eye, ev,
                                                           dtype=np.float64)
        np.testing.assert_allclose(s, s_o, rtol=1e-5)
        np.testing.assert_allclose(ms, ms_o, rtol=1e-5, atol=1e-7)
    finally:
        eks_b200.set_precision('float32')


# ----------------------------------------------------------------------------------------------- fused final pass
def _out(res):
    return res.out[0].permute(2, 0, 1).double().cpu().numpy()


@pytest.mark.parametrize('dtype,rtol', [(torch.float64, 1e-11), (torch.float32, 2e-5)])
def test_fused_smoother_equals_exact_scan(dtype, rtol):
    """the fused time-segmented final pass (halo frames absorb the boundary states, verified) against the exact
    Moebius / affine scan kernels over the whole sequence, at the same s"""
    raw = synth_singlecam(M=6, K=3, T=50_001, seed=13)
    s = [0.05, 0.2, 1.5]
    from eks_b200.pipeline import singlecam_smooth_sessions
    t = torch.as_tensor(raw).cuda().to(dtype)
    fused = singlecam_smooth_sessions(t[None], dtype=dtype, smooth_param=s)
    exact = singlecam_smooth_sessions(t[None], dtype=dtype, smooth_param=s, exact_scan=True)
    torch.cuda.synchronize()
    check_columns(_out(fused), _out(exact), rtol, f'fused vs exact {dtype}')


def test_fused_smoother_slow_forgetting_falls_back():
    """s / r tiny: the halos cannot absorb the boundary states, the verification flags every sequence and the exact
    kernels redo them -- the result is still the oracle's"""
    from oracle import oracle
    raw = _smooth_walk(M=4, K=2, T=30_000, seed=3, step=0.002, noise=0.8)
    s = [1e-3, 3e-3]
    ref = oracle.singlecam(raw, dtype=np.float64, smooth_param=s)
    from eks_b200.pipeline import singlecam_smooth_sessions
    t = torch.as_tensor(raw).cuda()[None]
    res = singlecam_smooth_sessions(t, dtype=torch.float64, smooth_param=s)
    exact = singlecam_smooth_sessions(t, dtype=torch.float64, smooth_param=s, exact_scan=True)
    torch.cuda.synchronize()
    assert torch.equal(res.out, exact.out), 'flagged sequences must come out of the exact scan kernels bit for bit'
    check_columns(_out(res), ref['out'], 1e-5, 'slow forgetting')


def test_fused_smoother_nan_observation_propagates_like_the_oracle():
    """a frame where every seed is NaN makes the ensemble median NaN: the reference's filter turns everything after it
    (and, through the RTS pass, everything before it) into NaN; the fused pass must not confine that to one segment"""
    from oracle import oracle
    raw = synth_singlecam(M=3, K=2, T=20_000, seed=5)
    raw[:, 0, 7777, 1, 0:2] = np.nan     # both coordinates (the reference's dense 2-D filter couples them through NaN)
    ref = oracle.singlecam(raw, dtype=np.float64, smooth_param=[0.1, 0.1])
    from eks_b200.pipeline import singlecam_smooth_sessions
    res = singlecam_smooth_sessions(torch.as_tensor(raw).cuda()[None], dtype=torch.float64, smooth_param=[0.1, 0.1])
    torch.cuda.synchronize()
    out = _out(res)
    for c in (0, 1, 7, 8):
        np.testing.assert_array_equal(np.isnan(out[..., c]), np.isnan(ref['out'][..., c]))
    ok = ~np.isnan(ref['out'])
    np.testing.assert_allclose(out[ok], ref['out'][ok], rtol=1e-5, atol=1e-7)
