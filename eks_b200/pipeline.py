"""Device-resident single-camera EKS pipeline, batched over sessions.

This is the B200-native form of ensemble_kalman_smoother_singlecam (eks/singlecam_smoother.py:105-243):
the raw seed predictions of S sessions stay on the device as one tensor, every stage is a kernel of
libeks_b200.so, and the result is ONE output block of frame-major planes

    out[s][k][c][t],  c indexing ops.OUT_COLS = the reference's nine output columns,

so that each input value is read once and each output value written once (data layout).
Stages: ensemble statistics (+ fused centring moments) -> initial guess / median R -> Adam on log s
(device-resident loop) -> filter + RTS smoother writing the smoothed / posterior-variance planes.
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from eks_b200 import ops
from eks_b200._lib import lib
from eks_b200.ops import Model, PlaneView


@dataclass
class SinglecamResult:
    out: torch.Tensor        # (S, K, 9, T) planes, see ops.OUT_COLS
    s_finals: torch.Tensor   # (S, K) float64
    iters: torch.Tensor | None    # (S, K) int32 (None when smooth_param was given)
    loss: torch.Tensor | None     # (S, K) last evaluated NLL
    means: torch.Tensor      # (S, K, 2) centring offsets


_SIDE_STREAMS: dict = {}
# stage -> stream priority inside a session group (0 = default, more negative = dispatched first).  Measured on B200
# (8 sessions x 20 keypoints x 10^6 frames, scripts/groups_bench.py): one stream 11.20 ms/step; 4 groups with equal
# priorities 10.96; "later stages outrank earlier ones" (0,-1,-3,-2) 11.66; only the Adam kernel raised 10.94;
# reversed 12.62.  The streaming kernels each fill the SMs on their own, so priorities only add hand-over gaps:
# all stages stay at the default priority and the mechanism is kept for experiments (EKS_STAGE_PRIO).
_STAGE_PRIO = {'ensemble': 0, 'prestage': 0, 'optimize_s': 0, 'filter_smooth': 0}
if os.environ.get('EKS_STAGE_PRIO'):   # experiments: "ensemble,prestage,optimize_s,filter_smooth" priorities
    _STAGE_PRIO = dict(zip(_STAGE_PRIO, [int(x) for x in os.environ['EKS_STAGE_PRIO'].split(',')]))


class _Chain:
    """All work of one session group is ONE dependency chain that hops between the group's streams of different
    priority (`stage`); memory stays safe because every launch of the group is ordered after all earlier ones."""

    def __init__(self, dev: torch.device, g: int):
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -5)
        key = (dev.type, dev.index, g)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = {p: torch.cuda.Stream(device=dev, priority=max(p, min(lo, hi)))
                                  for p in sorted(set(_STAGE_PRIO.values()), reverse=True)}
        self.streams = _SIDE_STREAMS[key]
        self.base = self.streams[0]
        self.where = self.base

    def hop(self, st):
        if st is not self.where:
            ev = torch.cuda.Event()
            ev.record(self.where)
            st.wait_event(ev)
            self.where = st

    def stage(self, name):
        chain, st = self, self.streams[_STAGE_PRIO.get(name, -1)]

        class _Ctx:
            def __enter__(self_):
                chain.hop(st)
                self_.inner = torch.cuda.stream(st)
                self_.inner.__enter__()

            def __exit__(self_, *exc):
                self_.inner.__exit__(*exc)
                chain.hop(chain.base)      # torch ops between the stages run on the group's base stream, in order
        return _Ctx()


class _NoChain:
    def stage(self, name):
        import contextlib
        return contextlib.nullcontext()

    def hop(self, st):
        pass


def _auto_groups(S: int, K: int, T: int) -> int:
    """Session groups of one call.  The stages of the path alternate between HBM-bound (ensemble, final pass),
    FP32-pipe-bound (lag statistics) and latency-bound (median select, the persistent Adam kernel: one warp per
    sequence for ~120 dependent evaluations) work; independent session groups on separate streams let one group's
    latency-bound stage run underneath another group's streaming stage.  A group must still fill the GPU on its own in
    the streaming stages (>= ~16 sequences of >= 10^5 frames)."""
    if S < 2 or K * T < 1_600_000:
        return 1
    return min(4, S)


def singlecam_smooth_sessions(raw: torch.Tensor, smooth_param=None, spans=None, blocks=None,
                              avg_mode='median', var_mode='confidence_weighted_var', dtype=torch.float32,
                              lr=0.25, s_bounds_log=(-8.0, 8.0), tol=1e-2, safety_cap=300, min_R_var=1e-4,
                              out: torch.Tensor | None = None, force_generic: bool = False,
                              trace_cap: int = 0, timers: dict | None = None,
                              opt_mode: str = 'lag', exact_scan: bool = False,
                              n_groups: int | None = None, _chain=None) -> SinglecamResult:
    """raw: (S, M, 1, T, K, 3) CUDA tensor (float32 or float64) in the MarkerArray layout.

    spans: validated [(start, end)] list (s_frames) or None; blocks: per-session keypoint blocks.
    opt_mode: 'lag' (one pass over the observations, closed-form NLL from lag statistics -- the default) or 'stream'
    (one streaming pass per Adam evaluation); both are the same optimisation (tests compare them).
    n_groups: sessions are independent problems; n_groups > 1 runs groups of sessions on separate streams (forked
    from and joined to the caller's stream) so that their stages overlap.  None = automatic (1 for small inputs and
    whenever per-stage timers or an optimiser trace are requested)."""
    assert raw.is_cuda and raw.dim() == 6 and raw.shape[2] == 1 and raw.shape[-1] == 3
    S_all = raw.shape[0]
    if n_groups is None:
        n_groups = 1 if (timers is not None or trace_cap or force_generic) else _auto_groups(S_all, raw.shape[4],
                                                                                           raw.shape[3])
    n_groups = max(1, min(int(n_groups), S_all))
    if n_groups > 1:
        raw = raw.contiguous()
        dev = raw.device
        if out is None:
            out = torch.empty((S_all, raw.shape[4], 9, raw.shape[3]), dtype=dtype, device=dev)
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        bounds = [(g * S_all) // n_groups for g in range(n_groups + 1)]
        parts = []
        for g in range(n_groups):
            chain = _Chain(dev, g)
            chain.base.wait_event(fork)
            with torch.cuda.stream(chain.base):
                parts.append(singlecam_smooth_sessions(
                    raw[bounds[g]:bounds[g + 1]], smooth_param=smooth_param, spans=spans, blocks=blocks,
                    avg_mode=avg_mode, var_mode=var_mode, dtype=dtype, lr=lr, s_bounds_log=s_bounds_log, tol=tol,
                    safety_cap=safety_cap, min_R_var=min_R_var, out=out[bounds[g]:bounds[g + 1]], opt_mode=opt_mode,
                    exact_scan=exact_scan, n_groups=1, _chain=chain))
            chain.hop(chain.base)
            main.wait_stream(chain.base)
        cat = lambda xs: None if xs[0] is None else torch.cat(xs, dim=0)
        return SinglecamResult(out, cat([p.s_finals for p in parts]), cat([p.iters for p in parts]),
                               cat([p.loss for p in parts]), cat([p.means for p in parts]))

    class _Stage:  # optional CUDA-event bracket per stage on the launching stream (bench.py)
        def __init__(self, name):
            self.name = name

        def __enter__(self):
            if timers is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *exc):
            if timers is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                timers.setdefault(self.name, []).append((self.e0, e1))
    raw = raw.contiguous()
    S, M, _, T, K, _ = raw.shape
    dev = raw.device
    B = S * K
    chain = _chain if _chain is not None else _NoChain()
    if out is None:
        out = torch.empty((S, K, 9, T), dtype=dtype, device=dev)
    assert out.shape == (S, K, 9, T) and out.dtype == dtype and out.is_contiguous()
    # 1) ensemble statistics straight into the output planes + centring moments
    plane_off = [c * T for c in ops.ENS_TO_OUT]
    with _Stage('ensemble'), chain.stage('ensemble'):
        partials = ops.ensemble_stats(raw, out, K * 9 * T, 0, 9 * T, plane_off, avg_mode=avg_mode,
                                      var_mode=var_mode, moments=True)
    pre = chain.stage('prestage')
    pre.__enter__()
    with _Stage('center_moments'):
        ymean, yvar = ops.center_moments(partials, T, dtype)      # (B,2) each
    # 2) model: m0 = 0, S0 = diag(var), A = C = Q = I (singlecam_smoother.py:246-284)
    eye = torch.eye(2, dtype=dtype, device=dev).expand(B, 2, 2).contiguous()
    model = Model(torch.zeros((B, 2), dtype=dtype, device=dev), torch.diag_embed(yvar).contiguous(), eye, eye, eye)
    yv = PlaneView(out, 9 * T, [3 * T, 4 * T])
    vv = PlaneView(out, 9 * T, [5 * T, 6 * T])
    iters = loss = None
    if smooth_param is not None:
        s = torch.as_tensor(smooth_param, dtype=torch.float64, device=dev)
        s_finals = (s.expand(K) if s.dim() == 0 or s.numel() == 1 else s).expand(S, K).contiguous()
        pre.__exit__(None, None, None)
    else:
        with _Stage('initial_guess'):
            guess, s_log0 = ops.initial_guess(vv, B, T)
        with _Stage('const_R_median'):
            Rconst = ops.const_R_median(vv, B, T, spans=spans, min_var=min_R_var)
        all_blocks = None
        if blocks:
            all_blocks = [[s_ * K + k for k in blk] for s_ in range(S) for blk in blocks]
            covered = {k for blk in blocks for k in blk}
            all_blocks += [[s_ * K + k] for s_ in range(S) for k in range(K) if k not in covered]
            g = guess.view(-1)
            s0 = torch.stack([g[torch.as_tensor(b, device=dev)].mean() for b in all_blocks])
            s_log0 = torch.log(s0.clamp(1e-6, 1e3)).float().to(dtype)
        pre.__exit__(None, None, None)
        with _Stage('optimize_s'), chain.stage('optimize_s'):
            opt = ops.optimize_s(model, yv, T, Rconst, s_log0, blocks=all_blocks, ymean=ymean, spans=spans, lr=lr,
                                 s_bounds_log=s_bounds_log, tol=tol, safety_cap=safety_cap, trace_cap=trace_cap,
                                 structure=ops.STRUCT_GENERAL if force_generic else (
                                     ops.STRUCT_DIAG_STREAM if opt_mode == 'stream' else ops.STRUCT_DIAG))
        s_blk = torch.exp(opt['s_log'].double().clamp(s_bounds_log[0], s_bounds_log[1]))
        if all_blocks is None:
            s_finals = s_blk.view(S, K)
            iters, loss = opt['iters'].view(S, K), opt['loss'].view(S, K)
        else:
            s_flat = torch.empty(B, dtype=torch.float64, device=dev)
            it_flat = torch.empty(B, dtype=torch.int32, device=dev)
            lo_flat = torch.empty(B, dtype=dtype, device=dev)
            for j, b in enumerate(all_blocks):
                idx = torch.as_tensor(b, device=dev)
                s_flat[idx] = s_blk[j]
                it_flat[idx] = opt['iters'][j]
                lo_flat[idx] = opt['loss'][j]
            s_finals, iters, loss = s_flat.view(S, K), it_flat.view(S, K), lo_flat.view(S, K)
        singlecam_smooth_sessions.last_opt = opt
    # 3) final filter + RTS smoother (time-varying R_t), outputs into planes 0,1,7,8
    if not force_generic:
        with _Stage('filter_smooth'), chain.stage('filter_smooth'):
            s_dev = s_finals.reshape(B).to(dtype)
            ops.diag_smooth(model, yv, vv, T, s_dev, ymean, out, 9 * T, [0, T, 7 * T, 8 * T], exact_scan=exact_scan)
    else:
        s_dev = s_finals.reshape(B).to(dtype)
        ms, Vs = ops.filter_smooth(model, yv, vv, T, s_dev, ymean=ymean)
        o = out.view(B, 9, T)
        o[:, 0, :] = ms[:, :, 0] + ymean[:, 0:1]
        o[:, 1, :] = ms[:, :, 1] + ymean[:, 1:2]
        o[:, 7, :] = Vs[:, :, 0, 0]
        o[:, 8, :] = Vs[:, :, 1, 1]
    return SinglecamResult(out, s_finals, iters, loss, ymean.view(S, K, 2))


# ===================================================================================================
# Multi-camera (linear PCA latent) pipeline, device resident
# ===================================================================================================
@dataclass
class MulticamResult:
    out: torch.Tensor        # (S, K, V, 9, T) planes per camera, see ops.OUT_COLS
    ms: torch.Tensor         # (S*K, T, L) smoothed latent means
    Vs: torch.Tensor         # (S*K, T, L, L) smoothed latent covariances
    s_finals: torch.Tensor   # (S, K) float64
    iters: torch.Tensor | None
    loss: torch.Tensor | None
    ymean: torch.Tensor      # (S*K, 2V) centring offsets
    C: torch.Tensor          # (S*K, 2V, L) observation matrices (PCA components)
    S0: torch.Tensor         # (S*K, L, L)
    Q: torch.Tensor          # (S*K, L, L)
    n_good: torch.Tensor     # (S*K, 2) int32: good frames, frames used for the PCA fit


def pca_from_moments(mom, O: int, L: int):
    """sklearn PCA(n_components=L).fit restated on the sufficient statistics (n, sum x, sum x x^T):
    covariance_eigh branch of sklearn/decomposition/_pca.py::_fit_full + svd_flip(u_based_decision=False).
    mom: (B, 1+O+O*O) float64 numpy.  Returns (pca_mean (B,O), components (B,L,O)) float64."""
    import numpy as np
    B = mom.shape[0]
    means = np.empty((B, O))
    comps = np.empty((B, L, O))
    for b in range(B):
        n = mom[b, 0]
        mean = mom[b, 1:1 + O] / n
        cov = (mom[b, 1 + O:].reshape(O, O) - n * np.outer(mean, mean)) / (n - 1.0)
        _, vec = np.linalg.eigh(cov)
        vt = vec[:, ::-1].T[:L].copy()
        idx = np.argmax(np.abs(vt), axis=1)
        sign = np.sign(vt[np.arange(L), idx])
        sign[sign == 0] = 1.0
        means[b], comps[b] = mean, vt * sign[:, None]
    return means, comps


def fa_from_moments(mom_b, O: int, L: int, tol: float = 1e-2, max_iter: int = 1000):
    """sklearn FactorAnalysis(n_components=L).fit restated on the sufficient statistics (n, sum x, sum x x^T)
    (sklearn/decomposition/_factor_analysis.py::fit): the SVD of X / (sqrt(psi) sqrt(n)) is taken through the
    eigen-decomposition of its O x O Gram matrix.  sklearn's default randomized solver sketches n_components + 10
    columns with 3 power iterations: for O <= n_components + 10 (up to 6 cameras at n_latent = 3) the sketch spans every
    feature and the two agree to rounding; for 7-8 cameras (O = 14, 16) sklearn's result carries the sketch's own
    approximation error and seed dependence, this one is the exact decomposition it approximates.
    Returns (mean (O,), loading (O,L) = components_.T) float64."""
    import numpy as np
    n = mom_b[0]
    mean = mom_b[1:1 + O] / n
    cov = mom_b[1 + O:].reshape(O, O) / n - np.outer(mean, mean)
    var = np.diag(cov).copy()
    psi = np.ones(O)
    old_ll, small = -np.inf, 1e-12
    llconst = O * np.log(2.0 * np.pi) + L
    W = np.zeros((L, O))
    for _ in range(max_iter):
        sqrt_psi = np.sqrt(psi) + small
        gram = cov / np.outer(sqrt_psi, sqrt_psi)
        lam, vec = np.linalg.eigh(gram)
        s = np.maximum(lam[::-1][:L], 0.0)
        vt = vec[:, ::-1][:, :L].T
        unexp_var = np.trace(gram) - s.sum()
        W = np.sqrt(np.maximum(s - 1.0, 0.0))[:, None] * vt * sqrt_psi
        with np.errstate(divide='ignore'):
            ll = (llconst + np.sum(np.log(s)) + unexp_var + np.sum(np.log(psi))) * (-n / 2.0)
        if ll - old_ll < tol:
            break
        old_ll = ll
        psi = np.maximum(var - np.sum(W ** 2, axis=0), small)
    return mean, W.T.copy()


def mc_inflate_variances(yv: PlaneView, vv: PlaneView, ymean: torch.Tensor, T: int, ws: torch.Tensor, n_latent: int,
                         lik: PlaneView | None = None, likelihood_threshold: float | None = 0.9,
                         v_quantile_threshold: float | None = 50.0, epsilon: float = 1e-6, loading_matrix=None,
                         mean=None, threshold: float = 5.0, scalar: float = 10.0, max_rounds: int = 64) -> int:
    """The while-loop of mA_compute_maha (eks/multicam_smoother.py:684-712) for all problems at once: the
    variance planes of `vv` are inflated in place.  Returns the number of rounds."""
    import numpy as np
    B, O = ymean.shape
    dev = ymean.device
    active = torch.ones(B, dtype=torch.int32, device=dev)
    fixed = loading_matrix is not None and mean is not None
    Wh = np.zeros((B, O, n_latent))
    muh = np.zeros((B, O))
    if fixed:
        Wh[:] = np.asarray(loading_matrix, dtype=np.float64)
        muh[:] = np.asarray(mean, dtype=np.float64)
    use_lik = lik is not None and likelihood_threshold is not None
    rounds = 0
    while rounds < max_rounds:
        rounds += 1
        if not fixed:
            mom = ops.mc_valid_moments(yv, ymean, vv, T, ws, v_quantile=v_quantile_threshold,
                                       lik=lik if use_lik else None,
                                       lik_threshold=likelihood_threshold if use_lik else 0.0, active=active).cpu().numpy()
            act = active.cpu().numpy()
            for b in range(B):
                if act[b]:
                    muh[b], Wh[b] = fa_from_moments(mom[b], O, n_latent)
        flags = ops.mc_inflate_step(yv, ymean, vv, T, torch.as_tensor(Wh, device=dev), torch.as_tensor(muh, device=dev),
                                    epsilon=epsilon, threshold=threshold, scalar=scalar, active=active)
        active = active * (flags != 0).to(torch.int32)
        if int(active.sum().item()) == 0:
            break
    else:   # the reference loops until no variance is inflated any more; x10 per round makes 64 rounds unreachable
        import logging
        logging.getLogger(__name__).warning(
            f'variance inflation stopped after {max_rounds} rounds with {int(active.sum().item())} keypoints still '
            f'inflating (the reference would continue)')
    return rounds


def multicam_smooth_sessions(raw: torch.Tensor, smooth_param=None, spans=None, quantile_keep_pca: float = 50.0,
                             n_latent: int = 3, avg_mode='median', var_mode='confidence_weighted_var',
                             dtype=torch.float32, lr=0.25, s_bounds_log=(-8.0, 8.0), tol=1e-2, safety_cap=300,
                             min_R_var=1e-4, out: torch.Tensor | None = None, timers: dict | None = None,
                             inflate_vars: bool = False, inflate_vars_kwargs: dict | None = None,
                             cams=None, trace_cap: int = 0, pca=None) -> MulticamResult:
    """ensemble_kalman_smoother_multicam (eks/multicam_smoother.py:279-551) for S sessions at once, every per-frame
    stage on the device.  raw: (S, M, V, T, K, 3) CUDA tensor.  cams = None: linear PCA-latent model; cams = (V, 29)
    packed camera parameters: calibrated pinhole EKF (triangulation on the device, geometric initialisation of the
    3-D state on the device: eks_geometric_init; variance inflation on the centred predictions as in the linear branch).

    The only host work is the O x O eigen-decomposition per keypoint (O = 2V <= 16) on the moments the device
    reduced -- one small device->host copy and one host->device copy of the components."""
    import numpy as np
    assert raw.is_cuda and raw.dim() == 6 and raw.shape[-1] == 3

    def stage(name):
        class _S:
            def __enter__(self_):
                if timers is not None:
                    self_.e0 = torch.cuda.Event(enable_timing=True)
                    self_.e0.record()

            def __exit__(self_, *exc):
                if timers is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    timers.setdefault(name, []).append((self_.e0, e1))
        return _S()
    raw = raw.contiguous()
    S, M, V, T, K, _ = raw.shape
    dev = raw.device
    B, O, L = S * K, 2 * V, n_latent
    if out is None:
        out = torch.empty((S, K, V, 9, T), dtype=dtype, device=dev)
    assert out.shape == (S, K, V, 9, T) and out.dtype == dtype and out.is_contiguous()
    with stage('ensemble'):
        ops.ensemble_stats(raw, out, K * V * 9 * T, 9 * T, V * 9 * T, [c * T for c in ops.ENS_TO_OUT],
                           avg_mode=avg_mode, var_mode=var_mode)
    yv = PlaneView(out, V * 9 * T, [v * 9 * T + (3 + j) * T for v in range(V) for j in range(2)])
    vv = PlaneView(out, V * 9 * T, [v * 9 * T + (5 + j) * T for v in range(V) for j in range(2)])
    if cams is not None:
        raw_var = None
        if inflate_vars:
            # eks/multicam_smoother.py:353-361 applies to both branches: Mahalanobis inflation on the CENTRED predictions;
            # the calibrated model then smooths the un-centred observations with the inflated variances but REPORTS the
            # raw ensemble variances (:474-477), so planes 5, 6 are saved and restored around the smoother
            with stage('center'):
                ymean_c, _, ws_c = ops.mc_center(yv, vv, S, K, T, quantile_keep_pca)
            raw_var = out[:, :, :, 5:7, :].clone()
            kw = dict(inflate_vars_kwargs or {})
            lik = None
            if kw.pop('likelihoods', None) is not None:
                lik = PlaneView(out, V * 9 * T, [v * 9 * T + 2 * T for v in range(V)])
            with stage('inflate_vars'):
                mc_inflate_variances(yv, vv, ymean_c, T, ws_c, kw.pop('n_latent', L), lik=lik, **kw)
        res = _multicam_pinhole(raw, out, yv, vv, cams, smooth_param, spans, dtype, lr, s_bounds_log, tol, safety_cap,
                                min_R_var, stage, trace_cap)
        if raw_var is not None:
            out[:, :, :, 5:7, :] = raw_var
        return res
    with stage('center'):
        ymean, n_good, ws = ops.mc_center(yv, vv, S, K, T, quantile_keep_pca)
    with stage('pca'):
        if pca is None:
            mom = ops.mc_pca_moments(yv, ymean, T, ws).cpu().numpy()
            pca_mean, comps = pca_from_moments(mom, O, L)
        else:   # a fitted PCA supplied by the caller (pca_object of the reference, eks/stats.py:52-56): one for all keypoints
            pm_, cp_ = np.asarray(pca[0], dtype=np.float64), np.asarray(pca[1], dtype=np.float64)
            assert pm_.shape == (O,) and cp_.shape == (L, O), 'pca = (mean_ (2V,), components_ (n_latent, 2V))'
            pca_mean, comps = np.tile(pm_, (B, 1)), np.tile(cp_, (B, 1, 1))
        C = torch.as_tensor(np.ascontiguousarray(np.swapaxes(comps, 1, 2)), device=dev).to(dtype).contiguous()
        pm = torch.as_tensor(pca_mean, device=dev).to(dtype).contiguous()
    with stage('latent_init'):
        S0, Q = ops.mc_latent_init(yv, ymean, pm, C, T, ws)
    if inflate_vars:   # in place on the ensemble-variance planes: the reference outputs the inflated variances
        kw = dict(inflate_vars_kwargs or {})
        lik = None
        if kw.pop('likelihoods', None) is not None:
            lik = PlaneView(out, V * 9 * T, [v * 9 * T + 2 * T for v in range(V)])
        with stage('inflate_vars'):
            mc_inflate_variances(yv, vv, ymean, T, ws, kw.pop('n_latent', L), lik=lik, **kw)
    eye = torch.eye(L, dtype=dtype, device=dev).expand(B, L, L).contiguous()
    model = Model(torch.zeros((B, L), dtype=dtype, device=dev), S0, eye, Q, C)
    iters = loss = None
    if smooth_param is not None:
        s = torch.as_tensor(smooth_param, dtype=torch.float64, device=dev)
        s_finals = (s.expand(K) if s.dim() == 0 or s.numel() == 1 else s).expand(S, K).contiguous()
    else:
        with stage('initial_guess'):
            _, s_log0 = ops.initial_guess(vv, B, T)
        with stage('const_R_median'):
            Rconst = ops.const_R_median(vv, B, T, spans=spans, min_var=min_R_var)
        with stage('optimize_s'):
            opt = ops.optimize_s(model, yv, T, Rconst, s_log0, ymean=ymean, spans=spans, lr=lr,
                                 s_bounds_log=s_bounds_log, tol=tol, safety_cap=safety_cap, trace_cap=trace_cap)
        multicam_smooth_sessions.last_opt = opt
        s_finals = torch.exp(opt['s_log'].double().clamp(s_bounds_log[0], s_bounds_log[1])).view(S, K)
        iters, loss = opt['iters'].view(S, K), opt['loss'].view(S, K)
    with stage('filter_smooth'):
        ms, Vs = ops.filter_smooth(model, yv, vv, T, s_finals.reshape(B).to(dtype), ymean=ymean)
    with stage('reproject'):
        ops.reproject(ms, Vs, V, out, V * 9 * T, 9 * T, [0, T, 7 * T, 8 * T], C=C, ymean=ymean, var=vv)
    return MulticamResult(out, ms, Vs, s_finals, iters, loss, ymean, C, S0, Q, n_good)


def _multicam_pinhole(raw, out, yv, vv, cams, smooth_param, spans, dtype, lr, s_bounds_log, tol, safety_cap, min_R_var,
                      stage, trace_cap=0) -> MulticamResult:
    """Calibrated branch of multicam_smooth_sessions (eks/multicam_smoother.py:380-405, 446-480): un-centred
    observations, 3-D state initialised from the triangulated ensemble mean, pinhole emission."""
    import numpy as np
    S, M, V, T, K, _ = raw.shape
    dev, B = raw.device, S * K
    cams64 = torch.as_tensor(np.asarray(cams, dtype=np.float64))
    assert cams64.shape == (V, 29), 'cams must be (n_cameras, 29) packed camera parameters'
    with stage('triangulate'):
        tri = torch.stack([ops.triangulate_mean(raw[s_], cams64) for s_ in range(S)])       # (S,K,T,3) float64
    with stage('geometric_init'):      # means / variances / MAD of the differences on the device (eks_geometric_init)
        m0d, S0d, Qd = ops.geometric_init(tri.reshape(B, T, 3))
        eye = torch.eye(3, dtype=dtype, device=dev).expand(B, 3, 3).contiguous()
        model_args = (m0d.to(dtype).contiguous(), torch.diag_embed(S0d).to(dtype).contiguous(), eye,
                      torch.diag_embed(Qd).to(dtype).contiguous())
    d_cams = cams64.to(device=dev, dtype=dtype).contiguous()
    model = Model(*model_args, None, d_cams)
    iters = loss = None
    if smooth_param is not None:
        s = torch.as_tensor(smooth_param, dtype=torch.float64, device=dev)
        s_finals = (s.expand(K) if s.dim() == 0 or s.numel() == 1 else s).expand(S, K).contiguous()
    else:
        with stage('initial_guess'):
            _, s_log0 = ops.initial_guess(vv, B, T)
        with stage('const_R_median'):
            Rconst = ops.const_R_median(vv, B, T, spans=spans, min_var=min_R_var)
        with stage('optimize_s'):
            opt = ops.optimize_s(model, yv, T, Rconst, s_log0, spans=spans, lr=lr, s_bounds_log=s_bounds_log, tol=tol,
                                 safety_cap=safety_cap, trace_cap=trace_cap)
        multicam_smooth_sessions.last_opt = opt
        s_finals = torch.exp(opt['s_log'].double().clamp(s_bounds_log[0], s_bounds_log[1])).view(S, K)
        iters, loss = opt['iters'].view(S, K), opt['loss'].view(S, K)
    with stage('filter_smooth'):
        ms, Vs = ops.filter_smooth(model, yv, vv, T, s_finals.reshape(B).to(dtype))
    with stage('reproject'):
        ops.reproject(ms, Vs, V, out, V * 9 * T, 9 * T, [0, T, 7 * T, 8 * T], cams=d_cams, var=vv,
                      pinhole_var_quirk=True)
    zeros = torch.zeros((B, 2 * V), dtype=dtype, device=dev)
    return MulticamResult(out, ms, Vs, s_finals, iters, loss, zeros, None, model.S0, model.Q, None)
