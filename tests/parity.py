"""Parity protocol shared by the GPU tests.

fp64 mode is the primary gate: identical Adam iteration counts, s and every output column within 1e-5 of the fp64
oracle.

fp32 mode (the production precision) cannot be held to "identical iteration count" by ANY float32 implementation:
the stop rule |loss_i - loss_{i-1}| < tol |log loss_{i-1}| + 1e-6 (eks/core.py:657-681) has a threshold of ~0.15 while
a float32 loss of ~4e6 (10^6 frames) has a resolution of 0.25-0.5, and the reference's own float32 arithmetic is
hundreds of units away from the float64 value.  `fp32_stop_protocol` therefore checks what CAN be checked and reports
the rest:
  1. loss values: at every common iterate the product's float32 loss equals the float64 oracle's loss -- after removing
     the first-order effect of the (tiny) difference in log s between the two trajectories, 1/2 (g_gpu + g_ref) ds_log --
     within `kappa` float32 ulps of the loss;
  2. the optimisation trajectories coincide: max |s_log_gpu[i] - s_log_oracle[i]| over the common iterations <= 1e-3;
  3. if the iteration counts differ, the dispute is a rounding knife edge: the float64 oracle's own stop margin
     | |loss_i - loss_{i-1}| - (tol |log loss_{i-1}| + 1e-6) | at the disputed iteration is below
     eps = `eps_ulps` float32 ulps of the loss (the stop statistic of a float32 run is quantised to 1 ulp); the margin
     is printed in absolute terms and in ulps;
  4. (caller) outputs within 1e-3 of the float64 oracle evaluated at the product's own s; |ds|/s is reported.
"""
import numpy as np


def ulp32(x):
    return float(np.spacing(np.float32(abs(x))))


def stop_threshold(prev, tol):
    return tol * abs(np.log(max(prev, 1e-12))) + 1e-6


def loss_deviation_ulp32(trace, n_trace, ref_trace, n_ref, lr=0.25):
    """max over the common iterates of |loss - loss_ref - first-order effect of the log-s difference| in float32 ulps of
    the reference loss; also the absolute deviations, the ulps and the log-s differences per iterate."""
    g = np.asarray(trace, dtype=np.float64)
    r = np.asarray(ref_trace, dtype=np.float64)
    n = int(min(n_trace, n_ref, g.shape[0], r.shape[0]))
    ds = g[:n, 0] - r[:n, 0]
    first_order = 0.5 * (g[:n, 2] + r[:n, 2]) / lr * ds
    d_loss = np.abs(g[:n, 1] - r[:n, 1] - first_order)
    u = np.array([ulp32(v) for v in r[:n, 1]])
    return (float((d_loss / u).max()) if n else 0.0), d_loss, u, ds


def fp32_stop_protocol(label, gpu_trace, n_gpu, ref_trace, n_ref, tol=1e-2, lr=0.25, kappa=16.0, eps_ulps=4.0,
                       traj_atol=1e-3, verbose=True, ref32_trace=None, n_ref32=None):
    """gpu_trace / ref_trace: (cap, 3) rows [s_log before the update, loss, lr * d loss / d s_log] per iteration (device
    trace of ops.optimize_s / oracle trace).  n_gpu / n_ref: iteration counts.  Returns the reported quantities.

    ref32_trace / n_ref32 (optional): the trace of the float32 ORACLE (the reference's own float32 arithmetic restated)
    on the same data.  Its own deviation from the float64 oracle then replaces `kappa` when it is larger: the product's
    float32 loss must be no further from the float64 value than max(kappa ulps, what the reference's float32 arithmetic
    manages on this data set) -- on short, ill-conditioned real data (mirror-mouse-separate, 501 frames, +-100 px
    coordinates through float32 PCA components) that is 25-135 ulps."""
    g = np.asarray(gpu_trace, dtype=np.float64)
    r = np.asarray(ref_trace, dtype=np.float64)
    n = int(min(n_gpu, n_ref, g.shape[0], r.shape[0]))
    assert n >= 2, f'{label}: traces too short'
    worst_ulps, d_loss, u, ds = loss_deviation_ulp32(g, n_gpu, r, n_ref, lr)
    worst = int(np.argmax(d_loss / u))
    rep = dict(n_gpu=int(n_gpu), n_ref=int(n_ref), max_loss_err=float(d_loss.max()),
               max_loss_err_ulp32=worst_ulps, traj_err=float(np.abs(ds).max()))
    if ref32_trace is not None:
        dev32 = loss_deviation_ulp32(ref32_trace, n_ref32, r, n_ref, lr)[0]
        rep['oracle_f32_loss_err_ulp32'] = dev32
        kappa = max(kappa, dev32)
    assert rep['max_loss_err_ulp32'] <= kappa, (
        f'{label}: fp32 loss off by {d_loss[worst]:.4g} = {rep["max_loss_err_ulp32"]:.1f} float32 ulps at iteration '
        f'{worst} (bound {kappa})')
    assert rep['traj_err'] <= traj_atol, f'{label}: log-s trajectories diverge by {rep["traj_err"]:.3e}'
    if n_gpu != n_ref and n < r.shape[0] and n < g.shape[0]:
        # iteration index n-1 is where the earlier run stopped; the later run's margin there was >= 0
        i = n - 1
        margin_ref = abs(r[i, 1] - r[i - 1, 1]) - stop_threshold(r[i - 1, 1], tol)
        eps = eps_ulps * ulp32(r[i, 1])
        rep.update(disputed_iteration=i + 1, oracle_stop_margin=float(margin_ref),
                   oracle_stop_margin_ulp32=float(margin_ref / ulp32(r[i, 1])), eps=float(eps))
        assert abs(margin_ref) <= eps, (
            f'{label}: iteration counts {n_gpu} vs {n_ref} but the fp64 oracle stop margin at iteration {i + 1} is '
            f'{margin_ref:.4g} > eps {eps:.4g} ({eps_ulps} float32 ulps): not a rounding knife edge')
    if verbose:
        print(f'[parity fp32] {label}: ' + ', '.join(f'{k}={v:.4g}' if isinstance(v, float) else f'{k}={v}'
                                                      for k, v in rep.items()))
    return rep


def check_columns(out, ref, rtol, label, var_cols=(5, 6, 7, 8), abs_floor=1.0, var_floor=1e-6):
    """Column-wise relative error of (..., C) arrays: coordinates relative to max(|ref|, abs_floor) (pixels), variance
    columns relative to max(|ref|, var_floor)."""
    for c in range(out.shape[-1]):
        a, b = out[..., c], ref[..., c]
        scale = np.maximum(np.abs(b), var_floor if c in var_cols else abs_floor)
        err = float(np.max(np.abs(a - b) / scale))
        assert err <= rtol, f'{label}: column {c} rel err {err:.3e} > {rtol}'
