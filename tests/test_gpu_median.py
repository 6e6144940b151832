"""GPU tests of eks_const_R_median (constant_R_from_timevarying, eks/core.py:702-709) on long sequences, where the
one-pass bracketed select is used: the result must be the exact nanmedian (bit for bit), with the three-pass radix
select taking over whenever the bracket misses or a candidate buffer overflows."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _median(planes, dtype, spans=None, min_var=1e-4):
    """planes (B, O, T) numpy -> (B, O) device result"""
    from eks_b200 import ops
    B, O, T = planes.shape
    d = torch.as_tensor(planes).to('cuda', dtype).contiguous()
    vv = ops.PlaneView(d, O * T, [o * T for o in range(O)])
    out = ops.const_R_median(vv, B, T, spans=spans, min_var=min_var)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _ref(planes, np_dtype, spans=None, min_var=1e-4):
    x = planes.astype(np_dtype)
    if spans:
        x = np.concatenate([x[..., a:b] for a, b in spans], axis=-1)
    with np.errstate(all='ignore'):
        med = np.nanmedian(x, axis=-1)
    return np.maximum(np.maximum(med, np_dtype(1e-12)), np_dtype(min_var)).astype(np_dtype)


@pytest.mark.parametrize('dtype,np_dtype', [(torch.float32, np.float32), (torch.float64, np.float64)])
@pytest.mark.parametrize('T', [131072, 300_001, 1_000_000])
def test_bracketed_median_is_exact(dtype, np_dtype, T):
    rng = np.random.default_rng(T)
    B, O = 3, 2
    planes = np.exp(rng.normal(-1.5, 0.8, size=(B, O, T))).astype(np_dtype)
    planes[0, 0, rng.random(T) < 0.02] *= 100.0                 # occlusion-like outliers
    planes[1, 1, rng.random(T) < 0.3] = np.nan                  # many NaNs (odd / even valid counts arise)
    planes[2, 0, ::7] = np.nan
    got = _median(planes, dtype)
    ref = _ref(planes, np_dtype)
    np.testing.assert_array_equal(got, ref)


def test_bracketed_median_with_spans_and_duplicates():
    rng = np.random.default_rng(1)
    T = 400_000
    planes = np.empty((4, 2, T), dtype=np.float32)
    planes[0] = rng.choice(np.array([0.1, 0.2, 0.3], dtype=np.float32), size=(2, T))     # three distinct values
    planes[1] = 0.25                                                                      # constant: everything in the bracket
    planes[2] = np.round(np.exp(rng.normal(-1, 0.5, size=(2, T))), 2)                     # heavy ties around the median
    planes[3] = np.sort(np.exp(rng.normal(-1, 0.5, size=(2, T))), axis=-1)                # sorted in time (sampling still fine)
    for spans in (None, [(1000, 399_000)], [(10, 150_000), (200_000, 399_999)]):
        got = _median(planes, torch.float32, spans=spans)
        np.testing.assert_array_equal(got, _ref(planes, np.float32, spans=spans))


def test_median_all_nan_and_floor():
    T = 200_000
    planes = np.full((2, 2, T), np.nan, dtype=np.float32)
    planes[1] = 1e-7                                            # below the min_var floor
    got = _median(planes, torch.float32)
    assert np.isnan(got[0]).all()
    np.testing.assert_array_equal(got[1], np.float32(1e-4))


def test_many_problems_exceed_grid_y_limit():
    """the batch dimension must not sit on gridDim.y (65535 limit): 70 000 short problems."""
    rng = np.random.default_rng(2)
    planes = np.exp(rng.normal(-1, 0.5, size=(35_000, 2, 64))).astype(np.float32)
    got = _median(planes, torch.float32)
    np.testing.assert_array_equal(got, _ref(planes, np.float32))
