import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, f'{name}.npz'), allow_pickle=False))
    if 'raw_from' in d:
        d['raw'] = np.load(os.path.join(GOLDEN, f'{str(d["raw_from"])}.npz'))['raw']
    return d


@pytest.fixture(scope='session')
def golden():
    return load_golden


def synth_singlecam(M=5, K=3, T=1500, seed=0, nan_frac=0.0, dtype=np.float64):
    """Seeded synthetic ensemble of the shape family in SURVEY 8(d): random-walk latent, seeds =
    truth + heteroscedastic noise (2% occlusion frames with 10x sigma and low likelihood)."""
    rng = np.random.default_rng(seed)
    truth = np.cumsum(rng.normal(0, 0.3, size=(T, K, 2)), axis=0) + rng.uniform(50, 300, size=(1, K, 2))
    occ = rng.random((T, K)) < 0.02
    sigma = np.where(occ, 5.0, 0.5)[None, :, :, None]
    pred = truth[None] + rng.normal(size=(M, T, K, 2)) * sigma
    lik = np.where(occ[None], rng.uniform(0.05, 0.5, size=(M, T, K)), rng.uniform(0.9, 1.0, size=(M, T, K)))
    raw = np.concatenate([pred, lik[..., None]], axis=-1)[:, None]  # (M,1,T,K,3)
    if nan_frac > 0:
        mask = rng.random(raw.shape[:-1]) < nan_frac
        raw[..., 0][mask] = np.nan
        raw[..., 1][mask] = np.nan
    return np.ascontiguousarray(raw.astype(np.float32).astype(dtype))
