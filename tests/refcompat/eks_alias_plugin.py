"""pytest plugin (test infrastructure): makes `import eks` resolve to eks_b200 so that the REFERENCE's own unit-test
files (read in place from /root/reference/tests, never copied) run unmodified against this package.  The reference
tests build their inputs with jax.numpy; jax is not installable here, so a minimal stand-in (`jax.numpy` = numpy,
`jax.jit` = identity, `jax.config.update` = no-op) is registered for the TEST process only -- the product never
imports it."""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as _np  # noqa: E402

if 'jax' not in sys.modules:
    try:
        import jax  # noqa: F401
    except Exception:
        jax = types.ModuleType('jax')
        jnp = types.ModuleType('jax.numpy')
        for _n in dir(_np):
            if not _n.startswith('__'):
                setattr(jnp, _n, getattr(_np, _n))
        jax.numpy = jnp
        jax.jit = lambda f, *a, **k: f
        jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
        sys.modules['jax'] = jax
        sys.modules['jax.numpy'] = jnp

import eks_b200  # noqa: E402

sys.modules['eks'] = eks_b200
for _sub in ('marker_array', 'utils', 'stats', 'core', 'ibl_pupil_smoother', 'singlecam_smoother',
             'multicam_smoother', 'ibl_paw_multicam_smoother'):
    sys.modules['eks.' + _sub] = importlib.import_module('eks_b200.' + _sub)
