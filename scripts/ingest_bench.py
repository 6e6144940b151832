"""Host I/O around the hot path at scale: seed-CSV ingest -> MarkerArray and smoothed-CSV writing.
CPU only (no device needed).  Usage: python scripts/ingest_bench.py [rows] [keypoints] [seeds]"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eks_b200 import io as eio  # noqa: E402
from eks_b200.marker_array import input_dfs_to_markerArray  # noqa: E402
from eks_b200.utils import make_dlc_pandas_index  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
M = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kps = [f'kp{k}' for k in range(K)]
rng = np.random.default_rng(0)
out = {'rows': T, 'keypoints': K, 'seeds': M, 'cores': os.cpu_count()}
with tempfile.TemporaryDirectory() as d:
    cols = pd.MultiIndex.from_product([['net'], kps, ['x', 'y', 'likelihood']], names=['scorer', 'bodyparts', 'coords'])
    for m in range(M):
        arr = rng.normal(100, 30, size=(T, K * 3)).astype(np.float32).astype(np.float64)
        eio.write_dlc_csv(pd.DataFrame(arr, columns=cols), os.path.join(d, f'seed{m}.csv'))
    out['input_MB'] = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) / 1e6
    t0 = time.perf_counter()
    dfs, names = eio.format_data(d)
    out['format_data_pyarrow_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ma = input_dfs_to_markerArray([dfs], names, [''])
    out['input_dfs_to_markerArray_s'] = time.perf_counter() - t0
    os.environ['EKS_B200_PANDAS_CSV'] = '1'
    t0 = time.perf_counter()
    dfs2, _ = eio.format_data(d)
    out['format_data_pandas_s'] = time.perf_counter() - t0
    del os.environ['EKS_B200_PANDAS_CSV']
    # pyarrow parses correctly rounded; pandas' default C parser may be off by one ulp on 17-digit cells
    out['max_rel_diff_pyarrow_vs_pandas_parser'] = float(max(np.max(np.abs(a.to_numpy() - b.to_numpy()) / np.maximum(np.abs(b.to_numpy()), 1.0)) for a, b in zip(dfs, dfs2)))
    assert out['max_rel_diff_pyarrow_vs_pandas_parser'] < 1e-12
    labels = ['x', 'y', 'likelihood', 'x_ens_median', 'y_ens_median', 'x_ens_var', 'y_ens_var', 'x_posterior_var',
              'y_posterior_var']
    res = pd.DataFrame(rng.normal(100, 30, size=(T, K * 9)), columns=make_dlc_pandas_index(kps, labels=labels))
    t0 = time.perf_counter()
    eio.write_dlc_csv(res, os.path.join(d, 'out_fast.csv'))
    out['write_fast_s'] = time.perf_counter() - t0
    out['output_MB'] = os.path.getsize(os.path.join(d, 'out_fast.csv')) / 1e6
    if T <= 200_000:
        t0 = time.perf_counter()
        res.to_csv(os.path.join(d, 'out_pandas.csv'))
        out['write_pandas_s'] = time.perf_counter() - t0
print(json.dumps(out))
