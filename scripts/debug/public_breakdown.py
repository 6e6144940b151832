import sys, os, time
import numpy as np, torch, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from eks_b200 import _xfer, ops
from eks_b200.marker_array import MarkerArray
from eks_b200.pipeline import singlecam_smooth_sessions
from eks_b200.utils import make_dlc_pandas_index
M, K, T = 10, 20, 1_000_000
host = bench.synth_session_host(M, K, T, 0)
dev = torch.device('cuda')
print('threads', torch.get_num_threads(), 'cores', os.cpu_count())
def tic(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = tic(); ma = MarkerArray(host, data_fields=['x', 'y', 'likelihood']); t1 = tic()
    raw = _xfer.to_device(ma.array, dev); t2 = tic()
    res = singlecam_smooth_sessions(raw.reshape(1, M, 1, T, K, 3)); t3 = tic()
    fd = torch.empty((T, K, 9), dtype=torch.float64, device=dev); fd.copy_(res.out[0].permute(2, 0, 1)); t4 = tic()
    final = _xfer.to_host(fd.view(T, K * 9)); t5 = tic()
    df = pd.DataFrame(final, columns=make_dlc_pandas_index([f'k{i}' for i in range(K)], labels=ops.OUT_COLS)); t6 = tic()
    s = res.s_finals[0].cpu().numpy(); t7 = tic()
    print(f'rep {rep}: markerarray {t1-t0:.3f} to_device {t2-t1:.3f} ({host.nbytes/1e9/(t2-t1):.1f} GB/s) pipeline {t3-t2:.3f} transpose {t4-t3:.3f} to_host {t5-t4:.3f} ({final.nbytes/1e9/(t5-t4):.1f} GB/s) dataframe {t6-t5:.3f} s {t7-t6:.3f}')
    # plain torch paths for comparison
    t0 = tic(); r2 = torch.as_tensor(host).to(dev); t1 = tic(); f2 = fd.cpu().numpy(); t2 = tic()
    print(f'   plain: .to(dev) {t1-t0:.3f}  .cpu() {t2-t1:.3f}')
    del raw, res, fd, final, df, r2, f2
