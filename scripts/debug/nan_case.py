import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import synth_singlecam
from eks_b200.pipeline import singlecam_smooth_sessions
raw = synth_singlecam(M=3, K=2, T=20_000, seed=5)
raw[:, 0, 7777, 1, 0:2] = np.nan
for exact in (False, True):
    res = singlecam_smooth_sessions(torch.as_tensor(raw).cuda()[None], dtype=torch.float64, smooth_param=[0.1, 0.1], exact_scan=exact)
    torch.cuda.synchronize()
    o = res.out[0]
    print('exact', exact, 'means', res.means.cpu().numpy().tolist())
    print('  nan counts per plane kp1:', [int(torch.isnan(o[1, c]).sum()) for c in range(9)], 'kp0:', [int(torch.isnan(o[0, c]).sum()) for c in range(9)])
    print('  x_med[7775:7780] kp1', o[1, 3, 7775:7780].cpu().numpy())
