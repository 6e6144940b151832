"""Time only the s-optimisation stage on resident planes (kernel tuning helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json
from bench import synth_session_device
from eks_b200.pipeline import singlecam_smooth_sessions
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda')
raw = torch.stack([synth_session_device(torch, 10, 20, 1_000_000, s, dev, torch.float32) for s in range(S)])
out = torch.empty((S, 20, 9, 1_000_000), device=dev)
for _ in range(2):
    singlecam_smooth_sessions(raw, out=out)
timers = {}
for _ in range(3):
    res = singlecam_smooth_sessions(raw, out=out, timers=timers)
torch.cuda.synchronize()
ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in timers.items()}
it = res.iters.double()
gb = it.sum().item() * 1e6 * 2 * 4 / 1e9
print(json.dumps({'tag': os.environ.get('TAG', ''), 'ms': {k: round(v, 3) for k, v in ms.items()},
                  'opt_TBps': round(gb / ms['optimize_s'], 3), 'iters_mean': it.mean().item()}))
# distribution of the Adam iteration counts: how many sequences are still active at each evaluation launch
import numpy as np
itn = res.iters.cpu().numpy().ravel()
act = np.array([(itn > k).sum() for k in range(int(itn.max()))]) / itn.size
print(json.dumps({'iters_min': int(itn.min()), 'iters_max': int(itn.max()),
                  'quantiles_10_50_90': [int(np.percentile(itn, q)) for q in (10, 50, 90)],
                  'launches_with_active_fraction_below': {str(f): int((act < f).sum()) for f in (0.9, 0.5, 0.25, 0.1)},
                  'ideal_ms_if_time_proportional_to_active': round(float(act.sum()) / len(act) * 100, 1)}))
