import sys, torch, time
sys.path.insert(0, '/root/repo')
import bench
from eks_b200.pipeline import singlecam_smooth_sessions
dev = torch.device('cuda:0')
S, M, K, T = 8, 10, 20, 1_000_000
raw = torch.empty((S, M, 1, T, K, 3), device=dev, dtype=torch.float32)
for s in range(S):
    raw[s] = bench.synth_session_device(torch, M, K, T, seed=s, device=dev, dtype=torch.float32)
out = torch.empty((S, K, 9, T), device=dev, dtype=torch.float32)
ref = None
for G in (1, 2, 4, 8, 4, 1):
    for _ in range(3):
        r = singlecam_smooth_sessions(raw, out=out, n_groups=G)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = singlecam_smooth_sessions(raw, out=out, n_groups=G)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = (r.s_finals.clone(), r.iters.clone(), out.clone())
    same = bool((r.s_finals == ref[0]).all() and (r.iters == ref[1]).all() and torch.equal(out, ref[2]))
    print(f'groups {G}: {e0.elapsed_time(e1) / 5:.3f} ms/step  identical to groups=1: {same}', flush=True)
