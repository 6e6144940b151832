import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import eks_b200
from eks_b200 import ops, core
from test_gpu_multicam import _linear_case
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
prec = sys.argv[2] if len(sys.argv) > 2 else 'float32'
eks_b200.set_precision(prec)
dtype = core.get_precision()
dev = torch.device('cuda')
args = _linear_case(4, T, seed=1)
ys, m0s, S0s, As, Cs, Qs, ev = args
model, yv, K, T, O = core._stage(ys, m0s, S0s, As, Cs, Qs, None, dev, dtype)
var = torch.as_tensor(ev, device=dev).to(dtype).permute(1, 2, 0).contiguous()
vv = ops.PlaneView(var, O * T, [o * T for o in range(O)])
guess, s_log0 = ops.initial_guess(vv, K, T)
Rc = ops.const_R_median(vv, K, T)
torch.cuda.synchronize(); t0 = time.perf_counter()
opt = ops.optimize_s(model, yv, T, Rc, s_log0)
torch.cuda.synchronize(); t1 = time.perf_counter()
ws = opt['_keep'][2].cpu().numpy()
isz = 4 if prec == 'float32' else 8
redo_off = 4 * isz + 8 + 2 * isz + 4
redo = [int(np.frombuffer(ws[b * 128 + redo_off: b * 128 + redo_off + 4].tobytes(), dtype=np.int32)[0]) for b in range(K)]
woff = (K * 128 + 255) // 256 * 256
warm = np.frombuffer(ws[woff: woff + 4 * K].tobytes(), dtype=np.int32)
print(json.dumps({'T': T, 'prec': prec, 'opt_sec': t1 - t0, 'iters': opt['iters'].cpu().tolist(), 'redo': redo, 'warm': warm.tolist()}))
s = torch.exp(opt['s_log'])
torch.cuda.synchronize(); t0 = time.perf_counter()
ms, Vs = ops.filter_smooth(model, yv, vv, T, s.to(dtype))
torch.cuda.synchronize(); t1 = time.perf_counter()
print('smooth sec', t1 - t0)
