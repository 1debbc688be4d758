"""Where does the forward mix kernel's MMA issuer wait?  (python scripts/mma_wait.py cfg3)"""
import ctypes as C, os, sys
os.environ["SAKE_DEBUG_WSPLITS"] = "7"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import sake_b200
from sake_b200 import runner as R, _lib
B, N, S, padded, n_min, mode, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
model = sake_b200.DenseSAKEModel(hidden_features=64, out_features=1, depth=4, engine=(sys.argv[2] if len(sys.argv) > 2 else "f16x2"))
run = R.ModelRunner(model, bench.init_params_cpu(4, S, 0), B, N, S, ragged=padded and N <= 128, train=False)
h, x, mask, am, y, n_real = bench.synth(2666, B, N, S, padded, n_min)
T = lambda a: None if a is None else torch.tensor(a, device="cuda")
run.load_inputs(T(h), T(x), target=T(y), n_real=(torch.tensor(n_real, device="cuda", dtype=torch.int32) if run.ragged else None))
for _ in range(2):
    run.forward()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 8)()
_lib.lib.sake_debug_counters(buf)
run.forward()
torch.cuda.synchronize()
_lib.lib.sake_debug_counters(buf)
tiles = buf[4]
print("tiles", tiles, "cycles/tile total", buf[3] / tiles, "wait acc", buf[0] / tiles, "wait weights", buf[1] / tiles,
      "wait pair-chunks", buf[2] / tiles, "issue+other", (buf[3] - buf[0] - buf[1] - buf[2]) / tiles)
