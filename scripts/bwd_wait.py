"""Where does the backward mix kernel wait?  (python scripts/bwd_wait.py cfg3 tf32x3)"""
import ctypes as C, os, sys
os.environ["SAKE_DEBUG_WSPLITS"] = "7"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import sake_b200
from sake_b200 import runner as R, _lib
B, N, S, padded, n_min, mode, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
eng = sys.argv[2] if len(sys.argv) > 2 else "f16x2"
model = sake_b200.DenseSAKEModel(hidden_features=64, out_features=1, depth=4, engine=eng)
run = R.ModelRunner(model, bench.init_params_cpu(4, S, 0), B, N, S, ragged=padded and N <= 128, train=False)
h, x, mask, am, y, n_real = bench.synth(2666, B, N, S, padded, n_min)
T = lambda a: None if a is None else torch.tensor(a, device="cuda")
run.load_inputs(T(h), T(x), target=T(y), n_real=(torch.tensor(n_real, device="cuda", dtype=torch.int32) if run.ragged else None))
for _ in range(2):
    run.energy_forces_step()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
_lib.lib.sake_debug_counters_bwd(buf)
run.energy_forces_step()
torch.cuda.synchronize()
_lib.lib.sake_debug_counters_bwd(buf)
b = [float(v) for v in buf]
tiles = b[6]
ctas = 148 * 4.0       # (CTAs x launches) the single-thread epilogue/builder counters were summed over
per = lambda v: round(v / tiles)
print(eng, "tiles", int(tiles), "| MMA issuer cycles/tile: total", per(b[5]), "wait d2_empty", per(b[0]), "G1 weights", per(b[1]),
      "G1 pair chunks", per(b[2]), "G2 weights", per(b[3]), "G2 dZ chunks", per(b[4]),
      "issue", per(b[5] - sum(b[0:5])))
print("   epilogue thread cycles/tile: wait d1_full", per(b[7]), "E1 total", per(b[11]), "(of which ring-slot wait", per(b[8]), ")",
      "bar1", per(b[9]), "wait d2_full", per(b[10]), "E2", per(b[12]))
print("   builder thread cycles/tile: ring waits", per(b[13]), "load+build", per(b[14]))
