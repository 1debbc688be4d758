"""Time the XtG kernel for the shapes one layer backward uses (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sake_b200 import _lib
torch.manual_seed(0)
shapes = [(215296, 256, 256), (215296, 64, 64), (215296, 64, 192), (215296, 64, 16),
          (7424, 64, 64), (7424, 256, 64), (7424, 128, 64), (7424, 64, 128), (7424, 64, 16)]
for eng in (2, 3):
    for (P, xw, gw) in shapes:
        X = torch.randn(P, xw, device="cuda"); G = torch.randn(P, gw, device="cuda")
        out = torch.zeros(xw, gw, device="cuda")
        for _ in range(3):
            _lib.lib.sake_selftest_xtg(eng, P, xw, gw, X.data_ptr(), G.data_ptr(), out.data_ptr(), None)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            _lib.lib.sake_selftest_xtg(eng, P, xw, gw, X.data_ptr(), G.data_ptr(), out.data_ptr(), None)
        b.record(); torch.cuda.synchronize()
        print("engine", eng, (P, xw, gw), "ms per call", a.elapsed_time(b) / 10)
