"""XtG kernel diagnostic: out = X^T G on the tcgen05 engine vs torch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sake_b200 import _lib
torch.manual_seed(0)
for eng in (2, 3):
    for (P, xw, gw) in [(64, 128, 64), (300, 256, 256), (1000, 64, 16), (5000, 256, 64), (777, 100, 128)]:
        X = torch.randn(P, xw, device="cuda"); G = torch.randn(P, gw, device="cuda")
        out = torch.zeros(xw, gw, device="cuda")
        rc = _lib.lib.sake_selftest_xtg(eng, P, xw, gw, X.data_ptr(), G.data_ptr(), out.data_ptr(), None)
        torch.cuda.synchronize()
        ref = (X.double().T @ G.double())
        err = (out.double() - ref).abs().max().item()
        print("engine", eng, (P, xw, gw), "rc", rc, "max err", err, "ref max", ref.abs().max().item(), "out absmax", out.abs().max().item())
        if err > 1e-2 * ref.abs().max().item() and eng == 2:
            # diagnose: is it a transposition / permutation?
            print(" out[0,:8]", out[0, :8].tolist()); print(" ref[0,:8]", ref[0, :8].float().tolist())
            print(" ref.T[0,:8]", ref.T[0, :8].float().tolist() if xw == gw else None)
