"""Run the tcgen05 self-test (descriptor / swizzle / TMEM layout) on cuda:0 and print the errors."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sake_b200 import _lib  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
err = (C.c_float * 2)()
rc = _lib.lib.sake_selftest_tcgen05(err, None)
torch.cuda.synchronize()
print("selftest rc", rc, "tf32x3 err", err[0], "bf16 err", err[1], _lib.lib.sake_last_error().decode())
sys.exit(0 if rc == 0 else 1)
