"""Developer tool: cycles per tcgen05.mma (M=128, K=16, bf16, operands in shared memory) vs N."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epipolar_transformers_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for mn in (0, 1):
    for N in (16, 32, 64, 128, 256):
        for reps in (64, 256):
            lib.epi_umma_bench(N, reps, mn, ctypes.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
            tot, issue = out[0].item(), out[1].item()
            print("A %s N=%3d reps=%3d: total %7d cyc (%.1f / mma), issue loop %6d cyc (%.1f / mma)" %
                  ("MN" if mn else "K ", N, reps, tot, tot / reps, issue, issue / reps))
