"""Developer tool: cycles per tcgen05.mma (M=128, K=16, bf16, operands in shared memory) vs N."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epipolar_transformers_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libepipolar_b200_timers.so")   # developer build (build.py --timers)
lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for M, mn in ((128, 0), (128, 1), (64, 0), (64, 1)):
    for N in (32, 64, 128, 256):
        for reps in (256,):
            lib.epi_umma_bench(M, N, reps, mn, ctypes.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
            tot, issue = out[0].item(), out[1].item()
            print("M=%3d A %s N=%3d reps=%3d: total %7d cyc (%.1f / mma), issue loop %6d cyc (%.1f / mma)" %
                  (M, "MN" if mn else "K ", N, reps, tot, tot / reps, issue, issue / reps))
