"""Developer tool: recover the TMEM lane mapping of an M=64 tcgen05 accumulator."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from epipolar_transformers_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libepipolar_b200_timers.so")   # developer build (build.py --timers)
lib = _lib.load()
for mn in (0, 1):
    N, K = 32, 128
    g = torch.Generator(device="cuda").manual_seed(5 + mn)
    if mn:
        At = torch.randn(K, 64, device="cuda", generator=g); A = At; ref_rows = At.bfloat16().double().T
    else:
        A = torch.randn(64, K, device="cuda", generator=g); ref_rows = A.bfloat16().double()
    B = torch.randn(N, K, device="cuda", generator=g)
    ref = (ref_rows @ B.bfloat16().double().T).float()          # [64, N]
    D = torch.full((128, N), float("nan"), device="cuda")
    lib.epi_umma_m64_probe(mn, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()), N, K, None)
    torch.cuda.synchronize()
    mapping = []
    for r in range(64):
        err = (D - ref[r][None]).abs().amax(1)
        lane = int(err.argmin()); mapping.append((r, lane, float(err[lane])))
    print("mn_major", mn, "row->lane:", [m[1] for m in mapping])
    print("   max err at matched lanes %.3g" % max(m[2] for m in mapping))
