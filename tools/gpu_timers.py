"""Developer tool: per-phase cycle totals of the tile kernel (library built with -DEPI_TILE_TIMERS)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epipolar_transformers_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libepipolar_b200_timers.so")
import epipolar_transformers_b200 as epi
from epipolar_transformers_b200 import synthetic as syn
lib = _lib.load()
N, C, H, W, K = 4, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 0, 64
VARIANT = sys.argv[2] if len(sys.argv) > 2 else "auto"
W = H
P1, P2 = syn.pairs_from_ring(N, 4 * H)
f1 = torch.relu(torch.randn(N, C, H, W, device="cuda")); f2 = torch.relu(torch.randn(N, C, H, W, device="cuda"))
P1 = torch.from_numpy(P1.astype(np.float32)).cuda(); P2 = torch.from_numpy(P2.astype(np.float32)).cuda()
buf = (ctypes.c_ulonglong * 16)()
for it in range(3):
    epi.epipolar_fusion(f1, f2, P1, P2, K=K, correct_normalize=True, variant=VARIANT)
    torch.cuda.synchronize()
    lib.epi_tile_timers_read(buf, 1)
v = np.array(list(buf), dtype=np.float64)
names = ["mark+prefix", "idx+Qstage", "phaseA(gather+mma)", "B1", "B2", "attnflush", "phaseC", "phaseD"]
groups = v[8]
tiles = N * ((H + 3) // 4) * ((W + 7) // 8)
print("variant", VARIANT)
print("groups %d tiles %d (%.2f groups/tile)" % (groups, tiles, groups / tiles))
tot = v[:8].sum()
for nm, x in zip(names, v[:8]):
    print("%-20s %8.0f cyc/group  %5.1f%%" % (nm, x / groups, 100 * x / tot))
print("total cyc/group %.0f  cyc/tile %.0f" % (tot / groups, tot / tiles))
for nm, i in (("A: acquire wait", 9), ("A: gather_store (waits loads)", 10), ("A: gather_load issue", 11), ("A: fence+sync", 12), ("A: mma issue", 13), ("A: loop overhead", 14)):
    print("%-32s %8.0f cyc/group" % (nm, v[i] / groups))
