"""Developer tool: per-role cycle totals of the pipelined fused kernel (library built with
`python -m epipolar_transformers_b200.build --timers`)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from epipolar_transformers_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libepipolar_b200_timers.so")
import epipolar_transformers_b200 as epi
from epipolar_transformers_b200 import synthetic as syn
lib = _lib.load()
N, C, K = 4, 256, 64
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = H
P1, P2 = syn.pairs_from_ring(N, 4 * H)
f1 = torch.relu(torch.randn(N, C, H, W, device="cuda")); f2 = torch.relu(torch.randn(N, C, H, W, device="cuda"))
P1 = torch.from_numpy(P1.astype(np.float32)).cuda(); P2 = torch.from_numpy(P2.astype(np.float32)).cuda()
wf = torch.randn(C, C, device="cuda") * 0.05; bf = torch.randn(C, device="cuda")
buf = (ctypes.c_ulonglong * 32)()
state = epi.FusionState() if "--nocache" not in sys.argv else None
zt = (ctypes.c_longlong * 64)(); zs = (ctypes.c_ulonglong * 4)()
lib.epi_zgemm_trace_read(zt, zs, 1)
for it in range(3):
    if it == 2: lib.epi_zgemm_trace_read(zt, zs, 1)
    epi.epipolar_fusion(f1, f2, P1, P2, K=K, correct_normalize=True, variant="pipe", z_folded=(wf, bf), z_residual=True, state=state)
    torch.cuda.synchronize()
    lib.epi_pipe_timers_read(buf, 1)
v = np.array(list(buf), dtype=np.float64)
items = v[8]
print("items %d (tiles %d)" % (items, N * ((H * W + 31) // 32)))
names = {0: "W bar(top)", 1: "W wait desc", 2: "W wait S", 3: "W B1+bar", 4: "W B2a sims", 5: "W B2b softmax+scatter", 6: "W argmax+wait O(j-1)", 7: "W beta conv+bar", 9: "W epilogue",
         10: "S wait desc_free", 28: "S claim+bar", 29: "S pix/ends/zero+bar", 30: "S mark+bar", 31: "S prefix+bar", 17: "S idx fill+bar", 18: "S pad+bar", 11: "S arrive/rest", 13: "G issue/work", 14: "G wait f_empty", 15: "G wait desc", 16: "G wait q_empty",
         20: "M issue/work", 21: "M wait beta", 22: "M wait o_empty", 23: "M wait f_full (G2)", 24: "M wait desc", 25: "M wait s_empty",
         26: "M wait q_full", 27: "M wait f_full (G1)"}
for k in sorted(names):
    print("%-24s %9.0f cyc/item" % (names[k], v[k] / max(items, 1)))
for lo, hi, nm in ((0, 8, "workers(excl epi)"), (10, 12, "setup(part)"), (13, 17, "gather"), (20, 28, "mma")):
    print("%-10s total %9.0f cyc/item" % (nm, v[lo:hi].sum() / max(items, 1)))

sb = (ctypes.c_ulonglong * 16)()
lib.epi_stage_timers_read(sb)
print("stage order block (cycles): geom+zero %d, hist %d, scan %d, place %d, binsort %d" % tuple(list(sb)[:5]))

# ---- timeline of CTA 0 (clock64, relative to the first event) ----
tr = (ctypes.c_longlong * 1024)()
lib.epi_pipe_trace_read(tr)
t = np.array(list(tr), dtype=np.int64).reshape(64, 16)
ev = ["setup start", "desc ready", "G1 gathers issued", "G1 mma issued", "workers start", "B1 done", "beta ready", "iter end (epi j-1)", "G2 gathers issued", "G2 mma issued"]
t0 = t[0, 0]
print("timeline of CTA 0 (kcycles since its first setup start):")
print("item " + " ".join("%11s" % e[:11] for e in ev))
for j in range(8):
    if t[j, 0] == 0 and j > 0: break
    print("%4d " % j + " ".join("%11.1f" % ((t[j, e] - t0) / 1e3) if t[j, e] else "%11s" % "-" for e in range(10)))

# ---- persistent z GEMM: CTA 0 timeline (cycles since kernel entry) and the spread of CTA start / end times (globaltimer) ----
lib.epi_zgemm_trace_read(zt, zs, 0)
z = np.array(list(zt), dtype=np.int64); t0 = z[0]
f = lambda i: (z[i] - t0) / 1e3 if z[i] else float("nan")
print("zgemm CTA 0 (kcycles): prologue done %.1f, after pdl_wait %.1f, W landed %.1f, exit %.1f" % (f(1), f(2), f(3), f(4)))
for k in range(4):
    if z[8 + k]: print("  unit %d: MMAs issued %.1f, epilogue start %.1f, epilogue end %.1f" % (k, f(8 + k), f(16 + k), f(24 + k)))
sp = list(zs)
print("zgemm CTA starts spread %.2f us, first start -> last end %.2f us, ends spread %.2f us" % ((sp[1] - sp[0]) / 1e3, (sp[3] - sp[0]) / 1e3, (sp[3] - sp[2]) / 1e3))

# ---- fused kernel: per-CTA start / end (globaltimer) and item counts: where the tail comes from ----
cb = (ctypes.c_ulonglong * 1024)()
lib.epi_pipe_cta_read(cb)
c = np.array(list(cb), dtype=np.float64).reshape(256, 4)[:148]
t0 = c[:, 0].min()
ent, dep, end, items = (c[:, 0] - t0) / 1e3, (c[:, 1] - t0) / 1e3, (c[:, 2] - t0) / 1e3, c[:, 3]
print("fused kernel CTAs: entry spread %.2f us, dependency wait ends %.2f..%.2f us, exits %.2f..%.2f us (median %.2f)" % (ent.max(), dep.min(), dep.max(), end.min(), end.max(), np.median(end)))
for n in sorted(set(items.astype(int))):
    m = items == n
    print("  %d CTAs with %d items: exit %.2f..%.2f us (mean %.2f)" % (m.sum(), n, end[m].min(), end[m].max(), end[m].mean()))
