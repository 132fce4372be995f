"""TEST INFRASTRUCTURE — generator of tests/golden/mpjpe_proxy.json.  Build container only (needs /root/reference):
    python -m oracle.make_golden_mpjpe
Runs the UNMODIFIED reference Epipolar (oracle/ref_harness.py) on the synthetic MPJPE-proxy inputs of
oracle/mpjpe_proxy.py and freezes the proxy's MPJPE with the reference's fused feature and with no fusion at all."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mpjpe_proxy as mp      # noqa: E402
from oracle import ref_harness as rh      # noqa: E402


def main(seeds=(0, 1, 2)):
    import torch
    out = {"meta": "MPJPE proxy (oracle/mpjpe_proxy.py); reference = unmodified /root/reference Epipolar on CPU, torch %s" % torch.__version__.split("+")[0][:4],
           "seeds": {}}
    for s in seeds:
        d = mp.build(s)
        r = rh.run_reference(d["cfg"], d["feat_ref"], d["feat_src"], d["P_ref"], d["P_src"])
        out["seeds"][str(s)] = {"mpjpe_no_fusion_mm": mp.mpjpe(d, np.zeros_like(d["feat_ref"])),
                                "mpjpe_reference_mm": mp.mpjpe(d, r["out"])}
        print(s, out["seeds"][str(s)])
    path = os.path.join(ROOT, "tests", "golden", "mpjpe_proxy.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
