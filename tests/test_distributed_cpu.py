"""CPU, world_size=2, gloo: the host logic of the one-view-per-rank mode (pairing table, all-gather of the
per-view maps, which source map each rank fuses against).  The CUDA op itself is covered by the gpu tests; here
the fusion callable is a stand-in so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epipolar_transformers_b200 import synthetic as syn
from epipolar_transformers_b200.distributed import ViewParallelFusion, source_view_table


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        KRT = syn.ring_cameras(world, 64)
        feat = torch.full((3, 4, 8, 8), float(rank + 1))
        seen = {}

        def fake_fuse(f_ref, f_src, P_ref, P_src):
            seen["src_val"] = float(f_src.mean())
            seen["P_ref"] = P_ref[0].numpy().copy(); seen["P_src"] = P_src[0].numpy().copy()
            return f_ref + f_src

        vp = ViewParallelFusion(KRT, fuse_fn=fake_fuse, exchange="allgather")
        y = vp(feat)
        vp2 = ViewParallelFusion(KRT, fuse_fn=fake_fuse, exchange="p2p")
        y2 = vp2(feat)
        assert torch.equal(y, y2)                                       # both exchange modes deliver the same source map
        src = int(source_view_table(KRT)[rank])
        assert vp.src == src and src != rank
        assert seen["src_val"] == float(src + 1)                       # fused against the gathered map of view src(v)
        np.testing.assert_allclose(seen["P_ref"], KRT[rank].astype(np.float32))
        np.testing.assert_allclose(seen["P_src"], KRT[src].astype(np.float32))
        assert torch.equal(y, torch.full_like(feat, float(rank + 1 + src + 1)))
        g = vp.gather(feat)
        assert g.shape == (world, 3, 4, 8, 8) and [float(g[v].mean()) for v in range(world)] == [1.0, 2.0]
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_view_parallel_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_source_view_table_ring():
    KRT = syn.ring_cameras(8, 256)
    src = source_view_table(KRT)
    assert all(abs(int(s) - v) in (1, 7) for v, s in enumerate(src))      # ring neighbours
