"""world_size-2 (and 3) CPU test of the multi-GPU host logic over the gloo backend: the
tile partition + single all-reduce(sum) of rpt_b200.distributed.assemble reproduces the
unsharded image bit for bit.  The shard renderer here is the CPU oracle (test
infrastructure) restricted to the rank's tiles -- the CUDA shard renderer obeys the same
tile ownership, which tests/test_gpu_parity.py checks on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpt_b200 import api, scenes
from rpt_b200.distributed import assemble, gather_permutation, gather_tiles, shard_tiles, tile_owner


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, spp, mb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc

    cfg = scenes.sphere_scene()
    flat = api.FlatScene(cfg.scene)
    osc = orc.OracleScene(flat)
    r = api.Renderer(cfg.scene, cfg.camera).width(w).height(h).max_bounces(mb).seed(4)
    seen = {}

    def shard(rk, wd):
        img, st = osc.render(cfg.camera, r.params(spp, 0, rk, wd), nthreads=2)
        own = tile_owner(w, h, wd).reshape(-1)
        assert (img[own != rk] == 0).all()
        seen["segments"] = st["segments"]
        return torch.from_numpy(img.copy())

    out = assemble(shard)

    # the other assembly: every rank contributes ONLY its own tiles, tile-major (what compact_out = 1 renders into), one
    # all-gather, a fixed permutation.  The compact buffer is cut out of the oracle's shard image with rptb_tile_pixel.
    from rpt_b200 import _capi as capi

    def compact(rk, wd):
        img, _ = osc.render(cfg.camera, r.params(spp, 0, rk, wd), nthreads=2)
        n = shard_tiles(w, h, rk, wd)
        buf = np.zeros((n * 128, 3))
        for k in range(n):
            for j in range(128):
                p = capi.lib().rptb_tile_pixel(w, h, rk, wd, k, j)
                if p >= 0:
                    buf[k * 128 + j] = img[p]
        return torch.from_numpy(buf.reshape(-1))

    out2 = gather_tiles(compact, w, h)
    assert torch.equal(out2, out)
    segs = torch.tensor([seen["segments"]], dtype=torch.int64)
    dist.all_reduce(segs)
    if rank == 0:
        q.put((out.numpy(), int(segs.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_equals_full_render_gloo(orc, world):
    w, h, spp, mb = 50, 27, 3, 2  # ragged against the 16x8 tiles
    cfg = scenes.sphere_scene()
    osc = orc.OracleScene(api.FlatScene(cfg.scene))
    r = api.Renderer(cfg.scene, cfg.camera).width(w).height(h).max_bounces(mb).seed(4)
    full, st = osc.render(cfg.camera, r.params(spp))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, w, h, spp, mb, q)) for rk in range(world)]
    for p in procs:
        p.start()
    got, segs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(got, full)  # x + 0 is exact: bit-identical for any world size
    assert segs == st["segments"]


def test_assemble_without_process_group_is_identity():
    t = torch.arange(6, dtype=torch.float32)
    assert assemble(lambda rk, wd: t) is t


def test_gather_permutation_matches_tile_pixel():
    from rpt_b200 import _capi as capi

    for (w, h, n) in ((50, 27, 3), (33, 9, 2), (16, 8, 1)):
        perm = gather_permutation(w, h, n)
        per = shard_tiles(w, h, 0, n) * 128
        assert len(np.unique(perm)) == w * h
        for s_ in range(n):
            for k in range(shard_tiles(w, h, s_, n)):
                for j in range(128):
                    p = capi.lib().rptb_tile_pixel(w, h, s_, n, k, j)
                    if p >= 0:
                        assert perm[p] == s_ * per + k * 128 + j
