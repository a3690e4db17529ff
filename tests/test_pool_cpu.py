"""CPU: host-side logic of the sharded acquisition round -- shard arithmetic, region budget, module surface,
drop-in installation, and the world_size-2 all-gather over gloo."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

import halo_b200
from halo_b200 import pool
from oracle import acquire as oacquire
from oracle import delta as odelta


def test_shard_range_partitions_the_pool():
    for n in (0, 1, 7, 8, 2975):
        for world in (1, 2, 4, 8):
            spans = [pool.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert pool.shard_range(2975, 7, 8) == (2604, 2975)


def test_region_budget_matches_reference_formula():
    for h, w, budget, rounds, rk in ((640, 1280, 0.05, 1, 1), (640, 1280, 0.05, 5, 1), (320, 640, 0.022, 1, 1),
                                     (640, 1280, 0.022, 1, 2), (640, 1280, 0.022, 1, 0), (1024, 2048, 0.05, 5, 1)):
        cfg = pool.AcquisitionConfig(radius_k=rk, budget=budget, n_rounds=rounds)
        assert cfg.regions_per_image(h, w) == oacquire.region_budget(h, w, budget, rounds, rk)
    assert pool.AcquisitionConfig(budget=0.05).regions_per_image(640, 1280) == 4552
    assert pool.AcquisitionConfig(budget=0.05, n_rounds=5).regions_per_image(640, 1280) == 911


def test_module_surface_matches_reference():
    m = halo_b200.HyperMLR(64, 19, c=1.0)
    assert set(m.state_dict()) == {"P_MLR", "A_MLR"}
    assert tuple(m.P_MLR.shape) == (19, 64) and m.num_classes == 19 and float(m.K) == 1.0
    # an fp64 checkpoint written by the reference loads (dtype-converting copy)
    sd = {"P_MLR": torch.randn(19, 64, dtype=torch.float64), "A_MLR": torch.randn(19, 64, dtype=torch.float64)}
    m.load_state_dict(sd)
    assert torch.allclose(m.P_MLR.double(), sd["P_MLR"], atol=1e-6)
    mapper = halo_b200.HyperMapper(c=0.5)
    assert mapper.c == 0.5 and float(mapper.K) == -0.5
    frs = halo_b200.FloatingRegionScore(in_channels=19, size=3, purity_type="hyper", K=100, curvature=1.0)
    assert frs.purity_size == 3 and frs.K == 100
    frs5 = halo_b200.FloatingRegionScore(in_channels=19, size=5, purity_type="hyper", K=50, curvature=1.0)
    assert frs5.size == 5 and frs5.purity_size == 3  # floating_region.py:54-55
    with pytest.raises(AssertionError):
        halo_b200.FloatingRegionScore(size=4, purity_type="radius", curvature=1.0)


def test_dropin_install_registers_reference_module_paths():
    saved = {k: v for k, v in sys.modules.items() if k == "core" or k.startswith("core.")}
    for k in saved:
        del sys.modules[k]
    try:
        patched = halo_b200.install()
        assert "core.utils.hyperbolic" in patched
        from core.utils.hyperbolic import HyperMapper, HyperMLR  # noqa
        from core.active.build import RegionSelection, select_pixels_to_label  # noqa
        from core.active.floating_region import FloatingRegionScore  # noqa

        assert HyperMLR is halo_b200.HyperMLR and RegionSelection is halo_b200.RegionSelection
    finally:
        for k in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_images, tmp):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = pool.shard_range(n_images, rank, world)
        H, W = 6, 10
        cnt = torch.arange(lo, hi, dtype=torch.int32) * 3 + 1
        msk = torch.stack([torch.full((H, W), i % 251, dtype=torch.uint8) for i in range(lo, hi)]) if hi > lo else None
        out = pool.gather_round(cnt if hi > lo else None, msk, n_images)
        torch.save({"n_picked": out["n_picked"], "active_mask": out["active_mask"]}, os.path.join(tmp, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 1])
def test_gather_round_world2_gloo(tmp_path, n_images):
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    for o in outs:
        assert o["n_picked"].tolist() == [3 * i + 1 for i in range(n_images)]
        assert o["active_mask"].shape == (n_images, 6, 10)
        for i in range(n_images):
            assert int(o["active_mask"][i, 0, 0]) == i % 251
    assert torch.equal(outs[0]["active_mask"], outs[1]["active_mask"])


def _delta_case(n_images, H, W, cap, r, seed=0):
    """Synthetic picks per image (distinct centres, some on the border), gt, and the dense masks the round produces."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.randint(0, 19, (n_images, H, W), generator=g, dtype=torch.int64).to(torch.uint8)
    gt[torch.rand((n_images, H, W), generator=g) < 0.1] = 255
    picks = torch.full((n_images, cap), -1, dtype=torch.int32)
    cnt = torch.zeros((n_images,), dtype=torch.int32)
    dense = torch.full((n_images, H, W), 255, dtype=torch.uint8)
    for i in range(n_images):
        n = int(torch.randint(0, cap + 1, (1,), generator=g))
        centres = torch.randperm(H * W, generator=g)[:n]
        centres[:2] = torch.tensor([0, H * W - 1])[:n]   # corner windows are clipped (build.py:45-48)
        picks[i, :n] = centres.int()
        cnt[i] = n
        for p in centres.tolist():
            h, w = divmod(p, W)
            h0, h1, w0, w1 = max(h - r, 0), min(h + r + 1, H), max(w - r, 0), min(w + r + 1, W)
            dense[i, h0:h1, w0:w1] = gt[i, h0:h1, w0:w1]
    return gt, picks, cnt, dense


def test_round_delta_oracle_and_gather_plumbing_cpu():
    for r in (0, 1, 2):
        gt, picks, cnt, dense = _delta_case(4, 9, 13, 7, r, seed=r)
        lab = odelta.pack_round_delta(picks, cnt, gt, r)
        assert lab.shape == (4, 7, (2 * r + 1) ** 2) and lab.dtype == torch.uint8
        masks = torch.full((4, 9, 13), 255, dtype=torch.uint8)
        odelta.apply_round_delta(masks, torch.arange(4, dtype=torch.int32), picks, cnt, lab, r)
        assert torch.equal(masks, dense)
        # padding rows and earlier labels are left alone
        masks2 = torch.full((6, 9, 13), 7, dtype=torch.uint8)
        odelta.apply_round_delta(masks2, torch.tensor([5, -1, 0, -1], dtype=torch.int32), picks, cnt, lab, r)
        assert torch.equal(masks2[1:5], torch.full((4, 9, 13), 7, dtype=torch.uint8))
        out = pool.gather_round_delta(cnt, picks, gt, torch.full((4, 9, 13), 255, dtype=torch.uint8), 4, r,
                                      pack=odelta.pack_round_delta, apply=odelta.apply_round_delta)   # no process group
        with pytest.raises(RuntimeError, match="no CPU path"):   # the product's own pack / apply are CUDA entry points
            pool.pack_round_delta(picks, cnt, gt, r)
        assert torch.equal(out["active_mask"], dense) and torch.equal(out["n_picked"], cnt)


def _delta_worker(rank, world, port, n_images, tmp):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, cap, r = 8, 12, 6, 1
        gt, picks, cnt, dense = _delta_case(n_images, H, W, cap, r, seed=3)
        lo, hi = pool.shard_range(n_images, rank, world)
        masks = torch.full((n_images, H, W), 255, dtype=torch.uint8)
        masks[:, 0, 5] = 3   # a label from an earlier round (on a pixel no window of this case may change: see below)
        out = pool.gather_round_delta(cnt[lo:hi] if hi > lo else None, picks[lo:hi] if hi > lo else None,
                                      gt[lo:hi] if hi > lo else None, masks, n_images, r,
                                      pack=odelta.pack_round_delta, apply=odelta.apply_round_delta)
        dense_all = pool.gather_round(cnt[lo:hi] if hi > lo else None, dense[lo:hi] if hi > lo else None, n_images)
        torch.save({"delta": out, "dense": dense_all}, os.path.join(tmp, "d%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 1])
def test_gather_round_delta_world2_gloo(tmp_path, n_images):
    """The compact exchange (picks + window labels) reproduces what the dense mask all-gather delivers."""
    world = 2
    mp.spawn(_delta_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "d%d.pt" % r)) for r in range(world)]
    for o in outs:
        assert torch.equal(o["delta"]["n_picked"], o["dense"]["n_picked"])
        got, ref = o["delta"]["active_mask"], o["dense"]["active_mask"]
        touched = ref != 255
        assert torch.equal(got[touched], ref[touched])
        untouched = ~touched
        untouched[:, 0, 5] = False
        assert bool((got[untouched] == 255).all())
        keep = ~touched[:, 0, 5]
        assert bool((got[:, 0, 5][keep] == 3).all())   # the earlier round's label survives where this round wrote nothing
    assert torch.equal(outs[0]["delta"]["active_mask"], outs[1]["delta"]["active_mask"])


def test_packed_rows_oracle_roundtrip_cpu():
    """oracle/delta.py restatement of halo_round_rows_pack / _apply: rows -> masks equals the dense result."""
    for r in (0, 1, 2):
        gt, picks, cnt, dense = _delta_case(5, 9, 13, 7, r, seed=10 + r)
        rb = odelta.row_bytes(7, r)
        assert rb % 16 == 0 and rb >= 4 + 4 * 7 + 7 * (2 * r + 1) ** 2
        rows = torch.randint(0, 255, (5, rb), dtype=torch.uint8)      # stale bytes past `count` must never matter
        odelta.pack_rows(rows, picks, cnt, gt, 7, r)
        masks = torch.full((6, 9, 13), 255, dtype=torch.uint8)
        got_cnt = torch.full((6,), -7, dtype=torch.int32)
        odelta.apply_rows(masks, torch.tensor([1, 2, 3, 4, 5], dtype=torch.int32), rows, got_cnt, 7, r)
        assert torch.equal(masks[1:], dense) and bool((masks[0] == 255).all())
        assert got_cnt.tolist() == [-7] + cnt.tolist()
    a = odelta.checksum64(torch.arange(100, dtype=torch.uint8).reshape(4, 5, 5), torch.tensor([1, 2, 3], dtype=torch.int32))
    b = torch.arange(100, dtype=torch.uint8).reshape(4, 5, 5)
    b[3, 4, 4] ^= 1
    assert int(a) != int(odelta.checksum64(b, torch.tensor([1, 2, 3], dtype=torch.int32)))   # position-sensitive
    assert int(a) != int(odelta.checksum64(torch.arange(100, dtype=torch.uint8).reshape(4, 5, 5),
                                           torch.tensor([1, 3, 2], dtype=torch.int32)))


def _exchange_worker(rank, world, port, n_images, tmp):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, cap, r = 8, 12, 6, 1
        gt, picks, cnt, dense = _delta_case(n_images, H, W, cap, r, seed=3)
        lo, hi = pool.shard_range(n_images, rank, world)
        ex = pool.RoundExchange(n_images, H, W, cap, r, "cpu", pack_fn=odelta.pack_rows, apply_fn=odelta.apply_rows)
        assert ex.per == (n_images + world - 1) // world and ex.row_bytes == odelta.row_bytes(cap, r)
        agree = []
        for rnd in range(2):     # the exchange object is reused round after round (buffers allocated once)
            masks = torch.full((n_images, H, W), 255, dtype=torch.uint8)
            n_picked = torch.zeros((n_images,), dtype=torch.int32)
            for b0 in range(lo, hi, 2):           # two batches per shard, packed as they finish
                b1 = min(b0 + 2, hi)
                ex.pack(b0 - lo, picks[b0:b1], cnt[b0:b1], gt[b0:b1])
            ex.exchange(masks, n_picked)
            ex.wait()
            cs, ok = ex.verify(masks, n_picked, checksum_fn=odelta.checksum64)
            agree.append(ok)
        torch.save({"masks": masks, "n_picked": n_picked, "dense": dense, "cnt": cnt, "checksum": cs, "agree": agree},
                   os.path.join(tmp, "x%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [5, 1, 4])
def test_round_exchange_world2_gloo(tmp_path, n_images):
    """ONE packed all-gather at the end of the round: every rank ends with the dense masks and counts of the whole pool,
    and the cross-rank checksum agrees (n_images=1: rank 1 owns no image and still joins the exchange)."""
    world = 2
    mp.spawn(_exchange_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "x%d.pt" % r)) for r in range(world)]
    for o in outs:
        assert torch.equal(o["masks"], o["dense"]) and torch.equal(o["n_picked"], o["cnt"])
        assert o["agree"] == [True, True]
    assert outs[0]["checksum"] == outs[1]["checksum"]
