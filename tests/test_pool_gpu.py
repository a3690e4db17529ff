"""GPU: the sharded round end to end (halo_b200.pool.acquire_pool / RoundExchange) -- packed-row pack / apply / checksum
kernels against the oracle's restatement, the single-process round against per-batch acquisition, and (when the box has
two GPUs) the real NCCL exchange with a cross-rank checksum, including a rank that owns no image.
Reference: core/active/build.py:58-62,92 (the loop the reference runs on rank 0 alone, core/train_learners.py:307-326)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import halo_b200
from halo_b200 import pool, synth
from oracle import delta as odelta

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _provider(C, O, H, W, device):
    def provider(lo, hi):
        return synth.batch(lo, hi, C, O, H, W, device=device)
    return provider


def test_packed_rows_kernels_match_oracle():
    C, O, H, W, B = 32, 19, 48, 80, 5
    for rk, budget in ((1, 0.03), (0, 0.004), (2, 0.05)):
        cfg = halo_b200.AcquisitionConfig(num_classes=O, radius_k=rk, mask_radius_k=3, budget=budget)
        P, A = synth.head_params(O, C, seed=2, device=DEV)
        d = synth.batch(0, B, C, O, H, W, device=DEV)
        d["active"][1] = 1
        d["active"][1, :2, :2] = 0                            # image 1 cannot meet its budget (one pickable corner): count < cap
        res = halo_b200.acquire_batch(d["feat"], P, A, cfg, d["gt"], d["active"], d["selected"], d["active_mask"], want_picks=True)
        cap = cfg.regions_per_image(H, W)
        assert int(res["n_picked"][1]) < cap
        lib_rb = halo_b200._native.load().halo_round_row_bytes(cap, rk)
        assert lib_rb == odelta.row_bytes(cap, rk)
        rows = torch.zeros((B, lib_rb), dtype=torch.uint8, device=DEV)
        pool.pack_rows(rows, res["picks"], res["n_picked"], d["gt"], cap, rk)
        ref_rows = torch.zeros((B, lib_rb), dtype=torch.uint8)
        odelta.pack_rows(ref_rows, res["picks"].cpu(), res["n_picked"].cpu(), d["gt"].cpu(), cap, rk)
        assert torch.equal(rows.cpu(), ref_rows)
        masks = torch.full((B + 1, H, W), 255, dtype=torch.uint8, device=DEV)
        cnt = torch.full((B + 1,), -1, dtype=torch.int32, device=DEV)
        row_image = torch.tensor([1, 2, 3, 4, 5], dtype=torch.int32, device=DEV)
        pool.apply_rows(masks, row_image, rows, cnt, cap, rk)
        assert torch.equal(masks[1:], d["active_mask"]) and bool((masks[0] == 255).all())
        assert cnt.tolist() == [-1] + res["n_picked"].tolist()
        cs = pool.checksum64(masks, cnt)
        assert int(cs) == int(odelta.checksum64(masks.cpu(), cnt.cpu()))
    odd = torch.arange(0, 203, dtype=torch.uint8, device=DEV)          # a tail that is not a multiple of 8 bytes
    assert int(pool.checksum64(odd, torch.tensor([5], dtype=torch.int32, device=DEV))) == \
        int(odelta.checksum64(odd.cpu(), torch.tensor([5], dtype=torch.int32)))


def test_acquire_pool_single_process_equals_batches():
    C, O, H, W, n = 32, 19, 48, 80, 7
    cfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.03)
    P, A = synth.head_params(O, C, seed=2, device=DEV)
    out = halo_b200.acquire_pool(_provider(C, O, H, W, DEV), n, P, A, cfg, batch_size=3, verify=True)
    ref = synth.batch(0, n, C, O, H, W, device=DEV)
    r = halo_b200.acquire_batch(ref["feat"], P, A, cfg, ref["gt"], ref["active"], ref["selected"], ref["active_mask"])
    assert torch.equal(out["n_picked"], r["n_picked"]) and torch.equal(out["active_mask"], ref["active_mask"])
    assert torch.equal(out["active_mask_local"], ref["active_mask"]) and torch.equal(out["selected_local"], ref["selected"])
    assert out["replicas_agree"] is True and out["checksum"] == int(odelta.checksum64(ref["active_mask"].cpu(), r["n_picked"].cpu())) & (2**64 - 1)
    # a second round onto the same replica keeps the first round's labels
    masks = out["active_mask"].clone()
    first = masks != 255
    state = {"active": out["active_local"], "selected": out["selected_local"], "active_mask": out["active_mask_local"]}

    def provider2(lo, hi):
        b = synth.batch(lo, hi, C, O, H, W, seed=77, device=DEV)
        for k in state:
            b[k] = state[k][lo:hi].clone()
        return b

    out2 = halo_b200.acquire_pool(provider2, n, P, A, cfg, batch_size=4, masks=masks)
    assert torch.equal(out2["active_mask"][first], out["active_mask"][first])
    assert int((out2["active_mask"] != 255).sum()) > int(first.sum())
    assert torch.equal(out2["active_mask"], out2["active_mask_local"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, n_images, tmp):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        C, O, H, W = 32, 19, 48, 80
        cfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.03)
        P, A = synth.head_params(O, C, seed=2, device=dev)
        ex = pool.RoundExchange(n_images, H, W, cfg.regions_per_image(H, W), cfg.radius_k, dev)
        out = halo_b200.acquire_pool(_provider(C, O, H, W, dev), n_images, P, A, cfg, batch_size=2, exchange=ex, verify=True)
        dense = pool.gather_round(out.get("n_picked_local"), out.get("active_mask_local"), n_images)   # the dense form agrees
        torch.save({"masks": out["active_mask"].cpu(), "n_picked": out["n_picked"].cpu(), "checksum": out["checksum"],
                    "agree": out["replicas_agree"], "dense": dense["active_mask"].cpu(), "dense_cnt": dense["n_picked"].cpu()},
                   os.path.join(tmp, "n%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("n_images", [5, 1])
def test_acquire_pool_nccl_world2(tmp_path, n_images):
    """The real exchange on hardware: two ranks, one NCCL all-gather of packed rows, replicas proven equal by the
    cross-rank checksum and equal to the single-process round (n_images=1: rank 1 owns nothing and still joins)."""
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "n%d.pt" % r)) for r in range(world)]
    C, O, H, W = 32, 19, 48, 80
    cfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.03)
    P, A = synth.head_params(O, C, seed=2, device=DEV)
    ref = synth.batch(0, n_images, C, O, H, W, device=DEV)
    r = halo_b200.acquire_batch(ref["feat"], P, A, cfg, ref["gt"], ref["active"], ref["selected"], ref["active_mask"])
    for o in outs:
        assert o["agree"] is True
        assert torch.equal(o["masks"], ref["active_mask"].cpu()) and torch.equal(o["n_picked"], r["n_picked"].cpu())
        assert torch.equal(o["dense"], o["masks"]) and torch.equal(o["dense_cnt"], o["n_picked"])
    assert outs[0]["checksum"] == outs[1]["checksum"]
