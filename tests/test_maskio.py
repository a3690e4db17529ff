"""On-disk mask / indicator round trip: the asynchronous writer produces the files the reference's
synchronous sequence (core/active/build.py:162-166) produces, and the readers mirror core/datasets/cityscapes.py:234,
245-251."""
import os

import numpy as np
import pytest
import torch

from halo_b200 import maskio


def _planes(seed, H, W, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    mask = torch.randint(0, 19, (H, W), generator=g, dtype=torch.int64)
    mask[torch.rand((H, W), generator=g) < 0.9] = 255
    active = torch.rand((H, W), generator=g) < 0.2
    selected = active & (torch.rand((H, W), generator=g) < 0.5)
    return mask.to(device), active.to(device), selected.to(device)


def _check_pair(tmp_path, i, mask, active, selected):
    from PIL import Image

    ref_png, ref_pth = tmp_path / ("ref_%d.png" % i), tmp_path / ("ref_%d.pth" % i)
    maskio.write_mask_sync(mask, active, selected, ref_png, ref_pth)
    got_png, got_pth = tmp_path / "out" / ("m_%d.png" % i), tmp_path / "ind" / ("i_%d.pth" % i)
    assert Image.open(got_png).mode == Image.open(ref_png).mode == "L"
    assert np.array_equal(np.array(Image.open(got_png)), np.array(Image.open(ref_png)))
    assert torch.equal(maskio.read_mask(got_png), mask.cpu().long())
    got, ref = torch.load(got_pth), torch.load(ref_pth)
    for k in ("active", "selected"):
        assert got[k].dtype == torch.bool and got[k].device.type == "cpu"
        assert torch.equal(got[k], ref[k])
    a, s = maskio.read_indicator(got_pth, mask.cpu())
    assert torch.equal(a, active.cpu().bool()) and torch.equal(s, selected.cpu().bool())


def test_async_writer_matches_reference_sequence_cpu(tmp_path):
    items = [_planes(i, 33 + 8 * (i % 3), 47 + i) for i in range(12)]   # changing plane sizes re-allocate staging slots
    with maskio.AsyncMaskWriter(workers=3, depth=4) as w:
        for i, (m, a, s) in enumerate(items):
            w.write(m, a, s, tmp_path / "out" / ("m_%d.png" % i), tmp_path / "ind" / ("i_%d.pth" % i))
    for i, (m, a, s) in enumerate(items):
        _check_pair(tmp_path, i, m, a, s)


def test_first_use_indicator_and_uint8_planes(tmp_path):
    p = tmp_path / "fresh.pth"
    torch.save({"active": torch.tensor([0]), "selected": torch.tensor([0])}, p)   # the files the reference starts from
    like = torch.zeros((5, 7), dtype=torch.int64)
    a, s = maskio.read_indicator(p, like)
    assert a.dtype == torch.bool and a.shape == (5, 7) and not a.any() and not s.any()
    # the batched device path keeps uint8 planes (0/1 flags, 255 = unlabeled): same files
    m, act, sel = _planes(3, 16, 24)
    with maskio.AsyncMaskWriter(workers=1) as w:
        w.write(m.to(torch.uint8), act.to(torch.uint8), sel.to(torch.uint8), tmp_path / "out" / "m_0.png", tmp_path / "ind" / "i_0.pth")
    _check_pair(tmp_path, 0, m, act, sel)


def test_writer_reports_io_errors(tmp_path):
    blocker = tmp_path / "file"
    blocker.write_text("x")
    m, a, s = _planes(0, 8, 8)
    w = maskio.AsyncMaskWriter(workers=1)
    w.write(m, a, s, blocker / "sub" / "m.png", tmp_path / "i.pth")   # a file stands where a directory is needed
    with pytest.raises(Exception):
        w.flush()
    w._errors.clear()
    w.close()


@pytest.mark.gpu
def test_async_writer_device_planes(tmp_path):
    dev = "cuda"
    items = [_planes(10 + i, 64, 96, device=dev) for i in range(6)]
    with maskio.AsyncMaskWriter(workers=2, depth=3) as w:
        for i, (m, a, s) in enumerate(items):
            # mixed residency like the reference: CUDA mask, CPU flags for odd items
            if i % 2:
                a, s = a.cpu(), s.cpu()
            w.write(m, a, s, tmp_path / "out" / ("m_%d.png" % i), tmp_path / "ind" / ("i_%d.pth" % i))
            m.fill_(7)   # the caller may reuse its planes right after write(): the copy was ordered before this fill
    for i in range(6):
        m, a, s = _planes(10 + i, 64, 96)
        _check_pair(tmp_path, i, m, a, s)
