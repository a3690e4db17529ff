"""On-disk round trip of the active-learning state.

The reference writes, per image and synchronously inside the acquisition loop (core/active/build.py:162-166),
  * the new label mask as an 8-bit PNG (255 = unlabeled) at `path_to_mask`, and
  * `{"active": bool (H,W), "selected": bool (H,W)}` with `torch.save` at `path_to_indicator`,
and the dataset reads them back in the DataLoader workers (core/datasets/cityscapes.py:234, 245-251).  The formats
are kept byte-compatible here; what changes is WHEN the work happens: the device planes are copied to pinned
staging buffers on a side stream and PNG encoding / pickling run on a small thread pool, so the GPU never waits for
zlib.  `flush()` (also called on context exit) restores the reference's post-condition: every file is on disk.
"""
import os
import queue
import threading

import numpy as np
import torch

UNLABELED = 255


def read_mask(path):
    """cityscapes.py:234 -- `np.array(Image.open(path), dtype=np.uint8)` as an int64 tensor (origin_mask, :242)."""
    from PIL import Image

    return torch.from_numpy(np.array(Image.open(path), dtype=np.uint8)).long()


def read_indicator(path, like):
    """cityscapes.py:245-251 -- the saved bool planes, or all-False planes shaped like `like` on first use
    (the initial files hold `torch.tensor([0])`)."""
    ind = torch.load(path)
    active, selected = ind["active"], ind["selected"]
    if tuple(active.size()) == (1,):
        active = torch.zeros_like(like, dtype=torch.bool)
        selected = torch.zeros_like(like, dtype=torch.bool)
    return active, selected


def write_mask_sync(active_mask, active, selected, path_mask, path_indicator):
    """The reference's own sequence (build.py:162-166) for CPU-resident results."""
    from PIL import Image

    Image.fromarray(np.array(active_mask.cpu().numpy(), dtype=np.uint8)).save(path_mask)
    torch.save({"active": active.cpu().bool(), "selected": selected.cpu().bool()}, path_indicator)


class AsyncMaskWriter:
    """Write (mask PNG, indicator .pth) pairs off the acquisition loop's critical path.

    write() accepts device or host planes of any integer/bool dtype: `active_mask` (H,W) label ids with 255 =
    unlabeled, `active` / `selected` (H,W) flags.  Device planes are copied device->host asynchronously on a private
    stream into pinned staging slots (`depth` of them; write() blocks when all are in flight)."""

    def __init__(self, workers=4, depth=8, device=None):
        self.depth = depth
        self._free = queue.Queue()
        self._jobs = queue.Queue()
        self._slots = {}
        self._errors = []
        self._pending = 0
        self._cv = threading.Condition()
        self._stream = None
        self._device = device
        self._threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()
        self._next_slot = 0

    # -- staging ---------------------------------------------------------------------------------------
    def _slot(self, shape, cuda):
        key = (tuple(shape), cuda)
        try:
            sid = self._free.get_nowait()
            if self._slots[sid]["key"] != key:      # plane size changed: re-allocate this slot
                self._slots[sid] = self._alloc(key)
            return sid
        except queue.Empty:
            pass
        if self._next_slot < self.depth:
            sid = self._next_slot
            self._next_slot += 1
            self._slots[sid] = self._alloc(key)
            return sid
        sid = self._free.get()                       # all slots in flight: wait for a worker
        if self._slots[sid]["key"] != key:
            self._slots[sid] = self._alloc(key)
        return sid

    @staticmethod
    def _alloc(key):
        shape, cuda = key
        mk = (lambda: torch.empty((3,) + shape, dtype=torch.uint8).pin_memory()) if cuda else (lambda: torch.empty((3,) + shape, dtype=torch.uint8))
        return {"key": key, "host": mk(), "event": torch.cuda.Event() if cuda else None}

    # -- producer side ---------------------------------------------------------------------------------
    def write(self, active_mask, active, selected, path_mask, path_indicator):
        if self._errors:
            raise self._errors[0]
        cuda = active_mask.is_cuda or active.is_cuda or selected.is_cuda
        sid = self._slot(active_mask.shape[-2:], cuda)
        slot = self._slots[sid]
        host = slot["host"]
        planes = (active_mask, active, selected)
        if cuda:
            dev = next(p.device for p in planes if p.is_cuda)
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream(dev)
            # device planes: snapshot as uint8 on the producer's stream (the caller may overwrite its planes right after
            # write() returns), then copy the snapshot on the side stream so the next image's kernels are not queued
            # behind the PCIe transfer
            staged = [p.reshape(p.shape[-2:]).to(torch.uint8, copy=True) if p.is_cuda else None for p in planes]
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                for k, p in enumerate(staged):
                    if p is not None:
                        host[k].copy_(p, non_blocking=True)
                        p.record_stream(self._stream)
                slot["event"].record(self._stream)
        for k, p in enumerate(planes):
            if not p.is_cuda:   # host planes (the reference keeps `active` / `selected` on the CPU): straight into the slot
                host[k].copy_(p.reshape(p.shape[-2:]).to(torch.uint8))
        with self._cv:
            self._pending += 1
        self._jobs.put((sid, str(path_mask), str(path_indicator)))

    # -- consumer side ---------------------------------------------------------------------------------
    def _worker(self):
        from PIL import Image

        while True:
            job = self._jobs.get()
            if job is None:
                return
            sid, path_mask, path_indicator = job
            try:
                slot = self._slots[sid]
                if slot["event"] is not None:
                    slot["event"].synchronize()
                host = slot["host"]
                for p in (path_mask, path_indicator):
                    d = os.path.dirname(p)
                    if d:
                        os.makedirs(d, exist_ok=True)
                Image.fromarray(host[0].numpy()).save(path_mask, format="PNG")      # build.py:162-164
                torch.save({"active": host[1].bool(), "selected": host[2].bool()}, path_indicator)  # build.py:165-166
            except Exception as e:  # surfaced by the next write() / flush()
                self._errors.append(e)
            finally:
                self._free.put(sid)
                with self._cv:
                    self._pending -= 1
                    self._cv.notify_all()

    def flush(self):
        with self._cv:
            while self._pending:
                self._cv.wait()
        if self._errors:
            raise self._errors[0]

    def close(self):
        self.flush()
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
