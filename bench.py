#!/usr/bin/env python
"""bench.py -- acquisition throughput of the HALO hyperbolic hot path on B200 (Mpixel/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

A "step" is one pass of the hot path (fused head K1 -> region score K2 -> budgeted selection K3, plus the
all-gather of counts/masks when N>1) over one batch of synthetic Cityscapes-shaped images that is resident in
HBM: BASELINE.json configs[1] (1280x640 px, 256-d features, 19 classes, 3x3 regions, mask radius 5, 5 % budget,
entropy x radius score, normalised).  The full 2 975-image pool (2.5 TB of features) is streamed through HBM in
such batches; `value` is the steady-state rate of that stream.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = dict(name="cfg2_gtav_cityscapes", pool_images=2975, H=640, W=1280, C=256, O=19, radius_k=1,
                mask_radius_k=5, budget=0.05, n_rounds=1, uncertainty="entropy", purity="radius", normalize=True,
                curvature=1.0, sigma=0.1)
KERNELS_PER_STEP = 7  # head_pack, head_pack_tc, head_fwd_tc, score_init, score_pass_a, score_pass_b, select (+1 rows_pack)


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints its version banner with
    printf when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the life of the
    process and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers of the
    end-to-end arm are first-touched on the GPU's own NUMA node (eight ranks otherwise share one node's memory
    controllers and cross the socket interconnect on every H2D copy)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return True
    except Exception:
        return False


def make_cfg():
    import halo_b200

    w = WORKLOAD
    return halo_b200.AcquisitionConfig(num_classes=w["O"], curvature=w["curvature"], radius_k=w["radius_k"],
                                       mask_radius_k=w["mask_radius_k"], budget=w["budget"], n_rounds=w["n_rounds"],
                                       uncertainty=w["uncertainty"], purity=w["purity"], normalize=w["normalize"])


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_sample(index, rows):
    """Synthetic inputs of one bounded sample: the first `rows` rows of pool image `index`."""
    from halo_b200 import synth

    w = WORKLOAD
    feat = synth.image_features(index, w["C"], w["H"], w["W"], sigma=w["sigma"])[:, :rows].unsqueeze(0).contiguous()
    gt = synth.image_labels(index, w["O"], w["H"], w["W"])[:rows].long()
    return feat, gt


def cpu_reference_step(sample, P, A):
    """One sample through the reference algorithm (fp64 expmap + HyperMLR, FloatingRegionScore, sequential
    arg-max select) -- oracle/acquire.py.  Returns pixels processed."""
    from oracle import acquire as oacquire

    w = WORKLOAD
    feat, gt = sample
    H, W = gt.shape
    oacquire.acquire_image(feat, P, A, gt, torch.zeros((H, W), dtype=torch.bool), torch.zeros((H, W), dtype=torch.bool),
                           torch.full((H, W), 255, dtype=torch.int64), c=w["curvature"], radius_k=w["radius_k"],
                           mask_radius_k=w["mask_radius_k"], budget=w["budget"], n_rounds=w["n_rounds"],
                           unc_type=w["uncertainty"], pur_type=w["purity"], normalize=w["normalize"], fast_select=False)
    return H * W


def cpu_baseline(max_seconds=25.0):
    """Time the oracle port on a bounded sample (whole images of the bench workload) for ~10-30 s."""
    from halo_b200 import synth

    w = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    P, A = synth.head_params(w["O"], w["C"], seed=0, dtype=torch.float64)
    rows = w["H"]
    strip = cpu_sample(0, 64)
    t0 = time.perf_counter()
    cpu_reference_step(strip, P, A)  # warm-up on a strip (thread pools, allocator)
    warm = time.perf_counter() - t0
    if warm * (w["H"] / 64) > max_seconds:  # slow host: shrink the sample to a strip of rows
        rows = max(64, int(w["H"] * max_seconds / (warm * (w["H"] / 64))) // 32 * 32)
    px, t, n = 0, 0.0, 0
    while n < 4:
        sample = cpu_sample(1 + n, rows)  # generation is not timed
        t1 = time.perf_counter()
        px += cpu_reference_step(sample, P, A)
        t += time.perf_counter() - t1
        n += 1
        if t > max_seconds * 0.6:
            break
    return {"value": round(px / t / 1e6, 5), "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d image(s) x %dx%d px of the bench workload through oracle/acquire.py (fp64 head, box-filter "
                      "score, sequential arg-max selection), %.1f s" % (n, rows, w["W"], t)}


def run_reference(args):
    """--impl reference: the reference's own algorithm for this path on the host cores (oracle port; the
    reference is pure Python/PyTorch and cannot travel to the GPU box -- see DESIGN.md)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from halo_b200 import synth

    w = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    P, A = synth.head_params(w["O"], w["C"], seed=0, dtype=torch.float64)
    rows = w["H"]
    strip = cpu_sample(0, 64)
    t0 = time.perf_counter()
    cpu_reference_step(strip, P, A)
    est_full = (time.perf_counter() - t0) * (w["H"] / 64)
    budget_s = 240.0
    total = args.steps + args.warmup
    if est_full * total > budget_s:
        rows = max(32, int(w["H"] * budget_s / (est_full * total)) // 32 * 32)
    ring = [cpu_sample(i, rows) for i in range(2)]  # inputs are generated before the timed region
    for i in range(args.warmup):
        cpu_reference_step(ring[i % 2], P, A)
    t1 = time.perf_counter()
    px = 0
    for i in range(args.steps):
        px += cpu_reference_step(ring[i % 2], P, A)
    dt = time.perf_counter() - t1
    value = px / dt / 1e6
    sample = "per step: first %d rows of one %dx%d image (%d px) through oracle/acquire.py" % (rows, w["H"], w["W"], rows * w["W"])
    line = {
        "impl": "reference", "metric": "acquisition_throughput", "value": round(value, 5), "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / max(args.steps, 1) * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(None, sample),
        "cpu_baseline": {"value": round(value, 5), "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 5), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(batch, note=None):
    w = WORKLOAD
    cfg = {"workload": "GTAV->Cityscapes-shaped acquisition round (BASELINE.json configs[1]): %d x %dx%d px pool, "
                       "%d-d features, %d classes, 3x3 regions, mask radius %d, %.0f %% budget single-shot (%d picks/img), "
                       "entropy x radius, normalised" % (w["pool_images"], w["W"], w["H"], w["C"], w["O"], w["mask_radius_k"],
                                                          w["budget"] * 100, 4552),
           "image": [w["H"], w["W"]], "channels": w["C"], "classes": w["O"], "budget": w["budget"],
           "l2_policy": "inputs larger than L2 (one batch of features >> 126 MB)"}
    if batch is not None:
        cfg["images_per_step_per_gpu"] = batch
    if note:
        cfg["sample"] = note
    return cfg


# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------
def traffic_from_profiles(batch):
    """DRAM bytes per launch of K1 from the committed ncu --set full capture, scaled to `batch` images -- only while the
    capture still describes the kernel: the json names the sha256 of csrc/head_fwd_tc.cu it was taken on."""
    import hashlib

    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            t = json.load(f)
        with open(os.path.join(ROOT, "halo_b200", "csrc", "head_fwd_tc.cu"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        if t.get("kernel_source_sha256") != sha:
            return None, "profiles/k1_traffic.json is stale (K1 source changed since the ncu capture)"
        return float(t["dram_bytes_per_image"]) * batch, "profiles/k1_traffic.json (%s)" % t.get("source", "ncu")
    except Exception as e:  # noqa: BLE001
        return None, "unavailable (%s)" % type(e).__name__


class PoolShard:
    """This rank's shard of the 2 975-image pool, streamed through ONE resident batch of HBM (124 GB of features per 148
    images: the pool itself is 2.5 TB).  `provider(lo, hi)` regenerates the features of pool images [lo, hi) in place,
    per image from its own seed (so any sharding of the pool sees the same images), resets the state planes, and brackets
    itself with CUDA events so that the generation stays OUT of the measured round time."""

    def __init__(self, B, C, O, H, W, sigma, dev):
        self.B, self.C, self.O, self.H, self.W, self.sigma, self.dev = B, C, O, H, W, sigma, dev
        self.feat = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        self.gt = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        self.active = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
        self.selected = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
        self.active_mask = torch.full((B, H, W), 255, dtype=torch.uint8, device=dev)
        self.gen = torch.Generator(device=dev)
        self.segments = []          # (start_event, end_event) of the timed stretches between two generations
        self._open = None

    def fill(self, lo, hi):
        for k, i in enumerate(range(lo, hi)):
            self.gen.manual_seed(1234 + i)
            self.feat[k].normal_(0.0, self.sigma, generator=self.gen)
            self.gt[k] = torch.randint(0, self.O, (self.H, self.W), generator=self.gen, device=self.dev, dtype=torch.int16).to(torch.uint8)
            self.gt[k].masked_fill_(torch.rand((self.H, self.W), generator=self.gen, device=self.dev) < 0.05, 255)

    def close_segment(self):
        if self._open is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.segments.append((self._open, e))
            self._open = None

    def provider(self, lo, hi):
        self.close_segment()
        b = hi - lo
        self.fill(lo, hi)
        self.active[:b].zero_(); self.selected[:b].zero_(); self.active_mask[:b].fill_(255)
        self._open = torch.cuda.Event(enable_timing=True)
        self._open.record()
        return dict(feat=self.feat[:b], gt=self.gt[:b], active=self.active[:b], selected=self.selected[:b],
                    active_mask=self.active_mask[:b])

    def timed_ms(self):
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.segments)


def run_full_round(shard, P, A, cfg, world, distributed, dev):
    """ONE real acquisition round over the whole 2 975-image pool (BASELINE.json configs[1] / [2]): every image generated
    from its own seed, scored and selected on the rank that owns it, ONE packed all-gather at the end, every rank replays
    all rows onto its replica of the pool's masks, and a 64-bit checksum of (masks, counts) is compared across ranks.
    Timed = the K1/K2/K3/pack stretches of every batch + the exchange and replay (feature generation excluded), max over
    ranks."""
    import torch.distributed as dist

    from halo_b200 import pool

    w = WORKLOAD
    n_images, H, W = w["pool_images"], w["H"], w["W"]
    ex = pool.RoundExchange(n_images, H, W, cfg.regions_per_image(H, W), cfg.radius_k, dev)
    masks = torch.full((n_images, H, W), 255, dtype=torch.uint8, device=dev)
    shard.segments.clear()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    out = pool.acquire_pool(shard.provider, n_images, P, A, cfg, batch_size=shard.B, exchange=ex, masks=masks, keep_local=False)
    shard.close_segment()
    ms = torch.tensor([shard.timed_ms()], dtype=torch.float64, device=dev)
    wall = time.perf_counter() - t0
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    checksum, agree = ex.verify(out["active_mask"], out["n_picked"])
    lo, hi = out["range"]
    cnt = out["n_picked"]
    secs = float(ms.item()) / 1e3
    return {"images": n_images, "images_this_rank": hi - lo, "batches_this_rank": len(shard.segments),
            "seconds": round(secs, 4), "Mpixel/s": round(n_images * H * W / secs / 1e6, 1),
            "wall_seconds_with_feature_generation": round(wall, 2), "exchanges": 1 if distributed else 0,
            "exchange_bytes_per_rank": ex.bytes_per_rank() if distributed else 0,
            "picks_min": int(cnt.min().item()), "picks_max": int(cnt.max().item()), "labelled_pixels": int((out["active_mask"] != 255).sum().item()),
            "checksum": "%016x" % checksum, "replicas_agree": agree,
            "note": "timed: K1+K2+K3+pack of every batch and the final all-gather + replay (CUDA events, max over ranks); "
                    "per-image feature generation between batches is not timed (the pool, 2.5 TB, does not fit HBM)"}


def run_ours(args):
    import torch.distributed as dist

    import halo_b200
    from halo_b200 import pool, synth
    from halo_b200 import _native as nat

    w = WORKLOAD
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    nat.load()
    cfg = make_cfg()
    B, C, O, H, W = args.batch, w["C"], w["O"], w["H"], w["W"]
    P, A = synth.head_params(O, C, seed=0, device=dev)
    # this rank's shard of the pool, streamed through one resident batch (B images = 124 GB of features at B = 148)
    n_images = w["pool_images"]
    lo, hi = pool.shard_range(n_images, rank, world)
    shard = PoolShard(B, C, O, H, W, w["sigma"], dev)
    shard.fill(lo, min(lo + B, hi))
    feat, gt = shard.feat, shard.gt
    # the steps walk through consecutive rounds over the shard: batches of B images (the last one of a round is ragged),
    # the round's ONE exchange after its last batch.  Features of the resident batch are reused by every step (the pool
    # does not fit HBM); the mask state is fresh per step, allocated up front so that no fill runs inside the timed region.
    sizes = [min(B, hi - b0) for b0 in range(lo, hi, B)]
    R = len(sizes)
    total_steps = args.steps + args.warmup
    n_state = min(total_steps, 16)
    state = [dict(active=torch.zeros((B, H, W), dtype=torch.uint8, device=dev),
                  selected=torch.zeros((B, H, W), dtype=torch.uint8, device=dev),
                  active_mask=torch.full((B, H, W), 255, dtype=torch.uint8, device=dev)) for _ in range(n_state)]

    def reset(s):
        s["active"].zero_(); s["selected"].zero_(); s["active_mask"].fill_(255)

    ex = pool.RoundExchange(n_images, H, W, cfg.regions_per_image(H, W), cfg.radius_k, dev)
    replica = torch.full((n_images, H, W), 255, dtype=torch.uint8, device=dev)
    counts = torch.zeros((n_images,), dtype=torch.int32, device=dev)
    launches = [0]

    k1_events = []     # (start, end) around the K1 launch of every full-size batch of the TIMED steps (rank 0)

    def step(s, k, timed=False):
        """Step k of the job = batch (k mod R) of this rank's shard; after a round's last batch, its exchange."""
        j = k % R
        b = sizes[j]
        ev = None
        if timed and rank == 0 and b == B:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            k1_events.append(ev)
        res = halo_b200.acquire_batch(feat[:b], P, A, cfg, gt[:b], s["active"][:b], s["selected"][:b], s["active_mask"][:b],
                                      want_picks=True, head_events=ev)
        ex.pack(j * B, res["picks"], res["n_picked"], gt[:b])
        launches[0] += KERNELS_PER_STEP + 1
        if j == R - 1:
            ex.exchange(replica, counts)     # side stream: overlaps the next round's first batch
            launches[0] += 1
        return res, b

    for i in range(args.warmup):
        res, _ = step(state[i % n_state], i)
    ex.wait()
    torch.cuda.synchronize()
    picks_ok = int(res["n_picked"].min().item()) if args.warmup else None
    for s in state:
        reset(s)
    launches[0] = 0
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    done, images = 0, 0
    step_starts = []     # (start event, batch size) of every timed step on rank 0: per-step durations for the roofline block
    while done < args.steps:
        chunk = min(n_state, args.steps - done)
        for j in range(chunk):
            if rank == 0:
                es = torch.cuda.Event(enable_timing=True)
                es.record()
                step_starts.append((es, sizes[(args.warmup + done + j) % R]))
            res, b = step(state[j], args.warmup + done + j, timed=True)
            images += b
        done += chunk
        if done < args.steps:  # recycle the state planes (outside the hot path; counted in the timed region)
            for j in range(min(n_state, args.steps - done)):
                reset(state[j])
    ex.wait()                  # the last exchange in flight belongs to the timed steps
    ev1.record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    clocks = sampler.finish()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    img_all = torch.tensor([images], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(img_all, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    px_total = float(img_all.item()) * H * W
    value = px_total / (ms_total / 1e3) / 1e6
    timed_launches = launches[0]

    # ---- roofline of the dominant kernel (K1, fused head), timed alone on its launch stream ----
    roof = None
    e2e = None
    cpu = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        reps = 5
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            halo_b200.head_forward(feat, P, A, cfg.curvature, want_logits=False, want_radius=True, want_pixunc=True)
        k0.record()
        for _ in range(reps):
            halo_b200.head_forward(feat, P, A, cfg.curvature, want_logits=False, want_radius=True, want_pixunc=True)
        k1.record()
        torch.cuda.synchronize()
        k_ms_alone = k0.elapsed_time(k1) / reps
        # the figure the roofline is quoted on: K1 as it ran INSIDE the timed steps (events on its launch stream around every
        # full-size batch); the back-to-back repetition above is kept beside it (it draws more power, so it clocks lower)
        k_ms = (sum(a_.elapsed_time(b_) for a_, b_ in k1_events) / len(k1_events)) if k1_events else k_ms_alone
        # whole-step time of the same full-size batches (start of the step to the start of the next one): at N > 1 the last
        # batch of a shard is smaller, so ms_per_step (mean over ALL steps) is not the step K1's launch belongs to
        full = [step_starts[i][0].elapsed_time(step_starts[i + 1][0]) for i in range(len(step_starts) - 1) if step_starts[i][1] == B]
        step_ms_full = (sum(full) / len(full)) if full else None
        alg_bytes = 4.0 * C * B * H * W  # features read once; SURVEY 8(d): radius/entropy planes are not algorithmic
        achieved = alg_bytes / (k_ms / 1e3) / 1e9
        traffic, traffic_src = traffic_from_profiles(B)
        roof = {"bound": "hbm", "kernel": "head_fwd_tc_kernel (K1 fused head, tcgen05)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms": round(k_ms, 4), "kernel_launches_timed": len(k1_events),
                "timing": "CUDA events on the launch stream around K1 inside the timed steps (mean over the full-size batches)"
                          if k1_events else "K1 repeated alone (no full-size batch in the timed steps)",
                "kernel_ms_repeated_alone": round(k_ms_alone, 4), "algorithmic_bytes_per_launch": alg_bytes,
                "step_ms_full_batch": round(step_ms_full, 4) if step_ms_full else None,
                "step_frac": round((4.0 * C + 13) * px_total / world / (ms_total / 1e3) / 1e9 / peak, 4)}
    if distributed:
        dist.barrier()

    # ---- the real round: the whole pool, sharded, ONE exchange, replicas checked against each other ----
    full_round = None
    del replica
    if not args.no_round:
        full_round = run_full_round(shard, P, A, cfg, world, distributed, dev)
        shard.fill(lo, min(lo + B, hi))
    torch.cuda.empty_cache()

    # ---- the other acquisition configurations of BASELINE.json / SURVEY 8(d), short runs on the same resident batch ----
    others = None
    if rank == 0 and world == 1 and not args.no_side_configs:
        others = run_side_configs(feat, gt, state[0], dev)

    # ---- BASELINE.json configs[4] beside it (not the headline): fused head forward + backward, batch 8, fp32 ----
    train = None
    if rank == 0 and world == 1 and not args.no_train_step:
        train = run_train_step(feat[:8], P, A, cfg.curvature, measured_peak_gbs()[0])
        torch.cuda.empty_cache()
    if distributed:
        dist.barrier()

    # ---- the drop-in path the Lightning learner calls, at the real pipeline shape (not the headline) ----
    dropin = None
    if rank == 0 and world == 1 and not args.no_dropin:
        dropin = run_dropin(dev)
    if distributed:
        dist.barrier()

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region ----
    bind_to_gpu_numa_node(local)   # pinned staging buffers first-touched on the GPU's own NUMA node (this leg only)
    e2e = run_e2e(args, cfg, P, A, dev, rank, world, distributed)
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, all_cpus)   # the CPU arm gets every host core again
        cpu = cpu_baseline()
    if rank == 0:
        line = {
            "metric": "acquisition_throughput", "value": round(value, 2), "unit": "Mpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": timed_launches,
            "picks_per_image": picks_ok, "round": full_round, "other_configs": others, "train_step": train, "dropin": dropin,
            "step_schedule": {"batches_per_round_per_gpu": R, "images_per_batch": sizes, "exchange": "one packed all-gather "
                              "per round (after its last batch, on a side stream)" if distributed else "none (one GPU)"},
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()


def run_side_configs(feat, gt, st, dev):
    """Side measurements (not the headline), 5 steps each after 2 warm-up steps, inputs resident:
    (a) the reference's default per-round budget BUDGET/len(SELECT_ITER) = 1 % -> 911 picks per image (build.py:78);
    (b) BASELINE.json configs[3], SYNTHIA->Cityscapes-shaped: 16 classes, 5x5 regions (radius_K=2), 2.2 % budget."""
    import halo_b200
    from halo_b200 import synth

    w = WORKLOAD
    B, C, H, W = feat.shape
    out = []
    cases = [("reference default budget: 5 %% over 5 rounds = 911 picks/img, 19 classes, 3x3", w["O"],
              halo_b200.AcquisitionConfig(num_classes=w["O"], curvature=w["curvature"], radius_k=1, mask_radius_k=w["mask_radius_k"],
                                          budget=w["budget"], n_rounds=5, uncertainty="entropy", purity="radius", normalize=True)),
             ("BASELINE.json configs[3] SYNTHIA-shaped: 16 classes, 5x5 regions, 2.2 %% budget = 721 picks/img", 16,
              halo_b200.AcquisitionConfig(num_classes=16, curvature=w["curvature"], radius_k=2, mask_radius_k=w["mask_radius_k"],
                                          budget=0.022, n_rounds=1, uncertainty="entropy", purity="radius", normalize=True))]
    if B % 4 == 0 and C % 4 == 0:
        # (c) the head width every shipped HALO config uses (REDUCED_CHANNELS 64, core/configs/defaults.py:14): the first
        #     quarter of the resident bytes viewed as B images of C/4 channels -- same pixels per step as the headline
        cases.append(("headline configuration at the shipped head width: %d-d features (a quarter of the bytes per pixel)" % (C // 4),
                      w["O"], cases[0][2].__class__(num_classes=w["O"], curvature=w["curvature"], radius_k=1,
                                                   mask_radius_k=w["mask_radius_k"], budget=w["budget"], n_rounds=1,
                                                   uncertainty="entropy", purity="radius", normalize=True)))
    for name, O, cfg in cases:
        narrow = "shipped head width" in name
        fx = feat[:B // 4].reshape(B, C // 4, H, W) if narrow else feat
        P, A = synth.head_params(O, fx.shape[1], seed=0, device=dev)
        g = gt if O == w["O"] else torch.where(gt == 255, gt, gt % O)

        def step():
            st["active"].zero_(); st["selected"].zero_(); st["active_mask"].fill_(255)
            return halo_b200.acquire_batch(fx, P, A, cfg, g, st["active"], st["selected"], st["active_mask"])

        for _ in range(2):
            res = step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            res = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out.append({"workload": name if narrow else name % (), "ms_per_step": round(ms, 3), "Mpixel/s": round(B * H * W / ms / 1e3, 1),
                    "picks_per_image": int(res["n_picked"].min().item()), "images_per_step": B,
                    "note": "includes the three state-plane resets per step"})
    return out


def run_train_step(feat8, P, A, c, peak):
    """configs[4]: hyperbolic MLR head fused fwd+bwd on a resident batch (8 x 1280x640 x 256-d, 19 classes, fp32), as the
    autograd Function runs it: the training forward writes the logits AND keeps the per-pixel contractions (164 B per
    pixel), the streaming backward reads them, dlogits and the features ONCE and writes du (+ dP, dA).
    Algorithmic bytes 12*C + 8*O per pixel.  Beside it: the same step at C = 64, the channel count every
    shipped HALO config uses (core/configs/defaults.py:14), where the per-pixel math, not HBM, bounds the step."""
    import halo_b200
    from halo_b200 import _native as nat
    from halo_b200 import synth

    def measure(feat, P, A):
        B, C, H, W = feat.shape
        O = P.shape[0]
        dl = torch.randn((B, O, H, W), device=feat.device, generator=torch.Generator(device=feat.device).manual_seed(7)) * 1e-3

        def step():
            r = halo_b200.head_forward(feat, P, A, c, want_logits=True, want_saved=True)
            return halo_b200.head_backward(feat, P, A, c, dl, saved=r["saved"])

        for _ in range(3):
            step()
        path = list(nat.last_path())
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        t0.record()
        for _ in range(reps):
            step()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        px = B * H * W
        alg = (12.0 * C + 8.0 * O) * px
        return {"workload": "head fwd+bwd, batch %d x %dx%d x %d-d, %d classes, fp32" % (B, W, H, C, O),
                "ms_per_step": round(ms, 3), "Mpixel/s": round(px / ms / 1e3, 1), "algorithmic_GB_per_step": round(alg / 1e9, 2),
                "GB/s": round(alg / ms / 1e6, 1), "frac_of_hbm_peak": round(alg / ms / 1e6 / peak, 4), "backward_path": path,
                "gpu_launches_per_step": 8}

    out = measure(feat8, P, A)
    out["workload"] = "BASELINE.json configs[4]: " + out["workload"]
    B, C, H, W = feat8.shape
    P64, A64 = synth.head_params(P.shape[0], 64, seed=0, device=feat8.device)
    out["c64"] = measure(feat8[:2].reshape(B, 64, H, W), P64, A64)   # the same bytes viewed as 8 images of 64 channels
    out["fused_loss"] = run_loss_step(feat8.device, P.shape[0])
    out["head_block"] = run_head_block_step(feat8.device, P.shape[0])
    return out


def run_loss_step(dev, O):
    """SURVEY 8f row 4 at the reference's training shape: logits 8 x O x 160x320 -> crop 640x1280 (classifier.py:556-557),
    CrossEntropy(ignore 255) on 5 % labelled pixels + negative-learning loss, forward AND backward to the low-resolution
    logits: `halo_seg_loss` (two kernels, nothing of size (N,O,H,W) materialised) beside the reference's torch sequence run
    eagerly on the same GPU (F.interpolate -> softmax -> CE + NegativeLearningLoss -> autograd)."""
    import torch.nn.functional as F

    from halo_b200.losses import fused_seg_loss

    N, h, w, H, W = 8, 160, 320, 640, 1280
    g = torch.Generator(device=dev).manual_seed(11)
    logits = torch.randn((N, O, h, w), device=dev, generator=g) * 3.0
    labels = torch.randint(0, O, (N, H, W), device=dev, generator=g)
    labels[torch.rand((N, H, W), device=dev, generator=g) > 0.05] = 255
    lab8 = labels.to(torch.uint8)

    def ours():
        x = logits.detach().requires_grad_(True)
        loss, _, _ = fused_seg_loss(x, lab8, (H, W), 1.0)
        loss.backward()
        return loss, x.grad

    def torch_seq():
        x = logits.detach().requires_grad_(True)
        out = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=True)
        p = torch.softmax(out, dim=1)
        m = (p < 0.05).detach()
        loss = F.cross_entropy(out, labels, ignore_index=255) + torch.sum(-1 * m * torch.log(1 - p + 1e-6)) / torch.sum(m)
        loss.backward()
        return loss, x.grad

    res = {}
    vals = {}
    for name, fn in (("halo_seg_loss", ours), ("torch_eager_sequence", torch_seq)):
        for _ in range(2):
            vals[name] = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    a, b = vals["halo_seg_loss"], vals["torch_eager_sequence"]
    res["loss_rel_diff"] = abs(float(a[0].detach()) - float(b[0].detach())) / abs(float(b[0].detach()))
    res["grad_rel_diff_of_max"] = float((a[1] - b[1]).abs().max() / b[1].abs().max())
    res["workload"] = "fwd+bwd of CE(ignore 255) + negative-learning loss on logits %dx%dx%dx%d up-sampled to %dx%d" % (N, O, h, w, H, W)
    return res


def run_head_block_step(dev, O):
    """SURVEY 8f rows 3 + 4 around the head at the reference's training shape (batch 4, decoder features 512 x 160x320,
    REDUCED_CHANNELS 64, labels 640x1280): conv_reduce + HFR in training mode -> fused head -> fused loss, forward AND
    backward to the decoder features and every parameter, as ONE autograd graph of this library's kernels.  Beside it the
    conv_reduce + HFR part alone against the reference's torch call sequence (classifier.py:527-550) run eagerly on the
    same GPU with the same modules."""
    import torch.nn as nn
    import torch.nn.functional as F

    import halo_b200
    from halo_b200.hfr import reduce_hfr
    from halo_b200.losses import fused_seg_loss

    N, Cin, C, h, w, H, W = 4, 512, 64, 160, 320, 640, 1280
    torch.manual_seed(5)
    conv = nn.Conv2d(Cin, C, 1).to(dev).train()
    mlp = nn.Sequential(nn.Linear(C, C), nn.BatchNorm1d(C), nn.ReLU(), nn.Linear(C, C)).to(dev).train()
    head = halo_b200.HyperMLR(C, O, c=1.0).to(dev)
    mapper = halo_b200.HyperMapper(c=1.0)
    g = torch.Generator(device=dev).manual_seed(12)
    feats = torch.randn((N, Cin, h, w), device=dev, generator=g) * 0.3
    labels = torch.randint(0, O, (N, H, W), device=dev, generator=g)
    labels[torch.rand((N, H, W), device=dev, generator=g) > 0.05] = 255
    lab8 = labels.to(torch.uint8)
    params = list(conv.parameters()) + list(mlp.parameters()) + list(head.parameters())

    def zero():
        for p in params:
            p.grad = None

    def block():
        zero()
        x = feats.detach().requires_grad_(True)
        z = reduce_hfr(x, conv, mlp)
        out = head(mapper.expmap(z, dim=1))
        loss, _, _ = fused_seg_loss(out, lab8, (H, W), 1.0)
        loss.backward()
        return x.grad

    def hfr_ours():
        zero()
        x = feats.detach().requires_grad_(True)
        z = reduce_hfr(x, conv, mlp)
        z.backward(dz)
        return z, x.grad, conv.weight.grad

    def hfr_torch():
        zero()
        x = feats.detach().requires_grad_(True)
        y = conv(x)
        t = mlp(y.permute(0, 2, 3, 1).contiguous().view(-1, C)).view(-1, h * w, C)
        wt = torch.clamp(torch.mean(t, dim=1).view(-1, C, 1, 1), min=1e-5)
        z = F.normalize(y.reshape(-1, C, h * w), dim=-1).reshape(-1, C, h, w) * wt
        z.backward(dz)
        return z, x.grad, conv.weight.grad

    dz = torch.randn((N, C, h, w), device=dev, generator=g)
    res, vals = {}, {}
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False     # the torch arm in fp32 like this library (cuDNN's default would run the conv in TF32)
    for name, fn in (("head_block_fwd_bwd", block), ("reduce_hfr_train", hfr_ours), ("reduce_hfr_torch_eager", hfr_torch)):
        for _ in range(2):
            vals[name] = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = round(e0.elapsed_time(e1) / 5, 3)
    torch.backends.cudnn.allow_tf32 = tf32_was
    a, b = vals["reduce_hfr_train"], vals["reduce_hfr_torch_eager"]
    res["z_rel_diff_of_max"] = float((a[0] - b[0]).abs().max() / b[0].abs().max())
    res["dfeat_rel_diff_of_max"] = float((a[1] - b[1]).abs().max() / b[1].abs().max())
    res["dWr_rel_diff_of_max"] = float((a[2] - b[2]).abs().max() / b[2].abs().max())
    res["note"] = ("the differences are single pixels whose hidden pre-activation sits on the ReLU kink (fp32 rounding decides the "
                   "side): against the float64 sequence both fp32 arms are at 1e-6 elsewhere (tools/hfr_accuracy.py)")
    res["workload"] = ("training step of the head block: conv_reduce %d->%d + HFR (BatchNorm on batch statistics), fused head, "
                       "fused loss; batch %d, decoder features %dx%d, labels %dx%d, %d classes" % (Cin, C, N, h, w, H, W, O))
    return res


class _Features(dict):
    """Stand-in for the image tensor of a loader item: the backbone is out of scope (BASELINE.json north_star), so the item
    carries the decoder features the backbone would have produced; `.cuda()` / `.shape` / `.device` are what
    RegionSelection touches (core/active/build.py:94,112)."""
    shape = (1, 3, 640, 1280)
    device = torch.device("cpu")

    def cuda(self, non_blocking=False):
        r = _Features({k: v.cuda(non_blocking=non_blocking) for k, v in self.items()})
        r.device = next(iter(r.values())).device
        return r


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def run_dropin(dev, n_images=24):
    """The path the Lightning learner calls (core/train_learners.py:318-322): `RegionSelection(cfg, feature_extractor,
    classifier, loader, round)` of the drop-in package over a fake loader at the REAL pipeline shape -- decoder features
    160x320 (C = 64, the shipped REDUCED_CHANNELS), labels 1024x2048, 19 classes, 1 % budget per round (BUDGET 0.05 over 5
    SELECT_ITER), entropy x radius, batch 1 -- including the int64 CPU planes the loader hands over, the fused up-sampling
    in front of the score, the batched selection and the PNG + indicator files on disk.  Wall clock, host side included.
    Beside it: the reference sequence (build.py:92-166) by the oracle on the host cores, on a bounded sample."""
    import shutil
    import tempfile

    import halo_b200
    from halo_b200 import synth

    C, O, h, w, H, W = 64, 19, 160, 320, 1024, 2048
    cfg = _Cfg(ACTIVE=_Cfg(RADIUS_K=1, MASK_RADIUS_K=5, BUDGET=0.05, SELECT_ITER=[0, 1, 2, 3, 4], UNCERTAINTY="entropy",
                           PURITY="radius", K=100, NORMALIZE=True, VIZ_MASK=False),
               MODEL=_Cfg(NUM_CLASSES=O, CURVATURE=1.0, HYPER=True))

    class Head(torch.nn.Module):   # ASPP_Classifier_V2_Hyper.forward (core/models/classifier.py:365-379) on precomputed features
        def __init__(self):
            super().__init__()
            self.mapper = halo_b200.HyperMapper(c=1.0)
            self.conv_seg = halo_b200.HyperMLR(C, O, c=1.0)

        def forward(self, x, size=None):
            embed = self.mapper.expmap(x["out"], dim=1)
            return self.conv_seg(embed.double()).float(), embed

    head = Head().to(dev)
    tmp = tempfile.mkdtemp(prefix="halo_dropin_")

    def items(lo, hi):
        out = []
        for i in range(lo, hi):
            out.append({"img": _Features({"out": synth.image_features(i, C, h, w, sigma=0.15)[None]}),
                        "path_to_mask": [os.path.join(tmp, "m%d.png" % i)], "path_to_indicator": [os.path.join(tmp, "i%d.pth" % i)],
                        "origin_mask": [torch.full((H, W), 255, dtype=torch.long)], "origin_label": [synth.image_labels(i, O, H, W).long()],
                        "size": [(H, W)], "active": [torch.zeros((H, W), dtype=torch.bool)], "selected": [torch.zeros((H, W), dtype=torch.bool)]})
        return out

    try:
        halo_b200.RegionSelection(cfg, torch.nn.Identity(), head, items(0, 4), round_number=1)      # warm-up
        loader = items(4, 4 + n_images)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        halo_b200.RegionSelection(cfg, torch.nn.Identity(), head, loader, round_number=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        from PIL import Image
        import numpy as np

        labelled = int((np.array(Image.open(loader[0]["path_to_mask"][0])) != 255).sum())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    res = {"workload": "drop-in RegionSelection, %d images: features %dx%dx%d -> labels %dx%d, %d classes, 1 %% budget per round "
                       "(2 331 picks), batch 1, files written" % (n_images, C, h, w, H, W, O),
           "ms_per_image": round(dt / n_images * 1e3, 2), "images/s": round(n_images / dt, 2),
           "Mpixel/s_label_res": round(n_images * H * W / dt / 1e6, 1), "labelled_pixels_image0": labelled}
    # reference sequence on the CPU (oracle port), one image
    try:
        from oracle import acquire as oacquire
        from oracle import head as ohead
        from oracle import select as oselect
        import numpy as np

        torch.set_num_threads(os.cpu_count() or 1)
        P = head.conv_seg.P_MLR.detach().cpu().double()
        A = head.conv_seg.A_MLR.detach().cpu().double()
        u = synth.image_features(4, C, h, w, sigma=0.15)[None]
        gt = synth.image_labels(4, O, H, W).long()
        t1 = time.perf_counter()
        lo, x, _ = ohead.head_forward(u, P, A, 1.0)
        sc, _, _ = oacquire.upsampled_score(lo, x, (H, W), unc_type="entropy", pur_type="radius", normalize=True, ground_truth=None,
                                            in_channels=O, size=3, ctor_purity_type="radius", K=100, c=1.0)
        n_regions = int(np.ceil(H * W * 0.01 / 9))
        oselect.select_sequential(sc, n_regions, 1, 5, torch.zeros((H, W), dtype=torch.bool), torch.zeros((H, W), dtype=torch.bool),
                                  torch.full((H, W), 255, dtype=torch.int64), gt)
        cpu_dt = time.perf_counter() - t1
        res["cpu_reference"] = {"ms_per_image": round(cpu_dt * 1e3, 1), "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "1 image through oracle head + F.interpolate (fp64 embedding) + score + sequential "
                                          "arg-max selection, no file I/O"}
    except Exception as e:  # noqa: BLE001
        res["cpu_reference"] = {"error": repr(e)[:200]}
    return res


def run_e2e(args, cfg, P, A, dev, rank, world, distributed):
    """Same metric through the public API from pinned HOST memory: per step the features, labels and mask state of
    `e2e_batch` images are copied host->device, scored and selected, and counts + updated masks copied back."""
    import torch.distributed as dist

    import halo_b200

    w = WORKLOAD
    Be, C, O, H, W = args.e2e_batch, w["C"], w["O"], w["H"], w["W"]
    g = torch.Generator().manual_seed(99 + rank)
    h_feat = torch.empty((Be, C, H, W), dtype=torch.float32).pin_memory()
    h_feat.normal_(0.0, w["sigma"], generator=g)
    h_gt = torch.randint(0, O, (Be, H, W), generator=g, dtype=torch.int64).to(torch.uint8).pin_memory()
    h_active = torch.zeros((Be, H, W), dtype=torch.uint8).pin_memory()
    h_mask_in = torch.full((Be, H, W), 255, dtype=torch.uint8).pin_memory()
    h_mask_out = torch.empty((Be, H, W), dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty((Be,), dtype=torch.int32).pin_memory()
    # double-buffered device staging: copy image i+1 while image i is scored (copy stream + compute stream)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [dict(feat=torch.empty((1, C, H, W), dtype=torch.float32, device=dev),
                  gt=torch.empty((1, H, W), dtype=torch.uint8, device=dev),
                  active=torch.empty((1, H, W), dtype=torch.uint8, device=dev),
                  selected=torch.zeros((1, H, W), dtype=torch.uint8, device=dev),
                  mask=torch.empty((1, H, W), dtype=torch.uint8, device=dev),
                  ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    d_cnt = torch.empty((Be,), dtype=torch.int32, device=dev)
    compute = torch.cuda.current_stream(dev)

    def one_step():
        for s in slots:
            s["free"].record(compute)
        for i in range(Be):
            s = slots[i % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(s["free"])
                s["feat"].copy_(h_feat[i:i + 1], non_blocking=True)
                s["gt"].copy_(h_gt[i:i + 1], non_blocking=True)
                s["active"].copy_(h_active[i:i + 1], non_blocking=True)
                s["mask"].copy_(h_mask_in[i:i + 1], non_blocking=True)
                s["ready"].record(copy_stream)
            compute.wait_event(s["ready"])
            s["selected"].zero_()
            res = halo_b200.acquire_batch(s["feat"], P, A, cfg, s["gt"], s["active"], s["selected"], s["mask"])
            d_cnt[i:i + 1].copy_(res["n_picked"])
            h_mask_out[i:i + 1].copy_(s["mask"], non_blocking=True)
            s["free"].record(compute)
        h_cnt.copy_(d_cnt, non_blocking=True)

    steps = max(2, min(args.steps, 5))
    for _ in range(2):
        one_step()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert int(h_cnt.min()) > 0
    px = Be * H * W * world * steps
    h2d = Be * (C * H * W * 4 + 3 * H * W)
    d2h = Be * (H * W + 4)
    return {"value": round(px / (float(ms.item()) / 1e3) / 1e6, 2), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "images_per_step_per_gpu": Be, "steps": steps,
            "note": "pinned host buffers -> PCIe H2D (double-buffered per image) -> K1/K2/K3 -> D2H of counts + masks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=148,
                    help="images resident per GPU per step (148 = one selection CTA per SM; 124 GB of features)")
    ap.add_argument("--e2e-batch", type=int, default=4, help="images per end-to-end step (pinned host memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the configs[4] fwd+bwd side measurement")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the 911-pick and configs[3] side measurements")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in RegionSelection leg (fake loader, files on disk)")
    ap.add_argument("--no-round", action="store_true", help="skip the full 2 975-image round (run once after the timed steps)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
