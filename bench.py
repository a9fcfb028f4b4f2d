#!/usr/bin/env python
"""bench.py -- labelled pixels/s of the fused upsample + OHEM-CE + confusion-matrix fwd+bwd path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU port of the reference path, host cores)

One step = one pass of the hot path over one Cityscapes-shaped batch per GPU (BASELINE.json configs[1]):
logits 8x19x256x512 fp32 -> labels 8x1024x2048 int64, OhemCrossEntropy forward + backward and the GT x pred
confusion matrix.  Through the C ABI that is sf_ce_onepass (one walk: loss, matrix, unscaled gradient) ->
sf_loss_finalize (OHEM decision on the device) -> sf_ce_onepass_backward (1/K scale, or the top-k re-walk);
`--route split` times the older sf_ce_forward -> sf_loss_finalize -> sf_ce_backward route instead.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "labelled pixels/s (fused upsample+OHEM-CE+mIoU fwd+bwd), % of HBM roofline"
UNIT = "px/s"
B, C, LH, LW, H, W = 8, 19, 256, 512, 1024, 2048          # per-GPU batch of configs[1]
N_SETS = 4                                                  # rotating input sets: 4 x 214 MB >> 126 MB L2
WORKLOAD = "cityscapes_ohem_fwd_bwd_hist: logits 8x19x256x512 fp32 -> labels 8x1024x2048 int64 per GPU"
CSRC = os.path.join(ROOT, "segmentation_factory_b200", "csrc")


def step_bytes_per_px(c, esz, lbytes, scale, ohem):
    """algorithmic bytes per label pixel of the fwd+bwd path (logits at 1/scale^2 density)."""
    return 3.0 * c * esz / (scale * scale) + 2 * lbytes + (12 if ohem else 0)


def onepass_min_bytes_per_px(c, esz, lbytes, scale, ohem):
    """What the one-pass walk itself must move: logits read, gradient sums written (fp32), label read, loss_px written."""
    return c * esz / (scale * scale) + c * 4.0 / (scale * scale) + lbytes + (4 if ohem else 0)


STEP_SOURCES = ("sf_onepass.cuh", "sf_onepass.cu", "sf_onepass_f32.cu", "sf_common.cuh", "sf_fast.cuh", "sf_select.cu")


def src_sha16():
    """Hash of the sources of the kernels one bench step launches (walk, finalize, backward): ncu-derived numbers under
    profiles/ are only quoted when they were captured on exactly these sources."""
    h = hashlib.sha256()
    for name in STEP_SOURCES:
        h.update(open(os.path.join(CSRC, name), "rb").read())
    return h.hexdigest()[:16]


def synth_batch(torch, seed, kind="early", b=B, c=C, lh=LH, lw=LW, h=H, w=W, dtype=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    coarse = torch.randint(0, c, (b, h // 32, w // 32), generator=g)
    coarse[torch.rand(coarse.shape, generator=g) < 0.1] = 255
    lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    if kind == "early":
        low = 3.0 * torch.randn((b, c, lh, lw), generator=g)
    else:
        l2 = coarse.repeat_interleave(lh * 32 // h, 1).repeat_interleave(lw * 32 // w, 2)
        l2 = torch.where(l2 == 255, torch.zeros_like(l2), l2)
        flip = torch.rand(l2.shape, generator=g) < 0.02
        l2 = torch.where(flip, torch.randint(0, c, l2.shape, generator=g), l2)
        low = torch.randn((b, c, lh, lw), generator=g) + 12.0 * torch.nn.functional.one_hot(l2, c).permute(0, 3, 1, 2).float()
    if dtype is not None:
        low = low.to(dtype)
    return low.contiguous(), lab


def torch_reference_step(torch, low, lab, c, ohem=True, hist=True):
    """The reference path as the reference issues it (models/build_models.py:65, util/losses.py:52-61, util/metrics.py:25-27,
    autograd): the same ATen calls on whatever device `low` lives on.  A timed comparator, never the product path."""
    F = torch.nn.functional
    x = low.detach().clone().requires_grad_(True)
    up = F.interpolate(x, size=lab.shape[1:], mode="bilinear", align_corners=False).float()
    if ohem:
        t = -torch.log(torch.tensor(0.7, dtype=torch.float))
        n_min = lab[lab != 255].numel() // 16
        loss = F.cross_entropy(up, lab, ignore_index=255, reduction="none").view(-1)
        hard = loss[loss > t]
        if hard.numel() < n_min:
            hard, _ = loss.topk(n_min)
        loss = torch.mean(hard)
    else:
        loss = F.cross_entropy(up, lab, ignore_index=255)
    loss.backward()
    if hist:
        with torch.no_grad():
            pred = up.argmax(1).flatten()
            tgt = lab.flatten()
            keep = tgt != 255
            torch.bincount(tgt[keep] * c + pred[keep], minlength=c * c)
    return loss.item()


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  The timed region is
    tens of milliseconds, shorter than nvidia-smi's start-up, so NVML is polled from a thread (every ~2 ms) while the
    main thread waits for the GPU; nvidia-smi is only the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, torch, local):
        import threading
        self.samples, self.mask, self.max_mhz, self.stop_flag = [], 0, None, False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(local).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while not self.stop_flag and self.h is not None:
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                break
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": len(self.samples), "source": "nvml"}
        if self.samples:
            v = sorted(self.samples)
            out["sm_mhz"] = v[len(v) // 2]
            out["reasons"] = sorted(n for bit, n in self.REASONS.items() if self.mask & bit)
            return out
        try:                                                          # fallback: one nvidia-smi query after the run
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True,
                               text=True, timeout=20).stdout.strip().splitlines()[0].split(",")
            out.update(sm_mhz=float(r[0]), sm_max_mhz=float(r[1]), source="nvidia-smi after the timed region")
        except Exception:
            pass
        return out


def cpu_port_step(orc, np, low, lab):
    """The reference path on the host (oracle port): forward + matrix, OHEM reduction, backward."""
    f = orc.path_forward(low, lab, 255, None, orc.OHEM_THRESH, want_arg=False)
    loss, kth, kept, branch = orc.ohem_reduce(f["loss_px"], int(f["sums"][2]), orc.OHEM_THRESH)
    lp = f["loss_px"]
    coef = np.where(lp > kth, np.float32(1.0 / max(kept, 1)), np.float32(0)).astype(np.float32)
    orc.path_backward(low, lab, f["lse_px"], coef)
    return loss


def base_config(kind, world):
    """The `config` object both arms print (identical keys and values, so the driver's same_config check holds)."""
    return {"workload": WORKLOAD, "logits": kind, "labels": "int64, 10% ignore=255",
            "batch_per_gpu": B, "classes": C, "parallelism": f"dp{world}"}


def torch_cpu_sample(torch, kind):
    """The reference's own torch path on the host cores (util/losses.py + F.interpolate + Metrics.update), bounded
    sample: ONE image of the configs[1] batch per step."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lo, la = synth_batch(torch, 1234, kind, b=1)
    torch_reference_step(torch, lo, la, C)
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or (time.perf_counter() - t0 < 4 and reps < 8):
        torch_reference_step(torch, lo, la, C)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": H * W / dt, "unit": UNIT, "cores": cores, "threads": torch.get_num_threads(), "kind": "reference torch ops (restated)",
            "sample": f"{reps} steps of 1 image (2.1 Mpx) of the configs[1] batch, {dt:.2f} s each"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python/torch and
    /root/reference does not exist on the GPU box, so this is the C port (oracle/, kind "port") with OpenMP over
    every host core.  Each step is --gpus full configs[1] batches (the work the N ranks of the other arm do per step),
    a bounded sample of the same workload.  Under torch.distributed.run only rank 0 works; the launcher's
    OMP_NUM_THREADS=1 is overridden explicitly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    orc.set_num_threads(cores)
    n = max(1, args.gpus)
    sets = [synth_batch(torch, 1234 + 97 * r, args.kind) for r in range(min(n, 2))]
    sets = [(lo.numpy(), la.numpy()) for lo, la in sets]
    steps, warm = min(args.steps, max(2, 12 // n)), min(args.warmup, 2)     # bounded: a step is n x ~0.5 s on 16 cores
    for _ in range(warm):
        cpu_port_step(orc, np, *sets[0])
    t0 = time.perf_counter()
    for _ in range(steps):
        for r in range(n):
            cpu_port_step(orc, np, *sets[r % len(sets)])
    dt = (time.perf_counter() - t0) / steps
    v = n * B * H * W / dt
    sample = f"{steps} step(s) of {n} full 8-image batch(es) (16.8 Mpx each) per step"
    extra = torch_cpu_sample(torch, args.kind) if not args.no_cpu_baseline else None
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.kind, n),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
        "torch_cpu_reference": extra,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa(local):
    """Pin this rank's host threads to the CPUs nearest its GPU (NVML affinity), so pinned staging buffers are
    first-touched on that NUMA node and the H2D copies do not all cross one socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * wi + bit for wi, wd in enumerate(words) for bit in range(64) if (wd >> bit) & 1 and 64 * wi + bit < n_cpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "n": len(cpus)}
    except Exception as e:                                           # best effort; report why
        return {"error": str(e)[:80]}
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kind", default="early", choices=["early", "late"],
                    help="logit distribution: early -> OHEM threshold branch, late -> top-k (radix select) branch")
    ap.add_argument("--route", default="onepass", choices=["onepass", "split"],
                    help="onepass: sf_ce_onepass + finalize + onepass_backward; split: sf_ce_forward + finalize + sf_ce_backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other-config / comparator legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local) if world > 1 else None           # before torch allocates pinned memory

    import numpy as np
    import torch
    import torch.distributed as dist

    import segmentation_factory_b200 as sf
    from segmentation_factory_b200 import _lib

    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL printf()s its version banner to stdout at communicator creation; stdout must carry ONE JSON line, so fd 1 is
        # pointed at stderr while the communicator is created and the first collective runs.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            ctypes.CDLL(None).fflush(None)
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    L = _lib.lib()
    thr = float(-torch.log(torch.tensor(0.7, dtype=torch.float)))
    cthr = ctypes.c_float(thr)
    n_px = B * H * W
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    ws_bytes = L.sf_workspace_bytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- inputs: N_SETS distinct batches resident in HBM, pinned host copies for the end-to-end leg ----
    host_sets = [synth_batch(torch, 1234 + 97 * rank + i, args.kind) for i in range(N_SETS)]
    dev_sets = [(lo.to(dev), la.to(dev)) for lo, la in host_sets]
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    loss_px = torch.empty(n_px, dtype=torch.float32, device=dev)
    lse_px = torch.empty(n_px, dtype=torch.float32, device=dev)      # split route only
    grad_acc = torch.empty((B, C, LH, LW), dtype=torch.float32, device=dev)
    grad = torch.empty((B, C, LH, LW), dtype=torch.float32, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    hist = torch.zeros(C * C + 2, dtype=torch.int64, device=dev)     # matrix || kept || valid, one all-reduce
    onepass = args.route == "onepass" and L.sf_ce_onepass_supported(0, dev_sets[0][0].data_ptr(), C, LH, LW, H, W) == 1
    geo = (B, C, LH, LW, H, W, 255)

    def launch_step(sets, i, evs=None, reduce_every_step=False, onepass=onepass):
        lo, la = sets[i % len(sets)]
        if evs:
            evs[0].record(stream)
        if onepass:
            rc = L.sf_ce_onepass(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, cthr, loss_px.data_ptr(), hist.data_ptr(),
                                 grad_acc.data_ptr(), ws.data_ptr(), ws_bytes, st)
        else:
            rc = L.sf_ce_forward(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, cthr, loss_px.data_ptr(), lse_px.data_ptr(),
                                 hist.data_ptr(), ws.data_ptr(), ws_bytes, st)
        if evs:
            evs[1].record(stream)
        rc |= L.sf_loss_finalize(loss_px.data_ptr(), n_px, cthr, ws.data_ptr(), ws_bytes, out.data_ptr(), st)
        if evs:
            evs[2].record(stream)
        work = None
        if world > 1 and reduce_every_step:
            work = dist.all_reduce(hist, async_op=True)          # runs on NCCL's stream underneath the backward
        if onepass:
            rc |= L.sf_ce_onepass_backward(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, loss_px.data_ptr(), grad_acc.data_ptr(),
                                           None, ws.data_ptr(), grad.data_ptr(), st)
        else:
            rc |= L.sf_ce_backward(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, loss_px.data_ptr(), lse_px.data_ptr(), None,
                                   ws.data_ptr(), grad.data_ptr(), st)
        if evs:
            evs[3].record(stream)
        if rc:
            raise RuntimeError(f"kernel launch failed rc={rc}")
        if work is not None:
            work.wait()                                          # stream-level wait, no host sync

    def timed(sets, steps, reduce_every_step=False, sample_clocks=False, onepass=onepass):
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(torch, local) if (sample_clocks and rank == 0) else None
        barrier()
        e0.record(stream)
        for i in range(steps):
            launch_step(sets, i, evs[i], reduce_every_step, onepass)
        if world > 1 and not reduce_every_step:
            # the path's one exchange (engine.py:100-101: once per evaluation, not per batch): the int64 matrix and
            # counters, one SUM over NVLink after the last step
            dist.all_reduce(hist)
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        avg = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in evs) / len(evs)
        return t.item() / steps, {"walk" if onepass else "fwd": avg(0, 1), "finalize": avg(1, 2), "backward": avg(2, 3)}, clocks

    for i in range(args.warmup):
        launch_step(dev_sets, i)
    barrier()
    ms_step, kern, clocks = timed(dev_sets, args.steps, sample_clocks=True)
    value = world * n_px / (ms_step * 1e-3)
    state = np.frombuffer(ws[:64].cpu().numpy().tobytes(), dtype=np.int32)
    branch = int(state[4])
    reduce_leg = None
    if world > 1:
        ms_r, _, _ = timed(dev_sets, args.steps, reduce_every_step=True)
        reduce_leg = {"ms_per_step": ms_r, "value": world * n_px / (ms_r * 1e-3), "note": "all_reduce(int64[C*C+2]) issued every step"}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(kern, key=kern.get)
    kname = {"walk": "sf::onepass_kernel<float,19,4,OP_FUSED,OHEM=1,HIST=1>", "fwd": "sf::fwd_fast_kernel<float,19,4,FF_SOFTMAX>",
             "backward": "sf::onepass_kernel<float,19,4,OP_BWD> (top-k re-walk)" if onepass else "sf::bwd_fast_kernel<float,19,4>",
             "finalize": "sf::finalize_kernel"}[dom]
    step_bpp = step_bytes_per_px(C, 4, 8, 4, True)
    dom_bpp = {"walk": step_bpp - 4.0 - 4.0,                     # the walk does the whole fwd+bwd except the select's and the scale's reads
               "fwd": C * 4 / 16.0 + 8 + 4, "backward": 2 * C * 4 / 16.0 + 8 + 4, "finalize": 4.0}[dom]
    achieved = dom_bpp * n_px / (kern[dom] * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture of this build under profiles/"
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")     # dram read+write bytes per launch from ncu --set full
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("src_sha16") == src_sha16():
            traffic, traffic_note = tj.get(dom), f"ncu --set full, {tj.get('source', 'profiles/')}"
        else:
            traffic_note = "profiles/traffic_latest.json was captured on different kernel sources (src_sha16 mismatch)"
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_px": dom_bpp,
                "frac_of_hbm3e_spec_8000GBs": achieved / 8000.0, "kernel_ms": kern,
                "step_frac": step_bpp * n_px / (ms_step * 1e-3) / 1e9 / peak, "step_bytes_per_px": step_bpp,
                "walk_min_traffic_bytes_per_px": onepass_min_bytes_per_px(C, 4, 8, 4, True) if onepass else None,
                "note": "frac = SURVEY 8(d) algorithmic bytes of the work the dominant kernel performs / its event-timed "
                        "duration / measured HBM peak; step_frac = 42.25 B/px x px / whole step.  At C=19 and 1/16 logit "
                        "density the path is issue/MUFU-bound, not HBM-bound (DESIGN.md section 5)"}

    extras = {}
    if world == 1 and not args.no_extras:
        extras = extra_legs(torch, sf, L, dev, args, launch_step, timed, peak, thr)

    # ---- end to end through the public API: pinned host inputs, H2D every step, loss + matrix read back ----
    pinned = [(lo.pin_memory(), la.pin_memory()) for lo, la in host_sets]
    crit_hist = torch.zeros(C, C, dtype=torch.int64, device=dev)
    hist_host = torch.empty((C, C), dtype=torch.int64).pin_memory()

    # Double-buffered like a DataLoader(pin_memory=True) + non_blocking consumer: the H2D copy of batch i+1 is issued on
    # a copy stream while batch i is computed; every step still copies its own 214 MB inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    slots = [None, None]

    # Second leg (reported beside `e2e`, never instead of it): the annotation bytes cross PCIe as raw uint8
    # ids and are mapped on the device by sf_label_lut; the loss / metric kernels then read uint8 labels.
    CS_TABLE = [255] * 35                                    # Cityscapes labelIds -> trainIds (public dataset definition)
    for t, r in enumerate([7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]):
        CS_TABLE[r] = t
    lut = sf.LabelLUT(CS_TABLE, device=dev)
    inv = torch.zeros(256, dtype=torch.uint8)                # a raw id for every train id; 255 -> raw id 0 (unlabeled)
    for r, t in enumerate(CS_TABLE):
        if t != 255:
            inv[t] = r
    pinned_raw = [inv[la.to(torch.int64)].contiguous().pin_memory() for _, la in host_sets]

    def stage(i, raw_ids):
        lo, la = pinned[i % N_SETS]
        with torch.cuda.stream(copy_stream):
            x = lo.to(dev, non_blocking=True)
            y = (pinned_raw[i % N_SETS] if raw_ids else la).to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        slots[i % 2] = (x, y, ev)

    def e2e_step(i, raw_ids):
        x, y, ev = slots[i % 2]
        stream.wait_event(ev)
        x.record_stream(stream); y.record_stream(stream)
        stage(i + 1, raw_ids)                                # next batch's copy overlaps this batch's kernels
        if raw_ids:
            y = lut(y)                                       # uint8 ids -> uint8 train ids on the device
        x.requires_grad_(True)
        loss = sf.fused_cross_entropy(x, y, None, 255, thr, crit_hist)
        loss.backward()
        hist_host.copy_(crit_hist, non_blocking=True)
        return loss.item()                                   # D2H + host sync, like engine.py:44

    def e2e_leg(raw_ids):
        stage(0, raw_ids)
        for i in range(3):
            e2e_step(i, raw_ids)
        barrier()
        n = max(5, min(args.steps, 20))
        t0 = time.perf_counter()
        for i in range(3, 3 + n):
            last = e2e_step(i, raw_ids)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = host_sets[0][0].numel() * 4 + host_sets[0][1].numel() * (1 if raw_ids else 8)
        return {"value": world * n_px / dt.item(), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 + C * C * 8,
                "steps": n, "ms_per_step": dt.item() * 1e3, "h2d_gbs_per_rank": h2d / dt.item() / 1e9}, last

    e2e, loss_i64 = e2e_leg(False)
    e2e_u8, loss_u8 = e2e_leg(True)
    e2e_u8["labels"] = "raw uint8 ids over PCIe + sf_label_lut on the device"
    e2e_u8["same_loss_as_int64_leg"] = bool(loss_i64 == loss_u8)
    if numa is not None:
        e2e["host_affinity"] = numa
    h2d = host_sets[0][0].numel() * 4 + host_sets[0][1].numel() * 8

    # ---- CPU baselines (rank 0, N = 1 only): the oracle port and the reference's torch ops on the box's host cores ----
    cpu = torch_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        orc.set_num_threads(os.cpu_count() or 1)
        lo, la = host_sets[0][0].numpy(), host_sets[0][1].numpy()
        cpu_port_step(orc, np, lo[:1], la[:1])
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            cpu_port_step(orc, np, lo, la)
            reps += 1
        cdt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_px / cdt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
               "sample": f"{reps} full 8-image batches (16.8 Mpx each), {cdt:.2f} s per batch"}
        torch_cpu = torch_cpu_sample(torch, args.kind)

    if rank == 0:
        cfg = base_config(args.kind, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "run": {"route": "onepass" if onepass else "split", "ohem_branch": branch,
                    "l2": f"inputs rotate over {N_SETS} distinct batches ({N_SETS * h2d / 1e6:.0f} MB > 126 MB L2)",
                    "collective": "one all_reduce(int64[C*C+2]) after the last step, inside the timed region" if world > 1 else "none",
                    "src_sha16": src_sha16()},
            "clocks": clocks, "e2e": e2e, "e2e_u8_ingest": e2e_u8,
            "gpu_launches": 3 * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "torch_cpu_reference": torch_cpu,
        }
        if reduce_leg:
            line["allreduce_every_step"] = reduce_leg
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_legs(torch, sf, L, dev, args, launch_step, timed, peak, thr):
    """Driver-witnessed numbers for the other BASELINE configs and the comparators, N = 1 only, each a short CUDA-event
    timed loop through the public Python API (or the same C-ABI step for the configs[1] variants)."""
    out = {}
    stream = torch.cuda.current_stream(dev)

    def time_fn(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    n_px2 = B * H * W
    step_bpp = step_bytes_per_px(C, 4, 8, 4, True)
    # configs[1] on the other OHEM branch (same C-ABI step, two rotating batches)
    other = "late" if args.kind == "early" else "early"
    sets = [tuple(t.to(dev) for t in synth_batch(torch, 4321 + i, other)) for i in range(2)]
    for i in range(3):
        launch_step(sets, i)
    ms, kern, _ = timed(sets, 10)
    out["cfg2_" + other] = {"ms": ms, "px_per_s": n_px2 / (ms * 1e-3), "kernel_ms": kern,
                           "frac": step_bpp * n_px2 / (ms * 1e-3) / 1e9 / peak,
                           "note": "late = top-k branch.  One-pass route: the walk's speculative sums are void, the backward re-walks "
                                   "(what a FIRST late step costs; the Python API then learns the branch and switches route)"}
    for i in range(3):
        launch_step(sets, i, onepass=False)
    ms, kern, _ = timed(sets, 10, onepass=False)
    out["cfg2_" + other + "_split_route"] = {"ms": ms, "px_per_s": n_px2 / (ms * 1e-3), "kernel_ms": kern,
                                            "frac": step_bpp * n_px2 / (ms * 1e-3) / 1e9 / peak,
                                            "note": "sf_ce_forward + sf_loss_finalize + sf_ce_backward: the route the Python API takes once "
                                                    "its branch hint says top-k"}
    del sets

    # configs[2] ADE20K plain CE fp32 C=150 B=16 and configs[3] COCO-Stuff bf16 C=171 B=16, fwd+bwd+matrix
    for key, c, dt, esz, ohem in (("cfg3_ade_ce_f32", 150, torch.float32, 4, False), ("cfg4_coco_ce_bf16", 171, torch.bfloat16, 2, False),
                                  ("cfg4_coco_ohem_bf16", 171, torch.bfloat16, 2, True)):
        b, lh, hh = 16, 128, 512
        sets = []
        for i in range(2):
            lo, la = synth_batch(torch, 2345 + i, "early", b=b, c=c, lh=lh, lw=lh, h=hh, w=hh, dtype=dt)
            sets.append((lo.to(dev).requires_grad_(True), la.to(dev)))
        hist = torch.zeros(c, c, dtype=torch.int64, device=dev)
        it = [0]

        def fn():
            x, y = sets[it[0] % 2]
            it[0] += 1
            x.grad = None
            sf.fused_cross_entropy(x, y, None, 255, thr if ohem else None, hist).backward()
        ms = time_fn(fn, 10)
        px = b * hh * hh
        bpp = step_bytes_per_px(c, esz, 8, 4, ohem)
        out[key] = {"ms": ms, "px_per_s": px / (ms * 1e-3), "frac": bpp * px / (ms * 1e-3) / 1e9 / peak, "bytes_per_px": bpp}
        del sets

    # configs[4]: validation sweep, 500 Cityscapes-shaped images, eval-only argmax + confusion matrix (batch 4)
    eb = 4
    sets = [tuple(t.to(dev) for t in synth_batch(torch, 4567 + i, "late", b=eb)) for i in range(4)]
    conf = sf.ConfusionMatrix(C)
    met = sf.Metrics(C, 255, dev)

    def sweep():
        for i in range(500 // eb):
            lo, la = sets[i % len(sets)]
            sf.fused_eval_update(lo, la, met, conf)
    ms = time_fn(sweep, 2, warm=1)
    bpp = C * 4 / 16.0 + 8
    out["cfg5_eval_sweep_500"] = {"ms": ms, "px_per_s": 500 * H * W / (ms * 1e-3), "frac": bpp * 500 * H * W / (ms * 1e-3) / 1e9 / peak,
                                  "bytes_per_px": bpp, "batch": eb, "launches": 2 * (500 // eb)}
    del sets

    # comparator: the reference's torch ops in eager mode on the same B200 (configs[1], includes its host syncs)
    lo, la = (t.to(dev) for t in synth_batch(torch, 1234, args.kind))
    try:
        torch_reference_step(torch, lo, la, C)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            torch_reference_step(torch, lo, la, C)
        torch.cuda.synchronize(dev)
        ms = (time.perf_counter() - t0) / 3 * 1e3
        out["torch_eager_b200"] = {"ms": ms, "px_per_s": n_px2 / (ms * 1e-3), "what": "F.interpolate + OhemCrossEntropy ops + backward + argmax/bincount, torch eager, same GPU"}
    except Exception as e:                                           # e.g. out of memory on a smaller part
        out["torch_eager_b200"] = {"error": str(e)[:120]}
    del lo, la
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
