#!/usr/bin/env python
"""bench.py -- labelled pixels/s of the fused upsample + OHEM-CE + confusion-matrix fwd+bwd path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU port of the reference path, host cores)

One step = one pass of the hot path over one Cityscapes-shaped batch per GPU (BASELINE.json configs[1]):
logits 8x19x256x512 fp32 -> labels 8x1024x2048 int64, OhemCrossEntropy forward + backward and the
GT x pred confusion matrix, i.e. sf_ce_forward -> sf_loss_finalize -> sf_ce_backward through the C ABI.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "labelled pixels/s (fused upsample+OHEM-CE+mIoU fwd+bwd), % of HBM roofline"
UNIT = "px/s"
B, C, LH, LW, H, W = 8, 19, 256, 512, 1024, 2048          # per-GPU batch of configs[1]
N_SETS = 4                                                  # rotating input sets: 4 x 214 MB >> 126 MB L2
WORKLOAD = "cityscapes_ohem_fwd_bwd_hist: logits 8x19x256x512 fp32 -> labels 8x1024x2048 int64 per GPU"


def algorithmic_bytes_per_px(kernel: str) -> float:
    """e = 4 B logits at 1/16 density, L = 8 B labels; per-pixel loss is 4 B."""
    ce16 = C * 4 / 16.0
    return {"fwd": ce16 + 8 + 4,                 # logits read, label read, loss_px write
            "bwd": 2 * ce16 + 8 + 4,             # logits re-read + grad write, label read, loss_px read
            "select": 4.0,                       # loss_px read (top-k branch only)
            "step": 3 * ce16 + 2 * 8 + 12}[kernel]


def synth_batch(torch, seed, device, kind="early"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    coarse = torch.randint(0, C, (B, H // 32, W // 32), generator=g)
    coarse[torch.rand(coarse.shape, generator=g) < 0.1] = 255
    lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    if kind == "early":
        low = 3.0 * torch.randn((B, C, LH, LW), generator=g)
    else:
        l2 = coarse.repeat_interleave(8, 1).repeat_interleave(8, 2)
        l2 = torch.where(l2 == 255, torch.zeros_like(l2), l2)
        flip = torch.rand(l2.shape, generator=g) < 0.02
        l2 = torch.where(flip, torch.randint(0, C, l2.shape, generator=g), l2)
        low = torch.randn((B, C, LH, LW), generator=g) + 12.0 * torch.nn.functional.one_hot(l2, C).permute(0, 3, 1, 2).float()
    return low.contiguous(), lab


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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python/torch and
    /root/reference does not exist on the GPU box, so this is the C port (oracle/, kind "port") with OpenMP over
    every host core; each step is one full configs[1] batch (8 images), a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import oracle as orc
    orc.build()
    low, lab = synth_batch(torch, 1234, "cpu")
    low, lab = low.numpy(), lab.numpy()
    for _ in range(args.warmup):
        cpu_port_step(orc, np, low, lab)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(orc, np, low, lab)
    dt = (time.perf_counter() - t0) / args.steps
    v = B * H * W / dt
    cores = orc.num_threads()
    sample = f"{args.steps} step(s) of one full 8-image batch (16.8 Mpx) per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "where": "host CPU, OpenMP"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kind", default="early", choices=["early", "late"],
                    help="logit distribution: early -> OHEM threshold branch, late -> top-k (radix select) branch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import segmentation_factory_b200 as sf
    from segmentation_factory_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    n_px = B * H * W

    # ---- inputs: N_SETS distinct batches resident in HBM, pinned host copies for the end-to-end leg ----
    host_sets = [synth_batch(torch, 1234 + 97 * rank + i, "cpu", args.kind) for i in range(N_SETS)]
    dev_sets = [(lo.to(dev), la.to(dev)) for lo, la in host_sets]
    ws_bytes = L.sf_workspace_bytes()
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    loss_px = torch.empty(n_px, dtype=torch.float32, device=dev)
    lse_px = torch.empty(n_px, dtype=torch.float32, device=dev)
    grad = torch.empty((B, C, LH, LW), dtype=torch.float32, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    hist = torch.zeros(C * C + 2, dtype=torch.int64, device=dev)     # matrix || kept || valid, one all-reduce
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream

    def launch_step(i, evs=None):
        lo, la = dev_sets[i % N_SETS]
        if evs:
            evs[0].record(stream)
        rc = L.sf_ce_forward(lo.data_ptr(), 0, la.data_ptr(), 0, B, C, LH, LW, H, W, 255, None, ctypes.c_float(thr),
                             loss_px.data_ptr(), lse_px.data_ptr(), hist.data_ptr(), ws.data_ptr(), ws_bytes, st)
        if evs:
            evs[1].record(stream)
        rc |= L.sf_loss_finalize(loss_px.data_ptr(), n_px, ctypes.c_float(thr), ws.data_ptr(), ws_bytes, out.data_ptr(), st)
        if evs:
            evs[2].record(stream)
        work = None
        if world > 1:
            # the path's one exchange: int64 matrix (+ counters) SUM over NVLink.  The matrix is final after the forward,
            # so the collective is issued here and runs on NCCL's stream underneath the backward kernel.
            work = dist.all_reduce(hist, async_op=True)
        rc |= L.sf_ce_backward(lo.data_ptr(), 0, la.data_ptr(), 0, B, C, LH, LW, H, W, 255, None, loss_px.data_ptr(),
                               lse_px.data_ptr(), None, ws.data_ptr(), grad.data_ptr(), st)
        if evs:
            evs[3].record(stream)
        if rc:
            raise RuntimeError(f"kernel launch failed rc={rc}")
        if work is not None:
            work.wait()                                      # stream-level wait, no host sync

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        launch_step(i)
    barrier()

    # ---- timed region: exactly K steps, device-timed, per-kernel events inside ----
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(torch, local) if rank == 0 else None
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        launch_step(i, evs[i])
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * n_px / (ms_step * 1e-3)

    k_fwd = sorted(e[0].elapsed_time(e[1]) for e in evs)
    k_sel = sorted(e[1].elapsed_time(e[2]) for e in evs)
    k_bwd = sorted(e[2].elapsed_time(e[3]) for e in evs)
    avg = lambda v: sum(v) / len(v)
    kern = {"fwd": avg(k_fwd), "select": avg(k_sel), "bwd": avg(k_bwd)}
    dom = "fwd" if kern["fwd"] >= kern["bwd"] else "bwd"

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = algorithmic_bytes_per_px(dom) * n_px / (kern[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")     # dram read+write bytes per launch from ncu --set full
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom)
    roofline = {"bound": "hbm", "kernel": f"sf::{dom}_fast_kernel<float,19,4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_hbm3e_spec_8000GBs": achieved / 8000.0,
                "kernel_ms": kern, "step_frac": algorithmic_bytes_per_px("step") * n_px / (ms_step * 1e-3) / 1e9 / peak,
                "note": "at C=19 and 1/16 logit density the path is issue/MUFU-bound, not HBM-bound (DESIGN.md)"}

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
                "steps": n, "ms_per_step": dt.item() * 1e3}, last

    e2e, loss_i64 = e2e_leg(False)
    e2e_u8, loss_u8 = e2e_leg(True)
    e2e_u8["labels"] = "raw uint8 ids over PCIe + sf_label_lut on the device"
    e2e_u8["same_loss_as_int64_leg"] = bool(loss_i64 == loss_u8)
    h2d = host_sets[0][0].numel() * 4 + host_sets[0][1].numel() * 8

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the box's host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        lo, la = host_sets[0][0].numpy(), host_sets[0][1].numpy()
        cpu_port_step(orc, np, lo[:1], la[:1])
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 20):
            cpu_port_step(orc, np, lo, la)
            reps += 1
        cdt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_px / cdt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
               "sample": f"{reps} full 8-image batches (16.8 Mpx each), {cdt:.2f} s per batch"}

    if rank == 0:
        state = np.frombuffer(ws[:64].cpu().numpy().tobytes(), dtype=np.int32)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "logits": args.kind, "ohem_branch": int(state[4]), "labels": "int64, 10% ignore=255",
                       "l2": f"inputs rotate over {N_SETS} distinct batches ({N_SETS * h2d / 1e6:.0f} MB > 126 MB L2)",
                       "parallelism": f"dp{world}", "collective": "all_reduce(int64[C*C+2]) per step, overlapped with the backward" if world > 1 else "none"},
            "clocks": clocks, "e2e": e2e, "e2e_u8_ingest": e2e_u8, "gpu_launches": 3 * args.steps, "roofline": roofline, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
