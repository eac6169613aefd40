#!/usr/bin/env python
"""bench.py -- ASM holograms/s (fwd+adjoint) @1024^2 on B200 (BASELINE.json metric).

A step = one pass of the hot path over one batch of synthetic fields: B forward propagations
(complex64 field -> fp32 |U|^2) plus B adjoint propagations (complex64 cotangent -> complex64), unpadded,
N = 1024, B = 512 per GPU (weak scaling: every rank owns its own 512 samples, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our CUDA path
  python bench.py --impl reference ...                          the reference's CPU torch path (oracle port)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the drop-in Python
API with pinned HOST buffers (H2D of the inputs and D2H of the results inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAMB, PX = 532e-9, 1.5e-6
N_FIELD = 1024
BATCH = 512
Z_MAX = 6e-3            # RBC-scale distances
BYTES_PER_UNIT = 28 * N_FIELD * N_FIELD      # 12 N^2 forward + 16 N^2 adjoint (SURVEY.md 8d / BASELINE.md 3)
DRAM_BYTES_PER_UNIT_MEASURED = 49.2e6        # ncu, round 1 final kernels (profiles/r01_dram_traffic.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML every 5 ms; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thr, self.stop_flag, self.nvml = index, [], None, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.nvml is not None:
            self.stop_flag = True
            self.thr.join(timeout=1.0)
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
            sm = [r[1] for r in rows]
            reasons = sorted({name for r in rows for (name, bit) in self.REASONS if r[2] & bit})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period, inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU torch path (oracle port; /root/reference does not exist on the GPU box)
# ---------------------------------------------------------------------------------------------------
def cpu_units_per_s(sample_units: int, repeats: int, warmup: int):
    import torch
    from oracle import torch_port as tp
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    O = torch.polar(0.5 + 0.5 * torch.rand(sample_units, 1, N_FIELD, N_FIELD, generator=g),
                    2 * torch.pi * torch.rand(sample_units, 1, N_FIELD, N_FIELD, generator=g))
    G = torch.randn(sample_units, 1, N_FIELD, N_FIELD, dtype=torch.complex64, generator=g)
    z = ((0.2 + 0.8 * torch.rand(sample_units, 1, 1, 1, generator=g)) * Z_MAX).float()
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            tp.forward_intensity_cpu(O, LAMB, z, PX, False)
            tp.adjoint_cpu(G, LAMB, z, PX)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 4     # units per step: ~1 s of CPU work per step on 16 cores
    times, threads = cpu_units_per_s(sample, args.steps, args.warmup)
    total = sum(times)
    v = sample * len(times) / total
    line = {"impl": "reference", "metric": "ASM holograms/s (fwd+adjoint) @1024^2", "value": v, "unit": "units/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (complex64 fft / complex128 ifft)",
            "data": "synthetic",
            "config": {"workload": f"configs[2]: ASM fwd+adjoint {N_FIELD}x{N_FIELD} complex64, unpadded; "
                                   f"each step = {sample} units on the host CPU", "n": N_FIELD, "zero_padding": False},
            "cpu_baseline": {"value": v, "unit": "units/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} fwd+adjoint units per step x {len(times)} steps, torch CPU, oracle/torch_port.py"},
            "e2e": {"value": v, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import style_transfer_based_holographic_imaging_b200 as pkg
    from style_transfer_based_holographic_imaging_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()

    B, n = args.batch, N_FIELD
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    O = torch.polar(0.5 + 0.5 * torch.rand(B, 1, n, n, device=dev, generator=g),
                    2 * torch.pi * torch.rand(B, 1, n, n, device=dev, generator=g))          # complex64 field
    G = torch.view_as_complex(torch.randn(B, 1, n, n, 2, device=dev, generator=g))            # complex64 cotangent
    z = ((0.2 + 0.8 * torch.rand(B, 1, 1, 1, device=dev, generator=g)) * Z_MAX).float()
    I = torch.empty(B, 1, n, n, device=dev, dtype=torch.float32)
    A = torch.empty(B, 1, n, n, device=dev, dtype=torch.complex64)

    def step():
        pkg.asm_forward_raw(O, z, LAMB, PX, False, out_mode=L.OUT_INTENSITY, out=I)
        pkg.asm_adjoint_raw(G, z, LAMB, PX, False, out=A)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    n0 = lib.asm_b200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t1 = time.time()
    launches = lib.asm_b200_launch_count() - n0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    units = B * world * args.steps
    value = units / (ms * 1e-3)

    # ---- per-pass share of the step (CUDA events inside the library, separate un-timed pass) ----
    lib.asm_b200_profile(1, None)
    step()
    torch.cuda.synchronize()
    ms3 = (ctypes.c_double * 3)()
    lib.asm_b200_profile(0, ms3)
    pass_ms = {"rows_fwd": ms3[0], "cols": ms3[1], "rows_inv": ms3[2]}

    # ---- end to end through the drop-in API with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        Be = min(B, args.e2e_batch)
        chunk = args.e2e_chunk
        hO = torch.empty(Be, 1, n, n, dtype=torch.complex64).pin_memory()
        hG = torch.empty(Be, 1, n, n, dtype=torch.complex64).pin_memory()
        hI = torch.empty(Be, 1, n, n, dtype=torch.float32).pin_memory()
        hA = torch.empty(Be, 1, n, n, dtype=torch.complex64).pin_memory()
        hO.copy_(O[:Be]); hG.copy_(G[:Be])
        hz = z[:Be].cpu().pin_memory()
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.e2e_streams)]

        def e2e_step():
            # chunks round-robin over the streams: H2D -> ASM (drop-in API) -> D2H; copies of one chunk overlap the
            # opposite-direction copies and the compute of the others (PCIe is full duplex)
            for ci, s0 in enumerate(range(0, Be, chunk)):
                st = streams[ci % len(streams)]
                with torch.cuda.stream(st):
                    o = hO[s0:s0 + chunk].to(dev, non_blocking=True)
                    gg = hG[s0:s0 + chunk].to(dev, non_blocking=True)
                    zz = hz[s0:s0 + chunk].to(dev, non_blocking=True)
                    inten = pkg.asm_forward_raw(o, zz, LAMB, PX, False, out_mode=L.OUT_INTENSITY)
                    adj = pkg.asm_adjoint_raw(gg, zz, LAMB, PX, False)
                    hI[s0:s0 + chunk].copy_(inten, non_blocking=True)
                    hA[s0:s0 + chunk].copy_(adj, non_blocking=True)
            for st in streams:
                st.synchronize()

        e2e_step()
        barrier()
        te = time.perf_counter()
        ke = max(1, min(args.steps, 3))
        for _ in range(ke):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - te], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": Be * world * ke / float(dt.item()), "unit": "units/s",
               "h2d_bytes_per_step": Be * (2 * n * n * 8 + 4), "d2h_bytes_per_step": Be * (n * n * 4 + n * n * 8),
               "batch_per_gpu": Be, "steps": ke, "note": f"pinned host buffers, {len(streams)}-stream pipeline over chunks of {chunk}, host wall clock"}

    if rank == 0:
        peak, peak_src = peaks()
        achieved = value / world * BYTES_PER_UNIT / 1e9          # per-GPU algorithmic GB/s
        line = {"metric": "ASM holograms/s (fwd+adjoint) @1024^2", "value": value, "unit": "units/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "configs[2]: synthetic ASM fwd+adjoint 1024x1024 complex64, unpadded (M = N)",
                           "n": n, "batch_per_gpu": B, "global_batch": B * world, "zero_padding": False,
                           "z_max_m": Z_MAX, "wavelength": LAMB, "pixel_size": PX, "parallelism": f"batch-sharded x{world}, no collective",
                           "l2": f"inputs {B * n * n * 16 / 2**30:.1f} GiB per step >> 126 MB L2 (no flush needed)"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             # DRAM bytes per step = 49.2 MB per unit measured with ncu (dram__bytes_read+write, application
                             # replay, no cache control; profiles/r01_dram_traffic.md) x units per step; algorithmic is 29.4 MB
                             "traffic": DRAM_BYTES_PER_UNIT_MEASURED * B, "traffic_source": "profiles/r01_dram_traffic.md (ncu, per unit) x batch",
                             "peak_source": peak_src,
                             "kernel": "whole fwd+adjoint step = 2 calls x (setup + 57 chunks x {k32_rows_fwd_bulk, k32_cols_pipe, k32_rows_inv_bulk}); "
                                       "dominant kernel k32_cols_pipe = 45 % of library time (profiles/r01_launches_bench_summary.md); "
                                       "algorithmic 28*N^2 B per unit",
                             "pass_ms_profiled": pass_ms},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
        if not args.no_cpu and world >= 1:
            times, threads = cpu_units_per_s(4, 3, 1)
            line["cpu_baseline"] = {"value": 4 * len(times) / sum(times), "unit": "units/s", "cores": threads, "kind": "port",
                                    "sample": "4 fwd+adjoint units x 3 repeats after 1 warm-up, torch CPU, oracle/torch_port.py"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--e2e-batch", type=int, default=128)
    ap.add_argument("--e2e-chunk", type=int, default=8)
    ap.add_argument("--e2e-streams", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
