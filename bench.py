#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (BASELINE.json):

    waveform windows/sec, MyCNN5 architecture at [4096, 3, 75000] bf16 per GPU
    (configs[1]; weak scaling: every rank scores its own 4096-window shard), independent
    windows (the semantics of bin/predictStream.py's per-row loop), synthetic N(0,1) data,
    seeded random-init weights (no trained weights exist for this shape).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU)
    python bench.py --impl reference ...                   # the reference's CPU path, same JSON

A "step" = one pass of the hot path over one batch.  `value` is measured with the inputs
resident in HBM; `e2e` goes through the reference-facing predict() call with HOST (pinned)
tensors, host<->device copies inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KIND, C, W, B_PER_GPU = "mycnn5", 3, 75000, 4096
METRIC = "waveform windows/sec (MyCNN5, W=75000)"
UNIT = "windows/s"


def workload(B):
    return {"workload": f"MyCNN5-arch forward, [{B},{C},{W}] bf16 per GPU, mode=independent",
            "arch": KIND, "batch_per_gpu": B, "channels": C, "window": W, "mode": "independent",
            "values": "N(0,1) seed 1234+rank", "weights": "default init, seed 0",
            "l2": "inputs (1.84 GB per GPU) are larger than the 126 MB L2; no flush needed"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks
    line), through NVML in a background thread: the timed region is only tens of milliseconds,
    too short for an `nvidia-smi -lms` child process to produce a sample."""

    def __init__(self, index):
        self.index, self.rows, self.run, self.thread, self.h = index, [], False, None, None
        self.t_begin = self.t_end = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.h is None:
            return
        self.run = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.run = False
        self.thread.join(timeout=2)
        nv = self.nv
        rows = [r for r in self.rows if self.t_begin is not None and self.t_begin <= r[0] <= self.t_end]
        where = "timed region"
        if len(rows) < 3:          # region shorter than a few NVML polls: include the warm-up just before it
            rows, where = self.rows, "warm-up + timed region"
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        reasons = sorted({n for r in rows for bit, n in names.items() if r[2] & bit})
        sm = [r[1] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm, "samples": len(sm),
                "power_w_max": max((r[3] for r in rows), default=None), "window": where, "reasons": reasons}


# ------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's path on the host cores: per-window model(x[i:i+1], age[i:i+1]) under
    no_grad (bin/predictStream.py:154-157), PyTorch-CPU fp32."""

    def __init__(self, n=16, seed=1234):
        import tskd_b200
        from oracle import mycnn_torch as O
        self.ref = O.make_ref(O.stretched(O.ARCH_MYCNN5, C, W), seed=0)
        self.n = n
        self.x = tskd_b200.synth.make_windows(n, C, W, "normal", seed=seed, dtype=torch.bfloat16).float()
        self.ages = tskd_b200.synth.make_ages(n, seed=seed)
        with torch.no_grad():
            for i in range(3):
                self.ref(self.x[i:i + 1], self.ages[i:i + 1])

    def run(self, budget_s, max_windows):
        done, t0 = 0, time.perf_counter()
        with torch.no_grad():
            while done < max_windows and time.perf_counter() - t0 < budget_s:
                i = done % self.n
                self.ref(self.x[i:i + 1], self.ages[i:i + 1])
                done += 1
        dt = time.perf_counter() - t0
        return done / dt, done, dt


def _cpu_worker(idx, barrier, seconds, max_windows, q):
    torch.set_num_threads(1)
    cpu = CpuReference(n=8, seed=1234 + idx)
    barrier.wait()
    _, n, dt = cpu.run(seconds, max_windows)
    q.put((n, dt))


def cpu_reference_all_cores(seconds, max_windows_per_worker=10 ** 9, workers=None):
    """All host cores on the reference's path.  One torch process with N intra-op threads is
    SLOWER than one thread on this per-window call (137 vs 461 windows/s measured on the 64-core
    box: the ops are too small to split), so the cores are used the way the path shards -- one
    single-threaded scorer process per core, each looping predictStream-style over its own windows
    (started together behind a barrier; rate = all windows / the slowest worker's time)."""
    import torch.multiprocessing as mp
    phys = max(1, (os.cpu_count() or 2) // 2)
    workers = workers or min(phys, 64)
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(workers), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(i, barrier, seconds, max_windows_per_worker, q)) for i in range(workers)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=600) for _ in procs]
    for p_ in procs:
        p_.join(timeout=60)
    wins = sum(n for n, _ in res)
    dt = max(t for _, t in res)
    return wins / dt, wins, dt, workers


def run_reference(args, rank):
    if rank != 0:
        return
    per_worker = max(2, int(args.ref_windows))
    wins, secs, workers = 0, 0.0, 0
    for s in range(args.warmup + args.steps):      # each step: every worker scores `per_worker` windows
        if s < args.warmup and s > 0:
            continue                                # one untimed warm-up pass is enough (process start-up dominates)
        r, n, dt, workers = cpu_reference_all_cores(1e9, per_worker)
        if s >= args.warmup:
            wins += n; secs += dt
    v = wins / secs
    sample = (f"{workers} single-thread worker processes x {per_worker} windows per step, per-window "
              f"model(x[i:i+1]) loop (bin/predictStream.py:154-157), torch {torch.__version__} CPU fp32")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload(B_PER_GPU),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": workers, "os_cpu_count": os.cpu_count(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="windows per GPU (default: the BASELINE config)")
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "tensorcore"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-windows", type=int, default=48, help="windows per worker process and step (reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"],
                    help="window dtype: bf16 is the BASELINE workload; f32 (the reference's native dtype) is an extra line")
    ap.add_argument("--tc-splits", type=int, default=3, help="bf16 pieces per fp32 conv1 weight on the tensor cores (3 = fp32-equivalent)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import tskd_b200
    from tskd_b200.dist import broadcast_weights

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    arch = tskd_b200.ARCH_PRESETS[KIND].with_shape(C, W)
    torch.manual_seed(0 if rank == 0 else 1000 + rank)     # only rank 0's weights survive
    model = tskd_b200.B200MyCNN(arch, path=args.path, tc_splits=args.tc_splits).to(dev)
    if world > 1:
        broadcast_weights(model, src=0)                    # the one init-time NCCL collective
    esz = 2 if args.dtype == "bf16" else 4
    x = tskd_b200.synth.make_windows(B, C, W, "normal", seed=1234 + rank, dtype=torch.bfloat16, device=dev)
    if args.dtype == "f32":
        x = x.float()                                      # same bf16-representable values, fp32 storage
    ages = tskd_b200.synth.make_ages(B, seed=1234 + rank, device=dev)
    model.set_profile(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        y = model.predict(x, ages)
    barrier()
    launches_per_step = model.gpu_launches
    path = model.last_path

    if sampler:
        sampler.mark_begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for s in range(args.steps):
        y = model.predict(x, ages)
        ev[s + 1].record()
    barrier()
    if sampler:
        sampler.mark_end()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    # the dominant kernel's own duration: events the library recorded around it on the same
    # stream inside the timed region (last step's value; steps are identical)
    k_ms = model.last_stage_ms(0)
    head_ms = model.last_stage_ms(1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3)
    assert torch.isfinite(y).all()

    # ---- e2e: the public predict() call with pinned HOST tensors, H2D + D2H timed ---------
    Be = min(B, 4096)                                       # e2e sample (pinned host copy); == B for the BASELINE batch
    xh = x[:Be].cpu().pin_memory()
    ah = ages[:Be].cpu().pin_memory()
    model.predict(xh[:256], ah[:256])                       # staging buffers allocated untimed
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        yh = model.predict(xh, ah)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * args.e2e_steps / float(te.item())
    assert torch.allclose(yh, y[:Be].cpu(), rtol=1e-5, atol=1e-6)

    if rank == 0:
        hbm_peak, peak_src = peaks()
        # algorithmic bytes per launch of the dominant kernel: every window's input
        # once (C*W*2 B) + its logit (4 B) + the weights once per launch
        n_w = sum(v.numel() for k, v in model.state_dict().items() if k in tskd_b200.arch.BLOB_KEYS)
        alg_bytes = B * (C * W * esz + 4) + n_w * 4
        achieved = alg_bytes / (k_ms / 1e3) / 1e9 if k_ms and k_ms > 0 else None
        roof = {"bound": "hbm", "kernel": {"generic": "front end (conv1+pool+conv2+pool)", "stream": "fp32 streaming front end (CUDA-core conv1 + tcgen05 projection)"}.get(path, "tcgen05 fused front end"),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None, "peak_source": peak_src,
                "kernel_ms": k_ms, "head_ms": head_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": TRAFFIC.get(path),
                "whole_step_frac": (alg_bytes / (total_ms / args.steps / 1e3) / 1e9) / hbm_peak,
                "compute_note": "co-bound by FP32/MUFU: 21.9 MFLOP + 169k tanh per window (DESIGN.md)"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16 in / f32 math" if args.dtype == "bf16" else "f32", "data": "synthetic",
               "config": dict(workload(B) if args.dtype == "bf16" else dict(workload(B), workload=workload(B)["workload"].replace("bf16", "fp32 (extra line, not the BASELINE dtype)")), path=path, parallelism=f"dp{world} (window shards, no data-path collective)"),
               "roofline": roof, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Be * (C * W * esz + 4),
                       "d2h_bytes_per_step": Be * 4, "steps": args.e2e_steps, "note": "per GPU; pinned host tensors through predict()"},
               "gpu_launches": launches_per_step * args.steps}
        if world == 1 and not args.no_cpu_baseline:
            r, n, dt, workers = cpu_reference_all_cores(args.cpu_seconds)
            out["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": workers,
                                   "os_cpu_count": os.cpu_count(), "kind": "port",
                                   "sample": f"{n} windows in {dt:.1f} s over {workers} single-thread worker processes, per-window "
                                             f"model(x[i:i+1]) loop (bin/predictStream.py:154-157), torch {torch.__version__} CPU fp32"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
# committed ncu captures under profiles/ (None until a capture exists for that path).
TRAFFIC = {"generic": 1.846642e9 + 0.297926e9,      # profiles/r01_generic_frontend_ncu_full.txt (features written to HBM)
           "tensorcore": 1.944850e9 + 0.034090e9,   # profiles/r01_final_fused_ncu_full.txt (halo re-reads + gate partials)
           "stream": 3.801502e9 + 0.034493e9}       # profiles/r01_final_stream_ncu_full.txt (fp32 windows)

if __name__ == "__main__":
    main()
