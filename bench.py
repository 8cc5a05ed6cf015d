#!/usr/bin/env python
"""bench.py -- headline benchmark of the FIR hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun)
    python bench.py --impl reference ...                   (CPU arm: the oracle port on host cores)

One "step" = one pass of ``multirate_FIR(b256).filter`` over a 2^28-sample complex64 stream
(BASELINE.json configs[1]); at N>1 every rank owns its own 2^28-sample segment of one long
stream (weak scaling) and the ranks exchange the (K-1)-sample overlap-save halo over NCCL.
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "scikit-dsp-comm_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

N_SAMPLES = 1 << 28
METRIC = "Msamples/s 256-tap FIR over 2^28 complex64"
UNIT = "Msamples/s"
WORKLOAD = "multirate_FIR.filter(), 256-tap Kaiser lowpass, 2^28 complex64 samples per GPU"
ALGO_BYTES_PER_SAMPLE = 16          # 8 B read + 8 B written; taps amortise to 0


def load_taps():
    return np.load(os.path.join(ROOT, "tests", "golden", "filters.npz"))["b256"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dram_traffic_per_launch():
    """ncu dram__bytes_read+write per launch of the dominant kernel, if a capture was summarised."""
    p = os.path.join(ROOT, "profiles", "fir_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.gpu = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """CPU arm: the oracle's restatement of the reference FIR call (np.convolve in complex128,
    what scipy.signal.lfilter executes for multirate_FIR.filter) on ALL host cores, each core
    filtering its own halo-chunk of the same synthetic workload.  A bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    b = load_taps()
    bl = oracle.FirCpuBaseline(b, chunk=1 << 21)
    for _ in range(max(args.warmup - 1, 0)):
        bl.run_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bl.run_pass()
    dt = time.perf_counter() - t0
    bl.close()
    val = bl.samples_per_pass * args.steps / dt / 1e6
    sample = "%d cores x 2^21 complex64 samples per step (halo-chunked np.convolve, complex128)" % bl.cores
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": bl.cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    # NCCL / torchrun chatter must not pollute stdout: the contract is ONE JSON line there.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    b = load_taps()

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        bl = oracle.FirCpuBaseline(b, chunk=1 << 21)
        t0 = time.perf_counter()
        passes = 0
        while True:
            bl.run_pass()
            passes += 1
            if time.perf_counter() - t0 > 8.0 or passes >= 40:
                break
        dt = time.perf_counter() - t0
        bl.close()
        cpu_baseline = {
            "value": bl.samples_per_pass * passes / dt / 1e6, "unit": UNIT, "cores": bl.cores,
            "kind": "port",
            "sample": "%d passes of %d cores x 2^21 complex64 samples (halo-chunked np.convolve in "
                      "complex128 = the reference's lfilter FIR branch)" % (passes, bl.cores)}

    import torch
    import torch.distributed as dist
    import sk_dsp_comm_b200.multirate_helper as mrh
    from sk_dsp_comm_b200 import _cabi, _engine, hostpipe
    from sk_dsp_comm_b200.sharded import ShardedFIR

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = N_SAMPLES
    torch.manual_seed(100 + rank)
    x = torch.randn(n, dtype=torch.complex64, device=dev)
    fir = mrh.multirate_FIR(b)
    sharded = ShardedFIR(b) if world > 1 else None

    def step():
        if sharded is not None:
            return sharded.filter(x)
        return fir.filter(x)

    for _ in range(args.warmup):
        y = step()
    barrier()

    # ---- timed region: exactly K steps, device timed, max over ranks --------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    stream = torch.cuda.current_stream(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    _cabi.launch_count_reset()
    ev[0].record(stream)
    for i in range(args.steps):
        y = step()
        ev[i + 1].record(stream)
    barrier()
    launches = _cabi.launch_count()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6

    # ---- dominant kernel vs the HBM roofline (per-launch CUDA-event average, this rank) -------
    peak, peak_src = measured_peak()
    k_ms = sum(per_step) / len(per_step)
    achieved = ALGO_BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": dram_traffic_per_launch(),
                "peak_source": peak_src, "kernel": "fir_tc2_kernel<96> (tcgen05 block-Toeplitz GEMM, taps in TMEM, fp16 hi/lo split)" if args.variant == 0 else "variant %d" % args.variant,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * n,
                "kernel_ms": k_ms, "kernel_ms_min": min(per_step)}

    # ---- e2e: public API with HOST buffers (pinned), H2D + D2H inside the timed region --------
    e2e = None
    if not args.no_e2e:
        del y
        xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        xh.copy_(x)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        # warm-up: device staging buffers + TWO pinned output blocks in torch's caching host
        # allocator (the loop below holds the previous result while the next one is produced)
        yh = fir.filter(xh)
        yh2 = fir.filter(xh)
        del yh, yh2
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            yh = fir.filter(xh)                 # returns a host tensor (synchronises the D2H)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d, d2h = hostpipe.bytes_per_call(n, 8, len(b))
        e2e = {"value": world * n * e2e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "multirate_FIR(b).filter(pinned host tensor) -> host tensor"}
        # sanity: host path result == device path result on a window
        if world == 1:
            yd = step()[:4096].cpu()
            assert (yh[:4096] - yd).abs().max().item() <= 1e-6 * yd.abs().max().item()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "taps": 256, "samples_per_gpu": n,
                       "l2": "inputs (2 GiB in + 2 GiB out per step) exceed the 126 MB L2; no flush needed",
                       "parallelism": "overlap-save segments, 1 per GPU, NCCL halo %d B" % (255 * 8)
                       if world > 1 else "single GPU",
                       "fir_variant": args.variant},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="FIR kernel variant (tuning)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.variant:
        from sk_dsp_comm_b200 import _cabi
        _cabi.lib.b200dsp_set_fir_variant(args.variant)
    run_gpu(args)


if __name__ == "__main__":
    main()
