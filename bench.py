#!/usr/bin/env python
"""bench.py -- headline benchmark of the FIR hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun)
    python bench.py --impl reference ...                   (CPU arm: the oracle port on host cores)

One "step" = one pass of ``multirate_FIR(b256).filter`` over a 2^28-sample complex64 stream
(BASELINE.json configs[1]); at N>1 every rank owns its own segment of one long stream and gets the
(K-1)-sample overlap-save halo of its left neighbour -- by default read by the FIR kernel itself from
the neighbour's memory over NVLink (one launch per step), ``--halo nccl`` for the send/recv variant.
Weak scaling by default (2^28 per rank); ``--total 2147483648`` runs BASELINE configs[4] as written
(2^31 samples split over the ranks, "scaling": "strong").
Prints ONE JSON line on rank 0 (see the task contract for the keys); at N=1 the line also carries a
``secondary`` list with the other BASELINE configs (cfg3 up/dn, cfg4 SOS) timed in the same process.
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
    bl = oracle.FirCpuBaseline(b, chunk=1 << 21)      # kind "scipy": scipy.signal.lfilter(b,[1],x), the literal reference call
    for _ in range(max(args.warmup - 1, 0)):
        bl.run_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bl.run_pass()
    dt = time.perf_counter() - t0
    bl.close()
    val = bl.samples_per_pass * args.steps / dt / 1e6
    how = ("scipy.signal.lfilter(b,[1],x), the call at multirate_helper.py:108" if bl.kind == "scipy"
           else "np.convolve restatement of lfilter's FIR branch")
    sample = "%d cores x 2^21 complex64 samples per step, halo-chunked (%s), complex128" % (bl.cores, how)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": bl.cores,
                         "kind": bl.kind, "sample": sample},
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
            "kind": bl.kind,
            "sample": "%d passes of %d cores x 2^21 complex64 samples, halo-chunked (%s, complex128)"
                      % (passes, bl.cores, "scipy.signal.lfilter(b,[1],x) = multirate_helper.py:108"
                         if bl.kind == "scipy" else "np.convolve = lfilter's FIR branch")}

    import torch
    import torch.distributed as dist
    import sk_dsp_comm_b200.multirate_helper as mrh
    from sk_dsp_comm_b200 import _cabi, _engine, hostpipe
    from sk_dsp_comm_b200.sharded import ShardedFIR

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = hostpipe.bind_to_gpu_numa(local_rank)      # pinned staging memory + copy threads next to this GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    strong = args.total > 0
    n = (args.total // world) if strong else N_SAMPLES
    if strong and args.total % world:
        raise SystemExit("--total must be a multiple of the number of ranks")
    fir = mrh.multirate_FIR(b)
    sharded = ShardedFIR(b) if world > 1 else None
    halo_mode = "none"
    x = None
    if world > 1:
        halo_mode = args.halo
        if halo_mode == "peer":
            try:
                x = sharded.attach_symmetric(n, torch.complex64, dev)      # segment lives in symmetric memory
            except Exception as e:                                         # noqa: BLE001
                sys.stderr.write("bench: symmetric memory unavailable (%s); NCCL halo instead\n" % (e,))
                halo_mode = "nccl"
        # every rank must agree on the mode
        flag = torch.tensor([1 if halo_mode == "peer" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and halo_mode == "peer":
            halo_mode, x = "nccl", None
    if x is None:
        x = torch.empty(n, dtype=torch.complex64, device=dev)
    torch.manual_seed(100 + rank)
    xr = torch.view_as_real(x)
    chunk = 1 << 26
    for c0 in range(0, n, chunk):                     # fill in place (a 2^30-sample segment is 8 GB)
        xr[c0:min(c0 + chunk, n)].normal_()
    barrier()                                         # all segments written before any neighbour reads a halo

    def step():
        if sharded is None:
            return fir.filter(x)
        if halo_mode == "peer":
            return sharded.filter_peer()              # ONE launch: the kernel reads the halo over NVLink
        return sharded.filter(x)                      # NCCL send/recv + interior / head launches

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
                "traffic_source": "profiles/ ncu capture of this kernel at 2^28 samples (not re-measured in this run)",
                "peak_source": peak_src, "kernel": "fir_tc2_kernel<96> (tcgen05 block-Toeplitz GEMM, taps in TMEM, fp16 hi/lo split)" if args.variant == 0 else "variant %d" % args.variant,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * n,
                "kernel_ms": k_ms, "kernel_ms_min": min(per_step)}
    if args.variant == 0 and len(b) == 256:
        # what the tensor cores execute for it: 80 UTCHMMA (128 x 96 x 16) per 6144-sample tile -- three fp16 hi/lo
        # products (a fourth against zeroed rows) on a 320-wide Toeplitz band for 256 taps -- against the measured
        # cuBLAS bf16 rates: the kernel sits at the chip's power-limited tensor rate while streaming at `achieved`
        tiles = (n + 6143) // 6144
        tflops = 80 * 128 * 96 * 16 * 2 * tiles / (k_ms * 1e-3) / 1e12
        roofline["tensor_executed"] = {"tflops": tflops, "unit": "TFLOP/s fp16 dense (executed, not algorithmic)"}
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roofline["tensor_executed"].update({
                "peak_burst": pk["bf16_tflops"], "frac_of_burst": tflops / pk["bf16_tflops"],
                "peak_sustained": pk.get("bf16_tflops_sustained"),
                "frac_of_sustained": (tflops / pk["bf16_tflops_sustained"]) if pk.get("bf16_tflops_sustained") else None})
        except Exception:
            pass

    # ---- parity inside the measured configuration: windows at the shard boundaries vs the oracle -----------
    parity = None
    if not args.no_parity:
        import oracle
        W = 1 << 16
        k1 = len(b) - 1
        errs = []
        # head window: first W outputs of this segment (they depend on the left neighbour's tail)
        hist = None
        if rank > 0:
            if halo_mode == "peer":
                t_, hdl_, n_ = sharded._symm
                hist = hdl_.get_buffer(rank - 1, (n_,), t_.dtype)[n_ - k1:].cpu().numpy().astype(np.complex128)
            else:
                halo, works = sharded.exchange_halo(x)
                for w_ in works:
                    w_.wait()
                torch.cuda.synchronize()
                hist = halo.cpu().numpy().astype(np.complex128)
        elif world > 1 and halo_mode == "nccl":
            halo, works = sharded.exchange_halo(x)
            for w_ in works:
                w_.wait()
        ref = oracle.fir_filter(b, x[:W].cpu().numpy().astype(np.complex128), hist=hist, backend="c")
        errs.append(float(np.abs(y[:W].cpu().numpy() - ref).max() / np.abs(ref).max()))
        # tail window: last W outputs (what the right neighbour's halo continues)
        xs = x[n - W - k1:].cpu().numpy().astype(np.complex128)
        ref = oracle.fir_filter(b, xs[k1:], hist=xs[:k1], backend="c")
        errs.append(float(np.abs(y[n - W:].cpu().numpy() - ref).max() / np.abs(ref).max()))
        e = torch.tensor([max(errs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
        parity = {"boundary_max_err": float(e.item()), "tolerance": 1e-6,
                  "what": "max over ranks of |y - oracle| / max|oracle| on the first and last 2^16 outputs of every "
                          "segment (oracle = float64 FIR of the same samples, left neighbour's tail as history)"}
        assert parity["boundary_max_err"] <= 1e-6, parity

    # ---- the other BASELINE configs, same process (N=1 only) ----------------------------------------------
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = run_secondary(torch, dev, _engine, b, peak)

    # ---- e2e: public API with HOST buffers (pinned), H2D + D2H inside the timed region --------
    e2e = None
    if not args.no_e2e and n <= (1 << 28):
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
               "api": "multirate_FIR(b).filter(pinned host tensor) -> host tensor",
               "numa": numa}
        # sanity: host path result == device path result on a window
        if world == 1:
            yd = step()[:4096].cpu()
            assert (yh[:4096] - yd).abs().max().item() <= 1e-6 * yd.abs().max().item()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if not strong else
                       "multirate_FIR.filter(), 256-tap Kaiser lowpass, %d complex64 samples split over %d GPUs "
                       "(overlap-save segments)" % (args.total, world),
                       "taps": 256, "samples_per_gpu": n,
                       "l2": "inputs (>= 2 GiB in + 2 GiB out per step and GPU) exceed the 126 MB L2; no flush needed",
                       "parallelism": ("overlap-save segments, 1 per GPU; halo %d B per boundary, %s"
                                       % (255 * 8, "read by the FIR kernel from the neighbour's memory over NVLink "
                                          "(one launch per step)" if halo_mode == "peer" else
                                          "NCCL send/recv overlapped with the interior launch"))
                       if world > 1 else "single GPU",
                       "fir_variant": args.variant},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "parity": parity,
            "secondary": secondary, "gpu_launches": launches, "clocks": clocks,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def run_secondary(torch, dev, _engine, b, peak):
    """cfg3 (multirate_FIR.up(4)/.dn(4), 2^26 float32) and cfg4 (multirate_IIR 6-SOS .filter, 2^28 float32) with
    the same timing rules as the headline (>= 3 warm-ups, CUDA events on the launching stream, inputs larger than
    L2).  Algorithmic bytes per input sample from SURVEY.md 8d: up(4) 20 B, dn(4) 5 B, SOS 8 B."""
    sos6 = np.load(os.path.join(ROOT, "tests", "golden", "filters.npz"))["sos6"]
    plan = _engine.FirPlan(b)
    splan = _engine.SosPlan(sos6)
    out = []

    def timed(name, kernel, fn, n_in, bytes_per_sample, reps=10):
        for _ in range(3):
            y = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        del y
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_per_sample * n_in / (ms * 1e-3) / 1e9
        out.append({"config": name, "kernel": kernel, "ms": ms, "Msamples_per_s_input": n_in / ms / 1e3,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                 "algorithmic_bytes_per_input_sample": bytes_per_sample}})

    n3 = 1 << 26
    x3 = torch.randn(n3, dtype=torch.float32, device=dev)
    timed("cfg3 multirate_FIR.up(4), 256 taps, 2^26 float32", "fir_tc_real_kernel<up,4> (tcgen05, 4 phase filters in TMEM)",
          lambda: _engine.fir_up(plan, x3, 4), n3, 20)
    timed("cfg3 multirate_FIR.dn(4), 256 taps, 2^26 float32", "fir_tc_real_kernel<dn,4> (tcgen05, 4 phase streams)",
          lambda: _engine.fir_dn(plan, x3, 4), n3, 5)
    del x3
    n4 = 1 << 28
    x4 = torch.randn(n4, dtype=torch.float32, device=dev)
    timed("cfg4 multirate_IIR.filter(), 6-SOS elliptic bandpass, 2^28 float32",
          "sos_tc_kernel<12> (single pass: Toeplitz + carry + correction GEMMs on tcgen05, float64 chain between tiles)",
          lambda: _engine.sos_filter(splan, x4), n4, 8)
    timed("cfg4 .dn(4) (multirate_IIR.dn), 2^28 float32", "sos_tc_kernel<12>, decimating stores",
          lambda: _engine.sos_filter(splan, x4, M=4), n4, 5)
    del x4
    x5 = torch.randn(1 << 26, dtype=torch.float32, device=dev)
    timed("cfg4 .up(4) (multirate_IIR.up), 2^26 float32 in", "sos_tc_kernel<12>, zero-stuffed stream staged once",
          lambda: _engine.sos_filter(splan, x5, L=4), 1 << 26, 20)
    del x5
    # the reference's default rate factors L_change = M_change = 12 (multirate_helper.py:112,121), both stream types
    x8 = torch.randn(1 << 27, dtype=torch.float32, device=dev)
    timed("multirate_FIR.dn(12), 256 taps, 2^27 float32", "fir_tc_real_kernel<filter> with decimating stores",
          lambda: _engine.fir_dn(plan, x8, 12), 1 << 27, 4 * 13 / 12)
    nu = (1 << 27) // 12 // 4 * 4
    timed("multirate_FIR.up(12), 256 taps, %d float32" % nu, "fir_up_short_kernel (CUDA cores, 22 taps per phase)",
          lambda: _engine.fir_up(plan, x8[:nu], 12), nu, 4 * 13)
    del x8
    x9 = torch.randn(1 << 26, dtype=torch.complex64, device=dev)
    timed("multirate_FIR.dn(12), 256 taps, 2^26 complex64", "fir_tc2_kernel<96, decimating stores>",
          lambda: _engine.fir_dn(plan, x9, 12), 1 << 26, 8 * 13 / 12)
    nu = (1 << 26) // 12 // 4 * 4
    timed("multirate_FIR.up(12), 256 taps, %d complex64" % nu, "fir_up_short_kernel (CUDA cores, 22 taps per phase)",
          lambda: _engine.fir_up(plan, x9[:nu], 12), nu, 8 * 13)
    timed("multirate_FIR.up(4), 256 taps, 2^23 complex64", "fir_tc2_kernel<96, zero-stuffed input>",
          lambda: _engine.fir_up(plan, x9[:1 << 23], 4), 1 << 23, 8 * 5)
    del x9
    # filters beyond the 256 taps of the tensor-core kernel: overlap-save FFT (sigsys.os_filter's method, sigsys.py:482)
    k = np.arange(1024) - 511.5
    plan_long = _engine.FirPlan(np.sinc(0.2 * k) * np.kaiser(1024, 8.0) * 0.2)
    x6 = torch.randn(1 << 26, dtype=torch.complex64, device=dev)
    timed("multirate_FIR.filter(), 1024 taps, 2^26 complex64", "fir_fft_os_kernel<complex> (4096-point FFT in shared memory, overlap-save)",
          lambda: _engine.fir_filter(plan_long, x6), 1 << 26, 16)
    del x6
    x7 = torch.randn(1 << 27, dtype=torch.float32, device=dev)
    timed("multirate_FIR.filter(), 1024 taps, 2^27 float32", "fir_fft_os_kernel<real> (two real frames per complex transform)",
          lambda: _engine.fir_filter(plan_long, x7), 1 << 27, 8)
    return out


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3 / cfg4 measurements (N=1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the shard-boundary parity windows")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N>1: how a segment gets its left neighbour's tail (default: peer-memory read by the kernel)")
    ap.add_argument("--total", type=int, default=0,
                    help="strong scaling: total samples split over the ranks (2147483648 = BASELINE configs[4])")
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
