#!/usr/bin/env python
"""bench.py — stereo stream-seconds rendered per second on B200 (BASELINE.json metric).

Workload (N = 1 and, weak-scaled, every rank at N > 1): BASELINE config 2 — 1024 independent stereo streams,
48 kHz, one shared 256-tap 4-path HRIR set, engine block 256, 10-band parametric EQ ("typical" preset), gain 0.5.
A step is one pass of the fused EQ -> 4-path convolution -> gain kernel over one batch: every stream advances by
FRAMES_PER_STEP frames (192 engine blocks = 1.024 s of audio) in ONE kernel launch.

  value      device-resident throughput: inputs already in HBM, CUDA events on the engine's stream, max over ranks
  e2e        the same work through the host-pointer C-ABI call (ohs_process) with pinned HOST buffers: H2D and D2H
             copies inside the timed region
  roofline   algorithmic HBM bytes of one launch (SURVEY.md §8d formula) / average launch duration, against the
             measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the CPU restatement of the reference (oracle/, the reference itself is Rust and cannot be built here),
             one stream per thread over all host cores, on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import _bootstrap  # noqa: E402

METRIC = "stereo stream-seconds rendered/sec"
UNIT = "stream-s/s"
N_STREAMS = 1024
BLOCK = 256
TAPS = 256
FS = 48000.0
BLOCKS_PER_STEP = int(os.environ.get("OHS_BENCH_BLOCKS", "192"))  # override only to keep ncu replays short
FRAMES_PER_STEP = BLOCK * BLOCKS_PER_STEP  # 49152 frames = 1.024 s
GAIN = 0.5
UNIQUE_STREAMS = 128  # distinct pink-noise streams generated on the host, tiled to N_STREAMS
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
FP32_FMA_TFLOPS_MEASURED = 70.88  # tools/microbench/fp32_pipes.cu on this pool's B200 (profiles/r01_fp32_pipes_microbench.jsonl)


def algorithmic_bytes_per_stream(k_blocks: int, block: int = BLOCK, parts: int = 1) -> int:
    """SURVEY.md §8d / BASELINE.md §3: bytes(K) for one stream through K blocks of one launch."""
    s = 8 * (block + 1)
    return 16 * block * k_blocks + 2 * s * ((parts - 1) + min(k_blocks, parts - 1)) + 16 * block + 320


def algorithmic_flops_per_stream_block(block: int = BLOCK, parts: int = 1) -> float:
    n = 2 * block
    return 4 * (2.5 * n * np.log2(n)) + 4 * parts * (block + 1) * 8 + 4 * block + 2 * block + 2 * 10 * 9 * block


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the render kernel, per launch, from the committed ncu capture of
    this same command (profiles/ncu_render_summary.json); None until a capture exists."""
    p = os.path.join(ROOT, "profiles", "ncu_render_summary.json")
    try:
        d = json.load(open(p))
        if d.get("frames_per_step") == FRAMES_PER_STEP and d.get("n_streams") == N_STREAMS:
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for i, nme in enumerate(names):
                    if r[3 + i].strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def make_workload(pkg):
    S = pkg.signals
    h = S.synthetic_hrir_set(TAPS, 40.0)
    coeffs = np.stack([pkg.eq_design(t, FS, fc, q, g) for (t, fc, q, g) in S.EQ_PRESET_TYPICAL])
    return h, coeffs


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "cfg2: %d independent stereo streams per GPU, 48 kHz, 256-tap HRIRs (4 paths), block 256, 10-band PEQ, gain" % N_STREAMS,
        "n_streams_per_gpu": N_STREAMS, "n_streams_total": N_STREAMS * n_gpus, "block": BLOCK, "taps": TAPS, "partitions": 1,
        "sample_rate": FS, "eq_bands": 10, "frames_per_step": FRAMES_PER_STEP, "blocks_per_launch": BLOCKS_PER_STEP,
        "audio_seconds_per_step_per_stream": FRAMES_PER_STEP / FS,
        "l2": "inputs larger than L2: %.0f MB in + %.0f MB out per step per GPU" % ((N_STREAMS * 2 * FRAMES_PER_STEP * 4 / 1e6,) * 2),
        "parallelism": "streams sharded across GPUs, no data-path collective (NCCL only broadcasts the HRIR spectra at set-up)",
    }


# --------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's algorithm, all host threads
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_run(pkg, steps: int, warmup: int, seconds_per_stream: float, streams_per_thread: int = 4):
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    h, coeffs = make_workload(pkg)
    n_streams = cores * streams_per_thread
    n = int(seconds_per_stream * FS) // BLOCK * BLOCK
    x = pkg.signals.stream_inputs(n_streams, n, unique=min(n_streams, 16))
    for _ in range(max(0, warmup)):
        O.render_batch(x[:cores], BLOCK, h, coeffs, [1] * 10, True, GAIN, n_threads=cores)
    dt = 0.0
    for _ in range(steps):
        # seconds = slowest thread's time inside its per-stream process loops (engine construction, set_ir and the
        # input copy are outside, as they would be for a long-running reference instance)
        dt += O.render_batch(x, BLOCK, h, coeffs, [1] * 10, True, GAIN, n_threads=cores)[1]
    value = steps * n_streams * (n / FS) / dt
    sample = "%d streams x %.3f s of audio per step, %d steps, one stream per thread at a time on %d threads (restated CPU baseline, not rustfft)" % (
        n_streams, n / FS, steps, cores)
    return value, cores, sample, dt / steps * 1e3


def run_reference(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step: 4 streams per host thread x 5.12 s of audio (~0.2 s of CPU work per step on 16 cores): long enough that
    # thread start-up and cold caches do not understate the CPU path, short enough for any --steps the driver picks
    value, cores, sample, ms = cpu_reference_run(pkg, args.steps, min(args.warmup, 1), seconds_per_stream=5.12)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded pink noise, synthetic 256-tap HRIR set)", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_gpu(args, pkg):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    h, coeffs = make_workload(pkg)
    eng = pkg.Engine(N_STREAMS, BLOCK, TAPS, device=local, sample_rate=FS)
    if rank == 0 or world == 1:
        eng.set_hrir_set(h)
        eng.commit_filters()
        eng.sync()
    if world > 1:
        # one HRIR spectra table for the whole job: rank 0 transforms the IRs, NCCL broadcasts the spectra
        pkg.parallel.broadcast_filters(eng, src=0, partitions=1)
    for b in range(10):
        eng.eq_set_band(b, coeffs[b], True)
    eng.set_eq_enable(True)
    eng.set_gain(GAIN)

    x_host = pkg.PinnedBuffer((N_STREAMS, 2, FRAMES_PER_STEP))
    y_host = pkg.PinnedBuffer((N_STREAMS, 2, FRAMES_PER_STEP))
    x_host.array[...] = pkg.signals.stream_inputs(N_STREAMS, FRAMES_PER_STEP, base_seed=1000 + 100000 * rank, unique=UNIQUE_STREAMS)
    d_in = torch.from_numpy(x_host.array).cuda()
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=torch.device("cuda", local))
    torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.process_device(d_in.data_ptr(), d_out.data_ptr(), FRAMES_PER_STEP)
    eng.sync(); torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        eng.process_device(d_in.data_ptr(), d_out.data_ptr(), FRAMES_PER_STEP)
    ev1.record(stream)
    eng.sync(); torch.cuda.synchronize()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    barrier()
    launches = eng.launch_count() - launches0
    ms_per_step = ms_total / args.steps
    stream_seconds_per_step = N_STREAMS * world * FRAMES_PER_STEP / FS
    value = stream_seconds_per_step / (ms_per_step * 1e-3)

    if args.device_only:
        if sampler:
            sampler.stop()  # a lingering nvidia-smi child would keep ncu (which waits for all children) from exiting
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "gpu_launches": int(launches),
                              "note": "device-only run (profiling helper)"}))
        return
    # ---- per-block API (K = 1): one launch per engine block, state round-trips through HBM every launch ----------
    k1_blocks = 64
    for _ in range(8):
        eng.process_device(d_in.data_ptr(), d_out.data_ptr(), BLOCK, FRAMES_PER_STEP)
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(k1_blocks):
        off = i * BLOCK * 4
        eng.process_device(d_in.data_ptr() + off, d_out.data_ptr() + off, BLOCK, FRAMES_PER_STEP)
    e1.record(stream)
    eng.sync()
    k1_ms = max_over_ranks(e0.elapsed_time(e1)) / k1_blocks
    launches += 0  # the K=1 probe is outside the headline timed region

    # ---- end-to-end arm: host buffers through ohs_process ----------------------------------------------------------
    e2e_steps = max(2, min(args.steps, 10))
    eng.process(x_host.array, out=y_host.array)  # warm-up: allocates the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.process(x_host.array, out=y_host.array)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_steps * stream_seconds_per_step / e2e_s
    clocks = sampler.stop() if sampler else None
    checksum = float(np.abs(y_host.array[0, 0, :4096]).sum())

    if rank == 0:
        peak, peak_src = hbm_peak()
        bytes_per_launch = algorithmic_bytes_per_stream(BLOCKS_PER_STEP) * N_STREAMS
        achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
        flops_per_launch = algorithmic_flops_per_stream_block() * BLOCKS_PER_STEP * N_STREAMS
        k1_bytes = algorithmic_bytes_per_stream(1) * N_STREAMS
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded pink noise, %d unique streams tiled to %d; synthetic 256-tap HRIR set; AutoEQ-like 10-band preset)"
                    % (UNIQUE_STREAMS, N_STREAMS),
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.array.nbytes), "d2h_bytes_per_step": int(y_host.array.nbytes),
                    "steps": e2e_steps, "api": "ohs_process (host pointers, pinned), 3-stage H2D/kernel/D2H pipeline"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(), "peak_source": peak_src, "kernel": "ohs::render_kernel<512,7>",
                         "algorithmic_bytes_per_launch": int(bytes_per_launch),
                         "note": "K=%d blocks per launch; formula bytes(K) of SURVEY.md 8d" % BLOCKS_PER_STEP},
            "roofline_fp32": {"achieved_tflops": flops_per_launch / (ms_per_step * 1e-3) / 1e12 / world,
                              "peak_tflops_fma_measured": FP32_FMA_TFLOPS_MEASURED,
                              "frac": flops_per_launch / (ms_per_step * 1e-3) / 1e12 / world / FP32_FMA_TFLOPS_MEASURED,
                              "note": "per GPU; the bit-exact EQ forbids FMA contraction, so its 180 flop/frame cost 180 FP32-pipe slots"},
            "per_block_api": {"blocks_per_launch": 1, "ms_per_launch": k1_ms, "value": N_STREAMS * world * (BLOCK / FS) / (k1_ms * 1e-3),
                              "unit": UNIT, "roofline_achieved_gbs": k1_bytes / (k1_ms * 1e-3) / 1e9,
                              "roofline_frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / peak},
            "checksum": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_reference_run(pkg, 1, 1, seconds_per_stream=10.24)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true", help="only the device-resident arm (used under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    pkg = _bootstrap.load_package()
    if args.impl == "reference":
        run_reference(args, pkg)
    else:
        run_gpu(args, pkg)


if __name__ == "__main__":
    main()
