#!/usr/bin/env python
"""Benchmark of the mastering hot path (BASELINE.json metric): stereo 44.1 kHz samples/sec as
x real-time, full pipeline (stages.main with need_default) on BASELINE config 2 -- a 3-minute
44.1 kHz stereo synthetic track mastered against a 3-minute reference -- one track per GPU per
step (tracks shard one-per-GPU, no data-path collective; NCCL only gathers the timings).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in HBM,
`--lanes` tracks in flight; the one-track-at-a-time figure is in `config`), `e2e` = the same job through
the C ABI's host-buffer batch entry (mgb_pipeline_submit/wait: pinned host float32 in and out, H2D +
four stages + D2H inside the timed region; `e2e.pcm16` with int16 buffers, `e2e.single_call` =
mgb_process_host).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 44100
METRIC = "stereo 44.1kHz samples/sec (x real-time)"
UNIT = "x real-time"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=180.0, help="track length (config 2: 180)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-sample-seconds", type=float, default=30.0)
    ap.add_argument("--lanes", type=int, default=3, help="tracks in flight per GPU for the device-resident number")
    ap.add_argument("--opt", action="append", default=[], help="library switch name=value (A/B measurements)")
    return ap.parse_args()


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import port
    return port


# --------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 100 ms in a child process that is started EARLY (its start-up can take
    seconds on a cold box); the samples that count are the ones time-stamped inside the timed window."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except (OSError, FileNotFoundError):
            self.proc = None

    def begin(self, wait_s: float = 15.0):
        """Start of the timed window; waits (bounded) until the poller has produced its first line."""
        if self.proc is not None:
            deadline = time.time() + wait_s
            while time.time() < deadline and os.path.getsize(self.path) == 0 and self.proc.poll() is None:
                time.sleep(0.05)
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    @staticmethod
    def _stamp(text: str):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        if self.t_end is None:
            self.end()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 10:
                    continue
                try:
                    rows.append((self._stamp(parts[0]), float(parts[2]), float(parts[3]),
                                 [n for n, v in zip(names, parts[6:10]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
        os.unlink(self.path)
        lo, hi = (self.t_begin or 0.0) - 0.05, (self.t_end or time.time()) + 0.15
        inside = [r for r in rows if r[0] is not None and lo <= r[0] <= hi]
        scope = "timed window"
        if not inside:  # clock skew or an unparsable stamp: fall back to everything the poller saw
            inside, scope = rows, "whole run"
        sm = sorted(r[1] for r in inside)
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": inside[-1][2] if inside else None,
                "samples": len(sm), "scope": scope, "reasons": reasons}


# --------------------------------------------------------------------------------------------------
# the reference arm: the reference's CPU algorithm (numpy/scipy, oracle/port.py) on all host cores
# --------------------------------------------------------------------------------------------------
_WORKER_INPUTS = {}


def _reference_inputs(seconds: float):
    """Per-process cache: synthesising the noise is not part of the measured hot path."""
    if seconds not in _WORKER_INPUTS:
        port = oracle()
        import numpy as np
        n = int(SAMPLE_RATE * seconds)
        seed = os.getpid() % 1000
        _WORKER_INPUTS[seconds] = (port.synth_target(n, seed).astype(np.float64),
                                   port.synth_reference(n, 1000 + seed).astype(np.float64))
    return _WORKER_INPUTS[seconds]


def _reference_worker(seconds):
    os.environ["OMP_NUM_THREADS"] = "1"
    port = oracle()
    t, r = _reference_inputs(seconds)
    t0 = time.perf_counter()
    port.main(t, r, port.OracleConfig(), True, False, False)
    return time.perf_counter() - t0


def run_reference(args) -> dict:
    """Reference arm.  /root/reference is pure Python over numpy/scipy and cannot travel to the GPU
    box, so this times oracle/port.py -- the same numpy/scipy native kernels in the same order --
    with one worker process per host core, each mastering one `sample`-second track per step."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    cores = os.cpu_count() or 1
    sample = args.reference_sample_seconds
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(max(1, args.warmup)):  # also fills every worker's input cache
            pool.map(_reference_worker, [sample] * cores, chunksize=1)
        t0 = time.perf_counter()
        for step in range(args.steps):
            pool.map(_reference_worker, [sample] * cores, chunksize=1)
        elapsed = time.perf_counter() - t0
    frames = args.steps * cores * int(SAMPLE_RATE * sample)
    value = frames / elapsed / SAMPLE_RATE
    desc = f"{cores} worker processes x one {sample:.0f}-s track (config-2 recipe) per step, float64 numpy/scipy"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "samples_per_sec": frames / elapsed,
        "config": {"workload": "config-2 recipe, bounded sample: full pipeline stages.main(need_default) on "
                               f"{sample:.0f}-s 44.1 kHz stereo tracks", "tracks_per_step": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
ALGORITHMIC_BYTES_PER_FRAME = {
    # compulsory HBM bytes per stereo frame of the signal a launch covers
    "analyze_kernel": 8,      # read L,R once
    "convolve_kernel": 16,    # read L,R, write result L,R
    "clip_sumsq_kernel": 4,   # re-read the mid plane
    "limiter_kernel": 16,     # read result, write final
}


def run_b200(args) -> dict:
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)  # (polls from now on; only samples inside the timed window are reported)
    if rank == 0:
        sampler.start()
    bound_cores = None
    if world > 1:
        # several ranks share the host: each stays on the cores (and memory) next to its own GPU
        from matchering_b200.sharding import bind_host_thread_near_gpu
        bound_cores = bind_host_thread_near_gpu(local_rank)
    if world > 1:
        # keep stdout for the one JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)

    port = oracle()  # synthetic-input recipes + the cpu_baseline leg only
    import matchering_b200 as mg
    from matchering_b200 import _native
    from matchering_b200.engine import TrackSession, get_plan

    cfg = mg.Config()
    plan = get_plan(cfg, device)
    lib = plan.lib
    for item in args.opt:
        name, value = item.split("=")
        _native.check(lib, lib.mgb_set_option(name.encode(), int(value)))
    n = int(SAMPLE_RATE * args.seconds)
    stream = torch.cuda.current_stream(device)
    sptr = C.c_void_p(stream.cuda_stream)

    # three distinct tracks per rank, rotated, so no step finds its inputs in the 126 MB L2
    n_sets = 3
    host_t, host_r, dev_t, dev_r = [], [], [], []
    for k in range(n_sets):
        seed = rank * 16 + k
        t = torch.from_numpy(port.synth_target(n, seed)).pin_memory()
        r = torch.from_numpy(port.synth_reference(n, 1000 + seed)).pin_memory()
        host_t.append(t)
        host_r.append(r)
        dev_t.append(t.to(device))
        dev_r.append(r.to(device))
    # `--lanes` tracks in flight on as many streams: one track's small latency-bound kernels (FIR design: four
    # CTAs) overlap the other's streaming kernels
    n_lanes = max(1, args.lanes)
    sessions = [TrackSession(plan, n, n) for _ in range(n_lanes)]
    lane_streams = [torch.cuda.Stream(device=device) for _ in range(n_lanes)]
    lane_out = [torch.empty((n, 2), dtype=torch.float32, device=device) for _ in range(n_lanes)]
    session = sessions[0]
    out_dev = lane_out[0]
    out_host = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    stage_t = torch.empty((n, 2), dtype=torch.float32, device=device)
    stage_r = torch.empty((n, 2), dtype=torch.float32, device=device)
    p_plan, p_layout = C.byref(plan.struct), C.byref(session.layout)
    ws, st = session.workspace.data_ptr(), session.state.data_ptr()

    def step_on(k, sess, out, stream_ptr):
        t, r = dev_t[k % n_sets], dev_r[k % n_sets]
        lay, w, s_ = C.byref(sess.layout), sess.workspace.data_ptr(), sess.state.data_ptr()
        _native.check(lib, lib.mgb_match_levels(p_plan, lay, t.data_ptr(), r.data_ptr(), w, s_, stream_ptr))
        _native.check(lib, lib.mgb_match_frequencies(p_plan, lay, t.data_ptr(), sess.result.data_ptr(), None, w, s_, stream_ptr))
        _native.check(lib, lib.mgb_correct_levels(p_plan, lay, w, s_, stream_ptr))
        _native.check(lib, lib.mgb_finalize(p_plan, lay, sess.result.data_ptr(), out.data_ptr(), None, None, w, s_, stream_ptr))

    def step_device(k):  # one stream, one track at a time (profiling pass, single-track latency)
        step_on(k, session, out_dev, sptr)

    def timed_lanes(steps, warmup):
        """K tracks, alternating over the lanes; CUDA events on the main stream bracket all of them."""
        for k in range(warmup):
            step_on(k, sessions[k % n_lanes], lane_out[k % n_lanes], C.c_void_p(lane_streams[k % n_lanes].cuda_stream))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.mgb_launch_count()
        e0.record(stream)
        for ls in lane_streams:
            ls.wait_event(e0)
        for k in range(steps):
            lane = k % n_lanes
            step_on(warmup + k, sessions[lane], lane_out[lane], C.c_void_p(lane_streams[lane].cuda_stream))
        for ls in lane_streams:
            stream.wait_stream(ls)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), lib.mgb_launch_count() - launches0

    def step_host(k):
        t, r = host_t[k % n_sets], host_r[k % n_sets]
        _native.check(lib, lib.mgb_process_host(p_plan, p_layout, t.data_ptr(), r.data_ptr(), out_host.data_ptr(), None, None,
                                                stage_t.data_ptr(), stage_r.data_ptr(), session.result.data_ptr(),
                                                out_dev.data_ptr(), ws, st, None, sptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, steps, warmup):
        for k in range(warmup):
            fn(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = lib.mgb_launch_count()
        e0.record(stream)
        for k in range(steps):
            fn(warmup + k)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        return ms, lib.mgb_launch_count() - launches0

    # end to end, batch entry: three tracks in flight per GPU (mgb_pipeline_*), host buffers in and out
    from matchering_b200.batch import MasteringPipeline
    depth = 3
    pipe = MasteringPipeline(cfg, n, n, depth, device)
    outs_host = [torch.empty((n, 2), dtype=torch.float32).pin_memory() for _ in range(depth)]
    s_h2d, _, s_d2h = (torch.cuda.ExternalStream(p, device=device) for p in pipe.streams())

    # the same with 16-bit PCM host buffers (what the files hold): a quarter of the H2D bytes of float64
    # arrays, half of float32
    pcm_t = [(t.numpy() * 32767.0).round().astype(np.int16) for t in host_t]
    pcm_r = [(r.numpy() * 32767.0).round().astype(np.int16) for r in host_r]
    pcm_t = [torch.from_numpy(a).pin_memory() for a in pcm_t]
    pcm_r = [torch.from_numpy(a).pin_memory() for a in pcm_r]
    pcm_out = [torch.empty((n, 2), dtype=torch.int16).pin_memory() for _ in range(depth)]

    def timed_pipeline(steps, warmup, pcm=False):
        if pcm:
            def submit(k, slot_k):
                pipe.submit_pcm(pcm_t[k % n_sets], pcm_r[k % n_sets], pcm_out[slot_k % depth])
        else:
            def submit(k, slot_k):
                pipe.submit(host_t[k % n_sets], host_r[k % n_sets], outs_host[slot_k % depth])
        return _timed_pipeline(submit, steps, warmup)

    def _timed_pipeline(submit, steps, warmup):
        for k in range(warmup):
            submit(k, k)
        pipe.wait_all()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_h2d)
        for k in range(steps):
            submit(warmup + k, k)
        pipe.wait_all()
        e1.record(s_d2h)
        barrier()
        return e0.elapsed_time(e1)

    sampler.begin()
    dev_ms, launches = timed_lanes(args.steps, args.warmup)
    dev_serial_ms, _ = timed(step_device, args.steps, args.warmup)
    e2e_single_ms, _ = timed(step_host, args.steps, max(3, args.warmup))
    e2e_ms = timed_pipeline(args.steps, max(3, args.warmup))
    e2e_pcm_ms = timed_pipeline(args.steps, max(3, args.warmup), pcm=True)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    pipe.close()

    # max over ranks of the device times
    if world > 1:
        tms = torch.tensor([dev_ms, e2e_ms, e2e_single_ms, dev_serial_ms, e2e_pcm_ms], dtype=torch.float64, device=device)
        gathered = [torch.zeros_like(tms) for _ in range(world)]
        dist.all_gather(gathered, tms)
        dev_ms = max(float(g[0]) for g in gathered)
        e2e_ms = max(float(g[1]) for g in gathered)
        e2e_single_ms = max(float(g[2]) for g in gathered)
        dev_serial_ms = max(float(g[3]) for g in gathered)
        e2e_pcm_ms = max(float(g[4]) for g in gathered)
    frames_total = world * args.steps * n
    value = frames_total / (dev_ms * 1e-3) / SAMPLE_RATE
    e2e_value = frames_total / (e2e_ms * 1e-3) / SAMPLE_RATE

    result = None
    if rank == 0:
        # ---- per-kernel durations: a separate profiled pass (CUDA events around every launch) ----
        lib.mgb_profile_enable(1)
        prof_steps = 5
        for k in range(prof_steps):
            step_device(k)
        cap = 4096
        names = C.create_string_buffer(1 << 16)
        ms = (C.c_float * cap)()
        got = lib.mgb_profile_collect(names, len(names), ms, cap)
        lib.mgb_profile_enable(0)
        per_kernel = {}
        for name, t in zip(names.value.decode().split("\n"), list(ms)[:got]):
            per_kernel.setdefault(name, []).append(float(t))
        summary = {k: {"launches_per_step": len(v) / prof_steps, "avg_ms": sum(v) / len(v),
                       "ms_per_step": sum(v) / prof_steps} for k, v in per_kernel.items()}
        dominant = max(summary, key=lambda k: summary[k]["ms_per_step"])
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bpf = ALGORITHMIC_BYTES_PER_FRAME.get(dominant)
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = None
        if os.path.exists(traffic_path):
            traffic = json.load(open(traffic_path)).get(dominant)
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                    "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": summary[dominant]["avg_ms"],
                    "algorithmic_bytes_per_launch": None}
        if bpf:
            alg = bpf * n
            ach = alg / (summary[dominant]["avg_ms"] * 1e-3) / 1e9
            roofline.update(achieved=ach, frac=ach / peak, algorithmic_bytes_per_launch=alg)
        # the FFT kernels sit under the FP32 roof, not the HBM one: real operations
        # per stereo frame (2*5*N*log2(N)/F + spectral product; 5*F*log2(F)/F + magnitudes) against one
        # operation per lane and clock -- reported next to the HBM figure, not instead of it
        fp32_ops = {"convolve_kernel": 290.0, "analyze_kernel": 75.0}.get(dominant)
        if fp32_ops:
            props = torch.cuda.get_device_properties(device)
            peak_ops = props.multi_processor_count * 128 * 1.965e9  # lanes x max SM clock
            ops = fp32_ops * n
            roofline["cuda_core"] = {"ops_per_launch": ops, "peak_top_per_s": peak_ops / 1e12,
                                     "frac": ops / (summary[dominant]["avg_ms"] * 1e-3) / peak_ops,
                                     "note": "non-FMA FP32 operations; floor of this kernel = ops / peak"}
        # whole pipeline against its compulsory bytes (56 T + 8 R)
        pipeline_bytes = 56 * n + 8 * n
        pipe_ach = pipeline_bytes / (dev_ms / args.steps * 1e-3) / 1e9
        roofline["pipeline"] = {"algorithmic_bytes_per_step": pipeline_bytes, "achieved": pipe_ach, "frac": pipe_ach / peak}

        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            # the oracle port, float64 numpy/scipy, one thread, on this box's host cores
            t64 = host_t[0].numpy().astype(np.float64)
            r64 = host_r[0].numpy().astype(np.float64)
            t0 = time.perf_counter()
            port.main(t64, r64, port.OracleConfig(), True, False, False)
            cpu_s = time.perf_counter() - t0
            cpu_baseline = {"value": args.seconds / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"one full {args.seconds:.0f}-s config-2 track, oracle/port.py (numpy/scipy float64), "
                                      f"{cpu_s:.2f} s wall; host has {os.cpu_count()} cores"}
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "samples_per_sec": frames_total / (dev_ms * 1e-3),
            "config": {"workload": f"config 2: {args.seconds:.0f}-s stereo 44.1 kHz synthetic track vs {args.seconds:.0f}-s reference, "
                                   "full pipeline stages.main(need_default), one track per GPU per step",
                       "frames_per_track": n, "tracks_per_step": world, "tracks_in_flight_per_gpu": n_lanes,
                       "one_track_at_a_time": {"value": frames_total / (dev_serial_ms * 1e-3) / SAMPLE_RATE,
                                               "ms_per_step": dev_serial_ms / args.steps},
                       "l2": "inputs larger than L2: 3 rotating tracks per rank, ~290 MB touched per step",
                       "host_cores_bound_per_rank": bound_cores,
                       "precision": "float32 I/O and FFTs, float64 reductions / FIR design / IIR state"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8,
                    "api": "mgb_pipeline_submit/wait (batch entry, 3 tracks in flight per GPU; pinned float32 host "
                           "buffers in and out; every step's H2D and D2H copies are inside the timed region)",
                    "pcm16": {"value": frames_total / (e2e_pcm_ms * 1e-3) / SAMPLE_RATE, "ms_per_step": e2e_pcm_ms / args.steps,
                              "h2d_bytes_per_step": 2 * n * 4, "d2h_bytes_per_step": n * 4,
                              "api": "mgb_pipeline_submit_pcm (int16 host buffers in and out, decoded / quantised on the device)"},
                    "single_call": {"value": frames_total / (e2e_single_ms * 1e-3) / SAMPLE_RATE,
                                    "ms_per_step": e2e_single_ms / args.steps,
                                    "api": "mgb_process_host (one track per call, copies and kernels back to back)"}},
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": summary, "cpu_baseline": cpu_baseline, "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result or {}


def main():
    args = parse_args()
    out = run_reference(args) if args.impl == "reference" else run_b200(args)
    if out:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
