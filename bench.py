#!/usr/bin/env python3
"""bench.py — aggregate audio-seconds decoded per second (BASELINE.json metric) on N B200s.

A step = one pass of the receiver hot path over one batch: `streams` synthetic 60 s s16le streams per GPU at 22050 Hz
(SAME bursts, AWGN 10 dB SNR, +-5 Hz offset), every stream decoded from a freshly reset receiver.
Weak scaling: every GPU gets its own `streams` streams (different seeds); no collective on the data path — NCCL is used
only for the barrier and the max-over-ranks of the measured time.

  value     device-resident inputs, CUDA-event time of K steps on the engine's own streams (reset + kernel + event read-back)
  e2e       same workload from pinned HOST memory through the C ABI (same_engine_submit_s16_2d), time-chunked so the
            host->device copy of chunk k+1 overlaps the kernel of chunk k; events copied back to the host every step
  roofline  algorithmic bytes (2 B/sample) of one receiver-kernel launch / its CUDA-event duration,
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (C++ restatement of sameold 0.6.0, one stream per host thread) on a bounded sample of the
            same corpus — `kind: "port"` (the Rust reference cannot be built here)

`--impl reference` times that CPU implementation alone (rank 0 only), same metric/config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RATE = 22050
METRIC = "audio_seconds_decoded_per_second"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU (config 3: 4096)")
    ap.add_argument("--seconds", type=float, default=60.0, help="stream duration")
    ap.add_argument("--e2e-chunks", type=int, default=24, help="time-chunks per step on the host-buffer path")
    ap.add_argument("--cpu-sample-streams", type=int, default=1024)
    ap.add_argument("--no-bursts", action="store_true", help="diagnostic: noise-only corpus (not a valid bench line)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_name(streams, seconds):
    return (f"config3: {streams} synthetic {seconds:g} s streams per GPU @22050 Hz s16le, 3 header + 3 EOM SAME bursts, "
            f"AWGN 10 dB SNR, +-5 Hz tone offset, seed 0x5A3E0000+stream_id")


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU during the timed region (B200_PROFILING.md recipe): NVML polled
    from a thread every 10 ms (no process start-up, so even a sub-second region gets samples); `nvidia-smi -lms` is the
    fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.thread, self.stop_flag = None, None, None, threading.Event()
        self.sm, self.max_sm, self.bits = [], None, 0
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            n = self.nvml
            masks = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
            reasons = [name for name, m in zip(self.NAMES, masks) if self.bits & m]
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "samples": len(self.sm), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "source": "nvidia-smi"}


def oracle_config_from(cfg):
    from oracle.pyoracle import OracleConfig
    o = OracleConfig()
    for name, _ in OracleConfig._fields_:
        setattr(o, name, getattr(cfg, name))
    return o


def cpu_baseline(host_samples, cfg, cores):
    """Oracle on the host cores: one receiver per stream, `cores` threads.  Best of 2."""
    from oracle import Oracle
    ocfg = oracle_config_from(cfg)
    best = None
    for _ in range(2):
        secs, nb, nm = Oracle.decode_batch(ocfg, host_samples, cores)
        best = secs if best is None else min(best, secs)
    audio = host_samples.shape[0] * host_samples.shape[1] / RATE
    return audio / best, best, int(nb.sum()), int(nm.sum())


def bind_to_gpu_numa_node(device):
    """Pin this process (and therefore its first-touch host allocations, the pinned sample buffer above all) to the CPU
    cores of the NUMA node the GPU hangs off, so that with one rank per GPU the host->device copies of different
    ranks do not cross the socket interconnect.  Best effort; returns the node or None."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_samples = int(args.seconds * RATE)
    cfgname = {"workload": workload_name(args.streams, args.seconds), "streams_per_gpu": args.streams,
               "seconds": args.seconds, "rate_hz": RATE, "receiver_config": "samedec (main.rs:29-37)",
               "l2": "inputs_exceed_l2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, cfgname, n_samples)

    import torch
    import torch.distributed as dist
    import sameold_b200 as sb
    from sameold_b200 import synth, _lib

    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic corpus, resident in HBM ----
    ns = args.streams
    stride = (n_samples + 7) // 8 * 8
    buf = torch.empty((ns, stride), dtype=torch.int16, device="cuda")
    plans = synth.plan_corpus(ns, RATE, args.seconds, first_stream=rank * ns)
    if args.no_bursts:
        for pl in plans:
            pl.burst_starts, pl.burst_payloads = [], []
    synth.generate_on_device(plans, buf.data_ptr(), stride, n_samples, RATE, device=local_rank)
    offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
    lengths = np.full(ns, n_samples, np.uint32)

    builder = sb.SameReceiverBuilder.samedec(RATE)
    rx = builder.build_batch(ns, device=local_rank)
    audio_per_step = ns * n_samples / RATE

    def step_device():
        rx.reset()
        rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
        rx.sync()
        return rx.drain_raw(reuse=True)

    # correctness guard inside the bench: every step must decode the corpus (no skipped work)
    for _ in range(args.warmup):
        evs, _pay = step_device()
    n_headers = int((evs["kind"] == 18).sum())
    n_events = int(evs.size)

    sampler = ClockSampler(local_rank)
    launches0 = rx.launch_count()
    kernel_ms = []
    barrier()
    sampler.start()
    rx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evs, _pay = step_device()
        kernel_ms.append(rx.last_timing()[1])
    dev_ms = rx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    launches = rx.launch_count() - launches0
    dev_ms = max_over_ranks(dev_ms)
    value = audio_per_step * args.steps * world / (dev_ms * 1e-3)
    assert int((evs["kind"] == 18).sum()) == n_headers and (n_headers > 0 or args.no_bursts), "bench step lost its work"
    if args.seconds >= 60.0 and not args.no_bursts:
        assert n_headers >= int(0.9 * ns), "corpus not decoded"

    # ---- roofline of the receiver kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    k_ms = statistics.mean(kernel_ms)
    achieved = (2.0 * ns * n_samples) / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "rx_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("streams") == ns and abs(tj.get("seconds", 0) - args.seconds) < 1e-9:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "same_rx_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                "kernel_ms_per_launch": round(k_ms, 3), "algorithmic_bytes_per_launch": 2 * ns * n_samples,
                "note": "fused per-lane receiver loop is issue/latency-bound, not HBM-bound (DESIGN.md §5)"}

    # ---- e2e: host buffers through the C ABI, H2D + event D2H inside the timed region ----
    e2e = None
    host_np = None
    if not args.no_e2e:
        lib = _lib.load()
        nbytes = ns * stride * 2
        hptr = lib.same_host_alloc(nbytes)
        if not hptr:
            raise RuntimeError("pinned host allocation failed")
        import ctypes as C
        host_np = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_int16)), shape=(ns, stride))
        # same corpus, now in pinned host memory (copied out once, untimed)
        host_t = torch.from_numpy(host_np)
        host_t.copy_(buf)
        torch.cuda.synchronize()
        nchunk = max(1, args.e2e_chunks)
        bounds = [int(round(i * n_samples / nchunk)) for i in range(nchunk + 1)]
        d2h_bytes = []

        def step_host():
            rx.reset()
            for c in range(nchunk):
                rx.submit_2d(hptr, stride, bounds[c], bounds[c + 1] - bounds[c])
            rx.sync()
            return rx.drain_raw(reuse=True)

        for _ in range(max(1, args.warmup - 1)):
            ev, pay = step_host()
        assert int((ev["kind"] == 18).sum()) == n_headers, "chunked host path decodes differently"
        barrier()
        rx.timer_start()
        for _ in range(args.steps):
            ev, pay = step_host()
            d2h_bytes.append(48 * int(ev.size) + int(pay.size) + 8 * nchunk)
        e2e_ms = rx.timer_stop()
        barrier()
        e2e_ms = max_over_ranks(e2e_ms)
        e2e = {"value": round(audio_per_step * args.steps * world / (e2e_ms * 1e-3), 1), "unit": UNIT,
               "h2d_bytes_per_step": int(ns * n_samples * 2 + nchunk * ns * 12) * world,
               "d2h_bytes_per_step": int(sum_over_ranks(statistics.mean(d2h_bytes))), "ms_per_step": round(e2e_ms / args.steps, 3),
               "chunks_per_step": nchunk, "api": "same_engine_submit_s16_2d + sync + drain_events (pinned host buffer)",
               "rank0_numa_node": numa_node}

    # ---- CPU baseline on rank 0 at N=1 ----
    cpu = None
    if not args.no_cpu and world == 1 and rank == 0:
        k = min(args.cpu_sample_streams, ns)
        sample = buf[:k, :n_samples].cpu().numpy()
        cores = os.cpu_count() or 1
        v, secs, nb, nm = cpu_baseline(sample, builder.config(), cores)
        cpu = {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {k} streams x {args.seconds:g} s of this workload, one receiver per stream, {cores} threads, best of 2 ({secs:.2f} s)",
               "what": "oracle/ C++ restatement of sameold 0.6.0 (link + transport layers), g++ -O2 -ffp-contract=off; the Rust crate cannot be built here"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgname,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_ms_per_step": round(wall_ms / args.steps, 3), "events_per_step": n_events,
            "headers_decoded_per_step": n_headers, "realtime_factor_per_gpu": round(value / world, 1),
        }
        print(json.dumps(line))
    if host_np is not None:
        del host_np
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(args, cfgname, n_samples):
    """The reference's CPU implementation of the path (oracle port) on all host threads; each step = a bounded sample
    (cpu_sample_streams streams) of the same workload."""
    from oracle import Oracle
    from sameold_b200 import synth, _lib
    import ctypes as C
    cores = os.cpu_count() or 1
    k = min(args.cpu_sample_streams, args.streams)
    cfg = _lib.SameConfig()
    _lib.load().same_config_samedec(C.byref(cfg), RATE)
    plans = synth.plan_corpus(k, RATE, args.seconds, first_stream=0)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        import torch
        stride = (n_samples + 7) // 8 * 8
        buf = torch.empty((k, stride), dtype=torch.int16, device="cuda")
        synth.generate_on_device(plans, buf.data_ptr(), stride, n_samples, RATE, device=0)
        sample = buf[:, :n_samples].cpu().numpy()
        del buf
        how = "device generator"
    else:
        k = min(k, 16)
        sample = np.stack([synth.render_numpy(p, n_samples, RATE) for p in plans[:k]])
        how = "numpy generator (no GPU visible)"
    ocfg = oracle_config_from(cfg)
    audio = sample.shape[0] * n_samples / RATE
    for _ in range(min(args.warmup, 1)):
        Oracle.decode_batch(ocfg, sample, cores)
    t0 = time.perf_counter()
    secs_total = 0.0
    for _ in range(args.steps):
        secs, nb, nm = Oracle.decode_batch(ocfg, sample, cores)
        secs_total += secs
    wall = time.perf_counter() - t0
    value = audio * args.steps / secs_total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs_total / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfgname,
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample.shape[0]} streams x {args.seconds:g} s per step ({how}), one receiver per stream, {cores} threads"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(wall, 2), "messages_per_step": int(nm.sum()),
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
