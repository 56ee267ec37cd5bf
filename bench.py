#!/usr/bin/env python
"""bench.py -- throughput of the Quisk receive-DSP hot path on B200 (libquisk_cuda) next to the
reference's own CPU implementation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rx_chain|panadapter|rx_chain+panadapter|rxa_usb|rxa_fm|channelizer]
                    [--channels C] [--block B] [--nco exact|closed]

One "step" = one pass of the hot path over one block of synthetic IQ for every channel:
  rx_chain    C independent receivers at 1.536 MS/s (BASELINE.json configs[0] batched): tune NCO ->
              4 x quisk_cDecim2HB45 -> quisk_cDecimate(98 taps, /2) -> 48 kS/s -> HB45 -> FIR /2 ->
              cRxFilterOut (164-tap I/Q, USB bw 2800) -> re-im -> quisk_dInterpolate x2 ->
              quisk_dInterp2HB45 -> 48 kS/s audio.
  panadapter  BASELINE.json configs[1] batched: 8192-point Hann/FFT/|X| accumulation of every input
              frame of every channel, one dB graph per channel per step.
  rxa_usb     BASELINE.json configs[2]: C WDSP RXA channels at 192 kS/s, nbp0 overlap-save band-pass (4096 taps) +
              wcpAGC + panel, 32 DSP blocks of 1024 samples per step.
  rxa_fm      configs[3]: C channels at 384 kS/s, resample /8 + nbp0 + fmd FM demodulator.
  channelizer configs[4]: one 98.304 MS/s stream -> 1024 receivers x 192 kS/s through the polyphase channelizer; with
              N > 1 every rank takes its own time block of the stream (halo in front, no inter-GPU traffic).
The default run (no --workload) is the headline rx_chain line PLUS short runs of the other four workloads under the extra
key `workloads` (each with its own value / roofline / cpu_baseline / e2e), so that one driver run records every BASELINE
config; --workload X runs X alone, --no-extra keeps the default line to rx_chain.
metric = complex input MS/s, whole job.  `value` is timed with inputs resident in HBM; `e2e` is the
same work through the host-buffer C-ABI entry point (H2D of the block + D2H of the audio inside the
timed region); `e2e_wire` (rx_chain only, extra key) is that step from int16 wire-format host blocks.  With N > 1 (torchrun) every rank runs its own C channels -- independent receivers,
no data-path collective -- and the time is the max over ranks ("weak" scaling); every rank also prints its own step
and kernel time on stderr.  `roofline` is the fused decimator against the measured HBM peak (CUDA events around the
kernel inside the library); `roofline.fp64` = the same launch against the measured FP64 pipe peak, `roofline.smem` =
its shared-memory traffic (bytes per sample from the committed ncu capture) against 128 B/clk/SM.

--impl reference times the reference's own C code (oracle/_ref: filter.c verbatim + the quisk.c RX
functions, gcc -O2, no -ffast-math) on the host cores, one private copy of the library per thread.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 1536000
RATE = [SAMPLE_RATE]                                     # the rx_chain / panadapter workloads' input rate (--rate; `pipeline` sets 192000)
FFT_SIZE = 8192
ALG_BYTES_RX = 16.0 + 8.0 * 48000 / SAMPLE_RATE          # 16 B in + 0.25 B out
ALG_BYTES_PAN = 16.0                                     # C2: 16 B in (+ 8 B/bin per returned graph)
REF_RX_LIB = "libquisk_rx_ref_O3.so"                     # the reference's filter.c + quisk.c RX functions, gcc -O3 (oracle/build_ref.sh)
REF_FLAGS = "gcc -O3, no -ffast-math"


def synth_block_torch(torch, C_, n, device, seed):
    """Multi-tone + noise IQ generated on the device: [C, n] complex128."""
    g = torch.Generator(device=device); g.manual_seed(1234 + seed)
    t = torch.arange(n, device=device, dtype=torch.float64)
    fr = torch.tensor([0.01, -0.01, 0.07, -0.07, 0.13, -0.13, 0.31, -0.31], device=device, dtype=torch.float64)
    amp = 2.0 ** torch.linspace(24, 30, 8, device=device, dtype=torch.float64)
    x = torch.zeros((C_, n), dtype=torch.complex128, device=device)
    CH = 64
    for c0 in range(0, C_, CH):
        c1 = min(C_, c0 + CH)
        ph = torch.rand((c1 - c0, 8), generator=g, device=device, dtype=torch.float64) * 2 * np.pi
        acc = torch.zeros((c1 - c0, n), dtype=torch.complex128, device=device)
        for k in range(8):
            acc += amp[k] * torch.exp(1j * (2 * np.pi * fr[k] * t[None, :] + ph[:, k:k + 1]))
        acc += (2.0 ** 20) * (torch.randn((c1 - c0, n), generator=g, device=device, dtype=torch.float64)
                              + 1j * torch.randn((c1 - c0, n), generator=g, device=device, dtype=torch.float64))
        x[c0:c1] = acc
    return x


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.sm = []; self.reasons = set(); self.sm_max = None; self.ok = False
        try:                        # NVML set-up happens here, outside the timed region, so that even a region of a
            import pynvml as nv     # few milliseconds gets its first sample at once
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                          nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                          nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                          nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                          nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            self.ok = True
        except Exception as e:      # NVML missing: report that, never fake a value
            self.err = str(e)

    def run(self):
        if not self.ok or os.environ.get("QUISK_BENCH_NO_SAMPLER"):     # diagnostic switch: the line then says so
            self.ok = False; self.err = "disabled by QUISK_BENCH_NO_SAMPLER"
            return
        nv, h = self.nv, self.h
        try:
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:
            self.err = str(e)

    def result(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["nvml_unavailable: %s" % getattr(self, "err", "no samples")]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# reference CPU arm
# --------------------------------------------------------------------------------------------

def ref_worker_setup(fi, fq):
    from oracle import ref_ctypes as R
    lib = R.load(REF_RX_LIB if R.have_ref(REF_RX_LIB) else "libquisk_rx_ref.so", private_copy=True)
    lib.ref_set_sample_rate(RATE[0]); lib.ref_init_chain()
    lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
    lib.ref_tune.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    return lib


def ref_run_block(lib, x, scratch, dbuf, vec, tune_hz):
    """One channel, one block, in <= 61440-sample calls (the reference's buffers hold 66000)."""
    n = len(x); pos = 0; total = 0
    while pos < n:
        m = min(61440, n - pos)
        scratch[:m] = x[pos:pos + m]
        if tune_hz:
            lib.ref_tune(scratch.ctypes.data, m, tune_hz, RATE[0], vec.ctypes.data)
        nd = lib.ref_process_decimate(scratch.ctypes.data_as(C.c_void_p), m, 0, 3)
        total += lib.ref_process_demodulate(scratch.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), nd, 0, 0, 3)
        pos += m
    return total


def ref_pan_block(x, window):
    """Reference panadapter math on the CPU for one channel block: numpy pocketfft stands in for
    FFTW (un-vendored dependency of the reference, absent from this image)."""
    fr = x.reshape(-1, FFT_SIZE) * window
    return np.abs(np.fft.fftshift(np.fft.fft(fr, axis=-1), axes=-1)).sum(axis=0)


def cpu_reference_rate(block, steps, warmup, workload, fi, fq, tune_hz):
    """Complex MS/s of the reference CPU code on all host cores: every core runs one channel."""
    from oracle import quisk_oracle as O
    cores = os.cpu_count() or 1
    libs = [ref_worker_setup(fi, fq) for _ in range(cores)]
    xs = [O.synth_iq(block, 1000 + i, 1.0) for i in range(cores)]
    window = O.hann_window(FFT_SIZE)
    state = [(np.zeros(66000, dtype=np.complex128), np.zeros(132000), np.array([1.0 + 0j])) for _ in range(cores)]

    def work(i, nsteps):
        for _ in range(nsteps):
            if "rx_chain" in workload:
                ref_run_block(libs[i], xs[i], state[i][0], state[i][1], state[i][2], tune_hz)
            if "panadapter" in workload:
                ref_pan_block(xs[i][: (block // FFT_SIZE) * FFT_SIZE], window)

    def run(nsteps):
        th = [threading.Thread(target=work, args=(i, nsteps)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return time.perf_counter() - t0

    run(warmup)
    dt = run(steps)
    return cores * block * steps / dt / 1e6, cores, dt


# --------------------------------------------------------------------------------------------
# WDSP RXA workloads (BASELINE.json configs[2] and configs[3])
# --------------------------------------------------------------------------------------------

RXA_CFG = {
    # 64 ch x 192 kS/s, dsp_size 1024, nbp0 nc 4096 (nfor 4, FFT 2048), wcpAGC mode 3, panel
    "rxa_usb": dict(channels=64, in_size=1024, dsp_size=1024, in_rate=192000, dsp_rate=192000, out_rate=192000, mode=1,
                    passband=(150.0, 2850.0), nc=4096, agc=3, alg_bytes=32.0, flops=330.0,
                    name="rxa_usb: C ch x 192 kS/s WDSP RXA nbp0 overlap-save bandpass (4096 taps) + wcpAGC + panel (BASELINE configs[2])"),
    # C4: 256 ch x 384 kS/s -> resample (1121 taps, /8) -> 48 k -> nbp0 -> fmd (PLL + 2 fircores + notch) -> panel
    "rxa_fm": dict(channels=256, in_size=2048, dsp_size=256, in_rate=384000, dsp_rate=48000, out_rate=48000, mode=5,
                   passband=(-8000.0, 8000.0), nc=2048, agc=None, alg_bytes=18.0, flops=650.0,
                   name="rxa_fm: C ch x 384 kS/s WDSP RXA resample + nbp0 + fmd FM demod (BASELINE configs[3])"),
}


def rxa_reference_rate(cfg, blocks, steps, warmup):
    """The reference's own stage functions (libwdsp_ref.so: wdsp/*.c + our FFTW-API shim -- FFTW3 itself is absent
    from this image) composed in xrxa's order, one channel per host core."""
    from tests.golden.make_golden_wdsp import wdsp, bandpass
    from quisk_b200.synth import sig, fm_sig
    lib = wdsp()
    cores = os.cpu_count() or 1
    n, m = cfg["dsp_size"], cfg["in_size"]
    objs = []
    for i in range(cores):
        inb = np.zeros(m, dtype=np.complex128); buf = np.zeros(2 * n, dtype=np.complex128)
        o = {"inb": inb, "buf": buf}
        if cfg["in_rate"] != cfg["dsp_rate"]:
            o["rs"] = lib.create_resample(1, m, inb.ctypes.data, buf.ctypes.data, cfg["in_rate"], cfg["dsp_rate"], 0.0, 0, 1.0)
        imp = bandpass(lib, cfg["nc"], cfg["passband"][0], cfg["passband"][1], float(cfg["dsp_rate"]), 0, 1, 1.0 / (2 * n))
        o["nbp"] = lib.create_fircore(n, buf.ctypes.data, buf.ctypes.data, cfg["nc"], 0, imp.ctypes.data)
        if cfg["mode"] == 5:
            o["fmd"] = lib.create_fmd(1, n, buf.ctypes.data, buf.ctypes.data, cfg["dsp_rate"], 5000.0, 300.0, 3000.0, -8000.0, 8000.0,
                                      1.0, 20000.0, 0.02, 0.5, 1, 254.1, cfg["nc"], 0, cfg["nc"], 0)
        else:
            o["agc"] = lib.create_wcpagc(1, 3, 1, buf.ctypes.data, buf.ctypes.data, n, cfg["dsp_rate"], 0.001, 0.250, 4, 10000.0, 1.5,
                                         1000.0, 1.0, 1.0, 0.250, 0.005, 5.0, 1, 0.500, 0.0, 1.0, 0.100)
        o["x"] = (fm_sig if cfg["mode"] == 5 else sig)(m * 4, 900 + i, float(cfg["in_rate"]))
        objs.append(o)

    def work(i, nblk):
        o = objs[i]
        for b in range(nblk):
            if "rs" in o:
                o["inb"][:] = o["x"][(b % 4) * m:(b % 4 + 1) * m]; lib.xresample(o["rs"])
            else:
                o["buf"][:n] = o["x"][(b % 4) * m:(b % 4 + 1) * m]
            lib.xfircore(o["nbp"])
            if "fmd" in o: lib.xfmd(o["fmd"])
            else: lib.xwcpagc(o["agc"])
            o["buf"][:n] *= 4.0

    def run(nblk):
        th = [threading.Thread(target=work, args=(i, nblk)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return time.perf_counter() - t0
    run(warmup * blocks)
    dt = run(steps * blocks)
    return cores * m * blocks * steps / dt / 1e6, cores, dt


class Ctx:
    """One process = one rank = one GPU: torch / torch.distributed (NCCL for the barrier and the max-over-ranks time only) and
    the library handle, set up once and shared by every workload of the run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from quisk_b200 import lib as L
        self.torch, self.dist, self.L = torch, dist, L
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.lib = L.require_device()                    # raises without a CUDA device: there is no CPU fallback
        pin_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        L.check(self.lib, self.lib.quisk_cuda_set_device(self.local_rank), "set_device")
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def pin_to_gpu_numa_node(local_rank):
    """Run this rank's host threads on the CPUs of the NUMA node its GPU hangs off (pinned staging buffers are then
    allocated there by first touch): the host side of the e2e path is memory-bandwidth bound, and with eight ranks
    feeding eight PCIe links every remote-node access counts.  Best effort; silently skipped where sysfs has no answer."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local_rank)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bus.lower()[-12:]
        if not os.path.exists(path):
            path = "/sys/bus/pci/devices/0000:%s/local_cpulist" % bus.lower()[-7:]
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                lo, hi = part.split("-"); cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        node = None
        try:
            node = int(open(path.replace("local_cpulist", "numa_node")).read().strip())
        except Exception:
            pass
        if cpus:
            os.sched_setaffinity(0, cpus)
            if os.environ.get("QUISK_BENCH_VERBOSE"):
                print("[rank %d] GPU %s on NUMA node %s: host threads on %d CPUs (%d..%d)" % (local_rank, bus, node, len(cpus), min(cpus), max(cpus)), file=sys.stderr, flush=True)
            return sorted(cpus)
    except Exception as ex:
        if os.environ.get("QUISK_BENCH_VERBOSE"):
            print("[rank %d] NUMA pinning skipped: %s" % (local_rank, ex), file=sys.stderr, flush=True)
    return None


def timed_steps(ctx, step, steps, warmup, after_warmup=None):
    """W untimed steps, then exactly `steps` steps between two events on the launching stream, barrier + synchronize on
    both sides; returns (ms on this rank, max over ranks, launches, clocks)."""
    torch = ctx.torch
    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    if after_warmup is not None:
        after_warmup()
    sampler = ClockSampler(ctx.local_rank); sampler.start()
    l0 = ctx.lib.quisk_cuda_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1)
    launches = int(ctx.lib.quisk_cuda_launch_count() - l0)
    sampler.stop_flag = True; sampler.join(timeout=2)
    return ms, ctx.max_over_ranks(ms), launches, sampler.result()


def rxa_reference_line(args, cfg, blocks):
    v, cores, dt = rxa_reference_rate(cfg, blocks, args.steps, args.warmup)
    return {"impl": "reference", "metric": "complex MS/s through RX chain", "value": v, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": cfg["name"], "channels": cores, "blocks_per_step": blocks},
            "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                             "sample": "%d channels x %d blocks x %d steps; libwdsp_ref.so stage functions (FFT = our shim, not FFTW3)" % (cores, blocks, args.steps)},
            "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_rxa(args, ctx, workload, steps, warmup, e2e_steps, cpu_steps):
    """BASELINE configs[2] / configs[3].  Returns the JSON dict on rank 0, None elsewhere (every rank runs its own channels)."""
    torch, lib, L = ctx.torch, ctx.lib, ctx.L
    from quisk_b200.synth import sig, fm_sig           # plain NumPy generators: the product arm imports nothing from oracle/
    cfg = dict(RXA_CFG[workload])
    C_ = args.channels if args.channels > 0 else cfg["channels"]
    blocks = 32
    dev = ctx.dev
    rxa = lib.quisk_cuda_rxa_create(C_, cfg["in_size"], cfg["dsp_size"], cfg["in_rate"], cfg["dsp_rate"], cfg["out_rate"])
    if not rxa:
        raise L.QuiskCudaError(lib.quisk_cuda_last_error().decode())
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    L.check(lib, lib.quisk_cuda_rxa_set_nc(rxa, cfg["nc"]), "set_nc")
    L.check(lib, lib.quisk_cuda_rxa_set_mode(rxa, cfg["mode"]), "set_mode")
    L.check(lib, lib.quisk_cuda_rxa_set_passband(rxa, *cfg["passband"]), "set_passband")
    if cfg["agc"] is not None:
        lib.quisk_cuda_rxa_set_agc_mode(rxa, cfg["agc"])
    m = cfg["in_size"]; osz = lib.quisk_cuda_rxa_out_size(rxa)
    base = np.stack([(fm_sig if cfg["mode"] == 5 else sig)(m * blocks, 900 + (c % 16), float(cfg["in_rate"])) for c in range(min(C_, 16))])
    x = torch.from_numpy(base).to(dev).repeat((C_ + 15) // 16, 1)[:C_].contiguous()
    y = torch.zeros((C_, osz * blocks), dtype=torch.complex128, device=dev)
    stream = ctx.stream
    multi = hasattr(lib, "quisk_cuda_rxa_xrxa_multi") and not args.rxa_per_block

    def step():
        if multi:
            L.check(lib, lib.quisk_cuda_rxa_xrxa_multi(rxa, x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), blocks, stream), "xrxa_multi")
        else:
            for b in range(blocks):
                L.check(lib, lib.quisk_cuda_rxa_xrxa(rxa, x.data_ptr() + b * m * 16, x.stride(0), y.data_ptr() + b * osz * 16, y.stride(0), stream), "xrxa")

    ms, ms_max, launches, clocks = timed_steps(ctx, step, steps, warmup)
    value = ctx.world * C_ * m * blocks * steps / (ms_max / 1e3) / 1e6
    # e2e: the fexchange0 entry (wdsp/iobuffs.c:464) with host buffers for all channels, one call per in_size block:
    # H2D of the block, the DSP turn, D2H of the block that leaves the output ring
    e2e = None
    if e2e_steps > 0:
        hxt = torch.empty((C_, m), dtype=torch.complex128).pin_memory(); hxt.copy_(x[:, :m].cpu())
        hyt = torch.zeros((C_, osz), dtype=torch.complex128).pin_memory()
        hx = hxt.numpy(); hy = hyt.numpy()
        err = C.c_int(0)
        lib.quisk_cuda_rxa_fexchange0(rxa, hx.ctypes.data, hy.ctypes.data, C.byref(err))
        ctx.barrier()
        t0 = time.perf_counter()
        nb = blocks * e2e_steps
        for _ in range(nb):
            lib.quisk_cuda_rxa_fexchange0(rxa, hx.ctypes.data, hy.ctypes.data, C.byref(err))
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": ctx.world * C_ * m * nb / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": C_ * m * 16 * blocks,
               "d2h_bytes_per_step": C_ * osz * 16 * blocks, "channels": C_, "note": "quisk_cuda_rxa_fexchange0 (the reference's exchange call for all channels at once), pinned host buffers, one call per DSP block: H2D, ring arithmetic, DSP turn, D2H, synchronous like the reference's blocking exchange"}
    lib.quisk_cuda_rxa_destroy(rxa)
    del x, y
    if ctx.rank != 0:
        return None
    peak, peak_src = load_peaks()
    alg = cfg["alg_bytes"] * C_ * m * blocks
    ach = alg * steps / (ms / 1e3) / 1e9
    cpu = None
    if not args.no_cpu_baseline and ctx.world == 1:       # a reported baseline: rank 0 at N = 1 only
        try:
            v, cores, dt = rxa_reference_rate(cfg, blocks, cpu_steps, 1)
            cpu = {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                   "sample": "%d channels x %d blocks x %d steps, %.1f s wall; libwdsp_ref.so stage functions, FFT = our shim (FFTW3 absent)" % (cores, blocks, cpu_steps, dt)}
        except Exception as ex:
            cpu = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}
    roofline = {"bound": "hbm", "kernel": "whole step (resampler + fircores + recurrent stages)", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_sample": cfg["alg_bytes"]}
    pk = C.c_double(0.0)
    if lib.quisk_cuda_fp64_peak(C.byref(pk)) == 0 and pk.value > 0:
        # these two chains are FP64-pipe bound, not HBM bound.  Algorithmic flop per input sample from
        # SURVEY.md 8(d) / DESIGN.md 4.5 (C3: 2 x 2048-point FFT + 4 partition MACs per 1024 samples + AGC + panel = 330;
        # C4: 1121-tap / 8 resampler = 560, + nbp0 + the two fmd fircores at 1/8 rate = 650) against 2 flop per measured DFMA slot.
        fl = cfg["flops"] * C_ * m * blocks * steps / (ms / 1e3)
        roofline["fp64"] = {"achieved": fl / 1e12, "peak": 2.0 * pk.value / 1e12, "unit": "TFLOP/s", "frac": fl / (2.0 * pk.value),
                            "flop_per_input_sample": cfg["flops"], "peak_source": "measured (quisk_cuda_fp64_peak: 8 independent DFMA chains per thread, x 2 flop)"}
    return {"metric": "complex MS/s through RX chain", "value": value, "unit": "MS/s", "n_gpus": ctx.world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "channels_per_gpu": C_, "blocks_per_step": blocks, "in_size": m, "dsp_size": cfg["dsp_size"],
                       "launch": "one launch chain per step of %d DSP blocks" % blocks if multi else "one launch chain per DSP block",
                       "l2": "working set %.0f MB per GPU; L2-resident by design for this config (FDL + masks), not flushed" % (C_ * (m + osz) * blocks * 16 / 1e6)},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu}

# --------------------------------------------------------------------------------------------
# C5: wideband polyphase channelizer (BASELINE.json configs[4])
# --------------------------------------------------------------------------------------------

PFB = dict(K=1024, D=512, P=16, fs=98.304e6, alg_bytes=48.0,
           name="channelizer: one 98.304 MS/s stream -> 1024 receivers x 192 kS/s (D=512, 16x1024-tap fir_bandpass prototype), "
                "time-block sharded with halos (BASELINE configs[4])")


def pfb_proto():
    from quisk_b200 import lib as L
    lib = L.load()
    h = np.zeros(PFB["K"] * PFB["P"])
    fc = 0.4 * PFB["fs"] / PFB["K"]
    rc = lib.quisk_cuda_fir_bandpass(len(h), -fc, fc, PFB["fs"], 1, 0, 1.0, h.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return h


def pfb_reference_rate(n_samples, steps, warmup, reps=1):
    """The reference's way to get receiver k out of the wideband stream: its tune loop (quisk.c:2477-2494, compiled
    reference) and quisk_cDecimate with the 16384-tap prototype, /512 (filter.c:203-229) -- per receiver.  One
    receiver per host core over n_samples; the wideband rate is what all 1024 receivers would sustain."""
    from oracle import quisk_oracle as O
    from oracle import ref_ctypes as R
    cores = os.cpu_count() or 1
    h = pfb_proto()
    libs, sts, xs = [], [], []
    for i in range(cores):
        lib = R.bind_filter_api(R.load("libquisk_rx_ref.so", private_copy=True))
        lib.ref_tune.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        st = R.cFilter()
        lib.quisk_filt_cInit(C.byref(st), h.ctypes.data_as(R.c_double_p), len(h))
        libs.append(lib); sts.append(st); xs.append(O.synth_iq(n_samples, 1100 + i, 1.0))
    scratch = [np.zeros(61440, dtype=np.complex128) for _ in range(cores)]
    vec = [np.array([1.0 + 0j]) for _ in range(cores)]

    def work(i, nsteps):
        for _ in range(nsteps * reps):          # one step = `reps` passes over the receiver's block (state carries on)
            pos = 0
            while pos < n_samples:
                m = min(61440, n_samples - pos)
                scratch[i][:m] = xs[i][pos:pos + m]
                libs[i].ref_tune(scratch[i].ctypes.data, m, (i + 1) * PFB["fs"] / PFB["K"], int(PFB["fs"]), vec[i].ctypes.data)
                libs[i].quisk_cDecimate(scratch[i].ctypes.data, m, C.byref(sts[i]), PFB["D"])
                pos += m

    def run(nsteps):
        th = [threading.Thread(target=work, args=(i, nsteps)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return time.perf_counter() - t0
    run(warmup)
    dt = run(steps)
    receiver_rate = cores * n_samples * reps * steps / dt / 1e6        # receiver-input MS/s over all cores
    return receiver_rate / PFB["K"], cores, dt


def pfb_reference_line(args):
    cfg = PFB
    ns, reps = 61440, 64
    v, cores, dt = pfb_reference_rate(ns, args.steps, args.warmup, reps)
    return {"impl": "reference", "metric": "complex MS/s through RX chain", "value": v, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": cfg["name"], "receivers_timed": cores, "samples_per_step": ns * reps},
            "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                             "sample": "%d receivers (one per core) x %d samples x %d steps through the reference's tune loop + quisk_cDecimate(16384 taps, /512); "
                                       "value = wideband rate at which all 1024 receivers would be served" % (cores, ns * reps, args.steps)},
            "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_pfb(args, ctx, steps, warmup, e2e_steps, cpu_steps):
    """BASELINE configs[4]: every rank takes its own time block of the wideband stream, halo in front, no inter-GPU traffic."""
    torch, lib, L = ctx.torch, ctx.lib, ctx.L
    from quisk_b200.rx import Channelizer
    from quisk_b200.shard import time_blocks
    cfg = PFB
    n = args.block if args.block != 32768 else (1 << 24)          # input samples per GPU per step (time block)
    n = (n // cfg["D"]) * cfg["D"]
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    K, D, T = cfg["K"], cfg["D"], cfg["K"] * cfg["P"]
    ch = Channelizer(K, D, pfb_proto())
    tb = time_blocks(world * n, world, D, T - 1)[rank]
    halo = tb.start - tb.halo_start
    x = synth_block_torch(torch, 1, halo + (tb.stop - tb.start), dev, 77 + rank)[0].contiguous()
    nf = (tb.stop - tb.start) // D
    y = torch.zeros((K, nf), dtype=torch.complex128, device=dev)
    stream = ctx.stream

    def step():
        ch.seek(tb.halo_start, stream)
        if halo:
            ch.prime(x.data_ptr(), halo, stream)
        got = ch.process(x.data_ptr() + halo * 16, tb.stop - tb.start, y.data_ptr(), nf, 0, stream)
        assert got == nf

    ms, ms_max, launches, clocks = timed_steps(ctx, step, steps, warmup)
    value = world * n * steps / (ms_max / 1e3) / 1e6
    # e2e: pinned host stream in, EVERY receiver's stream out (32 B per input sample back over PCIe) per step
    e2e = None
    if e2e_steps > 0:
        hx = torch.empty((x.numel(), 2), dtype=torch.float64).pin_memory(); hx.copy_(torch.view_as_real(x).cpu())
        hy = torch.empty((K, nf, 2), dtype=torch.float64).pin_memory()
        xr = torch.view_as_real(x); yr = torch.view_as_real(y)
        xr.copy_(hx, non_blocking=True)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            xr.copy_(hx, non_blocking=True)
            step()
            hy.copy_(yr, non_blocking=True)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n * e2e_steps / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": int(x.numel() * 16),
               "d2h_bytes_per_step": int(hy.numel() * 8),
               "note": "pinned host IQ -> H2D -> channelizer -> D2H of all 1024 receiver streams (32 B per input sample): PCIe bound"}
        del hx, hy
    ch.close()
    del x, y
    if rank != 0:
        return None
    peak, peak_src = load_peaks()
    alg = cfg["alg_bytes"] * n
    ach = alg * steps / (ms / 1e3) / 1e9
    cpu = None
    if not args.no_cpu_baseline and world == 1:       # a reported baseline: rank 0 at N = 1 only
        try:
            v, cores, dt = pfb_reference_rate(61440, cpu_steps, 1, 64)
            cpu = {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                   "sample": "%d receivers (one per core) x 61440 samples x 64 passes x %d steps, %.1f s wall: reference tune loop + quisk_cDecimate(16384 taps, /512) per receiver; "
                             "value = wideband rate at which all 1024 receivers would be served" % (cores, cpu_steps, dt)}
        except Exception as ex:
            cpu = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}
    return {"metric": "complex MS/s through RX chain", "value": value, "unit": "MS/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "samples_per_gpu_per_step": n, "halo": halo, "receivers": K, "decimation": D, "taps": T,
                       "l2": "input %.0f MB + output %.0f MB per step per GPU >> 126 MB L2, no flush needed" % (n * 16 / 1e6, n * 32 / 1e6)},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": "channelizer (branch FIRs + 1024-point FFT per frame)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_sample": cfg["alg_bytes"]},
            "cpu_baseline": cpu}


def tx_reference_rate(block, steps, warmup, mode=3, preemph=0.6, clip=2.5):
    """Microphone samples per second of the reference's own tx_filter (oracle/_ref/libquisk_tx_ref.so: microphone.c's
    tx_filter + CcmPeak extracted at build time + filter.c verbatim) on all host cores, one private copy per core."""
    import ctypes as C
    from oracle import ref_ctypes as R
    from tests.golden.make_golden_tx import mic_audio
    cores = os.cpu_count() or 1
    libs = []
    for _ in range(cores):
        lib = R.load("libquisk_tx_ref.so", private_copy=True)
        lib.ref_tx_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        lib.ref_tx_filter.argtypes = [C.c_void_p, C.c_int]
        lib.ref_tx_init(mode, 48000, preemph, clip)
        libs.append(lib)
    x = np.resize(mic_audio(), block).astype(np.complex128)
    bufs = [np.zeros(2 * block, dtype=np.complex128) for _ in range(cores)]

    def work(i, nsteps):
        for _ in range(nsteps):
            bufs[i][:block] = x
            libs[i].ref_tx_filter(bufs[i].ctypes.data_as(C.c_void_p), block)

    def run(nsteps):
        th = [threading.Thread(target=work, args=(i, nsteps)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return time.perf_counter() - t0

    run(warmup)
    dt = run(steps)
    return cores * block * steps / dt / 1e6, cores, dt


def run_tx(args, ctx, steps, warmup, e2e_steps, cpu_steps):
    """The TX mirror: C transmitters x 48 kS/s microphone audio through tx_filter (USB), 0.25 s blocks."""
    torch, lib = ctx.torch, ctx.lib
    from quisk_b200.rx import TxFilter, load_tables
    from tests.golden.make_golden_tx import mic_audio
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    C_ = args.channels if args.channels > 0 else 4096
    n = 12000
    tx = TxFilter(C_, "USB", load_tables(), mic_sample_rate=48000, preemphasis=0.6, clip=2.5)
    base = torch.from_numpy(np.resize(mic_audio(), n + 101)).to(dev)
    x = torch.stack([base[(c % 101):(c % 101) + n] for c in range(C_)]).to(torch.complex128).contiguous()
    y = torch.zeros((C_, tx.max_out(n) + 8), dtype=torch.complex128, device=dev)
    stream = ctx.stream             # torch's current stream: the copies of the e2e leg and the kernels share it

    def step():
        tx.process(x.data_ptr(), x.stride(0), n, y.data_ptr(), y.stride(0), stream)

    ms, ms_max, launches, clocks = timed_steps(ctx, step, steps, warmup)
    value = world * C_ * n * steps / (ms_max / 1e3) / 1e6
    e2e = None
    if e2e_steps > 0:
        hx = torch.empty((C_, n, 2), dtype=torch.float64).pin_memory(); hx.copy_(torch.view_as_real(x).cpu())
        hy = torch.empty((C_, n, 2), dtype=torch.float64).pin_memory()
        xr = torch.view_as_real(x); yr = torch.view_as_real(y)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            xr.copy_(hx, non_blocking=True)
            step()
            hy.copy_(yr[:, :n], non_blocking=True)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * C_ * n * e2e_steps / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": int(hx.numel() * 8), "d2h_bytes_per_step": int(hy.numel() * 8),
               "channels": C_, "note": "pinned host microphone blocks (complex double, as tx_filter takes them) -> H2D -> quisk_cuda_tx_filter_process -> D2H of the 48 kS/s I/Q"}
        del hx, hy
    tx.close()
    del x, y
    if rank != 0:
        return None
    peak, peak_src = load_peaks()
    alg = 32.0 * C_ * n
    ach = alg * steps / (ms / 1e3) / 1e9
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            v, cores, dt = tx_reference_rate(48000, cpu_steps, 1)
            cpu = {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                   "sample": "%d transmitters (one per host core) x 48000 samples x %d steps, %.1f s wall; oracle/_ref/libquisk_tx_ref.so = microphone.c tx_filter + CcmPeak + filter.c verbatim, gcc -O2" % (cores, cpu_steps, dt)}
        except Exception as ex:
            cpu = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}
    return {"metric": "complex MS/s through RX chain", "value": value, "unit": "MS/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tx_filter: C transmitters x 48 kS/s microphone audio -> microphone.c tx_filter (USB: decimate, band-limit, pre-emphasis, compressor, limiter, peak rounder, x6 interpolation) -> 48 kS/s I/Q (the TX mirror, SURVEY 8(f)4); value counts microphone samples",
                       "channels_per_gpu": C_, "block": n, "mic_rate": 48000,
                       "l2": "input %.0f MB + output %.0f MB per step per GPU >> 126 MB L2, no flush needed" % (C_ * n * 16 / 1e6, C_ * n * 16 / 1e6)},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": "whole step (exact FIR kernels + two per-transmitter gain recurrences at 8 kS/s)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_sample": 32.0},
            "cpu_baseline": cpu}


def run_chain(args, ctx, workload, steps, warmup, e2e_steps, cpu_blocks, fi, fq, tabs, wl_name):
    """rx_chain (BASELINE configs[0] batched), panadapter (configs[1] batched) or both on the same input."""
    torch, lib, L = ctx.torch, ctx.lib, ctx.L
    from quisk_b200.rx import RxChain
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    block = max(FFT_SIZE, (args.block // FFT_SIZE) * FFT_SIZE)
    C_ = args.channels if args.channels > 0 else 4096
    x = synth_block_torch(torch, C_, block, dev, rank)
    rx = pan = None
    stream = ctx.stream
    tune = [args.tune + 3.0 * (c % 101) for c in range(C_)] if args.tune else None
    if "rx_chain" in workload:
        rx = RxChain(C_, RATE[0], "USB", fi, fq, tabs, tune_hz=tune, fused=not args.unfused)
        for opt, val in ((2, args.chunk), (3, args.threads), (4, args.min_r), (8, args.deepk)):
            if val:
                rx.set_option(opt, val)
        for opt, val in ((5, args.dense), (6, args.plans), (11, args.split), (12, args.tailwarp), (19, args.async_load), (20, args.p3)):
            if val >= 0:
                rx.set_option(opt, val)
        if args.nco == "closed":
            rx.set_option(10, 0)
        acap = rx.max_out(block)
        audio = torch.zeros((C_, acap), dtype=torch.float64, device=dev)
    if "panadapter" in workload:
        pan = lib.quisk_cuda_pan_create(C_, FFT_SIZE)
        if not pan:
            raise L.QuiskCudaError(lib.quisk_cuda_last_error().decode())
        graph = torch.zeros((C_, 1024), dtype=torch.float64, device=dev)
    frames = block // FFT_SIZE

    def step():
        if pan:
            L.check(lib, lib.quisk_cuda_pan_accumulate(pan, x.data_ptr(), block, frames, stream), "pan_accumulate")
            L.check(lib, lib.quisk_cuda_pan_graph(pan, 1024, 1.0, 0.0, float(RATE[0]), graph.data_ptr(), stream), "pan_graph")
        if rx:
            rx.process(x.data_ptr(), block, block, audio.data_ptr(), acap, stream=stream)
        if args.sync_steps:
            torch.cuda.synchronize()

    def start_kernel_timing():
        if rx:
            rx.set_option(1, 1); rx.kernel_time()
    ms, ms_max, launches, clocks = timed_steps(ctx, step, steps, warmup, start_kernel_timing)
    kms, kn = (rx.kernel_time() if rx else (0.0, 0))
    kname = rx.fused_kernel_name() if rx else ""
    if rx:
        rx.set_option(1, 0)
    if world > 1:       # per-rank diagnostics (stderr): a rank that is slower than the others shows up here
        print("[rank %d] %s ms_per_step %.4f kernel_ms %.4f clocks %s" % (rank, workload, ms / steps, kms / max(kn, 1), clocks), file=sys.stderr, flush=True)
    value = world * C_ * block * steps / (ms_max / 1e3) / 1e6

    # ---- end to end through the host-buffer entry points, SAME channel count as `value`: H2D + chain + D2H inside the timed region
    e2e = None
    e2e_wire = None
    if rx and e2e_steps > 0:
        hx = torch.empty((C_, block), dtype=torch.complex128).pin_memory()
        hx.copy_(x.cpu())
        ha = torch.zeros((C_, acap), dtype=torch.float64).pin_memory()
        rx_h = RxChain(C_, RATE[0], "USB", fi, fq, tabs, tune_hz=tune, fused=not args.unfused)
        if args.host_chunks >= 0:
            rx_h.set_option(15, args.host_chunks)
        hx_np = hx.numpy(); ha_np = ha.numpy()
        rx_h.process_host(hx_np, block, ha_np)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            na = rx_h.process_host(hx_np, block, ha_np)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * C_ * block * e2e_steps / dt / 1e6, "unit": "MS/s",
               "h2d_bytes_per_step": C_ * block * 16, "d2h_bytes_per_step": C_ * na * 8,
               "channels": C_, "note": "quisk_cuda_rx_process_host: pinned host IQ (complex double) -> H2D -> chain -> D2H audio, per step; same channel count as `value`"}
        # the same step with the host block still in its wire format (int16 little-endian I/Q pairs, what
        # add_rx_samples receives, quisk.c:2922): the H2D copy carries 4 B/sample, the widening runs on the device
        rx_h.reset()
        hw = torch.empty((C_, block, 2), dtype=torch.int16).pin_memory()
        hw.copy_((torch.view_as_real(x) / 65536.0).round().clamp(-32768, 32767).to(torch.int16).cpu())
        hw_np = hw.numpy().view(np.uint8).reshape(C_, block * 4)
        rx_h.process_host_packed(hw_np, block, 2, False, ha_np)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            na = rx_h.process_host_packed(hw_np, block, 2, False, ha_np)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e_wire = {"value": world * C_ * block * e2e_steps / dt / 1e6, "unit": "MS/s",
                    "h2d_bytes_per_step": C_ * block * 4, "d2h_bytes_per_step": C_ * na * 8, "channels": C_,
                    "note": "quisk_cuda_rx_process_host_packed: pinned int16 LE I/Q pairs (add_rx_samples wire format) -> H2D -> unpack -> chain -> D2H audio"}
        rx_h.close()
        del hx, ha, hw
    elif pan and e2e_steps > 0:
        # panadapter end to end: pinned host frames -> H2D -> window/FFT/|X| accumulate + graph -> D2H of the dB graphs
        hx = torch.empty((C_, block, 2), dtype=torch.float64).pin_memory(); hx.copy_(torch.view_as_real(x).cpu())
        hg = torch.empty((C_, 1024), dtype=torch.float64).pin_memory()
        xr = torch.view_as_real(x)
        xr.copy_(hx, non_blocking=True); step(); ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            xr.copy_(hx, non_blocking=True)
            step()
            hg.copy_(graph, non_blocking=True)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * C_ * block * e2e_steps / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": C_ * block * 16,
               "d2h_bytes_per_step": C_ * 1024 * 8, "channels": C_,
               "note": "pinned host frames -> H2D -> quisk_cuda_pan_accumulate + quisk_cuda_pan_graph -> D2H of the 1024-pixel dB graphs"}
        del hx, hg
    if rx:
        rx.close()
    if pan:
        lib.quisk_cuda_pan_destroy(pan)
    del x
    if rank != 0:
        return None

    peak, peak_src = load_peaks()
    roofline = None
    if rx and kn > 0:
        per_launch_ms = kms / kn
        # fused kernel = tune + every half band / FIR down to the filter rate: 16 B in per input sample, 16 B out per Dtot inputs
        # (Dtot = 128 at 1.536 MS/s: 4xHB45 + FIR98/2 + HB45 + FIR98/2; 16 at 192 kS/s)
        dtot = RATE[0] // 12000
        alg = (16.0 + 16.0 / dtot) * C_ * block
        ach = alg / (per_launch_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": kname + " (tune + every half band / FIR of the cascade, %g kS/s -> 12 kS/s)" % (RATE[0] / 1e3), "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": kms / ms,
                    "alg_bytes_per_launch": alg}
        # dram__bytes of one `ncu --set full` capture of this kernel: only quoted when the committed capture names the
        # kernel that actually ran here (it goes stale when the kernel changes; then the key stays null)
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if kname and tj.get("kernel", "").replace(" ", "").find(kname.replace(" ", "")) >= 0:
                roofline["traffic"] = tj["dram_bytes_per_input_sample"] * C_ * block
                roofline["traffic_source"] = "profiles capture %s of %s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum), scaled to this launch" % (tj.get("capture", "?"), kname)
                sb = tj.get("smem_bytes_per_input_sample"); clk = clocks.get("sm_mhz")
                if sb and clk:
                    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                    spk = 128.0 * n_sm * clk * 1e6
                    sach = sb * C_ * block / (per_launch_ms / 1e3)
                    roofline["smem"] = {"achieved": sach / 1e9, "peak": spk / 1e9, "unit": "GB/s", "frac": sach / spk, "bytes_per_input_sample": sb,
                                        "peak_source": "128 B/clk/SM x %d SMs x sampled SM clock; bytes from the same capture (ncu shared-memory wavefronts x 128 B)" % n_sm}
        # C1 sits near the FP64 ridge -- quote the FP64 pipe next to HBM.  FMA instructions per input
        # sample of the cascade: half bands 46 per output at rates 1/2, 1/4, 1/8, 1/16 and 1/64 (43.8), the two 98-tap
        # FIRs at 1/32 and 1/128 (7.7), the tuning phasor (8): 59.5; the pipe's peak is measured on this device.
        pk = C.c_double(0.0)
        if lib.quisk_cuda_fp64_peak(C.byref(pk)) == 0 and pk.value > 0:
            fma_per = 59.5 if RATE[0] == SAMPLE_RATE else None
            fma = (fma_per or 0.0) * C_ * block / (per_launch_ms / 1e3)
            if fma_per:
                roofline["fp64"] = {"achieved": fma / 1e12, "peak": pk.value / 1e12, "unit": "TFMA/s", "frac": fma / pk.value,
                                    "fma_per_input_sample": 59.5, "peak_source": "measured (quisk_cuda_fp64_peak: 8 independent DFMA chains per thread)"}
    elif pan:
        alg = ALG_BYTES_PAN * C_ * block
        ach = alg * steps / (ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "pan_accumulate kernels + pan_graph_kernel (whole step)", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_sample": ALG_BYTES_PAN}

    cpu = None
    if not args.no_cpu_baseline and world == 1:       # a reported baseline: rank 0 at N = 1 only
        try:
            v, cores, dt = cpu_reference_rate(61440, cpu_blocks, 5, workload, fi, fq, args.tune)
            if "rx_chain" in workload:
                cpu = {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference",
                       "sample": "%d channels (one per host core) x 61440 samples x %d blocks, %.1f s wall (%.0f core-seconds); oracle/_ref (filter.c verbatim + quisk.c RX functions, %s)" % (cores, cpu_blocks, dt, dt * cores, REF_FLAGS)}
            else:
                cpu = {"value": v, "unit": "MS/s", "cores": cores, "kind": "port",
                       "sample": "%d streams (one per host core) x 61440 samples x %d blocks, %.1f s wall: the oracle's restatement of get_graph's window / FFT / |X| sum with numpy's pocketfft (FFTW3, the reference's FFT, is absent from this image)" % (cores, cpu_blocks, dt)}
        except Exception as ex:     # the compiled reference did not travel: say so
            cpu = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}

    line = {"metric": "complex MS/s through RX chain", "value": value, "unit": "MS/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "channels_per_gpu": C_, "block": block, "sample_rate": RATE[0], "tune_hz": args.tune, "tune_spread": "receiver c is tuned to tune_hz + 3 (c mod 101) Hz",
                       "fused": not args.unfused, "nco": args.nco if args.tune else "off", "l2": "input %.0f MB per step per GPU >> 126 MB L2, no flush needed" % (C_ * block * 16 / 1e6)},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu}
    if e2e_wire:
        line["e2e_wire"] = e2e_wire
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rate", type=int, default=SAMPLE_RATE, help="rx_chain / panadapter input sample rate")
    ap.add_argument("--workload", default="", choices=["", "rx_chain", "panadapter", "rx_chain+panadapter", "pipeline", "rxa_usb", "rxa_fm", "channelizer", "tx_filter"],
                    help="default: rx_chain as the headline line plus short runs of the other workloads under `workloads`")
    ap.add_argument("--no-extra", action="store_true", help="default run: rx_chain only, no `workloads` key")
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (0 = the workload's own: 4096 / 16 / 64 / 256)")
    ap.add_argument("--block", type=int, default=32768, help="input samples per channel per step (multiple of 8192)")
    ap.add_argument("--tune", type=float, default=12345.0, help="rx_tune_freq in Hz (0 = no tuning stage)")
    ap.add_argument("--unfused", action="store_true", help="run the one-kernel-per-stage exact path instead of the fused cascade")
    ap.add_argument("--chunk", type=int, default=0, help="fused decimator chunk (input samples), 0 = default")
    ap.add_argument("--threads", type=int, default=0, help="fused decimator CTA width (128/256), 0 = default")
    ap.add_argument("--min-r", type=int, default=0, help="fused decimator: min outputs per thread in half-band stages")
    ap.add_argument("--dense", type=int, default=-1, help="fused decimator: 1 = 128-register cap, 0 = 255")
    ap.add_argument("--plans", type=int, default=-1, help="fused decimator: 0 = generic kernel only")
    ap.add_argument("--split", type=int, default=-1, help="fused decimator: 1 = one lane per component in the half bands (default), 0 = complex lanes")
    ap.add_argument("--tailwarp", type=int, default=-1, help="fused decimator: 1 = low-rate stages on a fifth warp (default), 0 = all stages on the four main warps")
    ap.add_argument("--async-load", type=int, default=-1, help="fused decimator (tail-warp kernel): 1 = next chunk by cp.async into stage 0's buffer, 2 = same under a 128-register cap, 0 = register prefetch")
    ap.add_argument("--p3", type=int, default=-1, help="fused decimator: 1 = three-group pipeline kernel, 0 = tail-warp kernel")
    ap.add_argument("--deepk", type=int, default=0, help="fused decimator: low-rate stages every k chunks (1 or 4)")
    ap.add_argument("--nco", default="exact", choices=["exact", "closed"], help="tuning phasor at block starts: the reference's recurrence (default) or closed form only")
    ap.add_argument("--host-chunks", type=int, default=-1, help="host entry points: channel chunks pipelined over copy / compute streams (-1 = library default)")
    ap.add_argument("--rxa-per-block", action="store_true", help="rxa workloads: one xrxa call per DSP block instead of the multi-block entry")
    ap.add_argument("--sync-steps", action="store_true", help="diagnostic: synchronise after every step, so that each step starts on an idle GPU (the host never runs ahead)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    workload = args.workload or "rx_chain"
    if workload == "pipeline":
        # the north star's target pipeline: >= 1024 concurrent 192 kS/s receivers through decimate -> band-pass (cRxFilterOut) ->
        # SSB demodulation, with the panadapter on the same input; 1024 receivers per GPU here
        workload = "rx_chain+panadapter"; args.rate = 192000
        if args.channels <= 0:
            args.channels = 1024
    RATE[0] = args.rate
    wl_names = {"rx_chain": "rx_chain: C x 1.536 MS/s tune->4xHB45->FIR98/2->48k->HB45->FIR98/2->cRxFilterOut(164 I/Q, USB)->audio 48k (BASELINE configs[0], batched)",
                "panadapter": "panadapter: C streams x 8192-pt Hann+FFT+|X| average+dB graph (BASELINE configs[1], batched)",
                "rx_chain+panadapter": "rx_chain + panadapter on the same input (configs[0]+configs[1], batched)"}
    if RATE[0] != SAMPLE_RATE:
        wl_names = {k: v.replace("1.536 MS/s", "%g kS/s" % (RATE[0] / 1e3)).replace("tune->4xHB45->FIR98/2->48k", "tune->decimate->48k") for k, v in wl_names.items()}
        if args.workload == "pipeline":
            wl_names["rx_chain+panadapter"] = "pipeline (north-star target): C x 192 kS/s receivers, tune -> decimate to 48 k -> HB45 -> FIR98/2 -> cRxFilterOut band-pass (164 I/Q) -> USB demod -> audio 48 k, + 8192-pt panadapter on the same input"

    if args.impl == "reference":
        if rank != 0:
            return
        if workload.startswith("rxa_"):
            print(json.dumps(rxa_reference_line(args, dict(RXA_CFG[workload]), 32))); return
        if workload == "channelizer":
            print(json.dumps(pfb_reference_line(args))); return
        from quisk_b200.rx import get_filter_center, load_tables, make_filter_coef
        fi, fq = make_filter_coef(SAMPLE_RATE // 128, None, 2800, get_filter_center("USB", 2800), load_tables())
        ref_block = 61440
        blocks_per_step = 32
        v, cores, dt = cpu_reference_rate(ref_block, args.steps * blocks_per_step, args.warmup * blocks_per_step, workload, fi, fq, args.tune)
        line = {"impl": "reference", "metric": "complex MS/s through RX chain", "value": v, "unit": "MS/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl_names[workload], "channels": cores, "block": ref_block * blocks_per_step, "sample_rate": RATE[0], "tune_hz": args.tune},
                "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": "reference" if "rx_chain" in workload else "port",
                                 "sample": "%d channels (one per host core) x %d blocks of %d samples per step x %d steps; oracle/_ref = filter.c verbatim + quisk.c RX functions, " % (cores, blocks_per_step, ref_block, args.steps) + REF_FLAGS},
                "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line)); return

    ctx = Ctx()
    from quisk_b200.rx import get_filter_center, load_tables, make_filter_coef
    tabs = load_tables()
    # the C1 receive filter: USB, bandwidth 2800 at the 12 kS/s filter rate -> MakeFilterCoef's 164-tap I/Q pair
    fi, fq = make_filter_coef(SAMPLE_RATE // 128, None, 2800, get_filter_center("USB", 2800), tabs)
    if workload.startswith("rxa_"):
        line = run_rxa(args, ctx, workload, args.steps, args.warmup, args.e2e_steps, 8)
    elif workload == "channelizer":
        line = run_pfb(args, ctx, args.steps, args.warmup, args.e2e_steps, 30)
    elif workload == "tx_filter":
        line = run_tx(args, ctx, args.steps, args.warmup, args.e2e_steps, 40)
    else:
        line = run_chain(args, ctx, workload, args.steps, args.warmup, args.e2e_steps, 600, fi, fq, tabs, wl_names[workload])
        if not args.workload and not args.no_extra:
            # the other BASELINE configs, short: each <= a few seconds of GPU time and a bounded CPU sample
            extra = {}
            saved = (args.channels, args.block)
            args.channels, args.block = 0, 32768
            st = min(args.steps, 5)
            args.channels, args.block = 16, 1 << 20         # BASELINE configs[1]: 16 streams, 128 frames of 8192 each per step
            extra["panadapter"] = run_chain(args, ctx, "panadapter", st, 3, 2, 150, fi, fq, tabs, wl_names["panadapter"])
            args.channels, args.block = 0, 32768
            extra["rxa_usb"] = run_rxa(args, ctx, "rxa_usb", st, 3, 1, 3)
            extra["rxa_fm"] = run_rxa(args, ctx, "rxa_fm", st, 3, 1, 3)
            extra["channelizer"] = run_pfb(args, ctx, st, 3, 2, 8)
            extra["tx_filter"] = run_tx(args, ctx, st, 3, 2, 10)
            args.channels, args.block = 1024, 32768
            RATE[0] = 192000
            pname = "pipeline (north-star target): C x 192 kS/s receivers, tune -> decimate to 48 k -> HB45 -> FIR98/2 -> cRxFilterOut band-pass (164 I/Q) -> USB demod -> audio 48 k, + 8192-pt panadapter on the same input"
            extra["pipeline"] = run_chain(args, ctx, "rx_chain+panadapter", st, 3, 2, 100, fi, fq, tabs, pname)
            RATE[0] = args.rate
            args.channels, args.block = saved
            if line is not None:
                line["workloads"] = extra
    if line is not None:
        print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
