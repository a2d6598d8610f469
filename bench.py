#!/usr/bin/env python
"""bench.py -- GenPK P(k) hot path on B200: deposit -> r2c FFT -> |delta_k|^2 binning.

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N ...            the reference's CPU code (oracle/_ref)

One "step" = one pass of the hot path over one synthetic particle set
(gen-pk.cpp:208-234: zero the grid, deposit, FFT, bin, results on the host).
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "end-to-end P(k) throughput, CIC deposit -> r2c FFT -> |delta_k|^2 binning"
UNIT = "Mparticles/s"

# BASELINE.json configs.  `cpu` is the bounded replica the CPU arms time.
WORKLOADS = {
    # configs[1]: synthetic 256^3 DM particles uniform, 512^3 grid, single B200
    "c2": dict(n_side=256, dims=512, kind="uniform", label="synthetic 256^3 uniform-random particles -> 512^3 grid",
               cpu=dict(n_side=256, dims=512)),
    # configs[2]: synthetic 1024^3 clustered (Zel'dovich-displaced) particles, 1024^3 grid, 1/2/4/8 B200
    "c3": dict(n_side=1024, dims=1024, kind="clustered",
               label="synthetic 1024^3 clustered (Zel'dovich-displaced) particles -> 1024^3 grid",
               cpu=dict(n_side=256, dims=256)),
    # configs[4]: synthetic 2048^3 particles on a 2048^3 grid, slab-decomposed across 8 GPUs
    "c5": dict(n_side=2048, dims=2048, kind="clustered", label="synthetic 2048^3 clustered particles -> 2048^3 grid",
               cpu=dict(n_side=256, dims=256)),
    "tiny": dict(n_side=64, dims=128, kind="clustered", label="synthetic 64^3 clustered particles -> 128^3 grid",
                 cpu=dict(n_side=64, dims=128)),
}
BOX = 1000.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GENPK_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--fixed-point", action="store_true", help="deterministic int64 accumulation mode")
    ap.add_argument("--deposit", default="auto", choices=["auto", "direct", "sorted", "march", "sweep"])
    ap.add_argument("--no-sweep", action="store_true", help="lattice input: the march kernel instead of the sweep kernel")
    ap.add_argument("--no-zero-ahead", action="store_true", help="memset the grid instead of clearing it ahead of the sweep's front")
    ap.add_argument("--za-window", type=int, default=0, help="zero ahead: planes kept clear past the expected plane (0 = from the probe)")
    ap.add_argument("--za-slack", type=int, default=-1)
    ap.add_argument("--sweep-ry", type=int, default=0)
    ap.add_argument("--sweep-couple", type=int, default=-1, help="planes a sweep warp may lead the slowest one by (0 = uncoupled)")
    ap.add_argument("--za-zero-ctas", type=int, default=0)
    ap.add_argument("--sweep-couple-step", type=int, default=0)
    ap.add_argument("--sweep-rx", type=int, default=-1, help="lattice planes per sweep task (0 = one persistent sweep)")
    ap.add_argument("--sweep-poll-strong", action="store_true")
    ap.add_argument("--no-self-check", action="store_true", help="skip the comparison with the committed reference fixtures")
    ap.add_argument("--power", default="cached", choices=["cached", "fused"],
                    help="binning pass: geometry sums cached in the context, or recomputed every call")
    ap.add_argument("--lattice-hint", action="store_true", help="pass the lattice extents instead of probing them")
    ap.add_argument("--march-ry", type=int, default=0)
    ap.add_argument("--march-rx", type=int, default=0)
    ap.add_argument("--ghost-planes", type=int, default=16,
                    help="multi-GPU: ghost planes on each side of a slab (0 = always route particles)")
    ap.add_argument("--no-fused-xpass", action="store_true",
                    help="cuFFT x pass + bin_power_kernel instead of the fused x-pass/binning kernel")
    ap.add_argument("--xpass-narrow-tile", action="store_true", help="fused x pass with 4096-mode tiles at 1024 (two CTAs per SM)")
    ap.add_argument("--fft-yz-batch", type=int, default=-1, help="x planes per 2-D cuFFT call (0 = all, -1 = library default)")
    ap.add_argument("--fused-zy", type=int, default=-1, help="GENPK_OPT_FUSED_ZY: 1 one persistent kernel for the z and y passes "
                    "(default), 2 with 8192-mode tiles, 0 library z pass + column kernel")
    ap.add_argument("--zy-lag", type=int, default=0, help="GENPK_OPT_ZY_LAG: planes between a plane's z tiles and its y tiles")
    ap.add_argument("--zero-after", action="store_true", help="fused x pass: store zeros behind the tiles it reads (the next "
                    "step's genpk_grid_zero is free; measured slower than the memset, profiles/r02)")
    ap.add_argument("--no-tma", action="store_true", help="column kernels: per-thread cp.async tile fills instead of TMA bulk tensor copies")
    ap.add_argument("--no-own-ypass", action="store_true", help="cuFFT's 2-D (y,z) plan instead of cuFFT z + own y pass")
    ap.add_argument("--no-ghost-pull", action="store_true", help="multi-GPU: NCCL send/recv of the ghost planes instead of peer loads")
    ap.add_argument("--no-scatter", action="store_true", help="multi-GPU: pack + all-to-all instead of the y pass storing into peers")
    ap.add_argument("--check-mass", action="store_true",
                    help="before timing: one deposit (+ ghost exchange), the grid must sum to the particle count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--xpass-two-pass", action="store_true", help="fused x pass at 1024: two-pass plan, 32 modes per thread, one exchange")
    ap.add_argument("--xpass-halves", action="store_true", help="fused x pass at 1024: two halves of 256 threads out of step on one tile")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-spread", action="store_true", help="rank r uses GPU r instead of GPUs spread over the visible ones")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic_source():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return f"{d.get('report', 'ncu capture')} at commit {d.get('commit', 'unknown')}"
    except Exception:
        return None


def ncu_traffic(workload):
    """DRAM bytes per launch from the committed ncu --set full captures (tools/ncu_traffic.py), or {}."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        return d["kernels"] if d.get("workload") == workload else {}
    except Exception:
        return {}


def algorithmic_bytes(n_particles, dims, nranks=1):
    """particles read once (12 B), padded real grid written once (8 B/cell);
    spectrum read once (16 B/mode).  The (y,z) transform in between: grid read once, spectrum written once.  Per GPU."""
    fd = 2 * (dims // 2 + 1)
    deposit = 12 * n_particles / nranks + 8 * dims * dims * fd / nranks
    binning = 16 * dims * dims * (dims // 2 + 1) / nranks
    return deposit, binning


def fft_algorithmic_bytes(dims, nranks=1):
    return 2 * 8 * dims * dims * (2 * (dims // 2 + 1)) / nranks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.lines, self.gpu = None, [], gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def host_particles(kind, n_side, dims, seed=42):
    """Host-side synthetic particles of the same distributions as genpk_synth_particles
    (statistically, not bit for bit) for the CPU arms."""
    n = n_side ** 3
    if kind == "uniform":
        rng = np.random.default_rng(seed)
        return (rng.random((n, 3), dtype=np.float32) * np.float32(BOX))
    q1 = ((np.arange(n_side) + 0.5) / n_side)
    q = np.stack(np.meshgrid(q1, q1, q1, indexing="ij"), axis=-1).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    modes = []
    while len(modes) < 32:
        v = rng.integers(-8, 9, 3)
        ln = np.sqrt((v ** 2).sum())
        if 0.5 < ln <= 8.0:
            modes.append(v)
    modes = np.array(modes, np.float64)
    ln = np.sqrt((modes ** 2).sum(1))
    amp = ln ** -1.5
    amp *= (2.0 / dims) / np.sqrt((amp ** 2).sum() / 6.0)           # per-component rms of ~2 cells
    phase = rng.random(32)
    disp = np.zeros_like(q)
    for m in range(32):
        s = np.sin(2 * np.pi * (q @ modes[m] + phase[m]))
        disp += (amp[m] * modes[m] / ln[m])[None, :] * s[:, None]
    x = (q + disp) % 1.0
    return (x * BOX).astype(np.float32)


_HOST_CACHE = {}


def time_cpu_path(kind, wl, reps=1, want="reference"):
    """Times the reference's CPU implementation (oracle/_ref when built, else the oracle
    port) on wl['cpu'] = a bounded replica of the workload with the same particles per
    cell and distribution.  Returns per-stage seconds (best of reps)."""
    from oracle.oracle import Oracle, have_reference, padded_shape, rfftn_padded
    backend = "reference" if (want == "reference" and have_reference()) else "port"
    orc = Oracle(backend)
    n_side, dims = wl["cpu"]["n_side"], wl["cpu"]["dims"]
    key = (kind, n_side, dims)
    if key not in _HOST_CACHE:                       # generated once, outside every timed region
        _HOST_CACHE.clear()
        _HOST_CACHE[key] = host_particles(kind, n_side, dims)
    pos = _HOST_CACHE[key]
    n = n_side ** 3
    best = None
    for _ in range(reps):
        field = np.zeros(padded_shape(dims))
        t0 = time.perf_counter()
        orc.fieldize(BOX, dims, field, pos, None, 1.0, 1)
        t1 = time.perf_counter()
        spec = rfftn_padded(field, dims)
        t2 = time.perf_counter()
        orc.powerspectrum(dims, spec, None, dims, float(n), float(n))
        t3 = time.perf_counter()
        cur = dict(deposit=t1 - t0, fft=t2 - t1, binning=t3 - t2, total=t3 - t0)
        if best is None or cur["total"] < best["total"]:
            best = cur
    threads = int(os.environ.get("OMP_NUM_THREADS", cpu_threads()))
    return dict(kind=backend, n=n, n_side=n_side, dims=dims, cores=threads, **best)


def check_against_fixtures(workload, dims, power, cnt, keffs):
    """The run's own P(k) next to the committed reference fixtures: mode counts must equal the golden table
    (tests/golden/mode_counts.npz, from the reference's bin rule) and, for the configs that have a full-size
    fixture made from the reference's object code (tests/golden/fullsize_pk.npz), P(k) and k_eff must agree
    to 1e-5 relative.  A mismatch fails the run."""
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    try:
        mc = np.load(os.path.join(gold, "mode_counts.npz"))
        key = f"count{dims}"
        if key in mc.files:
            assert np.array_equal(cnt.astype(np.int64), mc[key]), "mode counts differ from the golden table"
            out["counts"] = f"bit-exact vs tests/golden/mode_counts.npz:{key}"
    except OSError:
        pass
    try:
        fs = np.load(os.path.join(gold, "fullsize_pk.npz"))
        if f"{workload}_power" in fs.files:
            pr, cr, kr = fs[f"{workload}_power"], fs[f"{workload}_count"], fs[f"{workload}_keffs"]
            assert np.array_equal(cnt.astype(np.int64), cr.astype(np.int64)), "mode counts differ from the reference fixture"
            nz = cr > 0
            dp = float(np.max(np.abs(power[nz] / pr[nz] - 1.0)))
            dk = float(np.max(np.abs(keffs[nz] / kr[nz] - 1.0)))
            assert dp <= 1e-5 and dk <= 1e-5, f"P(k) differs from the reference fixture: max rel {dp:.3e} (k_eff {dk:.3e})"
            out["pk"] = {"fixture": f"tests/golden/fullsize_pk.npz:{workload} (reference objects at full size)",
                         "max_rel_power": dp, "max_rel_keff": dk, "tolerance": 1e-5}
    except OSError:
        pass
    return out or None


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference's OpenMP code on
    ALL host cores.  Must run before the oracle's shared objects (libgomp) are loaded."""
    n = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def bind_to_gpu_numa(dev_index):
    """Run this rank's host side (and first-touch its page-locked shard) on the NUMA node the GPU hangs off, so that
    N ranks uploading at once do not all pull their shards across the socket link.  Returns what was done."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devid = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        if node < 0 or len(nodes) < 2:
            return f"single NUMA node ({len(nodes)} listed, GPU reports {node})"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: none of its CPUs are available to this process"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} CPUs)"
    except Exception as e:                                       # diagnostics only: never fail a run over this
        return f"unavailable ({type(e).__name__})"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------
# the reference arm: the reference's own CPU code on the box's host cores
# ----------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    use_all_host_threads()
    times = []
    for i in range(args.warmup + args.steps):
        t = time_cpu_path(wl["kind"], wl, reps=1)
        if i >= args.warmup:
            times.append(t)
    tot = float(np.mean([t["total"] for t in times]))
    n = times[0]["n"]
    value = n / tot / 1e6
    sample = (f"{wl['cpu']['n_side']}^3 {wl['kind']} particles -> {wl['cpu']['dims']}^3 grid per step "
              f"(same particles per cell and distribution as the workload); "
              f"FFT stage = pocketfft stand-in, FFTW3 is absent from the image")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": times[0]["cores"], "kind": times[0]["kind"],
                         "sample": sample, "cpu_model": cpu_model(),
                         "stage_s": {k: float(np.mean([t[k] for t in times])) for k in ("deposit", "fft", "binning")}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import genpk_b200 as gp
    from genpk_b200 import api
    from genpk_b200.distributed import CudaStages, SlabPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    # Ranks take GPUs spread evenly over the visible ones (N = 4 of 8: 0, 2, 4, 6): NVLink is uniform, but host memory
    # reaches the GPUs through shared PCIe bridges, and the e2e leg uploads from every rank at once (measured on this
    # pool: 53 / 105 / 111 / 180 GB/s aggregate with GPUs 0..N-1 at N = 1 / 2 / 4 / 8).
    visible = torch.cuda.device_count()
    stride = visible // world if (not args.no_spread and world > 1 and visible >= world and visible % world == 0) else 1
    local = local * stride
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    n_side, dims = wl["n_side"], wl["dims"]
    n_total = n_side ** 3
    nrbins = dims                                                   # gen-pk.cpp:173
    kind = {"uniform": api.SYNTH_UNIFORM_RANDOM, "clustered": api.SYNTH_CLUSTERED}[wl["kind"]]
    flags = api.FLAG_FIXED_POINT if args.fixed_point else 0
    mode = {"auto": api.DEPOSIT_AUTO, "direct": api.DEPOSIT_DIRECT, "sorted": api.DEPOSIT_SORTED,
            "march": api.DEPOSIT_MARCH, "sweep": api.DEPOSIT_SWEEP}[args.deposit]

    # this rank's shard: a contiguous index range of the set (an x-slab of the lattice for the
    # lattice-ordered kinds, an arbitrary subset for the random kind)
    first = rank * n_total // world
    count = (rank + 1) * n_total // world - first
    dpos = torch.empty(3 * count, dtype=torch.float32, device=dev)
    api.synth_particles_dev(kind, 42, n_side, first, count, BOX, dims, dpos.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    total_mass = float(n_total)

    if world == 1:
        ctx = gp.Context(dims, local, flags)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_deposit_mode(mode)
        ctx.set_power_mode(api.POWER_FUSED if args.power == "fused" else api.POWER_CACHED)
        pipe = None

        def step_device():
            ctx.grid_zero()
            ctx.deposit_dev(dpos.data_ptr(), count, 0, 1.0, BOX)
            return ctx.fft_power(nrbins, total_mass, total_mass)      # raw sums D2H + normalisation inside

        def step_host(hpos):
            return ctx.pk_from_particles_ptr(hpos.data_ptr(), count, 1.0, BOX, total_mass, nrbins)
    else:
        # wide-ghost slabs: index-range shards of lattice-ordered sets are deposited where they are
        ghost = min(args.ghost_planes, dims // world, (dims - dims // world) // 2)
        stages = CudaStages(dims, world, rank, dev, flags, ghost)
        stages.ctx.set_deposit_mode(mode)
        stages.ctx.set_power_mode(api.POWER_FUSED if args.power == "fused" else api.POWER_CACHED)
        ctx = stages.ctx
        pipe = SlabPipeline(dims, stages)
        if not (args.no_scatter or args.no_fused_xpass or args.no_own_ypass):
            stages.enable_scatter()
        if not args.no_ghost_pull:
            stages.enable_ghost_pull()

        def step_device():
            return pipe.pk(dpos, None, 1.0, BOX, total_mass, nrbins)

        def step_host(hpos):
            # the pinned shard goes up in chunks on the library's copy stream while earlier chunks are deposited
            return pipe.pk(hpos, None, 1.0, BOX, total_mass, nrbins)

    if args.no_fused_xpass:
        ctx.set_option(api.OPT_FUSED_XPASS, 0)
    elif args.xpass_narrow_tile:
        ctx.set_option(api.OPT_FUSED_XPASS, 2)
    elif args.xpass_halves:
        ctx.set_option(api.OPT_FUSED_XPASS, 3)
    elif args.xpass_two_pass:
        ctx.set_option(api.OPT_FUSED_XPASS, 4)
    if args.no_own_ypass:
        ctx.set_option(api.OPT_OWN_YPASS, 0)
    if args.no_tma:
        ctx.set_option(api.OPT_TMA, 0)
    if args.zero_after:
        ctx.set_option(api.OPT_ZERO_AFTER_POWER, 1)
    if args.fused_zy >= 0:
        ctx.set_option(api.OPT_FUSED_ZY, args.fused_zy)
    if args.zy_lag > 0:
        ctx.set_option(api.OPT_ZY_LAG, args.zy_lag)
    if args.fft_yz_batch >= 0:
        ctx.set_option(api.OPT_FFT_YZ_BATCH, args.fft_yz_batch)
    fused = ctx.fused_xpass_supported(nrbins)
    if args.lattice_hint and wl["kind"] != "uniform":
        ctx.set_lattice_hint(n_side, n_side)
    if args.no_sweep:
        ctx.set_option(api.OPT_SWEEP, 0)
    if args.no_zero_ahead:
        ctx.set_option(api.OPT_ZERO_AHEAD, 0)
    if args.za_window:
        ctx.set_option(api.OPT_ZA_WINDOW, args.za_window)
    if args.za_slack >= 0:
        ctx.set_option(api.OPT_ZA_SLACK, args.za_slack)
    if args.sweep_ry:
        ctx.set_option(api.OPT_SWEEP_RY, args.sweep_ry)
    if args.sweep_couple >= 0:
        ctx.set_option(api.OPT_SWEEP_COUPLE, args.sweep_couple)
    if args.za_zero_ctas:
        ctx.set_option(api.OPT_ZA_ZERO_CTAS, args.za_zero_ctas)
    if args.sweep_rx >= 0:
        ctx.set_option(api.OPT_SWEEP_RX, args.sweep_rx)
    if args.sweep_couple_step:
        ctx.set_option(api.OPT_SWEEP_COUPLE_STEP, args.sweep_couple_step)
    if args.sweep_poll_strong:
        ctx.set_option(api.OPT_SWEEP_POLL_WEAK, 0)
    if args.march_ry:
        ctx.set_option(api.OPT_MARCH_RY, args.march_ry)
    if args.march_rx:
        ctx.set_option(api.OPT_MARCH_RX, args.march_rx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.check_mass:
        from genpk_b200.distributed import _DevMem
        if world == 1:
            ctx.grid_zero()
            ctx.deposit_dev(dpos.data_ptr(), count, 0, 1.0, BOX)
        else:
            pipe.deposit(dpos, None, 1.0, BOX, 0, True, False)
            pipe.exchange_ghost(0)
        fd = 2 * (dims // 2 + 1)
        owned = (dims // world) * dims * fd
        off = ctx.owned_offset()
        g = torch.as_tensor(_DevMem(ctx.grid_ptr() + 8 * off, 8 * owned), device=dev)
        if args.fixed_point:
            tot = g.view(torch.int64).view(-1, fd).sum(dim=1).to(torch.float64).sum() / 2.0 ** 40
        else:
            tot = g.view(-1, fd).sum(dim=1).sum()
        tot = tot.reshape(1).clone()
        if world > 1:
            dist.all_reduce(tot)
        rel = abs(float(tot.item()) / n_total - 1.0)
        assert rel < 1e-9, f"deposited mass {float(tot.item())!r} != {n_total} particles (rel {rel:.3e})"
        if rank == 0:
            print(f"mass check: grid sums to {float(tot.item()):.6f} for {n_total} particles (rel {rel:.2e})", file=sys.stderr, flush=True)

    # ---- device-resident timing ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    ctx.stage_reset()
    if pipe is not None:
        pipe.profile = True
    launches0 = ctx.launch_count()
    libcalls0 = ctx.library_calls()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = ctx.launch_count() - launches0
    libcalls = ctx.library_calls() - libcalls0
    ms_step = ms_total / args.steps
    value = n_total / (ms_step * 1e-3) / 1e6
    stage_ms = {}
    for name, st in (("zero", api.STAGE_ZERO), ("deposit", api.STAGE_DEPOSIT), ("sort", api.STAGE_SORT),
                     ("fft", api.STAGE_FFT), ("binning", api.STAGE_POWER)):
        tot, nrec = ctx.stage_total_ms(st)
        stage_ms[name] = tot / args.steps if nrec else 0.0
    ctx.synchronize()
    if pipe is not None:
        pipe.profile = False
        for k, v in pipe.collect_timings().items():
            stage_ms[k] = v / args.steps                     # exchange steps of rank 0 (device time)
    power, cnt, keffs = out
    assert int(cnt.astype(np.int64).sum()) == dims ** 3 - 1, "mode counts do not sum to dims^3-1"
    self_check = None if args.no_self_check else check_against_fixtures(args.workload, dims, power, cnt, keffs)

    host_affinity = bind_to_gpu_numa(local) if (world > 1 and not args.no_e2e) else None
    # ---- end to end from pinned host memory ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        k_e2e = args.e2e_steps or min(args.steps, 5)
        hpos = torch.empty(3 * count, dtype=torch.float32, pin_memory=True)
        hpos.copy_(dpos)
        torch.cuda.synchronize()
        step_host(hpos)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(k_e2e):
            out2 = step_host(hpos)
        a1.record()
        barrier()
        ms_e2e = max_over_ranks(a0.elapsed_time(a1)) / k_e2e
        assert np.array_equal(out2[1], cnt)
        e2e = {"value": n_total / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_e2e, "steps": k_e2e,
               "h2d_bytes_per_step": 12 * n_total, "d2h_bytes_per_step": 3 * nrbins * 8 * world,
               "api": "genpk_pk_from_particles (host float32 positions in, power/count/keffs out)" if world == 1
               else "SlabPipeline.pk on pinned host shards"}
        if host_affinity:
            e2e["host_affinity"] = host_affinity
        del hpos

    # ---- roofline of the dominant kernel of ours ----------------------------------------------
    peak, peak_src = measured_peaks()
    dep_bytes, bin_bytes = algorithmic_bytes(n_total, dims, world)
    roof = {}
    # the deposit's algorithmic bytes include writing the grid once, so its time includes zeroing it:
    # genpk_grid_zero is lazy -- the grid is cleared inside the deposit stage, by the sweep kernel itself
    # (zero ahead) or by a memset (stage_ms["zero"], a part of stage_ms["deposit"])
    stage_ms["deposit_with_zero"] = stage_ms["deposit"]
    for name, nbytes in (("deposit", dep_bytes), ("binning", bin_bytes)):
        ms = stage_ms["deposit_with_zero"] if name == "deposit" else stage_ms[name]
        ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        roof[name] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "traffic": None, "ms": ms, "algorithmic_bytes": nbytes}
    if stage_ms.get("fft", 0) > 0:
        nb = fft_algorithmic_bytes(dims, world)
        ach = nb / (stage_ms["fft"] * 1e-3) / 1e9
        roof["fft_yz"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                          "ms": stage_ms["fft"], "algorithmic_bytes": nb}
    # DRAM traffic of the stage's kernels from the committed ncu capture of this workload (single GPU only)
    tr = ncu_traffic(args.workload) if world == 1 else {}
    if "deposit_march_kernel" in tr and ctx.last_order().get("lattice"):
        k = tr["deposit_march_kernel"]
        roof["deposit"]["traffic"] = k["dram_read_bytes"] + k["dram_write_bytes"] + 8 * dims * dims * (2 * (dims // 2 + 1))
        roof["deposit"]["traffic_note"] = "deposit_march_kernel (ncu) + the memset's writes (8 B/cell)"
    bk = "fftx_power_kernel" if fused else "bin_power_kernel"
    if bk in tr:
        roof["binning"]["traffic"] = tr[bk]["dram_read_bytes"] + tr[bk]["dram_write_bytes"]
    if "fft_yz" in roof and "fft_zy_kernel" in tr and libcalls == 0:
        roof["fft_yz"]["traffic"] = tr["fft_zy_kernel"]["dram_read_bytes"] + tr["fft_zy_kernel"]["dram_write_bytes"]
    if tr:
        for r in roof.values():
            if r.get("traffic") is not None:
                r["traffic_source"] = ncu_traffic_source()
    dom = "deposit" if stage_ms["deposit_with_zero"] >= stage_ms["binning"] else "binning"
    roofline = dict(roof[dom])
    roofline["kernel"] = {"deposit": "deposit stage (grid zero + order probe + deposit_march_kernel | brick sort + "
                                     "deposit_direct_kernel)",
                          "binning": ("fftx_power_kernel (x pass of the FFT fused with the binning)" if fused
                                      else "bin_power_kernel")}[dom]
    roofline["peak_source"] = peak_src

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "particles": n_total, "grid": dims, "nrbins": nrbins, "box": BOX,
                   "accumulation": "int64 fixed-point" if args.fixed_point else "fp64 red.add",
                   "deposit_mode": args.deposit, "order_probe": ctx.last_order(), "sweep": ctx.last_sweep(),
                   "binning_mode": args.power,
                   "x_pass": "fused with binning (fftx_power_kernel)" if fused else "cuFFT", "parallelism": (f"x-slab x{world}, {ghost} ghost planes, particles {pipe.placement}, transpose "
                                   + ("fused into the y pass (peer stores)" if stages.scatter_ready else "pack + all-to-all")
                                   + (", ghost planes pulled over NVLink (peer loads)" if stages.pull_ready else ", ghost planes by send/recv")
                                   if world > 1
                                   else "single GPU"),
                   "l2": "inputs larger than L2 (no flush needed)"},
        "pk_time_ms": ms_step,
        "stage_ms": stage_ms,
        "deposit_mparticles_per_s": (n_total / (stage_ms["deposit_with_zero"] * 1e-3) / 1e6) if stage_ms["deposit"] else None,
        "binning_gcells_per_s": (dims ** 2 * (dims // 2 + 1) / (stage_ms["binning"] * 1e-3) / 1e9)
        if stage_ms["binning"] else None,
        "roofline": roofline, "roofline_all": roof,
        "gpu_launches": int(launches), "cufft_execs_per_step": libcalls / args.steps,
        "e2e": e2e, "clocks": clocks, "self_check": self_check,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        t = time_cpu_path(wl["kind"], wl, reps=2)
        line["cpu_baseline"] = {
            "value": t["n"] / t["total"] / 1e6, "unit": UNIT, "cores": t["cores"], "kind": t["kind"],
            "sample": f"{t['n_side']}^3 {wl['kind']} particles -> {t['dims']}^3 grid (bounded replica, same particles "
                      f"per cell); FFT = pocketfft stand-in (FFTW3 absent)",
            "cpu_model": cpu_model(), "stage_s": {k: t[k] for k in ("deposit", "fft", "binning")}}
    if world > 1:
        line["config"]["gpus_used"] = [r * stride for r in range(world)]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        stages.close()
        dist.destroy_process_group()
    else:
        ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
