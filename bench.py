#!/usr/bin/env python
"""bench.py -- RADLite line ray-tracing hot path on N B200s of one node.

Metric (BASELINE.json): ray-channel integrations/s -- and the wall time of ONE spectrum at 1/2/4/8 GPUs.  A
"ray-channel integration" is one call the reference makes to charintline (telescope.F:3889), i.e. one ray
traced at one velocity channel of one line; both arms count the same units (R).  The workload is
configs[1] unless --config says otherwise: the CO fundamental 100-line LTE spectrum on the 200x80 (r,theta)
grid, 40 351 camera rays, 94 velocity channels per line.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path, one GPU
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference
  torchrun ... bench.py --gpus N ...                        # one rank per GPU: ONE spectrum over N GPUs
  python bench.py --config {1,3,4,5}                       # the other BASELINE configs (5: a 64-line slice)

A step = one pass of the hot path over one batch = ray geometry + all lines of the spectrum.
N = 1:  `value` times it with every input resident in HBM (rl_render_device); `e2e` times rl_set_* + rl_render
        through the C ABI from host buffers to the host flux array (cube for the cube config).
N > 1:  strong scaling -- the SAME spectrum, camera rings cut into N blocks of equal estimated work
        (rl_plan_costs), every rank builds the geometry of its own rays and integrates all lines on them
        (rl_render_rings_device); the one exchange is a sum reduction of the ring sums (disjoint rows: a
        concatenation) over NCCL to rank 0, which does the reference's index-ordered ring sum.  The result
        is compared bit for bit with the one-GPU render.  `replicas` additionally reports N independent
        spectra (weak scaling, no exchange).
The library's opaque-wall start is on, as it is for every caller: R and E are the reference's
counts for the workload; the roofline uses the element integrations the kernels executed.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from radlite_b200 import synth  # noqa: E402

METRIC = "ray-channel integrations/s"
UNIT = "ray-channel integrations/s"
FLOP_PER_ELEMENT = 64.0  # FP64 flop per element integration, exp excluded

# BASELINE.json configs: model keyword overrides, what is rendered, label
CONFIGS = {
    1: dict(kw={}, cube=False,
            label="configs[0]: single CO v=1-0 P(10) line, LTE, 100x40 (r,theta) grid, 25351 rays, 94 channels, "
                  "incl 15 deg, spectrum mode"),
    2: dict(kw={}, cube=False,
            label="configs[1]: CO fundamental 100-line LTE spectrum, 200x80 (r,theta) grid, nphi=150, "
                  "b_extra=-60, 40351 rays, 94 channels (70 km/s @ 1.5 km/s), incl 15 deg, spectrum mode"),
    3: dict(kw={}, cube=True,
            label="configs[2]: 13CO image cube (circular-polar pixels, 199 velocity channels), 400x160 grid, "
                  "70351 rays, incl 45 deg, cube mode"),
    4: dict(kw={}, cube=False,
            label="configs[3]: 12CO two-temperature (NLTE stand-in) populations, every 4.6-5.0 um line of "
                  "v<=9, J<=60 (442 lines, 4 batches), 200x80 grid, dust continuum, spectrum mode"),
    5: dict(kw=dict(nlines=64), cube=False,
            label="configs[4]: 1000x400 grid, line widths x0.2 (velocity sub-gridding on most segments), "
                  "160351 rays; a 64-line slice of the 2000 synthetic lines (per-line cost is the same), "
                  "spectrum mode"),
}


def model_for_bench(cfg=2, nlines=None):
    kw = dict(CONFIGS[cfg]["kw"])
    if nlines:
        kw["nlines"] = nlines
    return synth.config(cfg, **kw)


def input_arrays(m):
    return [m.r, m.theta, m.rho, m.abund, m.vel, m.linewidth, m.lev_up, m.lev_down, m.linefreq,
            m.aud, m.gdeg, m.popul, m.cont_freq_nu, m.kappa_abs, m.kappa_scat, m.dust_rho,
            m.dust_temp, m.starspec_cont]


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference; the Fortran binary cannot be built here)
# ------------------------------------------------------------------------------------------
def _oracle_flags():
    try:
        for ln in open(os.path.join(ROOT, "oracle", "Makefile")):
            if ln.startswith("CFLAGS"):
                return "gcc " + ln.split("=", 1)[1].strip()
    except OSError:
        pass
    return "gcc (flags unknown)"


def _cpu_worker(args):
    cfg, iline, ring_stride, nlines = args
    from oracle.oracle_py import Oracle
    m = model_for_bench(cfg, nlines)
    o = Oracle()
    o.load_model(m)
    if ring_stride > 1:
        o.set_ring_sample(1, m.nrr, ring_stride)
    t = time.perf_counter()
    o.render(iline, 1, m.nfr, m.passband, synth.PARSEC)
    dt = time.perf_counter() - t
    c = o.counters()
    return c["R"], c["E"], dt


def cpu_sample(cfg, ncores, ring_stride, nlines=None):
    """ncores concurrent single-thread processes, one line each (mirrors the drivers' one process
    per line chunk: radlite.py:508,593; line_run.pro:173).  Returns (R, E, slowest process [s], wall [s])."""
    nl = model_for_bench(cfg, nlines).nlines
    lines = [1 + (k * nl) // ncores for k in range(ncores)]
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(ncores) as pool:
        res = pool.map(_cpu_worker, [(cfg, il, ring_stride, nlines) for il in lines])
    wall = time.perf_counter() - t
    tmax = max(r[2] for r in res)
    return sum(r[0] for r in res), sum(r[1] for r in res), tmax, wall


# seconds one host core needs for one line of the config with every ring traced (measured with the oracle;
# only used to size the bounded CPU sample)
CPU_SECONDS_PER_LINE = {1: 5.0, 2: 25.0, 3: 250.0, 4: 25.0, 5: 900.0}

_OUT = sys.stdout


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    ncores = os.cpu_count() or 1
    nsteps = args.steps + args.warmup
    # bounded sample: one line per core per step; thin the rings so the whole run stays ~2-3 min
    budget = 150.0 / max(1, nsteps)
    stride = max(1, int(np.ceil(CPU_SECONDS_PER_LINE[cfg] / budget)))
    Rs, Es, ts = 0.0, 0.0, 0.0
    for s in range(nsteps):
        R, E, tmax, _ = cpu_sample(cfg, ncores, stride, args.lines)
        if s >= args.warmup:
            Rs += R
            Es += E
            ts += tmax
    value = Rs / ts
    nl = model_for_bench(cfg, args.lines).nlines
    sample = (f"{min(ncores, nl)} of {nl} lines per step (one per process, {ncores} processes), every "
              f"{stride}th camera ring of each, all channels")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ts / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIGS[cfg]["label"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "port",
                         "sample": sample, "compiler": _oracle_flags(),
                         "note": "C restatement of the reference path (oracle/); the Fortran binary "
                                 "cannot be built in this image (no Fortran compiler)",
                         "element_integrations_per_s": Es / ts},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples of one GPU.  The process is started before the warm-up (its start-up
    takes longer than a short timed region), every row carries nvidia-smi's own timestamp, and ``stop()`` keeps the
    rows that fall inside the timed region marked by ``mark()`` .. ``stop()``; if the region was shorter than the
    sampling period and holds no row, the rows of the warm-up steps right before it (the same workload) are used and
    ``window`` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=50):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = time.time()
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", str(period_ms), "-i", str(index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark(self):
        """Start of the timed region."""
        self.t0 = time.time()

    @staticmethod
    def _stamp(txt):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        t1 = time.time()
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        stamps = [self._stamp(r[0]) for r in rows]
        inside = [r for r, t in zip(rows, stamps) if t is not None and self.t0 <= t <= t1]
        window = "timed region"
        if not inside:  # shorter than the sampling period: the last warm-up steps (same workload) stand in
            before = [r for r, t in zip(rows, stamps) if t is not None and self.t0 - 1.0 <= t <= t1 + 0.05]
            inside = before or rows[-3:]
            window = "last warm-up steps + timed region (the timed region is shorter than the sampling period)"
        rows = inside
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if "Active" in v and "Not" not in v})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows), "window": window}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from radlite_b200 import shard
    from radlite_b200.api import Renderer

    cfg = args.config
    cube = CONFIGS[cfg]["cube"]
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    m = model_for_bench(cfg, args.lines)
    nl, nfr = m.nlines, m.nfr
    g = Renderer(local)
    g.load_model(m)
    arrays = input_arrays(m)
    h2d = int(sum(a.nbytes for a in arrays))
    nrr, nphi, nray = g.camera_dims()
    d2h = int(nl * nfr * 8) + (int(nl * (nrr + 1) * nphi * nfr * 8) if cube else 0)
    K = args.steps

    # ---- one GPU, inputs resident: the whole spectrum (also the reference for the sharded result) ----
    def step_device():
        g.invalidate_geometry()
        return g.render_device(1, nl, nfr, m.passband, synth.PARSEC)

    strong = world > 1
    sampler = None
    if not strong:
        sampler = ClockSampler(local, args.clock_ms)  # (started here: see the class)
        for _ in range(args.warmup):
            step_device()
        g.reset_counters()
        l0 = g.launch_count()
        barrier()
        sampler.mark()
        t0 = time.perf_counter()
        ms = np.zeros(5)
        for _ in range(K):
            ms += np.array(step_device())
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        launches = g.launch_count() - l0
        cnt = g.counters()
        executed = g.executed_elements()
        flux_dev = g.fetch_flux(nl, nfr)
        step_s = ms[4] * 1e-3 / K  # CUDA-event time per step on the library's stream
        R_step, E_step, ex_step = cnt["R"] / K, cnt["E"] / K, executed / K
        integ_ms = ms[2] / K
        phases = {k: float(v) / K for k, v in
                  zip(("geometry", "prep_select_scan", "integrate", "fill_flux", "total"), ms)}

        def step_e2e():
            g.load_model(m)
            out = g.render(1, nl, nfr, m.passband, synth.PARSEC, want_image=cube)
            return out["flux"]

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            flux = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / K
        assert np.array_equal(flux, flux_dev)
        sharded_equal = None
        blocks = None
        replicas = None
    else:
        # ---- N GPUs: ONE spectrum, camera rings in N blocks of equal estimated work ----
        ref_flux = g.render(1, nl, nfr, m.passband, synth.PARSEC)["flux"] if rank == 0 else None  # untimed
        cost = torch.zeros(nrr + 1, dtype=torch.float64, device=dev)
        if rank == 0:
            cost = torch.from_numpy(g.plan_costs(1, nl, nfr, m.passband)).to(dev)
        dist.broadcast(cost, src=0)
        blocks = shard.split_rings(nrr, world, cost.cpu().numpy())
        lo, hi = blocks[rank]
        rs = torch.zeros((nl, nrr + 1, nfr), dtype=torch.float64, device=dev)
        out = {}

        def step_strong(e2e=False):
            if e2e:
                g.load_model(m)
            else:
                g.invalidate_geometry()
            if hi >= lo:
                t = g.render_rings_device(1, nl, nfr, m.passband, synth.PARSEC, lo, hi, rs.data_ptr())
            else:
                rs.zero_()
                t = [0.0] * 5
            dist.reduce(rs, dst=0, op=dist.ReduceOp.SUM)  # disjoint rows, the others exactly 0: a concatenation
            if rank == 0:
                torch.cuda.current_stream().synchronize()
                out["flux"] = g.flux_from_rings_device(rs.data_ptr(), nl, nfr, synth.PARSEC)
            return t

        # warm-up, with the blocks re-cut from the measured device time per rank: the plan's work estimate is a
        # model; what a rank really needs per unit of estimated work corrects it (shard.rebalance_rings).  The
        # first step only allocates; then at least args.warmup steps, each followed by a correction while the
        # slowest rank is more than 2 % above the mean, at most 10; the cut with the smallest slowest-rank time
        # seen is the one that is timed (one more untimed step with it follows).
        def rank_times(t):
            tk = torch.tensor([t[4]], dtype=torch.float64, device=dev)
            allt = [torch.zeros_like(tk) for _ in range(world)]
            dist.all_gather(allt, tk)
            return np.array([float(x.item()) for x in allt])

        cost_np = cost.cpu().numpy().copy()
        step_strong()
        best = None
        recuts = 0
        for w in range(max(args.warmup, 10)):
            allt = rank_times(step_strong())
            if best is None or allt.max() < best[0]:
                best = (allt.max(), list(blocks))
            if allt.min() <= 0 or (w >= args.warmup - 1 and allt.max() <= 1.02 * allt.mean()):
                break
            cost_np, blocks = shard.rebalance_rings(cost_np, blocks, allt)
            lo, hi = blocks[rank]
            recuts += 1
        blocks = best[1]
        lo, hi = blocks[rank]
        sampler = ClockSampler(local, args.clock_ms) if rank == 0 else None  # (started here: see the class)
        for _ in range(3):
            step_strong()
        g.reset_counters()
        l0 = g.launch_count()
        barrier()
        if sampler:
            sampler.mark()
        t0 = time.perf_counter()
        ms = np.zeros(5)
        for _ in range(K):
            ms += np.array(step_strong())
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        launches = g.launch_count() - l0
        cnt = g.counters()
        executed = g.executed_elements()
        wall = allmax(wall)
        step_s = wall / K  # barrier + synchronize on both sides, max over ranks
        R_step, E_step, ex_step = allsum(cnt["R"]) / K, allsum(cnt["E"]) / K, allsum(executed) / K
        integ_ms = ms[2] / K
        phases = {k: float(v) / K for k, v in
                  zip(("geometry", "prep_select_scan", "integrate", "fill_flux", "total"), ms)}
        phases_max = {k: allmax(v) for k, v in phases.items()}
        sharded_equal = bool(np.array_equal(out["flux"], ref_flux)) if rank == 0 else None
        step_strong(e2e=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_strong(e2e=True)
        barrier()
        e2e_s = allmax(time.perf_counter() - t0) / K
        # N independent spectra (weak scaling of replicas: no exchange at all)
        step_device()
        g.reset_counters()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_device()
        barrier()
        rep_s = allmax(time.perf_counter() - t0) / K
        replicas = {"what": f"{world} independent spectra, one per GPU, no exchange (weak scaling)",
                    "value": allsum(g.counters()["R"]) / K / rep_s, "unit": UNIT, "ms_per_step": 1e3 * rep_s}
        g.reset_counters()

    if rank == 0:
        peak = g.fp64_peak_tflops()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture
        traffic, traffic_src = None, None
        try:
            import glob
            tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r2_cfg{cfg}_*traffic.json")))
            if tfiles and not strong and not args.lines:
                tj = json.load(open(tfiles[-1]))
                traffic, traffic_src = tj["traffic_bytes_per_launch"], os.path.relpath(tfiles[-1], ROOT)
        except (OSError, KeyError, ValueError):
            pass
        value = R_step / step_s
        ex_rank = executed / K  # this rank's executed element integrations per step
        achieved_tf = ex_rank * FLOP_PER_ELEMENT / (integ_ms * 1e-3) / 1e12 if integ_ms > 0 else 0.0
        nodes = g.total_nodes()
        ncell = len(m.r) * len(m.theta)
        # algorithmic HBM bytes of the integrate phase: the per-line cell tables and the
        # line-independent cell fields read once, the spectra (or the cube) written once; the 64-byte node
        # records are this design's own intermediate (built once per camera, re-read by every line tile)
        alg_bytes = (4 * nl + 6) * 8.0 * ncell + (nl * (nrr + 1) * nphi * nfr * 8.0 if cube else nl * nfr * 8.0)
        lb = min(nl, 128)
        kernel = ("ztile_kernel<9> (one warp = one ray x 16 lines across the lanes x 18 channels) + zcont_kernel "
                  "(continuum-only ray x line pairs) + center_kernel" if lb >= 8 else
                  "chan_kernel (one warp = one ray x one line x its channels across the lanes, channel-independent part "
                  "once per node) + center_kernel")
        cpu = None
        if world == 1 and not args.no_cpu:
            ncores = os.cpu_count() or 1
            stride = max(1, int(np.ceil(CPU_SECONDS_PER_LINE[cfg] / 30.0)))
            R, E, tmax, wall_cpu = cpu_sample(cfg, ncores, stride, args.lines)
            cpu = {"value": R / tmax, "unit": UNIT, "cores": ncores, "kind": "port",
                   "sample": f"{min(ncores, nl)} of {nl} lines (one per process, {ncores} processes), every "
                             f"{stride}th camera ring, all channels, {tmax:.1f} s",
                   "compiler": _oracle_flags(),
                   "element_integrations_per_s": E / tmax,
                   "note": "C restatement of the reference path (oracle/); Fortran binary not buildable here"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIGS[cfg]["label"], "lines": nl,
                       "l2": "inputs larger than L2: %.2f GB of ray nodes + %.0f MB of per-line cell "
                             "tables per step, geometry rebuilt every step" % (nodes * 64 / 1e9, nl * ncell * 32 / 1e6),
                       "parallelism": (f"ONE spectrum over {world} GPUs: camera-ring blocks of equal estimated work x "
                                       "all lines, ring sums reduced over NCCL to rank 0" if strong else
                                       "lines x rays independent; 1 GPU"),
                       "ring_blocks": blocks,
                       "opaque_wall": "library default (rl_set_wall_tau 64): ray segments whose contribution is provably "
                                      "below e^-64 of the front-side dust emission are not integrated; R and E are the "
                                      "reference's counts, the roofline uses the executed element integrations"},
            "element_integrations_per_s": E_step / step_s,
            "executed_element_fraction": ex_step / max(1.0, E_step),
            "wall_ms_per_step": 1e3 * wall / K,
            "phase_ms_per_step": phases,
            "e2e": {"value": R_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * world if strong else h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s},
            "gpu_launches": int(launches),
            "wall_time_per_spectrum_ms": 1e3 * step_s,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "kernel": kernel,
                         "how": "EXECUTED element integrations x 64 FP64 flop (SURVEY.md 8d, exp excluded) / CUDA-event "
                                "time of the integrate phase on the library's stream (rank 0); peak = DFMA "
                                "microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry); the "
                                "path is FP64 arithmetic on L2-resident data, the hbm figure is for completeness",
                         "achieved_exp22": ex_rank * (FLOP_PER_ELEMENT + 44.0) / (integ_ms * 1e-3) / 1e12 if integ_ms > 0 else 0.0,
                         "hbm": {"algorithmic_bytes": alg_bytes,
                                 "achieved": alg_bytes / (integ_ms * 1e-3) / 1e9 if integ_ms > 0 else 0.0,
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (integ_ms * 1e-3) / 1e9 / hbm_peak if integ_ms > 0 else 0.0,
                                 "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if strong:
            line["phase_ms_per_step_max_over_ranks"] = phases_max
            line["sharded_result_bitwise_equal_to_one_gpu"] = sharded_equal
            line["replicas"] = replicas
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based)")
    ap.add_argument("--lines", type=int, default=None, help="override the number of lines of the spectrum")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--clock-ms", type=int, default=50, help="nvidia-smi sampling period (B200_PROFILING.md recipe: "
                    "200 ms; 50 so that the 60 ms timed region of an 8-GPU run still holds a sample -- no effect on "
                    "the timing was measured between 100 ms and none at all)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line, the JSON result: anything libraries print there (NCCL's version
    # banner, for one) is routed to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
