#!/usr/bin/env python
"""bench.py -- RADLite line ray-tracing hot path on N B200s of one node.

Metric (BASELINE.json): ray-channel integrations/s on the CO fundamental 100-line LTE spectrum
(configs[1]: 200x80 (r,theta) grid, 40 351 camera rays, 94 velocity channels per line).  A "ray-
channel integration" is one call the reference makes to charintline (telescope.F:3889), i.e. one
ray traced at one velocity channel of one line; both arms count the same units (R).

  python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W  # CPU restatement of the reference
  torchrun ... bench.py --gpus N ...                      # one rank per GPU, lines are sharded

A step = one pass of the hot path over one batch = ray geometry + all lines of the rank's
spectrum.  `value` times it with every input resident in HBM (rl_render_device); `e2e` times
rl_set_* + rl_render through the C ABI from host buffers to the host flux array.
Multi-GPU is weak scaling: every rank renders one full 100-line spectrum (lines and rays are
independent: no data-path collective, only the final gather of the spectra to rank 0).  The
BASELINE "wall time per 100-line spectrum at N GPUs" (strong scaling, lines block-partitioned like
radlite.py:1163-1169) is measured in the same run and reported under "strong".
The library's opaque-wall start is on, as it is for every caller: R (ray-channel
integrations, the metric's unit) is the reference's count for the workload, the roofline uses the element
integrations the kernels executed, `roofline.*_reference_work` the reference's element count.
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
WORKLOAD = ("configs[1]: CO fundamental 100-line LTE spectrum, 200x80 (r,theta) grid, nphi=150, "
            "b_extra=-60, 40351 rays, 94 channels (70 km/s @ 1.5 km/s), incl 15 deg, spectrum mode")


def model_for_bench(nlines=100):
    return synth.config(2, nlines=nlines)


def input_arrays(m):
    return [m.r, m.theta, m.rho, m.abund, m.vel, m.linewidth, m.lev_up, m.lev_down, m.linefreq,
            m.aud, m.gdeg, m.popul, m.cont_freq_nu, m.kappa_abs, m.kappa_scat, m.dust_rho,
            m.dust_temp, m.starspec_cont]


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference; the Fortran binary cannot be built here)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    iline, ring_stride, nlines = args
    from oracle.oracle_py import Oracle
    m = model_for_bench(nlines)
    o = Oracle()
    o.load_model(m)
    if ring_stride > 1:
        o.set_ring_sample(1, m.nrr, ring_stride)
    t = time.perf_counter()
    o.render(iline, 1, m.nfr, m.passband, synth.PARSEC)
    dt = time.perf_counter() - t
    c = o.counters()
    return c["R"], c["E"], dt


def cpu_sample(ncores, ring_stride, nlines=100):
    """ncores concurrent single-thread processes, one line each (mirrors the drivers' one process
    per line chunk: radlite.py:508,593; line_run.pro:173).  Returns (R, E, wall seconds)."""
    lines = [1 + (k * nlines) // ncores for k in range(ncores)]
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(ncores) as pool:
        res = pool.map(_cpu_worker, [(il, ring_stride, nlines) for il in lines])
    wall = time.perf_counter() - t
    tmax = max(r[2] for r in res)
    return sum(r[0] for r in res), sum(r[1] for r in res), tmax, wall


_OUT = sys.stdout


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    nsteps = args.steps + args.warmup
    # bounded sample: one line per core per step; thin the rings so the whole run stays ~2-3 min
    per_line_s = 25.0
    budget = 150.0 / max(1, nsteps)
    stride = max(1, int(np.ceil(per_line_s / budget)))
    vals, Rs, Es, ts = [], 0.0, 0.0, 0.0
    for s in range(nsteps):
        R, E, tmax, _ = cpu_sample(ncores, stride)
        if s >= args.warmup:
            Rs += R
            Es += E
            ts += tmax
            vals.append(R / tmax)
    value = Rs / ts
    sample = (f"{ncores} of 100 lines per step (one per process), every {stride}th camera ring of "
              f"each, all channels")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ts / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "port",
                         "sample": sample,
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
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
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
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if "Active" in v and "Not" not in v})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from radlite_b200.api import Renderer

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = model_for_bench(args.lines)
    nl, nfr = m.nlines, m.nfr
    g = Renderer(local)
    g.load_model(m)
    arrays = input_arrays(m)
    h2d = int(sum(a.nbytes for a in arrays))
    d2h = int(nl * nfr * 8)

    # ---- kernel-only (inputs resident) ----
    def step_device():
        g.invalidate_geometry()
        return g.render_device(1, nl, nfr, m.passband, synth.PARSEC)

    for _ in range(args.warmup):
        step_device()
    g.reset_counters()
    l0 = g.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    ms = np.zeros(5)
    for _ in range(args.steps):
        ms += np.array(step_device())
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = g.launch_count() - l0
    cnt = g.counters()
    executed = g.executed_elements()  # element integrations the kernels performed (opaque-wall start skips some)
    flux_dev = g.fetch_flux(nl, nfr)
    dev_s = ms[4] * 1e-3  # CUDA-event time of the K steps on the library's stream

    # ---- end to end through the C ABI: host buffers in, host flux out ----
    def step_e2e():
        g.load_model(m)
        return g.render(1, nl, nfr, m.passband, synth.PARSEC)["flux"]

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flux = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(flux, flux_dev)

    # ---- strong scaling: ONE 100-line spectrum block-partitioned over the ranks ----
    per = -(-nl // world)  # ceil, like subN = ceil(nlines/ncores) (line_run.pro:79)
    i0 = min(nl, rank * per)
    n_loc = max(0, min(per, nl - i0))

    def step_strong():
        g.invalidate_geometry()
        if n_loc:
            g.render_device(i0 + 1, n_loc, nfr, m.passband, synth.PARSEC)
        if world > 1:  # the only exchange of the path: gather the spectra on rank 0
            loc = torch.zeros((per, nfr), dtype=torch.float64, device=dev)
            if n_loc:
                loc[:n_loc] = torch.from_numpy(g.fetch_flux(n_loc, nfr)).to(dev)
            parts = [torch.zeros_like(loc) for _ in range(world)] if rank == 0 else None
            dist.gather(loc, parts, dst=0)

    step_strong()  # (untimed: also opens NCCL's connections for the gather)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_strong()
    barrier()
    strong_s = time.perf_counter() - t0

    # ---- reduce over ranks: max time, sum of work ----
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

    dev_s_max, wall_max, e2e_max, strong_max = allmax(dev_s), allmax(wall), allmax(e2e_s), allmax(strong_s)
    R_tot, E_tot = allsum(cnt["R"]), allsum(cnt["E"])
    integ_s = allmax(ms[2] * 1e-3)
    if world > 1:
        flux_all = [torch.zeros((nl, nfr), dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(flux_dev).to(dev), flux_all, dst=0)

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
            tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
            if tfiles and nl == 100:
                tj = json.load(open(tfiles[-1]))
                traffic, traffic_src = tj["traffic_bytes_per_launch"], os.path.relpath(tfiles[-1], ROOT)
        except (OSError, KeyError, ValueError):
            pass
        value = R_tot / dev_s_max
        # dominant kernel: ztile_kernel, one launch per step (100 lines fit one batch)
        n_launch = args.steps * max(1, -(-nl // max(1, nl)))
        # the roofline counts the element integrations the kernel EXECUTED; the reference's own count E (it
        # integrates the segments behind opaque dust too) is reported next to it as *_reference_work
        E_rank = executed
        achieved_tf = E_rank * FLOP_PER_ELEMENT / (ms[2] * 1e-3) / 1e12
        achieved_ref_tf = cnt["E"] * FLOP_PER_ELEMENT / (ms[2] * 1e-3) / 1e12
        nodes = g.total_nodes()
        # algorithmic HBM bytes of one integrate launch: node lists once + per-line cell tables +
        # image written once (DESIGN.md "Kernels")
        alg_bytes = nodes * 60.0 + nl * len(m.r) * len(m.theta) * 32.0 + nl * (m.nrr + 1) * m.nphi * nfr * 8.0
        cpu = None
        if world == 1 and not args.no_cpu:
            ncores = os.cpu_count() or 1
            R, E, tmax, wall_cpu = cpu_sample(ncores, 1, args.lines)
            cpu = {"value": R / tmax, "unit": UNIT, "cores": ncores, "kind": "port",
                   "sample": f"{ncores} of {nl} lines (one per process, all rays, all channels), "
                             f"{tmax:.1f} s",
                   "element_integrations_per_s": E / tmax,
                   "note": "C restatement of the reference path (oracle/); Fortran binary not buildable here"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "lines_per_gpu": nl,
                       "l2": "inputs larger than L2: %.2f GB of ray nodes + %.0f MB of per-line cell "
                             "tables per step, geometry rebuilt every step" % (nodes * 60 / 1e9, nl * len(m.r) * len(m.theta) * 32 / 1e6),
                       "parallelism": f"lines x rays independent; {world} rank(s), one spectrum each",
                       "opaque_wall": "library default: ray segments behind tau_dust > 150 (seen from the observer) "
                                      "are not integrated; image bit-identical to the full walk (DESIGN.md 4.3); "
                                      "value counts the reference's ray-channel integrations, the roofline the "
                                      "executed element integrations"},
            "element_integrations_per_s": E_tot / dev_s_max,
            "executed_element_fraction": executed / max(1.0, cnt["E"]),
            "wall_ms_per_step": 1e3 * wall_max / args.steps,
            "phase_ms_per_step": {k: float(v) / args.steps for k, v in
                                  zip(("geometry", "prep_select_scan", "integrate", "fill_flux", "total"), ms)},
            "e2e": {"value": R_tot / e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_max / args.steps},
            "gpu_launches": int(launches),
            "strong": {"what": "wall time of ONE %d-line spectrum block-partitioned over %d GPU(s) incl. "
                               "final gather" % (nl, world), "ms": 1e3 * strong_max / args.steps},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src,
                         "kernel": "ztile_kernel<9> (formal solution: one warp = one ray x 16 lines across the lanes x 18 channels)",
                         "how": "EXECUTED element integrations x 64 FP64 flop (SURVEY.md 8d, exp excluded) / CUDA-event "
                                "time of the integrate phase (ztile_kernel + centre ray) on the library's stream; "
                                "peak = DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no "
                                "FP64 entry); the path is FP64 arithmetic on L2-resident data, so the hbm "
                                "figure below is reported for completeness only; *_reference_work counts the "
                                "element integrations the reference performs for the same result (it also "
                                "integrates the segments behind tau_dust > 150, which ztile_kernel skips)",
                         "achieved_reference_work": achieved_ref_tf, "frac_reference_work": achieved_ref_tf / peak,
                         "achieved_exp22": E_rank * (FLOP_PER_ELEMENT + 44.0) / (ms[2] * 1e-3) / 1e12,
                         "hbm": {"achieved": alg_bytes * args.steps / (ms[2] * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": alg_bytes * args.steps / (ms[2] * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
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
    ap.add_argument("--lines", type=int, default=100, help="lines per spectrum (BASELINE: 100)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
