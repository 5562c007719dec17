#!/usr/bin/env python
"""bench.py -- headline benchmark of the pair-counting hot path (BASELINE.json).

Workload (N=1): configs[1] of BASELINE.json -- FCFC_2PT_BOX xi(s, mu): 10^7 uniform points in a
2 Gpc/h periodic box, s in [0, 200) Mpc/h in 40 bins x 120 mu bins, DD auto count.
A "step" is one full count_pairs pass over that catalogue.

  value    pair evaluations / s with the catalogue resident (cell-sorted) in HBM
  e2e      same metric through the public C-ABI call sequence with HOST buffers:
           catalog_create (H2D + rescale + cell sort) + count + histogram D2H, every step
  roofline pair-evaluation roofline: FP32 issue peak (measured FFMA stream) / 6 instructions per evaluation
  cpu_baseline / --impl reference: the unmodified reference (oracle/_ref, its fastest SIMD build, all host
           cores) timed on count_pairs alone, on a bounded sample with the same number density and bins.
           Its value is job-equivalent: (in-range pairs of the sample x evaluations-per-in-range-pair of this
           workload) / seconds, so that the ratio of the two arms is the ratio of times for the same job.

Multi-GPU (torchrun, one rank per GPU): the catalogue is replicated, the primary work items are split
into WORLD_SIZE parts, and the per-rank histograms are combined by one NCCL all-reduce (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, L, bintype, nmu)  -- s in [0,200) step 5 everywhere
    "c2_box_smu_1e7": dict(n=10_000_000, box=2000.0, bintype=1, nmu=120, desc="FCFC_2PT_BOX xi(s,mu) 10^7 pts L=2000 40x120 bins DD"),
    "c1_box_iso_1e6": dict(n=1_000_000, box=1000.0, bintype=0, nmu=1, desc="FCFC_2PT_BOX xi(s) 10^6 pts L=1000 40 bins DD"),
    # single-GPU version of configs[3] (clustered mock): Neyman-Scott blobs + 20 % background
    "c4_box_smu_clustered_1e7": dict(n=10_000_000, box=2000.0, bintype=1, nmu=120, clustered=True,
                                     desc="FCFC_2PT_BOX xi(s,mu) 10^7 clustered pts (Neyman-Scott, sigma=1.5, 20% background) L=2000 40x120 bins DD"),
}
# configs[2] of BASELINE.json: FCFC_2PT survey mode, weighted DD/DR/RR, xi(s_perp, pi) (and w_p from it on the host).
# A step is the three pair counts.  Double precision: the reference's default build.
SURVEY_WORKLOADS = {
    "c3_svy_spi_wt_2e6_2e7": dict(nd=2_000_000, nr=20_000_000, smax=40.0, ds=2.0, pmax=80.0, dpi=1.0,
                                  desc="FCFC_2PT weighted DD+DR+RR, xi(s_perp,pi): 2x10^6 data + 2x10^7 randoms, 20 s_perp x 80 pi bins, "
                                       "60 deg x 30 deg patch at 1000-1700 Mpc/h comoving (coordinates already converted)"),
}
METRIC = "pair_evals_per_sec"
UNIT = "pair evaluations/s"
KAPPA_FILE = os.path.join(ROOT, "profiles", "workload_kappa.json")


_OUT = None


def claim_stdout():
    """Keep stdout for the one JSON line of the contract: whatever else writes to file descriptor 1 from here on
    (NCCL prints its version banner there) goes to stderr."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_box(n, L, seed=20261017):
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray(rng.random(n) * L) for _ in range(3)]


def make_clustered_box(n, L, seed=20261017, sigma=1.5, frac_bg=0.2):
    """Neyman-Scott style mock: n/50 parents, 40 Gaussian children each, 20 % uniform background."""
    rng = np.random.default_rng(seed)
    nbg = int(n * frac_bg)
    ncl = n - nbg
    par = rng.random((max(1, ncl // 40), 3)) * L
    x = np.concatenate([par[rng.integers(0, len(par), ncl)] + rng.normal(0, sigma, (ncl, 3)), rng.random((nbg, 3)) * L]) % L
    x[x >= L] = 0.0
    rng.shuffle(x)
    return [np.ascontiguousarray(x[:, k]) for k in range(3)]


def make_catalog(wl, n, L, seed=20261017):
    return make_clustered_box(n, L, seed) if wl.get("clustered") else make_box(n, L, seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_survey(n, seed, shrink=1.0):
    """Comoving coordinates + weights of a survey patch: RA in [120, 120 + 60 g) deg, sin(dec) in [0, 0.5 g), uniform in
    volume between 1000 Mpc/h and the radius that keeps the volume fraction g^3 of the full 1000-1700 Mpc/h shell
    (g = shrink: bounded CPU samples keep the number density of the full workload)."""
    rng = np.random.default_rng(seed)
    ra = np.deg2rad(rng.uniform(120.0, 120.0 + 60.0 * shrink, n))
    sd = rng.uniform(0.0, 0.5 * shrink, n)
    cd = np.sqrt(1.0 - sd * sd)
    d = np.cbrt(rng.uniform(1000.0 ** 3, 1000.0 ** 3 + shrink * (1700.0 ** 3 - 1000.0 ** 3), n))
    return [np.ascontiguousarray(a) for a in (d * cd * np.cos(ra), d * cd * np.sin(ra), d * sd, rng.uniform(0.75, 1.25, n))]


def run_survey_reference(args, wl):
    """The unmodified reference (count_pairs of DD, DR, RR) on a bounded sample of the survey workload: 1/20 of the
    objects in 1/20 of the volume (same number densities).  Job-equivalent value: evaluations of the full job per
    unit of weighted in-range pair sum (profiles/workload_kappa.json) x the sample's weighted sum / seconds."""
    from oracle import refdrv
    frac = args.cpu_sample / float(wl["nr"]) if args.cpu_sample < wl["nr"] else 1.0
    frac = min(frac, 1.0)
    nd, nr = max(1000, int(wl["nd"] * frac)), max(1000, int(wl["nr"] * frac))
    g = frac ** (1.0 / 3.0)
    D, R = make_survey(nd, 7, g), make_survey(nr, 8, g)
    prec = "flt" if args.prec == "float" else "dbl"
    flav = refdrv.best_simd_flavour(prec)
    p, isa = flav.split("_")
    cores = os.cpu_count() or 1
    nrep = max(1, args.steps if args.impl == "reference" else 1)
    warm = args.warmup if args.impl == "reference" else 0
    times, wsum = [], 0.0
    for it in range(warm + nrep):
        r = refdrv.run_reference([tuple(D), tuple(R)], periodic=False, prec=p, isa=isa, pairs=["DD", "DR", "RR"], bintype=2,
                                 smin=0.0, smax=wl["smax"], ds=wl["ds"], pmin=0.0, pmax=wl["pmax"], dpi=wl["dpi"], threads=cores)
        if it >= warm:
            times.append(sum(q.t_count for q in r.pairs))
        wsum = float(sum(q.cnt.sum() for q in r.pairs))
    t = float(np.mean(times))
    kappa = 10.0
    try:
        kappa = json.load(open(KAPPA_FILE)).get(args.workload, {}).get("evals_per_weighted_pair", kappa)
    except Exception:
        pass
    return dict(seconds=t, wsum=wsum, value=kappa * wsum / t, cores=cores, flavour=flav,
                sample=f"{nd} data + {nr} randoms in {frac:.3g} of the survey volume (same densities and bins), count_pairs of "
                       f"DD+DR+RR only, reference build {flav}, OMP_NUM_THREADS={cores}, weighted in-range sum={wsum:.6g}, {t:.3f} s; "
                       f"value = evals_per_weighted_pair({kappa:.3f}) x sum / s")


def bench_survey(args):
    """configs[2]: weighted survey counts in double precision; same JSON contract as the box workloads."""
    wl = SURVEY_WORKLOADS[args.workload]
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    prec = "double" if args.prec_given is None else args.prec
    args.prec = prec
    config = {"workload": f"{args.workload}: {wl['desc']}", "n_data": wl["nd"], "n_random": wl["nr"],
              "bins": f"{int(wl['smax'] / wl['ds'])} s_perp x {int(wl['pmax'] / wl['dpi'])} pi", "arith": "fma" if args.arith else "scalar",
              "l2": "inputs (cell-sorted double4 randoms: 640 MB) larger than the 126 MB L2",
              "parallelism": f"primary work items split over {world} rank(s), secondary replicated, NCCL all-reduce of the histograms"}
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_survey_reference(args, wl)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32" if prec == "float" else "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
                          "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    claim_stdout()
    import torch
    import fcfc_b200 as F
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    F.init(devices=[local])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    bins = F.Bins(periodic=False, prec=prec, arith=args.arith, bintype=2, smin=0.0, smax=wl["smax"], ds=wl["ds"],
                  pmin=0.0, pmax=wl["pmax"], dpi=wl["dpi"])
    npdt = np.float32 if prec == "float" else np.float64
    host = []
    for n, seed in ((wl["nd"], 1), (wl["nr"], 2)):
        pinned = [torch.from_numpy(a.astype(npdt)).pin_memory() for a in make_survey(n, seed)]
        host.append((pinned, [p.numpy() for p in pinned]))
    ntot = bins.ntot
    dev_hist = [torch.zeros(ntot, dtype=torch.float64, device=dev) for _ in range(3)]
    PAIRS = ((0, 0), (0, 1), (1, 1))          # DD, DR, RR

    def count_all(cats):
        ev, nl, kms, out = 0, 0, 0.0, []
        for k, (i, j) in enumerate(PAIRS):
            c = F.count_pairs(cats[i], None if i == j else cats[j], bins, withwt=True, part=rank, nparts=world,
                              dev_hist_ptr=dev_hist[k].data_ptr())
            st = F.stats()
            ev += st["pair_evals"]; nl += st["kernel_launches"]; kms += st["ms_count"]
            if dist is not None:
                dist.all_reduce(dev_hist[k])
                c = dev_hist[k].cpu().numpy()
            out.append(c)
        return out, ev, nl, kms, st

    cats = [F.Catalog(*h[1], bins=bins) for h in host]
    for _ in range(max(3, args.warmup)):
        counts, _, _, _, st = count_all(cats)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evals = launches = 0
    kern_ms = 0.0
    for _ in range(args.steps):
        counts, ev, nl, kms, st = count_all(cats)
        evals += ev; launches += nl; kern_ms += kms
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    evt = torch.tensor([float(evals)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(evt, op=dist.ReduceOp.SUM)
    ms, total_evals = float(t_ms.item()), float(evt.item())
    value = total_evals / (ms * 1e-3)
    for c in cats:
        c.destroy()

    # ---- end to end: upload both catalogues from pinned host memory, three counts, histograms back ----
    def step_e2e():
        cc = [F.Catalog(*h[1], bins=bins) for h in host]
        out, ev, _, _, _ = count_all(cc)
        for c in cc:
            c.destroy()
        return out, ev
    step_e2e()
    barrier()
    e0.record()
    ev_e2e = 0.0
    for _ in range(args.steps):
        c2, ev = step_e2e()
        ev_e2e += ev
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    ev2 = torch.tensor([ev_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev2, op=dist.ReduceOp.SUM)
    e2e_value = float(ev2.item()) / (float(t2.item()) * 1e-3)
    rel = max(float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) for a, b in zip(c2, counts))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    wsum = float(sum(c.sum() for c in counts))
    # roofline: FP64 issue peak / 5 instructions per evaluation up to the first range test (dot product 3, s1 + s2, s - t)
    peak64 = F.measure_fp64_peak() if prec == "double" else F.measure_fp32_peak()[0]
    achieved = (evals * world / args.steps) / (kern_ms / args.steps * 1e-3) * 5.0 * 1e-12 if world == 1 else None
    roofline = {"bound": "fp64_issue" if prec == "double" else "fp32_issue", "achieved": achieved, "peak": peak64 * 1e-12,
                "unit": "T FP instr/s (5 per pair evaluation)", "frac": (achieved / (peak64 * 1e-12)) if achieved else None, "traffic": None,
                "algorithmic_bytes": 32 * (wl["nd"] + wl["nr"]), "kernel": "fcfc::count_kernel (survey, weighted)", "kernel_ms": kern_ms / args.steps,
                "peak_source": "measured live: DFMA stream on all SMs (fcfc_gpu_measure_fp64_peak)"}
    kappa = total_evals / args.steps / max(wsum, 1e-300)
    try:
        allk = json.load(open(KAPPA_FILE)) if os.path.exists(KAPPA_FILE) else {}
        allk[args.workload] = {"evals_per_weighted_pair": kappa, "weighted_pairs_in": wsum}
        json.dump(allk, open(KAPPA_FILE, "w"), indent=1)
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            r = run_survey_reference(args, wl)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    isz = np.dtype(npdt).itemsize
    emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                      "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f32" if prec == "float" else "f64", "data": "synthetic", "config": config, "clocks": clocks,
                      "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * (wl["nd"] + wl["nr"]) * isz),
                              "d2h_bytes_per_step": int(3 * ntot * 8), "ms_per_step": float(t2.item()) / args.steps,
                              "max_rel_diff_vs_resident": rel},
                      "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                      "weighted_pairs_in": wsum, "evals_per_weighted_pair": kappa, "grid": st["ncell"], "work_items": st["nitem"]})
    if dist is not None:
        dist.destroy_process_group()


def run_reference_arm(args, wl, sample_n, prec):
    """Time the unmodified reference (count_pairs only) on a bounded sample of the workload."""
    from oracle import refdrv
    dens = wl["n"] / wl["box"] ** 3
    sample_n = max(sample_n, int(dens * (2.1 * 200.0) ** 3) + 1)      # the periodic sample box must exceed twice the maximum separation
    Ls = (sample_n / dens) ** (1.0 / 3.0)
    cat = make_catalog(wl, sample_n, Ls, seed=7)
    flav = refdrv.best_simd_flavour("flt" if prec == "float" else "dbl")
    p, isa = flav.split("_")
    kw = dict(bintype=wl["bintype"], smin=0.0, smax=200.0, ds=5.0)
    if wl["bintype"] == 1:
        kw["nmu"] = wl["nmu"]
    cores = os.cpu_count() or 1
    pairs = 0
    nrep = max(1, args.steps if args.impl == "reference" else 1)
    warm = args.warmup if args.impl == "reference" else 0
    times = []
    for it in range(warm + nrep):
        r = refdrv.run_reference([tuple(cat)], periodic=True, prec=p, isa=isa, pairs=["DD"], box=Ls, threads=cores, **kw)
        if it >= warm:
            times.append(r.pairs[0].t_count)
        pairs = int(r.pairs[0].cnt.sum())
    t = float(np.mean(times))
    kappa = 3.0
    try:
        kappa = json.load(open(KAPPA_FILE)).get(args.workload, {}).get("evals_per_inrange_pair", kappa)
    except Exception:
        pass
    return dict(seconds=t, pairs=pairs, value=kappa * pairs / t, kappa=kappa, cores=cores, flavour=flav,
                sample=f"{sample_n} {'clustered' if wl.get('clustered') else 'uniform'} points in a {Ls:.1f} Mpc/h box (same density and bins as the workload), "
                       f"count_pairs only, reference build {flav}, OMP_NUM_THREADS={cores}, in-range pairs={pairs}, "
                       f"{t:.3f} s; value = evals_per_inrange_pair({kappa:.3f}) x pairs / s")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_box_smu_1e7", choices=sorted(WORKLOADS) + sorted(SURVEY_WORKLOADS))
    ap.add_argument("--prec", default=None, choices=["float", "double"], help="default: float for the box workloads, double for the survey")
    ap.add_argument("--arith", type=int, default=1, help="0 scalar-parity order, 1 FMA order")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="points of the bounded CPU sample (default: 2x10^6 box points, about 5 s on 16 threads; survey: 10^6 randoms)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.prec_given = args.prec
    if args.workload in SURVEY_WORKLOADS:
        args.cpu_sample = args.cpu_sample or 1_000_000         # randoms of the bounded survey sample
        return bench_survey(args)
    args.cpu_sample = args.cpu_sample or 2_000_000
    args.prec = args.prec or "float"
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {wl['desc']}", "n_points": wl["n"], "box": wl["box"], "bins": f"40 s x {wl['nmu']} mu",
              "arith": "fma" if args.arith else "scalar", "l2": "inputs (160 MB cell-sorted float4) larger than the 126 MB L2",
              "parallelism": f"primary work items split over {world} rank(s), secondary replicated, NCCL all-reduce of the histogram"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference_arm(args, wl, args.cpu_sample, args.prec)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32" if args.prec == "float" else "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "in_range_pairs_per_sec": r["pairs"] / r["seconds"]}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    claim_stdout()
    import torch
    import fcfc_b200 as F
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F.init(devices=[local])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    bins = F.Bins(periodic=True, prec=args.prec, arith=args.arith, box=wl["box"], bintype=wl["bintype"], smin=0.0, smax=200.0,
                  ds=5.0, nmu=wl["nmu"])
    npdt = np.float32 if args.prec == "float" else np.float64
    xyz = make_catalog(wl, wl["n"], wl["box"])
    # pinned host buffers of the build's `real` type (what the FCFC host holds after reading the catalogue)
    pinned = [torch.from_numpy(a.astype(npdt)).pin_memory() for a in xyz]
    host = [p.numpy() for p in pinned]
    del xyz
    ntot = bins.ntot
    dev_hist = torch.zeros(ntot, dtype=torch.int64, device=dev)

    cat = F.Catalog(*host, bins=bins)
    launches = {"n": 0}
    evals = {"n": 0}

    def step_resident():
        c = F.count_pairs(cat, None, bins, part=rank, nparts=world, dev_hist_ptr=dev_hist.data_ptr())
        st = F.stats()
        launches["n"] += st["kernel_launches"]; evals["n"] += st["pair_evals"]
        if dist is not None:
            dist.all_reduce(dev_hist)
        return c, st

    def step_e2e():
        g = F.Catalog(*host, bins=bins)                       # H2D + rescale + stats
        c = F.count_pairs(g, None, bins, part=rank, nparts=world, dev_hist_ptr=dev_hist.data_ptr())   # sort + count + D2H
        st = F.stats()
        if dist is not None:
            dist.all_reduce(dev_hist)
            c = dev_hist.cpu().numpy()
        g.destroy()
        return c, st

    # ---- resident (value) ----
    for _ in range(max(3, args.warmup)):
        counts, st = step_resident()
    launches["n"] = 0; evals["n"] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kern_ms = []
    for _ in range(args.steps):
        counts, st = step_resident()
        kern_ms.append(st["ms_count"])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    ev = torch.tensor([float(evals["n"])], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item()); total_evals = float(ev.item())
    value = total_evals / (ms * 1e-3)
    total_counts = dev_hist.cpu().numpy() if dist is not None else counts
    pairs_in = int(total_counts.sum())

    # ---- end to end (host buffers, copies inside the timed region) ----
    for _ in range(1):
        step_e2e()
    barrier()
    e0.record()
    ev_e2e = 0.0
    for _ in range(args.steps):
        c2, st2 = step_e2e()
        ev_e2e += st2["pair_evals"]
    e1.record()
    barrier()
    ms2 = e0.elapsed_time(e1)
    t2 = torch.tensor([ms2], dtype=torch.float64, device=dev)
    ev2 = torch.tensor([ev_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev2, op=dist.ReduceOp.SUM)
    e2e_value = float(ev2.item()) / (float(t2.item()) * 1e-3)
    same = bool(np.array_equal(np.asarray(c2), total_counts))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (count_kernel): FP32 issue peak / 6 instr per evaluation ----
    peak_instr, implied_mhz = F.measure_fp32_peak()
    k_ms = float(np.mean(kern_ms))
    evals_per_launch = evals["n"] / args.steps
    achieved = evals_per_launch / (k_ms * 1e-3) * 6.0 * 1e-12           # T FP32 lane-instructions/s of algorithmic work
    traffic = None
    try:        # DRAM bytes of one launch of this workload's count kernel, from the committed ncu --set full capture
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
        if t and world == 1 and args.prec == "float":
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "fp32_issue", "achieved": achieved, "peak": peak_instr * 1e-12, "unit": "T FP32 instr/s (6 per pair evaluation)",
                "frac": achieved / (peak_instr * 1e-12), "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                "algorithmic_bytes": 16 * wl["n"], "kernel": "fcfc::count_kernel", "kernel_ms": k_ms,
                "peak_source": "measured live: FFMA stream on all SMs (fcfc_gpu_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                "evals_per_sec_kernel": evals_per_launch / (k_ms * 1e-3), "r_eval_peak": peak_instr / 6.0,
                # the same statement in flops: 8 flop per evaluation against 2 flop per FFMA lane
                "flop_equivalent": {"achieved_tflops": evals_per_launch / (k_ms * 1e-3) * 8e-12, "peak_tflops": 2.0 * peak_instr * 1e-12}}
    kappa = evals["n"] / args.steps * world / max(pairs_in, 1) if world == 1 else total_evals / args.steps / max(pairs_in, 1)
    try:
        os.makedirs(os.path.dirname(KAPPA_FILE), exist_ok=True)
        allk = json.load(open(KAPPA_FILE)) if os.path.exists(KAPPA_FILE) else {}
        allk[args.workload] = {"evals_per_inrange_pair": kappa, "pairs_in": pairs_in}
        json.dump(allk, open(KAPPA_FILE, "w"), indent=1)
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            r = run_reference_arm(args, wl, args.cpu_sample, args.prec)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                   "in_range_pairs_per_sec": r["pairs"] / r["seconds"]}
        except Exception as ex:     # the baseline is reported, never required for the product path
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    bytes_in = 3 * wl["n"] * np.dtype(npdt).itemsize
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.prec == "float" else "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_in), "d2h_bytes_per_step": int(ntot * 8),
                    "ms_per_step": float(t2.item()) / args.steps, "same_counts_as_resident": same},
            "gpu_launches": int(launches["n"]), "roofline": roofline, "cpu_baseline": cpu,
            "in_range_pairs": pairs_in, "in_range_pairs_per_sec": pairs_in / (ms / args.steps * 1e-3),
            "evals_per_inrange_pair": kappa, "dd_wall_time_s": float(t2.item()) / args.steps * 1e-3,
            "grid": st["ncell"], "work_items": st["nitem"]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
