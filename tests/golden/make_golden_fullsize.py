"""Golden vectors of the BASELINE configs at their real size, from the compiled, unmodified reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):
    python tests/golden/make_golden_fullsize.py [c1] [c2] [c4s] [c4] [c3]
It takes tens of minutes of CPU time (the 10^7-point (s,mu) count is ~2x10^11 in-range pairs per run), which is why
the outputs are committed as fixtures (fullsize_*.npz) instead of being recomputed by the tests:

  c1   BASELINE configs[0]: 10^6 uniform points, L = 1000, xi(s) 40 bins, DD     (bench.py workload c1_box_iso_1e6)
  c2   BASELINE configs[1]: 10^7 uniform points, L = 2000, xi(s,mu) 40 x 120, DD (bench.py workload c2_box_smu_1e7,
       the headline bench workload: bench.py checks its own step against these counts)
  c4s  a 2x10^6-point clustered (s,mu) box, the small-scale twin of configs[3]
  c4   the clustered (s,mu) box at the single-GPU size of configs[3]: 10^7 points, L = 2000 (bench.py workload
       c4_box_smu_clustered_1e7)
  c3   BASELINE configs[2]: survey, 2x10^6 data + 2x10^7 randoms, weighted DD + DR + RR, xi(s_perp,pi) (bench.py workload
       c3_svy_spi_wt_2e6_2e7; only when named: it is not in the default list)

For each: the reference's double AVX-512 build (k-d tree) -- which the survey found identical to the scalar double
build and to the ball tree -- and the float AVX-512 / float scalar builds with both trees, whose mutual differences
are the "reference's own spread" that bounds the float comparison (gate G2).
The catalogues are regenerated from bench.py's seeded generators; a checksum guards against generator drift.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import refdrv  # noqa: E402

JOBS = {
    "c1": dict(workload="c1_box_iso_1e6", n=None,
               runs=[("dbl", "avx512", 0), ("dbl", "scalar", 0), ("dbl", "avx512", 1), ("flt", "avx512", 0), ("flt", "avx512", 1),
                     ("flt", "scalar", 0), ("flt", "scalar", 1)]),
    "c2": dict(workload="c2_box_smu_1e7", n=None,
               runs=[("dbl", "avx512", 0), ("flt", "avx512", 0), ("flt", "avx512", 1), ("flt", "scalar", 0)]),
    "c4s": dict(workload="c4_box_smu_clustered_1e7", n=2_000_000, box=1169.607095285,
                runs=[("dbl", "avx512", 0), ("flt", "avx512", 0), ("flt", "avx512", 1), ("flt", "scalar", 0)]),
    "c4": dict(workload="c4_box_smu_clustered_1e7", n=None,
               runs=[("dbl", "avx512", 0), ("flt", "avx512", 0), ("flt", "avx512", 1), ("flt", "scalar", 0)]),
}


def survey_job():
    """c3: BASELINE configs[2], 2x10^6 data + 2x10^7 randoms, weighted DD + DR + RR, xi(s_perp,pi) 20 x 80 bins, the reference's
    default (double, AVX-512) build (bench.py workload c3_svy_spi_wt_2e6_2e7: bench.py checks its own step against these sums)."""
    wl = bench.SURVEY_WORKLOADS["c3_svy_spi_wt_2e6_2e7"]
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    D, R = bench.make_survey(wl["nd"], 1), bench.make_survey(wl["nr"], 2)
    out = {"input_sha256": np.array(checksum(D + R)), "nd": np.array(wl["nd"]), "nr": np.array(wl["nr"]), "workload": np.array("c3_svy_spi_wt_2e6_2e7")}
    t0 = time.time()
    r = refdrv.run_reference([tuple(D), tuple(R)], periodic=False, prec="dbl", isa="avx512", pairs=["DD", "DR", "RR"], threads=threads,
                             bintype=2, smin=0.0, smax=wl["smax"], ds=wl["ds"], pmin=0.0, pmax=wl["pmax"], dpi=wl["dpi"], timeout=6 * 3600)
    for k, p in enumerate(["DD", "DR", "RR"]):
        out[f"dbl_avx512_kd_{p}"] = r.pairs[k].cnt
        out[f"dbl_avx512_kd_{p}_seconds"] = np.array(r.pairs[k].t_count)
        print("c3", p, f"weighted sum {r.pairs[k].cnt.sum():.10g}, count_pairs {r.pairs[k].t_count:.1f} s on {threads} threads", flush=True)
    print("c3 wall", f"{time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(os.path.join(HERE, "fullsize_c3.npz"), **out)


def checksum(cols):
    h = hashlib.sha256()
    for a in cols:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    names = sys.argv[1:] or list(JOBS)
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    for name in names:
        if name == "c3":
            survey_job()
            continue
        job = JOBS[name]
        wl = bench.WORKLOADS[job["workload"]]
        n = job["n"] or wl["n"]
        L = job.get("box", wl["box"])
        cat = bench.make_catalog(wl, n, L)
        out = {"input_sha256": np.array(checksum(cat)), "n": np.array(n), "box": np.array(L), "bintype": np.array(wl["bintype"]),
               "nmu": np.array(wl["nmu"]), "workload": np.array(job["workload"])}
        kw = dict(bintype=wl["bintype"], smin=0.0, smax=200.0, ds=5.0)
        if wl["bintype"] == 1:
            kw["nmu"] = wl["nmu"]
        for prec, isa, ds in job["runs"]:
            t0 = time.time()
            r = refdrv.run_reference([tuple(cat)], periodic=True, prec=prec, isa=isa, pairs=["DD"], box=L, threads=threads,
                                     data_struct=ds, timeout=6 * 3600, **kw)
            key = f"{prec}_{isa}_{'kd' if ds == 0 else 'ball'}_DD"
            out[key] = r.pairs[0].cnt
            out[key + "_seconds"] = np.array(r.pairs[0].t_count)
            print(name, key, int(r.pairs[0].cnt.sum()), f"count_pairs {r.pairs[0].t_count:.1f} s on {threads} threads, wall {time.time() - t0:.1f} s",
                  flush=True)
            np.savez_compressed(os.path.join(HERE, f"fullsize_{name}.npz"), **out)


if __name__ == "__main__":
    main()
