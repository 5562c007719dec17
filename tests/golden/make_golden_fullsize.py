"""Golden vectors of the BASELINE configs at their real size, from the compiled, unmodified reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):
    python tests/golden/make_golden_fullsize.py [c1] [c2] [c4s]
It takes tens of minutes of CPU time (the 10^7-point (s,mu) count is ~2x10^11 in-range pairs per run), which is why
the outputs are committed as fixtures (fullsize_*.npz) instead of being recomputed by the tests:

  c1   BASELINE configs[0]: 10^6 uniform points, L = 1000, xi(s) 40 bins, DD     (bench.py workload c1_box_iso_1e6)
  c2   BASELINE configs[1]: 10^7 uniform points, L = 2000, xi(s,mu) 40 x 120, DD (bench.py workload c2_box_smu_1e7,
       the headline bench workload: bench.py checks its own step against these counts)
  c4s  a 2x10^6-point clustered (s,mu) box, the small-scale twin of configs[3]

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
}


def checksum(cols):
    h = hashlib.sha256()
    for a in cols:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    names = sys.argv[1:] or list(JOBS)
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    for name in names:
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
