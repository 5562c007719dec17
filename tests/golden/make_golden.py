"""Generate the golden vectors under tests/golden/ from the compiled, unmodified reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle` has been run):
    python tests/golden/make_golden.py
For every case of tests/cases.py and both precisions it stores the raw count_pairs output of the
reference's scalar build (the bit-exact parity oracle), the counts of the AVX-512 build (double: equal to scalar,
SURVEY.md section 7, and the pin of the engine's FMA order; float: vector and scalar-remainder formulas mixed, kept
to bound the float FMA order), and the tables cf_setup built.
The inputs are not stored: they are regenerated from the seeds in tests/cases.py, and a checksum of the
regenerated inputs is stored to detect generator drift.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import CASES, make_catalog  # noqa: E402
from oracle import refdrv  # noqa: E402


def checksum(cats):
    h = hashlib.sha256()
    for c in cats:
        for a in c:
            h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    for name, case in CASES.items():
        cats = [make_catalog(s, case["withwt"]) for s in case["cats"]]
        out = {"input_sha256": np.array(checksum(cats))}
        for prec in ("dbl", "flt"):
            r = refdrv.run_reference(cats, periodic=case["periodic"], prec=prec, isa="scalar", pairs=case["pairs"], **case["kw"])
            out[f"{prec}_rescale"] = np.array(r.rescale)
            out[f"{prec}_tabtype"] = np.array(r.tabtype)
            out[f"{prec}_s2bin"] = r.s2bin
            out[f"{prec}_stab"] = r.stab
            out[f"{prec}_bsize"] = r.bsize
            if r.pbin is not None:
                out[f"{prec}_pbin"] = r.pbin
                out[f"{prec}_ptab"] = r.ptab
            if r.mutab is not None:
                out[f"{prec}_mutab"] = r.mutab
            for p in r.pairs:
                out[f"{prec}_scalar_{p.label}"] = p.cnt
            # the AVX-512 build (the reference's default kind of build): double is the exact pin of the engine's FMA
            # order; float is stored too -- it mixes vector and scalar-remainder formulas, so it bounds, not pins
            r2 = refdrv.run_reference(cats, periodic=case["periodic"], prec=prec, isa="avx512", pairs=case["pairs"], **case["kw"])
            for p in r2.pairs:
                out[f"{prec}_avx512_{p.label}"] = p.cnt
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "ok", {k: (int(v.sum()) if v.dtype.kind == "i" else float(v.sum())) for k, v in out.items() if "_scalar_" in k})


if __name__ == "__main__":
    main()
