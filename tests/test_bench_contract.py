"""CPU: the reference arm of bench.py (the unmodified reference on a bounded sample, oracle/_ref) prints the JSON
line of the measurement contract for every workload; the synthetic generators keep their shapes and densities."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, have_ref

sys.path.insert(0, ROOT)
import bench  # noqa: E402

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.skipif(not (have_ref("flt_scalar", "box") and have_ref("dbl_scalar", "svy")), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload,sample", [("c2_box_smu_1e7", 20000), ("c4_box_smu_clustered_1e7", 20000),
                                             ("c3_svy_spi_wt_2e6_2e7", 40000)])
def test_reference_arm_line(workload, sample):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "1", "--warmup", "0", "--cpu-sample", str(sample)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert REQUIRED <= set(line), REQUIRED - set(line)
    assert line["impl"] == "reference" and line["metric"] == "pair_evals_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith(workload)


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_generators():
    x = bench.make_clustered_box(50000, 500.0, seed=3)
    assert all(len(a) == 50000 and a.min() >= 0 and a.max() < 500.0 for a in x)
    # clustered: many more close pairs than a uniform catalogue of the same density
    def close_pairs(c):
        p = np.stack(c, 1)[:4000]
        d = np.abs(p[:, None, :] - p[None, :, :])
        d = np.minimum(d, 500.0 - d)
        return int(((d ** 2).sum(-1) < 9.0).sum()) - len(p)
    assert close_pairs(x) > 20 * max(1, close_pairs(bench.make_box(50000, 500.0, seed=3)))
    full, part = bench.make_survey(20000, 1), bench.make_survey(20000, 1, shrink=0.5)
    for c in (full, part):
        r = np.sqrt(c[0] ** 2 + c[1] ** 2 + c[2] ** 2)
        assert r.min() >= 1000.0 - 1e-6 and r.max() <= 1700.0 + 1e-6 and 0.75 <= c[3].min() and c[3].max() <= 1.25
    rp = np.sqrt(sum(a ** 2 for a in part[:3]))
    assert rp.max() < 1000.0 * (1 + 0.5 * (1.7 ** 3 - 1)) ** (1 / 3) + 1e-6      # the shrunk sample keeps the inner radius


def test_blocked_generators_do_not_depend_on_the_split():
    """The 10^8-point workloads are generated in 64 seeded blocks so that every rank builds only the rows it uploads:
    any split into row ranges gives the same catalogue."""
    wl = dict(bench.WORKLOADS["c4_box_smu_clustered_1e8"], n=64 * 500)
    full = bench.box_rows(wl, 0, wl["n"])
    parts = [bench.box_rows(wl, lo, hi) for lo, hi in ((0, 7001), (7001, 20000), (20000, 32000))]
    assert all(np.array_equal(np.concatenate([p[k] for p in parts]), full[k]) for k in range(3))
    assert all(0 <= c.min() and c.max() < wl["box"] for c in full)
    sw = dict(bench.SURVEY_WORKLOADS["c5_svy_spi_wt_2e6_1e8"], nr=64 * 300, nd=1000)
    f = bench.survey_rows(sw, 1, 0, sw["nr"])
    q = [bench.survey_rows(sw, 1, a, b) for a, b in ((0, 999), (999, 19200))]
    assert all(np.array_equal(np.concatenate([p[k] for p in q]), f[k]) for k in range(4))
    assert len(bench.survey_rows(sw, 0, 100, 300)[0]) == 200
