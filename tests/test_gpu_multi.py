"""GPU (-m gpu), needs >= 2 devices (skipped otherwise): one process driving several GPUs through
fcfc_gpu_count -- primary work items sharded over the devices, secondary replicated, NCCL all-reduce of the
histograms -- must reproduce the single-device counts exactly (weighted: to 1e-12)."""
import numpy as np
import pytest

from cases import box_catalog, survey_catalog

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ndev() < 2, reason="needs at least two GPUs")
def test_in_process_multi_device_matches_single_device():
    import fcfc_b200 as F
    cat = box_catalog(300000, 600.0, 91)
    D, R = survey_catalog(30000, 92), survey_catalog(90000, 93)
    res = {}
    # "all": the per-device histograms are added on the host (the default of the in-process path);
    # "nccl": all-reduced on the devices (option nccl = 1)
    for tag, devs in (("one", [0]), ("all", None), ("nccl", None)):
        n = F.init(devices=devs)
        assert n == (1 if devs else _ndev())
        F.set_option("nccl", 1 if tag == "nccl" else 0)
        out = []
        b = F.Bins(periodic=True, prec="float", arith=1, box=600.0, bintype=1, smax=60.0, ds=1.5, nmu=60)
        g = F.Catalog(*cat[:3], bins=b)
        out.append(F.count_pairs(g, None, b)); g.destroy()
        b = F.Bins(periodic=True, prec="double", box=600.0, bintype=0, smax=60.0, ds=1.5)
        g = F.Catalog(*cat, bins=b)
        out.append(F.count_pairs(g, None, b, withwt=True)); g.destroy()
        b = F.Bins(periodic=False, prec="double", bintype=2, smax=40.0, ds=2.0, pmin=0.0, pmax=80.0, dpi=1.0)
        gd, gr = F.Catalog(*D, bins=b), F.Catalog(*R, bins=b)
        out.append(F.count_pairs(gd, gr, b, withwt=True)); out.append(F.count_pairs(gr, None, b, withwt=False))
        gd.destroy(); gr.destroy()
        res[tag] = out
    F.set_option("defaults", 0)
    F.init(devices=[0])
    np.testing.assert_array_equal(res["one"][0], res["nccl"][0])
    np.testing.assert_allclose(res["nccl"][1], res["one"][1], rtol=1e-12, atol=0)
    np.testing.assert_allclose(res["nccl"][2], res["one"][2], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(res["one"][3], res["nccl"][3])
    np.testing.assert_array_equal(res["one"][0], res["all"][0])
    np.testing.assert_allclose(res["all"][1], res["one"][1], rtol=1e-12, atol=0)
    np.testing.assert_allclose(res["all"][2], res["one"][2], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(res["one"][3], res["all"][3])
    assert res["one"][0].sum() > 0


@pytest.mark.skipif(_ndev() < 2, reason="needs at least two GPUs")
def test_streamed_catalogue_is_replicated_on_every_device():
    """Streamed ingest (fcfc_gpu_catalog_stream_*) lands on the first device and is copied to the others like a one-shot
    upload: the sharded count over all devices equals the single-device count."""
    import fcfc_b200 as F
    x, y, z, _ = box_catalog(200000, 600.0, 94)
    res = {}
    for tag, devs in (("one", [0]), ("all", None)):
        F.init(devices=devs)
        b = F.Bins(periodic=True, prec="float", arith=1, box=600.0, bintype=1, smax=60.0, ds=1.5, nmu=60)
        st = F.CatalogStream(bins=b, n_hint=1000)
        for lo in range(0, len(x), 30011):
            st.append(x[lo:lo + 30011], y[lo:lo + 30011], z[lo:lo + 30011])
        g = st.finish()
        res[tag] = F.count_pairs(g, None, b)
        g.destroy()
    F.init(devices=[0])
    np.testing.assert_array_equal(res["one"], res["all"])
    assert res["one"].sum() > 0
