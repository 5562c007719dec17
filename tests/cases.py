"""Shared definitions of the parity cases (inputs are regenerated from seeds; outputs are golden)."""
from __future__ import annotations

import numpy as np


def box_catalog(n, L, seed, weights=True):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3)) * L
    # keep text-like precision so that the float conversion is representative of ASCII catalogues
    x = np.round(x, 6)
    x[x >= L] = 0.0
    w = rng.uniform(0.75, 1.25, n)
    return (x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy(), w) if weights else (x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy())


def clustered_box_catalog(n, L, seed, sigma=1.5, frac_bg=0.2):
    """Neyman-Scott style mock: Gaussian blobs around uniform parents + background."""
    rng = np.random.default_rng(seed)
    nbg = int(n * frac_bg)
    ncl = n - nbg
    npar = max(1, ncl // 40)
    par = rng.random((npar, 3)) * L
    idx = rng.integers(0, npar, ncl)
    pts = par[idx] + rng.normal(0, sigma, (ncl, 3))
    x = np.concatenate([pts, rng.random((nbg, 3)) * L]) % L
    x = np.round(x, 6)
    x[x >= L] = 0.0
    rng.shuffle(x)
    return x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy(), rng.uniform(0.75, 1.25, n)


def survey_catalog(n, seed):
    """Comoving coordinates of a 60 deg x 30 deg patch at 1000-1700 Mpc/h, with weights."""
    rng = np.random.default_rng(seed)
    ra = np.deg2rad(rng.uniform(120, 180, n))
    sd = rng.uniform(0, 0.5, n)
    cd = np.sqrt(1 - sd * sd)
    d = np.cbrt(rng.uniform(1000.0 ** 3, 1700.0 ** 3, n))
    return d * cd * np.cos(ra), d * cd * np.sin(ra), d * sd, rng.uniform(0.75, 1.25, n)


def lattice_box_catalog(n, L, seed):
    """Coordinates on a 2^-3 lattice: with a power-of-two rescale every float operation is exact, so all
    evaluation orders (scalar, FMA, shift or wrap form) agree bit for bit (gate G2)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, int(L * 8), (n, 3)) / 8.0
    return x[:, 0].copy(), x[:, 1].copy(), x[:, 2].copy(), rng.integers(1, 4, n) / 2.0


# name -> (periodic, catalogue spec, pairs, withwt, binning keywords)
CASES = {
    "box_iso": dict(periodic=True, cats=[("box", 3000, 1000.0, 11), ("box", 2000, 1000.0, 12)], pairs=["DD", "DR"],
                    withwt=False, kw=dict(box=1000.0, bintype=0, smin=0.0, smax=200.0, ds=5.0)),
    "box_smu": dict(periodic=True, cats=[("box", 3000, 1000.0, 11), ("box", 2000, 1000.0, 12)], pairs=["DD", "DR"],
                    withwt=False, kw=dict(box=1000.0, bintype=1, smin=0.0, smax=200.0, ds=5.0, nmu=120)),
    "box_spi": dict(periodic=True, cats=[("box", 3000, 1000.0, 11), ("box", 2000, 1000.0, 12)], pairs=["DD", "DR"],
                    withwt=False, kw=dict(box=1000.0, bintype=2, smin=0.0, smax=100.0, ds=5.0, pmin=0.0, pmax=120.0, dpi=4.0)),
    "box_iso_wt": dict(periodic=True, cats=[("box", 3000, 1000.0, 11), ("box", 2000, 1000.0, 12)], pairs=["DD", "DR"],
                       withwt=True, kw=dict(box=1000.0, bintype=0, smin=0.0, smax=200.0, ds=5.0)),
    "box_iso_smin": dict(periodic=True, cats=[("box", 3000, 1000.0, 11)], pairs=["DD"], withwt=False,
                         kw=dict(box=1000.0, bintype=0, smin=10.0, smax=150.0, ds=2.5)),
    "box_smu_hybrid": dict(periodic=True, cats=[("box", 3000, 1000.0, 11)], pairs=["DD"], withwt=False,
                           kw=dict(box=1000.0, bintype=1, sbin_edges=list(np.logspace(0, np.log10(180), 21)), nmu=50)),
    "box_spi_hybrid": dict(periodic=True, cats=[("box", 3000, 1000.0, 11)], pairs=["DD"], withwt=False,
                           kw=dict(box=1000.0, bintype=2, sbin_edges=list(np.logspace(0, np.log10(180), 21)),
                                   pbin_edges=list(np.linspace(0, 99.5, 31)))),
    "box_cuboid_clustered": dict(periodic=True, cats=[("clustered", 4000, 600.0, 13)], pairs=["DD"], withwt=True,
                                 kw=dict(box=600.0, bintype=1, smin=0.0, smax=60.0, ds=2.0, nmu=30)),
    "box_lattice_smu": dict(periodic=True, cats=[("lattice", 3000, 1024.0, 14)], pairs=["DD"], withwt=False,
                            kw=dict(box=1024.0, bintype=2, smin=0.0, smax=128.0, ds=8.0, pmin=0.0, pmax=160.0, dpi=8.0)),
    "svy_iso": dict(periodic=False, cats=[("survey", 3000, 0, 21), ("survey", 5000, 0, 22)], pairs=["DD", "DR", "RR"],
                    withwt=False, kw=dict(bintype=0, smin=0.0, smax=200.0, ds=5.0)),
    "svy_smu_wt": dict(periodic=False, cats=[("survey", 3000, 0, 21), ("survey", 5000, 0, 22)], pairs=["DD", "DR"],
                       withwt=True, kw=dict(bintype=1, smin=0.0, smax=200.0, ds=5.0, nmu=100)),
    "svy_spi_wt": dict(periodic=False, cats=[("survey", 3000, 0, 21), ("survey", 5000, 0, 22)], pairs=["DD", "DR", "RR"],
                       withwt=True, kw=dict(bintype=2, smin=0.0, smax=40.0, ds=2.0, pmin=0.0, pmax=80.0, dpi=1.0)),
    "svy_spi_min": dict(periodic=False, cats=[("survey", 3000, 0, 21), ("survey", 5000, 0, 22)], pairs=["DD", "DR"],
                        withwt=False, kw=dict(bintype=2, smin=4.0, smax=80.0, ds=2.0, pmin=10.0, pmax=150.0, dpi=2.5)),
    "svy_spi_hybrid": dict(periodic=False, cats=[("survey", 3000, 0, 21)], pairs=["DD"], withwt=False,
                           kw=dict(bintype=2, sbin_edges=list(np.logspace(-1, 2, 16)), pbin_edges=list(np.linspace(0, 120, 25)))),
}


def make_catalog(spec, withwt):
    kind, n, L, seed = spec
    if kind == "box":
        c = box_catalog(n, L, seed)
    elif kind == "clustered":
        c = clustered_box_catalog(n, L, seed)
    elif kind == "lattice":
        c = lattice_box_catalog(n, L, seed)
    elif kind == "survey":
        c = survey_catalog(n, seed)
    else:
        raise ValueError(kind)
    return c if withwt else c[:3]
