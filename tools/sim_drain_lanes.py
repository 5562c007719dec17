"""CPU simulation behind the "rotating stack columns" of count_kernel.cuh (LaneQueue): how many of the pop slots of a drain are
useful when every lane pushes the accepted pairs of its own primaries on its own stack, for several assignments of the
primaries to lanes, and when the stack columns move on to the next lane after every step.

    python tools/sim_drain_lanes.py [tiles]

Model: one tile = the points of a cell of the bench workload (10^7 points, L = 2000, 49^3 cells, Morton order inside the cell,
up to 96 points, 3 per lane); the secondary points of the half stencil arrive row by row; points are classified against the
tile's bounding box as in count_kernel_cl.cuh and only the PARTIAL ones (neither dropped nor dense) go through the stacks, two
per step; a drain starts when the fullest stack has less than 2R free slots and pops (fullest - keep) entries from every stack.
Printed: useful pops / pop slots.  Round-2 result: 0.57-0.69 for every static assignment, 0.84-0.93 with rotating columns."""
import sys
import numpy as np

rng = np.random.default_rng(5)
cs = 2000 / 49; rmax = 200.0; dens = 1e7 / 2000 ** 3
ntiles = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def morton(p, lo):
    q = np.clip(((p - lo) / cs * 4).astype(int), 0, 3)
    m = np.zeros(len(p), int)
    for b in (1, 0):
        m = (m << 3) | (((q[:, 0] >> b) & 1) << 2) | (((q[:, 1] >> b) & 1) << 1) | ((q[:, 2] >> b) & 1)
    return m


def cell_points(ix, iy, iz):
    n = rng.poisson(dens * cs ** 3)
    lo = np.array([ix, iy, iz]) * cs
    p = lo + rng.random((n, 3)) * cs
    return p[np.argsort(morton(p, lo), kind="stable")]


K = 5
rows = []
for dx in range(-K - 1, K + 2):
    for dy in range(-K - 1, K + 2):
        zs = [dz for dz in range(-K - 1, K + 2) if (max(abs(dx) - 1, 0) ** 2 + max(abs(dy) - 1, 0) ** 2 + max(abs(dz) - 1, 0) ** 2) * cs * cs < rmax ** 2]
        if not zs or dx < 0 or (dx == 0 and dy < 0):
            continue
        if dx == 0 and dy == 0:
            zs = [z for z in zs if z > 0]
        if zs:
            rows.append((dx, dy, min(zs), max(zs)))


def simulate(assign, depth=24, keep=6, R=3, rotate=False):
    slots = useful = 0
    for _ in range(ntiles):
        P = cell_points(0, 0, 0)[: 32 * R]
        idx = assign(len(P), R)
        lo, hi = P.min(0), P.max(0); c = (lo + hi) / 2; h = (hi - lo) / 2
        fill = np.zeros(32, int)
        buf = []

        def process(block):
            nonlocal slots, useful, fill
            B = np.array(block)
            for s in range(0, len(B), 2):
                pair = B[s:s + 2]
                if fill.max() > depth - 1 - 2 * R:
                    mx = fill.max(); rounds = min((mx - keep + 3) & ~3, 32)
                    pop = np.minimum(fill, rounds); fill -= pop
                    slots += 32 * rounds; useful += pop.sum()
                for r in range(R):
                    ok = idx[r] >= 0
                    d2 = ((P[np.where(ok, idx[r], 0)][:, None, :] - pair[None, :, :]) ** 2).sum(-1)
                    fill += ((d2 < rmax ** 2) & ok[:, None]).sum(1)
                if rotate:
                    fill = np.roll(fill, 1)
        for dx, dy, zlo, zhi in rows:
            for dz in range(zlo, zhi + 1):
                S = cell_points(dx, dy, dz)
                t = np.abs(S - c)
                dmin2 = (np.maximum(t - h, 0) ** 2).sum(1); dmax2 = ((t + h) ** 2).sum(1)
                buf.extend(S[(dmin2 <= rmax ** 2) & (dmax2 >= rmax ** 2)])
                while len(buf) >= 32:
                    process(buf[:32]); buf = buf[32:]
        if buf:
            process(buf)
        slots += 32 * ((fill.max() + 3) & ~3); useful += fill.sum()
    return useful / slots


def a_plain(n, R):           # lane l holds the points l, 32 + l, 64 + l of the Morton order (what the kernels do)
    return np.array([[r * 32 + l if r * 32 + l < n else -1 for l in range(32)] for r in range(R)])


def a_reversed(n, R):        # odd groups reversed
    return np.array([[r * 32 + (l if r % 2 == 0 else 31 - l) if r * 32 + (l if r % 2 == 0 else 31 - l) < n else -1 for l in range(32)] for r in range(R)])


def a_mirrored(n, R):        # lane l holds l and its mirror image n - 1 - l, the rest in order
    idx = -np.ones((R, 32), int)
    for l in range(32):
        if l < n:
            idx[0, l] = l
        if R > 1 and n - 1 - l >= 32:
            idx[1, l] = n - 1 - l
    rest = [k for k in range(n) if k not in set(idx.flatten())]
    for r in range(1, R):
        for l in range(32):
            if idx[r, l] < 0 and rest:
                idx[r, l] = rest.pop(0)
    return idx


if __name__ == "__main__":
    for name, f in (("plain", a_plain), ("reversed", a_reversed), ("mirrored", a_mirrored)):
        print(f"static assignment {name:9s}: depth 24 keep 6 -> {simulate(f):.3f}   depth 48 keep 6 -> {simulate(f, 48, 6):.3f}", flush=True)
    for depth, keep in ((24, 6), (24, 8), (24, 12), (28, 14)):
        print(f"rotating columns, depth {depth} keep {keep}: {simulate(a_plain, depth, keep, rotate=True):.3f}", flush=True)
