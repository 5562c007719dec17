"""One-shot upload against streamed ingest of a 10^7-point catalogue (float and double columns).

    python tools/time_ingest.py [n] [rows_per_chunk]

One-shot: fcfc_gpu_catalog_create from pageable arrays (what the unmodified host does after the reader has finished).
Streamed: fcfc_gpu_catalog_stream_append per chunk of rows (a 1 MiB chunk of an ASCII catalogue holds about 3x10^4 lines),
then _finish.  The append calls are what the reader's chunk loop would pay (a host memcpy into pinned memory per chunk,
spread over the parsing); `finish` is what remains after the last line has been parsed."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import fcfc_b200 as F  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
rows = int(float(sys.argv[2])) if len(sys.argv) > 2 else 32768
F.init()
rng = np.random.default_rng(1)
for prec in ("float", "double"):
    b = F.Bins(periodic=True, prec=prec, arith=1, box=2000.0, bintype=1, smin=0.0, smax=200.0, ds=5.0, nmu=120)
    cols = [np.ascontiguousarray(rng.random(n) * 2000.0, dtype=b.dtype) for _ in range(3)]
    out = {"prec": prec, "n": n, "rows_per_chunk": rows}
    for rep in range(3):
        t0 = time.perf_counter()
        c = F.Catalog(*cols, bins=b)
        out["one_shot_ms"] = min(out.get("one_shot_ms", 1e9), 1e3 * (time.perf_counter() - t0))
        c.destroy()
    L = F.lib()
    for rep in range(3):
        t0 = time.perf_counter()
        h = L.fcfc_gpu_catalog_stream_begin(n, int(b.is_float), 0)
        t1 = time.perf_counter()
        item = cols[0].itemsize
        for lo in range(0, n, rows):
            m = min(rows, n - lo)
            rc = L.fcfc_gpu_catalog_stream_append(h, cols[0].ctypes.data + lo * item, cols[1].ctypes.data + lo * item,
                                                  cols[2].ctypes.data + lo * item, None, m)
            assert rc == 0, F.last_error()
        t2 = time.perf_counter()
        cat = L.fcfc_gpu_catalog_stream_finish(h, float(b.rescale), -1)
        t3 = time.perf_counter()
        assert cat, F.last_error()
        L.fcfc_gpu_catalog_destroy(cat)
        if rep == 0 or 1e3 * (t3 - t2) < out["stream_finish_ms"]:
            out.update(stream_begin_ms=1e3 * (t1 - t0), stream_append_total_ms=1e3 * (t2 - t1),
                       stream_append_us_per_chunk=1e6 * (t2 - t1) / ((n + rows - 1) // rows), stream_finish_ms=1e3 * (t3 - t2))
    print(json.dumps(out), flush=True)
