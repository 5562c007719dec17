"""Write the bench catalogue as an ASCII file for the FCFC command line: python tools/make_box_ascii.py N L path"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n, L, path = int(float(sys.argv[1])), float(sys.argv[2]), sys.argv[3]
x = bench.make_box(n, L)
a = np.stack(x, 1)
with open(path, "w") as f:
    for i in range(0, n, 1_000_000):
        np.savetxt(f, a[i:i + 1_000_000], fmt="%.6f")
