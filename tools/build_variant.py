"""Build an experimental variant of libfcfc_b200.so with extra nvcc flags (e.g. -DFCFC_FLAG_MODE=2) into
fcfc_b200/_variants/<name>/libfcfc_b200.so (git-ignored; travels to the GPU box with gpurun).

    python tools/build_variant.py <name> [nvcc flags ...]
"""
import os, sys
from pathlib import Path
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcfc_b200 import build as fb

name, extra = sys.argv[1], sys.argv[2:]
out = fb.PKG / "_variants" / name
out.mkdir(parents=True, exist_ok=True)
fb.OBJ = out
fb.LIB = out / "libfcfc_b200.so"
fb.CFLAGS = fb.CFLAGS + extra
print(fb.build(verbose=True))
