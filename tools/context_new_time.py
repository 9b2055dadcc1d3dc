#!/usr/bin/env python
"""Times futhark_context_new (second context of the process: CUDA itself is already initialised) for a given build of the
library, through ctypes only.  Usage: context_new_time.py [path/to/lib.so]"""
import ctypes as C
import os
import sys
import time

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytracers_b200", "libray_b200.so")
L = C.CDLL(path)
for f in ("futhark_context_config_new", "futhark_context_new"):
    getattr(L, f).restype = C.c_void_p
L.futhark_context_new.argtypes = [C.c_void_p]
L.futhark_context_free.argtypes = [C.c_void_p]
L.futhark_context_config_free.argtypes = [C.c_void_p]
ms = []
for i in range(4):
    cfg = L.futhark_context_config_new()
    t = time.perf_counter()
    ctx = L.futhark_context_new(cfg)
    ms.append((time.perf_counter() - t) * 1e3)
    L.futhark_context_free(ctx)
    L.futhark_context_config_free(cfg)
print(os.path.basename(path), "futhark_context_new ms: first", round(ms[0], 1), "later", [round(x, 2) for x in ms[1:]], "size MB", round(os.path.getsize(path) / 1e6, 2))
