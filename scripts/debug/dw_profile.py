"""Reads the dW kernel's per-CTA cycle counters (debug build: NRF_NVCC_EXTRA=-DNRF_PROFILE_DW python -m nerfpp_b200.build --force)."""
import ctypes
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from nerfpp_b200 import cabi  # noqa: E402

lib = cabi.lib()
buf = (ctypes.c_ulonglong * (148 * 8))()
bench.classic_nerf_leg(1664.5, reps=1)       # 2 backward launches (warm-up + 1)
lib.nrf_debug_dw_profile(buf)                # clears
bench.classic_nerf_leg(1664.5, reps=1)
lib.nrf_debug_dw_profile(buf)
a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 8).astype(np.float64) / 2     # per launch
names = ["issuer waits full", "issuer waits d_free", "producer waits empty", "kernel cycles (clock64)", "kernel ns (globaltimer)", "epi waits d_ready (+bias flush)",
         "flush total (incl. d_ready wait)", "items"]
for i, nm in enumerate(names):
    print(f"{nm:36s} mean {a[:, i].mean():12.0f}  min {a[:, i].min():12.0f}  max {a[:, i].max():12.0f}")
order = np.argsort(-a[:, 3])
print("cta   cycles   items  wait_full  wait_dfree  flush   dready")
for i in list(order[:12]) + list(order[-6:]):
    print(f"{i:3d} {a[i,3]:9.0f} {a[i,7]:6.0f} {a[i,0]:9.0f} {a[i,1]:9.0f} {a[i,6]:9.0f} {a[i,5]:9.0f}")
