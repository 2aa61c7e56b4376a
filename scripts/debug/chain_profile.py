"""Reads the chain kernel's per-CTA cycle counters (debug build: NRF_NVCC_EXTRA=-DNRF_PROFILE_CHAIN python -m nerfpp_b200.build --force)."""
import ctypes
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402
from nerfpp_b200 import cabi  # noqa: E402

lib = cabi.lib()
buf = (ctypes.c_ulonglong * (148 * 8))()
bench.classic_nerf_leg(1664.5, reps=1)
lib.nrf_debug_chain_profile(buf)
bench.classic_nerf_leg(1664.5, reps=1)
lib.nrf_debug_chain_profile(buf)
a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 8).astype(np.float64) / 2
names = ["issuer waits a_ready", "issuer waits full", "issuer issue+commit", "epilogue waits slab_ready", "epilogue work (ld/pack/st)", "producer expect_tx + TMA issue",
         "epilogue warp total", "producer waits empty"]
tiles = 1536 / 148
for i, nm in enumerate(names):
    print(f"{nm:30s} mean {a[:, i].mean():10.0f}  per tile-step {a[:, i].mean() / tiles / 10:8.0f}  min {a[:, i].min():10.0f}  max {a[:, i].max():10.0f}")
