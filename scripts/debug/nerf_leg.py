"""Runs bench.py's classic-NeRF leg alone (for ncu): python scripts/debug/nerf_leg.py [reps]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402

print(json.dumps(bench.classic_nerf_leg(1664.5, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 10)))
