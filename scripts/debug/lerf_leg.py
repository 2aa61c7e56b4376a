"""Runs the LeRF leg of bench.py alone (for ncu captures and quick timing): python scripts/debug/lerf_leg.py [reps]"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import bench  # noqa: E402

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    print(json.dumps(bench.lerf_leg(1666.0, reps=reps), indent=1))
