import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.argv = ['bench.py']
import bench
for w in ("dropin_cpp", "dropin_cpp_staged", "reference_cuda"):
    try:
        print(w, json.dumps(bench.cpp_render_leg(w)))
    except Exception as e:
        print(w, "ERR", type(e).__name__, e)
