import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        print(f, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "mlp_bwd ms", d["kernels_ms_per_step"].get("mlp_small_bwd"), d["final_loss"])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
