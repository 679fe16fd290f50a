"""Which encoder variant the final lines use: the fastest re-export phase among the A/B lines of scripts/r2_final.sh,
if it beats the default build by >= 2 %.  Prints the variant name, or nothing for the default build."""
import json
import sys


def load(tag, n):
    try:
        return json.loads(open(f"gpurun_out/{tag}_ab_{n}.json").read().strip().splitlines()[-1])["phases_ms"]["reexport"]
    except Exception:
        return None


tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
base = load(tag, "default")
best, best_t = "", base
for n in ("xla0", "xenc5", "xenc6"):
    t = load(tag, n)
    if base and t and t < 0.98 * best_t:
        best, best_t = n, t
print(best)
