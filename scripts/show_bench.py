"""One-line summaries of bench.py JSON lines: python scripts/show_bench.py FILE..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        print(path, d["config"].get("name"), d["config"]["docs_per_gpu"], round(d["value"] / 1e6, 1), "Mops/s",
              {k: round(v, 2) for k, v in d["phases_ms"].items()}, "dec_frac", round(d["decode_roofline"]["frac"], 4),
              "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 1))
    except Exception as e:   # a missing or truncated line must not stop the script that calls this
        print(path, "unreadable:", e)
