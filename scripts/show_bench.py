import json,sys
d=json.loads(sys.stdin.read()); print(d["config"]["docs_per_gpu"], round(d["value"]/1e6,1), "Mops/s", {k:round(v,2) for k,v in d["phases_ms"].items()}, "dec_frac", round(d["decode_roofline"]["frac"],4), "dev_GB", round(d["config"].get("device_table_bytes_per_step",0)/1e9,1))
