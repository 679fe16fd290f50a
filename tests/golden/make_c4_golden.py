"""Golden answers for BASELINE config C4 AT STATED SIZE (one Text document: 1 M base chars + 64 peers x 50 k concurrent
edits): the oracle imports the generated blob once (minutes on one core) and the digests of what it answers are committed
as tests/golden/c4_full.json, so that the GPU test can check the engine's state and exported bytes for the full
document without running the oracle on the GPU box.  Usage: python tests/golden/make_c4_golden.py [quarter]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
from oracle import OracleDoc                    # noqa: E402
from loro_b200.workload import C4Doc            # noqa: E402


def xxh32(b, seed=0):
    return oracle.i64s(oracle.codec("xxh32", bytes(b), seed))[0] & 0xFFFFFFFF


def main():
    # `quarter` = 500 k base chars + 64 peers x 12.5 k edits (the oracle needs ~5 minutes); default = the stated size (~1 h)
    quarter = len(sys.argv) > 1 and sys.argv[1] == "quarter"
    cfg = dict(base_chars=500000, n_peers=64, edits=12500, seed=0) if quarter else dict(base_chars=1000000, n_peers=64, edits=50000, seed=0)
    g = C4Doc(**cfg)
    blob = g.blob(0)
    t0 = time.time()
    d = OracleDoc(1)
    st = d.import_(blob)
    js = d.json_text()
    ex = d.export_updates()
    out = {"config": cfg, "blob_len": len(blob), "blob_xxh32": xxh32(blob), "atom_ops": int(g.atom_ops),
           "json_len": len(js), "json_xxh32": xxh32(js), "state_hash": (xxh32(js) << 32) | len(js),
           "export_len": len(ex), "export_xxh32": xxh32(ex), "vv": {str(k): v for k, v in d.oplog_vv().items()},
           "pending": st["pending"], "oracle_seconds": round(time.time() - t0, 1),
           "inconsistent_delete": d.inconsistent_delete()}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c4_quarter.json" if quarter else "c4_full.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
