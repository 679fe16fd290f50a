"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck) on a real GPU:
    compute-sanitizer --tool racecheck python tests/tools/sanitizer_probe.py
Covers the tracker (concurrent sites, version switches, splits), JSON, re-export, grouping by doc_id."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import loro_b200                      # noqa: E402
from loro_b200 import api             # noqa: E402
from loro_b200.workload import C3Batch  # noqa: E402
from tests import workloads           # noqa: E402


def main():
    blobs, js = [], []
    for i in range(6):
        blob, j, _, _ = workloads.make_doc_history(9100 + i, n_sites=3 + i % 3, n_ops=250)
        blobs.append(blob)
        js.append(j)
    gen = C3Batch(6, n_ops=1500, threads=2, want_json=True)
    blobs += gen.blobs()
    js += [gen.expected_json(i) for i in range(gen.n_docs)]
    b = loro_b200.import_batch(blobs, flags=api.LB_FLAG_EXPORT)
    for i in range(len(blobs)):
        assert b.status(i).code == 0
        assert b.json_bytes(i) == js[i], i
        b.export_updates(i)
    again = loro_b200.import_batch([b.export_updates(i) for i in range(len(blobs))], flags=api.LB_FLAG_EXPORT)
    assert again.counters()["state_hash"] == b.counters()["state_hash"]
    print("sanitizer probe ok", b.counters()["atom_ops"])


if __name__ == "__main__":
    main()
