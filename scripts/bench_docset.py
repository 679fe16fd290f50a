#!/usr/bin/env python3
"""Persistent documents (lb_docset_*): throughput of an update import against documents that already hold history.

Every document of config C3's shape exists twice: state A (the first half of its history, `--ops`/2 atom ops) and state
B (the whole history, what a peer that kept editing exports with all_updates; its changes overlap A's with other
boundaries, so the import dedupes and trims).  One step = a fresh docset takes A (untimed), then B (timed): host
blobs in, status + JSON + re-exported blobs resident / downloaded like bench.py's e2e leg.  Prints one JSON line.

  python scripts/bench_docset.py [--docs 8192] [--ops 10000] [--steps 3]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=8192)
    ap.add_argument("--ops", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import loro_b200
    from loro_b200.workload import C3Batch
    threads = len(os.sched_getaffinity(0))
    A = C3Batch(args.docs, n_ops=args.ops // 2, threads=threads)
    B = C3Batch(args.docs, n_ops=args.ops, threads=threads)
    a_blobs = [A.blob(i) for i in range(args.docs)]
    b_blobs = [B.blob(i) for i in range(args.docs)]
    ids = list(range(args.docs))
    fresh = loro_b200.import_batch(b_blobs, flags=loro_b200.api.LB_FLAG_EXPORT)
    want_hash, atoms_b = fresh.counters()["state_hash"], fresh.counters()["atom_ops"]
    fresh.close()
    t_fresh = []
    for _ in range(args.steps):
        t0 = time.time()
        fb = loro_b200.import_batch(b_blobs, flags=loro_b200.api.LB_FLAG_EXPORT)
        fb.json_bytes(0)
        torch.cuda.synchronize()
        t_fresh.append(time.time() - t0)
        fb.close()
    times, dev = [], []
    atoms_a = stored = 0
    for _ in range(args.steps + 1):
        ds = loro_b200.DocSet()
        b1 = ds.import_(a_blobs, ids)
        atoms_a = b1.counters()["atom_ops"]
        b1.close()
        stored = ds.stored_bytes
        torch.cuda.synchronize()
        t0 = time.time()
        b2 = ds.import_(b_blobs, ids)
        b2.json_bytes(0)
        torch.cuda.synchronize()
        times.append(time.time() - t0)
        c = b2.counters()
        assert c["docs_ok"] == args.docs and c["state_hash"] == want_hash, c   # same documents as a fresh import of B
        dev.append(b2.timings()["total_device"])
        b2.close()
        ds.close()
    times, dev = times[1:], dev[1:]      # first round = warm-up
    t = statistics.median(times)
    print(json.dumps({
        "what": "lb_docset_import: update blobs against documents resident in device memory",
        "docs": args.docs, "ops_per_doc": args.ops, "atoms_state_A": atoms_a, "atoms_state_B": atoms_b,
        "new_atom_ops_per_step": atoms_b - atoms_a, "stored_bytes_state_A": int(stored),
        "update_import_ms": t * 1e3, "update_import_device_ms": statistics.median(dev),
        "new_ops_per_s": (atoms_b - atoms_a) / t, "replayed_ops_per_s": atoms_b / t,
        "fresh_import_of_B_ms": statistics.median(t_fresh) * 1e3,
        "note": "an update import replays the stored history in front of the new blobs: it costs about a fresh import of state B plus the device-to-device copy of the stored blobs; state hash checked against the fresh import every step"}))


if __name__ == "__main__":
    main()
