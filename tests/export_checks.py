"""Shared export parity check: the engine's lb_doc_export_updates bytes against the oracle's
import -> export(all_updates) of the same blob (used by the emulated and the GPU test files)."""
from oracle import OracleDoc


def check_export_against_oracle(blobs, lib_path=None, reimport=True):
    import loro_b200
    from loro_b200 import api
    batch = loro_b200.import_batch(blobs, flags=api.LB_FLAG_EXPORT, lib_path=lib_path)
    outs = []
    for i, blob in enumerate(blobs):
        ref = OracleDoc(0xABCDEF)
        ref.import_(blob)
        want = ref.export_updates()
        got = batch.export_updates(i)
        if got != want:
            k = next((j for j in range(min(len(got), len(want))) if got[j] != want[j]), min(len(got), len(want)))
            raise AssertionError(f"doc {i}: export differs from the oracle at byte {k} (lens {len(got)} / {len(want)})")
        outs.append(got)
    if reimport:
        # round trip: what we exported imports to the same state and exports to the same bytes (idempotence)
        again = loro_b200.import_batch(outs, flags=api.LB_FLAG_EXPORT, lib_path=lib_path)
        for i in range(len(blobs)):
            assert again.status(i).code == 0
            assert again.json_bytes(i) == batch.json_bytes(i), i
            assert again.oplog_vv(i) == batch.oplog_vv(i), i
            assert again.export_updates(i) == outs[i], i
    return batch
