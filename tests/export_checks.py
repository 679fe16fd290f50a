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


def check_export_from_versions(blob, lib_path=None, seed=0, trials=6):
    """export(ExportMode::updates(from)) for random `from` version vectors: bytes equal to the oracle's export of a
    document that imported the same blob; importing them into a replica at `from` yields the full state."""
    import random
    import loro_b200
    from loro_b200 import api
    rnd = random.Random(seed)
    ref = OracleDoc(0xABCDEF)
    ref.import_(blob)
    vv = ref.oplog_vv()
    batch = loro_b200.import_batch([blob], flags=api.LB_FLAG_EXPORT, lib_path=lib_path)
    for trial in range(trials):
        frm = {p: rnd.randint(0, c) for p, c in vv.items() if rnd.random() < 0.8}
        if trial == 0:
            frm = dict(vv)                      # nothing to send: header-only blob
        elif trial == 1:
            frm = {p: 1 for p in vv}            # cut inside every peer's first op
        want = ref.export_updates(frm)
        got = batch.export_updates(0, frm)
        if got != want:
            k = next((j for j in range(min(len(got), len(want))) if got[j] != want[j]), min(len(got), len(want)))
            raise AssertionError(f"export from {frm}: differs from the oracle at byte {k} (lens {len(got)} / {len(want)})")
    assert batch.export_updates(0) == ref.export_updates()   # the import-time all_updates export is untouched
    return batch
