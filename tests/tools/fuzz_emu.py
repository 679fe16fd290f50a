"""Mutation fuzzer for the kernels (run by tests/test_fuzz_emu.py in a subprocess with LB_EMU_GUARD=1, where every
"device" allocation of the emulated build sits between inaccessible pages): blobs of valid documents get a few
bytes changed -- mostly the value bits of varints, so that the framing survives and the *values* (indices,
counters, positions, lengths) go wild -- and the header checksum is recomputed so that they pass the frame
phase.  The engine must answer with a per-document code or a state, never touch memory it does not own."""
import random
import struct
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import loro_b200                      # noqa: E402
import oracle                         # noqa: E402
from loro_b200 import api             # noqa: E402
from tests import workloads           # noqa: E402


def reseal(blob):
    h = oracle.i64s(oracle.codec("xxh32", bytes(blob[20:]), 0x4F524F4C))[0] & 0xFFFFFFFF
    return bytes(blob[:16]) + struct.pack("<I", h) + bytes(blob[20:])


def main(lib, seed, n):
    rnd = random.Random(seed)
    base = [workloads.make_doc_history(100 + i, n_sites=3, n_ops=120)[0] for i in range(4)]
    # documents whose blocks carry one kind of payload: the lane-parallel decode path (k_decode_warp.cuh) and the tree kernels
    from loro_b200.workload import C3Batch, C5Batch
    base += C3Batch(2, n_ops=400, prefix_ops=60, sync_every=80).blobs() + C5Batch(2, n_nodes=120, n_moves=40).blobs()
    base += [workloads.make_tree_history(7, n_sites=3, n_base=20, n_ops=60)[0]]
    if os.environ.get("LB_FUZZ_WIDE"):   # more shapes: many sites, longer histories, an insert larger than a block
        from oracle import OracleDoc
        base += [workloads.make_doc_history(300 + i, n_sites=2 + i, n_ops=200 + 50 * i, sync_prob=0.1)[0] for i in range(4)]
        big = OracleDoc(9)
        big.text_insert(big.get_text("t"), 0, "w\u00e9 " * 2500)
        big.list_insert(big.get_list("l"), 0, *list(range(1500)))
        base.append(big.export_updates())
    # several blobs per document (import_batch groups / docset imports): overlapping and repeated updates of one history,
    # one of them mutated -- the copy selection, the applied-order lists and the status pass see inconsistent copies
    multi = []
    if os.environ.get("LB_FUZZ_MULTI"):
        for i in range(3):
            _, _, tot, sites = workloads.make_doc_history(500 + i, n_sites=3, n_ops=150)
            full = sites[0]
            parts = [full.export_updates({p: rnd.randrange(0, c + 1) for p, c in tot.items()}) for _ in range(3)]
            parts += [full.export_updates({q: c for q, c in tot.items() if q != p}) for p in list(tot)[:2]]
            multi.append(parts)
    ok = 0
    for it in range(n):
        if multi:
            parts = [bytes(x) for x in rnd.choice(multi)]
            rnd.shuffle(parts)
            k = rnd.randrange(len(parts))
            b = bytearray(parts[k])
            for _ in range(rnd.randint(1, 3)):
                i = rnd.randrange(22, len(b))
                b[i] = (b[i] & 0x80) | rnd.randrange(128) if rnd.random() < 0.7 else rnd.randrange(256)
            parts[k] = reseal(b)
            if it % 2:
                r = loro_b200.import_batch(parts, doc_ids=[1] * len(parts), flags=api.LB_FLAG_EXPORT, lib_path=lib)
                rs = [r]
            else:       # the same through a docset, two calls
                ds = loro_b200.DocSet(lib_path=lib)
                rs = [ds.import_(parts[:2], [1, 1]), ds.import_(parts[2:], [1] * (len(parts) - 2))]
            for r in rs:
                if r.status(0).code == 0:
                    ok += 1
                    r.json_bytes(0)
                    try:
                        r.export_updates(0)
                    except api.EngineError:
                        pass
                r.close()
            if it % 2 == 0:
                ds.close()
            continue
        b = bytearray(rnd.choice(base))
        for _ in range(rnd.randint(1, 4)):
            i = rnd.randrange(22, len(b))
            mode = rnd.random()
            if mode < 0.6:
                b[i] = (b[i] & 0x80) | rnd.randrange(128)
            elif mode < 0.85:
                b[i] = rnd.randrange(256)
            else:
                b[i] ^= 1 << rnd.randrange(8)
        blob = reseal(b)
        if os.environ.get("LB_FUZZ_SAVE"):
            open(os.environ["LB_FUZZ_SAVE"], "wb").write(blob)
        r = loro_b200.import_batch([blob], flags=api.LB_FLAG_EXPORT, lib_path=lib)
        if r.status(0).code == 0:
            ok += 1
            r.json_bytes(0)
            try:
                r.export_updates(0)
            except api.EngineError:
                pass
        r.close()
    print("fuzz done", n, "imported", ok)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
