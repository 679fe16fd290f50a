"""Shared parity check for persistent documents (lb_docset_*): a stream of update blobs against documents that already
hold history, compared after EVERY import with an oracle document that took the same imports in the same order
(reference: LoroDoc::import / import_batch on an existing document, loro.rs:562-643, 1183-1290)."""
import random

from oracle import OracleDoc

from . import workloads


def _session(seed, n_sites, rounds, edits, stale_inside=False):
    """A live session of `n_sites` replicas of ONE document: every round each site edits, then publishes what it has
    that the server has not seen from it (export from the vv the site last published at).  Returns the blobs in
    publication order; some are held back and delivered late or twice, so imports see pending changes, overlaps,
    duplicates and changes the document already holds."""
    rnd = random.Random(seed)
    sites = [OracleDoc(1000 * seed + 7 + k) for k in range(n_sites)]
    handles = []
    for d in sites:
        handles.append((d.get_text("text"), d.get_list("list"), d.get_map("map")))
    published = [dict() for _ in sites]
    older = [[dict()] for _ in sites]   # versions a site published at before
    stream = []
    for r in range(rounds):
        for k, d in enumerate(sites):
            t, l, m = handles[k]
            for _ in range(rnd.randint(1, edits)):
                workloads.random_edit(rnd, d, t, l, m)
                if rnd.random() < 0.3:
                    d.commit()
            d.commit()
            if rnd.random() < 0.8:
                frm = dict(published[k])
                if rnd.random() < 0.3 and frm:
                    if stale_inside:                     # a sender whose idea of the receiver is off by a few ops: the
                        frm = {p: max(0, c - rnd.randint(1, 5)) for p, c in frm.items()}   # update starts INSIDE known changes
                    else:                                # overlap: re-send from a version published earlier
                        frm = dict(rnd.choice(older[k]))
                stream.append(d.export_updates(frm))
                older[k].append(dict(published[k]))
                published[k] = d.oplog_vv()
        if rnd.random() < 0.7:                           # sites talk to each other too: later updates depend on others'
            a, b = rnd.sample(range(n_sites), 2)
            workloads.merge(sites[a], sites[b])
    for k, d in enumerate(sites):
        stream.append(d.export_updates(published[k]))
    # delivery: mostly in order, some late (-> pending at the receiver), some twice
    out = []
    late = []
    for blob in stream:
        x = rnd.random()
        if x < 0.2:
            late.append(blob)
        else:
            out.append(blob)
            if x > 0.9:
                out.append(blob)
        if late and rnd.random() < 0.3:
            out.append(late.pop(rnd.randrange(len(late))))
    out.extend(late)
    return out


def check_docset_against_oracle(lib_path=None, n_docs=4, seed=0, rounds=6, edits=12, export_parity=True, compact=False,
                                stale_inside=False, require_coverage=True):
    """compact=False: the documents keep every blob (exported bytes equal the reference's after the same sequence of
    imports).  compact=True: every import carries LB_FLAG_COMPACT, i.e. a document without pending changes is replaced
    by a fresh one that imported its own export -- the oracle documents do exactly that after every import.
    stale_inside=True adds updates that start INSIDE changes the document already holds while a later part of the same
    range is known from another blob: which copy supplies an atom then decides where its payload sits in the arenas and
    so which ops of the export re-merge -- the engine takes every atom from the copy the reference would (k_resolve.cuh
    pick_copy_multi), exported bytes included."""
    import loro_b200
    from loro_b200 import api
    rnd = random.Random(77 + seed)
    streams = [_session(seed * 10 + d, 2 + d % 3, rounds, edits, stale_inside=stale_inside) for d in range(n_docs)]
    refs = [OracleDoc(0xD0C + d) for d in range(n_docs)]
    ds = loro_b200.DocSet(lib_path=lib_path)
    cursors = [0] * n_docs
    steps = 0
    saw_pending = saw_known = 0
    while any(c < len(s) for c, s in zip(cursors, streams)):
        # one call = a few documents, each getting one blob (import) or several (import_batch)
        blobs, ids, per_doc = [], [], {}
        for d in range(n_docs):
            if cursors[d] >= len(streams[d]) or rnd.random() < 0.3:
                continue
            k = 1 if rnd.random() < 0.6 else rnd.randint(2, 3)
            take = streams[d][cursors[d]:cursors[d] + k]
            cursors[d] += len(take)
            per_doc[d] = take
        order = [(d, b) for d, take in per_doc.items() for b in take]
        # interleave the documents of the call; the blobs of ONE document keep their order
        docs_in_call = list(per_doc)
        rnd.shuffle(docs_in_call)
        for d in docs_in_call:
            for b in per_doc[d]:
                blobs.append(b)
                ids.append(500 + d)
        if not blobs:
            continue
        batch = ds.import_(blobs, ids, flags=api.LB_FLAG_COMPACT if compact else 0)
        assert batch.n_docs == len(docs_in_call)
        for slot, d in enumerate(docs_in_call):
            ost = refs[d].import_batch(per_doc[d])
            st = batch.status(slot)
            assert st.code == 0, (d, st)
            assert st.success == ost["success"], (steps, d, st.success, ost["success"])
            assert st.pending == ost["pending"], (steps, d, st.pending, ost["pending"])
            saw_pending += ost["pending"] is not None
            saw_known += not ost["success"]
            assert batch.json_bytes(slot) == refs[d].json_text(), (steps, d)
            assert batch.oplog_vv(slot) == refs[d].oplog_vv(), (steps, d)
            assert batch.oplog_frontiers(slot) == sorted(refs[d].frontiers()), (steps, d)
            if export_parity:
                got, want = batch.export_updates(slot), refs[d].export_updates()
                assert got == want, (steps, d, len(got), len(want))
            else:   # same document either way: what the engine exports imports to the reference's state
                again = OracleDoc(0xA6A1)
                again.import_(batch.export_updates(slot))
                assert again.json_text() == refs[d].json_text() and again.oplog_vv() == refs[d].oplog_vv(), (steps, d)
            if compact and refs[d].pending_count() == 0:
                fresh = OracleDoc(0xD0C + d)
                fresh.import_(refs[d].export_updates())
                refs[d] = fresh
        batch.close()
        steps += 1
    assert ds.n_docs == n_docs
    if require_coverage:   # the stream must have exercised pending changes and imports that added nothing
        assert saw_pending > 0 and (compact or saw_known > 0), (saw_pending, saw_known)
    ds.close()
    return steps
