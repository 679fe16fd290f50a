"""Persistent documents on the emulated build: imports against existing document state (lb_docset_*)."""
import os
import subprocess

import pytest

from .docset_checks import check_docset_against_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu", "libloro_b200_emu.so")


@pytest.fixture(scope="module", autouse=True)
def _emu():
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])


@pytest.mark.parametrize("seed", range(3))
def test_docset_stream_of_updates_matches_persistent_oracle_documents(seed):
    steps = check_docset_against_oracle(lib_path=EMU, n_docs=4, seed=seed)
    assert steps > 5


def test_docset_same_shape_as_the_gpu_test():
    check_docset_against_oracle(lib_path=EMU, n_docs=12, seed=1, rounds=8, edits=16)


@pytest.mark.parametrize("seed", range(2))
def test_docset_updates_that_start_inside_known_changes(seed):
    """a sender whose idea of the receiver is stale: its update starts inside changes the document holds, and part of
    the rest is known from another blob -- every atom must come from the copy the reference applies (arrival order, not
    counter order), or the exported bytes differ by an op merge"""
    check_docset_against_oracle(lib_path=EMU, n_docs=6, seed=seed, rounds=8, edits=16, stale_inside=True)


@pytest.mark.parametrize("seed", range(2))
def test_docset_compaction_equals_recreating_the_document_from_its_export(seed):
    steps = check_docset_against_oracle(lib_path=EMU, n_docs=4, seed=20 + seed, compact=True)
    assert steps > 5


def test_docset_failed_import_keeps_the_earlier_state():
    """a damaged update (checksum) or a malformed one fails ITS import only: the reference rejects it before any state
    change (loro.rs:584); the next good import continues from the earlier state"""
    import loro_b200
    from oracle import OracleDoc
    from tests import workloads
    a, b = OracleDoc(1), OracleDoc(2)
    t = a.get_text("text")
    a.text_insert(t, 0, "hello world")
    a.commit()
    u1 = a.export_updates()
    workloads.merge(b, a)
    b.text_insert(b.get_text("text"), 5, ",")
    b.commit()
    u2 = b.export_updates(a.oplog_vv())
    ref = OracleDoc(9)
    ds = loro_b200.DocSet(lib_path=EMU)
    r = ds.import_([u1], [7])
    ref.import_(u1)
    assert r.status(0).code == 0 and r.json_bytes(0) == ref.json_text()
    stored = ds.stored_bytes
    bad = bytearray(u2)
    bad[-1] ^= 0x40                                   # checksum no longer matches
    r = ds.import_([bytes(bad)], [7])
    assert r.status(0).code == 2                      # LB_DOC_ERR_CHECKSUM, like LoroError::DecodeChecksumMismatchError
    assert ds.stored_bytes == stored and ds.n_docs == 1
    r = ds.import_([b"loro" + bytes(40)], [7])        # right magic, wrong everything else
    assert r.status(0).code != 0 and ds.stored_bytes == stored
    r = ds.import_([u2], [7])                         # the good update still applies on top of the earlier state
    ost = ref.import_(u2)
    st = r.status(0)
    assert st.code == 0 and st.success == ost["success"] and st.pending == ost["pending"]
    assert r.json_bytes(0) == ref.json_text() and r.oplog_vv(0) == ref.oplog_vv()
    assert r.export_updates(0) == ref.export_updates()
    assert ds.stored_bytes > stored
