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
