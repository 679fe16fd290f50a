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
