"""Seeded synthetic workloads (inputs for bench.py and the tests), in the reference's FastUpdates format.

Host-side input tooling: nothing here runs inside the measured path, and nothing here uses oracle/."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libloro_workload.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "gen.cpp")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = ctypes.CDLL(_LIB)
        L.lw_generate_c3.restype = ctypes.c_void_p
        L.lw_generate_c3.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64] + [ctypes.c_int] * 7
        L.lw_generate_c5.restype = ctypes.c_void_p
        L.lw_generate_c5.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64] + [ctypes.c_int] * 7
        L.lw_generate_c4.restype = ctypes.c_void_p
        L.lw_generate_c4.argtypes = [ctypes.c_uint64] + [ctypes.c_int] * 4
        L.lw_bytes.restype = ctypes.c_void_p
        L.lw_bytes.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.lw_offsets.restype = ctypes.c_void_p
        L.lw_offsets.argtypes = [ctypes.c_void_p]
        L.lw_lens.restype = ctypes.c_void_p
        L.lw_lens.argtypes = [ctypes.c_void_p]
        L.lw_atoms.restype = ctypes.c_uint64
        L.lw_atoms.argtypes = [ctypes.c_void_p]
        L.lw_json.restype = ctypes.c_void_p
        L.lw_json.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
        L.lw_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


class _Batch:
    """A generated batch: blobs at 16-byte aligned starts inside one buffer (the layout lb_import_batch_device takes)."""

    def _adopt(self, handle, n_docs, want_json):
        L = _load()
        self._h = handle
        total = ctypes.c_uint64()
        p = L.lw_bytes(self._h, ctypes.byref(total))
        self.n_docs = n_docs
        self.bytes = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(total.value,))
        self.offsets = np.ctypeslib.as_array(ctypes.cast(L.lw_offsets(self._h), ctypes.POINTER(ctypes.c_uint64)),
                                             shape=(n_docs,)) if n_docs else np.zeros(0, np.uint64)
        self.lens = np.ctypeslib.as_array(ctypes.cast(L.lw_lens(self._h), ctypes.POINTER(ctypes.c_uint32)),
                                          shape=(n_docs,)) if n_docs else np.zeros(0, np.uint32)
        self.atom_ops = L.lw_atoms(self._h)
        self.want_json = want_json

    def blob(self, i):
        o, n = int(self.offsets[i]), int(self.lens[i])
        return self.bytes[o:o + n].tobytes()

    def blobs(self):
        return [self.blob(i) for i in range(self.n_docs)]

    def expected_json(self, i):
        n = ctypes.c_uint64()
        p = _load().lw_json(self._h, i, ctypes.byref(n))
        return ctypes.string_at(p, n.value)

    def close(self):
        if getattr(self, "_h", None):
            _load().lw_free(self._h)
            self._h = None
            self.bytes = self.offsets = self.lens = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class C4Doc(_Batch):
    """Config C4: ONE Text document -- peer 0 inserts `base_chars` ASCII characters, then `n_peers` peers
    each make `edits` edits (70 % insert of 1-8 characters, 30 % delete of 1-8) on their own copy of that base without
    ever syncing; the blob holds all branches (what export(all_updates) of a replica that received everything yields)."""

    def __init__(self, base_chars=1000000, n_peers=64, edits=50000, txn_ops=10, seed=0):
        h = _load().lw_generate_c4(seed, base_chars, n_peers, edits, txn_ops)
        self._adopt(h, 1, False)
        self.config = dict(base_chars=base_chars, n_peers=n_peers, edits=edits, txn_ops=txn_ops, seed=seed)


class C5Batch(_Batch):
    """Config C5: per doc peer 0 builds a tree of `n_nodes` nodes (fan-out <= `max_fanout`), then each
    of the `n_peers` peers issues `n_moves` concurrent moves (random target, random new parent; moves of different
    peers may close cycles); seed = doc index.  One FastUpdates blob per doc."""

    def __init__(self, n_docs, n_nodes=5000, n_peers=3, n_moves=1000, max_fanout=8, txn_ops=10, first_doc=0,
                 seed_base=0, want_json=False, threads=None):
        L = _load()
        threads = threads or os.cpu_count() or 1
        h = L.lw_generate_c5(seed_base, first_doc, n_docs, n_nodes, n_peers, n_moves, max_fanout, txn_ops,
                             1 if want_json else 0, threads)
        self._adopt(h, n_docs, want_json)
        self.config = dict(n_docs=n_docs, n_nodes=n_nodes, n_peers=n_peers, n_moves=n_moves, max_fanout=max_fanout,
                           txn_ops=txn_ops, first_doc=first_doc, seed_base=seed_base)


class C3Batch(_Batch):
    """Config C3: per doc `n_peers` peers, `n_ops` mixed List/Map atom ops (60 % list insert of an
    I64 or short Str, 15 % list delete of 1-4, 25 % map set on 16 keys with 10 % deletes); peers fork from a
    common prefix of `prefix_ops`, edit concurrently in bursts and sync pairwise every ~`sync_every` ops;
    seed = doc index.  One FastUpdates blob per doc = export(all_updates) of a fully synced replica."""

    def __init__(self, n_docs, n_ops=10000, n_peers=3, prefix_ops=1000, sync_every=500, txn_ops=10,
                 first_doc=0, seed_base=0, want_json=False, threads=None):
        L = _load()
        threads = threads or os.cpu_count() or 1
        h = L.lw_generate_c3(seed_base, first_doc, n_docs, n_ops, n_peers, prefix_ops, sync_every, txn_ops,
                             1 if want_json else 0, threads)
        self._adopt(h, n_docs, want_json)
        self.config = dict(n_docs=n_docs, n_ops=n_ops, n_peers=n_peers, prefix_ops=prefix_ops,
                           sync_every=sync_every, txn_ops=txn_ops, first_doc=first_doc, seed_base=seed_base)
