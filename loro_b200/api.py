"""ctypes binding of include/loro_b200.h (the same stub a cgo/N-API/Rust `extern "C"` shim would bind)."""
import ctypes
import json
import os
from collections import namedtuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT_LIB = os.path.join(_HERE, "libloro_b200.so")

LB_FLAG_NO_JSON = 1
LB_FLAG_KEEP_DEVICE = 2
LB_FLAG_EXPORT = 4
LB_FLAG_COMPACT = 8

DOC_CODES = {0: "Ok", 1: "DecodeError", 2: "DecodeChecksumMismatchError", 3: "IncompatibleFutureEncodingError",
             4: "DecodeDataCorruptionError", 5: "Unsupported", 6: "CapacityExceeded"}

ImportStatus = namedtuple("ImportStatus", "code success pending")


class EngineUnavailable(RuntimeError):
    """The CUDA library is missing or no CUDA device is present (there is no CPU fallback)."""


class DocError(RuntimeError):
    def __init__(self, code):
        super().__init__(DOC_CODES.get(code, str(code)))
        self.code = code


class _Blob(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_char_p), ("len", ctypes.c_size_t), ("doc_id", ctypes.c_uint64)]


class _Options(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 6)]


class _IdSpan(ctypes.Structure):
    _fields_ = [("peer", ctypes.c_uint64), ("start", ctypes.c_int32), ("end", ctypes.c_int32)]


class _Status(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int), ("n_success", ctypes.c_size_t), ("success", ctypes.POINTER(_IdSpan)),
                ("n_pending", ctypes.c_size_t), ("pending", ctypes.POINTER(_IdSpan))]


class _Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("docs", "docs_ok", "blob_bytes", "blocks", "changes", "op_rows",
                                                "atom_ops", "pending_changes", "json_bytes", "state_hash")]


class _Timings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("h2d", "frame", "decode", "resolve", "classify", "integrate",
                                               "materialise", "d2h", "total_device", "reexport")] + \
               [("decode_bytes_read", ctypes.c_uint64), ("decode_bytes_written", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("export_bytes", ctypes.c_uint64),
                ("tree", ctypes.c_float), ("reserved0", ctypes.c_uint32), ("tree_ops", ctypes.c_uint64),
                ("decode_fast_blocks", ctypes.c_uint64), ("decode_lane_blocks", ctypes.c_uint64),
                ("decode_unstaged_blocks", ctypes.c_uint64),
                ("alloc_host_ms", ctypes.c_float), ("reserved1", ctypes.c_uint32), ("device_bytes", ctypes.c_uint64),
                ("host_call_ms", ctypes.c_float), ("host_tail_ms", ctypes.c_float)]


_libs = {}


def library_path():
    return _DEFAULT_LIB


def load_library(path=None):
    path = path or os.environ.get("LORO_B200_LIB") or _DEFAULT_LIB
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise EngineUnavailable(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). loro_b200 has no CPU fallback.")
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    L.lb_import_batch.argtypes = [ctypes.POINTER(_Blob), ctypes.c_size_t, ctypes.POINTER(_Options), ctypes.POINTER(vp)]
    L.lb_import_batch_device.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32),
                                         ctypes.c_size_t, ctypes.POINTER(_Options), ctypes.POINTER(vp)]
    L.lb_doc_count.restype = ctypes.c_size_t
    L.lb_doc_count.argtypes = [vp]
    L.lb_doc_status.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(_Status)]
    L.lb_doc_json.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_size_t)]
    L.lb_doc_export_updates.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(_IdSpan), ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
    L.lb_doc_vv.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ctypes.POINTER(_IdSpan)), ctypes.POINTER(ctypes.c_size_t)]
    L.lb_doc_frontiers.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ctypes.POINTER(_IdSpan)), ctypes.POINTER(ctypes.c_size_t)]
    L.lb_batch_counters.argtypes = [vp, ctypes.POINTER(_Counters)]
    L.lb_batch_timings.argtypes = [vp, ctypes.POINTER(_Timings)]
    L.lb_last_error.restype = ctypes.c_char_p
    L.lb_batch_free.argtypes = [vp]
    L.lb_debug_table.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                 ctypes.POINTER(ctypes.c_size_t)]
    L.lb_docset_new.argtypes = [ctypes.POINTER(_Options), ctypes.POINTER(vp)]
    L.lb_docset_import.argtypes = [vp, ctypes.POINTER(_Blob), ctypes.c_size_t, ctypes.POINTER(_Options), ctypes.POINTER(vp)]
    L.lb_docset_doc_count.restype = ctypes.c_size_t
    L.lb_docset_doc_count.argtypes = [vp]
    L.lb_docset_stored_bytes.restype = ctypes.c_uint64
    L.lb_docset_stored_bytes.argtypes = [vp]
    L.lb_docset_free.argtypes = [vp]
    _libs[path] = L
    return L


class EngineError(RuntimeError):
    """A C-ABI call returned a non-zero lb_status; .status carries it (6 = LB_ERR_UNSUPPORTED)."""

    def __init__(self, msg, status):
        super().__init__(msg)
        self.status = status


def _check(L, rc, what):
    if rc == 0:
        return
    msg = L.lb_last_error().decode(errors="replace")
    if rc == 2:
        raise EngineUnavailable(f"{what}: {msg}")
    raise EngineError(f"{what} failed (lb_status={rc}): {msg}", rc)


class Batch:
    """Result of one batched import; owns the engine-side outputs until closed."""

    def __init__(self, L, handle, keep=None):
        self._L = L
        self._h = handle
        self._keep = keep  # keeps input buffers alive for device-resident imports
        self.n_docs = L.lb_doc_count(handle)

    def close(self):
        if self._h:
            self._L.lb_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def status(self, i):
        st = _Status()
        _check(self._L, self._L.lb_doc_status(self._h, i, ctypes.byref(st)), "lb_doc_status")
        suc = {st.success[k].peer: (st.success[k].start, st.success[k].end) for k in range(st.n_success)}
        pen = {st.pending[k].peer: (st.pending[k].start, st.pending[k].end) for k in range(st.n_pending)}
        return ImportStatus(st.code, suc, pen or None)

    def export_updates(self, i, from_vv=None):
        """LoroDoc::export(ExportMode::updates(from_vv)) of document i, all_updates when from_vv is None
        (needs flags=LB_FLAG_EXPORT at import).  from_vv: {peer: first counter the receiver lacks}."""
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        spans, k = None, 0
        if from_vv:
            k = len(from_vv)
            spans = (_IdSpan * k)()
            for j, (peer, ctr) in enumerate(from_vv.items()):
                spans[j].peer = peer
                spans[j].start = 0
                spans[j].end = ctr
        _check(self._L, self._L.lb_doc_export_updates(self._h, i, spans, k, ctypes.byref(p), ctypes.byref(n)),
               "lb_doc_export_updates")
        return ctypes.string_at(p.value, n.value)

    def json_bytes(self, i):
        p = ctypes.c_char_p()
        n = ctypes.c_size_t()
        _check(self._L, self._L.lb_doc_json(self._h, i, ctypes.byref(p), ctypes.byref(n)), "lb_doc_json")
        return ctypes.string_at(p, n.value)

    def fetch_json(self):
        """make sure the JSON of every document of the batch is in host memory (one download of the whole buffer)"""
        if self.n_docs:
            self.json_bytes(0)

    def fetch_exports(self):
        """same for the re-exported blobs (needs LB_FLAG_EXPORT)"""
        if self.n_docs:
            self.export_updates(0)

    def get_deep_value(self, i):
        st = self.status(i)
        if st.code != 0:
            raise DocError(st.code)
        return json.loads(self.json_bytes(i))

    def oplog_vv(self, i):
        spans = ctypes.POINTER(_IdSpan)()
        n = ctypes.c_size_t()
        _check(self._L, self._L.lb_doc_vv(self._h, i, ctypes.byref(spans), ctypes.byref(n)), "lb_doc_vv")
        return {spans[k].peer: spans[k].end for k in range(n.value)}

    def oplog_frontiers(self, i):
        """LoroDoc::oplog_frontiers(): sorted list of (peer, counter) head ids."""
        spans = ctypes.POINTER(_IdSpan)()
        n = ctypes.c_size_t()
        _check(self._L, self._L.lb_doc_frontiers(self._h, i, ctypes.byref(spans), ctypes.byref(n)), "lb_doc_frontiers")
        return sorted((spans[k].peer, spans[k].start) for k in range(n.value))

    def counters(self):
        c = _Counters()
        _check(self._L, self._L.lb_batch_counters(self._h, ctypes.byref(c)), "lb_batch_counters")
        return {n: getattr(c, n) for n, _ in _Counters._fields_}

    def timings(self):
        t = _Timings()
        _check(self._L, self._L.lb_batch_timings(self._h, ctypes.byref(t)), "lb_batch_timings")
        return {n: getattr(t, n) for n, _ in _Timings._fields_}

    def debug_table(self, name):
        import numpy as np
        n = ctypes.c_size_t()
        es = ctypes.c_size_t()
        _check(self._L, self._L.lb_debug_table(self._h, name.encode(), None, 0, ctypes.byref(n), ctypes.byref(es)),
               "lb_debug_table")
        dt = {1: np.uint8, 2: np.uint16, 4: np.int32, 8: np.int64}[es.value]
        arr = np.empty(n.value, dtype=dt)
        _check(self._L, self._L.lb_debug_table(self._h, name.encode(), arr.ctypes.data, arr.nbytes, ctypes.byref(n),
                                               ctypes.byref(es)), "lb_debug_table")
        return arr


class MultiBatch:
    """A large host batch imported as consecutive sub-batches, two C-ABI calls in flight: while one sub-batch computes,
    the next one's blobs go through the pinned staging ring and the previous one's JSON / exported blobs come home
    (documents are independent, so the split changes no result).  Same accessors as Batch; document i lives in the
    sub-batch that holds it."""

    def __init__(self, parts):
        self._parts = parts
        self._bounds = [0]
        for p in parts:
            self._bounds.append(self._bounds[-1] + p.n_docs)
        self.n_docs = self._bounds[-1]

    def _loc(self, i):
        import bisect
        if not 0 <= i < self.n_docs:
            raise IndexError(i)
        k = bisect.bisect_right(self._bounds, i) - 1
        return self._parts[k], i - self._bounds[k]

    def status(self, i): p, j = self._loc(i); return p.status(j)
    def json_bytes(self, i): p, j = self._loc(i); return p.json_bytes(j)
    def get_deep_value(self, i): p, j = self._loc(i); return p.get_deep_value(j)
    def oplog_vv(self, i): p, j = self._loc(i); return p.oplog_vv(j)
    def oplog_frontiers(self, i): p, j = self._loc(i); return p.oplog_frontiers(j)
    def export_updates(self, i, from_vv=None): p, j = self._loc(i); return p.export_updates(j, from_vv)

    def fetch_json(self):
        for p in self._parts:
            p.fetch_json()

    def fetch_exports(self):
        for p in self._parts:
            p.fetch_exports()

    def counters(self):
        out = {}
        for p in self._parts:
            for k, v in p.counters().items():
                out[k] = (out.get(k, 0) ^ v) if k == "state_hash" else out.get(k, 0) + v
        return out

    def timings(self):
        """per-phase device times summed over the sub-batches (they overlap on the device: the sum is not a wall time)"""
        out = {}
        for p in self._parts:
            for k, v in p.timings().items():
                out[k] = out.get(k, 0) + v
        return out

    def close(self):
        for p in self._parts:
            p.close()
        self._parts = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


SPLIT_MIN_BYTES = 256 << 20   # host batches at least this large are imported as overlapping sub-batches ...
SPLIT_MIN_PART_DOCS = 2048    # ... of at least this many documents each (one warp per document: fewer would idle the SMs)
SPLIT_PARTS = 2


def auto_split(blobs):
    """number of sub-batches import_batch uses for `blobs` when the caller does not say"""
    n = len(blobs)
    forced = os.environ.get("LORO_B200_SPLIT")     # measurement hook: force the number of sub-batches
    if forced:
        return max(1, int(forced))
    if n < 2 * SPLIT_MIN_PART_DOCS:
        return 1
    total = 0
    for b in blobs:
        total += len(b)
        if total >= SPLIT_MIN_BYTES:
            return max(1, min(SPLIT_PARTS, n // SPLIT_MIN_PART_DOCS))
    return 1


def _import_one(L, blobs, device, flags, doc_ids):
    arr, keep = _blob_array(blobs, doc_ids)
    opt = _Options(device=device, flags=flags)
    h = ctypes.c_void_p()
    _check(L, L.lb_import_batch(arr, len(blobs), ctypes.byref(opt), ctypes.byref(h)), "lb_import_batch")
    return Batch(L, h.value)


def import_batch(blobs, device=0, flags=0, lib_path=None, doc_ids=None, split=None):
    """LoroDoc::import for a batch: one fresh document per blob (bytes-like), host buffers in.
    `doc_ids` (one int per blob) groups blobs into documents the way LoroDoc::import_batch takes several updates:
    blobs with the same id form one document; documents are numbered in order of first appearance.
    `split`: number of sub-batches (None = auto_split(blobs) for batches without doc_ids: up to SPLIT_PARTS once the
    batch holds SPLIT_MIN_BYTES, never fewer than SPLIT_MIN_PART_DOCS documents each): sub-batches are imported two
    at a time so that the host<->device transfers of one overlap the kernels of the other; the result is a MultiBatch."""
    L = load_library(lib_path)
    n = len(blobs)
    if split is None:
        split = auto_split(blobs) if doc_ids is None else 1
    if split > 1 and doc_ids is None and n >= 2 * split:
        import threading
        step = (n + split - 1) // split
        ranges = [(a, min(n, a + step)) for a in range(0, n, step)]
        parts = [None] * len(ranges)
        errs = []
        nxt = [0]
        lock = threading.Lock()

        def work():
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= len(ranges) or errs:
                    return
                a, b = ranges[k]
                try:
                    parts[k] = _import_one(L, blobs[a:b], device, flags, None)
                    if flags & LB_FLAG_EXPORT:
                        parts[k].fetch_exports()    # this sub-batch's blobs come home while the next one computes
                except Exception as e:   # noqa: BLE001 -- re-raised below
                    errs.append(e)

        ts = [threading.Thread(target=work) for _ in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            for p in parts:
                if p is not None:
                    p.close()
            raise errs[0]
        return MultiBatch(parts)
    n = len(blobs)
    arr, keep = _blob_array(blobs, doc_ids)
    opt = _Options(device=device, flags=flags)
    h = ctypes.c_void_p()
    _check(L, L.lb_import_batch(arr, n, ctypes.byref(opt), ctypes.byref(h)), "lb_import_batch")
    return Batch(L, h.value)


def _blob_array(blobs, doc_ids):
    n = len(blobs)
    arr = (_Blob * max(n, 1))()
    keep = []
    for i, b in enumerate(blobs):
        b = bytes(b)
        keep.append(b)
        arr[i].ptr = b
        arr[i].len = len(b)
        arr[i].doc_id = i if doc_ids is None else int(doc_ids[i])
    return arr, keep


class DocSet:
    """Persistent documents: LoroDoc::import / import_batch against documents that already hold history.  The documents
    live in device memory between calls (their change stores in wire form, include/loro_b200.h lb_docset_*); every
    import_() returns a Batch that answers for the documents it touched -- status of THIS import, state after it."""

    def __init__(self, device=0, lib_path=None):
        self._L = load_library(lib_path)
        self._device = device
        opt = _Options(device=device, flags=0)
        h = ctypes.c_void_p()
        _check(self._L, self._L.lb_docset_new(ctypes.byref(opt), ctypes.byref(h)), "lb_docset_new")
        self._h = h.value

    def import_(self, blobs, doc_ids, flags=0):
        """blobs[i] is imported into document doc_ids[i]; several blobs for one id = import_batch on that document.
        Documents of the returned Batch are numbered in order of first appearance of their id."""
        arr, keep = _blob_array(blobs, doc_ids)
        opt = _Options(device=self._device, flags=flags)
        h = ctypes.c_void_p()
        _check(self._L, self._L.lb_docset_import(self._h, arr, len(blobs), ctypes.byref(opt), ctypes.byref(h)), "lb_docset_import")
        return Batch(self._L, h.value)

    @property
    def n_docs(self):
        return self._L.lb_docset_doc_count(self._h)

    @property
    def stored_bytes(self):
        return self._L.lb_docset_stored_bytes(self._h)

    def close(self):
        if self._h:
            self._L.lb_docset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def import_batch_device(d_bytes_ptr, offsets, lens, device=0, flags=0, lib_path=None, keep=None):
    """Same with blobs already resident in HBM: `d_bytes_ptr` is a device pointer (int); blob i occupies
    [offsets[i], offsets[i] + lens[i]) with every offset a multiple of 16."""
    L = load_library(lib_path)
    n = len(offsets)
    assert len(lens) == n
    if hasattr(offsets, "ctypes"):  # numpy fast path
        import numpy as np
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        l_ = np.ascontiguousarray(lens, dtype=np.uint32)
        offs = o.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        ls = l_.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        keep = (keep, o, l_)
    else:
        offs = (ctypes.c_uint64 * max(n, 1))(*[int(x) for x in offsets])
        ls = (ctypes.c_uint32 * max(n, 1))(*[int(x) for x in lens])
    opt = _Options(device=device, flags=flags)
    h = ctypes.c_void_p()
    _check(L, L.lb_import_batch_device(ctypes.c_void_p(d_bytes_ptr), offs, ls, n, ctypes.byref(opt), ctypes.byref(h)),
           "lb_import_batch_device")
    return Batch(L, h.value, keep=keep)


def device_trim(device=0, lib_path=None):
    """Release the device blocks the engine keeps for the next batch (lb_device_trim)."""
    L = load_library(lib_path)
    _check(L, L.lb_device_trim(int(device)), "lb_device_trim")


def numa_bind(device=0, lib_path=None):
    """Pin this process (its current thread and the threads created from it) to the CPUs of the NUMA node of `device`.
    Returns True when the binding was applied."""
    L = load_library(lib_path)
    L.lb_numa_bind.argtypes = [ctypes.c_int]
    return L.lb_numa_bind(device) == 0


def pack_blobs(blobs):
    """Concatenate blobs at 16-byte aligned starts -> (bytes, offsets, lens) for import_batch_device."""
    import numpy as np
    lens = np.fromiter((len(b) for b in blobs), dtype=np.uint32, count=len(blobs))
    padded = (lens.astype(np.uint64) + 15) & ~np.uint64(15)
    offs = np.zeros(len(blobs), dtype=np.uint64)
    if len(blobs):
        offs[1:] = np.cumsum(padded)[:-1]
    total = int(padded.sum())
    buf = np.zeros(total + 64, dtype=np.uint8)
    for b, o in zip(blobs, offs):
        buf[int(o):int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf, offs, lens
