"""ctypes binding of the product's C ABI (include/ndp_b200.h -> ndp-5_b200/lib/libndp_b200.so).

This is plumbing for tests/, bench.py and __graft_entry__.py: it adds no logic of its own and
never falls back to anything -- if the CUDA library is missing, import fails; if no B200 is
present, every compute call raises NdpError(NDP_E_CUDA).
"""
import ctypes as C
import os

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # ndp-5_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libndp_b200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "ndp_b200.h")

NDP_OK, NDP_E_INVALID, NDP_E_UNSUPPORTED, NDP_E_CUDA, NDP_E_NOMEM = range(5)
UNSAT, SAT, NOT_RUN = 0, 1, 2
SIZE_MAX = C.c_size_t(-1).value

_i32p = C.POINTER(C.c_int32)
_szp = C.POINTER(C.c_size_t)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


class NdpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ndp_b200 error %d: %s" % (code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("clause_evals", C.c_uint64), ("rebuilds", C.c_uint64),
                ("kernel_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    sig = {
        "ndp_set_device": (C.c_int, [C.c_int]),
        "ndp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
        "ndp_set_engine": (C.c_int, [C.c_int]),
        "ndp_set_split": (C.c_int, [C.c_longlong]),
        "ndp_last_error": (C.c_char_p, []),
        "ndp_version": (C.c_char_p, []),
        "ndp_bfs": (C.c_int, [_i32p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                              C.POINTER(C.c_int), C.POINTER(C.c_int), _szp]),
        "ndp_frontier_bfs_stats": (C.c_int, [C.c_void_p, _u64p, C.POINTER(C.c_double), _u64p]),
        "ndp_frontier_bfs_bytes": (C.c_int, [C.c_void_p, _u64p, _u64p]),
        "ndp_frontier_size": (C.c_size_t, [C.c_void_p]),
        "ndp_frontier_get": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(_i32p), _szp, C.POINTER(_i32p), _szp]),
        "ndp_frontier_dims": (C.c_int, [C.c_void_p, C.c_size_t, _szp, _szp]),
        "ndp_frontier_from_host": (C.c_int, [_i32p, _szp, _i32p, _szp, C.c_size_t, C.POINTER(C.c_void_p)]),
        "ndp_frontier_attach_root": (C.c_int, [C.c_void_p, _i32p, C.c_size_t]),
        "ndp_frontier_free": (None, [C.c_void_p]),
        "ndp_dfs_shard": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, _szp, _u8p, _szp, _i32p,
                                    C.c_size_t, _szp, _u64p, _u64p, C.POINTER(Stats)]),
        "ndp_dfs_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _u8p, _szp, _i32p, C.c_size_t, _szp,
                                    C.POINTER(Stats)]),
        "ndp_exit_word_create": (C.c_int, [C.POINTER(C.c_void_p)]),
        "ndp_exit_word_export": (C.c_int, [C.c_void_p, _u8p]),
        "ndp_exit_word_import": (C.c_int, [_u8p, C.POINTER(C.c_void_p)]),
        "ndp_exit_word_reset": (C.c_int, [C.c_void_p]),
        "ndp_exit_word_read": (C.c_int, [C.c_void_p, _szp]),
        "ndp_exit_word_free": (None, [C.c_void_p]),
        "ndp_dfs_shard_shared": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                           C.POINTER(C.c_void_p), C.c_int, _u8p, _szp, _i32p, C.c_size_t, _szp,
                                           _u64p, _u64p, C.POINTER(Stats)]),
        "ndp_dfs_pieces_begin": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                           C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), _u32p,
                                           C.POINTER(Stats)]),
        "ndp_dfs_pieces_fold": (C.c_int, [C.c_void_p, _u32p, _u8p, _u64p, _u64p, _szp, _i32p, C.c_size_t, _szp,
                                          C.POINTER(Stats)]),
        "ndp_dfs_pieces_end": (None, [C.c_void_p]),
        "ndp_dfs_progress": (C.c_int, [_szp, _szp, C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _check(rc):
    if rc != NDP_OK:
        raise NdpError(rc, lib.ndp_last_error().decode(errors="replace"))


def set_device(d):
    _check(lib.ndp_set_device(int(d)))


ENGINE_AUTO, ENGINE_SCAN, ENGINE_INDEX = 0, 1, 2


def set_engine(e):
    """0 auto, 1 scan, 2 index (or the names)."""
    e = {"auto": 0, "scan": 1, "index": 2}.get(e, e)
    _check(lib.ndp_set_engine(int(e)))


def set_split(pieces):
    """-1 automatic, 0 never split a shard into pieces, > 0 piece target (ndp_set_split)."""
    pieces = {"auto": -1, "off": 0}.get(pieces, pieces)
    _check(lib.ndp_set_split(int(pieces)))


def dfs_progress():
    """(units pulled, units in total, devices searching) of the DFS call(s) in flight -- callable from another thread."""
    a, b, c = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    _check(lib.ndp_dfs_progress(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def device_count():
    n = C.c_int(0)
    rc = lib.ndp_device_count(C.byref(n))
    return n.value if rc == NDP_OK else 0


EXIT_HANDLE_BYTES = 64


class ExitWord:
    """ndp_exit_word: the device-visible early-exit word of one process (create), or another process's word
    mapped here through its 64-byte CUDA IPC handle (ExitWord.from_handle)."""

    def __init__(self, handle=None):
        self.h = C.c_void_p(None)
        if handle is None:
            _check(lib.ndp_exit_word_create(C.byref(self.h)))
        else:
            buf = np.frombuffer(bytes(handle), dtype=np.uint8).copy()
            assert buf.size == EXIT_HANDLE_BYTES
            _check(lib.ndp_exit_word_import(buf.ctypes.data_as(_u8p), C.byref(self.h)))

    @classmethod
    def from_handle(cls, handle):
        return cls(handle)

    def export(self):
        buf = np.zeros(EXIT_HANDLE_BYTES, dtype=np.uint8)
        _check(lib.ndp_exit_word_export(self.h, buf.ctypes.data_as(_u8p)))
        return buf.tobytes()

    def reset(self):
        _check(lib.ndp_exit_word_reset(self.h))

    def read(self):
        v = C.c_size_t(0)
        _check(lib.ndp_exit_word_read(self.h, C.byref(v)))
        return None if v.value == SIZE_MAX else v.value

    def free(self):
        if self.h:
            lib.ndp_exit_word_free(self.h)
            self.h = C.c_void_p(None)


class Frontier:
    """Owner of an ndp_frontier handle."""

    def __init__(self, handle, iterations=0, task_count=0, initial_queue_size=0):
        self.h = C.c_void_p(handle)
        self.iterations = iterations
        self.task_count = task_count
        self.initial_queue_size = initial_queue_size

    def __len__(self):
        return int(lib.ndp_frontier_size(self.h))

    def dims(self, k):
        ncl, nch = C.c_size_t(0), C.c_size_t(0)
        _check(lib.ndp_frontier_dims(self.h, k, C.byref(ncl), C.byref(nch)))
        return ncl.value, nch.value

    def get(self, k):
        """(clauses [n,3] int32, choices [m] int32) of queue element k, copied out."""
        cp, chp = _i32p(), _i32p()
        ncl, nch = C.c_size_t(0), C.c_size_t(0)
        _check(lib.ndp_frontier_get(self.h, k, C.byref(cp), C.byref(ncl), C.byref(chp), C.byref(nch)))
        cl = np.ctypeslib.as_array(cp, shape=(ncl.value * 3,)).copy().reshape(-1, 3) if ncl.value else \
            np.zeros((0, 3), np.int32)
        ch = np.ctypeslib.as_array(chp, shape=(nch.value,)).copy() if nch.value else np.zeros(0, np.int32)
        return cl, ch

    def bfs_stats(self):
        ev, ms, ln = C.c_uint64(0), C.c_double(0), C.c_uint64(0)
        _check(lib.ndp_frontier_bfs_stats(self.h, C.byref(ev), C.byref(ms), C.byref(ln)))
        h2d, d2h = C.c_uint64(0), C.c_uint64(0)
        _check(lib.ndp_frontier_bfs_bytes(self.h, C.byref(h2d), C.byref(d2h)))
        return dict(clause_evals=ev.value, kernel_ms=ms.value, launches=ln.value, h2d_bytes=h2d.value,
                    d2h_bytes=d2h.value)

    def dfs_shard(self, first=0, stride=1, early_exit=False, exit_flag=None, model_cap=1 << 16, counters=True):
        n = len(self)
        verdict = np.full(max(n, 1), NOT_RUN, dtype=np.uint8)
        nodes = np.zeros(max(n, 1), dtype=np.uint64)
        evals = np.zeros(max(n, 1), dtype=np.uint64)
        winner = C.c_size_t(SIZE_MAX)
        flag = C.c_size_t(SIZE_MAX if exit_flag is None else exit_flag)
        model = np.zeros(model_cap, dtype=np.int32)
        mlen = C.c_size_t(0)
        st = Stats()
        _check(lib.ndp_dfs_shard(self.h, first, stride, int(early_exit), C.byref(flag),
                                 verdict.ctypes.data_as(_u8p), C.byref(winner), model.ctypes.data_as(_i32p),
                                 model_cap, C.byref(mlen),
                                 nodes.ctypes.data_as(_u64p) if counters else None,
                                 evals.ctypes.data_as(_u64p) if counters else None, C.byref(st)))
        return dict(verdict=verdict[:n], winner=None if winner.value == SIZE_MAX else winner.value,
                    model=model[:mlen.value].copy(), nodes=nodes[:n], evals=evals[:n], stats=st.as_dict(),
                    exit_flag=None if flag.value == SIZE_MAX else flag.value)

    def dfs_shard_shared(self, own, peers, first=0, stride=1, early_exit=True, model_cap=1 << 16, counters=True):
        """ndp_dfs_shard_shared: this rank's shard with device-visible early-exit words (own + the other ranks')."""
        n = len(self)
        verdict = np.full(max(n, 1), NOT_RUN, dtype=np.uint8)
        nodes = np.zeros(max(n, 1), dtype=np.uint64)
        evals = np.zeros(max(n, 1), dtype=np.uint64)
        winner = C.c_size_t(SIZE_MAX)
        model = np.zeros(model_cap, dtype=np.int32)
        mlen = C.c_size_t(0)
        st = Stats()
        arr = (C.c_void_p * max(len(peers), 1))(*[p.h for p in peers])
        _check(lib.ndp_dfs_shard_shared(self.h, first, stride, int(early_exit), own.h, arr, len(peers),
                                        verdict.ctypes.data_as(_u8p), C.byref(winner), model.ctypes.data_as(_i32p),
                                        model_cap, C.byref(mlen),
                                        nodes.ctypes.data_as(_u64p) if counters else None,
                                        evals.ctypes.data_as(_u64p) if counters else None, C.byref(st)))
        return dict(verdict=verdict[:n], winner=None if winner.value == SIZE_MAX else winner.value,
                    model=model[:mlen.value].copy(), nodes=nodes[:n], evals=evals[:n], stats=st.as_dict())

    def dfs_pieces(self, rank=0, world=1, early_exit=False, own=None, peers=()):
        """ndp_dfs_pieces_begin: cut the frontier into DFS-order pieces (the same cut on every rank) and search
        pieces rank, rank + world, ...  Returns a PieceRun; combine with shard.fold_piece_runs (or call
        .fold(first_model) yourself after a MIN all-reduce of .first_model)."""
        n = len(self)
        first_model = np.full(max(n, 1), 0xffffffff, dtype=np.uint32)
        h = C.c_void_p(None)
        st = Stats()
        arr = (C.c_void_p * max(len(peers), 1))(*[p.h for p in peers])
        _check(lib.ndp_dfs_pieces_begin(self.h, rank, world, int(early_exit), own.h if own is not None else None, arr,
                                        len(peers), C.byref(h), first_model.ctypes.data_as(_u32p), C.byref(st)))
        return PieceRun(h, n, first_model[:n], st.as_dict())

    def dfs_batch(self, n_gpus=1, early_exit=False, model_cap=1 << 16):
        n = len(self)
        verdict = np.full(max(n, 1), NOT_RUN, dtype=np.uint8)
        winner = C.c_size_t(SIZE_MAX)
        model = np.zeros(model_cap, dtype=np.int32)
        mlen = C.c_size_t(0)
        st = Stats()
        _check(lib.ndp_dfs_batch(self.h, int(n_gpus), int(early_exit), verdict.ctypes.data_as(_u8p),
                                 C.byref(winner), model.ctypes.data_as(_i32p), model_cap, C.byref(mlen),
                                 C.byref(st)))
        return dict(verdict=verdict[:n], winner=None if winner.value == SIZE_MAX else winner.value,
                    model=model[:mlen.value].copy(), stats=st.as_dict())

    def attach_root(self, clauses):
        """ndp_frontier_attach_root: True if the resumed frontier descends from this clause array (index engine
        from here on), False if the library refused (NDP_E_INVALID) and left the frontier as it was."""
        a = np.ascontiguousarray(np.asarray(clauses, dtype=np.int32).reshape(-1, 3))
        rc = lib.ndp_frontier_attach_root(self.h, a.ctypes.data_as(_i32p), a.shape[0])
        if rc == NDP_E_INVALID:
            return False
        _check(rc)
        return True

    def free(self):
        if self.h:
            lib.ndp_frontier_free(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PieceRun:
    """Owner of an ndp_piece_run handle (this rank's share of a piece-level search)."""

    def __init__(self, handle, n, first_model, stats):
        self.h, self.n, self.first_model, self.stats = handle, n, first_model, stats

    def fold(self, first_model, model_cap=1 << 16):
        """ndp_dfs_pieces_fold: this rank's partial per-subproblem results given the job-wide first_model."""
        n = self.n
        fm = np.ascontiguousarray(np.asarray(first_model, dtype=np.uint32))
        if fm.size < max(n, 1):
            fm = np.concatenate([fm, np.full(max(n, 1) - fm.size, 0xffffffff, dtype=np.uint32)])
        verdict = np.zeros(max(n, 1), dtype=np.uint8)
        nodes = np.zeros(max(n, 1), dtype=np.uint64)
        evals = np.zeros(max(n, 1), dtype=np.uint64)
        winner = C.c_size_t(SIZE_MAX)
        model = np.zeros(model_cap, dtype=np.int32)
        mlen = C.c_size_t(0)
        st = Stats()
        _check(lib.ndp_dfs_pieces_fold(self.h, fm.ctypes.data_as(_u32p), verdict.ctypes.data_as(_u8p),
                                       nodes.ctypes.data_as(_u64p), evals.ctypes.data_as(_u64p), C.byref(winner),
                                       model.ctypes.data_as(_i32p), model_cap, C.byref(mlen), C.byref(st)))
        return dict(verdict=verdict[:n], nodes=nodes[:n], evals=evals[:n],
                    winner=None if winner.value == SIZE_MAX else winner.value, model=model[:mlen.value].copy(),
                    stats=dict(self.stats, nodes=st.nodes, clause_evals=st.clause_evals))

    def free(self):
        if self.h:
            lib.ndp_dfs_pieces_end(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bfs(clauses, max_iterations, override_max_iterations, max_queues, initial_queue_size=0):
    """ndp_bfs: same arguments as Satisfy_iterative_BFS (NDP-5_6_7.cpp:881-882)."""
    a = np.ascontiguousarray(np.asarray(clauses, dtype=np.int32).reshape(-1, 3))
    h = C.c_void_p(None)
    it, tc = C.c_int(0), C.c_int(0)
    iq = C.c_size_t(initial_queue_size)
    _check(lib.ndp_bfs(a.ctypes.data_as(_i32p), a.shape[0], int(max_iterations), int(bool(override_max_iterations)),
                       int(max_queues), C.byref(h), C.byref(it), C.byref(tc), C.byref(iq)))
    return Frontier(h.value, it.value, tc.value, iq.value)


def frontier_from_host(items):
    """items: iterable of (clauses [n,3], choices [m]) -- the -r resume path."""
    items = list(items)
    cl_off, ch_off = [0], [0]
    cls, chs = [], []
    for cl, ch in items:
        cl = np.asarray(cl, dtype=np.int32).reshape(-1, 3)
        ch = np.asarray(ch, dtype=np.int32).reshape(-1)
        cls.append(cl)
        chs.append(ch)
        cl_off.append(cl_off[-1] + cl.shape[0])
        ch_off.append(ch_off[-1] + ch.shape[0])
    cl_all = np.ascontiguousarray(np.concatenate(cls) if cls else np.zeros((0, 3), np.int32))
    ch_all = np.ascontiguousarray(np.concatenate(chs) if chs else np.zeros(0, np.int32))
    clo = np.asarray(cl_off, dtype=np.uintp)
    cho = np.asarray(ch_off, dtype=np.uintp)
    h = C.c_void_p(None)
    _check(lib.ndp_frontier_from_host(cl_all.ctypes.data_as(_i32p), clo.ctypes.data_as(_szp),
                                      ch_all.ctypes.data_as(_i32p), cho.ctypes.data_as(_szp), len(items),
                                      C.byref(h)))
    return Frontier(h.value)
