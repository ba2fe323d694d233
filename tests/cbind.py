"""ctypes bindings for the CHECKERS used by the tests (never by the product):

  * OracleLib -- oracle/libndp_oracle.so, the plain-C restatement (oracle/ndp_oracle.c)
  * RefLib    -- oracle/_ref/libndp_ref*.so, the unmodified reference TU behind a C ABI
                 (oracle/ref_harness.cpp); only exists where /root/reference was available
                 at build time.

Both expose the same Python surface so a test can run one against the other.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_i32p = C.POINTER(C.c_int32)
_szp = C.POINTER(C.c_size_t)
_ip = C.POINTER(C.c_int)


def _as_c3(clauses):
    a = np.ascontiguousarray(np.asarray(clauses, dtype=np.int32).reshape(-1, 3))
    return a, a.ctypes.data_as(_i32p), a.shape[0]


class _Common:
    """Shared surface: prefix-dependent symbol lookup."""

    prefix = ""

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def _bind_common(self):
        self._parse = self._fn("parse_dimacs", C.c_size_t, [C.c_char_p, _i32p, C.c_size_t])
        self._choice = self._fn("choice", C.c_int, [_i32p, C.c_size_t])
        self._rs = self._fn("resolution_step", None,
                            [_i32p, C.c_size_t, C.c_int, C.c_int, _i32p, _szp, _i32p, _szp, _ip, _ip])
        self._fsize = self._fn("frontier_size", C.c_size_t, [C.c_void_p])
        self._fncl = self._fn("frontier_nclauses", C.c_size_t, [C.c_void_p, C.c_size_t])
        self._fnch = self._fn("frontier_nchoices", C.c_size_t, [C.c_void_p, C.c_size_t])
        self._fcopy = self._fn("frontier_copy", None, [C.c_void_p, C.c_size_t, _i32p, _i32p])
        self._ffree = self._fn("frontier_free", None, [C.c_void_p])

    # -- parser / primitives
    def parse(self, text):
        b = text.encode() if isinstance(text, str) else text
        n = self._parse(b, None, 0)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self._parse(b, out.ctypes.data_as(_i32p), n)
        return out

    def choice(self, clauses):
        a, p, n = _as_c3(clauses)
        return int(self._choice(p, n))

    def resolution_step(self, clauses, i, with_conflict):
        a, p, n = _as_c3(clauses)
        la = np.zeros((max(n, 1), 3), dtype=np.int32)
        ra = np.zeros((max(n, 1), 3), dtype=np.int32)
        nla, nra = C.c_size_t(0), C.c_size_t(0)
        cla, cra = C.c_int(0), C.c_int(0)
        self._rs(p, n, int(i), int(with_conflict), la.ctypes.data_as(_i32p), C.byref(nla),
                 ra.ctypes.data_as(_i32p), C.byref(nra), C.byref(cla), C.byref(cra))
        return la[:nla.value].copy(), ra[:nra.value].copy(), cla.value, cra.value

    def frontier_list(self, handle):
        out = []
        for k in range(self._fsize(handle)):
            ncl, nch = self._fncl(handle, k), self._fnch(handle, k)
            cl = np.zeros((ncl, 3), dtype=np.int32)
            ch = np.zeros(nch, dtype=np.int32)
            self._fcopy(handle, k, cl.ctypes.data_as(_i32p), ch.ctypes.data_as(_i32p))
            out.append((cl, ch))
        return out

    def frontier_free(self, handle):
        self._ffree(handle)


class OracleLib(_Common):
    prefix = "orc_"

    def __init__(self, build=True):
        path = os.path.join(ORACLE_DIR, "libndp_oracle.so")
        if build and not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
        self.lib = C.CDLL(path)
        self._bind_common()
        self._bfs = self._fn("bfs", C.c_void_p, [_i32p, C.c_size_t, C.c_int, C.c_int, C.c_int, _ip, _ip, _szp,
                                                  C.POINTER(C.c_uint64)])
        self._dfs = self._fn("dfs", C.c_int, [_i32p, C.c_size_t, C.c_int, _i32p, C.c_size_t, _szp,
                                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])

    def bfs_handle(self, clauses, max_iterations, override, max_queues, initial_queue_size=0):
        a, p, n = _as_c3(clauses)
        it, tc = C.c_int(0), C.c_int(0)
        iq = C.c_size_t(initial_queue_size)
        ev = C.c_uint64(0)
        h = self._bfs(p, n, int(max_iterations), int(bool(override)), int(max_queues), C.byref(it), C.byref(tc),
                      C.byref(iq), C.byref(ev))
        return h, dict(iterations=it.value, task_count=tc.value, initial_queue_size=iq.value,
                       clause_evals=ev.value)

    def bfs(self, clauses, max_iterations, override, max_queues, materialise=True):
        h, info = self.bfs_handle(clauses, max_iterations, override, max_queues)
        info["size"] = self._fsize(h)
        info["frontier"] = self.frontier_list(h) if materialise else None
        self._ffree(h)
        return info

    def dfs(self, clauses, first=True, cap=1 << 16):
        a, p, n = _as_c3(clauses)
        model = np.zeros(cap, dtype=np.int32)
        ml = C.c_size_t(0)
        nodes, ev = C.c_uint64(0), C.c_uint64(0)
        nm = self._dfs(p, n, int(first), model.ctypes.data_as(_i32p), cap, C.byref(ml), C.byref(nodes), C.byref(ev))
        return dict(models=nm, model=model[:ml.value].copy(), nodes=nodes.value, clause_evals=ev.value)


class RefLib(_Common):
    prefix = "ref_"

    @staticmethod
    def available(prof=False):
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libndp_ref_prof.so" if prof else "libndp_ref.so"))

    def __init__(self, prof=False):
        path = os.path.join(ORACLE_DIR, "_ref", "libndp_ref_prof.so" if prof else "libndp_ref.so")
        self.lib = C.CDLL(path)
        self._bind_common()
        self._bfs = self._fn("bfs", C.c_void_p, [_i32p, C.c_size_t, C.c_int, C.c_int, C.c_int, _ip, _ip, _szp,
                                                  C.POINTER(C.c_double)])
        self._dfs = self._fn("dfs", C.c_int, [_i32p, C.c_size_t, C.c_int, _i32p, C.c_size_t, _szp,
                                              C.POINTER(C.c_long)])
        self._dfs_frontier = self._fn("dfs_frontier", C.c_double,
                                      [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                       C.POINTER(C.c_uint8), C.POINTER(C.c_long), _i32p, C.c_size_t, _szp,
                                       C.POINTER(C.c_long)])
        self._save = self._fn("save_bfs", C.c_size_t, [C.c_void_p] + [C.c_char_p] * 5 + [C.c_double, C.c_char_p,
                                                                                         C.c_size_t])
        self._read = self._fn("read_bfs", C.c_void_p, [C.c_char_p, C.POINTER(C.c_double)])

    def bfs_handle(self, clauses, max_iterations, override, max_queues, initial_queue_size=0):
        a, p, n = _as_c3(clauses)
        it, tc = C.c_int(0), C.c_int(0)
        iq = C.c_size_t(initial_queue_size)
        sec = C.c_double(0)
        h = self._bfs(p, n, int(max_iterations), int(bool(override)), int(max_queues), C.byref(it), C.byref(tc),
                      C.byref(iq), C.byref(sec))
        return h, dict(iterations=it.value, task_count=tc.value, initial_queue_size=iq.value, seconds=sec.value)

    def bfs(self, clauses, max_iterations, override, max_queues, materialise=True):
        h, info = self.bfs_handle(clauses, max_iterations, override, max_queues)
        info["size"] = self._fsize(h)
        info["frontier"] = self.frontier_list(h) if materialise else None
        self._ffree(h)
        return info

    def dfs(self, clauses, first=True, cap=1 << 16):
        a, p, n = _as_c3(clauses)
        model = np.zeros(cap, dtype=np.int32)
        ml = C.c_size_t(0)
        nodes = C.c_long(0)
        nm = self._dfs(p, n, int(first), model.ctypes.data_as(_i32p), cap, C.byref(ml), C.byref(nodes))
        return dict(models=nm, model=model[:ml.value].copy(), nodes=nodes.value)

    def dfs_frontier(self, handle, first_k, count, workers, early_exit, cap=1 << 16):
        size = self._fsize(handle)
        count = min(count, size - first_k)
        verdict = np.zeros(max(count, 1), dtype=np.uint8)
        winner = C.c_long(-1)
        model = np.zeros(cap, dtype=np.int32)
        ml = C.c_size_t(0)
        nodes = np.zeros(max(count, 1), dtype=np.int64)
        sec = self._dfs_frontier(handle, first_k, count, workers, int(early_exit),
                                 verdict.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(winner),
                                 model.ctypes.data_as(_i32p), cap, C.byref(ml),
                                 nodes.ctypes.data_as(C.POINTER(C.c_long)))
        return dict(seconds=sec, verdict=verdict[:count].copy(), winner=winner.value, model=model[:ml.value].copy(),
                    nodes=nodes[:count].copy())

    def last_busy_seconds(self, count):
        """Worker seconds each subproblem of the last dfs_frontier call took (0 for those not run)."""
        f = self._fn("last_busy_seconds", C.c_double, [C.POINTER(C.c_double), C.c_size_t])
        out = np.zeros(max(count, 1), dtype=np.float64)
        f(out.ctypes.data_as(C.POINTER(C.c_double)), count)
        return out[:count]

    def frontier_from_arrays(self, items):
        f = self._fn("frontier_from_arrays", C.c_void_p, [_i32p, _szp, _i32p, _szp, C.c_size_t])
        cl_off, ch_off = [0], [0]
        for cl, ch in items:
            cl_off.append(cl_off[-1] + len(cl))
            ch_off.append(ch_off[-1] + len(ch))
        cl_all = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1, 3) for c, _ in items])
                                      if items else np.zeros((0, 3), np.int32))
        ch_all = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1) for _, c in items])
                                      if items else np.zeros(0, np.int32))
        clo, cho = np.asarray(cl_off, dtype=np.uintp), np.asarray(ch_off, dtype=np.uintp)
        return f(cl_all.ctypes.data_as(_i32p), clo.ctypes.data_as(_szp), ch_all.ctypes.data_as(_i32p),
                 cho.ctypes.data_as(_szp), len(items))

    def save_bfs(self, handle, script, filename, pid, flag, outdir, bfs_time):
        buf = C.create_string_buffer(4096)
        self._save(handle, script.encode(), filename.encode(), pid.encode(), flag.encode(), outdir.encode(),
                   float(bfs_time), buf, 4096)
        return buf.value.decode()

    def read_bfs(self, path):
        t = C.c_double(0)
        h = self._read(path.encode(), C.byref(t))
        return h, t.value

    def _str(self, name, argtypes, *args):
        f = self._fn(name, C.c_size_t, argtypes + [C.c_char_p, C.c_size_t])
        buf = C.create_string_buffer(1 << 16)
        f(*args, buf, 1 << 16)
        return buf.value.decode()

    def create_problem_id(self, n, bits, world, utc):
        return self._str("create_problem_id", [C.c_char_p, C.c_int, C.c_int, C.c_char_p], n.encode(), bits, world,
                         utc.encode())

    def format_filename(self, script, file, pid, flag):
        return self._str("format_filename", [C.c_char_p] * 4, script.encode(), file.encode(), pid.encode(),
                         flag.encode())

    def create_bfs_filename(self, script, file, pid, flag):
        return self._str("create_bfs_filename", [C.c_char_p] * 4, script.encode(), file.encode(), pid.encode(),
                         flag.encode())

    def format_duration(self, s):
        return self._str("format_duration", [C.c_double], float(s))

    def format_percentage(self, a, b):
        return self._str("format_percentage", [C.c_double, C.c_double], float(a), float(b))

    def format_vector(self, v):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.int32))
        return self._str("format_vector", [_i32p, C.c_size_t], a.ctypes.data_as(_i32p), a.size)

    def calculate_max_iterations(self, world, clauses, nvars, bits):
        f = self._fn("calculate_max_iterations", C.c_int, [C.c_int] * 4)
        return f(world, clauses, nvars, bits)

    def scrape_header(self, text):
        f = self._fn("scrape_header", C.c_int, [C.c_char_p, _ip, _ip, _ip, C.c_char_p, C.c_size_t, _i32p, _szp,
                                                _i32p, _szp, C.c_size_t])
        nv, nc, nb = C.c_int(0), C.c_int(0), C.c_int(0)
        prod = C.create_string_buffer(4096)
        v1 = np.zeros(4096, dtype=np.int32)
        v2 = np.zeros(4096, dtype=np.int32)
        n1, n2 = C.c_size_t(0), C.c_size_t(0)
        rc = f(text.encode(), C.byref(nv), C.byref(nc), C.byref(nb), prod, 4096, v1.ctypes.data_as(_i32p),
               C.byref(n1), v2.ctypes.data_as(_i32p), C.byref(n2), 4096)
        return dict(rc=rc, num_vars=nv.value, num_clauses=nc.value, num_bits=nb.value, product=prod.value.decode(),
                    v1=v1[:n1.value].copy(), v2=v2[:n2.value].copy())

    def convert(self, model, v1, v2, input_number):
        f = self._fn("convert", C.c_int, [_i32p, C.c_size_t, _i32p, C.c_size_t, _i32p, C.c_size_t, C.c_char_p,
                                          C.c_char_p, C.c_char_p, C.c_size_t])
        m = np.ascontiguousarray(np.asarray(model, dtype=np.int32))
        a = np.ascontiguousarray(np.asarray(v1, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(v2, dtype=np.int32))
        d1 = C.create_string_buffer(4096)
        d2 = C.create_string_buffer(4096)
        ok = f(m.ctypes.data_as(_i32p), m.size, a.ctypes.data_as(_i32p), a.size, b.ctypes.data_as(_i32p), b.size,
               str(input_number).encode(), d1, d2, 4096)
        return int(d1.value), int(d2.value), bool(ok)

    def generate_output(self, solution_found, model, bfs_s, dfs_s, num_bits, num_vars, num_clauses, input_number,
                        v1, v2, script, filename, cli_flag, outdir, iterations, initial_queue_size, world_rank,
                        world_size, final_queue_size):
        f = self._fn("generate_output", C.c_size_t,
                     [C.c_int, _i32p, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p,
                      _i32p, C.c_size_t, _i32p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                      C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t])
        m = np.ascontiguousarray(np.asarray(model, dtype=np.int32))
        a = np.ascontiguousarray(np.asarray(v1, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(v2, dtype=np.int32))
        buf = C.create_string_buffer(1 << 20)
        f(int(solution_found), m.ctypes.data_as(_i32p), m.size, bfs_s, dfs_s, num_bits, num_vars, num_clauses,
          str(input_number).encode(), a.ctypes.data_as(_i32p), a.size, b.ctypes.data_as(_i32p), b.size,
          script.encode(), filename.encode(), cli_flag.encode(), outdir.encode(), iterations, initial_queue_size,
          world_rank, world_size, final_queue_size, buf, 1 << 20)
        return buf.value.decode()


class HostLib:
    """ctypes view of ndp-5_b200/lib/libndp_host.so (the C++17 host drop-in's test hooks)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "ndp-5_b200", "lib", "libndp_host.so"))

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, "ndph_" + name)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def _str(self, name, argtypes, *args):
        f = self._fn(name, C.c_size_t, argtypes + [C.c_char_p, C.c_size_t])
        buf = C.create_string_buffer(1 << 16)
        f(*args, buf, 1 << 16)
        return buf.value.decode()

    def dropped(self, text):
        """(clause lines dropped for their width, first such width) -- NDP-5_6_7.cpp:639."""
        f = self._fn("dimacs_dropped", C.c_size_t, [C.c_char_p, C.POINTER(C.c_size_t)])
        w = C.c_size_t(0)
        n = f(text.encode() if isinstance(text, str) else text, C.byref(w))
        return int(n), int(w.value)

    def parse(self, text):
        f = self._fn("parse_dimacs", C.c_size_t, [C.c_char_p, _i32p, C.c_size_t])
        b = text.encode() if isinstance(text, str) else text
        n = f(b, None, 0)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            f(b, out.ctypes.data_as(_i32p), n)
        return out

    def scrape_header(self, text):
        f = self._fn("scrape_header", C.c_int, [C.c_char_p, _ip, _ip, _ip, C.c_char_p, C.c_size_t, _i32p, _szp,
                                                _i32p, _szp, C.c_size_t])
        nv, nc, nb = C.c_int(0), C.c_int(0), C.c_int(0)
        prod = C.create_string_buffer(4096)
        v1 = np.zeros(4096, dtype=np.int32)
        v2 = np.zeros(4096, dtype=np.int32)
        n1, n2 = C.c_size_t(0), C.c_size_t(0)
        rc = f(text.encode(), C.byref(nv), C.byref(nc), C.byref(nb), prod, 4096, v1.ctypes.data_as(_i32p),
               C.byref(n1), v2.ctypes.data_as(_i32p), C.byref(n2), 4096)
        return dict(rc=rc, num_vars=nv.value, num_clauses=nc.value, num_bits=nb.value, product=prod.value.decode(),
                    v1=v1[:n1.value].copy(), v2=v2[:n2.value].copy())

    def calculate_max_iterations(self, world, clauses, nvars, bits):
        return self._fn("calculate_max_iterations", C.c_int, [C.c_int] * 4)(world, clauses, nvars, bits)

    def create_problem_id(self, n, bits, world, utc):
        return self._str("create_problem_id", [C.c_char_p, C.c_int, C.c_int, C.c_char_p], n.encode(), bits, world,
                         utc.encode())

    def format_filename(self, script, file, pid, flag):
        return self._str("format_filename", [C.c_char_p] * 4, script.encode(), file.encode(), pid.encode(),
                         flag.encode())

    def create_bfs_filename(self, script, file, pid, flag):
        return self._str("create_bfs_filename", [C.c_char_p] * 4, script.encode(), file.encode(), pid.encode(),
                         flag.encode())

    def format_duration(self, s):
        return self._str("format_duration", [C.c_double], float(s))

    def format_percentage(self, a, b):
        return self._str("format_percentage", [C.c_double, C.c_double], float(a), float(b))

    def format_vector(self, v):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.int32))
        return self._str("format_vector", [_i32p, C.c_size_t], a.ctypes.data_as(_i32p), a.size)

    def convert(self, model, v1, v2, input_number):
        f = self._fn("convert", C.c_int, [_i32p, C.c_size_t, _i32p, C.c_size_t, _i32p, C.c_size_t, C.c_char_p,
                                          C.c_char_p, C.c_char_p, C.c_size_t])
        m = np.ascontiguousarray(np.asarray(model, dtype=np.int32))
        a = np.ascontiguousarray(np.asarray(v1, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(v2, dtype=np.int32))
        d1 = C.create_string_buffer(4096)
        d2 = C.create_string_buffer(4096)
        ok = f(m.ctypes.data_as(_i32p), m.size, a.ctypes.data_as(_i32p), a.size, b.ctypes.data_as(_i32p), b.size,
               str(input_number).encode(), d1, d2, 4096)
        return int(d1.value), int(d2.value), bool(ok)

    def parse_cli(self, args):
        f = self._fn("parse_cli", C.c_int, [C.c_int, C.POINTER(C.c_char_p), _ip, _ip, C.c_char_p, C.c_size_t, _ip, _ip,
                                            _ip, C.c_char_p, C.c_size_t, _ip, _ip, C.c_char_p, C.c_size_t])
        argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
        save, resume, mq, depth, ovr, gpus, exh = (C.c_int(0) for _ in range(7))
        rf = C.create_string_buffer(4096)
        flag = C.create_string_buffer(256)
        err = C.create_string_buffer(4096)
        rc = f(len(args), argv, C.byref(save), C.byref(resume), rf, 4096, C.byref(mq), C.byref(depth), C.byref(ovr),
               flag, 256, C.byref(gpus), C.byref(exh), err, 4096)
        return dict(rc=rc, save=bool(save.value), resume=bool(resume.value), resume_file=rf.value.decode(),
                    max_queues=mq.value, depth=depth.value, override=bool(ovr.value), cli_flag=flag.value.decode(),
                    gpus=gpus.value, exhaustive=bool(exh.value), error=err.value.decode())

    def write_bfs_json(self, path, items, bfs_time):
        f = self._fn("write_bfs_json", C.c_int, [C.c_char_p, _i32p, _szp, _i32p, _szp, C.c_size_t, C.c_double])
        cl_off, ch_off = [0], [0]
        for cl, ch in items:
            cl_off.append(cl_off[-1] + len(cl))
            ch_off.append(ch_off[-1] + len(ch))
        cl_all = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1, 3) for c, _ in items])
                                      if items else np.zeros((0, 3), np.int32))
        ch_all = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32).reshape(-1) for _, c in items])
                                      if items else np.zeros(0, np.int32))
        clo, cho = np.asarray(cl_off, dtype=np.uintp), np.asarray(ch_off, dtype=np.uintp)
        return f(path.encode(), cl_all.ctypes.data_as(_i32p), clo.ctypes.data_as(_szp), ch_all.ctypes.data_as(_i32p),
                 cho.ctypes.data_as(_szp), len(items), float(bfs_time))

    def read_bfs_json(self, path):
        rd = self._fn("read_bfs_json", C.c_void_p, [C.c_char_p])
        h = rd(path.encode())
        if not h:
            return None
        cnt = self._fn("json_count", C.c_size_t, [C.c_void_p])(h)
        t = self._fn("json_time", C.c_double, [C.c_void_p])(h)
        ncl = self._fn("json_nclauses", C.c_size_t, [C.c_void_p, C.c_size_t])
        nch = self._fn("json_nchoices", C.c_size_t, [C.c_void_p, C.c_size_t])
        cp = self._fn("json_copy", None, [C.c_void_p, C.c_size_t, _i32p, _i32p])
        items = []
        for k in range(cnt):
            cl = np.zeros((ncl(h, k), 3), dtype=np.int32)
            ch = np.zeros(nch(h, k), dtype=np.int32)
            cp(h, k, cl.ctypes.data_as(_i32p), ch.ctypes.data_as(_i32p))
            items.append((cl, ch))
        self._fn("json_free", None, [C.c_void_p])(h)
        return items, t

    def report(self, solution_found, model, bfs_s, dfs_s, num_bits, num_vars, num_clauses, input_number, v1, v2,
               script, filename, cli_flag, iterations, initial_queue_size, world_rank, world_size, final_queue_size,
               utc):
        f = self._fn("report", C.c_size_t,
                     [C.c_int, _i32p, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, _i32p,
                      C.c_size_t, _i32p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_size_t, C.c_int,
                      C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p,
                      C.c_size_t])
        m = np.ascontiguousarray(np.asarray(model, dtype=np.int32))
        a = np.ascontiguousarray(np.asarray(v1, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(v2, dtype=np.int32))
        out = C.create_string_buffer(1 << 20)
        ft = C.create_string_buffer(1 << 20)
        fn = C.create_string_buffer(4096)
        f(int(solution_found), m.ctypes.data_as(_i32p), m.size, bfs_s, dfs_s, num_bits, num_vars, num_clauses,
          str(input_number).encode(), a.ctypes.data_as(_i32p), a.size, b.ctypes.data_as(_i32p), b.size,
          script.encode(), filename.encode(), cli_flag.encode(), iterations, initial_queue_size, world_rank,
          world_size, final_queue_size, utc.encode() if utc else None, out, 1 << 20, ft, 1 << 20, fn, 4096)
        return out.value.decode(), ft.value.decode(), fn.value.decode()
