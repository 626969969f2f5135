"""TEST INFRASTRUCTURE ONLY -- ctypes driver for ``oracle/liboracle.so`` (our plain-C restatement of the hot paths)."""
import ctypes
import json
import os
from ctypes import POINTER, Structure, byref, c_float, c_int, c_uint32, c_uint64

import numpy as np
import scipy.sparse as smat

from . import RESTATEMENT_LIB, build

_lib = None


class _Csc(Structure):
    _fields_ = [("rows", c_uint32), ("cols", c_uint32), ("col_ptr", POINTER(c_uint64)), ("row_idx", POINTER(c_uint32)),
                ("val", POINTER(c_float))]


class _Query(Structure):
    _fields_ = [("rows", c_uint32), ("cols", c_uint32), ("row_ptr", POINTER(c_uint64)), ("col_idx", POINTER(c_uint32)),
                ("val", POINTER(c_float))]


class _Result(Structure):
    _fields_ = [("indptr", POINTER(c_uint64)), ("indices", POINTER(c_uint32)), ("data", POINTER(c_float)),
                ("nnz", c_uint64), ("rows", c_uint32), ("cols", c_uint32)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(RESTATEMENT_LIB):
            build()
        L = ctypes.CDLL(RESTATEMENT_LIB)
        L.xlo_predict.restype = c_int
        L.xlo_predict.argtypes = [c_int, POINTER(_Csc), POINTER(_Csc), POINTER(c_float), POINTER(c_int), POINTER(c_int),
                                  POINTER(c_uint32), POINTER(_Query), POINTER(_Result)]
        L.xlo_predict_from.restype = c_int
        L.xlo_predict_from.argtypes = [c_int, POINTER(_Csc), POINTER(_Csc), POINTER(c_float), POINTER(c_int), POINTER(c_int),
                                       POINTER(c_uint32), POINTER(_Query), POINTER(_Query), POINTER(_Result)]
        L.xlo_predict_selected.restype = c_int
        L.xlo_predict_selected.argtypes = [c_int, POINTER(_Csc), POINTER(_Csc), POINTER(c_float), POINTER(c_int), POINTER(c_int),
                                           POINTER(_Query), POINTER(_Query), POINTER(_Result)]
        L.xlo_free_result.restype = None
        L.xlo_free_result.argtypes = [POINTER(_Result)]
        _lib = L
    return _lib


def parse_post_processor(name):
    """(kind, p) with the name rules of PostProcessor<T>::get (pecos/core/xmc/inference.hpp:192-240)."""
    if name == "noop":
        return 0, 0
    if name == "sigmoid":
        return 1, 0
    if name == "log-sigmoid":
        return 2, 0
    if name.startswith("log-l") and name.endswith("-hinge"):
        return 4, int(name[len("log-l"):-len("-hinge")] or 0)
    if name.startswith("l") and name.endswith("-hinge"):
        return 3, int(name[1:-len("-hinge")] or 0)
    return 0, 0


def _query_struct(X, keep):
    q = _Query()
    if isinstance(X, smat.csr_matrix):
        ip = np.ascontiguousarray(X.indptr, dtype=np.uint64)
        ix = np.ascontiguousarray(X.indices, dtype=np.uint32)
        dv = np.ascontiguousarray(X.data, dtype=np.float32)
        keep.extend([ip, ix, dv])
        q.rows, q.cols = X.shape
        q.row_ptr = ip.ctypes.data_as(POINTER(c_uint64))
        q.col_idx = ix.ctypes.data_as(POINTER(c_uint32))
        q.val = dv.ctypes.data_as(POINTER(c_float))
    else:
        Xd = np.ascontiguousarray(X, dtype=np.float32)
        keep.append(Xd)
        q.rows, q.cols = Xd.shape
        q.row_ptr = None
        q.col_idx = None
        q.val = Xd.ctypes.data_as(POINTER(c_float))
    return q


def _take_result(L, res):
    n = int(res.nnz)
    indptr = np.ctypeslib.as_array(res.indptr, shape=(res.rows + 1,)).astype(np.int64)
    indices = np.ctypeslib.as_array(res.indices, shape=(max(n, 1),))[:n].astype(np.int64)
    data = np.ctypeslib.as_array(res.data, shape=(max(n, 1),))[:n].copy()
    out = smat.csr_matrix((data, indices, indptr), shape=(res.rows, res.cols))
    L.xlo_free_result(byref(res))
    return out


def single_layer_predict(X, csr_codes, W, C, post_processor, only_topk, bias):
    """Restatement of c_xlinear_single_layer_predict_{csr,drm}_f32 (pecos/core/libpecos.cpp:201-235): one layer of the
    python prediction chain; `csr_codes` (previous layer's prediction, entries in stored order) or None."""
    L = lib()
    keep = []
    W = smat.csc_matrix(W, dtype=np.float32)
    W.sort_indices()
    C = smat.csc_matrix(C, dtype=np.float32)
    Ws = (_Csc * 1)(OracleXLinear._csc_struct(W, keep))
    Cs = (_Csc * 1)(OracleXLinear._csc_struct(C, keep))
    kind, p = parse_post_processor(post_processor)
    q = _query_struct(X, keep)
    codes = None
    if csr_codes is not None:
        codes = byref(_query_struct(smat.csr_matrix(csr_codes, dtype=np.float32), keep))
    res = _Result()
    rc = L.xlo_predict_from(1, Ws, Cs, (c_float * 1)(bias), (c_int * 1)(kind), (c_int * 1)(p), (c_uint32 * 1)(only_topk),
                            byref(q), codes, byref(res))
    assert rc == 0
    return _take_result(L, res)


class OracleXLinear(object):
    """Loads an npz model folder (the ``ranker/`` directory) with scipy and predicts with the C restatement."""

    def __init__(self, ranker_folder):
        param = json.load(open(os.path.join(ranker_folder, "param.json")))
        self.depth = int(param["depth"])
        self.layers = []
        for d in range(self.depth):
            sub = os.path.join(ranker_folder, f"{d}.model")
            p = json.load(open(os.path.join(sub, "param.json")))
            W = smat.load_npz(os.path.join(sub, "W.npz")).tocsc().astype(np.float32)
            W.sort_indices()
            c_path = os.path.join(sub, "C.npz")
            if d == 0 and not os.path.exists(c_path):
                C = smat.csc_matrix(np.ones((W.shape[1], 1), dtype=np.float32))
            else:
                C = smat.load_npz(c_path).tocsc().astype(np.float32)
            self.layers.append(dict(W=W, C=C, bias=float(p["bias"]), only_topk=int(p["pred_kwargs"]["only_topk"]),
                                    post_processor=str(p["pred_kwargs"]["post_processor"])))
        last = self.layers[-1]
        self.nr_features = last["W"].shape[0] - (1 if last["bias"] > 0 else 0)

    @staticmethod
    def _csc_struct(M, keep):
        ip = np.ascontiguousarray(M.indptr, dtype=np.uint64)
        ix = np.ascontiguousarray(M.indices, dtype=np.uint32)
        dv = np.ascontiguousarray(M.data, dtype=np.float32)
        keep.extend([ip, ix, dv])
        s = _Csc()
        s.rows, s.cols = M.shape
        s.col_ptr = ip.ctypes.data_as(POINTER(c_uint64))
        s.row_idx = ix.ctypes.data_as(POINTER(c_uint32))
        s.val = dv.ctypes.data_as(POINTER(c_float))
        return s

    def predict_on_selected_outputs(self, X, selected_outputs_csr, post_processor=None):
        """Restatement of c_xlinear_predict_on_selected_outputs_* (pecos/core/libpecos.cpp:179-198)."""
        L = lib()
        keep = []
        D = self.depth
        Ws = (_Csc * D)(*[self._csc_struct(l["W"], keep) for l in self.layers])
        Cs = (_Csc * D)(*[self._csc_struct(l["C"], keep) for l in self.layers])
        bias = (c_float * D)(*[l["bias"] for l in self.layers])
        kinds, ps = [], []
        for l in self.layers:
            kind, p = parse_post_processor(post_processor if post_processor else l["post_processor"])
            kinds.append(kind)
            ps.append(p)
        q = _query_struct(X, keep)
        sel = _query_struct(smat.csr_matrix(selected_outputs_csr, dtype=np.float32), keep)
        res = _Result()
        rc = L.xlo_predict_selected(D, Ws, Cs, bias, (c_int * D)(*kinds), (c_int * D)(*ps), byref(q), byref(sel), byref(res))
        assert rc == 0
        return _take_result(L, res)

    def predict(self, X, beam_size=0, post_processor=None, only_topk=0):
        L = lib()
        keep = []
        D = self.depth
        Ws = (_Csc * D)(*[self._csc_struct(l["W"], keep) for l in self.layers])
        Cs = (_Csc * D)(*[self._csc_struct(l["C"], keep) for l in self.layers])
        bias = (c_float * D)(*[l["bias"] for l in self.layers])
        kinds, ps, ks = [], [], []
        for d, l in enumerate(self.layers):
            kind, p = parse_post_processor(post_processor if post_processor else l["post_processor"])
            kinds.append(kind)
            ps.append(p)
            local = only_topk if d == D - 1 else beam_size   # inference.hpp:2471
            ks.append(local if local else l["only_topk"])    # inference.hpp:2055
        q = _query_struct(X, keep)
        res = _Result()
        rc = L.xlo_predict(D, Ws, Cs, bias, (c_int * D)(*kinds), (c_int * D)(*ps), (c_uint32 * D)(*ks), byref(q), byref(res))
        assert rc == 0
        return _take_result(L, res)


class _HnswIndex(Structure):
    _fields_ = [("num_node", c_uint32), ("maxM", c_uint32), ("maxM0", c_uint32), ("efC", c_uint32),
                ("max_level", c_uint32), ("init_node", c_uint32),
                ("feat_dim", c_uint32), ("l0_max_degree", c_uint32), ("l0_node_mem_size", c_uint32),
                ("l0_buffer", ctypes.c_void_p),
                ("l1_max_level", c_uint32), ("l1_max_degree", c_uint32), ("l1_node_mem_size", c_uint32),
                ("l1_level_mem_size", c_uint32), ("l1_buffer", ctypes.c_void_p),
                ("metric", c_int), ("isa", c_int), ("sparse", c_int), ("mem_start", ctypes.c_void_p)]


def _hnsw_lib():
    L = lib()
    if not hasattr(L, "_hnsw_ready"):
        L.hno_search.restype = c_int
        L.hno_search.argtypes = [POINTER(_HnswIndex), POINTER(c_float), c_uint32, c_uint32, c_uint32, POINTER(c_uint32),
                                 POINTER(c_float), POINTER(c_uint64)]
        L.hno_search_csr.restype = c_int
        L.hno_search_csr.argtypes = [POINTER(_HnswIndex), POINTER(c_uint64), POINTER(c_uint32), POINTER(c_float), c_uint32, c_uint32,
                                     c_uint32, POINTER(c_uint32), POINTER(c_float), POINTER(c_uint64)]
        L.hno_sparse_distance.restype = c_float
        L.hno_sparse_distance.argtypes = [ctypes.c_size_t, POINTER(c_float), POINTER(c_uint32), ctypes.c_size_t, POINTER(c_float),
                                          POINTER(c_uint32), c_int, c_int]
        L.hno_distance.restype = c_float
        L.hno_distance.argtypes = [POINTER(c_float), POINTER(c_float), c_uint32, c_int, c_int]
        L._hnsw_ready = True
    return L


def read_mmap_store(path):
    """Blocks of a PECOS ``*.mmap_store`` file as a list of uint8 arrays (pecos/core/utils/mmap_util.hpp:54-184)."""
    raw = np.fromfile(path, dtype=np.uint8)
    sig = raw[-16:]
    assert bytes(sig[:6]) == b"\x93PECOS" and sig[6] == ord("<") and sig[7] == 1, "not a PECOS mmap_store"
    meta_off = int(sig[8:16].view(np.uint64)[0])
    n_blocks = int(raw[meta_off:meta_off + 8].view(np.uint64)[0])
    info = raw[meta_off + 8: meta_off + 8 + 16 * n_blocks].view(np.uint64).reshape(n_blocks, 2)
    return [raw[int(o): int(o) + int(s)] for o, s in info]


def host_isa():
    """Which SIMD clone the reference's ifunc resolver picks on this CPU (distance_impl/x86.hpp:37-42, :84-85, :121-122)."""
    try:
        flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0].split()
    except Exception:
        return 0
    if "avx512f" in flags:
        return 0
    if "avx" in flags:
        return 1
    if "sse" in flags:
        return 2
    return 3


class OracleHNSW(object):
    """Parses ``<model>/c_model/index.mmap_store`` with numpy and searches with the C restatement."""

    def __init__(self, model_dir, isa=0):
        param = json.load(open(os.path.join(model_dir, "param.json")))
        self.sparse = param["data_type"] == "csr"
        self.metric = {"ip": 0, "l2": 1}[param["metric_type"]]
        blocks = read_mmap_store(os.path.join(model_dir, "c_model", "index.mmap_store"))
        u32 = lambda b: int(b.view(np.uint32)[0])  # noqa: E731
        it = iter(blocks)
        self.num_node, self.maxM, self.maxM0, self.efC, self.max_level, self.init_node = [u32(next(it)) for _ in range(6)]
        l0_num, self.feat_dim, self.l0_max_degree, self.l0_node_mem = [u32(next(it)) for _ in range(4)]
        next(it)            # mem_start_of_node: size,
        self.mem_start = np.ascontiguousarray(next(it)).view(np.uint64)  # data (num_node + 1 byte offsets)
        next(it)            # buffer size
        self.l0 = np.ascontiguousarray(next(it))
        l1_num, self.l1_max_level, self.l1_max_degree, self.l1_node_mem, self.l1_level_mem = [u32(next(it)) for _ in range(5)]
        next(it)
        self.l1 = np.ascontiguousarray(next(it))
        if self.l1.size == 0:
            self.l1 = np.zeros(4, dtype=np.uint8)
        self.isa = isa
        if self.sparse:
            assert self.mem_start.size == self.num_node + 1 and int(self.mem_start[-1]) == self.l0.size
        else:
            assert self.l0_node_mem == (1 + self.l0_max_degree) * 4 + 4 + 4 * self.feat_dim

    def _struct(self):
        s = _HnswIndex()
        s.num_node, s.maxM, s.maxM0, s.efC, s.max_level, s.init_node = (self.num_node, self.maxM, self.maxM0, self.efC,
                                                                            self.max_level, self.init_node)
        s.feat_dim, s.l0_max_degree, s.l0_node_mem_size = self.feat_dim, self.l0_max_degree, self.l0_node_mem
        s.l0_buffer = self.l0.ctypes.data
        s.l1_max_level, s.l1_max_degree, s.l1_node_mem_size, s.l1_level_mem_size = (self.l1_max_level, self.l1_max_degree,
                                                                                      self.l1_node_mem, self.l1_level_mem)
        s.l1_buffer = self.l1.ctypes.data
        s.metric, s.isa = self.metric, self.isa
        s.sparse = 1 if self.sparse else 0
        s.mem_start = self.mem_start.ctypes.data if self.sparse else None
        return s

    def vectors(self):
        if self.sparse:  # -> scipy csr of the stored rows
            import scipy.sparse as smat
            off = (1 + self.l0_max_degree) * 4
            indptr, idx, val = [0], [], []
            for i in range(self.num_node):
                b = int(self.mem_start[i]) + off
                n = int(self.l0[b:b + 4].view(np.uint32)[0])
                val.append(self.l0[b + 4:b + 4 + 4 * n].view(np.float32))
                idx.append(self.l0[b + 4 + 4 * n:b + 4 + 8 * n].view(np.uint32))
                indptr.append(indptr[-1] + n)
            return smat.csr_matrix((np.concatenate(val), np.concatenate(idx), indptr), shape=(self.num_node, self.feat_dim))
        rec = self.l0.reshape(self.num_node, self.l0_node_mem)
        off = (1 + self.l0_max_degree) * 4 + 4
        return np.ascontiguousarray(rec[:, off:]).view(np.float32).reshape(self.num_node, self.feat_dim)

    def predict(self, X, efS, topk, return_counters=False):
        L = _hnsw_lib()
        nq = X.shape[0]
        idx = np.zeros((nq, topk), dtype=np.uint32)
        val = np.zeros((nq, topk), dtype=np.float32)
        cnt = np.zeros((nq, 3), dtype=np.uint64)
        s = self._struct()
        if self.sparse:
            import scipy.sparse as smat
            X = smat.csr_matrix(X, dtype=np.float32)
            X.sort_indices()
            assert X.shape[1] == self.feat_dim
            indptr = np.ascontiguousarray(X.indptr, dtype=np.uint64)
            indices = np.ascontiguousarray(X.indices, dtype=np.uint32)
            data = np.ascontiguousarray(X.data, dtype=np.float32)
            rc = L.hno_search_csr(byref(s), indptr.ctypes.data_as(POINTER(c_uint64)), indices.ctypes.data_as(POINTER(c_uint32)),
                                  data.ctypes.data_as(POINTER(c_float)), nq, efS, topk, idx.ctypes.data_as(POINTER(c_uint32)),
                                  val.ctypes.data_as(POINTER(c_float)), cnt.ctypes.data_as(POINTER(c_uint64)))
            assert rc == 0
            return (idx, val, cnt) if return_counters else (idx, val)
        X = np.ascontiguousarray(X, dtype=np.float32)
        assert X.shape[1] == self.feat_dim
        rc = L.hno_search(byref(s), X.ctypes.data_as(POINTER(c_float)), nq, efS, topk, idx.ctypes.data_as(POINTER(c_uint32)),
                          val.ctypes.data_as(POINTER(c_float)), cnt.ctypes.data_as(POINTER(c_uint64)))
        assert rc == 0
        if return_counters:
            return idx, val, cnt
        return idx, val
