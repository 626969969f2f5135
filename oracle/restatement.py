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
        res = _Result()
        rc = L.xlo_predict(D, Ws, Cs, bias, (c_int * D)(*kinds), (c_int * D)(*ps), (c_uint32 * D)(*ks), byref(q), byref(res))
        assert rc == 0
        n = int(res.nnz)
        indptr = np.ctypeslib.as_array(res.indptr, shape=(res.rows + 1,)).astype(np.int64)
        indices = np.ctypeslib.as_array(res.indices, shape=(max(n, 1),))[:n].astype(np.int64)
        data = np.ctypeslib.as_array(res.data, shape=(max(n, 1),))[:n].copy()
        out = smat.csr_matrix((data, indices, indptr), shape=(res.rows, res.cols))
        L.xlo_free_result(byref(res))
        return out
