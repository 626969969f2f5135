"""``HNSW`` -- search-only mirror of ``pecos.ann.hnsw.HNSW`` running on a B200.

Same names, argument meaning and error behaviour as the reference for the load / search path:

* ``HNSW.load(model_folder, lazy_load=False)`` ........... pecos/ann/hnsw/model.py:152-175
* ``HNSW.predict(X, pred_params, searchers, ret_csr)`` ... pecos/ann/hnsw/model.py:219-269
* ``HNSW.searchers_create`` / ``HNSW.Searchers`` ......... pecos/ann/hnsw/model.py:65-78, :198-209
* ``HNSW.PredParams(efS, topk, threads)`` ................ pecos/ann/hnsw/model.py:51-63

Index construction (``train``) and ``save`` stay on the reference CPU library; only dense (``drm``) float32 indices with
the ``ip`` or ``l2`` metric are served (sparse ``csr`` indices raise ``NotImplementedError``).
"""
import copy
import dataclasses as dc
import json
import os
from ctypes import POINTER, byref, c_bool, c_char_p, c_float, c_uint32, c_void_p

import numpy as np
import scipy.sparse as smat

from .core import ScipyDrmF32, get_clib


class HNSW(object):
    @dc.dataclass
    class PredParams(object):
        efS: int = 100
        topk: int = 10
        threads: int = 1

        @classmethod
        def from_dict(cls, d):
            d = d or {}
            return cls(efS=int(d.get("efS", 100)), topk=int(d.get("topk", 10)), threads=int(d.get("threads", 1)))

    class Searchers(object):
        def __init__(self, model, num_searcher=1):
            self.searchers_ptr = c_void_p(model.fn_dict["searchers_create"](model.model_ptr, num_searcher))
            self.destruct_fn = model.fn_dict["searchers_destruct"]

        def __del__(self):
            try:
                if self.searchers_ptr is not None:
                    self.destruct_fn(self.searchers_ptr)
                    self.searchers_ptr = None
            except Exception:
                pass

        def ctypes(self):
            return self.searchers_ptr

    def __init__(self, model_ptr, num_item, feat_dim, fn_dict, pred_params=None, data_type="drm", metric_type="ip"):
        self.model_ptr = model_ptr
        self.num_item = num_item
        self.feat_dim = feat_dim
        self.fn_dict = fn_dict
        self.pred_params = self.PredParams() if pred_params is None else pred_params
        self.data_type_ = data_type
        self.metric_type_ = metric_type

    def __del__(self):
        try:
            if self.model_ptr and self.fn_dict:
                self.fn_dict["destruct"](self.model_ptr)
                self.model_ptr = None
        except Exception:
            pass

    @property
    def data_type(self):
        return self.data_type_

    @property
    def metric_type(self):
        return self.metric_type_

    @staticmethod
    def create_pymat(X):
        """Wrap the query matrix (pecos/ann/hnsw/model.py:100-121); sparse queries are outside this engine's scope."""
        if isinstance(X, ScipyDrmF32):
            return X, "drm"
        if isinstance(X, np.ndarray):
            return ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32)), "drm"
        if isinstance(X, smat.csr_matrix):
            return None, "csr"
        raise ValueError("type(X)={} is NOT supported!".format(type(X)))

    @classmethod
    def load(cls, model_folder, lazy_load=False):
        with open("{}/param.json".format(model_folder), "r") as fin:
            param = json.loads(fin.read())
        if param["model"] != cls.__name__:
            raise ValueError("param[model] != cls.__name__")
        if not ("data_type" in param and "metric_type" in param):
            raise ValueError("param.json did not have data_type or metric_type!")
        clib = get_clib()
        fn_dict = clib.ann_hnsw_init(param["data_type"], param["metric_type"])
        c_model_dir = f"{model_folder}/c_model"
        if not os.path.isdir(c_model_dir):
            raise ValueError(f"c_model_dir did not exist: {c_model_dir}")
        clib.require_gpu()
        model_ptr = c_void_p(fn_dict["load"](c_char_p(c_model_dir.encode("utf-8")), c_bool(lazy_load)))
        pred_params = cls.PredParams.from_dict(param.get("pred_kwargs"))
        return cls(model_ptr, param["num_item"], param["feat_dim"], fn_dict, pred_params, param["data_type"], param["metric_type"])

    def searchers_create(self, num_searcher=1):
        if not self.model_ptr:
            raise ValueError("self.model_ptr must exist before using self.create_searcher()")
        if num_searcher <= 0:
            raise ValueError("num_searcher={} <= 0 is NOT valid".format(num_searcher))
        return HNSW.Searchers(self, num_searcher)

    def get_pred_params(self):
        return copy.deepcopy(self.pred_params)

    def predict(self, X, pred_params=None, searchers=None, ret_csr=True):
        pred_params = self.get_pred_params() if pred_params is None else pred_params
        pX, data_type = self.create_pymat(X)
        if data_type != self.data_type:
            raise ValueError("data_type={} is NOT consistent with self.data_type={}".format(data_type, self.data_type))
        if pX.cols != self.feat_dim:
            raise ValueError("pX.cols={} is NOT consistent with self.feat_dim={}".format(pX.cols, self.feat_dim))
        indices = np.zeros(pX.rows * pred_params.topk, dtype=np.uint32)
        distances = np.zeros(pX.rows * pred_params.topk, dtype=np.float32)
        self.fn_dict["predict"](
            self.model_ptr,
            byref(pX),
            indices.ctypes.data_as(POINTER(c_uint32)),
            distances.ctypes.data_as(POINTER(c_float)),
            pred_params.efS,
            pred_params.topk,
            pred_params.threads,
            None if searchers is None else searchers.ctypes(),
        )
        if not ret_csr:
            return indices.reshape(pX.rows, pred_params.topk), distances.reshape(pX.rows, pred_params.topk)
        indptr = np.arange(0, pred_params.topk * (pX.rows + 1), pred_params.topk, dtype=np.int64)
        return smat.csr_matrix((distances, indices.astype(np.int64), indptr), shape=(pX.rows, self.num_item), dtype=np.float32)
