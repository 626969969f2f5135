"""TEST INFRASTRUCTURE ONLY -- ctypes driver for ``oracle/_ref/libpecos_float32.so`` (the unmodified reference library).

Binds the same symbols, with the same prototypes, that ``pecos/core/base.py`` binds (:799-976 XR-Linear,
:1865-1949 HNSW), without importing the reference's Python package (which does not exist on the GPU box).
"""
import ctypes
import os
from ctypes import POINTER, byref, c_bool, c_char_p, c_float, c_int, c_int32, c_uint32, c_void_p

import numpy as np
import scipy.sparse as smat

from pecos_b200.core import (
    XLINEAR_INFERENCE_MODEL_TYPES,
    ScipyCompressedSparseAllocator,
    ScipyCscF32,
    ScipyCsrF32,
    ScipyDrmF32,
)

from . import REF_LIB

_lib = None


def bind(L):
    """Declares the prototypes of the reference symbols used by the tests on a ctypes library object (the reference library,
    or -- tests/test_overlay_gpu.py -- a stand-in for ``pecos.core.clib.clib_float32`` that gets overlaid)."""
    L.c_xlinear_load_model_from_disk_ext.restype = c_void_p
    L.c_xlinear_load_model_from_disk_ext.argtypes = [c_char_p, c_int]
    L.c_xlinear_load_mmap_model_from_disk.restype = c_void_p
    L.c_xlinear_load_mmap_model_from_disk.argtypes = [c_char_p, c_bool]
    L.c_xlinear_compile_mmap_model.restype = None
    L.c_xlinear_compile_mmap_model.argtypes = [c_char_p, c_char_p]
    L.c_xlinear_destruct_model.restype = None
    L.c_xlinear_destruct_model.argtypes = [c_void_p]
    L.c_xlinear_get_int_attr.restype = c_uint32
    L.c_xlinear_get_int_attr.argtypes = [c_void_p, c_char_p]
    pred = [c_uint32, c_char_p, c_uint32, c_int, ScipyCompressedSparseAllocator.CFUNCTYPE]
    L.c_xlinear_predict_csr_f32.restype = None
    L.c_xlinear_predict_csr_f32.argtypes = [c_void_p, POINTER(ScipyCsrF32)] + pred
    L.c_xlinear_predict_drm_f32.restype = None
    L.c_xlinear_predict_drm_f32.argtypes = [c_void_p, POINTER(ScipyDrmF32)] + pred
    sel = [POINTER(ScipyCsrF32), c_char_p, c_int, ScipyCompressedSparseAllocator.CFUNCTYPE]  # pecos/core/base.py:846-876
    L.c_xlinear_predict_on_selected_outputs_csr_f32.restype = None
    L.c_xlinear_predict_on_selected_outputs_csr_f32.argtypes = [c_void_p, POINTER(ScipyCsrF32)] + sel
    L.c_xlinear_predict_on_selected_outputs_drm_f32.restype = None
    L.c_xlinear_predict_on_selected_outputs_drm_f32.argtypes = [c_void_p, POINTER(ScipyDrmF32)] + sel
    single = [POINTER(ScipyCsrF32), POINTER(ScipyCscF32), POINTER(ScipyCscF32), c_char_p, c_uint32, c_int, c_float,
              ScipyCompressedSparseAllocator.CFUNCTYPE]  # pecos/core/base.py:880-905
    L.c_xlinear_single_layer_predict_csr_f32.restype = None
    L.c_xlinear_single_layer_predict_csr_f32.argtypes = [POINTER(ScipyCsrF32)] + single
    L.c_xlinear_single_layer_predict_drm_f32.restype = None
    L.c_xlinear_single_layer_predict_drm_f32.argtypes = [POINTER(ScipyDrmF32)] + single
    single_sel = [POINTER(ScipyCsrF32), POINTER(ScipyCsrF32), POINTER(ScipyCscF32), POINTER(ScipyCscF32), c_char_p, c_int, c_float,
                  ScipyCompressedSparseAllocator.CFUNCTYPE]  # pecos/core/base.py:936-976
    L.c_xlinear_single_layer_predict_on_selected_outputs_csr_f32.restype = None
    L.c_xlinear_single_layer_predict_on_selected_outputs_csr_f32.argtypes = [POINTER(ScipyCsrF32)] + single_sel
    L.c_xlinear_single_layer_predict_on_selected_outputs_drm_f32.restype = None
    L.c_xlinear_single_layer_predict_on_selected_outputs_drm_f32.argtypes = [POINTER(ScipyDrmF32)] + single_sel
    # single-layer mmap handles (pecos/core/base.py:541-606)
    L.c_mlmodel_compile_mmap_model.restype = None
    L.c_mlmodel_compile_mmap_model.argtypes = [c_char_p, c_char_p]
    L.c_mlmodel_load_mmap_model.restype = c_void_p
    L.c_mlmodel_load_mmap_model.argtypes = [c_char_p, c_bool]
    L.c_mlmodel_destruct_model.restype = None
    L.c_mlmodel_destruct_model.argtypes = [c_void_p]
    L.c_mlmodel_get_int_attr.restype = c_uint32
    L.c_mlmodel_get_int_attr.argtypes = [c_void_p, c_char_p]
    ml = [POINTER(ScipyCsrF32), c_char_p, c_uint32, c_int, ScipyCompressedSparseAllocator.CFUNCTYPE]
    L.c_mlmodel_predict_csr_f32.restype = None
    L.c_mlmodel_predict_csr_f32.argtypes = [c_void_p, POINTER(ScipyCsrF32)] + ml
    L.c_mlmodel_predict_drm_f32.restype = None
    L.c_mlmodel_predict_drm_f32.argtypes = [c_void_p, POINTER(ScipyDrmF32)] + ml
    mls = [POINTER(ScipyCsrF32), POINTER(ScipyCsrF32), c_char_p, c_int, ScipyCompressedSparseAllocator.CFUNCTYPE]
    L.c_mlmodel_predict_on_selected_outputs_csr_f32.restype = None
    L.c_mlmodel_predict_on_selected_outputs_csr_f32.argtypes = [c_void_p, POINTER(ScipyCsrF32)] + mls
    L.c_mlmodel_predict_on_selected_outputs_drm_f32.restype = None
    L.c_mlmodel_predict_on_selected_outputs_drm_f32.argtypes = [c_void_p, POINTER(ScipyDrmF32)] + mls
    for data_type, metric in (("drm", "ip"), ("drm", "l2"), ("csr", "ip"), ("csr", "l2")):
        sfx = f"{data_type}_{metric}_f32"
        mat_t = ScipyDrmF32 if data_type == "drm" else ScipyCsrF32
        if not hasattr(L, "c_ann_hnsw_load_" + sfx):
            continue
        f = getattr(L, "c_ann_hnsw_train_" + sfx)
        f.restype = c_void_p
        f.argtypes = [POINTER(mat_t), c_uint32, c_uint32, c_int, c_int]
        f = getattr(L, "c_ann_hnsw_load_" + sfx)
        f.restype = c_void_p
        f.argtypes = [c_char_p, c_bool]
        f = getattr(L, "c_ann_hnsw_save_" + sfx)
        f.restype = None
        f.argtypes = [c_void_p, c_char_p]
        f = getattr(L, "c_ann_hnsw_destruct_" + sfx)
        f.restype = None
        f.argtypes = [c_void_p]
        f = getattr(L, "c_ann_hnsw_searchers_create_" + sfx)
        f.restype = c_void_p
        f.argtypes = [c_void_p, c_uint32]
        f = getattr(L, "c_ann_hnsw_searchers_destruct_" + sfx)
        f.restype = None
        f.argtypes = [c_void_p]
        f = getattr(L, "c_ann_hnsw_predict_" + sfx)
        f.restype = None
        f.argtypes = [c_void_p, POINTER(mat_t), POINTER(c_uint32), POINTER(c_float), c_uint32, c_uint32,
                      c_int32, c_void_p]
    return L


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(REF_LIB):
            raise RuntimeError(f"{REF_LIB} missing: run `make -C oracle` where /root/reference is available")
        _lib = bind(ctypes.CDLL(REF_LIB))
    return _lib


class RefXLinear(object):
    """Reference XR-Linear model handle (predict-only)."""

    def __init__(self, ranker_folder, weight_matrix_type="BINARY_SEARCH_CHUNKED", is_mmap=False):
        L = lib()
        if is_mmap:
            self.h = c_void_p(L.c_xlinear_load_mmap_model_from_disk(ranker_folder.encode(), False))
        else:
            self.h = c_void_p(L.c_xlinear_load_model_from_disk_ext(ranker_folder.encode(),
                                                                   XLINEAR_INFERENCE_MODEL_TYPES[weight_matrix_type]))

    def __del__(self):
        try:
            if self.h:
                lib().c_xlinear_destruct_model(self.h)
                self.h = None
        except Exception:
            pass

    def attr(self, name):
        return lib().c_xlinear_get_int_attr(self.h, name.encode())

    def predict(self, X, beam_size=0, post_processor=None, only_topk=0, threads=-1):
        L = lib()
        alloc = ScipyCompressedSparseAllocator()
        pp = post_processor.encode() if post_processor else None
        if isinstance(X, smat.csr_matrix):
            assert X.has_sorted_indices
            cx = ScipyCsrF32.init_from(X)
            L.c_xlinear_predict_csr_f32(self.h, byref(cx), beam_size or 0, pp, only_topk or 0, threads, alloc.cfunc)
        else:
            cx = ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32))
            L.c_xlinear_predict_drm_f32(self.h, byref(cx), beam_size or 0, pp, only_topk or 0, threads, alloc.cfunc)
        return alloc.get()


def predict_on_selected_outputs(model, X, selected_outputs_csr, post_processor=None, threads=-1):
    """c_xlinear_predict_on_selected_outputs_{csr,drm}_f32 (pecos/core/libpecos.cpp:179-198); `model` must be a RefXLinear
    loaded with weight_matrix_type="CSC" (inference.hpp:2143-2147)."""
    L = lib()
    alloc = ScipyCompressedSparseAllocator()
    cs = ScipyCsrF32.init_from(smat.csr_matrix(selected_outputs_csr, dtype=np.float32))
    pp = post_processor.encode() if post_processor else None
    if isinstance(X, smat.csr_matrix):
        assert X.has_sorted_indices
        cx = ScipyCsrF32.init_from(X)
        L.c_xlinear_predict_on_selected_outputs_csr_f32(model.h, byref(cx), byref(cs), pp, threads, alloc.cfunc)
    else:
        cx = ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32))
        L.c_xlinear_predict_on_selected_outputs_drm_f32(model.h, byref(cx), byref(cs), pp, threads, alloc.cfunc)
    return alloc.get()


def single_layer_predict(X, csr_codes, W, C, post_processor, only_topk, bias, threads=-1):
    """c_xlinear_single_layer_predict_{csr,drm}_f32 of the reference (pecos/core/libpecos.cpp:201-235): one layer of the
    python prediction chain, `csr_codes` = the previous layer's prediction (None for the first layer)."""
    L = lib()
    alloc = ScipyCompressedSparseAllocator()
    cw = ScipyCscF32.init_from(smat.csc_matrix(W, dtype=np.float32))
    cc = ScipyCscF32.init_from(smat.csc_matrix(C, dtype=np.float32))
    codes = None
    if csr_codes is not None:
        codes = byref(ScipyCsrF32.init_from(smat.csr_matrix(csr_codes, dtype=np.float32)))
    if isinstance(X, smat.csr_matrix):
        assert X.has_sorted_indices
        cx = ScipyCsrF32.init_from(X)
        fn = L.c_xlinear_single_layer_predict_csr_f32
    else:
        cx = ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32))
        fn = L.c_xlinear_single_layer_predict_drm_f32
    fn(byref(cx), codes, byref(cw), byref(cc), post_processor.encode(), only_topk, threads, bias, alloc.cfunc)
    return alloc.get()


def single_layer_predict_on_selected_outputs(X, selected_outputs_csr, csr_codes, W, C, post_processor, bias, threads=-1, clib=None):
    """c_xlinear_single_layer_predict_on_selected_outputs_{csr,drm}_f32 (pecos/core/libpecos.cpp:238-273) of the reference library,
    or of `clib` (a ctypes library with the same prototypes, e.g. the CUDA library)."""
    L = clib if clib is not None else lib()
    alloc = ScipyCompressedSparseAllocator()
    cw = ScipyCscF32.init_from(smat.csc_matrix(W, dtype=np.float32))
    cc = ScipyCscF32.init_from(smat.csc_matrix(C, dtype=np.float32))
    sel = smat.csr_matrix(selected_outputs_csr, dtype=np.float32)
    sel.sort_indices()
    cs = ScipyCsrF32.init_from(sel)
    codes = None
    if csr_codes is not None:
        codes = byref(ScipyCsrF32.init_from(smat.csr_matrix(csr_codes, dtype=np.float32)))
    if isinstance(X, smat.csr_matrix):
        assert X.has_sorted_indices
        cx = ScipyCsrF32.init_from(X)
        fn = L.c_xlinear_single_layer_predict_on_selected_outputs_csr_f32
    else:
        cx = ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32))
        fn = L.c_xlinear_single_layer_predict_on_selected_outputs_drm_f32
    fn(byref(cx), byref(cs), codes, byref(cw), byref(cc), post_processor.encode(), threads, bias, alloc.cfunc)
    return alloc.get()


class MLModelHandle(object):
    """Drives the c_mlmodel_* entry points (pecos/core/libpecos.cpp:37-113) of ANY library exporting them: the reference
    (default) or the CUDA library (`clib` = pecos_b200's ctypes handle) -- same prototypes, so the tests read the same."""

    def __init__(self, folder, clib=None, lazy_load=False):
        self.L = clib if clib is not None else lib()
        self.h = c_void_p(self.L.c_mlmodel_load_mmap_model(folder.encode(), lazy_load))

    def __del__(self):
        try:
            if self.h:
                self.L.c_mlmodel_destruct_model(self.h)
                self.h = None
        except Exception:
            pass

    def attr(self, name):
        return self.L.c_mlmodel_get_int_attr(self.h, name.encode())

    @staticmethod
    def _x(X):
        if isinstance(X, smat.csr_matrix):
            assert X.has_sorted_indices
            return ScipyCsrF32.init_from(X), "csr"
        return ScipyDrmF32.init_from(np.ascontiguousarray(X, dtype=np.float32)), "drm"

    def predict(self, X, csr_codes=None, post_processor=None, only_topk=0, threads=-1):
        alloc = ScipyCompressedSparseAllocator()
        cx, kind = self._x(X)
        codes = byref(ScipyCsrF32.init_from(smat.csr_matrix(csr_codes, dtype=np.float32))) if csr_codes is not None else None
        fn = getattr(self.L, f"c_mlmodel_predict_{kind}_f32")
        fn(self.h, byref(cx), codes, post_processor.encode() if post_processor else None, only_topk or 0, threads, alloc.cfunc)
        return alloc.get()

    def predict_on_selected_outputs(self, X, selected_outputs_csr, csr_codes=None, post_processor=None, threads=-1):
        alloc = ScipyCompressedSparseAllocator()
        cx, kind = self._x(X)
        cs = ScipyCsrF32.init_from(smat.csr_matrix(selected_outputs_csr, dtype=np.float32))
        codes = byref(ScipyCsrF32.init_from(smat.csr_matrix(csr_codes, dtype=np.float32))) if csr_codes is not None else None
        fn = getattr(self.L, f"c_mlmodel_predict_on_selected_outputs_{kind}_f32")
        fn(self.h, byref(cx), byref(cs), codes, post_processor.encode() if post_processor else None, threads, alloc.cfunc)
        return alloc.get()


def compile_mlmodel_mmap(npz_layer_folder, mmap_folder, clib=None):
    """c_mlmodel_compile_mmap_model: one `<d>.model` folder (npz) -> single-layer mmap folder."""
    (clib if clib is not None else lib()).c_mlmodel_compile_mmap_model(npz_layer_folder.encode(), mmap_folder.encode())


def compile_mmap_model(npz_ranker_folder, mmap_folder):
    lib().c_xlinear_compile_mmap_model(npz_ranker_folder.encode(), mmap_folder.encode())


class RefHNSW(object):
    """Reference HNSW index handle (float32; dense ``drm`` or sparse ``csr`` rows; ip or l2)."""

    def __init__(self, handle, metric, data_type="drm"):
        self.h, self.metric, self.data_type = handle, metric, data_type

    @staticmethod
    def _mat(X, data_type):
        """-> (ctypes matrix, keep-alive).  csr rows get sorted indices like pecos/ann/hnsw/model.py does."""
        if data_type == "drm":
            X = np.ascontiguousarray(X, dtype=np.float32)
            return ScipyDrmF32.init_from(X), X
        X = smat.csr_matrix(X, dtype=np.float32)
        X.sort_indices()
        return ScipyCsrF32.init_from(X), X

    @classmethod
    def train(cls, X, M=32, efC=100, metric="ip", threads=-1, max_level_upper_bound=-1):
        L = lib()
        data_type = "csr" if smat.issparse(X) else "drm"
        cx, keep = cls._mat(X, data_type)
        h = getattr(L, f"c_ann_hnsw_train_{data_type}_{metric}_f32")(byref(cx), M, efC, threads, max_level_upper_bound)
        return cls(c_void_p(h), metric, data_type)

    @classmethod
    def load(cls, c_model_dir, metric="ip", lazy_load=False, data_type="drm"):
        h = getattr(lib(), f"c_ann_hnsw_load_{data_type}_{metric}_f32")(c_model_dir.encode(), lazy_load)
        return cls(c_void_p(h), metric, data_type)

    def _fn(self, slot):
        return getattr(lib(), f"c_ann_hnsw_{slot}_{self.data_type}_{self.metric}_f32")

    def save(self, c_model_dir):
        os.makedirs(c_model_dir, exist_ok=True)
        self._fn("save")(self.h, c_model_dir.encode())

    def predict(self, X, efS, topk, threads=1):
        cx, keep = self._mat(X, self.data_type)
        idx = np.zeros((X.shape[0], topk), dtype=np.uint32)
        val = np.zeros((X.shape[0], topk), dtype=np.float32)
        searchers = c_void_p(self._fn("searchers_create")(self.h, max(1, threads)))
        self._fn("predict")(self.h, byref(cx), idx.ctypes.data_as(POINTER(c_uint32)), val.ctypes.data_as(POINTER(c_float)), efS,
                            topk, threads, searchers)
        self._fn("searchers_destruct")(searchers)
        return idx, val

    def __del__(self):
        try:
            if self.h:
                self._fn("destruct")(self.h)
                self.h = None
        except Exception:
            pass
