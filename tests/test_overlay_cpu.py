"""The INTEGRATION.md overlay against the REAL reference Python package (CPU box only: needs /root/reference and oracle/_ref).

A scratch copy of the reference's `pecos` package (never copied into this repo) gets the reference library dropped in and the
one scipy >= 1.14 shim of tests/golden/make_golden.py; `pecos.core.clib` is then overlaid by pecos_b200.integration.overlay.
Checked without a GPU: every hot-path symbol of `clib.clib_float32` / `clib.ann_hnsw_fn_dict` now resolves into
libpecos_b200_float32.so with the reference's own restype / argtypes, every other symbol still resolves into the reference
library, and the reference's unrelated entry points keep working (sparse_matmul).  Calling the re-pointed symbols needs a
GPU; that part is tests/test_overlay_gpu.py (stand-in corelib over oracle/_ref)."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as smat

REFERENCE = os.environ.get("REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dl_path_of(cfunc):
    """Path of the shared object a ctypes function pointer lives in (dladdr)."""
    class DlInfo(ctypes.Structure):
        _fields_ = [("dli_fname", ctypes.c_char_p), ("dli_fbase", ctypes.c_void_p), ("dli_sname", ctypes.c_char_p), ("dli_saddr", ctypes.c_void_p)]

    libdl = ctypes.CDLL(None)
    libdl.dladdr.argtypes = [ctypes.c_void_p, ctypes.POINTER(DlInfo)]
    info = DlInfo()
    addr = ctypes.cast(cfunc, ctypes.c_void_p).value
    assert libdl.dladdr(addr, ctypes.byref(info)) != 0
    return info.dli_fname.decode()


def test_overlay_on_the_reference_python_package(tmp_path, built, have_ref):
    if not os.path.isdir(os.path.join(REFERENCE, "pecos")):
        pytest.skip("the reference checkout is not on this box")
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    scratch = str(tmp_path / "refpy")
    shutil.copytree(os.path.join(REFERENCE, "pecos"), os.path.join(scratch, "pecos"))
    subprocess.run(["chmod", "-R", "u+w", scratch], check=True)
    shutil.copy(os.path.join(ROOT, "oracle", "_ref", "libpecos_float32.so"), os.path.join(scratch, "pecos", "core", "libpecos_float32.so"))
    p = os.path.join(scratch, "pecos", "utils", "smat_util.py")
    src = open(p).read().replace("smat.sputils.get_index_dtype", "smat._sputils.get_index_dtype").replace("copy=False", "copy=None")
    open(p, "w").write(src)
    code = r"""
import sys, ctypes, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, scipy.sparse as smat
from pecos.core import clib
from pecos_b200 import integration
names = integration.overlay(clib, require_gpu=False)
out = {"swapped": names}
def where(fn):
    class I(ctypes.Structure):
        _fields_ = [("f", ctypes.c_char_p), ("b", ctypes.c_void_p), ("s", ctypes.c_char_p), ("a", ctypes.c_void_p)]
    dl = ctypes.CDLL(None); dl.dladdr.argtypes = [ctypes.c_void_p, ctypes.POINTER(I)]
    i = I(); dl.dladdr(ctypes.cast(fn, ctypes.c_void_p).value, ctypes.byref(i)); return i.f.decode()
out["xlinear"] = {n: where(getattr(clib.clib_float32, n)) for n in integration.XLINEAR_SYMBOLS}
out["hnsw"] = {"%%s_%%s_%%s" %% (k[0], k[1], s): where(f) for k, d in clib.ann_hnsw_fn_dict.items() for s, f in d.items() if hasattr(f, "argtypes")}
out["other"] = {n: where(getattr(clib.clib_float32, n)) for n in ("c_sparse_matmul_csc_f32", "c_xlinear_single_layer_train_csr_f32", "c_xlinear_single_layer_train_csc_f32" if hasattr(clib.clib_float32, "c_xlinear_single_layer_train_csc_f32") else "c_sparse_matmul_csr_f32")}
out["argtypes_kept"] = all(getattr(clib.clib_float32, n).argtypes is not None for n in integration.XLINEAR_SYMBOLS)
A = smat.random(20, 30, 0.2, format="csr", dtype=np.float32, random_state=1); B = smat.random(30, 10, 0.3, format="csc", dtype=np.float32, random_state=2)
C = clib.sparse_matmul(A, B)
out["matmul_ok"] = bool(np.allclose(C.toarray(), (A @ B).toarray(), atol=1e-5))
print("RESULT" + json.dumps(out))
""" % (scratch, ROOT)
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    assert all(p.endswith("libpecos_b200_float32.so") for p in out["xlinear"].values()), out["xlinear"]
    for key, path in out["hnsw"].items():
        slot = key.split("_", 2)[2]
        if slot in ("load", "destruct", "searchers_create", "searchers_destruct", "predict", "save"):
            assert path.endswith("libpecos_b200_float32.so"), (key, path)  # dense (drm) and sparse (csr) indices
        else:  # train stays on the reference library
            assert path.endswith("libpecos_float32.so"), (key, path)
    assert all(p.endswith("libpecos_float32.so") for p in out["other"].values()), out["other"]
    assert out["argtypes_kept"] and out["matmul_ok"]
    assert len(out["swapped"]) == len(set(out["swapped"])) >= 19 + 24
