"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU checkers for the two hot paths:

* ``oracle/_ref/libpecos_float32.so`` -- the reference's own ``pecos/core/libpecos.cpp`` compiled unmodified
  (recipe: ``oracle/Makefile``), driven through ctypes by :mod:`oracle.ref`.
* ``oracle/liboracle.so`` -- our plain-C restatement (``xlinear_oracle.c``, ``hnsw_oracle.c``), driven by
  :mod:`oracle.restatement`.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package.  Nothing under ``pecos_b200/`` does.
"""
import os
import subprocess

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libpecos_float32.so")
RESTATEMENT_LIB = os.path.join(ORACLE_DIR, "liboracle.so")


def build(verbose=False):
    """Compile liboracle.so, and oracle/_ref when the reference sources are present (no-op on the GPU box)."""
    out = subprocess.run(["make", "-C", ORACLE_DIR, "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref():
    return os.path.exists(REF_LIB)
