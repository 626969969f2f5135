"""Multi-GPU XR-Linear prediction, one process per GPU (``torch.distributed``; NCCL on GPUs, gloo in CPU tests).

Two layouts:

* **query sharding** -- model replicated, rows of ``X`` split across ranks; no data-path collective
  (:func:`split_rows_by_nnz`, used by ``bench.py`` and by callers that scatter their own batches).
* **index sharding** -- the leaf layer's weight chunks are split across ranks (:class:`ShardedXLinearModel`); the upper
  layers are replicated so every rank walks the same global beam, each rank scores only the leaf chunks it owns and keeps
  a local top-k as ``(key, label, score)`` with ``key = (orderable(score) << 32) | ~global_position``; ONE all-gather of
  those lists and a merge kernel give the global top-k, bit-identical to the single-GPU result.

The reference has no inference-time model sharding (its parallelism is OpenMP threads inside one call,
pecos/core/xmc/inference.hpp:969-1005); this module is the multi-GPU counterpart of that loop.
"""
import json
import os
from ctypes import byref, c_uint32, c_void_p

import numpy as np

from .core import ScipyCompressedSparseAllocator, ScipyCsrF32, XLINEAR_INFERENCE_MODEL_TYPES, get_clib


def split_rows_by_nnz(indptr, world):
    """Contiguous row blocks with (nearly) equal non-zeros: returns ``world + 1`` row boundaries."""
    indptr = np.asarray(indptr, dtype=np.int64)
    rows = indptr.size - 1
    total = int(indptr[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(indptr, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), rows))
    bounds.append(rows)
    return bounds


class _TorchComm(object):
    """all_gather over torch.distributed for the fixed-shape local top-k buffers."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, local):
        import torch

        # concatenation along dim 0 is the layout both NCCL and gloo accept; rank r's block is out[r]
        shape = tuple(local.shape)
        out = torch.empty((self.world * shape[0],) + shape[1:], dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out.view((self.world,) + shape)


class ShardedXLinearModel(object):
    """Leaf-layer index-sharded XR-Linear model: ``predict`` returns the same CSR matrix on every rank."""

    def __init__(self, c_model, rank, world, comm, clib, depth_params):
        self.model_chain = c_model
        self.rank, self.world = rank, world
        self.comm = comm
        self._clib = clib
        self.pred_params = depth_params

    def __del__(self):
        try:
            if self.model_chain is not None:
                self._clib.xlinear_destruct_model(self.model_chain)
                self.model_chain = None
        except Exception:
            pass

    @classmethod
    def load(cls, model_folder, comm=None, weight_matrix_type="BINARY_SEARCH_CHUNKED", device=None):
        clib = get_clib()
        clib.require_gpu()
        comm = comm or _TorchComm()
        if device is not None:
            clib.set_device(device)
        ranker = os.path.join(model_folder, "ranker")
        param = json.load(open(os.path.join(ranker, "param.json")))
        is_mmap = bool(param.get("is_mmap", False))
        type_id = -1 if is_mmap else XLINEAR_INFERENCE_MODEL_TYPES[weight_matrix_type]
        h = c_void_p(clib.clib_float32.pb200_xlinear_load_sharded(ranker.encode("utf-8"), type_id, comm.rank, comm.world))
        depth = int(param["depth"])
        pp = [json.load(open(os.path.join(ranker, f"{d}.model", "param.json")))["pred_kwargs"] for d in range(depth)]
        return cls(h, comm.rank, comm.world, comm, clib, pp)

    @property
    def shard(self):
        out = (c_uint32 * 4)()
        self._clib.clib_float32.pb200_xlinear_get_shard(self.model_chain, out)
        return tuple(int(x) for x in out)

    def predict(self, X, beam_size=None, only_topk=None, post_processor=None):
        """Every rank passes the same ``X`` (float32 CSR with sorted indices)."""
        import torch

        assert X.dtype == np.float32 and X.has_sorted_indices
        c = self._clib.clib_float32
        k = int(only_topk or self.pred_params[-1]["only_topk"])
        rows = X.shape[0]
        dev = torch.device("cuda", c.pb200_get_device())
        # send buffer of the exchange: 16-byte {u64 key, u32 id, f32 value} records, viewed as int64 pairs for torch
        rec = torch.zeros((rows, k, 2), dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        cx = ScipyCsrF32.init_from(X)
        pp = post_processor.encode("utf-8") if post_processor else None
        stride = c.pb200_xlinear_sharded_local_csr_packed(self.model_chain, byref(cx), beam_size or 0, pp, only_topk or 0, k,
                                                          rec.data_ptr())
        if stride != k:  # fewer candidates than k can exist at all: the engine used a narrower stride
            rec = rec.view(-1)[: rows * stride * 2].view(rows, stride, 2)
        g_rec = self.comm.all_gather(rec.contiguous())  # THE exchange: one all-gather of one buffer
        torch.cuda.synchronize(dev)
        self.last_exchange_bytes = int(rec.numel() * 8)
        alloc = ScipyCompressedSparseAllocator()
        c.pb200_xlinear_sharded_merge_packed(self.model_chain, self.world, rows, stride, only_topk or 0, g_rec.data_ptr(), alloc.cfunc)
        return alloc.get()
