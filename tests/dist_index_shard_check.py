#!/usr/bin/env python
"""Multi-process check of index sharding over NCCL (not collected by pytest; launch with torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_index_shard_check.py

Every rank loads its leaf shard, all ranks predict the same batch through ShardedXLinearModel (ONE NCCL all-gather of the
per-rank top-k), and rank 0 compares the merged result with the unsharded single-GPU prediction (must be bit-identical).
Prints one JSON line with timings."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    from pecos_b200 import core, synth
    from pecos_b200.distributed import ShardedXLinearModel
    from pecos_b200.xlinear import XLinearModel

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = core.get_clib()
    lib.set_device(local)
    folder = os.path.join(tempfile.gettempdir(), "pb200_shard_check")
    if rank == 0:
        layers = synth.make_tree_model(5, [8, 64, 512, 20000], 20000, 64, bias=1.0, even=True)
        synth.save_xlinear_model(folder, layers, bias=1.0, only_topk=20)
    dist.barrier()
    X = synth.make_queries(6, 5000, 20000, 64)
    sharded = ShardedXLinearModel.load(folder, device=local)
    got = sharded.predict(X, beam_size=20, only_topk=10)
    t0 = time.perf_counter()
    for _ in range(5):
        got = sharded.predict(X, beam_size=20, only_topk=10)
    dt = (time.perf_counter() - t0) / 5
    ok = True
    if rank == 0:
        whole = XLinearModel.load(folder, is_predict_only=True)
        want = whole.predict(X, beam_size=20, only_topk=10)
        ok = (np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
              and np.array_equal(got.data.view(np.uint32), want.data.view(np.uint32)))
        print(json.dumps({"check": "index_shard_nccl", "world": world, "bit_identical": bool(ok), "shard": sharded.shard,
                          "queries": int(X.shape[0]), "e2e_ms_per_call": 1e3 * dt, "queries_per_s": X.shape[0] / dt}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
