#!/usr/bin/env python
"""BASELINE config 5 driver: PauliOp.apply on ONE state sharded by its high qubits across the GPUs of a box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/sharded_state_bench.py [--qubits 34] [--strings 16] [--chunk-mib 256]

Every rank owns 2^(qubits - log2 N) rows (complex128, one state), generated on the device by the counter-based
generator, and applies the operator through the C ABI (fp_sharded_op_apply: chunked ncclSend/ncclRecv exchange inside
the library, fast_pauli_b200.sharded).  torch.distributed is used for ONE thing: broadcasting the 128-byte NCCL unique
id.  Sampled output rows are checked against the closed form on regenerated inputs.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=None)
    ap.add_argument("--strings", type=int, default=16)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("FASTPAULI_DEVICE", str(local_rank))
    os.environ.setdefault("NCCL_DEBUG_FILE", f"/tmp/fp_nccl_debug_{os.getpid()}.log")
    real = os.dup(1)
    os.dup2(2, 1)  # NCCL banners go to stderr; stdout carries the JSON line only
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo")  # plumbing only (the id broadcast): no torch CUDA context at all
    fp = load_package()
    from fast_pauli_b200 import sharded

    ctx = fp.Context(local_rank)
    res = sharded.bench_config5(fp, ctx, dist, torch, rank, world, local_rank, n_qubits=args.qubits,
                                n_strings=args.strings)
    if rank == 0:
        os.write(real, (json.dumps(res) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
