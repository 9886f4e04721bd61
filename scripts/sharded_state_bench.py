#!/usr/bin/env python
"""BASELINE config 5 driver: PauliOp.apply on ONE state sharded by its high qubits across the GPUs of a box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/sharded_state_bench.py --qubits 34 --strings 16 [--iters 2]

Each rank owns 2^(qubits - log2 N) rows (complex128, batch 1), generated on the device by the counter-based
generator, applies the operator with pairwise NCCL shard swaps (fast_pauli_b200.distributed.ShardedStateOp) and
checks sampled output rows against the closed form evaluated on regenerated inputs.  Rank 0 prints one JSON line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def _claim_stdout():
    """Keep the real stdout for the single JSON line: everything else written to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr.  Returns a writer for the JSON line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(text: str) -> None:
        os.write(real, (text + "\n").encode())

    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--strings", type=int, default=16)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--samples", type=int, default=24)
    a = ap.parse_args()
    emit = _claim_stdout()
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["FASTPAULI_DEVICE"] = str(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", f"/tmp/fp_nccl_debug_{os.getpid()}.log")  # keep stdout to the JSON line
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fp = load_package()
    from fast_pauli_b200 import distributed as fpd
    from fast_pauli_b200.synth import random_strings, uniform_complex_at

    n = a.qubits
    n_loc = n - int(np.log2(world))
    rows = 1 << n_loc
    rng = np.random.default_rng(1234)
    strings = random_strings(rng, n, a.strings)
    h = rng.uniform(-1, 1, a.strings) + 1j * rng.uniform(-1, 1, a.strings)
    ctx = fp.default_context()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    psi = torch.empty(rows, dtype=torch.complex128, device="cuda")
    out = torch.empty_like(psi)
    bufs = [torch.empty_like(psi), torch.empty_like(psi)]
    fp._check(fp.lib.fp_fill_uniform(ctx._h, fp.FP_C128, C.c_void_p(psi.data_ptr()), C.c_uint64(rows),
                                     C.c_uint64(rank * rows), C.c_uint64(18)))
    op = fpd.ShardedStateOp(strings, h, world, rank)
    n_swaps = op.apply(out, psi, bufs)  # warm-up (also builds the plans, opens the NCCL pairs)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        op.apply(out, psi, bufs)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # ---- sampled parity: out[i] = sum_s h_s m_s[i] psi[i ^ x_s] with the GLOBAL row index
    from oracle import oracle as orc  # checker only

    srng = np.random.default_rng(99 + rank)
    idx = srng.integers(0, rows, size=a.samples)
    got = out[torch.from_numpy(idx).cuda()].cpu().numpy()
    worst = 0.0
    masks = [orc.masks(s) for s in strings]
    base = np.array([1, -1j, -1, 1j])
    for k, il in enumerate(idx):
        i = rank * rows + int(il)
        acc = 0j
        for (x, z, ny), hs in zip(masks, h):
            src = uniform_complex_at(np.array([i ^ x], dtype=np.uint64), np.complex128, 18)[0]
            sign = -1.0 if bin(i & z).count("1") & 1 else 1.0
            acc += (hs * (base[ny] * sign)) * src
        worst = max(worst, abs(got[k] - acc) / max(abs(acc), 1e-300))
    w = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        shard_bytes = rows * 16
        line = {"workload": f"PauliOp.apply, one {n}-qubit complex128 state sharded by {int(np.log2(world))} high qubits",
                "n_gpus": world, "n_qubits": n, "n_strings": a.strings, "shard_bytes": shard_bytes,
                "peer_swaps_per_apply": n_swaps, "ms_per_apply": ms,
                "amp_strings_per_s": (1 << n) * a.strings / (ms * 1e-3),
                "nvlink_bytes_per_gpu_per_apply": n_swaps * shard_bytes,
                "exchange_GBps_per_gpu_if_exchange_bound": n_swaps * shard_bytes / (ms * 1e-3) / 1e9,
                "sampled_parity_max_rel_err": float(w.item()), "samples_per_rank": a.samples}
        emit(json.dumps(line))
    assert float(w.item()) < 1e-12, f"sharded parity {float(w.item()):.3e}"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
