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


def sampled_parity(out, rank, rows, strings, h, samples) -> float:
    """max relative error of sampled output rows against out[i] = sum_s h_s m_s[i] psi[i ^ x_s] (GLOBAL row index),
    evaluated on inputs regenerated from the counter-based generator; max over ranks."""
    from fast_pauli_b200.synth import uniform_complex_at

    def masks_of(string):  # closed form of get_sparse_repr (PS:49-118): x, z masks and the number of Y
        nq = len(string)
        x = sum(1 << (nq - 1 - q) for q, ch in enumerate(string) if ch in "XY")
        z = sum(1 << (nq - 1 - q) for q, ch in enumerate(string) if ch in "YZ")
        return x, z, string.count("Y") & 3

    srng = np.random.default_rng(99 + rank)
    idx = srng.integers(0, rows, size=samples)
    got = out[torch.from_numpy(idx).cuda()].cpu().numpy()
    worst = 0.0
    masks = [masks_of(s) for s in strings]
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
    return float(w.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--strings", type=int, default=16)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--samples", type=int, default=24)
    ap.add_argument("--mode", default="both", choices=["nccl", "peer", "both"])
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
    results = {}
    n_swaps = len(op.plan.peer_offsets())

    def timed(fn):
        fn()  # warm-up (also builds the plans, opens the NCCL pairs / peer mappings)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def check():
        return sampled_parity(out, rank, rows, strings, h, a.samples)

    if a.mode in ("nccl", "both"):
        results["nccl_ms"] = timed(lambda: op.apply(out, psi, bufs))
        results["nccl_parity"] = check()
    if a.mode in ("peer", "both"):
        # fused exchange + apply: the kernels gather straight from the peers' shards over NVLink (CUDA IPC mappings)
        del bufs
        torch.cuda.empty_cache()
        psi_own = ctx.empty((rows,), np.complex128)  # plain cudaMalloc allocation: exportable through CUDA IPC
        fp._check(fp.lib.fp_memcpy(ctx._h, C.c_void_p(psi_own.ptr), C.c_void_p(psi.data_ptr()), C.c_size_t(rows * 16)))
        peers = fpd.PeerShards(psi_own, dist)
        out.zero_()

        def peer_apply():
            op.apply_peer(out.data_ptr(), peers.ptrs, np.complex128)

        results["peer_ms"] = timed(peer_apply)
        results["peer_parity"] = check()
        dist.barrier()
        peers.close()
    ms = results.get("peer_ms", results.get("nccl_ms"))

    worst = max(v for k, v in results.items() if k.endswith("_parity"))
    if rank == 0:
        shard_bytes = rows * 16
        line = {"workload": f"PauliOp.apply, one {n}-qubit complex128 state sharded by {int(np.log2(world))} high qubits",
                "n_gpus": world, "n_qubits": n, "n_strings": a.strings, "shard_bytes": shard_bytes,
                "peer_offsets_per_apply": n_swaps, "ms_per_apply": ms, "results": results,
                "amp_strings_per_s": (1 << n) * a.strings / (ms * 1e-3),
                "nvlink_bytes_per_gpu_per_apply": n_swaps * shard_bytes,
                "nvlink_GBps_per_gpu_per_direction": n_swaps * shard_bytes / (ms * 1e-3) / 1e9,
                "nvlink_frac_of_measured_770GBps": n_swaps * shard_bytes / (ms * 1e-3) / 1e9 / 770.0,
                "sampled_parity_max_rel_err": worst, "samples_per_rank": a.samples}
        emit(json.dumps(line))
    assert worst < 1e-12, f"sharded parity {worst:.3e}"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
