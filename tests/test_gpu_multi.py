"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
K/V rows sharded with owner_count/owner_disp (attention-mpi.c:19-27), merged with the
reference's three collectives over NCCL (attention-mpi.c:342,354,380) or the fused peer-memory
merge; one process driving several GPUs, and one process per GPU (torchrun model)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _need_gpus(sdpa, n):
    if sdpa.device_count() < n:
        pytest.skip(f"needs {n} GPUs, {sdpa.device_count()} visible")


@pytest.mark.parametrize("merge", ["nccl2", "nccl3", "peer"])
@pytest.mark.parametrize("prec,m,n,d,atol", [("f32", 300, 1001, 64, 1e-5), ("f32", 129, 3, 80, 1e-5),
                                             ("bf16", 700, 5000, 128, 1e-2), ("bf16", 9000, 4100, 128, 1e-2),
                                             ("bf16x3", 700, 5000, 128, 1e-5), ("auto", 300, 1001, 80, 1e-5), ("bf16", 8300, 40000, 128, 1e-2)])
def test_single_process_two_gpus(sdpa, oracle, merge, prec, m, n, d, atol):
    _need_gpus(sdpa, 2)
    Q, K, V = oracle.make_inputs(m, n, d, d, seed=m + n)
    ref = oracle.attention_f64_numpy(Q, K, V)
    with sdpa.Context(precision=prec, num_local=2, merge=merge) as ctx:
        ctx.load_kv_host_full(K, V)       # n=3 with 2 shards: ragged; n < shards elsewhere -> empty shard
        got = ctx.attention_host(Q)
    np.testing.assert_allclose(got, ref, rtol=0, atol=atol)
    assert oracle.verify_rule(got, ref)


@pytest.mark.parametrize("prec,atol", [("f32", 1e-5), ("bf16", 1e-2)])
def test_single_process_q_sharded_distribution(sdpa, oracle, prec, atol):
    """SDPA_DIST_Q: K/V replicated, Q rows sharded over the GPUs, no exchange (small-n policy, mpi.c:213-231);
    DIST_AUTO picks it below the reference's 64 MiB threshold and keeps K/V sharding above it."""
    _need_gpus(sdpa, 2)
    Q, K, V = oracle.make_inputs(1301, 777, 128, 128, seed=41)      # ragged on both sides, several 512-row batches
    ref = oracle.attention_f64(Q, K, V)
    for dist in ("q", "auto"):
        with sdpa.Context(precision=prec, num_local=2, q_batch=512, distribution=dist) as ctx:
            ctx.load_kv_host_full(K, V)
            for _ in range(2):
                got = ctx.attention_host(Q)
            np.testing.assert_allclose(got, ref, rtol=0, atol=atol)
            assert ctx.last_timings()["fused_launches"] == 4     # 651 rows on GPU 0 and 650 on GPU 1: two 512-row batches each
            np.testing.assert_allclose(ctx.attention_host(Q[:1]), ref[:1], rtol=0, atol=atol)   # fewer rows than GPUs


def test_single_process_empty_shard(sdpa, oracle):
    _need_gpus(sdpa, 2)
    Q, K, V = oracle.make_inputs(50, 1, 32, 32, seed=5)   # n=1 < 2 shards: GPU 1 owns nothing (lmax=-inf, mpi.c:172)
    ref = oracle.attention_f64(Q, K, V)
    for merge in ("nccl2", "nccl3", "peer"):
        with sdpa.Context(precision="f32", num_local=2, merge=merge) as ctx:
            ctx.load_kv_host_full(K, V)
            got = ctx.attention_host(Q)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _overlap_worker(rank, world, port, out_dir):
    """Queued passes with SDPA_OVERLAP_PASSES=1 (experimental): the exchange of pass i overlaps the compute of pass i+1."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), SDPA_OVERLAP_PASSES="1")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from sdpa_b200 import parallel
    from oracle import oracle as o

    m, d = 1500, 128
    cases = [o.make_inputs(m, 4000 + 300 * k, d, d, seed=90 + k) for k in range(3)]   # different K/V/Q per pass
    # the last two: enough keys per shard for the persistent kernel -> the casts of pass i+1 run on a side stream beside the
    # fused kernel of pass i (cast-ahead), with both root forms of the exchange
    big = [o.make_inputs(m, 40960 + 512 * k, d, d, seed=70 + k) for k in range(3)]
    for prec, atol, root_merge, cases in (("bf16", 1e-2, "overlap", cases), ("f32", 1e-5, "overlap", cases), ("bf16", 1e-2, "instream", cases),
                                          ("f32", 1e-5, "instream", cases), ("bf16", 1e-2, "push", cases), ("f32", 1e-5, "push", cases),
                                          ("bf16", 1e-2, "pushsync", cases), ("f32", 1e-5, "pushsync", cases),
                                          ("bf16", 1e-2, "overlap", big), ("bf16", 1e-2, "instream", big), ("bf16", 1e-2, "push", big),
                                          ("bf16", 1e-2, "pushsync", big), ("bf16", 1e-2, "auto", cases), ("bf16", 1e-2, "auto", big)):
        os.environ["SDPA_ROOT_MERGE"] = root_merge
        # big cases: one Q batch per pass (cast-ahead takes single-batch passes only); the others ping-pong three batches
        ctx = parallel.bootstrap_context(precision=prec, q_batch=2048 if cases is big else 512, local_rank=rank, merge="peer")
        dev, outs = [], []
        for Q, K, V in cases:
            first, count = parallel.shard_rows(K.shape[0], world, rank)
            dev.append([torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (K[first:first + count], V[first:first + count], Q)])
            outs.append(torch.zeros(m, d, dtype=torch.float64, device="cuda") if rank == 0 else None)
        for rep in range(3):
            for (Kd, Vd, Qd), out in zip(dev, outs):
                ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [Kd.shape[0]], d, d, [Qd.data_ptr()],
                                          out.data_ptr() if out is not None else None, m, blocking=False)
        ctx.synchronize()
        if rank == 0:
            for (Q, K, V), out in zip(cases, outs):
                np.testing.assert_allclose(out.cpu().numpy(), o.attention_f64_numpy(Q, K, V), rtol=0, atol=atol)
        ctx.close()
    dist.barrier()
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def _rank_worker(rank, world, port, out_dir, id_file):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sdpa_b200
    from sdpa_b200 import parallel
    from oracle import oracle as o

    m, n, d = 1500, 6001, 128
    Q, K, V = o.make_inputs(m, n, d, d, seed=77)
    ref = o.attention_f64_numpy(Q, K, V) if rank == 0 else None
    first, count = parallel.shard_rows(n, world, rank)
    for prec, atol, merge in (("f32", 1e-5, "nccl2"), ("bf16", 1e-2, "nccl3"), ("bf16", 1e-2, "peer"), ("f32", 1e-5, "peer"),
                              ("bf16", 1e-2, "peer-sliced"), ("f32", 1e-5, "peer-sliced"), ("bf16x3", 1e-5, "peer"), ("auto", 1e-5, "nccl2"),
                              ("bf16", 1e-2, "peer-instream"), ("f32", 1e-5, "peer-instream"), ("bf16x3", 1e-5, "peer-instream"),
                              ("bf16", 1e-2, "peer-push"), ("f32", 1e-5, "peer-push"), ("bf16x3", 1e-5, "peer-push"),
                              ("bf16", 1e-2, "peer-pushsync"), ("f32", 1e-5, "peer-pushsync"), ("bf16", 1e-2, "peer-overlap"), ("f32", 1e-5, "peer-overlap")):
        # (1) pre-sharded inputs, one context per rank (bench.py's model); merge="peer" = CUDA-IPC device-side exchange
        # (the root GPU merges all rows; "peer-sliced" = every rank merges its share of the rows from a pushed inbox)
        # "peer-instream" = the root merges its own partial states and the other shards' states in one kernel of its compute stream
        os.environ["SDPA_IPC_MERGE"] = "sliced" if merge == "peer-sliced" else "root"
        # "peer-push" = every shard pushes its state into the root's inbox, the root merges it with a small background kernel
        # "peer-pushsync" = the same pushes, final merge of the inbox on the root's compute stream
        os.environ["SDPA_ROOT_MERGE"] = {"peer-instream": "instream", "peer-push": "push", "peer-pushsync": "pushsync", "peer": "auto"}.get(merge, "overlap")
        merge = "peer" if merge.startswith("peer-") else merge
        ctx = parallel.bootstrap_context(precision=prec, q_batch=512, local_rank=rank, merge=merge)
        ctx.load_kv_host([K[first:first + count]], [V[first:first + count]])
        for _ in range(2):   # twice: slot reuse across calls
            got = ctx.attention_host(Q)
        if rank == 0:
            np.testing.assert_allclose(got, ref, rtol=0, atol=atol)
        else:
            assert got is None
        # one-batch calls between the three-batch ones: the default root form switches per call (inbox + in-stream final merge
        # for single-batch passes, comm-stream merge otherwise) on the same slots, epochs and flags
        for rows in (400, 512):
            part = ctx.attention_host(Q[:rows])
            if rank == 0:
                np.testing.assert_allclose(part, ref[:rows], rtol=0, atol=atol)
        again = ctx.attention_host(Q)
        if rank == 0:
            np.testing.assert_allclose(again, ref, rtol=0, atol=atol)
        # (1b) device-resident, queued passes (bench.py's timed loop): five passes back to back, one wait
        Kd, Vd, Qd = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (K[first:first + count], V[first:first + count], Q))
        out = torch.zeros(m, d, dtype=torch.float64, device="cuda") if rank == 0 else None
        for _ in range(5):
            ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [count], d, d, [Qd.data_ptr()],
                                      out.data_ptr() if out is not None else None, m, blocking=False)
        ctx.synchronize()
        if rank == 0:
            np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=atol)
        # (2) the reference's calling convention: data on rank 0 only (mpi.c:508-517)
        got2 = ctx.scatter_attention(Q, K, V) if rank == 0 else ctx.scatter_attention()
        if rank == 0:
            np.testing.assert_allclose(got2, ref, rtol=0, atol=atol)
        ctx.close()
    # (3) the drop-in entry point attention(..., mpi_rank, mpi_size) with the file rendezvous
    os.environ["SDPA_NCCL_ID_FILE"] = id_file
    os.environ["SDPA_PRECISION"] = "f32"
    got3 = sdpa_b200.attention(Q, K, V, mpi_rank=rank, mpi_size=world) if rank == 0 else \
        sdpa_b200.attention(None, None, None, mpi_rank=rank, mpi_size=world)
    if rank == 0:
        np.testing.assert_allclose(got3, ref, rtol=0, atol=1e-5)
    sdpa_b200.runtime_shutdown()
    dist.barrier()
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_one_process_per_gpu(sdpa, oracle, tmp_path):
    _need_gpus(sdpa, 2)
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_rank_worker, args=(world, _free_port(), str(tmp_path), str(tmp_path / "nccl_id")), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


@pytest.mark.timeout(600)
def test_overlapped_queued_passes(sdpa, oracle, tmp_path):
    _need_gpus(sdpa, 2)
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_overlap_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _deferred_worker(rank, world, port, out_dir):
    """Queued passes across two processes without the exact twin in the stream: the pass whose guard fires on ONE shard only
    must be repaired by both processes at sdpa_synchronize (MAX all-reduce of the guard verdicts)."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      SDPA_DEFER_TWIN="2")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from sdpa_b200 import parallel
    from oracle import oracle as o

    m, d, n = 700, 128, 3000
    good = o.make_inputs(m, n, d, d, seed=3)
    Q, K, V = o.make_inputs(m, n, d, d, seed=11)
    K = K.copy()
    K[:128] *= 0.01            # shard 0's first key tile is tiny ...
    K[128:1400] *= 30.0        # ... the rest of shard 0 enormous: only shard 0 raises its guard; shard 1 is ordinary
    bad = (Q * 3.0, K, V)
    cases = [tuple(o.bf16_round(a).astype(np.float64) for a in c) for c in (good, bad, good)]
    for merge in ("peer", "nccl2"):
        ctx = parallel.bootstrap_context(precision="bf16", q_batch=512, local_rank=rank, merge=merge)
        dev, outs = [], []
        for Qc, Kc, Vc in cases:
            first, count = parallel.shard_rows(n, world, rank)
            dev.append([torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Kc[first:first + count], Vc[first:first + count], Qc)])
            outs.append(torch.zeros(m, d, dtype=torch.float64, device="cuda") if rank == 0 else None)
        for rep in range(2):
            for (Kd, Vd, Qd), out in zip(dev, outs):
                ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [Kd.shape[0]], d, d, [Qd.data_ptr()],
                                          out.data_ptr() if out is not None else None, m, blocking=False)
        ctx.synchronize()
        if rank == 0:
            for (Qc, Kc, Vc), out, tol in zip(cases, outs, (2e-3, 2e-2, 2e-3)):
                got = out.cpu().numpy()
                assert np.isfinite(got).all()
                np.testing.assert_allclose(got, o.attention_f64_numpy(Qc, Kc, Vc), rtol=0, atol=tol)
        ctx.close()
    dist.barrier()
    Path(out_dir, f"okd{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_deferred_guard_repair_across_processes(sdpa, oracle, tmp_path):
    _need_gpus(sdpa, 2)
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_deferred_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"okd{r}").exists() for r in range(world))
