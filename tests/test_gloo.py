"""World-size-2 host logic on the gloo backend (CPU): the bootstrap broadcast of the
ncclUniqueId, the shard map used by every rank, the MAX/SUM/SUM merge choreography
(attention-mpi.c:340-380) and the max-over-ranks timing reduction (mpi.c:524).
The per-shard partial states come from the oracle (test infrastructure) because the
CUDA kernels cannot run here; on the GPU box tests/test_gpu_multi.py runs the real path."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sdpa_b200
    from sdpa_b200 import parallel
    from oracle import oracle as o

    # 1. bootstrap payload broadcast (stands for the 128-byte ncclUniqueId)
    payload = bytes(range(128)) if rank == 0 else None
    got = parallel.broadcast_bytes(payload, 128, src=0)
    assert got == bytes(range(128))

    # 2. every rank derives the same shard map through the C ABI
    m, n, dk, dv = 37, 101, 24, 20
    Q, K, V = o.make_inputs(m, n, dk, dv, seed=7)
    first, count = parallel.shard_rows(n, world, rank)
    assert (first, count) == (o.owner_disp(n, world, rank), o.owner_count(n, world, rank))

    # 3. merge choreography over the process group, partial states from the oracle
    contrib, lmax, lsum = o.online_softmax_partials_f32(o.cvt_d2f(Q), o.cvt_d2f(K[first:first + count]),
                                                        o.cvt_d2f(V[first:first + count]))
    c, mx, sm = torch.from_numpy(contrib), torch.from_numpy(lmax), torch.from_numpy(lsum)
    gmax = mx.clone()
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)                  # mpi.c:342
    corr = torch.exp(mx - gmax)                                  # mpi.c:347
    sm = sm * corr
    c = c * corr[:, None]
    gsum = sm.clone()
    dist.all_reduce(gsum, op=dist.ReduceOp.SUM)                  # mpi.c:354
    inv = torch.where(gsum == 0, torch.zeros_like(gsum), 1.0 / gsum)
    c = c * inv[:, None]                                         # mpi.c:358-362
    dist.reduce(c, dst=0, op=dist.ReduceOp.SUM)                  # mpi.c:380
    if rank == 0:
        ref = o.attention_f64(Q, K, V)
        np.testing.assert_allclose(c.numpy().astype(np.float64), ref, rtol=0, atol=2e-6)

    # 4. timing reduction
    assert parallel.max_over_ranks(1.0 + rank) == float(world)
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_host_logic(sdpa, oracle, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _bench_data_worker(rank, world, port, out_dir):
    """bench.py's oracle check regenerates every rank's K/V shard on rank 0: the regenerated shards must be the arrays the ranks
    built for themselves (same seeds, same owner_count split), for the weak-scaling headline and for a fixed-n config."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    for name in ("c3", "c5"):
        m, rows = bench.config_shape(name, world, 0, 1000 if name == "c3" else 0)
        if name == "c5":
            rows = [r // 512 for r in rows]          # same split rule, test-sized
        K, V = bench.make_shard(rank, rows[rank])
        mine = torch.tensor([float(K.sum()), float(V.sum()), float(K[0, 0]), float(V[-1, -1]), float(rows[rank])], dtype=torch.float64)
        got = [torch.zeros(5, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(got, mine)
        if rank == 0:
            for r in range(world):
                Kr, Vr = bench.make_shard(r, rows[r])
                want = torch.tensor([float(Kr.sum()), float(Vr.sum()), float(Kr[0, 0]), float(Vr[-1, -1]), float(rows[r])], dtype=torch.float64)
                assert torch.equal(got[r], want), (name, r)
        q = bench.make_q(64)
        qs = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(qs, q.sum().reshape(1))
        assert all(torch.equal(x, qs[0]) for x in qs)     # Q is replicated: identical on every rank
    Path(out_dir, f"okb{rank}").write_text("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_bench_shards_are_reproducible(tmp_path):
    world = 2
    mp.spawn(_bench_data_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"okb{r}").exists() for r in range(world))
