"""GPU parity: the CUDA path, called through the C ABI, against the oracle and the golden
fixtures produced by the reference itself.  Stated tolerances (max abs error vs the fp64
definition attention.c:20-75, N(0,1) inputs):
    f32 path  : 1e-5        (the reference's own fp32 path sits at ~4e-7)
    bf16 path : 1e-2        (bf16 operands, fp32 accumulation)
and always the reference's acceptance rule |err| <= 0.02 (attention-mpi.c:476)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32_ATOL = 1e-5
BF16_ATOL = 1e-2


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (run with -m gpu on the B200 box)")
    return torch


def _inputs(oracle, mt):
    return oracle.make_inputs(mt["m"], mt["n"], mt["dk"], mt["dv"], mt["seed"], mt["gain"])


def test_library_is_loaded_and_sees_the_gpu(sdpa, torch_cuda):
    assert sdpa.device_count() >= 1
    assert sdpa.LIB_PATH.exists()


# ---------------------------------------------------------------- casts
@pytest.mark.parametrize("count", [0, 1, 7, 8, 1000, 4099, 1 << 20, (1 << 22) + 3])
def test_casts_bit_exact(sdpa, oracle, torch_cuda, count):
    torch = torch_cuda
    rng = np.random.default_rng(count)
    x = rng.standard_normal(count) * 10.0 ** rng.integers(-20, 20, count)
    xd = torch.from_numpy(x).cuda()
    f = torch.empty(count, dtype=torch.float32, device="cuda")
    sdpa.cvt_d2f(f.data_ptr(), xd.data_ptr(), count)
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy(), oracle.cvt_d2f(x))            # mpi.c:31-64, RN-even
    d = torch.empty(count, dtype=torch.float64, device="cuda")
    sdpa.cvt_f2d(d.data_ptr(), f.data_ptr(), count)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle.cvt_f2d(oracle.cvt_d2f(x)))  # mpi.c:68-101, exact
    h = torch.empty(count, dtype=torch.bfloat16, device="cuda")
    sdpa.cvt_d2bf16(h.data_ptr(), xd.data_ptr(), count)
    torch.cuda.synchronize()
    assert np.array_equal(h.float().cpu().numpy(), oracle.bf16_round(x))


def test_casts_unaligned_pointers(sdpa, oracle, torch_cuda):
    torch = torch_cuda
    x = np.random.default_rng(0).standard_normal(5003)
    xd = torch.from_numpy(x).cuda()
    f = torch.zeros(5003, dtype=torch.float32, device="cuda")
    sdpa.cvt_d2f(f.data_ptr() + 4, xd.data_ptr() + 8, 5001)   # 8-byte / 4-byte aligned only
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy()[1:5002], oracle.cvt_d2f(x[1:5002]))


# ---------------------------------------------------------------- fp32 path vs golden
def test_f32_matches_reference_outputs(sdpa, oracle, golden, torch_cuda):
    meta, data = golden
    for name, mt in meta.items():
        Q, K, V = _inputs(oracle, mt)
        with sdpa.Context(precision="f32") as ctx:
            ctx.load_kv_host_full(K, V)
            got = ctx.attention_host(Q)
            assert ctx.last_kernel() == "f32_simt"
        atol = F32_ATOL * mt["gain"]
        np.testing.assert_allclose(got, data[name + "/serial"], rtol=0, atol=atol, err_msg=name)
        np.testing.assert_allclose(got, data[name + "/mpi"], rtol=0, atol=atol, err_msg=name)
        assert oracle.verify_rule(got, data[name + "/serial"])


@pytest.mark.parametrize("shape", [(512, 512, 64, 64), (300, 1000, 128, 128), (130, 257, 80, 80),
                                   (65, 129, 5, 7), (64, 64, 256, 256), (1, 4096, 128, 128), (777, 1, 64, 64)])
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_f32_vs_oracle_shapes(sdpa, oracle, torch_cuda, shape, splits):
    m, n, dk, dv = shape
    Q, K, V = oracle.make_inputs(m, n, dk, dv, seed=m + n)
    ref = oracle.attention_f64(Q, K, V)
    with sdpa.Context(precision="f32", kv_splits=splits) as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
    np.testing.assert_allclose(got, ref, rtol=0, atol=F32_ATOL)


def test_f32_small_q_batches_ping_pong(sdpa, oracle, torch_cuda):
    """m not a multiple of the batch, many batches: the ping-pong buffers (mpi.c:268-330)."""
    Q, K, V = oracle.make_inputs(1000, 700, 64, 64, seed=9)
    ref = oracle.attention_f64(Q, K, V)
    with sdpa.Context(precision="f32", q_batch=96) as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
        assert ctx.last_timings()["fused_launches"] == 11
    np.testing.assert_allclose(got, ref, rtol=0, atol=F32_ATOL)


def test_drop_in_attention_entry_point(sdpa, oracle, torch_cuda, monkeypatch):
    """attention(Q,K,V,result,m,n,dk,dv,0,1) -- c1 of BASELINE.json (serial config)."""
    monkeypatch.setenv("SDPA_PRECISION", "f32")
    Q, K, V = oracle.make_inputs(512, 512, 64, 64, seed=0)
    ref = oracle.attention_f64(Q, K, V)
    out = np.full((512, 64), np.nan)
    ret = sdpa.attention(Q, K, V, out, 512, 512, 64, 64, 0, 1)
    assert ret is out
    np.testing.assert_allclose(out, ref, rtol=0, atol=F32_ATOL)
    assert oracle.verify_rule(out, ref)
    sdpa.runtime_shutdown()


def test_online_softmax_partials_contract(sdpa, oracle, torch_cuda):
    """(contrib, lmax, lsum) of online_softmax_attention (mpi.c:168-189) for a batch of rows."""
    torch = torch_cuda
    Q, K, V = oracle.make_inputs(70, 333, 64, 48, seed=4)
    Qf = oracle.cvt_d2f(Q)
    c_ref, lmax_ref, lsum_ref = oracle.online_softmax_partials_f32(Qf, oracle.cvt_d2f(K), oracle.cvt_d2f(V))
    with sdpa.Context(precision="f32", kv_splits=3) as ctx:
        ctx.load_kv_host_full(K, V)
        qd = torch.from_numpy(Qf).cuda()
        c = torch.empty(70, 48, device="cuda")
        mx = torch.empty(70, device="cuda")
        sm = torch.empty(70, device="cuda")
        ctx.online_softmax_partials(qd.data_ptr(), 70, c.data_ptr(), mx.data_ptr(), sm.data_ptr())
    np.testing.assert_allclose(mx.cpu().numpy(), lmax_ref, rtol=0, atol=1e-5)
    np.testing.assert_allclose(sm.cpu().numpy(), lsum_ref, rtol=1e-5, atol=0)
    np.testing.assert_allclose(c.cpu().numpy(), c_ref, rtol=2e-5, atol=2e-5)


def test_empty_and_degenerate(sdpa, oracle, torch_cuda):
    Q, K, V = oracle.make_inputs(4, 3, 8, 8, seed=1)
    with sdpa.Context(precision="f32") as ctx:
        ctx.load_kv_host_full(K, V)
        assert ctx.attention_host(Q[:0]).shape == (0, 8)     # m = 0
        ctx.load_kv_host_full(K[:0], V[:0])                  # n = 0: gsum == 0 -> zeros (mpi.c:359)
        got = ctx.attention_host(Q)
        assert np.array_equal(got, np.zeros((4, 8)))


def test_device_resident_api(sdpa, oracle, torch_cuda):
    torch = torch_cuda
    Q, K, V = oracle.make_inputs(256, 512, 64, 64, seed=3)
    ref = oracle.attention_f64(Q, K, V)
    Qd, Kd, Vd = (torch.from_numpy(a).cuda() for a in (Q, K, V))
    out = torch.zeros(256, 64, dtype=torch.float64, device="cuda")
    with sdpa.Context(precision="f32") as ctx:
        ctx.load_kv_device_ptrs([Kd.data_ptr()], [Vd.data_ptr()], [512], 64, 64)
        ctx.attention_device_ptrs([Qd.data_ptr()], out.data_ptr(), 256)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=F32_ATOL)


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_queued_passes_match_blocking(sdpa, oracle, torch_cuda, precision):
    """sdpa_enqueue_device_full: passes queued back to back (different K/V/Q each) give the blocking results."""
    torch = torch_cuda
    cases = [oracle.make_inputs(384, 1024 + 256 * k, 128, 128, seed=20 + k) for k in range(3)]
    dev = [[torch.from_numpy(a).cuda() for a in c] for c in cases]
    outs_q = [torch.zeros(384, 128, dtype=torch.float64, device="cuda") for _ in cases]
    outs_b = [torch.zeros(384, 128, dtype=torch.float64, device="cuda") for _ in cases]
    with sdpa.Context(precision=precision) as ctx:
        for (Qd, Kd, Vd), o in zip(dev, outs_b):
            ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [Kd.shape[0]], 128, 128, [Qd.data_ptr()], o.data_ptr(), 384)
        for rep in range(2):
            for (Qd, Kd, Vd), o in zip(dev, outs_q):
                ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [Kd.shape[0]], 128, 128, [Qd.data_ptr()], o.data_ptr(), 384,
                                          blocking=False)
        ctx.synchronize()
        acc = ctx.accumulated_timings(reset=True)
        assert acc["calls"] == 9 and acc["total_ms"] >= acc["fused_ms"] > 0
    for (Q, K, V), a, b in zip(cases, outs_q, outs_b):
        assert torch.equal(a, b)
        tol = F32_ATOL if precision == "f32" else BF16_ATOL
        np.testing.assert_allclose(a.cpu().numpy(), oracle.attention_f64(Q, K, V), rtol=0, atol=tol)


def test_accumulated_stage_timings(sdpa, oracle, torch_cuda):
    """Stage times are recorded per call and only evaluated on demand; the accumulated view
    sums every call since the last reset (bench.py reads it once after its timed loop)."""
    Q, K, V = oracle.make_inputs(512, 2048, 128, 128, seed=5)
    with sdpa.Context(precision="bf16") as ctx:
        ctx.load_kv_host_full(K, V)
        ctx.accumulated_timings(reset=True)
        for _ in range(3):
            ctx.attention_host(Q)
        last = ctx.last_timings()
        acc = ctx.accumulated_timings(reset=True)
        assert acc["calls"] == 3 and acc["fused_launches"] == 3 * last["fused_launches"]
        assert 0 < last["fused_ms"] <= last["total_ms"]
        assert acc["fused_ms"] >= last["fused_ms"] and acc["total_ms"] >= acc["fused_ms"] > 0
        again = ctx.accumulated_timings()
        assert again["calls"] == 0 and again["total_ms"] == 0
        # many calls without a query: the event pools are folded, totals stay consistent
        for _ in range(700):
            ctx.attention_host(Q[:128])
        acc = ctx.accumulated_timings(reset=True)
        assert acc["calls"] == 700 and acc["total_ms"] >= acc["fused_ms"] > 0


def test_full_size_c2_properties(sdpa, oracle, torch_cuda):
    """BASELINE c2 (m=n=4096, d=128, fp32) at full size: a seeded row subset against the fp64
    oracle plus size-independent properties (rows are convex combinations of V rows; softmax
    shift invariance: adding a constant vector to every K row's score leaves the output)."""
    m = n = 4096
    Q, K, V = oracle.make_inputs(m, n, 128, 128, seed=1)
    with sdpa.Context(precision="f32") as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
        rows = np.random.default_rng(0).choice(m, 64, replace=False)
        ref = oracle.attention_f64_numpy(Q[rows], K, V)
        np.testing.assert_allclose(got[rows], ref, rtol=0, atol=F32_ATOL)
        assert np.all(got.max(axis=0) <= V.max(axis=0) + 1e-6) and np.all(got.min(axis=0) >= V.min(axis=0) - 1e-6)
        # permuting the keys (and values with them) must not change the result beyond fp32 rounding
        perm = np.random.default_rng(1).permutation(n)
        ctx.load_kv_host_full(K[perm], V[perm])
        got_p = ctx.attention_host(Q)
        np.testing.assert_allclose(got_p, got, rtol=0, atol=F32_ATOL)


def test_harness_protocol(sdpa, oracle, torch_cuda, tmp_path):
    """The C harness: `prog <file>` -> Correct!/Elapsed, Wrong! on a bad answer block."""
    import subprocess
    exe = sdpa.LIB_PATH.parent / "attention_b200"
    Q, K, V = oracle.make_inputs(512, 512, 64, 64, seed=0)
    ref = oracle.attention_f64(Q, K, V)
    good, bad = tmp_path / "good.bin", tmp_path / "bad.bin"
    oracle.write_data_file(good, Q, K, V, ref)
    oracle.write_data_file(bad, Q, K, V, ref + 0.05)
    r = subprocess.run([str(exe), str(good)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("Correct!\nElapsed time: ") and r.stdout.rstrip().endswith("us"), r
    # malloc'd matrices like the reference's read_matrix (mpi.c:417-423): the library stages them itself
    import os
    r = subprocess.run([str(exe), str(good)], capture_output=True, text=True, timeout=300, env={**os.environ, "HARNESS_PAGEABLE": "1"})
    assert r.returncode == 0 and r.stdout.startswith("Correct!\nElapsed time: "), r
    r = subprocess.run([str(exe), str(bad)], capture_output=True, text=True, timeout=300)
    assert "Wrong!" in r.stdout
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Usage" in r.stderr
