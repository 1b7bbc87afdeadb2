"""GPU parity of the general tcgen05 kernel (attn_umma_general.cu): every dk, dv the reference's masked-tail path takes
(attention-mpi.c:115-119) that is a multiple of 8 up to 256, the fp32-accurate split precision "bf16x3" (what AUTO
selects), and the exact two-phase repair behind the overflow guard.  Stated tolerances (max abs error vs the fp64
definition attention.c:20-75 on N(0,1) inputs):
    bf16x3 : 1e-5   -- the reference's own arithmetic is fp32 (attention-mpi.c:168-189), its measured error 4e-7
    bf16   : 1e-2   -- and 2e-3 against fp64 math on the bf16-rounded operands
and always the reference's acceptance rule |err| <= 0.02 (attention-mpi.c:476)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X3_ATOL = 1e-5
BF16_ATOL = 1e-2
BF16_KERNEL_ATOL = 2e-3


def _run(sdpa, Q, K, V, precision, expect_kernel=None, **cfg):
    with sdpa.Context(precision=precision, **cfg) as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
        if expect_kernel is not None:
            assert ctx.last_kernel() == expect_kernel
    return got


def test_split_cast_is_exact(sdpa, oracle):
    """hi = bf16(fp32(x)), lo = bf16(fp32(x) - hi): bit-exact against the NumPy restatement; hi + lo keeps 16 bits of x."""
    import torch
    for count in (1, 8, 1001, (1 << 20) + 6):
        rng = np.random.default_rng(count)
        x = rng.standard_normal(count) * 10.0 ** rng.integers(-6, 6, count)
        xd = torch.from_numpy(x).cuda()
        hi = torch.empty(count, dtype=torch.bfloat16, device="cuda")
        lo = torch.empty(count, dtype=torch.bfloat16, device="cuda")
        sdpa.cvt_d2bf16x2(hi.data_ptr(), lo.data_ptr(), xd.data_ptr(), count)
        torch.cuda.synchronize()
        x32 = x.astype(np.float32)
        h = oracle.bf16_round(x32)
        l = oracle.bf16_round(x32 - h)
        assert np.array_equal(hi.float().cpu().numpy(), h)
        assert np.array_equal(lo.float().cpu().numpy(), l)
        assert np.max(np.abs((h.astype(np.float64) + l) - x32) / np.abs(x32)) < 2.0 ** -16


@pytest.mark.parametrize("m,n", [(128, 128), (256, 128), (100, 200), (1, 130), (300, 1000), (777, 2049)])
@pytest.mark.parametrize("splits", [0, 1, 3])
def test_x3_vs_oracle_d128(sdpa, oracle, m, n, splits):
    Q, K, V = oracle.make_inputs(m, n, 128, 128, seed=m + 3 * n)
    got = _run(sdpa, Q, K, V, "bf16x3", "bf16x3_umma", kv_splits=splits)
    ref = oracle.attention_f64_numpy(Q, K, V)
    np.testing.assert_allclose(got, ref, rtol=0, atol=X3_ATOL)
    assert oracle.verify_rule(got, ref)


@pytest.mark.parametrize("dk,dv", [(64, 64), (80, 48), (8, 8), (128, 64), (120, 128), (32, 104), (16, 128)])
def test_x3_shapes(sdpa, oracle, dk, dv):
    """Split precision on widths that are not 128: TMA zero-fills the box columns beyond dk / dv (the reference masks the
    vector tail instead); n = 333 leaves a ragged last key tile and an uneven split, m = 150 a partial row block."""
    Q, K, V = oracle.make_inputs(150, 333, dk, dv, seed=dk * 1000 + dv)
    ref = oracle.attention_f64_numpy(Q, K, V)
    for splits in (0, 2):
        got = _run(sdpa, Q, K, V, "bf16x3", "bf16x3_umma", kv_splits=splits)
        np.testing.assert_allclose(got, ref, rtol=0, atol=X3_ATOL)


@pytest.mark.parametrize("dk,dv", [(64, 64), (80, 48), (32, 200), (256, 256), (128, 256), (256, 128), (8, 8), (96, 128), (200, 136)])
def test_bf16_general_shapes(sdpa, oracle, dk, dv):
    Q, K, V = oracle.make_inputs(150, 600, dk, dv, seed=dk * 1000 + dv)
    ref = oracle.attention_f64_numpy(Q, K, V)
    Qb, Kb, Vb = (oracle.bf16_round(a).astype(np.float64) for a in (Q, K, V))
    ref_b = oracle.attention_f64_numpy(Qb, Kb, Vb)
    for splits in (0, 1):
        got = _run(sdpa, Q, K, V, "bf16", "bf16_umma_general", kv_splits=splits)
        np.testing.assert_allclose(got, ref, rtol=0, atol=BF16_ATOL)
        np.testing.assert_allclose(got, ref_b, rtol=0, atol=BF16_KERNEL_ATOL)


def test_golden_cases_on_tensor_cores(sdpa, oracle, golden):
    """The reference's own outputs (tests/golden, produced by attention.c and attention-mpi.c): every case whose widths
    are multiples of 8 runs on tcgen05 -- split precision at the fp32 tolerance where dk, dv <= 128, bf16 beyond."""
    meta, data = golden
    ran = 0
    for name, mt in meta.items():
        if mt["dk"] % 8 or mt["dv"] % 8:
            continue
        Q, K, V = oracle.make_inputs(mt["m"], mt["n"], mt["dk"], mt["dv"], mt["seed"], mt["gain"])
        if mt["dk"] <= 128 and mt["dv"] <= 128:
            got = _run(sdpa, Q, K, V, "bf16x3", "bf16x3_umma")
            atol = X3_ATOL * max(1.0, mt["gain"])
        else:
            got = _run(sdpa, Q, K, V, "bf16")
            atol = BF16_ATOL * max(1.0, mt["gain"])
        np.testing.assert_allclose(got, data[name + "/serial"], rtol=0, atol=atol, err_msg=name)
        assert oracle.verify_rule(got, data[name + "/mpi"]), name
        ran += 1
    assert ran >= 4


def test_auto_precision_keeps_the_reference_accuracy(sdpa, oracle):
    """AUTO never selects plain bf16: split precision on tensor cores where the shape allows, else the fp32 kernel."""
    for (dk, dv), kernel in (((128, 128), "bf16x3_umma"), ((80, 80), "bf16x3_umma"), ((64, 256), "bf16x3_umma"), ((200, 64), "f32_simt"),
                             ((100, 100), "f32_simt"), ((128, 256), "f32_simt")):
        Q, K, V = oracle.make_inputs(64, 200, dk, dv, seed=dk + dv)
        got = _run(sdpa, Q, K, V, "auto", kernel)
        np.testing.assert_allclose(got, oracle.attention_f64_numpy(Q, K, V), rtol=0, atol=X3_ATOL)
    Q, K, V = oracle.make_inputs(64, 128, 100, 100, seed=2)
    for prec in ("bf16", "bf16x3"):
        with pytest.raises(sdpa.SdpaError):
            _run(sdpa, Q, K, V, prec)          # dk = 100: no tensor-core kernel, and no silent fallback
    Q, K, V = oracle.make_inputs(64, 128, 200, 64, seed=2)
    with pytest.raises(sdpa.SdpaError):
        _run(sdpa, Q, K, V, "bf16x3")          # split precision: dk <= 128 (and dv <= 128, or dk <= 64 with dv <= 256)
    _run(sdpa, Q, K, V, "bf16", "bf16_umma_general")


def test_default_precision_passes_the_reference_gate_on_peaky_scores(sdpa, oracle):
    """Keys 3x N(0,1) and a score gain of 8: scores reach magnitudes where bf16 operand rounding alone moves the output
    beyond the reference's 0.02 rule (attention-mpi.c:476).  The default precision must print `Correct!` there."""
    Q, K, V = oracle.make_inputs(512, 4000, 128, 128, seed=77, score_gain=8.0)
    K = K * 3.0
    ref = oracle.attention_f64_numpy(Q, K, V)
    got = sdpa.attention(Q, K, V)              # the drop-in entry point, SDPA_PRECISION unset
    assert oracle.verify_rule(got, ref)
    assert np.abs(got - ref).max() < 2e-3


@pytest.mark.parametrize("prec,kernel", [("bf16x3", "bf16x3_umma"), ("bf16", "bf16_umma_general")])
def test_exact_variant_repairs_an_overflowing_launch(sdpa, oracle, prec, kernel):
    """First 128 keys tiny, the rest enormous: the fast pass (reference fixed by the first key tile) raises the guard and
    the exact two-phase variant recomputes the launch; without it the result would be inf/NaN."""
    dk = dv = 64 if prec == "bf16" else 128
    Q, K, V = oracle.make_inputs(300, 1500, dk, dv, seed=11)
    Q = Q * 3.0
    K[:128] *= 0.01
    K[128:] *= 30.0
    if prec == "bf16":
        Q, K, V = (oracle.bf16_round(a).astype(np.float64) for a in (Q, K, V))
    ref = oracle.attention_f64_numpy(Q, K, V)
    for splits in (1, 0):
        got = _run(sdpa, Q, K, V, prec, kernel, kv_splits=splits)
        assert np.isfinite(got).all()
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-2)


@pytest.mark.parametrize("prec,dk,dv,atol", [("bf16x3", 128, 128, X3_ATOL), ("bf16x3", 80, 48, X3_ATOL), ("bf16", 64, 64, BF16_ATOL),
                                             ("bf16", 256, 256, BF16_ATOL)])
def test_exact_variant_alone(sdpa, oracle, monkeypatch, prec, dk, dv, atol):
    """SDPA_UMMA_SAFE=1 runs only the exact variant (phase 1: row maxima of the key range, phase 2: the pass itself)."""
    monkeypatch.setenv("SDPA_UMMA_SAFE", "1")
    Q, K, V = oracle.make_inputs(200, 1111, dk, dv, seed=31)
    ref = oracle.attention_f64_numpy(Q, K, V)
    for splits in (0, 1, 4):
        got = _run(sdpa, Q, K, V, prec, kv_splits=splits)
        np.testing.assert_allclose(got, ref, rtol=0, atol=atol)


def test_general_kernel_on_the_headline_shape(sdpa, oracle, monkeypatch):
    """SDPA_UMMA_GENERAL=1 sends dk = dv = 128 bf16 to the general kernel as well: same contract as v7."""
    monkeypatch.setenv("SDPA_UMMA_GENERAL", "1")
    Q, K, V = oracle.make_inputs(600, 2500, 128, 128, seed=21)
    got = _run(sdpa, Q, K, V, "bf16")
    Qb, Kb, Vb = (oracle.bf16_round(a).astype(np.float64) for a in (Q, K, V))
    np.testing.assert_allclose(got, oracle.attention_f64_numpy(Qb, Kb, Vb), rtol=0, atol=BF16_KERNEL_ATOL)


def test_x3_ping_pong_batches_device_and_host_agree(sdpa, oracle):
    """Several Q batches (mpi.c:268-330) through the host path and the queued device path."""
    import torch
    Q, K, V = oracle.make_inputs(1000, 3000, 128, 128, seed=5)
    ref = oracle.attention_f64_numpy(Q, K, V)
    with sdpa.Context(precision="bf16x3", q_batch=384) as ctx:
        ctx.load_kv_host_full(K, V)
        got = ctx.attention_host(Q)
        assert ctx.last_timings()["fused_launches"] == 3
        np.testing.assert_allclose(got, ref, rtol=0, atol=X3_ATOL)
        Qd, Kd, Vd = (torch.from_numpy(a).cuda() for a in (Q, K, V))
        out = torch.zeros(1000, 128, dtype=torch.float64, device="cuda")
        for _ in range(2):
            ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [3000], 128, 128, [Qd.data_ptr()], out.data_ptr(), 1000, blocking=False)
        ctx.synchronize()
        np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=X3_ATOL)


@pytest.mark.parametrize("prec,d", [("bf16", 128), ("bf16x3", 128), ("bf16", 64)])
def test_queued_passes_repair_a_fired_guard_at_synchronize(sdpa, oracle, prec, d):
    """Queued passes of a single-GPU context do not carry the exact twin in the stream: sdpa_synchronize reads the guard ring
    and re-runs (exact variant) the passes whose guard fired.  Three queued passes, the middle one with overflowing data."""
    import torch
    m = 300
    good = oracle.make_inputs(m, 1500, d, d, seed=3)
    bad = oracle.make_inputs(m, 1500, d, d, seed=11)
    bad = (bad[0] * 3.0, np.concatenate([bad[1][:128] * 0.01, bad[1][128:] * 30.0]), bad[2])
    if prec == "bf16":
        good = tuple(oracle.bf16_round(a).astype(np.float64) for a in good)
        bad = tuple(oracle.bf16_round(a).astype(np.float64) for a in bad)
    cases = [good, bad, good]
    with sdpa.Context(precision=prec) as ctx:
        dev = [[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (K, V, Q)] for (Q, K, V) in cases]
        outs = [torch.zeros(m, d, dtype=torch.float64, device="cuda") for _ in cases]
        for rep in range(2):
            for (Kd, Vd, Qd), out in zip(dev, outs):
                ctx.attention_device_full([Kd.data_ptr()], [Vd.data_ptr()], [Kd.shape[0]], d, d, [Qd.data_ptr()], out.data_ptr(), m, blocking=False)
        ctx.synchronize()
        tols = (2e-3, 2e-2, 2e-3) if prec == "bf16" else (1e-5, 2e-2, 1e-5)   # the overflowing case has scores of magnitude ~1000
        for (Q, K, V), out, tol in zip(cases, outs, tols):
            got = out.cpu().numpy()
            assert np.isfinite(got).all()
            np.testing.assert_allclose(got, oracle.attention_f64_numpy(Q, K, V), rtol=0, atol=tol)


def test_queued_passes_longer_than_the_guard_ring(sdpa, oracle):
    """More queued launches than the guard ring has words (4096): the pending passes are resolved before their words are reused;
    a pass with overflowing data early in the queue is still repaired."""
    import torch
    d, m = 64, 64
    good = oracle.make_inputs(m, 256, d, d, seed=5)
    bad = oracle.make_inputs(m, 256, d, d, seed=6)
    bad = (bad[0] * 3.0, np.concatenate([bad[1][:128] * 0.01, bad[1][128:] * 30.0]), bad[2])
    good, bad = (tuple(oracle.bf16_round(a).astype(np.float64) for a in t) for t in (good, bad))
    with sdpa.Context(precision="bf16") as ctx:
        g = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (good[1], good[2], good[0])]
        b = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (bad[1], bad[2], bad[0])]
        out_b = torch.zeros(m, d, dtype=torch.float64, device="cuda")
        out_g = torch.zeros(m, d, dtype=torch.float64, device="cuda")
        ctx.attention_device_full([b[0].data_ptr()], [b[1].data_ptr()], [256], d, d, [b[2].data_ptr()], out_b.data_ptr(), m, blocking=False)
        for _ in range(4500):
            ctx.attention_device_full([g[0].data_ptr()], [g[1].data_ptr()], [256], d, d, [g[2].data_ptr()], out_g.data_ptr(), m, blocking=False)
        ctx.synchronize()
        np.testing.assert_allclose(out_b.cpu().numpy(), oracle.attention_f64_numpy(*bad), rtol=0, atol=2e-2)
        np.testing.assert_allclose(out_g.cpu().numpy(), oracle.attention_f64_numpy(*good), rtol=0, atol=2e-3)
