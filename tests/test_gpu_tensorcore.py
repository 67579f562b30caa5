"""Parity of the tensor-core (tcgen05/TMA) precisions against the CPU oracle.
  fp16x3 : hi/lo-split fp16 operands, 3 MMAs, fp32 accumulate  -> held to the same <= 1e-3 gate
  fp16   : single-pass fp16 operands, fp32 accumulate          -> tolerance stated per test
           (fp16 operands cost ~5e-4 on O(1) outputs; it cannot meet 1e-3
            on the +-100 outputs of the Xavier-initialised regime A)."""
import os

import numpy as np
import pytest
import torch

from oracle import pfnl_ref as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.fixture(scope="module")
def tc_engines(built_lib):
    from pfnl_b200 import Engine
    out = {}
    for prec in ("fp16x3", "fp16", "fp16x3_nltc"):
        for reg in "AB":
            out[(prec, reg)] = Engine(R.make_weights(reg), device=0, precision=prec, graphs=False)
    return out


def _pfrb_ref(fr, W, blk):
    P = "nlvsr/"
    f64 = fr.astype(np.float64)
    k = lambda s: W[P + s].astype(np.float64)
    inp1 = [R.conv2d_same(f64[t:t + 1], k(f"conv1_{blk}/kernel"), k(f"conv1_{blk}/bias"), act=True) for t in range(7)]
    base = R.conv2d_same(np.concatenate(inp1, -1), k(f"conv10_{blk}/kernel"), k(f"conv10_{blk}/bias"), act=True)
    return np.concatenate([f64[t:t + 1] + R.conv2d_same(np.concatenate([base, inp1[t]], -1), k(f"conv2_{blk}/kernel"),
                                                        k(f"conv2_{blk}/bias"), act=True) for t in range(7)], 0)


@pytest.mark.parametrize("prec,tol", [("fp16x3", 2e-5), ("fp16", 2e-2)])
@pytest.mark.parametrize("shape", [(1, 32, 32), (1, 12, 20), (2, 6, 10), (1, 34, 18)])
def test_single_pfrb(tc_engines, prec, tol, shape):
    """One Progressive Fusion Residual Block; ragged shapes exercise TMA zero fill on every border."""
    n, h, w = shape
    W = R.make_weights("A")
    rng = np.random.default_rng(h * 10 + w)
    fr = rng.standard_normal((n * 7, h, w, 64)).astype(np.float32)
    out = tc_engines[(prec, "A")].pfrb(5, cu(fr), n, h, w).cpu().numpy()
    ref = _pfrb_ref(fr[:7], W, 5)
    err = np.abs(out[:7] - ref)
    print(f"{prec} {shape}: pfrb max-abs {err.max():.3e} (|ref|max {np.abs(ref).max():.2f})")
    assert err.max() <= tol * max(1.0, np.abs(ref).max())
    # borders separately
    assert err[:, [0, -1]].max() <= tol * max(1.0, np.abs(ref).max())
    assert err[:, :, [0, -1]].max() <= tol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("regime", ["A", "B"])
@pytest.mark.parametrize("shape", [(1, 8, 8), (2, 6, 10)])
def test_forward_small_fp16x3(tc_engines, regime, shape):
    n, h, w = shape
    z = np.load(os.path.join(GOLD, f"forward_{regime}_{n}x{h}x{w}.npz"))
    y = tc_engines[("fp16x3", regime)].forward(cu(z["x"])).cpu().numpy()
    err = np.abs(y - z["y64"]).max()
    print(f"fp16x3 regime {regime} {shape}: max-abs {err:.3e}")
    assert err <= 1e-3


@pytest.mark.parametrize("regime", ["A", "B"])
def test_forward_pr1_gate_fp16x3(tc_engines, regime):
    """BASELINE config 1 (1 clip x 7 x 32x32) on the tensor-core path, same 1e-3 gate."""
    x = R.make_input(1, 32, 32)
    y = tc_engines[("fp16x3", regime)].forward(cu(x)).cpu().numpy()
    ref64 = np.load(os.path.join(GOLD, f"forward_{regime}_1x32x32.npz"))["y64"]
    err = np.abs(y - ref64).max()
    print(f"fp16x3 regime {regime}: |out|max={np.abs(ref64).max():.2f} max-abs vs fp64 oracle {err:.3e}")
    assert err <= 1e-3


def test_forward_fp16_single_pass_tolerance(tc_engines):
    """Single-pass fp16 operands: reported, with its own stated bound (5e-3 abs on the O(1)
    outputs of regime B); regime A (outputs ~ +-100) is reported as a relative error."""
    x = R.make_input(1, 32, 32)
    yB = tc_engines[("fp16", "B")].forward(cu(x)).cpu().numpy()
    refB = np.load(os.path.join(GOLD, "forward_B_1x32x32.npz"))["y64"]
    eB = np.abs(yB - refB).max()
    yA = tc_engines[("fp16", "A")].forward(cu(x)).cpu().numpy()
    refA = np.load(os.path.join(GOLD, "forward_A_1x32x32.npz"))["y64"]
    eA = np.abs(yA - refA).max()
    print(f"fp16 single pass: regime B max-abs {eB:.3e}; regime A max-abs {eA:.3e} (rel {eA / np.abs(refA).max():.3e})")
    assert eB <= 5e-3
    assert eA <= 2e-2 * np.abs(refA).max()


def test_forward_fp16x3_with_tensor_core_nonlocal(tc_engines):
    """BASELINE configs[1] as worded ("fp16 tensor-core non-local") with fp32-grade convs: the fp16-operand
    non-local block costs ~5e-4 on its O(1) output, so the 1e-3 gate holds for trained-like weights
    (regime B); with Xavier weights the stack amplifies that error ~100x - reported, bounded relatively."""
    x = R.make_input(1, 32, 32)
    yB = tc_engines[("fp16x3_nltc", "B")].forward(cu(x)).cpu().numpy()
    refB = np.load(os.path.join(GOLD, "forward_B_1x32x32.npz"))["y64"]
    eB = np.abs(yB - refB).max()
    yA = tc_engines[("fp16x3_nltc", "A")].forward(cu(x)).cpu().numpy()
    refA = np.load(os.path.join(GOLD, "forward_A_1x32x32.npz"))["y64"]
    eA = np.abs(yA - refA).max()
    print(f"fp16x3 convs + tcgen05 non-local: regime B max-abs {eB:.3e}; regime A max-abs {eA:.3e} "
          f"(rel {eA / np.abs(refA).max():.3e})")
    assert eB <= 1e-3
    assert eA <= 5e-3 * np.abs(refA).max()
    # the conv stack is bit-identical to plain fp16x3: the two modes differ only through the non-local output
    y3 = tc_engines[("fp16x3", "B")].forward(cu(x)).cpu().numpy()
    assert 0 < np.abs(yB - y3).max() <= 2e-3


@pytest.mark.parametrize("hh,ww,n", [(16, 16, 2), (8, 8, 3), (10, 6, 2), (13, 10, 1), (32, 32, 1), (64, 64, 1), (1, 1, 1)])
def test_nonlocal_tcgen05_vs_oracle(tc_engines, hh, ww, n):
    """tcgen05 non-local block (fp16 operands, fp32 accumulate, online softmax) vs the fp64 oracle:
    L = 256, 64, 60 (ragged), 130 (ragged), 1024, 4096 and the degenerate L = 1.
    Tolerance 4e-3 abs on |y| <~ 2.5 (fp16 rounding of X perturbs the logits by ~1e-3 relative)."""
    W = R.make_weights("B")
    P = "nlvsr/nlblock_0/"
    rng = np.random.default_rng(hh * 100 + ww)
    t = rng.random((n, hh, ww, 84), dtype=np.float32)
    ref = R.nonlocal_block(t.astype(np.float64), *(W[P + s].astype(np.float64) for s in
                                                   ("g/g/kernel", "g/g/bias", "w/w/kernel", "w/w/bias")), stable=True)
    out = tc_engines[("fp16", "B")].nonlocal_block(cu(t.reshape(n, hh * ww, 84))).cpu().numpy().reshape(n, hh, ww, 84)
    err = np.abs(out - ref).max()
    print(f"tcgen05 non-local L={hh * ww}: max-abs {err:.3e} (|ref|max {np.abs(ref).max():.2f})")
    assert np.isfinite(out).all()
    assert err <= 4e-3


@pytest.mark.parametrize("hh,ww,n", [(16, 16, 2), (10, 6, 2), (13, 10, 1), (32, 32, 1), (64, 64, 1), (1, 1, 1)])
def test_nonlocal_tcgen05_split_operands_vs_oracle(tc_engines, hh, ww, n):
    """The hi/lo-split tcgen05 non-local block of the fp16x3 mode (fp32-grade logits and values, fp16 P):
    an order of magnitude closer to the fp64 oracle than the fp16-operand kernel."""
    L = hh * ww
    rng = np.random.default_rng(L)
    W = R.make_weights("A")
    x = rng.random((n, L, 84), dtype=np.float32)
    got = tc_engines[("fp16x3", "A")].nonlocal_block(cu(x)).cpu().numpy()
    P = "nlvsr/nlblock_0/"
    ref = R.nonlocal_block(x.astype(np.float64).reshape(n, hh, ww, 84), W[P + "g/g/kernel"].astype(np.float64),
                           W[P + "g/g/bias"].astype(np.float64), W[P + "w/w/kernel"].astype(np.float64),
                           W[P + "w/w/bias"].astype(np.float64), stable=True).reshape(n, L, 84)
    err = np.abs(got - ref).max()
    print(f"tcgen05 split non-local L={L}: max-abs {err:.3e} (|ref|max {np.abs(ref).max():.2f})")
    assert err <= 2.5e-4   # fp16 rounding of P (2^-12 relative per element); the fp16-operand kernel is at ~5e-4 to 8e-4


def test_nonlocal_tcgen05_large_logits_and_6480(tc_engines):
    e = tc_engines[("fp16", "B")]
    W = R.make_weights("B")
    P = "nlvsr/nlblock_0/"
    t = np.full((1, 2048, 84), 0.9, np.float32)          # logits 68: naive exp/sum would overflow
    out = e.nonlocal_block(cu(t)).cpu().numpy()
    assert np.isfinite(out).all()
    g = t[0, :1] @ W[P + "g/g/kernel"][0, 0] + W[P + "g/g/bias"]
    zrow = g @ W[P + "w/w/kernel"][0, 0] + W[P + "w/w/bias"]
    np.testing.assert_allclose(out[0], np.broadcast_to(zrow, (2048, 84)), atol=4e-3)
    rng = np.random.default_rng(9)
    t = (rng.random((1, 6480, 84), dtype=np.float32) * 0.5).astype(np.float32)   # 90x72 token grid
    out = e.nonlocal_block(cu(t)).cpu().numpy()
    rows = [0, 1234, 6479]
    x64 = t[0].astype(np.float64)
    s = x64[rows] @ x64.T
    p = np.exp(s - s.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    g = x64 @ W[P + "g/g/kernel"][0, 0].astype(np.float64) + W[P + "g/g/bias"]
    ref = (p @ g) @ W[P + "w/w/kernel"][0, 0].astype(np.float64) + W[P + "w/w/bias"]
    np.testing.assert_allclose(out[0, rows], ref, atol=4e-3)


def test_forward_batch16_tc_matches_per_clip(tc_engines):
    x = R.make_input(16, 32, 32, seed=99)
    e = tc_engines[("fp16x3", "B")]
    yb = e.forward(cu(x)).cpu().numpy()
    for i in (0, 9, 15):
        assert np.array_equal(yb[i:i + 1], e.forward(cu(x[i:i + 1])).cpu().numpy())
    ref = R.pfnl_forward(x[:2], R.make_weights("B"), backend="torch")
    assert np.abs(yb[:2] - ref).max() <= 1e-3


@pytest.mark.parametrize("prec,tol", [("fp16x3", 1e-3), ("fp16", 5e-3)])
def test_forward_more_units_than_sms(tc_engines, prec, tol):
    """2 clips x 7 x 128x96: 192 work units > 148 SMs, so persistent CTAs own several units per phase
    (conv1 -> conv10 and partial-sum -> conv2 run phase-ordered on the CTA's own units)."""
    x = R.make_input(2, 128, 96, seed=11)
    y = tc_engines[(prec, "B")].forward(cu(x)).cpu().numpy()
    ref = R.pfnl_forward(x, R.make_weights("B"), backend="torch")
    err = np.abs(y - ref).max()
    print(f"{prec} 2x128x96: max-abs {err:.3e}")
    assert err <= tol


def test_forward_128_fp16x3(tc_engines):
    x = R.make_input(1, 128, 128, seed=5)
    y = tc_engines[("fp16x3", "B")].forward(cu(x)).cpu().numpy()
    ref = R.pfnl_forward(x, R.make_weights("B"), backend="torch")
    assert np.abs(y - ref).max() <= 1e-3


# ---- the PFRB stack as one persistent dataflow kernel (csrc/pfrb_flow.cu) -----------------------------------
@pytest.mark.parametrize("prec", ["fp16x3", "fp16"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (16, 32, 32), (2, 6, 10), (1, 34, 18), (3, 48, 40), (1, 128, 128),
                                   (1, 144, 180), (40, 32, 32)])
def test_flow_matches_phase_kernels(tc_engines, prec, shape):
    """Same per-tile arithmetic, different scheduling: the dataflow kernel (148 persistent CTAs with one role each,
    ordered by arrival counters) must reproduce the two-launches-per-block kernels bit for bit - any missed
    dependency (a tile read before its producer finished) shows up as a difference.  Shapes: fewer tiles than
    CTAs, the bench shape, ragged borders, several tile rows/columns, config 4, a Vid4 frame and 40 clips (207 and
    320 units: the role split of the larger problems, planes that outgrow the L2)."""
    n, h, w = shape
    eng = tc_engines[(prec, "A")]
    x = cu(R.make_input(n, h, w))
    try:
        eng.set_flow(False)
        y_phase = eng.forward(x).clone()
        eng.set_flow(True)
        l0 = eng.launches
        y_flow = eng.forward(x).clone()
        n_flow = eng.launches - l0
        y_flow2 = eng.forward(x).clone()  # the counters were cleared by the first launch
    finally:
        eng.set_flow(True)
    torch.cuda.synchronize()
    # nl_prep, nl_tc, (nl_merge when the keys are split over CTAs,) nl_linear, conv0, pfrb_flow, convmerge1, tail
    assert n_flow in (7, 8), n_flow
    assert torch.equal(y_flow, y_phase), float((y_flow - y_phase).abs().max())
    assert torch.equal(y_flow2, y_phase)


def test_flow_repeated_and_shape_changes(tc_engines):
    """The dependency counters are handle-owned and self-clearing: alternate shapes and repeat."""
    eng = tc_engines[("fp16x3", "B")]
    shapes = [(2, 16, 16), (1, 32, 32), (2, 16, 16), (4, 8, 24), (1, 32, 32)]
    first = {}
    for n, h, w in shapes * 2:
        y = eng.forward(cu(R.make_input(n, h, w)))
        torch.cuda.synchronize()
        if (n, h, w) in first:
            assert torch.equal(y, first[(n, h, w)])
        else:
            first[(n, h, w)] = y.clone()
    ref = R.pfnl_forward(R.make_input(1, 32, 32), R.make_weights("B"), dtype=np.float64)
    assert np.abs(first[(1, 32, 32)].cpu().numpy() - ref).max() <= 1e-3


# ---- stage-level parity of the two tcgen05 kernels that were only covered end to end (a5, a8) -------
@pytest.mark.parametrize("prec,tol", [("fp16x3", 2e-5), ("fp16", 2e-2)])
@pytest.mark.parametrize("regime", ["A", "B"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 6, 10), (1, 34, 18), (3, 16, 8)])
def test_conv0_stage(tc_engines, prec, tol, regime, shape):
    """conv0 5x5 3->64 + leaky_relu on every frame (pfnl.py:48,61-62) through conv0_tc_kernel (explicit im2col on
    tcgen05) vs the fp64 oracle; ragged shapes put TMA-free zero padding on every border, negative pre-activations
    exercise the 0.2 slope."""
    n, h, w = shape
    W = R.make_weights(regime)  # regime B has non-zero biases
    rng = np.random.default_rng(100 + h * 10 + w)
    inp21 = rng.standard_normal((n, h, w, 21)).astype(np.float32)
    out = tc_engines[(prec, regime)].conv0(cu(inp21)).cpu().numpy().reshape(n, 7, h, w, 64)
    k = W["nlvsr/conv0/kernel"].astype(np.float64)
    b = W["nlvsr/conv0/bias"].astype(np.float64)
    ref = np.stack([R.conv2d_same(inp21[..., 3 * t:3 * t + 3].astype(np.float64), k, b, act=True)
                    for t in range(7)], 1)
    err = np.abs(out - ref)
    scale = max(1.0, np.abs(ref).max())
    print(f"{prec} {regime} {shape}: conv0 max-abs {err.max():.3e} (|ref|max {np.abs(ref).max():.2f})")
    assert (ref < 0).any()
    assert err.max() <= tol * scale
    assert err[:, :, [0, 1, -2, -1]].max() <= tol * scale and err[:, :, :, [0, 1, -2, -1]].max() <= tol * scale


@pytest.mark.parametrize("prec,tol", [("fp16x3", 2e-5), ("fp16", 2e-2)])
@pytest.mark.parametrize("regime", ["A", "B"])
@pytest.mark.parametrize("shape", [(1, 32, 32), (2, 6, 10), (1, 34, 18)])
def test_convmerge1_stage(tc_engines, prec, tol, regime, shape):
    """convmerge1 3x3 448->48 + leaky_relu over the concat of the 7 frames (pfnl.py:52,73-74) through the 7-phase
    tcgen05 launch (N = 48, fp32 partial sums accumulated in place) vs the fp64 oracle.  Regime B has non-zero
    biases."""
    n, h, w = shape
    W = R.make_weights(regime)
    rng = np.random.default_rng(200 + h * 10 + w)
    fr = rng.standard_normal((n * 7, h, w, 64)).astype(np.float32)
    out = tc_engines[(prec, regime)].convmerge1(cu(fr), n, h, w).cpu().numpy()
    cat = np.concatenate([fr.reshape(n, 7, h, w, 64)[:, t] for t in range(7)], -1).astype(np.float64)
    ref = R.conv2d_same(cat, W["nlvsr/convmerge1/kernel"].astype(np.float64),
                        W["nlvsr/convmerge1/bias"].astype(np.float64), act=True)
    err = np.abs(out - ref)
    scale = max(1.0, np.abs(ref).max())
    print(f"{prec} {regime} {shape}: convmerge1 max-abs {err.max():.3e} (|ref|max {np.abs(ref).max():.2f})")
    assert out.shape == (n, h, w, 48)
    assert err.max() <= tol * scale
    assert err[:, [0, -1]].max() <= tol * scale and err[:, :, [0, -1]].max() <= tol * scale


# ---- the 1e-3 gate of the headline precision, with margin (VERDICT r1 "what's weak" 1-2) -------------------------
GATE = 1e-3
MARGIN_TOL = 0.7 * GATE   # every check below must leave >= 30 % of the gate unused


def _oracle_clip_fp64(args):
    regime, clip = args
    import numpy as _np
    from oracle import pfnl_ref as _R
    return _R.pfnl_forward(clip[None], _R.make_weights(regime), dtype=_np.float64, backend="numpy")[0]


def oracle_fp64_parallel(x, regime):
    """fp64 numpy oracle, one process per clip (single-threaded BLAS each): 128 clips in tens of seconds."""
    import multiprocessing as mp
    keys = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")
    old = {k: os.environ.get(k) for k in keys}
    os.environ.update({k: "1" for k in keys})
    try:
        workers = max(1, min(32, len(os.sched_getaffinity(0)), x.shape[0]))
        with mp.get_context("spawn").Pool(workers) as pool:
            out = pool.map(_oracle_clip_fp64, [(regime, x[i]) for i in range(x.shape[0])], chunksize=1)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return np.stack(out)


@pytest.mark.parametrize("regime", ["A", "B"])
def test_gate_fp16x3_vs_fp32_and_fp64_oracle(tc_engines, regime):
    """BASELINE config 1: the reference computes in fp32, so the headline precision is held against the fp32 numpy
    oracle as well as against the fp64 evaluation, both with 30 % margin."""
    x = R.make_input(1, 32, 32)
    W = R.make_weights(regime)
    y = tc_engines[("fp16x3", regime)].forward(cu(x)).cpu().numpy()
    ref64 = np.load(os.path.join(GOLD, f"forward_{regime}_1x32x32.npz"))["y64"]
    ref32 = R.pfnl_forward(x, W, dtype=np.float32, backend="numpy")
    e64, e32 = np.abs(y - ref64).max(), np.abs(y - ref32).max()
    print(f"fp16x3 regime {regime}: vs fp64 {e64:.3e}, vs fp32 oracle {e32:.3e} (oracle fp32 vs fp64 {np.abs(ref32 - ref64).max():.3e})")
    assert e64 <= MARGIN_TOL and e32 <= MARGIN_TOL


def test_gate_fp16x3_config3_128_clips(tc_engines):
    """BASELINE config 3's batch: 128 distinct clips (16 per GPU x 8) in regime A - the max over all of them, not
    one seed.  Run as 8 batches of 16 (the per-GPU shape); also bit-equal to one batch of 128."""
    x = R.make_input(128, 32, 32, seed=4242)
    eng = tc_engines[("fp16x3", "A")]
    ys = torch.cat([eng.forward(cu(x[i:i + 16])) for i in range(0, 128, 16)])
    y_all = eng.forward(cu(x))
    torch.cuda.synchronize()
    assert torch.equal(ys, y_all)
    ref = oracle_fp64_parallel(x, "A")
    err = np.abs(ys.cpu().numpy() - ref).reshape(128, -1).max(1)
    print(f"fp16x3 regime A, 128 clips: max-abs per clip min {err.min():.3e} median {np.median(err):.3e} max {err.max():.3e}")
    assert err.max() <= MARGIN_TOL


def test_gate_fp16x3_config4_128x128_regime_A(tc_engines):
    """BASELINE config 4 (1 clip x 7 x 128x128, L = 4096 tokens) in regime A against the fp64 oracle: long
    accumulation chains in the non-local block (256 per row), key-split partials, 128 spatial tiles."""
    x = R.make_input(1, 128, 128, seed=77)
    y = tc_engines[("fp16x3", "A")].forward(cu(x)).cpu().numpy()
    ref = R.pfnl_forward(x, R.make_weights("A"), dtype=np.float64, backend="numpy")
    err = np.abs(y - ref).max()
    print(f"fp16x3 regime A 1x128x128: |out|max {np.abs(ref).max():.1f} max-abs vs fp64 {err:.3e}")
    assert err <= MARGIN_TOL
