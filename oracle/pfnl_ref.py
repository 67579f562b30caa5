"""CPU oracle for the PFNL 4x multi-frame forward hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (psychopa4/PFNL) ships no tests, golden vectors or
checkpoints for this path, and its arithmetic lives in an un-vendored dependency
(TensorFlow 1.12.0, README.md:23) that cannot be installed in this image (Python 3.12, no
network).  This file is therefore a *restatement* of the reference graph, following
  model/pfnl.py:39-80   (PFNL.forward)
  utils.py:18-71        (NonLocalBlock, nltype=1 "gaussian", sub_sample=1)
  modules/ps.py:3-15    (_PS periodic shuffle; index-identical to tf.depth_to_space)
  model/pfnl.py:90,139-141 (per-clip MSE -> PSNR)
plus the documented TF-1.12 op semantics (NHWC convs with HWIO kernels and 'same' zero
padding, DCR depth_to_space/space_to_depth, legacy ResizeBicubic with align_corners=False,
leaky_relu alpha=0.2).  The pins this project creates in place of reference goldens are the
known-answer tests in tests/test_oracle_kat.py and the frozen fixtures in tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path (pfnl_b200/) never does: it fails loudly when the
CUDA library is missing.

Two interchangeable back-ends compute the same graph:
  backend="numpy" - independent im2col+matmul restatement (the pinned checker, fp32 or fp64)
  backend="torch" - torch-CPU (oneDNN, multi-threaded) convs; used as the timed CPU baseline
                    and cross-checked against the numpy back-end in tests/.
"""
from __future__ import annotations

import numpy as np

NUM_FRAMES = 7      # model/pfnl.py:22
SCALE = 4           # model/pfnl.py:23
NUM_BLOCK = 20      # model/pfnl.py:43
MF = 64             # model/pfnl.py:40
LRELU_ALPHA = 0.2   # tf.nn.leaky_relu default (model/pfnl.py:42)
NL_CH = 3 * NUM_FRAMES * 4  # 84 (model/pfnl.py:58)

# ---------------------------------------------------------------------------------------
# Variable inventory (scope 'nlvsr', model/pfnl.py:47-53; utils.py:23-26,66-67)
# ---------------------------------------------------------------------------------------

def variable_shapes():
    """Ordered {tf_variable_name: shape}; kernels are HWIO.  3,003,156 parameters."""
    s = {}
    s["nlvsr/nlblock_0/g/g/kernel"] = (1, 1, NL_CH, NL_CH)
    s["nlvsr/nlblock_0/g/g/bias"] = (NL_CH,)
    s["nlvsr/nlblock_0/w/w/kernel"] = (1, 1, NL_CH, NL_CH)
    s["nlvsr/nlblock_0/w/w/bias"] = (NL_CH,)
    s["nlvsr/conv0/kernel"] = (5, 5, 3, MF)
    s["nlvsr/conv0/bias"] = (MF,)
    for i in range(NUM_BLOCK):
        s[f"nlvsr/conv1_{i}/kernel"] = (3, 3, MF, MF)
        s[f"nlvsr/conv1_{i}/bias"] = (MF,)
    for i in range(NUM_BLOCK):
        s[f"nlvsr/conv10_{i}/kernel"] = (1, 1, MF * NUM_FRAMES, MF)
        s[f"nlvsr/conv10_{i}/bias"] = (MF,)
    for i in range(NUM_BLOCK):
        s[f"nlvsr/conv2_{i}/kernel"] = (3, 3, 2 * MF, MF)
        s[f"nlvsr/conv2_{i}/bias"] = (MF,)
    s["nlvsr/convmerge1/kernel"] = (3, 3, MF * NUM_FRAMES, 48)
    s["nlvsr/convmerge1/bias"] = (48,)
    s["nlvsr/convmerge2/kernel"] = (3, 3, 12, 12)
    s["nlvsr/convmerge2/bias"] = (12,)
    return s


def num_params():
    return int(sum(int(np.prod(v)) for v in variable_shapes().values()))


def make_weights(regime="A", seed=4321):
    """Deterministic synthetic weights keyed by TF variable name (fp32, HWIO).

    regime "A" = what the reference runs with no checkpoint: Xavier/Glorot-uniform kernels
                 U(+-sqrt(6/(kh*kw*Cin + kh*kw*Cout))) (model/pfnl.py:45; TF default for the
                 NL 1x1 convs, utils.py:26,67) and zero biases.
    regime "B" = "trained-like": regime A with conv2_i kernels x0.1 and biases U(+-0.05).
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in variable_shapes().items():
        if name.endswith("kernel"):
            kh, kw, ci, co = shp
            lim = np.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
            w[name] = rng.uniform(-lim, lim, size=shp).astype(np.float32)
        else:
            w[name] = np.zeros(shp, np.float32)
    if regime == "B":
        rng2 = np.random.default_rng(seed + 1)
        for name in w:
            if name.endswith("bias"):
                w[name] = rng2.uniform(-0.05, 0.05, size=w[name].shape).astype(np.float32)
            elif "/conv2_" in name:
                w[name] = (w[name] * np.float32(0.1)).astype(np.float32)
    elif regime != "A":
        raise ValueError("regime must be 'A' or 'B'")
    return w


def make_input(n, h, w, seed=1234):
    """LR clips [n,7,h,w,3] fp32 U[0,1) (image range after /255., model/pfnl.py:209,270)."""
    rng = np.random.default_rng(seed)
    return rng.random((n, NUM_FRAMES, h, w, 3), dtype=np.float32)


def make_target(n, h, w, seed=1235):
    rng = np.random.default_rng(seed)
    return rng.random((n, 1, h * SCALE, w * SCALE, 3), dtype=np.float32)


# ---------------------------------------------------------------------------------------
# Op restatements (numpy)
# ---------------------------------------------------------------------------------------

def leaky_relu(x, alpha=LRELU_ALPHA):
    """tf.nn.leaky_relu: max(alpha*x, x)."""
    return np.maximum(x * x.dtype.type(alpha), x)


def conv2d_same(x, kernel, bias, act=False):
    """tf.layers.Conv2D(..., strides=1, padding='same') on NHWC with an HWIO kernel
    (model/pfnl.py:48-53): cross-correlation, zero pad (k-1)/2 per side, + bias, optional
    leaky_relu(0.2).  im2col + matmul in x.dtype."""
    n, h, w, ci = x.shape
    kh, kw, kci, co = kernel.shape
    assert kci == ci, (kci, ci)
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    dt = x.dtype
    xp = np.zeros((n, h + kh - 1, w + kw - 1, ci), dt)
    xp[:, ph:ph + h, pw:pw + w, :] = x
    win = np.lib.stride_tricks.sliding_window_view(xp, (kh, kw), axis=(1, 2))  # n,h,w,ci,kh,kw
    cols = np.ascontiguousarray(win.transpose(0, 1, 2, 4, 5, 3)).reshape(n * h * w, kh * kw * ci)
    y = cols @ kernel.astype(dt).reshape(kh * kw * ci, co)
    y = y + bias.astype(dt)
    y = y.reshape(n, h, w, co)
    return leaky_relu(y) if act else y


def space_to_depth(x, b):
    """tf.space_to_depth NHWC: out[n,h,w,(dy*b+dx)*C+c] = in[n,h*b+dy,w*b+dx,c]."""
    n, h, w, c = x.shape
    assert h % b == 0 and w % b == 0
    y = x.reshape(n, h // b, b, w // b, b, c).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(y).reshape(n, h // b, w // b, b * b * c)


def depth_to_space(x, b):
    """tf.depth_to_space NHWC (DCR): out[n,h*b+dy,w*b+dx,c] = in[n,h,w,(dy*b+dx)*Co+c]."""
    n, h, w, c = x.shape
    assert c % (b * b) == 0
    co = c // (b * b)
    y = x.reshape(n, h, w, b, b, co).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(y).reshape(n, h * b, w * b, co)


def periodic_shuffle(x, r, n_out_channel):
    """Restatement of modules/ps.py:_PS (ps.py:9-12): split channels into r groups,
    concatenate the groups along W, reshape to (r*H, r*W, n_out)."""
    n, a, b, c = x.shape
    assert c == r * r * n_out_channel
    xs = np.split(x, r, axis=3)
    xr = np.concatenate(xs, axis=2)
    return xr.reshape(n, r * a, r * b, n_out_channel)


_BICUBIC_A = -0.75
_TABLE = 1024


def _bicubic_table():
    """TF ResizeBicubic coefficient table (1024 entries, Keys cubic A=-0.75), fp32."""
    t = np.zeros((_TABLE + 1) * 2, np.float32)
    a = np.float32(_BICUBIC_A)
    for i in range(_TABLE + 1):
        x = np.float32(i) / np.float32(_TABLE)
        t[2 * i] = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + np.float32(1)
        x = x + np.float32(1)
        t[2 * i + 1] = ((a * x - np.float32(5) * a) * x + np.float32(8) * a) * x - np.float32(4) * a
    return t


def bicubic_weights_indices(out_size, in_size):
    """Per output index: 4 clamped input indices and 4 weights, as TF-1.12 ResizeBicubic
    (align_corners=False, no half-pixel centres) computes them."""
    tab = _bicubic_table()
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.zeros((out_size, 4), np.int64)
    wts = np.zeros((out_size, 4), np.float32)
    for o in range(out_size):
        in_f = np.float32(o) * scale
        loc = int(np.floor(in_f))
        delta = in_f - np.float32(loc)
        off = int(np.rint(delta * np.float32(_TABLE)))
        wts[o] = (tab[off * 2 + 1], tab[off * 2], tab[(_TABLE - off) * 2], tab[(_TABLE - off) * 2 + 1])
        idx[o] = np.clip([loc - 1, loc, loc + 1, loc + 2], 0, in_size - 1)
    return idx, wts


def resize_bicubic(img, out_h, out_w):
    """tf.image.resize_images(img,[out_h,out_w],method=2) (model/pfnl.py:63), NHWC.
    Separable: 4 taps along x first, then 4 taps along y, left-to-right sums."""
    n, h, w, c = img.shape
    dt = img.dtype
    ix, wx = bicubic_weights_indices(out_w, w)
    iy, wy = bicubic_weights_indices(out_h, h)
    wx = wx.astype(dt)
    wy = wy.astype(dt)
    # along x
    tmp = (img[:, :, ix[:, 0], :] * wx[None, None, :, 0, None]
           + img[:, :, ix[:, 1], :] * wx[None, None, :, 1, None]
           + img[:, :, ix[:, 2], :] * wx[None, None, :, 2, None]
           + img[:, :, ix[:, 3], :] * wx[None, None, :, 3, None])
    out = (tmp[:, iy[:, 0]] * wy[None, :, 0, None, None]
           + tmp[:, iy[:, 1]] * wy[None, :, 1, None, None]
           + tmp[:, iy[:, 2]] * wy[None, :, 2, None, None]
           + tmp[:, iy[:, 3]] * wy[None, :, 3, None, None])
    return out.astype(dt)


def nonlocal_block(x, wg, bg, ww, bw, stable=False):
    """utils.py:18-71 with nltype=1, sub_sample=1.  x: [N,h,w,C] (C=84)."""
    n, h, w, c = x.shape
    dt = x.dtype
    g = conv2d_same(x, wg, bg)                       # utils.py:26
    g_x = g.reshape(n, h * w, c)                     # utils.py:44
    theta_x = x.reshape(n, h * w, c)                 # utils.py:42,45 (theta = input_x)
    phi_x = x.reshape(n, h * w, c).transpose(0, 2, 1)  # utils.py:34,49-50 (phi = input_x)
    f = theta_x @ phi_x                              # utils.py:53
    if stable:
        f = f - f.max(axis=-1, keepdims=True)
    f = np.exp(f)                                    # utils.py:57
    f_softmax = f / f.sum(axis=-1, keepdims=True)    # utils.py:58
    y = f_softmax @ g_x                              # utils.py:64
    y = y.reshape(n, h, w, c)                        # utils.py:65
    return conv2d_same(y.astype(dt), ww, bw)         # utils.py:67,70


def tokens(x):
    """[N,7,H,W,3] -> token matrix [N,H/2,W/2,84] (model/pfnl.py:55-57)."""
    inp0 = np.concatenate([x[:, i] for i in range(x.shape[1])], axis=-1)
    return space_to_depth(inp0, 2)


# ---------------------------------------------------------------------------------------
# Whole forward
# ---------------------------------------------------------------------------------------

def pfnl_forward(x, weights, dtype=np.float32, backend="numpy", stable_softmax=False,
                 return_intermediates=False):
    """PFNL.forward (model/pfnl.py:39-80): x [N,7,H,W,3] -> [N,1,4H,4W,3]."""
    if backend == "torch":
        return _pfnl_forward_torch(x, weights, dtype, stable_softmax)
    assert backend == "numpy"
    x = np.asarray(x, dtype=dtype)
    W = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    n, f1, h, w, c = x.shape
    assert f1 == NUM_FRAMES and c == 3 and h % 2 == 0 and w % 2 == 0
    P = "nlvsr/"
    inter = {}

    inp0 = np.concatenate([x[:, i] for i in range(f1)], axis=-1)           # :55-56
    inp1 = space_to_depth(inp0, 2)                                         # :57
    inter["tokens"] = inp1
    inp1 = nonlocal_block(inp1, W[P + "nlblock_0/g/g/kernel"], W[P + "nlblock_0/g/g/bias"],
                          W[P + "nlblock_0/w/w/kernel"], W[P + "nlblock_0/w/w/bias"],
                          stable=stable_softmax)                           # :58
    inter["nl_out"] = inp1
    inp1 = depth_to_space(inp1, 2)                                         # :59
    inp0 = inp0 + inp1                                                     # :60
    inter["inp0_nl"] = inp0
    inp0 = np.split(inp0, f1, axis=-1)                                     # :61
    inp0 = [conv2d_same(f, W[P + "conv0/kernel"], W[P + "conv0/bias"], act=True) for f in inp0]  # :62
    inter["conv0"] = np.stack(inp0, 1)
    bic = resize_bicubic(x[:, f1 // 2], h * SCALE, w * SCALE)              # :63

    for i in range(NUM_BLOCK):                                             # :65
        k1, b1 = W[P + f"conv1_{i}/kernel"], W[P + f"conv1_{i}/bias"]
        k10, b10 = W[P + f"conv10_{i}/kernel"], W[P + f"conv10_{i}/bias"]
        k2, b2 = W[P + f"conv2_{i}/kernel"], W[P + f"conv2_{i}/bias"]
        inp1 = [conv2d_same(f, k1, b1, act=True) for f in inp0]            # :66
        base = np.concatenate(inp1, axis=-1)                               # :67
        base = conv2d_same(base, k10, b10, act=True)                       # :68
        inp2 = [np.concatenate([base, f], -1) for f in inp1]               # :69
        inp2 = [conv2d_same(f, k2, b2, act=True) for f in inp2]            # :70
        inp0 = [inp0[j] + inp2[j] for j in range(f1)]                      # :71
        if i == 0:
            inter["block0"] = np.stack(inp0, 1)
    inter["pfrb"] = np.stack(inp0, 1)

    merge = np.concatenate(inp0, axis=-1)                                  # :73
    merge = conv2d_same(merge, W[P + "convmerge1/kernel"], W[P + "convmerge1/bias"], act=True)  # :74
    inter["merge1"] = merge
    large1 = depth_to_space(merge, 2)                                      # :76
    out1 = conv2d_same(large1, W[P + "convmerge2/kernel"], W[P + "convmerge2/bias"])  # :77
    out = depth_to_space(out1, 2)                                          # :78
    res = np.stack([out + bic], axis=1)                                    # :80
    if return_intermediates:
        return res, inter
    return res


def _pfnl_forward_torch(x, weights, dtype, stable_softmax):
    """Same graph on torch-CPU (oneDNN convs, all host threads): the timed CPU baseline."""
    import torch
    import torch.nn.functional as F
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    xt = torch.as_tensor(np.asarray(x), dtype=tdt)
    W = {k: torch.as_tensor(np.asarray(v), dtype=tdt) for k, v in weights.items()}
    P = "nlvsr/"
    n, f1, h, w, c = xt.shape

    def conv(inp_nhwc, kname, act):
        k = W[P + kname + "/kernel"].permute(3, 2, 0, 1).contiguous()      # HWIO -> OIHW
        b = W[P + kname + "/bias"]
        y = F.conv2d(inp_nhwc.permute(0, 3, 1, 2), k, b, padding=(k.shape[2] - 1) // 2)
        y = y.permute(0, 2, 3, 1)
        return torch.maximum(y * LRELU_ALPHA, y) if act else y

    def s2d(t, b):
        nn_, hh, ww_, cc = t.shape
        return t.reshape(nn_, hh // b, b, ww_ // b, b, cc).permute(0, 1, 3, 2, 4, 5).reshape(
            nn_, hh // b, ww_ // b, b * b * cc)

    def d2s(t, b):
        nn_, hh, ww_, cc = t.shape
        co = cc // (b * b)
        return t.reshape(nn_, hh, ww_, b, b, co).permute(0, 1, 3, 2, 4, 5).reshape(nn_, hh * b, ww_ * b, co)

    inp0 = torch.cat([xt[:, i] for i in range(f1)], dim=-1)
    inp1 = s2d(inp0, 2)
    g = conv(inp1, "nlblock_0/g/g", False).reshape(n, -1, NL_CH)
    th = inp1.reshape(n, -1, NL_CH)
    f = th @ th.transpose(1, 2)
    if stable_softmax:
        f = f - f.amax(dim=-1, keepdim=True)
    f = torch.exp(f)
    p = f / f.sum(dim=-1, keepdim=True)
    y = (p @ g).reshape(n, h // 2, w // 2, NL_CH)
    inp1 = conv(y, "nlblock_0/w/w", False)
    inp0 = inp0 + d2s(inp1, 2)
    # frames folded into the batch: weights are shared across the 7 frames (pfnl.py:62,66,70)
    fr = inp0.reshape(n, h, w, f1, 3).permute(0, 3, 1, 2, 4).reshape(n * f1, h, w, 3)
    fr = conv(fr, "conv0", True)
    bic = torch.as_tensor(resize_bicubic(np.asarray(x[:, f1 // 2], dtype=dtype), h * SCALE, w * SCALE), dtype=tdt)
    for i in range(NUM_BLOCK):
        a1 = conv(fr, f"conv1_{i}", True)                                   # [n*7,h,w,64]
        cat = a1.reshape(n, f1, h, w, MF).permute(0, 2, 3, 1, 4).reshape(n, h, w, f1 * MF)
        base = conv(cat, f"conv10_{i}", True)                               # [n,h,w,64]
        baser = base[:, None].expand(n, f1, h, w, MF).reshape(n * f1, h, w, MF)
        a2 = conv(torch.cat([baser, a1], dim=-1), f"conv2_{i}", True)
        fr = fr + a2
    merge = fr.reshape(n, f1, h, w, MF).permute(0, 2, 3, 1, 4).reshape(n, h, w, f1 * MF)
    merge = conv(merge, "convmerge1", True)
    out1 = conv(d2s(merge, 2), "convmerge2", False)
    out = d2s(out1, 2)
    return (out + bic)[:, None].numpy()


def mse_per_clip(sr, hr):
    """eval_mse = mean((SR-H)^2, axis=[2,3,4]) (model/pfnl.py:90) -> [N,1]."""
    d = np.asarray(sr, np.float64) - np.asarray(hr, np.float64)
    return (d * d).mean(axis=(2, 3, 4))


def psnr_from_mse(mse):
    """10*log10(1/mse) (model/pfnl.py:139)."""
    return 10.0 * np.log10(1.0 / np.asarray(mse, np.float64))


def quantise_uint8(sr_frame):
    """round(clip(sr*255,0,255)).astype(uint8) (model/pfnl.py:255-257)."""
    img = np.asarray(sr_frame) * 255.0
    img = np.clip(img, 0, 255)
    return np.round(img, 0).astype(np.uint8)


def window_indices(max_frame, i, num_frames=NUM_FRAMES):
    """Sliding 7-frame window with edge clamping (model/pfnl.py:239-240,297-298)."""
    idx = np.arange(i - num_frames // 2, i + num_frames // 2 + 1)
    return np.clip(idx, 0, max_frame - 1).tolist()
