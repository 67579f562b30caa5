"""Weights for the PFNL forward, keyed by the reference's TF variable names
(scope 'nlvsr', model/pfnl.py:47-53; NonLocalBlock scopes utils.py:23-26,66-67).
Kernels are HWIO fp32, biases [Cout]."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

NUM_BLOCK = _lib.NUM_BLOCK
MF = _lib.MF
NL_CH = _lib.NL_CH
NUM_FRAMES = _lib.NUM_FRAMES


def variable_shapes():
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


def xavier_init(seed=4321):
    """What the reference runs with when no checkpoint is found (base_model.py:231-243 returns
    False and the session keeps its initialiser values): Xavier/Glorot-uniform kernels
    (model/pfnl.py:45; TF default for utils.py:26,67) and zero biases."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in variable_shapes().items():
        if name.endswith("kernel"):
            kh, kw, ci, co = shp
            lim = np.sqrt(6.0 / (kh * kw * ci + kh * kw * co))
            w[name] = rng.uniform(-lim, lim, size=shp).astype(np.float32)
        else:
            w[name] = np.zeros(shp, np.float32)
    return w


def validate(weights):
    shapes = variable_shapes()
    out = {}
    for name, shp in shapes.items():
        if name not in weights:
            raise KeyError(f"missing weight '{name}'")
        a = np.ascontiguousarray(np.asarray(weights[name], dtype=np.float32))
        if tuple(a.shape) != tuple(shp):
            raise ValueError(f"weight '{name}' has shape {a.shape}, expected {shp}")
        out[name] = a
    return out


def load_npz(path):
    with np.load(path) as z:
        return validate({k: z[k] for k in z.files})


def save_npz(path, weights):
    np.savez(path, **validate(weights))


def to_struct(weights):
    """-> (PfnlWeights, keepalive list of numpy arrays)."""
    w = validate(weights)
    keep = []

    def p(name):
        a = w[name]
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_float))

    s = _lib.PfnlWeights()
    P = "nlvsr/"
    s.nl_g_kernel, s.nl_g_bias = p(P + "nlblock_0/g/g/kernel"), p(P + "nlblock_0/g/g/bias")
    s.nl_w_kernel, s.nl_w_bias = p(P + "nlblock_0/w/w/kernel"), p(P + "nlblock_0/w/w/bias")
    s.conv0_kernel, s.conv0_bias = p(P + "conv0/kernel"), p(P + "conv0/bias")
    for i in range(NUM_BLOCK):
        s.conv1_kernel[i], s.conv1_bias[i] = p(P + f"conv1_{i}/kernel"), p(P + f"conv1_{i}/bias")
        s.conv10_kernel[i], s.conv10_bias[i] = p(P + f"conv10_{i}/kernel"), p(P + f"conv10_{i}/bias")
        s.conv2_kernel[i], s.conv2_bias[i] = p(P + f"conv2_{i}/kernel"), p(P + f"conv2_{i}/bias")
    s.merge1_kernel, s.merge1_bias = p(P + "convmerge1/kernel"), p(P + "convmerge1/bias")
    s.merge2_kernel, s.merge2_bias = p(P + "convmerge2/kernel"), p(P + "convmerge2/bias")
    return s, keep
