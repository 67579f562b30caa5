"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/pfnl_b200.h declares; without a GPU the compute entry points
fail loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pfnl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfnl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 18, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pfnl_b200.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from pfnl_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    assert _lib.lib.pfnl_version() == 100


def test_weights_struct_layout(built_lib):
    from pfnl_b200 import _lib, weights
    # 6 + 6*20 + 4 pointers
    assert C.sizeof(_lib.PfnlWeights) == (6 + 6 * 20 + 4) * C.sizeof(C.c_void_p)
    w = weights.xavier_init()
    assert sum(int(np.prod(v.shape)) for v in w.values()) == 3003156
    st, keep = weights.to_struct(w)
    assert len(keep) == len(w)
    k = np.ctypeslib.as_array(st.conv2_kernel[19], shape=(3, 3, 128, 64))
    np.testing.assert_array_equal(k, w["nlvsr/conv2_19/kernel"])


def test_product_weights_match_oracle_regime_a(built_lib):
    from oracle import pfnl_ref as R
    from pfnl_b200 import weights
    a, b = weights.xavier_init(4321), R.make_weights("A", 4321)
    assert list(a) == list(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    with pytest.raises(ValueError):
        bad = dict(a)
        bad["nlvsr/conv0/kernel"] = np.zeros((3, 3, 3, 64), np.float32)
        weights.validate(bad)
    with pytest.raises(KeyError):
        bad = dict(a)
        del bad["nlvsr/convmerge2/bias"]
        weights.validate(bad)


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the library refuses to create a handle and the Python class
    raises instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pfnl_b200 import PFNL, _lib, weights
    st, keep = weights.to_struct(weights.xavier_init())
    h = C.c_void_p()
    rc = _lib.lib.pfnl_create(C.byref(h), 0, C.byref(st), 0)
    assert rc < 0 and not h.value
    assert _lib.lib.pfnl_last_error()
    with pytest.raises(RuntimeError):
        PFNL().forward(np.zeros((1, 7, 8, 8, 3), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pfnl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} references the oracle"


def test_flow_role_split_covers_the_device(built_lib):
    """pfrb_flow.cu gives every SM exactly one persistent CTA with one role: the split must use all SMs, leave no
    role empty, and be the measured one on a 148-SM B200 - host arithmetic, no GPU needed."""
    from pfnl_b200 import _lib
    lib = _lib.lib
    out = (C.c_int * 4)()
    for sms in list(range(8, 200)):
        for units in (1, 8, 128, 191, 192, 480, 4000):
            assert lib.pfnl_debug_flow_split(sms, units, out) == 0
            n = list(out)
            assert sum(n) == sms and min(n) >= 1, (sms, units, n)
            assert n[0] + n[3] >= 3 * (n[1] + n[2]), (sms, units, n)  # the two 7-tiles-per-unit roles dominate
    lib.pfnl_debug_flow_split(148, 128, out)
    assert list(out) == [62, 12, 9, 65]
    lib.pfnl_debug_flow_split(148, 480, out)
    assert list(out) == [58, 12, 9, 69]
    assert lib.pfnl_debug_flow_split(4, 128, out) != 0 and lib.pfnl_debug_flow_split(148, 0, out) != 0
