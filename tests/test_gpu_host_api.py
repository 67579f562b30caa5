"""Host-facing call paths of the drop-in boundary on the GPU: the pipelined feed/fetch (pfnl_forward_host_submit /
_wait, float64 input like the reference's feed_dict), and the CUDA-graph cache of pfnl_forward (tensor-core
precision, legacy NULL stream, changing buffer addresses)."""
import numpy as np
import pytest
import torch

from oracle import pfnl_ref as R

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.fixture(scope="module")
def eng16(built_lib):
    from pfnl_b200 import Engine
    e = Engine(R.make_weights("B"), device=0, precision="fp16x3", graphs=True)
    yield e
    e.close()


def test_host_float64_feed_matches_float32(eng16):
    """The reference feeds float64 numpy arrays to a float32 placeholder (pfnl.py:209,252): narrowing inside the
    library must equal numpy's cast, pageable or pinned, and the host path must equal the device path."""
    x64 = R.make_input(2, 16, 16).astype(np.float64)
    x64 += 1e-9  # not exactly representable in float32
    x32 = x64.astype(np.float32)
    y_dev = eng16.forward(cu(x32)).cpu()
    y64 = eng16.forward_host(x64)
    y32 = eng16.forward_host(x32)
    y_pin = eng16.forward_host(torch.from_numpy(x32).pin_memory(), out=torch.empty(2, 1, 64, 64, 3).pin_memory())
    assert torch.equal(y64, y_dev) and torch.equal(y32, y_dev) and torch.equal(y_pin, y_dev)
    with pytest.raises(ValueError):
        eng16.forward_host(x32, out=torch.empty(2, 1, 64, 64, 4))
    with pytest.raises(ValueError):
        eng16.forward_host(cu(x32))


def test_pipelined_submit_wait_keeps_order(eng16):
    """Two batches in flight: every wait returns its own batch's result (slots are recycled after two submits)."""
    rng = np.random.default_rng(7)
    xs = [rng.random((2, 7, 16, 16, 3), dtype=np.float32) for _ in range(5)]
    ref = [eng16.forward(cu(x)).cpu() for x in xs]
    outs, tickets = [], []
    for k, x in enumerate(xs):
        t, o = eng16.forward_host_submit(x)
        tickets.append(t)
        outs.append(o)
        if k >= 1:
            eng16.forward_host_wait(tickets[k - 1])
            assert torch.equal(outs[k - 1], ref[k - 1])
    eng16.forward_host_wait(tickets[-1])
    assert torch.equal(outs[-1], ref[-1])
    assert tickets == [0, 1, 0, 1, 0] or tickets == [1, 0, 1, 0, 1]
    eng16.forward_host_wait(tickets[-1])  # waiting twice is harmless


def test_graph_cache_tensorcore_null_stream_and_new_buffers(built_lib):
    """fp16x3 under CUDA graphs: replay on the legacy NULL stream (private capture stream ordered with events) and
    on a torch stream equals plain launches bit for bit; outputs kept by the caller (a new address every call)
    re-point cached executables instead of instantiating one per call."""
    from pfnl_b200 import Engine
    W = R.make_weights("A")
    eg = Engine(W, 0, "fp16x3", graphs=True)
    ep = Engine(W, 0, "fp16x3", graphs=False)
    x = cu(R.make_input(2, 16, 16))
    ref = ep.forward(x)
    out = torch.empty_like(ref)
    l0 = eg.launches
    for _ in range(3):
        eg.forward(x, out=out)               # NULL stream: captured once, replayed twice
    per = (eg.launches - l0) // 3
    torch.cuda.synchronize()
    assert per == 7 and torch.equal(out, ref)
    assert eg.graph_stats()[0] == 1
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out_s = torch.empty_like(ref)
        eg.forward(x, out=out_s)
        eg.forward(x, out=out_s)
    s.synchronize()
    assert torch.equal(out_s, ref)
    kept = [eg.forward(x) for _ in range(12)]   # 12 live outputs = 12 different addresses
    torch.cuda.synchronize()
    assert all(torch.equal(k, ref) for k in kept)
    inst, upd, cached = eg.graph_stats()
    assert inst <= 6 and upd >= 6 and cached <= 6, (inst, upd, cached)
    # and a stream-ordered consumer right behind the replay sees the result (no missing dependency on the NULL stream)
    y = eg.forward(x, out=out)
    z = y * 2.0
    torch.cuda.synchronize()
    assert torch.equal(z, ref * 2.0)
    eg.close()
    ep.close()
