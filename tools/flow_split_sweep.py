"""Device time of one forward (16 clips x 7x32x32, fp16x3, L2 flushed between iterations) for a list of role splits of
the PFRB dataflow kernel (PFNL_FLOW_SPLIT = CTAs for conv1,conv10,conv2b,conv2f).
    python tools/flow_split_sweep.py 64,10,10,64 64,11,9,64 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from pfnl_b200 import Engine, weights as WT
    e = Engine(WT.xavier_init(), 0, os.environ.get("SWEEP_PREC", "fp16x3"), graphs=True)
    n, size = int(os.environ.get("SWEEP_N", "16")), int(os.environ.get("SWEEP_SIZE", "32"))
    h, w = int(os.environ.get("SWEEP_H", size)), int(os.environ.get("SWEEP_W", size))
    x = torch.rand(n, 7, h, w, 3, device="cuda")
    out = torch.empty(n, 1, 4 * h, 4 * w, 3, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        e.forward(x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get("SWEEP_ITERS", "30"))):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.forward(x, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("split", os.environ.get("PFNL_FLOW_SPLIT", "default"), "ms mean %.4f median %.4f min %.4f" % (sum(ts) / len(ts), ts[len(ts) // 2], ts[0]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one()
    else:
        for sp in sys.argv[1:] or ["default"]:
            env = dict(os.environ)
            if sp != "default":
                env["PFNL_FLOW_SPLIT"] = sp
            r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
            print("\n".join(l for l in (r.stdout + r.stderr).split("\n") if l.startswith("split") or "rror" in l)[:400], flush=True)
