"""Regime-A (Xavier weights, outputs ~ +-100) max-abs error of the fp16x3 mode against the CPU oracle (fp64 = the
"true" value, fp32 = what the reference computes in) over several input seeds, for a list of values of the TMEM
truncation-bias compensation kappa (PFNL_TC_TRUNC_COMP; conv_tc_dev.cuh).
    python tools/gate_margin.py [kappa ...]         e.g.  python tools/gate_margin.py 0 0.09 0.18 0.27
Every kappa runs in its own process (the library reads the variable once)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEEDS = list(range(100, 108))


def child(ref_path):
    import torch
    from oracle import pfnl_ref as R  # noqa: E402  (checker)
    from pfnl_b200 import Engine  # noqa: E402
    z = np.load(ref_path)
    e = Engine(R.make_weights("A"), 0, "fp16x3", graphs=False)
    e64, e32 = [], []
    for i, seed in enumerate(SEEDS):
        x = R.make_input(1, 32, 32, seed=seed)
        y = e.forward(torch.from_numpy(x).cuda()).cpu().numpy()
        e64.append(float(np.abs(y - z["r64"][i]).max()))
        e32.append(float(np.abs(y - z["r32"][i]).max()))
    print("kappa", os.environ.get("PFNL_TC_TRUNC_COMP", "default"), "NL", "ffma" if os.environ.get("PFNL_NL_FFMA") else "tcgen05-split",
          "| vs fp64:", " ".join(f"{v:.2e}" for v in e64), "worst", f"{max(e64):.2e}", "| vs fp32 oracle: worst", f"{max(e32):.2e}",
          "mean", f"{np.mean(e32):.2e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    from oracle import pfnl_ref as R
    W = R.make_weights("A")
    r64, r32 = [], []
    for seed in SEEDS:
        x = R.make_input(1, 32, 32, seed=seed)
        r64.append(R.pfnl_forward(x, W, dtype=np.float64, backend="torch"))
        r32.append(R.pfnl_forward(x, W, dtype=np.float32, backend="numpy"))
    print("oracle fp32 vs fp64 per seed:", " ".join(f"{float(np.abs(a - b).max()):.2e}" for a, b in zip(r32, r64)), flush=True)
    path = os.path.join(tempfile.gettempdir(), "gate_margin_refs.npz")
    np.savez(path, r64=np.stack(r64), r32=np.stack(r32))
    kappas = sys.argv[1:] or ["default"]
    for k in kappas:
        env = dict(os.environ)
        if k != "default":
            env["PFNL_TC_TRUNC_COMP"] = k
        r = subprocess.run([sys.executable, __file__, "--child", path], env=env, capture_output=True, text=True, timeout=600)
        print("\n".join(l for l in (r.stdout + r.stderr).split("\n") if l.startswith("kappa") or "Error" in l), flush=True)
