"""BASELINE.json configs 4 and 5 on one B200 (run through gpurun):
  config 4: PFNL 4x large-tile inference, 1 clip x 7x128x128 (non-local L = 4096), each precision
  config 5: non-local block isolation sweep, 7 x {16,32,64,128}^2 LR -> L = {64,256,1024,4096}, C = 84:
            tcgen05 kernel (fp16 operands) and fp32 FFMA kernel; useful FLOPs F = 4*84*L*(84+L) per clip
Prints one JSON object; timings are CUDA events, median of 20 after 3 warm-ups, L2 flushed between runs."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfnl_b200 import Engine  # noqa: E402
from pfnl_b200 import weights as WT  # noqa: E402


def timeit(fn, flush, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    W = WT.xavier_init()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"peaks": {"tensor_tflops_burst": peaks["bf16_tflops"], "hbm_gbs": peaks["hbm_gbs"]}, "config4": {}, "config5": []}
    eng = {p: Engine(W, 0, p, graphs=True) for p in ("fp32", "fp16x3", "fp16")}
    # ---- config 4
    x = torch.rand(1, 7, 128, 128, 3, device="cuda")
    for p, e in eng.items():
        ms = timeit(lambda: e.forward(x), flush)
        out["config4"][p] = {"ms": ms, "hr_px_per_s": 512 * 512 / (ms / 1e3)}
    # ---- config 5
    for hw in (16, 32, 64, 128):
        L = (hw // 2) ** 2
        n = max(1, min(64, 16384 // L))           # clips chosen to fill the GPU: N*L/128 query tiles >= 128 where possible
        t = torch.rand(n, L, 84, device="cuda")
        flops = n * 4.0 * 84 * L * (84 + L)
        row = {"lr": hw, "L": L, "clips": n, "useful_gflop": flops / 1e9}
        for p in ("fp16", "fp32"):
            ms = timeit(lambda: eng[p].nonlocal_block(t), flush)
            row[p] = {"ms": ms, "tflops": flops / (ms / 1e3) / 1e12,
                      "frac_of_tensor_peak": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"] if p == "fp16" else None}
        out["config5"].append(row)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
