#!/usr/bin/env python3
"""project_to_mel (reference model.py:224-226, 249) on the bench shape: (256 batches x 64 x 246 rows, 128) -> 768, bf16.

    python tools/bench_projection.py [--rows N]

Times ``adt_str_b200.ProjectToMel`` (csrc/project.cu: tcgen05, cast fused) against what the reference runs -
``nn.Linear`` under ``torch.autocast(bf16)`` (a cast kernel + cuBLAS) - on the same resident float32 log-mel matrix.
The GEMM has K = 128: it is bound by HBM (512 B read + 1536 B written per row), so the roofline is bytes / measured
copy bandwidth; the tensor-core share is reported beside it."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402


def measure(dev, rows=256 * 64 * 246, n_out=768, reps=5):
    from adt_str_b200 import ProjectToMel
    torch.manual_seed(0)
    lin = torch.nn.Linear(128, n_out).to(dev)
    proj = ProjectToMel.from_linear(lin).eval()
    x = torch.rand(rows, 128, device=dev)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    with torch.no_grad():
        ours = timed(lambda: proj(x))

        def library():
            with torch.autocast("cuda", torch.bfloat16):
                return lin(x)
        theirs = timed(library)
        same = float((proj(x) == library()).float().mean())
    bytes_alg = rows * (128 * 4 + n_out * 2)
    peak = 6550.0
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
            peak, tf_peak = float(mp["hbm_gbs"]), float(mp.get("bf16_tflops", 1684.0))
    except Exception:
        tf_peak = 1684.0
    flops = 2.0 * rows * 128 * n_out
    return {"what": "project_to_mel: (rows, 128) float32 log-mel -> (rows, 768) bf16, bias, bf16 autocast semantics",
            "rows": rows, "ms": ours, "library_ms": theirs, "speedup_vs_autocast_linear": theirs / ours,
            "bytes_alg": bytes_alg, "gbs": bytes_alg / ours / 1e6, "frac_of_hbm_peak": bytes_alg / ours / 1e6 / peak,
            "tflops": flops / ours / 1e9, "frac_of_bf16_peak": flops / ours / 1e9 / tf_peak,
            "bit_equal_fraction_vs_library": same}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256 * 64 * 246)
    ap.add_argument("--sweep", action="store_true", help="a ladder of row counts: 1 .. 256 training batches")
    args = ap.parse_args()
    if args.sweep:
        for rows in (15744, 2 * 15744, 4 * 15744, 8 * 15744, 16 * 15744, 64 * 15744, 256 * 15744):
            r = measure(torch.device("cuda", 0), rows)
            print(json.dumps({k: r[k] for k in ("rows", "ms", "library_ms", "speedup_vs_autocast_linear", "frac_of_hbm_peak",
                                                "bit_equal_fraction_vs_library")}))
    else:
        print(json.dumps(measure(torch.device("cuda", 0), args.rows)))
