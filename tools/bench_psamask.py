"""psa_mask: this repository's kernel against the reference's stock CUDA kernel on the same GPU, in the same run.

The stock kernel (lib/psa/src/gpu/psamask_cuda.cu:8-128, "the kernel the rewrite must beat") is the
reference's own extension compiled from its sources by oracle/build.py -> oracle/_ref/psamask_ref_gpu.so (test
infrastructure; nothing of it is on the product path). It is timed exactly as lib/psa/functions/psamask.py:17-35 calls it:
a zero-filled output allocation followed by the kernel. Algorithmic bytes per SURVEY.md §8(d): forward 2*4*N*(HW)^2,
backward 4*N*(HW)^2 + 4*N*mH*mW*HW. Prints one line per case and a JSON summary.
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from semseg_b200 import ops  # noqa: E402


def load_stock():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    for f in sorted(os.listdir(ref_dir)) if os.path.isdir(ref_dir) else []:
        if f.startswith("psamask_ref_gpu") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("psamask_ref_gpu", os.path.join(ref_dir, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


def timeit(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                     # > L2: every timed launch reads its input from HBM
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    stock = load_stock()
    peak = 6582.5
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    h = w = 30
    mh = mw = 59
    out = []
    for n in (2, 16):
        x = torch.randn((n, mh * mw, h, w), device="cuda")
        g = torch.randn((n, h * w, h, w), device="cuda")
        for t, name in ((0, "collect"), (1, "distribute")):
            fwd_bytes = 2 * 4 * n * (h * w) ** 2
            bwd_bytes = 4 * n * (h * w) ** 2 + 4 * n * mh * mw * h * w
            os.environ["SEMSEG_B200_PSA_ORDER"] = "0"           # A/B of the block order (i fastest vs h fastest)
            alt_f = timeit(lambda: ops.psamask_fwd(x, t, mh, mw), flush)
            alt_b = timeit(lambda: ops.psamask_bwd(g, t, mh, mw), flush)
            os.environ["SEMSEG_B200_PSA_ORDER"] = "1"
            ours_f = timeit(lambda: ops.psamask_fwd(x, t, mh, mw), flush)
            ours_b = timeit(lambda: ops.psamask_bwd(g, t, mh, mw), flush)
            print("   block order i-fastest: fwd %.4f bwd %.4f ms | h-fastest (default): fwd %.4f bwd %.4f ms" % (alt_f, alt_b, ours_f, ours_b))
            row = {"n": n, "type": name, "ours_fwd_ms": ours_f, "ours_bwd_ms": ours_b,
                   "ours_fwd_gbs": fwd_bytes / ours_f / 1e6, "ours_bwd_gbs": bwd_bytes / ours_b / 1e6}
            if stock is not None:
                def sf():
                    o = torch.zeros((n, h * w, h, w), device="cuda")            # lib/psa/functions/psamask.py:17
                    stock.psamask_forward(t, x, o, n, h, w, mh, mw, (mh - 1) // 2, (mw - 1) // 2)
                    return o

                def sb():
                    gi = torch.zeros((n, mh * mw, h, w), device="cuda")         # lib/psa/functions/psamask.py:31
                    stock.psamask_backward(t, g, gi, n, h, w, mh, mw, (mh - 1) // 2, (mw - 1) // 2)
                    return gi
                assert torch.equal(sf(), ops.psamask_fwd(x, t, mh, mw)) and torch.equal(sb(), ops.psamask_bwd(g, t, mh, mw))
                row["stock_fwd_ms"], row["stock_bwd_ms"] = timeit(sf, flush), timeit(sb, flush)
                row["speedup_fwd"] = row["stock_fwd_ms"] / ours_f
                row["speedup_bwd"] = row["stock_bwd_ms"] / ours_b
            row["ours_fwd_frac_hbm"] = row["ours_fwd_gbs"] / peak
            row["ours_bwd_frac_hbm"] = row["ours_bwd_gbs"] / peak
            out.append(row)
            print("N=%2d %-10s fwd: ours %.4f ms (%.0f GB/s = %.2f of HBM peak)%s | bwd: ours %.4f ms (%.0f GB/s = %.2f)%s" % (
                n, name, ours_f, row["ours_fwd_gbs"], row["ours_fwd_frac_hbm"],
                " stock %.4f ms (x%.1f)" % (row["stock_fwd_ms"], row["speedup_fwd"]) if stock else "",
                ours_b, row["ours_bwd_gbs"], row["ours_bwd_frac_hbm"],
                " stock %.4f ms (x%.1f)" % (row["stock_bwd_ms"], row["speedup_bwd"]) if stock else ""), flush=True)
    print(json.dumps({"psamask": out, "hbm_peak_gbs": peak, "stock_available": stock is not None}))


if __name__ == "__main__":
    main()
