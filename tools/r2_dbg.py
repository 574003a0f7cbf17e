#!/usr/bin/env python
"""GPU debug helper: runs single resample cases, each in its own process (a device trap poisons the CUDA context), and prints
OK / differs / the CUDA error per case.  usage: python tools/r2_dbg.py            (all cases)   |   python tools/r2_dbg.py one <spec>"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    dict(name="A ch3 1strip k0=0", iw=960, ih=540, ow=128, oh=128, alpha=0),
    dict(name="A4 ch4 1strip k0=0", iw=960, ih=540, ow=128, oh=128, alpha=1),
    dict(name="B ch3 1strip", iw=640, ih=480, ow=100, oh=150, alpha=0),
    dict(name="C ch4 2strips", iw=640, ih=480, ow=200, oh=150, alpha=1),
    dict(name="D ch3 2strips", iw=640, ih=480, ow=200, oh=150, alpha=0),
    dict(name="D1 ch3 2strips 1band", iw=640, ih=480, ow=200, oh=150, alpha=0, min_items=1),
    dict(name="E ch3 xoff=1", iw=960, ih=540, ow=128, oh=128, alpha=0, xoff=1),
    dict(name="F ch3 4K", iw=3840, ih=2160, ow=512, oh=512, alpha=0),
    dict(name="G ch3 lanczos", iw=960, ih=540, ow=128, oh=128, alpha=0, flt=6),
]


def one(spec):
    import numpy as np
    import torch
    import imageflow_b200 as ifb
    import oracle
    from tests import util
    iw, ih, ow, oh = spec["iw"], spec["ih"], spec["ow"], spec["oh"]
    alpha, xoff, flt = bool(spec.get("alpha", 0)), spec.get("xoff", 0), spec.get("flt", 2)
    inp = util.noise(iw, ih, seed=iw + oh, alpha_mode="mixed" if alpha else "opaque")
    exp = np.zeros((oh, ow, 4), np.uint8)
    oracle.scale_and_render(inp, exp, filter=flt, alpha_meaningful=alpha)
    pitch = ((iw + xoff) * 4 + 63) // 64 * 64
    t = torch.zeros((ih, pitch), dtype=torch.uint8, device="cuda")
    v = t.as_strided((ih, iw, 4), (pitch, 4, 1), storage_offset=xoff * 4)
    v.copy_(torch.from_numpy(inp))
    out = torch.zeros((oh, ow, 4), dtype=torch.uint8, device="cuda")
    b = ifb.Batch(0)
    if "min_items" in spec:
        b.set_option(ifb.Batch.OPT_MIN_ITEMS, spec["min_items"])
    wi = ifb.BitmapWindow(v.data_ptr(), iw, ih, pitch, alpha_meaningful=alpha)
    b.scale_and_render_many([(wi, ifb.BitmapWindow.from_torch(out), ifb.ScaleAndRenderParams(w=ow, h=oh, interpolation_filter=ifb.Filter(flt)))],
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d = np.abs(out.cpu().numpy().astype(np.int16) - exp.astype(np.int16))
    print(f"RESULT fused={b.fused_jobs} max_delta={int(d.max())} bad={int((d > 0).sum())}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        one(json.loads(sys.argv[2]))
        sys.exit(0)
    for c in CASES:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", json.dumps(c)], capture_output=True, text=True, timeout=120, cwd=ROOT)
        res = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        err = [l for l in r.stderr.splitlines() if "rror" in l][-1:] if r.returncode else []
        print(f"{c['name']:28s} rc={r.returncode} {res[0] if res else ''} {err[0][:160] if err else ''}", flush=True)
