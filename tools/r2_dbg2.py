"""debug helper: the decomposition that differed under compute-sanitizer memcheck, run several times, differences located"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import imageflow_b200 as ifb
from tests import util
from tests.test_gpu_parity import _gpu_batch, _oracle

inp = util.noise(1280, 720, seed=7, alpha_mode="mixed")
canvas = np.zeros((180, 320, 4), np.uint8)
exp = _oracle(inp, canvas, filter=2, alpha_meaningful=True)
for (sc, mi) in [(48, 100000), (48, 100000), (64, 100000), (48, 4096), (16, 100000)]:
    got, fused = _gpu_batch(ifb, torch, inp, canvas, filter=2, alpha_meaningful=True, strip_cols=sc, min_items=mi)
    d = np.abs(got.astype(np.int16) - exp.astype(np.int16)).max(axis=2)
    ys, xs = np.nonzero(d)
    print(sc, mi, "fused", fused, "max", int(d.max()), "count", len(ys), "rows", sorted(set(ys.tolist()))[:20], "cols", sorted(set(xs.tolist()))[:40])
