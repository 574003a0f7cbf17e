#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the reference checkout.

Run in the build container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_golden.py

Sources (relative to /root/reference):
  imageflow_core/tests/integration/weights.txt         -> weights_golden.json.gz["weights"]
  imageflow_core/tests/integration/weights_params.txt  -> weights_golden.json.gz["params"]
  imageflow_core/src/graphics/lut.rs:14 (LINEAR_TO_SRGB_LUT) -> lut16k_golden.bin.gz
The fixtures are the reference's own golden numbers re-encoded (parsed rows, 6-decimal
weights kept as the strings' float values); nothing is computed by our code here.
"""
import gzip
import json
import os
import re

REF = "/root/reference/imageflow_core"
HERE = os.path.dirname(os.path.abspath(__file__))

NAME_TO_ID = {  # weights.rs:45-78 (repr(C) ids) for the names used in weights_params.txt
    "Robidoux": 2, "RobidouxSharp": 3, "Mitchell": 14, "CatmullRom": 13, "Lanczos": 6, "Lanczos2": 8,
    "Lanczos2Sharp": 9, "Ginseng": 4, "Hermite": 16, "CubicFast": 10, "Triangle": 22, "Box": 24,
}


def parse_groups(s):
    out = []
    for _x, ws in re.findall(r"x=(\d+) from \(([^)]*)\)", s):
        out.append([float(v) for v in ws.split()])
    return out


def main():
    rows = []
    for line in open(f"{REF}/tests/integration/weights.txt").read().splitlines()[1:]:
        m = re.match(r"filter_(\d+) \(\s*(\d+)px to\s*(\d+)px\): (.*)", line)
        rows.append({"filter": int(m[1]), "in": int(m[2]), "out": int(m[3]), "w": parse_groups(m[4])})
    prow = []
    for line in open(f"{REF}/tests/integration/weights_params.txt").read().splitlines()[1:]:
        m = re.match(r"(\w+) (\S+) \(\s*(\d+)px to\s*(\d+)px\): (.*)", line)
        name, param = m[1], m[2]
        ks, lobe_mode, lobe_val = 1.0, 0, 0.0
        if param != "default":
            for p in param.split("+"):
                k, v = p.split("=")
                if k == "kernel_scale":
                    ks = float(v)
                elif k == "sharpen":
                    lobe_mode, lobe_val = 2, float(v)
                elif k == "lobe_exact":
                    lobe_mode, lobe_val = 1, float(v)
                else:
                    raise SystemExit("unknown param " + p)
        rec = {"filter": NAME_TO_ID[name], "name": name, "param": param, "kernel_scale": ks,
               "lobe_mode": lobe_mode, "lobe_value": lobe_val, "in": int(m[3]), "out": int(m[4])}
        if m[5].strip() == "ERROR":
            rec["error"] = True
        else:
            rec["w"] = parse_groups(m[5])
        prow.append(rec)
    with gzip.open(os.path.join(HERE, "weights_golden.json.gz"), "wt", compresslevel=9) as f:
        json.dump({"source": "imageflow_core/tests/integration/weights.txt + weights_params.txt @0ba1c9ea",
                   "weights": rows, "params": prow}, f, separators=(",", ":"))
    print(f"weights rows: {len(rows)}, params rows: {len(prow)}")

    src = open(f"{REF}/src/graphics/lut.rs").read()
    body = src[src.index("LINEAR_TO_SRGB_LUT: [u8; 16384] = [") + len("LINEAR_TO_SRGB_LUT: [u8; 16384] = ["):]
    body = body[:body.index("];")]
    vals = [int(v) for v in re.findall(r"\d+", body)]
    assert len(vals) == 16384, len(vals)
    with gzip.open(os.path.join(HERE, "lut16k_golden.bin.gz"), "wb", compresslevel=9) as f:
        f.write(bytes(vals))
    print("lut16k entries:", len(vals))


if __name__ == "__main__":
    main()
