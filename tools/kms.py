#!/usr/bin/env python
"""one-line digest of a bench.py JSON line (kernel time, roofline fraction, clocks, parity)"""
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    r = d["roofline"]; c = d.get("clocks") or {}
    print(f"ms {r['kernel_ms']:.3f} (min {r['kernel_ms_min']:.3f}) frac {r['frac']:.3f} GB/s {r['achieved']:.0f} clk {c.get('sm_mhz')} W {c.get('power_w_max')} {c.get('reasons')} "
          f"chk {d.get('parity_check')} jobs f{r['fused_jobs']}/g{r['generic_jobs']}/t{r['tile_jobs']} n={d['config']['images_per_gpu_per_step']}")
except Exception as e:
    print("FAILED", e, open(sys.argv[1]).read()[-400:].replace("\n", " "))
