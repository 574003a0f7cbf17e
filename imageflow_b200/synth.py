"""Seed-stable synthetic BGRA frames, numpy and torch flavours producing identical bytes.

gradient: B=x%256, G=y%256, R=(x+y)%256, A=255   (reference bench: benches/bench_graphics.rs:405-414)
noise   : counter hash of (0x1F2E3D4C+seed, x, y), uniform u8 per channel; alpha 'opaque' or 'mixed'
          (25 % exactly 0, 25 % exactly 255, rest uniform) -- worst case for the LUT gathers.
"""
from __future__ import annotations

import numpy as np

_M = 0xFFFFFFFF


def gradient_np(w: int, h: int, alpha: int = 255) -> np.ndarray:
    x = np.arange(w, dtype=np.uint32)[None, :]
    y = np.arange(h, dtype=np.uint32)[:, None]
    a = np.empty((h, w, 4), np.uint8)
    a[..., 0] = (x % 256).astype(np.uint8)
    a[..., 1] = (y % 256).astype(np.uint8)
    a[..., 2] = ((x + y) % 256).astype(np.uint8)
    a[..., 3] = alpha
    return a


def _mix_np(v):
    v = v.astype(np.uint64)
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x7FEB352D)) & np.uint64(_M)
    v ^= v >> np.uint64(15)
    v = (v * np.uint64(0x846CA68B)) & np.uint64(_M)
    v ^= v >> np.uint64(16)
    return v


def noise_np(w: int, h: int, seed: int = 0, alpha_mode: str = "opaque") -> np.ndarray:
    x = np.arange(w, dtype=np.uint64)[None, :]
    y = np.arange(h, dtype=np.uint64)[:, None]
    k = ((x * np.uint64(0x9E3779B1)) & np.uint64(_M)) ^ ((y * np.uint64(0x85EBCA77)) & np.uint64(_M)) ^ np.uint64((0x1F2E3D4C + seed) & _M)
    base = _mix_np(k)
    a = np.empty((h, w, 4), np.uint8)
    a[..., 0] = (base & np.uint64(0xFF)).astype(np.uint8)
    a[..., 1] = ((base >> np.uint64(8)) & np.uint64(0xFF)).astype(np.uint8)
    a[..., 2] = ((base >> np.uint64(16)) & np.uint64(0xFF)).astype(np.uint8)
    if alpha_mode == "opaque":
        a[..., 3] = 255
    else:
        h2 = _mix_np(base ^ np.uint64(0xA5A5A5A5))
        sel = (h2 >> np.uint64(24)) & np.uint64(3)
        al = ((h2 >> np.uint64(8)) & np.uint64(0xFF))
        a[..., 3] = np.where(sel == 0, 0, np.where(sel == 1, 255, al)).astype(np.uint8)
    return a


def _mix_t(v):
    v = v ^ (v >> 16)
    v = (v * 0x7FEB352D) & _M
    v = v ^ (v >> 15)
    v = (v * 0x846CA68B) & _M
    v = v ^ (v >> 16)
    return v


def noise_torch(w: int, h: int, seed: int = 0, alpha_mode: str = "opaque", device="cuda", out=None):
    """Same bytes as noise_np, computed on `device` (int64 arithmetic masked to 32 bits)."""
    import torch
    x = torch.arange(w, dtype=torch.int64, device=device)[None, :]
    y = torch.arange(h, dtype=torch.int64, device=device)[:, None]
    k = ((x * 0x9E3779B1) & _M) ^ ((y * 0x85EBCA77) & _M) ^ ((0x1F2E3D4C + seed) & _M)
    base = _mix_t(k)
    if out is None:
        out = torch.empty((h, w, 4), dtype=torch.uint8, device=device)
    out[..., 0] = (base & 0xFF).to(torch.uint8)
    out[..., 1] = ((base >> 8) & 0xFF).to(torch.uint8)
    out[..., 2] = ((base >> 16) & 0xFF).to(torch.uint8)
    if alpha_mode == "opaque":
        out[..., 3] = 255
    else:
        h2 = _mix_t(base ^ 0xA5A5A5A5)
        sel = (h2 >> 24) & 3
        al = (h2 >> 8) & 0xFF
        al = torch.where(sel == 0, torch.zeros_like(al), torch.where(sel == 1, torch.full_like(al, 255), al))
        out[..., 3] = al.to(torch.uint8)
    return out


def gradient_torch(w: int, h: int, alpha: int = 255, device="cuda", out=None):
    import torch
    x = torch.arange(w, dtype=torch.int64, device=device)[None, :]
    y = torch.arange(h, dtype=torch.int64, device=device)[:, None]
    if out is None:
        out = torch.empty((h, w, 4), dtype=torch.uint8, device=device)
    out[..., 0] = (x % 256).to(torch.uint8).expand(h, w)
    out[..., 1] = (y % 256).to(torch.uint8).expand(h, w)
    out[..., 2] = ((x + y) % 256).to(torch.uint8)
    out[..., 3] = alpha
    return out
