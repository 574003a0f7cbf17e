"""Python mirror of the reference's operator interface for the hot path (argument meaning and error
behaviour follow graphics/scaling.rs:19-90 and graphics/color_matrix.rs:5-28); thin marshalling over the C ABI."""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from ._lib import ResampleDesc, lib


class Filter(enum.IntEnum):
    """graphics/weights.rs:45-78 (repr(C) discriminants)."""
    RobidouxFast = 1; Robidoux = 2; RobidouxSharp = 3; Ginseng = 4; GinsengSharp = 5; Lanczos = 6
    LanczosSharp = 7; Lanczos2 = 8; Lanczos2Sharp = 9; CubicFast = 10; Cubic = 11; CubicSharp = 12
    CatmullRom = 13; Mitchell = 14; CubicBSpline = 15; Hermite = 16; Jinc = 17; RawLanczos3 = 18
    RawLanczos3Sharp = 19; RawLanczos2 = 20; RawLanczos2Sharp = 21; Triangle = 22; Linear = 23; Box = 24
    CatmullRomFast = 25; CatmullRomFastSharp = 26; Fastest = 27; MitchellFast = 28; NCubic = 29
    NCubicSharp = 30; LegacyIDCTFilter = 31


class WorkingFloatspace(enum.IntEnum):
    """graphics/color.rs:4-9 (Gamma is not reachable from scale_and_render)."""
    StandardRGB = 0
    LinearRGB = 1


class BitmapCompositing(enum.IntEnum):
    """graphics/bitmaps.rs:156-160."""
    ReplaceSelf = 0
    BlendWithSelf = 1
    BlendWithMatte = 2


class ErrorKind(enum.IntEnum):
    Ok = 0; InvalidArgument = 1; MethodNotImplemented = 2; InvalidState = 3
    TotalWeightZero = 10; SourcePixelCountTooLarge = 11; NoPixelInputs = 12; BadFilter = 13; Capacity = 14
    NoDevice = 20; CudaError = 21; OutOfMemory = 22


class FlowError(RuntimeError):
    """Mirror of imageflow's FlowError{kind, message}."""
    def __init__(self, code: int, message: str):
        try:
            self.kind = ErrorKind(code)
        except ValueError:
            self.kind = ErrorKind.InvalidState
        self.code = code
        super().__init__(f"{self.kind.name}: {message}")


def _check(rc: int, buf) -> None:
    if rc:
        raise FlowError(rc, buf.value.decode("utf-8", "replace"))


@dataclass
class ScaleAndRenderParams:
    """graphics/scaling.rs:9-17."""
    x: int = 0
    y: int = 0
    w: int = 0
    h: int = 0
    sharpen_percent_goal: float = 0.0
    interpolation_filter: Filter = Filter.Robidoux
    scale_in_colorspace: WorkingFloatspace = WorkingFloatspace.LinearRGB


@dataclass
class BitmapWindow:
    """BitmapWindowMut<u8> for a BGRA bitmap: host (numpy) or device (raw pointer) pixels + the bits of
    BitmapInfo that scale_and_render reads (alpha_meaningful, compose; bitmaps.rs)."""
    ptr: int
    w: int
    h: int
    stride: int
    alpha_meaningful: bool = False
    compose: BitmapCompositing = BitmapCompositing.ReplaceSelf
    matte_bgra: Sequence[int] = (0, 0, 0, 0)
    pixel_layout: str = "BGRA"
    _keep: object = field(default=None, repr=False)

    @staticmethod
    def from_numpy(a: np.ndarray, **kw) -> "BitmapWindow":
        assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4 and a.strides[2] == 1 and a.strides[1] == 4
        return BitmapWindow(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], _keep=a, **kw)

    @staticmethod
    def from_torch(t, **kw) -> "BitmapWindow":
        """t: uint8 CUDA tensor of shape (H, W, 4) (row stride may be padded)."""
        assert t.dim() == 3 and t.shape[2] == 4 and t.stride(2) == 1 and t.stride(1) == 4
        return BitmapWindow(t.data_ptr(), t.shape[1], t.shape[0], t.stride(0), _keep=t, **kw)

    def window(self, x1: int, y1: int, x2: int, y2: int) -> Optional["BitmapWindow"]:
        """bitmaps.rs:413-431: sub-rect view (None when out of bounds)."""
        if not (0 <= x1 < x2 <= self.w and 0 <= y1 < y2 <= self.h):
            return None
        return BitmapWindow(self.ptr + y1 * self.stride + x1 * 4, x2 - x1, y2 - y1, self.stride, self.alpha_meaningful,
                            self.compose, self.matte_bgra, self.pixel_layout, self._keep)


def _fill_desc(d: ResampleDesc, inp: BitmapWindow, canvas: BitmapWindow, info: ScaleAndRenderParams, color_matrix=None, keep=None) -> None:
    """Writes one job into a zero-initialised ResampleDesc (in place: descriptor arrays are filled without temporaries)."""
    if inp.pixel_layout != "BGRA" or canvas.pixel_layout != "BGRA":            # scaling.rs:43-48
        raise FlowError(ErrorKind.MethodNotImplemented, "scale_and_render only supports BGRA bitmaps")
    d.in_ = inp.ptr; d.in_w = inp.w; d.in_h = inp.h; d.in_stride = inp.stride
    d.canvas = canvas.ptr; d.cv_w = canvas.w; d.cv_h = canvas.h; d.cv_stride = canvas.stride
    d.x = info.x; d.y = info.y; d.w = info.w; d.h = info.h
    d.filter = int(info.interpolation_filter)
    d.sharpen_percent = float(info.sharpen_percent_goal)
    d.linear = 1 if info.scale_in_colorspace == WorkingFloatspace.LinearRGB else 0
    d.alpha_meaningful = 1 if inp.alpha_meaningful else 0
    d.compose = int(canvas.compose)
    m = canvas.matte_bgra
    if m[0] or m[1] or m[2] or m[3]:
        d.matte_bgra = (C.c_uint8 * 4)(*m)
    if color_matrix is not None:
        cm = np.ascontiguousarray(color_matrix, np.float32).reshape(25)
        d.color_matrix = cm.ctypes.data
        (keep if keep is not None else []).append(cm)
        d._cm = cm


def _desc(inp: BitmapWindow, canvas: BitmapWindow, info: ScaleAndRenderParams, color_matrix=None, keep=None) -> ResampleDesc:
    d = ResampleDesc()
    _fill_desc(d, inp, canvas, info, color_matrix, keep)
    return d


def scale_and_render(inp: BitmapWindow, canvas: BitmapWindow, info: ScaleAndRenderParams, color_matrix=None) -> None:
    """graphics/scaling.rs:19-90 with HOST bitmaps (the drop-in call). Raises FlowError."""
    d = _desc(inp, canvas, info, color_matrix)
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_scale_and_render(C.byref(d), buf, 512), buf)


def scale_and_render_many(jobs) -> None:
    """n independent scale_and_render calls with HOST bitmaps, pipelined inside the library
    (jobs: iterable of (input, canvas, params[, color_matrix]))."""
    jobs = list(jobs)
    arr = (ResampleDesc * len(jobs))()
    keep = []
    for i, j in enumerate(jobs):
        arr[i] = _desc(j[0], j[1], j[2], j[3] if len(j) > 3 else None, keep)
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_scale_and_render_many(arr, len(jobs), buf, 512), buf)


def window_bgra32_apply_color_matrix(window: BitmapWindow, m) -> None:
    """graphics/color_matrix.rs:5-28, in place on a HOST window; m is [[f32;5];5]."""
    mm = np.ascontiguousarray(m, np.float32).reshape(25)
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_color_matrix_bgra8(window.ptr, window.w, window.h, window.stride,
                                            mm.ctypes.data_as(C.POINTER(C.c_float)), buf, 512), buf)


def apply_matte(window: BitmapWindow, matte_bgra) -> None:
    """graphics/blend.rs:6-59 (Bitmap::apply_matte), in place on a HOST window."""
    mm = (C.c_uint8 * 4)(*matte_bgra)
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_apply_matte_bgra8(window.ptr, window.w, window.h, window.stride, mm, int(bool(window.alpha_meaningful)), buf, 512), buf)


def bitmap_window_transpose(from_window: BitmapWindow, to_window: BitmapWindow) -> None:
    """graphics/transpose.rs:95-121 on HOST windows: to[x][y] = from[y][x]; dimensions must be swapped, BGRA only."""
    if from_window.w != to_window.h or from_window.h != to_window.w or from_window.pixel_layout != to_window.pixel_layout:
        raise FlowError(ErrorKind.InvalidArgument,
                        "For transposition, canvas and input formats must be the same and dimensions must be swapped")
    if from_window.pixel_layout != "BGRA":
        raise FlowError(ErrorKind.InvalidArgument, "Only BGRA layout is supported")
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_transpose_bgra8(from_window.ptr, from_window.stride, from_window.w, from_window.h,
                                        to_window.ptr, to_window.stride, buf, 512), buf)


def flow_bitmap_bgra_flip_vertical_safe(bitmap: BitmapWindow) -> None:
    """graphics/flip.rs:10-22, in place on a HOST bitmap."""
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_flip_vertical_bgra8(bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, buf, 512), buf)


def flow_bitmap_bgra_flip_horizontal_safe(bitmap: BitmapWindow) -> None:
    """graphics/flip.rs:25-39, in place on a HOST bitmap."""
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_flip_horizontal_bgra8(bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, buf, 512), buf)


def white_balance_srgb_mut(bitmap: BitmapWindow, threshold: Optional[float] = None) -> None:
    """flow/nodes/white_balance.rs:93-121 (WhiteBalanceHistogramAreaThresholdSrgb { threshold }), in place on a HOST bitmap."""
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_white_balance_srgb_bgra8(bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, -1.0 if threshold is None else float(threshold), buf, 512), buf)


def flow_scale_spatial(samples: np.ndarray, n: int, srgb: bool = False) -> np.ndarray:
    """flow_scale_spatial[_srgb]_{n}x{n} (c_components/lib/codecs_jpeg_idct_fast.c) applied to every 8x8 block of a plane of
    samples (H x W uint8, both multiples of 8; host memory): returns the (H/8*n) x (W/8*n) plane."""
    a = np.ascontiguousarray(samples, np.uint8)
    h, w = a.shape
    if h % 8 or w % 8:
        raise FlowError(int(ErrorKind.InvalidArgument), "plane dimensions must be multiples of 8")
    out = np.zeros((h // 8 * n, w // 8 * n) if 1 <= n <= 7 else (1, 1), np.uint8)
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_block_scale_u8(a.ctypes.data, a.strides[0], w // 8, h // 8, out.ctypes.data, out.strides[0], n, int(bool(srgb)), buf, 512), buf)
    return out


def detect_content(bitmap: BitmapWindow, threshold: int = 1):
    """graphics/whitespace.rs:284-331 with a HOST bitmap -> (x1, y1, x2, y2)."""
    rect = (C.c_uint32 * 4)()
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_detect_content_bgra8(bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, int(bool(bitmap.alpha_meaningful)), int(threshold), rect, buf, 512), buf)
    return tuple(rect)


def detect_content_from_codes(codes: np.ndarray):
    """host half of detect_content alone: the reference's window walk over a (h, w) uint8 code map (ifb_whitespace.h);
    no CUDA call.  -> ((x1, y1, x2, y2), pixels visited)"""
    codes = np.ascontiguousarray(codes, np.uint8)
    rect = (C.c_uint32 * 4)(); n = C.c_uint64()
    rc = lib().ifb200_detect_content_from_codes(codes.ctypes.data, codes.shape[1], codes.shape[0], rect, C.byref(n))
    if rc:
        raise FlowError(rc, "empty or oversized code map")
    return tuple(rect), n.value


def color_filter_matrix(which: int, p: float = 0.0) -> np.ndarray:
    """flow/nodes/color.rs:86-225 presets (0 sepia ... 9 saturation)."""
    m = np.zeros(25, np.float32)
    rc = lib().ifb200_color_filter_matrix(which, p, m.ctypes.data_as(C.POINTER(C.c_float)))
    if rc:
        raise FlowError(rc, "unknown colour filter")
    return m.reshape(5, 5)


def populate_weights(filter: int, out_size: int, in_size: int, kernel_width_scale: float = 1.0,
                     lobe_mode: int = 0, lobe_value: float = 0.0):
    """graphics/weights.rs:681-788 -> [(left_pixel, right_pixel, weights f32[])] per output pixel."""
    left = np.zeros(out_size, np.uint32); right = np.zeros(out_size, np.uint32); off = np.zeros(out_size + 1, np.uint32)
    cap = out_size * (int(2 * (6.5 * max(1.0, in_size / max(out_size, 1)) * max(kernel_width_scale, 1.0))) + 8)
    w = np.zeros(cap, np.float32)
    u32p, f32p = C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    rc = lib().ifb200_weights(int(filter), kernel_width_scale, lobe_mode, lobe_value, out_size, in_size,
                              left.ctypes.data_as(u32p), right.ctypes.data_as(u32p), off.ctypes.data_as(u32p),
                              w.ctypes.data_as(f32p), cap)
    if rc:
        raise FlowError(rc, "populate_weights failed")
    return [(int(left[i]), int(right[i]), w[off[i]:off[i + 1]].copy()) for i in range(out_size)]


def plan_probe(geometries, threads: int = 1, want_hash: bool = False) -> dict:
    """Host-side cost of preparing kernel tables: builds (and discards) the plans of `geometries` =
    [(in_w, in_h, out_w, out_h[, filter[, sharpen_percent]])] on `threads` host threads.  No CUDA call."""
    arr = (ResampleDesc * len(geometries))()
    for d, g in zip(arr, geometries):
        d.in_w, d.in_h, d.w, d.h = g[:4]
        d.filter = g[4] if len(g) > 4 else int(Filter.Robidoux)
        d.sharpen_percent = g[5] if len(g) > 5 else 0.0
    sec, nb, hs = C.c_double(), C.c_uint64(), C.c_uint64()
    buf = C.create_string_buffer(512)
    _check(lib().ifb200_plan_probe(arr, len(geometries), threads, C.byref(sec), C.byref(nb), C.byref(hs) if want_hash else None, buf, 512), buf)
    return {"seconds": sec.value, "table_bytes": nb.value, "table_hash": hs.value if want_hash else None}


def device_count() -> int:
    return lib().ifb200_device_count()


class Batch:
    """Device-resident batch executor (ifb200_batch_*): many independent scale_and_render calls whose
    bitmaps already live in HBM of one GPU, enqueued on a CUDA stream."""
    OPT_FORCE_GENERIC, OPT_STRIP_COLUMNS, OPT_MIN_ITEMS = 1, 2, 3

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_create(device, C.byref(self._h), buf, 512), buf)
        self.device = device
        self._keep = []

    def ring_status(self):
        """(usable, reason): whether the streaming ring kernel (the fast path of down-scales) can run on this device"""
        buf = C.create_string_buffer(256)
        ok = lib().ifb200_batch_ring_status(self._h, buf, 256)
        return bool(ok), buf.value.decode()

    def set_option(self, option: int, value: int) -> None:
        rc = lib().ifb200_batch_set_option(self._h, option, value)
        if rc:
            raise FlowError(rc, f"bad option {option}={value}")

    def make_descs(self, jobs):
        """jobs: iterable of (input BitmapWindow, canvas BitmapWindow, ScaleAndRenderParams[, color_matrix])."""
        jobs = list(jobs)
        arr = (ResampleDesc * len(jobs))()
        keep = []
        for i, j in enumerate(jobs):
            _fill_desc(arr[i], j[0], j[1], j[2], j[3] if len(j) > 3 else None, keep)
        keep.append([(j[0]._keep, j[1]._keep) for j in jobs])          # the bitmaps stay alive until sync() / close()
        return arr, keep

    STREAM_OWN = C.c_void_p(-1)

    @classmethod
    def _stream(cls, stream):
        """An int is a cudaStream_t (0 = the legacy default stream); "own" = the batch's private non-blocking stream
        (IFB200_STREAM_OWN, the one sync() waits for).  None = the stream the caller's tensors live on: torch's current stream
        when torch is loaded with CUDA -- work enqueued there is ordered after whatever produced the inputs and before whatever
        reads the outputs -- otherwise the batch's own stream."""
        if stream == "own":
            return cls.STREAM_OWN
        if stream is None:
            import sys
            torch = sys.modules.get("torch")
            if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
            return cls.STREAM_OWN
        return C.c_void_p(stream)

    def enqueue(self, descs, stream=None, keep=None) -> None:
        if keep:
            self._keep.append(keep)
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_enqueue(self._h, descs, len(descs), self._stream(stream), buf, 512), buf)

    def scale_and_render_many(self, jobs, stream=None) -> None:
        descs, keep = self.make_descs(jobs)
        self.enqueue(descs, stream, keep)

    def color_matrix(self, dev_ptr: int, w: int, h: int, stride: int, m, stream=None) -> None:
        mm = np.ascontiguousarray(m, np.float32).reshape(25)
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_color_matrix(self._h, dev_ptr, w, h, stride, mm.ctypes.data_as(C.POINTER(C.c_float)),
                                                self._stream(stream), buf, 512), buf)

    def transpose(self, from_window: BitmapWindow, to_window: BitmapWindow, stream=None) -> None:
        """graphics/transpose.rs:95-121 on DEVICE windows (asynchronous on `stream`)."""
        if from_window.w != to_window.h or from_window.h != to_window.w:
            raise FlowError(ErrorKind.InvalidArgument,
                            "For transposition, canvas and input formats must be the same and dimensions must be swapped")
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_transpose(self._h, from_window.ptr, from_window.stride, from_window.w, from_window.h,
                                            to_window.ptr, to_window.stride, self._stream(stream), buf, 512), buf)

    def block_scale(self, dev_in: int, in_stride: int, blocks_x: int, blocks_y: int, dev_out: int, out_stride: int, n: int, srgb: bool, stream=None) -> None:
        """flow_scale_spatial[_srgb]_{n}x{n} over a device-resident plane of 8x8 sample blocks"""
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_block_scale(self._h, dev_in, in_stride, blocks_x, blocks_y, dev_out, out_stride, n, int(bool(srgb)), self._stream(stream), buf, 512), buf)

    def white_balance(self, bitmap: BitmapWindow, threshold: Optional[float] = None, stream=None) -> None:
        """flow/nodes/white_balance.rs:93-121 on a DEVICE bitmap, in place (three kernels, asynchronous on `stream`)."""
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_white_balance(self._h, bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride,
                                                -1.0 if threshold is None else float(threshold), self._stream(stream), buf, 512), buf)

    def detect_content(self, bitmap: BitmapWindow, threshold: int = 1, stream=None):
        """graphics/whitespace.rs:284-331 on a DEVICE bitmap -> (x1, y1, x2, y2); synchronises `stream`."""
        rect = (C.c_uint32 * 4)()
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_detect_content(self._h, bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, int(bool(bitmap.alpha_meaningful)),
                                                 int(threshold), rect, self._stream(stream), buf, 512), buf)
        return tuple(rect)

    def whitespace_codes(self, bitmap: BitmapWindow, dev_codes_ptr: int, threshold: int = 1, stream=None) -> None:
        """the GPU half of detect_content alone (graphics/whitespace.rs:426-613 per pixel): code map of a DEVICE bitmap into a
        DEVICE buffer of w*h bytes; asynchronous on `stream`."""
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_whitespace_codes(self._h, bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, int(bool(bitmap.alpha_meaningful)),
                                                   int(threshold), dev_codes_ptr, self._stream(stream), buf, 512), buf)

    def flip_vertical(self, bitmap: BitmapWindow, stream=None) -> None:
        """graphics/flip.rs:10-22 on a DEVICE bitmap, in place."""
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_flip_vertical(self._h, bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, self._stream(stream), buf, 512), buf)

    def flip_horizontal(self, bitmap: BitmapWindow, stream=None) -> None:
        """graphics/flip.rs:25-39 on a DEVICE bitmap, in place."""
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_flip_horizontal(self._h, bitmap.ptr, bitmap.w, bitmap.h, bitmap.stride, self._stream(stream), buf, 512), buf)

    def sync(self) -> None:
        buf = C.create_string_buffer(512)
        _check(lib().ifb200_batch_sync(self._h, buf, 512), buf)
        import sys
        torch = sys.modules.get("torch")
        if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
            torch.cuda.synchronize(self.device)          # work enqueued on torch's streams (the default for stream=None)
        self._keep.clear()

    def host_profile(self) -> dict:
        """seconds the calling thread has spent inside enqueue calls so far, by part (ifb200_batch_host_profile)"""
        v = (C.c_double * 8)()
        lib().ifb200_batch_host_profile(self._h, v, 8)
        return {"enqueue_s": v[0], "plans_s": v[1], "staging_slots_s": v[2], "staging_memcpy_s": v[3], "table_uploads_s": v[4],
                "table_uploads": int(v[5]), "staged_bytes": int(v[6]), "pinned_allocs": int(v[7])}

    @property
    def kernel_launches(self) -> int:
        return lib().ifb200_batch_kernel_launches(self._h)

    @property
    def fused_jobs(self) -> int:
        return lib().ifb200_batch_fused_jobs(self._h)

    @property
    def generic_jobs(self) -> int:
        return lib().ifb200_batch_generic_jobs(self._h)

    @property
    def tile_jobs(self) -> int:
        return lib().ifb200_batch_tile_jobs(self._h)

    def close(self) -> None:
        if self._h:
            lib().ifb200_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
