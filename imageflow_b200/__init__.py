"""imageflow_b200 -- B200-native stand-in for imageflow's BGRA resample hot path.

Host-side mirror of the reference seam (paths relative to the imageflow checkout):
  graphics/scaling.rs:9-23      ScaleAndRenderParams, scale_and_render(input, canvas, &params)
  graphics/color_matrix.rs:5    window_bgra32_apply_color_matrix(window, matrix)
  graphics/weights.rs:45-78     Filter
  graphics/bitmaps.rs:156-160   BitmapCompositing
  graphics/blend.rs:6-59        apply_matte
  graphics/transpose.rs:95      bitmap_window_transpose;  graphics/flip.rs:10,25  flips
  flow/nodes/white_balance.rs   white_balance_srgb_mut (histogram.rs, area thresholds, byte maps)
  graphics/whitespace.rs:284    detect_content (code map on the GPU, the reference's window walk on the host)
  c_components/lib/codecs_jpeg_idct_fast.c  flow_scale_spatial[_srgb]_NxN (decode-time JPEG block scalers)
All arithmetic runs in hand-written sm_100a kernels inside libifb200.so (include/ifb200.h);
this package only marshals arguments.  Nothing here imports the CPU oracle.
"""
from .graphics import (Batch, apply_matte, BitmapCompositing, BitmapWindow, ErrorKind, Filter, FlowError, ScaleAndRenderParams,
                       WorkingFloatspace, color_filter_matrix, device_count, plan_probe, populate_weights, scale_and_render, scale_and_render_many,
                       window_bgra32_apply_color_matrix, bitmap_window_transpose, flow_bitmap_bgra_flip_vertical_safe,
                       flow_bitmap_bgra_flip_horizontal_safe, white_balance_srgb_mut, detect_content, detect_content_from_codes, flow_scale_spatial)
from ._lib import LIB_PATH, ResampleDesc, lib

__all__ = [
    "Batch", "apply_matte", "BitmapCompositing", "BitmapWindow", "ErrorKind", "Filter", "FlowError", "ScaleAndRenderParams",
    "WorkingFloatspace", "color_filter_matrix", "device_count", "plan_probe", "populate_weights", "scale_and_render", "scale_and_render_many",
    "window_bgra32_apply_color_matrix", "bitmap_window_transpose", "flow_bitmap_bgra_flip_vertical_safe",
    "flow_bitmap_bgra_flip_horizontal_safe", "white_balance_srgb_mut", "detect_content", "detect_content_from_codes", "flow_scale_spatial", "LIB_PATH", "ResampleDesc", "lib",
]
