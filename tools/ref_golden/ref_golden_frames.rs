//! Golden frames of the REAL reference resample path, for pinning imageflow_b200's oracle and kernels.
//!
//! This file is NOT part of the product and cannot be built in the imageflow_b200 containers (no Rust toolchain, and
//! `zenresize` 0.3.1 is not vendored).  On any machine with the imazen/imageflow checkout (commit 0ba1c9ea) and its crates:
//!
//!     cp tools/ref_golden/ref_golden_frames.rs  <imageflow>/imageflow_core/tests/ref_golden_frames.rs
//!     cd <imageflow> && IFB_GOLDEN_OUT=/path/to/imageflow_b200/tests/golden/frames \
//!         cargo test -p imageflow_core --release --test ref_golden_frames -- --nocapture
//!
//! It calls `imageflow_core::graphics::scaling::scale_and_render` exactly as `benches/bench_graphics.rs:429-450` does, on the
//! seeded frames of `imageflow_b200/synth.py` (re-stated below: `noise` = counter hash of (0x1F2E3D4C + seed, x, y), `gradient` =
//! bench_graphics.rs:405-414), and writes for every case
//!     <name>.bgra      raw canvas bytes after the call, rows packed (w * 4 bytes)
//! plus `manifest.json` describing the cases.  `tests/test_ref_golden_frames.py` then compares the oracle and the GPU path with
//! these frames at |delta| <= 1 per channel (the reference's own band: tests/integration/visuals/scaling.rs:18) and prints the
//! difference histogram.
use imageflow_core::graphics::bitmaps::*;
use imageflow_core::graphics::color::WorkingFloatspace;
use imageflow_core::graphics::scaling::{scale_and_render, ScaleAndRenderParams};
use imageflow_core::graphics::weights::Filter;
use imageflow_types::*;
use std::io::Write;

fn mix(mut v: u32) -> u32 {
    v ^= v >> 16;
    v = v.wrapping_mul(0x7FEB_352D);
    v ^= v >> 15;
    v = v.wrapping_mul(0x846C_A68B);
    v ^= v >> 16;
    v
}

/// imageflow_b200/synth.py noise_np / noise_torch, byte for byte.  alpha_mixed: 25 % exactly 0, 25 % exactly 255, rest uniform.
fn noise_px(x: u32, y: u32, seed: u32, alpha_mixed: bool) -> [u8; 4] {
    let k = x.wrapping_mul(0x9E37_79B1) ^ y.wrapping_mul(0x85EB_CA77) ^ 0x1F2E_3D4Cu32.wrapping_add(seed);
    let base = mix(k);
    let a = if !alpha_mixed {
        255
    } else {
        let h2 = mix(base ^ 0xA5A5_A5A5);
        match (h2 >> 24) & 3 {
            0 => 0,
            1 => 255,
            _ => ((h2 >> 8) & 0xFF) as u8,
        }
    };
    [(base & 0xFF) as u8, ((base >> 8) & 0xFF) as u8, ((base >> 16) & 0xFF) as u8, a]
}

fn gradient_px(x: u32, y: u32) -> [u8; 4] {
    [(x % 256) as u8, (y % 256) as u8, ((x + y) % 256) as u8, 255]
}

struct Case {
    name: &'static str,
    in_w: u32,
    in_h: u32,
    out_w: u32,
    out_h: u32,
    filter: Filter,
    filter_id: u32,          // weights.rs:45-78 repr(C) value, for the manifest
    sharpen: f32,
    linear: bool,
    alpha: bool,             // input alpha meaningful (and mixed alpha content)
    compose: u8,             // 0 ReplaceSelf, 1 BlendWithSelf (canvas = noise seed 100000+seed), 2 BlendWithMatte
    matte: [u8; 4],          // r, g, b, a of the matte colour (compose == 2)
    content: &'static str,   // "noise" | "gradient"
    seeds: u32,              // seeds 0 .. seeds-1, one frame each
}

const CASES: &[Case] = &[
    Case { name: "c1_640x480_to_200x150_robidoux", in_w: 640, in_h: 480, out_w: 200, out_h: 150, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "bench_800x600_to_400x300_robidoux_alpha", in_w: 800, in_h: 600, out_w: 400, out_h: 300, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: true, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "c2q_960x540_to_128x128_robidoux", in_w: 960, in_h: 540, out_w: 128, out_h: 128, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "c2q_960x540_to_128x128_lanczos", in_w: 960, in_h: 540, out_w: 128, out_h: 128, filter: Filter::Lanczos, filter_id: 6, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "c2_3840x2160_to_512x512_robidoux", in_w: 3840, in_h: 2160, out_w: 512, out_h: 512, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 16 },
    Case { name: "c2_3840x2160_to_512x512_lanczos", in_w: 3840, in_h: 2160, out_w: 512, out_h: 512, filter: Filter::Lanczos, filter_id: 6, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 16 },
    Case { name: "c2_3840x2160_to_512x512_robidoux_gradient", in_w: 3840, in_h: 2160, out_w: 512, out_h: 512, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "gradient", seeds: 1 },
    Case { name: "c3q_1920x1080_to_480x270_robidoux_sharpen50", in_w: 1920, in_h: 1080, out_w: 480, out_h: 270, filter: Filter::Robidoux, filter_id: 2, sharpen: 50.0, linear: true, alpha: false, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "c4q_480x270_to_960x540_mitchell_over_canvas", in_w: 480, in_h: 270, out_w: 960, out_h: 540, filter: Filter::Mitchell, filter_id: 14, sharpen: 0.0, linear: true, alpha: true, compose: 1, matte: [0; 4], content: "noise", seeds: 4 },
    Case { name: "matte_640x400_to_200x125_robidoux", in_w: 640, in_h: 400, out_w: 200, out_h: 125, filter: Filter::Robidoux, filter_id: 2, sharpen: 0.0, linear: true, alpha: true, compose: 2, matte: [250, 120, 40, 200], content: "noise", seeds: 4 },
    Case { name: "srgb_512x384_to_128x96_mitchell", in_w: 512, in_h: 384, out_w: 128, out_h: 96, filter: Filter::Mitchell, filter_id: 14, sharpen: 0.0, linear: false, alpha: true, compose: 0, matte: [0; 4], content: "noise", seeds: 4 },
];

fn fill(bitmap: &mut Bitmap, w: u32, h: u32, f: &dyn Fn(u32, u32) -> [u8; 4]) {
    let mut window = bitmap.get_window_u8().unwrap();
    for y in 0..h {
        let row = window.row_mut(y as usize).unwrap();
        for x in 0..w {
            row[(x * 4) as usize..(x * 4 + 4) as usize].copy_from_slice(&f(x, y));
        }
    }
}

#[test]
fn write_golden_frames() {
    let out_dir = std::env::var("IFB_GOLDEN_OUT").unwrap_or_else(|_| "ifb_golden_frames".to_string());
    std::fs::create_dir_all(&out_dir).unwrap();
    let mut manifest = String::from("{\n  \"reference\": \"imazen/imageflow scale_and_render (zenresize streaming resize)\",\n  \"frames\": [\n");
    let mut first = true;
    for c in CASES {
        for seed in 0..c.seeds {
            let mut src = Bitmap::create_u8(c.in_w, c.in_h, PixelLayout::BGRA, false, c.alpha, ColorSpace::StandardRGB, BitmapCompositing::ReplaceSelf).unwrap();
            if c.content == "gradient" {
                fill(&mut src, c.in_w, c.in_h, &|x, y| gradient_px(x, y));
            } else {
                fill(&mut src, c.in_w, c.in_h, &|x, y| noise_px(x, y, seed, c.alpha));
            }
            let compose = match c.compose {
                0 => BitmapCompositing::ReplaceSelf,
                1 => BitmapCompositing::BlendWithSelf,
                _ => BitmapCompositing::BlendWithMatte(Color::Srgb(ColorSrgb::Hex(format!("{:02x}{:02x}{:02x}{:02x}", c.matte[0], c.matte[1], c.matte[2], c.matte[3])))),
            };
            let mut dst = Bitmap::create_u8(c.out_w, c.out_h, PixelLayout::BGRA, false, true, ColorSpace::StandardRGB, compose).unwrap();
            if c.compose == 1 {
                fill(&mut dst, c.out_w, c.out_h, &|x, y| noise_px(x, y, 100_000 + seed, true));
            }
            let params = ScaleAndRenderParams {
                x: 0, y: 0, w: c.out_w, h: c.out_h,
                sharpen_percent_goal: c.sharpen,
                interpolation_filter: c.filter,
                scale_in_colorspace: if c.linear { WorkingFloatspace::LinearRGB } else { WorkingFloatspace::StandardRGB },
            };
            scale_and_render(src.get_window_u8().unwrap(), dst.get_window_u8().unwrap(), &params).unwrap();
            let file = format!("{}_s{}.bgra", c.name, seed);
            let mut f = std::fs::File::create(std::path::Path::new(&out_dir).join(&file)).unwrap();
            let mut window = dst.get_window_u8().unwrap();
            for y in 0..c.out_h {
                f.write_all(&window.row_mut(y as usize).unwrap()[..(c.out_w * 4) as usize]).unwrap();
            }
            if !first { manifest.push_str(",\n"); }
            first = false;
            manifest.push_str(&format!(
                "    {{\"file\": \"{}\", \"in_w\": {}, \"in_h\": {}, \"out_w\": {}, \"out_h\": {}, \"filter\": {}, \"sharpen\": {}, \"linear\": {}, \"alpha\": {}, \"compose\": {}, \"matte_rgba\": [{}, {}, {}, {}], \"content\": \"{}\", \"seed\": {}}}",
                file, c.in_w, c.in_h, c.out_w, c.out_h, c.filter_id, c.sharpen, c.linear, c.alpha, c.compose, c.matte[0], c.matte[1], c.matte[2], c.matte[3], c.content, seed));
        }
    }
    manifest.push_str("\n  ]\n}\n");
    std::fs::write(std::path::Path::new(&out_dir).join("manifest.json"), manifest).unwrap();
}
