// ifb_whitespace.h -- host half of detect_content (graphics/whitespace.rs:284-421): the reference's window walk, replayed
// over the per-pixel code map that whitespace_codes_kernel produces.  Product code: must not include or link anything from oracle/.
#pragma once
#include <cstdint>

namespace ifb {

// Code of pixel (x, y) as the centre of its 3x3 neighbourhood (whitespace.rs:525-613, the part of sobel_scharr_detect
// that does not depend on the window): 0xFF = border pixel or Scharr value <= threshold; otherwise
// bits 1:0 local_min_x (0..2), 3:2 local_max_x - 1, 5:4 local_min_y, 7:6 local_max_y - 1, relative to (x - 1, y - 1).
constexpr uint8_t kNoEdge = 0xFF;

// detect_content over a w*h code map (row pitch w).  rect = {x1, y1, x2, y2}.  `centres`, if not null, receives the number of
// window-interior pixels visited.  Returns false for an empty or oversized bitmap.
bool detect_content_from_codes(const uint8_t* codes, uint32_t w, uint32_t h, uint32_t rect[4], uint64_t* centres);

}  // namespace ifb
