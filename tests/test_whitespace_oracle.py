"""Oracle restatement of graphics/whitespace.rs (detect_content) against the reference's own known answers
(tests/integration/visuals/smoke.rs:420-592).  SURVEY.md section 8(f) item 4: the oracle comes first; the GPU side is
the window-independent per-pixel code map (oracle.whitespace_codes) with the reference's window walk replayed over it."""
import numpy as np

import oracle

RED, BLUE, CLEAR = (0, 0, 255, 255), (255, 0, 0, 255), (0, 0, 0, 0)        # BGRA bytes of "FF0000FF", "0000FFFF", Transparent


def _fill(a, x1, y1, x2, y2, bgra):
    a[y1:y2, x1:x2] = bgra


def test_detect_whitespace_basic():                                       # smoke.rs:420-475
    a = np.zeros((10, 10, 4), np.uint8)
    _fill(a, 1, 1, 9, 9, RED)
    assert oracle.detect_content(a, 1)[0] == (1, 1, 9, 9)
    b = np.zeros((100, 100, 4), np.uint8)
    _fill(b, 2, 3, 70, 70, RED)
    assert oracle.detect_content(b, 1)[0] == (2, 3, 70, 70)


def test_detect_whitespace_all_small_images():                             # smoke.rs:477-592
    combos = []
    for w in range(3, 12):
        for h in range(3, 12):
            on = [(x, y, sw, sh) for x in range(w) for y in range(h) for sw in (1, 2) for sh in (1, 2)
                  if not (x == 1 and y == 1 and w == 3 and h == 3) and x + sw <= w and y + sh <= h]
            combos.append((w, h, on))
    for (w, h) in [(3000, 2000), (1370, 1370), (1896, 1896), (3000, 3000)]:
        on = [(x, y, rw, rh) for x in (67, 0, 1, 881) for y in (67, 0, 1, 881) for (rw, rh) in ((1, 1), (1896, 1370))
              if x + rw <= w and y + rh <= h]
        combos.append((w, h, on))
    failures, count = [], 0
    for (w, h, on) in combos:
        a = np.zeros((h, w, 4), np.uint8)
        for (x, y, sw, sh) in on:
            a[...] = CLEAR
            _fill(a, x, y, x + sw, y + sh, RED)
            if sw > 2:
                _fill(a, x + 1, y + 1, x + sw - 1, y + sh - 1, BLUE)
            r = oracle.detect_content(a, 1)[0]
            if r != (x, y, x + sw, y + sh):
                failures.append((w, h, x, y, sw, sh, r))
            count += 1
    assert count > 10000
    assert len(failures) <= 3, failures[:10]           # the reference test tolerates up to three misses (smoke.rs:588-590)


def test_blank_and_tiny_bitmaps():
    assert oracle.detect_content(np.zeros((2, 7, 4), np.uint8), 1)[0] == (0, 0, 7, 2)          # whitespace.rs:288-290
    assert oracle.detect_content(np.zeros((40, 30, 4), np.uint8), 1)[0] == (0, 0, 30, 40)       # nothing found: the whole image (:324-326)
    full = np.full((40, 30, 4), 255, np.uint8)
    assert oracle.detect_content(full, 1)[0] == (0, 0, 30, 40)


def test_window_walk_can_be_replayed_over_the_code_map():
    """The per-pixel code depends only on the 3x3 neighbourhood, never on the window that visits it, so the scan can be
    replayed over a precomputed code map (what a GPU produces in one pass) with the identical result and visit count.
    The bounding box of ALL codes is only an outer bound: the full-image region stops at w-1 / h-1 (whitespace.rs:221-236)
    and windows are shrunk or skipped by the box found so far, so the ordered scan misses some centres on purpose."""
    rng = np.random.default_rng(7)
    differs = 0
    for it in range(600):
        mode = it % 4
        w, h = [(int(rng.integers(3, 1200)), int(rng.integers(3, 40))), (int(rng.integers(3, 40)), int(rng.integers(3, 1200))),
                (int(rng.integers(3, 400)), int(rng.integers(3, 400))), (int(rng.integers(280, 700)), int(rng.integers(3, 30)))][mode]
        a = np.zeros((h, w, 4), np.uint8)
        if rng.random() < 0.3:
            a[...] = rng.integers(0, 256, 4)
        for _ in range(int(rng.integers(1, 8))):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            if rng.random() < 0.5:
                a[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 4))] = rng.integers(0, 256, 4)
            else:
                x2, y2 = int(rng.integers(x, w)) + 1, int(rng.integers(y, h)) + 1
                a[y:y2, x:x2] = rng.integers(0, 256, (y2 - y, x2 - x, 4))
        thr, am = int(rng.choice([0, 1, 5, 30, 80])), bool(rng.integers(0, 2))
        direct = oracle.detect_content(a, thr, am)
        codes = oracle.whitespace_codes(a, thr, am)
        assert oracle.detect_content_from_codes(codes) == direct, (w, h, thr, am)
        ys, xs = np.nonzero(codes != 0xFF)
        if len(ys):
            c = codes[ys, xs].astype(np.int64)
            box = (int((xs - 1 + (c & 3)).min()), int((ys - 1 + ((c >> 4) & 3)).min()),
                   int((xs - 1 + ((c >> 2) & 3) + 1).max()), int((ys - 1 + ((c >> 6) & 3) + 1).max()))
            rect = direct[0]
            assert box[0] <= rect[0] and box[1] <= rect[1] and box[2] >= rect[2] and box[3] >= rect[3]
            differs += rect != box
    assert differs > 0            # the order really matters (about one image in seven here)
