"""Seeded synthetic 752x480 test images: multi-octave value noise + random rectangles, FAST-rich."""
import numpy as np


def make_image(seed=0, w=752, h=480):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float64)
    for octave, amp in [(6, 60.0), (12, 40.0), (24, 30.0), (48, 20.0), (96, 12.0)]:
        gh, gw = h // octave + 2, w // octave + 2
        g = rng.random((gh, gw))
        ys = np.arange(h) / octave; xs = np.arange(w) / octave
        y0 = ys.astype(int); x0 = xs.astype(int); fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
        v = (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0][:, x0 + 1] * (1 - fy) * fx +
             g[y0 + 1][:, x0] * fy * (1 - fx) + g[y0 + 1][:, x0 + 1] * fy * fx)
        img += amp * v
    for _ in range(140):
        x, y = rng.integers(0, w - 8), rng.integers(0, h - 8)
        ww, hh = rng.integers(6, 90), rng.integers(6, 70)
        img[y:y + hh, x:x + ww] += rng.uniform(-70, 70)
    img += rng.normal(0, 2.0, size=img.shape)
    img = (img - img.min()) / (img.max() - img.min()) * 255.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
