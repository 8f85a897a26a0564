"""Seeded synthetic inputs of the benchmark and parity tests.

The reference's datasets (Oxford, Hannover) are not vendored, so every input is
generated: a "blob field" image whose SIFT feature count is controlled by the
number of blobs, and random / detector-made descriptor sets for the matcher.
"""
import numpy as np


def blob_image(width, height, n_blobs, seed):
    """u8 grayscale (height,width): 0.5 + sum of n_blobs signed Gaussians, clipped."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 0.5, np.float32)
    cx = rng.uniform(0, width, n_blobs)
    cy = rng.uniform(0, height, n_blobs)
    sig = np.exp(rng.uniform(np.log(2.0), np.log(12.0), n_blobs))
    amp = rng.uniform(0.15, 0.6, n_blobs) * rng.choice([-1.0, 1.0], n_blobs)
    for i in range(n_blobs):
        r = int(np.ceil(4 * sig[i]))
        x0, x1 = max(0, int(cx[i]) - r), min(width, int(cx[i]) + r + 1)
        y0, y1 = max(0, int(cy[i]) - r), min(height, int(cy[i]) + r + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        xs = np.arange(x0, x1, dtype=np.float32) - np.float32(cx[i])
        ys = np.arange(y0, y1, dtype=np.float32) - np.float32(cy[i])
        g = np.exp(-(ys[:, None] ** 2 + xs[None, :] ** 2) / np.float32(2 * sig[i] ** 2))
        img[y0:y1, x0:x1] += np.float32(amp[i]) * g
    return np.round(np.clip(img, 0, 1) * 255).astype(np.uint8)


# calibrated workloads (feature counts are asserted in tests/ against the oracle)
C1 = dict(width=640, height=480, n_blobs=500, seed=42)      # BASELINE configs[0]
C2 = dict(width=1920, height=1080, n_blobs=2400, seed=7)    # BASELINE configs[1]


def random_descriptors(n, seed):
    """(n,128) u8 uniform random descriptors (BASELINE configs[3])."""
    return np.random.default_rng(seed).integers(0, 256, (n, 128), dtype=np.uint8)
