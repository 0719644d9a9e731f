"""Seeded inputs shared by the CPU and GPU parity tests."""
import numpy as np

from oracle.oracle import synthetic_image


def make_image(kind, H, W, seed=1, sigma=12.0):
    if kind == "syn":
        return synthetic_image(H, W, seed, sigma)
    if kind == "flat":
        return np.full((H, W, 3), 77, np.uint8)
    if kind == "noise":
        return np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8)
    if kind == "blocks":  # piecewise-constant patches: many exact distance ties
        rng = np.random.RandomState(seed)
        small = rng.randint(0, 4, (H // 8 + 1, W // 8 + 1, 3)) * 60
        return np.ascontiguousarray(np.kron(small, np.ones((8, 8, 1)))[:H, :W].astype(np.uint8))
    if kind == "tiled":  # large images without float64 H x W x 3 temporaries: a synthetic tile repeated + a coarse ramp
        base = synthetic_image(540, 960, seed, sigma)
        img = np.tile(base, ((H + 539) // 540, (W + 959) // 960, 1))[:H, :W]
        ramp = ((np.arange(H)[:, None] // 7 + np.arange(W)[None, :] // 5) % 23).astype(np.uint16)
        return np.ascontiguousarray(np.minimum(img + ramp[..., None], 255).astype(np.uint8))
    raise ValueError(kind)


# name, kind, H, W, K, kwargs
PIPELINE_CASES = [
    ("A_640x480_K200", "syn", 480, 640, 200, {}),
    ("odd_97x131_K37_msf.1", "syn", 97, 131, 37, dict(min_size_factor=0.1)),
    ("flat_97x131_K37", "flat", 97, 131, 37, dict(min_size_factor=0.5)),
    ("noise_120x160_K48_msf0", "noise", 120, 160, 48, dict(min_size_factor=0.0)),
    ("noise_120x160_K48_msf.25", "noise", 120, 160, 48, dict(min_size_factor=0.25)),
    ("blocks_200x300_K150_msf0", "blocks", 200, 300, 150, dict(min_size_factor=0.0)),
    ("thin_300x10_K4", "syn", 300, 10, 4, {}),
    ("thin_10x400_K5_it2", "syn", 10, 400, 5, dict(max_iter=2)),
    ("thin_301x17_K6_it1", "syn", 301, 17, 6, dict(max_iter=1)),
    ("one_cluster_64x64", "syn", 64, 64, 1, {}),
    ("rgb_path_150x200_K300", "syn", 150, 200, 300, dict(convert_to_lab=False, min_size_factor=0.0)),
    ("compact37.5_150x200_K30", "syn", 150, 200, 30, dict(compactness=37.5)),
    ("compact1_180x240_K150", "syn", 180, 240, 150, dict(compactness=1.0)),
    ("stride2_it1", "syn", 150, 200, 30, dict(subsample_stride=2, max_iter=1)),
    ("stride1_it3", "syn", 90, 120, 20, dict(subsample_stride=1, max_iter=3)),
    ("stride5_it7", "syn", 150, 200, 60, dict(subsample_stride=5, max_iter=7)),
    ("it0", "syn", 120, 160, 40, dict(max_iter=0)),
    ("bigS_generic_300x400_K2", "syn", 300, 400, 2, {}),
    ("dense_K_64x64_K1500", "noise", 64, 64, 1500, dict(min_size_factor=0.0)),
    ("speckle_240x320_K100_msf0", "syn", 240, 320, 100, dict(min_size_factor=0.0, sigma=40.0)),
]

# shapes and parameters off the beaten path (each verified oracle == compiled reference when added)
EDGE_CASES = [
    ("tiny_5x7_K3", "noise", 5, 7, 3, {}),
    ("row_1x200_K7", "syn", 1, 200, 7, {}),
    ("col_200x1_K7", "syn", 200, 1, 7, {}),
    ("denseK_96x128_K6000", "noise", 96, 128, 6000, dict(min_size_factor=0.0)),
    ("S1_20x20_K300", "noise", 20, 20, 300, dict(min_size_factor=0.0)),
    ("stride255_it3", "syn", 300, 200, 40, dict(subsample_stride=255, max_iter=3)),
    ("compact0.01", "syn", 120, 160, 30, dict(compactness=0.01)),
    ("compact2000", "syn", 120, 160, 30, dict(compactness=2000.0)),
    ("it25", "syn", 100, 140, 25, dict(max_iter=25)),
    ("msf3_everything_absorbed", "noise", 100, 140, 25, dict(min_size_factor=3.0)),
    ("w33", "syn", 70, 33, 9, {}),
    ("w31", "syn", 70, 31, 9, {}),
]

# shapes aimed at the TMA-staged assign kernel (W % 8 == 0): widths that are / are not multiples of the 32- and
# 128-column tiles, heights that leave ragged sub-row groups, small S (single-tile super tiles, list overflow),
# uncovered pixels, warm start is covered separately
TMA_CASES = [
    ("tma_97x136_K37", "syn", 97, 136, 37, dict(min_size_factor=0.1)),
    ("tma_123x200_K60_msf0", "syn", 123, 200, 60, dict(min_size_factor=0.0)),
    ("tma_250x264_K100", "noise", 250, 264, 100, {}),
    ("tma_131x128_K50_flat", "flat", 131, 128, 50, dict(min_size_factor=0.5)),
    ("tma_200x328_blocks_K150", "blocks", 200, 328, 150, dict(min_size_factor=0.0)),
    ("tma_S8_160x160_K400", "syn", 160, 160, 400, dict(min_size_factor=0.0)),
    ("tma_S16_256x256_K256", "noise", 256, 256, 256, dict(min_size_factor=0.0)),
    ("tma_S5_96x104_K350", "syn", 96, 104, 350, dict(min_size_factor=0.0)),
    ("tma_thin_13x520_K6", "syn", 13, 520, 6, {}),
    ("tma_thin_500x8_K5", "syn", 500, 8, 5, {}),
    ("tma_1row_1x256_K9", "syn", 1, 256, 9, {}),
    ("tma_compact40_300x400_K200", "syn", 300, 400, 200, dict(compactness=40.0)),
    ("tma_rgb_222x344_K99", "syn", 222, 344, 99, dict(convert_to_lab=False)),
    ("tma_it1_150x200_K30", "syn", 150, 200, 30, dict(max_iter=1)),
    ("tma_it2_150x200_K30", "syn", 150, 200, 30, dict(max_iter=2, min_size_factor=0.0)),
    ("tma_S60_700x1000_K190", "syn", 700, 1000, 190, dict(min_size_factor=0.0)),
    ("tma_S90_900x1200_K130", "syn", 900, 1200, 130, dict(min_size_factor=0.0)),
]

BIG_CASES = [
    ("B_1280x720_K1600_msf0", "syn", 720, 1280, 1600, dict(min_size_factor=0.0)),
    ("B_1280x720_K1600_msf.1_s40", "syn", 720, 1280, 1600, dict(min_size_factor=0.1, sigma=40.0)),
    ("C_1920x1080_K2000_msf0", "syn", 1080, 1920, 2000, dict(min_size_factor=0.0)),
    ("D_3840x2160_K4000_msf0", "tiled", 2160, 3840, 4000, dict(min_size_factor=0.0)),
    # > 2^24 pixels: more than 32 of the 1024-pixel blocks per numbering warp (k_ccl_number's outer chunk loop)
    ("huge_4100x4200_K3000_msf.1", "tiled", 4100, 4200, 3000, dict(min_size_factor=0.1, sigma=20.0)),
]


def split_kwargs(kw):
    kw = dict(kw)
    sigma = kw.pop("sigma", 12.0)
    args = dict(max_iter=10, compactness=10.0, min_size_factor=0.25, subsample_stride=3, convert_to_lab=True)
    args.update(kw)
    return sigma, args
