"""Where the compact benchmark-shape goldens sample the big arrays: shared by the generator
(make_golden_bench_gpu.py) and the test (tests/test_gpu_bench_shapes.py)."""


def crop_origins(shape, size):
    """Origins (y, x) of the 3x3 grid of size x size crops (corners, edge centres, centre) of a [H,W,...] array."""
    H, W = shape[:2]
    return [(y, x) for y in (0, (H - size) // 2, H - size) for x in (0, (W - size) // 2, W - size)]
