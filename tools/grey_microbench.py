"""Grey image (Alg. 3) on one frame: the library's own FFT passes against the cuFFT route, CUDA-event times.
Under `ncu --metrics gpu__time_duration.sum` the launch list gives the per-pass durations."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200"))


def main():
    from handheld_super_resolution import utils_image as UI
    from handheld_super_resolution.synthetic import synth_burst
    iters = int(os.environ.get("ITERS", "20"))
    res = {}
    sizes = [tuple(int(v) for v in x.split("x")) for x in os.environ.get("SIZES", "3000x4000,6144x8192").split(",")]
    for (H, W) in sizes:
        burst, _ = synth_burst(2, H, W, seed=0, device="cuda", as_numpy=False)
        img = burst[1]

        def timeit(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        res["%dx%d_native_us" % (H, W)] = timeit(lambda: UI.compute_grey_images(img, "FFT"))
        UI.GREY_FFT_NATIVE = False
        res["%dx%d_cufft_us" % (H, W)] = timeit(lambda: UI.compute_grey_images(img, "FFT"))
        UI.GREY_FFT_NATIVE = True
        del burst, img
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
