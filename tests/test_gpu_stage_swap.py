"""Stage-swap rig (`-m gpu`): the UNMODIFIED reference (baseline/_ref, real Numba-CUDA) runs its
own main() with ONE stage replaced by libhhsr.so through the ctypes stub printed in INTEGRATION.md
(tools/integration_stub_merge.py) — proof that the C ABI drops in behind the reference's function surface: same device
arrays (Numba DeviceNDArray / torch tensors through __cuda_array_interface__), same in-place convention, same stream.
Skipped where the reference copy or Numba-CUDA is not available."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "baseline", "_ref", "handheld_super_resolution")


@pytest.mark.gpu
def test_reference_pipeline_with_merge_swapped_for_libhhsr():
    if not os.path.isdir(REF):
        pytest.skip("baseline/_ref (verbatim copy of the reference) is not present")
    try:
        from numba import cuda
        assert cuda.is_available()
    except Exception:
        pytest.skip("Numba-CUDA not available")
    import sys
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k.split(".")[0] == "handheld_super_resolution"}
    for k in saved_mods:                      # the product package has the same name: import the reference afresh
        del sys.modules[k]
    try:
        spec = importlib.util.spec_from_file_location("make_golden_gpu_rig", os.path.join(ROOT, "tests", "golden", "make_golden_gpu.py"))
        G = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(G)            # stubs for absent third-party modules; baseline/_ref first on sys.path
        G.install_copysign_shim()
        from handheld_super_resolution import super_resolution as SR
        assert os.path.abspath(SR.__file__).startswith(os.path.join(ROOT, "baseline", "_ref")), SR.__file__
        os.environ["HHSR_LIB"] = os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200", "handheld_super_resolution", "libhhsr.so")
        spec = importlib.util.spec_from_file_location("integration_stub_merge", os.path.join(ROOT, "tools", "integration_stub_merge.py"))
        stub = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(stub)
        synth = G.load_synth()
        burst, _ = synth.synth_burst(4, 352, 416, seed=12, max_shift=2.5, quantize_bits=14)
        std = np.load(os.path.join(G.REF, "data", "noise_model_std_ISO_100.npy"))
        diff = np.load(os.path.join(G.REF, "data", "noise_model_diff_ISO_100.npy"))
        outs = []
        calls = {"n": 0}

        def counted(*a, **k):
            calls["n"] += 1
            return stub.merge(*a, **k)
        for swapped in (False, True):
            cfg = G.make_config(2, 32, [1, 2, 2], burst[0], std, diff)
            original = SR.merge
            if swapped:
                SR.merge = counted
            try:
                out, _ = SR.main(burst[0], burst[1:], cfg)
                cuda.synchronize()
                outs.append(out.copy_to_host())
            finally:
                SR.merge = original
        assert calls["n"] == 3                                     # one call per comp frame went through libhhsr.so
        a, b = outs
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = np.isfinite(a)
        d = float(np.abs(a[m] - b[m]).max())
        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "stage_swap_report.txt"), "w") as f:
            f.write("reference main() with merge.merge -> libhhsr hhsr_merge_accumulate (ctypes stub of INTEGRATION.md): "
                    "max |difference| of the final image = %.3g on %d finite values, identical NaN set\n" % (d, int(m.sum())))
        assert d < 1e-4
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] == "handheld_super_resolution"]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
