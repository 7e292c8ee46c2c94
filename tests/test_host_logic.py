"""CPU tests: the C-ABI library loads and exports everything include/hhsr.h declares (no compute without a GPU),
argument validation, and the host-side logic (config object, SNR-derived parameters, sanitize_config, noise
curves, frame sharding)."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import curves

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hhsr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hhsr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from handheld_super_resolution import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libhhsr.so does not export %s" % s
    # the ctypes table binds exactly the declared compute entry points
    assert sorted(_lib.SIGNATURES) == [s for s in syms if s not in ("hhsr_version", "hhsr_last_error_string")]
    assert L.hhsr_version() == 100


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, with the error convention of include/hhsr.h."""
    from handheld_super_resolution import _lib
    L = _lib.lib()
    rc = L.hhsr_divide(None, None, 16, None)
    assert rc == -1 and b"null pointer" in L.hhsr_last_error_string()
    rc = L.hhsr_bm_l2_search(ctypes.c_void_p(16), 64, 64, ctypes.c_void_p(16), 64, 64, ctypes.c_void_p(16), 2, 2, 24, 4, None)
    assert rc == -2 and b"tile size" in L.hhsr_last_error_string()
    rc = L.hhsr_ica(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 64, 64, ctypes.c_void_p(16),
                    ctypes.c_void_p(16), 64, 64, ctypes.c_void_p(16), 4, 4, 12, 3, None)
    assert rc == -2
    cfa = (ctypes.c_int * 4)(0, 1, 1, 2)
    rc = L.hhsr_merge_accumulate(ctypes.c_void_p(16), 8, 8, ctypes.c_void_p(16), 1, 1, 32, ctypes.c_void_p(16),
                                 ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 16, 16, 0.5, cfa, 0, None)
    assert rc == -1 and b"scale" in L.hhsr_last_error_string()
    with pytest.raises(RuntimeError):
        _lib.call("hhsr_add_f64_f32", None, None, 0, None)


def test_missing_library_fails_loudly(monkeypatch):
    from handheld_super_resolution import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhhsr.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_config_object():
    from handheld_super_resolution.config import Config, load_config, to_plain
    cfg = load_config(overrides={"scale": 2, "merging": {"kernel": "iso"}})
    assert cfg.scale == 2 and cfg.merging.kernel == "iso" and cfg.merging.selection_law == "linear"
    assert cfg.accumulated_robustness_denoiser.merge.rad_max == 2          # key named like a dict-ish method
    assert cfg.noise_model.get("alpha", None) is None
    cfg.noise_model.update({"alpha": 1.0})
    cfg.exif = {"cfa_pattern": [[0, 1], [1, 2]]}
    assert isinstance(cfg.exif, Config) and cfg.exif.cfa_pattern[1][1] == 2
    assert to_plain(cfg)["noise_model"]["alpha"] == 1.0
    with pytest.raises(AttributeError):
        cfg.nope
    # same keys as the reference's configs/default.yaml
    for path in ["block_matching.tuning.factors", "block_matching.tuning.tile_size_factors", "ica.tuning.n_iter",
                 "robustness.tuning.Mt", "merging.tuning.k_shrink", "accumulated_robustness_denoiser.gauss.sigma_max",
                 "postprocessing.sharpening.amount", "grey_method", "mode"]:
        node = cfg
        for k in path.split("."):
            node = node[k]


def test_update_snr_config_and_lerp():
    from handheld_super_resolution.config import load_config
    from handheld_super_resolution.params import lerp, update_snr_config
    assert lerp(18, [6, 30], [0.33, 0.25]) == pytest.approx(0.29)
    assert lerp(100, [6, 30], [5.0, 3.0]) == 3.0 and lerp(0, [6, 30], [5.0, 3.0]) == 5.0
    for snr, ts in [(5, 64), (14, 64), (14.1, 32), (22, 32), (22.5, 16), (80, 16)]:
        cfg = load_config()
        update_snr_config(cfg, snr)
        assert cfg.block_matching.tuning.tile_size == ts
        assert cfg.block_matching.tuning.tile_sizes == [ts, ts, ts, ts // 2]
        assert 0.25 <= cfg.merging.tuning.k_detail <= 0.33 and 1 <= cfg.merging.tuning.D_tr <= 1.24
    cfg = load_config(overrides={"block_matching": {"tuning": {"tile_size": 32}}, "merging": {"tuning": {"k_detail": 0.3}}})
    update_snr_config(cfg, 40)
    assert cfg.block_matching.tuning.tile_size == 32 and cfg.merging.tuning.k_detail == 0.3
    assert cfg.merging.tuning.k_denoise == 3.0 and cfg.merging.tuning.D_th == pytest.approx(0.71)
    cfg = load_config(overrides={"merging": {"tuning": {"k_detail": 1}}})
    with pytest.raises(AssertionError):
        update_snr_config(cfg, 10)


def test_sanitize_config():
    from handheld_super_resolution.config import load_config
    from handheld_super_resolution.params import pyramid_shapes, sanitize_config, update_snr_config

    def cfg_for(ts=32, **over):
        c = load_config(overrides={"block_matching": {"tuning": {"tile_size": ts}}})
        c.merge_with(over)
        update_snr_config(c, 30)
        return c
    sanitize_config(cfg_for(), (3000, 4000))
    assert pyramid_shapes((3008, 4000), [1, 2, 4, 4]) == [(3008, 4000), (1500, 1996), (371, 495), (88, 119)]
    with pytest.raises(ValueError):
        sanitize_config(cfg_for(), (100, 100))
    with pytest.raises(ValueError, match="valid Gaussian"):
        sanitize_config(cfg_for(ts=16), (256, 256))          # passes the reference's check, fails for real
    with pytest.raises(AssertionError):
        sanitize_config(cfg_for(scale=0.5), (3000, 4000))
    with pytest.raises(ValueError):
        sanitize_config(cfg_for(robustness={"enabled": False}), (3000, 4000))       # save_mask still on
    with pytest.raises(AssertionError):
        sanitize_config(cfg_for(merging={"kernel": "box"}), (3000, 4000))
    with pytest.raises(ValueError):
        sanitize_config(cfg_for(accumulated_robustness_denoiser={"median": {"enabled": True}, "merge": {"enabled": True}}),
                        (3000, 4000))
    with pytest.raises(AssertionError):
        sanitize_config(cfg_for(block_matching={"tuning": {"flow_upscale_mode": "cubic"}}), (3000, 4000))


def test_noise_curve_host_logic_close_to_reference():
    """run_fast_MC's host part (non-linearity bounds, placement of the Monte-Carlo levels, interpolation in between,
    fast_monte_carlo.py:157-230) with the NumPy estimator standing in for the device kernel (tested in test_gpu_parity)."""
    from handheld_super_resolution.noise_model import regular_MC_numpy, run_fast_MC
    s1, d1 = run_fast_MC(1.80710882e-4, 3.1937599182128e-6, seed=0, n_patches=20000, mc=regular_MC_numpy)
    s2, d2 = run_fast_MC(1.80710882e-4, 3.1937599182128e-6, seed=0, n_patches=20000, mc=regular_MC_numpy)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and s1.shape == (1001,)
    std, diff = curves()          # the reference's data/noise_model_*_ISO_100.npy
    assert np.abs(s1 - std).max() / std.max() < 0.03 and np.abs(d1 - diff).max() / diff.max() < 0.06


def test_synthetic_burst_deterministic():
    from handheld_super_resolution.synthetic import synth_burst
    a, sa = synth_burst(3, 64, 96, seed=3)
    b, sb = synth_burst(3, 64, 96, seed=3)
    assert np.array_equal(a, b) and sa == sb and a.dtype == np.float32 and a.min() >= 0 and a.max() <= 1
    assert sa[0] == (0.0, 0.0)


def test_frame_sharding():
    from handheld_super_resolution.distributed import shard_frames
    for n, g in [(19, 8), (12, 8), (7, 2), (3, 4), (1, 1), (0, 2)]:
        parts = [shard_frames(n, r, g) for r in range(g)]
        assert sorted(sum(parts, [])) == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [len(shard_frames(19, r, 8)) for r in range(8)] == [3, 3, 3, 2, 2, 2, 2, 2]


def test_row_slices_of_the_row_sharded_merge():
    from handheld_super_resolution.distributed import equal_row_slices, p2p_row_slices
    for Hs, world in [(6000, 8), (6000, 2), (9000, 8), (12288, 4), (100, 3), (24, 8), (8, 8), (7, 2)]:
        for sl in (equal_row_slices(Hs, world), p2p_row_slices(Hs, world)):
            assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == Hs
            assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1)) and all(b >= a for a, b in sl)
        sl = equal_row_slices(Hs, world)
        assert all(a % 8 == 0 for a, _ in sl)                       # cuts on the 8-row blocks of the merge kernels
        if Hs >= 8 * world:
            sizes = [b - a for a, b in sl]
            assert min(sizes) > 0 and max(sizes) - min(sizes) <= 8 + Hs % 8


def test_raw_normalisation_host_matches_reference_loop():
    """RawNormalization.apply_numpy (the constants the device kernel uses, rounded to float32 up front) against the
    literal reference loop with Python-scalar operands (oracle.normalize_raw, utils_dng.py:146-160): bit-identical."""
    import hhsr_oracle as O
    from handheld_super_resolution.utils_dng import RawNormalization
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 16384, size=(3, 64, 96), dtype=np.uint16)
    for cfa, black, white, wb in [([[0, 1], [1, 2]], [1024, 1023, 1025, 1023], 16383, [2.1, 1.0, 1.63, 0.0]),
                                  ([[1, 2], [0, 1]], [64, 64, 64, 64], 1023, [1.91, 1.07, 1.4, 1.07]),
                                  ([[2, 1], [1, 0]], [0, 0, 0, 0], 4095, [1.0, 1.0, 1.0, 1.0])]:
        want = O.normalize_raw(raw, cfa, black, white, wb)
        got = RawNormalization(cfa, black, white, wb).apply_numpy(raw)
        assert got.dtype == np.float32 and np.array_equal(got, want)


def test_api_surface_matches_the_reference():
    """The drop-in boundary: every host-side stage function of the reference's hot path exists
    under the same module and name and takes the reference's positional parameters in the reference's order
    (tests/golden/api_signatures.json is extracted from the reference sources by make_api_signatures.py); extra
    keyword parameters with defaults are B200 additions."""
    import importlib
    import inspect
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "api_signatures.json")))
    assert len(ref) == 10
    for module, sigs in ref.items():
        mod = importlib.import_module("handheld_super_resolution." + module)
        for name, sig in sigs.items():
            assert hasattr(mod, name), "%s.%s is missing" % (module, name)
            params = inspect.signature(getattr(mod, name)).parameters
            ours = list(params)
            n = len(sig["args"])
            assert ours[:n] == sig["args"], (module, name, sig["args"], ours)
            for extra in ours[n:]:
                assert params[extra].default is not inspect.Parameter.empty, (module, name, extra)


def test_params_against_reference_golden():
    """update_snr_config / sanitize_config against what the reference's own params.py does on the same configs
    (tests/golden/params_cases.json, generated by make_golden_params.py): derived values identical, same accept /
    reject decisions with the same exception types.  Documented deviation: we additionally reject
    pyramids whose coarsest level holds no tile, which the reference only notices later."""
    import json
    from handheld_super_resolution.config import load_config
    from handheld_super_resolution.params import sanitize_config, update_snr_config
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "params_cases.json")))
    for c in g["snr"]:
        cfg = load_config(overrides=c["over"])
        update_snr_config(cfg, c["snr"])
        t, m = cfg.block_matching.tuning, cfg.merging.tuning
        assert t.tile_size == c["tile_size"] and list(t.tile_sizes) == c["tile_sizes"], c
        for k in ("k_detail", "k_denoise", "D_th", "D_tr"):
            assert m[k] == c[k], (c["snr"], k, m[k], c[k])          # bit-identical floats
    stricter = 0
    for c in g["sanitize"]:
        cfg = load_config(overrides=c["over"])
        try:
            update_snr_config(cfg, 30.0)
            sanitize_config(cfg, tuple(c["shape"]))
            res = "ok"
        except Exception as e:      # noqa: BLE001
            res = type(e).__name__
        if res != c["result"]:
            assert c["result"] == "ok" and res == "ValueError", (c, res)     # only ever stricter, and only the Q12 check
            stricter += 1
    assert stricter <= 1


def test_default_config_has_the_reference_keys_and_values():
    """configs/defaults.yaml carries exactly the keys and default values of the reference's configs/default.yaml
    (flattened in tests/golden/params_cases.json by make_golden_params.py)."""
    import json
    from handheld_super_resolution.config import load_config, to_plain
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "params_cases.json")))["default_config"]

    def flat(d, prefix=""):
        out = {}
        for k, v in d.items():
            if isinstance(v, dict):
                out.update(flat(v, prefix + k + "."))
            else:
                out[prefix + k] = v
        return out
    assert flat(to_plain(load_config())) == want
