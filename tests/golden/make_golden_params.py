"""Golden vectors for the host parameter logic: run the REFERENCE's own params.py (it only imports numpy) on the
reference's configs/default.yaml and record what update_snr_config derives for a sweep of SNR values and what
sanitize_config accepts / rejects for a list of configurations -> tests/golden/params_cases.json.

    python tests/golden/make_golden_params.py        # needs /root/reference (build container only)
"""
import copy
import importlib.util
import json
import os

import yaml

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class Cfg(dict):
    """attribute/item mapping standing in for an OmegaConf DictConfig"""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        return Cfg({k: Cfg.wrap(v) for k, v in d.items()}) if isinstance(d, dict) else d


def merge(c, over):
    for k, v in over.items():
        if isinstance(v, dict):
            merge(c[k], v)
        else:
            c[k] = v


def main():
    spec = importlib.util.spec_from_file_location("ref_params", os.path.join(REF, "handheld_super_resolution", "params.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    base = yaml.safe_load(open(os.path.join(REF, "configs", "default.yaml")))
    def flat(d, prefix=""):
        r = {}
        for k, v in d.items():
            if isinstance(v, dict):
                r.update(flat(v, prefix + k + "."))
            else:
                r[prefix + k] = v
        return r
    out = {"default_config": flat(base), "snr": [], "sanitize": []}
    for snr in [0.5, 6, 6.01, 10, 13.99, 14, 14.01, 18, 22, 22.01, 25.5, 30, 31, 1000]:
        for over in ({}, {"block_matching": {"tuning": {"tile_size": 32}}}, {"merging": {"tuning": {"k_detail": 0.3, "D_tr": 1.1}}}):
            c = Cfg.wrap(copy.deepcopy(base))
            merge(c, over)
            P.update_snr_config(c, snr)
            t, m = c.block_matching.tuning, c.merging.tuning
            out["snr"].append({"snr": snr, "over": over, "tile_size": t.tile_size, "tile_sizes": list(t.tile_sizes),
                               "k_detail": m.k_detail, "k_denoise": m.k_denoise, "D_th": m.D_th, "D_tr": m.D_tr})
    cases = [({}, (3000, 4000)), ({}, (1200, 1600)), ({}, (100, 100)), ({"scale": 0.5}, (3000, 4000)), ({"scale": 1.5}, (3000, 4000)),
             ({"robustness": {"enabled": False}}, (3000, 4000)),
             ({"robustness": {"enabled": False, "save_mask": False}}, (3000, 4000)),
             ({"merging": {"kernel": "box"}}, (3000, 4000)), ({"merging": {"kernel": "iso"}}, (3000, 4000)),
             ({"accumulated_robustness_denoiser": {"median": {"enabled": True}, "merge": {"enabled": True}}}, (3000, 4000)),
             ({"accumulated_robustness_denoiser": {"merge": {"enabled": True}}}, (3000, 4000)),
             ({"block_matching": {"tuning": {"flow_upscale_mode": "cubic"}}}, (3000, 4000)),
             ({"block_matching": {"tuning": {"flow_upscale_mode": "bilinear"}}}, (3000, 4000)),
             ({"ica": {"tuning": {"n_iter": 0}}}, (3000, 4000)), ({"mode": "rgb"}, (3000, 4000)),
             ({"block_matching": {"tuning": {"tile_size": 16}}}, (700, 740)), ({"block_matching": {"tuning": {"tile_size": 64}}}, (700, 740))]
    for over, shape in cases:
        c = Cfg.wrap(copy.deepcopy(base))
        merge(c, over)
        try:
            P.update_snr_config(c, 30.0)
            P.sanitize_config(c, shape)
            res = "ok"
        except Exception as e:      # noqa: BLE001 - the exception type is the recorded behaviour
            res = type(e).__name__
        out["sanitize"].append({"over": over, "shape": list(shape), "result": res})
    json.dump(out, open(os.path.join(HERE, "params_cases.json"), "w"), indent=1)
    print("wrote params_cases.json:", len(out["snr"]), "snr cases,", [c["result"] for c in out["sanitize"]])


if __name__ == "__main__":
    main()
