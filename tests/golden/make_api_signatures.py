"""Extract the public function signatures of the reference's hot-path modules (parsed with `ast`, nothing is imported)
into tests/golden/api_signatures.json — the drop-in surface tests/test_host_logic.py::test_api_surface checks against.

    python tests/golden/make_api_signatures.py            # needs /root/reference (build container only)
"""
import ast
import json
import os

REF = "/root/reference/handheld_super_resolution"
MODULES = ["super_resolution", "alignment", "block_matching", "ICA", "kernels", "robustness", "merge", "utils", "utils_image",
           "params"]
# launchers / host functions on the path; device kernels (@cuda.jit) are implementation details
WANTED = {
    "super_resolution": ["main", "process"],
    "alignment": ["init_alignment", "build_gaussian_pyramid", "align", "align_lvl", "upscale_lvl"],
    "block_matching": ["align_lvl_block_matching_L2", "align_lvl_block_matching_L1"],
    "ICA": ["init_ica", "align_lvl_ica"],
    "kernels": ["estimate_kernels"],
    "robustness": ["init_robustness", "compute_robustness", "compute_guide_image", "compute_local_stats", "upscale_warp_stats",
                   "local_min"],
    "merge": ["merge", "merge_ref"],
    "utils": ["divide", "add"],
    "utils_image": ["compute_grey_images", "GAT", "cuda_downsample", "computeRMSE", "computePSNR"],
    "params": ["sanitize_config", "update_snr_config"],
}


def main():
    out = {}
    for m in MODULES:
        tree = ast.parse(open(os.path.join(REF, m + ".py")).read())
        sigs = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in WANTED[m]:
                a = node.args
                sigs[node.name] = {"args": [x.arg for x in a.args], "n_defaults": len(a.defaults)}
        missing = [f for f in WANTED[m] if f not in sigs]
        assert not missing, (m, missing)
        out[m] = sigs
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")
    json.dump(out, open(p, "w"), indent=1, sort_keys=True)
    print("wrote", p)


if __name__ == "__main__":
    main()
