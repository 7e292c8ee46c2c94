"""Diagnostic: medium golden burst through main(); where does the output differ most from the B200 golden?"""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "handheld-multi-frame-super-resolution_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import attr_cfg, load
from handheld_super_resolution import main
import handheld_super_resolution.super_resolution as SR
m = load("medium_pipeline.npz")
burst = m["burst_u16"].astype(np.float32) / np.float32(16383.0)
cfg = attr_cfg(m["cfg_json"])
keep = {}
orig_ref = SR.merge_ref
def spy_ref(ref_img, covs, num, den, *a, **k):
    keep["num_comp"], keep["den_comp"] = num.clone(), den.clone()
    k2 = dict(k); k2["fuse_divide"] = False
    n2, d2 = num.clone(), den.clone()
    orig_ref(ref_img, covs, n2, d2, *a, **k2)
    keep["num_final"], keep["den_final"] = n2, d2
    return orig_ref(ref_img, covs, num, den, *a, **k)
SR.merge_ref = spy_ref
out, _ = main(burst[0], burst[1:], cfg)
out = out.cpu().numpy()
H, W = out.shape[:2]; size = 48
cy, cx = (H - size) // 2, (W - size) // 2
sl = dict(tl=(slice(0, size), slice(0, size)), tr=(slice(0, size), slice(W - size, W)), bl=(slice(H - size, H), slice(0, size)),
          br=(slice(H - size, H), slice(W - size, W)), c=(slice(cy, cy + size), slice(cx, cx + size)))
for k, s in sl.items():
    got, want = out[s], m["out__" + k]
    d = np.abs(np.nan_to_num(got) - np.nan_to_num(want))
    i = np.unravel_index(np.argmax(d), d.shape)
    nf, df = keep["num_final"].cpu().numpy()[s], keep["den_final"].cpu().numpy()[s]
    print(k, "max", d.max(), "at", i, "got", got[i], "want", want[i], "| num got/want", nf[i], m["num_final__" + k][i],
          "| den got/want", df[i], m["den_final__" + k][i], "| count>1e-5:", int((d > 1e-5).sum()))
