"""Golden vectors for the format helpers from the UNMODIFIED reference utils.py.  Build container only:

    python tests/golden/make_golden_formats.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SOVITS_REF_DIR", "/root/reference")
for m in ("faiss", "librosa", "matplotlib", "matplotlib.pylab"):
    sys.modules[m] = types.ModuleType(m)
sys.path.insert(0, REF)
import utils as ref_utils  # noqa: E402

g = torch.Generator().manual_seed(4321)
out = {}
cases = [(7, 12), (50, 86), (431, 862), (20, 20), (30, 17), (3, 100)]      # (src_len, target_len): up, equal, down
for i, (s, t) in enumerate(cases):
    x = torch.randn((5, s), generator=g)
    out[f"re_in_{i}"] = x.numpy()
    out[f"re_left_{i}"] = ref_utils.repeat_expand_2d(x, t).numpy()
    out[f"re_nearest_{i}"] = ref_utils.repeat_expand_2d(x, t, "nearest").numpy()
    out[f"re_linear_{i}"] = ref_utils.repeat_expand_2d(x, t, "linear").numpy()
out["re_cases"] = np.array(cases)
for i, n in enumerate((512 * 9 + 100, 44100, 1000)):
    a = torch.randn((1, n), generator=g) * 0.3
    out[f"vol_in_{i}"] = a.numpy()
    out[f"vol_out_{i}"] = ref_utils.Volume_Extractor(512).extract(a).numpy()
np.savez_compressed(os.path.join(HERE, "ref_formats.npz"), **out)
print("wrote", os.path.join(HERE, "ref_formats.npz"), {k: v.shape for k, v in out.items() if k.startswith(("re_left", "vol_out"))})
