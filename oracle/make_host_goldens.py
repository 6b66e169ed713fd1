"""ORACLE tooling (build container only): run the reference's own host-side helpers (token alignment, alpha tables,
eta schedule) and write tests/golden/host_logic.json.  Reference functions: modules/utils/seq_aligner.py:100-201,
modules/utils/ptp_utils.py:305-357, modules/inversion/eta_inversion.py:52-58,115-137."""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
os.system = lambda *a, **k: 0
os.chdir(tempfile.mkdtemp(prefix="etai_oracle_"))
sys.path[:0] = [str(REPO / "oracle" / "shim"), str(REPO), "/root/reference"]

from modules.utils import ptp_utils, seq_aligner  # noqa: E402
from modules.inversion.eta_inversion import EtaInversion, _create_eta_func_pow  # noqa: E402
from oracle.sd15 import SyntheticTokenizer  # noqa: E402

tok = SyntheticTokenizer()
PAIRS = [("a cat sitting next to a mirror", "a tiger sitting next to a mirror"),
         ("a photo of a house on a hill", "a photo of a red house on a green hill"),
         ("a dog", "a small brown dog running"),
         ("two birds on a wire at sunset", "two birds on a wire")]
out = {"pairs": PAIRS, "refine": [], "replace": [], "alpha": [], "word_inds": [], "eta": {}}
for s, t in PAIRS:
    m, a = seq_aligner.get_refinement_mapper([s, t], tok)
    out["refine"].append({"mapper": m[0].tolist(), "alphas": a[0].tolist()})
    if len(s.split(" ")) == len(t.split(" ")):
        out["replace"].append(seq_aligner.get_replacement_mapper([s, t], tok)[0].tolist())
    else:
        out["replace"].append(None)
    for crs in (0.8, {"default_": 0.4, t.split(" ")[1]: (0.2, 0.6)}):
        al = ptp_utils.get_time_words_attention_alpha([s, t], 50, dict(crs) if isinstance(crs, dict) else crs, tok)
        out["alpha"].append({"crs": crs if not isinstance(crs, dict) else {k: list(v) if isinstance(v, tuple) else v for k, v in crs.items()},
                             "sum_per_step": al.reshape(51, -1).sum(1).tolist(), "shape": list(al.shape)})
    out["word_inds"].append([ptp_utils.get_word_inds(t, w, tok).tolist() for w in t.split(" ")])
for name, eta in (("linear", (0.0, 0.4)), ("paper", ((0.6, 0.0), (1.0, 0.7))), ("pow2", ((0.2, 0.1), (0.9, 0.6), 2)), ("scalar", 0.3)):
    e = eta if isinstance(eta, tuple) else (eta, eta)
    if len(e) == 3 or isinstance(e[0], tuple):
        f, _ = _create_eta_func_pow(*e)
        v = np.clip(f(np.linspace(0, 1, 1000)), 0, None)
    else:
        v = np.clip(np.linspace(e[0], e[1], 1000), 0, None)
    out["eta"][name] = {"spec": eta, "values_every_37": v[::37].tolist()}
(REPO / "tests" / "golden" / "host_logic.json").write_text(json.dumps(out))
print("wrote host_logic.json")
