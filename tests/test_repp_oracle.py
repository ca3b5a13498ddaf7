"""CPU test: oracle/repp_oracle.py (numpy restatement of tools/REPP.py) against the outputs of the reference's own REPP class
(tests/golden/repp.json, tools/make_goldens_host.py) -- identical predictions, floats bit for bit."""
import json
import os

import numpy as np

from conftest import GOLDEN


def typed_video(v):
    """Detections typed like Predictor.to_repp_heavy's (numpy float32 scalars), as in the golden generator."""
    return {k: [{"image_id": p["image_id"], "bbox": [x if isinstance(x, int) else np.float32(x) for x in p["bbox"]],
                 "bbox_center": [np.float32(x) for x in p["bbox_center"]], "scores": np.asarray(p["scores"], dtype=np.float32)}
                for p in plist] for k, plist in v.items()}


def load():
    return json.load(open(os.path.join(GOLDEN, "repp.json")))


def test_repp_oracle_matches_reference_outputs():
    from oracle import repp_oracle
    g = load()
    for case in g["cases"]:
        o = repp_oracle.REPPOracle(logreg=g["logreg"], **case["cfg"])
        out = o(typed_video(case["video"]))
        assert len(out) == len(case["out"]) > 50
        for a, b in zip(out, case["out"]):
            assert a == b, (a, b)
