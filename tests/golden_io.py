"""Loader for tests/golden/case_*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"))
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        self.que = {k[4:]: t(z[k]) for k in z.files if k.startswith("que_") and k not in ("que_depth", "que_depth_fine")}
        self.ref = {k[4:]: t(z[k]) for k in z.files if k.startswith("ref_")}
        self.W = {k[2:]: t(z[k]) for k in z.files if k.startswith("W_")}
        self.cfg = json.loads(bytes(z["cfg_json"]).decode())
        self.is_train = bool(int(z["is_train"]))
        self.out = {k[4:]: t(z[k]) for k in z.files if k.startswith("out_")}
        self.que_depth = t(z["que_depth"])
        self.que_depth_fine = t(z["que_depth_fine"])
        self.fine_u = t(z["fine_u"]) if "fine_u" in z.files else None
        self.stage_sel = t(z["stage_sel"])
        self.stage = {"c": {}, "f": {}}
        for k in z.files:
            if k.startswith("stage_c_"):
                self.stage["c"][k[8:]] = t(z[k])
            elif k.startswith("stage_f_"):
                self.stage["f"][k[8:]] = t(z[k])

    def flat_cfg(self):
        from gen_golden import flat_cfg
        return flat_cfg(self.cfg)

    def stage_que(self):
        q = dict(self.que)
        q["coords"] = self.que["coords"][:, self.stage_sel].contiguous()
        return q
