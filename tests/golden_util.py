"""Loader for tests/golden/jump_golden.json (generated from the reference by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from oracle import kgo

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jump_golden.json")


def load_cases():
    with open(PATH) as f:
        data = json.load(f)
    out = []
    for c in data["cases"]:
        c = dict(c)
        c["dp_mask"] = int(c["dp_mask"], 16)
        c["key"] = (int(c["key"][0], 16), int(c["key"][1], 16))
        c["width_div2"] = int(c["width_div2"], 16)
        jt = c["jump_table"]
        c["table"] = (kgo.ints_to_array([int(r[0], 16) for r in jt], 2), kgo.ints_to_array([int(r[1], 16) for r in jt]),
                      kgo.ints_to_array([int(r[2], 16) for r in jt]))
        for k in ("start", "end"):
            c[k] = [tuple(int(v, 16) for v in row) for row in c[k]]
        c["dps"] = [(int(x, 16), int(d, 16), k, j) for x, d, k, j in c["dps"]]
        out.append(c)
    return out


def arrays(rows):
    return (kgo.ints_to_array([r[0] for r in rows]), kgo.ints_to_array([r[1] for r in rows]),
            kgo.ints_to_array([r[2] for r in rows]))
