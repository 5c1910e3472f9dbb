#!/usr/bin/env python
"""Quick A/B timing of a few sweep points (forward, forward+backward; CUDA-event medians) without the whole sweep.
usage: quick_ab.py [cfg1|cfg2|cfg3|k256d64|c4hw256 ...]"""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import sweep  # noqa: E402

PTS = {
    "cfg1": (1 << 20, 64, 512, 1, 256, "init"),
    "cfg2": (1 << 20, 128, 64, 4, 64, "trained"),
    "cfg3": (1 << 20, 128, 64, 1, 64, "trained"),
    "k256d64": (1 << 20, 64, 256, 1, 256, "trained"),
    "k256d32": (1 << 20, 32, 256, 1, 256, "trained"),
    "k256d128": (1 << 20, 128, 256, 1, 256, "trained"),
    "k1024d64": (1 << 20, 64, 1024, 1, 256, "trained"),
    "c4hw256": (1 << 20, 128, 64, 4, 256, "trained"),
    "c2hw64": (1 << 20, 64, 64, 2, 64, "trained"),
    "cfg3hw256": (1 << 20, 128, 64, 1, 256, "trained"),
    "k4096d128": (1 << 18, 128, 4096, 1, 256, "trained"),
    "k16384d32": (1 << 20, 32, 16384, 1, 256, "trained"),
    "k16384d256": (1 << 16, 256, 16384, 1, 256, "trained"),
    "k1024d256": (1 << 18, 256, 1024, 1, 256, "trained"),
    "k4096d256": (1 << 17, 256, 4096, 1, 256, "trained"),
    "k1024d128": (1 << 19, 128, 1024, 1, 256, "trained"),
    "k16384d128": (1 << 17, 128, 16384, 1, 256, "trained"),
}
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for name in (sys.argv[1:] or list(PTS)):
    r = sweep.point(*PTS[name], dev, flush)
    print(json.dumps({"pt": name, "fwd_ms": round(r["fwd_ms"], 4), "fwd_frac": round(r["fwd_frac"], 3),
                      "fwdbwd_ms": round(r["fwdbwd_ms"], 4), "fwdbwd_frac": round(r["fwdbwd_frac"], 3),
                      "bwd_ms": round(r["fwdbwd_ms"] - r["fwd_ms"], 4), "idx_exact": r["idx_exact"]}), flush=True)
