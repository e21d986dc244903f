#!/usr/bin/env python
"""Device time per stage (cattus_b200_time_stage: CUDA events, L2 flushed between iterations) at small batch sizes --
where a single leaf's latency goes (exploration tool; the judged numbers come from bench.py)."""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="chess10x128")
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 4, 8, 64, 256])
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()

    from cattus_b200 import CudaNetwork
    from cattus_b200.export import export_blob
    from oracle import net
    from tests.util import synth_inputs

    cfg = net.CONFIGS[args.game]
    words, bitmaps, _ = synth_inputs(args.game, max(args.batch), 1)
    with CudaNetwork(export_blob(net.make_state_dict(cfg, 0), cfg.game), cfg.game, batch_size=max(64, max(args.batch)), n_streams=1) as nw:
        for n in args.batch:
            nw.resident_upload(words[:n], None if bitmaps is None else bitmaps[:n])
            row = []
            for stage, name in ((1, "trunk"), (2, "heads"), (3, "tail"), (4, "all")):
                ms = nw.time_stage(stage, n, args.iters)
                row.append(f"{name} {float(np.median(ms)) * 1000:7.1f} us")
            print(f"{args.game} n={n:5d}: " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
