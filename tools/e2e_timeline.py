#!/usr/bin/env python
"""Host-side timeline of one end-to-end eval_batch call of several device batches (CATTUS_B200_TRACE_BATCH=1 makes the
library print a stamp per pipeline event on stderr) -- where `e2e` loses against `value` (exploration tool)."""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="chess10x128")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--streams", type=int, default=2)
    args = ap.parse_args()
    from cattus_b200 import CudaNetwork
    from cattus_b200.export import export_blob
    from oracle import net
    from tests.util import synth_inputs

    cfg = net.CONFIGS[args.game]
    n = args.batch * args.batches
    words, bitmaps, _ = synth_inputs(args.game, n, 1)
    with CudaNetwork(export_blob(net.make_state_dict(cfg, 0), cfg.game), cfg.game, batch_size=args.batch, n_streams=args.streams) as nw:
        for _ in range(3):
            nw.eval_batch(words, bitmaps)
        print("---- timed call", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        nw.eval_batch(words, bitmaps)
        dt = time.perf_counter() - t0
        print(f"{args.game}: {n} positions in {dt * 1e3:.2f} ms = {n / dt / 1e6:.3f} M positions/s ({args.streams} streams)")


if __name__ == "__main__":
    main()
