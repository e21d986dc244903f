#!/usr/bin/env python
"""Evaluator capacity in the self-play regime: T host threads each keep one batch of b positions in flight through
cattus_b200_eval_batch (host buffers in, host buffers out), no MCTS.  Prints positions/s per (T, b) -- the ceiling the
self-play driver's evaluations/s can be compared with (exploration tool; the judged numbers come from bench.py)."""
from __future__ import annotations

import argparse
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="chess10x128")
    ap.add_argument("--threads", type=int, nargs="+", default=[16, 32])
    ap.add_argument("--batch", type=int, nargs="+", default=[216, 296, 430, 592, 650, 888, 1290, 4096])
    ap.add_argument("--seconds", type=float, default=1.5)
    args = ap.parse_args()

    from cattus_b200 import CudaNetwork
    from cattus_b200.export import export_blob
    from oracle import net
    from tests.util import synth_inputs

    cfg = net.CONFIGS[args.game]
    blob = export_blob(net.make_state_dict(cfg, 0), cfg.game)
    words, bitmaps, _ = synth_inputs(args.game, max(args.batch), 1)
    for th in args.threads:
        with CudaNetwork(blob, cfg.game, batch_size=max(args.batch), n_streams=min(32, th)) as nw:
            for b in args.batch:
                w = words[:b]
                bm = None if bitmaps is None else bitmaps[:b]
                nw.eval_batch(w, bm)  # builds the graph of this bucket
                counts = [0] * th
                stop = time.perf_counter() + args.seconds

                def work(i):
                    while time.perf_counter() < stop:
                        nw.eval_batch(w, bm)
                        counts[i] += 1

                ts = [threading.Thread(target=work, args=(i,)) for i in range(th)]
                t0 = time.perf_counter()
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
                dt = time.perf_counter() - t0
                print(json.dumps({"threads": th, "batch": b, "positions_per_sec": round(sum(counts) * b / dt), "batches_per_sec": round(sum(counts) / dt)}), flush=True)


if __name__ == "__main__":
    main()
