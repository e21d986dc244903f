#!/usr/bin/env python
"""Prints the per-layer clock64 trace of trunk_small's CTA 0 (diagnostic; CATTUS_B200_TRACE_TRUNK=1)."""
import os, sys
from pathlib import Path
os.environ["CATTUS_B200_TRACE_TRUNK"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cattus_b200 import CudaNetwork
from cattus_b200.export import export_blob
from oracle import games, net
name, n = sys.argv[1], int(sys.argv[2])
cfg = net.CONFIGS[name]
words, _ = games.synth_hex_positions(n, cfg.board_size, 1)
with CudaNetwork(export_blob(net.make_state_dict(cfg, 0), cfg.game), cfg.game, batch_size=n, n_streams=1) as nw:
    nw.resident_upload(words)
    print("ms", nw.time_stage(1, n, 3))
