#!/usr/bin/env python
"""Smallest end-to-end case for `compute-sanitizer --tool memcheck`: a few positions through every production kernel
(trunk_small, trunk_fused, tc_gemm_dual with all three fused epilogues, softmax_compact).  On this pool the sanitizer
is closed by the operators ("runs under it have left GPUs needing a reset"), so the script only serves as the plain
smallest case; memory safety rests on the parity tests (every output element is compared with the oracle)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from cattus_b200 import CudaNetwork  # noqa: E402
from cattus_b200.export import export_blob  # noqa: E402
from oracle import games, net  # noqa: E402

for name, n in (("hex5", 40), ("hex11", 9), ("chess_dev", 12), ("chess_2x128", 6)):
    cfg = net.CONFIGS[name]
    sd = net.make_state_dict(cfg, 0)
    if cfg.game == "chess":
        words, bitmaps = games.synth_chess_positions(n, 1)
    else:
        words, _ = games.synth_hex_positions(n, cfg.board_size, 1)
        bitmaps = None
    with CudaNetwork(export_blob(sd, cfg.game), cfg.game, batch_size=64, n_streams=1) as nw:
        probs, offsets, values = nw.eval_batch(words, bitmaps)
        p1, v1 = nw.eval_planes(words[0], None if bitmaps is None else bitmaps[0])
    assert np.isfinite(probs).all() and np.isfinite(values).all() and np.array_equal(p1, probs[offsets[0]:offsets[1]])
    print(name, "ok", len(probs))
