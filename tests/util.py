"""Shared helpers for the parity tests: build evaluators for the oracle's named configurations."""
from __future__ import annotations

from functools import lru_cache
from pathlib import Path

import numpy as np

from oracle import games, net

GOLDEN = Path(__file__).resolve().parent / "golden"


@lru_cache(maxsize=None)
def state_dict(name: str, seed: int = 0):
    return net.make_state_dict(net.CONFIGS[name], seed)


@lru_cache(maxsize=None)
def blob(name: str, seed: int = 0) -> bytes:
    from cattus_b200.export import export_blob

    return export_blob(state_dict(name, seed), net.CONFIGS[name].game)


def make_network(name: str, precision: str = "bf16", batch_size: int = 64, n_streams: int = 2, cache=None, fused_trunk: bool = True):
    from cattus_b200 import CudaNetwork

    return CudaNetwork(blob(name), net.CONFIGS[name].game, batch_size=batch_size, n_streams=n_streams, precision=precision, cache=cache,
                       fused_trunk=fused_trunk)


def synth_inputs(name: str, n: int, seed: int):
    """(words [n, planes*wpp] u64, bitmaps [n, 235] u8 or None, legal index lists)."""
    cfg = net.CONFIGS[name]
    if cfg.game == "chess":
        words, bitmaps = games.synth_chess_positions(n, seed)
        legal = [games.legal_from_bitmap(bitmaps[i], cfg.moves) for i in range(n)]
        return words, bitmaps, legal
    if cfg.game == "ttt":
        rng = np.random.default_rng(seed)
        samples, legal = [], []
        for _ in range(n):
            k = int(rng.integers(0, 9))
            cells = rng.permutation(9)[:k]
            x = sum(1 << int(c) for c in cells[0::2])
            o = sum(1 << int(c) for c in cells[1::2])
            samples.append(games.ttt_position_to_planes(x, o))
            legal.append([i for i in range(9) if not ((x | o) >> i) & 1])
        return games.pack_planes(samples, 3), None, legal
    words, legal = games.synth_hex_positions(n, cfg.board_size, seed)
    return words, None, legal


def oracle_eval(name: str, words: np.ndarray, legal, threads=None):
    """Oracle end to end: planes_to_tensor -> ConvNetV1 forward (fp32) -> clamp -> calc_moves_probs."""
    cfg = net.CONFIGS[name]
    x = games.planes_to_tensor_fast(words, cfg.board_size, cfg.planes)
    logits, values = net.convnet_forward(state_dict(name), cfg, x, threads=threads)
    probs = [games.calc_moves_probs(legal[i], games.clamp_non_finite(logits[i])) for i in range(len(legal))]
    return logits, values.reshape(-1), probs
