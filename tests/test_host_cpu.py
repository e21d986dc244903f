"""CPU: host logic of the product (exporter, position types, cache, library surface).  No compute calls."""
import ctypes as C
import re
import struct
from pathlib import Path

import numpy as np
import pytest

from oracle import games as ogames
from oracle import net
from tests.util import blob, state_dict

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_every_declared_symbol():
    from cattus_b200 import _lib

    lib = _lib.load()
    header = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    declared = set(re.findall(r"\b(cattus_b200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cattus_b200_abi_version() == 1
    assert C.sizeof(_lib.Desc) == 64 and C.sizeof(_lib.Metrics) == 48 and C.sizeof(_lib.Info) == 64


def test_create_fails_loudly_without_a_device_or_with_a_bad_blob():
    from cattus_b200 import CudaNetwork, _lib
    from tests.conftest import HAS_GPU

    with pytest.raises(_lib.CattusB200Error) as ei:
        CudaNetwork(b"\0" * 128, "hex")
    assert ei.value.code == _lib.EINVAL and "magic" in str(ei.value)
    good = blob("hex4")
    with pytest.raises(_lib.CattusB200Error) as ei:
        CudaNetwork(good[:-4], "hex")
    assert ei.value.code == _lib.EINVAL
    with pytest.raises(_lib.CattusB200Error) as ei:
        CudaNetwork(good, "chess")  # game id mismatch
    assert ei.value.code == _lib.EINVAL
    if not HAS_GPU:
        with pytest.raises(_lib.CattusB200Error) as ei:
            CudaNetwork(good, "hex")
        assert ei.value.code == _lib.ENODEV and "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("name", ["hex5_2x2", "hex4", "chess_dev", "ttt"])
def test_export_folds_batchnorm_exactly(name):
    """A plain conv + bias network rebuilt from the blob must reproduce the oracle's conv + BN forward."""
    import torch
    import torch.nn.functional as F

    cfg = net.CONFIGS[name]
    b = blob(name)
    hdr = struct.unpack("<16I", b[:64])
    assert hdr[0] == 0x00324243 and hdr[1] == 1
    assert hdr[2:11] == ({"ttt": 0, "hex": 1, "chess": 2}[cfg.game], cfg.board_size, cfg.planes, cfg.moves, cfg.filters, cfg.blocks,
                         cfg.value_channels, cfg.policy_channels, 128)
    data = np.frombuffer(b, dtype="<f4", offset=64)
    pos = 0

    def take(*shape):
        nonlocal pos
        cnt = int(np.prod(shape))
        a = torch.from_numpy(data[pos:pos + cnt].reshape(shape).copy())
        pos += cnt
        return a

    s2, f = cfg.board_size ** 2, cfg.filters
    words, _ = ogames.synth_chess_positions(4, 3) if cfg.game == "chess" else (ogames.synth_hex_positions(4, cfg.board_size, 3) if cfg.game == "hex"
                                                                                else (np.array([[8, 0x140, 0x1FF]] * 4, dtype=np.uint64), None))
    x = torch.from_numpy(ogames.planes_to_tensor_fast(words, cfg.board_size, cfg.planes))
    w, bias = take(f, cfg.planes, 3, 3), take(f)
    flow = F.relu(F.conv2d(x, w, bias, padding=1))
    for _ in range(cfg.blocks):
        w1, b1, w2, b2 = take(f, f, 3, 3), take(f), take(f, f, 3, 3), take(f)
        h = F.relu(F.conv2d(flow, w1, b1, padding=1))
        flow = F.relu(flow + F.conv2d(h, w2, b2, padding=1))
    vw, vb = take(cfg.value_channels, f, 1, 1), take(cfg.value_channels)
    v = F.relu(F.conv2d(flow, vw, vb)).flatten(1)
    v = F.relu(F.linear(v, take(128, cfg.value_channels * s2), take(128)))
    v = torch.tanh(F.linear(v, take(1, 128), take(1)))
    pw, pb = take(cfg.policy_channels, f, 1, 1), take(cfg.policy_channels)
    p = F.linear(F.relu(F.conv2d(flow, pw, pb)).flatten(1), take(cfg.moves, cfg.policy_channels * s2), take(cfg.moves))
    assert pos == data.size
    logits, values = net.convnet_forward(state_dict(name), cfg, x.numpy())
    np.testing.assert_allclose(p.numpy(), logits, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(v.numpy(), values, rtol=2e-4, atol=2e-5)


def test_position_types_match_oracle():
    from cattus_b200.games import ChessPosition, HexPosition, TttPosition

    rng = np.random.default_rng(9)
    for s in (4, 5, 7, 9, 11):
        for _ in range(8):
            cells = rng.permutation(s * s)[: int(rng.integers(0, s * s))]
            red = sum(1 << int(c) for c in cells[0::2])
            blue = sum(1 << int(c) for c in cells[1::2])
            for turn in (1, 2):
                p = HexPosition(s, red, blue, turn)
                planes, legal, flipped = ogames.hex_evaluate_inputs(red, blue, turn, s)
                q = p.flipped() if turn == 2 else p
                assert flipped == (turn == 2)
                assert np.array_equal(q.to_planes_words(), ogames.pack_planes([planes], s)[0])
                assert q.legal_moves() == legal
                assert p.flipped().flipped() == p
                assert all(p.flip_move(p.flip_move(m)) == m for m in range(s * s))
                assert [p.flip_move(m) for m in range(s * s)] == [ogames.hex_move_flipped(m, s) for m in range(s * s)]
    t = TttPosition(*ogames.ttt_position_from_str("o_xx_x__ox"))
    assert t.to_planes_words().tolist() == [0x02C, 0x101, 0x1FF] and t.flipped().to_planes_words().tolist() == [0x101, 0x02C, 0x1FF]
    planes = tuple(ogames.chess_planes_from_fen("4k2r/6r1/8/8/8/8/3R4/R3K3 w Qk - 0 1"))
    c = ChessPosition(planes, (((0, 8, ""), 7),), 2)
    assert list(c.flipped().planes) == ogames.chess_flip_planes(planes) and c.flipped().flipped().planes == planes
    assert ChessPosition.flip_move((12, 28, "")) == (12 ^ 56, 28 ^ 56, "")


def test_value_func_cache_fifo_and_counters():
    from cattus_b200 import ValueFuncCache

    cache = ValueFuncCache(2)
    calls = []

    def mk(k):
        return lambda: calls.append(k) or k * 10

    assert cache.get_or_compute(1, mk(1)) == 10 and cache.get_or_compute(2, mk(2)) == 20
    assert cache.get_or_compute(1, mk(1)) == 10  # hit; FIFO order is insertion order, not recency (cache.rs:60-72)
    assert cache.get_or_compute(3, mk(3)) == 30  # evicts key 1 (oldest insertion)
    assert cache.get_or_compute(2, mk(2)) == 20  # still cached
    assert cache.get_or_compute(1, mk(1)) == 10  # recomputed
    assert calls == [1, 2, 3, 1] and cache.metrics() == {"cache.hits": 2, "cache.misses": 4}


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under cattus_b200/ (Python or C++) may import, include, link or execute
    it, and the only other users are tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs."""
    pkg = ROOT / "cattus_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.hpp")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
        assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f
        assert "liboracle" not in text, f
    entry = (ROOT / "__graft_entry__.py").read_text()
    build_fn = entry[entry.index("def build()"):entry.index("def smoke()")]
    assert "import oracle" not in build_fn and "from oracle" not in build_fn  # build() only compiles the checker
