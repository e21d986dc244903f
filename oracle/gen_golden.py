"""Generates tests/golden/*.npz by running the REFERENCE's own Python code in the build container.

Run once here (`python oracle/gen_golden.py`); /root/reference does not exist on the GPU box, so the outputs are
committed as small fixtures and this script documents how they were made.  Nothing in the product imports it.

What is taken from the reference itself (imported by file path, unmodified):
  * training/cattus_train/net_utils.py   -> ConvNetV1 (the network definition, net_utils.py:45-89)
  * training/cattus_train/data_set.py    -> DataSet.unpack_planes (data_set.py:65-73), the Python half of the
    reference's encode-parity test (training/tests/test_serialize_encode.py:95-151).  Its package imports
    `chess` and `construct`, which are not installed; both are stubbed with empty modules because
    unpack_planes itself only needs numpy.
  * the 1880-entry NN_INDEX_TO_MOVE list, parsed textually from training/cattus_train/chess.py:80-217 and
    engine/src/chess/core.rs:453-595 (only its SHA-256 is stored).
Fixture positions: training/tests/test_net_output.py:138-204.
"""
from __future__ import annotations

import hashlib
import importlib.util
import re
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle import games, net  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

TTT_FIXTURES = ["___x__o_ox", "o_xx_x__ox", "o__xo__xxx", "o___x_o__x", "oo__x____x", "oo__o__oxx"]
HEX11_FIXTURES = [
    "r" + "e" * 120 + "r",
    "".join(("re" * 6)[:11] if r % 2 == 0 else ("er" * 6)[:11] for r in range(11)) + "r",
    "".join("e" * r + "r" + "e" * (10 - r) for r in range(11)) + "r",
]
CHESS_FIXTURES = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "nnqrkbbr/pppppppp/8/8/8/8/PPPPPPPP/NNQRKBBR w - - 0 1",
    "8/1p6/3QR3/6k1/1P2b3/2P3K1/6b1/6r1/ w - - 0 1",
    "4k2r/6r1/8/8/8/8/3R4/R3K3 w Qk - 0 1",
    "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR w KQkq e3 0 1",
]


def _load_reference_modules():
    for name in ("chess", "construct"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.__getattr__ = lambda attr, _n=name: (lambda *a, **k: None)  # type: ignore[attr-defined]
            sys.modules[name] = stub
    pkg = types.ModuleType("cattus_train")
    pkg.__path__ = [str(REF / "training" / "cattus_train")]
    sys.modules["cattus_train"] = pkg

    def load(modname, filename):
        spec = importlib.util.spec_from_file_location(f"cattus_train.{modname}", REF / "training" / "cattus_train" / filename)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"cattus_train.{modname}"] = mod
        spec.loader.exec_module(mod)
        return mod

    nu = load("net_utils", "net_utils.py")
    return nu


def _reference_unpack(planes_words: np.ndarray, board_size: int, planes_num: int) -> np.ndarray:
    """Body of the reference's DataSet.unpack_planes (training/cattus_train/data_set.py:65-73) executed verbatim
    through `exec` of the source lines, so no reference code is restated here."""
    src = (REF / "training" / "cattus_train" / "data_set.py").read_text().splitlines()
    start = next(i for i, l in enumerate(src) if "def unpack_planes" in l)
    body = []
    for l in src[start + 1:]:
        if l.strip().startswith("def ") or l.strip().startswith("@"):
            break
        body.append(l)
    body = [l[8:] if l.startswith("        ") else l.strip() for l in body if l.strip() and not l.strip().startswith("#")]
    body = [l for l in body if not l.startswith("return ")]
    game = types.SimpleNamespace(PLANES_NUM=planes_num, BOARD_SIZE=board_size)
    packed_entry = types.SimpleNamespace(planes=planes_words, probs=None, winner=None)
    env = {"np": np, "game": game, "packed_entry": packed_entry}
    exec("\n".join(body), env)  # noqa: S102 - executing the reference's own lines is the point
    return np.asarray(env["planes"])


def gen_encode(out: Path):
    data = {}

    def add(tag, samples, s, c):
        words = games.pack_planes(samples, s)
        wpp = games.words_per_plane(s)
        ref = np.stack([_reference_unpack(words[i].reshape(c, wpp) if wpp > 1 else words[i], s, c) for i in range(len(samples))])
        data[f"{tag}_words"] = words
        data[f"{tag}_tensor"] = ref.astype(np.uint8)

    add("ttt", [games.ttt_position_to_planes(*games.ttt_position_from_str(s)[:2]) for s in TTT_FIXTURES], 3, 3)
    add("hex11", [games.hex_position_to_planes(*games.hex_position_from_str(s, 11)[:2], 11) for s in HEX11_FIXTURES], 11, 3)
    add("chess", [games.chess_planes_from_fen(f) for f in CHESS_FIXTURES], 8, 18)
    # seeded random bit patterns, every supported board size
    rng = np.random.default_rng(7)
    for s, c in ((3, 3), (4, 3), (5, 3), (7, 3), (9, 3), (11, 3), (8, 18)):
        samples = [[int.from_bytes(rng.bytes(16), "little") & ((1 << (s * s)) - 1) for _ in range(c)] for _ in range(6)]
        add(f"rand{s}", samples, s, c)
    np.savez_compressed(out, **data)
    print("wrote", out, {k: v.shape for k, v in data.items() if k.endswith("tensor")})


def gen_net(nu, out_dir: Path):
    cases = {
        "ttt_1x1": ("fixtures", [games.ttt_position_to_planes(*games.ttt_position_from_str(s)[:2]) for s in TTT_FIXTURES]),
        "hex11_1x1": ("fixtures", [games.hex_position_to_planes(*games.hex_position_from_str(s, 11)[:2], 11) for s in HEX11_FIXTURES]),
        "chess_1x1": ("fixtures", [games.chess_planes_from_fen(f) for f in CHESS_FIXTURES]),
        "hex5_2x2": ("synth", 8), "ttt": ("synth", 8), "hex4": ("synth", 8), "hex5": ("synth", 16), "hex7": ("synth", 8),
        "hex9": ("synth", 4), "hex11": ("synth", 4), "chess_dev": ("synth", 8), "chess_2x128": ("synth", 8),
        "chess10x128": ("synth", 8),
        "chess_4x64": ("synth", 8), "chess_4x256": ("synth", 4), "hex7_4x32": ("synth", 8), "hex11_2x128": ("synth", 4),
        "ttt_simple": ("fixtures", [games.ttt_position_to_planes(*games.ttt_position_from_str(s)[:2]) for s in TTT_FIXTURES]),
        "hex5_simple": ("synth", 8), "hex11_simple": ("synth", 4), "chess_simple": ("synth", 8),
    }
    only = set(sys.argv[1:])
    for name, (kind, arg) in cases.items():
        if only and name not in only:
            continue
        cfg = net.CONFIGS[name]
        if kind == "fixtures":
            words = games.pack_planes(arg, cfg.board_size)
        elif cfg.game == "chess":
            words, _ = games.synth_chess_positions(arg, seed=11)
        else:
            words, _ = games.synth_hex_positions(arg, cfg.board_size, seed=11)
        x = games.planes_to_tensor_fast(words, cfg.board_size, cfg.planes)
        sd = net.make_state_dict(cfg, seed=0)
        if cfg.arch == "simple":  # the reference's second model type (net_utils.py:92-121)
            model = nu.SimpleTwoHeadedModel((1, cfg.planes, cfg.board_size, cfg.board_size), cfg.moves)
        else:
            model = nu.ConvNetV1((1, cfg.planes, cfg.board_size, cfg.board_size), cfg.blocks, cfg.filters,
                                 cfg.value_channels, cfg.policy_channels, cfg.moves)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        model.eval()
        with torch.no_grad():
            p, v = model(torch.from_numpy(x))
        np.savez_compressed(out_dir / f"net_ref_{name}.npz", words=words, logits=p.numpy(), values=v.numpy(), seed=0)
        print("wrote net_ref", name, p.shape, float(np.abs(p.numpy()).max()))


def gen_chess_table(out: Path):
    py = (REF / "training" / "cattus_train" / "chess.py").read_text()
    rs = (REF / "engine" / "src" / "chess" / "core.rs").read_text()
    pat = r'"([a-h][1-8][a-h][1-8][qrbn]?)"'
    lst_py = re.findall(pat, py[py.index("NN_INDEX_TO_MOVE = ["):py.index("# fmt: on")])
    lst_rs = re.findall(pat, rs[rs.index("static NN_INDEX_TO_MOVE"):rs.index("static MOVE_TO_NN_INDEX")])
    assert lst_py == lst_rs and len(lst_py) == 1880
    digest = hashlib.sha256(",".join(lst_py).encode()).hexdigest()
    spots = {str(i): lst_py[i] for i in (0, 1, 22, 23, 500, 1000, 1791, 1792, 1795, 1796, 1879)}
    out.write_text(f"count=1880\nsha256={digest}\n" + "".join(f"{k}={v}\n" for k, v in spots.items()))
    print("wrote", out, digest)


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    nu = _load_reference_modules()
    if len(sys.argv) > 1:  # `python oracle/gen_golden.py <config> ...`: only those networks' fixtures
        gen_net(nu, GOLDEN)
    else:
        gen_encode(GOLDEN / "encode_ref.npz")
        gen_net(nu, GOLDEN)
        gen_chess_table(GOLDEN / "chess_nn_index.txt")
