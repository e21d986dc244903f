"""ctypes wrapper of tests/emul/libdsearch_emul.so (TEST INFRASTRUCTURE): the device-resident search compiled for the
host with a one-lane warp, driven by the same host driver as on the GPU, over a Python evaluator callback."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from cattus_b200 import _lib
from cattus_b200.selfplay import GameRecord, SelfPlayRunner, _eval_thunk, parse_game

HERE = Path(__file__).resolve().parent
SO = HERE / "libdsearch_emul.so"
SRC = HERE / "dsearch_emul.cpp"
CSRC = HERE.parent.parent / "cattus_b200" / "csrc"


def build() -> Path:
    deps = [SRC] + sorted(CSRC.glob("*.hpp")) + sorted((HERE.parent.parent / "include").glob("*.h"))
    if SO.exists() and all(d.stat().st_mtime <= SO.stat().st_mtime for d in deps):
        return SO
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-o", str(SO), str(SRC)],
                   check=True)
    return SO


_cached = None


def load():
    global _cached
    if _cached is None:
        lib = C.CDLL(str(build()))
        lib.dsearch_emul_run.restype = C.c_void_p
        lib.dsearch_emul_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.dsearch_emul_error.restype = C.c_char_p
        lib.dsearch_emul_error.argtypes = [C.c_void_p]
        lib.dsearch_emul_game_count.restype = C.c_uint32
        lib.dsearch_emul_game_count.argtypes = [C.c_void_p]
        lib.dsearch_emul_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.dsearch_emul_game_info.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.dsearch_emul_game_moves.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint16)]
        lib.dsearch_emul_entry.restype = C.c_uint32
        lib.dsearch_emul_entry.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint32)]
        lib.dsearch_emul_free.argtypes = [C.c_void_p]
        _cached = lib
    return _cached


def run(game: str, cfg: dict, eval1, eval2, games_num: int, *, n_slots: int, pool_words: int, depth: int = 2, first_game: int = 0,
        game_stride: int = 1):
    """-> (counters dict, [GameRecord]); raises RuntimeError with the emulation's message on failure."""
    lib = load()
    runner = SelfPlayRunner(game, cfg)
    c = runner._fill(games_num, None, None, True, first_game, game_stride)
    g, s = parse_game(game)
    chess = g == _lib.GAME_CHESS
    words = 18 if chess else 3 * ((s * s + 63) // 64)
    errors: list = []
    t1 = _eval_thunk(eval1, words, chess, errors)
    t2 = _eval_thunk(eval2, words, chess, errors) if eval2 is not None else None
    h = lib.dsearch_emul_run(C.cast(t1, C.c_void_p), None, C.cast(t2, C.c_void_p) if t2 else None, None, C.byref(c), n_slots, pool_words, depth)
    try:
        if errors:
            raise errors[0]
        err = lib.dsearch_emul_error(h)
        if err:
            raise RuntimeError(err.decode())
        cnt = (C.c_uint64 * 8)()
        lib.dsearch_emul_counters(h, cnt)
        counters = dict(zip(("simulations", "evaluations", "terminal", "searches", "cache_hits", "w1", "w2", "draws"), [int(x) for x in cnt]))
        records = []
        for k in range(lib.dsearch_emul_game_count(h)):
            gi, w, nm = C.c_uint32(), C.c_uint32(), C.c_uint32()
            lib.dsearch_emul_game_info(h, k, C.byref(gi), C.byref(w), C.byref(nm))
            mv = (C.c_uint16 * max(1, nm.value))()
            lib.dsearch_emul_game_moves(h, k, mv)
            entries, dirs = [], []
            for pi in range(nm.value):
                d = C.c_uint32()
                nb = lib.dsearch_emul_entry(h, k, pi, None, 0, C.byref(d))
                buf = (C.c_uint8 * nb)()
                lib.dsearch_emul_entry(h, k, pi, buf, nb, C.byref(d))
                entries.append(bytes(buf))
                dirs.append(d.value)
            records.append(GameRecord(gi.value, w.value or None, list(mv)[: nm.value], entries, dirs))
        return counters, records
    finally:
        lib.dsearch_emul_free(h)
