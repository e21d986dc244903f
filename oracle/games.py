"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the integer/byte half of the Cattus NN-evaluation path.

This module is a plain numpy / pure-Python restatement of what the reference does on the host
around `Model::run`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
/ `--impl reference` legs may import it; the product path (`cattus_b200/`) never does.

Every function cites the reference lines it restates (paths relative to /root/reference):

* planes_to_tensor            engine/src/net/mod.rs:121-156
* clamp_non_finite            engine/src/net/mod.rs:57-61
* calc_moves_probs            engine/src/net/mod.rs:106-119
* flip_pos / flip_score       engine/src/net/mod.rs:158-182
* hex bitboard / planes       engine/src/hex/core.rs:52-110, :297-305, :324-334; engine/src/hex/net.rs:14-24
* ttt bitboard / planes       engine/src/ttt/core.rs:60-97; engine/src/ttt/net.rs:14-24
* chess planes                engine/src/chess/net/mod.rs:19-60; engine/src/chess/core.rs:105-153
* chess move <-> nn index     engine/src/chess/core.rs:55-72, :93-95, :453-605
* packed wire formats         training/self-play/src/serialize/{hex.rs:16-28,chess.rs:18-57}

Parity pinning: tests/test_oracle_golden.py checks this file against (a) the reference's own
`DataSet.unpack_planes` run in the build container (tests/golden/encode_ref.npz, produced by
oracle/gen_golden.py), (b) the hand-derived fixture words of SURVEY.md Appendix B, and (c) the
SHA-256 of the reference's 1880-entry NN_INDEX_TO_MOVE list.
"""
from __future__ import annotations

import math
import struct
from typing import Iterable, List, Sequence, Tuple

import numpy as np

F32_MIN = np.float32(-3.4028234663852886e38)  # Rust f32::MIN

# --------------------------------------------------------------------------------------------
# Bitboards -> dense tensor
# --------------------------------------------------------------------------------------------


def words_per_plane(board_size: int) -> int:
    """u64 words the C ABI uses per plane: hex u128 -> [lo, hi] (serialize/hex.rs:16-28); chess u64; ttt u16 -> 1."""
    return (board_size * board_size + 63) // 64


def planes_to_tensor(samples: Sequence[Sequence[int]], batch_size: int, board_size: int) -> np.ndarray:
    """t[b,c,h,w] = bit(plane_c, h*S+w) ? 1.0 : 0.0; rows >= len(samples) are zero (net/mod.rs:121-156)."""
    assert 1 <= len(samples) <= batch_size, f"invalid sample len {len(samples)}, 1..={batch_size}"
    planes_num = len(samples[0])
    s = board_size
    t = np.empty((batch_size, planes_num, s, s), dtype=np.float32)
    for b, sample in enumerate(samples):
        for c, plane in enumerate(sample):
            for h in range(s):
                for w in range(s):
                    t[b, c, h, w] = 1.0 if (int(plane) >> (h * s + w)) & 1 else 0.0
    t[len(samples):] = 0.0
    return t


def planes_to_tensor_fast(words: np.ndarray, board_size: int, planes_num: int) -> np.ndarray:
    """Vectorised equivalent of planes_to_tensor for packed u64 words [n, planes*wpp] (same bit rule)."""
    n = words.shape[0]
    wpp = words_per_plane(board_size)
    w = np.ascontiguousarray(words, dtype="<u8").reshape(n, planes_num, wpp)
    bits = np.unpackbits(w.view(np.uint8).reshape(n, planes_num, wpp * 8), axis=-1, bitorder="little")
    s2 = board_size * board_size
    return bits[..., :s2].reshape(n, planes_num, board_size, board_size).astype(np.float32)


def pack_planes(samples: Sequence[Sequence[int]], board_size: int) -> np.ndarray:
    """Python-int bitboards -> u64 words [n, planes*wpp], little end first (serialize/hex.rs:16-28)."""
    wpp = words_per_plane(board_size)
    out = np.zeros((len(samples), len(samples[0]) * wpp), dtype=np.uint64)
    for b, sample in enumerate(samples):
        for c, plane in enumerate(sample):
            for k in range(wpp):
                out[b, c * wpp + k] = (int(plane) >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


# --------------------------------------------------------------------------------------------
# Net output post-processing
# --------------------------------------------------------------------------------------------


def clamp_non_finite(scores: np.ndarray) -> np.ndarray:
    """Non-finite logit -> f32::MIN (net/mod.rs:57-61)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    s[~np.isfinite(s)] = F32_MIN
    return s


def calc_moves_probs(legal_nn_idx: Sequence[int], move_scores: np.ndarray) -> np.ndarray:
    """Softmax over the legal indices only, f32, sequential sum (net/mod.rs:106-119).

    Returns probabilities in the order of `legal_nn_idx`.
    """
    ms = np.asarray(move_scores, dtype=np.float32)
    sc = [np.float32(ms[i]) for i in legal_nn_idx]
    max_p = F32_MIN
    for p in sc:
        max_p = max(max_p, p)
    with np.errstate(over="ignore", under="ignore"):
        ex = [np.float32(np.exp(np.float32(p - max_p), dtype=np.float32)) for p in sc]
    p_sum = np.float32(0.0)
    for e in ex:
        p_sum = np.float32(p_sum + e)
    return np.array([np.float32(e / p_sum) for e in ex], dtype=np.float32)


def legal_from_bitmap(bitmap: bytes | np.ndarray, moves_num: int) -> List[int]:
    """Ascending nn indices set in a little-bit-order bitmap (serialize/chess.rs:34-41 writes the same layout)."""
    b = np.frombuffer(bytes(bitmap), dtype=np.uint8) if not isinstance(bitmap, np.ndarray) else bitmap.astype(np.uint8)
    bits = np.unpackbits(b, bitorder="little")[:moves_num]
    return [int(i) for i in np.nonzero(bits)[0]]


def bitmap_from_legal(legal_nn_idx: Iterable[int], moves_num: int) -> np.ndarray:
    out = np.zeros(((moves_num + 7) // 8,), dtype=np.uint8)
    for i in legal_nn_idx:
        out[i // 8] |= np.uint8(1 << (i % 8))
    return out


def evaluate_from_net_output(legal_nn_idx: Sequence[int], logits_row: np.ndarray, value: float, flipped: bool):
    """run_net tail + evaluate_impl tail + flip_score_if_needed (net/mod.rs:51-64, :99-101, :166-182).

    Move un-flipping is game specific and done by the caller (it is an index remap)."""
    probs = calc_moves_probs(legal_nn_idx, clamp_non_finite(logits_row))
    v = np.float32(value)
    return probs, (np.float32(-v) if flipped else v)


# --------------------------------------------------------------------------------------------
# Hex (engine/src/hex/core.rs)
# --------------------------------------------------------------------------------------------


def hex_full(board_size: int) -> int:
    """HexBitboard::full(true): low S^2 bits (hex/core.rs:84-93)."""
    return (1 << (board_size * board_size)) - 1


def hex_flip_bitboard(bb: int, board_size: int) -> int:
    """HexBitboard::flip: transpose, idx r*S+c -> c*S+r (hex/core.rs:61-71)."""
    s = board_size
    f = 0
    for r in range(s):
        for c in range(s):
            if (bb >> (r * s + c)) & 1:
                f |= 1 << (c * s + r)
    return f


def hex_move_flipped(idx: int, board_size: int) -> int:
    """HexMove::flipped: (r,c) -> (c,r) (hex/core.rs:36-38)."""
    r, c = divmod(idx, board_size)
    return c * board_size + r


def hex_position_from_str(s: str, board_size: int) -> Tuple[int, int, int]:
    """training/self-play/src/test_util.rs:36-66.  Returns (red, blue, turn) with turn 1 = Player1 (red)."""
    n = board_size * board_size
    assert len(s) == n + 1, "unexpected string length"
    red = blue = 0
    for idx, ch in enumerate(s[:n]):
        if ch == "r":
            red |= 1 << idx
        elif ch == "b":
            blue |= 1 << idx
        elif ch != "e":
            raise ValueError(f"unknown board char: {ch!r}")
    turn = {"r": 1, "b": 2}[s[n]]
    return red, blue, turn


def hex_flip_position(red: int, blue: int, turn: int, board_size: int) -> Tuple[int, int, int]:
    """HexPosition::flipped: red' = flip(blue), blue' = flip(red), turn' = opposite (hex/core.rs:324-334)."""
    return hex_flip_bitboard(blue, board_size), hex_flip_bitboard(red, board_size), 3 - turn


def hex_position_to_planes(red: int, blue: int, board_size: int) -> List[int]:
    """[red, blue, ones] (hex/net.rs:14-24)."""
    return [red, blue, hex_full(board_size)]


def hex_legal_moves(red: int, blue: int, board_size: int) -> List[int]:
    """Ascending empty cells (hex/core.rs:297-305); to_nn_idx is the identity (hex/core.rs:40-42)."""
    occ = red | blue
    return [i for i in range(board_size * board_size) if not (occ >> i) & 1]


def hex_evaluate_inputs(red: int, blue: int, turn: int, board_size: int):
    """flip_pos_if_needed + position_to_planes + legal_moves (net/mod.rs:158-164, :94-100)."""
    flipped = turn != 1
    if flipped:
        red, blue, turn = hex_flip_position(red, blue, turn, board_size)
    return hex_position_to_planes(red, blue, board_size), hex_legal_moves(red, blue, board_size), flipped


# --------------------------------------------------------------------------------------------
# TicTacToe (engine/src/ttt)
# --------------------------------------------------------------------------------------------


def ttt_position_from_str(s: str) -> Tuple[int, int, int]:
    """training/self-play/src/test_util.rs:7-34.  Returns (x, o, turn)."""
    assert len(s) == 10, "unexpected string length"
    x = o = 0
    for idx, ch in enumerate(s[:9]):
        if ch == "x":
            x |= 1 << idx
        elif ch == "o":
            o |= 1 << idx
        elif ch != "_":
            raise ValueError(f"unknown board char: {ch!r}")
    turn = {"x": 1, "o": 2}[s[9]]
    return x, o, turn


def ttt_position_to_planes(x: int, o: int) -> List[int]:
    """[x, o, ones(9 bits)] (ttt/net.rs:14-24, ttt/core.rs:77-81)."""
    return [x, o, (1 << 9) - 1]


# --------------------------------------------------------------------------------------------
# Chess planes from a FEN (engine/src/chess/net/mod.rs:19-60)
# --------------------------------------------------------------------------------------------

_PIECE_PLANE = {"P": 0, "N": 1, "B": 2, "R": 3, "Q": 4, "K": 5, "p": 6, "n": 7, "b": 8, "r": 9, "q": 10, "k": 11}
U64_ALL = 0xFFFFFFFFFFFFFFFF


def chess_planes_from_fen(fen: str, ep_pawn_square: int | None = None) -> List[int]:
    """18 planes: 0-5 white PNBRQK, 6-11 black, 12-15 castle WK,WQ,BK,BQ (all-0/all-1), 16 EP, 17 ones.

    Bit = square = rank*8+file, a1 = 0 (chess/core.rs:135-138).  Plane 16 holds `1 << sq` of crate-`chess`
    `en_passant()`; that crate (chess 3.2.0, not vendored) reports the *capturable pawn's* square and only when
    a capture is actually possible, which cannot be derived from the FEN field alone, so the caller passes it
    (None -> empty plane).  Host side of the ABI; flagged "unverified" in SURVEY.md Appendix D.1.
    """
    fields = fen.split()
    board = fields[0].rstrip("/")
    castle = fields[2] if len(fields) > 2 else "-"
    planes = [0] * 18
    ranks = board.split("/")
    assert len(ranks) == 8, fen
    for i, row in enumerate(ranks):
        rank = 7 - i
        file = 0
        for ch in row:
            if ch.isdigit():
                file += int(ch)
            else:
                planes[_PIECE_PLANE[ch]] |= 1 << (rank * 8 + file)
                file += 1
        assert file == 8, fen
    planes[12] = U64_ALL if "K" in castle else 0
    planes[13] = U64_ALL if "Q" in castle else 0
    planes[14] = U64_ALL if "k" in castle else 0
    planes[15] = U64_ALL if "q" in castle else 0
    planes[16] = 0 if ep_pawn_square is None else 1 << ep_pawn_square
    planes[17] = U64_ALL
    return planes


def chess_flip_planes(planes: Sequence[int]) -> List[int]:
    """ChessPosition::flipped seen through position_to_planes: rank mirror (sq ^ 56), colours and castle
    rights swapped, EP file kept (chess/core.rs:366-399)."""

    def mirror(bb: int) -> int:
        return int.from_bytes(int(bb).to_bytes(8, "little")[::-1], "little")

    out = [0] * 18
    for p in range(6):
        out[p] = mirror(planes[6 + p])
        out[6 + p] = mirror(planes[p])
    out[12], out[13], out[14], out[15] = planes[14], planes[15], planes[12], planes[13]
    out[16] = mirror(planes[16])
    out[17] = planes[17]
    return out


def chess_square_flipped(sq: int) -> int:
    """ChessMove::flipped on one square: rank mirrored, file kept (chess/core.rs:82-91)."""
    return sq ^ 56


# --------------------------------------------------------------------------------------------
# Chess move <-> NN index (engine/src/chess/core.rs:55-72, :453-605)
# --------------------------------------------------------------------------------------------

CHESS_MOVES_NUM = 1880
_PROMO = "qrbn"  # offsets 0..3 (chess/core.rs:61-67)


def _sq_name(sq: int) -> str:
    return "abcdefgh"[sq % 8] + "12345678"[sq // 8]


def chess_nn_index_to_move() -> List[str]:
    """The 1880-entry policy layout, regenerated from its rule instead of copied:
    every (from, to) pair a queen or a knight can travel, from-square ascending then to-square ascending
    (1792 entries), followed by rank-7 -> rank-8 promotions ordered by from-file, to-file, then q, r, b, n
    (22 file pairs x 4).  tests/test_oracle_golden.py pins the SHA-256 of the joined list to the reference's."""
    out: List[str] = []
    for f in range(64):
        fr, ff = divmod(f, 8)
        for t in range(64):
            if t == f:
                continue
            tr, tf = divmod(t, 8)
            dr, df = abs(tr - fr), abs(tf - ff)
            if dr == 0 or df == 0 or dr == df or (dr, df) in ((1, 2), (2, 1)):
                out.append(_sq_name(f) + _sq_name(t))
    for ff in range(8):
        for tf in (ff - 1, ff, ff + 1):
            if 0 <= tf < 8:
                for p in _PROMO:
                    out.append(_sq_name(48 + ff) + _sq_name(56 + tf) + p)
    assert len(out) == CHESS_MOVES_NUM
    return out


def chess_move_to_idx(lan: str) -> int:
    """ChessMove::to_idx (chess/core.rs:55-72)."""
    sf, sr = ord(lan[0]) - 97, ord(lan[1]) - 49
    df, dr = ord(lan[2]) - 97, ord(lan[3]) - 49
    if len(lan) == 5:
        return 64 * 64 + (sf * 2 + df) * 4 + _PROMO.index(lan[4])
    return (sr * 8 + sf) * 64 + (dr * 8 + df)


def chess_move_to_nn_index_table() -> np.ndarray:
    """MOVE_TO_NN_INDEX (chess/core.rs:597-605): u16[64*64 + 22*4], 0xFFFF = no move."""
    res = np.full((64 * 64 + 22 * 4,), 0xFFFF, dtype=np.uint16)
    for idx, m in enumerate(chess_nn_index_to_move()):
        i = chess_move_to_idx(m)
        assert res[i] == 0xFFFF
        res[i] = idx
    return res


# --------------------------------------------------------------------------------------------
# .traindata wire format (next row); used by fixtures to feed the reference parser
# --------------------------------------------------------------------------------------------


def serialize_hex_entry(planes: Sequence[int], probs_dense: np.ndarray, winner: int) -> bytes:
    """u64 LE x 6 | f32 LE x M (-1 illegal) | i8 (self_play.rs:33-61, serialize/hex.rs:16-28)."""
    words = []
    for p in planes:
        words += [int(p) & U64_ALL, (int(p) >> 64) & U64_ALL]
    return struct.pack(f"<{len(words)}Q", *words) + np.asarray(probs_dense, "<f4").tobytes() + struct.pack("<b", winner)


def serialize_chess_entry(planes: Sequence[int], legal_probs: Sequence[Tuple[int, float]], winner: int) -> bytes:
    """18 x u64 | 235 B bitmap | 225 x f32 sorted by nn_idx, -1 pad | i8 (serialize/chess.rs:18-57)."""
    lp = sorted(legal_probs, key=lambda t: t[0])
    bitmap = bitmap_from_legal([i for i, _ in lp], CHESS_MOVES_NUM)
    probs = np.full((225,), -1.0, dtype="<f4")
    for k, (_, p) in enumerate(lp):
        probs[k] = p
    return struct.pack("<18Q", *[int(p) for p in planes]) + bitmap.tobytes() + probs.tobytes() + struct.pack("<b", winner)


# --------------------------------------------------------------------------------------------
# Synthetic positions -- shared by tests and bench so both sides see the same inputs
# --------------------------------------------------------------------------------------------


def synth_hex_positions(n: int, board_size: int, seed: int) -> Tuple[np.ndarray, List[List[int]]]:
    """k ~ U[0, S^2-1] stones on distinct cells alternating red/blue, Player1-to-move view.
    Returns (packed planes [n, 6] u64, legal index lists)."""
    rng = np.random.default_rng(seed)
    s2 = board_size * board_size
    samples, legal = [], []
    for _ in range(n):
        k = int(rng.integers(0, s2))
        cells = rng.permutation(s2)[:k]
        red = blue = 0
        for j, c in enumerate(cells):
            if j % 2 == 0:
                red |= 1 << int(c)
            else:
                blue |= 1 << int(c)
        samples.append(hex_position_to_planes(red, blue, board_size))
        legal.append(hex_legal_moves(red, blue, board_size))
    return pack_planes(samples, board_size), legal


def synth_chess_positions(n: int, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """2 kings + random men on distinct squares (no pawns on ranks 1/8), Bernoulli castle planes, rare EP bit,
    ones plane; legal bitmap = random subset of the 1880 indices with popcount ~ clamp(N(32,10^2),1,218).
    Returns (planes [n,18] u64, bitmaps [n,235] u8).  Vectorised so bench-sized batches are cheap."""
    rng = np.random.default_rng(seed)
    planes = np.zeros((n, 18), dtype=np.uint64)
    perm = np.argsort(rng.random((n, 64)), axis=1)  # distinct squares per position
    one = np.uint64(1)
    planes[:, 5] = one << perm[:, 0].astype(np.uint64)
    planes[:, 11] = one << perm[:, 1].astype(np.uint64)
    men = rng.integers(6, 31, size=n)
    kinds = rng.integers(0, 10, size=(n, 30))  # piece planes without kings: 0-4 white, 6-10 black
    kinds = np.where(kinds < 5, kinds, kinds + 1)
    for j in range(30):
        sq = perm[:, 2 + j]
        k = kinds[:, j]
        rank = sq // 8
        ok = (j < men) & ~(((k == 0) | (k == 6)) & ((rank == 0) | (rank == 7)))
        rows = np.nonzero(ok)[0]
        planes[rows, k[rows]] |= one << sq[rows].astype(np.uint64)
    castle = rng.random((n, 4)) < 0.5
    planes[:, 12:16] = np.where(castle, np.uint64(U64_ALL), np.uint64(0))
    ep = rng.random(n) < 0.05
    ep_sq = (rng.integers(3, 5, size=n) * 8 + rng.integers(0, 8, size=n)).astype(np.uint64)
    planes[:, 16] = np.where(ep, one << ep_sq, np.uint64(0))
    planes[:, 17] = np.uint64(U64_ALL)
    cnt = np.clip(np.rint(rng.normal(32.0, 10.0, size=n)), 1, 218).astype(np.int64)
    # rank of an independent uniform matrix picks exactly cnt distinct indices per position
    order = np.argsort(rng.random((n, CHESS_MOVES_NUM)), axis=1)
    bits = np.zeros((n, CHESS_MOVES_NUM), dtype=np.uint8)
    mask = np.arange(CHESS_MOVES_NUM)[None, :] < cnt[:, None]
    np.put_along_axis(bits, order, mask.astype(np.uint8), axis=1)
    bitmaps = np.packbits(bits, axis=1, bitorder="little")
    assert bitmaps.shape[1] == 235
    return planes, bitmaps


def is_close_outputs(probs1, val1, probs2, val2) -> bool:
    """The reference's own parity predicate (training/tests/test_net_output.py:28-33)."""
    return math.isclose(float(val1), float(val2), rel_tol=1e-5, abs_tol=1e-6) and bool(
        np.isclose(probs1, probs2, rtol=1e-3, atol=1e-6).all()
    )
