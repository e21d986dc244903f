"""Position types for the host mirror: just enough of the reference's `Position`/`Move`/`Bitboard` traits
(engine/src/game/mod.rs:8-107) for `CudaNetwork.evaluate`: turn, flipped, legal_moves, to_nn_idx, planes.

Hex  : engine/src/hex/core.rs:6-43 (move), :52-110 (bitboard), :297-305 (legal), :324-334 (flipped); hex/net.rs:14-24
Ttt  : engine/src/ttt/core.rs:60-97; ttt/net.rs:14-24
Chess: planes and the legal set come from the caller's move generator (crate `chess` on the Rust side), so
       ChessPosition carries them explicitly; flipping is chess/core.rs:82-91, :366-399.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

_M64 = (1 << 64) - 1


def _words(planes: Sequence[int], wpp: int) -> np.ndarray:
    out = np.empty(len(planes) * wpp, dtype=np.uint64)
    for c, p in enumerate(planes):
        for k in range(wpp):
            out[c * wpp + k] = (p >> (64 * k)) & _M64
    return out


def _transpose(bb: int, s: int) -> int:
    out = 0
    while bb:
        low = bb & -bb
        i = low.bit_length() - 1
        r, c = divmod(i, s)
        out |= 1 << (c * s + r)
        bb ^= low
    return out


@dataclass(frozen=True)
class HexPosition:
    size: int
    red: int
    blue: int
    turn: int  # 1 = Player1 (red), 2 = Player2 (blue)

    def flipped(self) -> "HexPosition":
        return HexPosition(self.size, _transpose(self.blue, self.size), _transpose(self.red, self.size), 3 - self.turn)

    def legal_moves(self) -> List[int]:
        occ = self.red | self.blue
        return [i for i in range(self.size * self.size) if not (occ >> i) & 1]

    @staticmethod
    def move_to_nn_idx(m: int) -> int:
        return m

    def flip_move(self, m: int) -> int:
        r, c = divmod(m, self.size)
        return c * self.size + r

    def to_planes_words(self) -> np.ndarray:
        s2 = self.size * self.size
        return _words([self.red, self.blue, (1 << s2) - 1], (s2 + 63) // 64)

    def key(self):
        return (self.size, self.red, self.blue, self.turn)


@dataclass(frozen=True)
class TttPosition:
    x: int
    o: int
    turn: int

    def flipped(self) -> "TttPosition":
        return TttPosition(self.o, self.x, 3 - self.turn)

    def legal_moves(self) -> List[int]:
        occ = self.x | self.o
        return [i for i in range(9) if not (occ >> i) & 1]

    @staticmethod
    def move_to_nn_idx(m: int) -> int:
        return m

    @staticmethod
    def flip_move(m: int) -> int:
        return m

    def to_planes_words(self) -> np.ndarray:
        return _words([self.x, self.o, 0x1FF], 1)

    def key(self):
        return (self.x, self.o, self.turn)


def _mirror(bb: int) -> int:
    return int.from_bytes(int(bb).to_bytes(8, "little")[::-1], "little")


@dataclass(frozen=True)
class ChessPosition:
    """18 bitboard planes (chess/net/mod.rs:19-60) + the legal moves as (from_sq, to_sq, promo) with their nn indices."""
    planes: Tuple[int, ...]
    legal: Tuple[Tuple[Tuple[int, int, str], int], ...]  # ((from, to, promo), nn_idx) in MoveGen order
    turn: int

    def flipped(self) -> "ChessPosition":
        p = self.planes
        out = [0] * 18
        for k in range(6):
            out[k] = _mirror(p[6 + k])
            out[6 + k] = _mirror(p[k])
        out[12], out[13], out[14], out[15] = p[14], p[15], p[12], p[13]
        out[16] = _mirror(p[16])
        out[17] = p[17]
        # the caller supplies nn indices for the side-to-move view, so the legal list is carried over unchanged
        return ChessPosition(tuple(out), self.legal, 3 - self.turn)

    @staticmethod
    def from_fen(fen: str, moves=()) -> "ChessPosition":
        """ChessPosition::from_fen (chess/core.rs:170-172) followed by `moved_position` for each move (from | to << 6 |
        promotion << 12, real board coordinates), through the library's chess rules (include/cattus_b200_chess.h)."""
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        info = _lib.ChessInfo()
        info.struct_size = C.sizeof(_lib.ChessInfo)
        arr = (C.c_uint16 * max(1, len(moves)))(*moves)
        if lib.cattus_b200_chess_position(fen.encode(), arr, len(moves), C.byref(info)) != 0:
            raise ValueError((lib.cattus_b200_chess_last_error() or b"").decode(errors="replace"))
        promo = (None, "q", "n", "r", "b")
        view = ChessPosition(tuple(int(x) for x in info.planes), (), 1)  # the evaluator's view: side to move plays white
        legal = []
        for k in range(info.n_legal):
            m = int(info.moves[k])
            real = (m & 63, (m >> 6) & 63, promo[m >> 12])
            legal.append((real if info.turn == 1 else ChessPosition.flip_move(real), int(info.nn_index[k])))
        planes = view.planes if info.turn == 1 else view.flipped().planes
        return ChessPosition(planes, tuple(legal), int(info.turn))

    def legal_moves(self):
        return [m for m, _ in self.legal]

    def move_to_nn_idx(self, m) -> int:
        return dict(self.legal)[m]

    @staticmethod
    def flip_move(m):
        f, t, promo = m
        return (f ^ 56, t ^ 56, promo)

    def to_planes_words(self) -> np.ndarray:
        return _words(self.planes, 1)

    def key(self):
        return (self.planes, self.turn)
