"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the chess rules behind the evaluation path.  Plain Python on a 64-entry
mailbox board -- deliberately NOT the bitboard arrangement of the product (cattus_b200/csrc/chess_rules.hpp), and in
real board colours with explicit flips where the product keeps one side-to-move view.  Only `tests/` and bench.py's
cpu_baseline leg may import it.

Restated from (paths relative to /root/reference):

* ChessMove::to_idx / to_nn_idx / flipped   engine/src/chess/core.rs:55-72, :82-95 (table :453-605 via oracle/games.py)
* ChessPosition                             engine/src/chess/core.rs:155-400 (equality :292-309 ignores the fifty-move
                                            count, moved_position :326-346, status :348-364, flipped :366-399)
* ChessGame (threefold repetition)          engine/src/chess/core.rs:402-451
* position_to_planes                        engine/src/chess/net/mod.rs:19-60
* ChessSerializer                           training/self-play/src/serialize/chess.rs:18-57

Third-party behaviour modelled explicitly -- crate `chess` 3.2.0 (engine/Cargo.lock), NOT in the tree, so parity is
UNPINNED for everything in this list except the legal move SET, which the published perft counts pin
(tests/test_chess_cpu.py):

* `Board::en_passant()` is the square of the pawn that just advanced two ranks, recorded only if an enemy pawn stands
  beside it (`set_ep`); a FEN's en-passant field goes through the same rule (BoardBuilder keeps only the file).
* `MoveGen::new_legal` order: pawns, knights, bishops, rooks, queens, king; inside a piece type the unpinned pieces by
  ascending square, then the pinned ones; en-passant captures after every other pawn move; per piece the destinations
  by ascending square; a promoting destination yields queen, knight, rook, bishop; castling is one of the king's
  destinations.  Here the legal set is found the slow, obviously-correct way (make the move, look whether the king is
  attacked) and then SORTED into that order.
* castle rights are lost when anything leaves e1/a1/h1 (e8/a8/h8) or lands on the opponent's a/h corner.
* `Board::status()`: no legal move => checkmate if in check else stalemate; nothing else ends a game.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import games as og

P1, P2 = 1, 2
Move = Tuple[int, int, Optional[str]]  # (from, to, promotion in "qnrb" or None)

_KNIGHT = ((1, 2), (2, 1), (-1, 2), (-2, 1), (1, -2), (2, -1), (-1, -2), (-2, -1))
_KING = ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1))
_ROOK = ((1, 0), (-1, 0), (0, 1), (0, -1))
_BISHOP = ((1, 1), (1, -1), (-1, 1), (-1, -1))
_TYPE_RANK = {"p": 0, "n": 1, "b": 2, "r": 3, "q": 4, "k": 5}
_PROMO_RANK = {None: 0, "q": 0, "n": 1, "r": 2, "b": 3}
_NN_TABLE = None


def _nn_table():
    global _NN_TABLE
    if _NN_TABLE is None:
        _NN_TABLE = og.chess_move_to_nn_index_table()
    return _NN_TABLE


def _is_white(piece: str) -> bool:
    return piece.isupper()


def _on_board(r: int, f: int) -> bool:
    return 0 <= r < 8 and 0 <= f < 8


def move_to_lan(m: Move) -> str:
    return og._sq_name(m[0]) + og._sq_name(m[1]) + (m[2] or "")


def move_to_u16(m: Move) -> int:
    """The product's move encoding (include/cattus_b200_chess.h)."""
    return m[0] | (m[1] << 6) | ({None: 0, "q": 1, "n": 2, "r": 3, "b": 4}[m[2]] << 12)


def move_from_u16(x: int) -> Move:
    return (x & 63, (x >> 6) & 63, (None, "q", "n", "r", "b")[x >> 12])


class ChessPosition:
    REPETITION_LIMIT = 3  # ChessGame::REPETITION_LIMIT, core.rs:414
    moves_num = og.CHESS_MOVES_NUM

    __slots__ = ("board", "turn", "castle", "ep", "fifty", "_legal", "_hash")

    def __init__(self, board: Sequence[Optional[str]], turn: int, castle: str, ep: Optional[int], fifty: int = 0):
        self.board = tuple(board)
        self.turn = turn
        self.castle = "".join(c for c in "KQkq" if c in castle)
        self.ep = ep  # square of the pawn capturable en passant (crate semantics), or None
        self.fifty = fifty
        self._legal = None
        self._hash = None

    # ---- construction
    @staticmethod
    def new() -> "ChessPosition":
        return ChessPosition.from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -")

    @staticmethod
    def from_fen(fen: str) -> "ChessPosition":
        fields = fen.split()
        board: List[Optional[str]] = [None] * 64
        for i, row in enumerate(fields[0].rstrip("/").split("/")):
            rank, file = 7 - i, 0
            for ch in row:
                if ch.isdigit():
                    file += int(ch)
                else:
                    board[rank * 8 + file] = ch
                    file += 1
            assert file == 8, fen
        turn = P1 if fields[1] == "w" else P2
        castle = fields[2] if len(fields) > 2 and fields[2] != "-" else ""
        ep = None
        if len(fields) > 3 and fields[3] != "-":
            file = ord(fields[3][0]) - 97
            # the pawn that just moved stands on its own fourth rank; kept only if a pawn of the side to move is beside it
            sq = (3 if turn == P2 else 4) * 8 + file
            ep = ChessPosition._ep_if_capturable(board, sq, mover_white=(turn == P2))
        return ChessPosition(board, turn, castle, ep)

    @staticmethod
    def _ep_if_capturable(board, sq: int, mover_white: bool) -> Optional[int]:
        if board[sq] != ("P" if mover_white else "p"):
            return None
        enemy = "p" if mover_white else "P"
        f = sq % 8
        for df in (-1, 1):
            if 0 <= f + df < 8 and board[sq + df] == enemy:
                return sq
        return None

    # ---- identity: boards, castle rights, en passant, side to move (core.rs:292-309)
    def _key(self):
        return (self.board, self.turn, self.castle, self.ep)

    def __eq__(self, other):
        return isinstance(other, ChessPosition) and self._key() == other._key()

    def __hash__(self):
        if self._hash is None:
            self._hash = hash(self._key())
        return self._hash

    # ---- attacks
    def _attacked(self, board, sq: int, by_white: bool) -> bool:
        r, f = divmod(sq, 8)
        pawn, knight, bishop, rook, queen, king = ("P", "N", "B", "R", "Q", "K") if by_white else ("p", "n", "b", "r", "q", "k")
        pr = r - 1 if by_white else r + 1  # a white pawn attacks upwards, so it stands one rank below
        for df in (-1, 1):
            if _on_board(pr, f + df) and board[pr * 8 + f + df] == pawn:
                return True
        for dr, df in _KNIGHT:
            if _on_board(r + dr, f + df) and board[(r + dr) * 8 + f + df] == knight:
                return True
        for dr, df in _KING:
            if _on_board(r + dr, f + df) and board[(r + dr) * 8 + f + df] == king:
                return True
        for dirs, slider in ((_ROOK, rook), (_BISHOP, bishop)):
            for dr, df in dirs:
                nr, nf = r + dr, f + df
                while _on_board(nr, nf):
                    pc = board[nr * 8 + nf]
                    if pc is not None:
                        if pc == slider or pc == queen:
                            return True
                        break
                    nr, nf = nr + dr, nf + df
        return False

    def _king_square(self, board, white: bool) -> int:
        return board.index("K" if white else "k")

    def in_check(self) -> bool:
        white = self.turn == P1
        return self._attacked(self.board, self._king_square(self.board, white), not white)

    def _pinned_squares(self) -> set:
        """Own pieces standing alone between the own king and an enemy slider that moves along that line."""
        white = self.turn == P1
        ksq = self._king_square(self.board, white)
        kr, kf = divmod(ksq, 8)
        out = set()
        for dirs, kinds in ((_ROOK, "rq"), (_BISHOP, "bq")):
            for dr, df in dirs:
                nr, nf = kr + dr, kf + df
                first = None
                while _on_board(nr, nf):
                    pc = self.board[nr * 8 + nf]
                    if pc is not None:
                        if first is None:
                            first = (nr * 8 + nf, pc)
                        else:
                            if _is_white(pc) != white and pc.lower() in kinds and _is_white(first[1]) == white:
                                out.add(first[0])
                            break
                    nr, nf = nr + dr, nf + df
        return out

    # ---- make move on a raw board; returns (board, captured_something)
    def _apply(self, m: Move):
        frm, to, promo = m
        board = list(self.board)
        piece = board[frm]
        white = _is_white(piece)
        captured = board[to] is not None
        board[frm] = None
        if piece.lower() == "p" and self.ep is not None and to == self.ep + (8 if white else -8) and frm % 8 != to % 8:
            board[self.ep] = None  # en passant: the captured pawn is not on the destination square
        if piece.lower() == "k" and abs(to - frm) == 2:
            if to > frm:
                board[to - 1], board[to + 1] = board[to + 1], None
            else:
                board[to + 1], board[to - 2] = board[to - 2], None
        board[to] = (promo.upper() if white else promo) if promo else piece
        return board, captured

    # ---- legal moves in the crate's order
    def legal_moves(self) -> List[Move]:
        if self._legal is None:
            self._legal = self._gen()
        return list(self._legal)

    def _gen(self) -> List[Move]:
        white = self.turn == P1
        board = self.board
        pseudo: List[Tuple[Move, bool]] = []  # (move, is_en_passant)
        for sq in range(64):
            pc = board[sq]
            if pc is None or _is_white(pc) != white:
                continue
            r, f = divmod(sq, 8)
            kind = pc.lower()
            if kind == "p":
                step = 1 if white else -1
                start, last = (1, 6) if white else (6, 1)
                promos = ("q", "n", "r", "b") if r == last else (None,)
                if board[sq + 8 * step] is None:
                    for p in promos:
                        pseudo.append(((sq, sq + 8 * step, p), False))
                    if r == start and board[sq + 16 * step] is None:
                        pseudo.append(((sq, sq + 16 * step, None), False))
                for df in (-1, 1):
                    if not 0 <= f + df < 8:
                        continue
                    to = sq + 8 * step + df
                    tp = board[to]
                    if tp is not None and _is_white(tp) != white:
                        for p in promos:
                            pseudo.append(((sq, to, p), False))
                    elif tp is None and self.ep is not None and self.ep == sq + df:
                        pseudo.append(((sq, to, None), True))
            elif kind in "nk":
                for dr, df in (_KNIGHT if kind == "n" else _KING):
                    if _on_board(r + dr, f + df):
                        to = (r + dr) * 8 + f + df
                        if board[to] is None or _is_white(board[to]) != white:
                            pseudo.append(((sq, to, None), False))
            else:
                dirs = {"b": _BISHOP, "r": _ROOK, "q": _ROOK + _BISHOP}[kind]
                for dr, df in dirs:
                    nr, nf = r + dr, f + df
                    while _on_board(nr, nf):
                        to = nr * 8 + nf
                        if board[to] is None:
                            pseudo.append(((sq, to, None), False))
                        else:
                            if _is_white(board[to]) != white:
                                pseudo.append(((sq, to, None), False))
                            break
                        nr, nf = nr + dr, nf + df
        legal: List[Tuple[Move, bool]] = []
        for m, is_ep in pseudo:
            nb, _ = self._apply(m)
            if not self._attacked(nb, self._king_square(nb, white), not white):
                legal.append((m, is_ep))
        # castling: rights, empty squares between, king not in check and not passing through or landing on an attacked square
        home = 0 if white else 56
        ksq = home + 4
        rights = self.castle
        if board[ksq] == ("K" if white else "k") and not self._attacked(board, ksq, not white):
            if ("K" if white else "k") in rights and board[ksq + 1] is None and board[ksq + 2] is None:
                if not self._attacked(board, ksq + 1, not white) and not self._attacked(board, ksq + 2, not white):
                    legal.append(((ksq, ksq + 2, None), False))
            if ("Q" if white else "q") in rights and board[ksq - 1] is None and board[ksq - 2] is None and board[ksq - 3] is None:
                if not self._attacked(board, ksq - 1, not white) and not self._attacked(board, ksq - 2, not white):
                    legal.append(((ksq, ksq - 2, None), False))
        pinned = self._pinned_squares()

        def order(item):
            (frm, to, promo), is_ep = item
            kind = board[frm].lower()
            group = 2 if is_ep else (1 if frm in pinned else 0)
            return (_TYPE_RANK[kind], group, frm, to, _PROMO_RANK[promo])

        legal.sort(key=order)
        return [m for m, _ in legal]

    # ---- Position trait
    def moved_position(self, m: Move) -> "ChessPosition":
        frm, to, _ = m
        piece = self.board[frm]
        white = _is_white(piece)
        board, captured = self._apply(m)
        castle = self.castle
        mine, theirs = ("KQ", "kq") if white else ("kq", "KQ")
        home, far = (0, 56) if white else (56, 0)
        if frm == home + 4:
            castle = castle.replace(mine[0], "").replace(mine[1], "")
        if frm == home + 7:
            castle = castle.replace(mine[0], "")
        if frm == home:
            castle = castle.replace(mine[1], "")
        if to == far + 7:
            castle = castle.replace(theirs[0], "")
        if to == far:
            castle = castle.replace(theirs[1], "")
        ep = None
        if piece.lower() == "p" and abs(to - frm) == 16:
            ep = ChessPosition._ep_if_capturable(board, to, mover_white=white)
        is_pawn = piece.lower() == "p"
        if is_pawn or captured:  # core.rs:334-343 (an en-passant capture is a pawn move anyway)
            fifty = 0
        elif self.turn == P1:
            fifty = self.fifty + 1
        else:
            fifty = self.fifty
        return ChessPosition(board, 3 - self.turn, castle, ep, fifty)

    def status(self):
        """('ongoing', None) | ('finished', winner or None) -- core.rs:348-364."""
        n = len(self._legal) if self._legal is not None else len(self.legal_moves())
        if n == 0:
            return ("finished", (3 - self.turn) if self.in_check() else None)
        if self.fifty >= 50:
            return ("finished", None)
        return ("ongoing", None)

    def is_finished(self) -> bool:
        return self.status()[0] == "finished"

    def flipped(self) -> "ChessPosition":
        board: List[Optional[str]] = [None] * 64
        for sq, pc in enumerate(self.board):
            if pc is not None:
                board[sq ^ 56] = pc.swapcase()
        castle = self.castle.swapcase()
        ep = None if self.ep is None else self.ep ^ 56
        return ChessPosition(board, 3 - self.turn, castle, ep, self.fifty)

    @staticmethod
    def flip_move(m: Move) -> Move:
        return (m[0] ^ 56, m[1] ^ 56, m[2])

    # ---- network view
    def planes(self) -> List[int]:
        planes = [0] * 18
        for sq, pc in enumerate(self.board):
            if pc is not None:
                planes[og._PIECE_PLANE[pc]] |= 1 << sq
        for k, c in enumerate("KQkq"):
            planes[12 + k] = og.U64_ALL if c in self.castle else 0
        planes[16] = 0 if self.ep is None else 1 << self.ep
        planes[17] = og.U64_ALL
        return planes

    @staticmethod
    def to_nn_idx(m: Move) -> int:
        idx = int(_nn_table()[og.chess_move_to_idx(move_to_lan(m))])
        assert idx != 0xFFFF, m
        return idx


def perft(pos: ChessPosition, depth: int) -> int:
    moves = pos.legal_moves()
    if depth <= 1:
        return len(moves) if depth == 1 else 1
    return sum(perft(pos.moved_position(m), depth - 1) for m in moves)


def serialize_entry(pos: ChessPosition, probs: Sequence[Tuple[Move, float]], winner: Optional[int]) -> bytes:
    """write_data_entry + ChessSerializer (self_play.rs:248-276, serialize/chess.rs:18-57): Player1's view, 18 planes,
    235-byte bitmap over the nn indices, 225 probabilities sorted by nn index (-1 padded), winner as i8."""
    w = 0 if winner is None else (1 if winner == P1 else -1)
    if pos.turn != P1:
        pos = pos.flipped()
        probs = [(ChessPosition.flip_move(m), p) for m, p in probs]
        w = -w
    legal = sorted(((ChessPosition.to_nn_idx(m), np.float32(p)) for m, p in probs), key=lambda t: t[0])
    return og.serialize_chess_entry(pos.planes(), legal, w)
