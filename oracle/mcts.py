"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the caller side of the evaluation path: MctsPlayer, the Hex / TicTacToe
rules it needs, ValueFuncCache and the self-play game loop.  Pure Python + numpy float32 scalars, small cases only.
Only `tests/` and `bench.py`'s cpu_baseline leg may import it; the product (cattus_b200/csrc/selfplay.cpp) never does.

Restated from (paths relative to /root/reference):

* MctsPlayer                    engine/src/mcts/mod.rs:105-454  (select :199-231, heuristic :233-244, create_children
                                :246-262, backpropagate :270-281, tree reuse :283-352, move probabilities :335-385,
                                move choice :387-417, Dirichlet noise :419-446), TemperaturePolicy :456-489
* HexPosition                   engine/src/hex/core.rs:112-335  (update_reach :215-264, make_move :272-285,
                                legal_moves :297-305, status :314-322, flipped :324-334)
* TttPosition                   engine/src/ttt/core.rs:101-246
* NNetwork::evaluate            engine/src/net/mod.rs:74-103, flip helpers :158-182
* ValueFuncCache                engine/src/mcts/cache.rs:31-75
* self-play game loop           training/self-play/src/self_play.rs:179-276, serializers serialize/{hex,ttt}.rs

Third-party behaviour this restatement models explicitly (parity UNPINNED -- the crates are not in the tree):

* petgraph 0.8 `DiGraph::edges(n)` walks the per-node linked list newest-edge-first, so iteration order is the
  REVERSE of insertion order; `remove_all_but_subtree` (mod.rs:303-333) re-inserts edges in iteration order, which
  reverses every kept node's child order at each tree reuse.  Children are kept here in insertion order and
  `_edges()` yields them reversed.
* `Iterator::max_by` returns the LAST maximal element; with the reversed iteration that is the maximal child that
  was inserted FIRST.
* `rand::rng()` (noise, temperature sampling) is unseeded in the reference; here and in the product both draw from
  the SplitMix64 stream defined below so that whole games are reproducible and comparable between the two.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
P1, P2 = 1, 2


def opposite(c: int) -> int:
    return 3 - c


def to_signed_one(winner: Optional[int]) -> int:
    """GameColor::to_signed_one (game/mod.rs)."""
    return 0 if winner is None else (1 if winner == P1 else -1)


# --------------------------------------------------------------------------------------------------------------
# Hex rules (engine/src/hex/core.rs)
# --------------------------------------------------------------------------------------------------------------
_HEX_DIRS = ((0, 1), (-1, 0), (-1, -1), (0, -1), (1, 0), (1, 1))  # hex/core.rs:204


def _transpose(bb: int, s: int) -> int:
    f = 0
    for r in range(s):
        for c in range(s):
            if (bb >> (r * s + c)) & 1:
                f |= 1 << (c * s + r)
    return f


@dataclass(frozen=True)
class HexPosition:
    s: int
    red: int = 0
    blue: int = 0
    turn: int = P1
    left_red_reach: int = 0
    top_blue_reach: int = 0
    empty: int = -1
    winner: Optional[int] = None

    @staticmethod
    def new(s: int, starting: int = P1) -> "HexPosition":
        return HexPosition(s=s, turn=starting, empty=s * s)

    @staticmethod
    def from_board(s: int, red: int, blue: int, turn: int) -> "HexPosition":
        """new_from_board (hex/core.rs:143-173)."""
        st = _HexMutable(s, red, blue, turn, 0, 0, s * s, None)
        for r in range(s):
            for c in range(s):
                idx = r * s + c
                color = P1 if (red >> idx) & 1 else (P2 if (blue >> idx) & 1 else None)
                if color is not None:
                    st.empty -= 1
                    if (c == 0) if color == P1 else (r == 0):
                        st.update_reach(r, c, color)
        return st.freeze()

    def status(self):
        """('finished', winner) | ('ongoing', None)  (hex/core.rs:314-322)."""
        if self.winner is not None:
            return ("finished", self.winner)
        if self.empty == 0:
            return ("finished", None)
        return ("ongoing", None)

    def is_finished(self) -> bool:
        return self.status()[0] == "finished"

    def legal_moves(self) -> List[int]:
        occ = self.red | self.blue
        return [i for i in range(self.s * self.s) if not (occ >> i) & 1]

    def moved_position(self, m: int) -> "HexPosition":
        assert not self.is_finished() and not ((self.red | self.blue) >> m) & 1
        st = _HexMutable(self.s, self.red, self.blue, self.turn, self.left_red_reach, self.top_blue_reach, self.empty, self.winner)
        if st.turn == P1:
            st.red |= 1 << m
        else:
            st.blue |= 1 << m
        st.update_reach(m // self.s, m % self.s, st.turn)
        st.empty -= 1
        st.turn = opposite(st.turn)
        return st.freeze()

    def flipped(self) -> "HexPosition":
        s = self.s
        return HexPosition(s, _transpose(self.blue, s), _transpose(self.red, s), opposite(self.turn), _transpose(self.top_blue_reach, s),
                           _transpose(self.left_red_reach, s), self.empty, None if self.winner is None else opposite(self.winner))

    def flip_move(self, m: int) -> int:
        r, c = divmod(m, self.s)
        return c * self.s + r

    def planes(self) -> List[int]:
        return [self.red, self.blue, (1 << (self.s * self.s)) - 1]

    @property
    def moves_num(self) -> int:
        return self.s * self.s


class _HexMutable:
    def __init__(self, s, red, blue, turn, lrr, tbr, empty, winner):
        self.s, self.red, self.blue, self.turn, self.lrr, self.tbr, self.empty, self.winner = s, red, blue, turn, lrr, tbr, empty, winner

    def freeze(self) -> HexPosition:
        return HexPosition(self.s, self.red, self.blue, self.turn, self.lrr, self.tbr, self.empty, self.winner)

    def _neighbors(self, r, c):
        for dr, dc in _HEX_DIRS:
            nr, nc = r + dr, c + dc
            if 0 <= nr < self.s and 0 <= nc < self.s:
                yield nr, nc

    def update_reach(self, r: int, c: int, player: int) -> None:
        """hex/core.rs:215-264."""
        s = self.s
        board = self.red if player == P1 else self.blue
        reach = self.lrr if player == P1 else self.tbr
        begin = (lambda rr, cc: cc == 0) if player == P1 else (lambda rr, cc: rr == 0)
        end = (lambda rr, cc: cc == s - 1) if player == P1 else (lambda rr, cc: rr == s - 1)
        layer = 0
        upd = begin(r, c)
        for nr, nc in self._neighbors(r, c):
            upd = upd or bool((reach >> (nr * s + nc)) & 1)
        if upd:
            reach |= 1 << (r * s + c)
            layer |= 1 << (r * s + c)
        while layer:
            idx = (layer & -layer).bit_length() - 1
            layer &= ~(1 << idx)
            rr, cc = divmod(idx, s)
            if end(rr, cc):
                self.winner = player
            else:
                for nr, nc in self._neighbors(rr, cc):
                    n = nr * s + nc
                    if not (reach >> n) & 1 and (board >> n) & 1:
                        reach |= 1 << n
                        layer |= 1 << n
        if player == P1:
            self.lrr = reach
        else:
            self.tbr = reach


# --------------------------------------------------------------------------------------------------------------
# TicTacToe rules (engine/src/ttt/core.rs)
# --------------------------------------------------------------------------------------------------------------
_TTT_WINS = (0b111000000, 0b000111000, 0b000000111, 0b100100100, 0b010010010, 0b001001001, 0b100010001, 0b001010100)


@dataclass(frozen=True)
class TttPosition:
    x: int = 0
    o: int = 0
    turn: int = P1
    winner: Optional[int] = None
    s: int = 3

    @staticmethod
    def new() -> "TttPosition":
        return TttPosition()

    @staticmethod
    def _winner(x: int, o: int) -> Optional[int]:
        for w in _TTT_WINS:  # ttt/core.rs:170-193: x is checked before o for each line
            if x & w == w:
                return P1
            if o & w == w:
                return P2
        return None

    def status(self):
        if self.winner is not None:
            return ("finished", self.winner)
        if (self.x | self.o) == 0x1FF:
            return ("finished", None)
        return ("ongoing", None)

    def is_finished(self) -> bool:
        return self.status()[0] == "finished"

    def legal_moves(self) -> List[int]:
        occ = self.x | self.o
        return [i for i in range(9) if not (occ >> i) & 1]

    def moved_position(self, m: int) -> "TttPosition":
        assert not self.is_finished() and not ((self.x | self.o) >> m) & 1
        x, o = self.x, self.o
        if self.turn == P1:
            x |= 1 << m
        else:
            o |= 1 << m
        return TttPosition(x, o, opposite(self.turn), self._winner(x, o))

    def flipped(self) -> "TttPosition":
        return TttPosition(self.o, self.x, opposite(self.turn), None if self.winner is None else opposite(self.winner))

    def flip_move(self, m: int) -> int:
        return m

    def planes(self) -> List[int]:
        return [self.x, self.o, 0x1FF]

    @property
    def moves_num(self) -> int:
        return 9


# --------------------------------------------------------------------------------------------------------------
# Shared random stream (the reference uses the unseeded thread-local rand::rng(); see the module docstring)
# --------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.state = seed & _M64

    def next_u64(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & _M64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def next_f64(self) -> float:
        """uniform in [0, 1) with 53 bits."""
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)

    def next_open_f64(self) -> float:
        """uniform in (0, 1)."""
        return ((self.next_u64() >> 12) + 0.5) * (1.0 / 4503599627370496.0)

    def normal(self) -> float:
        """Box-Muller, one variate per two uniforms (libm log / sqrt / cos)."""
        u1 = self.next_open_f64()
        u2 = self.next_f64()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)

    def gamma(self, alpha: float) -> float:
        """Marsaglia-Tsang; alpha < 1 boosted with U^(1/alpha)."""
        if alpha < 1.0:
            u = self.next_open_f64()
            return self.gamma(alpha + 1.0) * math.pow(u, 1.0 / alpha)
        d = alpha - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        while True:
            x = self.normal()
            v = 1.0 + c * x
            if v <= 0.0:
                continue
            v = v * v * v
            u = self.next_open_f64()
            if math.log(u) < 0.5 * x * x + d - d * v + d * math.log(v):
                return d * v


def game_seed(base_seed: int, game_idx: int) -> int:
    return (base_seed ^ (0xD1B54A32D192ED03 * (game_idx + 1))) & _M64


def dirichlet(rng: SplitMix64, alpha: float, n: int) -> List[np.float32]:
    """n iid Gamma(alpha) normalised to sum 1, returned as f32 (the reference's util/dirichlet.rs samples f32)."""
    g = [rng.gamma(alpha) for _ in range(n)]
    tot = 0.0
    for x in g:
        tot += x
    return [f32(x / tot) for x in g]


# --------------------------------------------------------------------------------------------------------------
# TemperaturePolicy (mcts/mod.rs:456-489)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class TemperaturePolicy:
    temperatures: List[Tuple[int, float]] = field(default_factory=list)
    last_temperature: float = 1.0

    @staticmethod
    def from_config(policy: Sequence[Tuple[int, float]]) -> "TemperaturePolicy":
        """self_play_cmd.rs:69-72: all but the last entry are scheduled, the last entry's temperature is the tail."""
        assert policy
        return TemperaturePolicy([(int(n), float(t)) for n, t in policy[:-1]], float(policy[-1][1]))

    def get(self, move_num: int) -> float:
        for threshold, t in self.temperatures:
            if move_num < threshold:
                return t
        return self.last_temperature


# --------------------------------------------------------------------------------------------------------------
# ValueFuncCache + NNetwork::evaluate (mcts/cache.rs:31-75, net/mod.rs:74-103)
# --------------------------------------------------------------------------------------------------------------
class ValueFuncCache:
    def __init__(self, max_size: int):
        assert max_size > 0
        self.max_size = max_size
        self.map: "OrderedDict" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get_or_compute(self, pos, compute):
        if pos in self.map:
            self.hits += 1
            return self.map[pos]
        val = compute(pos)
        while len(self.map) >= self.max_size:
            self.map.popitem(last=False)
        self.map[pos] = val
        self.misses += 1
        return val


NetFn = Callable[[object], Tuple[List[float], float]]
"""net_fn(position with Player1 to move) -> (probabilities over position.legal_moves() in that order, value)."""


class Evaluator:
    """NNetwork::evaluate: flip -> cache -> net -> un-flip (net/mod.rs:74-87, :158-182)."""

    def __init__(self, net_fn: NetFn, cache: Optional[ValueFuncCache] = None):
        self.net_fn = net_fn
        self.cache = cache
        self.calls = 0

    def _impl(self, pos):
        self.calls += 1
        probs, val = self.net_fn(pos)
        moves = pos.legal_moves()
        assert len(probs) == len(moves)
        return [(m, f32(p)) for m, p in zip(moves, probs)], f32(val)

    def evaluate(self, position):
        flipped = position.turn != P1
        pos = position.flipped() if flipped else position
        res = self.cache.get_or_compute(pos, self._impl) if self.cache is not None else self._impl(pos)
        if not flipped:
            return res
        moves_probs, val = res
        return [(pos.flip_move(m), p) for m, p in moves_probs], f32(-val)


# --------------------------------------------------------------------------------------------------------------
# MctsPlayer (mcts/mod.rs)
# --------------------------------------------------------------------------------------------------------------
class _Edge:
    __slots__ = ("m", "init_score", "n", "w", "target")

    def __init__(self, m, init_score, target):
        self.m = m
        self.init_score = f32(init_score)
        self.n = 0
        self.w = f32(0.0)
        self.target = target


class _Node:
    __slots__ = ("position", "children")

    def __init__(self, position):
        self.position = position
        self.children: List[_Edge] = []  # insertion order


def _edges(node: _Node):
    """petgraph `edges()`: newest first."""
    return reversed(node.children)


@dataclass
class MctsParams:
    sim_num: int
    explore_factor: float = math.sqrt(2.0)
    temperature: TemperaturePolicy = field(default_factory=lambda: TemperaturePolicy([], 1.0))
    prior_noise_alpha: float = 0.0
    prior_noise_epsilon: float = 0.0


class MctsPlayer:
    def __init__(self, params: MctsParams, evaluator: Evaluator, rng: SplitMix64):
        assert params.sim_num > 0 and params.explore_factor >= 0 and params.prior_noise_alpha >= 0
        assert 0.0 <= params.prior_noise_epsilon <= 1.0
        self.p = params
        self.explore_factor = f32(params.explore_factor)
        self.evaluator = evaluator
        self.rng = rng
        self.root: Optional[_Node] = None
        self.sims_done = 0
        self.terminal_leaves = 0   # simulations that ended on a finished position or a repetition (no evaluation)
        self.repetition_hits = 0

    # -- mod.rs:233-244
    def _heuristic(self, e: _Edge, parent_simcount: int) -> np.float32:
        exploit = f32(0.0) if e.n == 0 else f32(e.w / f32(e.n))
        explore = f32(f32(self.explore_factor * e.init_score) * f32(np.sqrt(f32(parent_simcount)) / f32(1 + e.n)))
        return f32(exploit + explore)

    # -- mod.rs:199-231
    def _select(self) -> List[Tuple[_Node, _Edge]]:
        path = []
        node = self.root
        while True:
            if node.position.is_finished() or not node.children:
                return path
            simcount = 1 + sum(e.n for e in node.children)
            best, best_val = None, None
            for e in _edges(node):
                v = self._heuristic(e, simcount)
                if best is None or not (v < best_val):  # max_by keeps the LAST maximum; NaN compares Equal
                    best, best_val = e, v
            path.append((node, best))
            node = best.target

    # -- mod.rs:133-154
    @staticmethod
    def _detect_repetition(pos_history: Sequence, path) -> bool:
        limit = getattr(type(pos_history[-1]), "REPETITION_LIMIT", None)
        if limit is None or limit <= 1:
            return False
        counts: dict = {}
        for pos in list(pos_history) + [e.target.position for _, e in path]:
            counts[pos] = counts.get(pos, 0) + 1
            if counts[pos] >= limit:
                return True
        return False

    # -- mod.rs:156-196
    def _develop_tree(self, pos_history: Sequence) -> None:
        assert self.p.sim_num > 1
        for _ in range(self.p.sim_num):
            path = self._select()
            repetition_reached = self._detect_repetition(pos_history, path)
            leaf = path[-1][1].target if path else self.root
            st, winner = (None, None) if repetition_reached else leaf.position.status()
            if repetition_reached:
                ev = f32(0.0)
                self.terminal_leaves += 1
                self.repetition_hits += 1
            elif st == "finished":
                ev = f32(to_signed_one(winner))
                self.terminal_leaves += 1
            else:
                per_move, ev = self.evaluator.evaluate(leaf.position)
                for m, p in per_move:  # create_children, mod.rs:246-262
                    leaf.children.append(_Edge(m, p, _Node(leaf.position.moved_position(m))))
                if leaf is self.root:
                    self._add_dirichlet_noise(leaf)
            for src, e in path:  # backpropagate, mod.rs:270-281
                e.n += 1
                e.w = f32(e.w + (ev if src.position.turn == P1 else f32(-ev)))
            self.sims_done += 1

    # -- mod.rs:283-301
    def _find_node_with_position(self, position, depth_limit: int) -> Optional[_Node]:
        layer = [self.root]
        for _ in range(depth_limit):
            nxt = []
            for node in layer:
                if node.position == position:
                    return node
                for e in _edges(node):
                    nxt.append(e.target)
            layer = nxt
        return None

    # -- mod.rs:303-333
    def _remove_all_but_subtree(self, sub_root: _Node) -> None:
        if self.root is sub_root:
            return
        new_root = _Node(sub_root.position)
        stack = [(sub_root, new_root)]
        while stack:
            old, new = stack.pop()
            for e in _edges(old):  # re-inserted in iteration order => child order reversed
                child_new = _Node(e.target.position)
                ne = _Edge(e.m, e.init_score, child_new)
                ne.n, ne.w = e.n, e.w
                new.children.append(ne)
                stack.append((e.target, child_new))
        self.root = new_root
        if new_root.children:
            self._add_dirichlet_noise(new_root)

    # -- mod.rs:419-446
    def _add_dirichlet_noise(self, node: _Node) -> None:
        if self.p.prior_noise_alpha == 0.0 or self.p.prior_noise_epsilon == 0.0:
            return
        moves = list(_edges(node))
        if len(moves) < 2:
            return
        noise = dirichlet(self.rng, self.p.prior_noise_alpha, len(moves))
        eps = f32(self.p.prior_noise_epsilon)
        for e, nz in zip(moves, noise):
            e.init_score = f32(f32(f32(1.0) - eps) * e.init_score + f32(eps * nz))

    # -- mod.rs:335-385
    def calc_moves_probabilities(self, pos_history: Sequence) -> List[Tuple[int, np.float32]]:
        position = pos_history[-1]
        if self.root is not None:
            node = self._find_node_with_position(position, 3)
            if node is not None:
                self._remove_all_but_subtree(node)
            else:
                self.root = None
        if self.root is None:
            self.root = _Node(position)
        assert self.root.position == position
        self._develop_tree(pos_history)
        ms = [(e.m, e.n) for e in _edges(self.root)]
        total = sum(n for _, n in ms)
        return [(m, f32(f32(n) / f32(total))) for m, n in ms]

    # -- mod.rs:387-417
    def choose_move_from_probabilities(self, pos_history: Sequence, moves_probs) -> Optional[int]:
        if not moves_probs:
            return None
        temperature = f32(self.p.temperature.get(len(pos_history) // 2))
        if temperature == 0.0:
            best = None
            for m, p in moves_probs:
                if best is None or not (p < best[1]):  # max_by(total_cmp): last maximum
                    best = (m, p)
            return best[0]
        inv = f32(f32(1.0) / temperature)
        pw = [powf32(p, inv) for _, p in moves_probs]
        tot = f32(0.0)
        for x in pw:
            tot = f32(tot + x)
        pw = [f32(x / tot) for x in pw]
        return moves_probs[weighted_index(self.rng, pw)][0]


def powf32(x: np.float32, y: np.float32) -> np.float32:
    """f32::powf: evaluated in double by libm's pow and rounded once (glibc powf is correctly rounded in practice)."""
    return f32(math.pow(float(x), float(y)))


def weighted_index(rng: SplitMix64, weights: Sequence[np.float32]) -> int:
    """First i with cumsum_f32(weights)[i] > u * total, u uniform in [0, 1) (stand-in for rand's WeightedIndex)."""
    cum = []
    tot = f32(0.0)
    for w in weights:
        tot = f32(tot + w)
        cum.append(tot)
    x = rng.next_f64() * float(tot)
    for i, c in enumerate(cum):
        if float(c) > x:
            return i
    return len(cum) - 1


# --------------------------------------------------------------------------------------------------------------
# Self-play game loop (training/self-play/src/self_play.rs:179-276)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class GameRecord:
    game_idx: int
    winner: Optional[int]
    moves: List[int]
    entries: List[Tuple[object, List[Tuple[int, np.float32]]]]  # (position before the move, MCTS probabilities)
    sims: int
    terminal_leaves: int = 0
    repetition_hits: int = 0


def play_game(game_idx: int, new_position: Callable[[], object], params1: MctsParams, params2: MctsParams, eval1: Evaluator,
              eval2: Evaluator, base_seed: int, max_moves: int = 0) -> GameRecord:
    rng = SplitMix64(game_seed(base_seed, game_idx))
    player1 = MctsPlayer(params1, eval1, rng)
    player2 = MctsPlayer(params2, eval2, rng)
    history = [new_position()]
    entries, moves = [], []
    switch = game_idx % 2 == 1
    # ChessGame (chess/core.rs:402-451): a position seen REPETITION_LIMIT times ends the game as a draw
    limit = getattr(type(history[0]), "REPETITION_LIMIT", None)
    seen = {history[0]: 1} if limit else None
    while True:
        st, winner = history[-1].status()
        if seen is not None and seen[history[-1]] >= limit:
            st, winner = "finished", None
        if st != "finished" and max_moves and len(moves) >= max_moves:  # the product's bounded-run option (not in the reference)
            st, winner = "finished", None
        if st == "finished":
            break
        who = history[-1].turn
        if switch:
            who = opposite(who)
        player = player1 if who == P1 else player2
        probs = player.calc_moves_probabilities(history)
        mv = player.choose_move_from_probabilities(history, probs)
        entries.append((history[-1], probs))
        moves.append(mv)
        history.append(history[-1].moved_position(mv))
        if seen is not None:
            seen[history[-1]] = seen.get(history[-1], 0) + 1
    return GameRecord(game_idx, winner, moves, entries, player1.sims_done + player2.sims_done,
                      player1.terminal_leaves + player2.terminal_leaves, player1.repetition_hits + player2.repetition_hits)


def data_entry_bytes(pos, probs, winner: Optional[int]) -> bytes:
    """write_data_entry + the hex/ttt serializers + SerializerBase::write_entry (self_play.rs:33-61,248-276;
    serialize/hex.rs:16-28; serialize/ttt.rs:17-22): flip to Player1's view, planes as u64 LE (hex: lo, hi per plane),
    dense f32 probabilities with -1 for illegal moves, winner as i8."""
    import struct

    from . import chess as _chess

    if isinstance(pos, _chess.ChessPosition):
        return _chess.serialize_entry(pos, probs, winner)
    w = f32(to_signed_one(winner))
    flipped = pos.turn != P1
    if flipped:
        pos2 = pos.flipped()
        probs = [(pos2.flip_move(m), p) for m, p in probs]
        w = f32(-w)
        pos = pos2
    dense = np.full((pos.moves_num,), -1.0, dtype="<f4")
    for m, p in probs:
        dense[m] = p
    words = []
    for pl in pos.planes():
        if isinstance(pos, HexPosition):
            words += [pl & _M64, (pl >> 64) & _M64]
        else:
            words.append(pl & _M64)
    return struct.pack(f"<{len(words)}Q", *words) + dense.tobytes() + struct.pack("<b", int(w))


def data_entry_dir(pos_turn: int, game_idx: int) -> int:
    """self_play.rs:256-259: which of (out_dir1, out_dir2) an entry goes to; returns 1 or 2."""
    pair = (1, 2) if pos_turn == P1 else (2, 1)
    return pair[game_idx % 2]
