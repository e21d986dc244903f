"""CPU tests of the chess side of the self-play driver: the rules (csrc/chess_rules.hpp through
include/cattus_b200_chess.h) against the published perft counts and against the independent mailbox restatement in
oracle/chess.py, and whole chess self-play games (csrc/selfplay.cpp) against oracle/mcts.py.

The reference's move generator is the crate `chess` 3.2.0, which is not in the tree: the legal move SET is pinned by
perft; the ORDER of `MoveGen::new_legal` is restated twice (bitboards in C++, sort key over a mailbox board in Python)
and the two must agree on every position visited below.
"""
from __future__ import annotations

import ctypes as C
import math
import random

import numpy as np
import pytest

from cattus_b200 import _lib
from cattus_b200.selfplay import SelfPlayRunner
from oracle import chess as oc
from oracle import games as og
from oracle import mcts as om

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -"
KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
# the standard move-generator test positions with their published node counts
PERFT = [
    (START, [20, 400, 8902, 197281, 4865609]),
    (KIWIPETE, [48, 2039, 97862, 4085603]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -", [14, 191, 2812, 43238, 674624]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467, 422333]),
    ("r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1", [6, 264, 9467, 422333]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379, 2103487]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890, 3894594]),
]


def chess_info(fen: str, moves=()):
    lib = _lib.load()
    info = _lib.ChessInfo()
    info.struct_size = C.sizeof(_lib.ChessInfo)
    arr = (C.c_uint16 * max(1, len(moves)))(*moves)
    rc = lib.cattus_b200_chess_position(fen.encode(), arr, len(moves), C.byref(info))
    assert rc == 0, lib.cattus_b200_chess_last_error()
    return info


@pytest.mark.parametrize("fen,counts", PERFT)
def test_perft_matches_published_counts(fen, counts):
    lib = _lib.load()
    for depth, want in enumerate(counts, 1):
        n = C.c_uint64()
        assert lib.cattus_b200_chess_perft(fen.encode(), depth, C.byref(n)) == 0
        assert n.value == want, (fen, depth)
    pos = oc.ChessPosition.from_fen(fen)  # the oracle, as deep as pure Python goes in a moment
    for depth, want in enumerate(counts[:2], 1):
        assert oc.perft(pos, depth) == want


def test_nn_index_table_matches_oracle_table():
    lib = _lib.load()
    table = (C.c_uint16 * (64 * 64 + 88))()
    assert lib.cattus_b200_chess_nn_table(table, len(table)) == 0
    assert np.array_equal(np.frombuffer(table, dtype=np.uint16), og.chess_move_to_nn_index_table())
    assert lib.cattus_b200_chess_nn_table(table, 10) != 0


def test_fixture_planes():
    """planes of the reference's chess fixtures (training/tests/test_net_output.py:199-203)."""
    info = chess_info(START)
    assert list(info.planes)[:12] == [0xFF00, 0x42, 0x24, 0x81, 0x08, 0x10, 0x00FF000000000000, 0x4200000000000000, 0x2400000000000000,
                                      0x8100000000000000, 0x0800000000000000, 0x1000000000000000]
    assert list(info.planes)[12:] == [og.U64_ALL] * 4 + [0, og.U64_ALL]
    assert info.n_legal == 20 and info.turn == 1 and info.status == 0
    info = chess_info("4k2r/6r1/8/8/8/8/3R4/R3K3 w Qk -")
    pl = list(info.planes)
    assert pl[3] == 0x801 and pl[5] == 0x10 and pl[9] == 0x8040000000000000 and pl[11] == 0x1000000000000000
    assert pl[12:16] == [0, og.U64_ALL, og.U64_ALL, 0] and pl[16] == 0
    # en passant: recorded only when a pawn of the side to move stands beside the pawn that just advanced
    assert chess_info("rnbqkbnr/pppp1ppp/8/8/4pP2/8/PPPPP1PP/RNBQKBNR b KQkq f3").planes[16] == 1 << (32 + 5)  # black's view: rank mirrored
    assert chess_info("rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3").planes[16] == 0
    for bad in ("", "8/8/8/8/8/8/8/8 w - -", "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBN w KQkq -", "4k3/8/8/8/8/8/8/4K3 w K -"):
        info = _lib.ChessInfo()
        info.struct_size = C.sizeof(_lib.ChessInfo)
        assert _lib.load().cattus_b200_chess_position(bad.encode(), None, 0, C.byref(info)) != 0


def _eval_order(p: oc.ChessPosition):
    """legal_moves() as NNetwork::evaluate hands them to the tree (net/mod.rs:74-87, :166-182) and the evaluated view."""
    if p.turn == oc.P1:
        return p.legal_moves(), p
    f = p.flipped()
    return [oc.ChessPosition.flip_move(m) for m in f.legal_moves()], f


def test_rules_match_oracle_on_random_playouts():
    rng = random.Random(11)
    fens = [f for f, _ in PERFT]
    ends = {}
    for g in range(24):
        fen = fens[g % len(fens)]
        p = oc.ChessPosition.from_fen(fen)
        moves = []
        for _ply in range(220):
            info = chess_info(fen, moves)
            mv, view = _eval_order(p)
            st, w = p.status()
            want = 0 if st == "ongoing" else (3 if w is None else w)
            assert (info.turn, info.status, info.fifty_rule_count, info.in_check) == (p.turn, want, p.fifty, int(p.in_check()))
            assert list(info.planes) == view.planes()
            assert view.flipped() == (p if p.turn != oc.P1 else p.flipped())
            if want:
                assert info.n_legal == 0
                ends[want] = ends.get(want, 0) + 1
                break
            assert [oc.move_from_u16(info.moves[i]) for i in range(info.n_legal)] == mv
            nn = [oc.ChessPosition.to_nn_idx(m if p.turn == oc.P1 else oc.ChessPosition.flip_move(m)) for m in mv]
            assert [info.nn_index[i] for i in range(info.n_legal)] == nn
            assert bytes(info.legal_bitmap) == og.bitmap_from_legal(nn, og.CHESS_MOVES_NUM).tobytes()
            m = rng.choice(mv)
            moves.append(oc.move_to_u16(m))
            p = p.moved_position(m)
    assert ends  # some playouts reach mate, stalemate or the fifty-move draw
    bad = _lib.ChessInfo()
    bad.struct_size = C.sizeof(_lib.ChessInfo)
    illegal = (C.c_uint16 * 1)(oc.move_to_u16((12, 36, None)))  # e2e5
    assert _lib.load().cattus_b200_chess_position(START.encode(), illegal, 1, C.byref(bad)) != 0


def test_oracle_flip_and_fifty_rule():
    """flip_rand / fifty-move bookkeeping (chess/core.rs:607-730 tests the same properties with a random player)."""
    rng = random.Random(3)
    p = oc.ChessPosition.new()
    for _ in range(120):
        if p.is_finished():
            break
        f = p.flipped()
        assert f.flipped() == p and f.turn == 3 - p.turn and f.fifty == p.fifty
        assert {oc.ChessPosition.flip_move(m) for m in f.legal_moves()} == set(p.legal_moves())
        p = p.moved_position(rng.choice(p.legal_moves()))
    # knights out and back: no pawn move or capture, so only white's moves count (core.rs:334-343)
    p = oc.ChessPosition.new()
    for lan in ["g1f3", "g8f6", "f3g1", "f6g8"] * 3:
        m = (og.chess_move_to_idx(lan) // 64, og.chess_move_to_idx(lan) % 64, None)
        p = p.moved_position(m)
    assert p.fifty == 6 and p == oc.ChessPosition.new()  # equality ignores the counter (core.rs:292-309)


# --------------------------------------------------------------------------------------------------------------
# whole games: the C++ driver against oracle/mcts.py, both driven by the same deterministic "network"
# --------------------------------------------------------------------------------------------------------------
def _mix(x: int) -> int:
    x &= (1 << 64) - 1
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & ((1 << 64) - 1)
    x ^= x >> 33
    return x


def chess_fake_net(kind: str, salt: int = 0):
    """net(planes[18], legal nn indices ascending) -> (probabilities in that order, value)."""

    def net(planes, legal):
        if kind == "uniform":
            return np.full(len(legal), np.float32(1.0) / np.float32(len(legal)), dtype=np.float32), np.float32(0.0)
        h = salt
        for w in planes[:17]:
            h = _mix(h * 0x9E3779B97F4A7C15 + int(w) + 1)
        logits = np.array([((_mix(h + 977 * i) % 1000) / 250.0) - 2.0 for i in legal], dtype=np.float32)
        e = np.exp(logits - logits.max()).astype(np.float32)
        p = (e / e.sum(dtype=np.float32)).astype(np.float32)
        v = np.float32(((h >> 7) % 2001) / 1000.0 - 1.0)
        if kind == "coarse":
            v = np.float32(round(float(v) * 2) / 2)
        return p, v

    return net


def chess_cb(net):
    def cb(words: np.ndarray, n: int, legal: np.ndarray):
        probs, values = [], []
        for i in range(n):
            idx = og.legal_from_bitmap(legal[i], og.CHESS_MOVES_NUM)
            p, v = net([int(w) for w in words[i]], idx)
            probs.append(p)
            values.append(v)
        return probs, values

    return cb


def chess_oracle_fn(net):
    def fn(pos: oc.ChessPosition):
        moves = pos.legal_moves()
        nn = [oc.ChessPosition.to_nn_idx(m) for m in moves]
        order = sorted(nn)
        p, v = net(pos.planes(), order)
        rank = {idx: k for k, idx in enumerate(order)}
        return [p[rank[i]] for i in nn], v  # calc_moves_probs gathers per legal move (net/mod.rs:106-119)

    return fn


def _params(cfg):
    mc = cfg["mcts"]
    return om.MctsParams(sim_num=mc["sim_num"], explore_factor=mc.get("explore_factor", math.sqrt(2.0)),
                         temperature=om.TemperaturePolicy.from_config(mc.get("temperature_policy", [[0, 1.0]])),
                         prior_noise_alpha=mc.get("prior_noise_alpha", 0.0), prior_noise_epsilon=mc.get("prior_noise_epsilon", 0.0))


def check_chess_against_oracle(cfg: dict, kind: str, games_num: int, kind2=None):
    net1 = chess_fake_net(kind)
    net2 = chess_fake_net(kind2, salt=99) if kind2 else None
    summary, records = SelfPlayRunner("chess", cfg).run_with(chess_cb(net1), chess_cb(net2) if net2 else None, games_num, keep_records=True)
    mc = cfg["mcts"]
    e1 = om.Evaluator(chess_oracle_fn(net1), om.ValueFuncCache(mc["cache_size"]) if mc.get("cache_size") else None)
    e2 = e1 if net2 is None else om.Evaluator(chess_oracle_fn(net2), om.ValueFuncCache(mc["cache_size"]) if mc.get("cache_size") else None)
    params = _params(cfg)
    ref = [om.play_game(g, oc.ChessPosition.new, params, params, e1, e2, cfg.get("seed", 0)) for g in range(games_num)]
    how = []
    for rec, o in zip(records, ref):
        assert rec.game_idx == o.game_idx
        assert rec.moves == [oc.move_to_u16(m) for m in o.moves], (rec.game_idx, len(rec.moves), len(o.moves))
        assert rec.winner == o.winner
        assert len(rec.entries) == len(o.entries)
        for k, (pos, probs) in enumerate(o.entries):
            assert rec.entries[k] == om.data_entry_bytes(pos, probs, o.winner), (rec.game_idx, k)
            assert len(rec.entries[k]) == 18 * 8 + 235 + 225 * 4 + 1
            assert rec.entry_dirs[k] == om.data_entry_dir(pos.turn, o.game_idx)
        last = oc.ChessPosition.new()
        for m in o.moves:
            last = last.moved_position(m)
        st, _ = last.status()
        how.append("repetition" if st == "ongoing" else ("fifty" if last.legal_moves() else ("mate" if last.in_check() else "stalemate")))
    m = summary["metrics"]
    assert m["selfplay.simulations"] == sum(o.sims for o in ref) == m["selfplay.searches"] * cfg["mcts"]["sim_num"]
    assert summary["player1_wins"] + summary["player2_wins"] + summary["draws"] == games_num
    assert m["selfplay.terminal_leaves"] == sum(o.terminal_leaves for o in ref)
    summary["oracle.repetition_hits"] = sum(o.repetition_hits for o in ref)
    return summary, records, how


CHESS_BASE = {"mcts": {"sim_num": 12, "explore_factor": 1.41421, "temperature_policy": [[9999, 0.0]], "prior_noise_alpha": 0.0,
                       "prior_noise_epsilon": 0.0, "cache_size": 0}, "threads": 1, "games_per_thread": 1, "seed": 5}


def chess_cfg(**kw):
    c = {"mcts": dict(CHESS_BASE["mcts"]), "threads": 1, "games_per_thread": 1, "seed": 5}
    for k, v in kw.items():
        if k in c["mcts"]:
            c["mcts"][k] = v
        else:
            c[k] = v
    return c


def test_chess_games_match_oracle_and_end_by_repetition():
    # a deterministic net at temperature 0 shuffles pieces: these games end by threefold repetition, which exercises
    # MctsPlayer::detect_repetition inside the searches on the way
    s, _, how = check_chess_against_oracle(chess_cfg(), "hash", games_num=2)
    assert "repetition" in how and s["oracle.repetition_hits"] > 0


def test_chess_games_with_noise_temperature_and_cache_match_oracle():
    cfg = chess_cfg(sim_num=10, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[30, 1.0], [9999, 0.0]],
                    cache_size=50000, seed=77)
    s, _, _ = check_chess_against_oracle(cfg, "coarse", games_num=2)
    assert s["metrics"]["cache.hits"] > 0


def test_chess_uniform_net_ties_match_oracle():
    check_chess_against_oracle(chess_cfg(sim_num=8, temperature_policy=[[9999, 1.0]], seed=9), "uniform", games_num=2)


def test_chess_results_do_not_depend_on_threads_or_batching():
    net = chess_fake_net("hash")
    kw = dict(sim_num=8, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[20, 1.0], [9999, 0.0]])
    _, ref = SelfPlayRunner("chess", chess_cfg(**kw)).run_with(chess_cb(net), None, 6, keep_records=True)
    for threads, gpt, cache in ((3, 1, 0), (2, 3, 1000), (1, 6, 7)):
        summary, got = SelfPlayRunner("chess", chess_cfg(threads=threads, games_per_thread=gpt, cache_size=cache, **kw)).run_with(
            chess_cb(net), None, 6, keep_records=True)
        assert [(r.game_idx, r.moves, r.winner, r.entries) for r in got] == [(r.game_idx, r.moves, r.winner, r.entries) for r in ref]


def test_chess_max_moves_bounds_a_game_like_the_oracle():
    """`max_moves` (bench / tool option, not in the reference): the game stops as a draw after that many moves."""
    net = chess_fake_net("hash")
    cfg = chess_cfg(sim_num=6, max_moves=5)
    summary, records = SelfPlayRunner("chess", cfg).run_with(chess_cb(net), None, 2, keep_records=True)
    ev = om.Evaluator(chess_oracle_fn(net), None)
    ref = [om.play_game(g, oc.ChessPosition.new, _params(cfg), _params(cfg), ev, ev, cfg["seed"], max_moves=5) for g in range(2)]
    for rec, o in zip(records, ref):
        assert len(rec.moves) == 5 and rec.moves == [oc.move_to_u16(m) for m in o.moves] and rec.winner is None and o.winner is None
    assert summary["draws"] == 2


# --------------------------------------------------------------------------------------------------------------
# one search at a time: the UCI loop's player (cattus_b200_chess_search_*, cattus_b200/uci.py)
# --------------------------------------------------------------------------------------------------------------
def test_search_session_matches_oracle_player_with_tree_reuse():
    """`ucinewgame`, then `position ... moves ...` + `go` with the engine's move and a reply appended each time: the
    tree of the previous search is found two plies down and reused (mcts/mod.rs:335-352).  Same moves as the oracle's
    MctsPlayer fed the same history."""
    from cattus_b200.selfplay import ChessSearch, move_from_lan

    net = chess_fake_net("hash")
    cfg = chess_cfg(sim_num=30, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], cache_size=10000, seed=13)
    player = om.MctsPlayer(_params(cfg), om.Evaluator(chess_oracle_fn(net), om.ValueFuncCache(10000)), om.SplitMix64(om.game_seed(13, 0)))
    rng = random.Random(2)
    fen = KIWIPETE
    history = [oc.ChessPosition.from_fen(fen)]
    moves = []
    with ChessSearch(cfg, eval_fn=chess_cb(net)) as search:
        for _ in range(6):
            best, stats = search.go(fen, moves)
            probs = player.calc_moves_probabilities(history)
            want = player.choose_move_from_probabilities(history, probs)
            assert best == oc.move_to_lan(want)
            assert stats["simulations"] == 30 and stats["root_children"] == len(history[-1].legal_moves())
            moves.append(best)
            history.append(history[-1].moved_position(want))
            if history[-1].is_finished():
                break
            reply = rng.choice(history[-1].legal_moves())
            moves.append(oc.move_to_lan(reply))
            history.append(history[-1].moved_position(reply))
        assert move_from_lan("e7e8q") == 52 | (60 << 6) | (1 << 12)
        with pytest.raises(Exception, match="not legal"):
            search.go(None, ["e2e5"])
        with pytest.raises(Exception, match="over"):
            search.go("7k/5Q2/6K1/8/8/8/8/8 b - -", [])  # stalemate: nothing to search


def test_uci_loop_commands():
    """The command set of engine/src/chess/uci.rs:27-73 over a deterministic evaluator."""
    import io

    from cattus_b200.uci import UCI

    out = io.StringIO()
    cfg = chess_cfg(sim_num=20, seed=3)
    uci = UCI(cfg, eval_fn=chess_cb(chess_fake_net("hash")), out=out)
    script = ["uci", "isready", "setoption name Hash value 16", "ucinewgame", "position startpos moves e2e4 e7e5", "go movetime 1000",
              "position fen r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - moves e1g1", "go", "stop", "bogus", "quit", "go"]
    uci.run(io.StringIO("\n".join(script) + "\n"))
    lines = out.getvalue().split("\n")
    assert lines[:4] == ["id name cattus_b200 v1.0.0", "id author cattus_b200", "uciok", "readyok"]
    best = [ln.split()[1] for ln in lines if ln.startswith("bestmove")]
    assert len(best) == 2  # nothing after quit
    start = oc.ChessPosition.new()
    for lan in ("e2e4", "e7e5"):
        start = start.moved_position(next(m for m in start.legal_moves() if oc.move_to_lan(m) == lan))
    assert best[0] in [oc.move_to_lan(m) for m in start.legal_moves()]
    kiwi = oc.ChessPosition.from_fen(KIWIPETE)
    kiwi = kiwi.moved_position(next(m for m in kiwi.legal_moves() if oc.move_to_lan(m) == "e1g1"))
    assert best[1] in [oc.move_to_lan(m) for m in kiwi.legal_moves()]
    assert uci.options == {"Hash": "16"} and uci.last_stats["simulations"] == 20


def test_host_mirror_position_from_fen_matches_oracle():
    """cattus_b200.games.ChessPosition.from_fen: what CudaNetwork.evaluate needs (real planes, the legal moves of the
    evaluated view with their nn indices), against the oracle for a white-to-move and a black-to-move position."""
    from cattus_b200.games import ChessPosition

    for fen, moves in ((KIWIPETE, []), (START, [oc.move_to_u16((12, 28, None))])):  # the second: after 1. e4, black to move
        host = ChessPosition.from_fen(fen, moves)
        p = oc.ChessPosition.from_fen(fen)
        for m in moves:
            p = p.moved_position(oc.move_from_u16(m))
        assert host.turn == p.turn and list(host.planes) == p.planes()
        view = p if p.turn == oc.P1 else p.flipped()
        assert host.legal_moves() == view.legal_moves()
        assert [host.move_to_nn_idx(m) for m in host.legal_moves()] == [oc.ChessPosition.to_nn_idx(m) for m in view.legal_moves()]
        assert list((host if host.turn == 1 else host.flipped()).planes) == view.planes()


def test_chess_partition_by_stride_is_the_union():
    """Multi-GPU arrangement (first_game = rank, game_stride = world): the ranks' games are exactly the whole job's."""
    net = chess_fake_net("hash")
    cfg = chess_cfg(sim_num=6, max_moves=10, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[9999, 1.0]])
    _, whole = SelfPlayRunner("chess", cfg).run_with(chess_cb(net), None, 4, keep_records=True)
    parts = []
    for r in range(2):
        _, p = SelfPlayRunner("chess", cfg).run_with(chess_cb(net), None, 4, keep_records=True, first_game=r, game_stride=2)
        assert [x.game_idx for x in p] == [r, r + 2]
        parts += p
    parts.sort(key=lambda x: x.game_idx)
    assert [(r.moves, r.entries) for r in parts] == [(r.moves, r.entries) for r in whole]


def test_speculative_evaluation_does_not_change_the_search():
    """cfg.speculate: extra positions ride along with the waiting leaf and only land in the cache, so the chosen moves
    and the simulation counts are those of the plain search; fewer evaluator CALLS are needed for the same search."""
    from cattus_b200.selfplay import ChessSearch

    net = chess_fake_net("hash")
    calls = {"plain": 0, "spec": 0}

    def counting(tag):
        inner = chess_cb(net)

        def cb(words, n, legal):
            calls[tag] += 1
            return inner(words, n, legal)

        return cb

    base = dict(sim_num=300, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, cache_size=100000, seed=8)
    moves, out = [], {}
    for tag, speculate in (("plain", 0), ("spec", 7)):
        got = []
        with ChessSearch(chess_cfg(speculate=speculate, **base), eval_fn=counting(tag)) as search:
            played = []
            for _ in range(3):
                best, stats = search.go(KIWIPETE, played)
                got.append((best, stats["simulations"], stats["terminal_leaves"]))
                played = played + [best]
                reply = oc.ChessPosition.from_fen(KIWIPETE)
                for lan in played:
                    reply = reply.moved_position(next(m for m in reply.legal_moves() if oc.move_to_lan(m) == lan))
                played.append(oc.move_to_lan(reply.legal_moves()[0]))
            out[tag] = (got, stats)
    assert out["plain"][0] == out["spec"][0]
    assert out["plain"][1]["speculative_evaluations"] == 0 and out["spec"][1]["speculative_evaluations"] > 0
    assert calls["spec"] < 0.8 * calls["plain"], calls


def test_chess_speculation_in_self_play_same_games_fewer_calls():
    net = chess_fake_net("hash")
    kw = dict(sim_num=40, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], cache_size=100000,
              threads=2, games_per_thread=2, max_moves=12)
    s0, plain = SelfPlayRunner("chess", chess_cfg(**kw)).run_with(chess_cb(net), None, 4, keep_records=True)
    s1, spec = SelfPlayRunner("chess", chess_cfg(speculate=8, **kw)).run_with(chess_cb(net), None, 4, keep_records=True)
    assert [(r.game_idx, r.moves, r.winner, r.entries) for r in spec] == [(r.game_idx, r.moves, r.winner, r.entries) for r in plain]
    assert s1["metrics"]["selfplay.speculative_evaluations"] > 0
    assert s1["metrics"]["model.activation_count"] < 0.8 * s0["metrics"]["model.activation_count"]


# --------------------------------------------------------------------------------------------------------------
# the reference's own chess rule tests (engine/src/chess/core.rs:617-730), restated for both rule sets
# --------------------------------------------------------------------------------------------------------------
def _san_to_move(pos: oc.ChessPosition, san: str):
    """Enough SAN for the reference's test games: piece letter, optional from-file for pawn captures, destination."""
    san = san.rstrip("+#")
    kind = san[0].lower() if san[0] in "NBRQK" else "p"
    dest = og.chess_move_to_idx("a1" + san[-2:]) % 64
    from_file = ord(san[0]) - 97 if kind == "p" and "x" in san else None
    found = [m for m in pos.legal_moves()
             if m[1] == dest and pos.board[m[0]].lower() == kind and (from_file is None or m[0] % 8 == from_file)]
    assert len(found) == 1, (san, found)
    return found[0]


def _play_san(moves):
    """Plays the SAN moves on the oracle and on the library's rules; yields (oracle position, library info) BEFORE each move
    and finally after the last one."""
    pos = oc.ChessPosition.new()
    played = []
    for san in moves:
        yield pos, chess_info(START, played)
        m = _san_to_move(pos, san)
        played.append(oc.move_to_u16(m))
        pos = pos.moved_position(m)
    yield pos, chess_info(START, played)


def test_reference_simple_game_and_mate():
    """core.rs:617-631: ongoing before every move, Finished(Some(Player1)) at the end."""
    moves = ["e4", "e5", "d4", "exd4", "Qxd4", "Nc6", "Qa4", "a6", "Bg5", "h6", "Bc4", "Rb8", "Qb3", "Ra8", "Bxf7"]
    states = list(_play_san(moves))
    for pos, info in states[:-1]:
        assert pos.status() == ("ongoing", None) and info.status == 0
    pos, info = states[-1]
    assert pos.status() == ("finished", oc.P1) and info.status == 1 and info.n_legal == 0 and info.in_check == 1


def test_reference_fifty_rule_count():
    """core.rs:633-670: two pawn moves, then the kings shuffle; the 50th quiet white move ends the game as a draw, and not
    one move earlier (the count only advances on white's moves, core.rs:334-343; positions repeat, but repetition is the
    game's business, not the position's)."""
    moves = ["e4", "e5"] + ["Ke2", "Ke7", "Ke1", "Ke8"] * 24 + ["Ke2", "Ke7", "Ke1"]
    states = list(_play_san(moves))
    for pos, info in states[:-1]:
        assert pos.status() == ("ongoing", None) and info.status == 0
    pos, info = states[-1]
    assert pos.status() == ("finished", None) and pos.fifty == 50
    assert info.status == 3 and info.fifty_rule_count == 50 and info.n_legal == 0


def test_reference_flip_positions():
    """core.rs:672-693: the flipped position has the other side to move and flips back to the original; here also: the
    library's view of a position (side to move plays white) is the oracle's flipped position for black to move."""
    for fen in ["7r/2B3n1/K5R1/3nPP2/P1k2Pp1/4p1p1/2p4P/8 w - - 0 1", "8/5pB1/1P4P1/1p3q2/BK1pP2P/pQ1pP3/7k/8 w - - 0 1",
                "2k5/2b1p1P1/1p5P/P1pnp3/QPK1b2p/8/5r2/8 b - - 0 1", "2b5/3N4/2B4p/1KP1R2r/n5p1/3N3P/k2B3b/q7 w - - 0 1",
                "8/2P1r3/4k1p1/4n1p1/3Rq3/P3Pp2/P4PP1/5K1N b - - 0 1", "1N6/5pP1/P1pK3P/P3p3/2R3p1/k3pQ2/6r1/6N1 b - - 0 1",
                "1B6/1r6/Q1ppKP2/qP1P1N1k/3p4/1pb5/7p/8 w - - 0 1", "3r4/1b2p3/7k/1P3R2/K4nrN/1N5P/n1Pp3P/8 b - - 0 1"]:
        pos = oc.ChessPosition.from_fen(fen)
        assert pos.flipped().turn == 3 - pos.turn and pos.flipped().flipped() == pos
        info = chess_info(fen)
        view = pos if pos.turn == oc.P1 else pos.flipped()
        assert info.turn == pos.turn and list(info.planes) == view.planes()
        assert {oc.move_from_u16(info.moves[i]) for i in range(info.n_legal)} == set(pos.legal_moves())


def test_uci_plays_a_whole_game_with_legal_moves():
    """training/tests/test_uci.py drives the reference's UCI binary with python-chess until the game is over; here the
    oracle's rules play that part: every `bestmove` must be legal, the tree is reused from `go` to `go`."""
    import io

    from cattus_b200.uci import UCI

    out = io.StringIO()
    uci = UCI(chess_cfg(sim_num=25, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, temperature_policy=[[9999, 1.0]], cache_size=100000, seed=2),
              eval_fn=chess_cb(chess_fake_net("hash")), out=out)
    assert uci.cfg["speculate"] == 31
    uci.handle("uci")
    uci.handle("ucinewgame")
    board = oc.ChessPosition.new()
    seen = {board: 1}
    moves = []
    for _ply in range(80):
        if board.is_finished() or seen[board] >= 3:
            break
        uci.handle("position startpos" + (" moves " + " ".join(moves) if moves else ""))
        uci.handle("go movetime 20000")
        best = out.getvalue().strip().split("\n")[-1].split()[1]
        legal = {oc.move_to_lan(m): m for m in board.legal_moves()}
        assert best in legal, (best, sorted(legal))
        moves.append(best)
        board = board.moved_position(legal[best])
        seen[board] = seen.get(board, 0) + 1
    assert len(moves) >= 20
    uci.handle("quit")


def test_chess_entries_parse_like_the_trainer_parser():
    """The consumer's side of the wire format: training/cattus_train/chess.py:22-49 (`Chess.load_data_entry`) restated without
    `construct` -- 18 u64 planes, 235-byte bitmap unpacked little-endian bit order, the first popcount probabilities scattered
    to the set indices, i8 winner -- applied to entries the driver produced."""
    import struct

    net = chess_fake_net("hash")
    cfg = chess_cfg(sim_num=10, max_moves=9, temperature_policy=[[9999, 1.0]], seed=4)
    _, records = SelfPlayRunner("chess", cfg).run_with(chess_cb(net), None, 2, keep_records=True)
    for rec in records:
        pos = oc.ChessPosition.new()
        for k, entry in enumerate(rec.entries):
            assert len(entry) == 18 * 8 + 235 + 225 * 4 + 1
            planes = np.array(struct.unpack("<18Q", entry[:144]), dtype=np.uint64)
            moves_bitmap = np.frombuffer(entry[144:379], dtype=np.uint8)
            probs = np.frombuffer(entry[379:1279], dtype="<f4")
            winner = float(struct.unpack("<b", entry[1279:])[0])
            probs_all = np.full((1880,), -1.0, dtype=np.float32)
            move_indices = np.where(np.unpackbits(moves_bitmap, count=1880, bitorder="little"))[0]
            probs_all[move_indices] = probs[: len(move_indices)]
            view = pos if pos.turn == oc.P1 else pos.flipped()  # entries are always written as Player1 to move
            assert planes.tolist() == view.planes()
            assert sorted(move_indices.tolist()) == sorted(oc.ChessPosition.to_nn_idx(m) for m in view.legal_moves())
            assert abs(float(probs_all[move_indices].sum()) - 1.0) < 1e-5 and (probs[len(move_indices):] == -1.0).all()
            assert winner == 0.0  # stopped by max_moves: recorded as a draw
            pos = pos.moved_position(oc.move_from_u16(rec.moves[k]))


@pytest.mark.parametrize("trial", range(8))
def test_speculation_never_changes_a_game_randomised(trial):
    """Random draws over game, evaluator (incl. exact ties), noise, temperature, cache size (incl. tiny caches that evict all the
    time), one or two evaluators, threads x games per thread and rows per game: with and without cfg.speculate the records are
    byte-identical."""
    from tests.test_selfplay_cpu import cb_for, cfg_with, fake_net

    rng = random.Random(100 + trial)
    game = rng.choice(["hex4", "hex5", "hex7", "ttt", "chess"])
    kind = rng.choice(["hash", "coarse", "uniform"])
    kw = dict(sim_num=rng.choice([8, 30, 60]), prior_noise_alpha=rng.choice([0.0, 0.3]), prior_noise_epsilon=0.25,
              temperature_policy=rng.choice([[[9999, 0.0]], [[3, 1.0], [9999, 0.0]], [[9999, 1.0]]]), cache_size=rng.choice([50, 5000, 100000]),
              threads=rng.choice([1, 2, 3]), games_per_thread=rng.choice([1, 2, 5]), seed=rng.randrange(1000))
    two = rng.random() < 0.3
    spec = rng.choice([1, 4, 9, 31])
    if game == "chess":
        net1, net2 = chess_fake_net(kind), chess_fake_net(kind, salt=7)

        def run(s):
            return SelfPlayRunner("chess", chess_cfg(speculate=s, max_moves=10, **kw)).run_with(chess_cb(net1), chess_cb(net2) if two else None, 4,
                                                                                                 keep_records=True)
    else:
        wpp = 1 if game == "ttt" else (int(game[3:]) ** 2 + 63) // 64
        net1, net2 = fake_net(kind), fake_net(kind, salt=7)

        def run(s):
            return SelfPlayRunner(game, cfg_with(speculate=s, **kw)).run_with(cb_for(net1, wpp), cb_for(net2, wpp) if two else None, 4, keep_records=True)

    _, plain = run(0)
    summary, speculating = run(spec)
    assert [(r.game_idx, r.moves, r.winner, r.entries) for r in speculating] == [(r.game_idx, r.moves, r.winner, r.entries) for r in plain]
    assert summary["metrics"]["selfplay.speculative_evaluations"] > 0
