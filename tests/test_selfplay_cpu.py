"""CPU tests of the self-play driver (csrc/selfplay.cpp through include/cattus_b200_selfplay.h) against oracle/mcts.py.

No GPU: the driver is bound to a deterministic evaluator callback (the reference's `dyn ValueFunction` seam), and the
oracle's MctsPlayer is driven by the very same function, so whole games -- every move, every visit distribution, every
.traindata byte -- must be identical.  Also pins the oracle's Hex rules to the reference's own rule tests
(engine/src/hex/core.rs:386-490).
"""
from __future__ import annotations

import math
import struct

import numpy as np
import pytest

from cattus_b200.selfplay import SelfPlayRunner, SelfPlayError
from oracle import mcts as om

# --------------------------------------------------------------------------------------------------------------
# the oracle's rules against the reference's rule tests
# --------------------------------------------------------------------------------------------------------------


def _hex_from_str(s: str, size: int = 11) -> om.HexPosition:
    s = "".join(s.split())
    n = size * size
    red = sum(1 << i for i, ch in enumerate(s[:n]) if ch == "r")
    blue = sum(1 << i for i, ch in enumerate(s[:n]) if ch == "b")
    return om.HexPosition.from_board(size, red, blue, om.P1 if s[n] == "r" else om.P2)


def _diag(ch: str, anti: bool, drop_first=False, drop_last=False) -> str:
    rows = []
    for r in range(11):
        row = ["e"] * 11
        c = 10 - r if anti else r
        if not ((drop_first and r == 0) or (drop_last and r == 10)):
            row[c] = ch
        rows.append("".join(row))
    return "".join(rows)


def test_oracle_hex_rules_match_reference_tests():
    # short_diagonal_wins (hex/core.rs:386-420)
    assert _hex_from_str(_diag("r", False) + "b").status() == ("finished", om.P1)
    assert _hex_from_str(_diag("b", False) + "r").status() == ("finished", om.P2)
    # almost_short_diagonal_doesnt_win (:422-455)
    assert _hex_from_str(_diag("r", False, drop_first=True) + "b").status() == ("ongoing", None)
    assert _hex_from_str(_diag("b", False, drop_last=True) + "r").status() == ("ongoing", None)
    # long_diagonal_doesnt_win (:457-490)
    assert _hex_from_str(_diag("r", True) + "b").status() == ("ongoing", None)
    assert _hex_from_str(_diag("b", True) + "r").status() == ("ongoing", None)
    # flip (:492-511)
    pos = _hex_from_str("eebeeeeeeer eeeeeeeeeee eeeebeeeree eeeeeeereee eeeeeereeee eeeeereeeee eeeerebeeee eeereeeeeee eereeereeee ereeeeeeeee reeeeebeeee b")
    assert pos.turn == om.P2 and pos.flipped().turn == om.P1 and pos.flipped().flipped() == pos


def test_oracle_hex_flip_rand():
    """flip_rand (hex/core.rs:513-548) with a seeded random player."""
    rng = np.random.default_rng(7)
    for _ in range(10):
        pos = om.HexPosition.new(11)
        while not pos.is_finished():
            t = pos.flipped()
            assert t.flipped() == pos
            assert set(pos.legal_moves()) == {t.flip_move(m) for m in t.legal_moves()}
            a, b = pos.status(), t.status()
            assert a[0] == b[0]
            moves = pos.legal_moves()
            pos = pos.moved_position(moves[int(rng.integers(len(moves)))])
        w, wt = pos.status()[1], pos.flipped().status()[1]
        assert wt == (None if w is None else om.opposite(w))


def test_oracle_incremental_reach_equals_from_board():
    rng = np.random.default_rng(3)
    for s in (4, 5, 7):
        for _ in range(20):
            pos = om.HexPosition.new(s)
            while not pos.is_finished():
                moves = pos.legal_moves()
                pos = pos.moved_position(moves[int(rng.integers(len(moves)))])
                if not pos.is_finished():
                    assert om.HexPosition.from_board(s, pos.red, pos.blue, pos.turn) == pos


# --------------------------------------------------------------------------------------------------------------
# a deterministic "network": the same function drives the oracle and (through the callback) the C++ driver
# --------------------------------------------------------------------------------------------------------------
def _mix(x: int) -> int:
    x &= (1 << 64) - 1
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & ((1 << 64) - 1)
    x ^= x >> 33
    return x


def fake_net(kind: str, salt: int = 0):
    """Returns net(words_of_one_position) -> (probs over the legal cells ascending, value)."""

    def net(a: int, b: int, ones: int):
        legal = [i for i in range(ones.bit_length()) if (ones >> i) & 1 and not ((a | b) >> i) & 1]
        if kind == "uniform":  # exact ties everywhere: exercises the petgraph edge-order / max_by rules
            p = np.full(len(legal), np.float32(1.0) / np.float32(len(legal)), dtype=np.float32)
            return p, np.float32(0.0)
        h = _mix(a * 0x9E3779B97F4A7C15 + _mix(b + salt) + salt)
        logits = np.array([((_mix(h + 977 * i) % 1000) / 250.0) - 2.0 for i in legal], dtype=np.float32)
        e = np.exp(logits - logits.max()).astype(np.float32)
        p = (e / e.sum(dtype=np.float32)).astype(np.float32)
        v = np.float32(((h >> 7) % 2001) / 1000.0 - 1.0)
        if kind == "coarse":  # few distinct values: ties between some children but not all
            v = np.float32(round(float(v) * 2) / 2)
        return p, v

    return net


def _words_to_ints(row: np.ndarray, wpp: int):
    vals = []
    for c in range(3):
        v = 0
        for k in range(wpp):
            v |= int(row[c * wpp + k]) << (64 * k)
        vals.append(v)
    return vals


def cb_for(net, wpp: int):
    def cb(words: np.ndarray, n: int):
        probs, values = [], []
        for i in range(n):
            a, b, ones = _words_to_ints(words[i], wpp)
            p, v = net(a, b, ones)
            probs.append(p)
            values.append(v)
        return probs, values

    return cb


def oracle_net_fn(net):
    def fn(pos):
        pl = pos.planes()
        p, v = net(pl[0], pl[1], pl[2])
        return list(p), v

    return fn


def oracle_games(game: str, cfg: dict, net1, net2, games):
    mc = cfg["mcts"]
    params = om.MctsParams(sim_num=mc["sim_num"], explore_factor=mc.get("explore_factor", math.sqrt(2.0)),
                           temperature=om.TemperaturePolicy.from_config(mc.get("temperature_policy", [[0, 1.0]])),
                           prior_noise_alpha=mc.get("prior_noise_alpha", 0.0), prior_noise_epsilon=mc.get("prior_noise_epsilon", 0.0))
    cache1 = om.ValueFuncCache(mc["cache_size"]) if mc.get("cache_size") else None
    e1 = om.Evaluator(oracle_net_fn(net1), cache1)
    if net2 is None:
        e2 = e1
    else:
        e2 = om.Evaluator(oracle_net_fn(net2), om.ValueFuncCache(mc["cache_size"]) if mc.get("cache_size") else None)
    if game == "ttt":
        new_pos = om.TttPosition.new
    else:
        s = int(game[3:])
        new_pos = lambda: om.HexPosition.new(s)  # noqa: E731
    return [om.play_game(g, new_pos, params, params, e1, e2, cfg.get("seed", 0)) for g in games], (e1, e2)


def check_against_oracle(game: str, cfg: dict, kind1: str, kind2=None, games_num: int = 4):
    wpp = 1 if game == "ttt" else (int(game[3:]) ** 2 + 63) // 64
    net1 = fake_net(kind1)
    net2 = fake_net(kind2, salt=99) if kind2 else None
    runner = SelfPlayRunner(game, cfg)
    summary, records = runner.run_with(cb_for(net1, wpp), cb_for(net2, wpp) if net2 else None, games_num, keep_records=True)
    ref, _ = oracle_games(game, cfg, net1, net2, range(games_num))
    assert len(records) == games_num
    for rec, o in zip(records, ref):
        assert rec.game_idx == o.game_idx
        assert rec.moves == o.moves, (rec.game_idx, rec.moves, o.moves)
        assert rec.winner == o.winner
        assert len(rec.entries) == len(o.entries)
        for k, (pos, probs) in enumerate(o.entries):
            assert rec.entries[k] == om.data_entry_bytes(pos, probs, o.winner), (rec.game_idx, k)
            assert rec.entry_dirs[k] == om.data_entry_dir(pos.turn, o.game_idx)
    m = summary["metrics"]
    assert m["selfplay.simulations"] == sum(o.sims for o in ref) == m["selfplay.searches"] * cfg["mcts"]["sim_num"]
    assert summary["player1_wins"] + summary["player2_wins"] + summary["draws"] == games_num
    return summary, records, ref


BASE = {"mcts": {"sim_num": 40, "explore_factor": 1.41421, "temperature_policy": [[9999, 0.0]], "prior_noise_alpha": 0.0,
                 "prior_noise_epsilon": 0.0, "cache_size": 0}, "threads": 1, "games_per_thread": 1, "seed": 5}


def cfg_with(**kw):
    c = {"mcts": dict(BASE["mcts"]), "threads": BASE["threads"], "games_per_thread": BASE["games_per_thread"], "seed": BASE["seed"]}
    for k, v in kw.items():
        if k in c["mcts"]:
            c["mcts"][k] = v
        else:
            c[k] = v
    return c


@pytest.mark.parametrize("kind", ["uniform", "coarse", "hash"])
def test_hex4_games_match_oracle(kind):
    check_against_oracle("hex4", cfg_with(), kind)


def test_hex5_tree_reuse_and_cache_match_oracle():
    s, _, _ = check_against_oracle("hex5", cfg_with(sim_num=60, cache_size=100000), "coarse", games_num=2)
    assert s["metrics"]["cache.hits"] > 0 and s["metrics"]["cache.misses"] > 0


def test_hex9_two_word_planes_match_oracle():
    check_against_oracle("hex9", cfg_with(sim_num=12), "hash", games_num=2)


def test_noise_and_temperature_match_oracle():
    cfg = cfg_with(prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[3, 1.0], [5, 0.5], [9999, 0.0]], seed=1234)
    check_against_oracle("hex4", cfg, "hash", games_num=6)


def test_model_compare_two_evaluators_match_oracle():
    s, _, _ = check_against_oracle("hex4", cfg_with(cache_size=1000), "hash", kind2="hash", games_num=4)
    assert s["player1_wins"] + s["player2_wins"] == 4


def test_ttt_games_match_oracle_including_draws():
    s, records, _ = check_against_oracle("ttt", cfg_with(sim_num=50), "uniform", games_num=2)
    check_against_oracle("ttt", cfg_with(sim_num=30, prior_noise_alpha=1.0, prior_noise_epsilon=0.5, temperature_policy=[[9999, 1.0]]), "hash", games_num=8)


def test_results_do_not_depend_on_threads_batching_or_cache():
    wpp = 1
    net = fake_net("hash")
    base = cfg_with(prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]])
    _, ref = SelfPlayRunner("hex5", base).run_with(cb_for(net, wpp), None, 12, keep_records=True)
    for threads, gpt, cache in ((3, 1, 0), (2, 4, 0), (1, 12, 5000), (4, 2, 7)):
        cfg = cfg_with(prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], threads=threads,
                       games_per_thread=gpt, cache_size=cache)
        summary, got = SelfPlayRunner("hex5", cfg).run_with(cb_for(net, wpp), None, 12, keep_records=True)
        assert [(r.game_idx, r.moves, r.winner, r.entries) for r in got] == [(r.game_idx, r.moves, r.winner, r.entries) for r in ref]
        if gpt > 1:
            assert summary["metrics"]["selfplay.evaluations"] / summary["metrics"]["model.activation_count"] > 1.5  # real batches


def test_partition_by_stride_is_the_union():
    net = fake_net("hash")
    cfg = cfg_with()
    _, whole = SelfPlayRunner("hex4", cfg).run_with(cb_for(net, 1), None, 6, keep_records=True)
    parts = []
    for r in range(2):
        _, p = SelfPlayRunner("hex4", cfg).run_with(cb_for(net, 1), None, 6, keep_records=True, first_game=r, game_stride=2)
        assert [x.game_idx for x in p] == [r, r + 2, r + 4]
        parts += p
    parts.sort(key=lambda x: x.game_idx)
    assert [(r.moves, r.entries) for r in parts] == [(r.moves, r.entries) for r in whole]


def test_traindata_files(tmp_path):
    net = fake_net("hash")
    d1, d2 = tmp_path / "a" / "p1", tmp_path / "b" / "p2"
    _, recs = SelfPlayRunner("hex4", cfg_with()).run_with(cb_for(net, 1), None, 2, d1, d2, keep_records=True)
    n_files = 0
    for r in recs:
        for k, (e, d) in enumerate(zip(r.entries, r.entry_dirs)):
            f = (d1 if d == 1 else d2) / f"{r.game_idx:08d}_{k:03d}.traindata"
            assert f.read_bytes() == e
            n_files += 1
            # layout: 3 planes x (lo, hi) u64 | 16 x f32 | i8  (self_play.rs:33-61, serialize/hex.rs:16-28)
            assert len(e) == 6 * 8 + 16 * 4 + 1
            planes = struct.unpack("<6Q", e[:48])
            probs = np.frombuffer(e[48:112], "<f4")
            assert planes[4] == 0xFFFF and planes[5] == 0 and planes[1] == planes[3] == 0
            occ = planes[0] | planes[2]
            for i in range(16):
                assert (probs[i] == -1.0) == bool((occ >> i) & 1)
            assert abs(float(probs[probs >= 0].sum()) - 1.0) < 1e-5
            assert struct.unpack("<b", e[112:])[0] in (-1, 1)
    assert n_files == len(list(d1.glob("*.traindata"))) + len(list(d2.glob("*.traindata")))


def test_bad_arguments_are_reported():
    net = fake_net("hash")
    with pytest.raises(SelfPlayError, match="multiple of 2"):
        SelfPlayRunner("hex4", cfg_with()).run_with(cb_for(net, 1), None, 3)
    with pytest.raises(SelfPlayError, match="sim_num"):
        SelfPlayRunner("hex4", cfg_with(sim_num=1)).run_with(cb_for(net, 1), None, 2)
    with pytest.raises(ValueError):
        SelfPlayRunner("go", cfg_with())

    def broken(words, n):
        return [np.zeros(1, np.float32)] * n, [0.0] * n

    with pytest.raises(SelfPlayError, match="wrong number"):
        SelfPlayRunner("hex4", cfg_with()).run_with(broken, None, 2)


def test_self_player_cli_rejects_other_engines_and_foreign_models(tmp_path):
    import json

    from cattus_b200 import self_player
    from tests.util import blob

    model = tmp_path / "m.cb2"
    model.write_bytes(blob("hex4"))
    assert self_player.game_of_blob(model) == ("hex4", "hex")
    (tmp_path / "c.json").write_text(json.dumps({"mcts": {"sim_num": 10}, "model": {"batch_size": 4, "inference": {"engine": "onnx-ort"}}, "threads": 1}))
    argv = [f"--model1-path={model}", f"--model2-path={model}", "--games-num=2", f"--out-dir1={tmp_path}", f"--out-dir2={tmp_path}",
            f"--config-file={tmp_path / 'c.json'}"]
    with pytest.raises(SystemExit, match="no CPU fallback"):
        self_player.main(argv)
    bad = tmp_path / "bad.cb2"
    bad.write_bytes(b"\0" * 128)
    with pytest.raises(ValueError, match="not a .cb2"):
        self_player.game_of_blob(bad)
    chess = tmp_path / "chess.cb2"
    chess.write_bytes(blob("chess_dev"))
    assert self_player.game_of_blob(chess) == ("chess", "chess")


@pytest.mark.parametrize("game,threads,gpt", [("hex5", 1, 1), ("hex5", 2, 3), ("hex9", 1, 2), ("ttt", 1, 2)])
def test_speculation_in_self_play_changes_nothing_but_the_number_of_calls(game, threads, gpt):
    """cfg.speculate in the self-play driver (trainer-sized jobs: few games per thread, so batches are tiny and the device
    time is flat): likely next leaves ride along into the cache.  Same games, same .traindata bytes, fewer evaluator calls."""
    wpp = 1 if game == "ttt" else (int(game[3:]) ** 2 + 63) // 64
    net = fake_net("hash")
    kw = dict(sim_num=80, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], cache_size=100000,
              threads=threads, games_per_thread=gpt)
    s0, plain = SelfPlayRunner(game, cfg_with(**kw)).run_with(cb_for(net, wpp), None, 6, keep_records=True)
    s1, spec = SelfPlayRunner(game, cfg_with(speculate=8, **kw)).run_with(cb_for(net, wpp), None, 6, keep_records=True)
    assert [(r.game_idx, r.moves, r.winner, r.entries) for r in spec] == [(r.game_idx, r.moves, r.winner, r.entries) for r in plain]
    m0, m1 = s0["metrics"], s1["metrics"]
    assert m0["selfplay.speculative_evaluations"] == 0 and m1["selfplay.speculative_evaluations"] > 0
    assert m1["selfplay.simulations"] == m0["selfplay.simulations"]
    if game != "ttt":  # tic-tac-toe trees are exhausted within a few dozen simulations either way
        assert m1["model.activation_count"] < 0.8 * m0["model.activation_count"], (m0["model.activation_count"], m1["model.activation_count"])


def test_speculation_with_two_evaluators_and_groups_of_games():
    """Model comparison (two networks, two caches): a game's speculative rows go to the evaluator and cache of the side
    that is searching."""
    net1, net2 = fake_net("hash"), fake_net("hash", salt=99)
    kw = dict(sim_num=60, cache_size=50000, threads=2, games_per_thread=3, prior_noise_alpha=0.3, prior_noise_epsilon=0.25)
    _, plain = SelfPlayRunner("hex5", cfg_with(**kw)).run_with(cb_for(net1, 1), cb_for(net2, 1), 6, keep_records=True)
    s, spec = SelfPlayRunner("hex5", cfg_with(speculate=6, **kw)).run_with(cb_for(net1, 1), cb_for(net2, 1), 6, keep_records=True)
    assert [(r.game_idx, r.moves, r.winner, r.entries) for r in spec] == [(r.game_idx, r.moves, r.winner, r.entries) for r in plain]
    assert s["metrics"]["selfplay.speculative_evaluations"] > 0


def test_oracle_ttt_rules_match_reference_tests():
    """engine/src/ttt/core.rs:290-339: simple_game_and_mate and flip."""
    from oracle import games as og

    def to_pos(s: str) -> om.TttPosition:
        x, o, turn = og.ttt_position_from_str(s)
        return om.TttPosition(x, o, turn, om.TttPosition._winner(x, o))

    for s, winner in (("xxxoo____o", om.P1), ("oo_xxx___o", om.P1), ("oo____xxxo", om.P1), ("oxxo__ox_x", om.P2), ("xox_o_xo_x", om.P2),
                      ("xxo__o_xox", om.P2), ("xxoooxxxoo", None)):
        assert to_pos(s).status() == ("finished", winner), s
    for s in ("oxx_o_o__o", "o_____xx_o", "xx_xx_xo_o", "ox___x_xox", "_x_o__o_xo", "ox__o____x", "_o__o_oxxo", "__xx_x__ox"):
        pos = to_pos(s)
        assert pos.flipped().turn == om.opposite(pos.turn) and pos.flipped().flipped() == pos
