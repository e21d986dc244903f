"""CPU tests of the DEVICE-RESIDENT search's logic (cattus_b200/csrc/dsearch_core.hpp + dsearch_host.hpp).

The per-slot code that runs one warp per game on the GPU is compiled for the host with a one-lane warp into a test-only
library (tests/emul) and driven by the same host driver as on the GPU.  Whole games must equal the HOST driver's
(csrc/selfplay.cpp, itself pinned to oracle/mcts.py by tests/test_selfplay_cpu.py and tests/test_chess_cpu.py): every
move, winner and .traindata byte, for any number of concurrent slots and any pipeline depth.  The cross-lane parts
(shuffles, __syncwarp) are covered by the -m gpu tests.
"""
import pytest

from cattus_b200.selfplay import SelfPlayRunner
from tests.emul import emul
from tests.test_chess_cpu import chess_cb, chess_cfg, chess_fake_net
from tests.test_selfplay_cpu import cb_for, cfg_with, fake_net


def same_games(got, ref):
    assert [r.game_idx for r in got] == [r.game_idx for r in ref]
    for a, b in zip(got, ref):
        assert a.moves == b.moves, (a.game_idx, a.moves, b.moves)
        assert a.winner == b.winner
        assert a.entries == b.entries and a.entry_dirs == b.entry_dirs


def host_games(game, cfg, cb1, cb2, games_num, **kw):
    return SelfPlayRunner(game, cfg).run_with(cb1, cb2, games_num, keep_records=True, **kw)


@pytest.mark.parametrize("game,kind,sim_num", [("hex4", "uniform", 40), ("hex4", "coarse", 40), ("hex5", "hash", 60), ("hex9", "hash", 12),
                                               ("ttt", "uniform", 50), ("ttt", "hash", 30)])
def test_emulated_device_search_equals_host_driver(game, kind, sim_num):
    wpp = 1 if game == "ttt" else (int(game[3:]) ** 2 + 63) // 64
    net = fake_net(kind)
    cfg = cfg_with(sim_num=sim_num, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[3, 1.0], [5, 0.5], [9999, 0.0]], seed=21)
    summary, ref = host_games(game, cfg, cb_for(net, wpp), None, 6)
    counters, got = emul.run(game, cfg, cb_for(net, wpp), None, 6, n_slots=4, pool_words=1 << 16)
    same_games(got, ref)
    m = summary["metrics"]
    assert counters["simulations"] == m["selfplay.simulations"] == counters["searches"] * sim_num
    assert counters["terminal"] == m["selfplay.terminal_leaves"]
    assert counters["w1"] == summary["player1_wins"] and counters["w2"] == summary["player2_wins"] and counters["draws"] == summary["draws"]


def test_results_do_not_depend_on_slots_or_pipeline_depth():
    net = fake_net("hash")
    cfg = cfg_with(sim_num=30, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], seed=8)
    _, ref = host_games("hex5", cfg, cb_for(net, 1), None, 10)
    for n_slots, depth in ((1, 1), (3, 2), (16, 3), (10, 4), (3, 2 | 256), (16, 1 | 256), (4, 2 | (1 << 16)),
                           (4, 2 | 256 | (1000 << 16)), (5, 2 | 256 | 512), (16, 3 | 512), (2, 2 | 256 | 512)):  # | 512: two populations taking waves in turn  # | 256: begin overlapped (a wave later); << 16: node-visit budget per wave (1: one simulation)
        _, got = emul.run("hex5", cfg, cb_for(net, 1), None, 10, n_slots=n_slots, pool_words=1 << 15, depth=depth)
        same_games(got, ref)


@pytest.mark.parametrize("game,cache_size", [("hex5", 100000), ("hex5", 16), ("ttt", 1000)])
def test_device_side_cache_changes_no_game(game, cache_size):
    """ValueFuncCache in HBM: a hit is expanded on the spot inside select.  Same games; fewer evaluator rows."""
    net = fake_net("coarse")
    cfg = cfg_with(sim_num=60, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], seed=4)
    _, ref = host_games(game, cfg, cb_for(net, 1), None, 6)
    plain, got0 = emul.run(game, cfg, cb_for(net, 1), None, 6, n_slots=3, pool_words=1 << 16)
    cached, got = emul.run(game, dict(cfg, mcts=dict(cfg["mcts"], cache_size=cache_size)), cb_for(net, 1), None, 6, n_slots=3, pool_words=1 << 16)
    same_games(got0, ref)
    same_games(got, ref)
    assert plain["cache_hits"] == 0 and cached["cache_hits"] > 0
    assert cached["evaluations"] + cached["cache_hits"] == plain["evaluations"] == cached["simulations"] - cached["terminal"]


def test_device_side_cache_chess_and_two_models():
    net = chess_fake_net("hash")
    cfg = chess_cfg(max_moves=30, sim_num=12, cache_size=5000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[10, 1.0], [9999, 0.0]])
    _, ref = host_games("chess", cfg, chess_cb(net), None, 2)
    counters, got = emul.run("chess", cfg, chess_cb(net), None, 2, n_slots=2, pool_words=1 << 16)
    same_games(got, ref)
    assert counters["cache_hits"] > 0
    n1, n2 = fake_net("hash"), fake_net("hash", salt=99)
    cfg = cfg_with(sim_num=30, temperature_policy=[[9999, 1.0]], seed=3, cache_size=1000)
    _, ref = host_games("hex4", cfg, cb_for(n1, 1), cb_for(n2, 1), 6)
    counters, got = emul.run("hex4", cfg, cb_for(n1, 1), cb_for(n2, 1), 6, n_slots=3, pool_words=1 << 14)
    same_games(got, ref)
    assert counters["cache_hits"] > 0
    counters, got = emul.run("hex4", cfg, cb_for(n1, 1), cb_for(n2, 1), 6, n_slots=4, pool_words=1 << 14, depth=2 | 256 | 512)  # two populations, each its own caches
    same_games(got, ref)
    assert counters["cache_hits"] > 0


def test_no_noise_temperature_zero_and_tree_reuse():
    net = fake_net("coarse")
    cfg = cfg_with(sim_num=80)
    _, ref = host_games("hex5", cfg, cb_for(net, 1), None, 4)
    _, got = emul.run("hex5", cfg, cb_for(net, 1), None, 4, n_slots=2, pool_words=1 << 16)
    same_games(got, ref)


def test_two_evaluators_model_compare():
    n1, n2 = fake_net("hash"), fake_net("hash", salt=99)
    cfg = cfg_with(sim_num=30, temperature_policy=[[9999, 1.0]], seed=3)
    _, ref = host_games("hex4", cfg, cb_for(n1, 1), cb_for(n2, 1), 6)
    _, got = emul.run("hex4", cfg, cb_for(n1, 1), cb_for(n2, 1), 6, n_slots=3, pool_words=1 << 14)
    same_games(got, ref)


def test_partition_by_stride_and_max_moves():
    net = fake_net("hash")
    cfg = cfg_with(sim_num=20, max_moves=7)
    _, ref = host_games("hex5", cfg, cb_for(net, 1), None, 8, first_game=1, game_stride=2)
    _, got = emul.run("hex5", cfg, cb_for(net, 1), None, 8, n_slots=8, pool_words=1 << 14, first_game=1, game_stride=2)
    same_games(got, ref)
    assert all(len(r.moves) <= 7 for r in got)


def test_pool_exhaustion_is_an_error_not_a_wrong_game():
    net = fake_net("hash")
    with pytest.raises(RuntimeError, match="error bits"):
        emul.run("hex5", cfg_with(sim_num=200), cb_for(net, 1), None, 2, n_slots=2, pool_words=2048)


@pytest.mark.parametrize("kind,kw", [("hash", {}), ("coarse", dict(sim_num=10, prior_noise_alpha=0.3, prior_noise_epsilon=0.25,
                                                                   temperature_policy=[[30, 1.0], [9999, 0.0]], seed=77)),
                                     ("uniform", dict(sim_num=8, temperature_policy=[[9999, 1.0]], seed=9))])
def test_emulated_chess_equals_host_driver(kind, kw):
    net = chess_fake_net(kind)
    cfg = chess_cfg(max_moves=40, **kw)
    summary, ref = host_games("chess", cfg, chess_cb(net), None, 2)
    counters, got = emul.run("chess", cfg, chess_cb(net), None, 2, n_slots=2, pool_words=1 << 16)
    same_games(got, ref)
    assert counters["terminal"] == summary["metrics"]["selfplay.terminal_leaves"]


def test_emulated_chess_repetition_inside_the_search():
    # a deterministic net at temperature 0 shuffles pieces: games end by threefold repetition, and detect_repetition fires
    # inside the searches on the way (tests/test_chess_cpu.py asserts that for the host driver)
    net = chess_fake_net("hash")
    cfg = chess_cfg()
    summary, ref = host_games("chess", cfg, chess_cb(net), None, 2)
    counters, got = emul.run("chess", cfg, chess_cb(net), None, 2, n_slots=1, pool_words=1 << 16, depth=1)
    same_games(got, ref)
    assert counters["terminal"] == summary["metrics"]["selfplay.terminal_leaves"] > 0
