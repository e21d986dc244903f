"""GPU (-m gpu): the DEVICE-RESIDENT search (cfg["device_games"] > 0: trees in HBM, one warp per game, leaves written
straight into the evaluator's device batch) must play exactly the games the host-tree driver plays on the same
evaluator: every move, every winner, every .traindata byte -- for any number of concurrent device games and any number of
waves in flight.  The host driver is pinned to oracle/mcts.py by tests/test_gpu_selfplay.py; one test here also replays
the oracle directly.
"""
import numpy as np
import pytest

from cattus_b200.selfplay import SelfPlayRunner
from oracle import mcts as om
from tests.test_selfplay_cpu import cfg_with, oracle_games
from tests.util import make_network

pytestmark = pytest.mark.gpu


def games_of(records):
    return [(r.game_idx, r.moves, r.winner, r.entries, r.entry_dirs) for r in records]


@pytest.mark.parametrize("name,sim_num,games_num", [("hex4", 60, 8), ("hex5", 50, 6), ("hex7", 40, 4), ("hex9", 16, 4), ("hex11", 12, 2), ("ttt", 40, 8)])
def test_device_games_equal_host_driver_games(name, sim_num, games_num):
    base = dict(sim_num=sim_num, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [6, 0.5], [9999, 0.0]], seed=77)
    with make_network(name, batch_size=64) as nw:
        s_host, host = SelfPlayRunner(name, cfg_with(cache_size=10000, threads=2, games_per_thread=3, **base)).generate_data(nw, None, games_num, keep_records=True)
        s_dev, dev = SelfPlayRunner(name, cfg_with(device_games=5, **base)).generate_data(nw, None, games_num, keep_records=True)
    assert games_of(dev) == games_of(host)
    mh, md = s_host["metrics"], s_dev["metrics"]
    assert md["selfplay.simulations"] == mh["selfplay.simulations"] == md["selfplay.searches"] * sim_num
    assert md["selfplay.terminal_leaves"] == mh["selfplay.terminal_leaves"]
    assert md["selfplay.evaluations"] == md["selfplay.simulations"] - md["selfplay.terminal_leaves"]  # no cache on the device
    for k in ("player1_wins", "player2_wins", "draws"):
        assert s_dev[k] == s_host[k]


@pytest.mark.parametrize("name,game,kw", [("hex5", "hex5", dict(sim_num=120)), ("hex7", "hex7", dict(sim_num=60)),
                                          ("chess_dev", "chess", dict(sim_num=24, max_moves=24)), ("ttt", "ttt", dict(sim_num=40))])
def test_device_side_cache_same_games_fewer_rows(name, game, kw):
    """cfg cache_size > 0 puts ValueFuncCache in HBM: hits are expanded inside the select kernel.  Same games."""
    base = dict(prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=13)
    base.update(kw)
    with make_network(name, batch_size=64) as nw:
        _, host = SelfPlayRunner(game, cfg_with(cache_size=10000, threads=2, games_per_thread=4, **base)).generate_data(nw, None, 8, keep_records=True)
        s0, plain = SelfPlayRunner(game, cfg_with(device_games=8, **base)).generate_data(nw, None, 8, keep_records=True)
        s1, cached = SelfPlayRunner(game, cfg_with(device_games=8, cache_size=50000, **base)).generate_data(nw, None, 8, keep_records=True)
        s2, tiny = SelfPlayRunner(game, cfg_with(device_games=8, cache_size=8, **base)).generate_data(nw, None, 8, keep_records=True)
    assert games_of(plain) == games_of(host) and games_of(cached) == games_of(host) and games_of(tiny) == games_of(host)
    m0, m1 = s0["metrics"], s1["metrics"]
    assert m0["cache.hits"] == 0 and m1["cache.hits"] > 0
    assert m1["selfplay.evaluations"] + m1["cache.hits"] == m0["selfplay.evaluations"]


def test_cache_on_and_off_same_games_at_scale():
    """Thousands of concurrent games hammer the HBM cache's buckets (concurrent inserts take the bucket lock; probes run in
    another kernel): with and without the cache, and from run to run, every game must be the same game."""
    base = dict(sim_num=100, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=31)
    with make_network("hex5", batch_size=4096, n_streams=2) as nw:
        s0, plain = SelfPlayRunner("hex5", cfg_with(device_games=4096, **base)).generate_data(nw, None, 4096, keep_records=True)
        s1, cached = SelfPlayRunner("hex5", cfg_with(device_games=4096, cache_size=4096, **base)).generate_data(nw, None, 4096, keep_records=True)
        s2, again = SelfPlayRunner("hex5", cfg_with(device_games=2048, cache_size=200000, **base)).generate_data(nw, None, 4096, keep_records=True)
    key = lambda recs: [(r.game_idx, r.moves, r.winner) for r in recs]  # noqa: E731
    assert key(cached) == key(plain) and key(again) == key(plain)
    assert s1["metrics"]["cache.hits"] > 0 and s1["metrics"]["selfplay.simulations"] == s0["metrics"]["selfplay.simulations"]


def test_device_games_equal_the_oracle_directly():
    from tests.test_gpu_selfplay import gpu_net_fn

    cfg = cfg_with(sim_num=40, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[4, 1.0], [9999, 0.0]], seed=5, device_games=4)
    with make_network("hex5", batch_size=64) as nw:
        _, dev = SelfPlayRunner("hex5", cfg).generate_data(nw, None, 4, keep_records=True)
        ref, _ = oracle_games("hex5", cfg, gpu_net_fn(nw, 5), None, range(4))
    for rec, o in zip(dev, ref):
        assert rec.moves == o.moves and rec.winner == o.winner
        for k, (pos, probs) in enumerate(o.entries):
            assert rec.entries[k] == om.data_entry_bytes(pos, probs, o.winner)


def test_results_do_not_depend_on_device_games_or_waves_in_flight():
    base = dict(sim_num=80, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, temperature_policy=[[10, 1.0], [9999, 0.0]], seed=3)
    results = []
    with make_network("hex5", batch_size=256, n_streams=2) as nw:
        for dg, waves in ((1, 1), (7, 2), (64, 3), (256, 4)):
            s, recs = SelfPlayRunner("hex5", cfg_with(device_games=dg, device_waves_in_flight=waves, **base)).generate_data(nw, None, 64, keep_records=True)
            results.append(games_of(recs))
            assert s["player1_wins"] + s["player2_wins"] == 64
    for r in results[1:]:
        assert r == results[0]


def test_two_populations_experiment_knob_same_games(monkeypatch):
    """CATTUS_B200_DSEARCH_TWO_POPULATIONS: two halves of the slots take waves in turn on two streams (measured: no gain)."""
    base = dict(sim_num=60, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=21, cache_size=20000)
    with make_network("hex5", batch_size=256, n_streams=2) as nw:
        _, one = SelfPlayRunner("hex5", cfg_with(device_games=128, **base)).generate_data(nw, None, 160, keep_records=True)
        monkeypatch.setenv("CATTUS_B200_DSEARCH_TWO_POPULATIONS", "1")
        _, two = SelfPlayRunner("hex5", cfg_with(device_games=128, **base)).generate_data(nw, None, 160, keep_records=True)
    assert games_of(two) == games_of(one)


def test_device_search_many_games_long_searches():
    """Tree reuse over whole games with hundreds of simulations per move and hundreds of games in flight."""
    base = dict(sim_num=300, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=11)
    with make_network("hex5", batch_size=512, n_streams=4) as nw:
        _, host = SelfPlayRunner("hex5", cfg_with(threads=8, games_per_thread=32, cache_size=100000, **base)).generate_data(nw, None, 256, keep_records=True)
        s, dev = SelfPlayRunner("hex5", cfg_with(device_games=512, **base)).generate_data(nw, None, 256, keep_records=True)
    assert [(r.game_idx, r.moves, r.winner) for r in dev] == [(r.game_idx, r.moves, r.winner) for r in host]
    assert s["metrics"]["model.activation_count"] > 0


def test_device_search_two_models_and_files(tmp_path):
    from cattus_b200 import CudaNetwork
    from tests.util import blob

    base = dict(sim_num=40, temperature_policy=[[9999, 1.0]], seed=2)
    with make_network("hex4") as nw1, CudaNetwork(blob("hex4", 1), "hex", batch_size=64) as nw2:
        _, host = SelfPlayRunner("hex4", cfg_with(cache_size=1000, threads=2, games_per_thread=4, **base)).generate_data(nw1, nw2, 8, keep_records=True)
        s, dev = SelfPlayRunner("hex4", cfg_with(device_games=8, **base)).generate_data(nw1, nw2, 8, tmp_path / "d1", tmp_path / "d2", keep_records=True)
    assert games_of(dev) == games_of(host)
    files = list((tmp_path / "d1").glob("*.traindata")) + list((tmp_path / "d2").glob("*.traindata"))
    assert len(files) == sum(len(r.entries) for r in dev)
    for r in dev:
        for k, (e, d) in enumerate(zip(r.entries, r.entry_dirs)):
            assert (tmp_path / f"d{d}" / f"{r.game_idx:08d}_{k:03d}.traindata").read_bytes() == e


@pytest.mark.parametrize("name,kw", [("chess_dev", dict(sim_num=24, max_moves=30)), ("chess_dev", dict(sim_num=12, max_moves=0, temperature_policy=[[9999, 0.0]])),
                                     ("chess10x128", dict(sim_num=16, max_moves=12))])
def test_device_chess_games_equal_host_driver_games(name, kw):
    base = dict(prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[20, 1.0], [9999, 0.0]], seed=9)
    base.update(kw)
    with make_network(name, batch_size=64) as nw:
        s_host, host = SelfPlayRunner("chess", cfg_with(cache_size=50000, threads=2, games_per_thread=2, **base)).generate_data(nw, None, 4, keep_records=True)
        s_dev, dev = SelfPlayRunner("chess", cfg_with(device_games=4, **base)).generate_data(nw, None, 4, keep_records=True)
    assert games_of(dev) == games_of(host)
    assert s_dev["metrics"]["selfplay.terminal_leaves"] == s_host["metrics"]["selfplay.terminal_leaves"]
    assert s_dev["metrics"]["selfplay.simulations"] == s_host["metrics"]["selfplay.simulations"]


def test_device_chess_whole_games_many_slots():
    """64 chess games played to their end (mate, stalemate, fifty moves or threefold repetition) with the cache on: tree reuse,
    repetition detection inside the search and the 128-deep history ring over hundreds of plies."""
    base = dict(sim_num=32, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[30, 1.0], [9999, 0.0]], seed=5, max_moves=0)
    with make_network("chess_dev", batch_size=64) as nw:
        s_host, host = SelfPlayRunner("chess", cfg_with(cache_size=200000, threads=4, games_per_thread=16, **base)).generate_data(nw, None, 64, keep_records=True)
        s_dev, dev = SelfPlayRunner("chess", cfg_with(device_games=64, cache_size=100000, **base)).generate_data(nw, None, 64, keep_records=True)
    assert games_of(dev) == games_of(host)
    assert max(len(r.moves) for r in dev) > 100
    assert s_dev["metrics"]["selfplay.terminal_leaves"] == s_host["metrics"]["selfplay.terminal_leaves"]


def test_device_games_beyond_max_batch_is_an_error():
    from cattus_b200.selfplay import SelfPlayError

    with make_network("hex4", batch_size=16) as nw:
        with pytest.raises(SelfPlayError, match="max_batch"):
            SelfPlayRunner("hex4", cfg_with(device_games=64)).generate_data(nw, None, 128)
