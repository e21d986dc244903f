"""GPU (-m gpu): self-play through the C ABI with the B200 evaluator.

North-star contract: "given identical net outputs, MCTS move choices must be identical".  The C++ driver plays whole
games with batched leaf evaluation on the GPU; the oracle's MctsPlayer (oracle/mcts.py, restating
engine/src/mcts/mod.rs) replays the same games fed with the SAME network's outputs taken one leaf at a time through
`cattus_b200_eval` -- the evaluator is batch invariant, so every move, visit distribution and .traindata byte must
be identical.
"""
import numpy as np
import pytest

from cattus_b200.selfplay import SelfPlayRunner
from oracle import games, mcts as om
from tests.test_selfplay_cpu import cfg_with, oracle_games
from tests.util import make_network

pytestmark = pytest.mark.gpu


def gpu_net_fn(nw, s: int):
    wpp = (s * s + 63) // 64

    def net(a: int, b: int, ones: int):
        words = games.pack_planes([[a, b, ones]], s).reshape(-1)
        assert len(words) == 3 * wpp
        probs, value = nw.eval_planes(words)
        return np.asarray(probs, dtype=np.float32), np.float32(value)

    return net


@pytest.mark.parametrize("name,sim_num,games_num", [("hex4", 60, 4), ("hex5", 50, 2), ("hex9", 16, 2), ("ttt", 40, 4)])
def test_gpu_selfplay_move_choices_identical_to_oracle(name, sim_num, games_num):
    cfg = cfg_with(sim_num=sim_num, cache_size=10000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25,
                   temperature_policy=[[4, 1.0], [9999, 0.0]], threads=2, games_per_thread=3, seed=77)
    s = 3 if name == "ttt" else int(name[3:])
    with make_network(name, batch_size=64) as nw:
        summary, records = SelfPlayRunner(name, cfg).generate_data(nw, None, games_num, keep_records=True)
        ref, _ = oracle_games(name, cfg, gpu_net_fn(nw, s), None, range(games_num))
    for rec, o in zip(records, ref):
        assert rec.moves == o.moves and rec.winner == o.winner
        for k, (pos, probs) in enumerate(o.entries):
            assert rec.entries[k] == om.data_entry_bytes(pos, probs, o.winner)
    m = summary["metrics"]
    assert m["selfplay.simulations"] == m["selfplay.searches"] * sim_num
    assert m["selfplay.evaluations"] > 0 and m["model.activation_count"] > 0


def test_gpu_selfplay_is_independent_of_scheduling():
    base = dict(sim_num=80, cache_size=50000, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, temperature_policy=[[10, 1.0], [9999, 0.0]], seed=3)
    results = []
    with make_network("hex5", batch_size=256, n_streams=3) as nw:
        for threads, gpt, leaf_queue in ((1, 1, 0), (4, 16, 0), (2, 64, 0), (8, 1, 1)):
            cfg = cfg_with(threads=threads, games_per_thread=gpt, leaf_queue=leaf_queue, **base)
            summary, recs = SelfPlayRunner("hex5", cfg).generate_data(nw, None, 16, keep_records=True)
            results.append([(r.game_idx, r.moves, r.winner, r.entries) for r in recs])
            assert summary["player1_wins"] + summary["player2_wins"] == 16  # hex has no draws
    for r in results[1:]:
        assert r == results[0]


def test_gpu_selfplay_many_threads_many_lanes_same_games():
    """Stress version of the above: more worker threads than evaluator lanes, fresh handles (so graphs and launch
    sequences are built concurrently by several threads), hundreds of games -- every game must still be identical."""
    base = dict(sim_num=40, cache_size=20000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=11)
    results = []
    for threads, gpt, streams in ((32, 8, 8), (1, 512, 1), (24, 32, 4)):
        with make_network("hex5", batch_size=512, n_streams=streams) as nw:
            cfg = cfg_with(threads=threads, games_per_thread=gpt, **base)
            _, recs = SelfPlayRunner("hex5", cfg).generate_data(nw, None, 512, keep_records=True)
            results.append([(r.game_idx, r.moves, r.winner) for r in recs])
    assert results[1] == results[0] and results[2] == results[0]


def test_gpu_selfplay_two_models_and_files(tmp_path):
    cfg = cfg_with(sim_num=40, cache_size=1000, threads=2, games_per_thread=4)
    from cattus_b200 import CudaNetwork
    from tests.util import blob

    with make_network("hex4") as nw1, CudaNetwork(blob("hex4", 1), "hex", batch_size=64) as nw2:
        summary, recs = SelfPlayRunner("hex4", cfg).generate_data(nw1, nw2, 8, tmp_path / "d1", tmp_path / "d2", keep_records=True)
    assert summary["player1_wins"] + summary["player2_wins"] == 8
    files = list((tmp_path / "d1").glob("*.traindata")) + list((tmp_path / "d2").glob("*.traindata"))
    assert len(files) == sum(len(r.entries) for r in recs)
    for r in recs:
        for k, (e, d) in enumerate(zip(r.entries, r.entry_dirs)):
            assert (tmp_path / f"d{d}" / f"{r.game_idx:08d}_{k:03d}.traindata").read_bytes() == e


def test_self_player_cli_is_a_drop_in_for_the_trainer(tmp_path):
    """The command line and config file the trainer uses (train_process.py:151-170), the summary keys it reads
    (:174-186) and the data-entry files its DataSet globs (data_set.py:48)."""
    import json

    from cattus_b200 import self_player
    from tests.util import blob

    model = tmp_path / "model.cb2"
    model.write_bytes(blob("hex4"))
    cfg = {"mcts": {"sim_num": 30, "explore_factor": 1.41421, "temperature_policy": [[4, 1.0], [9999, 0.0]], "prior_noise_alpha": 0.03,
                    "prior_noise_epsilon": 0.25, "cache_size": 1000},
           "model": {"batch_size": 4, "inference": {"engine": "cuda-b200"}}, "threads": 2}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    out = tmp_path / "games" / "run0"
    summary_file = tmp_path / "summary.json"
    rc = self_player.main([f"--model1-path={model}", f"--model2-path={model}", "--games-num=6", f"--out-dir1={out}", f"--out-dir2={out}",
                           f"--summary-file={summary_file}", f"--config-file={tmp_path / 'config.json'}"])
    assert rc == 0
    s = json.loads(summary_file.read_text())
    assert s["player1_wins"] + s["player2_wins"] + s["draws"] == 6
    for key in ("model.activation_count", "model.run_duration", "mcts.search_duration", "cache.hits", "cache.misses"):
        assert key in s["metrics"]
    assert s["metrics"]["model.activation_count"] > 0 and s["metrics"]["cache.hits"] + s["metrics"]["cache.misses"] > 0
    files = sorted(out.rglob("*.traindata"))
    assert len(files) == s["metrics"]["selfplay.searches"] and files[0].name == "00000000_000.traindata"
    assert all(f.stat().st_size == 6 * 8 + 16 * 4 + 1 for f in files)
    # the same job with "device_games" in the config file: trees in HBM, the very same files
    cfg["device_games"] = 4
    (tmp_path / "config2.json").write_text(json.dumps(cfg))
    out2 = tmp_path / "games" / "run1"
    rc = self_player.main([f"--model1-path={model}", f"--model2-path={model}", "--games-num=6", f"--out-dir1={out2}", f"--out-dir2={out2}",
                           f"--summary-file={tmp_path / 'summary2.json'}", f"--config-file={tmp_path / 'config2.json'}"])
    assert rc == 0
    files2 = sorted(out2.rglob("*.traindata"))
    assert [f.name for f in files2] == [f.name for f in files] and all(a.read_bytes() == b.read_bytes() for a, b in zip(files, files2))
    s2 = json.loads((tmp_path / "summary2.json").read_text())
    assert (s2["player1_wins"], s2["player2_wins"], s2["draws"]) == (s["player1_wins"], s["player2_wins"], s["draws"])


def test_gpu_selfplay_groups_with_batches_in_flight_same_games():
    """groups_per_thread > 1: several batches of one worker in flight through eval_batch_submit / _wait, including more
    groups than evaluator streams (the worker then takes its own oldest batch back first) -- same games as one group."""
    base = dict(sim_num=40, cache_size=20000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=11)
    results = []
    for threads, gpt, groups, streams in ((2, 64, 1, 2), (2, 64, 4, 8), (4, 32, 8, 2), (1, 128, 16, 3)):
        with make_network("hex5", batch_size=128, n_streams=streams) as nw:
            cfg = cfg_with(threads=threads, games_per_thread=gpt, groups_per_thread=groups, **base)
            summary, recs = SelfPlayRunner("hex5", cfg).generate_data(nw, None, 128, keep_records=True)
            results.append([(r.game_idx, r.moves, r.winner) for r in recs])
            assert summary["player1_wins"] + summary["player2_wins"] == 128
    for r in results[1:]:
        assert r == results[0]


def test_eval_batch_submit_wait_matches_eval_batch():
    import ctypes as C

    from cattus_b200 import _lib
    from tests.util import synth_inputs

    words, _, legal = synth_inputs("hex7", 100, 5)
    with make_network("hex7", batch_size=64, n_streams=2) as nw:
        ref_p, ref_o, ref_v = nw.eval_batch(words[:64])
        lib = _lib.load()
        t1, t2, t3 = C.c_int32(-5), C.c_int32(-5), C.c_int32(-5)
        w = np.ascontiguousarray(words, dtype=np.uint64)
        ptr = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
        _lib.check(lib.cattus_b200_eval_batch_submit(nw._h, ptr(w[:64], _lib._u64p), None, 64, 0, C.byref(t1)))
        _lib.check(lib.cattus_b200_eval_batch_submit(nw._h, ptr(w[64:], _lib._u64p), None, 36, 0, C.byref(t2)))
        _lib.check(lib.cattus_b200_eval_batch_submit(nw._h, ptr(w[:8], _lib._u64p), None, 8, 0, C.byref(t3)))
        assert t1.value >= 0 and t2.value >= 0 and t1.value != t2.value and t3.value == -1  # both streams busy, non-blocking
        probs = np.empty(64 * 49, np.float32)
        offs = np.empty(65, np.uint32)
        vals = np.empty(64, np.float32)
        _lib.check(lib.cattus_b200_eval_batch_wait(nw._h, t1.value, ptr(probs, _lib._f32p), probs.size, ptr(offs, _lib._u32p), ptr(vals, _lib._f32p)))
        assert np.array_equal(offs, ref_o) and np.array_equal(probs[: offs[64]], ref_p) and np.array_equal(vals, ref_v)
        _lib.check(lib.cattus_b200_eval_batch_wait(nw._h, t2.value, ptr(probs, _lib._f32p), probs.size, ptr(offs, _lib._u32p), ptr(vals, _lib._f32p)))
        assert offs[36] == sum(len(l) for l in legal[64:])
        assert lib.cattus_b200_eval_batch_wait(nw._h, t2.value, ptr(probs, _lib._f32p), probs.size, ptr(offs, _lib._u32p), ptr(vals, _lib._f32p)) == _lib.EINVAL


def chess_gpu_net(nw):
    """net(planes[18], legal nn indices ascending) through one blocking cattus_b200_eval per position."""

    def net(planes, legal):
        probs, value = nw.eval_planes(np.array(planes, dtype=np.uint64), games.bitmap_from_legal(legal, games.CHESS_MOVES_NUM))
        return np.asarray(probs, dtype=np.float32), np.float32(value)

    return net


@pytest.mark.parametrize("name,sim_num,max_moves", [("chess_dev", 24, 0), ("chess_2x128", 12, 60)])
def test_gpu_chess_selfplay_move_choices_identical_to_oracle(name, sim_num, max_moves):
    """Chess: 18 planes + the 235-byte legal bitmap per leaf, probabilities mapped from nn-index order back to the move
    generator's order, threefold repetition inside the searches -- whole games against oracle/mcts.py + oracle/chess.py."""
    from oracle import chess as oc
    from tests.test_chess_cpu import _params, chess_cfg, chess_oracle_fn

    cfg = chess_cfg(sim_num=sim_num, cache_size=20000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[12, 1.0], [9999, 0.0]],
                    threads=2, games_per_thread=2, seed=21, max_moves=max_moves)
    with make_network(name, batch_size=64) as nw:
        summary, records = SelfPlayRunner("chess", cfg).generate_data(nw, None, 4, keep_records=True)
        ev = om.Evaluator(chess_oracle_fn(chess_gpu_net(nw)), om.ValueFuncCache(20000))
        params = _params(cfg)
        ref = [om.play_game(g, oc.ChessPosition.new, params, params, ev, ev, cfg["seed"], max_moves=max_moves) for g in range(4)]
    for rec, o in zip(records, ref):
        assert rec.moves == [oc.move_to_u16(m) for m in o.moves] and rec.winner == o.winner
        for k, (pos, probs) in enumerate(o.entries):
            assert rec.entries[k] == om.data_entry_bytes(pos, probs, o.winner)
    m = summary["metrics"]
    assert m["selfplay.simulations"] == m["selfplay.searches"] * sim_num == sum(o.sims for o in ref)
    assert m["selfplay.terminal_leaves"] == sum(o.terminal_leaves for o in ref)
    assert m["selfplay.evaluations"] > 0 and m["model.activation_count"] > 0


def test_gpu_chess_selfplay_is_independent_of_scheduling():
    from tests.test_chess_cpu import chess_cfg

    base = dict(sim_num=40, cache_size=50000, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, temperature_policy=[[30, 1.0], [9999, 0.0]], seed=3,
                max_moves=24)
    results = []
    with make_network("chess_dev", batch_size=256, n_streams=4) as nw:
        for threads, gpt, groups in ((1, 1, 0), (4, 16, 0), (2, 32, 2)):
            cfg = chess_cfg(threads=threads, games_per_thread=gpt, groups_per_thread=groups, **base)
            _, recs = SelfPlayRunner("chess", cfg).generate_data(nw, None, 64, keep_records=True)
            results.append([(r.game_idx, r.moves, r.winner, r.entries) for r in recs])
    assert results[1] == results[0] and results[2] == results[0]


def test_gpu_uci_search_matches_oracle_player():
    """`position` + `go` through cattus_b200_chess_search_* with the B200 evaluator behind it (per-leaf cattus_b200_eval),
    against the oracle's MctsPlayer fed the same network's outputs; the second `go` reuses the first one's tree."""
    import io

    from cattus_b200.selfplay import ChessSearch
    from cattus_b200.uci import UCI
    from oracle import chess as oc
    from tests.test_chess_cpu import KIWIPETE, _params, chess_cfg, chess_oracle_fn

    cfg = chess_cfg(sim_num=200, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, cache_size=100000, seed=4)
    with make_network("chess_2x128", batch_size=8, n_streams=1) as nw:
        player = om.MctsPlayer(_params(cfg), om.Evaluator(chess_oracle_fn(chess_gpu_net(nw)), om.ValueFuncCache(100000)), om.SplitMix64(om.game_seed(4, 0)))
        history = [oc.ChessPosition.from_fen(KIWIPETE)]
        moves = []
        with ChessSearch(cfg, model=nw) as search:
            for _ in range(3):
                best, stats = search.go(KIWIPETE, moves)
                want = player.choose_move_from_probabilities(history, player.calc_moves_probabilities(history))
                assert best == oc.move_to_lan(want) and stats["simulations"] == 200 and stats["evaluations"] > 0
                history.append(history[-1].moved_position(want))
                reply = history[-1].legal_moves()[0]
                history.append(history[-1].moved_position(reply))
                moves += [best, oc.move_to_lan(reply)]
        out = io.StringIO()
        UCI(cfg, model=nw, out=out).run(io.StringIO("uci\nucinewgame\nposition startpos\ngo\nquit\n"))
        assert any(ln.startswith("bestmove ") for ln in out.getvalue().split("\n"))


def test_gpu_uci_speculation_same_moves_fewer_round_trips():
    """cfg.speculate on the real evaluator: likely next leaves ride along with the waiting leaf into the cache.  The
    evaluator is batch invariant, so every `go` must return the move of the plain search."""
    from cattus_b200.selfplay import ChessSearch
    from tests.test_chess_cpu import KIWIPETE, chess_cfg

    base = dict(sim_num=600, prior_noise_alpha=0.03, prior_noise_epsilon=0.25, cache_size=100000, seed=6)
    results = {}
    for speculate in (0, 7):
        with make_network("chess_2x128", batch_size=8, n_streams=1) as nw:
            with ChessSearch(chess_cfg(speculate=speculate, **base), model=nw) as search:
                moves, out = [], []
                for _ in range(3):
                    best, stats = search.go(KIWIPETE, moves)
                    out.append((best, stats["simulations"], stats["terminal_leaves"]))
                    moves = moves + [best]
                results[speculate] = (out, stats, nw.metrics()["model.activation_count"])
    assert results[0][0] == results[7][0]
    assert results[7][1]["speculative_evaluations"] > 0 and results[7][2] < 0.8 * results[0][2]


def test_self_player_cli_chess(tmp_path):
    """The self-play executable's command line for chess (training/self-play/src/self_play_cmd.rs:14-31 with
    ChessSerializer, serialize/chess.rs:18-57): 1280-byte entries that the trainer's parser layout accepts
    (cattus_train/chess.py:22-49: 18 planes | 235-byte bitmap | 225 probabilities | winner)."""
    import json
    import struct

    from cattus_b200 import self_player
    from tests.util import blob

    model = tmp_path / "model.cb2"
    model.write_bytes(blob("chess_dev"))
    cfg = {"mcts": {"sim_num": 20, "explore_factor": 1.41421, "temperature_policy": [[30, 1.0], [9999, 0.0]], "prior_noise_alpha": 0.03,
                    "prior_noise_epsilon": 0.25, "cache_size": 1000},
           "model": {"batch_size": 1, "inference": {"engine": "cuda-b200"}}, "threads": 2, "max_moves": 6}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    out = tmp_path / "games"
    summary_file = tmp_path / "summary.json"
    assert self_player.main([f"--model1-path={model}", f"--model2-path={model}", "--games-num=4", f"--out-dir1={out}", f"--out-dir2={out}",
                             f"--summary-file={summary_file}", f"--config-file={tmp_path / 'config.json'}"]) == 0
    s = json.loads(summary_file.read_text())
    assert s["player1_wins"] + s["player2_wins"] + s["draws"] == 4
    files = sorted(out.rglob("*.traindata"))
    assert len(files) == s["metrics"]["selfplay.searches"] == 24
    for f in files:
        e = f.read_bytes()
        assert len(e) == 18 * 8 + 235 + 225 * 4 + 1
        planes = struct.unpack("<18Q", e[:144])
        bitmap = np.frombuffer(e[144:379], dtype=np.uint8)
        probs = np.frombuffer(e[379:1279], dtype="<f4")
        n_legal = int(np.unpackbits(bitmap).sum())
        assert planes[17] == 0xFFFFFFFFFFFFFFFF and bin(planes[5]).count("1") == 1 and bin(planes[11]).count("1") == 1  # one king each
        assert (probs[:n_legal] >= 0).all() and (probs[n_legal:] == -1).all() and abs(float(probs[:n_legal].sum()) - 1.0) < 1e-4
        assert struct.unpack("<b", e[1279:])[0] in (-1, 0, 1)


@pytest.mark.parametrize("name,game,kw", [("hex5", "hex5", dict(sim_num=200)), ("chess_dev", "chess", dict(sim_num=60, max_moves=10))])
def test_gpu_selfplay_speculation_same_games_fewer_round_trips(name, game, kw):
    """cfg.speculate in the self-play driver on the real evaluator (async batches, several lanes): the trainer-sized
    arrangement -- a few games per worker -- with likely next leaves riding along into the cache.  Games must be identical."""
    base = dict(cache_size=100000, prior_noise_alpha=0.3, prior_noise_epsilon=0.25, temperature_policy=[[6, 1.0], [9999, 0.0]], seed=5,
                threads=4, games_per_thread=4, **kw)
    out = {}
    for speculate in (0, 14):
        with make_network(name, batch_size=256, n_streams=4) as nw:
            summary, recs = SelfPlayRunner(game, cfg_with(speculate=speculate, **base)).generate_data(nw, None, 16, keep_records=True)
            out[speculate] = ([(r.game_idx, r.moves, r.winner, r.entries) for r in recs], summary["metrics"])
    assert out[0][0] == out[14][0]
    assert out[14][1]["selfplay.speculative_evaluations"] > 0 and out[0][1]["selfplay.speculative_evaluations"] == 0
    assert out[14][1]["model.activation_count"] < 0.8 * out[0][1]["model.activation_count"]
