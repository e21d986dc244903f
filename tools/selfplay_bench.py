#!/usr/bin/env python
"""Sweeps the self-play driver's threads x games_per_thread on one GPU and prints sims/s (exploration tool; the
judged number comes from bench.py)."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="hex5")
    ap.add_argument("--sim-num", type=int, default=1400)
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--threads", type=int, nargs="+", default=[16])
    ap.add_argument("--gpt", type=int, nargs="+", default=[16])
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--cache", type=int, default=1000000)
    ap.add_argument("--leaf-queue", type=int, default=0)
    ap.add_argument("--groups", type=int, nargs="+", default=[0])
    ap.add_argument("--max-moves", type=int, default=0)
    ap.add_argument("--speculate", type=int, nargs="+", default=[0], help="rows per game evaluated ahead into the cache (small batches only)")
    ap.add_argument("--device-games", type=int, nargs="+", default=[0], help="> 0: device-resident search with this many concurrent games")
    ap.add_argument("--device-cache", action="store_true", help="device search: ValueFuncCache in HBM with --cache entries")
    ap.add_argument("--waves", type=int, default=0, help="device search: waves in flight (0 = default)")
    ap.add_argument("--tree-kwords", type=int, default=0, help="device search: words (x1024) per tree buffer (0 = auto)")
    ap.add_argument("--batch", type=int, default=0, help="evaluator max batch (0 = games per thread, clamped to 64..4096)")
    args = ap.parse_args()

    from cattus_b200 import CudaNetwork
    from cattus_b200.export import export_blob
    from cattus_b200.selfplay import SelfPlayRunner
    from oracle import net

    cfg_net = net.CONFIGS[args.game]
    blob = export_blob(net.make_state_dict(cfg_net, 0), cfg_net.game)
    for dg in [d for d in args.device_games if d > 0]:
        with CudaNetwork(blob, cfg_net.game, batch_size=max(64, dg), n_streams=args.streams if args.streams != 4 else 2) as nw:
            cfg = {"mcts": {"sim_num": args.sim_num, "explore_factor": 1.41421, "temperature_policy": [[10, 1.0], [9999, 0.0]],
                            "prior_noise_alpha": 0.03, "prior_noise_epsilon": 0.25, "cache_size": args.cache if args.cache != 1000000 or args.device_cache else 0},
                   "seed": 1, "max_moves": args.max_moves, "device_games": dg, "device_waves_in_flight": args.waves, "device_tree_kwords": args.tree_kwords}
            games = max(2, (args.games + 1) // 2 * 2)
            summary, _ = SelfPlayRunner("chess" if args.game.startswith("chess") else args.game, cfg).generate_data(nw, None, games)
            m = summary["metrics"]
            print(json.dumps({"device_games": dg, "games": games, "sims_per_sec": round(m["selfplay.sims_per_sec"]), "seconds": round(m["selfplay.seconds"], 3),
                              "simulations": m["selfplay.simulations"], "evals": m["selfplay.evaluations"], "waves": m["model.activation_count"],
                              "mean_batch": round(m["selfplay.evaluations"] / max(1, m["model.activation_count"]), 1), "cache_hits": m["cache.hits"],
                              "ms_per_wave": round(1e3 * m["selfplay.seconds"] / max(1, m["model.activation_count"]), 4),
                              "p1": summary["player1_wins"], "p2": summary["player2_wins"], "draws": summary["draws"]}), flush=True)
    if args.device_games != [0]:
        return
    for th in args.threads:
        for gpt, groups, speculate in [(g, k, sp) for g in args.gpt for k in args.groups for sp in args.speculate]:
            with CudaNetwork(blob, cfg_net.game, batch_size=args.batch or max(256 if speculate else 64, min(4096, gpt)), n_streams=args.streams) as nw:
                cfg = {"mcts": {"sim_num": args.sim_num, "explore_factor": 1.41421, "temperature_policy": [[10, 1.0], [9999, 0.0]],
                                "prior_noise_alpha": 0.03, "prior_noise_epsilon": 0.25, "cache_size": args.cache},
                       "threads": th, "games_per_thread": gpt, "groups_per_thread": groups, "leaf_queue": args.leaf_queue, "seed": 1,
                       "max_moves": args.max_moves, "speculate": speculate}
                games = max(2, (args.games + 1) // 2 * 2)
                summary, _ = SelfPlayRunner("chess" if args.game.startswith("chess") else args.game, cfg).generate_data(nw, None, games)
                m = summary["metrics"]
                print(json.dumps({"threads": th, "gpt": gpt, "groups": groups, "speculate": speculate, "games": games,
                                  "spec_evals": m["selfplay.speculative_evaluations"], "sims_per_sec": round(m["selfplay.sims_per_sec"]),
                                  "seconds": round(m["selfplay.seconds"], 3), "evals": m["selfplay.evaluations"], "batches": m["model.activation_count"],
                                  "mean_batch": round(m["selfplay.evaluations"] / max(1, m["model.activation_count"]), 1),
                                  "hit_rate": round(m["cache.hits"] / max(1, m["cache.hits"] + m["cache.misses"]), 3),
                                  "eval_wait_frac": round(m["selfplay.eval_wait_seconds"] / max(1e-9, m["selfplay.seconds"] * th), 3),
                                  "p1": summary["player1_wins"], "p2": summary["player2_wins"], "draws": summary["draws"]}), flush=True)


if __name__ == "__main__":
    main()
