"""UCI front-end with the B200 evaluator behind the search: the reference's `cattus` binary
(engine/src/bin/cattus.rs, engine/src/chess/uci.rs) with `CudaNetwork<ChessGame>` as the value function instead of
the `StockfishNet` placeholder it is built with today (cattus.rs:57-64).

    python -m cattus_b200.uci --config-file cfg.json

`cfg.json` is the reference's config (cattus.rs:16-42): {"model": {"model_path", "inference", "batch_size"},
"mcts": {"sim_num", "explore_factor", "temperature_policy", "prior_noise_alpha", "prior_noise_epsilon", "cache_size"},
"threads"}; `model.model_path` is a `.cb2` blob (cattus_b200.export) and `model.inference` may carry
{"engine": "cuda-b200", "device": 0, "precision": "bf16"}.

`speculate` (top-level key, default 31; 0 = the reference's behaviour) lets up to that many likely next leaves ride along
with the leaf the search is waiting for; they only land in the value-function cache, so the moves played are unchanged
and a `go` needs about an eighth of the evaluator round trips.

Commands handled as in uci.rs:27-73: uci, isready, setoption, ucinewgame, position [fen <FEN> | startpos] [moves ...],
go (its arguments are parsed and ignored, as there), stop, quit.  One MctsPlayer per game: the tree of the previous
`go` is reused when the new position is in it; every search runs `mcts.sim_num` simulations with one leaf in flight.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path
from typing import List, Optional, TextIO

from .network import CudaNetwork
from .selfplay import ChessSearch

GO_KEYS = ("searchmoves", "ponder", "wtime", "btime", "winc", "binc", "movestogo", "depth", "nodes", "mate", "movetime", "infinite")


def parse_args(args: List[str], keys) -> dict:
    """UCI::parse_args (uci.rs:177-192): words after a key belong to it; a repeated key or a word before any key is an error."""
    out: dict = {}
    key = None
    for a in args:
        if a in keys:
            if a in out:
                raise ValueError(f"arg '{a}' appears multiple times")
            out[a] = []
            key = a
        else:
            if key is None:
                raise ValueError(f"arg '{a}' has no key")
            out[key].append(a)
    return out


class UCI:
    def __init__(self, cfg: dict, model=None, eval_fn=None, out: TextIO = sys.stdout):
        cfg = dict(cfg)
        cfg.setdefault("speculate", 31 if (cfg.get("mcts") or {}).get("cache_size") else 0)
        self.cfg = cfg
        self.model = model
        self.eval_fn = eval_fn
        self.out = out
        self.options: dict = {}
        self.player: Optional[ChessSearch] = None
        self.fen: Optional[str] = None
        self.moves: Optional[List[str]] = None
        self.last_stats: Optional[dict] = None

    def send(self, s: str) -> None:
        print(s, file=self.out, flush=True)

    def handle(self, line: str) -> bool:
        """One command; False after `quit`."""
        words = line.split()
        if not words:
            return True
        command, args = words[0], words[1:]
        if command == "uci":
            self.send("id name cattus_b200 v1.0.0")
            self.send("id author cattus_b200")
            self.send("uciok")
        elif command == "isready":
            self.send("readyok")
        elif command == "setoption":
            a = parse_args(args, ("name", "value"))
            self.options[" ".join(a["name"])] = " ".join(a["value"])
        elif command == "ucinewgame":
            if self.player is not None:
                self.player.close()
            self.player = ChessSearch(self.cfg, self.model, self.eval_fn)
        elif command == "position":
            a = parse_args(args, ("fen", "startpos", "moves"))
            if ("fen" in a) == ("startpos" in a):
                raise ValueError("position cmd requires either fen or startpos")
            self.fen = " ".join(a["fen"]) if "fen" in a else None
            self.moves = list(a.get("moves", []))
        elif command == "go":
            parse_args(args, GO_KEYS)
            if self.player is None:  # the reference unwraps None here; a GUI that skips ucinewgame still gets a player
                self.player = ChessSearch(self.cfg, self.model, self.eval_fn)
            if self.moves is None:
                raise ValueError("go before position")
            best, stats = self.player.go(self.fen, self.moves)
            self.last_stats = stats
            self.send(f"info nodes {stats['simulations']} time {int(stats['seconds'] * 1000)} nps {int(stats['simulations'] / max(stats['seconds'], 1e-9))}")
            self.send(f"bestmove {best}")
        elif command in ("stop",):
            pass
        elif command in ("ponderhit", "start", "fen", "xyzzy"):
            self.send("uciok")
        elif command == "quit":
            return False
        else:
            print(f"unknown command {command}", file=sys.stderr)
        return True

    def run(self, inp: TextIO = sys.stdin) -> None:
        for line in inp:
            try:
                if not self.handle(line.strip()):
                    break
            except Exception as e:  # a malformed command must not take the engine down mid-game
                print(f"error: {e}", file=sys.stderr)
        if self.player is not None:
            self.player.close()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="cattus_b200.uci", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-file", required=True, type=Path)
    args = ap.parse_args(argv)
    cfg = json.loads(args.config_file.read_text())
    model_cfg = cfg["model"]
    inference = model_cfg.get("inference") or {}
    if inference.get("engine", "cuda-b200") != "cuda-b200":
        raise SystemExit(f"this executable is the cuda-b200 engine; config.model.inference.engine is {inference.get('engine')!r} (there is no CPU fallback)")
    # room for the speculative rows: a device batch of up to 64 positions costs what one position costs
    with CudaNetwork(Path(model_cfg["model_path"]), "chess", device=int(inference.get("device", 0)), batch_size=max(64, int(model_cfg.get("batch_size", 1))),
                     n_streams=1, precision=inference.get("precision", "bf16")) as nw:
        UCI(cfg, model=nw).run()
    return 0


if __name__ == "__main__":
    sys.exit(main())
