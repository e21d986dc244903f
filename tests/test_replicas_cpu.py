"""CPU: the N > 1 plumbing (replicas only, no data-path collective) with the gloo backend, world_size 2."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    from cattus_b200 import replicas

    rep = replicas.init_from_env("gloo")
    rep.barrier()
    res = {
        "rank": rep.rank, "world": rep.world,
        "max": rep.max_over_ranks(1.0 + rank),            # slowest rank decides the time
        "sum": rep.sum_over_ranks(100.0 * (rank + 1)),    # every rank's units count
        "agg": rep.aggregate_throughput(1000.0, 0.5 * (rank + 1)),
        "seed": replicas.rank_seed(0xCA7705, rep.rank),
    }
    rep.close()
    Path(out_dir, f"r{rank}.json").write_text(json.dumps(res))


def test_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r = [json.loads((tmp_path / f"r{k}.json").read_text()) for k in range(2)]
    assert [x["rank"] for x in r] == [0, 1] and all(x["world"] == 2 for x in r)
    assert all(x["max"] == 2.0 and x["sum"] == 300.0 for x in r)
    assert all(x["agg"] == pytest.approx(2000.0 / 1.0) for x in r)  # 2 x 1000 units / max(0.5, 1.0) s
    assert r[0]["seed"] != r[1]["seed"]


def test_single_replica_needs_no_process_group(monkeypatch):
    from cattus_b200 import replicas

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    rep = replicas.init_from_env("nccl")
    assert (rep.rank, rep.world, rep.dist) == (0, 1, None)
    assert rep.max_over_ranks(3.0) == 3.0 and rep.aggregate_throughput(10.0, 2.0) == 5.0
    rep.barrier()
    rep.close()


def test_worker_partition_covers_every_worker_once():
    from cattus_b200.replicas import partition_workers

    parts = partition_workers(19, 8)
    assert sorted(w for p in parts for w in p) == list(range(19)) and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert partition_workers(4, 1) == [[0, 1, 2, 3]]


def test_reference_arm_prints_on_rank0_only():
    """bench.py --impl reference under N ranks: rank 0 alone runs and prints, the others exit 0 without work."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--workload", "hex4", "--cpu-sample", "64"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "positions/s" and line["config"]["workload"] == "hex4"
