"""World-size-2 check (gloo, CPU) of the only multi-rank logic on this path: each rank owns a
column, ranks exchange nothing but a barrier, the max of their timings and the sum of their
evaluation counts (bench.py); and the per-device layer sharding of ``Gas``."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r})
    import numpy as np
    import torch.distributed as dist
    import bench
    from pylbl_b200 import synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ranks = bench.Ranks(dist, "cpu")
    col = bench.rank_column(rank)
    ranks.barrier()
    total_t = ranks.sum(float(col.t.sum()))
    slowest = ranks.max(10.0 + rank)
    # the shards of the other two modes: columns round-robin (configs[4]), one band of the grid
    # per rank (configs[3]); every rank contributes its share to the same reductions
    mine = bench.rank_columns(11, world, rank)
    n_cols = ranks.sum(float(len(mine)))
    col_ids = ranks.sum(float(sum(mine)))
    lo, hi = bench.rank_band([0, 1700, 3491], rank)
    cells = ranks.sum(float(hi - lo))
    if rank == 0:
        print(json.dumps(dict(world=world, total_t=total_t, slowest=slowest, n_cols=n_cols,
                              col_ids=col_ids, cells=cells)))
    dist.destroy_process_group()
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_agree_on_totals(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        procs.append(subprocess.Popen([sys.executable, str(script)],
                                      env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)),
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    import json
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    import bench
    want = float(bench.rank_column(0).t.sum() + bench.rank_column(1).t.sum())
    assert line["world"] == 2
    assert abs(line["total_t"] - want) < 1e-9
    assert line["slowest"] == 11.0
    assert line["n_cols"] == 11 and line["col_ids"] == sum(range(11)) and line["cells"] == 3491
    # columns differ between ranks (weak scaling: each rank has its own work)
    assert not np.array_equal(bench.rank_column(0).t, bench.rank_column(1).t)


def test_single_process_ranks_object():
    import bench
    r = bench.Ranks()
    r.barrier()
    assert r.max(3.5) == 3.5 and r.sum(2) == 2.0


def test_layer_shards_cover_all_layers():
    # the split Gas.absorption_coefficients uses for `devices=[...]`
    for n_layers in (1, 2, 7, 60, 61):
        for ndev in (1, 2, 4, 8):
            nd = max(1, min(ndev, n_layers))
            edges = np.linspace(0, n_layers, nd + 1).astype(int)
            sizes = np.diff(edges)
            assert edges[0] == 0 and edges[-1] == n_layers and np.all(sizes >= 1)
            assert sizes.max() - sizes.min() <= 1


def test_no_collective_is_conditional_on_the_rank():
    """Every rank must reach every collective: a reduction or barrier inside `if rank == 0:` (or
    any branch on the rank) leaves the other ranks waiting until the watchdog fires -- a bench at
    N > 1 that hangs for minutes instead of failing.  Static check of bench.py: no call of the
    rank-wide helpers (Ranks.barrier / max / sum and the names bound to them, torch.distributed
    collectives) inside a branch whose condition reads the rank."""
    import ast
    source = (Path(__file__).resolve().parent.parent / "bench.py").read_text()
    tree = ast.parse(source)
    collective = {"barrier", "max_over_ranks", "sum_over_ranks", "all_reduce", "all_gather", "broadcast",
                  "reduce", "gather", "scatter", "all_to_all"}
    rank_wide_methods = {"barrier", "max", "sum", "_reduce"}

    def reads_rank(node):
        return any(isinstance(n, ast.Name) and n.id in ("rank", "local_rank") for n in ast.walk(node)) or \
            any(isinstance(n, ast.Attribute) and n.attr in ("rank", "local_rank") for n in ast.walk(node))

    def is_collective(call):
        f = call.func
        if isinstance(f, ast.Name):
            return f.id in collective
        if isinstance(f, ast.Attribute):
            owner = f.value.id if isinstance(f.value, ast.Name) else getattr(f.value, "attr", "")
            if owner == "ranks" and f.attr in rank_wide_methods:
                return True
            return owner in ("dist", "distributed") and f.attr in collective
        return False

    offenders = []
    for node in ast.walk(tree):
        if isinstance(node, (ast.If, ast.IfExp, ast.While)) and reads_rank(node.test):
            for branch in (node.body, node.orelse):
                for stmt in (branch if isinstance(branch, list) else [branch]):
                    for inner in ast.walk(stmt):
                        if isinstance(inner, ast.Call) and is_collective(inner):
                            offenders.append(inner.lineno)
    assert offenders == [], f"collectives under a condition on the rank at bench.py lines {offenders}"
    # and the check does see the helpers where they are legitimately used
    assert sum(1 for n in ast.walk(tree) if isinstance(n, ast.Call) and is_collective(n)) >= 10
