"""CPU, world_size 2, gloo: the multi-GPU host logic of bench.py - ARFCN sharding (each rank owns
its own ARFCNs, distinct seeds, no data-path collective) and the max-over-ranks time reduction."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharding_and_time_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        # per-rank workload parameters: distinct payloads / channel parameters per rank (weak scaling)
        p = bench.burst_params(64, "bcch", 1000 + rank + bench.BT["bcch"])
        digest = torch.tensor([float(p["l2"].astype(np.int64).sum()), float(p["toa"].sum())], dtype=torch.float64)
        allg = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allg, digest)
        t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)      # fake per-rank elapsed ms
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({{"distinct": bool((allg[0] != allg[1]).any()), "max": t.tolist(), "world": world}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["distinct"] and out["max"] == [11.0, 20.0]


def test_shared_capture_slices_and_arfcn_shares(tmp_path):
    """the strong-scaling wideband leg of bench.py: every rank feeds a time slice of ONE recording, an all-gather
    (int16 I/Q pairs travel as int32) rebuilds the recording on every rank, the ARFCNs and their burst rows are split
    a mod world == rank - every sample and every burst exactly once, ragged sizes included"""
    script = tmp_path / "s.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n_wide, n_arfcn, per = 10001, 7, 3                      # ragged: odd sample count, 7 ARFCNs over 2 ranks
        rec = torch.from_numpy(np.random.default_rng(5).integers(-30000, 30000, (n_wide, 2)).astype(np.int16))
        sh = bench.capture_shares(n_wide, n_arfcn, world, rank)
        mine = torch.zeros((sh["n_slice"], 2), dtype=torch.int16)
        mine[:sh["hi"] - sh["lo"]] = rec[sh["lo"]:sh["hi"]]
        parts = [torch.zeros((sh["n_slice"], 1), dtype=torch.int32) for _ in range(world)]
        dist.all_gather(parts, mine.view(torch.int32))
        full = torch.cat(parts).view(torch.int16)[:n_wide]
        rows = bench.burst_rows(sh["own"], per)
        got = [None] * world
        dist.all_gather_object(got, (sh["own"].tolist(), rows.tolist(), bool((full == rec).all())))
        if rank == 0:
            print(json.dumps(got))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29732", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert all(ok for _, _, ok in got)                          # every rank rebuilt the whole recording
    assert sorted(got[0][0] + got[1][0]) == list(range(7)) and got[0][0] == [0, 2, 4, 6]
    assert sorted(got[0][1] + got[1][1]) == list(range(21)) and got[1][1][:3] == [3, 4, 5]
