"""world_size-2 gloo test of the rank plumbing: sharding + the end-of-run count all-reduce."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r)
    from ccsmeth_b200 import parallel
    rank, world, local = parallel.init_from_env(backend="gloo")
    s, e = parallel.shard_range(1001, rank, world)
    owned = sum(1 for b in range(7) if parallel.owns_holebatch(b, rank, world))
    tot = parallel.allreduce_counts([e - s, owned, rank, 1])
    mx = parallel.allreduce_max(10.0 + rank)
    parallel.barrier()
    assert tot == [1001, 7, 1, 2], tot
    assert mx == 11.0, mx
    parallel.finalize()
    print("rank" + str(rank) + "-ok", flush=True)
""") % ROOT


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 400)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_shard_range_covers_everything():
    from ccsmeth_b200 import parallel
    for n in (0, 1, 7, 64, 1001):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
