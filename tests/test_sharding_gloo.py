"""N>1 path on CPU: two gloo ranks shard the document batch (weak scaling: each rank generates and owns its own
docs) and exchange the per-shard summary counters with one all-gather -- the only collective of the path.
The per-shard work runs through the emulated engine here (tests/emu); on the GPU box bench.py does the same
with NCCL and the real library."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import loro_b200
from loro_b200.workload import C3Batch
from loro_b200.shard import shard_range, gather_counters
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total_docs = 10
lo, hi = shard_range(total_docs, rank, world)
gen = C3Batch(hi - lo, n_ops=300, first_doc=lo, want_json=True, threads=2)
batch = loro_b200.import_batch(gen.blobs(), lib_path={emu!r})
c = batch.counters()
assert c["docs_ok"] == hi - lo
for i in range(hi - lo):
    assert batch.json_bytes(i) == gen.expected_json(i)
allc = gather_counters(c, device="cpu")
if rank == 0:
    print(json.dumps({{"docs": [int(x["docs_ok"]) for x in allc], "ops": [int(x["atom_ops"]) for x in allc]}}))
dist.destroy_process_group()
'''


def test_two_rank_shard_and_counter_allgather(tmp_path):
    subprocess.check_call([os.path.join(HERE, "emu", "build_emu.sh")])
    emu = os.path.join(HERE, "emu", "libloro_b200_emu.so")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, emu=emu))
    env = dict(os.environ, LB_EMU_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["docs"] == [5, 5] and all(o == 300 * 5 for o in r["ops"])


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    from loro_b200.shard import shard_range
    for total in (0, 1, 7, 100, 100003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
