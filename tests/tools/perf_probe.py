"""Scratch performance probe (test infrastructure, uses oracle-built docs): replicate a few seeded histories
to a large batch, run the device-resident entry point a few times, print per-phase device timings."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import loro_b200
from tests import workloads


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    distinct = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    t0 = time.time()
    base = [workloads.make_doc_history(100 + k, n_sites=3, n_ops=n_ops, sync_prob=0.004, unicode_=False)[0] for k in range(distinct)]
    print("gen", time.time() - t0, "s; blob sizes", [len(b) for b in base][:4])
    blobs = [base[i % distinct] for i in range(n_docs)]
    buf, offs, lens = loro_b200.pack_blobs(blobs)
    t = torch.from_numpy(buf).cuda()
    for r in range(reps):
        torch.cuda.synchronize()
        t1 = time.time()
        b = loro_b200.import_batch_device(t.data_ptr(), offs, lens, keep=t)
        torch.cuda.synchronize()
        wall = time.time() - t1
        c, tm = b.counters(), b.timings()
        print(f"rep {r}: wall {wall*1e3:.1f} ms  ops {c['atom_ops']}  -> {c['atom_ops']/wall/1e6:.1f} Mops/s (wall) "
              f"{c['atom_ops']/(tm['total_device']*1e-3)/1e6:.1f} Mops/s (device)")
        print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()})
        print("   ", c)
        b.close()


if __name__ == "__main__":
    main()
