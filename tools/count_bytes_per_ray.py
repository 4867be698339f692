#!/usr/bin/env python
"""Counts the algorithmic bytes per ray of a workload under the reference's own traversal
 by running the CPU oracle over the FULL frame with its work counters.
Output goes into bench.py's BYTES_PER_RAY table and DESIGN.md.  (Test/measurement tooling:
may use oracle/.)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob  # noqa: E402
import bench  # noqa: E402

names = sys.argv[1:] or ["c1", "c2", "c3", "c4-terrain", "c4-soup"]
out = {}
for name in names:
    sc, cam, desc = bench.workload(name)
    t = time.time()
    o = ob.OracleScene(sc)
    tb = time.time() - t
    _, _, _, st = o.render(cam)
    o.close()
    bpr, rays, total = ob.bytes_per_ray(st)
    out[name] = dict(bytes_per_ray=bpr, rays=rays, bytes=total, oracle_seconds=st["seconds"], threads=st["threads"],
                     build_seconds=tb, **{k: st[k] for k in st if k.startswith("n_")})
    print(name, json.dumps(out[name]), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bytes_per_ray.json"), "w"), indent=1)
