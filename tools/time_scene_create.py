#!/usr/bin/env python
"""Times yahr_b200_scene_create (device vs host BVH builder) on repeated builds of a workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from yahr_b200 import api
name = sys.argv[1] if len(sys.argv) > 1 else "c4-terrain"
sc, cam, desc = bench.workload(name)
for host in (False, True, False, False, True):
    if host: os.environ["YAHR_B200_HOST_BUILD"] = "1"
    else: os.environ.pop("YAHR_B200_HOST_BUILD", None)
    t = time.perf_counter(); s = api.Scene(sc); dt = (time.perf_counter() - t) * 1e3
    i = s.info(); s.close()
    print("%s builder=%s create %.1f ms (build %.1f ms, upload %.1f ms) nodes %d depth %d" %
          (name, "host" if host else "device", dt, i["build_ms"], i["upload_ms"], i["n_nodes"], i["depth"]), flush=True)
