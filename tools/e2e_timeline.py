import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from yahr_b200 import api
sc, cam, _ = bench.workload("c4-terrain")
w, h = api.image_size(cam)
s = api.Scene(sc)
host = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in range(6):
    flush.zero_(); torch.cuda.synchronize()
    t = time.perf_counter()
    s.render(cam, want_primid=False, out=(host.numpy(), None))
    print("call %.3f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
