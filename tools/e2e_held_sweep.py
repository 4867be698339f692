#!/usr/bin/env python
"""Host-buffer entry, streamed rows: call time of yahr_b200_render (float frame) and yahr_b200_render_rgb8 on C4 terrain
for several values of the copy-run size limit (YAHR_B200_STREAM_MAX_HELD_KB, capi.cu renderStreamedRows)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402

sc, cam, _ = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c4-terrain")
w, h = api.image_size(cam)
s = api.Scene(sc)
host = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
host8 = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for held in ("512", "2048", "6144", "12288", "24576", "49152", "200000"):
    os.environ["YAHR_B200_STREAM_MAX_HELD_KB"] = held
    tf, t8 = [], []
    for i in range(14):
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.render(cam, want_primid=False, out=(host.numpy(), None))
        dt = time.perf_counter() - t
        flush.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.render_rgb8(cam, out=host8.numpy())
        d8 = time.perf_counter() - t
        if i >= 4:
            tf.append(dt)
            t8.append(d8)
    print("max held %7s KB: float %.3f ms (min %.3f)   rgb8 %.3f ms (min %.3f)"
          % (held, np.mean(tf) * 1e3, np.min(tf) * 1e3, np.mean(t8) * 1e3, np.min(t8) * 1e3), flush=True)
os.environ["YAHR_B200_HOST_STREAM"] = "0"
tf = []
for i in range(12):
    flush.zero_()
    torch.cuda.synchronize()
    t = time.perf_counter()
    s.render(cam, want_primid=False, out=(host.numpy(), None))
    if i >= 4:
        tf.append(time.perf_counter() - t)
print("copy-engine bands: float %.3f ms (min %.3f)" % (np.mean(tf) * 1e3, np.min(tf) * 1e3))
s.close()
