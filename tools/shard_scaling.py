#!/usr/bin/env python
"""Kernel-only time of ONE GPU's share of the frame for shard counts 1, 2, 4, 8 (strided tiles or strided tile
rows), with the L2 flushed or left warm: where single-GPU strong scaling stops being linear."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4-terrain"
sc, cam, _ = bench.workload(name)
w, h = api.image_size(cam)
s = api.Scene(sc)
rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for by_rows in (False, True):
    for stride in (1, 2, 4, 8):
        for do_flush in (True, False):
            ms, ph = [], []
            for i in range(7):
                if do_flush:
                    flush.zero_()
                st = s.render_device(cam, rgb.data_ptr(), None, tile_stride=stride, tile_offset=0, by_rows=by_rows)
                if i >= 2:
                    ms.append(st["gpu_ms"])
                    ph.append(st["phase_ms"][:3])
            p = np.mean(np.asarray(ph), axis=0)
            print("%-5s 1/%d of the frame  L2 %-7s  %.3f ms (ideal %.3f)  primary %.3f shadow %.3f  rays %d"
                  % ("rows" if by_rows else "tiles", stride, "flushed" if do_flush else "warm", np.mean(ms),
                     0, p[0], p[2], st["n_primary"] + st["n_shadow"]), flush=True)
s.close()
