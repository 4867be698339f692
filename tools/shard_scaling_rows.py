import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from yahr_b200 import api
sc, cam, _ = bench.workload("c4-terrain")
w, h = api.image_size(cam)
s = api.Scene(sc)
rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
for stride in (8, 16, 32, 64, 135):
    for off in (0, stride // 2):
        ms, ph = [], []
        for i in range(6):
            st = s.render_device(cam, rgb.data_ptr(), None, tile_stride=stride, tile_offset=off, by_rows=True)
            if i >= 2:
                ms.append(st["gpu_ms"]); ph.append(st["phase_ms"][:3])
        p = np.mean(np.asarray(ph), axis=0)
        print("rows 1/%d off %d: %.3f ms primary %.3f shadow %.3f rays %d" % (stride, off, np.mean(ms), p[0], p[2], st["n_primary"] + st["n_shadow"]), flush=True)
# single tile rows: which rows are slow?
for r in range(0, 135, 9):
    ms = []
    for i in range(4):
        st = s.render_device(cam, rgb.data_ptr(), None, tile_stride=135, tile_offset=r, by_rows=True)
        if i >= 1: ms.append(st["phase_ms"][0])
    print("tile row %d alone: primary %.3f ms, rays %d" % (r, np.mean(ms), st["n_primary"] + st["n_shadow"]), flush=True)
