#!/usr/bin/env python
"""Times the wavefront kernel set on the bench workloads for a list of tuning words (opts.reserved[0],
see csrc/capi.cu planFrame): bits 0-7 leaf-parking threshold, bit 8 uncapped registers, bit 9 generic
node step, bit 10 binary tree instead of the 4-wide one, bits 16-23 CTAs per SM.
  python tools/sweep_tune.py [--workloads c4-terrain,c4-soup] [--tunes 0,0x400,...] [--reps 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="c4-terrain,c4-soup,c3,c2")
ap.add_argument("--tunes", default="0,0x400")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between repetitions")
args = ap.parse_args()
api.build_library()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for name in args.workloads.split(","):
    sc, cam, _ = bench.workload(name)
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    info = s.info()
    rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    ref = None
    for t in args.tunes.split(","):
        tune = int(t, 0)
        ms, ph = [], []
        for i in range(args.reps + 2):
            if not args.no_flush:
                flush.zero_()
            st = s.render_device(cam, rgb.data_ptr(), None, tune=tune, kernel=2)
            if i >= 2:
                ms.append(st["gpu_ms"])
                ph.append(st["phase_ms"][:3])
        rays = st["n_primary"] + st["n_shadow"]
        img = rgb.view(torch.int32).clone()
        same = "-" if ref is None else str(bool(torch.equal(img, ref)))
        if ref is None:
            ref = img
        p = np.mean(np.asarray(ph), axis=0)
        print("%-8s %-11s %8.3f ms %8.1f Mrays/s  primary %.3f shadow %.3f  bit-equal-to-first %s  (nodes %d wide %d)"
              % (t, name, np.mean(ms), rays / np.mean(ms) / 1e3, p[0], p[2], same, info["n_nodes"], info["n_wide_nodes"]),
              flush=True)
    s.close()
