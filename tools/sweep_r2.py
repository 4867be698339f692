#!/usr/bin/env python
"""Round-2 sweep: kernel-only time (library CUDA events, L2 flushed) of tuning words on full frames and on one GPU's
1/N share of a frame (strided tile rows), with a bit-equality check of every variant against the first.
  python tools/sweep_r2.py --workloads c4-terrain,c4-soup --tunes 0,0x8000 --shares 1,8 [--reps 5]
Tuning word (opts.reserved[0], csrc/capi.cu planFrame): bit 10 binary tree, bit 14 4-wide tree, bit 15 k_wf_persist,
bit 29 two-kernel set pinned (tune 0 picks k_wf_fused for frames of <= 2.5 M work items), bit 12 k_wf_fused, bits 24-27 shared-memory stack entries (1 = none, 8, 12), bit 28 keep consumed queue lines in L2."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="c4-terrain")
ap.add_argument("--tunes", default="0,0x8000")
ap.add_argument("--shares", default="1")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--spp", type=int, default=1)
args = ap.parse_args()
api.build_library()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for name in args.workloads.split(","):
    sc, cam, _ = bench.workload(name)
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    info = s.info()
    for share in [int(x) for x in args.shares.split(",")]:
        ref = None
        for t in args.tunes.split(","):
            tune = int(t, 0)
            rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
            ms, ph = [], []
            for i in range(args.reps + 2):
                flush.zero_()
                st = s.render_device(cam, rgb.data_ptr(), None, tune=tune, spp=args.spp, seed=9, tile_stride=share,
                                     tile_offset=0, by_rows=share > 1)
                if i >= 2:
                    ms.append(st["gpu_ms"])
                    ph.append(st["phase_ms"][:3])
            rays = st["n_primary"] + st["n_shadow"]
            img = rgb.view(torch.int32)
            same = "-" if ref is None else str(bool(torch.equal(img, ref)))
            if ref is None:
                ref = img.clone()
            p = np.mean(np.asarray(ph), axis=0)
            print("%-11s 1/%d  tune %-11s %8.3f ms (min %.3f) %8.1f Mrays/s  phases %.3f / %.3f  launches %d  bit-equal %s"
                  % (name, share, t, np.mean(ms), np.min(ms), rays / np.mean(ms) / 1e3, p[0], p[2], st["launches"], same),
                  flush=True)
    s.close()
