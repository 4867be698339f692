"""Diagnostic for the multi-GPU host-buffer entry (yahr_b200_render_shard), run under torchrun with N ranks:
where does the call time go?  Prints, per rank: the call's wall time into (a) the shared pinned host frame and
(b) a private pinned frame, alone and concurrently with the other ranks, and the pure D2H copy time of this
rank's share into both kinds of buffer.  With YAHR_B200_TIMELINE=1 the library adds per-band render/copy times."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from yahr_b200 import api  # noqa: E402
from yahr_b200.dist import SharedHostFrame  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    sc, cam, _ = bench.workload("c4-terrain")
    w, h = api.image_size(cam)
    s = api.Scene(sc)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    shared = SharedHostFrame(w, h, rank, world, barrier=barrier)
    private = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def calls(label, buf, solo):
        ts = []
        for i in range(6):
            flush.zero_()
            barrier()
            if solo is not None and rank != solo:
                barrier()
                continue
            t = time.perf_counter()
            st = s.render_shard(cam, rank, world, (buf, None))
            ts.append((time.perf_counter() - t) * 1e3)
            if solo is not None:
                barrier()
        if ts:
            print("rank %d %-28s call ms %s  (gpu_ms %.3f wall_ms %.3f launches %d d2h %d)" % (
                rank, label, " ".join("%.3f" % x for x in ts), st["gpu_ms"], st["wall_ms"], st["launches"], st["d2h_bytes"]),
                flush=True)

    if "copyonly" not in sys.argv[1:]:                    # "copyonly": just the concurrent pure copies below
        calls("shared, all ranks", shared.array, None)
        calls("private, all ranks", private.numpy(), None)
        for r in range(min(world, 2)):
            calls("shared, rank %d alone" % r, shared.array, r)
            calls("private, rank %d alone" % r, private.numpy(), r)

    # pure copies of this rank's share (contiguous block of the same size)
    share = (h // world) * w * 3
    dev = torch.zeros(share, dtype=torch.float32, device="cuda")
    sh_t = torch.from_numpy(shared.array.reshape(-1))
    for label, dst in (("shared", sh_t[rank * share:(rank + 1) * share]), ("private", private.view(-1)[:share])):
        tt = []
        for i in range(6):
            barrier()
            t = time.perf_counter()
            dst.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            tt.append((time.perf_counter() - t) * 1e3)
        print("rank %d copy of %d bytes into %s: ms %s" % (rank, share * 4, label, " ".join("%.3f" % x for x in tt)), flush=True)
    barrier()
    shared.close()
    s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
