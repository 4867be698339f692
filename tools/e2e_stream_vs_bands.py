"""Host-buffer entry: streamed rows (the kernels publish finished tile rows, the host thread copies them while the rest
is traced) against the copy-engine bands.  Single process, or under torchrun with N ranks (yahr_b200_render_shard into
one shared pinned frame).  Prints per workload and mode the mean call time (max over ranks); checks both frames agree."""
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank = 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    names = sys.argv[1:] or ["c4-terrain", "c2", "c3"]
    for name in names:
        sc, cam, _ = bench.workload(name)
        w, h = api.image_size(cam)
        s = api.Scene(sc)
        if world > 1:
            shared = SharedHostFrame(w, h, rank, world, barrier=barrier)
            frame = shared.array
        else:
            pinned = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
            frame = pinned.numpy()
        results = {}
        for mode in ("0", "1", "1f", "1q", "auto", "0", "1f", "auto", "auto"):
            if mode == "auto":                                   # the entry's own measured choice
                os.environ.pop("YAHR_B200_HOST_STREAM", None)
                os.environ.pop("YAHR_B200_HOST_FUSED", None)
            else:
                os.environ["YAHR_B200_HOST_STREAM"] = mode[0]
                os.environ["YAHR_B200_HOST_FUSED"] = {"f": "1", "q": "3"}.get(mode[-1], "0")
            frame[...] = np.nan
            ts = []
            for i in range(9):
                flush.zero_()
                barrier()
                t = time.perf_counter()
                if world > 1:
                    st = s.render_shard(cam, rank, world, (frame, None))
                else:
                    _, _, st = s.render(cam, want_primid=False, out=(frame, None))
                dt = (time.perf_counter() - t) * 1e3
                if i >= 3:
                    ts.append(dt)
            tt = torch.tensor(ts, dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            barrier()
            if rank == 0:
                print("%-11s world %d HOST_STREAM=%-4s  call %.3f ms (min %.3f)  gpu_ms(rank0) %.3f launches %d" % (
                    name, world, mode, float(tt.mean()), float(tt.min()), st["gpu_ms"], st["launches"]), flush=True)
                results.setdefault(mode, frame.copy())
            barrier()                                            # the other ranks reset the shared frame only now
        if rank == 0:
            same = True
            for k in ("1", "1f", "1q", "auto"):
                a, b = results["0"], results[k]
                same = same and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
            print("%-11s frames of both modes identical: %s (NaN left: %d)" % (name, same, int(np.isnan(b).sum())), flush=True)
        barrier()
        if world > 1:
            shared.close()
        s.close()
    os.environ.pop("YAHR_B200_HOST_STREAM", None)
    os.environ.pop("YAHR_B200_HOST_FUSED", None)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
