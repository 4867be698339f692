"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank renders its strided tile subset; rank 0's gathered frame (both exchange modes) must be
bit-identical to a single-GPU render of the whole frame, and equal to the oracle on the sampled
tiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from yahr_b200 import api, scenes  # noqa: E402
from yahr_b200.dist import SharedHostFrame, TileShardedRenderer  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # YAHR_DIST_CHECK_SAME_DEVICE=1: every rank on cuda:0 (a one-GPU box).  The data plane is the same -- CUDA IPC
    # mappings of rank 0's frame, remote stores / device-to-device pushes, the shared pinned host frame -- only the
    # fence and the "reduce" baseline go through gloo, because NCCL refuses two ranks on one device.
    same_device = os.environ.get("YAHR_DIST_CHECK_SAME_DEVICE") == "1"
    if same_device:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    if rank == 0:
        print("dist_check: world=%d backend=%s devices=%d%s" % (world, dist.get_backend(), torch.cuda.device_count(),
                                                                " (all ranks on cuda:0)" if same_device else ""), flush=True)
    # a third scene with TWO point lights: the compacted two-slot path, whose probes add to the pixel with float atomics --
    # into rank 0's frame over NVLink in the p2p mode
    two_sc, two_cam = scenes.c1_scene_yahrr(400, 300)
    two_sc = dict(two_sc)
    two_sc["lights"] = np.concatenate([np.asarray(two_sc["lights"], np.float32),
                                       np.array([[-6.0, 8.0, 4.0, 300.0, 260.0, 220.0]], np.float32)])
    for name, (sc, cam) in {"bunny": scenes.c2_bunny_proxy(960, 540, nu=80, nv=60),
                            "c1": scenes.c1_scene_yahrr(517, 389), "c1-two-lights": (two_sc, two_cam)}.items():
        w, h = api.image_size(cam)
        ref = None
        if rank == 0:
            s = api.Scene(sc)
            full = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
            fpid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            s.render_device(cam, full.data_ptr(), fpid.data_ptr())
            s.close()
            ref = (full, fpid)
            # the single-GPU frame itself against the CPU oracle on every 16th tile
            from oracle import binding as ob
            o = ob.OracleScene(sc)
            orgb = np.full((h, w, 3), np.nan, np.float32)
            opid = np.full((h, w), 0xFFFFFFFE, np.uint32)
            o.render(cam, tile_stride=16, tile_offset=3, out=(orgb, opid, np.zeros((h, w), np.float32)))
            o.close()
            sel = opid != 0xFFFFFFFE
            ids_ok = np.array_equal(fpid.cpu().numpy().view(np.uint32)[sel], opid[sel])
            err = float(np.nanmax(np.abs(full.cpu().numpy()[sel].astype(np.float64) - orgb[sel].astype(np.float64))))
            print("%s single-GPU frame vs oracle on %d sampled pixels: ids equal=%s max abs err %.2e"
                  % (name, int(sel.sum()), ids_ok, err), flush=True)
            ok = ok and ids_ok and err <= 1e-3
        for mode in ("rows", "p2p", "reduce"):
            R = TileShardedRenderer(sc, cam, mode=mode, want_primid=True)
            for _ in range(3):
                R.render()
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                same_rgb = torch.equal(R.frame.view(torch.int32), ref[0].view(torch.int32))
                same_pid = torch.equal(R.primid, ref[1])
                print("%s world=%d mode=%s rgb bit-equal=%s primid equal=%s" % (name, world, mode, same_rgb, same_pid),
                      flush=True)
                ok = ok and same_rgb and same_pid
            R.close()
            dist.barrier()
        # host-buffer multi-GPU entry: every rank copies its own tile rows into one shared pinned host frame
        def barrier():
            torch.cuda.synchronize()
            dist.barrier()
        host = SharedHostFrame(w, h, rank, world, barrier=barrier)
        if rank == 0:
            host.array[...] = np.nan
        barrier()
        s = api.Scene(sc)
        for _ in range(2):
            s.render_shard(cam, rank, world, (host.array, None))
        s.close()
        barrier()
        if rank == 0:
            same = np.array_equal(host.array.view(np.uint32), ref[0].cpu().numpy().view(np.uint32))
            print("%s world=%d host shards rgb bit-equal=%s pinned=%s" % (name, world, same, host.pinned), flush=True)
            ok = ok and same
        barrier()
        host.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist_check OK")


if __name__ == "__main__":
    main()
