"""One-process-per-GPU tile-sharded rendering.

Every rank holds a replica of the (read-only) scene and renders the reference's own tile windows
strided over ranks (tile i -> rank i mod G, Sampling.hs:9-21).  The frame ends up on rank 0.
The only exchange is at the end of a frame -- the reference's `concat` of per-tile sample lists
(main.hs:83,95) -- and there are two ways to do it:

  "p2p"    tile i -> rank i mod G; rank 0's frame buffer is mapped into every rank through CUDA IPC and
           the render kernel stores each finished pixel straight into it over NVLink: the gather
           is fused into the kernel's epilogue and overlaps traversal.
  "rows"   the frame is cut into whole rows of the reference's tile grid, row r -> rank r mod G; every rank renders
           its rows into a local frame with ONE persistent kernel that publishes finished tile rows, and each
           finished row is copied into rank 0's frame (same IPC mapping) by the copy engine over NVLink while the
           rest of the share is traced (yahr_b200_render_device_shard, streamed push): bulk transfers that overlap
           the rendering.
  Both end with the flag fence (yahr_b200_flag_signal / yahr_b200_flags_wait): a sequence number per rank in rank 0's
  memory that rank 0's stream waits for -- no collective on the frame path (round 1 used a one-element NCCL
  all-reduce; fence="nccl" keeps it).
  "auto"   (default) "p2p" (C4, round 2: 0.86 / 0.56 / 0.43 ms at 2 / 4 / 8 GPUs; rows 1.00 / 0.63 / 0.47 ms).
  "reduce" every rank renders into a zeroed local full frame and the frames are summed onto rank 0
           with one NCCL reduce -- exact, because every pixel has exactly one owner and x + 0 = x.
           This is the plain-library baseline.

torch / torch.distributed are plumbing (device memory, streams, process group); all rendering goes
through the C ABI of libyahr_b200.so.
"""
import numpy as np

from . import api


class TileShardedRenderer:
    def __init__(self, scene, cam, mode="auto", want_primid=False, group=None, fence="flags"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if mode == "auto":
            # Round 2: peer stores at every GPU count.  With the per-batch kernel on a rank's share (one launch, every
            # pixel stored once, straight into rank 0's frame) and the flag fence, p2p measured 0.86 / 0.56 / 0.43 ms at
            # 2 / 4 / 8 GPUs on C4 against 1.00 / 0.63 / 0.47 ms for the streamed row pushes (profiles/r2y, r2g).
            # (Round 1, two kernels + NCCL fence: p2p 0.65 ms, rows 0.49 ms at 8 GPUs, hence "rows from 6".)
            mode = "p2p"
        self.group = group
        self.cam = cam
        self.width, self.height = api.image_size(cam)
        self.mode = mode if self.world > 1 else "local"
        import time
        t0 = time.time()
        self.scene = api.Scene(scene)
        torch.cuda.synchronize()
        t1 = time.time()
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = (self.height, self.width, 3)
        self.want_primid = want_primid
        self.frame = None         # rank 0 (and every rank in "reduce" mode): the frame tensor
        self.primid = None
        self._peer_rgb = None     # "p2p": rank 0's frame as seen from this rank
        self._peer_pid = None
        self._opened_blocks = []
        self._fence = torch.zeros(1, dtype=torch.float32, device=dev)
        if self.mode in ("local", "reduce") or self.rank == 0:
            self.frame = torch.zeros(shape, dtype=torch.float32, device=dev)
            if want_primid:
                self.primid = torch.zeros((self.height, self.width), dtype=torch.int32, device=dev)
        self.local = self.local_pid = None      # "rows": this rank's own frame (rank 0 renders straight into `frame`)
        if self.mode == "rows" and self.rank != 0:
            self.local = torch.zeros(shape, dtype=torch.float32, device=dev)
            if want_primid:
                self.local_pid = torch.zeros((self.height, self.width), dtype=torch.int32, device=dev)
        # End-of-frame fence.  "flags": every pushing rank stores the frame's sequence number into its word of a small
        # array in rank 0's memory once its rows have landed, and rank 0's stream waits for the words
        # (yahr_b200_flag_signal / yahr_b200_flags_wait: no collective, no host round trip).  "nccl": round 1's
        # one-element all-reduce on the render stream.
        self.fence = fence
        self._seq = 0
        self._flags = None          # rank 0: the flag words (one per rank)
        self._peer_flag = None      # other ranks: device pointer of this rank's word in rank 0's array
        self._peer_base_flags = None
        if self.mode in ("p2p", "rows"):
            handles = [None]
            if self.rank == 0:
                L = api.lib()
                import ctypes as C

                def export(t):
                    b = C.create_string_buffer(64)
                    api._check(L.yahr_b200_ipc_export(C.c_void_p(t.data_ptr()), b))
                    return b.raw
                # its own cudaMalloc block (>= 1 MB requests are never packed with other tensors by torch's allocator)
                self._flags = torch.zeros(1 << 18, dtype=torch.int32, device=dev)
                # torch's caching allocator sub-allocates: ship the offset inside the IPC block too
                handles = [(export(self.frame), export(self.primid) if want_primid else None,
                            self._alloc_offset(self.frame), self._alloc_offset(self.primid),
                            export(self._flags), self._alloc_offset(self._flags))]
            dist.broadcast_object_list(handles, src=0, group=group)
            if self.rank != 0:
                import ctypes as C
                L = api.lib()
                hb, hp, off_rgb, off_pid, hf, off_flags = handles[0]
                opened = {}

                def open_block(h):
                    if h not in opened:
                        p = C.c_void_p()
                        api._check(L.yahr_b200_ipc_open(h, C.byref(p)))
                        opened[h] = p.value
                    return opened[h]
                self._peer_rgb = open_block(hb) + off_rgb
                if hp is not None:
                    self._peer_pid = open_block(hp) + off_pid
                self._peer_flag = open_block(hf) + off_flags + 4 * self.rank
                self._opened_blocks = list(opened.values())
            dist.barrier(group=group)
        torch.cuda.synchronize()
        # scene_ms: BVH build + upload through the C ABI; exchange_ms: frame buffers, IPC handles, handshake
        self.timing = {"scene_ms": (t1 - t0) * 1e3, "exchange_ms": (time.time() - t1) * 1e3}

    def work_counts(self, **kw):
        """The GPU's own work counters for this rank's share (counting build of the default kernels)."""
        torch = self.torch
        scratch = torch.empty((self.height, self.width, 3), dtype=torch.float32, device="cuda")
        return self.scene.render_device_counted(self.cam, scratch.data_ptr(), tile_stride=self.world, tile_offset=self.rank,
                                                stream=torch.cuda.current_stream().cuda_stream,
                                                by_rows=(self.mode == "rows"), **kw)

    def _alloc_offset(self, t):
        """Offset of a tensor inside its cudaMalloc block (cudaIpcGetMemHandle refers to the block)."""
        if t is None:
            return 0
        import ctypes as C
        cudart = self.torch.cuda.cudart()
        # cuMemGetAddressRange through the runtime is not exposed; use the driver API via ctypes
        drv = C.CDLL("libcuda.so.1")
        base = C.c_uint64()
        size = C.c_size_t()
        rc = drv.cuMemGetAddressRange_v2(C.byref(base), C.byref(size), C.c_uint64(t.data_ptr()))
        if rc != 0:
            raise RuntimeError("cuMemGetAddressRange failed: %d" % rc)
        del cudart
        return t.data_ptr() - base.value

    def render(self, recursion_depth=1, spp=1, seed=0, traversal=api.TRAVERSAL_REFERENCE, kernel=0, tune=0):
        """Enqueue one frame on the current stream; the frame is complete on rank 0 once the
        stream has drained.  Returns nothing (no host sync).

        Frame ownership: there is ONE gather frame on rank 0 and the other ranks write into it directly ("p2p": remote
        stores from the kernels, "rows": device-to-device pushes).  The fence at the end orders the completion of THIS
        frame only, so rank 0 must have finished reading frame N (on the render stream, or synchronised with it) before
        any rank calls render() for frame N + 1 -- e.g. a barrier after the consumer, as bench.py and tests/dist_check.py
        do.  The "reduce" mode has no such constraint (every rank owns its frame)."""
        torch, dist = self.torch, self.dist
        stream = torch.cuda.current_stream().cuda_stream
        kw = dict(recursion_depth=recursion_depth, spp=spp, seed=seed, traversal=traversal, stream=stream,
                  stats=False, kernel=kernel, tune=tune)
        if self.mode == "local":
            self.scene.render_device(self.cam, self.frame.data_ptr(),
                                     self.primid.data_ptr() if self.primid is not None else None, **kw)
            return
        if self.mode == "rows":
            kw.pop("stats")
            if self.rank == 0:
                api.render_device_shard(self.scene, self.cam, 0, self.world, self.frame.data_ptr(), None,
                                        self.primid.data_ptr() if self.primid is not None else None, None, **kw)
            else:
                api.render_device_shard(self.scene, self.cam, self.rank, self.world, self.local.data_ptr(), self._peer_rgb,
                                        self.local_pid.data_ptr() if self.local_pid is not None else None,
                                        self._peer_pid, **kw)
            self._end_of_frame(stream)
            return
        kw.update(tile_stride=self.world, tile_offset=self.rank)
        if self.mode == "p2p":
            if self.rank == 0:
                rgb, pid = self.frame.data_ptr(), (self.primid.data_ptr() if self.primid is not None else None)
            else:
                rgb, pid = self._peer_rgb, self._peer_pid
            self.scene.render_device(self.cam, rgb, pid, **kw)
            self._end_of_frame(stream)
        elif self.mode == "reduce":
            self.frame.zero_()
            if self.primid is not None:
                self.primid.zero_()
            self.scene.render_device(self.cam, self.frame.data_ptr(),
                                     self.primid.data_ptr() if self.primid is not None else None, **kw)
            dist.reduce(self.frame, dst=0, op=dist.ReduceOp.SUM, group=self.group)
            if self.primid is not None:
                dist.reduce(self.primid, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        else:
            raise ValueError("unknown mode " + self.mode)

    def _end_of_frame(self, stream):
        """Completion fence on the render stream: rank 0's stream does not pass it before every rank's rows are in."""
        if self.fence == "nccl":
            self.dist.all_reduce(self._fence, group=self.group)
            return
        import ctypes as C
        L = api.lib()
        self._seq += 1
        if self.rank == 0:
            L.yahr_b200_flags_wait.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
            api._check(L.yahr_b200_flags_wait(C.c_void_p(self._flags.data_ptr()), 1, self.world - 1,
                                              C.c_uint32(self._seq), C.c_void_p(stream)))
        else:
            L.yahr_b200_flag_signal.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
            api._check(L.yahr_b200_flag_signal(C.c_void_p(self._peer_flag), C.c_uint32(self._seq), C.c_void_p(stream)))

    def stats_render(self, **kw):
        """One synchronous local render of this rank's tiles into a scratch frame, returning the
        library's stats (ray counts, kernel ms)."""
        torch = self.torch
        scratch = torch.empty((self.height, self.width, 3), dtype=torch.float32, device="cuda")
        st = self.scene.render_device(self.cam, scratch.data_ptr(), None, tile_stride=self.world,
                                      tile_offset=self.rank, stream=torch.cuda.current_stream().cuda_stream,
                                      by_rows=(self.mode == "rows"), **kw)
        return st

    def close(self):
        if self._opened_blocks:
            import ctypes as C
            self.torch.cuda.synchronize()
            for base in self._opened_blocks:             # every IPC mapping this rank opened (frame, IDs, flags)
                api.lib().yahr_b200_ipc_close(C.c_void_p(base))
            self._opened_blocks = []
            self._peer_rgb = self._peer_pid = self._peer_flag = None
        self.scene.close()


class SharedHostFrame:
    """One pinned host frame (H x W x 3 float32) shared by the per-GPU processes of a node, for the host-buffer
    multi-GPU entry (yahr_b200_render_shard): every rank copies its own tile rows into it over its own PCIe link,
    so the frame reaches host memory without any inter-GPU exchange.  POSIX shared memory, page-locked in every
    process with cudaHostRegister."""

    def __init__(self, width, height, rank, world, name=None, barrier=None):
        import os
        from multiprocessing import shared_memory
        self.rank, self.world = rank, world
        self.nbytes = int(width) * int(height) * 12
        self._owner = rank == 0
        if name is None:
            # a name nobody else can hold: rank 0 draws it (pid + random suffix) and tells the others, so that no
            # existing segment is ever unlinked on a guess
            names = [None]
            if self._owner:
                import uuid
                names = ["yahr_b200_frame_%d_%s" % (os.getpid(), uuid.uuid4().hex[:12])]
            if world > 1:
                import torch.distributed as dist
                dist.broadcast_object_list(names, src=0)
            name = names[0]
        if self._owner:
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.nbytes)
        if barrier:
            barrier()
        if not self._owner:
            self.shm = shared_memory.SharedMemory(name=name)
            # only the owner unlinks the segment: keep this process's resource tracker from "cleaning it up" (and
            # warning about it) at exit
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray((int(height), int(width), 3), np.float32, buffer=self.shm.buf)
        import torch
        rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
        self.pinned = int(rc) == 0
        if barrier:
            barrier()

    def close(self):
        import torch
        if getattr(self, "shm", None) is None:
            return
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
        self.array = None
        self.shm.close()
        if self._owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
        self.shm = None


def gather_tiles_reference(width, height, world_size, per_rank_frames):
    """Host-side statement of what the exchange must produce (used by the gloo CPU tests):
    pixel (u, v) comes from the rank that owns its tile."""
    from . import tiles
    out = np.zeros_like(per_rank_frames[0])
    wins = tiles.tile_windows(width, height)
    for i, (x0, y0, x1, y1) in enumerate(wins):
        out[y0:y1, x0:x1] = per_rank_frames[i % world_size][y0:y1, x0:x1]
    return out
