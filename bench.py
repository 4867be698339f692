#!/usr/bin/env python
"""bench.py -- headline benchmark of yahr_b200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full frame of the workload: every primary ray of the image plus the shadow rays the
reference's integrator needs (Integrators.hs:59).  Metric: Mrays/s (primary + shadow), whole job.
  value : frame rendered with the scene resident in HBM and the frame left in HBM (rank 0)
  e2e   : the same frame through the host-buffer C-ABI call (yahr_b200_render): kernel parameters
          up, frame down to pinned host memory, inside the timed region
The reference arm (--impl reference) times the CPU oracle port of the reference's render loop,
tiles dealt to one std::thread per host core (the GHC reference cannot be built in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
UNIT = "Mrays/s"

# Algorithmic bytes per ray of each workload under the REFERENCE's own traversal:
#   32 B per box test + 36 B per triangle test + 36 B per triangle candidate (normals)
#   + 16 B per sphere test + 28 B per material fetch + 24 B per light fetch + 12 B per primary ray,
# counted by the oracle over the full frame (tools/count_bytes_per_ray.py; table in DESIGN.md).
def _load_bytes_per_ray():
    p = os.path.join(ROOT, "tools", "bytes_per_ray.json")
    try:
        return {k: float(v["bytes_per_ray"]) for k, v in json.load(open(p)).items()}
    except Exception:
        return {}


# `ncu --set full` capture of the default kernels on a workload, summarised by tools/ncu_summary.py from the .ncu-rep of
# this same command.  It supplies the per-frame COUNTS that do not depend on timing -- executed warp instructions, L1
# data-pipe wavefronts, DRAM bytes -- and bench.py turns them into live fractions with the kernel durations and the SM
# clock it measures itself.  A capture is only trusted when its source fingerprint is the current tree's.
NCU_SUMMARIES = {"c4-terrain": "profiles/r2_ncu_full_default_c4terrain.json",
                 # the per-batch kernel (frames / shares of <= 2.5 M work items), captured on a whole C4 frame
                 "c4-terrain/fused": "profiles/r2_ncu_full_k_wf_fused_c4terrain.json"}


def ncu_capture(workload_name):
    """(capture dict, None) or (None, reason)."""
    from yahr_b200 import api
    path = NCU_SUMMARIES.get(workload_name)
    if not path:
        return None, "no ncu capture committed for this workload"
    try:
        js = json.load(open(os.path.join(ROOT, path)))
    except Exception as e:
        return None, "cannot read %s: %s" % (path, e)
    fp = api.source_fingerprint()
    if js.get("fingerprint") != fp:
        return None, ("capture %s refused: it was taken on sources %s, the loaded library is built from %s"
                      % (path, js.get("fingerprint"), fp))
    by = {}
    for k in js["kernels"]:
        short = k["name"].replace("void ", "").split("<")[0].split("(")[0].split("::")[-1]
        if short in ("k_wf_primary", "k_wf_shadow", "k_wf_fused"):
            by[short] = k
    if "k_wf_primary" not in by and "k_wf_fused" not in by:
        return None, "capture %s holds no k_wf_primary / k_wf_fused launch" % path
    return {"path": path, "fingerprint": fp, "kernels": by}, None


BYTES_PER_RAY = _load_bytes_per_ray()


def workload(name):
    from yahr_b200 import scenes
    if name == "c4-terrain":
        sc, cam = scenes.c4_terrain()
        desc = "C4: 1M-triangle terrain (1001x501 height field), 3840x2160, 1 spp, 1 point light, BVH 32 Midpoint, depth 1"
    elif name == "c4-soup":
        sc, cam = scenes.c4_soup()
        desc = "C4: 1M-triangle random soup, 3840x2160, 1 spp, 1 point light, BVH 32 Midpoint, depth 1"
    elif name == "c3":
        sc, cam = scenes.c3_sphere_grid()
        desc = "C3: 32^3 sphere grid, 2048x2048, 1 point light, BVH 16 Midpoint, depth 1"
    elif name == "c2":
        sc, cam = scenes.c2_bunny_proxy()
        desc = "C2: bunny proxy (69 566 triangles) + floor, 1920x1080, 1 point light, BVH 24 Midpoint, depth 1"
    elif name == "c2-area":
        sc, cam = scenes.c2_bunny_proxy(area_samples=1)
        desc = ("C2 as named in BASELINE.json: bunny proxy + floor, 1920x1080, point light + quad area light (extension, 1 light "
                "sample per pixel sample; run with --spp 16), BVH 24 Midpoint, depth 1")
    elif name == "c5-area":
        sc, cam = scenes.c5_replicated_bunny(area_samples=1)
        desc = ("C5 as named in BASELINE.json: 144 copies of the bunny proxy (10M triangles), 3840x2160, point light + quad area "
                "light (extension; run with --spp 64), BVH 40 Midpoint, depth 1")
    elif name == "c1":
        sc, cam = scenes.c1_scene_yahrr()
        desc = "C1: corrected scene.yahrr (7 spheres + floor), 512x512, 1 spp, BVH 16 Midpoint, depth 1"
    elif name == "c5":
        sc, cam = scenes.c5_replicated_bunny()
        desc = "C5: 144 copies of the bunny proxy (10M triangles), 3840x2160, BVH 40 Midpoint, depth 1"
    else:
        raise SystemExit("unknown workload " + name)
    return sc, cam, desc


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline_run(sc, cam, target_seconds=15.0, threads=0, steps=1, warmup=0):
    """Times the CPU oracle port on a bounded sample of the workload's tiles, all host threads.
    Returns (dict, oracle stats of the last run)."""
    from oracle import binding as ob
    from yahr_b200 import api
    w, h = api.image_size(cam)
    if threads <= 0:
        threads = _host_cores()       # every host core this process may use (the oracle runs one std::thread per core;
                                      # it does not depend on OMP_NUM_THREADS, which torchrun sets to 1)
    t0 = time.time()
    o = ob.OracleScene(sc)
    build_s = time.time() - t0
    n_tiles = int(ob.lib().yo_num_batches(1, w, h))
    # calibrate on a thin sample, then choose a stride so one step is about target_seconds
    stride = max(1, n_tiles // 64)
    out = (np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32))
    _, _, _, st = o.render(cam, threads=threads, tile_stride=stride, tile_offset=stride // 2, out=out)
    per_tile = st["seconds"] / max(st["tiles"], 1)
    want_tiles = max(8, int(target_seconds / max(per_tile, 1e-9)))
    stride = max(1, n_tiles // want_tiles)
    vals, last = [], None
    for i in range(warmup + steps):
        _, _, _, st = o.render(cam, threads=threads, tile_stride=stride, tile_offset=stride // 2, out=out)
        rays = st["n_primary"] + st["n_shadow"] + st["n_secondary"]
        if i >= warmup:
            vals.append((rays / st["seconds"] / 1e6, st["seconds"]))
        last = st
    o.close()
    bpr, rays, _ = ob.bytes_per_ray(last)
    v = float(np.mean([x[0] for x in vals]))
    sample = ("%d of %d reference tiles (every %dth, %d rays) of the same frame; oracle C++ port of the reference, "
              "tiles pulled by one std::thread per core from a shared counter (renderPar analogue); BVH build %.1f s excluded"
              % (last["tiles"], n_tiles, stride, rays, build_s))
    d = {"value": v, "unit": UNIT, "cores": int(last["threads"]), "kind": "port", "sample": sample,
         "seconds_per_step": float(np.mean([x[1] for x in vals])), "bytes_per_ray_sample": bpr,
         "frames_per_s_extrapolated": v * 1e6 / (rays / last["tiles"] * n_tiles) if rays else None}
    return d, last


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def build_roofline(args, rays_local, rays_total, kernel_ms, phase_ms, clocks, launches_per_step, world, gpu_counts):
    """The roofline object of the JSON line.  The traversal is irregular graph work served from L1 / L2 (real DRAM
    traffic is a few % of the HBM peak), so the resource that binds is on the SM: instruction issue slots, or the L1
    data pipe.  `frac` is the larger of the two for the dominant kernel (k_wf_primary), <= 1 by construction:
        issue : warp instructions executed per launch / (4 schedulers x SMs x SM clock x kernel duration)
        l1    : L1 data-pipe wavefronts per launch    / (1 per cycle x SMs x SM clock x kernel duration)
    Counts come from the committed ncu capture of the same sources (refused when stale); durations and the clock
    are measured live in this run.  The SURVEY 8(d) algorithmic-bytes figure is kept as `algorithmic`."""
    import torch
    peak, peak_src = peaks()
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else None
    wf = launches_per_step > 1
    if not k_ms:
        return None
    ph = [float(x) for x in np.mean(np.asarray(phase_ms), axis=0)] if phase_ms else [k_ms, 0.0, 0.0, 0.0]
    n_sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    share = rays_local / max(rays_total, 1.0)
    default_run = args.depth == 1 and args.spp == 1 and args.kernel == 0 and args.tune == 0 and args.traversal == "reference"
    fused_run = wf and default_run and launches_per_step == 2
    if fused_run:
        # a frame (or a rank's share) of at most 2.5 M work items runs the per-batch kernel k_wf_fused: its own capture
        cap, why = ncu_capture(args.workload + "/fused")
    else:
        cap, why = ncu_capture(args.workload) if (wf and default_run) else (None, "not the default kernel set / configuration")
    kernels = {}
    if cap:
        for short, ms in ((("k_wf_fused", ph[0]),) if fused_run else (("k_wf_primary", ph[0]), ("k_wf_shadow", ph[2]))):
            k = cap["kernels"].get(short)
            if not k or ms <= 0:
                continue
            cyc = ms * 1e-3 * mhz * 1e6
            inst = (k.get("inst_executed") or 0.0) * share
            wavef = (k.get("l1_lsu_wavefronts") or 0.0) * share
            kernels[short] = {
                "live_ms": ms, "ms_under_ncu": k["ms"],
                "issue_frac": inst / (cyc * n_sm * 4.0), "l1_frac": wavef / (cyc * n_sm) if wavef else None,
                "warp_inst_per_launch": inst, "l1_wavefronts_per_launch": wavef or None,
                "threads_per_inst": k.get("threads_per_inst"), "l1_hit": k.get("l1_hit"), "l2_hit": k.get("l2_hit"),
                "warps_active_pct": k.get("warps_active_pct"), "registers": k.get("registers"),
                "dram_bytes_per_launch": k.get("dram_bytes") if world == 1 else None,
                "hbm_frac_actual": (k["dram_bytes"] / (ms * 1e-3) / 1e9 / peak) if (world == 1 and k.get("dram_bytes")) else None,
            }
    bpr = BYTES_PER_RAY.get(args.workload) if (args.depth == 1 and args.spp == 1) else None
    algorithmic = None
    if bpr:
        a = rays_local * bpr / (k_ms * 1e-3) / 1e9
        algorithmic = {"bytes_per_ray": bpr, "achieved_gbs": a, "frac_of_hbm_peak": a / peak,
                       "note": "SURVEY 8(d): every box / primitive fetch of every ray under the REFERENCE's traversal at "
                               "32 B per box; the scene is cache-resident and the 4-wide walk skips ancestor boxes, so "
                               "this is a work-rate figure, not an HBM fraction (it can exceed 1)"}
    dom = kernels.get("k_wf_fused" if fused_run else "k_wf_primary")
    if dom:
        bound = "issue" if dom["issue_frac"] >= (dom["l1_frac"] or 0.0) else "l1"
        frac = dom["issue_frac"] if bound == "issue" else dom["l1_frac"]
        achieved = (dom["warp_inst_per_launch"] if bound == "issue" else dom["l1_wavefronts_per_launch"]) / (dom["live_ms"] * 1e-3) / 1e9
        rpeak = n_sm * (4.0 if bound == "issue" else 1.0) * mhz * 1e6 / 1e9
        unit = "G warp-inst/s" if bound == "issue" else "G wavefronts/s"
        traffic = sum(v["dram_bytes_per_launch"] or 0.0 for v in kernels.values()) if world == 1 else None
    else:
        bound, frac, achieved, rpeak, unit, traffic = "issue", None, None, n_sm * 4.0 * mhz * 1e6 / 1e9, "G warp-inst/s", None
    return {"bound": bound, "achieved": achieved, "peak": rpeak, "unit": unit, "frac": frac, "traffic": traffic,
            "kernel": ("k_wf_fused (per-batch kernel: this share has <= 2.5 M work items)" if fused_run else
                       "k_wf_primary (primary trace + shading, dominant); k_wf_shadow listed beside it") if wf else "k_render_mega",
            "kernel_ms": k_ms, "phase_ms": {"primary_trace_and_shade": ph[0], "shadow_trace": ph[2]} if wf else None,
            "peak_source": "%d SMs x %s per cycle x %.0f MHz (median SM clock sampled during this run)"
                           % (n_sm, "4 issue slots" if bound == "issue" else "1 L1 wavefront", mhz),
            "counts_source": (cap["path"] + " (fingerprint %s = current sources)" % cap["fingerprint"]) if cap else None,
            "counts_refused": why, "kernels": kernels or None,
            "hbm": {"peak_gbs": peak, "peak_source": peak_src, "traffic_bytes_per_frame": traffic,
                    "frac_actual": (traffic / (k_ms * 1e-3) / 1e9 / peak) if traffic else None},
            "algorithmic": algorithmic, "gpu_counted": gpu_counts, "rays_per_launch": rays_local,
            "note": "HBM is not the binding roofline of this path (irregular traversal served from L1/L2; see hbm.frac_actual): "
                    "`frac` is the binding SM-side resource of the dominant kernel, counts from ncu on the same sources, "
                    "durations and clock live.  DESIGN.md section 4."}


def check_frames(R, sc, cam, args, trav, world, rank, barrier, host_frames):
    """Not timed.  (i) N > 1: rank 0's assembled frame (the exchange `value` times) bit for bit against a single-GPU
    render of the whole frame on rank 0 -- the reference's `concat` + scatter (main.hs:83,95,98-107) must not lose or
    misplace a pixel; (ii) the host frames the e2e calls filled, against the same single-GPU frame; (iii) the
    single-GPU frame against the CPU oracle on sampled reference tiles: primitive IDs bit-exact, radiance within the
    north-star tolerance.  The oracle is the checker here, never the thing measured."""
    import torch
    from yahr_b200 import api
    w, h = api.image_size(cam)
    kw = dict(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
    R.render(**kw)                                   # one more exchange: the frame to check
    barrier()
    out = None
    if rank == 0:
        full = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        fpid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        R.scene.render_device(cam, full.data_ptr(), fpid.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, **kw)
        torch.cuda.synchronize()
        out = {"n_gpus": world}
        if world > 1:
            out["vs_single_gpu"] = "bit-equal" if torch.equal(R.frame.view(torch.int32), full.view(torch.int32)) else "DIFFERS"
        ref = full.cpu().numpy()
        for name, a in host_frames.items():
            if a.dtype == np.uint8:
                same = np.array_equal(a, api.quantize_rgb8_host(ref))
            else:
                same = np.array_equal(a.view(np.uint32), ref.view(np.uint32))
            out[name + "_vs_single_gpu"] = "bit-equal" if same else "DIFFERS"
        if args.spp == 1 or args.check_oracle:
            from oracle import binding as ob
            n_tiles = int(api.num_batches(1, w, h))
            stride = max(1, n_tiles // 96)
            o = ob.OracleScene(sc)
            orgb = np.full((h, w, 3), np.nan, np.float32)
            opid = np.full((h, w), 0xFFFFFFFE, np.uint32)
            _, _, _, ost = o.render(cam, recursion_depth=args.depth, spp=args.spp, tile_stride=stride, tile_offset=stride // 3,
                                    out=(orgb, opid, np.zeros((h, w), np.float32)))
            o.close()
            sel = opid != 0xFFFFFFFE
            pid = fpid.cpu().numpy().view(np.uint32)
            a, b = ref[sel].astype(np.float64), orgb[sel].astype(np.float64)
            nan = np.isnan(a) | np.isnan(b)
            d = np.where(nan, 0.0, a - b)
            out.update({"oracle_tiles": int(ost["tiles"]), "oracle_pixels": int(sel.sum()),
                        "id_match": float((pid[sel] == opid[sel]).mean()),
                        "nan_pixels_equal": bool(np.array_equal(np.isnan(a), np.isnan(b))),
                        "max_abs_err": float(np.abs(d).max()), "rmse": float(np.sqrt((d ** 2).mean()))})
    barrier()
    return out


def extra_workloads(args, world, rank, barrier):
    """The other BASELINE.json configs as they are named there, device-resident `value` only, a few frames each: C5 (10 M
    triangles, 3840x2160, 64 spp, point + quad area light -- the config quoted for 8 GPUs; one frame is ~1.5 G rays),
    C2 (bunny proxy, 1920x1080, 16 spp, point + area light), C3 (32^3 spheres, 2048x2048, 4 spp) and the soup half of
    C4 (1 M random triangles, 3840x2160, 1 spp).  Tiles sharded over the GPUs like the headline workload."""
    import torch
    import torch.distributed as dist
    from yahr_b200.dist import TileShardedRenderer
    out = []
    for name, spp in (("c5-area", 64), ("c2-area", 16), ("c3", 4), ("c4-soup", 1)):
        sc, cam, desc = workload(name)
        R = TileShardedRenderer(sc, cam, mode=args.exchange)
        st = R.stats_render(recursion_depth=1, spp=spp)
        rays = torch.tensor([st["n_primary"] + st["n_shadow"] + st["n_secondary"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(rays)
        R.render(recursion_depth=1, spp=spp)
        barrier()
        steps = max(1, args.extra_steps)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            if world > 1:
                dist.barrier()
            ev[i][0].record()
            R.render(recursion_depth=1, spp=spp)
            ev[i][1].record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        info = R.scene.info()
        mode = R.mode
        R.close()
        del R
        m = float(ms.mean())
        out.append({"name": name, "workload": desc.replace("run with --spp 64", "64 spp").replace("run with --spp 16", "16 spp"),
                    "spp": spp, "n_gpus": world, "steps": steps, "ms_per_step": m, "value": float(rays) / (m * 1e-3) / 1e6,
                    "unit": UNIT, "rays_per_frame": float(rays), "exchange": mode, "scene_bytes": int(info["device_bytes"]),
                    "bvh_build_ms": info["build_ms"], "scaling": "strong",
                    "note": ("area lights and spp > 1 are extensions (no reference counterpart): parity against the repo's own oracle"
                             if (spp > 1 or "area" in name) else "the reference's own configuration (1 spp, point light)")})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc, cam, desc = workload(args.workload)
    d, _ = cpu_baseline_run(sc, cam, target_seconds=args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": d["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": d["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": args.workload},
            "cpu_baseline": {k: d[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": d["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "GHC reference cannot be built in this image (no ghc/cabal); this is the restated C++ oracle port"}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from yahr_b200 import api
    from yahr_b200.dist import TileShardedRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: yahr_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.build_library()
    sc, cam, desc = workload(args.workload)
    w, h = api.image_size(cam)
    trav = api.TRAVERSAL_ORDERED if args.traversal == "ordered" else api.TRAVERSAL_REFERENCE

    barrier0 = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    barrier0()                                         # process-group / context start-up is not scene creation
    t0 = time.time()
    R = TileShardedRenderer(sc, cam, mode=args.exchange)
    create_s = time.time() - t0
    info = R.scene.info()
    exchange_mode = R.mode
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ray counts and kernel-only time of this rank's share (library stats, synchronous)
    st = R.stats_render(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
    rays_local = st["n_primary"] + st["n_shadow"] + st["n_secondary"]
    rays_t = torch.tensor([rays_local, st["n_primary"], st["n_shadow"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(rays_t)
    rays_total, n_primary, n_shadow = (float(x) for x in rays_t.tolist())
    launches_per_step = st["launches"]

    # ---- value: device-resident ------------------------------------------------------------
    for _ in range(args.warmup):
        R.render(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, phase_ms = [], []
    with ClockSampler(local) as clk:
        barrier()
        for i in range(args.steps):
            flush.zero_()                                    # L2 flush between timed iterations
            if world > 1:
                dist.barrier()
            ev[i][0].record()
            R.render(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
            ev[i][1].record()
        barrier()
        # the timed region lasts only a few milliseconds (nvidia-smi samples every 200 ms): keep the same load up for
        # about 1.6 s more (at most 4000 frames), untimed, so that several clock / throttle samples are taken under this very workload
        # (the frame count is derived from the max-over-ranks step time, so it is the same on every rank: the exchange
        # contains a collective)
        probe = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev]))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(probe, op=dist.ReduceOp.MAX)
        extra = int(min(4000, max(1, 1600.0 / max(float(probe), 1e-3))))
        for k in range(extra):
            R.render(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
            if k % 20 == 19:
                torch.cuda.synchronize()
        barrier()
        # kernel-only time of the dominant kernel, measured live with CUDA events (library stats)
        for i in range(min(args.steps, 5)):
            flush.zero_()
            stk = R.stats_render(recursion_depth=args.depth, spp=args.spp, traversal=trav, kernel=args.kernel, tune=args.tune)
            kernel_ms.append(stk["gpu_ms"])
            phase_ms.append(stk["phase_ms"])
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)       # max over ranks, per step
    ms_per_step = float(step_ms.mean())
    value = rays_total / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host-buffer C-ABI call, pinned output, copies inside the timed region ----------
    e2e = None
    host_frames = {}                                  # name -> numpy view of a host frame to check afterwards (rank 0)
    if world == 1:
        host_rgb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
        out = (host_rgb.numpy(), None)
        for _ in range(args.warmup):
            _, _, ste = R.scene.render(cam, recursion_depth=args.depth, spp=args.spp, want_primid=False, out=out)
        times = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t = time.perf_counter()
            _, _, ste = R.scene.render(cam, recursion_depth=args.depth, spp=args.spp, want_primid=False, out=out)
            times.append(time.perf_counter() - t)
        e2e_ms = float(np.mean(times)) * 1e3
        # the same call with the reference's 8-bit output stage (savePngImage's quantisation, main.hs:142) done on the
        # GPU: what the CLI uses; 4x fewer bytes cross PCIe
        host_rgb8 = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
        t8 = []
        for i in range(max(1, args.warmup) + args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t = time.perf_counter()
            _, st8 = R.scene.render_rgb8(cam, recursion_depth=args.depth, spp=args.spp, out=host_rgb8.numpy())
            if i >= max(1, args.warmup):
                t8.append(time.perf_counter() - t)
        rgb8_ms = float(np.mean(t8)) * 1e3
        host_frames = {"e2e_float": host_rgb.numpy(), "e2e_rgb8": host_rgb8.numpy()}
        # pure device-to-host copy of one frame from pinned memory, for scale
        dev_frame = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
        tc = []
        for i in range(6):
            torch.cuda.synchronize()
            t = time.perf_counter()
            host_rgb.copy_(dev_frame, non_blocking=True)
            torch.cuda.synchronize()
            if i >= 2:
                tc.append(time.perf_counter() - t)
        e2e = {"value": rays_total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(ste["h2d_bytes"]), "d2h_bytes_per_step": int(ste["d2h_bytes"]),
               "api": "yahr_b200_render (host buffers; kernel parameters up, RGB32F frame down to pinned memory)",
               "launches_per_call": int(ste["launches"]),
               "strategy": ("streamed rows (fused kernel, finished tile rows copied while the frame is traced)"
                            if ste["launches"] <= 3 else "copy-engine bands") + " -- static rule of the entry (1 light slot, 1 spp -> streamed)",
               "frame_d2h_copy_alone_ms": float(np.mean(tc)) * 1e3,
               "rgb8": {"value": rays_total / (rgb8_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": rgb8_ms,
                        "d2h_bytes_per_step": int(st8["d2h_bytes"]),
                        "api": "yahr_b200_render_rgb8 (8-bit output stage on the GPU, as the yahr CLI does)"},
               "scene_create_ms": R.timing["scene_ms"], "exchange_setup_ms": R.timing["exchange_ms"],
               "renderer_init_ms": create_s * 1e3, "scene_upload_bytes": int(info["device_bytes"])}
    else:
        # every rank renders its own tile rows through the host-buffer shard entry and copies them into ONE pinned
        # host frame shared by the ranks (POSIX shared memory): N PCIe links, no inter-GPU exchange
        from yahr_b200.dist import SharedHostFrame
        host = SharedHostFrame(w, h, rank, world, barrier=barrier)
        times = []
        ste = None
        e2e_warmup = args.warmup
        for i in range(e2e_warmup + args.steps):
            flush.zero_()
            barrier()                                  # ranks leave the barrier together: a common start
            t = time.perf_counter()
            ste = R.scene.render_shard(cam, rank, world, (host.array, None), recursion_depth=args.depth, spp=args.spp)
            dt = time.perf_counter() - t               # the call returns when this rank's rows are in the host frame
            if i >= e2e_warmup:
                times.append(dt)
        # a frame is complete when the slowest rank's call has returned: max over ranks, per step
        tt = torch.tensor(times, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bytes_t = torch.tensor([float(ste["h2d_bytes"]), float(ste["d2h_bytes"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(bytes_t)
        e2e_ms = float(tt.mean()) * 1e3
        if rank == 0:
            host_frames = {"e2e_float": host.array.copy()}
        # the same with the reference's 8-bit output stage on the GPU (what the CLI writes): a quarter of the bytes
        frame8 = host.array.reshape(-1).view(np.uint8)[:h * w * 3].reshape(h, w, 3)
        t8 = []
        for i in range(e2e_warmup + args.steps):
            flush.zero_()
            barrier()
            t = time.perf_counter()
            st8 = R.scene.render_shard_rgb8(cam, rank, world, frame8, recursion_depth=args.depth, spp=args.spp)
            dt = time.perf_counter() - t
            if i >= e2e_warmup:
                t8.append(dt)
        tt8 = torch.tensor(t8, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt8, op=dist.ReduceOp.MAX)
        b8 = torch.tensor([float(st8["d2h_bytes"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(b8)
        barrier()
        if rank == 0:
            host_frames["e2e_rgb8"] = frame8.copy()
        rgb8_ms = float(tt8.mean()) * 1e3
        e2e = {"value": rays_total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(bytes_t[0]), "d2h_bytes_per_step": int(bytes_t[1]),
               "api": "yahr_b200_render_shard on every rank (tile rows r mod N == rank) into one shared pinned host frame"
                      + ("" if host.pinned else " (cudaHostRegister failed: pageable)"),
               "launches_per_call": int(ste["launches"]),
               "strategy": ("streamed rows (fused kernel, finished tile rows copied while the frame is traced)"
                            if ste["launches"] <= 3 else "copy-engine bands") + " -- static rule of the entry (1 light slot, 1 spp -> streamed)",
               "rgb8": {"value": rays_total / (rgb8_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": rgb8_ms,
                        "d2h_bytes_per_step": int(b8[0]),
                        "api": "yahr_b200_render_shard_rgb8 on every rank (8-bit output stage on the GPU, as the yahr CLI does)"},
               "scene_create_ms": R.timing["scene_ms"], "exchange_setup_ms": R.timing["exchange_ms"],
               "renderer_init_ms": create_s * 1e3, "scene_upload_bytes": int(info["device_bytes"])}
        host.close()

    # ---- frame check (outside every timed region): the frame this run produced is the right one ----------------
    frame_check = check_frames(R, sc, cam, args, trav, world, rank, barrier, host_frames)

    # the GPU's own work counters (counting build of the default kernels): visited nodes, primitive tests, probes
    gpu_counts = None
    if args.depth == 1 and args.spp == 1 and args.kernel == 0 and not args.no_gpu_counts:
        try:
            gpu_counts = R.work_counts(recursion_depth=args.depth, spp=args.spp, traversal=trav)
        except Exception as e:        # noqa: BLE001 -- a diagnostic, never fatal for the measurement
            gpu_counts = {"error": str(e)}

    # ---- roofline of the dominant kernel ----------------------------------------------------
    roofline = build_roofline(args, rays_local, rays_total, kernel_ms, phase_ms, clk.summary(), launches_per_step, world,
                              gpu_counts)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_run(sc, cam, target_seconds=args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "bytes_per_ray_sample")}

    R.close()
    del R
    extra = None
    if not args.no_extra and args.workload == "c4-terrain":
        extra = extra_workloads(args, world, rank, barrier)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "name": args.workload, "rays_per_frame": rays_total,
                           "primary": n_primary, "shadow": n_shadow, "traversal": args.traversal,
                           "exchange": exchange_mode, "l2": "flushed between timed iterations (256 MiB write)",
                           "bvh_nodes": info["n_nodes"], "bvh_depth": info["depth"],
                           "scene_bytes": info["device_bytes"], "bvh_build_ms": info["build_ms"]},
                "frames_per_s": 1e3 / ms_per_step, "clocks": clk.summary(), "e2e": e2e,
                "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu,
                "frame_check": frame_check, "extra_workloads": extra}
        print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4-terrain")
    ap.add_argument("--traversal", default="reference", choices=["reference", "ordered"])
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "rows", "reduce"])
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--tune", type=lambda x: int(x, 0), default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads block (C5, 64 spp, area light)")
    ap.add_argument("--no-gpu-counts", action="store_true", help="skip the counting-build pass")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--check-oracle", action="store_true", help="frame_check against the oracle also for spp > 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 themselves (NCCL prints its
    # version there) are sent to stderr; the JSON line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
