"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Gates (BASELINE.json north_star):
  * primary-hit primitive IDs bit-exact (reference-order traversal: 100 %; ordered traversal:
    >= 99.99 % of pixels),
  * radiance max-abs error <= 1e-3 and RMSE <= 1e-4 per channel in linear space (powf differs
    in the last ulps between glibc and CUDA, so radiance is tolerance-checked only).
"""
import numpy as np
import pytest

from oracle import binding as ob
from yahr_b200 import api, scenes

pytestmark = pytest.mark.gpu

MAX_ABS = 1e-3
RMSE = 1e-4


def torch_mod():
    import torch
    return torch


def compare(rgb, pid, orgb, opid, id_fraction=1.0, what=""):
    match = float((pid == opid).mean())
    assert match >= id_fraction, "%s: primitive IDs match on %.6f of pixels" % (what, match)
    nan_a, nan_b = np.isnan(rgb), np.isnan(orgb)
    same = pid == opid
    assert np.array_equal(nan_a[same], nan_b[same]), what + ": NaN pixels differ"
    a = np.where(nan_a | nan_b, 0, rgb).astype(np.float64)
    b = np.where(nan_a | nan_b, 0, orgb).astype(np.float64)
    sel = same if id_fraction < 1.0 else np.ones_like(same)
    diff = (a - b)[sel]
    assert np.abs(diff).max() <= MAX_ABS, "%s: max abs radiance error %g" % (what, np.abs(diff).max())
    rmse = np.sqrt((diff ** 2).mean(axis=0))
    assert (rmse <= RMSE).all(), "%s: RMSE %s" % (what, rmse)
    return match


def gpu_render_modes(sc, cam, depth=1, spp=1, seed=0):
    """Host-buffer entry + device entry in both traversal modes."""
    torch = torch_mod()
    s = api.Scene(sc)
    rgb, pid, st = s.render(cam, recursion_depth=depth, spp=spp, seed=seed)
    w, h = api.image_size(cam)
    outs = {"host": (rgb, pid, st)}
    # kernel 1 = megakernel, 2 = persistent wavefront set (recursion depth 1 only)
    kernels = (("mega", 1), ("wavefront", 2)) if depth == 1 else (("mega", 1),)
    for kname, kernel in kernels:
        for name, mode in (("reference", api.TRAVERSAL_REFERENCE), ("ordered", api.TRAVERSAL_ORDERED)):
            d_rgb = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda")
            d_pid = torch.full((h, w), 12345, dtype=torch.int32, device="cuda")
            st2 = s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr(), recursion_depth=depth, spp=spp,
                                  seed=seed, traversal=mode, stream=torch.cuda.current_stream().cuda_stream,
                                  kernel=kernel)
            outs[kname + "/" + name] = (d_rgb.cpu().numpy(), d_pid.cpu().numpy().view(np.uint32), st2)
    if depth == 1:
        # tuning bit 10: the wavefront set on the BINARY tree (default: its 4-wide collapse, wide_bvh.cu)
        d_rgb = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda")
        d_pid = torch.full((h, w), 12345, dtype=torch.int32, device="cuda")
        st2 = s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr(), recursion_depth=depth, spp=spp, seed=seed,
                              stream=torch.cuda.current_stream().cuda_stream, kernel=2, tune=0x400 | 0x20000000)
        outs["wavefront/binary"] = (d_rgb.cpu().numpy(), d_pid.cpu().numpy().view(np.uint32), st2)
        # tuning bit 29: the TWO-kernel set pinned (small frames use the per-batch kernel k_wf_fused by default), and
        # bit 12: k_wf_fused pinned
        for key, tune in (("wavefront/two-kernel", 0x20000000), ("wavefront/fused", 0x1000)):
            d_rgb = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda")
            d_pid = torch.full((h, w), 12345, dtype=torch.int32, device="cuda")
            st2 = s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr(), recursion_depth=depth, spp=spp, seed=seed,
                                  stream=torch.cuda.current_stream().cuda_stream, kernel=2, tune=tune)
            outs[key] = (d_rgb.cpu().numpy(), d_pid.cpu().numpy().view(np.uint32), st2)
        # tuning bit 30: the wavefront set on the COMPRESSED wide nodes (conservative inner boxes, exact leaf boxes)
        d_rgb = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda")
        d_pid = torch.full((h, w), 12345, dtype=torch.int32, device="cuda")
        st2 = s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr(), recursion_depth=depth, spp=spp, seed=seed,
                              stream=torch.cuda.current_stream().cuda_stream, kernel=2, tune=0x40000000 | 0x20000000)
        outs["wavefront/compressed"] = (d_rgb.cpu().numpy(), d_pid.cpu().numpy().view(np.uint32), st2)
    outs["reference"] = outs["mega/reference"]
    outs["ordered"] = outs["mega/ordered"]
    outs["info"] = s.info()
    s.close()
    return outs


SMALL = {
    "c1": lambda: scenes.c1_scene_yahrr(),
    "c1-native-aspect": lambda: scenes.c1_scene_yahrr(320, 240),
    "bunny": lambda: scenes.c2_bunny_proxy(384, 216, nu=60, nv=40),
    "bunny-depth6": lambda: _with(scenes.c2_bunny_proxy(192, 108, nu=40, nv=20), bvh_max_depth=6),
    "grid": lambda: scenes.c3_sphere_grid(12, 320, 320),
    "grid-sah": lambda: _with(scenes.c3_sphere_grid(8, 256, 256), split_mode=1),
    "terrain": lambda: scenes.c4_terrain(121, 61, 384, 216),
    "terrain-sah": lambda: _with(scenes.c4_terrain(61, 31, 192, 108), split_mode=1),
    "soup": lambda: scenes.c4_soup(30000, 384, 216),
    "adversarial": lambda: scenes.adversarial_shared_edges(),
    "adversarial-sah": lambda: scenes.adversarial_shared_edges(split_mode=1),
    "adversarial-depth3": lambda: _with(scenes.adversarial_shared_edges(128, 128), bvh_max_depth=3),
    # extension (no reference counterpart): quad area lights, alone and next to point lights
    "bunny-area": lambda: scenes.c2_bunny_proxy(256, 144, nu=40, nv=20, area_samples=4),
    "c1-area-only": lambda: _area_only(scenes.c1_scene_yahrr(192, 192)),
    # IEEE corner cases: zero-area triangles, camera inside a sphere, NaN shading frames, exact t ties, a light on a surface
    "degenerate": lambda: scenes.degenerate_mix(),
    "degenerate-sah": lambda: _with(scenes.degenerate_mix(96, 72), split_mode=1),
}


def _area_only(sc_cam):
    sc, cam = sc_cam
    sc = dict(sc)
    sc["lights"] = np.zeros((0, 6), np.float32)
    scenes.add_area_light(sc, [2, 9, -12], [4, 0, 0], [0, 0, 4], [30, 30, 30], 1)
    return sc, cam


def _with(sc_cam, **kw):
    sc, cam = sc_cam
    sc = dict(sc)
    sc.update(kw)
    return sc, cam


@pytest.mark.parametrize("name", list(SMALL.keys()))
def test_parity_small_scenes(name):
    sc, cam = SMALL[name]()
    o = ob.OracleScene(sc)
    orgb, opid, _, ost = o.render(cam)
    o.close()
    outs = gpu_render_modes(sc, cam)
    for mode in ("host", "mega/reference", "wavefront/reference", "wavefront/binary", "wavefront/compressed",
                 "wavefront/two-kernel", "wavefront/fused"):
        rgb, pid, st = outs[mode]
        compare(rgb, pid, orgb, opid, 1.0, "%s/%s" % (name, mode))
        assert st["n_primary"] == ost["n_primary"]
        assert st["n_shadow"] == ost["n_shadow"], "shadow-ray count differs from Integrators.hs:59 semantics"
        assert st["launches"] >= 1
    # the compressed walk visits a superset of inner nodes and the same leaves: frames and ray counts bit-identical
    a, b = outs["wavefront/two-kernel"], outs["wavefront/compressed"]
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
    assert a[2]["n_shadow"] == b[2]["n_shadow"]
    for mode in ("mega/ordered", "wavefront/ordered"):
        rgb, pid, st = outs[mode]
        compare(rgb, pid, orgb, opid, 0.9999, name + "/" + mode)
    # the host entry (default kernel set) and the device entry are the same computation, and the two
    # kernel sets agree bit for bit on IDs (radiance: same arithmetic, so bit-equal as well)
    assert np.array_equal(outs["host"][1], outs["wavefront/reference"][1])
    assert np.array_equal(outs["mega/reference"][1], outs["wavefront/reference"][1])
    for other in ("wavefront/reference", "wavefront/binary"):
        a, b = outs["mega/reference"][0], outs[other][0]
        nan = np.isnan(a) | np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(np.where(nan, 0, a), np.where(nan, 0, b))
    if outs["info"]["n_nodes"] > 0:
        assert outs["info"]["n_wide_nodes"] > 0, "the 4-wide tree was not built"


def _wide_leaf_sequence(wide, root_ref):
    """Leaf refs of the 4-wide tree in left-first order + (ref -> box) of every child slot."""
    refs = wide[:, 24:28].view(np.uint32)
    seq, boxes = [], {}
    stack = [("n", root_ref)]
    while stack:
        _, ref = stack.pop()
        if ref & 0x80000000:
            seq.append(int(ref))
            continue
        node = wide[ref]
        n = int(node[28:29].view(np.uint32)[0])
        assert 2 <= n <= 4
        kids = []
        for k in range(4):
            r = int(refs[ref, k])
            if k >= n:
                assert r == 0xFFFFFFFF and node[4 * k] == np.inf and node[4 * k + 2] == -np.inf
                continue
            lo = (node[4 * k], node[4 * k + 1], node[16 + 2 * k])
            hi = (node[4 * k + 2], node[4 * k + 3], node[16 + 2 * k + 1])
            if r & 0x80000000:
                boxes[r] = lo + hi
            kids.append(r)
        for r in reversed(kids):
            stack.append(("n", r))
    return seq, boxes


def _binary_leaf_sequence(nodes, root_ref):
    refs = nodes[:, 12:14].view(np.uint32)
    seq, boxes = [], {}
    stack = [root_ref]
    while stack:
        ref = stack.pop()
        if ref & 0x80000000:
            seq.append(int(ref))
            continue
        nd = nodes[ref]
        l, r = int(refs[ref, 0]), int(refs[ref, 1])
        if l & 0x80000000:
            boxes[l] = (nd[0], nd[1], nd[8], nd[2], nd[3], nd[9])
        if r & 0x80000000:
            boxes[r] = (nd[4], nd[5], nd[10], nd[6], nd[7], nd[11])
        stack.append(r)
        stack.append(l)
    return seq, boxes


@pytest.mark.parametrize("name", ["terrain", "grid-sah", "bunny-depth6", "soup"])
def test_wide_tree_is_a_collapse_of_the_reference_tree(name):
    """The 4-wide tree must present the same leaves, in the same left-first order, with the same leaf boxes
    as the binary tree restating Culling.hs -- that is all the traversal result depends on (wide_bvh.cu)."""
    sc, cam = SMALL[name]()
    s = api.Scene(sc)
    _, nodes, _, root, _ = s.download_bvh()
    wide = s.download_wide()
    s.close()
    assert len(wide) > 0 and len(wide) < len(nodes)
    bseq, bbox = _binary_leaf_sequence(nodes, root)
    wseq, wbox = _wide_leaf_sequence(wide, root)
    assert bseq == wseq
    assert set(bbox) == set(wbox)
    for r in bbox:
        assert tuple(map(float, bbox[r])) == tuple(map(float, wbox[r]))


def _decode_compressed(cw):
    """Child boxes of the compressed wide nodes, decoded with the kernels' own arithmetic in binary32:
    fma(2^23 + q, 2^e, origin') -- exact by construction (csrc/wide_bvh.cu)."""
    org = cw[:, 0:3].view(np.float32)
    eb = cw[:, 3]
    step = np.stack([((eb >> (8 * d)) & 0xFF).astype(np.uint32) << 23 for d in range(3)], axis=1).view(np.float32)
    words = {"lo": cw[:, 4:7], "hi": np.stack([cw[:, 7], cw[:, 8], cw[:, 9]], axis=1)}
    out = {}
    for side in ("lo", "hi"):
        box = np.zeros((len(cw), 4, 3), np.float64)
        for k in range(4):
            q = ((words[side] >> (8 * k)) & 0xFF).astype(np.uint32)
            f23 = (q | 0x4B000000).view(np.float32).astype(np.float64)           # 2^23 + q
            exact = f23 * step.astype(np.float64) + org.astype(np.float64)       # exact in binary64
            assert np.array_equal(exact.astype(np.float32).astype(np.float64), exact), "decoded coordinate not a float"
            box[:, k, :] = exact
        out[side] = box
    return out["lo"], out["hi"]


@pytest.mark.parametrize("name", ["terrain", "grid-sah", "bunny-depth6", "soup", "c1", "adversarial"])
def test_compressed_nodes_contain_the_exact_boxes_and_leaf_boxes_are_exact(name):
    """The compressed walk (64-byte nodes, csrc/wide_bvh.cu) may only change WHICH inner nodes are visited, never a
    hit: every decoded child box must contain the exact child box of the 4-wide node (supersets only add visits),
    every decoded coordinate must be exactly representable in binary32, empty slots must decode to inverted boxes,
    the child refs must be the 4-wide node's, and the boxes tested at the leaves must be the reference's own leaf
    boxes bit for bit (Culling.hs:33,52; AABBs.hs:42-43)."""
    sc, cam = SMALL[name]()
    s = api.Scene(sc)
    _, nodes, _, root, _ = s.download_bvh()
    wide = s.download_wide()
    cw, leaf_box, multi_box = s.download_compressed()
    s.close()
    assert len(cw) == len(wide) > 0
    assert np.array_equal(cw[:, 12:16], wide[:, 24:28].view(np.uint32))
    lo, hi = _decode_compressed(cw)
    n = wide[:, 28].view(np.uint32)
    for k in range(4):
        live = n > k
        exact_lo = np.stack([wide[:, 4 * k], wide[:, 4 * k + 1], wide[:, 16 + 2 * k]], axis=1).astype(np.float64)
        exact_hi = np.stack([wide[:, 4 * k + 2], wide[:, 4 * k + 3], wide[:, 16 + 2 * k + 1]], axis=1).astype(np.float64)
        assert (lo[live, k] <= exact_lo[live]).all() and (hi[live, k] >= exact_hi[live]).all()
        assert (lo[~live, k] > hi[~live, k]).all()                   # empty slot: inverted on every axis
        # the grid is tight: a decoded side is less than two grid steps away from the exact one
        eb = cw[:, 3]
        step = np.stack([(((eb >> (8 * d)) & 0xFF).astype(np.uint32) << 23).view(np.float32) for d in range(3)], axis=1)
        assert ((exact_lo[live] - lo[live, k]) < 2 * step[live]).all() and ((hi[live, k] - exact_hi[live]) < 2 * step[live]).all()
    _, bbox = _binary_leaf_sequence(nodes, root)
    assert len(bbox) > 0
    for ref, box in bbox.items():
        src = multi_box if (ref & 0xC0000000) == 0xC0000000 else leaf_box
        got = src[ref & 0x3FFFFFFF]
        want = np.array([box[0], box[1], box[2], box[3], box[4], box[5]], np.float32)
        assert np.array_equal(got[:6].view(np.uint32), want.view(np.uint32)), "leaf %x" % ref


@pytest.mark.parametrize("depth", [0, 2, 3])
def test_parity_recursion_depth(depth):
    """Whitted recursion (Integrators.hs:37-43); scene.yahrr ships with depth 3."""
    sc, cam = scenes.c1_scene_yahrr(256, 192)
    if depth == 2:        # also with an area light (extension): the sample points depend on the recursion level
        sc = dict(sc)
        scenes.add_area_light(sc, [2, 9, -12], [4, 0, 0], [0, 0, 4], [30, 30, 30], 3)
    o = ob.OracleScene(sc)
    orgb, opid, _, ost = o.render(cam, recursion_depth=depth)
    o.close()
    outs = gpu_render_modes(sc, cam, depth=depth)
    rgb, pid, st = outs["reference"]
    compare(rgb, pid, orgb, opid, 1.0, "c1 depth %d" % depth)
    assert st["n_secondary"] == ost["n_secondary"]
    assert st["n_shadow"] == ost["n_shadow"]
    # the host-buffer entry: depth 3 with the one point light goes through the recursion kernel (k_wf_fused_depth)
    hrgb, hpid, hst = outs["host"]
    compare(hrgb, hpid, orgb, opid, 1.0, "c1 depth %d, host entry" % depth)
    assert hst["n_secondary"] == ost["n_secondary"] and hst["n_shadow"] == ost["n_shadow"]


def test_recursion_kernel_is_for_point_lights_only():
    """Area lights (extension) with recursion stay on the megakernel; asking for the wavefront kernels is an error."""
    torch = torch_mod()
    sc, cam = SMALL["bunny-area"]()
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    d_rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    with pytest.raises(api.YahrError):
        s.render_device(cam, d_rgb.data_ptr(), None, recursion_depth=2, kernel=2)
    a = s.render_device(cam, d_rgb.data_ptr(), None, recursion_depth=2, kernel=0)
    b = s.render_device(cam, d_rgb.data_ptr(), None, recursion_depth=2, kernel=1)
    assert a["launches"] == b["launches"]
    s.close()


@pytest.mark.parametrize("name", ["c1", "bunny", "grid", "terrain", "adversarial", "adversarial-sah", "degenerate",
                                  "bunny-depth6"])
def test_recursion_kernel_equals_megakernel(name):
    """recursionDepth >= 2 with point lights (one: k_wf_fused_depth, several: k_wf_fused_depth_lights -- the default):
    the per-batch wavefront kernels against the megakernel: same primitive IDs, same ray counts, bit-identical radiance (NaN pixels in the same places), on the
    device-resident frame, through the host-buffer entry (bands and streamed rows) and against the oracle."""
    torch = torch_mod()
    sc, cam = SMALL[name]()
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    o = ob.OracleScene(sc)

    def same(a, b):
        na, nb = np.isnan(a), np.isnan(b)
        return np.array_equal(na, nb) and np.array_equal(np.where(na, 0, a).view(np.uint32), np.where(nb, 0, b).view(np.uint32))

    for depth in (2, 3, 5):
        frames = {}
        for kernel in (1, 2, 0):
            d_rgb = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda")
            d_pid = torch.full((h, w), 12345, dtype=torch.int32, device="cuda")
            st = s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr(), recursion_depth=depth, kernel=kernel)
            frames[kernel] = (d_rgb.cpu().numpy(), d_pid.cpu().numpy().view(np.uint32), st)
        assert frames[0][2]["launches"] == frames[2][2]["launches"] != frames[1][2]["launches"], "default is the recursion kernel"
        for k in (2, 0):
            assert np.array_equal(frames[k][1], frames[1][1]), "%s depth %d: primitive IDs" % (name, depth)
            assert same(frames[k][0], frames[1][0]), "%s depth %d: radiance" % (name, depth)
            for key in ("n_primary", "n_shadow", "n_secondary"):
                assert frames[k][2][key] == frames[1][2][key], key
        for _ in range(5):                       # host-buffer entry: bands three times, then streamed rows
            rgb, pid, _ = s.render(cam, recursion_depth=depth)
            assert np.array_equal(pid, frames[1][1]) and same(rgb, frames[1][0])
        if depth == 3:
            orgb, opid, _, ost = o.render(cam, recursion_depth=depth)
            compare(frames[2][0], frames[2][1], orgb, opid, 1.0, "%s depth %d vs oracle" % (name, depth))
            assert frames[2][2]["n_secondary"] == ost["n_secondary"] and frames[2][2]["n_shadow"] == ost["n_shadow"]
    o.close()
    s.close()


def test_parity_spp_extension():
    """spp > 1 (extension): same counter-based jitter in oracle and kernel; sample 0 = reference ray."""
    sc, cam = scenes.c2_bunny_proxy(160, 90, nu=40, nv=20, area_samples=2)
    o = ob.OracleScene(sc)
    orgb, opid, _, _ = o.render(cam, spp=4, seed=0x1234ABCD5678)
    o1rgb, o1pid, _, _ = o.render(cam, spp=1)
    o.close()
    outs = gpu_render_modes(sc, cam, spp=4, seed=0x1234ABCD5678)
    for mode in ("mega/reference", "wavefront/reference", "host"):
        rgb, pid, st = outs[mode]
        compare(rgb, pid, orgb, opid, 1.0, "bunny spp4 " + mode)
        assert np.array_equal(pid, o1pid)          # primitive ID is sample 0's
        assert st["n_primary"] == 4 * 160 * 90


def test_tile_sharding_partitions_the_image():
    """Rendering the strided tile subsets of G ranks into one buffer equals the full render."""
    torch = torch_mod()
    sc, cam = scenes.c2_bunny_proxy(384, 216, nu=40, nv=20)
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    full = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    fpid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    s.render_device(cam, full.data_ptr(), fpid.data_ptr())
    for G in (2, 3, 8):
        acc = torch.full((h, w, 3), float("nan"), dtype=torch.float32, device="cuda")
        apid = torch.full((h, w), -7, dtype=torch.int32, device="cuda")
        tiles = 0
        for r in range(G):
            st = s.render_device(cam, acc.data_ptr(), apid.data_ptr(), tile_stride=G, tile_offset=r)
            tiles += st["tiles"]
        assert tiles == api.num_batches(1, w, h)
        assert torch.equal(apid, fpid)
        assert torch.equal(acc.view(torch.int32), full.view(torch.int32))
    s.close()


@pytest.mark.parametrize("stream", ["0", "1"])
def test_host_buffer_shards_assemble_the_frame(stream, monkeypatch):
    """yahr_b200_render_shard: the row shards of G GPUs written into one host frame equal the single call, with the
    copy-engine bands and with the streamed rows."""
    sc, cam = scenes.c2_bunny_proxy(384, 216, nu=40, nv=20)
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    full, fpid, fst = s.render(cam)
    monkeypatch.setenv("YAHR_B200_HOST_STREAM", stream)
    for G in (1, 2, 3, 8, 64):
        rgb = np.full((h, w, 3), np.nan, np.float32)
        pid = np.full((h, w), 0xABCDEF01, np.uint32)
        rays = 0
        for k in range(G):
            st = s.render_shard(cam, k, G, (rgb, pid))
            rays += st["n_primary"]
        assert rays == w * h
        assert np.array_equal(pid, fpid)
        assert np.array_equal(rgb.view(np.uint32), full.view(np.uint32))
    with pytest.raises(api.YahrError):
        s.render_shard(cam, 2, 2, (rgb, pid))
    s.close()


def test_streamed_rows_equal_the_banded_copies_and_the_device_frame(monkeypatch):
    """Host-buffer entry: the streamed-row path (kernels publish finished tile rows, the host thread copies them while
    the rest is traced) against the copy-engine bands and against the device-resident frame: bit-identical, for
    pageable and pinned destinations, repeated calls (row sequence numbers) and changing image sizes."""
    torch = torch_mod()
    sc, _ = scenes.c2_bunny_proxy(384, 216, nu=40, nv=20)
    s = api.Scene(sc)
    monkeypatch.setenv("YAHR_B200_POISON_FRAME", "1")      # a row copied before its pixels landed would carry NaN
    for (w, h) in [(384, 216), (517, 389), (33, 70), (384, 216)]:
        _, cam = scenes.c2_bunny_proxy(w, h, nu=40, nv=20)
        dev = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        dpid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
        s.render_device(cam, dev.data_ptr(), dpid.data_ptr())
        ref, rpid = dev.cpu().numpy(), dpid.cpu().numpy().view(np.uint32)
        monkeypatch.setenv("YAHR_B200_HOST_STREAM", "0")
        b_rgb, b_pid, b_st = s.render(cam)
        monkeypatch.delenv("YAHR_B200_HOST_STREAM")
        monkeypatch.setenv("YAHR_B200_HOST_STREAM", "1")
        for fused in ("0", "1", "2", "3"):        # streamed rows with the two-kernel set / k_wf_fused (both builds) / k_wf_persist
            monkeypatch.setenv("YAHR_B200_HOST_FUSED", fused)
            t_rgb, t_pid, _ = s.render(cam)
            assert np.array_equal(t_rgb.view(np.uint32), ref.view(np.uint32)) and np.array_equal(t_pid, rpid)
        monkeypatch.delenv("YAHR_B200_HOST_FUSED")
        monkeypatch.delenv("YAHR_B200_HOST_STREAM")
        for tune in (0x1000, 0x3000, 0x8000, 0x8200):     # k_wf_fused / k_wf_persist on the device-resident frame
            fdev = torch.full((h, w, 3), float("nan"), dtype=torch.float32, device="cuda")
            s.render_device(cam, fdev.data_ptr(), None, tune=tune)
            assert torch.equal(fdev.view(torch.int32), dev.view(torch.int32))
        pinned = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
        for rep in range(8):                   # unpinned strategy: three calls with the bands, three streamed, then the faster
            rgb, pid, _ = s.render(cam)
            assert np.array_equal(rgb.view(np.uint32), ref.view(np.uint32)) and np.array_equal(pid, rpid)
        monkeypatch.setenv("YAHR_B200_HOST_STREAM", "1")
        for rep in range(3):
            rgb, pid, st = s.render(cam)                                       # pageable destination
            assert np.array_equal(rgb.view(np.uint32), ref.view(np.uint32)) and np.array_equal(pid, rpid)
            pinned.fill_(float("nan"))
            s.render(cam, want_primid=False, out=(pinned.numpy(), None))       # pinned destination
            assert np.array_equal(pinned.numpy().view(np.uint32), ref.view(np.uint32))
        monkeypatch.delenv("YAHR_B200_HOST_STREAM")
        assert np.array_equal(b_rgb.view(np.uint32), ref.view(np.uint32)) and np.array_equal(b_pid, rpid)
        assert st["launches"] < b_st["launches"] or h < 40       # one launch of the kernel set instead of one per band
        assert st["d2h_bytes"] == b_st["d2h_bytes"] == w * h * 16
    s.close()


def test_streamed_rows_full_size_stress(monkeypatch):
    """BASELINE config C4 at full size through the streamed-row path, 40 frames with the device frame poisoned before
    every call: every frame must equal the device-resident render bit for bit (a tile row published before all its
    pixel stores had landed would show up as NaN)."""
    torch = torch_mod()
    sc, cam = scenes.c4_terrain()
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    dev = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    s.render_device(cam, dev.data_ptr(), None)
    ref = dev.cpu().numpy().view(np.uint32)
    pinned = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    monkeypatch.setenv("YAHR_B200_POISON_FRAME", "1")
    monkeypatch.setenv("YAHR_B200_HOST_STREAM", "1")
    for rep in range(40):
        pinned.fill_(float("nan"))
        s.render(cam, want_primid=False, out=(pinned.numpy(), None))
        assert np.array_equal(pinned.numpy().view(np.uint32), ref), "frame %d differs" % rep
    s.close()


def test_full_size_c4_terrain_properties_and_sampled_oracle():
    """BASELINE config C4 (1M triangles, 3840x2160): determinism, and oracle parity on every
    64th tile of the full-size frame."""
    torch = torch_mod()
    sc, cam = scenes.c4_terrain()
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    info = s.info()
    assert info["n_primitives"] == 1_000_000
    a = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    ap = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    b = torch.zeros_like(a)
    bp = torch.zeros_like(ap)
    st = s.render_device(cam, a.data_ptr(), ap.data_ptr())
    s.render_device(cam, b.data_ptr(), bp.data_ptr(), traversal=api.TRAVERSAL_ORDERED)
    assert st["n_primary"] == w * h
    # ordered traversal agrees with reference order on >= 99.99 % of pixels
    assert float((ap == bp).float().mean()) >= 0.9999
    # idempotence / determinism
    c = torch.zeros_like(a)
    cp = torch.zeros_like(ap)
    s.render_device(cam, c.data_ptr(), cp.data_ptr())
    assert torch.equal(ap, cp) and torch.equal(a.view(torch.int32), c.view(torch.int32))
    s.close()
    # sampled oracle parity at full size
    o = ob.OracleScene(sc)
    stride = 64
    orgb = np.full((h, w, 3), np.nan, np.float32)
    opid = np.full((h, w), 0xFFFFFFFE, np.uint32)
    ot = np.zeros((h, w), np.float32)
    o.render(cam, tile_stride=stride, tile_offset=5, out=(orgb, opid, ot))
    o.close()
    sel = opid != 0xFFFFFFFE
    assert sel.sum() > 100000
    rgb = a.cpu().numpy()
    pid = ap.cpu().numpy().view(np.uint32)
    assert np.array_equal(pid[sel], opid[sel])
    d = (rgb[sel].astype(np.float64) - orgb[sel].astype(np.float64))
    assert np.abs(d).max() <= MAX_ABS
    assert (np.sqrt((d ** 2).mean(axis=0)) <= RMSE).all()


@pytest.mark.parametrize("area", [False, True])
def test_full_size_c5_sampled_oracle(area):
    """BASELINE config C5 (144 bunny copies, 10 M triangles, BVH 40, 3840x2160): oracle parity on every 256th tile of
    the full-size frame; with the area light + 2 spp (both extensions) as the config names them."""
    sc, cam = scenes.c5_replicated_bunny(area_samples=1 if area else 0)
    spp = 2 if area else 1
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    info = s.info()
    assert info["n_primitives"] == 10_017_218 and info["n_wide_nodes"] > 0
    rgb, pid, st = s.render(cam, spp=spp, seed=77)
    s.close()
    assert st["n_primary"] == spp * w * h
    o = ob.OracleScene(sc)
    orgb = np.full((h, w, 3), np.nan, np.float32)
    opid = np.full((h, w), 0xFFFFFFFE, np.uint32)
    ot = np.zeros((h, w), np.float32)
    o.render(cam, spp=spp, seed=77, tile_stride=256, tile_offset=7, out=(orgb, opid, ot))
    o.close()
    sel = opid != 0xFFFFFFFE
    assert sel.sum() > 30000
    assert np.array_equal(pid[sel], opid[sel])
    d = (rgb[sel].astype(np.float64) - orgb[sel].astype(np.float64))
    assert np.abs(d).max() <= MAX_ABS
    assert (np.sqrt((d ** 2).mean(axis=0)) <= RMSE).all()


def _sampled_oracle(sc, cam, spp, seed, stride, offset):
    w, h = api.image_size(cam)
    o = ob.OracleScene(sc)
    orgb = np.full((h, w, 3), np.nan, np.float32)
    opid = np.full((h, w), 0xFFFFFFFE, np.uint32)
    o.render(cam, spp=spp, seed=seed, tile_stride=stride, tile_offset=offset, out=(orgb, opid, np.zeros((h, w), np.float32)))
    o.close()
    return orgb, opid, opid != 0xFFFFFFFE


def _check_sampled(rgb, pid, orgb, opid, sel, what):
    assert np.array_equal(pid[sel], opid[sel]), what + ": primitive IDs differ from the oracle"
    a, b = rgb[sel], orgb[sel]
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), what + ": NaN pixels differ"
    d = np.where(nan_a, 0, a).astype(np.float64) - np.where(nan_b, 0, b).astype(np.float64)
    assert np.abs(d).max() <= MAX_ABS, "%s: max abs radiance error %g" % (what, np.abs(d).max())
    assert (np.sqrt((d ** 2).mean(axis=0)) <= RMSE).all(), what + ": RMSE"


def test_full_size_c2_sampled_oracle():
    """BASELINE config C2 as named: bunny proxy (69 566 triangles), 1920x1080, 16 spp, point + area light; oracle
    parity on every 64th tile of the full-size frame, plus the 1 spp point-light frame (the reference's own
    configuration) through the host entry and the device entry."""
    torch = torch_mod()
    sc, cam = scenes.c2_bunny_proxy(area_samples=1)
    w, h = api.image_size(cam)
    assert (w, h) == (1920, 1080)
    s = api.Scene(sc)
    assert s.info()["n_primitives"] >= 69_000
    rgb, pid, st = s.render(cam, spp=16, seed=2024)
    assert st["n_primary"] == 16 * w * h
    s.close()
    orgb, opid, sel = _sampled_oracle(sc, cam, 16, 2024, 64, 11)
    assert sel.sum() > 30000
    _check_sampled(rgb, pid, orgb, opid, sel, "C2 16 spp area")
    sc1, _ = scenes.c2_bunny_proxy()
    s = api.Scene(sc1)
    rgb, pid, _ = s.render(cam)
    d_rgb = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    d_pid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    s.render_device(cam, d_rgb.data_ptr(), d_pid.data_ptr())
    s.close()
    assert np.array_equal(rgb.view(np.uint32), d_rgb.cpu().numpy().view(np.uint32))
    assert np.array_equal(pid, d_pid.cpu().numpy().view(np.uint32))
    orgb, opid, sel = _sampled_oracle(sc1, cam, 1, 0, 16, 5)
    _check_sampled(rgb, pid, orgb, opid, sel, "C2 1 spp")


def test_full_size_c3_sampled_oracle():
    """BASELINE config C3 as named: 32^3 sphere grid (generalises Spec.hs:221-262), 2048x2048, 4 spp, shadow rays;
    oracle parity on every 128th tile of the full-size frame, and at 1 spp on every 64th."""
    sc, cam = scenes.c3_sphere_grid()
    w, h = api.image_size(cam)
    assert (w, h) == (2048, 2048)
    s = api.Scene(sc)
    assert s.info()["n_primitives"] == 32 ** 3
    rgb4, pid4, st4 = s.render(cam, spp=4, seed=5)
    rgb1, pid1, st1 = s.render(cam)
    s.close()
    assert st4["n_primary"] == 4 * w * h and st1["n_primary"] == w * h and st1["n_shadow"] > 0
    orgb, opid, sel = _sampled_oracle(sc, cam, 4, 5, 128, 9)
    assert sel.sum() > 20000
    _check_sampled(rgb4, pid4, orgb, opid, sel, "C3 4 spp")
    orgb, opid, sel = _sampled_oracle(sc, cam, 1, 0, 64, 3)
    _check_sampled(rgb1, pid1, orgb, opid, sel, "C3 1 spp")


def test_full_size_c4_soup_sampled_oracle():
    """BASELINE config C4, soup half: 1 M random triangles at 3840x2160 -- the hardest case for the reference's
    left-first order (Culling.hs:24-25: 235 box tests per ray, deep stacks).  Oracle parity on every 256th tile; the
    wavefront set on the binary tree and on its 4-wide collapse must agree bit for bit on the whole frame."""
    torch = torch_mod()
    sc, cam = scenes.c4_soup()
    w, h = api.image_size(cam)
    assert (w, h) == (3840, 2160)
    s = api.Scene(sc)
    assert s.info()["n_primitives"] == 1_000_000
    a = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    ap = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    b = torch.zeros_like(a)
    bp = torch.zeros_like(ap)
    st = s.render_device(cam, a.data_ptr(), ap.data_ptr())
    s.render_device(cam, b.data_ptr(), bp.data_ptr(), kernel=2, tune=0x400)          # binary tree
    assert torch.equal(ap, bp) and torch.equal(a.view(torch.int32), b.view(torch.int32))
    s.render_device(cam, b.data_ptr(), bp.data_ptr(), kernel=2, tune=0x4000)         # 4-wide tree pinned
    assert torch.equal(ap, bp) and torch.equal(a.view(torch.int32), b.view(torch.int32))
    for tune in (0x40000000, 0x80000000 - 2 ** 32):                                  # compressed / exact wide nodes pinned
        s.render_device(cam, b.data_ptr(), bp.data_ptr(), kernel=2, tune=tune)
        assert torch.equal(ap, bp) and torch.equal(a.view(torch.int32), b.view(torch.int32))
    s.close()
    assert st["n_primary"] == w * h
    orgb, opid, sel = _sampled_oracle(sc, cam, 1, 0, 256, 17)
    assert sel.sum() > 20000
    _check_sampled(a.cpu().numpy(), ap.cpu().numpy().view(np.uint32), orgb, opid, sel, "C4 soup")


def test_full_size_c4_terrain_host_entries(monkeypatch):
    """C4 terrain at 3840x2160 through every host-buffer path: streamed rows and copy-engine bands, each with the
    primitive-ID plane, and the 8-bit entry (streamed and banded) -- all against the device-resident frame."""
    torch = torch_mod()
    sc, cam = scenes.c4_terrain()
    w, h = api.image_size(cam)
    s = api.Scene(sc)
    dev = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    dpid = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    s.render_device(cam, dev.data_ptr(), dpid.data_ptr())
    ref = dev.cpu().numpy()
    rpid = dpid.cpu().numpy().view(np.uint32)
    ref8 = api.quantize_rgb8_host(ref)
    for stream in ("1", "0"):
        monkeypatch.setenv("YAHR_B200_HOST_STREAM", stream)
        monkeypatch.setenv("YAHR_B200_POISON_FRAME", "1")
        rgb, pid, st = s.render(cam)
        assert np.array_equal(rgb.view(np.uint32), ref.view(np.uint32)), "float frame, stream=" + stream
        assert np.array_equal(pid, rpid), "primid plane, stream=" + stream
        assert st["d2h_bytes"] == w * h * 16
        rgb8, st8 = s.render_rgb8(cam)
        assert np.array_equal(rgb8, ref8), "8-bit frame, stream=" + stream
        assert st8["d2h_bytes"] == w * h * 3
    monkeypatch.delenv("YAHR_B200_HOST_STREAM")
    rgb8, st8 = s.render_rgb8(cam)                      # the static rule: first call of this kind, streamed
    assert np.array_equal(rgb8, ref8) and st8["launches"] <= 3
    s.close()


def test_spp_scratch_regrows_for_a_larger_image(monkeypatch):
    """One scene handle, a small image with many samples per launch, then a larger image with fewer: the accumulator
    must grow with the image even when the per-sample scratch of the first call is still large enough."""
    sc, _ = scenes.c1_scene_yahrr()
    small = scenes._camera(200, 150, 1.5, [0.4, -0.3, 1], [0, 1, 0], [-4, 3, 2])
    large = scenes._camera(300, 200, 1.5, [0.4, -0.3, 1], [0, 1, 0], [-4, 3, 2])
    s = api.Scene(sc)
    monkeypatch.setenv("YAHR_B200_SAMPLES_PER_LAUNCH", "4")
    s.render(small, spp=4, seed=3)
    monkeypatch.setenv("YAHR_B200_SAMPLES_PER_LAUNCH", "1")
    rgb, pid, _ = s.render(large, spp=4, seed=3)
    s.close()
    fresh = api.Scene(sc)
    rgb2, pid2, _ = fresh.render(large, spp=4, seed=3)
    fresh.close()
    assert np.array_equal(rgb.view(np.uint32), rgb2.view(np.uint32)) and np.array_equal(pid, pid2)
    o = ob.OracleScene(sc)
    orgb, opid, _, _ = o.render(large, spp=4, seed=3)
    o.close()
    compare(rgb, pid, orgb, opid, 1.0, "small-then-large spp")


def test_out_buffers_are_validated():
    sc, cam = scenes.c1_scene_yahrr(64, 48)
    s = api.Scene(sc)
    with pytest.raises(ValueError):
        s.render(cam, out=(np.zeros((48, 64, 3), np.float64), None))
    with pytest.raises(ValueError):
        s.render(cam, out=(np.zeros((64, 48, 3), np.float32), None))
    with pytest.raises(ValueError):
        s.render(cam, out=(np.zeros((48, 128, 3), np.float32)[:, ::2], None))
    with pytest.raises(ValueError):
        s.render_rgb8(cam, out=np.zeros((48, 64, 3), np.float32))
    s.close()


def test_empty_scene_renders_black():
    sc = scenes._empty_scene()
    cam = scenes._camera(64, 48, 1.0, [0, 0, 1], [0, 1, 0], [0, 0, 0])
    s = api.Scene(sc)
    rgb, pid, st = s.render(cam)
    assert (rgb == 0).all() and (pid == api.PRIM_MISS).all()
    assert st["n_shadow"] == 0
    s.close()


def test_ragged_image_sizes():
    """Image sizes that do not divide into equal tiles, 1-pixel-wide images, tiny images."""
    sc, _ = scenes.c1_scene_yahrr()
    o = ob.OracleScene(sc)
    for (w, h) in [(1, 1), (7, 5), (33, 1), (1, 40), (250, 130), (517, 389)]:
        cam = scenes._camera(w, h, 1.5, [0.4, -0.3, 1], [0, 1, 0], [-4, 3, 2])
        orgb, opid, _, _ = o.render(cam)
        s = api.Scene(sc)
        rgb, pid, _ = s.render(cam)
        s.close()
        compare(rgb, pid, orgb, opid, 1.0, "c1 %dx%d" % (w, h))
    o.close()


def test_invalid_arguments_are_errors():
    sc, cam = scenes.c1_scene_yahrr(64, 64)
    s = api.Scene(sc)
    with pytest.raises(api.YahrError):
        s.render(cam, spp=0)
    with pytest.raises(api.YahrError):
        s.render(cam, recursion_depth=99)
    bad = dict(cam)
    bad["imW"] = 0
    with pytest.raises(api.YahrError):
        s.render(bad)
    s.close()
