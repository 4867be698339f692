"""GPU tests of the output stage and the command line: our `yahr` binary
obeys the reference's argv contract (main.hs:28-38), prints its status line (main.hs:140-141) and writes
the PNG JuicyPixels would write (main.hs:142).  The first test is the reference's own
compat/test_yahr.py scenario run against our binary."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from yahr_b200 import api, scenes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAHR = os.path.join(ROOT, "yahr_b200", "bin", "yahr")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(args, cwd):
    return subprocess.run([YAHR] + args, cwd=cwd, capture_output=True, text=True, timeout=120)


def oracle_png(text):
    sc, cam, depth = api.load_yahrr(text)
    o = ob.OracleScene(sc)
    rgb, _, _, _ = o.render(cam, recursion_depth=depth)
    o.close()
    return api.quantize_rgb8_host(rgb), rgb


def test_reference_cli_test_scenario(tmp_path):
    """compat/test_yahr.py:10-66: `$YAHR_CMD .testscene.yahr .testout.png` exits 0 and writes a PNG."""
    from PIL import Image
    text = open(os.path.join(GOLD, "testscene.yahr")).read()
    (tmp_path / ".testscene.yahr").write_text(text)
    r = run([".testscene.yahr", ".testout.png"], str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "1 threads, 1 batches, parallel sequential"
    out = tmp_path / ".testout.png"
    assert out.stat().st_size > 0
    img = np.asarray(Image.open(str(out)).convert("RGB"))
    want, _ = oracle_png(text)
    assert img.shape == (10, 10, 3)
    assert np.abs(img.astype(int) - want.astype(int)).max() <= 1     # powf ulps can flip a truncation


def test_cli_modes_rts_options_and_repo_scene(tmp_path):
    from PIL import Image
    scene = os.path.join(ROOT, "scenes", "scene.yahrr")
    text = open(scene).read()
    want, _ = oracle_png(text)                                       # recursionDepth 3, 1024x768
    for args, line in ((["-p", "eval", "+RTS", "-N4"], "4 threads, 4096 batches, parallel eval"),
                       (["--parallel-mode", "gpu"], "1 threads, 4096 batches, parallel gpu"),
                       (["+RTS", "-N8", "-RTS", "-p", "par"], "8 threads, 4096 batches, parallel par")):
        out = str(tmp_path / "o.png")
        r = run([scene, out] + args, str(tmp_path))
        assert r.returncode == 0, r.stderr
        assert r.stdout.strip() == line
        img = np.asarray(Image.open(out).convert("RGB"))
        assert img.shape == want.shape
        d = np.abs(img.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_cli_errors(tmp_path):
    scene = os.path.join(ROOT, "scenes", "scene.yahrr")
    assert run([scene, "o.png", "-p", "bogus"], str(tmp_path)).returncode != 0
    assert run(["missing.yahrr", "o.png"], str(tmp_path)).returncode != 0
    (tmp_path / "bad.yahrr").write_text("Scene { nonsense }")
    r = run(["bad.yahrr", "o.png"], str(tmp_path))
    assert r.returncode != 0 and "no parse" in r.stderr
    assert run([scene], str(tmp_path)).returncode != 0


def test_render_rgb8_equals_quantised_float_frame():
    for sc, cam in (scenes.c1_scene_yahrr(320, 200), scenes.c2_bunny_proxy(384, 216, nu=40, nv=20)):
        s = api.Scene(sc)
        rgb, _, _ = s.render(cam)
        rgb8, st = s.render_rgb8(cam)
        s.close()
        assert np.array_equal(rgb8, api.quantize_rgb8_host(rgb))
        assert st["d2h_bytes"] == rgb8.size
