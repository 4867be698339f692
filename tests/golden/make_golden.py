#!/usr/bin/env python
"""Generates the golden `.yahrr` fixtures of tests/golden/ by IMPORTING the reference's own Python
scene writer (compat/yahr.py) from /root/reference -- it cannot travel to the GPU box, so its outputs
are committed here together with this script.

  python tests/golden/make_golden.py          # needs /root/reference

Fixtures:
  testscene.yahr      exactly the scene compat/test_yahr.py:10-60 writes (the reference's own CLI test)
  mixed.yahr          spheres + smooth/flat mesh + nested, negative numbers, escaped material ids
  *.expected.json     what `read` + `objects >>= expand` must produce (computed here with numpy float32,
                      following Scene.hs:61-86), used to check the C++ reader
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/compat")
from yahr import (BVH, BlinnPhongMaterial, Camera, PointLight, Scene, TriangleMesh, Vec3,  # noqa: E402
                  WhittedIntegrator)

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32


def f32(x):
    """What `read :: Float` makes of the text '%f' % x."""
    return float(F(float("%f" % x)))


def v3(v):
    return [f32(v.x1), f32(v.x2), f32(v.x3)]


def flat_normal(p0, p1, p2):
    p0, p1, p2 = (np.asarray(p, F) for p in (p0, p1, p2))
    a, b = p2 - p0, p1 - p0
    c = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F)
    d = F((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2])
    with np.errstate(all="ignore"):
        return [float(x) for x in (F(1.0) / np.sqrt(d)) * c]


def expected(scene):
    mats = {}
    for i, m in enumerate(scene.materials):
        mats[m.id] = i                       # Map.fromList: last duplicate wins
    prims = []
    for o in scene.objects:
        pts = [v3(p) for p in o.triangleMeshPoints]
        nrm = None if o.triangleMeshNormals is None else [v3(n) for n in o.triangleMeshNormals]
        smooth = o.triangleMeshSmooth
        tris = o.triangleMeshTriangles
        if smooth is not None:
            tris = tris[:len(smooth)]        # zipWith
        for k, (i0, i1, i2) in enumerate(tris):
            n = flat_normal(pts[i0], pts[i1], pts[i2])
            if nrm is not None and smooth is not None and smooth[k]:
                ns = [nrm[i0], nrm[i1], nrm[i2]]
            else:
                ns = [n, n, n]
            prims.append(dict(kind="triangle", p=[pts[i0], pts[i1], pts[i2]], n=ns, material=mats[o.materialId]))
    return dict(
        recursionDepth=scene.integrator.recursionDepth, bvhMaxDepth=scene.cullingMode.maxDepth,
        splitMode=scene.cullingMode.splitMode,
        camera=dict(imW=float(scene.camera.imW), imH=float(scene.camera.imH),
                    focalLength=float(int(scene.camera.focalLength)), lookDir=v3(scene.camera.lookDir),
                    upDir=v3(scene.camera.upDir), position=v3(scene.camera.position)),
        materials=[dict(id=m.id, diffuse=v3(m.diffuse), specular=v3(m.specular), shininess=f32(m.shininess))
                   for m in scene.materials],
        lights=[v3(l.position) + v3(l.spectrum) for l in scene.lights],
        prims=prims)


def test_yahr_scene():
    """compat/test_yahr.py:10-60, verbatim arguments."""
    return Scene(
        WhittedIntegrator(3), BVH(16),
        Camera(imW=10, imH=10, focalLength=5, lookDir=Vec3(0, -0.1, -1), upDir=Vec3(0, 1, 0), position=Vec3(0, 0.2, 1)),
        [BlinnPhongMaterial(id="testmat", ambient=Vec3(0, 0, 0), diffuse=Vec3(0.8, 0.8, 0.8), specular=Vec3(0, 0, 0),
                            shininess=1)],
        [PointLight(Vec3(4, 10, 10), Vec3(1, 1, 1))],
        [TriangleMesh([Vec3(-1000, -2.2, -1000), Vec3(1000, -2.2, -1000), Vec3(1000, -2.2, 1000),
                       Vec3(-1000, -2.2, 1000)],
                      [Vec3(0, 1, 0), Vec3(0, 1, 0), Vec3(0, 1, 0), Vec3(0, 1, 0)],
                      [(0, 1, 2), (0, 2, 3)], [True, False], "testmat")])


def mixed_scene():
    rng = np.random.default_rng(7)
    pts = [Vec3(*[float(x) for x in rng.uniform(-3, 3, 3)]) for _ in range(12)]
    nrm = [Vec3(*[float(x) for x in rng.normal(size=3)]) for _ in range(12)]
    tris = [tuple(int(x) for x in rng.choice(12, 3, replace=False)) for _ in range(9)]
    return Scene(
        WhittedIntegrator(2), BVH(7, 'SurfaceAreaHeuristic'),
        Camera(imW=64, imH=48, focalLength=2, lookDir=Vec3(0.1, -0.2, 1), upDir=Vec3(0, 1, 0), position=Vec3(0, 1, -9)),
        [BlinnPhongMaterial(id='a "quoted" \\ id', ambient=Vec3(0.1, 0, 0), diffuse=Vec3(0.5, 0.25, 0.125),
                            specular=Vec3(1, 1, 1), shininess=12.5),
         BlinnPhongMaterial(id="dup", ambient=Vec3(0, 0, 0), diffuse=Vec3(1, 0, 0), specular=Vec3(0, 0, 0), shininess=1),
         BlinnPhongMaterial(id="dup", ambient=Vec3(0, 0, 0), diffuse=Vec3(0, 1, 0), specular=Vec3(0, 0, 0), shininess=2)],
        [PointLight(Vec3(-4, 10, -10), Vec3(50, 40, 30)), PointLight(Vec3(5, 5, -5), Vec3(10, 10, 10))],
        [TriangleMesh(pts, nrm, tris, [bool(k % 2) for k in range(7)], 'a "quoted" \\ id'),    # smooth list shorter
         TriangleMesh(pts[:6], None, [(0, 1, 2), (3, 4, 5)], None, "dup")])


for name, sc in (("testscene", test_yahr_scene()), ("mixed", mixed_scene())):
    open(os.path.join(HERE, name + ".yahr"), "w").write(sc.repr())
    json.dump(expected(sc), open(os.path.join(HERE, name + ".expected.json"), "w"), indent=1)
    print("wrote", name)
