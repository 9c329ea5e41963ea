#!/usr/bin/env python
"""Authoring-container script: concave prism meshes -> OBJ -> robovat's V-HACD binary -> URDF + hull OBJs under
robovat_b200/data/urdf/.  The outputs are committed (small text files); the GPU box only reads them.

    python tools/make_vhacd_assets.py [--vhacd /root/reference/bin/vhacd] [--meshconv /root/reference/bin/meshconv]
"""
import argparse
import os
import shutil
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robovat_b200 import mesh_io  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'robovat_b200', 'data', 'urdf')

# outlines (counter-clockwise, metres) of the concave movables, extruded to +-H
H = 0.025
OUTLINES = {
    'L': [(-0.06, -0.06), (0.06, -0.06), (0.06, -0.02), (-0.02, -0.02), (-0.02, 0.08), (-0.06, 0.08)],
    'T': [(-0.02, -0.08), (0.02, -0.08), (0.02, 0.02), (0.06, 0.02), (0.06, 0.06), (-0.06, 0.06), (-0.06, 0.02), (-0.02, 0.02)],
    'U': [(-0.06, -0.06), (0.06, -0.06), (0.06, 0.06), (0.02, 0.06), (0.02, -0.02), (-0.02, -0.02), (-0.02, 0.06), (-0.06, 0.06)],
    'plus': [(-0.02, -0.06), (0.02, -0.06), (0.02, -0.02), (0.06, -0.02), (0.06, 0.02), (0.02, 0.02), (0.02, 0.06), (-0.02, 0.06),
             (-0.02, 0.02), (-0.06, 0.02), (-0.06, -0.02), (-0.02, -0.02)],
}


def ear_clip(poly):
    """Triangulation of a simple CCW polygon: list of index triples."""
    idx = list(range(len(poly)))
    tris = []

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    def inside(p, a, b, c):
        return cross(a, b, p) > 1e-12 and cross(b, c, p) > 1e-12 and cross(c, a, p) > 1e-12

    guard = 0
    while len(idx) > 3 and guard < 1000:
        guard += 1
        for k in range(len(idx)):
            i0, i1, i2 = idx[k - 1], idx[k], idx[(k + 1) % len(idx)]
            a, b, c = poly[i0], poly[i1], poly[i2]
            if cross(a, b, c) <= 1e-12:
                continue
            if any(inside(poly[j], a, b, c) for j in idx if j not in (i0, i1, i2)):
                continue
            tris.append((i0, i1, i2))
            idx.pop(k)
            break
    tris.append(tuple(idx))
    return tris


def prism_mesh(outline, h):
    poly = [tuple(p) for p in outline]
    n = len(poly)
    verts = [(x, y, -h) for x, y in poly] + [(x, y, h) for x, y in poly]
    tris = []
    for i in range(n):
        j = (i + 1) % n
        tris += [(i, j, n + j), (i, n + j, n + i)]
    for a, b, c in ear_clip(poly):
        tris.append((n + a, n + b, n + c))     # top, CCW seen from +z
        tris.append((c, b, a))                 # bottom
    return np.array(verts, float), np.array(tris, int)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--vhacd', default='/root/reference/bin/vhacd')
    ap.add_argument('--meshconv', default='/root/reference/bin/meshconv')
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for name, outline in OUTLINES.items():
        verts, tris = prism_mesh(outline, H)
        with tempfile.TemporaryDirectory() as scratch:
            src = os.path.join(scratch, name + '.obj')
            mesh_io.write_obj(src, verts, tris)
            vol = mesh_io.compute_volume(*mesh_io.read_from_obj(src))
            urdf = mesh_io.convert_obj_to_urdf(src, OUT, vhacd_bin=args.vhacd, meshconv_bin=args.meshconv,
                                               scratch_dir=os.path.join(scratch, 'tmp'))
            shutil.copy(src, os.path.join(OUT, name, name + '.obj'))
        body = mesh_io.load_urdf(urdf)
        print('%-5s volume %.6g  hulls %d  verts %s  com %s' % (name, vol, len(body['hulls']), [len(h) for h in body['hulls']],
                                                           np.round(body['com'], 4).tolist()))


if __name__ == '__main__':
    main()
