#!/usr/bin/env python
"""OBJ -> V-HACD -> URDF, the asset step of robovat (reference tools/convert_obj_to_urdf.py:35-90 CLI).

    python tools/convert_obj_to_urdf.py --input mesh.obj --output out_dir --vhacd /path/to/vhacd [--meshconv ...]

`--input` may be a file or a directory of .obj files.  The V-HACD binary is the one robovat ships (bin/vhacd);
it is run from a scratch directory with explicit output paths.
"""
import argparse
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robovat_b200 import mesh_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--input', required=True, help='input .obj file or directory')
    ap.add_argument('--output', default=None, help='output directory (default: next to the input)')
    ap.add_argument('--rgba', default='0.50 0.50 0.50 1.00')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--mass', type=float, default=0.1)
    ap.add_argument('--density', type=float, default=None)
    ap.add_argument('--vhacd', default=os.environ.get('VHACD_BIN', '/root/reference/bin/vhacd'))
    ap.add_argument('--meshconv', default=os.environ.get('MESHCONV_BIN'))
    args = ap.parse_args()
    paths = sorted(glob.glob(os.path.join(args.input, '*.obj'))) if os.path.isdir(args.input) else [args.input]
    for path in paths:
        urdf = mesh_io.convert_obj_to_urdf(path, args.output, rgba=args.rgba, scale=args.scale, mass=args.mass,
                                           density=args.density, vhacd_bin=args.vhacd, meshconv_bin=args.meshconv)
        print(urdf)


if __name__ == '__main__':
    main()
