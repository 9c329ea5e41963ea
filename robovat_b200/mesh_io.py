"""On-disk formats feeding the path: OBJ meshes and the single-link URDFs of robovat's asset pipeline.

SURVEY.md section 8(f) rank 1.  Mirrors, file for file:

  robovat/utils/mesh_utils.py:12-136      read_from_obj / compute_volume / compute_surface_area / compute_centroid
  tools/convert_obj_to_urdf.py:111-344    V-HACD -> per-hull OBJ -> URDF (`process_object`)
  tools/templates/*.xml                   the three URDF text templates
  BulletPhysics.add_body -> pybullet.loadURDF (bullet_physics.py:143-186): `load_urdf` + AssetLibrary.add_asset

Quirks of the reference that are kept on purpose:
  * read_from_obj reverses the face list, keeps only the first index of `v/vt/vn` triples and silently skips
    lines that fail to parse (mesh_utils.py:31-57);
  * the centroid is the area-weighted mean of the triangle VERTICES (not the volume centroid), :113-136;
  * the URDF carries ixx = iyy = izz = 1 (convert_obj_to_urdf.py:325-330), which pybullet ignores anyway
    (inertia is recomputed from the collision shape), so the loader ignores <inertia> too;
  * `--density` is unusable (`process_object` raises when mass is None, :270-272).
"""
import os
import re
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

# tools/templates/collision_template.xml, visual_template.xml, urdf_template.xml -- same text, same format specs
COLLISION_TEMPLATE = (
    '    <collision>\n'
    '      <origin rpy="0 0 0" xyz="0 0 0"/>\n'
    '      <geometry>\n'
    '        <mesh filename="{filename:s}" scale="{scale:g} {scale:g} {scale:g}"/>\n'
    '      </geometry>\n'
    '    </collision>\n')
VISUAL_TEMPLATE = (
    '    <visual>\n'
    '      <origin rpy="0 0 0" xyz="0 0 0"/>\n'
    '      <geometry>\n'
    '        <mesh filename="{filename:s}" scale="{scale:g} {scale:g} {scale:g}"/>\n'
    '      </geometry>\n'
    '      <material name="color"/>\n'
    '    </visual>\n')
URDF_TEMPLATE = (
    '<?xml version="1.0" ?>\n'
    '\n'
    '<robot name="{body_name:s}">\n'
    '\n'
    '  <material name="color">\n'
    '    <color rgba="{rgba:s}"/>\n'
    '  </material>\n'
    '\n'
    '  <link name="base_link">\n'
    '\n'
    '    <contact>\n'
    '      <lateral_friction value="1.0"/>\n'
    '      <rolling_friction value="0.001"/>\n'
    '      <spinning_friction value="0.001"/>\n'
    '      <inertia_scaling value="1.0"/>\n'
    '    </contact>\n'
    '\n'
    '    <inertial>\n'
    '      <origin rpy="0 0 0" xyz="{cx:g} {cy:g} {cz:g}"/>\n'
    '       <mass value="{mass:g}"/>\n'
    '       <inertia ixx="{ixx:g}" ixy="{ixy:g}" ixz="{ixz:g}" iyy="{iyy:g}" iyz="{iyz:g}" izz="{izz:g}"/>\n'
    '    </inertial>\n'
    '\n'
    '{visual:s}\n'
    '{collision:s}\n'
    '  </link>\n'
    '\n'
    '</robot>\n')

MAX_HULL_VERTS = 64          # bin/vhacd --maxNumVerticesPerCH default


# ---- robovat/utils/mesh_utils.py ---------------------------------------------------------------------

def read_from_obj(filename):
    """(vertices [V,3] float64, triangles [F,3] int) with the reference's parsing rules (mesh_utils.py:12-61)."""
    vertices, triangles = [], []
    with open(filename, 'rb') as f:
        for line in f:
            line = line.decode('UTF-8').strip()
            try:
                vals = line.split()
                if vals[0] == 'v':
                    vertices.append([float(x) for x in vals[1:4]])
                elif vals[0] == 'f':
                    if vals[1].find('/') == -1:
                        vi = [int(x) - 1 for x in vals[1:]]
                    else:
                        vi = [int(val.split('/')[0]) - 1 for val in vals[1:]]
                    triangles.append(vi)
            except Exception:       # the reference logs and skips (blank lines raise IndexError on vals[0])
                pass
    triangles.reverse()
    return np.array(vertices), np.array(triangles)


def _tri(vertices, triangles):
    v = np.asarray(vertices, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64)
    return v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]


def compute_volume(vertices, triangles):
    """Sum of signed tetrahedron volumes, sign corrected (mesh_utils.py:64-88).  Sequential sum like the reference."""
    v0, v1, v2 = _tri(vertices, triangles)
    total = 0.0
    for s in np.einsum('ij,ij->i', v0, np.cross(v1, v2)) / 6.0:
        total += s
    return -total if total < 0 else total


def compute_surface_area(vertices, triangles):
    """mesh_utils.py:91-110."""
    v0, v1, v2 = _tri(vertices, triangles)
    total = 0.0
    for a in 0.5 * np.linalg.norm(np.cross(v1 - v0, v2 - v0), axis=1):
        total += a
    return total


def compute_centroid(vertices, triangles):
    """Area-weighted mean of the triangle vertices (mesh_utils.py:113-136)."""
    v0, v1, v2 = _tri(vertices, triangles)
    areas = 0.5 * np.linalg.norm(np.cross(v1 - v0, v2 - v0), axis=1)
    total_area = 0
    centroid = np.zeros((3))
    for i in range(len(areas)):
        centroid += v0[i] * areas[i]
        centroid += v1[i] * areas[i]
        centroid += v2[i] * areas[i]
        total_area += areas[i]
    return centroid / (total_area * 3)


def write_obj(filename, vertices, triangles):
    with open(filename, 'w') as f:
        for v in np.asarray(vertices, dtype=np.float64):
            f.write('v %.9g %.9g %.9g\n' % (v[0], v[1], v[2]))
        for t in np.asarray(triangles, dtype=np.int64):
            f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))


# ---- tools/convert_obj_to_urdf.py ------------------------------------------------------------------------

def urdf_text(body_name, hull_filenames, mass, centroid, scale=1.0, rgba='0.50 0.50 0.50 1.00'):
    """The URDF `process_object` writes (convert_obj_to_urdf.py:288-334), character for character."""
    visual = ''.join(VISUAL_TEMPLATE.format(filename=fn, scale=scale) for fn in hull_filenames)
    collision = ''.join(COLLISION_TEMPLATE.format(filename=fn, scale=scale) for fn in hull_filenames)
    return URDF_TEMPLATE.format(body_name=body_name, mass=mass, ixx=1, iyy=1, izz=1, ixy=0, ixz=0, iyz=0,
                                cx=centroid[0], cy=centroid[1], cz=centroid[2], visual=visual, collision=collision,
                                rgba=rgba)


def count_output_groups(wrl_path):
    """convert_obj_to_urdf.py:93-108."""
    with open(wrl_path, 'r') as f:
        return sum(1 for line in f if line.startswith('Group'))


def split_wrl_text(text):
    """Pieces of a V-HACD .wrl, one per `#VRML` header (split_wrl_file, convert_obj_to_urdf.py:131-164)."""
    data = text.splitlines()
    pieces, i = [], 0
    while i < len(data):
        piece = data[i]          # the reference writes the header line without a newline
        i += 1
        while i < len(data) and data[i][:5] != '#VRML':
            piece += data[i] + '\n'
            i += 1
        pieces.append(piece)
    return pieces


def parse_wrl_piece(text):
    """Vertices and triangles of one IndexedFaceSet piece (what bin/meshconv turns into an OBJ)."""
    m = re.search(r'point\s*\[(.*?)\]', text, re.S)
    n = re.search(r'coordIndex\s*\[(.*?)\]', text, re.S)
    if not m or not n:
        return np.zeros((0, 3)), np.zeros((0, 3), np.int64)
    pts = np.array([float(x) for x in re.split(r'[\s,]+', m.group(1).strip()) if x]).reshape(-1, 3)
    idx = [int(x) for x in re.split(r'[\s,]+', n.group(1).strip()) if x]
    faces, cur = [], []
    for k in idx:
        if k == -1:
            if len(cur) >= 3:
                faces.append(cur[:3])
            cur = []
        else:
            cur.append(k)
    return pts, np.array(faces, dtype=np.int64).reshape(-1, 3)


def convert_obj_to_urdf(input_path, output_dir=None, rgba='0.50 0.50 0.50 1.00', scale=1.0, mass=0.1, density=None,
                        vhacd_bin=None, meshconv_bin=None, scratch_dir=None):
    """`process_object` (convert_obj_to_urdf.py:211-344): V-HACD with its defaults -> one OBJ per hull -> URDF.

    The closed `vhacd` binary of the reference (bin/vhacd) does the decomposition when `vhacd_bin` points at it; it
    is always run with explicit --output/--log paths from `scratch_dir` (never from inside the reference tree).  The
    VRML pieces are converted to OBJ by `meshconv_bin` when given, else by the built-in IndexedFaceSet reader (same
    vertices and faces).  Returns the URDF path.
    """
    if mass is None:
        raise ValueError('The volume is problematic. Do not use the density.')      # reference behaviour (:270-272)
    del density
    body_name = os.path.splitext(os.path.basename(input_path))[0]
    if output_dir is None:
        output_dir = os.path.dirname(input_path)
    output_dir = os.path.join(output_dir, body_name)
    os.makedirs(output_dir, exist_ok=True)
    tmp_dir = scratch_dir or os.path.join(output_dir, 'tmp')
    os.makedirs(tmp_dir, exist_ok=True)
    wrl = os.path.join(tmp_dir, 'output.wrl')
    if vhacd_bin is None or not os.path.exists(vhacd_bin):
        raise OSError('convert_obj_to_urdf needs the V-HACD binary (reference bin/vhacd); pass vhacd_bin=...')
    subprocess.run([os.path.abspath(vhacd_bin), '--input', os.path.abspath(input_path), '--output', wrl,
                    '--log', os.path.join(tmp_dir, 'log.txt')], cwd=tmp_dir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(wrl) as f:
        pieces = split_wrl_text(f.read())
    hull_filenames = []
    for i, piece in enumerate(pieces):
        basename = '%s_vhacd_%d_of_%d' % (body_name, i, len(pieces))
        hull_filenames.append(basename + '.obj')
        piece_path = os.path.join(tmp_dir, 'tmp_vhacd_%d.wrl' % i)
        with open(piece_path, 'w') as f:
            f.write(piece)
        if meshconv_bin and os.path.exists(meshconv_bin):
            subprocess.run([os.path.abspath(meshconv_bin), piece_path, '-c', 'obj', '-o', os.path.join(output_dir, basename)],
                           cwd=tmp_dir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        else:
            pts, faces = parse_wrl_piece(piece)
            write_obj(os.path.join(output_dir, basename + '.obj'), pts, faces)
    vertices, triangles = read_from_obj(input_path)
    centroid = compute_centroid(vertices, triangles)
    urdf_path = os.path.join(output_dir, body_name + '.urdf')
    with open(urdf_path, 'w') as f:
        f.write(urdf_text(body_name, hull_filenames, mass, centroid, scale, rgba))
    if scratch_dir is None:
        for fn in os.listdir(tmp_dir):
            os.remove(os.path.join(tmp_dir, fn))
        os.rmdir(tmp_dir)
    return urdf_path


# ---- pybullet.loadURDF for the template's single-link bodies ------------------------------------------------------

def convex_hull_vertices(points, max_verts=MAX_HULL_VERTS):
    """Vertices of the convex hull of `points` (what Bullet keeps of a URDF collision mesh)."""
    pts = np.unique(np.round(np.asarray(points, dtype=np.float64), 9), axis=0)
    if len(pts) > 4:
        try:
            from scipy.spatial import ConvexHull
            pts = pts[np.sort(ConvexHull(pts).vertices)]
        except Exception:           # flat / degenerate input: keep the unique points
            pass
    if len(pts) > max_verts:
        raise ValueError('hull with %d vertices exceeds the V-HACD cap of %d' % (len(pts), max_verts))
    return pts


def load_urdf(path):
    """Parse a URDF written from tools/templates/urdf_template.xml.

    Returns a dict: name, mass, com (inertial origin), lateral/rolling/spinning friction, hulls = list of [n,3]
    vertex arrays in the LINK frame (mesh scale applied).  Like pybullet without URDF_USE_INERTIA_FROM_FILE the
    <inertia> element is ignored; the body frame the simulator uses is the inertial frame, so callers centre the
    hulls on `com` (AssetLibrary.add_asset(center_on_com=...) / `urdf_asset`).
    """
    root = ET.parse(path).getroot()
    links = root.findall('link')
    if len(links) != 1:
        raise ValueError('%s: %d links; the PushEnv movables are single-link bodies' % (path, len(links)))
    link = links[0]
    out = {'name': root.get('name'), 'mass': 0.0, 'com': np.zeros(3), 'lateral_friction': 1.0,
           'rolling_friction': 0.0, 'spinning_friction': 0.0, 'hulls': []}
    contact = link.find('contact')
    if contact is not None:
        for key in ('lateral_friction', 'rolling_friction', 'spinning_friction'):
            el = contact.find(key)
            if el is not None:
                out[key] = float(el.get('value'))
    inertial = link.find('inertial')
    if inertial is not None:
        if inertial.find('mass') is not None:
            out['mass'] = float(inertial.find('mass').get('value'))
        if inertial.find('origin') is not None:
            out['com'] = np.array([float(x) for x in inertial.find('origin').get('xyz', '0 0 0').split()])
    base = os.path.dirname(os.path.abspath(path))
    for col in link.findall('collision'):
        mesh = col.find('geometry/mesh')
        if mesh is None:
            raise NotImplementedError('%s: only <mesh> collision geometry is supported' % path)
        scale = np.array([float(x) for x in mesh.get('scale', '1 1 1').split()])
        origin = col.find('origin')
        xyz = np.array([float(x) for x in origin.get('xyz', '0 0 0').split()]) if origin is not None else np.zeros(3)
        vertices, _ = read_from_obj(os.path.join(base, mesh.get('filename')))
        out['hulls'].append(convex_hull_vertices(vertices * scale + xyz))
    if not out['hulls']:
        raise ValueError('%s: no collision geometry' % path)
    return out


def urdf_asset(path):
    """Hull list of a URDF body in its inertial (centre-of-mass) frame, ready for AssetLibrary.add_asset."""
    body = load_urdf(path)
    return [h - body['com'] for h in body['hulls']], body
