"""Asset pipeline: OBJ reader / mesh statistics / URDF text against outputs of the reference's
own code (tests/golden/mesh.json, written by oracle/gen_golden.py), URDF loading, and a V-HACD scene on the oracle."""
import json
import os

import numpy as np
import pytest

from robovat_b200 import assets, config, mesh_io
from tests import helpers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mesh.json')


@pytest.fixture(scope='module')
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_obj_reader_and_mesh_statistics_match_reference(golden, tmp_path):
    assert set(golden['objs']) >= {'L', 'T', 'U', 'plus', 'sample'}
    for name, ref in golden['objs'].items():
        path = tmp_path / (name + '.obj')
        path.write_text(ref['text'])
        v, t = mesh_io.read_from_obj(str(path))
        np.testing.assert_array_equal(v, np.array(ref['vertices']))
        np.testing.assert_array_equal(t, np.array(ref['triangles']))          # incl. the reversed face order
        assert mesh_io.compute_volume(v, t) == pytest.approx(ref['volume'], rel=1e-12, abs=1e-18)
        assert mesh_io.compute_surface_area(v, t) == pytest.approx(ref['surface_area'], rel=1e-12)
        np.testing.assert_allclose(mesh_io.compute_centroid(v, t), ref['centroid'], rtol=1e-12, atol=1e-15)


def test_sample_obj_keeps_first_index_of_slash_triples(golden):
    tri = np.array(golden['objs']['sample']['triangles'])
    # file order: f 1/1/1 3/1/1 2/1/1 | f 1//1 2//1 4//1 | f 2 3 4 | f 1 4 3 ; the reference reverses the list
    np.testing.assert_array_equal(tri, [[0, 3, 2], [1, 2, 3], [0, 1, 3], [0, 2, 1]])


def test_urdf_text_is_character_identical_to_the_reference_templates(golden):
    for case in golden['urdf']:
        text = mesh_io.urdf_text(case['body_name'], case['files'], case['mass'], case['centroid'], case['scale'], case['rgba'])
        assert text == case['text']


def test_wrl_split_matches_reference(golden):
    assert mesh_io.split_wrl_text(golden['wrl']['text']) == golden['wrl']['pieces']
    pts, faces = mesh_io.parse_wrl_piece(golden['wrl']['pieces'][1])
    np.testing.assert_array_equal(pts, [[2, 0, 0], [3, 0, 0], [2, 1, 0], [2, 0, 1]])
    np.testing.assert_array_equal(faces, [[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]])


def test_urdf_round_trip(tmp_path):
    hull = assets.box_vertices(0.03, 0.02, 0.01, (0.1, 0.0, 0.0))
    from scipy.spatial import ConvexHull
    mesh_io.write_obj(str(tmp_path / 'b_vhacd_0_of_1.obj'), hull, ConvexHull(hull).simplices)
    (tmp_path / 'b.urdf').write_text(mesh_io.urdf_text('b', ['b_vhacd_0_of_1.obj'], 0.2, [0.1, 0.0, 0.0], scale=2.0))
    body = mesh_io.load_urdf(str(tmp_path / 'b.urdf'))
    assert body['name'] == 'b' and body['mass'] == pytest.approx(0.2)
    assert (body['lateral_friction'], body['rolling_friction'], body['spinning_friction']) == (1.0, 0.001, 0.001)
    np.testing.assert_allclose(body['com'], [0.1, 0, 0])
    assert len(body['hulls']) == 1 and body['hulls'][0].shape == (8, 3)
    np.testing.assert_allclose(np.abs(body['hulls'][0] - [0.2, 0, 0]).max(axis=0), [0.06, 0.04, 0.02])     # scale applied
    hulls, _ = mesh_io.urdf_asset(str(tmp_path / 'b.urdf'))
    np.testing.assert_allclose(hulls[0].mean(axis=0), [0.1, 0, 0], atol=1e-12)       # link frame -> inertial frame


def test_committed_vhacd_assets_load_and_decompose():
    data = os.path.join(assets.DATA_DIR, 'urdf')
    for name, min_hulls in (('L', 2), ('T', 2), ('U', 3), ('plus', 3)):
        body = mesh_io.load_urdf(os.path.join(data, name, name + '.urdf'))
        assert len(body['hulls']) >= min_hulls
        assert all(4 <= len(h) <= mesh_io.MAX_HULL_VERTS for h in body['hulls'])
        v, t = mesh_io.read_from_obj(os.path.join(data, name, name + '.obj'))
        # the hulls cover the source mesh: every source vertex is within 5 mm of a hull's bounding box (V-HACD works on a
        # voxel grid, so its hulls hug the surface only approximately), and the hulls do not stick out of the mesh's box
        lo = np.array([h.min(axis=0) for h in body['hulls']]) - 0.005
        hi = np.array([h.max(axis=0) for h in body['hulls']]) + 0.005
        assert all(((p >= lo) & (p <= hi)).all(axis=1).any() for p in v)
        assert (lo.min(axis=0) >= v.min(axis=0) - 0.0101).all() and (hi.max(axis=0) <= v.max(axis=0) + 0.0101).all()
        np.testing.assert_allclose(body['com'], mesh_io.compute_centroid(v, t), rtol=1e-5, atol=1e-9)    # %g in the URDF


def test_converter_refuses_without_the_vhacd_binary(tmp_path):
    src = tmp_path / 'x.obj'
    src.write_text('v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 3 2\nf 1 2 4\nf 2 3 4\nf 1 4 3\n')
    with pytest.raises(OSError):
        mesh_io.convert_obj_to_urdf(str(src), str(tmp_path / 'out'), vhacd_bin=str(tmp_path / 'missing'))
    with pytest.raises(ValueError):
        mesh_io.convert_obj_to_urdf(str(src), str(tmp_path / 'out'), mass=None, density=100.0)    # reference :270-272


def test_vhacd_scene_drops_and_settles_on_the_oracle():
    """PushEnv scene whose movables come from the URDF files: bodies land on the table and come to rest."""
    cfg, cpu = helpers.make_oracle(8, threads=4, MOVABLE_NAME='vhacd', MIN_MOVABLE_BODIES=3, MAX_MOVABLE_BODIES=3)
    cpu.reset(seed=4)
    cpu.settle(0.1, 0.1, 500)
    cpu.settle()
    st = np.array(cpu.body_state)
    assert np.isfinite(st).all()
    z = st[2]
    assert (z > -0.01).all() and (z < 0.08).all(), z
    speed = np.linalg.norm(st[7:10], axis=0)
    assert (speed < 0.05).all(), speed
    assert int(cpu.array('error_flags' if False else 13).max()) == 0
