"""Host front-end (librt_host.so): scene TOML, OBJ, BVH order, camera — CPU only.

Checked against (a) the BVH dumps of SURVEY.md Appendix B committed under tests/golden/
and (b) oracle/scene_oracle.py, an independent numpy-float32 restatement of
src/scene.rs, src/mesh.rs and src/bvh.rs.
"""
import json
import math
import os
import tomllib

import numpy as np
import pytest

from conftest import ASSETS, ROOT


def _golden(name):
    with open(os.path.join(ROOT, "tests", "golden", f"bvh_{name}.json")) as fh:
        return json.load(fh)


def _nodes_as_golden(scene):
    out = []
    for i, n in enumerate(scene.bvh_nodes_words()):
        idx, length, axis = int(n[8]), int(n[9]), int(n[10])
        out.append([i, "L", idx, length] if length else [i, "N", idx, axis])
    return out


@pytest.mark.parametrize("name", ["house", "default"])
def test_bvh_matches_survey_golden(rt, name):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    gold = _golden(name)
    assert _nodes_as_golden(scene) == gold["nodes"]
    assert scene.primitives_words().tolist() == gold["primitives"]
    root = scene.bvh_nodes_words()[0].view(np.float32)
    assert root[0:3].tolist() == pytest.approx(gold["root_bounds"][0], abs=1e-6)
    assert root[4:7].tolist() == pytest.approx(gold["root_bounds"][1], abs=1e-6)


def test_scene_counts(house, default_scene, suzanne):
    # SURVEY.md §4 table.
    assert house.counts() == {"materials": 8, "spheres": 2, "planes": 2, "vertices": 51,
                              "normals": 16, "triangles": 52, "primitives": 56, "bvh_nodes": 39}
    assert house.bvh_depth == 6
    c = default_scene.counts()
    assert (c["spheres"], c["planes"], c["vertices"], c["triangles"], c["primitives"],
            c["bvh_nodes"]) == (10, 1, 8, 12, 23, 15)
    assert default_scene.bvh_depth == 5
    c = suzanne.counts()
    assert (c["vertices"], c["triangles"], c["primitives"], c["bvh_nodes"]) == (507, 968, 968, 545)
    assert suzanne.bvh_depth == 12
    assert suzanne.primitives_words()[:8].tolist() == [[2, 845], [2, 779], [2, 297], [2, 781],
                                                       [2, 299], [2, 351], [2, 831], [2, 279]]


@pytest.mark.parametrize("name", ["house", "default", "suzanne"])
def test_bvh_bit_exact_vs_numpy_restatement(rt, scene_oracle, name):
    path = os.path.join(ASSETS, "scenes", f"{name}.toml")
    scene = rt.Scene.load_toml(path)
    ref = scene_oracle.load_scene_toml(path)
    nodes, ordered = scene_oracle.build_bvh(ref)

    assert scene.primitives_words().tolist() == [list(p) for p in ordered]
    words = scene.bvh_nodes_words()
    assert len(words) == len(nodes)
    for w, (mn, mx, idx, length, axis) in zip(words, nodes):
        f = w.view(np.float32)
        assert f[0:3].tobytes() == mn.tobytes() and f[4:7].tobytes() == mx.tobytes()
        assert (int(w[8]), int(w[9]), int(w[10])) == (idx, length, axis)

    # Packed mesh arrays (mesh.rs:92-113) and triangles (7 x u32).
    assert scene.vertices().tobytes() == ref.vertices.tobytes()
    assert scene.normals().tobytes() == ref.normals.tobytes()
    assert scene.triangles_words().tolist() == [list(t) for t in ref.triangles]


def test_planes_and_materials_vs_restatement(rt, scene_oracle):
    path = os.path.join(ASSETS, "scenes", "house.toml")
    scene = rt.Scene.load_toml(path)
    ref = scene_oracle.load_scene_toml(path)
    planes = scene.planes_words()
    for row, (pos, fwd, right, mat) in zip(planes, ref.planes):
        normal, cols = scene_oracle.plane_to_uniform(pos, fwd, right)
        assert row[0:3].tobytes() == pos.tobytes()
        assert row[4:7].tobytes() == normal.tobytes()
        for k in range(3):
            assert row[8 + 4 * k: 11 + 4 * k].tobytes() == cols[k].tobytes()
        assert int(row[20:21].view(np.uint32)[0]) == mat
    # Ground plane: normal = cross((0,0,60),(60,0,0)) normalised = +y; M maps the far corner to (1, ., 1).
    assert planes[0][4:7].tolist() == [0.0, 1.0, 0.0]
    mats = scene.materials_words()
    for row, (color, rough, metal, emission) in zip(mats, ref.materials):
        assert row[0:3].tobytes() == color.tobytes()
        assert row[3] == rough and row[4] == metal
        assert row[8:11].tobytes() == emission.tobytes()
    spheres = scene.spheres_words()
    for row, (pos, r, mat) in zip(spheres, ref.spheres):
        assert row[0:3].tobytes() == pos.tobytes() and row[3] == r
        assert int(row[4:5].view(np.uint32)[0]) == mat


def test_toml_reader_agrees_with_tomllib(rt, tmp_path):
    # The same document through the C++ TOML subset reader and through tomllib.
    text = """
# comment
[[material]]
name = "a b"   # trailing comment
color = [ 1, 0.5,
          0.25 ]
roughness = 1
metallic = 0.0
emission = [0,0,0,]

[[material]]
name = 'lit'
color = [1e-1, +2.5E-1, 3_0.0]
roughness = 0.5
metallic = 1
emission = [ 1, 2, 3 ]

[[object]]
[object.Sphere]
material = "lit"
pos = [ 1, 2, 3 ]
radius = 0.5
unknown_field = "ignored like serde does"

[[object]]
Plane = { material = "a b", pos = [0,0,0], forward = [0,0,2], right = [2,0,0] }

[[object]]
[object.Sphere]
material = "a b"
pos = [ -1, 0.25, -3 ]
radius = 1

[camera]
pos = [ 0.0, 1.0, 3.0 ]
pitch = -10
yaw = 45.0
fov_y = 90
"""
    p = tmp_path / "scene.toml"
    p.write_text(text)
    scene = rt.Scene.load_toml(str(p))
    doc = tomllib.loads(text)
    mats = scene.materials_words()
    for row, m in zip(mats, doc["material"]):
        assert row[0:3].tolist() == [np.float32(x) for x in m["color"]]
        assert row[3] == np.float32(m["roughness"]) and row[4] == np.float32(m["metallic"])
        assert row[8:11].tolist() == [np.float32(x) for x in m["emission"]]
    assert scene.counts()["spheres"] == 2 and scene.counts()["planes"] == 1
    sph = scene.spheres_words()
    assert sph[0][0:4].tolist() == [1.0, 2.0, 3.0, 0.5]
    assert int(sph[0][4:5].view(np.uint32)[0]) == 1  # material "lit"
    k = np.float32(math.pi / 180.0)
    assert np.float32(scene.camera.yaw) == np.float32(45.0) * k
    assert np.float32(scene.camera.pitch) == np.float32(-10.0) * k
    assert np.float32(scene.camera.fov_y) == np.float32(90.0) * k


def test_scene_error_messages(rt, tmp_path):
    # scene.rs:236-249, 334-351, 413-427
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_toml(str(tmp_path / "missing.toml"))
    assert str(e.value).startswith(f"Couldn't open scene {tmp_path / 'missing.toml'}:\n  ")
    assert "No such file or directory (os error 2)" in str(e.value)

    bad = tmp_path / "bad.toml"
    bad.write_text("[[material]\nname = 3")
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_toml(str(bad))
    assert str(e.value).startswith(f"Couldn't parse scene {bad}:\n  ")

    base = ('[[material]]\nname = "m"\ncolor=[1,1,1]\nroughness=1\nmetallic=0\nemission=[0,0,0]\n'
            '[camera]\npos=[0,0,0]\nyaw=0\npitch=0\nfov_y=90\n')
    nomat = tmp_path / "nomat.toml"
    nomat.write_text(base + '[[object]]\n[object.Sphere]\nmaterial="m"\npos=[0,0,0]\nradius=1\n'
                            '[[object]]\n[object.Plane]\nmaterial="nope"\npos=[0,0,0]\n'
                            'forward=[0,0,1]\nright=[1,0,0]\n')
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_toml(str(nomat))
    assert str(e.value) == f"Error in object 1 (Plane): Material 'nope' does not exist.\n  --> {nomat}"

    nomesh = tmp_path / "nomesh.toml"
    nomesh.write_text(base + '[[object]]\n[object.Mesh]\nmaterial="m"\npath="gone.obj"\n')
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_toml(str(nomesh))
    assert str(e.value).startswith("Error in object 0 (Mesh): Cannot open 'gone.obj': ")
    assert str(e.value).endswith(f"\n  --> {nomesh}")

    missing_field = tmp_path / "field.toml"
    missing_field.write_text(base.replace("roughness=1\n", "") +
                             '[[object]]\n[object.Sphere]\nmaterial="m"\npos=[0,0,0]\nradius=1\n')
    with pytest.raises(rt.RtError) as e:
        rt.Scene.load_toml(str(missing_field))
    assert "missing field `roughness`" in str(e.value)


def test_obj_fan_triangulation_and_normals(rt):
    # wavefront_obj: polygon c0..c(n-1) -> Triangle(c(n-1), c(i), c(i+1)); a plain
    # triangle a b c -> (c, a, b). Indices are global and 0-based after loading.
    obj = """
o quad
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 2 2
vn 0 0 1
vn 0 1 0
vt 0.5 0.5
f 1/1/1 2/1/1 3/1/1 4/1/1
f 1//2 2//2 5//2
f -5//1 -4//1 -1//1
l 1 2
"""
    scene = (rt.SceneBuilder().add_material((1, 1, 1), 1, 0)
             .add_mesh_obj(obj, 0).set_camera((0, 0, 3), 0, 0, 1.0).build())
    tris = scene.triangles_words().tolist()
    assert tris == [
        [3, 0, 1, 0, 0, 0, 0],   # quad: (d, a, b)
        [3, 1, 2, 0, 0, 0, 0],   #       (d, b, c)
        [4, 0, 1, 1, 1, 1, 0],   # triangle a b c -> (c, a, b)
        [4, 0, 1, 0, 0, 0, 0],   # negative (relative) indices
    ]
    with pytest.raises(rt.RtError) as e:
        rt.SceneBuilder().add_material((1, 1, 1), 1, 0).add_mesh_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", 0)
    assert "Object must include baked normals" in str(e.value)


def test_mesh_packing_offsets(rt):
    # pack_meshes (mesh.rs:92-113): later meshes are rebased by the running vertex/normal counts.
    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n"
    scene = (rt.SceneBuilder().add_material((1, 1, 1), 1, 0).add_material((0, 1, 0), 1, 0)
             .add_mesh_obj(tri, 0).add_mesh_obj(tri.replace("0 0 0", "5 5 5"), 1)
             .set_camera((0, 0, 3), 0, 0, 1.0).build())
    assert scene.triangles_words().tolist() == [[2, 0, 1, 0, 0, 0, 0], [5, 3, 4, 1, 1, 1, 1]]
    assert scene.vertices()[3].tolist() == [5.0, 5.0, 5.0]
    # Two triangles -> a single leaf holding both, in spheres/planes/triangles order.
    assert scene.primitives_words().tolist() == [[2, 0], [2, 1]]
    n = scene.bvh_nodes_words()
    assert len(n) == 1 and int(n[0][9]) == 2


def test_degenerate_centroid_leaf(rt):
    # bvh.rs:238-244: > 5 primitives with identical centroids collapse into one leaf.
    b = rt.SceneBuilder().add_material((1, 1, 1), 1, 0)
    for r in (0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 2.0):  # exact in binary -> equal centroids
        b.add_sphere((1, 2, 3), r, 0)
    scene = b.set_camera((0, 0, 9), 0, 0, 1.0).build()
    n = scene.bvh_nodes_words()
    assert len(n) == 1 and int(n[0][9]) == 7 and scene.bvh_depth == 0


def test_camera_uniform_and_state_string(rt, scene_oracle):
    cam = rt.Camera((0.5, 1.0, -2.0), 0.7, -0.3, 1.2)
    u = cam.uniform()
    cols = scene_oracle.rot_transform(cam.yaw, cam.pitch)
    got = np.array([[u.rot[c][r] for r in range(3)] for c in range(3)], dtype=np.float32)
    assert np.allclose(got, np.array(cols), atol=2e-7)  # libm sin/cos: last-ulp slack
    assert list(u.pos) == [0.5, 1.0, -2.0] and u.fov_y == np.float32(1.2)
    # Identity for the shipped scenes (yaw = pitch = 0) must be exact.
    ident = rt.Camera((0, 1, 3), 0.0, 0.0, 1.0).uniform()
    assert [[ident.rot[c][r] for r in range(3)] for c in range(3)] == [[1, 0, 0], [0, 1, 0], [0, 0, 1]]

    s = cam.serialize()  # camera.rs:30-49
    import base64
    import struct
    assert base64.b64decode(s) == struct.pack("<6f", 0.5, 1.0, -2.0, 0.7, -0.3, 1.2)
    back = rt.Camera.deserialize(s)
    assert (np.float32(back.yaw), np.float32(back.pitch), np.float32(back.fov_y)) == (
        np.float32(0.7), np.float32(-0.3), np.float32(1.2))
    with pytest.raises(rt.RtError) as e:
        rt.Camera.deserialize(base64.b64encode(b"\0" * 20).decode())
    assert str(e.value) == "Couldn't deserialize camera: binary data (20 bytes) not 24 bytes"


def test_shipped_scene_cameras(house, default_scene):
    for s in (house, default_scene):
        assert s.camera.pos == (0.0, 1.0, 3.0)
        assert s.camera.yaw == 0.0 and s.camera.pitch == 0.0
        assert np.float32(s.camera.fov_y) == np.float32(100.0) * np.float32(math.pi / 180.0)


def test_parallel_bvh_build_is_byte_identical(rt, monkeypatch):
    # Subtrees of >= 65536 primitives are built on their own threads and spliced in
    # pre-order (host/bvh.cpp); the arrays must be the serial build's byte for byte.
    dumps = {}
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("RT_BVH_THREADS", threads)
        scene = rt.Scene.synthetic_field(ASSETS, 14)  # 200 k triangles: three levels fork
        assert scene.counts()["triangles"] > 150000
        dumps[threads] = (scene.bvh_nodes_words().tobytes(), scene.primitives_words().tobytes(), scene.bvh_depth)
    assert dumps["1"] == dumps["3"] == dumps["8"]


def test_bvh_can_be_built_elsewhere_and_installed(rt):
    """rt_scene_synthetic_field_ex(build_bvh = 0) + rt_scene_bvh_input + rt_scene_set_bvh /
    rt_scene_bvh_storage + commit: the seams the device builder (rt_bvh_build) uses, on the CPU."""
    import ctypes as C
    lib = rt._ffi.host_lib()
    built = rt.Scene.synthetic_field(ASSETS, 2)
    bare = rt.Scene.synthetic_field(ASSETS, 2, build_bvh=False)
    assert bare.counts()["bvh_nodes"] == 0 and bare.counts()["primitives"] == 0
    assert bare.counts()["triangles"] == built.counts()["triangles"] == 4080
    inp = bare.bvh_input()
    assert (inp.sphere_count, inp.plane_count, inp.triangle_count) == (0, 1, 4080)
    pb = bare.plane_bounds()          # Plane::bounds = box of pos and pos + forward + right (scene.rs:203-207)
    assert pb.shape == (1, 6) and np.all(pb[0, :3] <= pb[0, 3:]) and pb[0, 1] == pb[0, 4] == 0.0
    nodes, prims = built.bvh_nodes_words(), built.primitives_words()
    # (a) copy in
    st = lib.rt_scene_set_bvh(bare._handle, nodes.ctypes.data, len(nodes), prims.ctypes.data, len(prims), built.bvh_depth)
    assert st == 0
    bare.refresh()
    assert bare.bvh_nodes_words().tobytes() == nodes.tobytes() and bare.bvh_depth == built.bvh_depth
    assert lib.rt_scene_set_bvh(bare._handle, nodes.ctypes.data, len(nodes), prims.ctypes.data, len(prims) - 1, 3) != 0
    # (b) write straight into the scene's own arrays
    other = rt.Scene.synthetic_field(ASSETS, 2, build_bvh=False)
    pn, pp = C.c_void_p(), C.c_void_p()
    assert lib.rt_scene_bvh_storage(other._handle, len(nodes), C.byref(pn), C.byref(pp)) == 0
    C.memmove(pn, nodes.ctypes.data, nodes.nbytes)
    C.memmove(pp, prims.ctypes.data, prims.nbytes)
    assert lib.rt_scene_bvh_commit(other._handle, built.bvh_depth) == 0
    other.refresh()
    assert other.primitives_words().tobytes() == prims.tobytes() and other.counts()["bvh_nodes"] == len(nodes)
    # (c) and the host builder on request
    other.build_bvh_host()
    assert other.bvh_nodes_words().tobytes() == nodes.tobytes()
