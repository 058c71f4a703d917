"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/rpt_b200.h declares, the host-side entry points work without a GPU, and error
behaviour is status + message (never an abort)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rpt_b200 import _capi as capi
from rpt_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.lib()
    header = open(os.path.join(ROOT, "include", "rpt_b200.h")).read()
    declared = set(re.findall(r"\b(rptb_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in capi.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)


def test_struct_layouts_match_header_sizes(tmp_path):
    """The ctypes mirrors against the header itself: a C program compiled from include/rpt_b200.h
    prints sizeof() of every struct that crosses the boundary."""
    import subprocess
    names = {"rptb_material": capi.Material, "rptb_kdnode": capi.KdNode, "rptb_mesh": capi.Mesh,
             "rptb_object": capi.Object, "rptb_group": capi.Group, "rptb_light": capi.Light, "rptb_env": capi.Env,
             "rptb_scene_desc": capi.SceneDesc, "rptb_camera": capi.Camera, "rptb_render_params": capi.RenderParams,
             "rptb_stats": capi.Stats, "rptb_kdtree_out": capi.KdTreeOut, "rptb_obj_group": capi.ObjGroup,
             "rptb_obj_groups_out": capi.ObjGroupsOut}
    src = tmp_path / "sizes.c"
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    src.write_text('#include <stdio.h>\n#include "rpt_b200.h"\nint main(void){' + body + 'return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for n, t in names.items():
        assert int(got[n]) == C.sizeof(t), n
    # and the sizes the docs quote
    assert C.sizeof(capi.Material) == 64 and C.sizeof(capi.KdNode) == 32 and C.sizeof(capi.Camera) == 96
    assert C.sizeof(capi.RenderParams) == 64 and C.sizeof(capi.Stats) == 88
    assert C.sizeof(capi.Object) == 16 + 128 + 32 + 16


def test_build_kdtree_runs_on_host_and_matches_oracle(orc):
    tris = scenes.teapot_triangles()
    mesh = api.Mesh(tris)
    nodes, refs, depth, max_leaf = orc.build_kdtree(tris)
    got = np.frombuffer(C.string_at(mesh.nodes, C.sizeof(capi.KdNode) * len(mesh.nodes)), dtype=orc.KDNODE_DTYPE)
    for k in ("split", "kind", "left", "right", "first_ref", "num_refs"):
        np.testing.assert_array_equal(got[k], nodes[k])
    np.testing.assert_array_equal(mesh.refs, refs)
    assert (mesh.depth, mesh.max_leaf) == (depth, max_leaf)


def test_build_kdtree_small_and_degenerate_inputs(orc):
    # < 16 triangles -> a single leaf (src/kdtree.rs:236)
    tris = np.stack([api.Triangle.from_vertices([i, 0, 0], [i + 1, 0, 0], [i, 1, 0]) for i in range(5)])
    m = api.Mesh(tris)
    assert len(m.nodes) == 1 and m.nodes[0].kind == 3 and list(m.refs) == [0, 1, 2, 3, 4]
    # many identical triangles: every split scores n >= 0.85 n -> leaf, no infinite recursion
    same = np.tile(api.Triangle.from_vertices([0, 0, 0], [1, 0, 0], [0, 1, 0]), (64, 1))
    m = api.Mesh(same)
    on, orf, _, _ = orc.build_kdtree(same)
    assert len(m.nodes) == len(on) == 1 and len(m.refs) == 64
    # random soup, larger than the parallel-task threshold is exercised elsewhere; here 3000 tris
    rng = np.random.default_rng(1)
    c = rng.uniform(-1, 1, (3000, 1, 3))
    v = c + rng.normal(0, 0.05, (3000, 3, 3))
    soup = np.stack([api.Triangle.from_vertices(*t) for t in v])
    m = api.Mesh(soup)
    on, orf, d, ml = orc.build_kdtree(soup)
    assert len(m.nodes) == len(on) and (m.refs == orf).all() and m.depth == d


def test_errors_are_statuses_with_messages():
    lib = capi.lib()
    out = capi.KdTreeOut()
    rc = lib.rptb_build_kdtree(None, 0, C.byref(out))
    assert rc == -1 and b"triangle" in lib.rptb_last_error()
    rc = lib.rptb_scene_create(None, 0, None)
    assert rc == -1
    flat = api.FlatScene(scenes.sphere_scene().scene)
    h = C.c_void_p()
    flat.objects[0].material = 99
    rc = lib.rptb_scene_create(C.byref(flat.desc), 0, C.byref(h))
    assert rc == -1 and b"material" in lib.rptb_last_error() and not h
    with pytest.raises(capi.RptbError):
        capi.check(rc, "rptb_scene_create")


def test_no_gpu_means_no_device_error_not_a_fallback():
    """On a box without a GPU the product path must fail loudly (no CPU fallback)."""
    lib = capi.lib()
    if lib.rptb_device_count() > 0:
        pytest.skip("a GPU is present")
    flat = api.FlatScene(scenes.sphere_scene().scene)
    h = C.c_void_p()
    rc = lib.rptb_scene_create(C.byref(flat.desc), 0, C.byref(h))
    assert rc == -3, lib.rptb_last_error()
    r = api.Renderer(scenes.sphere_scene().scene, api.Camera.default()).width(8).height(8)
    with pytest.raises(capi.RptbError):
        r.render()


def test_tile_pixel_is_the_ownership_map_of_sharded_renders():
    """rptb_tile_pixel: element j of the k-th tile owned by shard s of n <-> pixel.  Over all shards it is a bijection
    onto the image (ragged edges included) and agrees with rpt_b200.distributed.tile_owner."""
    from rpt_b200.distributed import tile_owner

    lib = capi.lib()
    for (w, h, n) in ((37, 21, 1), (37, 21, 3), (64, 16, 8), (5, 3, 2)):
        own = tile_owner(w, h, n)
        seen = np.full(w * h, -1, np.int64)
        ntiles = ((w + 15) // 16) * ((h + 7) // 8)
        for s in range(n):
            mine = (ntiles - s + n - 1) // n if ntiles > s else 0
            for k in range(mine):
                for j in range(128):
                    p = lib.rptb_tile_pixel(w, h, s, n, k, j)
                    if p >= 0:
                        assert seen[p] == -1
                        seen[p] = s
            assert lib.rptb_tile_pixel(w, h, s, n, mine, 0) == -1      # past this shard's last tile
        assert (seen == own.ravel()).all()
    assert lib.rptb_tile_pixel(16, 8, 0, 1, 0, 128) == -1 and lib.rptb_tile_pixel(0, 8, 0, 1, 0, 0) == -1
