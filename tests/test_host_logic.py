"""Host-side (numpy) helpers of the package against the oracle, no GPU: the synthetic workload generators used by bench.py's
end-to-end leg and by the multi-GPU slabs must describe exactly the problem the oracle (and the device generators) build."""
import numpy as np
import pytest


def test_numpy_kuhn_mesh_equals_oracle_mesh(orc):
    from del_fem_b200 import synthetic
    for shape in [(4, 4, 4), (5, 3, 6)]:
        e_o, x_o = orc.kuhn_tet_mesh(*shape, 0.25)
        e_n, x_n = synthetic.kuhn_tet_mesh(*shape, 0.25)
        assert np.array_equal(e_n.astype(np.uint64), e_o) and np.array_equal(x_n, x_o)
        for p in x_o[e_o[::7].astype(np.int64)]:
            assert orc.tet_volume(*p) > 0.0  # positively oriented


def test_numpy_kuhn_slab_is_a_window_of_the_global_mesh(orc):
    """planes [k_begin, k_end) of the global grid, vertex ids local to the window: what every rank's e2e leg uploads"""
    from del_fem_b200 import synthetic
    nx, ny, nz, kb, ke = 4, 3, 7, 2, 5
    e_g, x_g = orc.kuhn_tet_mesh(nx, ny, nz, 0.5)
    e_w, x_w = synthetic.kuhn_tet_mesh(nx, ny, nz, 0.5, kb, ke)
    assert np.array_equal(x_w, x_g[kb * nx * ny: ke * nx * ny])
    inside = ((e_g >= kb * nx * ny) & (e_g < ke * nx * ny)).all(axis=1)
    assert np.array_equal(e_w.astype(np.int64) + kb * nx * ny, e_g[inside].astype(np.int64))


def test_numpy_tri_grid_equals_oracle_mesh(orc):
    """bench.py's config-1 workload generator (numpy) against the oracle's serial generator: connectivity equal, coordinates bit-exact"""
    from del_fem_b200 import synthetic
    for disk in (False, True):
        t2v, xy = synthetic.grid_tri_mesh(9, disk=disk)
        t2v_o, xy_o = orc.grid_tri_mesh(9, disk=disk)
        assert np.array_equal(t2v.astype(np.uint64), t2v_o)
        assert np.array_equal(xy, xy_o)


def test_numpy_splitmix_equals_oracle_splitmix(orc):
    from del_fem_b200 import synthetic
    for seed, first, n in [(0, 0, 1000), (3, 12345, 257)]:
        a = synthetic.splitmix_uniform(seed, first, n, -0.1, 0.1)
        b = orc.splitmix_fill(seed, first, n, -0.1, 0.1)
        assert np.array_equal(a, b)
        assert a.min() >= -0.1 and a.max() <= 0.1 and abs(a.mean()) < 0.02


def test_clamp_flags_mark_the_global_plane_z0():
    from del_fem_b200 import synthetic
    f = synthetic.clamp_flags_z0(4, 3, 5)
    assert f.shape == (60, 3) and f[:12].all() and not f[12:].any()
    g = synthetic.clamp_flags_z0(4, 3, 2, k_begin=2)  # a slab that does not contain plane 0
    assert g.shape == (24, 3) and not g.any()


def test_cloth_topology_helpers():
    from del_fem_b200.cloth import edges_of_grid_cloth, hinges_of_tri_mesh, tris_of_grid_cloth
    nx, ny = 7, 5
    tris = tris_of_grid_cloth(nx, ny).astype(np.int64)
    assert tris.shape == (2 * (nx - 1) * (ny - 1), 3)
    x, y = np.meshgrid(np.arange(nx), np.arange(ny))
    p = np.stack([x.ravel(), y.ravel()], 1).astype(float)
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    assert np.allclose(area, 0.5)  # counter-clockwise, covers the grid once
    edges = edges_of_grid_cloth(nx, ny).astype(np.int64)
    tri_edges = {tuple(sorted(e)) for t in tris for e in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0]))}
    assert {tuple(sorted(e)) for e in edges} == tri_edges and len(edges) == len(tri_edges)
    hinges = hinges_of_tri_mesh(tris).astype(np.int64)
    interior = len(tri_edges) - 2 * (nx - 1) - 2 * (ny - 1)
    assert hinges.shape == (interior, 4)
    tri_set = {tuple(int(v) for v in np.roll(t, k)) for t in tris for k in range(3)}
    for v0, v1, v2, v3 in hinges:
        assert (v0, v2, v3) in tri_set and (v1, v3, v2) in tri_set  # the numbering quadratic_bending::wdw expects
