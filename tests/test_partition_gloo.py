"""Multi-rank host logic on CPU: world_size-2/3 `gloo` runs of the slab partition (halo lists, interior ranges, local
numbering) driving a distributed Jacobi-PCG whose local matrices are assembled by the ORACLE on each rank's slab.  The
distributed result must equal the single-domain oracle result: this is the multi-GPU contract exercised
without a GPU; the device path uses exactly the same partition description."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, nz, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import _bootstrap
    _bootstrap.load()
    import oracle as orc
    from del_fem_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h = 1.0 / (nx - 1)
    s = partition.make_slab(nx, ny, nz, rank, world)
    e2v, xyz = partition.local_mesh_host(s, h, index_dtype=np.uint64)
    nloc = xyz.shape[0]
    assert nloc == s.n_own + s.n_ghost
    r2i_all, i2c_all = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nloc, False)
    r2i = r2i_all[: s.n_own + 1].copy()
    i2c = i2c_all[: int(r2i[-1])].copy()
    # local assembly of the OWNED rows only: merge elements in ascending (= global) order, drop ghost rows
    rv_all, iv_all = orc.assemble("solid_linear_tet", 0, e2v, xyz, 1.0, 1.0, r2i_all, i2c_all)
    rv, iv = rv_all[: s.n_own].copy(), iv_all[: int(r2i[-1])].copy()
    flag = s.flags_clamp_z0()
    g = s.local_to_global()

    def halo(v):  # v: [nloc, 3]; fills the ghost tail
        reqs = []
        bufs = []
        for p, peer in enumerate(s.peers):
            sb = torch.from_numpy(np.ascontiguousarray(v[s.send_idx[int(s.send_ptr[p]):int(s.send_ptr[p + 1])].astype(np.int64)]))
            rb = torch.empty((int(s.recv_ptr[p + 1] - s.recv_ptr[p]), 3), dtype=torch.float64)
            reqs.append(dist.isend(sb, peer))
            reqs.append(dist.irecv(rb, peer))
            bufs.append((p, rb))
        for q in reqs:
            q.wait()
        for p, rb in bufs:
            v[s.n_own + int(s.recv_ptr[p]): s.n_own + int(s.recv_ptr[p + 1])] = rb.numpy()

    def spmv(y, alpha, x, rvv, ivv):  # y[:n_own] = alpha * A_local x
        halo(x)
        yy = np.zeros((s.n_own, 3))
        # oracle mult_vec on the rectangular local matrix: x has nloc blocks, rows n_own
        orc.mult_vec(yy, 0.0, alpha, x, r2i, i2c, ivv, rvv)
        y[: s.n_own] = yy

    def gdot(a, b):
        t = torch.tensor([float(np.sum(a[: s.n_own] * b[: s.n_own]))], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    u0 = np.zeros((nloc, 3))
    u0[: s.n_own] = orc.splitmix_fill(0, 3 * s.first_global_vtx, 3 * s.n_own, -0.1, 0.1).reshape(-1, 3)
    u0[: s.n_own][flag[: s.n_own] != 0] = 0.0
    rhs = np.zeros((nloc, 3))
    spmv(rhs, -1.0, u0, rv, iv)
    rhs[: s.n_own][flag[: s.n_own] != 0] = 0.0
    # BC on the local rectangular matrix needs the ghost flags for the column pass
    rvb, ivb = rv.copy(), iv.copy()
    import ctypes as C
    orc.lib().orc_set_fixed_bc(C.c_uint64(3), C.c_double(1.0), flag.ctypes.data_as(C.c_void_p), rvb.ctypes.data_as(C.c_void_p),
                               ivb.ctypes.data_as(C.c_void_p), r2i.ctypes.data_as(C.c_void_p), C.c_uint64(s.n_own),
                               i2c.ctypes.data_as(C.c_void_p), C.c_uint64(i2c.size))
    dinv = 1.0 / rvb[:, [0, 4, 8]]
    # distributed Jacobi-PCG (A.6)
    x = np.zeros((nloc, 3))
    r = rhs.copy()
    hist = [np.sqrt(gdot(r, r))]
    z = np.zeros_like(r)
    z[: s.n_own] = dinv * r[: s.n_own]
    p = z.copy()
    rho = gdot(r, z)
    q = np.zeros_like(r)
    rr0 = hist[0] ** 2
    for it in range(3000):
        spmv(q, 1.0, p, rvb, ivb)
        alpha = rho / gdot(p, q)
        r[: s.n_own] -= alpha * q[: s.n_own]
        x[: s.n_own] += alpha * p[: s.n_own]
        rr = gdot(r, r)
        hist.append(np.sqrt(rr))
        if np.sqrt(rr / rr0) < 1e-8:
            break
        z[: s.n_own] = dinv * r[: s.n_own]
        rho1 = gdot(r, z)
        p[: s.n_own] = z[: s.n_own] + (rho1 / rho) * p[: s.n_own]
        rho = rho1
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x[: s.n_own], g=g[: s.n_own], rv=rv, hist=np.array(hist),
             r2i=r2i, i2c_global=g[i2c.astype(np.int64)], iv=iv)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (5, 4, 9)), (3, (4, 4, 10))])
def test_distributed_pcg_equals_single_domain(tmp_path, orc, world, shape):
    nx, ny, nz = shape
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nx, ny, nz, str(tmp_path)), nprocs=world, join=True)
    # single-domain oracle
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_elasticity_problem
    prob = oracle_elasticity_problem(orc, nx, nz=nz) if nx == ny else None
    if prob is None:
        e2v, xyz = orc.kuhn_tet_mesh(nx, ny, nz, 1.0 / (nx - 1))
        nv = xyz.shape[0]
        r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
        rv, iv = orc.assemble("solid_linear_tet", 0, e2v, xyz, 1.0, 1.0, r2i, i2c)
        flag = np.zeros((nv, 3), dtype=np.int32)
        flag[xyz[:, 2] == 0.0] = 1
        u0 = orc.splitmix_fill(0, 0, 3 * nv, -0.1, 0.1).reshape(nv, 3)
        u0[flag != 0] = 0
        rhs = np.zeros((nv, 3))
        rv0, iv0 = rv.copy(), iv.copy()
        orc.mult_vec(rhs, 0.0, -1.0, u0, r2i, i2c, iv, rv)
        orc.set_fix_dof_to_rhs_vector(rhs, flag)
        orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
        prob = dict(nv=nv, r2i=r2i, i2c=i2c, rv=rv, iv=iv, rv0=rv0, iv0=iv0, rhs=rhs, u0=u0)
    ilu = orc.Ilu("jacobi", 3, prob["r2i"], prob["i2c"])
    ilu.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    ilu.decompose()
    r = prob["rhs"].copy()
    x, pr, p = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist = orc.preconditioned_conjugate_gradient(r, x, pr, p, 1e-8, 3000, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"], ilu)
    xs = np.zeros_like(x)
    seen = np.zeros(prob["nv"], dtype=bool)
    for rk in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{rk}.npz"))
        xs[d["g"]] = d["x"]
        assert not seen[d["g"]].any()
        seen[d["g"]] = True
        # owned rows of the local assembly are BIT-identical to the global assembly (same ascending element order)
        assert np.array_equal(d["rv"], prob["rv0"][d["g"]])
        lo, hi = int(prob["r2i"][d["g"][0]]), int(prob["r2i"][d["g"][-1] + 1])
        assert np.array_equal(d["i2c_global"], prob["i2c"][lo:hi].astype(np.int64))  # same column order, local ids mapped back
        assert np.array_equal(d["iv"], prob["iv0"][lo:hi])
        assert abs(len(d["hist"]) - len(hist)) <= 1
    assert seen.all()
    assert np.linalg.norm(xs - x) / np.linalg.norm(x) < 1e-7


def test_remote_offsets_match_the_receivers_ghost_ranges():
    """peer push: remote_off[p] on the sender must be the receiver's recv_ptr entry for the sender, and sizes must agree"""
    from del_fem_b200 import partition
    for nranks in (2, 3, 5, 8):
        slabs = [partition.make_slab(5, 4, 23, r, nranks) for r in range(nranks)]
        for s in slabs:
            for p, peer in enumerate(s.peers):
                t = slabs[peer]
                q = t.peers.index(s.rank)
                assert int(s.remote_off[p]) == int(t.recv_ptr[q])
                assert int(s.send_ptr[p + 1] - s.send_ptr[p]) == int(t.recv_ptr[q + 1] - t.recv_ptr[q])
                # and the packet is the receiver's ghost copy of exactly those global vertices
                g_send = s.local_to_global()[s.send_idx[int(s.send_ptr[p]):int(s.send_ptr[p + 1])]]
                g_recv = t.local_to_global()[t.n_own + int(t.recv_ptr[q]): t.n_own + int(t.recv_ptr[q + 1])]
                assert np.array_equal(g_send, g_recv)


def _shuffled_rcm_mesh(orc, shape, seed=0):
    from del_fem_b200 import partition
    e2v, xyz = orc.kuhn_tet_mesh(*shape, 1.0 / (shape[0] - 1))
    nv = xyz.shape[0]
    shuf = np.random.default_rng(seed).permutation(nv)
    e1, x1, _ = partition.renumber_mesh(e2v, xyz, shuf)          # destroy the structured numbering ...
    e2, x2, _ = partition.renumber_mesh(e1, x1, partition.rcm_order(e1, nv))  # ... and recover locality with RCM
    return e2, x2


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_general_partition_description(orc, nranks):
    """partition_general on an unstructured numbering: packets match the receivers' ghost ranges, interior rows read no ghost,
    and the owned rows of the local oracle assembly equal the global assembly bit for bit"""
    from del_fem_b200 import partition
    e2v, xyz = _shuffled_rcm_mesh(orc, (6, 5, 8))
    nv = xyz.shape[0]
    r2i_g, i2c_g = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv_g, iv_g = orc.assemble("solid_linear_tet", 0, e2v, xyz, 1.0, 1.0, r2i_g, i2c_g)
    parts = [partition.partition_general(e2v, nv, nranks, r) for r in range(nranks)]
    assert sum(s.n_own for s in parts) == nv
    for s in parts:
        for p, q in enumerate(s.peers):
            t = parts[q]
            k = t.peers.index(s.rank)
            assert int(s.remote_off[p]) == int(t.recv_ptr[k])
            gs = s.local2global[s.send_idx[int(s.send_ptr[p]):int(s.send_ptr[p + 1])]]
            gr = t.local2global[t.n_own + int(t.recv_ptr[k]): t.n_own + int(t.recv_ptr[k + 1])]
            assert np.array_equal(gs, gr)
        nloc = s.n_own + s.n_ghost
        r2i, i2c = orc.vtx2vtx_from_uniform_mesh(s.elem2vtx, 4, nloc, False)
        for r in range(s.int_begin, s.int_end):
            assert (i2c[r2i[r]:r2i[r + 1]] < s.n_own).all()
        rv, iv = orc.assemble("solid_linear_tet", 0, s.elem2vtx, xyz[s.local2global], 1.0, 1.0, r2i, i2c)
        g = s.local2global
        for r in (0, s.n_own // 2, s.n_own - 1):
            gr_ = int(g[r])
            assert np.array_equal(rv[r], rv_g[gr_])
            assert np.array_equal(g[i2c[r2i[r]:r2i[r + 1]].astype(np.int64)], i2c_g[r2i_g[gr_]:r2i_g[gr_ + 1]].astype(np.int64))
            assert np.array_equal(iv[r2i[r]:r2i[r + 1]], iv_g[r2i_g[gr_]:r2i_g[gr_ + 1]])
