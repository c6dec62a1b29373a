"""GPU, one process per GPU (needs >= 2 devices; skipped otherwise): the row-block distributed SpMV / dot / CG
of libsmb200 (dist.cu) against the oracle's single-address-space results, on both data paths: peer memory over NVLink
(ghost entries stored into the neighbour's HBM by the product kernel, all-reduce through peer slots — halo.cuh) and the
NCCL send/recv + all-reduce fallback.
Partition contract: sparsemat_par.rs:20-35 (contiguous row blocks, every block needs the x entries its columns touch)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import n_gpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q, path):
    try:
        os.environ["SMB200_DIST_P2P"] = "1" if path == "p2p" else "0"
        os.environ.setdefault("SMB200_P2P_TIMEOUT_MS", "20000")
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist

        import cases
        import sparsemat_b200 as smb
        from oracle import oracle_py as orc
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        ctx = smb.Context(rank)
        box = [smb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, 0)
        ctx.comm_init(rank, world, box[0])

        # ---- (1) z-slab Laplacian generated on the device, halo planes exchanged over NCCL -------------------------
        for vdt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
            nx, ny, nz = 24, 20, 4 * world + 3
            n = nx * ny * nz
            a = smb.DistCRS.laplace(ctx, vdt, np.uint32, nx, ny, nz)
            assert a.info()["p2p"] == (path == "p2p"), a.info()
            d = a.dims()
            lo, hi = d["row_lo"], d["row_lo"] + d["n_local"]
            vals, cols, offs = orc.laplace(vdt, np.uint32, nx, ny, nz)
            xg = orc.uniform(vdt, 4, n)
            want = orc.mvp(vals, cols, offs, xg)
            x = a.new_vec()
            x.upload(xg[lo:hi])
            y = a.mvp(x)
            assert y.dim() == d["n_local"]
            assert np.array_equal(y.to_numpy(), want[lo:hi]), "dist laplace spmv is not bit-exact"
            # distributed dot (all-reduced) against the global sequential fold
            got = a.dot(x, y)
            ref = float(orc.dot(xg, want))
            scale = float(np.sum(np.abs(xg.astype(np.float64) * want.astype(np.float64))))
            assert abs(got - ref) <= max(tol, 0.5 * n * float(np.finfo(vdt).eps)) * scale, (got, ref)
            # distributed CG: same iteration count (+-2%) and solution as the oracle's single-process solver
            b_g = orc.mvp(vals, cols, offs, orc.uniform(vdt, 6, n))
            b = a.new_vec()
            b.upload(b_g[lo:hi])
            xs = a.new_vec()
            rtol = 1e-9 if vdt == np.float64 else 1e-4
            st = smb.ConjugateGradient(rtol, 3000, relative=True).solve_with_stats(a, b, xs)
            xo = np.zeros(n, vdt)
            so = orc.cg(n, n, vals, cols, offs, b_g, xo, tol=rtol, relative=True, iter_max=3000)
            assert st["converged"] and so["converged"], (st, so)
            assert abs(int(st["iterations"]) - so["iterations"]) <= max(2, so["iterations"] // 50), (st, so["iterations"])
            assert st["final_residual"] <= rtol * np.linalg.norm(b_g.astype(np.float64)) * 1.0000001
            assert np.allclose(xs.to_numpy(), xo[lo:hi], rtol=0, atol=1e-7 if vdt == np.float64 else 2e-3)
            # the single-reduction rearrangement (one all-reduce per iteration, cg_sr.cuh): same answer, same count +-2%
            xs.fill(0.0)
            st = smb.ConjugateGradient(rtol, 3000, relative=True, single_reduce=True).solve_with_stats(a, b, xs)
            assert st["converged"], st
            assert abs(int(st["iterations"]) - so["iterations"]) <= max(2, so["iterations"] // 50), (st, so["iterations"])
            assert np.allclose(xs.to_numpy(), xo[lo:hi], rtol=0, atol=1e-7 if vdt == np.float64 else 2e-3)
            rl = b_g[lo:hi].astype(np.float64) - a.mvp(xs).to_numpy().astype(np.float64)
            rr = torch.tensor([float(rl @ rl)], dtype=torch.float64)
            dist.all_reduce(rr)
            assert float(rr.item()) ** 0.5 <= (10 * rtol if vdt == np.float64 else 5e-3) * np.linalg.norm(b_g.astype(np.float64))

        # ---- (2) general matrix: host ghost plan, packed sends, nnz-balanced bounds ---------------------------------
        vdt, idt = np.float64, np.uint64
        n, _, vals, cols, offs = cases.ragged(31, 6000, 6000, 14, vdt, idt, empty_frac=0.1)
        bounds = smb.partition_rows_by_nnz(offs, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        o64 = offs.astype(np.int64)
        a = smb.DistCRS.from_local_block(ctx, n, bounds, vals[o64[lo]:o64[hi]], cols[o64[lo]:o64[hi]],
                                         (o64[lo:hi + 1] - o64[lo]).astype(idt))
        xg = orc.uniform(vdt, 5, n)
        want = orc.mvp(vals, cols, offs, xg)
        x = a.new_vec()
        x.upload(xg[lo:hi])
        for _ in range(3):                                               # repeated exchanges reuse the plan
            y = a.mvp(x)
            assert np.array_equal(y.to_numpy(), want[lo:hi]), "dist general spmv is not bit-exact"
        assert not a.info()["peer_timeout"]

        # ---- (3) banded short rows, uneven nnz-balanced blocks: the ring kernel with packed (non-contiguous) sends,
        #          ghost windows and windows that straddle the end of the owned part; x changes between products --------
        for vdt, idt in ((np.float32, np.uint32), (np.float64, np.uint64)):
            n, _, vals, cols, offs = cases.banded(7, 30011, 700, 9, vdt, idt)
            bounds = smb.partition_rows_by_nnz(offs, world)
            lo, hi = int(bounds[rank]), int(bounds[rank + 1])
            o64 = offs.astype(np.int64)
            a = smb.DistCRS.from_local_block(ctx, n, bounds, vals[o64[lo]:o64[hi]], cols[o64[lo]:o64[hi]],
                                             (o64[lo:hi + 1] - o64[lo]).astype(idt))
            assert a.local.plan_info()["variant_name"] == "ring", a.local.plan_info()
            x, y = a.new_vec(), a.new_vec()
            for rep in range(5):
                xg = orc.uniform(vdt, 40 + rep, n)
                want = orc.mvp(vals, cols, offs, xg)
                x.upload(xg[lo:hi])
                a.mvp(x, out=y)
                assert np.array_equal(y.to_numpy(), want[lo:hi]), f"dist banded ring spmv is not bit-exact (rep {rep})"
            inf = a.info()
            assert not inf["peer_timeout"] and (path != "p2p" or inf["products"] == 5), inf
            a.barrier()
        # ---- (4) SparseMatPar with a working mvp_par (sparsemat_par.rs:37-68): 6 blocks on `world` GPUs, replicated x, the
        #          slices of y gathered on every rank; the last block is short (and an empty one follows) ----------------------
        if path == "p2p":
            n_rows, n_cols, vals, cols, offs = cases.ragged(71, 1100, 900, 11, np.float64, np.uint32, empty_frac=0.0)
            i = np.repeat(np.arange(n_rows), np.diff(offs.astype(np.int64)))
            keep = i < 1100 - 150                                        # rows 950.. stay empty: block 4 holds 150 of 200 rows, block 5 none
            edge = np.arange(0, 800, 200) + 199                          # make sure blocks 0..3 are full (their last row exists)
            ii, jj, vv = np.append(i[keep], edge), np.append(cols[keep], np.zeros(4, np.uint32)), np.append(vals[keep], np.full(4, 0.25))
            ii, jj, vv = np.append(ii, 949), np.append(jj, 1), np.append(vv, 0.5)
            par = smb.SparseMatPar(6, 1200, np.float64, np.uint32)
            par.set(ii, jj, vv)
            ref = orc.IndexListMat(np.float64, np.uint32)
            ref.set(ii, jj, vv)
            _, _, wv, wc, wo = ref.to_crs()
            xg = orc.uniform(np.float64, 9, 900)
            y = par.mvp(smb.DenseVec.from_vec(ctx, xg))
            assert par.n_rows() == 950 and y.dim() == 950
            assert len({par.owner(b) for b in range(6)}) == min(world, 6)            # the blocks are spread over the ranks
            assert np.array_equal(y.to_numpy(), orc.mvp(wv, wc, wo, xg)), "SparseMatPar mvp_par over the ranks is not bit-exact"
        ctx.sync()
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))


@pytest.mark.parametrize("path", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_block_distributed_spmv_dot_cg(world, path):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs, have {n_gpus()}")
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, q, path)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in range(world):
            results.append(q.get(timeout=600))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert len(results) == world and all(msg == "ok" for _, msg in results), results
