"""CPU, world_size = 2 over gloo: the host-side logic of the N > 1 path.

Each rank owns a contiguous row block (SparseMatPar's contract, sparsemat_par.rs:20-35), runs libsmb200's ghost
plan on its block, learns from the peers which of its rows they need (the same count all-gather + id exchange
that dist.cu performs over NCCL), ships the values, and multiplies its block in local numbering
[owned | ghosts].  The local arithmetic is done by the oracle here (no GPU in this container); what is under
test is the partition / ghost plan / exchange schedule, which is shared with the GPU path."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, out_q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist

        import cases
        import sparsemat_b200 as smb
        from oracle import oracle_py as orc
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        vdt, idt = np.float64, np.uint32
        if kind == "laplace":
            nx, ny, nz = 6, 5, 8
            n = nx * ny * nz
            bounds = smb.partition_rows(n, world, nx * ny)               # whole z-planes per rank
            vals, cols, offs = orc.laplace(vdt, idt, nx, ny, nz)
        else:
            n, _, vals, cols, offs = cases.ragged(7, 3000, 3000, 12, vdt, idt, empty_frac=0.1)
            bounds = smb.partition_rows_by_nnz(offs, world)
        o64 = offs.astype(np.int64)
        x = orc.uniform(vdt, 4, n)
        want = orc.mvp(vals, cols, offs, x)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        lv, lc = vals[o64[lo]:o64[hi]], cols[o64[lo]:o64[hi]]
        lofs = (o64[lo:hi + 1] - o64[lo]).astype(idt)
        local_cols, ghosts, per_owner = smb.ghost_plan(lc, world, rank, bounds)
        if kind == "laplace":                                             # one plane from each neighbour, nothing else
            expect = np.zeros(world, np.uint64)
            if rank > 0:
                expect[rank - 1] = nx * ny
            if rank + 1 < world:
                expect[rank + 1] = nx * ny
            assert np.array_equal(per_owner, expect), (per_owner, expect)
        # 1. all-gather the world x world matrix of receive counts
        counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.from_numpy(per_owner.astype(np.int64)))
        counts = torch.stack(counts).numpy()                              # counts[q][p] = what q receives from p
        send_count = counts[:, rank].copy()
        send_count[rank] = 0
        # 2. ghost ids go to their owners, 3. owners answer with the values
        recv_off = np.concatenate([[0], np.cumsum(per_owner.astype(np.int64))])
        reqs, wants = [], {}
        for q in range(world):
            if q == rank:
                continue
            if per_owner[q]:
                ids = torch.from_numpy(ghosts[recv_off[q]:recv_off[q + 1]].astype(np.int64))
                reqs.append(dist.isend(ids, q))
            if send_count[q]:
                wants[q] = torch.zeros(int(send_count[q]), dtype=torch.int64)
                reqs.append(dist.irecv(wants[q], q))
        for r in reqs:
            r.wait()
        x_owned = x[lo:hi]
        ghost_vals = np.zeros(ghosts.size, vdt)
        reqs, bufs = [], {}
        for q in range(world):
            if q == rank:
                continue
            if send_count[q]:
                ids = wants[q].numpy()
                assert np.all((ids >= lo) & (ids < hi))                   # peers only ask for rows this rank owns
                reqs.append(dist.isend(torch.from_numpy(x_owned[ids - lo].copy()), q))
            if per_owner[q]:
                bufs[q] = torch.zeros(int(per_owner[q]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[q], q))
        for r in reqs:
            r.wait()
        for q, b in bufs.items():
            ghost_vals[recv_off[q]:recv_off[q + 1]] = b.numpy()
        assert np.array_equal(ghost_vals, x[ghosts.astype(np.int64)])
        # 4. local product in local numbering == the rank's slice of the global product, bit for bit
        y_local = orc.mvp(lv, local_cols, lofs, np.concatenate([x_owned, ghost_vals]))
        assert np.array_equal(y_local, want[lo:hi])
        # 5. the CG scalars: local dot + all-reduce(sum) vs the global sequential fold (vector.rs:50-53)
        part = torch.tensor([float(orc.dot(x_owned, y_local))], dtype=torch.float64)
        dist.all_reduce(part)
        glob = float(orc.dot(x, want))
        scale = float(np.sum(np.abs(x * want)))
        assert abs(part.item() - glob) <= 1e-12 * scale
        # 6. bench.py's timing plumbing: max over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))


@pytest.mark.parametrize("kind", ["laplace", "ragged"])
def test_row_block_spmv_world_size_2(kind):
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    world, port = 2, _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results
