"""Sweep sharding: the frequency / imaginary-frequency / Bloch-vector points of a sweep are independent
units — every point needs only the immutable geometry and the frequency-independent FIPPI table, both replicated on
every GPU — so point j goes to rank j mod world with NO data-path collective; only the per-point scalar results are
gathered on rank 0 (host objects through torch.distributed, backend nccl on GPU boxes, gloo in the CPU tests).
This mirrors how the reference's applications loop over frequencies (scuff-scatter.cc:373-472,
scuff-cas3D/CasimirIntegrand.cc:296-312); one process per GPU replaces the reference's single process."""
import torch.distributed as dist


def points_for_rank(points, rank, world):
    """round-robin: the (global index, point) pairs rank `rank` evaluates"""
    return [(j, p) for j, p in enumerate(points) if j % world == rank]


def gather_results(local, npoints, rank, world):
    """local: list of (global index, result) from this rank -> on rank 0 the list of results in point order
    (None on other ranks)."""
    if world == 1:
        parts = [local]
    else:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(local, parts, dst=0)
        if rank != 0:
            return None
    out = [None] * npoints
    for part in parts:
        for j, r in part:
            if out[j] is not None:
                raise RuntimeError("sweep point %d evaluated twice" % j)
            out[j] = r
    missing = [j for j, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError("sweep points not evaluated: %s" % missing[:8])
    return out


def run_sweep(points, evaluate, rank=0, world=1):
    """evaluate(point) -> picklable result (e.g. a few floats: power, log-det, trace).  Returns the ordered results on
    rank 0.  `evaluate` is where RWGGeometry.AssembleBEMMatrix(+LU) runs on this rank's GPU."""
    local = [(j, evaluate(p)) for j, p in points_for_rank(points, rank, world)]
    return gather_results(local, len(points), rank, world)


# ---- the frequency-independent FIPPI table, built once by all ranks together (third row) ------------
def allreduce_records(rec, world, device=None):
    """rec: this rank's [count, 36] float64 table with zeros where other ranks computed; summed over the ranks IN PLACE
    (the table is the union: every record is computed by exactly one rank).  NCCL needs the data on the device."""
    if world == 1 or rec.size == 0:
        return rec
    import numpy as np
    import torch
    t = torch.from_numpy(np.ascontiguousarray(rec))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    rec[...] = t.cpu().numpy()
    return rec


def share_fippi_table(ps, G, rank, world, device=None):
    """Cold start of a sweep: instead of every rank building the whole FIPPI table of G (seconds), rank r builds the records
    with index = r (mod world) and the tables are summed with one all-reduce (36 doubles per record: 90 MB at N = 20 280).
    Every rank classifies the pairs itself (milliseconds), so the record indices agree without any exchange."""
    if world == 1:
        G.PrepareNearPlans()
        return ps.cache_records()
    ps.fippi_set_partition(rank, world)
    try:
        G.PrepareNearPlans()
    finally:
        ps.fippi_set_partition(0, 1)
    rec = ps.fippi_host_records()
    counts = [None] * world
    dist.all_gather_object(counts, int(rec.shape[0]))
    if len(set(counts)) != 1:
        raise RuntimeError("FIPPI tables of the ranks differ in size (%s): the ranks do not hold the same geometry" % counts)
    allreduce_records(rec, world, device)
    ps.fippi_commit()
    return rec.shape[0]


# ---- one very large matrix across GPUs (second row) ----------------------------------------------
def column_ranges(n, world):
    """Column blocks [c_r, c_{r+1}) of an n x n SYMMETRIC matrix with equal upper-triangle area (work ~ c_{r+1}^2 - c_r^2):
    c_r = round(n sqrt(r / world))."""
    cuts = [int(round(n * (r / world) ** 0.5)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_columns(flat, n, elem_doubles, ranges, rank, world, dst=0):
    """flat: this rank's zero-copy float64 view of its n x n column-major device matrix.  Column block r (contiguous in
    column-major storage) travels from rank r straight into rank dst's matrix at the same offset (NCCL send/recv over
    NVLink); nothing is staged."""
    if world == 1:
        return
    ops = []
    if rank == dst:
        for r, (c0, c1) in enumerate(ranges):
            if r != dst and c1 > c0:
                ops.append(dist.P2POp(dist.irecv, flat[c0 * n * elem_doubles:c1 * n * elem_doubles], r))
    else:
        c0, c1 = ranges[rank]
        if c1 > c0:
            ops.append(dist.P2POp(dist.isend, flat[c0 * n * elem_doubles:c1 * n * elem_doubles], dst))
    if ops:
        # one batch: the receives from all peers progress concurrently (unbatched NCCL P2P ops are serialised)
        for q in dist.batch_isend_irecv(ops):
            q.wait()
