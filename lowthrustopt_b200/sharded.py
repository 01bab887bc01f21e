"""Multi-GPU form of the hot path: one process per GPU (torch.distributed), independent units
(segments, or whole trajectories) sharded across the ranks, and ONE kind of collective --
an all-gather of each rank's output slab (defects, Jacobian / STM blocks, status words) so that
the rank running the host-side Newton step holds the full set (BASELINE.json north_star;
SURVEY.md 8(e)).  There is no halo and no reduction: segment i depends only on its own node data
(multiShoot_CRTBP_direct.jl:82-101, multiShoot_CRTBP_indirect.jl:75-82).

Work assignment.  The unit range is cut into `n_chunks` chunks of world*cs units; inside chunk c
rank r owns units [(c*world + r)*cs, (c*world + r + 1)*cs).  Consequences:
  * the all-gather of chunk c lands directly in the final, globally ordered arrays
    (no re-ordering pass, no second copy);
  * chunk c's all-gather (NCCL stream) overlaps chunk c+1's propagation kernel (library stream);
  * every rank takes an interleaved sample of the batch, which evens out the spread of adaptive
    step counts along a trajectory (4..67 accepted steps per segment).
The tail is padded up to a whole chunk; padded rows are never computed and are sliced off.

torch is plumbing here (device buffers, streams, the process group); the propagation itself is
liblto_b200's kernels enqueued through lto_*_dev.  On CPU (gloo) the same machinery runs with a
caller-supplied `compute` -- that is how tests/ cover the N > 1 logic without GPUs.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import capi


class ShardPlan:
    def __init__(self, n_units, world, n_chunks=2):
        if n_units < 0 or world < 1 or n_chunks < 1:
            raise ValueError("bad shard plan")
        self.n_units, self.world = int(n_units), int(world)
        n_chunks = max(1, min(int(n_chunks), -(-self.n_units // self.world) or 1))
        self.cs = max(1, -(-self.n_units // (self.world * n_chunks)))                   # units per rank per chunk
        self.n_chunks = max(1, -(-self.n_units // (self.world * self.cs)))
        self.padded = self.n_chunks * self.world * self.cs

    def local(self, rank, c):
        """(first unit, number of real units) of rank `rank` in chunk c."""
        u0 = (c * self.world + rank) * self.cs
        return u0, max(0, min(self.cs, self.n_units - u0))

    def owner(self, unit):
        c, rem = divmod(int(unit), self.world * self.cs)
        return c, rem // self.cs


class ShardedRunner:
    """spec: name -> (per-unit shape tuple, torch dtype).  run() returns {name: tensor (n_units, *shape)} on every rank."""

    def __init__(self, spec, device, group=None, n_chunks=2):
        self.spec, self.device, self.group = dict(spec), torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_chunks = n_chunks
        self.cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._bufs = {}
        self.timing = None          # set to a list to collect (start, end) CUDA event pairs around every compute call

    def _buffers(self, plan):
        key = (plan.padded, plan.cs)
        if self._bufs.get("key") != key:
            full = {k: torch.zeros((plan.padded,) + tuple(s), dtype=dt, device=self.device) for k, (s, dt) in self.spec.items()}
            loc = {k: [torch.zeros((plan.cs,) + tuple(s), dtype=dt, device=self.device) for _ in range(plan.n_chunks)]
                   for k, (s, dt) in self.spec.items()}
            self._bufs = {"key": key, "full": full, "loc": loc}
        return self._bufs["full"], self._bufs["loc"]

    def run(self, n_units, compute, compute_stream=None, sync=True):
        """compute(unit0, count, outs) fills outs[name][:count] for units [unit0, unit0+count) -- on `compute_stream`
        (a torch.cuda stream object wrapping the library's stream) when on CUDA, synchronously on CPU.
        sync=False (CUDA): return without waiting on the host; the results are ordered on `self.comm_stream`
        (world > 1) / `compute_stream` (world == 1) -- call finish() before reading them from another stream."""
        plan = ShardPlan(n_units, self.world, self.n_chunks)
        full, loc = self._buffers(plan)
        works = []
        if self.cuda and self.world > 1 and compute_stream is not None:
            compute_stream.wait_stream(self.comm_stream)       # the previous pass's gathers still read the local chunk buffers
        for c in range(plan.n_chunks):
            u0, cnt = plan.local(self.rank, c)
            outs = {k: loc[k][c] for k in self.spec}
            if cnt > 0:
                if self.cuda and self.timing is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(compute_stream); compute(u0, cnt, outs); e1.record(compute_stream)
                    self.timing.append((e0, e1, cnt))
                else:
                    compute(u0, cnt, outs)
            if self.world == 1:
                for k in self.spec:
                    if self.cuda:
                        with torch.cuda.stream(compute_stream):
                            full[k][c * plan.cs:(c + 1) * plan.cs].copy_(outs[k], non_blocking=True)
                    else:
                        full[k][c * plan.cs:(c + 1) * plan.cs].copy_(outs[k])
                continue
            g0 = c * self.world * plan.cs
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record(compute_stream)
                self.comm_stream.wait_event(ev)
                with torch.cuda.stream(self.comm_stream):
                    for k in self.spec:
                        works.append(dist.all_gather_into_tensor(full[k][g0:g0 + self.world * plan.cs], outs[k], group=self.group, async_op=True))
            else:
                for k in self.spec:
                    dist.all_gather_into_tensor(full[k][g0:g0 + self.world * plan.cs], outs[k], group=self.group)
        if self.cuda:
            with torch.cuda.stream(self.comm_stream):
                for w in works:
                    w.wait()
            self._last_stream = compute_stream
            if sync:
                self.finish()
        return {k: v[:n_units] for k, v in full.items()}, plan

    def finish(self):
        if self.cuda:
            self.comm_stream.synchronize()
            if getattr(self, "_last_stream", None) is not None:
                self._last_stream.synchronize()


def _bcast(t, src, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


class ShardedIndirect:
    """Indirect method, trajectory form, across the ranks of `group`: the multi-GPU counterpart of
    lto_indirect_defect[_jac]_traj.  Unit = one trajectory (its n_nodes-1 segments), so each rank's slab is a
    contiguous block of whole trajectories of the solver's arrays."""

    def __init__(self, handle, n_traj, n_nodes, ndim, device, group=None, n_chunks=2, compute=None):
        self.h, self.n_traj, self.n_nodes, self.nd = handle, int(n_traj), int(n_nodes), int(ndim)
        self.device, self.group = torch.device(device), group
        spu = n_nodes - 1
        f64, i32 = torch.float64, torch.int32
        self.spec_def = {"defect": ((spu, ndim), f64), "status": ((spu,), i32), "nsteps": ((spu, 2), i32)}
        self.spec_jac = dict(self.spec_def, phi=((spu, ndim, ndim), f64))
        self.spec_ls = {"sumsq": ((), f64), "bad": ((), i32)}
        self.run_def = ShardedRunner(self.spec_def, device, group, n_chunks)
        self.run_jac = ShardedRunner(self.spec_jac, device, group, n_chunks)
        self.run_ls = ShardedRunner(self.spec_ls, device, group, 1)
        self._ls_scratch = None
        self._compute = compute
        self.XC = torch.empty((n_traj, n_nodes, ndim), dtype=f64, device=self.device)
        self.t = torch.empty((n_traj, n_nodes), dtype=f64, device=self.device)
        self.tl = torch.empty(n_traj, dtype=f64, device=self.device)
        self.rho = torch.empty(n_traj, dtype=f64, device=self.device)
        self.stream = None
        if self.device.type == "cuda" and handle is not None:
            self.stream = torch.cuda.ExternalStream(handle.stream, device=self.device)

    def load(self, XC_all=None, t_TU=None, thrustLimit=None, rho=None, src=0):
        """The solver rank (`src`) supplies host arrays; every rank ends up with the inputs on its device."""
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        for dst, a in ((self.XC, XC_all), (self.t, t_TU), (self.tl, thrustLimit), (self.rho, rho)):
            if rank == src:
                a = np.asarray(a, dtype=np.float64)
                if a.shape != tuple(dst.shape):
                    a = np.broadcast_to(a, tuple(dst.shape)).copy()                 # scalars allowed for tl / rho
                a = np.ascontiguousarray(a)
                dst.copy_(torch.from_numpy(a), non_blocking=True)
            _bcast(dst, src, self.group)

    def _gpu_compute(self, params, jac):
        h, nn, nd = self.h, self.n_nodes, self.nd

        def compute(u0, cnt, outs):
            h.indirect_dev(params, cnt * (nn - 1), nn, nd, self.XC[u0].data_ptr(), self.t[u0].data_ptr(), None, None,
                           self.tl[u0:].data_ptr(), self.rho[u0:].data_ptr(), outs["defect"].data_ptr(), outs["status"].data_ptr(),
                           outs["nsteps"].data_ptr(), outs["phi"].data_ptr() if jac else None)
        return compute

    def _gpu_compute_sumsq(self, params):
        """Line-search form: propagate, then reduce sum(defect^2) per trajectory on the device (lto_sumsq_dev);
        only one double (+ a status flag) per trajectory is gathered."""
        h, nn, nd = self.h, self.n_nodes, self.nd
        spu = nn - 1

        def compute(u0, cnt, outs):
            if self._ls_scratch is None or self._ls_scratch[0].shape[0] < cnt:
                self._ls_scratch = (torch.empty((cnt, spu, nd), dtype=torch.float64, device=self.device),
                                    torch.empty((cnt, spu), dtype=torch.int32, device=self.device))
            d, st = self._ls_scratch
            h.indirect_dev(params, cnt * spu, nn, nd, self.XC[u0].data_ptr(), self.t[u0].data_ptr(), None, None,
                           self.tl[u0:].data_ptr(), self.rho[u0:].data_ptr(), d.data_ptr(), st.data_ptr(), None, None)
            h.sumsq_dev(d.data_ptr(), cnt, spu * nd, outs["sumsq"].data_ptr())
            with torch.cuda.stream(self.stream):
                outs["bad"][:cnt].copy_(st[:cnt].amax(dim=1))
        return compute

    def run(self, params, jac=True, mode=None, sync=True):
        """One pass over all trajectories.  mode "jac" / "defect": returns {defect, status, nsteps[, phi]} as full,
        globally ordered tensors (n_traj, n_nodes-1, ...) on every rank's device; mode "sumsq": {sumsq, bad} (n_traj,)."""
        mode = mode or ("jac" if jac else "defect")
        runner = {"jac": self.run_jac, "defect": self.run_def, "sumsq": self.run_ls}[mode]
        if self._compute is not None:
            compute = lambda u0, cnt, outs: self._compute(self, params, mode if mode == "sumsq" else (mode == "jac"), u0, cnt, outs)   # noqa: E731
        else:
            if self.h is None:
                raise capi.LtoError("ShardedIndirect needs a liblto_b200 handle: there is no CPU propagation path")
            compute = self._gpu_compute_sumsq(params) if mode == "sumsq" else self._gpu_compute(params, mode == "jac")
            if self.stream is not None:                       # inputs were produced on torch's current stream
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
        out, plan = runner.run(self.n_traj, compute, self.stream, sync=sync)
        return out, plan

    def finish(self):
        for r in (self.run_jac, self.run_def, self.run_ls):
            r.finish()


class ShardedDirect:
    """Direct method, pairs form (BASELINE config 3 shape) across ranks: unit = one segment."""

    def __init__(self, handle, n_seg, nstate, device, group=None, n_chunks=2, nsteps=10, compute=None):
        self.h, self.n_seg, self.ns, self.nsteps = handle, int(n_seg), int(nstate), int(nsteps)
        self.device, self.group = torch.device(device), group
        f64, i32 = torch.float64, torch.int32
        nv = 2 * (nstate + 3)
        self.spec_def = {"defect": ((nstate,), f64), "errors": ((), f64), "status": ((), i32)}
        self.spec_jac = dict(self.spec_def, jac=((nv, nstate), f64))
        self.run_def = ShardedRunner(self.spec_def, device, group, n_chunks)
        self.run_jac = ShardedRunner(self.spec_jac, device, group, n_chunks)
        self._compute = compute
        self.inp = {k: torch.empty((n_seg,) + s, dtype=f64, device=self.device)
                    for k, s in (("Xa", (nstate,)), ("Xb", (nstate,)), ("ua", (3,)), ("ub", (3,)), ("ta", ()), ("tb", ()))}
        self.stream = None
        if self.device.type == "cuda" and handle is not None:
            self.stream = torch.cuda.ExternalStream(handle.stream, device=self.device)

    def load(self, batch=None, src=0):
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        for k, dst in self.inp.items():
            if rank == src:
                dst.copy_(torch.from_numpy(np.ascontiguousarray(batch[k], dtype=np.float64)).reshape(dst.shape), non_blocking=True)
            _bcast(dst, src, self.group)

    def run(self, params, jac=True):
        runner = self.run_jac if jac else self.run_def
        if self._compute is not None:
            compute = lambda u0, cnt, outs: self._compute(self, params, jac, u0, cnt, outs)   # noqa: E731
        else:
            if self.h is None:
                raise capi.LtoError("ShardedDirect needs a liblto_b200 handle: there is no CPU propagation path")
            h, ns, i = self.h, self.ns, self.inp

            def compute(u0, cnt, outs):
                h.direct_dev(params, cnt, 0, ns, self.nsteps, i["Xa"][u0:].data_ptr(), i["Xb"][u0:].data_ptr(), i["ua"][u0:].data_ptr(),
                             i["ub"][u0:].data_ptr(), i["ta"][u0:].data_ptr(), i["tb"][u0:].data_ptr(), outs["defect"].data_ptr(),
                             outs["errors"].data_ptr(), outs["status"].data_ptr(), outs["jac"].data_ptr() if jac else None)
            if self.stream is not None:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
        out, plan = runner.run(self.n_seg, compute, self.stream)
        return out, plan


# --------------------------------------------------------------------------- fused compute + gather over NVLink peer memory
class _DevArray:
    """Raw device memory seen as a torch tensor (no copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2, "strides": None}


_TYPESTR = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}


class PeerGather:
    """The solver rank (`root`) owns the full output arrays; every other rank maps them over NVLink (CUDA IPC) and the
    propagation kernels are launched with those mapped addresses as their OUTPUT pointers, so each rank's slab is stored
    straight into the solver rank's HBM by the kernel's own epilogue -- no collective, no SM taken from the propagation,
    the transfer overlaps the computation segment by segment.  Completion is a stream-ordered 64-bit flag per rank in the
    solver rank's memory (lto_signal_dev / lto_wait_dev): nothing in the steady state touches the host.

    Units are split into contiguous, equal slabs (rank r: [n*r/W, n*(r+1)/W)).  spec: name -> (per-unit shape, dtype).

    Pass separation: a rank that starts pass k+1 overwrites its slab of the solver rank's arrays, and nothing here tells it that
    the solver rank has finished READING pass k.  Callers must put a collective (or any root -> ranks synchronisation) between the
    solver rank's use of one pass and the next `run` -- in this package that is the broadcast of the next iterate in `load()`."""

    def __init__(self, handle, spec, n_units, device, group=None, root=0):
        self.h, self.spec, self.n_units, self.device, self.group, self.root = handle, dict(spec), int(n_units), torch.device(device), group, root
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.unit_bytes = {k: int(np.prod(s, dtype=np.int64)) * torch.empty((), dtype=dt).element_size() for k, (s, dt) in self.spec.items()}
        self.seq = 0
        self._owned, self._mapped = {}, {}
        if self.rank == root:
            for k in self.spec:
                self._owned[k] = handle.dev_alloc(max(256, self.n_units * self.unit_bytes[k]))
            self._owned["__flags"] = handle.dev_alloc(max(256, 8 * self.world))          # one 64-bit completion flag per rank
            self.flags_t = torch.as_tensor(_DevArray(self._owned["__flags"], (max(32, self.world),), "<i8"), device=self.device)
            self.flags_t.zero_()
            torch.cuda.synchronize(self.device)
            payload = [{k: handle.ipc_export(p) for k, p in self._owned.items()}]
        else:
            payload = [None]
        if self.world > 1:
            dist.broadcast_object_list(payload, src=root, group=group)
        if self.rank == root:
            self.ptr = dict(self._owned)
        else:
            self._mapped = {k: handle.ipc_open(hd) for k, hd in payload[0].items()}
            self.ptr = dict(self._mapped)

    def slab(self, rank=None):
        r = self.rank if rank is None else rank
        return self.n_units * r // self.world, self.n_units * (r + 1) // self.world

    def out_ptr(self, name, unit0):
        return self.ptr[name] + unit0 * self.unit_bytes[name]

    def signal(self):
        """Enqueue on the library stream: "everything this rank enqueued so far has been written"."""
        self.h.signal_dev(self.ptr["__flags"] + 8 * self.rank, self.seq)

    def begin(self):
        self.seq += 1

    def wait_all(self):
        """Solver rank: hold the library stream until every rank's slab of pass `seq` has landed."""
        if self.rank == self.root:
            for r in range(self.world):
                self.h.wait_dev(self.ptr["__flags"] + 8 * r, self.seq)

    def tensors(self):
        """Solver rank: the full arrays as torch tensors (views of the gathered memory)."""
        assert self.rank == self.root
        return {k: torch.as_tensor(_DevArray(self._owned[k], (self.n_units,) + tuple(s), _TYPESTR[dt]), device=self.device)
                for k, (s, dt) in self.spec.items()}

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)
        for p in self._mapped.values():
            self.h.ipc_close(p)
        self._mapped = {}
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)
        for p in self._owned.values():
            self.h.dev_free(p)
        self._owned = {}


class PeerIndirect(ShardedIndirect):
    """ShardedIndirect with the results delivered by PeerGather instead of an all-gather: one kernel launch per rank and
    pass over the rank's contiguous block of trajectories, output pointers in the solver rank's memory."""

    def __init__(self, handle, n_traj, n_nodes, ndim, device, group=None, root=0):
        super().__init__(handle, n_traj, n_nodes, ndim, device, group, n_chunks=1)
        self.root = root
        self.pg = PeerGather(handle, self.spec_jac, n_traj, device, group, root)
        self.pg_ls = None

    def run(self, params, jac=True, wait=True, mode=None):
        """Enqueue one pass.  With wait=True the solver rank's library stream is held until all slabs have landed, so work
        enqueued afterwards on that stream (or after lto_sync) sees the complete arrays.  Returns them on the solver rank.
        mode "sumsq": the line-search form, one sum(defect^2) (+ a status flag) per trajectory."""
        nn, nd = self.n_nodes, self.nd
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if mode == "sumsq":
            if self.pg_ls is None:
                self.pg_ls = PeerGather(self.h, self.spec_ls, self.n_traj, self.device, self.group, self.root)
                self._bad_view = torch.as_tensor(_DevArray(self.pg_ls.ptr["bad"], (self.n_traj,), "<i4"), device=self.device)
            pg = self.pg_ls
            u0, u1 = pg.slab()
            pg.begin()
            if u1 > u0:
                cnt, spu = u1 - u0, nn - 1
                if self._ls_scratch is None or self._ls_scratch[0].shape[0] < cnt:
                    self._ls_scratch = (torch.empty((cnt, spu, nd), dtype=torch.float64, device=self.device),
                                        torch.empty((cnt, spu), dtype=torch.int32, device=self.device))
                d, st = self._ls_scratch
                self.h.indirect_dev(params, cnt * spu, nn, nd, self.XC[u0].data_ptr(), self.t[u0].data_ptr(), None, None,
                                    self.tl[u0:].data_ptr(), self.rho[u0:].data_ptr(), d.data_ptr(), st.data_ptr(), None, None)
                self.h.sumsq_dev(d.data_ptr(), cnt, spu * nd, pg.out_ptr("sumsq", u0))
                with torch.cuda.stream(self.stream):
                    self._bad_view[u0:u1].copy_(st[:cnt].amax(dim=1))
            pg.signal()
            if wait:
                pg.wait_all()
            return pg.tensors() if pg.rank == pg.root else None
        pg = self.pg
        u0, u1 = pg.slab()
        pg.begin()
        if u1 > u0:
            self.h.indirect_dev(params, (u1 - u0) * (nn - 1), nn, nd, self.XC[u0].data_ptr(), self.t[u0].data_ptr(), None, None,
                                self.tl[u0:].data_ptr(), self.rho[u0:].data_ptr(), pg.out_ptr("defect", u0), pg.out_ptr("status", u0),
                                pg.out_ptr("nsteps", u0), pg.out_ptr("phi", u0) if jac else None)
        pg.signal()
        if wait:
            pg.wait_all()
        return pg.tensors() if pg.rank == pg.root else None


class PeerDirect(ShardedDirect):
    def __init__(self, handle, n_seg, nstate, device, group=None, nsteps=10, root=0):
        super().__init__(handle, n_seg, nstate, device, group, n_chunks=1, nsteps=nsteps)
        self.pg = PeerGather(handle, self.spec_jac, n_seg, device, group, root)

    def run(self, params, jac=True, wait=True):
        pg, ns, i = self.pg, self.ns, self.inp
        u0, u1 = pg.slab()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        pg.begin()
        if u1 > u0:
            self.h.direct_dev(params, u1 - u0, 0, ns, self.nsteps, i["Xa"][u0:].data_ptr(), i["Xb"][u0:].data_ptr(), i["ua"][u0:].data_ptr(),
                              i["ub"][u0:].data_ptr(), i["ta"][u0:].data_ptr(), i["tb"][u0:].data_ptr(), pg.out_ptr("defect", u0),
                              pg.out_ptr("errors", u0), pg.out_ptr("status", u0), pg.out_ptr("jac", u0) if jac else None)
        pg.signal()
        if wait:
            pg.wait_all()
        return pg.tensors() if pg.rank == pg.root else None
