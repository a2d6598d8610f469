"""Slab-decomposed P(k) across the GPUs of one box: one process per GPU,
`torch.distributed` for the three exchange steps, libgenpk_cuda.so for every
compute stage.

    particles sharded by rank (any order)
      route      bucket by owner x-slab            -> all-to-all-v of particle runs
      deposit    CIC into the local slab + one ghost plane on the high-x side
      ghost      ring shift: ghost plane -> rank+1, added into its first plane
      FFT        batched 2-D D2Z over local planes -> all-to-all transpose -> 1-D Z2Z along x
      binning    this rank's ky rows of the spectrum -> all-reduce of 3*nrbins raw sums

The choreography below only moves tensors; what a stage computes is delegated
to a `stages` object (CudaStages for the product).  The same code runs over
gloo with CPU tensors, which is how the exchange logic is tested without GPUs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api


class _DevMem:
    """Exposes a raw device allocation of the library to torch (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class CudaStages:
    """Compute stages of one rank, on its GPU, through the C ABI."""

    def __init__(self, dims: int, nranks: int, rank: int, device: torch.device, flags: int = 0, ghost_planes: int = 0):
        self.device = device
        torch.cuda.set_device(device)
        self.ctx = api.Context(dims, device.index, flags, nranks, rank, ghost_planes)
        self.ghost_planes = self.ctx.ghost_planes          # > 0: ghosts on both sides, slab-local shards need no routing
        self.ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.dims, self.nranks, self.rank = dims, nranks, rank
        self.nc = dims // 2 + 1
        self.nx = dims // nranks
        self._spec = [None, None]
        self._send = None
        self._sums = None
        self.scatter_ready = False     # every rank's transposed block is known: the y pass stores into them
        self.pull_ready = False        # the ring neighbours' grids are mapped: ghost planes are pulled over NVLink
        self._recv = None

    def close(self):
        self.ctx.close()

    def bind_stream(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    # -- particles ---------------------------------------------------------------
    def route(self, pos: torch.Tensor, mass, boxsize: float):
        n = pos.numel() // 3
        spos = torch.empty_like(pos)
        smass = torch.empty_like(mass) if mass is not None else None
        counts = torch.zeros(self.nranks, dtype=torch.int64, device=self.device)
        self.ctx.route_particles(pos.data_ptr(), mass.data_ptr() if mass is not None else 0, n, boxsize,
                                 spos.data_ptr(), smass.data_ptr() if smass is not None else 0, counts.data_ptr())
        return spos, smass, counts

    def zero(self, which=0):
        self.ctx.grid_zero(which)

    def deposit(self, pos: torch.Tensor, mass, cmass: float, boxsize: float, which=0):
        n = pos.numel() // 3
        if not n:
            return
        if pos.device.type == "cpu":
            # host shard (pinned for the overlap): uploaded in chunks on the library's copy stream while
            # earlier chunks are deposited
            self.ctx.deposit_host_ptr(pos.data_ptr(), n, mass.data_ptr() if mass is not None else 0, cmass, boxsize, which)
        else:
            self.ctx.deposit_dev(pos.data_ptr(), n, mass.data_ptr() if mass is not None else 0, cmass, boxsize, which)

    # -- ghost plane ---------------------------------------------------------------
    def ghost_plane(self, which=0, side=1) -> torch.Tensor:
        """Ghost planes on the high-x (side 1) or low-x (side 0) side of the slab."""
        ptr, nbytes = self.ctx.ghost_side_ptr(side, which)
        return torch.as_tensor(_DevMem(ptr, nbytes), device=self.device)

    def ghost_accumulate(self, recv: torch.Tensor, which=0, side=0):
        """side 0: planes from rank-1 into my first owned planes; side 1: from rank+1 into my last."""
        self.ctx.ghost_side_accumulate(side, recv.data_ptr(), which)

    def rejected(self) -> int:
        return self.ctx.take_rejected()

    def rejected_to(self, slot: torch.Tensor):
        """Stream-ordered: this rank's rejected-particle count into a device double (no host round trip)."""
        self.ctx.rejected_to(slot.data_ptr())

    # -- ghost exchange by peer loads ----------------------------------------------------
    def enable_ghost_pull(self, group=None, which=0) -> bool:
        """Exchange the CUDA IPC handles of every rank's grid and map the two ring neighbours', so that
        genpk_ghost_pull can read their ghost planes over NVLink (no send/recv, no staging buffers)."""
        if self.pull_ready:
            return True
        good = 1
        try:
            mine = torch.frombuffer(bytearray(self.ctx.ipc_export_grid(which)), dtype=torch.uint8).to(self.device)
        except Exception:
            good = 0
            mine = torch.zeros(self.ctx.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        handles = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(handles, mine, group=group)
        if good:
            try:
                prv, nxt = (self.rank - 1) % self.nranks, (self.rank + 1) % self.nranks
                self.ctx.slab_set_grid_peer(0, bytes(handles[prv].cpu().numpy().tobytes()), 0, which)
                if self.ghost_planes > 0:
                    self.ctx.slab_set_grid_peer(1, bytes(handles[nxt].cpu().numpy().tobytes()), 0, which)
            except Exception:
                good = 0
        ok = torch.tensor([good], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.pull_ready = bool(int(ok.item()))
        return self.pull_ready

    def ghost_pull(self, which=0):
        self.ctx.ghost_pull(which)

    # -- FFT -----------------------------------------------------------------------
    def fft_yz(self, which=0):
        self.ctx.slab_fft_yz(which)

    def pack(self, which=0) -> torch.Tensor:
        nd = self.ctx.slab_spectrum_bytes() // 8
        if self._send is None:
            self._send = torch.empty(nd, dtype=torch.float64, device=self.device)
        self.ctx.slab_pack(self._send.data_ptr(), which)
        return self._send

    def spectrum_buffer(self, which=0) -> torch.Tensor:
        if self._spec[which] is None:
            self._spec[which] = torch.empty(self.ctx.slab_spectrum_bytes() // 8, dtype=torch.float64, device=self.device)
        return self._spec[which]

    def fft_x(self, spec: torch.Tensor):
        self.ctx.slab_fft_x(spec.data_ptr())

    # -- binning ---------------------------------------------------------------------
    def power_partial(self, spec_a: torch.Tensor, spec_b, nrbins: int) -> torch.Tensor:
        if self._sums is None or self._sums.numel() != 3 * nrbins + 1:
            self._sums = torch.zeros(3 * nrbins + 1, dtype=torch.float64, device=self.device)   # [-1]: rejected particles
        self.ctx.slab_power_partial(spec_a.data_ptr(), spec_b.data_ptr() if spec_b is not None else 0, nrbins,
                                    self._sums.data_ptr())
        return self._sums

    # -- transpose fused into the y pass ----------------------------------------------------
    def recv_block(self) -> torch.Tensor:
        if self._recv is None:
            ptr, nbytes = self.ctx.slab_recv_buffer()
            self._recv = torch.as_tensor(_DevMem(ptr, nbytes), device=self.device)
        return self._recv

    def enable_scatter(self, group=None) -> bool:
        """Exchange the CUDA IPC handles of every rank's transposed block (one process per GPU,
        same node) so that genpk_slab_fft_yz_scatter can store into them over NVLink."""
        if self.scatter_ready:
            return True
        ok = torch.tensor([1 if self.ctx.slab_scatter_supported() else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            return False
        good = 1
        try:
            self.recv_block()
            mine = torch.frombuffer(bytearray(self.ctx.ipc_export()), dtype=torch.uint8).to(self.device)
        except Exception:                                   # no IPC export on this platform: every rank falls back
            good = 0
            mine = torch.zeros(self.ctx.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        handles = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(handles, mine, group=group)
        if good:
            try:
                for r in range(self.nranks):
                    if r == self.rank:
                        self.ctx.slab_set_peer(r, None, self.ctx.slab_recv_buffer()[0])
                    else:
                        self.ctx.slab_set_peer(r, bytes(handles[r].cpu().numpy().tobytes()))
            except Exception:                               # a peer mapping failed (IPC namespace, no peer access)
                good = 0
        ok.fill_(good)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # all ranks or none: the paths must match
        self.scatter_ready = bool(int(ok.item()))
        return self.scatter_ready

    def fft_yz_scatter(self, which=0):
        self.ctx.slab_fft_yz_scatter(which)

    def fused_xpass(self, nrbins: int) -> bool:
        return self.ctx.fused_xpass_supported(nrbins)

    def sums_buffer(self, nrbins: int) -> torch.Tensor:
        if self._sums is None or self._sums.numel() != 3 * nrbins + 1:
            self._sums = torch.zeros(3 * nrbins + 1, dtype=torch.float64, device=self.device)
        return self._sums

    def fftx_power_partial(self, spec_yz: torch.Tensor, nrbins: int) -> torch.Tensor:
        """x transform + binning of the transposed (y,z)-transformed block in one kernel."""
        if self._sums is None or self._sums.numel() != 3 * nrbins + 1:
            self._sums = torch.zeros(3 * nrbins + 1, dtype=torch.float64, device=self.device)   # [-1]: rejected particles
        self.ctx.slab_fftx_power_partial(spec_yz.data_ptr(), nrbins, self._sums.data_ptr())
        return self._sums

    def check(self):
        self.ctx.synchronize()


class SlabPipeline:
    """P(k) of a particle set sharded over the ranks of `group`."""

    def __init__(self, dims: int, stages, group=None):
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.r = dist.get_rank(group) if dist.is_initialized() else 0
        if dims % self.P:
            raise ValueError(f"grid side {dims} is not divisible by {self.P} ranks")
        self.dims, self.stages = dims, stages
        self.timings = {}
        self.profile = False         # True: CUDA-event timing of the exchange steps into self.timings (ms, summed)
        self._events = []
        # wide-ghost stages: deposit the shard as it is and only route when a rank had stragglers
        # beyond its ghosts ("local" until that happens once, then "route" for good)
        self.placement = "local" if getattr(stages, "ghost_planes", 0) > 0 and self.P > 1 else "route"

    # ---- optional timing of the exchange steps (device time on the current stream) --------------
    class _Timed:
        def __init__(self, pipe, name):
            self.pipe, self.name = pipe, name

        def __enter__(self):
            if self.pipe.profile:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()

        def __exit__(self, *exc):
            if self.pipe.profile:
                self.e1.record()
                self.pipe._events.append((self.name, self.e0, self.e1))

    def collect_timings(self):
        """Sum the recorded event pairs into self.timings (call after a synchronize)."""
        for name, e0, e1 in self._events:
            self.timings[name] = self.timings.get(name, 0.0) + e0.elapsed_time(e1)
        self._events = []
        return self.timings

    # ---- exchange steps ---------------------------------------------------------------
    def exchange_particles(self, spos, smass, counts):
        """all-to-all-v of the routed particle runs; returns this slab's particles."""
        if self.P == 1:
            return spos, smass
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        send_n = [int(v) for v in counts.cpu().tolist()]
        recv_n = [int(v) for v in recv_counts.cpu().tolist()]
        rpos = torch.empty(3 * sum(recv_n), dtype=spos.dtype, device=spos.device)
        dist.all_to_all_single(rpos, spos, [3 * v for v in recv_n], [3 * v for v in send_n], group=self.group)
        rmass = None
        if smass is not None:
            rmass = torch.empty(sum(recv_n), dtype=smass.dtype, device=smass.device)
            dist.all_to_all_single(rmass, smass, recv_n, send_n, group=self.group)
        return rpos, rmass

    def exchange_ghost(self, which=0):
        """Ring shift of the high-x ghost plane into the next rank's first plane."""
        if self.P == 1:
            return
        if getattr(self.stages, "pull_ready", False):
            # every rank has deposited (barrier), then each pulls what its neighbours' deposits left in their
            # ghost planes straight from their memory: no send/recv, only the planes that were touched
            self._barrier(self.stages.device)
            self.stages.ghost_pull(which)
            return
        nxt, prv = (self.r + 1) % self.P, (self.r - 1) % self.P
        if self.group is not None:
            nxt, prv = dist.get_global_rank(self.group, nxt), dist.get_global_rank(self.group, prv)
        up = self.stages.ghost_plane(which, 1)                     # my high ghosts belong to rank+1
        from_prev = torch.empty_like(up)
        ops = [dist.P2POp(dist.isend, up, nxt, self.group), dist.P2POp(dist.irecv, from_prev, prv, self.group)]
        from_next = None
        if getattr(self.stages, "ghost_planes", 0) > 0:            # wide slabs: low ghosts go down as well
            down = self.stages.ghost_plane(which, 0)
            from_next = torch.empty_like(down)
            ops += [dist.P2POp(dist.isend, down, prv, self.group), dist.P2POp(dist.irecv, from_next, nxt, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.stages.ghost_accumulate(from_prev, which, 0)
        if from_next is not None:
            self.stages.ghost_accumulate(from_next, which, 1)

    def transpose(self, which=0):
        """[x_local][y][kz] on every rank -> [x][y_local][kz] on every rank."""
        send = self.stages.pack(which)
        spec = self.stages.spectrum_buffer(which)
        if self.P == 1:
            spec.copy_(send)
        else:
            dist.all_to_all_single(spec, send, group=self.group)
        return spec

    # ---- stages in order -----------------------------------------------------------------
    def deposit(self, pos, mass=None, cmass=1.0, boxsize=1.0, which=0, zero=True, routed=False, defer_check=None):
        """defer_check: a device double that receives this rank's rejected-particle count instead of a
        host round trip here (the caller reduces it with its sums and looks at it once per step)."""
        if zero:
            self.stages.zero(which)
        if routed or self.P == 1:
            self.stages.deposit(pos, mass, cmass, boxsize, which)
            return
        if self.placement == "local":
            # optimistic: every particle of the shard lies in this rank's slab or its ghosts
            self.stages.deposit(pos, mass, cmass, boxsize, which)
            if defer_check is not None and hasattr(self.stages, "rejected_to"):
                self.stages.rejected_to(defer_check)
                return
            bad = torch.tensor([self.stages.rejected()], dtype=torch.int64, device=self._device_of(pos))
            dist.all_reduce(bad, group=self.group)
            if int(bad.item()) == 0:
                return
            if not zero:
                raise RuntimeError(f"{int(bad.item())} particles lie outside their rank's slab and ghost planes and the "
                                   "grid already holds earlier deposits: route the particles (placement='route')")
            self.placement = "route"                                # redo this deposit, and route from now on
            self.stages.zero(which)
        if pos.device.type == "cpu":
            pos = pos.to(self._device_of(pos), non_blocking=True)
            mass = mass.to(pos.device, non_blocking=True) if mass is not None else None
        spos, smass, counts = self.stages.route(pos, mass, boxsize)
        rpos, rmass = self.exchange_particles(spos, smass, counts)
        self.stages.deposit(rpos, rmass, cmass, boxsize, which)

    def _device_of(self, t):
        return getattr(self.stages, "device", t.device) if t.device.type == "cpu" else t.device

    def spectrum(self, which=0, x_pass=True):
        with self._Timed(self, "ghost_exchange"):
            self.exchange_ghost(which)
        self.stages.fft_yz(which)
        with self._Timed(self, "transpose"):
            spec = self.transpose(which)
        if x_pass:
            self.stages.fft_x(spec)
        return spec

    def _barrier(self, device):
        """Stream-ordered rendezvous of all ranks (a 1-element all-reduce)."""
        if getattr(self, "_flag", None) is None or self._flag.device != device:
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        dist.all_reduce(self._flag, group=self.group)

    def _reduce_finalize(self, sums, nrbins, total_mass, total_mass2):
        """All-reduce of the raw sums (+ the rejected-particle count riding in the last slot when the
        deposit deferred its check), one D2H, normalisation on the host."""
        with self._Timed(self, "allreduce_d2h"):
            if self.P > 1:
                dist.all_reduce(sums, group=self.group)
            host = sums.cpu().numpy()
        self.last_rejected = int(host[3 * nrbins]) if host.size > 3 * nrbins else 0
        return api.power_finalize(host[:3 * nrbins], nrbins, total_mass, total_mass2)

    def power(self, spec_a, spec_b, nrbins, total_mass, total_mass2):
        sums = self.stages.power_partial(spec_a, spec_b, nrbins)
        return self._reduce_finalize(sums, nrbins, total_mass, total_mass2)

    def pk(self, pos, mass=None, cmass=1.0, boxsize=1.0, total_mass=1.0, nrbins=None, routed=False):
        """The per-type step of gen-pk.cpp:208-234 on this rank's particle shard (a device tensor, or a
        pinned host tensor that is uploaded in chunks while earlier chunks are deposited)."""
        nrbins = self.dims if nrbins is None else nrbins
        fused = getattr(self.stages, "fused_xpass", None)
        dev = self._device_of(pos)
        if fused is not None and fused(nrbins) and getattr(self.stages, "scatter_ready", False) and self.P > 1:
            # the y pass stores its results straight into their owner's transposed block (peer
            # stores over NVLink): no pack, no all-to-all.  Barriers: nobody still reads its block
            # (the previous call's x pass) / every rank has finished storing.  No host round trip
            # before the sums come back: the deposit's rejected-particle count rides with them.
            can_defer = hasattr(self.stages, "rejected_to") and hasattr(self.stages, "sums_buffer")
            sums = self.stages.sums_buffer(nrbins) if can_defer else None
            optimistic = can_defer and self.placement == "local" and not routed
            self.deposit(pos, mass, cmass, boxsize, 0, True, routed, defer_check=sums[3 * nrbins:] if optimistic else None)
            pulled = getattr(self.stages, "pull_ready", False)
            with self._Timed(self, "ghost_exchange"):
                self.exchange_ghost(0)                        # (with peer pulls: starts with the barrier)
            if not pulled:
                with self._Timed(self, "barriers"):
                    self._barrier(dev)
            self.stages.fft_yz_scatter(0)
            with self._Timed(self, "barriers"):
                self._barrier(dev)
            if can_defer and not optimistic:
                sums[3 * nrbins:] = 0
            sums = self.stages.fftx_power_partial(self.stages.recv_block(), nrbins)      # (writes the first 3*nrbins slots only)
            out = self._reduce_finalize(sums, nrbins, total_mass, total_mass)
            if optimistic and self.last_rejected:
                # stragglers beyond the ghost planes: this result is incomplete -- route from now on and redo
                self.placement = "route"
                return self.pk(pos, mass, cmass, boxsize, total_mass, nrbins, routed)
            return out
        self.deposit(pos, mass, cmass, boxsize, 0, True, routed)
        if fused is not None and fused(nrbins):
            # the x transform and the binning of the transposed block share one kernel
            spec = self.spectrum(0, x_pass=False)
            sums = self.stages.fftx_power_partial(spec, nrbins)
            return self._reduce_finalize(sums, nrbins, total_mass, total_mass)
        spec = self.spectrum(0)
        return self.power(spec, None, nrbins, total_mass, total_mass)
