"""Oracle-backed stand-in for genpk_b200.distributed.CudaStages on CPU tensors.

TEST INFRASTRUCTURE: lets the exchange choreography of SlabPipeline (all-to-all-v
routing, ghost ring shift, transpose, all-reduce) run over gloo without GPUs.
Every stage is computed with numpy / the CPU oracle."""
import numpy as np
import torch

from oracle.oracle import Oracle, padded_shape


class NumpyStages:
    def __init__(self, dims, nranks, rank, ghost_planes=0, fused=False, scatter=False):
        """fused: offer the x pass + binning as one stage (CudaStages.fftx_power_partial);
        scatter: also offer the y pass that delivers its rows to their owners (here an
        all-to-all inside the stage stands in for the peer stores)."""
        self.dims, self.nranks, self.rank = dims, nranks, rank
        if fused:
            self.fused_xpass = lambda nrbins: True
        self.scatter_ready = bool(scatter)
        self._recv = None
        self.ghost_planes = ghost_planes if nranks > 1 else 0     # > 0: ghosts on both sides (wide slabs)
        self.glo = self.ghost_planes
        self.ghi = (self.ghost_planes or 1) if nranks > 1 else 1
        self._rejected = 0
        self.nc, self.fd = dims // 2 + 1, 2 * (dims // 2 + 1)
        self.nx = dims // nranks
        self.x0 = rank * self.nx
        self.orc = Oracle("port")
        self.grid = None
        self.spec2d = None

    def _cell_x(self, pos, boxsize):
        x = pos.reshape(-1, 3)[:, 0].astype(np.float64) * (self.dims / boxsize)
        return np.mod(np.floor(x).astype(np.int64), self.dims)

    def route(self, pos, mass, boxsize):
        p = pos.numpy().reshape(-1, 3)
        dest = self._cell_x(p, boxsize) // self.nx
        order = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=self.nranks).astype(np.int64)
        spos = torch.from_numpy(np.ascontiguousarray(p[order]).reshape(-1))
        smass = torch.from_numpy(np.ascontiguousarray(mass.numpy()[order])) if mass is not None else None
        return spos, smass, torch.from_numpy(counts)

    def zero(self, which=0):
        self.grid = np.zeros((self.glo + self.nx + self.ghi, self.dims, self.fd))

    def deposit(self, pos, mass, cmass, boxsize, which=0):
        p = pos.numpy().reshape(-1, 3)
        if len(p) == 0:
            return
        m = None if mass is None else mass.numpy()
        # local plane of the low-x corner: periodic image of X - x0 nearest to the slab
        d = self._cell_x(p, boxsize) - self.x0
        d = np.where(d < -self.glo, d + self.dims, np.where(d >= self.dims - self.glo, d - self.dims, d))
        ok = (d >= -self.glo) & (d <= self.nx + self.ghi - 2) if self.nranks > 1 else np.ones(len(p), bool)
        self._rejected += int((~ok).sum())
        p, m = p[ok], (None if m is None else m[ok])
        if len(p) == 0:
            return
        full = np.zeros(padded_shape(self.dims))
        self.orc.fieldize(boxsize, self.dims, full, np.ascontiguousarray(p), None if m is None else np.ascontiguousarray(m), cmass, 1)
        if self.nranks == 1:            # the periodic wrap already landed in plane 0
            self.grid[: self.nx] += full
            return
        # every plane the kept particles can touch, [x0 - glo, x0 + nx + ghi), is distinct modulo dims
        # as long as glo + nx + ghi <= dims; the caller guarantees it (ghost_planes <= nx, P >= 2 ... see test)
        planes = (np.arange(-self.glo, self.nx + self.ghi) + self.x0) % self.dims
        assert len(set(planes.tolist())) == len(planes)
        self.grid += full[planes]

    def rejected(self):
        r, self._rejected = self._rejected, 0
        return r

    def ghost_plane(self, which=0, side=1):
        g = self.grid[self.glo + self.nx:] if side else self.grid[: self.glo]
        return torch.from_numpy(np.ascontiguousarray(g).reshape(-1))

    def ghost_accumulate(self, recv, which=0, side=0):
        r = recv.numpy().reshape(-1, self.dims, self.fd)
        if side == 0:
            self.grid[self.glo: self.glo + len(r)] += r
        else:
            self.grid[self.glo + self.nx - len(r): self.glo + self.nx] += r

    def fft_yz(self, which=0):
        self.spec2d = np.fft.rfft2(self.grid[self.glo: self.glo + self.nx, :, : self.dims], axes=(1, 2))

    def pack(self, which=0):
        ny = self.dims // self.nranks
        blocks = self.spec2d.reshape(self.nx, self.nranks, ny, self.nc).transpose(1, 0, 2, 3)
        return torch.from_numpy(np.ascontiguousarray(blocks).view(np.float64).reshape(-1))

    def spectrum_buffer(self, which=0):
        ny = self.dims // self.nranks
        return torch.empty(self.dims * ny * self.nc * 2, dtype=torch.float64)

    def fft_x(self, spec):
        ny = self.dims // self.nranks
        a = spec.numpy().view(np.complex128).reshape(self.dims, ny, self.nc)
        a[...] = np.fft.fft(a, axis=0)

    def power_partial(self, spec_a, spec_b, nrbins):
        """Raw sums of this rank's ky rows: the oracle bins a full-size spectrum that is
        zero outside them; the geometry-only sums (|k|, counts) are contributed by rank 0."""
        ny = self.dims // self.nranks
        full = np.zeros((self.dims, self.dims, self.nc), np.complex128)
        full[:, self.rank * ny: (self.rank + 1) * ny] = spec_a.numpy().view(np.complex128).reshape(self.dims, ny, self.nc)
        _, p, c, k = self.orc.powerspectrum(self.dims, full, None, nrbins, 1.0, 1.0)
        sums = np.zeros(3 * nrbins)
        sums[:nrbins] = p * c
        if self.rank == 0:
            sums[nrbins: 2 * nrbins] = k * c
            sums[2 * nrbins:] = c
        return torch.from_numpy(sums)

    # -- the fused stages of CudaStages --------------------------------------------------
    def fftx_power_partial(self, spec_yz, nrbins):
        spec = spec_yz.clone()
        self.fft_x(spec)
        return self.power_partial(spec, None, nrbins)

    def fft_yz_scatter(self, which=0):
        import torch.distributed as dist
        self.fft_yz(which)
        send = self.pack(which)
        self._recv = self.spectrum_buffer(which)
        if self.nranks == 1:
            self._recv.copy_(send)
        else:
            dist.all_to_all_single(self._recv, send)

    def recv_block(self):
        return self._recv

    def check(self):
        pass
