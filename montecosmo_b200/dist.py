"""
Slab-decomposed particle-mesh engine over torch.distributed (one process per GPU, NCCL) -- SURVEY section 8e.

Decomposition ("Lagrangian slabs"): the mesh [nx, ny, nz] is split along x; rank r owns planes [r*xl, (r+1)*xl) and,
permanently, the particles whose LATTICE x-coordinate lies there.  Particles of a PM run move a bounded distance from
their lattice site (rms ~2 cells, max ~15 at 2.5 Mpc/h cells), so each rank works on a local mesh extended by H halo
planes on both sides and no particle ever migrates: perfect load balance, fixed-size buffers, a trivial tape.  A guard
(checked once per call) raises if any particle comes within a cell of the edge of its extended slab.
HBM is plentiful (180 GB): the halo costs (2H / xl) extra mesh memory and NVLink traffic, nothing else.

Exchange steps per force evaluation (everything else is the single-GPU kernels on the local extended mesh):
  * after paint:      halo planes are SENT to the neighbours and ADDED to their owned planes        (halo_reduce)
  * before readout:   halo planes are FETCHED from the neighbours' owned planes                      (halo_gather)
  * each 3-D FFT:     2-D FFT over (y,z) of the local planes -> all-to-all (x-planes for ky rows) -> 1-D FFT along x;
                      k-space stays split along ky, layout [nx, kyl, nz/2+1]                           (rfftn / irfftn)
halo_gather is the exact transpose of halo_reduce and irfftn o diag(m) o rfftn transposes as on one GPU, so the reverse
sweep is the same sequence with scatter and gather swapping roles.

Positions handed to / returned by this class are DISPLACEMENTS from the particles' own lattice sites (float32; the
kernels are given the frame `self.frame`: site (i, j, k) of the rank's xl x ny x nz lattice sits at cell (H + i, j, k) of
its halo-extended mesh, include/mcpm.h: mcpm_frame).  The CIC fraction then resolves ~1e-7 cell whatever the mesh size,
where an absolute float32 coordinate near 1000 resolves 6e-5.
Reference semantics: montecosmo/nbody.py:583-604 (pm_forces), 634-667 (lpt), 933-951 (BullFrog step).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from ._capi import check, fd_code, frame as _frame

INF = float("inf")


class _Sections:
    """CUDA-event section timers (MCPM_SLAB_PROFILE=1): where a slab evaluation spends its time, per named section."""

    def __init__(self):
        self.on = os.environ.get("MCPM_SLAB_PROFILE", "0") == "1"
        self.pending, self.total, self.count = [], {}, {}

    def wrap(self, obj, names):
        if not self.on:
            return
        for name in names:
            fn = getattr(obj, name)

            def timed(*a, _fn=fn, _name=name, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = _fn(*a, **kw)
                e1.record()
                self.pending.append((_name, e0, e1))
                return out
            setattr(obj, name, timed)

    def report(self, reset=True):
        torch.cuda.synchronize()
        for name, e0, e1 in self.pending:
            self.total[name] = self.total.get(name, 0.0) + e0.elapsed_time(e1)
            self.count[name] = self.count.get(name, 0) + 1
        self.pending = []
        out = {k: {"ms": round(v, 3), "calls": self.count[k]} for k, v in sorted(self.total.items(), key=lambda kv: -kv[1])}
        if reset:
            self.total, self.count = {}, {}
        return out


class SlabPM:
    def __init__(self, ops, mesh_shape, halo=24, group=None, p2p=True):
        self.o, self.lib, self.A = ops, ops.lib, ops.A
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        nx, ny, nz = (int(s) for s in mesh_shape)
        if nx % self.P or ny % self.P or nz % 2:
            raise ValueError("mesh sides nx, ny must be divisible by the number of ranks and nz even")
        self.nx, self.ny, self.nz, self.nzc = nx, ny, nz, nz // 2 + 1
        self.xl, self.kyl = nx // self.P, ny // self.P
        self.H = int(halo)
        if not (1 <= self.H <= self.xl):
            raise ValueError(f"halo must be in [1, {self.xl}] (neighbour-only exchange)")
        self.ext = self.xl + 2 * self.H
        self.x0, self.y0 = self.rank * self.xl, self.rank * self.kyl
        self.N = nx * ny * nz
        self.npl = self.xl * ny * nz  # local particles (lattice == mesh)
        h = C.c_void_p()
        check(self.lib, self.lib.mcpm_slabfft_create(nx, ny, nz, self.P, C.byref(h)))
        self._fft = h
        # fused x-transform (xfft.cu) between the all-to-alls where this build has it for nx; set False to compare
        self.xfuse = bool(self.lib.mcpm_xfuse_supported(nx))
        # ... and, on CUDA with more than one rank, with the all-to-all folded into that kernel: every rank maps its
        # peers' spectrum buffers (torch symmetric memory = CUDA IPC over NVLink) and the kernel loads / stores x-planes
        # at their owners.  MCPM_SLAB_P2P=0 keeps NCCL all-to-alls; any failure to set it up falls back to them.
        # brick-tiled scatters (brick.cu) of the rank's lattice particles into its halo-extended mesh: CUDA builds only;
        # the first call decides (MCPM_EUNSUP -> generic kernels)
        self.brick = os.environ.get("MCPM_SLAB_BRICK", "1") != "0" and bool(self.lib.mcpm_xfuse_supported(64))
        self.p2p, self.p2p_note = False, "off"
        # p2p=False: a geometry that only paints and transforms (the finer paint mesh of the final nufft) needs no peer
        # buffers -- they are 6 half spectra of symmetric memory
        self.prev, self.next = (self.rank - 1) % self.P, (self.rank + 1) % self.P
        self._want_p2p = bool(p2p)
        if p2p and self.xfuse and self.P > 1 and os.environ.get("MCPM_SLAB_P2P", "1") != "0":
            self._setup_p2p()
        if p2p:
            self._agree_p2p()
        ax = [np.arange(self.xl, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(nz, dtype=np.float32)]
        self.q_own = self.A.prepare(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3))  # owned-slab coords
        self.frame = _frame((self.xl, ny, nz), origin=(self.H, 0, 0))  # relative positions on the halo-extended mesh
        self._fr = C.byref(self.frame)
        self._maxdx = None  # running maxima of |x displacement| seen by the guard (device floats, _guard)
        self._hsched = None  # active halo planes per step (set_halo_schedule); None: the whole halo
        self._h = self.H     # active halo planes of the exchange in flight
        # tape_forces = False: the step loop tapes kick positions only and the reverse sweep recomputes each step's force
        # mesh (one more paint + force evaluation per step instead of 16 bytes per extended cell and step: at 1024^3 on 8
        # GPUs with 20 steps that is 59 GB per GPU) -- the trade of the reference's checkpointed adjoint (nbody.py:999)
        self.tape_forces = os.environ.get("MCPM_SLAB_TAPE_FORCES", "1") != "0"
        # two_field: the step loop's distributed x-transforms carry two fields instead of three (xfft_kernel.h: FORCE2)
        self.two_field = os.environ.get("MCPM_SLAB_TWO_FIELD", "1") != "0"
        self.sections = _Sections()
        if self.sections.on and self.A.device.type == "cuda":  # nested sections: inner times are included in outer ones
            self.sections.wrap(self, ["halo_reduce", "halo_gather", "forces_from_density", "density_cotangent",
                                      "lpt_forward", "lpt_backward", "steps_forward", "steps_backward", "rfftn", "irfftn"])

    def __del__(self):
        try:
            if getattr(self, "_fft", None):
                self.lib.mcpm_slabfft_destroy(self._fft)
                self._fft = None
        except Exception:
            pass

    def _setup_p2p(self):
        try:
            import torch.distributed._symmetric_memory as symm
            dev = self.A.device
            if getattr(dev, "type", str(dev)) != "cuda" and not str(dev).startswith("cuda"):
                self.p2p_note = "not a CUDA device"
                return
            grp = self.group if self.group is not None else dist.group.WORLD
            # interleaved complex64 as float32 pairs; 6 components: the 2LPT Hessian set goes through the same buffers
            shape = (6, self.xl, self.ny, self.nzc, 2)
            self._sym_in = symm.empty(shape, dtype=torch.float32, device=dev)
            self._sym_out = symm.empty(shape, dtype=torch.float32, device=dev)
            self._h_in = symm.rendezvous(self._sym_in, grp)
            self._h_out = symm.rendezvous(self._sym_out, grp)
            pin, pout = list(self._h_in.buffer_ptrs), list(self._h_out.buffer_ptrs)
            if len(pin) != self.P or len(pout) != self.P or self.P > 8:
                self.p2p_note = "unexpected peer table"
                return
            self._peer_in = (C.c_void_p * 8)(*(pin + [0] * (8 - self.P)))
            self._peer_out = (C.c_void_p * 8)(*(pout + [0] * (8 - self.P)))
            # the meshes that take part in a halo exchange, in symmetric memory too (csrc/halo.cu: one kernel per exchange
            # reads the neighbours' planes where they lie)
            self._halo = {}
            for key, shape in (("rho", (self.ext, self.ny, self.nz)), ("m3", (3, self.ext, self.ny, self.nz)),
                               ("F", (3, self.xl, self.ny, self.nz)), ("rb", (self.ext, self.ny, self.nz))):
                buf = symm.empty(shape, dtype=torch.float32, device=dev)
                hd = symm.rendezvous(buf, grp)
                ptrs = list(hd.buffer_ptrs)
                if len(ptrs) != self.P:
                    self.p2p_note = "unexpected peer table"
                    return
                self._halo[key] = (buf, hd, ptrs)
            self.p2p, self.p2p_note = True, "symmetric memory: peer loads/stores inside the x-transform and halo kernels"
        except Exception as e:  # no NVLink peer access, missing permissions, older torch: keep the NCCL path
            self.p2p, self.p2p_note = False, f"unavailable ({type(e).__name__}: {e})"

    def _agree_p2p(self):
        """Every rank must take the same path: one that fell back to NCCL while its peers wait in a symmetric-memory
        barrier would deadlock the job.  p2p stays on only if it came up on ALL ranks."""
        if self.P > 1:
            flag = torch.tensor([1.0 if self.p2p else 0.0], device=self.A.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if self.p2p and float(flag) == 0.0:
                self.p2p, self.p2p_note = False, "off: peer memory did not come up on every rank"

    def _halo_buf(self, key, shape):
        """The mesh of an exchange: my symmetric buffer and (mine, prev's, next's) pointers; a plain tensor with my own
        pointer three times when this is the only rank (the periodic wrap of a single slab)."""
        if self.p2p:
            buf, hd, ptrs = self._halo[key]
            return buf, hd, (ptrs[self.rank], ptrs[self.prev], ptrs[self.next])
        buf = self.A.empty(shape)
        return buf, None, (buf.data_ptr(),) * 3

    @property
    def peer_halos(self):
        """Halo exchanges as csrc/halo.cu kernels: over symmetric memory (p2p), or on one CUDA / host rank."""
        return self.p2p or self.P == 1

    def _peer_force(self, real_in, nb_in, transpose, out=None):
        """real_in [nb_in, xl, ny, nz] -> [nb_out, xl, ny, nz]: local 2-D R2C into the symmetric buffer, barrier, the
        fused kernel on my ky block reading / writing every rank's buffer, barrier, local 2-D C2R."""
        st = self._st()
        self._call("mcpm_slabfft_r2c_yz", self._fft, st, real_in.data_ptr(), self._sym_in.data_ptr(), nb_in)
        two = self.two_field  # move (F_x, potential) / (C_x, g_y C_y + g_z C_z) through the exchange instead of 3 fields:
        # the y and z gradient factors do not depend on x, so they are applied on the local (y,z) spectra
        if two and transpose:
            self._call("mcpm_yz_gradients", st, self._sym_in.data_ptr(), self.xl, self.ny, self.nz, 0, 1)
        self._h_in.barrier()
        self._call("mcpm_xfuse_force_peer", st, self._peer_in, self._peer_out, self.P, int(transpose) + (2 if two else 0),
                   self.nx, self.ny, self.nz, self.kyl, self.y0, 0, 0, 0.0, 0, 1.0 / self.N)
        self._h_out.barrier()
        if two and not transpose:
            self._call("mcpm_yz_gradients", st, self._sym_out.data_ptr(), self.xl, self.ny, self.nz, 0, 0)
        nb_out = 1 if transpose else 3
        out = self.A.empty((nb_out, self.xl, self.ny, self.nz)) if out is None else out
        self._call("mcpm_slabfft_c2r_yz", self._fft, st, self._sym_out.data_ptr(), out.data_ptr(), nb_out)
        return out

    def _peer_k2x(self, mode, dk, nb_out):
        """Local ky block of a 3-D spectrum -> nb_out real meshes of my planes [nb_out, xl, ny, nz]: the multiply, the
        inverse x-transforms and the transpose in one kernel that stores every x-plane at its owner (mcpm_xfuse_peer
        mode 2 FORCE_K / 3 HESS_K), then my local 2-D C2R.  lpt's force and Hessian meshes (nbody.py:595-603, 611-627)."""
        st = self._st()
        self._h_in.barrier()  # every rank has finished the C2R that read its output buffer last
        self._call("mcpm_xfuse_peer", st, mode, None, self._peer_out, dk.data_ptr(), self.P, self.nx, self.ny, self.nz,
                   self.kyl, self.y0, 0, 0, 0, 0, 1.0 / self.N)
        self._h_out.barrier()
        out = self.A.empty((nb_out, self.xl, self.ny, self.nz))
        self._call("mcpm_slabfft_c2r_yz", self._fft, st, self._sym_out.data_ptr(), out.data_ptr(), nb_out)
        return out

    def _peer_x2k(self, mode, real_in, dkbar, accumulate):
        """nb real meshes of my planes -> (added into) my ky block of the cotangent spectrum: local 2-D R2C, then one
        kernel that loads every x-plane from its owner, transforms along x and applies the transposed operator with the
        Hermitian weights (mode 4 FORCE_TK / 5 HESS_TK).  lpt_vjp (engine.cu)."""
        st = self._st()
        nb = real_in.shape[0]
        self._call("mcpm_slabfft_r2c_yz", self._fft, st, real_in.data_ptr(), self._sym_in.data_ptr(), nb)
        self._h_in.barrier()
        self._call("mcpm_xfuse_peer", st, mode, self._peer_in, None, dkbar.data_ptr(), self.P, self.nx, self.ny, self.nz,
                   self.kyl, self.y0, 0, 0, 1, int(accumulate), 1.0)
        self._h_out.barrier()  # nobody overwrites its input buffer while a peer still reads it
        return dkbar

    # ------------------------------------------------------------------------------------------------ helpers
    def _call(self, name, *args):
        check(self.lib, getattr(self.lib, name)(*args))

    def _st(self):
        return self.A.stream()

    def _global_rank(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def scatter_real(self, full):
        """Global real mesh [.., nx, ny, nz] -> my owned planes (test / setup helper)."""
        return self.A.prepare(full[..., self.x0:self.x0 + self.xl, :, :])

    def scatter_spectrum(self, full_k):
        """Global half spectrum [nx, ny, nzc] -> my ky block [nx, kyl, nzc]."""
        return self.A.prepare(full_k[..., :, self.y0:self.y0 + self.kyl, :], "c64")

    # ------------------------------------------------------------------------------------------------ distributed FFT
    def _a2a(self, send):
        if self.P == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=self.group)
        return recv

    def rfftn_yz(self, a):
        """[nb, xl, ny, nz] real (owned planes) -> [nb, nx, kyl, nzc] complex (my ky block) transformed along (y,z)
        only: batched 2-D R2C of the local planes + the all-to-all that trades x-planes for ky rows."""
        a = self.A.prepare(a)
        nb = a.shape[0]
        b = self.A.empty((nb, self.xl, self.ny, self.nzc), "c64")
        self._call("mcpm_slabfft_r2c_yz", self._fft, self._st(), a.data_ptr(), b.data_ptr(), nb)
        if self.P == 1:  # one rank: the two layouts coincide
            return b
        # send[q] = my planes restricted to rank q's ky rows
        send = b.view(nb, self.xl, self.P, self.kyl, self.nzc).permute(2, 0, 1, 3, 4).contiguous()
        recv = self._a2a(send)  # recv[q] = rank q's planes, my ky rows
        return recv.permute(1, 0, 2, 3, 4).contiguous().view(nb, self.nx, self.kyl, self.nzc)

    def rfftn(self, a):
        """[nb, xl, ny, nz] real (owned planes) -> [nb, nx, kyl, nzc] complex (my ky block), unnormalised."""
        c = self.rfftn_yz(a)
        self._call("mcpm_slabfft_c2c_x", self._fft, self._st(), c.data_ptr(), c.shape[0], 0)
        return c

    def _project_yz(self, b):
        """Hermitian projection along ky of the self-conjugate planes kz = 0 and kz = Nyquist of b [nb, xl, ny, nzc], in
        place.  After the inverse x-transform this equals the 3-D projection jnp.fft.irfftn applies implicitly (it
        returns the real part of the full inverse transform); cuFFT's 2-D C2R on inconsistent input is algorithm
        dependent (Hermitian consistency).  The engine's projection kernel with nx = 1: every local x-plane is one
        2-D spectrum.  Two planes out of nz/2+1: negligible traffic."""
        self._call("mcpm_hermitian_project", self._st(), b.data_ptr(), 1, self.ny, self.nz, b.shape[0] * b.shape[1])

    def irfftn(self, c, overwrite=False, x_done=False, project=False):
        """[nb, nx, kyl, nzc] complex -> [nb, xl, ny, nz] real, UNNORMALISED (fold 1/N into the preceding Fourier pass).
        x_done: the inverse transform along x has already been applied (fused x-transform kernels).
        project: the input is not Hermitian-consistent by construction (a cotangent, an interlaced sum): project it."""
        c = self.A.prepare(c, "c64")
        if not overwrite:
            c = c.clone()
        nb = c.shape[0]
        if not x_done:
            self._call("mcpm_slabfft_c2c_x", self._fft, self._st(), c.data_ptr(), nb, 1)
        if self.P == 1:
            b = c
        else:
            send = c.view(nb, self.P, self.xl, self.kyl, self.nzc).permute(1, 0, 2, 3, 4).contiguous()
            recv = self._a2a(send)  # recv[q] = my planes, rank q's ky rows
            b = recv.permute(1, 2, 0, 3, 4).contiguous().view(nb, self.xl, self.ny, self.nzc)
        if project:
            self._project_yz(b)
        out = self.A.empty((nb, self.xl, self.ny, self.nz))
        self._call("mcpm_slabfft_c2r_yz", self._fft, self._st(), b.data_ptr(), out.data_ptr(), nb)
        return out

    # ------------------------------------------------------------------------------------------------ halos
    def _exchange(self, to_prev, to_next, from_next, from_prev):
        if self.P == 1:
            from_next.copy_(to_prev)
            from_prev.copy_(to_next)
            return
        ops = [dist.P2POp(dist.isend, to_prev, self._global_rank(self.prev), self.group),
               dist.P2POp(dist.isend, to_next, self._global_rank(self.next), self.group),
               dist.P2POp(dist.irecv, from_next, self._global_rank(self.next), self.group),
               dist.P2POp(dist.irecv, from_prev, self._global_rank(self.prev), self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def halo_reduce(self, ext, lead=False):
        """ext [ext, ny, nz(,4)] (lead: [c, ext, ny, nz]): send both halos to the neighbours, add theirs into my owned
        planes."""
        H, xl, h = self.H, self.xl, self._h  # h <= H active planes: the ones next to the owned region
        e = ext.movedim(1, 0) if lead else ext  # a view with the plane index first
        a, b = torch.empty_like(e[:h].contiguous()), torch.empty_like(e[:h].contiguous())
        self._exchange(e[H - h:H].contiguous(), e[H + xl:H + xl + h].contiguous(), a, b)
        e[xl + H - h:xl + H] += a  # next's left halo covers my last owned planes
        e[H:H + h] += b  # prev's right halo covers my first owned planes

    def halo_gather(self, ext):
        """ext [ext, ny, nz(,4)]: fill both halos from the neighbours' owned planes (transpose of halo_reduce)."""
        H, xl, h = self.H, self.xl, self._h
        right, left = torch.empty_like(ext[:h]), torch.empty_like(ext[:h])
        self._exchange(ext[H:H + h].contiguous(), ext[xl + H - h:xl + H].contiguous(), right, left)
        ext[H + xl:H + xl + h] = right
        ext[H - h:H] = left

    NSLOT = 129  # slot 0: every guarded position; slot 1 + s: the kick positions of step s (s < 128)

    def _guard(self, pos, step=None):
        # site i + displacement d must keep the CIC stencil inside the extended slab: 1 <= H + i + d <= ext - 2 for every
        # owned i in [0, xl).  One pass over the strided x column (mcpm_absmax) into running maxima on the device; nothing
        # synchronises.  The kick positions of step s are also tracked on their own: they size that step's active halo.
        if self._maxdx is None:
            self._maxdx = self.A.zeros((self.NSLOT,))
        self._call("mcpm_absmax", self._st(), pos.data_ptr(), pos.shape[0], 3, self._maxdx.data_ptr())
        if step is not None and step + 1 < self.NSLOT:
            self._call("mcpm_absmax", self._st(), pos.data_ptr(), pos.shape[0], 3, self._maxdx[step + 1:].data_ptr())

    def _maxima(self):
        """Host copy of the running maxima, all-reduced over the ranks (one read, one collective)."""
        m = self._maxdx.detach().clone()
        if self.P > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        return m.cpu().numpy()

    def halo_needed(self, factor=1.25, margin=2):
        """Halo planes that would have held every particle seen by the guard since this object was built (max over ranks
        of |x displacement|, times `factor`, plus `margin` planes; the CIC stencil needs one more than the displacement).
        One device -> host read and, with several ranks, one all-reduce: call it between evaluations, e.g. after a
        warm-up one, to size the halo of a run from a measurement instead of the a = 1 default of 24 planes."""
        import math
        if self._maxdx is None:
            return self.H
        return int(math.ceil(float(self._maxima()[0]) * factor)) + int(margin)

    def halo_schedule(self, n_steps, factor=1.25, margin=2):
        """Active halo planes per step of the loop, from the largest |x displacement| seen at each step's kick positions
        (same rule as halo_needed, capped at the allotted halo): early steps exchange a few planes, the last ones all."""
        import math
        if self._maxdx is None:
            return None
        m = self._maxima()
        return [max(1, min(self.H, int(math.ceil(float(m[1 + s]) * factor)) + int(margin))) for s in range(n_steps)]

    def set_halo_schedule(self, sched):
        """sched[s] = active halo planes of step s (None: the whole halo every step).  The guard checks each step against
        its own entry, so a later evaluation that outgrows the schedule raises instead of losing deposits."""
        self._hsched = None if sched is None else [int(v) for v in sched]

    def _active(self, s):
        return self.H if self._hsched is None or s >= len(self._hsched) else self._hsched[s]

    def check_guard(self):
        """Raise if any particle came within a cell of the edge of its extended slab since the last check (one sync)."""
        if self._maxdx is None:
            return
        m = self._maxdx.detach().cpu().numpy()  # one read
        if not float(m[0]) <= self.H - 1.0:  # also catches NaN
            raise RuntimeError(f"a particle left its extended slab: increase halo (= {self.H} planes)")
        if self._hsched is not None:
            for s, h in enumerate(self._hsched):
                if 1 + s < self.NSLOT and not float(m[1 + s]) <= h - 1.0:
                    raise RuntimeError(f"step {s}: a particle left the {h} active halo planes of the schedule "
                                       f"(largest x-displacement {float(m[1 + s]):.2f}): set_halo_schedule(None) or a wider one")

    # ------------------------------------------------------------------------------------------------ Fourier passes
    def force_spectra(self, dk, lap_fd=INF, grad_fd=INF, deconv_order=0):
        out = self.A.empty((3, self.nx, self.kyl, self.nzc), "c64")
        self._call("mcpm_force_spectra_slab", self._st(), dk.data_ptr(), out.data_ptr(), self.nx, self.ny, self.nz,
                   self.kyl, self.y0, fd_code(lap_fd), fd_code(grad_fd), 0.0, deconv_order, 1.0 / self.N)
        return out

    def force_spectra_T(self, in3, out=None, half_weights=False, lap_fd=INF, grad_fd=INF, deconv_order=0):
        acc = out is not None
        out = self.A.empty((self.nx, self.kyl, self.nzc), "c64") if out is None else out
        self._call("mcpm_force_spectra_T_slab", self._st(), in3.data_ptr(), out.data_ptr(), self.nx, self.ny, self.nz,
                   self.kyl, self.y0, fd_code(lap_fd), fd_code(grad_fd), 0.0, deconv_order, int(half_weights), int(acc),
                   1.0 if half_weights else 1.0 / self.N)
        return out

    # density planes -> three force meshes, and the transpose (3 meshes -> 1), each with ONE kernel between the
    # all-to-alls where the fused x-transform exists for nx (xfft.cu), else c2c_x + streaming multiply + c2c_x
    def forces_from_density(self, rho_owned, out=None):
        if self.p2p:
            return self._peer_force(rho_owned.contiguous().unsqueeze(0), 1, False, out=out)
        if self.xfuse:
            c = self.rfftn_yz(rho_owned.unsqueeze(0))
            out = self.A.empty((3, self.nx, self.kyl, self.nzc), "c64")
            self._call("mcpm_xfuse_force_slab", self._st(), c.data_ptr(), out.data_ptr(), self.nx, self.ny, self.nz,
                       self.kyl, self.y0, 0, 0, 0.0, 0, 1.0 / self.N)
            return self.irfftn(out, overwrite=True, x_done=True)
        rk = self.rfftn(rho_owned.unsqueeze(0))
        return self.irfftn(self.force_spectra(rk[0]), overwrite=True)

    def density_cotangent(self, planar3, out=None):
        if self.p2p:
            return self._peer_force(planar3.contiguous(), 3, True, out=None if out is None else out.unsqueeze(0))[0]
        if self.xfuse:
            c = self.rfftn_yz(planar3)
            out = self.A.empty((1, self.nx, self.kyl, self.nzc), "c64")
            self._call("mcpm_xfuse_force_T_slab", self._st(), c.data_ptr(), out.data_ptr(), self.nx, self.ny, self.nz,
                       self.kyl, self.y0, 0, 0, 0.0, 0, 1.0 / self.N)
            return self.irfftn(out, overwrite=True, x_done=True)[0]
        rk = self.force_spectra_T(self.rfftn(planar3))
        return self.irfftn(rk.unsqueeze(0), overwrite=True)[0]

    # ------------------------------------------------------------------------------------------------ force step
    def force_mesh4(self, pos, order=2):
        """Extended float4 force mesh {Fx,Fy,Fz,0} [ext, ny, nz, 4] at the local positions (nbody.py:583-603)."""
        A, lib, st = self.A, self.lib, self._st()
        peer = self.peer_halos
        plane = self.ny * self.nz
        if peer:
            rho, hd_rho, p_rho = self._halo_buf("rho", (self.ext, self.ny, self.nz))
            rho.zero_()
        else:
            rho = A.zeros((self.ext, self.ny, self.nz))
        one = (C.c_float * 3)(1.0, 1.0, 1.0)
        # brick-tiled scatter of my xl x ny x nz lattice particles into the halo-extended mesh where this build has it
        if not (order == 2 and self.brick and self._brick_ok(lib.mcpm_paint_brick_f(
                st, self._fr, self.xl, self.ny, self.nz, pos.data_ptr(), 0, 1.0, 0.0, pos.shape[0], self.ext, self.ny,
                self.nz, rho.data_ptr()))):
            self._call("mcpm_paint_f", st, self._fr, pos.data_ptr(), 0, 1.0, pos.shape[0], self.ext, self.ny, self.nz,
                       order, one, 0.0, rho.data_ptr(), 1)
        fm4 = A.empty((self.ext, self.ny, self.nz, 4))
        if peer:  # one kernel per exchange over the neighbours' memory (csrc/halo.cu); barriers order the ranks
            if hd_rho is not None:
                hd_rho.barrier()
            self._call("mcpm_halo_reduce_peer", st, p_rho[0], p_rho[1], p_rho[2], self.H, self.xl, plane, 1, self._h)
            Fb, hd_F, p_F = self._halo_buf("F", (3, self.xl, self.ny, self.nz))
            F = self.forces_from_density(rho[self.H:self.H + self.xl], out=Fb if self.p2p else None)
            if not self.p2p:
                p_F = (F.data_ptr(),) * 3
            if hd_F is not None:
                hd_F.barrier()
            # interleave {Fx, Fy, Fz, 0} and fetch the halo planes from the neighbours' owned planes in one pass
            self._call("mcpm_halo_gather4_peer", st, fm4.data_ptr(), p_F[0], p_F[1], p_F[2], self.H, self.xl, plane, self._h)
            return fm4
        self.halo_reduce(rho)
        F = self.forces_from_density(rho[self.H:self.H + self.xl])  # [3, xl, ny, nz]
        own = fm4[self.H:self.H + self.xl]
        self._call("mcpm_interleave3", st, F.data_ptr(), own.data_ptr(), self.xl * self.ny * self.nz)
        self.halo_gather(fm4)
        return fm4

    def _brick_ok(self, code):
        """True: the brick kernel ran.  False: this geometry / build has none (MCPM_EUNSUP), take the generic kernel.
        Any other code is a real error and is raised -- a failed launch must not be papered over by the fallback."""
        if code == 0:
            return True
        if code == 5:  # MCPM_EUNSUP
            return False
        check(self.lib, code)
        return False

    def steps_forward(self, pos, vel, alpha, beta, drift_pre, drift_post, tape=True):
        """BullFrog DKD steps in place on local (pos, vel); CIC.  Returns the tape [(x_kick, fm4), ...]."""
        ns = len(alpha)
        st = self._st()
        self._call("mcpm_drift", st, pos.data_ptr(), vel.data_ptr(), float(drift_pre[0]), pos.shape[0])
        out = []
        for s in range(ns):
            self._guard(pos, step=s)
            self._h = self._active(s)
            fm4 = self.force_mesh4(pos)
            if tape:
                out.append((pos.clone(), fm4 if self.tape_forces else None))
            dcomb = float(drift_post[s]) + (float(drift_pre[s + 1]) if s + 1 < ns else 0.0)
            self._call("mcpm_kick_drift4_f", st, self._fr, pos.data_ptr(), vel.data_ptr(), fm4.data_ptr(), pos.shape[0],
                       self.ext, self.ny, self.nz, float(alpha[s]), float(beta[s]), dcomb)
        self._h = self.H
        self._guard(pos)
        return out

    def steps_backward(self, tape, posbar, velbar, alpha, beta, drift_pre, drift_post):
        """Reverse sweep of steps_forward, in place on (posbar, velbar)."""
        A, st = self.A, self._st()
        ns = len(alpha)
        n = posbar.shape[0]
        cells = self.xl * self.ny * self.nz
        dcomb = lambda s: float(drift_post[s]) + (float(drift_pre[s + 1]) if s + 1 < ns else 0.0)
        # vbar += xbar * dcomb leads every reverse step; for all but the first it rides in the previous gather (dnext)
        if ns:
            self._call("mcpm_drift", st, velbar.data_ptr(), posbar.data_ptr(), dcomb(ns - 1), n)
        for s in reversed(range(ns)):
            x1, fm4 = tape[s]
            self._h = self._active(s)  # the exchanges of a reverse step are the transposes of its forward ones
            if fm4 is None:  # not taped: recompute from the kick positions
                fm4 = self.force_mesh4(x1)
            dnext = dcomb(s - 1) if s > 0 else float(drift_pre[0])
            peer = self.peer_halos
            plane = self.ny * self.nz
            m3 = hd3 = p3 = None
            if self.brick:
                if peer:
                    m3, hd3, p3 = self._halo_buf("m3", (3, self.ext, self.ny, self.nz))
                    m3.zero_()
                else:
                    m3 = A.zeros((3, self.ext, self.ny, self.nz))
            if m3 is not None and self._brick_ok(self.lib.mcpm_paint3_brick_f(
                    st, self._fr, self.xl, self.ny, self.nz, x1.data_ptr(), velbar.data_ptr(), 0, 0.0,
                    float(beta[s]), n, self.ext, self.ny, self.nz, m3.data_ptr())):
                if peer:
                    if hd3 is not None:
                        hd3.barrier()
                    self._call("mcpm_halo_reduce_peer", st, p3[0], p3[1], p3[2], self.H, self.xl, plane, 3, self._h)
                else:
                    self.halo_reduce(m3, lead=True)  # three planar extended meshes
                planar = m3[:, self.H:self.H + self.xl].contiguous()
            else:
                m4 = A.zeros((self.ext, self.ny, self.nz, 4))
                self._call("mcpm_paint3v4_f", st, self._fr, x1.data_ptr(), velbar.data_ptr(), 0, 0.0,
                           float(beta[s]), n, self.ext, self.ny, self.nz, m4.data_ptr())
                if self.P == 1:  # one rank: the periodic wrap, same kernel on the float4 mesh
                    self._call("mcpm_halo_reduce_peer", st, m4.data_ptr(), m4.data_ptr(), m4.data_ptr(), self.H, self.xl,
                               4 * plane, 1, self._h)
                else:
                    self.halo_reduce(m4)
                planar = A.empty((3, self.xl, self.ny, self.nz))
                self._call("mcpm_deinterleave3", st, m4[self.H:self.H + self.xl].data_ptr(), planar.data_ptr(), cells)
            if peer:
                rhobar, hdr, pr = self._halo_buf("rb", (self.ext, self.ny, self.nz))
                own = rhobar[self.H:self.H + self.xl]
                res = self.density_cotangent(planar, out=own if self.p2p else None)
                if not self.p2p:
                    own.copy_(res)
                if hdr is not None:
                    hdr.barrier()
                self._call("mcpm_halo_gather_peer", st, pr[0], pr[1], pr[2], self.H, self.xl, plane, 1, self._h)
            else:
                rhobar = A.empty((self.ext, self.ny, self.nz))
                rhobar[self.H:self.H + self.xl] = self.density_cotangent(planar)
                self.halo_gather(rhobar)
            self._call("mcpm_read_grad4v_step_f", st, self._fr, x1.data_ptr(), fm4.data_ptr(), rhobar.data_ptr(),
                       velbar.data_ptr(), float(beta[s]), float(alpha[s]), dnext, n, self.ext, self.ny, self.nz,
                       posbar.data_ptr())
        self._h = self.H

    # ------------------------------------------------------------------------------------------------ LPT
    def _lattice_read3(self, planar3):
        """NGP read of three owned-slab meshes at the owned lattice sites -> [npl, 3] (nbody.py:602 note)."""
        out = self.A.empty((self.npl, 3))
        one = (C.c_float * 3)(1.0, 1.0, 1.0)
        self._call("mcpm_read", self._st(), self.q_own.data_ptr(), planar3.data_ptr(), 3, self.npl, self.xl, self.ny,
                   self.nz, 1, one, 0.0, out.data_ptr())
        return out

    def _lattice_paint3(self, vals3):
        out = self.A.empty((3, self.xl, self.ny, self.nz))
        self._call("mcpm_paint3", self._st(), self.q_own.data_ptr(), vals3.data_ptr(), 1.0, self.npl, self.xl, self.ny,
                   self.nz, 1, out.data_ptr(), 0)
        return out

    def lpt_forward(self, dk, d1, d2, dv2, lpt_order=2):
        """delta_k block [nx, kyl, nzc] -> (displacement, vel) of the owned lattice particles and the tape (nbody.py:634-667)."""
        A, st = self.A, self._st()
        dk = A.prepare(dk, "c64")
        f2 = h6 = None
        if self.p2p:  # every transform's transpose inside the fused x-transform kernel, over peer memory
            f1 = self._lattice_read3(self._peer_k2x(2, dk, 3))
            if lpt_order == 2:
                h6 = self._peer_k2x(3, dk, 6)
                d2m = A.empty((1, self.xl, self.ny, self.nz))
                self._call("mcpm_lpt2_source", st, h6.data_ptr(), d2m.data_ptr(), self.xl * self.ny * self.nz)
                f2 = self._lattice_read3(self._peer_force(d2m, 1, False))
        else:
            f1 = self._lattice_read3(self.irfftn(self.force_spectra(dk), overwrite=True))
            if lpt_order == 2:
                h6k = A.empty((6, self.nx, self.kyl, self.nzc), "c64")
                self._call("mcpm_hessian_spectra_slab", st, dk.data_ptr(), h6k.data_ptr(), self.nx, self.ny, self.nz,
                           self.kyl, self.y0, 0, 0, 1.0 / self.N)
                h6 = self.irfftn(h6k, overwrite=True)
                d2m = A.empty((1, self.xl, self.ny, self.nz))
                self._call("mcpm_lpt2_source", st, h6.data_ptr(), d2m.data_ptr(), self.xl * self.ny * self.nz)
                f2 = self._lattice_read3(self.irfftn(self.force_spectra(self.rfftn(d2m)[0]), overwrite=True))
        dpos, vel = A.empty((self.npl, 3)), A.empty((self.npl, 3))
        self._call("mcpm_lpt_combine", st, 0, f1.data_ptr(), 0 if f2 is None else f2.data_ptr(), float(d1), float(d2),
                   float(dv2), self.npl, dpos.data_ptr(), vel.data_ptr(), 0)
        return dpos, vel, (h6, (float(d1), float(d2), float(dv2)), lpt_order)  # dpos IS the relative position

    def lpt_backward(self, tape, posbar, velbar):
        """(posbar, velbar) -> cotangent of the delta_k block, convention dL/dRe + i dL/dIm (engine.cu:lpt_vjp)."""
        A, st, o = self.A, self._st(), self.o
        h6, (d1, d2, dv2), lpt_order = tape
        dkbar = None
        if self.p2p:
            dkbar = A.empty((self.nx, self.kyl, self.nzc), "c64")
            if lpt_order == 2:
                f2bar = o.axpby(posbar, -d2, velbar, -dv2)
                d2bar = self._peer_force(self._lattice_paint3(f2bar), 3, True)
                hbar = A.empty((6, self.xl, self.ny, self.nz))
                self._call("mcpm_lpt2_source_vjp", st, h6.data_ptr(), d2bar.data_ptr(), hbar.data_ptr(),
                           self.xl * self.ny * self.nz)
                self._peer_x2k(5, hbar, dkbar, False)
            f1bar = o.axpby(posbar, d1, velbar, 1.0)
            return self._peer_x2k(4, self._lattice_paint3(f1bar), dkbar, lpt_order == 2)
        if lpt_order == 2:
            f2bar = o.axpby(posbar, -d2, velbar, -dv2)
            rk = self.force_spectra_T(self.rfftn(self._lattice_paint3(f2bar)))
            d2bar = self.irfftn(rk.unsqueeze(0), overwrite=True)
            hbar = A.empty((6, self.xl, self.ny, self.nz))
            self._call("mcpm_lpt2_source_vjp", st, h6.data_ptr(), d2bar.data_ptr(), hbar.data_ptr(),
                       self.xl * self.ny * self.nz)
            hk = self.rfftn(hbar)
            dkbar = A.empty((self.nx, self.kyl, self.nzc), "c64")
            self._call("mcpm_hessian_spectra_T_slab", st, hk.data_ptr(), dkbar.data_ptr(), self.nx, self.ny, self.nz,
                       self.kyl, self.y0, 0, 0, 1, 0, 1.0)
        f1bar = o.axpby(posbar, d1, velbar, 1.0)
        return self.force_spectra_T(self.rfftn(self._lattice_paint3(f1bar)), out=dkbar, half_weights=True)

    # ------------------------------------------------------------------------------------------------ nbody_bf
    def nbody_forward(self, dk, cosmo, a0=0.0, a1=1.0, n_steps=5, lpt_order=2):
        """lpt at a0 then n_steps BullFrog steps (nbody.py:967-1002) on this rank's particles.  Returns local (pos, vel)
        and the tape for nbody_backward."""
        from . import cosmo as _c
        d1, d2, dv2 = float(_c.a2g(cosmo, a0)), float(_c.a2g2(cosmo, a0)), float(_c.a2dg2dg(cosmo, a0))
        al, be, pre, post, _, _ = _c.bullfrog_coefficients(cosmo, a0, a1, n_steps)
        co = [t.tolist() for t in (al, be, pre, post)]
        pos, vel, ltape = self.lpt_forward(dk, d1, d2, dv2, lpt_order)
        stape = self.steps_forward(pos, vel, *co)
        self.check_guard()
        return pos, vel, (ltape, stape, co)

    def nbody_backward(self, tape, posbar, velbar):
        ltape, stape, co = tape
        posbar, velbar = posbar.clone(), velbar.clone()
        self.steps_backward(stape, posbar, velbar, *co)
        return self.lpt_backward(ltape, posbar, velbar)


class SlabResize:
    """chreshape (utils.py:975-1013) from the spectrum of one slab geometry to a coarser one -- the Fourier crop between
    the paint mesh and the final mesh of `nufft` (nbody.py:575-576) when both are split along ky over the same ranks.

    Per axis the crop keeps the low frequencies and folds the two old rows at +-s/2 into the new Nyquist row with
    weights 1/sqrt2.  On x and kz that is local to a rank (mcpm_chreshape_crop_xz_slab); kz additionally needs the
    mirrored element conj in[-kx, -ky] on ONE plane, which is all-gathered (nx * ny complex numbers); on ky it is a
    redistribution of whole rows (one all-to-all with uneven splits) in which the two rows that merge carry 1/sqrt2.
    `backward` is the transpose in the real inner product (the convention of mcpm_chreshape_vjp).
    """

    def __init__(self, pm_in, pm_out):
        if pm_in.P != pm_out.P or pm_in.rank != pm_out.rank:
            raise ValueError("both slab geometries must live on the same ranks")
        if pm_out.nx > pm_in.nx or pm_out.ny > pm_in.ny or pm_out.nz > pm_in.nz:
            raise ValueError("SlabResize crops: the output mesh must not exceed the input mesh on any axis")
        self.i, self.o, self.A, self.lib = pm_in, pm_out, pm_in.A, pm_in.lib
        a, b, P, rank = pm_in, pm_out, pm_in.P, pm_in.rank
        self.scale = float(b.N) / float(a.N)
        r2 = 0.7071067811865476
        # row plan: input row j2 (frequency f) -> output row and weight
        plan = []  # (src rank, dst rank, j2, j_out, w)
        for j2 in range(a.ny):
            f = j2 if j2 < (a.ny + 1) // 2 else j2 - a.ny
            if b.ny == a.ny:
                jo, w = j2, 1.0
            elif -b.ny // 2 < f < b.ny // 2:
                jo, w = f % b.ny, 1.0
            elif abs(f) == b.ny // 2:
                jo, w = b.ny // 2, r2
            else:
                continue
            plan.append((j2 // a.kyl, jo // b.kyl, j2, jo, w))
        mine = sorted((p for p in plan if p[0] == rank), key=lambda p: (p[1], p[3], p[2]))
        to_me = sorted((p for p in plan if p[1] == rank), key=lambda p: (p[0], p[3], p[2]))
        dev = self.A.device
        self.send_rows = torch.tensor([p[2] - a.y0 for p in mine], dtype=torch.long, device=dev)
        self.send_split = [sum(1 for p in mine if p[1] == d) for d in range(P)]
        self.recv_rows = torch.tensor([p[3] - b.y0 for p in to_me], dtype=torch.long, device=dev)
        self.recv_split = [sum(1 for p in to_me if p[0] == s) for s in range(P)]
        w = np.ones(a.kyl, dtype=np.float32)
        for p in mine:
            w[p[2] - a.y0] = p[4]
        self.row_w = self.A.prepare(w)
        self.crop_z = b.nzc < a.nzc

    def _rows_a2a(self, send, send_split, recv_split):
        """[R_send, nx, nzc] complex, rows grouped by destination -> [R_recv, nx, nzc], rows grouped by source."""
        if self.i.P == 1:
            return send
        recv = torch.empty((sum(recv_split), *send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send.contiguous()), recv_split, send_split,
                               group=self.i.group)
        return recv

    def _gather_plane(self, blk):
        """Plane l = nz_out/2 of the whole input spectrum, [nx_in, ny_in], from every rank's rows."""
        a = self.i
        mine = blk[:, :, self.o.nz // 2].contiguous()
        if a.P == 1:
            return mine
        parts = [torch.empty_like(mine) for _ in range(a.P)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(mine), group=a.group)
        return torch.cat(parts, dim=1).contiguous()

    def forward(self, blk):
        """[nx_in, kyl_in, nzc_in] (my ky rows of the input spectrum) -> [nx_out, kyl_out, nzc_out]."""
        a, b, A = self.i, self.o, self.A
        blk = A.prepare(blk, "c64")
        plane = self._gather_plane(blk) if self.crop_z else None
        xz = A.empty((b.nx, a.kyl, b.nzc), "c64")
        a._call("mcpm_chreshape_crop_xz_slab", a._st(), blk.data_ptr(), a.nx, a.ny, a.nz, a.kyl, a.y0,
                0 if plane is None else plane.data_ptr(), self.row_w.data_ptr(), xz.data_ptr(), b.nx, b.nz, self.scale)
        send = xz[:, self.send_rows, :].permute(1, 0, 2).contiguous()
        recv = self._rows_a2a(send, self.send_split, self.recv_split)
        out = torch.zeros((b.nx, b.kyl, b.nzc), dtype=blk.dtype, device=blk.device)
        # the two rows of the new ky Nyquist land on one index; accumulated on the float pairs (every backend has that)
        torch.view_as_real(out).index_add_(1, self.recv_rows, torch.view_as_real(recv.permute(1, 0, 2).contiguous()))
        return out

    def backward(self, outbar):
        """Transpose of forward: [nx_out, kyl_out, nzc_out] -> [nx_in, kyl_in, nzc_in]."""
        a, b, A = self.i, self.o, self.A
        outbar = A.prepare(outbar, "c64")
        send = outbar[:, self.recv_rows, :].permute(1, 0, 2).contiguous()
        got = self._rows_a2a(send, self.recv_split, self.send_split)
        xzbar = torch.zeros((b.nx, a.kyl, b.nzc), dtype=outbar.dtype, device=outbar.device)
        xzbar[:, self.send_rows, :] = got.permute(1, 0, 2)
        inbar = A.empty((a.nx, a.kyl, a.nzc), "c64")
        pbar = torch.zeros((a.nx, a.ny), dtype=outbar.dtype, device=outbar.device) if self.crop_z else None
        a._call("mcpm_chreshape_crop_xz_slab_vjp", a._st(), xzbar.data_ptr(), b.nx, b.nz, a.kyl, a.y0,
                self.row_w.data_ptr(), inbar.data_ptr(), a.nx, a.ny, a.nz, 0 if pbar is None else pbar.data_ptr(),
                self.scale)
        if pbar is not None:
            if a.P > 1:
                dist.all_reduce(torch.view_as_real(pbar), group=a.group)
            inbar[:, :, b.nz // 2] += pbar[:, a.y0:a.y0 + a.kyl]
        return inbar
