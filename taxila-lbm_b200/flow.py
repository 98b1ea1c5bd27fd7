"""Host-side mirror of the reference's flow interface over the C ABI.

Method names follow the reference's module procedures (src/lbm/lbm_flow.F90 public list :84-102,
src/lbm/lbm.F90 LBMInit2/LBMRun2) so a driver written against `Flow` reads like lbm.F90:

    flow = Flow(cfg)                      # FlowCreate + FlowSetUp
    flow.walls_set_values(walls_rg)       # WallsSetValues/WallsCommunicate result
    flow.initialize_state(rho_rg)         # LBMInitializeState result
    flow.fi_init()                        # FlowFiInit            lbm.F90:212
    flow.update_moments()                 # FlowUpdateMoments     lbm.F90:238
    flow.run(istep, kstep)                # LBMRun2 loop body x (kstep-istep)
    flow.get_arrays() / update_diagnostics()

All arrays are numpy float64 in the reference's local ghosted layouts (see geometry.ghosted).
Everything computes on the GPU through libtaxila_gpu.so; nothing here does arithmetic.
"""
import ctypes as C

import numpy as np

from . import capi
from .config import TxgConfig


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _c(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Flow:
    def __init__(self, cfg: TxgConfig, device=0, nccl_id=None):
        self.lib = capi.load()
        self.cfg = cfg.copy()
        self.h = C.c_void_p()
        rc = self.lib.txg_create(C.byref(self.h), C.byref(self.cfg), int(device))
        if rc:
            msg = self.lib.txg_last_error(None)
            raise capi.TaxilaGpuError(rc, msg.decode() if msg else "")
        c = self.cfg
        self.S, self.Q, self.D, self.R = c.ncomponents, c.Q, c.ndims, c.stencil_size_rho
        self.NZl = c.zl if c.ndims == 3 else 1
        self.NY, self.NX = c.NY, c.NX
        if c.nranks > 1:
            if nccl_id is None:
                raise ValueError("nranks > 1 needs the NCCL unique id broadcast from rank 0")
            buf = (C.c_ubyte * 128).from_buffer_copy(bytes(nccl_id))
            self._check(self.lib.txg_comm_init(self.h, buf))

    # ---- plumbing
    def _check(self, rc):
        capi.check(self.lib, self.h, rc)

    def close(self):
        if self.h:
            self.lib.txg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id():
        lib = capi.load()
        buf = (C.c_ubyte * 128)()
        rc = lib.txg_nccl_unique_id(buf)
        if rc:
            raise capi.TaxilaGpuError(rc, lib.txg_last_error(None).decode())
        return bytes(buf)

    # ---- local ghosted shapes (C order == the reference's Fortran arrays)
    def _gz(self, w):
        return w if self.D == 3 else 0

    def shape_walls(self):
        R = self.R
        return (self.NZl + 2 * self._gz(R), self.NY + 2 * R, self.NX + 2 * R)

    def shape_rho(self):
        return self.shape_walls() + (self.S,)

    def shape_fi(self):
        return (self.NZl + 2 * self._gz(1), self.NY + 2, self.NX + 2, self.Q, self.S)

    def shape_u(self):
        return (self.NZl + 2 * self._gz(1), self.NY + 2, self.NX + 2, self.D, self.S)

    def _expect(self, a, shape, what):
        if a.shape != shape:
            raise ValueError("%s: expected local ghosted shape %s, got %s" % (what, shape, a.shape))

    # ---- set-up (lbm.F90:140-175,183-253,444-453)
    def walls_set_values(self, walls_rg):
        w = _c(walls_rg)
        self._expect(w, self.shape_walls(), "walls")
        self._check(self.lib.txg_set_walls(self.h, _dp(w)))

    def shape_bc(self, boundary):
        """Local face array of boundary 0..5 (xm, xp, ym, yp, zm, zp) in the reference's layout
        (lbm_bc.F90:1100-1106): xm/xp (zl, NY, nbcs), ym/yp (zl, NX, nbcs), zm/zp (NY, NX, nbcs),
        nbcs = ndims * ncomponents viewed as (ndims, S) in C order."""
        n = {0: (self.NZl, self.NY), 1: (self.NZl, self.NX), 2: (self.NY, self.NX)}[boundary // 2]
        if self.D == 2:
            n = n[1:]
        return n + (self.D, self.S)

    def bc_set_values(self, boundary, vals):
        """BCSetValues (lbm_bc.F90:215-228) for one face."""
        v = _c(vals)
        self._expect(v, self.shape_bc(boundary), "bc values of boundary %d" % boundary)
        self._check(self.lib.txg_set_bc_values(self.h, int(boundary), _dp(v)))

    def bc_set_pressure_outlet(self, boundary, pressure):
        """flow%bc_flags(boundary) = BC_PRESSURE_OUTLET, flow%bc_data(1,boundary) = pressure (lbm_flow.F90:1170-1189)."""
        self._check(self.lib.txg_set_bc_pressure_outlet(self.h, int(boundary), float(pressure)))

    def initialize_state(self, rho_rg, u_g=None):
        r = _c(rho_rg)
        self._expect(r, self.shape_rho(), "rho")
        u = _c(u_g)
        if u is not None:
            self._expect(u, self.shape_u(), "u")
        self._check(self.lib.txg_set_rho_u(self.h, _dp(r), _dp(u)))

    def set_fi(self, fi_g):
        f = _c(fi_g)
        self._expect(f, self.shape_fi(), "fi")
        self._check(self.lib.txg_set_fi(self.h, _dp(f)))

    def initialize_state_restarted(self, prefix, counter):
        """LBMInitializeStateRestarted (lbm.F90:524-544): fi from <prefix>fiNNN.dat as FlowOutputDiagnostics wrote it."""
        from . import petsc_io

        self.initialize_state_from_file(petsc_io.output_name(prefix, "fi", counter), "fi")

    def initialize_state_from_file(self, path, kind="fi"):
        """LBMInitializeStateFromFile (lbm.F90:482-522): -ic_file (kind 'fi': populations, FlowFiInit is then skipped,
        lbm.F90:208-213) or -ic_file_rho (kind 'rho': densities, followed by the usual fi_init())."""
        from . import petsc_io

        if kind == "fi":
            self.set_fi(petsc_io.load_local(path, self.cfg, (self.Q, self.S), 1))
        elif kind == "rho":
            self.initialize_state(petsc_io.load_local(path, self.cfg, (self.S,), self.R))
        else:
            raise ValueError("kind must be 'fi' or 'rho'")

    def fi_init(self):
        self._check(self.lib.txg_fi_init(self.h))

    def update_moments(self):
        self._check(self.lib.txg_update_moments(self.h))

    # ---- time stepping (lbm.F90:262-422)
    def step(self, nsteps=1):
        self._check(self.lib.txg_step(self.h, int(nsteps)))

    def run(self, istep, kstep):
        """LBMRun2's loop `do lcv_step = istep+1, kstep`."""
        self.step(kstep - istep)

    def collision(self):
        self._check(self.lib.txg_collision(self.h))

    def communicate_fi(self):
        self._check(self.lib.txg_communicate_fi(self.h))

    def stream(self):
        self._check(self.lib.txg_stream(self.h))

    def bounceback(self):
        self._check(self.lib.txg_bounceback(self.h))

    def apply_bcs(self):
        self._check(self.lib.txg_apply_bcs(self.h))

    def update_flux(self):
        self._check(self.lib.txg_update_flux(self.h))

    def synchronize(self):
        self._check(self.lib.txg_synchronize(self.h))

    # ---- state out
    def get_fi(self, out=None):
        f = np.zeros(self.shape_fi()) if out is None else out
        self._check(self.lib.txg_get_fi(self.h, _dp(f)))
        return f

    def get_arrays(self, rho=True, u=True, forces=True):
        """FlowGetArrays view: (rho_rg, u_g, forces_g); ghost entries are left as passed (zero)."""
        r = np.zeros(self.shape_rho()) if rho else None
        uu = np.zeros(self.shape_u()) if u else None
        ff = np.zeros(self.shape_u()) if forces else None
        self._check(self.lib.txg_get_state(self.h, _dp(r), _dp(uu), _dp(ff)))
        return r, uu, ff

    def shape_diagnostics(self):
        n = (self.NZl, self.NY, self.NX)
        return n, n, n + (self.D,)

    def update_diagnostics(self, out=None):
        """FlowUpdateDiagnostics: (rhot[z,y,x], prs[z,y,x], velt[z,y,x,d]) owned only.  `out`: the caller's
        own (rhot, prs, velt) arrays -- the reference writes into its existing Vecs; page-locked
        arrays make the device-to-host copy run at the full PCIe rate."""
        if out is None:
            out = tuple(np.zeros(sh) for sh in self.shape_diagnostics())
        rhot, prs, velt = out
        for a, sh, nm in zip(out, self.shape_diagnostics(), ("rhot", "prs", "velt")):
            self._expect(a, sh, nm)
        self._check(self.lib.txg_get_diagnostics(self.h, _dp(rhot), _dp(prs), _dp(velt)))
        return rhot, prs, velt

    def output_diagnostics(self, prefix, counter, fi=True, rho=True, velt=True, rhot=True, prs=True):
        """FlowUpdateDiagnostics + FlowOutputDiagnostics (lbm_flow.F90:576-601, called from LBMOutput,
        lbm.F90:424-438): <prefix>{fi,rho,u,rhot,prs}NNN.dat as PETSc binary Vecs in DMDA natural ordering
        (petsc_io.py), so src/testing/check_solution.py and petsc2tec.py read them unchanged.  The velocity
        file is named `u` like the reference's (it holds velt).  With several ranks every rank writes its
        z-slab -- one contiguous byte range of the natural ordering -- into the same file; returns the paths.
        The file is complete once EVERY rank has returned: synchronise the ranks (a barrier) before reading it."""
        import os

        from . import geometry as geo
        from . import petsc_io

        D, S = self.D, self.S
        NZg = self.cfg.NZ if D == 3 else 1
        plane = self.NY * self.NX
        zs = self.cfg.zs if D == 3 else 0
        fields = {}
        if fi:
            fields["fi"] = geo.owned(self.get_fi(), 1, D)
        if rho:
            fields["rho"] = geo.owned(self.get_arrays(u=False, forces=False)[0], self.R, D)
        if velt or rhot or prs:
            rt, pr, vt = self.update_diagnostics()
            if velt:
                fields["u"] = vt
            if rhot:
                fields["rhot"] = rt
            if prs:
                fields["prs"] = pr
        paths = []
        for name, a in fields.items():
            a = np.ascontiguousarray(a, dtype=np.float64)
            dof = a.size // (self.NZl * plane)
            path = petsc_io.output_name(prefix, name, counter)
            fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
            try:
                # exact size of the PETSc Vec file: a longer file left under this name by another box size or dof
                # count would keep its stale tail (every rank sets the same length: idempotent, no rank order)
                os.ftruncate(fd, 8 + NZg * plane * dof * 8)
                if self.cfg.rank == 0:
                    os.pwrite(fd, np.array([petsc_io.VEC_CLASSID, NZg * plane * dof], dtype=">i4").tobytes(), 0)
                os.pwrite(fd, a.astype(">f8").tobytes(), 8 + zs * plane * dof * 8)
            finally:
                os.close(fd)
            paths.append(path)
        return paths

    def node_class(self):
        out = np.zeros(self.shape_walls(), dtype=np.uint8)
        self._check(self.lib.txg_get_node_class(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def delta_norm(self):
        v = C.c_double()
        self._check(self.lib.txg_delta_norm(self.h, C.byref(v)))
        return v.value

    # ---- measurement
    def last_step_ms(self):
        ms, n = C.c_float(), C.c_int64()
        self._check(self.lib.txg_last_step_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def enable_kernel_timing(self, on=True):
        self._check(self.lib.txg_enable_kernel_timing(self.h, int(on)))

    def reset_kernel_times(self):
        self._check(self.lib.txg_reset_kernel_times(self.h))

    def kernel_times(self):
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        ln = (C.c_int64 * cap)()
        n = C.c_int()
        self._check(self.lib.txg_kernel_times(self.h, cap, names, ms, ln, C.byref(n)))
        return {names[i].decode(): (ms[i], ln[i]) for i in range(n.value)}
