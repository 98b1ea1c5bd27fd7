"""ctypes mirror of `txg_config` (include/taxila_gpu.h) plus the reference's option
defaults and its `input_data` options-file syntax.

The reference reads a PETSc options file given as argv[1] (src/lbm/main.F90:45-46):
`-key value` lines, `#` comments (tests/bubble_2D/input_data).  `parse_options_file`
reads the same syntax and `config_from_options` applies the same keys the
reference's *SetFromOptions procedures read for the flow hot path:
  lbm_info.F90:116-238        -NX -NY -NZ -bc_periodic_{x,y,z} -stencil_size_rho
  lbm_options.F90:163-365     -ncomponents -nminerals -discretization
                              -flow_relaxation_mode -isotropy_order/-derivative_order
                              -component<i>_name -mineral<i>_name -flow_use_nonideal_eos
  lbm_relaxation.F90:117-151  -tau_<c> | -s_c_<c> -s_e_<c> -s_e2_<c> -s_q_<c> -s_nu_<c> -s_pi_<c> -s_m_<c>
  lbm_component.F90:131-176   -mm_<c> -g_<i><j>
  lbm_mineral.F90:103-134     -gw_<mineral>_<component>
  lbm_flow.F90:226-239        -gvt
"""
import ctypes as C

NMAX_COMPONENTS = 5
MAX_MINERALS = 100

D3Q19_DISCRETIZATION = 1
D2Q9_DISCRETIZATION = 2
RELAXATION_MODE_SRT = 0
RELAXATION_MODE_MRT = 1
EOS_NULL, EOS_DENSITY, EOS_SC, EOS_PR, EOS_THERMO = 0, 1, 2, 3, 4

BC_NULL, BC_PERIODIC, BC_REFLECTING, BC_DIRICHLET, BC_NEUMANN, BC_VELOCITY = 0, 1, 2, 3, 4, 5
BOUNDARY_XM, BOUNDARY_XP, BOUNDARY_YM, BOUNDARY_YP, BOUNDARY_ZM, BOUNDARY_ZP = range(6)

WALL_PORESPACE = 0.0
WALL_NONREACTIVE = 800.0
WALL_NORMAL_X = 900.0
WALL_NORMAL_Y = 901.0
WALL_NORMAL_Z = 902.0
WALL_GHOST = 999.0


class TxgConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("ndims", C.c_int32),
        ("discretization", C.c_int32),
        ("ncomponents", C.c_int32),
        ("NX", C.c_int32),
        ("NY", C.c_int32),
        ("NZ", C.c_int32),
        ("zs", C.c_int32),
        ("zl", C.c_int32),
        ("periodic", C.c_int32 * 3),
        ("stencil_size_rho", C.c_int32),
        ("relaxation_mode", C.c_int32),
        ("isotropy_order", C.c_int32),
        ("nminerals", C.c_int32),
        ("fluidfluid_forces", C.c_int32),
        ("fluidsolid_forces", C.c_int32),
        ("body_forces", C.c_int32),
        ("use_nonideal_eos", C.c_int32),
        ("eos_type", C.c_int32 * NMAX_COMPONENTS),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
        ("bc_flags", C.c_int32 * 6),
        ("tau", C.c_double * NMAX_COMPONENTS),
        ("s_c", C.c_double * NMAX_COMPONENTS),
        ("s_e", C.c_double * NMAX_COMPONENTS),
        ("s_e2", C.c_double * NMAX_COMPONENTS),
        ("s_q", C.c_double * NMAX_COMPONENTS),
        ("s_nu", C.c_double * NMAX_COMPONENTS),
        ("s_pi", C.c_double * NMAX_COMPONENTS),
        ("s_m", C.c_double * NMAX_COMPONENTS),
        ("mm", C.c_double * NMAX_COMPONENTS),
        ("gf", (C.c_double * NMAX_COMPONENTS) * NMAX_COMPONENTS),
        ("eos_rho0", C.c_double * NMAX_COMPONENTS),
        ("gw", (C.c_double * NMAX_COMPONENTS) * MAX_MINERALS),
        ("gvt", C.c_double * 3),
        ("null_pressure", C.c_double),
        ("reserved_d", C.c_double * 8),
        ("eos_psi0", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_a", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_b", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_R", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_T", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_Tc", C.c_double * NMAX_COMPONENTS),
        ("eos_pr_omega", C.c_double * NMAX_COMPONENTS),
    ]

    def copy(self):
        out = TxgConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(TxgConfig))
        return out

    @property
    def Q(self):
        return 19 if self.discretization == D3Q19_DISCRETIZATION else 9

    @property
    def owned_shape(self):
        """(NZ_local, NY, NX) of the owned slab; NZ_local = 1 in 2-D."""
        return (self.zl if self.ndims == 3 else 1, self.NY, self.NX)


def stencil_size_rho_for(isotropy_order):
    """lbm_grid.F90:107-120"""
    return {0: 0, 4: 1, 8: 2, 10: 3}.get(isotropy_order, 1)


def default_config(ndims=3, ncomponents=2, NX=1, NY=1, NZ=1):
    """Reference defaults (lbm_options.F90:93-152, lbm_relaxation.F90:60-83)."""
    c = TxgConfig()
    c.struct_bytes = C.sizeof(TxgConfig)
    c.ndims = ndims
    c.discretization = D3Q19_DISCRETIZATION if ndims == 3 else D2Q9_DISCRETIZATION
    c.ncomponents = ncomponents
    c.NX, c.NY, c.NZ = NX, NY, (NZ if ndims == 3 else 1)
    c.zs, c.zl = 0, c.NZ
    c.isotropy_order = 4
    c.stencil_size_rho = 1
    c.relaxation_mode = RELAXATION_MODE_SRT
    c.nminerals = 1
    c.rank, c.nranks = 0, 1
    for m in range(NMAX_COMPONENTS):
        c.tau[m] = c.s_c[m] = c.s_e[m] = c.s_e2[m] = c.s_q[m] = 1.0
        c.s_nu[m] = c.s_pi[m] = c.s_m[m] = 1.0
        c.mm[m] = 1.0
        c.eos_rho0[m] = 1.0
        c.eos_type[m] = EOS_DENSITY
        set_eos_pr(c, m)
        c.eos_psi0[m] = 1.0
    c.null_pressure = 0.0
    return c


def set_eos_pr(c, m, a=2.0 / 49.0, b=2.0 / 21.0, R=1.0, T=None, reduced_T=None, omega=None):
    """EOSSetFromOptions_PR (lbm_eos.F90:272-320).  0.0778, 0.45724, 0.9 and 0.344 are default-real
    literals there: their single-precision values enter the double arithmetic."""
    import numpy as np

    f32 = lambda v: float(np.float32(v))  # noqa: E731
    c.eos_pr_a[m], c.eos_pr_b[m], c.eos_pr_R[m] = a, b, R
    Tc = a / b * f32(0.0778) / f32(0.45724) / R
    c.eos_pr_Tc[m] = Tc
    if T is None:
        T = (f32(0.9) if reduced_T is None else reduced_T) * Tc
    c.eos_pr_T[m] = T
    c.eos_pr_omega[m] = f32(0.344) if omega is None else omega


def finalize_flags(c):
    """Derive the flags the reference derives while parsing options:
    fluidfluid_forces if any |g| > 1e-15 (lbm_flow.F90:233-239), fluidsolid_forces
    if any |gw| > 1e-15 (lbm_walls.F90:117-125)."""
    eps = 1.0e-15
    S = c.ncomponents
    c.fluidfluid_forces = int(any(abs(c.gf[m][k]) > eps for m in range(S) for k in range(S)))
    c.fluidsolid_forces = int(any(abs(c.gw[k][m]) > eps for k in range(c.nminerals) for m in range(S)))
    c.stencil_size_rho = max(1, stencil_size_rho_for(c.isotropy_order))
    return c


def parse_options_file(path):
    """PETSc options-file syntax: `-key [value]`, `#` starts a comment."""
    opts = {}
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line or not line.startswith("-"):
                continue
            parts = line.split(None, 1)
            opts[parts[0][1:]] = parts[1].strip() if len(parts) > 1 else ""
    return opts


def _flag(opts, key):
    if key not in opts:
        return False
    return opts[key].lower() not in ("0", "false", "no")


def config_from_options(opts):
    """Build a TxgConfig from parsed reference options.  Returns (config, names)
    where names = {'components': [...], 'minerals': [...]}."""
    disc = opts.get("discretization", "d3q19").lower()
    ndims = 3 if disc == "d3q19" else 2
    S = int(opts.get("ncomponents", 1))
    c = default_config(ndims, S, int(opts.get("NX", 1)), int(opts.get("NY", 1)), int(opts.get("NZ", 1)))
    c.periodic[0] = int(_flag(opts, "bc_periodic_x"))
    c.periodic[1] = int(_flag(opts, "bc_periodic_y"))
    c.periodic[2] = int(_flag(opts, "bc_periodic_z")) if ndims == 3 else 0
    c.relaxation_mode = int(opts.get("flow_relaxation_mode", 0))
    # -isotropy_order and -derivative_order set the same field (lbm_options.F90:293-299)
    if "isotropy_order" in opts:
        c.isotropy_order = int(opts["isotropy_order"])
    if "derivative_order" in opts:
        c.isotropy_order = int(opts["derivative_order"])
    c.nminerals = int(opts.get("nminerals", 1))
    c.use_nonideal_eos = int(_flag(opts, "flow_use_nonideal_eos"))
    comp_names = [opts.get("component%d_name" % (m + 1), "component%d" % (m + 1)) for m in range(S)]
    min_names = [opts.get("mineral%d_name" % (k + 1), "mineral%d" % (k + 1)) for k in range(c.nminerals)]
    for m, name in enumerate(comp_names):
        if c.relaxation_mode == RELAXATION_MODE_SRT:
            c.tau[m] = float(opts.get("tau_" + name, 1.0))
            c.s_c[m] = 1.0 / c.tau[m]
        else:
            for key in ("s_c", "s_e", "s_e2", "s_q", "s_nu", "s_pi", "s_m"):
                getattr(c, key)[m] = float(opts.get("%s_%s" % (key, name), 1.0))
        c.mm[m] = float(opts.get("mm_" + name, 1.0))
        for k in range(S):
            c.gf[m][k] = float(opts.get("g_%d%d" % (m + 1, k + 1), 0.0))
    for k, mname in enumerate(min_names):
        for m, cname in enumerate(comp_names):
            c.gw[k][m] = float(opts.get("gw_%s_%s" % (mname, cname), 0.0))
    if "gvt" in opts:
        c.body_forces = 1
        vals = [float(v) for v in opts["gvt"].split(",")]
        for d, v in enumerate(vals[:3]):
            c.gvt[d] = v
    finalize_flags(c)
    if "stencil_size_rho" in opts:
        c.stencil_size_rho = int(opts["stencil_size_rho"])
    return c, {"components": comp_names, "minerals": min_names}
