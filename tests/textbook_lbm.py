"""A second, independently written CPU model of the flow step -- TEST INFRASTRUCTURE ONLY.

Purpose: SURVEY.md 8c lists what no reference vector pins (D3Q19 MRT with general rates, the order-8/10
force stencils, bounce-back, mineral and body forces, the non-ideal EOS kinds).  For those the oracle (oracle/taxila_oracle.c)
is a line-by-line restatement of the Fortran; this file restates the same *equations* from the
literature in whole-array numpy form, sharing no code, table or loop structure with the oracle:

  * moment matrices built from the d'Humieres (D3Q19) / Lallemand-Luo (D2Q9) polynomials, collision as
    f -= M^-1 diag(s) M (f - feq_bar) with an explicit inverse (the reference: a table of rows and
    rank-1 updates with the row norms, lbm_relaxation.F90:182-200, lbm_discretization_d3q19.F90:177-197);
  * streaming + half-way bounce-back in pull form with np.roll (the reference: push into a temporary,
    then a sweep over wall nodes, lbm_distribution_function.F90:560-784);
  * the Shan-Chen gradient from a generated offset list and a generic line-of-sight rule (the
    reference: 92 + 36 hand-written blocks, lbm_forcing.F90:51-1299);
  * fluid-solid force from the mineral id of each lattice neighbour (lbm_forcing.F90:1326-1421, with its
    single-precision weights 1./6., 1./12., 1./3.).

Fully periodic boxes only (a non-periodic face is a ring of 999 nodes, with optional density / flux / velocity face BCs
on the nodes next to it); free-slip walls (900-902) only as whole planes.  Agreement with the oracle is to accumulated round-off (different summation
order), checked in tests/test_oracle_textbook.py."""
import itertools

import numpy as np

F32 = lambda x: float(np.float32(x))  # noqa: E731  a Fortran default-real literal


class Lattice:
    def __init__(self, ndims):
        if ndims == 2:
            # 0; E N W S; NE NW SW SE
            c = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
            w = [4 / 9] + [1 / 9] * 4 + [1 / 36] * 4
        else:
            # 0; E N W S U D; NE NW SW SE; EU WU WD ED; NU SU SD ND
            c = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
                 (1, 1, 0), (-1, 1, 0), (-1, -1, 0), (1, -1, 0),
                 (1, 0, 1), (-1, 0, 1), (-1, 0, -1), (1, 0, -1),
                 (0, 1, 1), (0, -1, 1), (0, -1, -1), (0, 1, -1)]
            w = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12
        self.D = ndims
        self.c = np.array(c, dtype=np.int64)
        self.w = np.array(w)
        self.Q = len(c)
        self.opp = np.array([c.index(tuple(-v for v in cc)) for cc in c])
        self.M = self._moments()

    def _moments(self):
        c = self.c.astype(np.float64)
        c2 = (c * c).sum(axis=1)
        one = np.ones(self.Q)
        if self.D == 2:
            x, y = c[:, 0], c[:, 1]
            rows = [one, -4 + 3 * c2, 4 - 10.5 * c2 + 4.5 * c2 * c2, x, (-5 + 3 * c2) * x, y, (-5 + 3 * c2) * y,
                    x * x - y * y, x * y]
        else:
            x, y, z = c[:, 0], c[:, 1], c[:, 2]
            rows = [one, 19 * c2 - 30, (21 * c2 * c2 - 53 * c2 + 24) / 2,
                    x, (5 * c2 - 9) * x, y, (5 * c2 - 9) * y, z, (5 * c2 - 9) * z,
                    3 * x * x - c2, (3 * c2 - 5) * (3 * x * x - c2), y * y - z * z, (3 * c2 - 5) * (y * y - z * z),
                    x * y, y * z, x * z, (y * y - z * z) * x, (z * z - x * x) * y, (x * x - y * y) * z]
        return np.array(rows)

    def rates(self, s):
        """relaxation rate of every moment row; s: dict s_c s_e s_e2 s_q s_nu s_pi s_m"""
        if self.D == 2:
            keys = ["s_c", "s_e", "s_e2", "s_c", "s_q", "s_c", "s_q", "s_nu", "s_nu"]
        else:
            keys = ["s_c", "s_e", "s_e2", "s_c", "s_q", "s_c", "s_q", "s_c", "s_q", "s_nu", "s_pi", "s_nu", "s_pi",
                    "s_nu", "s_nu", "s_nu", "s_m", "s_m", "s_m"]
        return np.array([s[k] for k in keys])


FFW = {
    (3, 4): {1: 1 / 6, 2: 1 / 12},
    (3, 8): {1: 4 / 45, 2: 1 / 21, 3: 2 / 105, 4: 5 / 504, 5: 1 / 315, 6: 1 / 630, 8: 1 / 5040},
    (2, 4): {1: 1 / 3, 2: 1 / 12},
    (2, 8): {1: 4 / 21, 2: 4 / 45, 4: 1 / 60, 5: 2 / 315, 8: 1 / 5040},
    (2, 10): {1: 262 / 1785, 2: 93 / 1190, 4: 7 / 340, 5: 6 / 595, 8: 9 / 9520, 9: 2 / 5355, 10: 1 / 7140},
}


def stencil(ndims, order):
    """[(offset, weight, sight)]: sight = list of alternatives, each a list of intermediate offsets that must
    all be fluid (empty list of alternatives: target only)."""
    wts = FFW[(ndims, order)]
    rmax = 3 if order == 10 else (2 if order == 8 else 1)
    out = []
    for off in itertools.product(range(-rmax, rmax + 1), repeat=ndims):
        L = sum(v * v for v in off)
        a = sorted((abs(v) for v in off), reverse=True)
        if L not in wts or L == 0:
            continue
        if L == 9 and a[0] != 3:
            continue  # (2,2,1) is not part of any stencil
        sgn = [int(np.sign(v)) for v in off]
        alts = []
        if a[0] == 2 and (len(a) < 2 or a[1] <= 1):
            # one long axis: half a step along it, plus any subset of the unit components
            long_ax = [i for i, v in enumerate(off) if abs(v) == 2][0]
            units = [i for i, v in enumerate(off) if abs(v) == 1]
            for k in range(len(units) + 1):
                for sub in itertools.combinations(units, k):
                    mid = [0] * ndims
                    mid[long_ax] = sgn[long_ax]
                    for i in sub:
                        mid[i] = sgn[i]
                    alts.append([tuple(mid)])
        elif a[0] == 2 and a[1] == 2:
            alts.append([tuple(sgn[i] if abs(off[i]) == 2 else 0 for i in range(ndims))])
        elif a[0] == 3:
            long_ax = [i for i, v in enumerate(off) if abs(v) == 3][0]
            units = [i for i, v in enumerate(off) if abs(v) == 1]
            for side in ([0] if not units else [0, 1]):
                path = []
                for step in (1, 2):
                    mid = [0] * ndims
                    mid[long_ax] = step * sgn[long_ax]
                    if side:
                        mid[units[0]] = sgn[units[0]]
                    path.append(tuple(mid))
                alts.append(path)
        out.append((off, wts[L], alts))
    return out


def shift(a, off):
    """a(X + off) on the periodic box; a is [..., z, y, x] (3-D) or [..., y, x] (2-D); off is (dx, dy[, dz])"""
    axes = tuple(range(a.ndim - len(off), a.ndim))
    return np.roll(a, tuple(-v for v in reversed(off)), axis=axes)


class Model:
    """p: dict with ndims, mrt (bool), tau[S] or rate dicts s[S], mm[S], gf[S][S], gw[nminerals][S], gvt or None,
    order.  walls: [z,y,x] / [y,x] array of node codes (0 pore, k mineral id)."""

    def __init__(self, p, walls, rho0):
        self.p = p
        self.lat = lat = Lattice(p["ndims"])
        self.S = S = len(p["mm"])
        self.walls = np.asarray(walls)
        self.fluid = self.walls == 0
        self.mm = np.array(p["mm"], dtype=np.float64)
        self.d_k = 1 - 2 / (3 * self.mm)
        if p["mrt"]:
            Minv = np.linalg.inv(lat.M)
            self.omega = [Minv @ np.diag(lat.rates(p["s"][m])) @ lat.M for m in range(S)]
            self.s_c = np.array([p["s"][m]["s_c"] for m in range(S)])
        else:
            self.omega = [np.eye(lat.Q) / p["tau"][m] for m in range(S)]
            self.s_c = np.array([1 / p["tau"][m] for m in range(S)])
        self.sten = stencil(p["ndims"], p["order"])
        self.rho = np.where(self.fluid, np.asarray(rho0, dtype=np.float64), 0.0)  # [S, (z,) y, x]
        self.F = self.forces(self.rho)
        zero = np.zeros((lat.D,) + self.fluid.shape)
        self.f = (1 - 0.5 * self.prefactor(self.rho, self.F, zero)) * self.feq(self.rho, zero)
        self.f *= self.fluid
        self.moments(apply_bcs=False)  # LBMInit2: FlowUpdateMoments, no FlowApplyBCs

    # -- pieces
    def feq(self, rho, u):
        lat = self.lat
        usq = (u * u).sum(axis=0)
        out = np.empty((self.S, lat.Q) + self.fluid.shape)
        for m in range(self.S):
            if lat.D == 3:
                out[m, 0] = rho[m] * (self.d_k[m] - usq / 2)
            else:
                out[m, 0] = rho[m] * ((1 + 5 * self.d_k[m]) / 6 - 2 * usq / 3)
            for n in range(1, lat.Q):
                cu = sum(lat.c[n, d] * u[d] for d in range(lat.D))
                out[m, n] = lat.w[n] * rho[m] * (1.5 * (1 - self.d_k[m]) + 3 * cu + 4.5 * cu * cu - 1.5 * usq)
        return out

    def prefactor(self, rho, F, u):
        lat = self.lat
        out = np.zeros((self.S, lat.Q) + self.fluid.shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            for m in range(self.S):
                for n in range(lat.Q):
                    acc = sum(F[m, d] * (lat.c[n, d] - u[d]) for d in range(lat.D))
                    out[m, n] = np.where(self.fluid, acc / (rho[m] / 3), 0.0)
        return out

    def forces(self, rho):
        lat, p = self.lat, self.p
        F = np.zeros((self.S, lat.D) + self.fluid.shape)
        gw = np.array(p.get("gw", []), dtype=np.float64)
        if gw.size and np.abs(gw).max() > 1e-15:
            wa, wd = (F32(1.0) / F32(6.0), F32(1.0) / F32(12.0)) if lat.D == 3 else (F32(1.0) / F32(3.0), F32(1.0) / F32(12.0))
            wa, wd = F32(wa), F32(wd)
            for n in range(1, lat.Q):
                code = shift(self.walls, tuple(lat.c[n]))
                wgt = wa if (lat.c[n] ** 2).sum() == 1 else wd
                for k in range(gw.shape[0]):
                    hit = code == (k + 1)
                    for m in range(self.S):
                        for d in range(lat.D):
                            if lat.c[n, d]:
                                F[m, d] -= np.where(hit, wgt * rho[m] * gw[k, m] * lat.c[n, d], 0.0)
        if p.get("gvt") is not None:
            for m in range(self.S):
                for d in range(lat.D):
                    F[m, d] += p["gvt"][d] * self.mm[m] * rho[m]
        gf = np.array(p["gf"], dtype=np.float64)
        if np.abs(gf).max() > 0:
            psi = self.psi(rho, gf)  # the fluid-fluid term alone sees psi(rho) (identity without -flow_use_nonideal_eos)
            G = np.zeros((self.S, lat.D) + self.fluid.shape)
            W = np.zeros((lat.D,) + self.fluid.shape)
            for off, wgt, alts in self.sten:
                ok = shift(self.fluid, off)
                if alts:
                    anyalt = np.zeros_like(ok)
                    for path in alts:
                        allp = np.ones_like(ok)
                        for mid in path:
                            allp &= shift(self.fluid, mid)
                        anyalt |= allp
                    ok = ok & anyalt
                for d in range(lat.D):
                    if off[d]:
                        W[d] += np.where(ok, wgt * off[d] * off[d], 0.0)
                        for m in range(self.S):
                            G[m, d] += np.where(ok, wgt * off[d] * (shift(psi[m], off) - psi[m]), 0.0)
            with np.errstate(divide="ignore", invalid="ignore"):
                for d in range(lat.D):
                    on = W[d] > 1e-12
                    for m in range(self.S):
                        acc = sum(gf[m, k] * np.where(on, G[k, d] / W[d], 0.0) for k in range(self.S))
                        F[m, d] -= 6.0 * psi[m] * acc
        return F * self.fluid

    def psi(self, rho, gf):
        """Shan-Chen '93 / '94 and Peng-Robinson pseudo-potentials (Yuan & Schaefer 2006 form
        psi = sqrt(2 (p_EOS - rho c_s^2) / (c_0 g_mm))); p["eos"][m] = None | ("sc", rho0) | ("thermo", psi0, rho0) |
        ("pr", a, b, R, T, Tc, omega)."""
        eos = self.p.get("eos")
        if not eos:
            return rho
        out = np.array(rho, dtype=np.float64, copy=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            for m, e in enumerate(eos):
                r = rho[m]
                if e is None:
                    continue
                if e[0] == "sc":
                    out[m] = e[1] * (1 - np.exp(-r / e[1]))
                elif e[0] == "thermo":
                    out[m] = np.where(r > 0, e[1] * np.exp(-e[2] / np.where(r > 0, r, 1.0)), 0.0)
                elif e[0] == "pr":
                    _, a, b, R, T, Tc, om = e
                    kappa = F32(0.37464) + F32(1.54226) * om - F32(0.26992) * om ** 2  # default-real literals in the reference
                    alpha = (1 + kappa * (1 - np.sqrt(T / Tc))) ** 2
                    p_eos = r * R * T / (1 - b * r) - a * alpha * r ** 2 / (1 + 2 * b * r - (b * r) ** 2)
                    out[m] = np.sqrt(np.maximum(2 * (p_eos - r / 3) / (6.0 * gf[m, m]), 0.0))
        return out

    def moments(self, apply_bcs=True):
        lat = self.lat
        outlets = self.p.get("outlets") if apply_bcs else None  # {face index: pressure}: bc_pressure_outlet faces
        targets = {}
        if outlets:
            # FlowUpdateBCPressureOutlet: the densities prescribed on the face follow the phase fraction found (in the
            # densities of the previous step) one node inside, at the given pressure p = (r1 + r2)/3 + c_0 g_21 r1 r2
            g21 = self.p["gf"][1][0]
            for i, press in outlets.items():
                kind, axis, side, mask, vals = self.p["faces"][i]
                step = [0] * lat.D
                step[axis] = 1 if side == 0 else -1
                inside_fluid = shift(self.fluid, tuple(step))
                r1 = np.where(inside_fluid, shift(self.rho[0], tuple(step)), self.rho[0])
                r2 = np.where(inside_fluid, shift(self.rho[1], tuple(step)), self.rho[1])
                with np.errstate(divide="ignore", invalid="ignore"):
                    frac = r1 / (r1 + r2)
                    alpha = 1 / (1 / frac - 1)
                    a = (1 + alpha) / 3
                    t1 = (-a + np.sqrt(a * a + 4 * 6.0 * g21 * alpha * press)) / (2 * 6.0 * g21)
                    t2 = t1 / alpha
                t1 = np.where(frac < 1e-10, 0.0, np.where(frac > 1 - 1e-10, 3 * press, t1))
                t2 = np.where(frac < 1e-10, 3 * press, np.where(frac > 1 - 1e-10, 0.0, t2))
                targets[i] = [t1, t2]
        self.rho = self.f.sum(axis=1) * self.fluid
        # external face BCs: [(kind, axis, side, mask, vals[D][S])], kind "dirichlet" (vals[0][m] = density), "neumann"
        # (vals[d][m] = momentum) or "velocity" (vals[d][0] = velocity); given in BCApply's order (by kind, then xm .. zp)
        faces = self.p.get("faces") if apply_bcs else None
        if faces:
            for i, (kind, axis, side, mask, vals) in enumerate(faces):  # BCApplyDirichletToRho: the stencil sees the prescribed densities
                if kind == "dirichlet":
                    for m in range(self.S):
                        self.rho[m] = np.where(mask & self.fluid, targets[i][m] if i in targets else vals[0][m], self.rho[m])
        self.F = self.forces(self.rho)
        if faces:
            # Chang, Liu & Lin (2009) as the reference applies it (BCApply{Dirichlet,Neumann,Velocity}Node): the unknown
            # (incoming) populations of a face node get w_n c_n.Q with Q chosen so that the density, the momentum (+ F/2)
            # or the velocity is met; an edge node of two faces is corrected twice
            for i, (kind, axis, side, mask, vals) in enumerate(faces):
                inward = 1 if side == 0 else -1
                inc = [n for n in range(1, lat.Q) if lat.c[n][axis] == inward]
                wsum = [sum(lat.w[n] for n in inc if lat.c[n][d] != 0) for d in range(lat.D)]
                on = mask & self.fluid
                for m in range(self.S):
                    fm = self.f[m]
                    tot = fm.sum(axis=0)
                    mom = [sum(fm[n] * lat.c[n][d] for n in range(1, lat.Q)) for d in range(lat.D)]
                    Q = [None] * lat.D
                    if kind == "dirichlet":
                        for d in range(lat.D):
                            want = targets[i][m] if i in targets else vals[0][m]
                            Q[d] = inward * (want - tot) / wsum[d] if d == axis else -mom[d] / wsum[d]
                    elif kind == "neumann":
                        for d in range(lat.D):
                            Q[d] = (vals[d][m] - self.F[m, d] / 2 - mom[d]) / wsum[d]
                    else:
                        ua = vals[axis][0]
                        Q[axis] = (tot * ua - mom[axis] - self.F[m, axis] / 2) / (1 - inward * ua) / wsum[axis]
                        rho_new = tot + wsum[axis] * Q[axis] * inward
                        for d in range(lat.D):
                            if d != axis:
                                Q[d] = (rho_new * vals[d][0] - self.F[m, d] / 2 - mom[d]) / wsum[d]
                    for n in inc:
                        fm[n] = np.where(on, fm[n] + lat.w[n] * sum(lat.c[n][d] * Q[d] for d in range(lat.D)), fm[n])
            for kind, axis, side, mask, vals in faces:  # BCUpdateRho
                self.rho = np.where(mask & self.fluid, self.f.sum(axis=1), self.rho)
        j = np.einsum("mn...,nd->md...", self.f, lat.c.astype(np.float64))
        ue = j + 0.5 * self.F
        wgt = (self.mm * self.s_c).reshape((self.S,) + (1,) * self.fluid.ndim)
        den = (self.rho * wgt).sum(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.u = np.where(self.fluid, (ue * wgt[:, None]).sum(axis=0) / den, 0.0)

    def step(self, n=1):
        lat = self.lat
        for _ in range(n):
            feq = self.feq(self.rho, self.u)
            pref = self.prefactor(self.rho, self.F, self.u)
            d = self.f - (1 - 0.5 * pref) * feq
            post = np.empty_like(self.f)
            for m in range(self.S):
                post[m] = self.f[m] - np.einsum("ab,b...->a...", self.omega[m], d[m]) + pref[m] * feq[m]
            post *= self.fluid
            new = np.empty_like(post)
            for k in range(lat.Q):
                back = tuple(-lat.c[k])
                src_fluid = shift(self.fluid, back)
                val = np.where(src_fluid, shift(post[:, k], back), post[:, lat.opp[k]])
                # planar free-slip walls (codes 900 + a, normal along axis a): the population arriving along c_k off such a
                # wall left the node one tangential step back with the normal component of its velocity reversed
                src_code = shift(self.walls, back)
                for a in range(lat.D):
                    if lat.c[k][a] == 0:
                        continue
                    mirrored = lat.c[k].copy()
                    mirrored[a] = -mirrored[a]
                    n = [tuple(v) for v in lat.c].index(tuple(mirrored))
                    tang = lat.c[k].copy()
                    tang[a] = 0
                    val = np.where(src_code == 900 + a, shift(post[:, n], tuple(-tang)), val)
                new[:, k] = val
            self.f = new * self.fluid
            self.moments()

    # -- the oracle's array conventions ([z][y][x][Q][S] etc.)
    def fi_natural(self):
        f = self.f if self.lat.D == 3 else self.f[:, :, None]
        return np.ascontiguousarray(np.moveaxis(f, (0, 1), (4, 3)))

    def rho_natural(self):
        r = self.rho if self.lat.D == 3 else self.rho[:, None]
        return np.ascontiguousarray(np.moveaxis(r, 0, 3))

    def u_natural(self):
        u = self.u if self.lat.D == 3 else self.u[:, None]
        return np.ascontiguousarray(np.moveaxis(u, 0, 3))

    def forces_natural(self):
        F = self.F if self.lat.D == 3 else self.F[:, :, None]
        return np.ascontiguousarray(np.moveaxis(F, (0, 1), (4, 3)))


def from_config(cfg, walls, rho):
    """Build the model from a TxgConfig + natural-order arrays (walls [z][y][x], rho [z][y][x][S])."""
    D, S = cfg.ndims, cfg.ncomponents
    keys = ["s_c", "s_e", "s_e2", "s_q", "s_nu", "s_pi", "s_m"]
    p = dict(ndims=D, mrt=bool(cfg.relaxation_mode), order=cfg.isotropy_order,
             tau=[cfg.tau[m] for m in range(S)], s=[{k: getattr(cfg, k)[m] for k in keys} for m in range(S)],
             mm=[cfg.mm[m] for m in range(S)], gf=[[cfg.gf[m][k] for k in range(S)] for m in range(S)],
             gw=[[cfg.gw[k][m] for m in range(S)] for k in range(cfg.nminerals)],
             gvt=[cfg.gvt[d] for d in range(D)] if cfg.body_forces else None)
    if cfg.use_nonideal_eos:
        kinds = {1: lambda m: None, 2: lambda m: ("sc", cfg.eos_rho0[m]), 4: lambda m: ("thermo", cfg.eos_psi0[m], cfg.eos_rho0[m]),
                 3: lambda m: ("pr", cfg.eos_pr_a[m], cfg.eos_pr_b[m], cfg.eos_pr_R[m], cfg.eos_pr_T[m], cfg.eos_pr_Tc[m],
                               cfg.eos_pr_omega[m])}
        p["eos"] = [kinds[cfg.eos_type[m]](m) for m in range(S)]
    w = np.asarray(walls)
    r = np.moveaxis(np.asarray(rho, dtype=np.float64), -1, 0)
    if D == 2:
        w = w.reshape(w.shape[-2:])
        r = r.reshape((S,) + w.shape)
    return Model(p, w, r)
