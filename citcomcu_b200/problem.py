"""Host-side problem setup from a reference-format input file (Cartesian box).

Mirrors, for one subdomain of the reference's `nprocx x nprocy x nprocz` block decomposition,
what `read_instructions` (Instructions.c:55) leaves in `struct All_variables` and the hot path
reads:

* per-level mesh sizes                      (Instructions.c:430-575 `global_derived_values`)
* node coordinates `XX[lev][1..3]`          (Nodal_mesh.c:53 `node_locations`: float accumulation
                                             `XX[i] = XX[i-1] + dx`, coarser levels by injection)
* `NODE[lev]` boundary-condition flag bits  (Boundary_conditions.c:44-170 `velocity_boundary_conditions`,
                                             :259 `velocity_refl_vert_bc`, :177 `temperature_boundary_conditions`,
                                             :516 `temperature_refl_vert_bc`)
* initial temperature                       (Convection.c:325-350 `convection_initial_temperature`)
* material groups `mat`                     (Construct_arrays.c:742 `construct_mat_group`, Viscosity_structures.c:1389 `layers`)
* buoyancy                                  (Pan_problem_misc_functions.c:83 `thermal_buoyancy`,
                                             Global_operations.c:55,133 `remove_horiz_ave`)

Coordinates, flags, temperature and material groups are bit-exact against the reference's own
arrays (tests/test_problem_setup.py compares them with oracle dumps).  Only the subset of the
input language the BASELINE configurations use is honoured: cart3d, reflecting side walls,
free-slip or no-slip top/bottom, fixed-temperature top/bottom.
"""
from __future__ import annotations

import numpy as np

# flag bits, global_defs.h:65-89
LIDN, VBX, VBZ, VBY = 0x1, 0x2, 0x4, 0x8
TBX, TBZ, TBY = 0x10, 0x20, 0x40
INTX, INTZ, INTY = 0x2000, 0x4000, 0x8000
SBX, SBZ, SBY = 0x10000, 0x20000, 0x40000
FBX, FBZ, FBY = 0x80000, 0x100000, 0x200000
OFFSIDE, SIDEE = 0x400000, 0x800000
BC_MASK = VBX | VBZ | VBY | TBX | TBZ | TBY | SBX | SBZ | SBY | FBX | FBZ | FBY

f32 = np.float32


def parse_input(text: str) -> dict:
    """`name=value` lines (Parsing.c:101 setup_parser); '#' starts a comment; later keys win."""
    out = {}
    for line in text.split("\n"):
        line = line.split("#", 1)[0].strip()
        if "=" not in line:
            continue
        k, v = line.split("=", 1)
        out[k.strip()] = v.strip()
    return out


def _on(v) -> bool:
    return str(v).strip().lower() in ("on", "1", "true", "yes")


def _fvec(s, n=None):
    v = [f32(float(x)) for x in str(s).split(",") if x.strip()]
    return v if n is None else v[:n]


def _ivec(s):
    return [int(x) for x in str(s).split(",") if x.strip()]


class CartesianProblem:
    """One rank's view of a cart3d input file.  `me_loc` = (x, y, z) position in the processor grid;
    rank = z + nprocz*x + nprocz*nprocx*y (Parallel_related.c:108-121)."""
    GEOMETRY = "cart3d"

    def __init__(self, text: str, me_loc=(0, 0, 0)):
        self.text = text
        p = self.params = parse_input(text)
        g = lambda k, d=None: p.get(k, d)  # noqa: E731
        if g("Geometry", "cart3d") != self.GEOMETRY:
            raise ValueError(f"{type(self).__name__}: Geometry={self.GEOMETRY} expected")
        self.nproc = (int(g("nprocx", 1)), int(g("nprocy", 1)), int(g("nprocz", 1)))
        self.me_loc = tuple(me_loc)
        self.levels = int(g("levels"))
        self.levmin, self.levmax = 0, self.levels - 1
        mg = (int(g("mgunitx")), int(g("mgunity")), int(g("mgunitz")))
        for m, n in zip(mg, self.nproc):
            if m % n:
                raise ValueError("mgunit must be divisible by nproc in each direction (README:149-151)")
        self.mgunit = mg
        # global / local sizes per level
        self.NOX, self.NOY, self.NOZ = {}, {}, {}         # global nodes
        self.nox, self.noy, self.noz = {}, {}, {}         # local nodes
        self.NXS, self.NYS, self.NZS = {}, {}, {}         # 1-based global index of the first local node
        for lev in range(self.levels):
            f = 2 ** lev
            gx, gy, gz = mg[0] * f, mg[1] * f, mg[2] * f   # global elements
            self.NOX[lev], self.NOY[lev], self.NOZ[lev] = gx + 1, gy + 1, gz + 1
            lx, ly, lz = gx // self.nproc[0], gy // self.nproc[1], gz // self.nproc[2]
            self.nox[lev], self.noy[lev], self.noz[lev] = lx + 1, ly + 1, lz + 1
            self.NXS[lev] = self.me_loc[0] * lx + 1
            self.NYS[lev] = self.me_loc[1] * ly + 1
            self.NZS[lev] = self.me_loc[2] * lz + 1
        self.layer = (f32(float(g("dimenx", 1.0))), f32(float(g("dimeny", 1.0))), f32(float(g("dimenz", 1.0))))
        # solver controls (Instructions.c:1034-1060)
        self.control = dict(
            v_steps_low=int(g("vlowstep", 20)), v_steps_high=int(g("vhighstep", 3)), down_heavy=int(g("down_heavy", 1)),
            up_heavy=int(g("up_heavy", 1)), mg_cycle=int(g("mg_cycle", 1)), p_iterations=int(g("piterations", 100)),
            accuracy=float(g("accuracy", 1.0e-4)), augmented_Lagr=int(_on(g("aug_lagr", "off"))),
            augmented=float(g("aug_number", 0.0)), precondition=int(_on(g("precond", "off"))))
        self.rayleigh = f32(float(g("rayleigh")))
        self.topvbc, self.botvbc = int(g("topvbc", 0)), int(g("botvbc", 0))
        # imposed boundary velocities (Instructions.c:980-992; the top x value is plate_velocity, Boundary_conditions.c:92)
        self.plate_vel, self.topvby = f32(float(g("plate_velocity", 0.0))), f32(float(g("topvbyval", 0.0)))
        self.botvbx, self.botvby = f32(float(g("botvbxval", 0.0))), f32(float(g("botvbyval", 0.0)))
        self.toptbc, self.bottbc = int(g("toptbc", 1)), int(g("bottbc", 1))
        self.toptbcval, self.bottbcval = f32(float(g("toptbcval", 0.0))), f32(float(g("bottbcval", 1.0)))
        if _on(g("periodicx", "off")) or _on(g("periodicy", "off")):
            raise ValueError("CartesianProblem: periodic side walls are not mirrored")
        # viscosity law (Viscosity_structures.c:57-326)
        self.num_mat = int(g("num_mat", 1))
        self.visc = dict(
            tdepv=int(_on(g("TDEPV", "off"))), rheol=int(g("rheol", 0)),
            N0=_fvec(g("visc0", "1"), self.num_mat), E=_fvec(g("viscE", "0"), self.num_mat),
            T=_fvec(g("viscT", "0"), self.num_mat), Z=_fvec(g("viscZ", "0"), self.num_mat),
            vmin=int(_on(g("VMIN", "off"))), min_value=float(g("visc_min", 0.0)),
            vmax=int(_on(g("VMAX", "off"))), max_value=float(g("visc_max", 0.0)),
            smooth_cycles=int(g("visc_smooth_cycles", 0)),
            # stress-dependent viscosity (Viscosity_structures.c:150-310)
            sdepv=int(_on(g("SDEPV", "off"))), sdepv_rheology=int(g("sdepv_rheology", 2)),
            sdepv_expt=_fvec(g("sdepv_expt", "1"), self.num_mat), sdepv_trns=_fvec(g("sdepv_trns", "1"), self.num_mat),
            sdepv_misfit=float(g("sdepv_misfit", 0.001)), sdepv_iter_damp=float(g("sdepv_iter_damp", 1.0)),
            sdepv_max_iter=int(g("max_sdep_visc_iter", 50)), sdepv_start_from_newtonian=int(_on(g("sdepv_start_from_newtonian", "off"))),
            sdepv_trns_T=float(g("sdepv_trns_T", 3000.0)), sdepv_trns_c=float(g("sdepv_trns_c", 2.0)))
        for k in ("N0", "E", "T", "Z", "sdepv_expt", "sdepv_trns"):
            v = self.visc[k]
            self.visc[k] = v + [v[-1]] * (self.num_mat - len(v))
        self.zbase_layer = [f32(float(g("z_lith", 0.0))), f32(float(g("z_410", 1.0))), f32(float(g("z_lmantle", 1.0))), f32(0.55)]
        self.perturb_mag = _fvec(g("perturbmag", "0.001"))[0]
        self.perturb_k = _fvec(g("perturbk", "1.0"))[0]
        self._coords1d = None

    # ---------------------------------------------------------------- mesh
    def dims(self, lev):
        return self.nox[lev], self.noy[lev], self.noz[lev]

    def nno(self, lev):
        return self.nox[lev] * self.noy[lev] * self.noz[lev]

    def nel(self, lev):
        return (self.nox[lev] - 1) * (self.noy[lev] - 1) * (self.noz[lev] - 1)

    def _axis(self, d):
        """Global 1-D node positions in direction d (0 x, 1 y, 2 z), float accumulation as Nodal_mesh.c:81-166."""
        p = self.params
        lm = self.levmax
        nnx = (self.NOX[lm], self.NOY[lm], self.NOZ[lm])[d]
        name = "xyz"[d]
        X = np.zeros(nnx + 2, dtype=f32)
        dx = f32(self.layer[d] / f32(nnx - 1))
        X[1] = 0.0
        X[nnx] = self.layer[d]
        for i in range(2, nnx):
            X[i] = f32(X[i - 1] + dx)
        nl = int(p.get(f"{name}_grid_layers", 1))
        zz = _fvec(p.get(name * 2, "0,1"), nl)
        nz = _ivec(p.get("n" + name, f"1,{nnx}"))[:nl]
        dxx = [f32(0)] + [f32(f32(zz[j] - zz[j - 1]) / f32(nz[j] - nz[j - 1])) for j in range(1, nl)]
        j = 1
        for i in range(2, nnx):
            if j < nl and i <= nz[j]:
                X[i] = f32(X[i - 1] + dxx[j])
            if j < nl and i == nz[j]:
                j += 1
        return X[1:nnx + 1].copy()

    def coords1d(self):
        if self._coords1d is None:
            self._coords1d = [self._axis(d) for d in range(3)]
        return self._coords1d

    def coordinates(self, lev):
        """(X1, X2, X3) float32[nno] in the reference's node order n = k + noz*(j + nox*i)."""
        ax = self.coords1d()
        s = 2 ** (self.levmax - lev)
        lm = self.levmax
        xs = ax[0][self.NXS[lm] - 1: self.NXS[lm] - 1 + self.nox[lm]][::s]
        ys = ax[1][self.NYS[lm] - 1: self.NYS[lm] - 1 + self.noy[lm]][::s]
        zs = ax[2][self.NZS[lm] - 1: self.NZS[lm] - 1 + self.noz[lm]][::s]
        nox, noy, noz = self.dims(lev)
        assert len(xs) == nox and len(ys) == noy and len(zs) == noz
        X1 = np.broadcast_to(xs[None, :, None], (noy, nox, noz)).reshape(-1).astype(f32)
        X2 = np.broadcast_to(ys[:, None, None], (noy, nox, noz)).reshape(-1).astype(f32)
        X3 = np.broadcast_to(zs[None, None, :], (noy, nox, noz)).reshape(-1).astype(f32)
        return X1, X2, X3

    def velocity_bcs(self):
        """E->VB at levmax (velocity_boundary_conditions, Boundary_conditions.c:44-100) where a velocity flag reads it: the no-slip rows
        are written last, so the lid / base value also holds on the wall edges; every other flagged dof is held at zero."""
        nox, noy, noz = self.dims(self.levmax)
        VB = [np.zeros((noy, nox, noz), dtype=np.float32) for _ in range(3)]
        if self.topvbc == 1 and self.me_loc[2] == self.nproc[2] - 1:
            VB[0][:, :, noz - 1], VB[1][:, :, noz - 1] = self.plate_vel, self.topvby
        if self.botvbc == 1 and self.me_loc[2] == 0:
            VB[0][:, :, 0], VB[1][:, :, 0] = self.botvbx, self.botvby
        return [v.reshape(-1) for v in VB]

    # ---------------------------------------------------------------- boundary-condition flags
    def node_flags(self, lev):
        """NODE[lev] (BC bits only; at levmax also the temperature bits and INTX|INTZ|INTY)."""
        nox, noy, noz = self.dims(lev)
        F = np.zeros((noy, nox, noz), dtype=np.uint32)
        px, py, pz = self.me_loc
        first_x, last_x = px == 0, px == self.nproc[0] - 1
        first_y, last_y = py == 0, py == self.nproc[1] - 1
        bot, top = pz == 0, pz == self.nproc[2] - 1
        gz = np.arange(noz) + self.NZS[lev]                   # global 1-based z index
        gx = np.arange(nox) + self.NXS[lev]
        zin = (gz != 1) & (gz != self.NOZ[lev])
        xin = (gx != 1) & (gx != self.NOX[lev])

        def setb(sl, on, off):
            F[sl] = (F[sl] | np.uint32(on)) & ~np.uint32(off)

        rows = []
        if bot:
            rows.append((self.botvbc, (slice(None), slice(None), 0)))
        if top:
            rows.append((self.topvbc, (slice(None), slice(None), noz - 1)))
        for vbc, sl in rows:                                  # free slip first (Boundary_conditions.c:50-71)
            if vbc != 1:
                setb(sl, VBZ | SBX | SBY, VBX | VBY | SBZ)
        # reflecting side walls (velocity_refl_vert_bc, :259-365): x walls then y walls
        for cond, j in ((first_x, 0), (last_x, nox - 1)):
            if cond:
                F[:, j, :] = (F[:, j, :] | np.uint32(VBX)) & ~np.uint32(SBX)
                sub = F[:, j, :]
                sub[:, zin] = (sub[:, zin] | np.uint32(SBY | SBZ)) & ~np.uint32(VBY | VBZ)
                F[:, j, :] = sub
        for cond, i in ((first_y, 0), (last_y, noy - 1)):
            if cond:
                sub = F[i, :, :]
                sub[:] = (sub | np.uint32(VBY)) & ~np.uint32(SBY)
                sub[:, zin] = (sub[:, zin] | np.uint32(SBZ)) & ~np.uint32(VBZ)
                m = np.outer(xin, zin)
                sub[m] = (sub[m] | np.uint32(SBX)) & ~np.uint32(VBX)
                F[i, :, :] = sub
        for vbc, sl in rows:                                  # no slip last (:79-100)
            if vbc == 1:
                setb(sl, VBX | VBZ | VBY, SBX | SBZ | SBY)
        if lev == self.levmax:
            if bot:
                if self.bottbc in (1, 2):
                    setb((slice(None), slice(None), 0), TBZ, FBZ)
                elif self.bottbc == 0:
                    setb((slice(None), slice(None), 0), FBZ, TBZ)
            if top:
                if self.toptbc >= 1:
                    setb((slice(None), slice(None), noz - 1), TBZ, FBZ)
                elif self.toptbc == 0:
                    setb((slice(None), slice(None), noz - 1), FBZ, TBZ)
            for cond, j in ((first_x, 0), (last_x, nox - 1)):
                if cond:
                    F[:, j, :] = (F[:, j, :] | np.uint32(FBX)) & ~np.uint32(TBX)
            for cond, i in ((first_y, 0), (last_y, noy - 1)):
                if cond:
                    F[i, :, :] = (F[i, :, :] | np.uint32(FBY)) & ~np.uint32(TBY)
            F |= np.uint32(INTX | INTZ | INTY)
        return F.reshape(-1)

    # ---------------------------------------------------------------- fields
    def initial_temperature(self):
        """convection_initial_temperature (Convection.c:325-350), restart=0, then temperatures_conform_bcs."""
        lm = self.levmax
        X1, X2, X3 = self.coordinates(lm)
        x1, y1, z1 = X1.astype(np.float64), X2.astype(np.float64), X3.astype(np.float64)
        T = (1 - z1).astype(f32)
        k = float(self.perturb_k)
        pert = float(self.perturb_mag) * np.sin(np.pi * (1.0 - z1)) * np.cos(k * np.pi * x1) * np.cos(k * np.pi * y1)
        T = (T.astype(np.float64) + pert).astype(f32)
        nox, noy, noz = self.dims(lm)
        T3 = T.reshape(noy, nox, noz)
        if self.me_loc[2] == 0 and self.bottbc in (1, 2):
            T3[:, :, 0] = self.bottbcval
        if self.me_loc[2] == self.nproc[2] - 1 and self.toptbc >= 1:
            T3[:, :, noz - 1] = self.toptbcval
        return T

    def _depth_coordinate(self, lev):
        """Xtmp[3] of construct_mat_group / visc_from_T: z of the box (r of the regional sphere in SphericalProblem)."""
        return self.coordinates(lev)[2]

    def material(self):
        """construct_mat_group: layer index from the float mean of the 8 node depths."""
        lm = self.levmax
        nox, noy, noz = self.dims(lm)
        zs = self._depth_coordinate(lm).reshape(noy, nox, noz)[0, 0, :]
        zlo, zhi = zs[:-1], zs[1:]
        # x3 = sum over local nodes 1..8 (four at the lower z, then four at the upper z), in float
        x3 = np.zeros(noz - 1, dtype=f32)
        for z in (zlo, zlo, zlo, zlo, zhi, zhi, zhi, zhi):
            x3 = (x3 + z).astype(f32)
        x3 = (x3 / f32(8)).astype(f32)
        lay = np.full(noz - 1, self.num_mat, dtype=np.int32)
        done = np.zeros(noz - 1, dtype=bool)
        for i in range(self.num_mat - 1):
            hit = (~done) & (x3 >= self.zbase_layer[i])
            lay[hit] = i + 1
            done |= hit
        return np.broadcast_to(lay[None, None, :], (noy - 1, nox - 1, noz - 1)).reshape(-1).astype(np.int32).copy()

    def buoyancy(self, T):
        """thermal_buoyancy, thermal only, unit expansivity: Ra*T minus its horizontal (layer) average.
        The layer average is the reference's 2-D Gauss quadrature (return_horiz_ave), evaluated here in
        float64 with trapezoid-equivalent bilinear weights on this rank's layer."""
        lm = self.levmax
        nox, noy, noz = self.dims(lm)
        b = (self.rayleigh * T.astype(f32) * f32(1.0)).astype(f32).reshape(noy, nox, noz)
        ax = self.coords1d()
        xs = ax[0][self.NXS[lm] - 1: self.NXS[lm] - 1 + nox].astype(np.float64)
        ys = ax[1][self.NYS[lm] - 1: self.NYS[lm] - 1 + noy].astype(np.float64)

        def w1(x):
            w = np.zeros_like(x)
            d = np.diff(x)
            w[:-1] += 0.5 * d
            w[1:] += 0.5 * d
            return w

        W = np.outer(w1(ys), w1(xs))
        H = np.tensordot(W, b.astype(np.float64), axes=([0, 1], [0, 1])) / W.sum()
        if self.nproc[0] * self.nproc[1] != 1:
            raise NotImplementedError("buoyancy(): the layer average spans ranks; use global_problem().buoyancy + local_slice")
        return (b - H.astype(f32)[None, None, :]).astype(f32).reshape(-1)

    # ---------------------------------------------------------------- global <-> subdomain views
    def global_problem(self) -> "CartesianProblem":
        """The same input file seen by a single rank owning the whole mesh."""
        if self.nproc == (1, 1, 1):
            return self
        return type(self)(self.text + "nprocx=1\nnprocy=1\nnprocz=1\n")

    def local_slice(self, field, lev=None, per_node=1):
        """This subdomain's part (duplicated faces included) of a nodal field given on the global mesh."""
        lev = self.levmax if lev is None else lev
        nox, noy, noz = self.dims(lev)
        a = np.asarray(field).reshape(self.NOY[lev], self.NOX[lev], self.NOZ[lev], per_node)
        i0, j0, k0 = self.NYS[lev] - 1, self.NXS[lev] - 1, self.NZS[lev] - 1
        return np.ascontiguousarray(a[i0:i0 + noy, j0:j0 + nox, k0:k0 + noz]).reshape(-1)

    def local_slice_elements(self, field, lev=None, per_elt=1):
        lev = self.levmax if lev is None else lev
        nox, noy, noz = self.dims(lev)
        a = np.asarray(field).reshape(self.NOY[lev] - 1, self.NOX[lev] - 1, self.NOZ[lev] - 1, per_elt)
        i0, j0, k0 = self.NYS[lev] - 1, self.NXS[lev] - 1, self.NZS[lev] - 1
        return np.ascontiguousarray(a[i0:i0 + noy - 1, j0:j0 + nox - 1, k0:k0 + noz - 1]).reshape(-1)


class SphericalProblem(CartesianProblem):
    """One rank's view of a Geometry=Rsphere input file (regional spherical block, BASELINE config 4): mesh and boundary flags.
    Directions 1, 2, 3 are colatitude, longitude (degrees in the file) and radius; `spherical_coordinates` returns E->SXX
    (theta, phi in radians, r), `coordinates` the Cartesian node positions E->XX the element routines integrate on
    (Nodal_mesh.c:87-96, 188-203); initial temperature and material groups as the reference sets them up."""
    GEOMETRY = "Rsphere"

    def __init__(self, text: str, me_loc=(0, 0, 0)):
        super().__init__(text, me_loc)
        g = self.params.get
        self.corner = ((f32(float(g("theta_north"))), f32(float(g("theta_south")))), (f32(float(g("fi_west"))), f32(float(g("fi_east")))),
                       (f32(float(g("radius_inner"))), f32(float(g("radius_outer")))))
        # material layers by radius (Viscosity_structures.c:122-141)
        self.zbase_layer = [f32(float(g("r_lith", 0.0))), f32(float(g("r_410", 1.0))), f32(float(g("r_lmantle", 1.0))), f32(0.55)]

    def _axis(self, d):
        """Global 1-D node positions in direction d (0 theta [deg], 1 phi [deg], 2 r), float accumulation as Nodal_mesh.c:87-166."""
        p = self.params
        lm = self.levmax
        nnx = (self.NOX[lm], self.NOY[lm], self.NOZ[lm])[d]
        lo, hi = self.corner[d]
        X = np.zeros(nnx + 2, dtype=f32)
        dx = f32(f32(hi - lo) / f32(nnx - 1))
        X[1] = lo
        X[nnx] = hi
        for i in range(2, nnx):
            X[i] = f32(X[i - 1] + dx)
        name = "tfr"[d]
        nl = int(p.get(f"{name}_grid_layers", 1))
        zz = _fvec(p.get(name * 2, f"{lo},{hi}"), nl)
        nz = _ivec(p.get("n" + name, f"1,{nnx}"))[:nl]
        dxx = [f32(0)] + [f32(f32(zz[j] - zz[j - 1]) / f32(nz[j] - nz[j - 1])) for j in range(1, nl)]
        j = 1
        for i in range(2, nnx):
            if j < nl and i <= nz[j]:
                X[i] = f32(X[i - 1] + dxx[j])
            if j < nl and i == nz[j]:
                j += 1
        return X[1:nnx + 1].copy()

    def spherical_coordinates(self, lev):
        """(theta, phi, r) float32[nno] = E->SXX[lev][1..3]."""
        T, F, R = super().coordinates(lev)                       # the 1-D arrays broadcast over the block; angles still in degrees
        rad = np.pi / 180
        return (T.astype(np.float64) * rad).astype(f32), (F.astype(np.float64) * rad).astype(f32), R

    def coordinates(self, lev):
        """Cartesian node positions E->XX[lev] (float products of double sines / cosines of the float angles, :197-199)."""
        t, f, r = [a.astype(np.float64) for a in self.spherical_coordinates(lev)]
        return (r * np.sin(t) * np.cos(f)).astype(f32), (r * np.sin(t) * np.sin(f)).astype(f32), (r * np.cos(t)).astype(f32)

    def _depth_coordinate(self, lev):
        return self.spherical_coordinates(lev)[2]

    def initial_temperature(self):
        """convection_initial_temperature, Rsphere branch (Convection.c:364-408), restart=0, then temperatures_conform_bcs."""
        lm = self.levmax
        t, f, r = [a.astype(np.float64) for a in self.spherical_coordinates(lm)]
        (ti, to), (fi, fo), (ri, ro) = self.corner
        rad, eps = np.pi / 180, 1.0e-6                               # CITCOM_TRACER_EPS_MARGIN (global_defs.h:93)
        xg1 = (float(ti) * rad + eps, float(fi) * rad + eps)
        xg2 = (float(to) * rad - eps, float(fo) * rad - eps)
        beta = float(f32(ri / f32(ri - ro)))
        T = (beta * (1.0 - 1.0 / r)).astype(f32)
        k = float(self.perturb_k)
        pert = float(self.perturb_mag) * np.sin(np.pi * (float(ro) - r) / float(f32(ro - ri))) * \
            np.cos(k * np.pi * (t - xg1[0]) / (xg2[0] - xg1[0])) * np.cos(k * np.pi * (f - xg1[1]) / (xg2[1] - xg1[1]))
        T = (T.astype(np.float64) + pert).astype(f32)
        nox, noy, noz = self.dims(lm)
        T3 = T.reshape(noy, nox, noz)
        if self.me_loc[2] == 0 and self.bottbc in (1, 2):
            T3[:, :, 0] = self.bottbcval
        if self.me_loc[2] == self.nproc[2] - 1 and self.toptbc >= 1:
            T3[:, :, noz - 1] = self.toptbcval
        return T

    def buoyancy(self, T):
        raise NotImplementedError("SphericalProblem: the shell averages of thermal_buoyancy live on the device (StokesContext.thermal_buoyancy)")
