"""Oracle: RectilinearGrid (regular spacing), Field storage, halo filling.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates:
  * src/Grids/rectilinear_grid.jl:264-291, grid_generation.jl:98-155 (regular spacing,
    Flat => N=1, H=0, Δ=1), input_validation.jl:71-77 (default halo min(3, N)),
    grid_utils.jl:66-72 + new_data.jl:15-73 (parent size: N + 2H, +1 for Face in Bounded)
  * src/BoundaryConditions/fill_halo_regions*.jl, boundary_condition_ordering.jl:17-46,113-128
"""
from fractions import Fraction

import numpy as np

PERIODIC, BOUNDED, FLAT = "P", "B", "F"


def _rational(x):
    return Fraction(x) if not isinstance(x, Fraction) else x


class Grid:
    """Regular RectilinearGrid.  ``topology`` is a 3-tuple of 'P'/'B'/'F'.

    ``size`` and ``extent``/``x,y,z`` list only the non-Flat dimensions, like the
    reference constructor (rectilinear_grid.jl:264-291).
    """

    def __init__(self, FT=np.float64, size=None, extent=None, x=None, y=None, z=None,
                 topology=(PERIODIC, PERIODIC, BOUNDED), halo=None):
        self.FT = np.dtype(FT).type
        self.topo = tuple(topology)
        nonflat = [d for d in range(3) if self.topo[d] != FLAT]
        size = (size,) if np.isscalar(size) else tuple(size)
        assert len(size) == len(nonflat), "size must list the non-Flat dimensions"
        bounds = [None, None, None]
        if extent is not None:
            extent = (extent,) if np.isscalar(extent) else tuple(extent)
            for n, d in enumerate(nonflat):
                # extent => x=(0,Lx), y=(0,Ly), z=(-Lz,0)   (input_validation.jl)
                bounds[d] = (0.0, float(extent[n])) if d < 2 else (-float(extent[n]), 0.0)
        else:
            for d, b in enumerate((x, y, z)):
                if b is not None and isinstance(b, tuple) and len(b) == 2:
                    bounds[d] = (float(b[0]), float(b[1]))
        if halo is not None:
            halo = (halo,) if np.isscalar(halo) else tuple(halo)
        N, H, L, D, X0 = [1, 1, 1], [0, 0, 0], [1.0, 1.0, 1.0], [None] * 3, [0.0] * 3
        # z = vector of Nz+1 faces or a function of the face index: vertically stretched grid
        # (rectilinear_grid.jl:264-291 -> generate_coordinate(FT, topo, N, H, node_generator, …), grid_generation.jl:33-94)
        self.z_faces = None
        if z is not None and (callable(z) or not isinstance(z, tuple) or len(z) != 2):   # a 2-tuple is an interval (regular)
            nz = int(size[nonflat.index(2)])
            zf = [z(i) for i in range(1, nz + 2)] if callable(z) else list(z)
            assert len(zf) == nz + 1, "z must list Nz+1 faces"
            self.z_faces = np.asarray(zf, dtype=self.FT)
            assert np.all(np.diff(self.z_faces) > 0), "The elements of z must be increasing!"
            bounds[2] = (float(self.z_faces[0]), float(self.z_faces[-1]))
        for n, d in enumerate(nonflat):
            N[d] = int(size[n])
            H[d] = int(halo[n]) if halo is not None else min(3, N[d])
            a, b = bounds[d]
            # Δ = FT((BigFloat(b) - BigFloat(a)) / N)   grid_generation.jl:105-110
            delta = (_rational(b) - _rational(a)) / N[d]
            D[d] = self.FT(float(delta))
            L[d] = self.FT(float(_rational(b) - _rational(a)))
            X0[d] = a
        for d in range(3):
            if self.topo[d] == FLAT:
                D[d] = self.FT(1)      # grid_generation.jl:138-155
                L[d] = self.FT(1)
        self.N, self.H, self.L, self.D, self.x0 = tuple(N), tuple(H), tuple(L), tuple(D), tuple(X0)
        self.Nx, self.Ny, self.Nz = self.N
        self.Hx, self.Hy, self.Hz = self.H
        self.dx, self.dy, self.dz = self.D
        FTc = self.FT
        # spacings_and_areas_and_volumes.jl:309-333,376 ; reciprocal_metric_operators.jl:7,13
        self.Ax = FTc(self.dy * self.dz)
        self.Ay = FTc(self.dx * self.dz)
        self.Az = FTc(self.dx * self.dy)
        self.V = FTc(self.Az * self.dz)
        self.rV = FTc(FTc(1) / self.V)
        self.rD = tuple(FTc(FTc(1) / dd) for dd in self.D)
        self.A = (self.Ax, self.Ay, self.Az)
        self.stretched = self.z_faces is not None
        if self.stretched:
            assert self.topo[2] == BOUNDED, "only a Bounded z can be stretched (fourier_tridiagonal_poisson_solver.jl:86-90)"
            self.L = (self.L[0], self.L[1], FTc(self.z_faces[-1] - self.z_faces[0]))
            self.D = (self.D[0], self.D[1], None)
            self.dz = None
            # level-dependent metrics: use Ctx.dz / Ctx.area / Ctx.rvol / Ctx.rdelta (oracle/operators.py)
            self.Ax = self.Ay = self.V = self.rV = None
            self.A = (None, None, self.Az)
            self.rD = (self.rD[0], self.rD[1], None)
            self._generate_z()

    def _generate_z(self):
        """generate_coordinate for a Bounded, variably spaced coordinate (grid_generation.jl:33-94), in FT arithmetic.
        Sets zF[k], zC[k], dzc[k] = Δzᵃᵃᶜ, dzf[k] = Δzᵃᵃᶠ as dicts-by-offset arrays: logical index k <-> array[k + Hz + 1]
        for dzf (k = -Hz … Nz+Hz) and array[k + Hz - 1]-style accessors below."""
        FT, N, H = self.FT, self.N[2], self.H[2]
        Fi = self.z_faces
        dlo = [FT(Fi[1] - Fi[0])] * H                       # lower_exterior_Δcoordᶠ(::BoundedTopology)
        dhi = [FT(Fi[-1] - Fi[-2])] * H                     # reverse(upper_exterior_Δcoordᶠ)
        def ssum(v):
            s = FT(0)
            for x in v:
                s = FT(s + x)
            return s
        Flo = [FT(Fi[0] - ssum(dlo[i:H])) for i in range(H)]
        Fhi = [FT(Fi[-1] + ssum(dhi[i:H])) for i in range(H)][::-1]
        F = np.asarray(Flo + list(Fi) + Fhi, dtype=FT)      # N + 1 + 2H faces, logical index k <-> F[k + H - 1]
        TC = N + 2 * H
        C = np.asarray([FT(FT(F[i + 1] + F[i]) / FT(2)) for i in range(TC)], dtype=FT)
        dF_raw = [FT(C[i] - C[i - 1]) for i in range(1, TC)]
        dC_ = np.asarray([FT(F[i + 1] - F[i]) for i in range(len(F) - 1)], dtype=FT)   # Δᶜ, logical k <-> [k + H - 1]
        dFp = [dF_raw[0]] + dF_raw + [dF_raw[-1]]
        for i in range(len(dFp) - 1, 0, -1):
            dFp[i] = dFp[i - 1]
        dF_ = np.asarray(dFp, dtype=FT)                     # Δᶠ, OffsetArray(-H-1): logical k <-> [k + H]
        self._zF, self._zC, self._dzc, self._dzf = F, C, dC_, dF_

    def dz_at(self, zloc, k):
        """Δzᵃᵃᶜ[k] / Δzᵃᵃᶠ[k] for 1-based logical (possibly array) index k"""
        if not self.stretched:
            return self.D[2]
        k = np.asarray(k)
        return self._dzc[k + self.H[2] - 1] if zloc == "c" else self._dzf[k + self.H[2]]

    def flat(self, d):
        return self.topo[d] == FLAT

    def bounded(self, d):
        return self.topo[d] == BOUNDED

    def with_halo(self, halo):
        g = Grid.__new__(Grid)
        g.__dict__.update(self.__dict__)
        H = tuple(0 if self.flat(d) else int(halo[d]) for d in range(3))
        g.H = H
        g.Hx, g.Hy, g.Hz = H
        if g.stretched:
            g._generate_z()
        return g

    def nodes(self, d, loc):
        """Cell centres ('c') or faces ('f') along dimension d (interior)."""
        n = self.N[d] + (1 if (loc == "f" and self.bounded(d)) else 0)
        if d == 2 and self.stretched:
            H = self.H[2]
            return (self._zC[H:H + n] if loc == "c" else self._zF[H:H + n]).astype(np.float64)
        i = np.arange(n, dtype=np.float64)
        off = 0.5 if loc == "c" else 0.0
        return self.x0[d] + (i + off) * float(self.D[d])


# ---------------------------------------------------------------------------------
# Boundary conditions  (src/BoundaryConditions/boundary_condition.jl:8,83-110)
# ---------------------------------------------------------------------------------
class BC:
    """kind in {'periodic','flux','value','gradient','open', None}; value scalar or 2-D array or None"""

    def __init__(self, kind, value=None):
        self.kind = kind
        self.value = value

    def get(self, FT):
        if self.value is None:
            return FT(0)
        if np.isscalar(self.value):
            return FT(self.value)
        return np.asarray(self.value, dtype=FT)


SIDES = ("west", "east", "south", "north", "bottom", "top")


def default_bcs(grid, loc):
    """field_boundary_conditions.jl:15-60: Periodic -> periodic; Bounded+Center -> NoFlux (Flux, nothing);
    Bounded+Face (wall-normal) -> Impenetrable (Open, nothing -> 0); Flat -> None."""
    bcs = {}
    for d in range(3):
        for s in (0, 1):
            name = SIDES[2 * d + s]
            t = grid.topo[d]
            if t == PERIODIC:
                bcs[name] = BC("periodic")
            elif t == FLAT:
                bcs[name] = BC(None)
            elif loc[d] == "f":
                bcs[name] = BC("open", None)
            else:
                bcs[name] = BC("flux", None)
    return bcs


class Field:
    """Field{LX,LY,LZ}: loc is a 3-string like 'fcc'.  ``data`` is the dense parent array
    (x fastest in the reference; here a C-ordered numpy array indexed [i, j, k] — only the
    logical indexing matters for the oracle).  Logical index i (1-based) <-> data[i + Hx - 1]."""

    def __init__(self, grid, loc="ccc", bcs=None, name=""):
        self.grid, self.loc, self.name = grid, loc, name
        shape = []
        for d in range(3):
            extra = 1 if (loc[d] == "f" and grid.bounded(d)) else 0
            shape.append(grid.N[d] + 2 * grid.H[d] + extra)
        self.data = np.zeros(shape, dtype=grid.FT)
        self.bcs = default_bcs(grid, loc)
        if bcs:
            self.bcs.update(bcs)

    # number of interior points along d (Face+Bounded has N+1)
    def n(self, d):
        return self.grid.N[d] + (1 if (self.loc[d] == "f" and self.grid.bounded(d)) else 0)

    @property
    def interior(self):
        g = self.grid
        sl = tuple(slice(g.H[d], g.H[d] + self.n(d)) for d in range(3))
        return self.data[sl]

    def set(self, value):
        """set!(field, array|number|function(x,y,z))  src/Fields/set!.jl:34-121"""
        g = self.grid
        if callable(value):
            nodes = [g.nodes(d, self.loc[d]) if not g.flat(d) else np.zeros(1) for d in range(3)]
            X, Y, Z = np.meshgrid(*nodes, indexing="ij")
            args = [A for d, A in enumerate((X, Y, Z)) if not g.flat(d)]
            value = np.broadcast_to(np.asarray(value(*args), dtype=np.float64), X.shape)
        if np.isscalar(value):
            self.interior[...] = g.FT(value)
        else:
            a = np.asarray(value)
            self.interior[...] = a.reshape(self.interior.shape).astype(g.FT)

    def like(self, name=""):
        return Field(self.grid, self.loc, bcs=dict(self.bcs), name=name)


def _plane(f, d, idx):
    """view of the parent plane with logical index ``idx`` (1-based) along d"""
    sl = [slice(None)] * 3
    sl[d] = idx + f.grid.H[d] - 1
    return f.data[tuple(sl)]


def _interior_tangential(f, d, arr):
    """Restrict a plane (2-D view, dim d removed) to the tangential interior range 1:N of the *grid*
    (fill_halo_regions.jl:119-128 — size symbols :xy/:xz/:yz mean the grid size, not the field's)."""
    g = f.grid
    others = [e for e in range(3) if e != d]
    sl = tuple(slice(g.H[e], g.H[e] + g.N[e]) for e in others)
    return arr[sl]


def fill_halo_regions(f, fill_open_bcs=True):
    """fill_halo_regions!(field)   fill_halo_regions.jl:25-36.

    Order (boundary_condition_ordering.jl:17-46,113-128): dimensions whose (west/south/bottom) BC is
    non-periodic first (x, y, z order, stable sort), then the periodic ones.  Non-periodic fills touch
    ONE halo plane over the interior tangential range; periodic fills copy H planes over the whole
    parent extent of the other two dimensions (fill_halo_regions_periodic.jl:5-32)."""
    g = f.grid
    FT = g.FT
    order = sorted(range(3), key=lambda d: 1 if f.bcs[SIDES[2 * d]].kind == "periodic" else 0)
    for d in order:
        left, right = f.bcs[SIDES[2 * d]], f.bcs[SIDES[2 * d + 1]]
        N, H = g.N[d], g.H[d]
        if left.kind is None and right.kind is None:
            continue
        if left.kind == "periodic":
            sl_dst_w, sl_src_w, sl_dst_e, sl_src_e = ([slice(None)] * 3 for _ in range(4))
            sl_dst_w[d] = slice(0, H)
            sl_src_w[d] = slice(N, N + H)
            sl_dst_e[d] = slice(N + H, N + 2 * H)
            sl_src_e[d] = slice(H, 2 * H)
            f.data[tuple(sl_dst_w)] = f.data[tuple(sl_src_w)]
            f.data[tuple(sl_dst_e)] = f.data[tuple(sl_src_e)]
            continue
        for side, bc in ((0, left), (1, right)):
            if bc.kind is None:
                continue
            # Δ between the interior and the halo point, at the boundary face (fill_halo_regions_value_gradient.jl:44,60)
            delta = g.dz_at("f", 1 if side == 0 else N + 1) if (d == 2 and g.stretched) else g.D[d]
            if bc.kind == "open":
                # fill_halo_regions_open.jl:2-14 — sets the boundary face itself
                if not fill_open_bcs:
                    continue
                idx = 1 if side == 0 else N + 1
                tgt = _interior_tangential(f, d, _plane(f, d, idx))
                tgt[...] = bc.get(FT)
                continue
            iI = 1 if side == 0 else N          # interior cell
            iH = 0 if side == 0 else N + 1      # halo cell
            cI = _interior_tangential(f, d, _plane(f, d, iI))
            cH = _interior_tangential(f, d, _plane(f, d, iH))
            if bc.kind == "flux":
                cH[...] = cI                       # fill_halo_regions_flux.jl:9-27
            elif bc.kind == "gradient":
                grad = bc.get(FT)                  # fill_halo_regions_value_gradient.jl:9-10
                cH[...] = cI + grad * (-delta if side == 0 else delta)
            elif bc.kind == "value":
                val = bc.get(FT)                   # :12-13
                if side == 0:
                    grad = (cI - val) / (delta / FT(2))
                    cH[...] = cI + grad * (-delta)
                else:
                    grad = (val - cI) / (delta / FT(2))
                    cH[...] = cI + grad * delta
            else:
                raise ValueError(f"unsupported BC kind {bc.kind}")
