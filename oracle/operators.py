"""Oracle: staggered-grid operators evaluated over an index window.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates src/Operators/
difference_operators.jl:7-14,33-49, derivative_operators.jl:20-22,
interpolation_operators.jl:8-15,45-71,87-112, divergence_operators.jl:16-19.

A *quantity* is a callable ``q(o)`` that returns the array of values at logical indices
``(i+o[0], j+o[1], k+o[2])`` for all (i,j,k) of the evaluation window of a ``Ctx``.
Operators build new quantities from quantities, which mirrors the reference's style of
passing point functions ``f(i, j, k, grid, args...)`` into ``δ``/``ℑ``/``∂``.
"""
import numpy as np


def sh(o, d, n):
    o = list(o)
    o[d] += n
    return tuple(o)


class Ctx:
    """Evaluation window: inclusive 1-based logical ranges in each dimension."""

    def __init__(self, grid, ir, jr, kr):
        self.g = grid
        self.r = (tuple(ir), tuple(jr), tuple(kr))
        self.shape = tuple(r[1] - r[0] + 1 for r in self.r)
        self.FT = grid.FT

    def field(self, f):
        """quantity for field data"""
        g, r = self.g, self.r
        data = f.data

        def q(o):
            sl = []
            for d in range(3):
                if g.flat(d):
                    assert o[d] == 0 or True
                    sl.append(slice(0, 1))  # Flat: single plane, any offset maps to it (ψ is uniform along d)
                else:
                    a = r[d][0] + o[d] + g.H[d] - 1
                    b = r[d][1] + o[d] + g.H[d]
                    assert a >= 0 and b <= data.shape[d], (f.name, d, o, a, b, data.shape)
                    sl.append(slice(a, b))
            return data[tuple(sl)]
        return q

    def index(self, d, o):
        """logical index array (1-based) along d at offset o, broadcastable to the window"""
        shape = [1, 1, 1]
        shape[d] = self.shape[d]
        return (np.arange(self.r[d][0], self.r[d][1] + 1) + o[d]).reshape(shape)

    def zeros(self):
        return np.zeros(self.shape, dtype=self.FT)

    # ---- metrics as quantities (spacings_and_areas_and_volumes.jl:309-376, reciprocal_metric_operators.jl:7-13).
    # x and y are regularly spaced; z may be stretched, in which case Δz depends on the level AND on the z-location
    # ('c': Δzᵃᵃᶜ[k], 'f': Δzᵃᵃᶠ[k]) of the point where the metric is evaluated.
    def dz(self, zloc):
        g = self.g
        if not g.stretched:
            return lambda o: g.D[2]
        return lambda o: g.dz_at(zloc, self.index(2, o))

    def area(self, d, zloc):
        """Ax = Δy·Δz, Ay = Δx·Δz, Az = Δx·Δy at a point whose z-location is zloc"""
        g = self.g
        if d == 2 or not g.stretched:
            A = g.A[d]
            return lambda o: A
        other = g.D[1] if d == 0 else g.D[0]
        dz = self.dz(zloc)
        return lambda o: other * dz(o)

    def rvol(self, zloc):
        """V⁻¹ = 1 / (Az·Δz)"""
        g = self.g
        if not g.stretched:
            return lambda o: g.rV
        dz = self.dz(zloc)
        one = self.FT(1)
        return lambda o: one / (g.Az * dz(o))

    def vol(self, zloc):
        g = self.g
        if not g.stretched:
            return lambda o: g.V
        dz = self.dz(zloc)
        return lambda o: g.Az * dz(o)

    def rdelta(self, d, loc):
        """Δ⁻¹ along d at a point whose location along d is loc"""
        g = self.g
        if d != 2 or not g.stretched:
            r = g.rD[d]
            return lambda o: r
        dz = self.dz(loc)
        one = self.FT(1)
        return lambda o: one / dz(o)

    def const(self, v):
        return lambda o: self.FT(v)


O = (0, 0, 0)


# difference_operators.jl:7-14 ; Flat => 0 (:33-49)
def dC(ctx, q, d):
    """δᶜ: difference landing on a Center from Face data:  q[i+1] - q[i]"""
    if ctx.g.flat(d):
        return lambda o: ctx.zeros()
    return lambda o: q(sh(o, d, 1)) - q(o)


def dF(ctx, q, d):
    """δᶠ: difference landing on a Face from Center data:  q[i] - q[i-1]"""
    if ctx.g.flat(d):
        return lambda o: ctx.zeros()
    return lambda o: q(o) - q(sh(o, d, -1))


# interpolation_operators.jl:8-15 ; Flat => identity (:87-112)
def iC(ctx, q, d):
    """ℑᶜ: Face -> Center:  0.5 (q[i] + q[i+1])"""
    if ctx.g.flat(d):
        return q
    h = ctx.FT(0.5)
    return lambda o: h * (q(o) + q(sh(o, d, 1)))


def iF(ctx, q, d):
    """ℑᶠ: Center -> Face:  0.5 (q[i-1] + q[i])"""
    if ctx.g.flat(d):
        return q
    h = ctx.FT(0.5)
    return lambda o: h * (q(sh(o, d, -1)) + q(o))


# derivative_operators.jl:20-22 : ∂ = δ * Δ⁻¹
def ddC(ctx, q, d):
    r = ctx.rdelta(d, "c")
    dq = dC(ctx, q, d)
    return lambda o: dq(o) * r(o)


def ddF(ctx, q, d):
    r = ctx.rdelta(d, "f")
    dq = dF(ctx, q, d)
    return lambda o: dq(o) * r(o)


def scaled(q, a):
    """a * q   (e.g. Ax_qᶠᶜᶜ = Ax * u); a is a metric quantity"""
    return lambda o: a(o) * q(o)


def div_ccc(ctx, u, v, w):
    """divᶜᶜᶜ  divergence_operators.jl:16-19"""
    qx = dC(ctx, scaled(ctx.field(u), ctx.area(0, "c")), 0)
    qy = dC(ctx, scaled(ctx.field(v), ctx.area(1, "c")), 1)
    qz = dC(ctx, scaled(ctx.field(w), ctx.area(2, "f")), 2)
    return ctx.rvol("c")(O) * (qx(O) + qy(O) + qz(O))
