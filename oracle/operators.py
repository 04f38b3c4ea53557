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
    r = ctx.g.rD[d]
    dq = dC(ctx, q, d)
    return lambda o: dq(o) * r


def ddF(ctx, q, d):
    r = ctx.g.rD[d]
    dq = dF(ctx, q, d)
    return lambda o: dq(o) * r


def scaled(q, a):
    """a * q   (e.g. Ax_qᶠᶜᶜ = Ax * u)"""
    return lambda o: a * q(o)


def div_ccc(ctx, u, v, w):
    """divᶜᶜᶜ  divergence_operators.jl:16-19"""
    g = ctx.g
    qx = dC(ctx, scaled(ctx.field(u), g.Ax), 0)
    qy = dC(ctx, scaled(ctx.field(v), g.Ay), 1)
    qz = dC(ctx, scaled(ctx.field(w), g.Az), 2)
    return g.rV * (qx(O) + qy(O) + qz(O))
