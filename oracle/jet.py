"""Second-order forward-mode AD ("jets") in NumPy.  Oracle-only.

Stands in for the CasADi SX ``jacobian`` calls the reference uses to obtain
``fAd/fBd`` and ``fEd/fFd/fGd`` (``DGSQP/dynamics/dynamics_models.py:128-144``).
A jet carries, for a batch of K evaluation points and nv independent
variables, the value ``v[K]``, gradient ``g[K,nv]`` and (optionally) Hessian
``h[K,nv,nv]`` of a scalar expression.
"""
import numpy as np


class Jet:
    __slots__ = ("v", "g", "h")

    def __init__(self, v, g, h=None):
        self.v, self.g, self.h = v, g, h

    # -- construction -----------------------------------------------------
    @staticmethod
    def variables(x, order=2):
        """x: (K, nv) -> list of nv jets, one per independent variable."""
        x = np.asarray(x, dtype=np.float64)
        K, nv = x.shape
        out = []
        for i in range(nv):
            g = np.zeros((K, nv))
            g[:, i] = 1.0
            h = np.zeros((K, nv, nv)) if order >= 2 else None
            out.append(Jet(x[:, i].copy(), g, h))
        return out

    def _const_like(self, c):
        return Jet(np.full_like(self.v, c), np.zeros_like(self.g),
                   None if self.h is None else np.zeros_like(self.h))

    def _lift(self, o):
        return o if isinstance(o, Jet) else self._const_like(float(o))

    # -- chain rule for a univariate function f(a) -------------------------
    def _chain(self, f, df, d2f):
        g = df[:, None] * self.g
        h = None
        if self.h is not None:
            h = df[:, None, None] * self.h + d2f[:, None, None] * (self.g[:, :, None] * self.g[:, None, :])
        return Jet(f, g, h)

    # -- arithmetic ---------------------------------------------------------
    def __add__(self, o):
        if not isinstance(o, Jet):
            return Jet(self.v + o, self.g, self.h)
        return Jet(self.v + o.v, self.g + o.g, None if self.h is None else self.h + o.h)

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.g, None if self.h is None else -self.h)

    def __sub__(self, o):
        if not isinstance(o, Jet):
            return Jet(self.v - o, self.g, self.h)
        return Jet(self.v - o.v, self.g - o.g, None if self.h is None else self.h - o.h)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if not isinstance(o, Jet):
            o = np.asarray(o, dtype=np.float64)
            og, oh = (o[:, None], o[:, None, None]) if o.ndim == 1 else (o, o)
            return Jet(self.v * o, self.g * og, None if self.h is None else self.h * oh)
        v = self.v * o.v
        g = self.v[:, None] * o.g + o.v[:, None] * self.g
        h = None
        if self.h is not None:
            cross = self.g[:, :, None] * o.g[:, None, :]
            h = self.v[:, None, None] * o.h + o.v[:, None, None] * self.h + cross + np.swapaxes(cross, 1, 2)
        return Jet(v, g, h)

    __rmul__ = __mul__

    def recip(self):
        r = 1.0 / self.v
        return self._chain(r, -r * r, 2.0 * r * r * r)

    def __truediv__(self, o):
        if not isinstance(o, Jet):
            return self * (1.0 / o)
        return self * o.recip()

    def __rtruediv__(self, o):
        return self.recip() * o

    def square(self):
        return self._chain(self.v * self.v, 2.0 * self.v, np.full_like(self.v, 2.0))

    # -- elementary functions ----------------------------------------------
    def sin(self):
        s, c = np.sin(self.v), np.cos(self.v)
        return self._chain(s, c, -s)

    def cos(self):
        s, c = np.sin(self.v), np.cos(self.v)
        return self._chain(c, -s, -c)

    def tan(self):
        t = np.tan(self.v)
        sec2 = 1.0 + t * t
        return self._chain(t, sec2, 2.0 * t * sec2)

    def atan(self):
        d = 1.0 / (1.0 + self.v * self.v)
        return self._chain(np.arctan(self.v), d, -2.0 * self.v * d * d)

    def abs_ifelse(self):
        """CasADi ``if_else(x > 0, x, -x)``: derivative is that of the taken branch
        (``dynamics_models.py:228-234``)."""
        sgn = np.where(self.v > 0, 1.0, -1.0)
        return Jet(sgn * self.v, sgn[:, None] * self.g,
                   None if self.h is None else sgn[:, None, None] * self.h)
