"""
Optical materials (`rtm`) -- host side of the drop-in API.

Mirrors the public surface of the reference's ``raytrace.materials`` (reference
``src/raytrace/materials.py:6-227``): ``Material(b_coeffs, c_coeffs).n(wavelength)``, ``Vacuum``,
``Constant(n)`` and the glass catalogue.  Wavelengths are in micrometres.

A material does two jobs here:

* it answers ``n(wavelength)`` on the host exactly like the reference does (used by the paraxial helpers and to
  build the per-wavelength refractive-index table that is shipped to the GPU), and
* it describes itself to the device through :meth:`Material.device_record`, so the fused trace kernel can
  evaluate the three-term Sellmeier formula in registers when a batch has too many distinct wavelengths for a table.

Anything that is not a plain Sellmeier/constant medium (``Ebaf11`` or a user subclass overriding ``n``) is only ever
evaluated by calling the user's own ``n`` on the batch's unique wavelengths: there is no
device-side guess at what an overridden ``n`` does.
"""
from __future__ import annotations

import numpy as np

# device material kinds, must match include/rtb.h
KIND_CONSTANT = 0
KIND_SELLMEIER = 1
KIND_TABLE_ONLY = 2


class Material:
    """
    Dispersive medium described by the Sellmeier formula
    ``n^2 - 1 = sum_i b_i w^2 / (w^2 - c_i)`` (reference ``materials.py:39-51``).

    ``vd`` is the Abbe number ``(n_d - 1) / (n_F - n_C)`` evaluated at the class-level Fraunhofer lines.
    """

    wd = 0.5876  # helium d
    wf = 0.4861  # hydrogen F
    wc = 0.6563  # hydrogen C
    vd = None

    def __init__(self, b_coeffs, c_coeffs):
        b = np.array(b_coeffs).squeeze()
        c = np.array(c_coeffs).squeeze()
        self.b1, self.b2, self.b3 = b
        self.c1, self.c2, self.c3 = c
        with np.errstate(invalid="ignore", divide="ignore"):
            nd, nf, nc = (self.n(w) for w in (self.wd, self.wf, self.wc))
            self.vd = (nd - 1) / (nf - nc)

    def n(self, wavelength):
        """Refractive index at ``wavelength`` (um); scalar or array in, same shape out."""
        w2 = wavelength ** 2
        # the three terms are summed left to right; the device kernel and the oracle use the same order
        acc = self.b1 * w2 / (w2 - self.c1) + self.b2 * w2 / (w2 - self.c2) + self.b3 * w2 / (w2 - self.c3)
        return np.sqrt(acc + 1)

    # ------------------------------------------------------------------
    # device description
    # ------------------------------------------------------------------
    def device_record(self):
        """
        ``(kind, b[3], c[3], n_const)`` for ``rtb_material`` (include/rtb.h).

        The in-kernel formula is only claimed for objects whose ``n`` is *this class's* ``n``; a subclass that
        overrides ``n`` is reported as table-only.
        """
        if type(self).n is not Material.n:
            return KIND_TABLE_ONLY, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), float("nan")
        b = (float(self.b1), float(self.b2), float(self.b3))
        c = (float(self.c1), float(self.c2), float(self.c3))
        return KIND_SELLMEIER, b, c, float("nan")

    def __repr__(self):
        return f"{type(self).__name__}()"


class Vacuum(Material):
    """n = 1 exactly (all Sellmeier coefficients zero, reference ``materials.py:54-56``)."""

    def __init__(self):
        super().__init__([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])


class Constant(Material):
    """Wavelength-independent index (reference ``materials.py:59-79``)."""

    def __init__(self, n):
        self._n = float(n)
        self.b1 = self.b2 = self.b3 = None
        self.c1 = self.c2 = self.c3 = None

    def n(self, wavelength):
        if isinstance(wavelength, float):
            return self._n
        shape = np.atleast_1d(np.array(wavelength)).shape
        return np.ones(shape) * self._n

    def device_record(self):
        if type(self).n is not Constant.n:
            return KIND_TABLE_ONLY, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), float("nan")
        return KIND_CONSTANT, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), self._n

    def __repr__(self):
        return f"Constant({self._n!r})"


class Ebaf11(Material):
    """
    HIKARI E-BAF11, given as a Laurent polynomial in w^2 rather than Sellmeier terms
    (reference ``materials.py:128-144``).  Table-only on the device: ``w**-2 ... w**-8`` go through ``np.power``,
    whose last bit is platform dependent, so the index is always taken from this host method.
    """

    def __init__(self):
        self.params = [2.71954649, -0.0100472501, 0.0200301385,
                       0.00046586302, -7.51633336e-6, 1.77544989e-6]

    def n(self, wavelength):
        p = self.params
        n_sqr = (p[0] + p[1] * wavelength**2 + p[2] * wavelength**-2 + p[3] * wavelength**-4 +
                 p[4] * wavelength**-6 + p[5] * wavelength**-8)
        return np.sqrt(n_sqr)

    def device_record(self):
        return KIND_TABLE_ONLY, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), float("nan")


# ----------------------------------------------------------------------
# Sellmeier catalogue.  Coefficients are data carried over from the reference tables
# (reference ``materials.py:82-227``); c_i in um^2.
# ----------------------------------------------------------------------
_CATALOGUE = {
    # name: (b1, b2, b3), (c1, c2, c3), note
    "FusedSilica": ((0.6961663, 0.4079426, 0.8974794),
                    (0.0684043**2, 0.1162414**2, 9.896161**2), "fused silica"),
    "Bk7": ((1.03961212, 0.231792344, 1.01046945),
            (0.00600069867, 0.0200179144, 103.560653), "crown"),
    "Nbak4": ((1.28834642, 0.132817724, 0.945395373),
              (0.00779980626, 0.0315631177, 105.965875), "crown"),
    "Nbaf10": ((1.5851495, 0.143559385, 1.08521269),
               (0.00926681282, 0.0424489805, 105.613573), "crown"),
    "Nlak22": ((1.14229781, 0.535138441, 1.040883850),
               (0.00585778594, 0.0198546147, 100.8340170), "crown"),
    "Nsk11": ((1.17963631, 0.229817295, 0.935789652),
              (0.00680282081, 0.0219737205, 101.513232), "crown"),
    "Sf10": ((1.62153902, 0.256287842, 1.64447552),
             (0.0122241457, 0.0595736775, 147.468793), "flint"),
    "Nsf11": ((1.737596950, 0.313747346, 1.898781010),
              (0.013188707, 0.0623068142, 155.23629000), "flint"),
    "Nsf6": ((1.77931763, 0.338149866, 2.087344740),
             (0.01337141820, 0.0617533621, 174.0175900), "flint"),
    "Sf6": ((1.72448482, 0.390104889, 1.045728580),
            (0.01348719470, 0.0569318095, 118.5571850), "flint"),
    "Nsf6ht": ((1.77931763, 0.338149866, 2.087344740),
               (0.01337141820, 0.0617533621, 174.0175900), "flint"),
    "Sf2": ((1.40301821, 0.231767504, 0.939056586),
            (0.0105795466, 0.0493226978, 112.405955), "flint"),
    "Nsf19": ((1.52005444, 0.17573947, 1.43623424),
              (0.01096144, 0.0593248486, 126.795151), "flint"),
}


def _make_glass(name, bs, cs, note):
    def __init__(self):
        Material.__init__(self, list(bs), list(cs))

    return type(name, (Material,), {"__init__": __init__,
                                    "__doc__": f"{name}: Sellmeier {note} glass.",
                                    "__module__": __name__})


for _name, (_bs, _cs, _note) in _CATALOGUE.items():
    globals()[_name] = _make_glass(_name, _bs, _cs, _note)
del _name, _bs, _cs, _note

__all__ = ["Material", "Vacuum", "Constant", "Ebaf11", *_CATALOGUE.keys()]
