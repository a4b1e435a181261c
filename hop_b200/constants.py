"""Label conventions and unit factors of the hop hot path.

``SITELABEL`` is the hard contract of hop/constants.py:70.  The density unit factors are those of
``MDAnalysis.units`` (re-exported by hop/constants.py:45,63); MDAnalysis is not a dependency here,
so the 0.18-release values are pinned (hop's setup.py:42 requires MDAnalysis>=0.18.0).
"""
from __future__ import annotations

#: pre-defined integer labels for sites -- hop/constants.py:70
SITELABEL = dict(outlier=-1, interstitial=0, bulk=1)

#: Angstrom; conventional water van der Waals radius (hop/constants.py:48)
r_water = 1.4

constants = {"N_Avogadro": 6.02214179e+23, "elementary_charge": 1.602176487e-19, "calorie": 4.184}
N_Avogadro = constants["N_Avogadro"]

water = {"exp": 0.997, "SPC": 0.985, "TIP3P": 1.002, "TIP4P": 1.001, "MolarMass": 18.016}

lengthUnit_factor = {"Angstrom": 1.0, "A": 1.0, "angstrom": 1.0, "nm": 1.0 / 10, "nanometer": 1.0 / 10,
                     "pm": 1e2, "picometer": 1e2, "fm": 1e5, "femtometer": 1e5}

densityUnit_factor = {
    "Angstrom^{-3}": 1 / 1.0, "A^{-3}": 1 / 1.0,
    "nm^{-3}": 1 / 1e-3, "nanometer^{-3}": 1 / 1e-3,
    "Molar": 1 / (1e-27 * N_Avogadro),
    "SPC": 1 / (1e-24 * N_Avogadro * water["SPC"] / water["MolarMass"]),
    "TIP3P": 1 / (1e-24 * N_Avogadro * water["TIP3P"] / water["MolarMass"]),
    "TIP4P": 1 / (1e-24 * N_Avogadro * water["TIP4P"] / water["MolarMass"]),
    "water": 1 / (1e-24 * N_Avogadro * water["exp"] / water["MolarMass"]),
}

timeUnit_factor = {"ps": 1.0, "picosecond": 1.0, "fs": 1e3, "femtosecond": 1e3, "ns": 1e-3, "nanosecond": 1e-3,
                   "second": 1e-12, "sec": 1e-12, "s": 1e-12, "AKMA": 1 / 4.888821e-2}

#: conversion_factor used by get_conversion_factor(type, unit1, unit2); 'threshold' is a density
#: (hop/constants.py:63)
conversion_factor = {"length": lengthUnit_factor, "density": densityUnit_factor, "time": timeUnit_factor,
                     "threshold": densityUnit_factor}


def get_conversion_factor(unit_type, u1, u2):
    """Factor f such that X[u2] = f * X[u1] (MDAnalysis.units.get_conversion_factor)."""
    return conversion_factor[unit_type][u2] / conversion_factor[unit_type][u1]


def density_conversion_factor(u1, u2):
    return get_conversion_factor("density", u1, u2)
