"""Lennard-Jones reduced units (argon), host-side only.

Mirrors the unit system of the reference (`enflow/utils/constants.py:2-5`,
`enflow/utils/conversion.py:5-62`): length sigma = 3.4 A, energy eps = 238 J/mol,
mass = argon atomic weight.  The reference pulls the argon mass from rdkit's
periodic table (`constants.py:2`); the value it returns is 39.948 amu and is
pinned here so rdkit is not required.
"""
import math

AR_MASS_AMU = 39.948          # rdkit GetAtomicWeight('Ar'), constants.py:2
SIGMA_M = 3.4e-10             # constants.py:3
EPS_J_PER_MOL = 0.238e3       # constants.py:4
KB_J_PER_K_MOL = 8.3144621    # constants.py:5

ATOM_TYPES = {'H': 0, 'C': 1, 'N': 2, 'O': 3, 'F': 4}   # constants.py:7

_LEN = {'ang': 1e-10, 'nm': 1e-9}
_TIME = {'pico': 1e-12, 'femto': 1e-15}


def meter_to_lj(x):
    return x / SIGMA_M


def second_to_lj(t):
    return t * math.sqrt(EPS_J_PER_MOL / AR_MASS_AMU) / SIGMA_M


def time_to_lj(t, unit='pico'):
    return second_to_lj(t * _TIME[unit])


def time_to_lj_physical(t, unit='pico'):
    """Time in the physical reduced unit tau = sigma sqrt(M / eps) with M in kg/mol (4.405 ps for argon).
    `second_to_lj` above keeps the reference's convention (conversion.py:14-15 divides J/mol by amu, i.e. g/mol) and is
    sqrt(1000) = 31.6 times smaller; the reference applies it to the flow's dt, while its prior sampler (OpenMM,
    simulated.py:109) advances physical time.  The GPU sampler follows OpenMM: it integrates with THIS step."""
    return t * _TIME[unit] / (SIGMA_M * math.sqrt(AR_MASS_AMU * 1e-3 / EPS_J_PER_MOL))


def dist_to_lj(x, unit='ang'):
    return meter_to_lj(x * _LEN[unit])


def meter_per_sec_to_lj(v):
    return v * math.sqrt(AR_MASS_AMU / EPS_J_PER_MOL)


def vel_to_lj(v, unit1='ang', unit2='pico'):
    # conversion.py:32-35 maps both 'pico' and 'femto' to 1e-12; kept as is.
    return meter_per_sec_to_lj(v * _LEN[unit1] / 1e-12)


def kelvin_to_lj(T):
    return T * KB_J_PER_K_MOL / EPS_J_PER_MOL


def lj_to_kelvin(kBT):
    return kBT * EPS_J_PER_MOL / KB_J_PER_K_MOL


def lj_to_dist(x, unit='ang'):
    return x / _LEN[unit] * SIGMA_M


def lj_to_vel(v, unit1='ang', unit2='pico'):
    return v * 1e-12 / _LEN[unit1] * math.sqrt(EPS_J_PER_MOL / AR_MASS_AMU)
