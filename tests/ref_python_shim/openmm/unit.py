"""STAND-IN for openmm.unit: the mirror works in OpenMM's default units (nm, kJ/mol, e) with plain floats."""


def is_quantity(x):
    return False
