"""Synthetic workloads shaped like BASELINE.json's configs (no benchmark files exist in the
reference and there is no network): a lattice of 3-site waters with a carved-out "solute" made
of small bonded residues linked into a chain (1-2/1-3 exclusions, scaled 1-4 exceptions).

Every workload returns float32-representable positions and box lengths widened to float64, so
the CPU oracle (double) and the CUDA path (float coordinates) see identical inputs.
"""
import math

import numpy as np

from .context import System
from .force import SlicedNonbondedForce

# TIP3P-like water
_QO, _QH = -0.834, 0.417
_SIG_O, _EPS_O = 0.315075, 0.635968
_ROH, _HOH = 0.09572, math.radians(104.52)

_CENTER_TYPES = np.array([(0.339967, 0.359824), (0.325, 0.71128), (0.295992, 0.87864), (0.35, 0.276144)])
_LIGAND_TYPES = np.array([(0.106908, 0.0656888), (0.12, 0.07), (0.09, 0.05), (0.11, 0.0656888)])


class Workload:
    def __init__(self, name, system, force, positions, box, parameters, meta):
        self.name, self.system, self.force = name, system, force
        self.positions, self.box, self.parameters, self.meta = positions, box, parameters, meta

    @property
    def num_particles(self):
        return self.positions.shape[0]


def _f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    return np.stack([np.stack([1-2*(y*y+z*z), 2*(x*y-z*w), 2*(x*z+y*w)], -1),
                     np.stack([2*(x*y+z*w), 1-2*(x*x+z*z), 2*(y*z-x*w)], -1),
                     np.stack([2*(x*z-y*w), 2*(y*z+x*w), 1-2*(x*x+y*y)], -1)], 1)


_TETRA = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]])/math.sqrt(3.0)


def _snake_order(m):
    """Boustrophedon order of the sites of an m[0] x m[1] x m[2] lattice: consecutive sites adjacent."""
    ix, iy, iz = np.meshgrid(np.arange(m[0]), np.arange(m[1]), np.arange(m[2]), indexing="ij")
    iy2 = np.where(ix % 2 == 0, iy, m[1]-1-iy)
    iz2 = np.where((ix*m[1]+iy2) % 2 == 0, iz, m[2]-1-iz)
    key = (ix*m[1]+iy2)*m[2]+iz2
    return np.argsort(key.ravel(), kind="stable")


def solvated(name, box_lengths, lattice, num_atoms, solute_sites, solute_groups, method, seed,
             cutoff=0.9, pme=None, ljpme=None, dispersion_correction=False, scaling=(), derivatives=(),
             switch=None, solute_region="sphere"):
    """Build a water box of `lattice` sites with `solute_sites` sites replaced by solute residues so
    that the total atom count is exactly `num_atoms`.

    solute_groups: fractions of the solute chain assigned to subsets 1..S-1 (solvent is subset 0).
    scaling: list of (name, default, subset1, subset2, includeCoulomb, includeLJ).
    """
    rng = np.random.default_rng(seed)
    box_lengths = _f32(box_lengths)
    m = np.asarray(lattice)
    num_sites = int(m.prod())
    num_waters = num_sites-solute_sites
    solute_atoms = num_atoms-3*num_waters
    assert solute_sites == 0 or solute_atoms >= 2*solute_sites, "too few solute atoms for the carved sites"
    assert solute_sites > 0 or solute_atoms == 0
    spacing = box_lengths/m
    ix, iy, iz = np.meshgrid(np.arange(m[0]), np.arange(m[1]), np.arange(m[2]), indexing="ij")
    sites = (np.stack([ix.ravel(), iy.ravel(), iz.ravel()], -1)+0.5)*spacing
    if solute_region == "sphere":
        dist = np.linalg.norm(sites-0.5*box_lengths, axis=1)
    else:   # slab-ish blob near one corner, exercises periodic wrap of the solute
        dist = np.linalg.norm((sites-0.1*box_lengths+0.5*box_lengths) % box_lengths-0.5*box_lengths, axis=1)
    is_solute = np.zeros(num_sites, dtype=bool)
    is_solute[np.argsort(dist, kind="stable")[:solute_sites]] = True
    order = _snake_order(m)
    solute_idx = order[is_solute[order]]
    water_idx = order[~is_solute[order]]

    # residue sizes: base b everywhere, b+1 on the first r residues
    sizes = np.zeros(0, dtype=np.int64)
    if solute_sites:
        b, r = divmod(solute_atoms, solute_sites)
        assert 2 <= b and b+1 <= 5, "residue size out of range (%d)" % b
        sizes = np.full(solute_sites, b)
        sizes[:r] += 1
    n_sol = int(sizes.sum())
    first = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    res_of = np.repeat(np.arange(solute_sites), sizes)
    slot = np.arange(n_sol)-first[res_of]                    # 0 = centre atom, 1.. = ligands

    pos = np.zeros((num_atoms, 3))
    charge = np.zeros(num_atoms)
    sigma = np.ones(num_atoms)
    epsilon = np.zeros(num_atoms)
    subset = np.zeros(num_atoms, dtype=np.int64)

    if solute_sites:
        rot = _random_rotations(rng, solute_sites)
        centre = sites[solute_idx]+rng.uniform(-0.015, 0.015, size=(solute_sites, 3))
        lig_dir = np.einsum("rij,kj->rki", rot, _TETRA)      # [res][4][3]
        off = np.where(slot[:, None] > 0, 0.09*lig_dir[res_of, np.maximum(slot-1, 0)], 0.0)
        pos[:n_sol] = centre[res_of]+off
        ctype = rng.integers(0, len(_CENTER_TYPES), size=n_sol)
        ltype = rng.integers(0, len(_LIGAND_TYPES), size=n_sol)
        is_c = slot == 0
        sigma[:n_sol] = np.where(is_c, _CENTER_TYPES[ctype, 0], _LIGAND_TYPES[ltype, 0])
        epsilon[:n_sol] = np.where(is_c, _CENTER_TYPES[ctype, 1], _LIGAND_TYPES[ltype, 1])
        q = np.where(is_c, rng.uniform(-0.6, 0.6, n_sol), rng.uniform(-0.3, 0.3, n_sol))
        charge[:n_sol] = np.round(q-q.mean(), 6)
        charge[0] -= charge[:n_sol].sum()                    # exactly neutral solute
        bounds = np.floor(np.cumsum([0.0]+list(solute_groups))*solute_sites).astype(int)
        bounds[-1] = solute_sites
        for g in range(len(solute_groups)):
            subset[:n_sol][(res_of >= bounds[g]) & (res_of < bounds[g+1])] = g+1

    rotw = _random_rotations(rng, num_waters)
    o = sites[water_idx]+rng.uniform(-0.015, 0.015, size=(num_waters, 3))
    h1 = np.array([_ROH*math.sin(_HOH/2), _ROH*math.cos(_HOH/2), 0.0])
    h2 = np.array([-_ROH*math.sin(_HOH/2), _ROH*math.cos(_HOH/2), 0.0])
    w0 = n_sol+3*np.arange(num_waters)
    pos[w0] = o
    pos[w0+1] = o+rotw@h1
    pos[w0+2] = o+rotw@h2
    charge[w0], charge[w0+1], charge[w0+2] = _QO, _QH, _QH
    sigma[w0], epsilon[w0] = _SIG_O, _EPS_O

    force = SlicedNonbondedForce(1+len(solute_groups))
    force.setNonbondedMethod(method)
    force.setCutoffDistance(cutoff)
    force.setUseDispersionCorrection(dispersion_correction)
    if switch is not None:
        force.setUseSwitchingFunction(True)
        force.setSwitchingDistance(switch)
    if pme is not None:
        force.setPMEParameters(*pme)
    if ljpme is not None:
        force.setLJPMEParameters(*ljpme)
    force._particles = np.stack([charge, sigma, epsilon], -1).tolist()
    force._subsets = {int(i): int(s) for i, s in zip(np.nonzero(subset)[0], subset[np.nonzero(subset)[0]])}

    # exceptions: waters (3 exclusions each) + solute chain via sparse graph powers
    exc = []
    if solute_sites:
        import scipy.sparse as sp
        c_idx = first
        bonds_a = [first[res_of[slot > 0]], c_idx[:-1]]
        bonds_b = [np.nonzero(slot > 0)[0], c_idx[1:]]
        a = np.concatenate(bonds_a)
        b_ = np.concatenate(bonds_b)
        adj = sp.coo_matrix((np.ones(2*len(a)), (np.concatenate([a, b_]), np.concatenate([b_, a]))), shape=(n_sol, n_sol)).tocsr()
        adj.data[:] = 1
        a2 = (adj@adj).tocsr()
        a3 = (a2@adj).tocsr()
        within2 = ((adj+a2) > 0).astype(np.int8)
        within3 = ((adj+a2+a3) > 0).astype(np.int8)
        within2.setdiag(0)
        within3.setdiag(0)
        w2 = sp.triu(within2, k=1).tocoo()
        only3 = sp.triu(within3-within3.multiply(within2), k=1).tocoo()
        only3.eliminate_zeros()
        w2.eliminate_zeros()
        excl = np.stack([w2.row, w2.col, np.zeros(w2.nnz), np.ones(w2.nnz), np.zeros(w2.nnz)], -1)
        i14, j14 = only3.row, only3.col
        e14 = np.stack([i14, j14, 0.8333*charge[i14]*charge[j14], 0.5*(sigma[i14]+sigma[j14]),
                        0.5*np.sqrt(epsilon[i14]*epsilon[j14])], -1)
        exc += [excl, e14]
    wexc = np.zeros((num_waters, 3, 5))
    wexc[:, 0, 0], wexc[:, 0, 1] = w0, w0+1
    wexc[:, 1, 0], wexc[:, 1, 1] = w0, w0+2
    wexc[:, 2, 0], wexc[:, 2, 1] = w0+1, w0+2
    wexc[:, :, 3] = 1.0
    exc.append(wexc.reshape(-1, 5))
    exc = np.concatenate(exc)
    force._exceptions = [[int(r[0]), int(r[1]), float(r[2]), float(r[3]), float(r[4])] for r in exc]
    force._exceptionMap = {(min(e[0], e[1]), max(e[0], e[1])): k for k, e in enumerate(force._exceptions)}

    params = {}
    for (pname, default, s1, s2, coul, lj) in scaling:
        if pname not in params:
            force.addGlobalParameter(pname, default)
            params[pname] = default
        force.addScalingParameter(pname, s1, s2, coul, lj)
    for pname in derivatives:
        force.addScalingParameterDerivative(pname)

    system = System()
    system.setNumParticles(num_atoms)
    box = np.diag(box_lengths)
    system.setDefaultPeriodicBoxVectors(box[0], box[1], box[2])
    system.addForce(force)
    meta = {"num_waters": num_waters, "solute_atoms": n_sol, "num_exceptions": len(force._exceptions),
            "lattice": [int(x) for x in m], "seed": seed}
    return Workload(name, system, force, _f32(pos), box, params, meta)


ALPHA = 2.9203      # sqrt(-ln(2*5e-4))/0.9, SURVEY.md section 8


def water_box(method=SlicedNonbondedForce.PME, seed=2024, dispersion_correction=True, subset_molecules=8):
    """C1: 648-atom water box, 2 subsets (first 8 molecules -> subset 1), lambda on slice (0,1)."""
    w = solvated("water648", [1.86206]*3, [6, 6, 6], 648, 0, [], method, seed, pme=(ALPHA, 18, 18, 18),
                 dispersion_correction=dispersion_correction)
    f = w.force
    f._numSubsets = 2
    for mol in range(subset_molecules):
        for k in range(3):
            f.setParticleSubset(3*mol+k, 1)
    f.addGlobalParameter("lambda", 0.8)
    f.addScalingParameter("lambda", 0, 1, True, True)
    f.addScalingParameterDerivative("lambda")
    w.parameters = {"lambda": 0.8}
    return w


def dhfr_like(seed=2025, method=SlicedNonbondedForce.PME):
    """C2: DHFR-size, 23,558 atoms, cubic 6.2208 nm, solute (2,489 atoms, subset 1) + 7,023 waters."""
    return solvated("dhfr23558", [6.2208]*3, [20, 20, 20], 23558, 977, [1.0], method, seed, pme=(ALPHA, 56, 56, 56),
                    scaling=[("lambda", 0.7, 0, 1, True, True)], derivatives=["lambda"])


def apoa1_like(seed=2026):
    """C3: ApoA1-size, 92,224 atoms, 4 subsets / 10 slices, cubic PME grid 98^3 (nx == nz)."""
    return solvated("apoa1_92224", [10.8861, 10.8861, 7.7758], [35, 35, 25], 92224, 6125, [0.3, 0.3, 0.4],
                    SlicedNonbondedForce.PME, seed, pme=(ALPHA, 98, 98, 98),
                    scaling=[("lambda_c", 0.6, 0, 1, True, False), ("lambda_lj", 0.85, 0, 2, False, True),
                             ("lambda_both", 0.4, 1, 2, True, True)],
                    derivatives=["lambda_c", "lambda_lj", "lambda_both"])


def fep_ligand(seed=2027):
    """C4: ligand (30 atoms, subset 1) in water, ~3,000 atoms, LJPME, separate Coulomb / LJ lambdas."""
    alpha_d = math.sqrt(-math.log(2*5e-4))/0.9
    return solvated("fep3000", [3.1]*3, [10, 10, 10], 3000, 10, [1.0], SlicedNonbondedForce.LJPME, seed,
                    pme=(ALPHA, 28, 28, 28), ljpme=(alpha_d, 14, 14, 14),
                    scaling=[("lambda_coul", 0.5, 0, 1, True, False), ("lambda_lj", 0.9, 0, 1, False, True)],
                    derivatives=["lambda_coul", "lambda_lj"])


def stmv_like(seed=2028):
    """C5: STMV-size, 1,066,628 atoms, cubic 21.6832 nm, ~13% "capsid" (subset 1), PME 196^3."""
    return solvated("stmv1066628", [21.6832]*3, [70, 70, 70], 1066628, 34300, [1.0], SlicedNonbondedForce.PME, seed,
                    pme=(ALPHA, 196, 196, 196), scaling=[("lambda", 0.7, 0, 1, True, True)], derivatives=["lambda"])


def small_mixed(n_side=6, solute_sites=12, method=SlicedNonbondedForce.PME, seed=7, num_groups=2, **kw):
    """A small solute+water box (a few hundred atoms) for fast parity tests of every method."""
    num_sites = n_side**3
    atoms = 3*(num_sites-solute_sites)+4*solute_sites-1
    L = 0.31034*n_side
    groups = [1.0/num_groups]*num_groups
    g = max(6, int(math.ceil(2*ALPHA*L/(3*5e-4**0.2))))
    kw.setdefault("pme", (ALPHA, g, g, g))
    return solvated("small%d" % atoms, [L]*3, [n_side]*3, atoms, solute_sites, groups, method, seed, **kw)


WORKLOADS = {"water": water_box, "dhfr": dhfr_like, "apoa1": apoa1_like, "fep": fep_ligand, "stmv": stmv_like}
# atom counts, known without building the workload (bench.py picks the multi-GPU mode from them)
WORKLOAD_ATOMS = {"water": 648, "dhfr": 23558, "apoa1": 92224, "fep": 3000, "stmv": 1066628}
