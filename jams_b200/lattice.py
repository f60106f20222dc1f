"""Host-side lattice and interaction-template construction (pure integer / small float logic).

Mirrors what JAMS does before the solver ever runs, so that the exchange template handed to
``jb_set_exchange_template`` is exactly the one the reference would build:

* site numbering ``((i*Ny + j)*Nz + k)*M + m``          — reference core/lattice.cc:622-657
* boundary wrap / open-boundary rejection               — core/lattice.cc:987-1007
* interaction template processing                        — core/interactions.cc:24-124,292-347
* point-group expansion                                  — core/lattice.cc:1015-1035,1127-1153
  (spglib's operation list is replaced by ``find_space_group_operations``: metric-preserving integer
  rotations of the cell x translations that map the motif onto itself)
* neighbour list                                          — core/interactions.cc:349-395

Everything here is setup code; none of it is on the per-step path.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

from .consts import kBohrMagnetonIU, kGyromagneticRatioIU, kJoule2meV, ENERGY_UNITS

LATTICE_TOLERANCE = 1e-4  # reference helpers/defaults.h:44


# ---- tolerant comparisons (reference helpers/maths.h:16-46) -------------------------------------
def approximately_equal(a: float, b: float, eps: float) -> bool:
    if abs(a - b) <= eps:
        return True
    return abs(a - b) <= max(abs(a), abs(b)) * eps


def approximately_zero(a: float, eps: float) -> bool:
    return abs(a) <= eps


def definately_greater_than(a: float, b: float, eps: float) -> bool:
    return (a - b) > max(abs(a), abs(b)) * eps


def definately_less_than(a: float, b: float, eps: float) -> bool:
    return (b - a) > max(abs(a), abs(b)) * eps


def vec_approximately_equal(a, b, eps) -> bool:
    return all(approximately_equal(float(x), float(y), eps) for x, y in zip(a, b))


def cubic_point_group():
    """48 signed permutation matrices (fractional basis) with zero translations."""
    rots = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1.0, -1.0), repeat=3):
            R = np.zeros((3, 3))
            for r in range(3):
                R[r, perm[r]] = signs[r]
            rots.append(R)
    return np.array(rots), np.zeros((len(rots), 3))


def find_space_group_operations(cell, motif_frac, motif_types, symprec=LATTICE_TOLERANCE):
    """The symmetry operations {R|t} of the crystal in the basis of the given cell -- what the reference asks spglib for
    (spg_get_dataset, core/lattice.cc:783-822: rotations / translations of the INPUT cell).  R runs over the integer matrices
    that preserve the cell's metric (the lattice holohedry, at most 48); for each, every translation that maps the first atom
    onto an atom of its type is tried and kept if the whole motif maps onto itself type by type (cartesian tolerance
    ``symprec`` in lattice parameters, like spglib's symprec).  Returns (rotations (n,3,3), translations (n,3) in [0,1))."""
    A = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    P = np.asarray(motif_frac, dtype=np.float64).reshape(-1, 3)
    types = np.asarray(motif_types)
    G = A.T @ A
    tol = symprec * max(1.0, float(np.sqrt(G.max())))
    lens = np.sqrt(np.diag(G))
    cand = np.array(list(itertools.product(range(-2, 3), repeat=3)), dtype=np.float64)
    clen = np.linalg.norm(cand @ A.T, axis=1)
    cols = [cand[np.abs(clen - lens[k]) <= tol] for k in range(3)]
    rots = []
    for c0 in cols[0]:
        for c1 in cols[1]:
            if abs((A @ c0) @ (A @ c1) - G[0, 1]) > tol * max(lens):
                continue
            for c2 in cols[2]:
                if abs((A @ c0) @ (A @ c2) - G[0, 2]) > tol * max(lens) or abs((A @ c1) @ (A @ c2) - G[1, 2]) > tol * max(lens):
                    continue
                R = np.column_stack([c0, c1, c2])
                if abs(abs(np.linalg.det(R)) - 1.0) < 1e-9:
                    rots.append(R)
    out_R, out_t = [], []
    same = [np.nonzero(types == types[a])[0] for a in range(len(P))]
    for R in rots:
        RP = P @ R.T
        for j in same[0]:
            t = P[j] - RP[0]
            t = t - np.floor(t)
            ok = True
            for a in range(len(P)):
                d = RP[a] + t - P[same[a]]
                d = d - np.rint(d)
                if np.min(np.linalg.norm(d @ A.T, axis=1)) > symprec:
                    ok = False
                    break
            if ok:
                t = np.where(np.abs(t - 1.0) <= symprec, 0.0, t)
                if not any(np.array_equal(R, r) and np.allclose(t, u, atol=symprec) for r, u in zip(out_R, out_t)):
                    out_R.append(R)
                    out_t.append(t)
    return np.array(out_R), np.array(out_t)


def normalise_fractional_coordinate(r, eps=LATTICE_TOLERANCE):
    """reference core/lattice.cc:48-64"""
    r = [float(v) for v in r]
    for n in range(3):
        if r[n] < 0.0:
            r[n] = r[n] + 1.0
        if approximately_equal(r[n], 1.0, eps):
            r[n] = 0.0
    return r


def lattice_translation_vector(r_frac, tolerance):
    """reference core/interactions.cc:58-76 (round to nearest even like std::nearbyint)"""
    T = []
    for n in range(3):
        x = float(r_frac[n])
        nearest = float(np.rint(x))
        floored = math.floor(x)
        T.append(nearest if approximately_zero(x - nearest, tolerance) else float(floored))
    return T


@dataclass
class Material:
    """reference containers/material.h:16-60 (moment in Bohr magnetons, gyro as a fraction of gamma_e)"""
    name: str
    moment: float
    gyro: float = 1.0
    alpha: float = 0.01
    spin: tuple = (0.0, 0.0, 1.0)


def read_interaction_file(path):
    """``exc_file`` (hamiltonian/exchange.cc:118-133): the text table of discover_interaction_file_format /
    interactions_from_file (core/interactions.cc:126-171,205-252).  Lines that are empty or start with ``#`` / ``//`` are
    skipped (helpers/utils.h:115-126); the first data line fixes the format for the whole file: 6 columns = scalar J,
    14 columns = 3x3 tensor (row-major), anything else is an error; the first two columns are material names (JAMS
    format) or 1-based motif indices (KKR format, both must be unsigned integers).  Returns the list that the
    ``interactions`` setting would hold."""
    def is_comment(line):
        t = line.split()
        return (not t) or t[0][0] == "#" or t[0][:2] == "//"

    def is_int(tok):   # string_is_int (helpers/utils.h:158-160)
        return all(ch in "0123456789" for ch in tok)

    try:
        lines = open(path).read().splitlines()
    except OSError:
        raise RuntimeError(f"{path}: failed to open file")
    ncols, kkr = None, None
    for line in lines:
        if is_comment(line):
            continue
        tok = line.split()
        if len(tok) not in (6, 14):
            raise RuntimeError("interaction file has an incorrect number of columns")
        ncols = len(tok)
        if is_int(tok[0]) != is_int(tok[1]):
            break
        kkr = is_int(tok[0])
        break
    if kkr is None:
        raise RuntimeError("failed to discover interaction file format")
    out = []
    for n, line in enumerate(lines):
        if is_comment(line):
            continue
        tok = line.split()
        try:
            ti, tj = (int(tok[0]), int(tok[1])) if kkr else (tok[0], tok[1])
            nums = [float(v) for v in tok[2:ncols]]
            if len(nums) != ncols - 2:
                raise ValueError
        except (ValueError, IndexError):
            raise RuntimeError(f"failed to read line {n} of interaction file")
        out.append((ti, tj, nums[:3], nums[3] if ncols == 6 else nums[3:12]))
    return out


class Pcg32:
    """pcg32 = setseq_xsh_rr_64_32 with the default stream (M. O'Neill, pcg-random.org, pcg_random.hpp): 64-bit LCG state,
    output = rotr32(((old >> 18) ^ old) >> 27, old >> 59) of the state BEFORE the step; seeding state = bump(seed + increment).
    uniform_real(): libstdc++'s std::generate_canonical<double, 53> over a 32-bit generator: two draws, (g0 + g1 2^32) / 2^64."""
    MULT, INC, MASK = 6364136223846793005, 1442695040888963407, (1 << 64) - 1

    def __init__(self, seed):
        self.state = (((int(seed) + self.INC) & self.MASK) * self.MULT + self.INC) & self.MASK

    def __call__(self):
        old = self.state
        self.state = (old * self.MULT + self.INC) & self.MASK
        x = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        return ((x >> rot) | (x << ((-rot) & 31))) & 0xffffffff

    def uniform_real(self):
        total, tmp = 0.0, 1.0
        for _ in range(2):
            total += float(self()) * tmp
            tmp *= 4294967296.0
        ret = total / tmp
        return ret if ret < 1.0 else float(np.nextafter(1.0, 0.0))


@dataclass
class Lattice:
    """Materials + unit cell + supercell (reference core/lattice.h)."""
    materials: list
    cell: np.ndarray                  # 3x3, columns are a, b, c (core/lattice.cc:356-367)
    motif: list                       # [(material name, (fx, fy, fz)), ...]
    dims: tuple
    periodic: tuple = (True, True, True)
    gilbert_prefactor: bool = False
    symops: tuple | None = None       # (rotations (n,3,3) fractional, translations (n,3)); None = found from the cell and motif
    impurities: list | None = None    # lattice.impurities: [(materialA, materialB, fraction), ...] (core/lattice.cc:424-427,1077-1109)
    impurities_seed: int = 0          # lattice.impurities_seed (the reference draws one from its global generator if absent)
    # "Rotating the system" (core/lattice.cc:434-454,515-575): orientation first (the lattice vector / cartesian vector is turned onto
    # orientation_axis), then global_rotation; both rotate the unit-cell vectors, a_k <- R a_k
    orientation_axis: tuple | None = None
    orientation_lattice_vector: tuple | None = None
    orientation_cartesian_vector: tuple | None = None
    global_rotation: np.ndarray | None = None

    def __post_init__(self):
        self.cell = np.asarray(self.cell, dtype=np.float64).reshape(3, 3)
        if self.orientation_axis is not None:
            if self.orientation_lattice_vector is not None and self.orientation_cartesian_vector is not None:
                raise RuntimeError("Only one of 'orientation_lattice_vector' or 'orientation_cartesian_vector' can be defined")
            vec = None
            if self.orientation_lattice_vector is not None:
                vec = np.asarray(self.orientation_lattice_vector, dtype=np.float64)
            elif self.orientation_cartesian_vector is not None:
                vec = np.linalg.inv(self.cell) @ np.asarray(self.orientation_cartesian_vector, dtype=np.float64)
            if vec is not None:   # global_reorientation (:535-575)
                cart = self.cell @ vec
                cart = cart / np.sqrt(cart @ cart)
                self._rotate_cell(rotation_matrix_between_vectors(cart, np.asarray(self.orientation_axis, dtype=np.float64)))
        if self.global_rotation is not None:   # global_rotation (:515-533)
            self._rotate_cell(np.asarray(self.global_rotation, dtype=np.float64).reshape(3, 3))
        self.cell_inv = np.linalg.inv(self.cell)
        self.material_index = {m.name: i for i, m in enumerate(self.materials)}
        self.motif_material = np.array([self.material_index[name] for name, _ in self.motif], dtype=np.int32)
        self.motif_frac = np.array([normalise_fractional_coordinate(p) for _, p in self.motif], dtype=np.float64)
        self.dims = tuple(int(d) for d in self.dims)
        self.periodic = tuple(bool(p) for p in self.periodic)
        if self.symops is None:   # the reference asks spglib (core/lattice.cc:783-822); here: find_space_group_operations
            self.symops = find_space_group_operations(self.cell, self.motif_frac, self.motif_material)
        # read_impurities_from_config (core/lattice.cc:1077-1109)
        self.impurity_map = {}
        for n, (a, b, fraction) in enumerate(self.impurities or []):
            if a not in self.material_index:
                raise RuntimeError(f"impurity {n} materialA ({a}) does not exist")
            if b not in self.material_index:
                raise RuntimeError(f"impurity {n} materialB ({b}) does not exist")
            if fraction < 0.0 or fraction >= 1.0:
                raise RuntimeError(f"impurity {n} fraction must be 0 =< x < 1")
            if self.material_index[a] in self.impurity_map:
                raise RuntimeError(f"impurity {n} redefines a previous impurity")
            self.impurity_map[self.material_index[a]] = (self.material_index[b], float(fraction))
        self._site_material_all = None

    def _rotate_cell(self, R):
        before = abs(np.linalg.det(self.cell))
        self.cell = R @ self.cell          # columns are a, b, c: a_k <- R a_k (containers/cell.cc:62-68)
        if abs(abs(np.linalg.det(self.cell)) - before) > max(before, 1.0) * LATTICE_TOLERANCE ** 3:
            raise RuntimeError("unitcell volume has changed after rotation")

    # -- sizes
    @property
    def M(self):
        return len(self.motif)

    @property
    def num_spins(self):
        return self.dims[0] * self.dims[1] * self.dims[2] * self.M

    def site_index(self, i, j, k, m):
        return ((i * self.dims[1] + j) * self.dims[2] + k) * self.M + m

    def apply_boundary_conditions(self, abc):
        """reference core/lattice.cc:987-1007; returns wrapped cell or None if outside an open boundary"""
        out = []
        for l in range(3):
            if not self.periodic[l] and (abc[l] < 0 or abc[l] >= self.dims[l]):
                return None
            out.append((abc[l] + self.dims[l]) % self.dims[l])
        return out

    # -- per-site arrays in reference site order for the slab x in [x0, x0 + nx)
    def _tile(self, per_motif, x0=0, nx=None):
        nx = self.dims[0] if nx is None else nx
        cells = nx * self.dims[1] * self.dims[2]
        return np.tile(np.asarray(per_motif), (cells,) + (1,) * (np.ndim(per_motif) - 1))

    @property
    def has_impurities(self):   # core/lattice.cc:1115-1117
        return bool(self.impurity_map)

    def _site_materials(self):
        """generate_supercell's substitution loop (core/lattice.cc:614-640): sites in the order (i, j, k, m); a site whose motif
        material has an impurity entry draws one uniform number (std::uniform_real_distribution<> over pcg32(impurity_seed)) and
        becomes materialB if it is below the fraction.  pcg32 and libstdc++'s generate_canonical are restated from their
        published definitions (the reference fetches pcg at configure time; it is not in the tree): the STREAM is unpinned, the
        structure that follows from a given set of site materials is checked against the oracle."""
        if self._site_material_all is None:
            mat = self._tile(self.motif_material).astype(np.int32)
            if self.impurity_map:
                rng = Pcg32(self.impurities_seed)
                for site in np.nonzero(np.isin(mat, list(self.impurity_map)))[0]:
                    b, fraction = self.impurity_map[int(mat[site])]
                    if rng.uniform_real() < fraction:
                        mat[site] = b
            self._site_material_all = mat
        return self._site_material_all

    def _per_site(self, per_material, x0=0, nx=None):
        return np.asarray(per_material)[self.site_material(x0, nx)]

    def site_material(self, x0=0, nx=None):
        if not self.impurity_map:
            return self._tile(self.motif_material, x0, nx).astype(np.int32)
        nx = self.dims[0] if nx is None else nx
        per_plane = self.dims[1] * self.dims[2] * self.M
        return self._site_materials()[x0 * per_plane:(x0 + nx) * per_plane]

    def site_motif(self, x0=0, nx=None):
        return self._tile(np.arange(self.M, dtype=np.int32), x0, nx)

    def mus(self, x0=0, nx=None):
        """globals::mus = moment * mu_B (containers/material.h:32)"""
        return self._per_site([m.moment * kBohrMagnetonIU for m in self.materials], x0, nx)

    def alpha(self, x0=0, nx=None):
        return self._per_site([m.alpha for m in self.materials], x0, nx)

    def gyro(self, x0=0, nx=None):
        """globals::gyro (containers/material.h:33, core/lattice.cc:91-97,709-713)"""
        vals = []
        for mat in self.materials:
            g = mat.gyro * kGyromagneticRatioIU
            if self.gilbert_prefactor:
                g = g / (1.0 + mat.alpha * mat.alpha)
            vals.append(g)
        return self._per_site(vals, x0, nx)

    def positions(self, x0=0, nx=None):
        """cartesian site positions in lattice constants (core/lattice.cc:751-756)"""
        nx = self.dims[0] if nx is None else nx
        ii, jj, kk, mm = np.meshgrid(np.arange(x0, x0 + nx), np.arange(self.dims[1]), np.arange(self.dims[2]),
                                     np.arange(self.M), indexing="ij")
        frac = self.motif_frac[mm.reshape(-1)] + np.stack([ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)], axis=1)
        return frac @ self.cell.T

    def initial_spins(self, x0=0, nx=None, seed=None):
        """material.spin normalised (core/lattice.cc:715-733), or seeded uniform-on-sphere spins (seed given)"""
        nx = self.dims[0] if nx is None else nx
        n = nx * self.dims[1] * self.dims[2] * self.M
        if seed is None:
            per = []
            for mat in self.materials:
                s = np.asarray(mat.spin, dtype=np.float64)
                nrm = np.sqrt(s @ s)
                per.append(s / nrm if nrm > np.finfo(float).eps else s)
            return np.ascontiguousarray(self._per_site(per, x0, nx))
        # decomposition-independent: generate the whole lattice stream and slice (test / bench inputs only)
        rng = np.random.default_rng(seed)
        v = rng.standard_normal((self.num_spins, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        per_plane = self.dims[1] * self.dims[2] * self.M
        return np.ascontiguousarray(v[x0 * per_plane:(x0 + nx) * per_plane])

    # -- interaction template (core/interactions.cc:292-347)
    def point_group_of_motif(self, m):
        """reference core/lattice.cc:1127-1153"""
        rots, trans = self.symops
        out = []
        p = self.motif_frac[m]
        for R, t in zip(rots, trans):
            if not all(approximately_zero(float(x), LATTICE_TOLERANCE) for x in t):
                continue
            new_position = normalise_fractional_coordinate(R @ p)
            if vec_approximately_equal(p, new_position, LATTICE_TOLERANCE):
                out.append(R)
        return out

    def expand_interactions(self, interactions, *, energy_units="joules", coordinate_format="cartesian", use_symops=True,
                            energy_cutoff=0.0, radius_cutoff=100.0, distance_tolerance=LATTICE_TOLERANCE,
                            interaction_prefactor=1.0):
        """``interactions``: [(type_i, type_j, (rx,ry,rz), J)], type names (JAMS format) or 1-based motif
        indices (KKR format); J scalar or 9 numbers in ``energy_units``.  Returns the processed template
        dict(mi, mj, T, J9) with J9 in meV (scaled as hamiltonian/exchange.cc:165-167)."""
        unit = ENERGY_UNITS[energy_units]
        kkr = isinstance(interactions[0][0], (int, np.integer))
        entries = []
        for ti, tj, r, J in interactions:
            r = np.asarray(r, dtype=np.float64)
            if coordinate_format.lower() == "fractional":
                r = self.cell @ r
            J = np.asarray(J, dtype=np.float64)
            J9 = (float(J) * np.eye(3)).reshape(9) if J.ndim == 0 else J.reshape(9)
            if kkr:
                entries.append(dict(mi=int(ti) - 1, mj=int(tj) - 1, r=r, J9=J9))
            else:
                entries.append(dict(ti=self.material_index[ti], tj=self.material_index[tj], r=r, J9=J9))
        if not kkr:  # complete_interaction_unitcell_positions (:98-124)
            new = []
            for e in entries:
                for i in range(self.M):
                    if self.motif_material[i] != e["ti"]:
                        continue
                    q = self.cell_inv @ e["r"] + self.motif_frac[i]
                    T = lattice_translation_vector(q, distance_tolerance)
                    offset = q - np.asarray(T)
                    partner = None
                    for k in range(self.M):
                        if vec_approximately_equal(self.motif_frac[k], offset, distance_tolerance):
                            partner = k
                            break
                    if partner is None or self.motif_material[partner] != e["tj"]:
                        continue
                    new.append(dict(mi=i, mj=partner, r=e["r"], J9=e["J9"]))
            entries = new
        if use_symops:  # apply_symops (:24-37) + generate_symmetric_points (core/lattice.cc:1015-1035)
            groups = [self.point_group_of_motif(m) for m in range(self.M)]
            new = []
            for e in entries:
                r_frac = self.cell_inv @ e["r"]
                pts = [e["r"]]
                for R in groups[e["mi"]]:
                    r_sym = self.cell @ (R @ r_frac)
                    if not any(vec_approximately_equal(r_sym, v2, LATTICE_TOLERANCE) for v2 in pts):
                        pts.append(r_sym)
                for p in pts:
                    new.append(dict(mi=e["mi"], mj=e["mj"], r=p, J9=e["J9"]))
            entries = new
        if energy_cutoff > 0.0:
            entries = [e for e in entries if not definately_less_than(float(np.max(np.abs(e["J9"]))), energy_cutoff, np.finfo(float).eps)]
        if radius_cutoff > 0.0:
            entries = [e for e in entries if not definately_greater_than(float(np.sqrt(e["r"] @ e["r"])), radius_cutoff, LATTICE_TOLERANCE)]
        mi, mj, Ts, J9s = [], [], [], []
        for e in entries:
            q = self.cell_inv @ e["r"] + self.motif_frac[e["mi"]] - self.motif_frac[e["mj"]]
            T = lattice_translation_vector(q, distance_tolerance)
            Jij = interaction_prefactor * unit * e["J9"]
            # hamiltonian/exchange.cc:166: keep only if max_abs(Jij) > energy_cutoff * unit
            if not (float(np.max(np.abs(Jij))) > energy_cutoff * unit):
                continue
            mi.append(e["mi"]); mj.append(e["mj"]); Ts.append([int(T[0]), int(T[1]), int(T[2])]); J9s.append(Jij)
        return dict(mi=np.array(mi, np.int32), mj=np.array(mj, np.int32), T=np.array(Ts, np.int32).reshape(-1, 3),
                    J9=np.array(J9s, np.float64).reshape(-1, 9))

    # -- exchange-functional: J(r_ij) within a cutoff (hamiltonian/exchange_functional.cc:94-252)
    def max_interaction_radius(self):
        """Lattice::max_interaction_radius / jams::maximum_interaction_length (core/lattice.cc:1009-1011,1155-1200), in lattice
        parameters: inradius of the supercell over the periodic directions, longest body diagonal for an open system"""
        a = [self.cell[:, k] * self.dims[k] for k in range(3)]
        per = self.periodic
        height = lambda u, v, w: abs(np.dot(np.cross(u, v), w)) / np.linalg.norm(np.cross(u, v))   # noqa: E731  parallelepiped_height
        pheight = lambda u, v: np.linalg.norm(np.cross(u, v)) / np.linalg.norm(u)                   # noqa: E731  parallelogram_height
        if all(per):
            return 0.5 * min(height(a[0], a[1], a[2]), height(a[2], a[0], a[1]), height(a[1], a[2], a[0]))
        if sum(per) == 2:
            u, v = [a[k] for k in range(3) if per[k]]
            return 0.5 * min(pheight(u, v), pheight(v, u))
        if sum(per) == 1:
            return 0.5 * np.linalg.norm(a[per.index(True)])
        return max(np.linalg.norm(a[0] + a[1] + a[2]), np.linalg.norm(-a[0] + a[1] + a[2]), np.linalg.norm(a[0] - a[1] + a[2]),
                   np.linalg.norm(a[0] + a[1] - a[2]))

    def functional_template(self, functionals, tolerance=LATTICE_TOLERANCE):
        """``functionals``: {(material i, material j): (r_cutoff, J(r_ij) -> meV)} with lengths in lattice parameters.  The
        reference walks every site's neighbours inside the largest cutoff through a near-tree over the supercell and inserts
        J(r_ij) for each ordered pair (i, j != i) whose material pair has a functional and whose distance is within that
        pair's cutoff (less_than_approx_equal, relative tolerance 1e-4; exchange_functional.cc:206-243).  On a lattice
        without impurities that list is translation invariant, so it is generated here as a template (motif i, motif j,
        cell offset T, J) -- the form the kernels consume."""
        if not functionals:
            return dict(mi=np.zeros(0, np.int32), mj=np.zeros(0, np.int32), T=np.zeros((0, 3), np.int32), J9=np.zeros((0, 9)))
        rmax = max(rc for rc, _ in functionals.values())
        names = [m.name for m in self.materials]
        mi_l, mj_l, T_l, J_l = [], [], [], []
        for mi, mj, T, rij, r in self.motif_pairs_within(rmax * (1 + tolerance)):
            key = (names[self.motif_material[mi]], names[self.motif_material[mj]])
            if key not in functionals:
                continue
            rc, fn = functionals[key]
            if (r - rc) < max(abs(r), abs(rc)) * tolerance:   # less_than_approx_equal (helpers/maths.h:37-40)
                mi_l.append(mi); mj_l.append(mj); T_l.append(T); J_l.append(float(fn(rij)) * np.eye(3).reshape(9))
        return dict(mi=np.array(mi_l, np.int32), mj=np.array(mj_l, np.int32), T=np.array(T_l, np.int32).reshape(-1, 3),
                    J9=np.array(J_l, np.float64).reshape(-1, 9))

    def motif_pairs_within(self, radius):
        """every (motif i, motif j, cell offset T != self) with |r_ij| possibly <= radius: yields (mi, mj, T, r_ij, |r_ij|), lengths
        in lattice parameters.  |T_k| <= radius / (spacing of the lattice planes normal to a_k) + 1 covers the sphere."""
        vol = abs(np.linalg.det(self.cell))
        nmax = []
        for k in range(3):
            u, v = self.cell[:, (k + 1) % 3], self.cell[:, (k + 2) % 3]
            nmax.append(int(np.ceil(radius / (vol / np.linalg.norm(np.cross(u, v))))) + 1)
        for mi in range(self.M):
            ri = self.cell @ self.motif_frac[mi]
            for mj in range(self.M):
                for tx in range(-nmax[0], nmax[0] + 1):
                    for ty in range(-nmax[1], nmax[1] + 1):
                        for tz in range(-nmax[2], nmax[2] + 1):
                            if mi == mj and tx == 0 and ty == 0 and tz == 0:
                                continue   # no self interaction (exchange_functional.cc:213-215, exchange_neartree.cc:119-121)
                            rij = self.cell @ (self.motif_frac[mj] + np.array([tx, ty, tz], dtype=np.float64)) - ri
                            yield mi, mj, (tx, ty, tz), rij, float(np.linalg.norm(rij))

    def shell_template(self, shells, shell_width, energy_cutoff):
        """exchange-neartree (hamiltonian/exchange_neartree.cc:100-128): ``shells`` = [(material id A, material id B, radius, J meV)]
        (already mirrored for A != B).  A site of material A couples with J to every site of material B whose distance lies in
        the annulus radius -+ shell_width / 2 (InteractionNearTree::shell -> NearTree::in_annulus, containers/neartree.h:359-376,
        relative tolerance shell_width / 10); a pair reached twice is an error; |J| <= energy_cutoff is dropped."""
        if not shells:
            return dict(mi=np.zeros(0, np.int32), mj=np.zeros(0, np.int32), T=np.zeros((0, 3), np.int32), J9=np.zeros((0, 9)))
        eps = shell_width / 10.0
        gt = lambda a, b: (a - b) > max(abs(a), abs(b)) * eps   # noqa: E731  definately_greater_than
        rmax = max(sh[2] for sh in shells) + shell_width
        seen, mi_l, mj_l, T_l, J_l = set(), [], [], [], []
        for mi, mj, T, rij, r in self.motif_pairs_within(rmax * (1 + eps)):
            for A, B, radius, J in shells:
                if self.motif_material[mi] != A or self.motif_material[mj] != B:
                    continue
                inner, outer = radius - 0.5 * shell_width, radius + 0.5 * shell_width
                if gt(r, outer) or not gt(r, inner):
                    continue
                if (mi, mj, T) in seen:
                    raise RuntimeError(f"multiple interactions between spins of motif positions {mi} and {mj}")
                seen.add((mi, mj, T))
                if abs(J) > energy_cutoff:
                    mi_l.append(mi); mj_l.append(mj); T_l.append(T); J_l.append(J * np.eye(3).reshape(9))
        return dict(mi=np.array(mi_l, np.int32), mj=np.array(mj_l, np.int32), T=np.array(T_l, np.int32).reshape(-1, 3),
                    J9=np.array(J_l, np.float64).reshape(-1, 9))

    def neighbour_list(self, template):
        """neighbour_list_from_interactions (core/interactions.cc:349-395) for a lattice without impurities.
        Returns (i, j, value_id, values9) in jams::InteractionList order (pairs sorted by {i,j}, values in
        first-insertion order).  Vectorised over cells; raises on duplicate pairs like the reference."""
        nx, ny, nz = self.dims
        M = self.M
        ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        ii, jj, kk = ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)
        all_i, all_j, all_v = [], [], []
        smat = self._site_materials() if self.impurity_map else None
        raw_i, raw_j = [], []   # with impurities: all pairs before the material check (the reference looks for duplicates first)
        n_t = len(template["mi"])
        first_pos = np.full(n_t, -1, np.int64)   # position of each entry's first insertion in generation order
        for n in range(n_t):
            d = [ii + template["T"][n, 0], jj + template["T"][n, 1], kk + template["T"][n, 2]]
            ok = np.ones(ii.shape, bool)
            for l, size in enumerate(self.dims):
                if not self.periodic[l]:
                    ok &= (d[l] >= 0) & (d[l] < size)
                d[l] = (d[l] + size) % size
            local = ((ii * ny + jj) * nz + kk) * M + int(template["mi"][n])
            nbr = ((d[0] * ny + d[1]) * nz + d[2]) * M + int(template["mj"][n])
            if smat is not None:
                # "catch if the site has a different material (presumably an impurity site)" (core/interactions.cc:381-385): the
                # entry's types are those of its motif positions
                raw_i.append(local[ok]); raw_j.append(nbr[ok])
                ok = ok & (smat[local] == self.motif_material[int(template["mi"][n])]) & (smat[nbr] == self.motif_material[int(template["mj"][n])])
            all_i.append(local[ok]); all_j.append(nbr[ok])
            all_v.append(np.full(int(ok.sum()), n, np.int64))
            if ok.any():
                first_pos[n] = int(np.argmax(ok)) * n_t + n   # generation order: cell-major, entry-minor
        if not all_i:
            return (np.zeros(0, np.int32),) * 3 + (np.zeros((0, 9)),)
        I = np.concatenate(all_i); Jn = np.concatenate(all_j); E = np.concatenate(all_v)
        # unique values in first-insertion order (containers/unordered_vector_set.h:38-45)
        values, value_of = [], {}
        for n in sorted((n for n in range(n_t) if first_pos[n] >= 0), key=lambda n: first_pos[n]):
            key = tuple(template["J9"][n])
            if key not in value_of:
                value_of[key] = len(values)
                values.append(key)
        vid_of_entry = np.array([value_of.get(tuple(template["J9"][n]), -1) for n in range(n_t)], np.int32)
        order = np.lexsort((Jn, I))
        I, Jn, V = I[order], Jn[order], vid_of_entry[E[order]]
        if smat is not None and raw_i:
            Ri, Rj = np.concatenate(raw_i), np.concatenate(raw_j)
            ro = np.lexsort((Rj, Ri))
            Ri, Rj = Ri[ro], Rj[ro]
            rdup = (Ri[1:] == Ri[:-1]) & (Rj[1:] == Rj[:-1])
            if rdup.any():
                p = int(np.nonzero(rdup)[0][0])
                raise RuntimeError(f"Multiple interactions for sites {int(Ri[p])} and {int(Rj[p])}")
        dup = (I[1:] == I[:-1]) & (Jn[1:] == Jn[:-1])
        if dup.any():
            p = int(np.nonzero(dup)[0][0])
            raise RuntimeError(f"Multiple interactions for sites {int(I[p])} and {int(Jn[p])}")
        return I.astype(np.int32), Jn.astype(np.int32), V.astype(np.int32), np.array(values, np.float64).reshape(-1, 9)


def bloch_domain_wall(positions, spins, width, center, normal=(1, 0, 0), domain=(0, 0, 1)):
    """InitBlochDomainWall::execute (reference initializer/init_bloch_domain_wall.cc:10-32): rotate every spin
    by the rotation that takes ``domain`` to m(x) = (0, sech(pi x/w), tanh(pi x/w))."""
    normal = np.asarray(normal, float); normal = normal / np.linalg.norm(normal)
    domain = np.asarray(domain, float); domain = domain / np.linalg.norm(domain)
    out = np.empty_like(spins)
    x = positions @ normal - center
    my = 1.0 / np.cosh(np.pi * x / width)
    mz = np.tanh(np.pi * x / width)
    for i in range(len(spins)):
        out[i] = rotation_matrix_between_vectors(domain, np.array([0.0, my[i], mz[i]])) @ spins[i]
    return out


def load_spins_tsv(path, num_spins):
    """``lattice.spins = "file"`` through the text route of the reference's loader (core/lattice.cc:738-748,
    helpers/load.h:21-61): whitespace-separated numbers, empty lines and lines starting with ``#`` or ``//`` skipped,
    the element count must match.  Returns N x 3."""
    if str(path).endswith(".h5"):
        raise RuntimeError("lattice.spins: HDF5 is not available in this build; give the whitespace-separated text form")
    vals = []
    try:
        fh = open(path)
    except OSError:
        raise RuntimeError("failed to open file: " + str(path))
    with fh:
        for line in fh:
            t = line.strip()
            if not t or t.startswith("#") or t.startswith("//"):
                continue
            vals.extend(float(v) for v in t.split())
    if len(vals) != 3 * num_spins:
        raise RuntimeError(f"loading array from file: '{path}' expected size: {3 * num_spins} actual size: {len(vals)}")
    return np.asarray(vals, dtype=np.float64).reshape(num_spins, 3)


def write_spins_tsv(path, s_aos, iteration=0, time_ps=0.0):
    """the snapshot format of the C++ host's ``hdf5`` / ``spins-tsv`` monitor stand-in (jams_b200/host/jams_host.h):
    one spin per line, 17 significant digits (round-trips exactly), ``#`` header"""
    s = np.asarray(s_aos, dtype=np.float64).reshape(-1, 3)
    with open(path, "w") as fh:
        fh.write(f"# spins {len(s)} x 3   iteration {iteration}   time_ps {time_ps:.17g}\n")
        for row in s:
            fh.write("%.17g %.17g %.17g\n" % tuple(row))


def rotation_matrix_between_vectors(a, b):
    """reference containers/mat3.h:334-366"""
    def unit(v):
        n = np.sqrt(v @ v)
        return v if n <= np.finfo(float).eps else v / n

    def ssc(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)

    ua, ub = unit(np.asarray(a, float)), unit(np.asarray(b, float))
    c = float(ua @ ub)
    if approximately_equal(c, 1.0, 1e-12):
        return np.eye(3)
    if approximately_equal(c, -1.0, 1e-12):
        ortho = np.array([1.0, 0, 0]) if abs(ua[0]) < 0.9 else np.array([0, 1.0, 0])
        axis = unit(np.cross(ua, ortho))
        vx = ssc(unit(axis))
        return np.eye(3) + math.sin(math.pi) * vx + ((1.0 - math.cos(math.pi)) * vx) @ vx
    v = np.cross(ua, ub)
    s = np.sqrt(v @ v)
    vx = ssc(v)
    k = (1.0 - c) / (s * s)
    return np.eye(3) + vx + (k * vx) @ vx
