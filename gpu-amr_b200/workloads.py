"""Synthetic meshes and initial conditions of the reference's drivers, as host-side
numpy: leaf flags for the scripted refinements and the cell-centre initial conditions that
amr_solver::initialize evaluates (solver/amr_solver.hpp:105-145, physics_system.hpp:88-138).
Used by bench.py and the tests to put the same workload on the device pool."""
import numpy as np

from . import binding as B


class Config:
    """A reference instantiation (patch shape, halo, morton depth, equation) as runtime values."""

    def __init__(self, rank, size, halo, depth, eq, length=None, gamma=1.4, cfl=0.3):
        self.rank, self.size, self.halo, self.depth, self.eq = rank, size, halo, depth, eq
        self.length = (1000.0 if eq == B.EQ_EULER else 1.0) if length is None else length
        self.gamma, self.cfl = gamma, cfl
        self.nvar = 1 if eq == B.EQ_ADVECTION else rank + 2
        self.psize = size + 2 * halo
        self.flat = self.psize ** rank
        self.data = size ** rank
        self.ndir = 2 * rank
        self.kf = 1 << (rank - 1)

    @property
    def name(self):
        return "r%d_s%d_h%d_d%d_%s" % (self.rank, self.size, self.halo, self.depth,
                                        "euler" if self.eq == B.EQ_EULER else "adv")


def morton_decode(ids, rank):
    """ids -> (anchor coords [n, rank] in finest-level units, x first; level [n])
    (morton/morton_id.hpp:21-229: id = interleave(x,y[,z]) << 6 | level)."""
    ids = np.asarray(ids, dtype=np.uint64)
    level = (ids & np.uint64(63)).astype(np.int64)
    m = ids >> np.uint64(6)
    coords = np.zeros((len(ids), rank), dtype=np.uint64)
    for b in range(20):
        for a in range(rank):
            coords[:, a] |= ((m >> np.uint64(rank * b + a)) & np.uint64(1)) << np.uint64(b)
    return coords.astype(np.int64), level


def patch_centres(ids, cfg):
    """patch centre / L per physical axis: [n, rank]"""
    coords, lvl = morton_decode(ids, cfg.rank)
    span = float(1 << cfg.depth)
    ext = (2.0 ** (cfg.depth - lvl)) / span
    return coords / span + 0.5 * ext[:, None], lvl


def flags_refine_all(ids):
    return np.full(len(ids), B.REFINE, dtype=np.int8)


def flags_ball(ids, cfg, radius, max_level, centre):
    """Refine leaves whose centre lies within `radius`*L of `centre`*L (static multi-level meshes
    of SURVEY 8d C2/C3)."""
    c, lvl = patch_centres(ids, cfg)
    r2 = ((c - np.asarray(centre)[None, :cfg.rank]) ** 2).sum(axis=1)
    f = np.zeros(len(ids), dtype=np.int8)
    f[(r2 < radius * radius) & (lvl < max_level)] = B.REFINE
    return f


def cell_centres(ids, cfg):
    """[P, S.., rank] physical cell-centre coordinates; physical axis d <-> layout dim rank-1-d."""
    coords, lvl = morton_decode(ids, cfg.rank)
    R, S, L = cfg.rank, cfg.size, cfg.length
    out = np.zeros((len(ids),) + (S,) * R + (R,))
    for d in range(R):
        org = L * coords[:, d].astype(np.float64) / float(1 << cfg.depth)
        dx = (L * (2.0 ** (cfg.depth - lvl)) / float(1 << cfg.depth)) / float(S)
        shape = [len(ids)] + [1] * R
        kshape = [1] * (R + 1)
        kshape[1 + (R - 1 - d)] = S
        k = np.arange(S, dtype=np.float64).reshape(kshape)
        out[..., d] = org.reshape(shape) + (k + 0.5) * dx.reshape(shape)
    return out


def acoustic_pulse(ids, cfg, chunk=65536):
    """Conservative interior state [nvar, P, S..] of the benchmark pulse
    (benchmark/bench_fvm_solver_integration.b.cpp:146-178): rho = 0.5 + 2 g, u = 0,
    p = 1 + 10 g, g = exp(-r^2 / (0.01 L^2)) about the domain centre."""
    R, L = cfg.rank, cfg.length
    out = np.zeros((cfg.nvar, len(ids)) + (cfg.size,) * R)
    for s in range(0, len(ids), chunk):
        x = cell_centres(ids[s:s + chunk], cfg)
        r2 = sum((x[..., d] - 0.5 * L) ** 2 for d in range(R))
        g = np.exp(-r2 / (0.01 * L * L))
        out[0, s:s + chunk] = 0.5 + 2.0 * g
        out[R + 1, s:s + chunk] = (1.0 + 10.0 * g) / (cfg.gamma - 1.0)
    return out


def gaussian_scalar(ids, cfg, chunk=65536):
    """examples/fvm_solver_advection.e.cpp:57-63: exp(-|x - 0.2 L|^2 / (0.005 L^2))"""
    R, L = cfg.rank, cfg.length
    out = np.zeros((1, len(ids)) + (cfg.size,) * R)
    for s in range(0, len(ids), chunk):
        x = cell_centres(ids[s:s + chunk], cfg)
        r2 = sum((x[..., d] - 0.2 * L) ** 2 for d in range(R))
        out[0, s:s + chunk] = np.exp(-r2 / (0.005 * L * L))
    return out


def initial_condition(ids, cfg):
    return acoustic_pulse(ids, cfg) if cfg.eq == B.EQ_EULER else gaussian_scalar(ids, cfg)


def build_static_tree(cfg, base_level, ball_radii=(), centre=(0.5, 0.5, 0.5)):
    """Host topology of a static multi-level mesh: uniform `base_level`, then one ring of
    refinement per entry of ball_radii (each pass refines leaves inside the ball by one level)."""
    t = B.HostTree(cfg.rank, cfg.depth)
    for _ in range(base_level):
        t.reconstruct(flags_refine_all(t.ids()))
    for r in ball_radii:
        t.reconstruct(flags_ball(t.ids(), cfg, r, cfg.depth, centre))
    return t


C2 = dict(cfg=("r2_s64_h1_d7_euler", 2, 64, 1, 7, B.EQ_EULER), base_level=5,
          ball_radii=(0.25, 0.125))


def c2_config():
    return Config(2, 64, 1, 7, B.EQ_EULER)


# BASELINE config C3 (benchmark/bench_fvm_solver_integration3D.b.cpp:31-64 patch shape: 8^3 Euler, halo 1)
# as the static multi-level tree of ~1e9 cells SURVEY 8d names: uniform level 6 (262 144 patches), every
# leaf within 0.42 L of the centre refined to level 7, every leaf within 0.27 L to level 8
# (2:1-balanced by construction) -> ~2.1e6 patches, ~1.07e9 cells, levels 6-8, morton_id<8,3>.
# base_level < 6 gives the geometrically similar mesh with 8x fewer cells per level dropped (the
# bounded samples of the CPU / reference-CUDA baseline legs).
C3 = dict(base_level=6, ball_radii=(0.42, 0.27), depth=8, patches=2037232, cells=1043062784)


def c3_config():
    return Config(3, 8, 1, C3["depth"], B.EQ_EULER)


def c3_script(base_level=6):
    d = C3["depth"]
    return "\n".join(["A\nX"] * base_level + ["B %g 99 %d 0.5 0.5 0.5" % (C3["ball_radii"][0], d), "X",
                                              "B %g 99 %d 0.5 0.5 0.5" % (C3["ball_radii"][1], d), "X"])


def device_initial_condition(torch, ids, cfg, device, chunk=32768):
    """The same cell-centre initial conditions evaluated ON THE DEVICE (torch elementwise ops), for meshes
    whose host image would not fit a benchmark's time or memory budget (1e9 cells): yields
    (first_patch, [nvar] tensors of shape [n, S..]) per chunk of patches.  Synthetic-input plumbing only."""
    R, S, L = cfg.rank, cfg.size, cfg.length
    coords, lvl = morton_decode(ids, R)
    span = float(1 << cfg.depth)
    k = (torch.arange(S, dtype=torch.float64, device=device) + 0.5)
    c0 = 0.5 * L if cfg.eq == B.EQ_EULER else 0.2 * L
    for s in range(0, len(ids), chunk):
        n = min(chunk, len(ids) - s)
        ext = torch.from_numpy((2.0 ** (cfg.depth - lvl[s:s + n])) / span).to(device)
        r2 = torch.zeros((n,) + (S,) * R, dtype=torch.float64, device=device)
        for d in range(R):                       # physical axis d <-> layout dim R-1-d
            org = torch.from_numpy(coords[s:s + n, d].astype(np.float64) / span * L).to(device)
            dx = ext * L / S
            shape = [n] + [1] * R
            kshape = [1] * (R + 1)
            kshape[1 + (R - 1 - d)] = S
            x = org.reshape(shape) + k.reshape(kshape) * dx.reshape(shape)
            r2 = r2 + (x - c0) ** 2
        if cfg.eq == B.EQ_EULER:
            g = torch.exp(-r2 / (0.01 * L * L))
            z = torch.zeros_like(g)
            yield s, [0.5 + 2.0 * g] + [z] * R + [(1.0 + 10.0 * g) / (cfg.gamma - 1.0)]
        else:
            yield s, [torch.exp(-r2 / (0.005 * L * L))]


def c2_script(base_level=5):
    """the same mesh in the oracle/ref_dump script language (for the CPU baseline legs);
    base_level < 5 gives the geometrically similar mesh with 4x fewer cells per level dropped"""
    return "\n".join(["A\nX"] * base_level + ["B 0.25 99 7 0.5 0.5", "X", "B 0.125 99 7 0.5 0.5", "X"])
