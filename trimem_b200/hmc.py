"""Hamiltonian Monte Carlo with the trajectory on the device.

Host-side mirror of the reference's ``trimem.mc.hmc.HMC`` / ``MeshHMC`` (src/trimem/mc/hmc.py:48-231)
for the caller row next to the hot path: same options, same ``step`` / ``run`` /
``info`` protocol, same use of numpy's global random state (one ``normal`` draw of the momenta, one
``uniform`` draw for the Metropolis test, in that order), so a chain started from the same seed visits
the same states. What changes is ``_step``: the reference integrates in Python, crossing the pybind
boundary for every force evaluation (hmc.py:14-26,133-155); here ``core.vv_integrate`` runs the
``num_integration_steps`` velocity-Verlet steps and both Hamiltonian evaluations on the GPU and
returns the end point, its momenta and the four energies.

The reference's per-move callback contract is kept: ``callback(x, counter)`` runs after every step and
is where ``estore.update_repulsion`` / ``update_reference_properties`` belong (mc/evaluators.py:163-183).
"""
from collections import Counter

import numpy as np

from . import core as m

_hmc_default_options = {
    "mass":                  1.0,
    "time_step":             1.0e-4,
    "num_integration_steps": 100,
    "initial_temperature":   1.0,
    "minimal_temperature":   1.0e-6,
    "cooling_factor":        0.0,
    "cooling_start_step":    0,
    "info_step":             100,
}


def get_step_counters():
    """Counter with keys for vertex moves and edge flips (hmc.py:28-35)."""
    return Counter(move=0, flip=0)


class MeshHMC:
    """HMC with a mesh as state and the trajectory resident on the GPU.

    Args:
        mesh: an object with ``.trimesh`` (``core.TriMesh``) and a settable ``.x`` — the reference's
            ``trimem.mc.mesh.Mesh`` — or a bare ``core.TriMesh``.
        estore (core.EnergyManager): defines nlog_prob = ``estore.energy`` and its gradient.

    Keyword Args:
        callback: ``callback(x, counter)`` after every step (default: no-op).
        counter: step counters shared with a flip proposal (default: fresh counters).
        options: see ``trimem.mc.hmc.HMC``.
        device_momenta: ``None`` (default: momenta from ``rng``, like the reference) or an integer seed: momenta drawn on the
            device (extension, see below).
        rng: ``numpy.random.RandomState`` or ``numpy.random.Generator`` of this chain (default: numpy's global state, as
            in the reference; a ``Generator`` draws the momenta outside the interpreter lock, which is what lets
            several chains driven from threads of one process overlap their host work).
    """

    def __init__(self, mesh, estore, callback=None, counter=None, options={}, verbose=True, rng=None, device_momenta=None):
        self.trimesh = getattr(mesh, "trimesh", mesh)
        self.mesh = mesh
        self.estore = estore
        self.cb = (lambda x, s: None) if callback is None else callback
        options = {**_hmc_default_options, **options}
        self.m = options["mass"]
        self.dt = options["time_step"]
        self.L = options["num_integration_steps"]
        self.Tinit = options["initial_temperature"]
        self.Tmin = options["minimal_temperature"]
        self.fT = options["cooling_factor"]
        self.cN = options["cooling_start_step"]
        self.istep = options["info_step"]
        if verbose:
            print("\n---------------------------------------")
            print("Hamiltonian Monte Carlo Initialization:")
            width = max(len(str(k)) for k in options)
            for k, v in options.items():
                print(f"  {k: <{width}}: {v}")
        self.counter = get_step_counters() if counter is None else counter
        self.i = 0
        self.acc = 0
        self.T = self.Tinit
        self.x = self.trimesh.points().copy()
        self.last = None            # (dh, accepted, [U0, K0, U1, K1]) of the last step
        # rebuild trigger (extension): largest vertex displacement of the CURRENT state from the
        # positions the verlet list was built from, as measured on the device behind each trajectory (None: not
        # measured -- cell list or no repulsion), and rlist - lc1. See LazyRepulsion.
        self.displacement = 0.0
        self.margin = None
        # source of the momenta and the Metropolis variate: numpy's global state like the reference, or the chain's
        # own numpy.random.RandomState (independent chains driven from several threads, ranks of a partitioned mesh)
        self.rng = np.random if rng is None else rng
        # extension: an integer seed -> the momenta of move k are drawn ON THE DEVICE from (seed, k) (counter-based
        # generator, core.vv_integrate_random) instead of by numpy on the host, where the draw is two thirds of a move at
        # 40,962 vertices. Such a chain is reproducible from its seed but is not the reference's chain for any numpy seed.
        self.device_momenta = device_momenta
        self.n_draws = 0

    def _set_mesh(self, x):
        # (np.copyto(self.trimesh.points(), x) in the reference, mesh.py:39-42; set_points copies AND uploads once)
        self.trimesh.set_points(x)

    def _step(self):
        """Metropolis step (hmc.py:133-155)."""
        p_var = self.m * self.T
        self._set_mesh(self.x)
        if self.device_momenta is None:
            p = self.rng.normal(size=self.x.shape) * np.sqrt(p_var)
            xn, pn, st = m.vv_integrate(self.trimesh, self.estore, p, self.m, self.dt, self.L)
        else:
            xn, st = m.vv_integrate_random(self.trimesh, self.estore, int(self.device_momenta), self.n_draws,
                                           float(np.sqrt(p_var)), self.m, self.dt, self.L)
            self.n_draws += 1
        # (nlog_prob(xn) + K(pn)) - (nlog_prob(x) + K(p)), in the reference's order of operations
        dh = ((st[2] + st[3]) - (st[0] + st[1])) / self.T
        a = 1.0 if dh <= 0 else np.exp(-dh)
        u = self.rng.uniform()
        acc = u <= a
        try:
            d_end, self.margin = m.vv_displacement(self.trimesh, self.estore)
        except RuntimeError:
            d_end, self.margin = None, None
        if acc:
            self.x = xn
            self.acc += 1
            self.displacement = d_end       # a rejected move returns to a state whose displacement is known
        self.i += 1
        self.last = (dh, bool(acc), st)

    def info(self):
        i_total = sum(self.counter.values())
        if self.istep and i_total % self.istep == 0:
            ar = self.acc / self.i if not self.i == 0 else 0.0
            print("\n-- HMC-Step ", self.counter["move"])
            print("----- acc-rate:   ", ar)
            print("----- temperature:", self.T)
            self.acc = 0
            self.i = 0

    def step(self):
        """One step; the mesh vertices are updated explicitly afterwards (MeshHMC.step, hmc.py:230-233)."""
        i = sum(self.counter.values())
        Tn = np.exp(-self.fT * (i - self.cN)) * self.Tinit
        self.T = max(min(Tn, self.Tinit), self.Tmin)
        self._step()
        self.counter["move"] += 1
        self._set_mesh(self.x)

    def run(self, N):
        for _ in range(N):
            self.step()
            self.info()
            self.cb(self.x, self.counter)


class LazyRepulsion:
    """Per-move callback that rebuilds the verlet list only when the device-side trigger says so.

    The reference rebuilds the list in its per-move callback every ``refresh`` moves whatever has happened
    (mc/evaluators.py:181-182). A verlet list built with ``rlist > lc1`` stays COMPLETE -- no pair inside the
    repulsion cut-off is missing, so energies and gradients are those of a fresh list -- until some vertex has moved
    more than ``(rlist - lc1) / 2`` from where it was at the build. ``MeshHMC`` tracks that displacement for the
    chain's current state (measured on the device behind every trajectory, no extra synchronisation); this callback
    rebuilds when it exceeds ``fraction`` of the bound (default one half: the other half is left to the motion inside
    the next trajectory), after accepted flips (a verlet list freezes its exclusions), and always
    for lists that keep no build positions (cell lists: those are cheap to rebuild).

        hmc = MeshHMC(mesh, estore, ...)
        upd = LazyRepulsion(hmc)
        hmc.cb = upd                              # or call upd(x, counter) from an existing callback
        ...; upd.flips_accepted(core.pflip(...))  # after a flip sweep
    """

    def __init__(self, hmc, fraction=0.5):
        self.hmc = hmc
        self.fraction = fraction
        self.force = False
        self.rebuilds = 0
        self.skipped = 0

    def flips_accepted(self, n):
        self.force = self.force or n > 0
        return n

    def needed(self):
        h = self.hmc
        if self.force or h.margin is None or h.displacement is None:
            return True
        return h.displacement > self.fraction * 0.5 * h.margin

    def __call__(self, x=None, counter=None):
        if self.needed():
            self.hmc.estore.update_repulsion(self.hmc.trimesh)
            self.hmc.displacement = 0.0
            self.force = False
            self.rebuilds += 1
        else:
            self.skipped += 1
