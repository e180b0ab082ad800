"""The synthetic workloads of BASELINE.json / SURVEY.md section 8(d).

The functions take the module to parametrise as an argument (trimem_b200.core, or any module with
the trimem.core API such as the reference's own build): same API, which is the point of the drop-in
boundary.
"""
import numpy as np

from . import meshgen

N_VERTICES = {4: 2562, 6: 40962, 8: 655362, 10: 10485762}

# algorithmic bytes per vertex; P = mean directed list entries/vertex
def bytes_energy(P=0.0, rep=False):
    return 52.0 + ((4.0 * P + 4.0) if rep else 0.0)


def bytes_grad_props(P=0.0, rep=False):
    return 60.0 + ((4.0 * P + 4.0) if rep else 0.0)


def bytes_grad_sweep(P=0.0, rep=False):
    return 84.0 + ((4.0 * P + 4.0) if rep else 0.0)


def bytes_eg(P=0.0, rep=False):
    return bytes_energy(P, rep) + bytes_grad_props(P, rep) + bytes_grad_sweep(P, rep)


def energy_params(mod, mesh, repulsion="none", n_search="verlet-list", exclusion=2, bond="Edge",
                  kappa=None, fracs=(1.0, 0.9, 1.0)):
    """EnergyParams of the benchmark configurations, relative to the mesh's mean edge length.

    repulsion: "none"     kappa_r = 0, lc1 = 0, rlist = 4/3 l   (config 1)
               "physical" kappa_r = 1e3, lc1 = 1.0 l, rlist = 4/3 l
               "stress"   kappa_r = 1e3, lc1 = 3.0 l, rlist = 3.2 l  (configs 2-4)
    """
    a, l = mod.avg_tri_props(mesh)
    bp = mod.BondParams()
    bp.type = getattr(mod.BondType, bond)
    bp.r = 2
    bp.lc0 = 1.15 * l
    bp.lc1 = 0.85 * l
    bp.a0 = a
    ep = mod.EnergyParams()
    kr = 0.0 if repulsion == "none" else 1.0e3
    kb, ka, kv, kc, kt, kr = kappa or (30.0, 1.0e6, 1.0e6, 0.0, 1.0e5, kr)
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = kb, ka, kv, kc, kt, kr
    ep.area_frac, ep.volume_frac, ep.curvature_frac = fracs
    ep.bond_params = bp
    rp = mod.SurfaceRepulsionParams()
    rp.n_search = n_search
    rp.exclusion_level = exclusion
    rp.r = 2
    if repulsion == "stress":
        rp.lc1, rp.rlist = 3.0 * l, 3.2 * l
    elif repulsion == "physical":
        rp.lc1, rp.rlist = 1.0 * l, 4.0 / 3.0 * l
    else:
        rp.lc1, rp.rlist = 0.0, 4.0 / 3.0 * l
    ep.repulse_params = rp
    return ep


def make_case(mod, subdiv, repulsion="none", n_search="verlet-list", seed=42, **kw):
    """(mesh, estore, x) for a jittered icosphere; the EnergyManager is built after the jitter."""
    v, f = meshgen.jittered_icosphere(subdiv, seed=seed)
    mesh = mod.TriMesh(v, f)
    ep = energy_params(mod, mesh, repulsion=repulsion, n_search=n_search, **kw)
    estore = mod.EnergyManager(mesh, ep)
    return mesh, estore, v


def abi_params(estore):
    """tm_energy_params for the current state of a trimem_b200.core.EnergyManager (ctypes)."""
    from .cabi import EnergyParamsABI, BOND_EDGE, BOND_AREA
    ep = estore.eparams
    lam = ep.continuation_params.lam
    ip = estore.initial_props
    p = EnergyParamsABI()
    p.kappa_b, p.kappa_a, p.kappa_v = ep.kappa_b, ep.kappa_a, ep.kappa_v
    p.kappa_c, p.kappa_t, p.kappa_r = ep.kappa_c, ep.kappa_t, ep.kappa_r
    p.ref_area = (1.0 - lam * (1.0 - ep.area_frac)) * ip.area
    p.ref_volume = (1.0 - lam * (1.0 - ep.volume_frac)) * ip.volume
    p.ref_curvature = (1.0 - lam * (1.0 - ep.curvature_frac)) * ip.curvature
    p.bond_lc0, p.bond_lc1, p.bond_a0 = ep.bond_params.lc0, ep.bond_params.lc1, ep.bond_params.a0
    p.bond_r = ep.bond_params.r
    p.bond_type = BOND_EDGE if ep.bond_params.type == type(ep.bond_params.type).Edge else BOND_AREA
    p.rep_lc1, p.rep_r = ep.repulse_params.lc1, ep.repulse_params.r
    return p
