"""Checkpoint hand-off for the hot path.

What a restart needs from the force-evaluation path is small: the vertex positions, the CURRENT
face definitions (flip sweeps change them on the device; `TriMesh.fv_indices()` reads them back) and
the continuation state of the `EnergyManager` (`eparams.continuation_params.lam / .delta`). This
module mirrors the reference's interfaces for exactly that and nothing else of its run orchestration:

  * `Mesh`                 src/trimem/mc/mesh.py:13-60   (`x`, `f`, `copy`, on trimem_b200.core.TriMesh)
  * `CheckpointWriter`     src/trimem/mc/output/checkpoint.py:23-100  (`write(points, cells, config)`)
  * `CheckpointReader`     src/trimem/mc/output/checkpoint.py:103-172 (`read() -> points, cells, config`)
  * `write_checkpoint_handle` / `read_checkpoint`   src/trimem/mc/util.py:50-132

File layout as in the reference: `<prefix>.p<N>.cpt` is the XML index (points/cells shape, dtype and
`file:/dataset` locations, the configuration as a JSON string), enumerated per write; the arrays live
in `<prefix>.cpt.p<N>.h5` (gzip level 4) when `h5py` is importable -- files of the two
implementations are then interchangeable -- and otherwise in `<prefix>.cpt.p<N>.npz`.

Restart (src/trimem/mc/util.py:153-160): the EnergyManager is built on the INITIAL mesh (its
reference properties come from there), the continuation parameters come from the checkpoint, the mesh
comes from the checkpoint, and `estore.update_repulsion(mesh.trimesh)` rebuilds the neighbour list on it
before the first evaluation (a list built on another TriMesh object is refused by the device layer).
"""
import configparser
import copy
import io
import json
import pathlib
import re
from xml.etree import ElementTree as ET

import numpy as np

try:                                    # optional, exactly the reference's container when present
    import h5py
except ImportError:                     # pragma: no cover - depends on the image
    h5py = None


class Mesh:
    """Positions and faces of a `core.TriMesh` with the reference wrapper's accessors."""

    def __init__(self, points=None, cells=None, core=None):
        if core is None:
            from . import core
        self._core = core
        if points is not None and cells is not None:
            self.trimesh = core.TriMesh(np.ascontiguousarray(points, dtype=np.float64),
                                        np.ascontiguousarray(cells, dtype=np.int32))
        else:
            self.trimesh = core.TriMesh()

    @property
    def x(self):
        return self.trimesh.points()

    @x.setter
    def x(self, values):
        np.copyto(self.x, values)

    @property
    def f(self):
        return self.trimesh.fv_indices()

    def copy(self):
        return Mesh(self.x.copy(), self.f.copy(), core=self._core)


def _pattern(fname):
    return f"{fname.stem}.p{{}}{fname.suffix}"


def _get_parts(fname):
    """Number of existing parts `<stem>.p<N><suffix>` beside fname (output/_common.py:15-35)."""
    rx = re.compile(re.escape(_pattern(fname)).replace(r"\{\}", "[0-9]+") + "$")
    parent = fname.parent if str(fname.parent) else pathlib.Path(".")
    return sum(1 for f in parent.iterdir() if rx.match(f.name))


def _create_part(fname):
    return fname.with_name(_pattern(fname).format(_get_parts(fname)))


class CheckpointWriter:
    def __init__(self, fname):
        self.fname = pathlib.Path(fname).with_suffix(".cpt")
        self.fdata = self.fname.with_suffix(".cpt.h5" if h5py is not None else ".cpt.npz")
        self.fname = _create_part(self.fname)          # <prefix>.p<N>.cpt
        self.fdata = _create_part(self.fdata)          # <prefix>.cpt.p<N>.h5 / .npz
        self.cpt = ET.Element("cpt", Version="0.1")

    def _write_arrays(self, points, cells):
        if h5py is not None:
            with h5py.File(self.fdata, "w") as h5:
                h5.create_dataset("points", data=points, compression="gzip", compression_opts=4)
                h5.create_dataset("cells", data=cells, compression="gzip", compression_opts=4)
        else:
            np.savez_compressed(self.fdata, points=points, cells=cells)

    def write(self, points, cells, config, **kwargs):
        points = np.asarray(points)
        cells = np.asarray(cells)
        for tag, a in (("points", points), ("cells", cells)):
            el = ET.SubElement(self.cpt, tag, shape="{} {}".format(*a.shape), dtype=a.dtype.name)
            el.text = self.fdata.name + ":/" + tag
        self._write_arrays(points, cells)
        with io.StringIO() as fp:
            config.write(fp)
            config_str = fp.getvalue()
        ET.SubElement(self.cpt, "config").text = json.dumps(config_str)
        ET.ElementTree(self.cpt).write(self.fname)


class CheckpointReader:
    def __init__(self, fname, fnum):
        if fnum == -1:
            fnum = _get_parts(pathlib.Path(fname).with_suffix(".cpt")) - 1
        self.fname = pathlib.Path(fname).with_suffix(f".p{fnum}.cpt")
        self.root = ET.parse(self.fname).getroot()

    def _array(self, el, tag):
        if el.tag != tag:
            raise ValueError("Checkpoint file corrupt.")
        shape = tuple(int(i) for i in el.attrib["shape"].split())
        dtype = el.attrib["dtype"]
        fname, dataset = el.text.split(":/")[0], el.text.split(":/")[-1]
        path = self.fname.with_name(fname)
        out = np.zeros(shape, dtype=dtype)
        if path.suffix == ".h5":
            if h5py is None:
                raise RuntimeError("this checkpoint stores its arrays in HDF5 and h5py is not installed")
            with h5py.File(path, "r") as h5:
                h5[dataset].read_direct(out)
        else:
            with np.load(path, allow_pickle=False) as z:
                a = z[dataset]
            if a.shape != shape or a.dtype != out.dtype:
                raise ValueError("Checkpoint file corrupt.")
            out[...] = a
        return out

    def read(self):
        points = self._array(self.root[0], "points")
        cells = self._array(self.root[1], "cells")
        if self.root[2].tag != "config":
            raise ValueError("Checkpoint file corrupt.")
        config = configparser.ConfigParser()
        config.read_string(json.loads(self.root[2].text))
        return points, cells, config


def write_checkpoint_handle(config):
    """(mesh, estore, step) -> writes `<restart_prefix>.p<N>.cpt` (src/trimem/mc/util.py:50-93)."""
    conf = copy.deepcopy(config)

    def _write_checkpoint(mesh, estore, step={}):
        if not isinstance(step, dict):
            raise TypeError("kwarg 'step' must be dict-like")
        cp = estore.eparams.continuation_params
        conf.read_dict({"ENERGY": {"continuation_lambda": repr(cp.lam), "continuation_delta": repr(cp.delta)},
                        "HMC": {"init_step": json.dumps(step)}})
        cpt = CheckpointWriter(config["GENERAL"]["restart_prefix"])
        cpt.write(mesh.x, mesh.f, conf)
        return cpt.fname

    return _write_checkpoint


def read_checkpoint(config, restartnum, core=None):
    """(Mesh, config) with the continuation parameters and `init_step` of checkpoint `restartnum`
    (src/trimem/mc/util.py:95-132)."""
    cpt = CheckpointReader(config["GENERAL"]["restart_prefix"], restartnum)
    points, cells, conf = cpt.read()
    config.read_dict({"ENERGY": {"continuation_delta": conf["ENERGY"]["continuation_delta"],
                                 "continuation_lambda": conf["ENERGY"]["continuation_lambda"]},
                      "DEFAULT": {"init_step": conf["HMC"]["init_step"]}})
    return Mesh(points, cells, core=core), config
