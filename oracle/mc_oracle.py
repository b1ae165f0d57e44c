"""ctypes loader for the CPU marching-cubes checker (oracle/mc_oracle.c).  TEST INFRASTRUCTURE ONLY.

``marching_cubes_lewiner(volume, level)`` mirrors the call shape of the function the
reference uses (lib/mesh_util.py:40): returns (verts [V,3] f32 in (axis0,axis1,axis2)
index coordinates, faces [F,3] i32, normals [V,3] f32, values [V] f32) and raises the
same errors as skimage 0.17.2 (ValueError when level is outside the data range,
RuntimeError when no surface is found).  PARITY UNPINNED vs skimage -- see mc_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmc_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "libmc_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        P = ctypes.c_void_p
        lib.mc_oracle_run.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      P, P, P, P, P, P, P]
        lib.mc_oracle_run.restype = ctypes.c_int
        lib.mc_oracle_interior_stats.argtypes = [P, P]
        lib.mc_oracle_interior_stats.restype = None
        _LIB = lib
    return _LIB


def marching_cubes_lewiner(volume, level, return_stats=False):
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    if vol.ndim != 3 or min(vol.shape) < 2:
        raise ValueError("Input volume should be a 3D numpy array with at least 2 nodes per axis.")
    if level < vol.min() or level > vol.max():
        raise ValueError("Surface level must be within volume data range.")
    lib = _lib()
    nv, nf, na = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    R0, R1, R2 = vol.shape
    rc = lib.mc_oracle_run(vol.ctypes.data, R0, R1, R2, float(level), None, None, None, None,
                           ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(na))
    if rc != 0:
        raise MemoryError("mc_oracle_run")
    if nv.value == 0 or nf.value == 0:
        raise RuntimeError("No surface found at the given iso value.")
    verts = np.empty((nv.value, 3), np.float32)
    faces = np.empty((nf.value, 3), np.int32)
    normals = np.empty((nv.value, 3), np.float32)
    values = np.empty((nv.value,), np.float32)
    lib.mc_oracle_run(vol.ctypes.data, R0, R1, R2, float(level), verts.ctypes.data, faces.ctypes.data,
                      normals.ctypes.data, values.ctypes.data,
                      ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(na))
    if return_stats:
        ni, nt = ctypes.c_int64(), ctypes.c_int64()
        lib.mc_oracle_interior_stats(ctypes.byref(ni), ctypes.byref(nt))
        return verts, faces, normals, values, {"ambiguous_cells": int(na.value), "interior_ambiguous_cells": int(ni.value),
                                               "tunnel_cells": int(nt.value)}
    return verts, faces, normals, values
