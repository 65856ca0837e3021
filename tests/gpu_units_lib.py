"""ctypes loader of the stage-level test kernel (tests/gpu_units/units.cu).  TEST ONLY."""
import ctypes as C
import pathlib
import subprocess

import numpy as np

from dgsqp_b200.games import params_to_struct

HERE = pathlib.Path(__file__).resolve().parent / "gpu_units"
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
        "-Xcompiler", "-fPIC"]


def build():
    out = HERE / "libdgsqp_units.so"
    srcs = [HERE / "units.cu"] + sorted((HERE.parents[1] / "dgsqp_b200" / "csrc").glob("*.cuh")) \
        + sorted((HERE.parents[1] / "dgsqp_b200" / "csrc").glob("*.h"))
    if out.exists() and all(out.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return out
    subprocess.check_call([*NVCC, "-o", str(out), str(HERE / "units.cu")])
    return out


def run_stages(game, params, x0, u, l, threads=128, smem_limit_doubles=0):
    """evaluate -> nearestPD -> QP at (u, l), and the LSQR dual initialisation at u, for a batch."""
    import os
    lib = C.CDLL(os.environ.get("DG_UNITS_LIB") or str(build()))
    B, n, m = x0.shape[0], game.n, game.m
    x0, u, l = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, u, l))
    out = dict(Q=np.zeros((B, n, n)), H=np.zeros((B, n, n)), q=np.zeros((B, n)), gtl=np.zeros((B, n)),
               du=np.zeros((B, n)), g=np.zeros((B, m)), lam=np.zeros((B, m)), l0=np.zeros((B, m)),
               nneg=np.zeros(B, np.int32), qpst=np.zeros(B, np.int32), qpit=np.zeros(B, np.int32),
               lsqr_it=np.zeros(B, np.int32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    gs, ps = game.to_struct(), params_to_struct(params)
    rc = lib.units_run(C.byref(gs), C.byref(ps), B, threads, C.c_long(smem_limit_doubles), p(x0), p(u), p(l), p(out["Q"]), p(out["q"]),
                       p(out["gtl"]), p(out["g"]), p(out["H"]), p(out["du"]), p(out["lam"]), p(out["l0"]),
                       p(out["nneg"]), p(out["qpst"]), p(out["qpit"]), p(out["lsqr_it"]))
    assert rc == 0, f"units_run failed: {rc}"
    return out
