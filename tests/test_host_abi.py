"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol the header declares,
fails loudly without a device, and the Python mirror of the reference surface behaves like the reference."""
import ctypes as C
import pathlib
import re

import numpy as np
import pytest

import dgsqp_b200 as dg
from dgsqp_b200 import _abi
from dgsqp_b200.games import params_to_struct

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _lib():
    if not _abi.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    return _abi.load()


def test_header_symbols_exported():
    header = (ROOT / "include" / "dgsqp_b200.h").read_text()
    declared = set(re.findall(r"\b(dgsqp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS)
    lib = _lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dgsqp_version()


def test_struct_layout_matches_header():
    # field order/types are mirrored by hand in _abi.py; sizes must match the C compiler's layout
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "dgsqp_b200.h"
        int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(dgsqp_racing_game), sizeof(dgsqp_params), sizeof(dgsqp_v2_params), sizeof(dgsqp_merge_game), sizeof(dgsqp_lane_row)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        p = pathlib.Path(d)
        (p / "t.c").write_text(src)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), "-o", str(p / "t"), str(p / "t.c")])
        a, b, c2, d2, e2 = map(int, subprocess.check_output([str(p / "t")]).split())
    assert (a, b, c2) == (C.sizeof(_abi.RacingGameStruct), C.sizeof(_abi.ParamsStruct), C.sizeof(_abi.ParamsV2Struct))
    assert (d2, e2) == (C.sizeof(_abi.MergeGameStruct), C.sizeof(_abi.LaneRowStruct))


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib()
    gs, ps = dg.chicane_game().to_struct(), params_to_struct(dg.chicane_params())
    h = C.c_void_p()
    rc = lib.dgsqp_create(C.byref(gs), C.byref(ps), 0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.dgsqp_last_error()
    with pytest.raises(_abi.DgsqpLibraryError):
        dg.DGSQP(dg.chicane_game(), dg.chicane_params(), print_method=None, mu_vio_thresh=1e-10)


def test_invalid_arguments_rejected():
    lib = _lib()
    g = dg.chicane_game().to_struct()
    g.M = 7
    ps = params_to_struct(dg.chicane_params())
    h = C.c_void_p()
    assert lib.dgsqp_create(C.byref(g), C.byref(ps), 0, C.byref(h)) == -1
    assert b"invalid" in lib.dgsqp_last_error()
    assert lib.dgsqp_solve_batch(None, 1, None, None, None, None, None, None, None, None, None, None, None, 0, None) == -1


def test_params_and_types_surface():
    p = dg.DGSQPParams(N=25, reg=1e-3, nonmono_ls=True)
    assert (p.sqp_iters, p.line_search_iters, p.merit_function, p.qp_solver, p.conv_approx) == (50, 50, "stat_l1", "osqp", True)
    v2 = dg.DGSQPV2Params()
    assert (v2.reg, v2.reg_decay, v2.nms_frequency, v2.nms_memory_size, v2.sqp_iters, v2.p_tol) == (1e2, 0.95, 5, 3, 500, 1e-4)
    with pytest.raises(TypeError):
        p.not_a_field = 1
    with pytest.raises(ValueError):
        params_to_struct(dg.DGSQPParams(merit_function="bogus"))
    s = dg.VehicleState(t=0.0)
    s.p.s, s.p.x_tran, s.v.v_long = 0.5, 0.2, 2.5
    g = dg.chicane_game()
    g.track.local_to_global_typed(s)
    q = g.state2q([s, s])
    assert q.shape == (12,) and np.allclose(q[:6], [s.x.x, s.x.y, 2.5, 0.0, 0.5, 0.2])


def test_game_validation():
    with pytest.raises(ValueError):
        dg.RacingGame(track=dg.chicane_game().track, M=2, obs_r=[0.4])
    with pytest.raises(ValueError):
        dg.RacingGame(track=dg.chicane_game().track, M=5, obs_r=[0.4] * 5)


def test_warm_start_layout_roundtrip():
    """set_warm_start's stage-major -> agent-major reshuffle (DGSQP.py:271-281) and its inverse (:477-482),
    exercised without constructing a GPU handle."""
    from dgsqp_b200.solver import DGSQP
    obj = DGSQP.__new__(DGSQP)
    obj.N, obj.M, obj.n_u = 5, 3, 6
    u = np.arange(2 * 5 * 6, dtype=float).reshape(2, 5, 6)
    am = obj.stage_to_agent_major(u)
    ref = np.concatenate([u[0][:, 2 * a:2 * a + 2].ravel() for a in range(3)])
    assert np.array_equal(am[0], ref)
    assert np.array_equal(obj.agent_to_stage_major(am), u)
    with pytest.raises(RuntimeError, match="incompatible with required shape"):
        obj.set_warm_start(np.zeros((4, 6)))


def test_montecarlo_sampler_properties():
    from dgsqp_b200.montecarlo import sample_head_to_head, sample_agents
    g = dg.chicane_game()
    x0, u = sample_head_to_head(g, 300, seed=1)
    assert x0.shape == (300, 12) and u.shape == (300, 100)
    assert np.all(np.abs(x0[:, [5, 11]]) <= 1.0) and np.all(x0[:, [4, 10]] >= 0) and np.all((x0[:, [2, 8]] >= 2) & (x0[:, [2, 8]] <= 3))
    d = np.linalg.norm(x0[:, 0:2] - x0[:, 6:8], axis=1)
    assert np.all(d >= 0.8 - 1e-9)                               # not colliding at k = 0
    ua = u.reshape(300, 2, 25, 2)
    assert np.all(np.abs(ua[..., 0]) <= 2.1 + 1e-12) and np.all(np.abs(ua[..., 1]) <= 0.436 + 1e-12)
    x0b, ub = sample_head_to_head(g, 300, seed=1)
    assert np.array_equal(x0, x0b) and np.array_equal(u, ub)     # seeded
    g3 = dg.agents_game(M=3, N=10)
    x3, u3 = sample_agents(g3, 50, seed=0)
    assert x3.shape == (50, 18) and u3.shape == (50, 60)
