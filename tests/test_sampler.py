"""Warm-start generation: the PID lane-follower roll-out of the Monte-Carlo drivers
(scripts/DGSQP_ALGAMES_monte_carlo_chicane.py:411-467) as device code (csrc/pid_rollout.cuh) against the vectorised
NumPy statement of the same roll-out (dgsqp_b200/montecarlo.py: pid_rollout).  CPU: the device source compiled for the
host; GPU: the kernel through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import dgsqp_b200 as dg
from dgsqp_b200.montecarlo import pid_rollout, pid_rollout_device, sample_head_to_head


def _agents(K, seed=1):
    rng = np.random.default_rng(seed)
    s0, xt0, v0 = np.maximum(0.1, rng.random(K) * 3.0), rng.random(K) * 2 - 1, rng.random(K) + 2
    s0[:5] = [14.5, 13.9, 0.1, 5.0, 9.99]          # close to the wrap-around and to segment boundaries
    return s0, xt0, v0


@pytest.mark.parametrize("mk", [lambda: dg.chicane_game(), lambda: dg.curve_game(75.0, 20), lambda: dg.agents_game(3, 90.0, 25)])
def test_pid_rollout_device_source_on_host(mk):
    import hostsim_lib
    lib = C.CDLL(str(hostsim_lib.build(False, False)))
    game = mk()
    K = 300
    s0, xt0, v0 = _agents(K)
    q0, xy, u = pid_rollout(game, s0, xt0, v0)
    q0h, xyh, uh = np.empty((K, 6)), np.empty((K, game.N + 1, 2)), np.empty((K, game.N, 2))
    kp, gs = np.ascontiguousarray(game.track.key_pts), game.to_struct()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.hs_pid_rollout(C.byref(gs), p(kp), K, p(s0), p(xt0), p(v0), p(q0h), p(xyh), p(uh)) == 0
    assert np.abs(q0 - q0h).max() < 1e-13 and np.abs(xy - xyh).max() < 1e-12 and np.abs(u - uh).max() < 1e-12
    # the roll-out respects the input box and rate limits of the controllers (PID.py:104-113)
    assert np.all(uh[:, :, 0] <= game.u_ub[0] + 1e-15) and np.all(uh[:, :, 1] >= game.u_lb[1] - 1e-15)


@pytest.mark.gpu
def test_pid_rollout_on_device():
    game = dg.chicane_game()
    K = 5000
    s0, xt0, v0 = _agents(K, seed=3)
    q0, xy, u = pid_rollout(game, s0, xt0, v0)
    q0d, xyd, ud = pid_rollout_device(game, s0, xt0, v0)
    assert np.abs(q0 - q0d).max() < 1e-12 and np.abs(xy - xyd).max() < 1e-10 and np.abs(u - ud).max() < 1e-10
    # the sampler built on it: same accepted instances as the host sampler (a borderline collision check may flip)
    xh, uh = sample_head_to_head(game, 512, seed=5)
    xd, udv = sample_head_to_head(game, 512, seed=5, device=0)
    assert xd.shape == xh.shape and udv.shape == uh.shape
    same = np.all(np.abs(xd - xh) < 1e-9, axis=1)
    assert same.mean() > 0.98 and np.abs(udv[same] - uh[same]).max() < 1e-9
    # and the solver accepts them
    res = dg.DGSQP(game, dg.chicane_params(), print_method=None, mu_vio_thresh=1e-10).solve_batch(xd[:64], udv[:64])
    assert set(np.unique(res.status)) <= {0, 1, 2, 3, 4}


def test_pid_rollout_against_the_oracle_rk45_sampler():
    """The product roll-out (fixed-step RK4 x 4 per stage, the form the device kernel implements) against the ORACLE's
    faithful restatement of the script's sampler (oracle/sampler.py: SciPy RK45 at its default rtol = 1e-3 / atol = 1e-6,
    dynamics_models.py:161-186).  Same initial state bit for bit; the warm-start trajectories and inputs differ by the
    reference integrator's own tolerance (measured on 40 starts: 1.1e-2 in position, 2.3e-2 in the inputs) -- they are
    solver INPUTS, the parity fixtures draw theirs from the oracle sampler."""
    from oracle import sampler as osamp
    from oracle.racing_game import RacingGame
    from oracle.track import chicane_track
    og, game = RacingGame(chicane_track(), M=2, N=25), dg.chicane_game()
    rng = np.random.default_rng(1)
    K = 24
    s0, xt0, v0 = np.maximum(0.1, rng.random(K) * 3.0), rng.random(K) * 2 - 1, rng.random(K) + 2
    q0, xy, u = pid_rollout(game, s0, xt0, v0)
    for i in range(K):
        q0o, qws, uo = osamp.pid_rollout(og, s0[i], xt0[i], v0[i])
        assert np.array_equal(q0o, q0[i])
        assert np.abs(qws[:, :2] - xy[i]).max() < 3e-2 and np.abs(uo - u[i]).max() < 6e-2
