import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def chicane_small():
    """Oracle game + matching product game at a short horizon (fast on CPU)."""
    import dgsqp_b200 as dg
    from oracle.track import chicane_track
    from oracle.racing_game import RacingGame
    N = 8
    return RacingGame(chicane_track(), M=2, N=N), dg.chicane_game(N=N), dg.chicane_params(N=N)


@pytest.fixture(scope="session")
def chicane_full():
    import dgsqp_b200 as dg
    from oracle.track import chicane_track
    from oracle.racing_game import RacingGame
    return RacingGame(chicane_track(), M=2, N=25), dg.chicane_game(), dg.chicane_params()
