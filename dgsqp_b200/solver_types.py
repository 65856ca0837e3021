"""Solver parameter records.

API-compatible (same class names, field names and defaults) with the records the reference's
drivers construct: ``PIDParams`` (``DGSQP/solvers/solver_types.py:12-55``), ``DGSQPParams``
(``:92-127``) and ``DGSQPV2Params`` (``:130-174``).  They are generated from the compact tables
below; unknown fields are rejected exactly like the reference's ``PythonMsg`` base does.
"""
from dataclasses import field, make_dataclass

from .types import PythonMsg

_COMMON_BUILD = [("code_gen", bool, False), ("jit", bool, False), ("opt_flag", str, "O0"),
                 ("enable_jacobians", bool, True), ("solver_dir", str, None), ("so_name", str, None),
                 ("qp_interface", str, "casadi"), ("qp_solver", str, "osqp"),
                 ("hessian_approximation", str, "none")]
_COMMON_DEBUG = [("debug", bool, False), ("debug_plot", bool, False), ("pause_on_plot", bool, False),
                 ("local_pos", bool, False)]

_TABLES = {
    "PIDParams": [("dt", float, 0.1), ("Kp", float, 2.0), ("Ki", float, 0.0), ("Kd", float, 0.0),
                  ("int_e_max", float, 100), ("int_e_min", float, -100),
                  ("u_max", float, None), ("u_min", float, None), ("du_max", float, None), ("du_min", float, None),
                  ("u_ref", float, 0.0), ("x_ref", float, 0.0),
                  ("noise", bool, False), ("noise_max", float, 0.1), ("noise_min", float, -0.1),
                  ("periodic_disturbance", bool, False), ("disturbance_amplitude", float, 0.1),
                  ("disturbance_period", float, 1.0)],
    # v1 (DGSQP.py): watchdog non-monotone line search, stationarity-l1 merit
    "DGSQPParams": [("dt", float, 0.1), ("N", int, 10), ("beta", float, 0.25), ("tau", float, 0.5),
                    ("p_tol", float, 1e-3), ("d_tol", float, 1e-3), ("reg", float, 1e-3),
                    ("line_search_iters", int, 50), ("nonmono_ls", bool, False), ("sqp_iters", int, 50),
                    ("merit_function", str, "stat_l1"), ("verbose", bool, False), ("save_iter_data", bool, True),
                    ("solver_name", str, "DGSQP"), ("time_limit", float, None), ("conv_approx", bool, True)]
                   + _COMMON_BUILD + _COMMON_DEBUG,
    # v2 (DGSQP_v2.py): decaying regularisation, d-step / m-step non-monotone strategy
    "DGSQPV2Params": [("dt", float, 0.1), ("N", int, 10), ("beta", float, 0.25), ("tau", float, 0.5),
                      ("p_tol", float, 1e-4), ("d_tol", float, 1e-4), ("reg", float, 1e2), ("reg_decay", float, 0.95),
                      ("line_search_iters", int, 50), ("nms", bool, True), ("nms_frequency", int, 5),
                      ("nms_memory_size", int, 3), ("sqp_iters", int, 500), ("merit_function", str, "stat_l1"),
                      ("merit_parameter", float, None), ("merit_decrease", float, 0.01),
                      ("merit_decrease_condition", str, "armijo"), ("approximation_eval", str, "always"),
                      ("delta_decay", float, 0.95), ("verbose", bool, False), ("save_iter_data", bool, False),
                      ("save_qp_data", bool, False), ("time_limit", float, None), ("solver_name", str, "DGSQP")]
                     + _COMMON_BUILD + _COMMON_DEBUG + [("save_plot", bool, False), ("show_ts", bool, False)],
}


def _build(name):
    cls = make_dataclass(name, [(n, t, field(default=d)) for n, t, d in _TABLES[name]], bases=(PythonMsg,))
    cls.__module__ = __name__
    return cls


PIDParams = _build("PIDParams")
DGSQPParams = _build("DGSQPParams")
DGSQPV2Params = _build("DGSQPV2Params")
