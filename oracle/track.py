"""Constant-curvature-segment track.  Oracle-only restatement.

Follows ``DGSQP/tracks/radius_arclength_track.py`` (key points ``:361-408``,
curvature ``:199-205``, tangent angle ``:207-225``, ``local_to_global``
``:752-807``) and ``DGSQP/tracks/track_lib.py`` (``CurveTrack:27-52``,
``ChicaneTrack:54-87``).
"""
import math
import numpy as np


def wrap_angle(theta):
    # radius_arclength_track.py:809-817
    if theta < -np.pi:
        return 2 * np.pi + theta
    elif theta > np.pi:
        return theta - 2 * np.pi
    return theta


def _sign(a):
    # radius_arclength_track.py:820-826 (sign(0) = +1)
    return 1 if a >= 0 else -1


class RadiusArclengthTrack:
    def __init__(self, track_width, slack, cl_segs):
        self.track_width = track_width
        self.slack = slack
        self.cl_segs = np.asarray(cl_segs, dtype=np.float64)
        self.half_width = track_width / 2
        self.key_pts = self._key_pts(self.cl_segs, (0.0, 0.0, 0.0))
        self.track_length = self.key_pts[-1, 3]
        # tables used by the CasADi functions the dynamics embed
        self.cum_len = self.key_pts[:, 3].copy()               # tval of pw_lin  (:221)
        self.curv_breaks = self.key_pts[1:-1, 3].copy()        # tval of pw_const (:204)
        self.curv_vals = self.key_pts[1:, 5].copy()            # val  of pw_const (:204)
        seg_len, curv = self.key_pts[:, 4], self.key_pts[:, 5]
        abs_angs = np.zeros(self.key_pts.shape[0] + 1)
        for i in range(self.key_pts.shape[0]):                  # :211-216
            abs_angs[i + 1] = abs_angs[i] if curv[i] == 0 else abs_angs[i] + seg_len[i] * curv[i]
        self.cum_ang = abs_angs[1:]
        # pw_lin = pw_const over interior breakpoints of the per-segment lines
        n = len(self.cum_len)
        self.slopes = np.array([(self.cum_ang[i + 1] - self.cum_ang[i]) / (self.cum_len[i + 1] - self.cum_len[i])
                                for i in range(n - 1)])

    @staticmethod
    def _key_pts(cl_segs, init_pos):
        n_segs = cl_segs.shape[0]
        kp = np.zeros((n_segs + 1, 6))
        kp[0, :3] = init_pos
        for i in range(1, n_segs + 1):
            x_prev, y_prev, psi_prev, cum_s_prev = kp[i - 1, :4]
            l, r = cl_segs[i - 1]
            if r == 0:
                psi = psi_prev
                x = x_prev + l * np.cos(psi_prev)
                y = y_prev + l * np.sin(psi_prev)
                curvature = 0
            else:
                x_c = x_prev - r * np.sin(psi_prev)
                y_c = y_prev + r * np.cos(psi_prev)
                theta = l / r
                x = x_c + r * np.sin(psi_prev + theta)
                y = y_c - r * np.cos(psi_prev + theta)
                curvature = 1 / r
                psi = wrap_angle(psi_prev + theta)
            kp[i] = [x, y, psi, cum_s_prev + l, l, curvature]
        return kp

    # -- the CasADi lookups (scalar or ndarray s) ---------------------------
    def s_bar(self, s):
        L = self.track_length
        return np.fmod(np.fmod(s, L) + L, L)          # ca.fmod == C fmod

    def seg_index(self, s):
        """Index selected by CasADi ``pw_const``: number of interior breakpoints with
        ``s_bar >= break``."""
        sb = self.s_bar(np.asarray(s, dtype=np.float64))
        return (sb[..., None] >= self.curv_breaks).sum(axis=-1)

    def curvature(self, s):
        # pw_const(t, tv, v) = v0 + sum_i (v_{i+1}-v_i)*(t>=tv_i); zero derivative wrt t
        sb = self.s_bar(np.asarray(s, dtype=np.float64))
        ret = np.full(sb.shape, self.curv_vals[0])
        for i, b in enumerate(self.curv_breaks):
            ret = ret + (self.curv_vals[i + 1] - self.curv_vals[i]) * (sb >= b)
        return ret

    def tangent(self, s):
        """Returns (psi_t, dpsi_t/ds).  pw_lin: per-segment line selected by pw_const over the
        interior breakpoints; d fmod(a,b)/da = 1 so the slope passes straight through."""
        sb = self.s_bar(np.asarray(s, dtype=np.float64))
        lseg = [self.cum_ang[i] + self.slopes[i] * (sb - self.cum_len[i]) for i in range(len(self.slopes))]
        ret, dret = lseg[0], np.full(sb.shape, self.slopes[0])
        for i, b in enumerate(self.curv_breaks):
            ind = (sb >= b)
            ret = ret + (lseg[i + 1] - lseg[i]) * ind
            dret = dret + (self.slopes[i + 1] - self.slopes[i]) * ind
        return ret, dret

    # -- geometry -------------------------------------------------------------
    def local_to_global(self, cl_coord):
        s = cl_coord[0]
        while s < 0:
            s += self.track_length
        while s >= self.track_length:
            s -= self.track_length
        e_y, e_psi = cl_coord[1], cl_coord[2]
        kp = self.key_pts
        i_s = np.where(s >= kp[:, 3])[0][-1]
        i_f = i_s + 1
        x_s, y_s, psi_s = kp[i_s, 0], kp[i_s, 1], kp[i_s, 2]
        x_f, y_f, psi_f, curve_f = kp[i_f, 0], kp[i_f, 1], kp[i_f, 2], kp[i_f, 5]
        l = kp[i_f, 4]
        d = s - kp[i_s, 3]
        if curve_f == 0:
            x = x_s + (x_f - x_s) * d / l + e_y * np.cos(psi_f + np.pi / 2)
            y = y_s + (y_f - y_s) * d / l + e_y * np.sin(psi_f + np.pi / 2)
            psi = wrap_angle(psi_f + e_psi)
        else:
            r = 1 / curve_f
            dir = _sign(r)
            x_c = x_s + np.abs(r) * np.cos(psi_s + dir * np.pi / 2)
            y_c = y_s + np.abs(r) * np.sin(psi_s + dir * np.pi / 2)
            span_ang = d / np.abs(r)
            psi_d = wrap_angle(psi_s + dir * span_ang)
            ang_norm = wrap_angle(psi_s + dir * np.pi / 2)
            ang = -_sign(ang_norm) * (np.pi - np.abs(ang_norm))
            x = x_c + (np.abs(r) - dir * e_y) * np.cos(ang + dir * span_ang)
            y = y_c + (np.abs(r) - dir * e_y) * np.sin(ang + dir * span_ang)
            psi = wrap_angle(psi_d + e_psi)
        return (x, y, psi)


def chicane_track(enter=1.0, curve1_len=4.0, curve1_angle=math.pi / 4, mid=1.0, curve2_len=4.0,
                  curve2_angle=math.pi / 4, exit=5.0, width=2.0, slack=0.8, mirror=False):
    s1, s2 = (1, -1) if mirror else (-1, 1)
    segs = [[enter, 0], [curve1_len, s1 * curve1_len / curve1_angle], [mid, 0],
            [curve2_len, s2 * curve2_len / curve2_angle], [exit, 0]]
    return RadiusArclengthTrack(width, slack, segs)


def curve_track(enter=1.0, curve_len=8.0, curve_angle=math.pi / 4, exit=5.0, width=2.0, slack=0.8, ccw=True):
    s = 1 if ccw else -1
    segs = [[enter, 0], [curve_len, s * curve_len / curve_angle], [exit, 0]]
    return RadiusArclengthTrack(width, slack, segs)
