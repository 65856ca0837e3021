"""Constant-curvature-segment tracks (host side).

Same constructor arguments as the reference's ``ChicaneTrack`` / ``CurveTrack`` / ``StraightTrack``
(``DGSQP/tracks/track_lib.py:14-87``) on top of a ``RadiusArclengthTrack``
(``DGSQP/tracks/radius_arclength_track.py``).  Only what the solve path and the instance samplers
need is provided: key points (``:361-408``), the segment tables behind the curvature / tangent-angle
look-ups (``:199-225``; the look-ups themselves run on the GPU), and ``local_to_global`` (``:752-807``),
here vectorised over instances.
"""
import numpy as np


class RadiusArclengthTrack:
    def __init__(self, track_width=None, slack=None, cl_segs=None):
        self.track_width, self.slack = track_width, slack
        self.cl_segs = None if cl_segs is None else np.asarray(cl_segs, dtype=np.float64)
        self.circuit = False
        if self.cl_segs is not None:
            self.initialize()

    def initialize(self, track_width=None, slack=None, cl_segs=None, init_pos=(0.0, 0.0, 0.0)):
        if track_width is not None:
            self.track_width = track_width
        if slack is not None:
            self.slack = slack
        if cl_segs is not None:
            self.cl_segs = np.asarray(cl_segs, dtype=np.float64)
        self.half_width = self.track_width / 2
        self.n_segs = self.cl_segs.shape[0]
        self.key_pts = self.get_track_key_pts(self.cl_segs, init_pos)
        self.track_length = self.key_pts[-1, 3]
        self.circuit = bool(np.isclose(self.key_pts[0, 0], self.key_pts[-1, 0])
                            and np.isclose(self.key_pts[0, 1], self.key_pts[-1, 1]))

    @staticmethod
    def get_track_key_pts(cl_segs, init_pos):
        """Rows [x, y, psi, cumulative length, segment length, signed curvature] at segment ends."""
        pts = [np.array([init_pos[0], init_pos[1], init_pos[2], 0.0, 0.0, 0.0])]
        for length, radius in np.asarray(cl_segs, dtype=np.float64):
            x0, y0, psi0, cum0 = pts[-1][:4]
            if radius == 0:
                x1, y1, psi1, curv = x0 + length * np.cos(psi0), y0 + length * np.sin(psi0), psi0, 0.0
            else:
                sweep = length / radius
                xc, yc = x0 - radius * np.sin(psi0), y0 + radius * np.cos(psi0)
                x1, y1 = xc + radius * np.sin(psi0 + sweep), yc - radius * np.cos(psi0 + sweep)
                psi1, curv = _wrap(psi0 + sweep), 1.0 / radius
            pts.append(np.array([x1, y1, psi1, cum0 + length, length, curv]))
        return np.vstack(pts)

    # segment tables consumed by the C ABI (dgsqp_racing_game.track_seg_len / track_seg_curv)
    def segment_lengths(self):
        return self.key_pts[1:, 4].copy()

    def segment_curvatures(self):
        return self.key_pts[1:, 5].copy()

    def get_halfwidth(self, s=None):
        return self.half_width

    def local_to_global(self, cl_coord):
        """(s, e_y, e_psi) -> (x, y, psi); scalars or equally shaped arrays."""
        s, e_y, e_psi = (np.asarray(v, dtype=np.float64) for v in cl_coord)
        scalar = s.ndim == 0
        s, e_y, e_psi = np.atleast_1d(s).copy(), np.atleast_1d(e_y), np.atleast_1d(e_psi)
        L = self.track_length
        # the reference wraps by repeated +-L
        s = np.where(s < 0, s + L * np.ceil(-s / L), s)
        s = np.where(s >= L, s - L * np.floor(s / L), s)
        kp = self.key_pts
        i_s = np.clip(np.searchsorted(kp[:, 3], s, side="right") - 1, 0, kp.shape[0] - 2)
        i_f = i_s + 1
        x_s, y_s, psi_s = kp[i_s, 0], kp[i_s, 1], kp[i_s, 2]
        x_f, y_f, psi_f, curv = kp[i_f, 0], kp[i_f, 1], kp[i_f, 2], kp[i_f, 5]
        seg_len, d = kp[i_f, 4], s - kp[i_s, 3]
        straight = curv == 0
        # straight segments
        xs = x_s + (x_f - x_s) * d / seg_len + e_y * np.cos(psi_f + np.pi / 2)
        ys = y_s + (y_f - y_s) * d / seg_len + e_y * np.sin(psi_f + np.pi / 2)
        ps = _wrap(psi_f + e_psi)
        # curved segments
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(straight, 1.0, 1.0 / np.where(straight, 1.0, curv))
        sgn = np.where(r >= 0, 1.0, -1.0)
        ra = np.abs(r)
        xc = x_s + ra * np.cos(psi_s + sgn * np.pi / 2)
        yc = y_s + ra * np.sin(psi_s + sgn * np.pi / 2)
        span = d / ra
        psi_d = _wrap(psi_s + sgn * span)
        ang_n = _wrap(psi_s + sgn * np.pi / 2)
        ang = -np.where(ang_n >= 0, 1.0, -1.0) * (np.pi - np.abs(ang_n))
        xcv = xc + (ra - sgn * e_y) * np.cos(ang + sgn * span)
        ycv = yc + (ra - sgn * e_y) * np.sin(ang + sgn * span)
        pc = _wrap(psi_d + e_psi)
        x, y, psi = np.where(straight, xs, xcv), np.where(straight, ys, ycv), np.where(straight, ps, pc)
        if scalar:
            return float(x[0]), float(y[0]), float(psi[0])
        return x, y, psi

    def local_to_global_typed(self, data):
        x, y, psi = self.local_to_global((data.p.s, data.p.x_tran, data.p.e_psi))
        data.x.x, data.x.y, data.e.psi = x, y, psi
        return -1


def _wrap(theta):
    theta = np.asarray(theta, dtype=np.float64)
    return np.where(theta < -np.pi, theta + 2 * np.pi, np.where(theta > np.pi, theta - 2 * np.pi, theta))


class StraightTrack(RadiusArclengthTrack):
    def __init__(self, length, width, slack, phase_out=False):
        segs = [[length, 0], [10, 0]] if phase_out else [[length, 0]]
        super().__init__(width, slack, segs)
        self.phase_out = phase_out


class CurveTrack(RadiusArclengthTrack):
    def __init__(self, enter_straight_length, curve_length, curve_swept_angle, exit_straight_length, width, slack,
                 phase_out=False, ccw=True):
        radius = (1 if ccw else -1) * curve_length / curve_swept_angle
        segs = [[enter_straight_length, 0], [curve_length, radius], [exit_straight_length, 0]]
        if phase_out:
            segs.append([10, 0])
        super().__init__(width, slack, segs)
        self.phase_out = phase_out


class ChicaneTrack(RadiusArclengthTrack):
    def __init__(self, enter_straight_length, curve1_length, curve1_swept_angle, mid_straight_length, curve2_length,
                 curve2_swept_angle, exit_straight_length, width, slack, phase_out=False, mirror=False):
        s1, s2 = (1, -1) if mirror else (-1, 1)
        segs = [[enter_straight_length, 0], [curve1_length, s1 * curve1_length / curve1_swept_angle],
                [mid_straight_length, 0], [curve2_length, s2 * curve2_length / curve2_swept_angle],
                [exit_straight_length, 0]]
        if phase_out:
            segs.append([10, 0])
        super().__init__(width, slack, segs)
        self.phase_out = phase_out
