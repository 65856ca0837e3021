// Host-side translation of the C-ABI descriptors into the device structs.
// The track tables restate RadiusArclengthTrack.get_track_key_pts / get_curvature_casadi_fn /
// get_tangent_angle_casadi_fn (DGSQP/tracks/radius_arclength_track.py:361-408,199-225).
#pragma once
#include <cstdlib>
#include "../../include/dgsqp_b200.h"
#include "game.cuh"
#include "sqp_v2.cuh"

#ifdef DG_GAME_MERGE
typedef dgsqp_merge_game dg_game_struct;
static inline int dg_fill_game(const dgsqp_merge_game* g, GameDesc* G) {
  if (!g || g->M < 2 || g->M > DG_MAX_AGENTS || g->N < 2 || g->N > 63) return -1;
  if (!(g->mass > 0.0) || !(g->dt > 0.0)) return -1;
  G->M = g->M; G->N = g->N;
  G->veh.inv_m = 1.0 / g->mass; G->veh.dt = g->dt;
  for (int i = 0; i < 2; ++i) { G->w_u[i] = g->input_weight[i]; G->u_ub[i] = g->u_ub[i]; G->u_lb[i] = g->u_lb[i]; }
  for (int i = 0; i < 4; ++i) G->w_q[i] = g->state_weight[i];
  G->term_scale = g->term_scale; G->v_ub = g->v_ub; G->v_lb = g->v_lb; G->lane_r = g->lane_r;
  for (int a = 0; a < DG_MAX_AGENTS; ++a) {
    G->obs_r[a] = g->obs_r[a];
    for (int i = 0; i < 4; ++i) G->goal[a][i] = g->goal[a][i];
    for (int j = 0; j < 2; ++j) {
      const dgsqp_lane_row& s = g->lane[a][j];
      LaneRow& L = G->lane[a][j];
      L.brk = s.brk;
      for (int i = 0; i < 2; ++i) { L.na[i] = s.n_lo[i]; L.nb[i] = s.n_hi[i]; L.pt[i] = s.pt[i]; }
    }
  }
  return 0;
}
#else
typedef dgsqp_racing_game dg_game_struct;
static inline int dg_fill_game(const dgsqp_racing_game* g, GameDesc* G) {
  if (!g || g->M < 2 || g->M > DG_MAX_AGENTS || g->N < 2 || g->N > 63) return -1;   // the packed row table keeps the stage in 6 bits
  if (g->track_nseg < 1 || g->track_nseg > DG_MAX_SEGS) return -1;
  G->M = g->M; G->N = g->N;
  G->veh.Lr = g->L_r; G->veh.L = g->L_f + g->L_r; G->veh.c_da = g->c_da; G->veh.c_dr = g->c_dr; G->veh.c_s = g->c_s;
  G->veh.inv_m = 1.0 / g->mass; G->veh.dt = g->dt;
  for (int i = 0; i < 2; ++i) {
    G->w_u[i] = g->input_weight[i]; G->w_du[i] = g->rate_weight[i];
    G->u_ub[i] = g->u_ub[i]; G->u_lb[i] = g->u_lb[i]; G->rate_ub[i] = g->rate_ub[i]; G->rate_lb[i] = g->rate_lb[i];
  }
  G->c_prog = g->comp_weights[0]; G->c_comp = g->comp_weights[1];
  G->half_width = g->half_width;
  for (int a = 0; a < DG_MAX_AGENTS; ++a) G->obs_r[a] = g->obs_r[a];
  TrackTable& T = G->trk;
  T.nseg = g->track_nseg;
  T.cum_len[0] = 0.0; T.cum_ang[0] = 0.0;
  for (int i = 0; i < T.nseg; ++i) {
    double len = g->track_seg_len[i], cv = g->track_seg_curv[i];
    if (!(len > 0.0)) return -1;
    T.curv[i] = cv;
    T.cum_len[i + 1] = T.cum_len[i] + len;
    T.cum_ang[i + 1] = cv == 0.0 ? T.cum_ang[i] : T.cum_ang[i] + len * cv;
  }
  T.L = T.cum_len[T.nseg];
  for (int i = 0; i < T.nseg; ++i) {
    T.slope[i] = (T.cum_ang[i + 1] - T.cum_ang[i]) / (T.cum_len[i + 1] - T.cum_len[i]);
    if (i + 1 < T.nseg) T.brk[i] = T.cum_len[i + 1];
  }
  return 0;
}
#endif

static inline void dg_fill_common(double time_limit, int qp_warm_start, int iter_log, SolverParams* P) {
  P->time_limit_ns = time_limit > 0.0 ? time_limit * 1e9 : 0.0;
  P->qp_warm = qp_warm_start != 0;
  { const char* e = getenv("DGSQP_QP_WARM"); if (e && (e[0] == '0' || e[0] == '1')) P->qp_warm = e[0] == '1'; }   // A/B override
  P->iter_log = iter_log != 0;
}

static inline int dg_fill_params(const dgsqp_params* p, SolverParams* P) {
  if (!p || p->line_search_iters < 1 || p->sqp_iters < 1 || !(p->mu_vio_thresh >= 0.0)) return -1;
  P->reg = p->reg; P->p_tol = p->p_tol; P->d_tol = p->d_tol; P->beta = p->beta; P->tau = p->tau;
  P->eig_floor = 1e-10;            // DGSQP.py:1294
  P->merit_max = 1e6;              // DGSQP.py:1174
  P->diverge_tol = 1e5;            // DGSQP.py:383
  P->line_search_iters = p->line_search_iters; P->sqp_iters = p->sqp_iters; P->nonmono_ls = p->nonmono_ls;
  P->merit_l1 = p->merit_function == 0; P->conv_approx = p->conv_approx;
  P->rel_tol_req = 3;              // DGSQP.py:56
  P->t_hat = 5;                    // DGSQP.py:1179
  P->dbg_l0_perturb = 0.0;
  P->mu_vio_thresh = p->mu_vio_thresh;
  P->merit_obj = 0;
  dg_fill_common(p->time_limit, p->qp_warm_start, p->iter_log, P);
  P->policy = 1; P->nms = 0; P->nms_frequency = 0; P->nms_memory = 1; P->armijo = 1; P->has_merit_parameter = 0;
  P->reg_decay = 1.0; P->sigma = 0.0; P->gamma = 1.0; P->merit_parameter = 0.0;
  return 0;
}

static inline int dg_fill_params_v2(const dgsqp_v2_params* p, SolverParams* P) {
  if (!p || p->line_search_iters < 1 || p->sqp_iters < 1 || !(p->mu_vio_thresh >= 0.0)) return -1;
  if (p->nms_memory_size < 1 || p->nms_memory_size > DG_V2_MEM_MAX || p->nms_frequency < 0) return -1;
  if (p->merit_function != 0 && p->merit_function != 1) return -1;
  if (p->merit_decrease_condition != 0 && p->merit_decrease_condition != 1) return -1;
  P->reg = p->reg; P->p_tol = p->p_tol; P->d_tol = p->d_tol; P->beta = p->beta; P->tau = p->tau;
  P->eig_floor = 1e-9;             // DGSQP_v2.py:1273
  P->merit_max = 1e6;
  P->diverge_tol = 1e10;           // DGSQP_v2.py:394
  P->line_search_iters = p->line_search_iters; P->sqp_iters = p->sqp_iters; P->nonmono_ls = 0;
  P->merit_l1 = 1; P->conv_approx = 1;
  P->rel_tol_req = 10;             // DGSQP_v2.py:86
  P->t_hat = 0;
  P->dbg_l0_perturb = 0.0;
  P->mu_vio_thresh = p->mu_vio_thresh;
  P->merit_obj = p->merit_function == 1;
  dg_fill_common(p->time_limit, p->qp_warm_start, p->iter_log, P);
  P->policy = 2; P->nms = p->nms != 0; P->nms_frequency = p->nms_frequency; P->nms_memory = p->nms_memory_size;
  P->armijo = p->merit_decrease_condition == 0; P->has_merit_parameter = p->has_merit_parameter != 0;
  P->reg_decay = p->reg_decay; P->sigma = p->merit_decrease; P->gamma = p->delta_decay; P->merit_parameter = p->merit_parameter;
  return 0;
}
