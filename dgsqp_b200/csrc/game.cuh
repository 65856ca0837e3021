// Selects the game the solver stack is compiled for.  The solver (qp_gi.cuh, lsqr.cuh, sqp_v1.cuh, sqp_v2.cuh) sees a
// game only through: GameDesc, Dims/make_dims, EvalBuf, DG_AB_SZ / DG_T2_SZ / DG_HC_SZ (per-(stage, agent) sizes of
// the derivative tables), game_row_table, game_rollout, game_linearize, game_constraints, game_costates, game_sens,
// game_contract, game_gradients, game_hessian, game_G_times, game_GT_times, game_G_row, game_costs.
// One translation unit = one game (dgsqp_abi.cu: racing games; dgsqp_merge_abi.cu: merge game, -DDG_GAME_MERGE).
#pragma once
#ifdef DG_GAME_MERGE
#include "merge_game.cuh"
#else
#include "racing_game.cuh"
#endif
