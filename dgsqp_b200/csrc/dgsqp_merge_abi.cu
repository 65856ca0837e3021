// C-ABI constructors of the merge game (include/dgsqp_b200.h: dgsqp_create_merge[_v2]) and its engine: the same solver
// stack and engine source as dgsqp_abi.cu, compiled against merge_game.cuh (see game.cuh, engine.inc).  All other entry
// points are game-independent and dispatch through the handle (dgsqp_abi.cu).
#define DG_GAME_MERGE 1
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <new>

#include "abi_common.h"

namespace {
#include "sqp_v2.cuh"
#include "host_setup.h"
#include "engine.inc"
}  // namespace

extern "C" {

int dgsqp_create_merge(const dgsqp_merge_game* game, const dgsqp_params* params, int device, dgsqp_handle** out) {
  return eng_create(game, params, nullptr, device, out);
}

int dgsqp_create_merge_v2(const dgsqp_merge_game* game, const dgsqp_v2_params* params, int device, dgsqp_handle** out) {
  if (!params) return dg_set_err(DGSQP_EINVAL, "NULL parameters");
  return eng_create(game, nullptr, params, device, out);
}

}  // extern "C"
