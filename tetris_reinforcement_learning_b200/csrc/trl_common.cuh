// trl_common.cuh — error plumbing and the library-owned device workspace.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/trl.h"

// Records the error text for trl_last_error(); returns TRL_OK or TRL_E_CUDA.
int trl_check(cudaError_t e);

// Library-owned, grow-only device scratch (used only by the *_host entry points and by
// trl_movegen when the caller passes no mask buffer).  Returns nullptr on failure.
enum TrlWorkspaceSlot { TRL_WS_MOVEGEN_MASK = 0, TRL_WS_HOST_STAGE = 1, TRL_WS_TRUNK_COUNTER = 2, TRL_WS_TRUNK_COUNTER_TAPS = 3, TRL_WS_SLOTS = 4 };
void* trl_workspace(int slot, size_t bytes);

// Streams (0 or 1) used by the *_host entry points (created on first use, non-blocking).
cudaStream_t trl_host_stream(int which = 0);
