// Scratch slots of the multi-kernel reductions (grad_norm -> adam_apply, episode_stats, shac_loss).  Their partials and
// tickets live in device globals; each stream that calls them owns one slot, so calls on DISTINCT streams do not share state
// (include/sgbpo.h: "thread-safe with respect to distinct streams").  Up to N_SLOTS streams are told apart; a process using
// more streams than that for these entry points has the extra ones share slots round-robin (documented limit).
#pragma once
#include <cuda_runtime.h>

namespace sgbpo {
constexpr int N_SLOTS = 16;
int stream_slot(cudaStream_t st);      // host, thread-safe (sgbpo_optim.cu)
}  // namespace sgbpo
