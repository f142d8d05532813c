// tensor-core rollout kernels (forward, step Jacobians, reverse pass) instantiated for one environment
#include "sgbpo_rollout_tc.cuh"
namespace sgbpo {
namespace tcr {
SGBPO_INSTANTIATE_ROLLOUT_TC(ENV_BALANCE_CARTPOLE)
}
}
