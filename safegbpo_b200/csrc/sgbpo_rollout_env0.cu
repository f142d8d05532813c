// fused rollout kernels instantiated for one environment (parallel compilation unit)
#include "sgbpo_rollout_impl.cuh"
namespace sgbpo {
SGBPO_INSTANTIATE_ROLLOUT_ENV(ENV_BALANCE_PENDULUM)
}
