// fused rollout kernels instantiated for one environment and one dtype (parallel compilation unit)
#include "sgbpo_rollout_impl.cuh"
namespace sgbpo {
SGBPO_INSTANTIATE_ROLLOUT_ENV_T(ENV_BALANCE_QUADROTOR, float)
}
