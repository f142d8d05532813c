// explicit instantiation of the warp-MMA forward kernel for one environment (build parallelism)
#include "sgbpo_rollout_wm.cuh"
namespace sgbpo { namespace wm {
SGBPO_INSTANTIATE_ROLLOUT_WM(3)
} }
