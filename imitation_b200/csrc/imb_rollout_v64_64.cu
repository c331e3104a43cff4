// one instantiation of the rollout kernel: policy tower width 64, reward-net width 64
#include "imb_rollout_impl.cuh"
int imb_rl_64_64(const void* A, const void* L, const float* env_params, float* env_obs, const float* pol_params,
                 const float* pol_norm, const float* disc_params, float* rollout, float* ring, float* flat_out,
                 float* aux, const float* noise, const int64_t* state, cudaStream_t st) {
  return launch_rollout<64, 64>(*static_cast<const RolloutArgs*>(A), *static_cast<const DiscLaunch*>(L), env_params,
                                env_obs, pol_params, pol_norm, disc_params, rollout, ring, flat_out, aux, noise, state,
                                st);
}
