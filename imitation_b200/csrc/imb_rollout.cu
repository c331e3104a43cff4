// imb_rollout.cu -- C-ABI entry points of stage 1 (kernels: imb_rollout_impl.cuh).
#include "imb_rollout_impl.cuh"

extern "C" int imb_rollout_row_width(const imb_policy_desc* pol) {
  // padded to a multiple of 4 floats: rows are 16-byte aligned, so the PPO update can stage a minibatch row
  // with one bulk-async copy
  return (pol->d_obs + (pol->discrete ? 1 : pol->d_act) + 5 + 3) / 4 * 4;
}

extern "C" int imb_rollout(const imb_env_desc* env, const float* env_params, float* env_obs,
                           const imb_policy_desc* pol, const float* pol_params, const float* pol_norm,
                           const imb_disc_desc* disc, const float* disc_params, const float* disc_norm,
                           int reward_mode, const imb_ppo_hparams* hp, int64_t n_envs, int64_t n_steps,
                           float* rollout, float* ring, int64_t ring_capacity, float* flat_out, float* aux,
                           const float* noise, int flags, const int64_t* state, void* stream) {
  IMB_REQUIRE(n_envs >= 1 && n_steps >= 1, "rollout needs n_envs, n_steps >= 1");
  IMB_REQUIRE(env->d_obs == pol->d_obs && env->d_act == pol->d_act && env->discrete == pol->discrete,
              "env / policy space mismatch");
  IMB_REQUIRE(pol->hidden >= 1 && pol->hidden <= 64, "policy tower width must be <= 64");
  IMB_REQUIRE(env->d_obs <= IMB_MAX_DIN && env->d_act <= IMB_MAX_DIN, "d_obs/d_act must be <= %d", IMB_MAX_DIN);
  RolloutArgs A;
  A.env = *env;
  A.pol = *pol;
  A.hp = *hp;
  A.reward_mode = reward_mode;
  A.deterministic = (flags & IMB_RF_DETERMINISTIC) ? 1 : 0;
  A.E = n_envs;
  A.T = n_steps;
  A.rw = imb_rollout_row_width(pol);
  A.ring_capacity = ring ? ring_capacity : 0;
  DiscLaunch L;
  memset(&L, 0, sizeof(L));
  if (reward_mode != 0) {
    IMB_REQUIRE(disc && disc_params, "reward_mode != 0 needs a reward net");
    const int onehot = env->discrete ? env->d_act : env->d_act;
    IMB_REQUIRE(disc->d_obs == env->d_obs && disc->d_act == onehot, "reward net / env space mismatch");
    imb_disc_desc dd = *disc;
    dd.subtract_logp = 0;  // reward_train.predict_processed never subtracts log pi (airl.py:121-124)
    if (int rc = build_launch(&dd, disc_norm, nullptr, L)) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return launch_rollout(A, L, env_params, env_obs, pol_params, pol_norm, disc_params, rollout, ring, flat_out, aux,
                        noise, state, st);
}

extern "C" int imb_rollout_advance(int64_t* state, int64_t n_envs, int64_t n_steps, int32_t horizon,
                                   int64_t ring_capacity, void* stream) {
  k_rollout_advance<<<1, 1, 0, (cudaStream_t)stream>>>(state, n_envs, n_steps, horizon, ring_capacity);
  IMB_CHECK_LAUNCH("k_rollout_advance");
  return 0;
}

extern "C" int imb_gae(float* rollout, int32_t rw, int32_t col_value, int64_t n_envs, int64_t n_steps,
                       const float* aux, float gamma, float gae_lambda, const int64_t* state_before,
                       int32_t horizon, void* stream) {
  k_gae<<<(int)((n_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rollout, rw, col_value, n_envs, n_steps, aux,
                                                                      gamma, gae_lambda, state_before, horizon);
  IMB_CHECK_LAUNCH("k_gae");
  return 0;
}

extern "C" int imb_env_reset(float* env_obs, int64_t n_envs, const imb_env_desc* env, const int64_t* state,
                             void* stream) {
  k_env_reset<<<(int)((n_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(env_obs, n_envs, env->d_obs, env->seed,
                                                                            env->env_id_offset, state);
  IMB_CHECK_LAUNCH("k_env_reset");
  return 0;
}
