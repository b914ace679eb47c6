// Per-sample gradients of the PPO loss (reference agent/ppo.py:173-204) with respect
// to the network outputs, exactly as autograd derives them:
//   ratio = exp(logp_new - logp_old); pg = mean(max(-A ratio, -A clamp(ratio, 1-c, 1+c)))
//   v_loss = 0.5 mean(max((v - R)^2, (clamp(v - v_old, -c, c) + v_old - R)^2))
// torch.maximum splits the gradient evenly on ties and clamp passes it on its closed
// interval; both conventions are reproduced.  The entropy bonus has no gradient
// because log_std is a buffer (agent/ppo.py:24).  Shared by ppo_loss_grad_kernel and
// the fused minibatch-gradient kernel.
#pragma once

namespace rk {

// d loss / d mu (2 components) and this sample's (logp_old - logp_new); n = minibatch size of the mean
__device__ __forceinline__ void ppo_policy_grad(float m0, float m1, float a0, float a1, float old_logp, float adv_raw,
                                                float adv_mean, float adv_std, float ls0, float ls1, float clip, int n,
                                                float& dmu0, float& dmu1, float& kl) {
    const float s0 = expf(ls0), s1 = expf(ls1);
    const float kLogSqrt2Pi = 0.9189385332046727f;
    const float d0 = a0 - m0, d1 = a1 - m1;
    const float lp = (-(d0 * d0) / (2.f * s0 * s0) - ls0 - kLogSqrt2Pi) + (-(d1 * d1) / (2.f * s1 * s1) - ls1 - kLogSqrt2Pi);
    const float logratio = lp - old_logp;
    kl = -logratio;
    const float ratio = expf(logratio);
    const float A = (adv_raw - adv_mean) / (adv_std + 1e-8f);
    const float lo = 1.f - clip, hi = 1.f + clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float pg1 = -A * ratio, pg2 = -A * rc;
    const bool inside = ratio >= lo && ratio <= hi;
    const float w1 = pg1 > pg2 ? 1.f : (pg1 == pg2 ? 0.5f : 0.f);  // share of the max() gradient going to pg1
    const float dratio = w1 * (-A) + (1.f - w1) * (inside ? -A : 0.f);
    const float g_lp = dratio * ratio / (float)n;                   // d loss / d logp_new
    dmu0 = g_lp * d0 / (s0 * s0);
    dmu1 = g_lp * d1 / (s1 * s1);
}

// d loss / d v of the clipped value loss
__device__ __forceinline__ float ppo_value_grad(float v, float R, float v_old, float clip, float vf_coef, int n) {
    const float dvv = v - v_old;
    const float vclip = v_old + fminf(fmaxf(dvv, -clip), clip);
    const float l1 = (v - R) * (v - R), l2 = (vclip - R) * (vclip - R);
    const float g1 = 2.f * (v - R), g2 = (dvv >= -clip && dvv <= clip) ? 2.f * (vclip - R) : 0.f;
    const float u1 = l1 > l2 ? 1.f : (l1 == l2 ? 0.5f : 0.f);
    return vf_coef * 0.5f * (u1 * g1 + (1.f - u1) * g2) / (float)n;
}

}  // namespace rk
