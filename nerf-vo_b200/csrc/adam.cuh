// torch.optim.Adam's single-tensor update (torch/optim/adam.py, amsgrad = False, weight_decay = 0, maximize = False), term by term:
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
//   bias_correction{1,2} = 1 - beta{1,2} ** step; step_size = lr / bias_correction1
//   denom = (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps); param.addcdiv_(exp_avg, denom, value = -step_size)
// The scalar constants are evaluated in double precision like torch's python floats and rounded to fp32 once (1 - 0.999 in fp32 would
// already be off by 1.3e-5 relative).
#pragma once

struct AdamC {
    float w1, b2, w2;     // 1 - beta1, beta2, 1 - beta2
    float neg_step;       // -lr / bias_correction1
    float bc2_sqrt;       // sqrt(bias_correction2)
    float eps;
    float grad_scale;     // 1 / world_size for the data-parallel mean
};

__device__ __forceinline__ AdamC adam_constants(int step, double lr, double b1, double b2, double eps, float grad_scale) {
    AdamC c;
    c.w1 = (float)(1.0 - b1);
    c.b2 = (float)b2;
    c.w2 = (float)(1.0 - b2);
    c.neg_step = (float)(-lr / (1.0 - pow(b1, (double)step)));
    c.bc2_sqrt = (float)sqrt(1.0 - pow(b2, (double)step));
    c.eps = (float)eps;
    c.grad_scale = grad_scale;
    return c;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamC& c) {
    const float gr = g * c.grad_scale;
    m = m + c.w1 * (gr - m);
    v = v * c.b2 + (c.w2 * gr) * gr;
    p = p + c.neg_step * (m / (sqrtf(v) / c.bc2_sqrt + c.eps));
}
