// ppo_loss.cu — the elementwise part of the PPO-Lagrangian minibatch loss, fused: advantage
// normalisation (given the all-reduced moments), probability ratio, clipped surrogate, clipped value / cost-value losses,
// entropy, valid-action loss, cost term — forward values AND the gradients with respect to the network outputs — in ONE
// pass over the minibatch.  Reference: Model.train, model.py:104-164 (restated in PyTorch in ppo/loss.py, which stays the
// checked reference of this kernel: tests/test_gpu_ppo.py compares values and gradients).
//
// Per (row, agent) element the reference's ~40 eager tensor ops read and write the same few dozen floats over and over;
// here an element is read once (21 floats + 1 byte), its 12 gradient floats are written once, and the ten scalar sums
// leave through per-block partials in double precision (deterministic: no floating-point atomics).
//
// Gradient conventions are PyTorch's, so that autograd through ppo/loss.py gives the same numbers: clamp passes the
// gradient on the closed interval [min, max]; minimum / maximum split the gradient equally on ties.
#include "common.cuh"

namespace mapf {

namespace {

constexpr int PL_THREADS = 256;
constexpr int PL_STATS = MAPF_PPO_LOSS_STATS;

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float in_closed(float x, float lo, float hi) { return (x >= lo && x <= hi) ? 1.0f : 0.0f; }

template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NS], double *__restrict__ partials) {
    __shared__ double sh[PL_THREADS / 32][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double x = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0) sh[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double t = 0.0;
        for (int w = 0; w < PL_THREADS / 32; ++w) t += sh[w][threadIdx.x];     // fixed order: deterministic
        partials[(size_t)blockIdx.x * NS + threadIdx.x] = t;
    }
}

// moments of (returns - old_v) and (cost_returns - old_cv): [sum a, sum a^2, sum c, sum c^2] per block
__global__ void __launch_bounds__(PL_THREADS)
adv_moments_kernel(const float *__restrict__ ret, const float *__restrict__ cret, const float *__restrict__ old_v,
                   const float *__restrict__ old_cv, const long long n, double *__restrict__ partials) {
    double acc[4] = {0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = (double)__fsub_rn(ret[i], old_v[i]), c = (double)__fsub_rn(cret[i], old_cv[i]);
        acc[0] += a; acc[1] += a * a; acc[2] += c; acc[3] += c * c;
    }
    block_reduce_store<4>(acc, partials);
}

__global__ void __launch_bounds__(PL_THREADS)
ppo_loss_kernel(const MapfPpoLossConfig cfg, const long long n, const float *__restrict__ policy, const float *__restrict__ value,
                const float *__restrict__ cost_value, const float *__restrict__ sig, const float *__restrict__ ret,
                const float *__restrict__ cret, const float *__restrict__ old_v, const float *__restrict__ old_cv,
                const int8_t *__restrict__ actions, const float *__restrict__ old_ps, const float *__restrict__ tv,
                float *__restrict__ g_policy, float *__restrict__ g_value, float *__restrict__ g_cost_value,
                float *__restrict__ g_sig, double *__restrict__ partials) {
    const float clip = cfg.clip_range, lam = cfg.lagrangian;
    const float s = (float)(1.0 / cfg.n_global);                     // every mean is a local sum / global count
    const float a_mean = (float)cfg.adv_mean, a_den = (float)cfg.adv_std + 1e-6f;
    const float c_mean = (float)cfg.cadv_mean, c_den = (float)cfg.cadv_std + 1e-6f;
    double acc[PL_STATS];
#pragma unroll
    for (int k = 0; k < PL_STATS; ++k) acc[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float p[NA], sg[NA], t[NA], op[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) { p[k] = policy[i * NA + k]; sg[k] = sig[i * NA + k]; t[k] = tv[i * NA + k]; op[k] = old_ps[i * NA + k]; }
        const float v = value[i], cv = cost_value[i], R = ret[i], CR = cret[i], ov = old_v[i], ocv = old_cv[i];
        int a = actions[i];
        a = a < 0 ? 0 : (a >= NA ? NA - 1 : a);
        // advantages (model.py:106-113)
        float adv = (R - ov - a_mean) / a_den;
        const float cadv = (CR - ocv - c_mean) / c_den;
        if (cfg.minus_adv_with_cadv) adv = (adv - lam * cadv) / (lam + 1.0f);
        // ratio (model.py:119)
        float new_p = p[0], old_p = op[0];
#pragma unroll
        for (int k = 1; k < NA; ++k) { if (k == a) { new_p = p[k]; old_p = op[k]; } }
        const float cn = clampf(new_p, 1e-6f, 1.0f), co = clampf(old_p, 1e-6f, 1.0f);
        const float ratio = expf(logf(cn) - logf(co));
        const float dratio = ratio / cn * in_closed(new_p, 1e-6f, 1.0f);            // d ratio / d new_p
        // clipped surrogate (model.py:139-143): min(adv*ratio, adv*clamp(ratio))
        const float rc = clampf(ratio, 1.0f - clip, 1.0f + clip);
        const float t1 = adv * ratio, t2 = adv * rc;
        const float inside = in_closed(ratio, 1.0f - clip, 1.0f + clip);
        float dsur;                                                                     // d min / d ratio
        if (t1 < t2) dsur = adv; else if (t1 > t2) dsur = adv * inside; else dsur = 0.5f * adv + 0.5f * adv * inside;
        acc[0] += (double)fminf(t1, t2);
        // entropy (model.py:121)
        float ent = 0.0f;
        float gp[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            const float ck = clampf(p[k], 1e-6f, 1.0f), lk = logf(ck);
            ent -= p[k] * lk;
            gp[k] = cfg.entropy_coef * s * (lk + p[k] / ck * in_closed(p[k], 1e-6f, 1.0f));     // d(-ec * entropy) / d p_k
        }
        acc[1] += (double)ent;
        // value losses (model.py:124-136)
        auto clipped = [&](float nv, float old, float target, float &loss, float &grad) {
            const float d = nv - old, vcl = old + clampf(d, -clip, clip);
            const float e1 = (nv - target) * (nv - target), e2 = (vcl - target) * (vcl - target);
            const float pass = in_closed(d, -clip, clip);                               // d vcl / d nv
            const float g1 = 2.0f * (nv - target), g2 = 2.0f * (vcl - target) * pass;
            loss = fmaxf(e1, e2);
            grad = e1 > e2 ? g1 : (e1 < e2 ? g2 : 0.5f * g1 + 0.5f * g2);
        };
        float lv, gv, lcv, gcv;
        clipped(v, ov, R, lv, gv);
        clipped(cv, ocv, CR, lcv, gcv);
        acc[2] += (double)lv; acc[3] += (double)lcv;
        // valid-action loss (model.py:146-148)
        float vl = 0.0f, gs[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            const float q = 1.0f - sg[k];
            const float c1 = clampf(sg[k], 1e-6f, 1.0f - 1e-6f), c2 = clampf(q, 1e-6f, 1.0f - 1e-6f);
            vl += logf(c1) * t[k] + logf(c2) * (1.0f - t[k]);
            const float d1 = t[k] / c1 * in_closed(sg[k], 1e-6f, 1.0f - 1e-6f), d2 = -(1.0f - t[k]) / c2 * in_closed(q, 1e-6f, 1.0f - 1e-6f);
            gs[k] = -cfg.valid_coef * (s / NA) * (d1 + d2);
        }
        acc[4] += (double)vl;
        // cost term (model.py:155) and statistics
        acc[5] += (double)(ratio * cadv);
        acc[6] += fabsf(ratio - 1.0f) > clip ? 1.0 : 0.0;
        acc[7] += (double)adv; acc[8] += (double)cadv;
        // gradients of  loss = -policy - ec*entropy + vc*critic + validc*valid + cvc*cost_critic + cc*lam*cost
        const float gpa = (-dsur + cfg.cost_coef * lam * cadv) * s * dratio;
        if (g_policy) {
#pragma unroll
            for (int k = 0; k < NA; ++k) g_policy[i * NA + k] = gp[k] + (k == a ? gpa : 0.0f);
        }
        if (g_value) g_value[i] = cfg.value_coef * s * gv;
        if (g_cost_value) g_cost_value[i] = cfg.cost_value_coef * s * gcv;
        if (g_sig) {
#pragma unroll
            for (int k = 0; k < NA; ++k) g_sig[i * NA + k] = gs[k];
        }
    }
    block_reduce_store<PL_STATS>(acc, partials);
}

}  // namespace

cudaError_t launch_adv_moments(const float *ret, const float *cret, const float *old_v, const float *old_cv, long long n,
                               double *partials, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(partials, 0, sizeof(double) * 4 * MAPF_PPO_LOSS_MAX_BLOCKS, s);
    if (e != cudaSuccess || n <= 0) return e;
    const long long need = (n + PL_THREADS - 1) / PL_THREADS;
    const int blocks = (int)(need < MAPF_PPO_LOSS_MAX_BLOCKS ? need : MAPF_PPO_LOSS_MAX_BLOCKS);
    adv_moments_kernel<<<blocks, PL_THREADS, 0, s>>>(ret, cret, old_v, old_cv, n, partials);
    return cudaGetLastError();
}

cudaError_t launch_ppo_loss(const MapfPpoLossConfig &cfg, long long n, const float *policy, const float *value,
                            const float *cost_value, const float *sig, const float *ret, const float *cret, const float *old_v,
                            const float *old_cv, const int8_t *actions, const float *old_ps, const float *tv, float *g_policy,
                            float *g_value, float *g_cost_value, float *g_sig, double *partials, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(partials, 0, sizeof(double) * PL_STATS * MAPF_PPO_LOSS_MAX_BLOCKS, s);
    if (e != cudaSuccess || n <= 0) return e;
    const long long need = (n + PL_THREADS - 1) / PL_THREADS;
    const int blocks = (int)(need < MAPF_PPO_LOSS_MAX_BLOCKS ? need : MAPF_PPO_LOSS_MAX_BLOCKS);
    ppo_loss_kernel<<<blocks, PL_THREADS, 0, s>>>(cfg, n, policy, value, cost_value, sig, ret, cret, old_v, old_cv, actions, old_ps, tv,
                                                  g_policy, g_value, g_cost_value, g_sig, partials);
    return cudaGetLastError();
}

}  // namespace mapf
