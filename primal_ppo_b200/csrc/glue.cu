// glue.cu — policy-side glue of the rollout loop that the reference runs on the host between the network and the env.
//
// sample_actions_kernel: categorical sampling of the joint action from the policy's probabilities.  Reference:
//   `actions[i] = np.random.choice(range(N_ACTIONS), p=ps[i].ravel())` per agent on the host (model.py:38-40, 58-59),
//   which forces a device->host copy of the probabilities and a host->device copy of the actions every step.  Here one
//   thread draws one agent's action with Philox4x32-10 keyed by (seed; row, draw counter): inverse-CDF on the f32
//   probabilities.  Equal to np.random.choice in distribution (not in bits: NumPy uses the global MT19937 stream);
//   bit-identical to the oracle's restatement of the same draw (oracle/mapf_oracle.c: orc_sample_actions).
#include "common.cuh"

namespace mapf {

namespace {

constexpr uint64_t SAMPLE_SALT = 0x53414D504C455221ull;   // keeps this stream apart from the fixActions draws

__global__ void __launch_bounds__(256) sample_actions_kernel(const float *__restrict__ ps, const long long rows,
                                                             const unsigned long long seed, const uint32_t draw,
                                                             int8_t *__restrict__ actions, float *__restrict__ chosen_p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const float *p = ps + i * NA;
        float q[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) q[k] = __ldg(p + k);
        float total = q[0];
#pragma unroll
        for (int k = 1; k < NA; ++k) total = __fadd_rn(total, q[k]);
        const uint32_t r = philox_draw(seed ^ SAMPLE_SALT, (uint32_t)i, draw, (uint32_t)((unsigned long long)i >> 32));
        const float u = __fmul_rn((float)(r >> 8), 5.9604644775390625e-08f);        // [0, 1): 24 random bits * 2^-24
        const float x = __fmul_rn(u, total);
        int a = 0;
        float c = q[0];
#pragma unroll
        for (int k = 1; k < NA; ++k) {
            if (x >= c && a == k - 1) { a = k; c = __fadd_rn(c, q[k]); }
        }
        while (a > 0 && !(q[a] > 0.0f)) --a;       // rounding at the top of the CDF must not pick a zero-probability action
        actions[i] = (int8_t)a;
        if (chosen_p) chosen_p[i] = q[a];
    }
}

}  // namespace

cudaError_t launch_sample_actions(const float *ps, long long rows, unsigned long long seed, uint32_t draw, int8_t *actions,
                                  float *chosen_p, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const long long need = (rows + 255) / 256;
    const int blocks = (int)(need < 148 * 8 ? need : 148 * 8);
    sample_actions_kernel<<<blocks, 256, 0, s>>>(ps, rows, seed, draw, actions, chosen_p);
    return cudaGetLastError();
}

}  // namespace mapf
