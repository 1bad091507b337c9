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

// checksum_rows_kernel: 64-bit position-keyed checksum of every row of a [rows, words] u32 array (mapf_checksum_rows).
// One CTA per row at a time (grid-stride over rows), 16-byte loads when the row allows, order-independent sum.
__device__ __forceinline__ unsigned long long mix_word(uint32_t x, unsigned long long i) {
    unsigned long long z = ((unsigned long long)x + 1ull) * 0x9E3779B97F4A7C15ull + i * 0xC2B2AE3D27D4EB4Full;
    z = (z ^ (z >> 29)) * 0xBF58476D1CE4E5B9ull;
    return z ^ (z >> 32);
}

__global__ void __launch_bounds__(256) checksum_rows_kernel(const uint32_t *__restrict__ data, const long long rows,
                                                            const long long words, unsigned long long *__restrict__ out) {
    __shared__ unsigned long long part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const uint32_t *row = data + r * words;
        unsigned long long acc = 0;
        if (((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (words & 3) == 0) {
            const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
            for (long long q = threadIdx.x; q < (words >> 2); q += blockDim.x) {
                const uint4 x = __ldcs(row4 + q);
                acc += mix_word(x.x, 4 * q) + mix_word(x.y, 4 * q + 1) + mix_word(x.z, 4 * q + 2) + mix_word(x.w, 4 * q + 3);
            }
        } else {
            for (long long i = threadIdx.x; i < words; i += blockDim.x) acc += mix_word(__ldcs(row + i), i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += part[k];
            out[r] = t;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_checksum_rows(const uint32_t *data, long long rows, long long words, unsigned long long *out, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const int blocks = (int)(rows < 148 * 8 ? rows : 148 * 8);
    checksum_rows_kernel<<<blocks, 256, 0, s>>>(data, rows, words, out);
    return cudaGetLastError();
}

cudaError_t launch_sample_actions(const float *ps, long long rows, unsigned long long seed, uint32_t draw, int8_t *actions,
                                  float *chosen_p, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const long long need = (rows + 255) / 256;
    const int blocks = (int)(need < 148 * 8 ? need : 148 * 8);
    sample_actions_kernel<<<blocks, 256, 0, s>>>(ps, rows, seed, draw, actions, chosen_p);
    return cudaGetLastError();
}

}  // namespace mapf
