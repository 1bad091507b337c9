// gae.cu — GAE / return computation as a reverse scan over the rollout buffer.
//
// Replaces the NumPy loop of runner.py:120-149 for one stream (rewards/values or costRewards/costValues):
//     delta  = (r[t] + g * v[t+1]) - v[t]            g  = float(GAMMA * next_nonterminal)
//     adv[t] = delta + gl * adv[t+1]                 gl = float(GAMMA * LAM * next_nonterminal)
//     ret[t] = adv[t] + v[t]
// NumPy rounds every *, +, - separately in f32, so the kernel uses __fmul_rn/__fadd_rn/__fsub_rn (never contracted
// into FMA) and is bit-exact with the reference.  Thread = 4 adjacent columns (one column = one (world, agent) pair),
// serial over t = T-1..0 with the loads of 8 time steps in flight; 12 B of HBM traffic per element and stream.
// The rollout has two streams (rewards/values and costRewards/costValues, runner.py:146-149): mapf_gae2 scans both in ONE
// launch (blockIdx.y = stream), which doubles the columns in flight and halves the launch / tail overhead.
#include "common.cuh"

namespace mapf {

namespace {

constexpr int GAE_UNROLL = 8;

template <int V>
struct Vec;
template <>
struct Vec<4> { using T = float4; };
template <>
struct Vec<1> { using T = float; };

template <int V>
__device__ __forceinline__ void load(const float *p, float (&o)[V]) {
    if (V == 4) { const float4 x = __ldcs(reinterpret_cast<const float4 *>(p)); o[0] = x.x; o[1 % V] = x.y; o[2 % V] = x.z; o[3 % V] = x.w; }
    else o[0] = __ldcs(p);
}
template <int V>
__device__ __forceinline__ void store(float *p, const float (&o)[V]) {
    if (V == 4) __stcs(reinterpret_cast<float4 *>(p), make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]));
    else __stcs(p, o[0]);
}

struct GaeStream {
    const float *r, *v, *last_v;
    float *ret, *adv;
};

template <int V>
__global__ void __launch_bounds__(128)
gae_kernel(const GaeStream s0, const GaeStream s1, const uint8_t *__restrict__ nonterminal, const float g, const float gl,
           const int T, const long long cols) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (col >= cols) return;
    const GaeStream &S = blockIdx.y ? s1 : s0;
    const float *__restrict__ r = S.r, *__restrict__ v = S.v, *__restrict__ last_v = S.last_v;
    float *__restrict__ ret = S.ret, *__restrict__ adv = S.adv;
    float nv[V], last[V];
    load<V>(last_v + col, nv);
#pragma unroll
    for (int k = 0; k < V; ++k) last[k] = 0.0f;
    int t = T - 1;
    while (t >= 0) {
        const int nb = t + 1 < GAE_UNROLL ? t + 1 : GAE_UNROLL;
        float rr[GAE_UNROLL][V], vv[GAE_UNROLL][V];
        uint8_t nt[GAE_UNROLL][V];
#pragma unroll
        for (int u = 0; u < GAE_UNROLL; ++u) {
            if (u < nb) {
                const size_t o = (size_t)(t - u) * cols + col;
                load<V>(r + o, rr[u]);
                load<V>(v + o, vv[u]);
#pragma unroll
                for (int k = 0; k < V; ++k) nt[u][k] = nonterminal ? nonterminal[o + k] : (uint8_t)1;
            }
        }
#pragma unroll
        for (int u = 0; u < GAE_UNROLL; ++u) {
            if (u < nb) {
                float a[V], rt[V];
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const float m1 = nt[u][k] ? __fmul_rn(g, nv[k]) : 0.0f;
                    const float delta = __fsub_rn(__fadd_rn(rr[u][k], m1), vv[u][k]);
                    const float m2 = nt[u][k] ? __fmul_rn(gl, last[k]) : 0.0f;
                    a[k] = __fadd_rn(delta, m2);
                    rt[k] = __fadd_rn(a[k], vv[u][k]);
                    last[k] = a[k];
                    nv[k] = vv[u][k];
                }
                const size_t o = (size_t)(t - u) * cols + col;
                store<V>(ret + o, rt);
                if (adv) store<V>(adv + o, a);
            }
        }
        t -= nb;
    }
}

}  // namespace

cudaError_t launch_gae2(const float *r, const float *v, const float *last_v, const float *cr, const float *cv,
                        const float *last_cv, const uint8_t *nonterminal, float g, float gl, int T, long long cols, float *ret,
                        float *cret, float *adv, float *cadv, cudaStream_t stream) {
    if (cols <= 0 || T <= 0) return cudaSuccess;
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool two = cr != nullptr;
    bool vec = (cols % 4 == 0) && al(r) && al(v) && al(last_v) && al(ret) && (adv == nullptr || al(adv));
    if (two) vec = vec && al(cr) && al(cv) && al(last_cv) && al(cret) && (cadv == nullptr || al(cadv));
    const GaeStream s0{r, v, last_v, ret, adv}, s1{cr, cv, last_cv, cret, cadv};
    if (vec) {
        const long long threads = cols / 4;
        const dim3 grid((unsigned)((threads + 127) / 128), two ? 2 : 1);
        gae_kernel<4><<<grid, 128, 0, stream>>>(s0, s1, nonterminal, g, gl, T, cols);
    } else {
        const dim3 grid((unsigned)((cols + 127) / 128), two ? 2 : 1);
        gae_kernel<1><<<grid, 128, 0, stream>>>(s0, s1, nonterminal, g, gl, T, cols);
    }
    return cudaGetLastError();
}

cudaError_t launch_gae(const float *r, const float *v, const float *last_v, const uint8_t *nonterminal, float g, float gl,
                       int T, long long cols, float *ret, float *adv, cudaStream_t stream) {
    return launch_gae2(r, v, last_v, nullptr, nullptr, nullptr, nonterminal, g, gl, T, cols, ret, nullptr, adv, nullptr, stream);
}

}  // namespace mapf
