// step_observe.cu — one launch for the rollout loop's whole per-step env work (runner.py:64-100): joint-step
// resolution (step_world.cuh) followed, for the same world and by the same warp, by the observation build of the
// post-step state (observe_world.cuh).
//
// Why fuse: step resolution is issue/latency-bound (~1000 warp instructions per world, almost no HBM traffic) while the
// observation store is HBM-write-bound with the SM issue slots ~26 % busy.  Run back to back as two kernels they add
// up (0.13 ms + 0.68 ms for 65 536 worlds); inside one persistent kernel the step work of some warps hides under the
// stores of the others, the post-step cells / goals / human tick go from the step phase to the observe phase in
// registers (no state re-read), and the obstacle bit rows are staged once.
// N <= 32 only (lane = agent).  Results are bit-identical to mapf_step followed by mapf_observe.
#include "common.cuh"
#include "observe_world.cuh"
#include "step_world.cuh"

namespace mapf {

namespace {

using namespace sw;
using namespace ow;

template <int C_T, int F_T, bool VEC4, int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB)
step_observe_kernel(const EnvView v, const int8_t *__restrict__ actions, const MapfStepOut out, float *__restrict__ obs,
                    float *__restrict__ vec, const ObsLayout L, const int per_warp, const int step_off,
                    int *__restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint4 lut[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HP = v.HP, GS = v.GS, nob = v.HP * v.RW, npw = v.PW;
    unsigned char *base = smem_raw + (size_t)warp * per_warp;
    const ObsSmem m = obs_carve(base, L, v.N);
    // the step phase shares the obstacle bit rows and the agent-id grid with the observe phase; its own scratch
    // (trainValid staging, conflict masks, fixActions queue) sits behind the observe layout
    WarpSmem s;
    {
        unsigned char *b = base + step_off;
        s.obits = m.obits;
        s.grid = m.grid;
        s.tv = reinterpret_cast<float *>(b); b += 32 * 5 * 4;
        s.mmask = reinterpret_cast<uint32_t *>(b); b += 32 * 4;
        s.act = reinterpret_cast<int8_t *>(b); b += 32;
        s.cls = reinterpret_cast<int8_t *>(b); b += 32;
        s.st = reinterpret_cast<int8_t *>(b); b += 32;
        s.commit = reinterpret_cast<int8_t *>(b); b += 32;
        s.rep = reinterpret_cast<int8_t *>(b); b += 32;
        s.queue = reinterpret_cast<int8_t *>(b);
    }
    if (threadIdx.x < 16) {
        const uint32_t one = 0x3f800000u, t = threadIdx.x;
        lut[t] = make_uint4((t & 1u) ? one : 0u, (t & 2u) ? one : 0u, (t & 4u) ? one : 0u, (t & 8u) ? one : 0u);
    }
    __syncthreads();

    int w = claim_work(work_counter, 1, lane);      // claim-then-load: see observe.cu
    const uint64_t pol = policy_evict_last();
    const int pf_ahead = prefetch_ahead(v), pf_batch = prefetch_batch(v);
    StepRegs cur, nxt;
    load_step_world<MODE_FUSED>(v, actions, nullptr, w, lane, npw, pol, cur);
    const bool direct_ob = npw > SOBW * 32;
    bool first = true;

    while (w < v.W) {
        const int w1 = claim_work(work_counter, 1, lane);
        load_step_world<MODE_FUSED>(v, actions, nullptr, w1, lane, npw, pol, nxt);   // in flight during this world
        if (lane == 0 && pf_ahead >= 0 && (w1 & (pf_batch - 1)) == 0) prefetch_world_batch(v, actions, w1 + pf_ahead, pol);

        // packed obstacle words: registers -> scratch (the agent-presence rows, zeroed right after) -> padded rows
        if (!direct_ob) {
#pragma unroll
            for (int k = 0; k < SOBW; ++k) if (k * 32 + lane < npw) m.abits[k * 32 + lane] = cur.ob[k];
            __syncwarp();
            expand_obstacle_rows(m.obits, m.abits, v, lane, 32);
        } else {
            expand_obstacle_rows(m.obits, v.obst_pack + (size_t)w * npw, v, lane, 32);
        }
        __syncwarp();
        for (int k = lane; k < nob; k += 32) m.abits[k] = 0;
        if (L.alias || first) {          // the chunk bit string of the previous world overlays the id grid when L.alias
            uint4 *g4 = reinterpret_cast<uint4 *>(m.grid);
            for (int k = lane; k < (HP * GS) / 16; k += 32) g4[k] = make_uint4(0, 0, 0, 0);
            first = false;
        }
        __syncwarp();
        uint32_t new_pw, new_gw;
        resolve_world<MODE_FUSED>(v, out, s, w, lane, cur, pol, new_pw, new_gw);      // leaves the id grid clean
        __syncwarp();
        // the human has ticked: its getNextPos() is the `next` field of the following trace entry
        const int nr = (int16_t)(cur.ht2.y & 0xffff), nc = (int16_t)((uint32_t)cur.ht2.y >> 16);
        observe_world<C_T, F_T, VEC4>(v, L, m, lut, w, lane, new_pw, new_gw, nr, nc, obs, vec);
        w = w1;
        cur = nxt;
    }
    finish_work(work_counter, gridDim.x * (blockDim.x >> 5), lane);
}

template <int C_T, int F_T, bool VEC4, int MINB>
cudaError_t launch_tb(const EnvView &v, const int8_t *actions, const MapfStepOut &out, float *obs, float *vec,
                      const ObsLayout &L, int *counter, cudaStream_t stream) {
    const int step_off = (int)L.total;
    const int per_warp = step_off + 32 * 5 * 4 + 32 * 4 + 5 * 32 + QRING;
    const int wpb = WARPS_PER_BLOCK;
    const size_t smem = (size_t)per_warp * wpb;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = step_observe_kernel<C_T, F_T, VEC4, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > MINB) per_sm = MINB;
    const int need = (v.W + wpb - 1) / wpb;
    const int blocks = need < sms * per_sm ? need : sms * per_sm;
    kern<<<blocks, wpb * 32, smem, stream>>>(v, actions, out, obs, vec, L, per_warp, step_off, counter);
    return cudaGetLastError();
}

// As in observe.cu fewer write streams help store-dominated worlds (>= 32 KB of observations each), but here the step
// phase needs warps to hide behind: in-process A/B at 40x40x32: 2 / 3 / 4 CTAs per SM = 0.772 / 0.722 / 0.731 ms.  The
// 3-CTA instantiation is also COMPILED for 3 CTAs (80 registers instead of 64, no spills): 0.706 vs 0.721 ms.
template <int C_T, int F_T, bool VEC4>
cudaError_t launch_t(const EnvView &v, const int8_t *actions, const MapfStepOut &out, float *obs, float *vec,
                     const ObsLayout &L, int *counter, cudaStream_t stream) {
    const bool big = (size_t)v.N * L.PB * (L.out_bf16 ? 2 : 4) >= 32768;
    if (big) return launch_tb<C_T, F_T, VEC4, 3>(v, actions, out, obs, vec, L, counter, stream);
    return launch_tb<C_T, F_T, VEC4, 4>(v, actions, out, obs, vec, L, counter, stream);
}

}  // namespace

// chunk of agents handled per phase-1 pass of the observation build (same rule as launch_observe)
static int obs_chunk(const EnvView &v) {
    const int PB = v.C * v.F * v.F;
    int CH = v.N < 32 ? v.N : 32;
    while (CH > 4 && (size_t)CH * ((PB + 31) / 32 + 2) * 4 * 2 > 24 * 1024) CH >>= 1;
    return CH;
}

// The fused kernel covers lane = agent worlds whose observation is built in one chunk (every training configuration);
// anything else (N > 32, FOV 31 with 32 agents, ...) is served by step_kernel + observe_kernel back to back.
bool step_observe_fusable(const EnvView &v) {
    if (v.N > 32 || obs_chunk(v) < v.N) return false;
    const ObsLayout L = make_layout(v.HP, v.RW, v.GS, v.N, v.C, v.F, obs_chunk(v));
    return (L.total + 32 * 5 * 4 + 32 * 4 + 5 * 32 + QRING) * WARPS_PER_BLOCK <= 200 * 1024;
}

cudaError_t launch_step_observe(const EnvView &v, const int8_t *actions, const MapfStepOut &out, float *obs, float *vec,
                                int *work_counter, cudaStream_t stream, int out_bf16) {
    const int PB = v.C * v.F * v.F;
    const int CH = obs_chunk(v);
    const bool vec4 = ((size_t)v.N * PB) % (out_bf16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
    ObsLayout L = make_layout(v.HP, v.RW, v.GS, v.N, v.C, v.F, CH);
    L.out_bf16 = out_bf16;
    if (v.C == 6 && v.F == 9) {
        return vec4 ? launch_t<6, 9, true>(v, actions, out, obs, vec, L, work_counter, stream)
                    : launch_t<6, 9, false>(v, actions, out, obs, vec, L, work_counter, stream);
    }
    return vec4 ? launch_t<0, 0, true>(v, actions, out, obs, vec, L, work_counter, stream)
                : launch_t<0, 0, false>(v, actions, out, obs, vec, L, work_counter, stream);
}

}  // namespace mapf
