// observe.cu — observation builder: one warp per world, persistent grid-stride loop over worlds.
//
// Replaces MapfGym.getAllObservations / observe / worldWithAgents (mapf_gym.py:192-198, 246-336) for W worlds:
//   obs f32 [W,N,C,F,F]  (0/1 valued)   and   vec f32 [W,N,4] = [dx/d, dy/d, d, 0]
// written straight into the policy's input tensors.  The store of obs is ~95 % of all bytes of a step, so the
// kernel is organised around it:
//   phase 1  lane = agent: build the agent's C*F*F observation as a BIT string in shared memory.  Channels 0/1 are
//            F-bit windows cut out of padded obstacle / agent bit rows with one funnel shift per row (no per-cell
//            work); channels 2-5 are a handful of single bits (own goal, clamped goals of visible agents, human).
//   phase 1b compact the per-agent bit strings into one contiguous bit string of the chunk (N*C*F*F bits).
//   phase 2  all 32 lanes expand bits to floats: 4 bits -> one 16-byte streaming store, 512 contiguous bytes per
//            warp instruction, world after world (each world's block N*C*F*F*4 B is contiguous in HBM).
// Algorithmic HBM traffic per world: write N*(4*C*F*F + 16) B, read N*8 B state + HP*RW*4 B obstacle bits + 8 B human.
#include "common.cuh"

namespace mapf {

namespace {

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct ObsLayout {
    int PB;      // bits per agent = C*F*F
    int AST;     // u32 stride of one agent's padded bit string (odd -> conflict-free lane-strided access)
    int CH;      // agents per chunk (<= 32)
    int WB;      // u32 words of the chunk bit string
    size_t off_abits, off_grid, off_goal, off_aw, off_wb, total;
};

__host__ __device__ inline ObsLayout make_layout(int HP, int RW, int GS, int N, int C, int F, int CH) {
    ObsLayout L;
    L.PB = C * F * F;
    int aw = (L.PB + 31) / 32 + 1;
    if ((aw & 1) == 0) aw++;
    L.AST = aw;
    L.CH = CH;
    L.WB = (CH * L.PB + 31) / 32 + 2;
    size_t o = align16((size_t)HP * RW * 4);
    L.off_abits = o; o += align16((size_t)HP * RW * 4);
    L.off_grid = o; o += align16((size_t)HP * GS);
    L.off_goal = o; o += align16((size_t)N * 4);
    L.off_aw = o; o += align16((size_t)CH * L.AST * 4);
    L.off_wb = o; o += align16((size_t)L.WB * 4);
    L.total = o;
    return L;
}

__device__ __forceinline__ void or_bits(uint32_t *words, int p, uint32_t val, int nbits) {
    const int k = p >> 5, s = p & 31;
    words[k] |= val << s;
    if (s + nbits > 32) words[k + 1] |= val >> (32 - s);
}
__device__ __forceinline__ void or_bit(uint32_t *words, int p) { words[p >> 5] |= 1u << (p & 31); }

template <bool VEC4>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
observe_kernel(const EnvView v, float *__restrict__ obs, float *__restrict__ vec, const ObsLayout L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = v.N, P = v.P, GS = v.GS, RW = v.RW, HP = v.HP, F = v.F, C = v.C, half = v.F >> 1;
    const int FF = F * F, PB = L.PB, AST = L.AST, CH = L.CH;
    unsigned char *base = smem_raw + (size_t)warp * L.total;
    uint32_t *obits = reinterpret_cast<uint32_t *>(base);
    uint32_t *abits = reinterpret_cast<uint32_t *>(base + L.off_abits);
    uint8_t *grid = base + L.off_grid;
    uint32_t *sgoal = reinterpret_cast<uint32_t *>(base + L.off_goal);
    uint32_t *aw = reinterpret_cast<uint32_t *>(base + L.off_aw);
    uint32_t *wb = reinterpret_cast<uint32_t *>(base + L.off_wb);

    // one-time clear of the agent bit rows and the id grid; every world un-scatters what it scattered
    for (int k = lane; k < HP * RW; k += 32) abits[k] = 0;
    {
        uint4 *g4 = reinterpret_cast<uint4 *>(grid);
        for (int k = lane; k < (HP * GS) / 16; k += 32) g4[k] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    const int wpb = blockDim.x >> 5;
    for (int w = blockIdx.x * wpb + warp; w < v.W; w += gridDim.x * wpb) {
        // ---- stage: obstacle bit rows, agents (ids into the byte grid, presence into bit rows), goals -----------
        {
            const uint32_t *src = v.obst_bits + (size_t)w * HP * RW;
            for (int k = lane; k < HP * RW; k += 32) obits[k] = __ldg(src + k);
        }
        const uint32_t *posw = reinterpret_cast<const uint32_t *>(v.pos) + (size_t)w * N;
        const uint32_t *goalw = reinterpret_cast<const uint32_t *>(v.goal) + (size_t)w * N;
        for (int i = lane; i < N; i += 32) {
            const uint32_t pw = __ldg(posw + i);
            const int r = (int16_t)(pw & 0xffff), c = (int16_t)(pw >> 16);
            grid[(r + P) * GS + c + P] = (uint8_t)(i + 1);
            atomicOr(&abits[(r + P) * RW + ((c + P) >> 5)], 1u << ((c + P) & 31));
            sgoal[i] = __ldg(goalw + i);
        }
        const int tick = v.htick[w];
        const int2 ht = *reinterpret_cast<const int2 *>(v.htrace + ((size_t)w * v.L + tick) * 4);
        const int nr = (int16_t)(ht.y & 0xffff), nc = (int16_t)((uint32_t)ht.y >> 16);   // human.getNextPos()
        const int rows = v.dims ? v.dims[2 * w] : v.H, cols = v.dims ? v.dims[2 * w + 1] : v.Wd;
        __syncwarp();

        for (int c0 = 0; c0 < N; c0 += CH) {
            const int nch = min(CH, N - c0);
            const int i = c0 + lane;
            const bool act = lane < nch;
            // ---- phase 1: per-agent bit strings -------------------------------------------------------------
            if (act) {
                uint32_t *my = aw + lane * AST;
                for (int k = 0; k < AST; ++k) my[k] = 0;
                const uint32_t pw = __ldg(posw + i), gw = sgoal[i];
                const int r = (int16_t)(pw & 0xffff), c = (int16_t)(pw >> 16);
                const int gr = (int16_t)(gw & 0xffff), gc = (int16_t)(gw >> 16);
                const int top = r - half, left = c - half;                                    // :251
                for (int y = 0; y < F; ++y) {
                    const int prow = top + y + P, off = left + P;
                    uint32_t o = row_window(obits + prow * RW, off, F);      // OOB or obstacle  (:270-276)
                    uint32_t g = row_window(abits + prow * RW, off, F);      // agents           (:278-285)
                    if (y == half) { o |= 1u << half; g &= ~(1u << half); }  // own cell goes to channel 0 (:278-280)
                    or_bits(my, y * F, o, F);
                    or_bits(my, FF + y * F, g, F);
                    while (g) {                                              // visible agents' goals, clamped (:302-308)
                        const int x = __ffs(g) - 1; g &= g - 1;
                        const int j = grid[prow * GS + off + x] - 1;
                        const uint32_t jw = sgoal[j];
                        const int jr = (int16_t)(jw & 0xffff), jc = (int16_t)(jw >> 16);
                        const int mr = max(top, min(top + F - 1, jr)), mc = max(left, min(left + F - 1, jc));
                        or_bit(my, 3 * FF + (mr - top) * F + (mc - left));
                    }
                    if (v.use_da) {                                          // danger disc |cell - H'| <= 5 (:289-290)
                        const int rr = top + y, dy = rr > nr ? rr - nr : nr - rr;
                        if (rr >= 0 && rr < rows && dy <= 5) {
                            const int hw = dy == 0 ? 5 : dy <= 3 ? 4 : dy == 4 ? 3 : 0;
                            const int lo = max(max(nc - hw, 0), left), hi = min(min(nc + hw, cols - 1), left + F - 1);
                            if (lo <= hi) or_bits(my, 4 * FF + y * F + (lo - left), (1u << (hi - lo + 1)) - 1u, hi - lo + 1);
                        }
                    }
                }
                if (gr >= top && gr < top + F && gc >= left && gc < left + F)                  // own goal (:298-300)
                    or_bit(my, 2 * FF + (gr - top) * F + (gc - left));
                if (nr >= top && nr < top + F && nc >= left && nc < left + F)                  // human (:310-312)
                    or_bit(my, 4 * FF + (nr - top) * F + (nc - left));
                if (v.use_hp && C == 6 && v.hp5) {                                             // (:293-297)
                    const int16_t *p5 = v.hp5 + (v.hp5_per_tick ? ((size_t)w * v.L + tick) * 10 : (size_t)w * 10);
                    for (int k = 0; k < 5; ++k) {
                        const int pr = p5[2 * k], pc = p5[2 * k + 1];
                        if (pr >= 0 && pr < rows && pc >= 0 && pc < cols && pr >= top && pr < top + F && pc >= left && pc < left + F)
                            or_bit(my, 5 * FF + (pr - top) * F + (pc - left));
                    }
                }
                // vector (:316-323): f64 sqrt / divide, then cast
                const double dx = (double)(gr - r), dy_ = (double)(gc - c);
                const double d = sqrt(dx * dx + dy_ * dy_);
                float4 o4;
                o4.x = (float)(d != 0.0 ? dx / d : dx);
                o4.y = (float)(d != 0.0 ? dy_ / d : dy_);
                o4.z = (float)d;
                o4.w = 0.0f;
                reinterpret_cast<float4 *>(vec)[(size_t)w * N + i] = o4;
            }
            __syncwarp();
            // ---- phase 1b: compact to one contiguous bit string ------------------------------------------------
            const int TB = nch * PB;
            const int nwords = (TB + 31) >> 5;
            for (int m = lane; m < nwords; m += 32) {
                const int b0 = m << 5;
                const int n = b0 / PB, e = b0 - n * PB;
                const uint32_t *src = aw + n * AST;
                uint32_t x = __funnelshift_r(src[e >> 5], src[(e >> 5) + 1], e & 31);
                const int valid = PB - e;
                if (valid < 32) {
                    x &= (1u << valid) - 1u;
                    if (n + 1 < nch) x |= aw[(n + 1) * AST] << valid;
                }
                wb[m] = x;
            }
            __syncwarp();
            // ---- phase 2: bits -> floats, streaming stores -----------------------------------------------------
            float *dst = obs + ((size_t)w * N + c0) * PB;
            if (VEC4) {
                const int n4 = TB >> 2;
                const int sh = (lane & 7) << 2;
                const uint32_t *wp = wb + (lane >> 3);
                float *d4 = dst + (lane << 2);
                for (int q = lane; q < n4; q += 32, wp += 4, d4 += 128) {
                    const uint32_t nib = *wp >> sh;
                    st_stream_v4(d4, (nib & 1u) ? 0x3f800000u : 0u, (nib & 2u) ? 0x3f800000u : 0u,
                                 (nib & 4u) ? 0x3f800000u : 0u, (nib & 8u) ? 0x3f800000u : 0u);
                }
            } else {
                for (int f = lane; f < TB; f += 32) dst[f] = ((wb[f >> 5] >> (f & 31)) & 1u) ? 1.0f : 0.0f;
            }
            __syncwarp();
        }
        // ---- un-scatter this world's agents so the next world starts from a clean grid ----------------------------
        for (int i = lane; i < N; i += 32) {
            const uint32_t pw = __ldg(posw + i);
            const int r = (int16_t)(pw & 0xffff), c = (int16_t)(pw >> 16);
            grid[(r + P) * GS + c + P] = 0;
            abits[(r + P) * RW + ((c + P) >> 5)] = 0;
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_observe(const EnvView &v, float *obs, float *vec, cudaStream_t stream) {
    const int PB = v.C * v.F * v.F;
    // chunk of agents handled per phase-1 pass: as many as fit ~24 KB of bit-string scratch per warp
    int CH = v.N < 32 ? v.N : 32;
    while (CH > 4 && (size_t)CH * ((PB + 31) / 32 + 2) * 4 * 2 > 24 * 1024) CH >>= 1;
    const bool vec4 = ((size_t)v.N * PB) % 4 == 0 && (CH == v.N || ((size_t)CH * PB) % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
    const ObsLayout L = make_layout(v.HP, v.RW, v.GS, v.N, v.C, v.F, CH);
    int wpb = WARPS_PER_BLOCK;
    while (wpb > 1 && L.total * wpb > 200 * 1024) wpb >>= 1;
    if (L.total * wpb > 227 * 1024) return cudaErrorInvalidConfiguration;
    const size_t smem = L.total * wpb;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e;
    int per_sm = 1;
    if (vec4) {
        e = cudaFuncSetAttribute(observe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, observe_kernel<true>, wpb * 32, smem);
    } else {
        e = cudaFuncSetAttribute(observe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, observe_kernel<false>, wpb * 32, smem);
    }
    if (per_sm < 1) per_sm = 1;
    const int need = (v.W + wpb - 1) / wpb;
    const int blocks = need < sms * per_sm ? need : sms * per_sm;
    if (vec4) observe_kernel<true><<<blocks, wpb * 32, smem, stream>>>(v, obs, vec, L);
    else observe_kernel<false><<<blocks, wpb * 32, smem, stream>>>(v, obs, vec, L);
    return cudaGetLastError();
}

}  // namespace mapf
