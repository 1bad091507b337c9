// bfs.cu — BFS distance-to-goal maps: bit-parallel frontier wavefront, one warp per map.
//
// Replaces MapfGym.makeBfsMap (mapf_gym.py:211-244): 4-connected BFS from the agent's goal over free cells;
// -1 obstacle, -2 unreached, >= 0 distance (the goal cell is written 0 even when it is not free, as the reference
// does).  The reference computes one map per agent at reset and on every goal arrival (:183, :627).
//
// Layout: the world's rows are bit masks held in registers; G lanes share one map (G = 8 for H <= 40, so a warp runs
// four maps at once), R consecutive rows per lane, NW 32-bit words per row.  One BFS level is
//     next = ((f << 1) | (f >> 1) | f_row_above | f_row_below) & free & ~visited
// with the rows above/below a lane's block fetched by warp shuffles.  Newly reached cells get the level written into
// an int16 tile in shared memory; when the wavefront dies the tile (H*Wd*2 bytes, 3200 B for 40x40) leaves with ONE
// TMA bulk store (cp.async.bulk.global.shared::cta) issued by lane 0, so the warp spends no store instructions on
// the output.  Algorithmic HBM traffic per map: 2*H*Wd B written, the world's obstacle bit rows read (shared by its N maps).
#include "common.cuh"

namespace mapf {

namespace {

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_g2s_commit_wait(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// G lanes cooperate on one map (32/G maps per warp), R consecutive rows per lane, NW 32-bit words per row.
template <int G, int R, int NW>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
bfs_kernel(const EnvView v, const int32_t *__restrict__ agent_list, const long long n_maps_in,
           const int32_t *__restrict__ n_dev, int16_t *__restrict__ out, const int tile_bytes, const int use_tma,
           const int scatter, int *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int MPW = 32 / G;                                   // maps per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;                      // map slot within the warp, lane within the map
    const int H = v.H, Wd = v.Wd;
    int16_t *tile = reinterpret_cast<int16_t *>(smem_raw + ((size_t)warp * MPW + grp) * tile_bytes);
    const int cells = H * Wd;
    const long long n_maps = n_dev ? (long long)*n_dev : n_maps_in;   // device-side count: no host sync for refreshes

    for (;;) {
        const long long base = claim_work(work_counter, MPW, lane);
        if (base >= n_maps) break;
        const long long m = base + grp;
        const bool valid = m < n_maps;
        const long long fid = valid ? (agent_list ? (long long)agent_list[m] : m) : 0;
        const int w = (int)(fid / v.N);
        const uint32_t gw = reinterpret_cast<const uint32_t *>(v.goal)[fid];
        const int goal_r = valid ? (int16_t)(gw & 0xffff) : -1, goal_c = (int16_t)(gw >> 16);

        // tile := -1 everywhere (obstacles keep it; reached cells get their level; the rest -2 at the end)
        {
            uint32_t *t32 = reinterpret_cast<uint32_t *>(tile);
            for (int k = gl; k < (cells + 1) / 2; k += G) t32[k] = 0xffffffffu;
        }
        uint32_t freeb[R][NW], vis[R][NW], fr[R][NW];
        const uint32_t *ob = v.obst_pack + (size_t)w * v.PW;          // packed bit matrix, bit r*Wd + c
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int r = gl * R + k;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                uint32_t x = 0;
                const int nb = Wd - 32 * q;
                if (r < H && valid && nb > 0) {
                    const int n = nb < 32 ? nb : 32, off = r * Wd + 32 * q, kk = off >> 5;
                    const uint32_t lo = __ldg(ob + kk), hi = (kk + 1 < v.PW) ? __ldg(ob + kk + 1) : 0u;
                    x = ~__funnelshift_r(lo, hi, off & 31);
                    if (n < 32) x &= (1u << n) - 1u;
                }
                freeb[k][q] = x;
                const bool here = (r == goal_r) && ((goal_c >> 5) == q);
                fr[k][q] = here ? (1u << (goal_c & 31)) : 0u;
                vis[k][q] = fr[k][q];
            }
        }
        __syncwarp();
        for (int level = 0;; ++level) {
            // write the level of the current frontier
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int r = gl * R + k;
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    uint32_t b = fr[k][q];
                    while (b) { const int c = __ffs(b) - 1; b &= b - 1; tile[r * Wd + 32 * q + c] = (int16_t)level; }
                }
            }
            // rows adjacent to this lane's block (lanes of the same map only)
            uint32_t up[NW], dn[NW];
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                up[q] = __shfl_up_sync(FULL, fr[R - 1][q], 1, G);
                dn[q] = __shfl_down_sync(FULL, fr[0][q], 1, G);
                if (gl == 0) up[q] = 0;
                if (gl == G - 1) dn[q] = 0;
            }
            uint32_t nx[R][NW];
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    uint32_t x = (fr[k][q] << 1) | (fr[k][q] >> 1);
                    if (q > 0) x |= fr[k][q - 1] >> 31;
                    if (q + 1 < NW) x |= fr[k][q + 1] << 31;
                    x |= (k > 0) ? fr[k - 1][q] : up[q];
                    x |= (k + 1 < R) ? fr[k + 1][q] : dn[q];
                    x &= freeb[k][q] & ~vis[k][q];
                    nx[k][q] = x;
                    any |= x;
                }
            }
#pragma unroll
            for (int k = 0; k < R; ++k)
#pragma unroll
                for (int q = 0; q < NW; ++q) { fr[k][q] = nx[k][q]; vis[k][q] |= nx[k][q]; }
            if (!__any_sync(FULL, any != 0)) break;
        }
        // free cells the wavefront never reached: -2
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int r = gl * R + k;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                uint32_t b = freeb[k][q] & ~vis[k][q];
                while (b) { const int c = __ffs(b) - 1; b &= b - 1; tile[r * Wd + 32 * q + c] = (int16_t)-2; }
            }
        }
        int16_t *dst = out + (size_t)(scatter ? fid : m) * cells;
        if (use_tma) {
            fence_proxy_async_smem();
            __syncwarp();
            if (gl == 0 && valid) bulk_store_g2s_commit_wait(dst, tile, (uint32_t)(cells * 2));
            __syncwarp();
        } else {
            __syncwarp();
            if (valid) for (int k = gl; k < cells; k += G) dst[k] = tile[k];
            __syncwarp();
        }
    }
    finish_work(work_counter, gridDim.x * (blockDim.x >> 5), lane);
}

// Compaction of arrivals: flat ids (w*N+i) of agents with goals_reached == 1.  Order is irrelevant (each refreshed
// map is written to its own slot), so one atomicAdd per warp on a device counter is enough.
__global__ void arrivals_kernel(const uint8_t *__restrict__ goals, const long long n, int32_t *__restrict__ list,
                                int32_t *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    for (long long start = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; start < n;
         start += (long long)gridDim.x * blockDim.x) {
        const long long i = start + lane;
        const bool f = i < n && goals[i] != 0;
        const unsigned bal = __ballot_sync(FULL, f);
        if (bal == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(count, __popc(bal));
        base = __shfl_sync(FULL, base, 0);
        if (f) list[base + __popc(bal & ((1u << lane) - 1u))] = (int32_t)i;
    }
}

template <int G, int R, int NW>
cudaError_t launch_bfs_t(const EnvView &v, const int32_t *agent_list, long long n, const int32_t *n_dev, int16_t *out,
                         int scatter, int *work_counter, cudaStream_t stream) {
    constexpr int MPW = 32 / G;
    const int cells = v.H * v.Wd;
    const int tile_bytes = ((cells * 2 + 127) / 128) * 128;
    const int use_tma = ((cells * 2) % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int wpb = WARPS_PER_BLOCK;
    while (wpb > 1 && (size_t)tile_bytes * MPW * wpb > 110 * 1024) wpb >>= 1;
    const size_t smem = (size_t)tile_bytes * MPW * wpb;
    cudaError_t e = cudaFuncSetAttribute(bfs_kernel<G, R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bfs_kernel<G, R, NW>, wpb * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const long long need = (n + (long long)wpb * MPW - 1) / ((long long)wpb * MPW);
    const int blocks = (int)(need < (long long)sms * per_sm ? need : (long long)sms * per_sm);
    if (blocks <= 0) return cudaSuccess;
    bfs_kernel<G, R, NW><<<blocks, wpb * 32, smem, stream>>>(v, agent_list, n, n_dev, out, tile_bytes, use_tma, scatter, work_counter);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_bfs(const EnvView &v, const int32_t *agent_list, long long n, const int32_t *n_dev, int16_t *out,
                       int scatter, int *work_counter, cudaStream_t stream) {
    // lanes per map: the smallest of 8/16/32 that keeps <= 5 rows per lane AND at most ~6.5 KB of int16 tiles per warp,
    // so that 32 warps fit an SM.  The kernel is issue/latency-bound (serial wavefront), and occupancy buys more than
    // the extra maps per warp: 40x40 runs 1.31x faster with 16 lanes per map than with 8, 80x80 1.7x faster with 32.
    const int tile_b = ((v.H * v.Wd * 2 + 127) / 128) * 128;
    int G = 8;
    while (G < 32 && ((v.H + G - 1) / G > 5 || (32 / G) * tile_b > 6656)) G *= 2;
    const int R = (v.H + G - 1) / G, NW = (v.Wd + 31) / 32;
#define CASE(g, r, q) if (G == g && R == r && NW == q) return launch_bfs_t<g, r, q>(v, agent_list, n, n_dev, out, scatter, work_counter, stream);
#define CASES_NW(g, r) CASE(g, r, 1) CASE(g, r, 2) CASE(g, r, 3) CASE(g, r, 4)
    CASES_NW(8, 1) CASES_NW(8, 2) CASES_NW(8, 3) CASES_NW(8, 4) CASES_NW(8, 5)
    CASES_NW(16, 1) CASES_NW(16, 2) CASES_NW(16, 3) CASES_NW(16, 4) CASES_NW(16, 5)
    CASES_NW(32, 1) CASES_NW(32, 2) CASES_NW(32, 3) CASES_NW(32, 4)
#undef CASES_NW
#undef CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_arrivals(const EnvView &v, const uint8_t *goals, int32_t *list, int32_t *n_dev, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(n_dev, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) return e;
    const long long n = (long long)v.W * v.N;
    const int blocks = (int)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
    arrivals_kernel<<<blocks, 256, 0, stream>>>(goals, n, list, n_dev);
    return cudaGetLastError();
}

}  // namespace mapf
