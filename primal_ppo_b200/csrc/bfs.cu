// bfs.cu — BFS distance-to-goal maps, bit-parallel frontier wavefront.
//
// Replaces MapfGym.makeBfsMap (mapf_gym.py:211-244): 4-connected BFS from the agent's goal over free cells;
// -1 obstacle, -2 unreached, >= 0 distance (the goal cell is written 0 even when it is not free, as the reference
// does).  The reference computes one map per agent at reset and on every goal arrival (:183, :627).
//
// Two formulations live here:
//   bfs_gray_kernel  (shipped path)  the map as one cell string, Gray-coded level planes, no per-cell level stores;
//                                    see the comment above the kernel.
//   bfs_kernel       (fallback)      rows as words, an int16 tile in shared memory, per-bit level writes, one TMA bulk
//                                    store per map; takes the maps that cannot leave as 16-byte vectors.
//
// bfs_kernel layout: the world's rows are bit masks held in registers; G lanes share one map, R consecutive rows per
// lane, NW 32-bit words per row.  One BFS level is
//     next = ((f << 1) | (f >> 1) | f_row_above | f_row_below) & free & ~visited
// with the rows above/below a lane's block fetched by warp shuffles.  Newly reached cells get the level written into
// an int16 tile in shared memory; when the wavefront dies the tile (H*Wd*2 bytes, 3200 B for 40x40) leaves with ONE
// TMA bulk store (cp.async.bulk.global.shared::cta) issued by lane 0.
// Algorithmic HBM traffic per map (both kernels): 2*H*Wd B written, the world's obstacle bit matrix read (shared by its N maps).
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

// prmt.b32 in its default mode: bit 3 of a selector nibble replicates the sign of the selected byte (__byte_perm only
// documents the low three bits)
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
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

// ---- second formulation: cell-string layout, Gray-coded level planes, no per-cell stores in the level loop ----------
//
// ncu of the kernel above (40x40, `profiles/r01_ncu_full_raw_final_bfs.csv`): 60 % of its warp instructions are the
// per-bit `tile[cell] = level` loops (19 % / 10 % of the lanes active), 35 % the wavefront update.  This kernel never
// writes a level per cell.  Layout: a map is ONE bit string, bit i = cell i = r*Wd + c (exactly the layout of
// `obst_pack`, so the free mask is a plain load); G lanes share a map, each lane owns NWL consecutive 32-bit words.
// Column neighbours are the string shifted by 1 (masked at row ends), row neighbours the string shifted by Wd
// = 32*WO + sb: funnel shifts over the lane's words plus the WO+1 words of each neighbouring lane (shuffles).
// Levels are kept bit-sliced: plane b of a cell = bit b of the Gray code of its distance.  All unvisited cells carry
// the code of the current level; advancing the level flips ONE plane (b = ctz(level+1)) for the unvisited cells
// (NWL load-xor-stores in shared memory per level), a visited cell simply stops being flipped.  When the wavefront
// dies: Gray -> binary by a suffix XOR over the planes, obstacles forced to 0xffff (-1) and unreached cells to 0xfffe
// (-2) in plane space, a 16x16 bit-matrix transpose per word turns 16 planes x 32 cells into 32 int16, and the lane
// stores its 64-byte run with 16-byte vector stores straight from registers (no tile, no TMA store).
template <int NWL, int WO>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
bfs_gray_kernel(const EnvView v, const int32_t *__restrict__ agent_list, const long long n_maps_in,
                const int32_t *__restrict__ n_dev, int16_t *__restrict__ out, const int G, const int NB,
                const int scatter, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) uint32_t planes_all[];
    static_assert(NWL >= WO + 1, "a row neighbour must live in the lane itself or the adjacent lane");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MPW = 32 / G, grp = lane / G, gl = lane % G;
    const int Wd = v.Wd, cells = v.H * v.Wd, sb = Wd & 31;
    // plane b, word j of this lane: pl[(b * NWL + j) * 32]; planes NB and NB+1 stash the free / unreached masks
    uint32_t *pl = planes_all + (size_t)warp * (NB + 2) * NWL * 32 + lane;
    const long long n_maps = n_dev ? (long long)*n_dev : n_maps_in;
    const int cell0 = gl * NWL * 32;                                  // first cell of this lane's words
    // The level loop is bound by the ALU pipe (shifts and logic ops issue at half rate; ncu: ALU pipe 90 %, FMA pipe 8 %
    // busy), so the operations that have an exact integer multiply-add form are moved to the idle FMA pipe one for one:
    //   fm &= ~nw   ->  fm = nw * 0xffffffff + fm     (nw is a subset of fm: subtracting clears exactly those bits)
    //   edge-lane selects  ->  multiplication by 0 / 1
    // The multipliers live in opaque registers so that the compiler does not turn them back into logic ops.  Measured
    // (profiles/r02_bfs_variants_ab.txt): 40x40 -6 %, 80x80 -13 %, 128x128 -8 %.  Two-for-one replacements lose: shifts as
    // hi * 2^s + mulhi(lo, 2^s) were 3 % slower at 40x40, the any-flag as a mad.wide sum 5 % slower.
    uint32_t minus1 = 0xffffffffu, keep_prev = gl != 0 ? 1u : 0u, keep_next = gl != G - 1 ? 1u : 0u;
    asm volatile("" : "+r"(minus1), "+r"(keep_prev), "+r"(keep_next));
    // masks of the cells that are not in column 0 / not in column Wd-1 (the +-1 shifts must not wrap between rows)
    uint32_t nc0[NWL], ncl[NWL];
#pragma unroll
    for (int j = 0; j < NWL; ++j) {
        const int i0 = cell0 + 32 * j, c0 = i0 % Wd;
        uint32_t a = 0, b = 0;
        for (int t = (Wd - c0) % Wd; t < 32; t += Wd) a |= 1u << t;              // column 0
        for (int t = (Wd - 1 - c0 + Wd) % Wd; t < 32; t += Wd) b |= 1u << t;     // column Wd-1
        nc0[j] = ~a; ncl[j] = ~b;
    }

    for (;;) {
        const long long base = claim_work(work_counter, MPW, lane);
        if (base >= n_maps) break;
        const long long m = base + grp;
        const bool valid = m < n_maps;
        const long long fid = valid ? (agent_list ? (long long)agent_list[m] : m) : 0;
        const int w = (int)(fid / v.N);
        const uint32_t gw = reinterpret_cast<const uint32_t *>(v.goal)[fid];
        const int gi = valid ? (int)(int16_t)(gw & 0xffff) * Wd + (int)(int16_t)(gw >> 16) : -1;   // goal cell
        const uint32_t *ob = v.obst_pack + (size_t)w * v.PW;

        uint32_t freeb[NWL], fm[NWL], fr[NWL];
#pragma unroll
        for (int j = 0; j < NWL; ++j) {
            const int k = gl * NWL + j, i0 = cell0 + 32 * j;
            uint32_t x = 0;
            if (valid && k < v.PW && i0 < cells) {
                x = ~__ldg(ob + k);
                if (cells - i0 < 32) x &= (1u << (cells - i0)) - 1u;
            }
            const uint32_t g = (gi >= i0 && gi < i0 + 32) ? (1u << (gi - i0)) : 0u;
            fr[j] = g;
            freeb[j] = x | g;            // the goal cell gets its 0 even when it is not free (mapf_gym.py:216)
            fm[j] = x & ~g;              // free and not visited yet
        }
        int level = 0;                   // distance of the current frontier
        for (;;) {
            // words of the neighbouring lanes of the same map
            uint32_t pv[WO + 1], nx_[WO + 1];                         // pv[t] = word NWL-1-t of lane-1, nx_[t] = word t of lane+1
#pragma unroll
            for (int t = 0; t <= WO; ++t) {
                pv[t] = __shfl_up_sync(FULL, fr[NWL - 1 - t], 1, G);
                nx_[t] = __shfl_down_sync(FULL, fr[t], 1, G);
                pv[t] *= keep_prev;                                   // lane 0 / G-1 of a map have no neighbour: x * 0 (FMA pipe)
                nx_[t] *= keep_next;
            }
            auto word = [&](int idx) -> uint32_t {                    // idx is a compile-time constant after unrolling
                if (idx >= 0 && idx < NWL) return fr[idx];
                if (idx < 0 && -idx - 1 <= WO) return pv[-idx - 1];
                if (idx >= NWL && idx - NWL <= WO) return nx_[idx - NWL];
                return 0u;
            };
            uint32_t nw[NWL], any = 0;
#pragma unroll
            for (int j = 0; j < NWL; ++j) {
                const uint32_t left = __funnelshift_l(word(j - 1), fr[j], 1) & nc0[j];        // from cell i-1
                const uint32_t right = __funnelshift_r(fr[j], word(j + 1), 1) & ncl[j];       // from cell i+1
                const uint32_t up = __funnelshift_l(word(j - WO - 1), word(j - WO), sb);      // from cell i-Wd
                const uint32_t down = __funnelshift_r(word(j + WO), word(j + WO + 1), sb);    // from cell i+Wd
                nw[j] = (left | right | up | down) & fm[j];
                any |= nw[j];
            }
            if (!__any_sync(FULL, any != 0)) break;
            // level -> level+1: one Gray plane flips for every cell that was unvisited (the new frontier included)
            const int nl = level + 1, b = __ffs(nl) - 1;
            uint32_t *pb = pl + (size_t)b * NWL * 32;
            if (nl == (1 << b)) {                                     // first touch of this plane: it was all zero
#pragma unroll
                for (int j = 0; j < NWL; ++j) pb[j * 32] = fm[j];
            } else {
#pragma unroll
                for (int j = 0; j < NWL; ++j) pb[j * 32] ^= fm[j];
            }
#pragma unroll
            for (int j = 0; j < NWL; ++j) {
                fr[j] = nw[j];
                fm[j] = nw[j] * minus1 + fm[j];                       // = fm & ~nw (nw is a subset of fm), on the FMA pipe
            }
            level = nl;
        }
        // planes in use: those of codes up to `level` (warp-uniform: the loop runs until the slowest map is done)
        const int nb = 32 - __clz(level);
#pragma unroll
        for (int j = 0; j < NWL; ++j) { pl[((size_t)NB * NWL + j) * 32] = freeb[j]; pl[((size_t)(NB + 1) * NWL + j) * 32] = fm[j]; }
        int16_t *dst = out + (size_t)(scatter ? fid : m) * cells;
#pragma unroll 1
        for (int j = 0; j < NWL; ++j) {
            const int i0 = cell0 + 32 * j;
            if (!valid || i0 >= cells) continue;
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + i0);
            const int nvec = (cells - i0 >= 32) ? 4 : (cells - i0) >> 3;              // cells % 8 == 0 (launcher)
            const uint32_t ob_ = ~pl[((size_t)NB * NWL + j) * 32], un = pl[((size_t)(NB + 1) * NWL + j) * 32];
            if (nb <= 7) {
                // levels < 128: 8 planes, an 8x8 bit transpose of the four bytes at once, then sign-extending byte
                // permutes widen level / 0xff (-1) / 0xfe (-2) to int16
                uint32_t Q[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) Q[q] = q < nb ? pl[((size_t)q * NWL + j) * 32] : 0u;
#pragma unroll
                for (int q = 6; q >= 0; --q) Q[q] ^= Q[q + 1];                        // Gray -> binary
                Q[0] = (Q[0] | ob_) & ~un;
#pragma unroll
                for (int q = 1; q < 8; ++q) Q[q] |= ob_ | un;
#define MAPF_TSTAGE8(S, MK)                                                     \
                _Pragma("unroll") for (int q = 0; q < 8; ++q) {                     \
                    if ((q & S) == 0) {                                             \
                        const uint32_t t = ((Q[q] >> S) ^ Q[q + S]) & MK;           \
                        Q[q + S] ^= t;                                              \
                        Q[q] ^= t << S;                                             \
                    }                                                               \
                }
                MAPF_TSTAGE8(4, 0x0f0f0f0fu) MAPF_TSTAGE8(2, 0x33333333u) MAPF_TSTAGE8(1, 0x55555555u)
#undef MAPF_TSTAGE8
                // now byte h of Q[q] = level of cell 8h + q
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (h >= nvec) continue;
                    const uint32_t sel = (uint32_t)h | ((uint32_t)(h | 8) << 4) | ((uint32_t)(4 + h) << 8) | ((uint32_t)((4 + h) | 8) << 12);
                    uint4 o;
                    o.x = prmt_sx(Q[0], Q[1], sel);
                    o.y = prmt_sx(Q[2], Q[3], sel);
                    o.z = prmt_sx(Q[4], Q[5], sel);
                    o.w = prmt_sx(Q[6], Q[7], sel);
                    d4[h] = o;
                }
                continue;
            }
            uint32_t P[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) P[q] = q < nb ? pl[((size_t)q * NWL + j) * 32] : 0u;
#pragma unroll
            for (int q = 14; q >= 0; --q) P[q] ^= P[q + 1];                           // Gray -> binary
            P[0] = (P[0] | ob_) & ~un;                                                // -1 = 0xffff, -2 = 0xfffe
#pragma unroll
            for (int q = 1; q < 16; ++q) P[q] |= ob_ | un;
            // 16x16 bit transpose of both halves at once: afterwards P[q] = level(cell q) | level(cell 16+q) << 16
#define MAPF_TSTAGE(S, MK)                                                      \
            _Pragma("unroll") for (int q = 0; q < 16; ++q) {                        \
                if ((q & S) == 0) {                                                 \
                    const uint32_t t = ((P[q] >> S) ^ P[q + S]) & MK;               \
                    P[q + S] ^= t;                                                  \
                    P[q] ^= t << S;                                                 \
                }                                                                   \
            }
            MAPF_TSTAGE(8, 0x00ff00ffu) MAPF_TSTAGE(4, 0x0f0f0f0fu) MAPF_TSTAGE(2, 0x33333333u) MAPF_TSTAGE(1, 0x55555555u)
#undef MAPF_TSTAGE
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h >= nvec) continue;
                const int q = (h & 1) * 8;                                            // cells 8h .. 8h+7
                const uint32_t sel = (h < 2) ? 0x5410u : 0x7632u;                     // low halves: cells 0..15; high: 16..31
                uint4 o;
                o.x = __byte_perm(P[q + 0], P[q + 1], sel);
                o.y = __byte_perm(P[q + 2], P[q + 3], sel);
                o.z = __byte_perm(P[q + 4], P[q + 5], sel);
                o.w = __byte_perm(P[q + 6], P[q + 7], sel);
                d4[h] = o;
            }
        }
        __syncwarp();
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

template <int NWL, int WO>
cudaError_t launch_bfs_gray_t(const EnvView &v, const int32_t *agent_list, long long n, const int32_t *n_dev, int16_t *out,
                              int G, int scatter, int *work_counter, cudaStream_t stream) {
    const int cells = v.H * v.Wd, MPW = 32 / G;
    int NB = 1;
    while ((1 << NB) <= cells) ++NB;                                  // levels < cells < 2^NB
    const size_t per_warp = (size_t)(NB + 2) * NWL * 128;
    cudaError_t e = cudaFuncSetAttribute(bfs_gray_kernel<NWL, WO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(per_warp * WARPS_PER_BLOCK > 200 * 1024 ? 200 * 1024 : per_warp * WARPS_PER_BLOCK));
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the planes are the occupancy limit: take the CTA size that keeps the most warps resident
    int wpb = 1, per_sm = 1, best = 0;
    for (int cand = WARPS_PER_BLOCK; cand >= 1; --cand) {
        if (per_warp * cand > 200 * 1024) continue;
        int k = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, bfs_gray_kernel<NWL, WO>, cand * 32, per_warp * cand) != cudaSuccess) continue;
        if (k * cand > best) { best = k * cand; wpb = cand; per_sm = k; }
    }
    if (best == 0) return cudaErrorInvalidConfiguration;
    const size_t smem = per_warp * wpb;
    const long long need = (n + (long long)wpb * MPW - 1) / ((long long)wpb * MPW);
    const int blocks = (int)(need < (long long)sms * per_sm ? need : (long long)sms * per_sm);
    if (blocks <= 0) return cudaSuccess;
    bfs_gray_kernel<NWL, WO><<<blocks, wpb * 32, smem, stream>>>(v, agent_list, n, n_dev, out, G, NB, scatter, work_counter);
    return cudaGetLastError();
}

// Lanes per map and words per lane of the cell-string kernel: the smallest G of 8/16/32 with <= 8 words per lane
// (measured at 40x40: 8 lanes x 7 words 0.90 ms, 16 x 4 1.07 ms, 32 x 2 1.20 ms for 262 144 maps)
// (G = 32 takes whatever is left, up to 16 words for 128x128), words rounded up to an instantiated count.
bool bfs_gray_shape(const EnvView &v, int &G, int &NWL, int &WO) {
    const int cells = v.H * v.Wd;
    WO = v.Wd >> 5;
    const int force = (v.dbg_flags >> 17) & 3;                        // MAPF_DBG_FLAGS bits 17..18: force G = 8/16/32
    G = 8;
    while (G < 32 && (cells + 32 * G - 1) / (32 * G) > 8) G *= 2;
    if (force) G = 4 << force;
    NWL = (cells + 32 * G - 1) / (32 * G);
    if (NWL < WO + 1) NWL = WO + 1;
    if (NWL > 8) NWL = NWL <= 10 ? 10 : NWL <= 12 ? 12 : 16;
    return NWL * 32 * G >= cells && NWL <= 16;
}

}  // namespace

cudaError_t launch_bfs(const EnvView &v, const int32_t *agent_list, long long n, const int32_t *n_dev, int16_t *out,
                       int scatter, int *work_counter, cudaStream_t stream) {
    // lanes per map: the smallest of 8/16/32 that keeps <= 5 rows per lane AND at most ~6.5 KB of int16 tiles per warp,
    // so that 32 warps fit an SM.  The kernel is issue/latency-bound (serial wavefront), and occupancy buys more than
    // the extra maps per warp: 40x40 runs 1.31x faster with 16 lanes per map than with 8, 80x80 1.7x faster with 32.
    // cell-string kernel whenever the maps can leave as 16-byte vectors (H*Wd % 8 == 0, aligned output);
    // MAPF_DBG_FLAGS bit 16 keeps the row-word kernel for A/B runs
    {
        int G, NWL, WO;
        const bool vec_ok = ((v.H * v.Wd) % 8 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        if (vec_ok && !(v.dbg_flags & (1 << 16)) && bfs_gray_shape(v, G, NWL, WO)) {
            // row shifts as integer multiply-adds (FMA pipe) unless Wd is a multiple of 32; MAPF_DBG_FLAGS bit 21 = funnel shifts
#define GCASE(nwl, wo) if (NWL == nwl && WO == wo) return launch_bfs_gray_t<nwl, wo>(v, agent_list, n, n_dev, out, G, scatter, work_counter, stream);
#define GCASES0(nwl) GCASE(nwl, 0)
#define GCASES1(nwl) GCASES0(nwl) GCASE(nwl, 1)
#define GCASES2(nwl) GCASES1(nwl) GCASE(nwl, 2)
#define GCASES3(nwl) GCASES2(nwl) GCASE(nwl, 3)
#define GCASES4(nwl) GCASES3(nwl) GCASE(nwl, 4)
            GCASES0(1) GCASES1(2) GCASES2(3) GCASES3(4) GCASES4(5) GCASES4(6) GCASES4(7) GCASES4(8) GCASES4(10) GCASES4(12) GCASES4(16)
#undef GCASES4
#undef GCASES3
#undef GCASES2
#undef GCASES1
#undef GCASES0
#undef GCASE
        }
    }
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
