// step_common.cuh — pieces shared by the two joint-step kernels (step.cu: N <= 32, step_wide.cu: N <= 128).
#pragma once
#include "common.cuh"

namespace mapf {

constexpr int C_INV0 = 0, C_INV1 = 1, C_GOOD = 2, C_E = 3;   // class of an agent's chosen action (getActionStatus if-chain)
constexpr int FIX_CAP = 256;  // fixActions iteration cap (reference: unbounded while loop, mapf_gym.py:563); same in the oracle

// Scan the radius-2 diamond around (gr, gc) [padded grid coordinates] for other agents.
//   restr  : bit a set iff another agent lies within Manhattan distance 1 of T(a)
//   confl  : bit a set iff some neighbour j with b = other_act[j] >= 0 has conflict(i,a; j,b)
//   selmask: agents j with conflict(i, a_sel; j, other_act[j]) — a bitmask (PACK = false, N <= 32) or up to four
//            agent ids packed one per byte, 0xFF = none (PACK = true, any N <= 254)
template <bool PACK>
__device__ __forceinline__ void scan_diamond(const uint8_t *grid, int GS, int gr, int gc, int self_code,
                                             const int8_t *other_act, int a_sel, uint32_t &restr, uint32_t &confl,
                                             uint32_t &selmask) {
    restr = 0; confl = 0; selmask = PACK ? 0xffffffffu : 0u;
#pragma unroll
    for (int dr = -2; dr <= 2; ++dr) {
#pragma unroll
        for (int dc = -2; dc <= 2; ++dc) {
            const int md = (dr < 0 ? -dr : dr) + (dc < 0 ? -dc : dc);
            if (md == 0 || md > 2) continue;
            const int code = grid[(gr + dr) * GS + (gc + dc)];
            if (code == 0 || code == self_code) continue;
            const int j = code - 1;
            const int b = other_act[j];
            const int tr = dr + (b >= 0 ? dr_of(b) : 0), tc = dc + (b >= 0 ? dc_of(b) : 0);  // T_j - pos_i
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const int ar = (a == 2) - (a == 4), ac = (a == 1) - (a == 3);
                const int d1 = (dr - ar < 0 ? ar - dr : dr - ar) + (dc - ac < 0 ? ac - dc : dc - ac);
                if (d1 > 1) continue;                       // compile-time: j not adjacent to T_i(a)
                restr |= 1u << a;
                bool c = (b >= 0) && (tr == ar) && (tc == ac);                                    // vertex
                if (a != 0 && dr == ar && dc == ac) c = c || (b == (a == 1 ? 3 : a == 2 ? 4 : a == 3 ? 1 : 2));  // swap
                if (c) {
                    confl |= 1u << a;
                    if (a == a_sel) selmask = PACK ? ((selmask << 8) | (uint32_t)j) : (selmask | (1u << j));
                }
            }
        }
    }
}


}  // namespace mapf
