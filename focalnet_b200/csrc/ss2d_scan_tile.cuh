// ss2d_scan_tile.cuh — cp.async staging of a group's B/C rows, shared by the scan forward and backward.
#pragma once
#include "ss2d_common.cuh"

namespace ss2d {

template <typename in_t, int T, int SB> struct BCTile {
    using RL = RowLayout<in_t, T>;
    static constexpr int chunk = kWarp * T;
    static constexpr int pieces_per_row = chunk / RL::per;
    static constexpr int tile_bytes = 2 * SB * RL::row_bytes;  // SB rows of B then SB rows of C
};

// Stage the B and C rows of states [n0, n0+SB) for timesteps [t0, t0+chunk) into `buf`
// (lane-blocked padded rows, see RowLayout).  Steps past L are zero-filled so that a masked
// (identity) element never multiplies smem garbage.
template <typename in_t, int T, int SB, int NTHREADS>
__device__ __forceinline__ void stage_bc(unsigned char *buf, const in_t *__restrict__ Bg, const in_t *__restrict__ Cg,
                                         int64_t B_nstride, int64_t C_nstride, int n0, int N, int64_t t0, int64_t L,
                                         bool vec) {
    using FT = BCTile<in_t, T, SB>;
    using RL = typename FT::RL;
    if (vec) {
        constexpr int total = 2 * SB * FT::pieces_per_row;
        for (int idx = threadIdx.x; idx < total; idx += NTHREADS) {
            const int row = idx / FT::pieces_per_row, q = idx % FT::pieces_per_row;
            const int r = row % SB, n = n0 + r;
            if (n >= N) continue;
            const in_t *base = row < SB ? Bg + (int64_t)n * B_nstride : Cg + (int64_t)n * C_nstride;
            const int64_t t = t0 + (int64_t)q * RL::per;
            const int64_t rem = (L - t) * (int64_t)sizeof(in_t);
            const int bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
            cp_async16(buf + row * RL::row_bytes + RL::unit_of_piece(q) * 16, bytes > 0 ? base + t : base, bytes);
        }
    } else {  // unaligned rows: element-wise, zero-filled tail
        constexpr int total = 2 * SB * FT::chunk;
        for (int idx = threadIdx.x; idx < total; idx += NTHREADS) {
            const int row = idx / FT::chunk, e = idx % FT::chunk;
            const int r = row % SB, n = n0 + r;
            if (n >= N) continue;
            const in_t *base = row < SB ? Bg + (int64_t)n * B_nstride : Cg + (int64_t)n * C_nstride;
            const int64_t t = t0 + e;
            in_t v = t < L ? base[t] : from_f32<in_t>(0.f);
            reinterpret_cast<in_t *>(buf + row * RL::row_bytes)[RL::unit_of_piece(e / RL::per) * RL::per + e % RL::per] = v;
        }
    }
}

}  // namespace ss2d
