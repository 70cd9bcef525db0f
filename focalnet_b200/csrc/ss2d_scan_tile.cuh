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
        // Every thread copies the SAME 16-byte column of rows r0, r0+rpp, ... : the source/destination
        // addresses differ by compile-time multiples, and the tail clamp is computed once per tile.
        static_assert(NTHREADS % FT::pieces_per_row == 0, "threads must tile the row pieces");
        constexpr int rpp = NTHREADS / FT::pieces_per_row;          // rows covered per pass
        static_assert(SB % rpp == 0, "state block must be a multiple of the rows per pass");
        const int q = threadIdx.x % FT::pieces_per_row, r0 = threadIdx.x / FT::pieces_per_row;
        const int64_t t = t0 + (int64_t)q * RL::per;
        const int64_t rem = (L - t) * (int64_t)sizeof(in_t);
        const int bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        const int64_t toff = bytes > 0 ? t : 0;
        unsigned char *dst = buf + r0 * RL::row_bytes + RL::unit_of_piece(q) * 16;
        const in_t *srcB = Bg + (int64_t)(n0 + r0) * B_nstride + toff;
        const in_t *srcC = Cg + (int64_t)(n0 + r0) * C_nstride + toff;
#pragma unroll
        for (int k = 0; k < SB / rpp; ++k) {
            if (n0 + r0 + k * rpp < N) {
                cp_async16(dst + (k * rpp) * RL::row_bytes, srcB + (int64_t)(k * rpp) * B_nstride, bytes);
                cp_async16(dst + (SB + k * rpp) * RL::row_bytes, srcC + (int64_t)(k * rpp) * C_nstride, bytes);
            }
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
