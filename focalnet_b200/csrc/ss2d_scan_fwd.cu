// ss2d_scan_fwd.cu — selective-scan forward for sm_100a (seam S1, scan-order operands).
//
// Replaces selective_scan_fwd_kernel + its launcher/host code
// (reference: kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_fwd_kernel_oflex.cuh:67-211,
//  selective_scan_oflex.cpp:157-243).  Same maths:
//     dl_t = softplus(delta_t + bias_c)            a_{t,n} = exp2(dl_t * A_{c,n} * log2 e)
//     h_{t,n} = a_{t,n} h_{t-1,n} + dl_t u_t B_{n,t}        out_t = D_c u_t + sum_n C_{n,t} h_{t,n}
// Different organisation (see ss2d_common.cuh): one warp per channel, CTA = NW channels of one
// (batch, group); B/C rows of the group staged once per CTA by cp.async in a 2-deep ping-pong over
// blocks of SB states; no block-wide scan, one __syncthreads per state block.
//
// HBM traffic per launch (the figure bench.py uses):  s_in*(2*B*Dm*L + 2*B*G*N*L) + s_out*B*Dm*L
//   + 4*B*Dm*(ceil(L/2048)*2N + ceil(L/256)*N)   [checkpoints]
#include "ss2d_common.cuh"
#include "ss2d_scan_tile.cuh"
#include "../../include/ss2d_b200.h"
#include <cstdlib>
#include <cstring>

namespace ss2d {

struct FwdFlags {
    bool vec_u, vec_delta, vec_bc, vec_out, vec_z;
};

// BATCH states of one chunk for this lane's T steps, processed together so that the BATCH warp scans
// (5 dependent shuffle rounds each) are in flight at the same time — the scan's shuffle latency is what
// bounds this kernel, not its instruction count.
//   pass 1 keeps the running product and the local state of every step (Pcum_i, hloc_i) in the registers
//   that held a_i and b_i, so pass 2 (h_i = Pcum_i * h_in + hloc_i) has no serial dependency.
#ifndef SS2D_KNOCK
#define SS2D_KNOCK 0  // diagnosis only: 1 = no shuffles, 2 = no MUFU, 3 = no B/C LDS, 4 = no pass-1/2 chains
#endif
template <typename in_t, int T, int SLOTS, int BATCH>
__device__ __forceinline__ void fwd_states(const unsigned char *Brow, const unsigned char *Crow, int lane,
                                           const float (&dl)[T], const float (&du)[T], float (&y)[T],
                                           const float *sA2, float *ck /* [SLOTS] strided by ck_stride */,
                                           int ck_stride, float *sP) {
    using RL = RowLayout<in_t, T>;
    float a[BATCH][T], hl[BATCH][T];
#pragma unroll
    for (int s = 0; s < BATCH; ++s) {
        float Bv[T];
#if SS2D_KNOCK == 3 || SS2D_KNOCK == 8
#pragma unroll
        for (int i = 0; i < T; ++i) Bv[i] = dl[i] + s;
#else
        lds_block<in_t, T>(Brow + s * RL::row_bytes, lane, Bv);
#endif
        const float A2 = sA2[s];
#pragma unroll
        for (int i = 0; i < T; ++i) {
#if SS2D_KNOCK == 2 || SS2D_KNOCK == 8
            a[s][i] = fmaf(dl[i], A2, 1.f);
#else
            a[s][i] = ex2(dl[i] * A2);
#endif
            hl[s][i] = du[i] * Bv[i];
        }
    }
#pragma unroll
    for (int i = 1; i < T; ++i) {
#pragma unroll
        for (int s = 0; s < BATCH; ++s) { hl[s][i] = fmaf(a[s][i], hl[s][i - 1], hl[s][i]); a[s][i] *= a[s][i - 1]; }
    }
    float P[BATCH], H[BATCH];
#pragma unroll
    for (int s = 0; s < BATCH; ++s) { P[s] = a[s][T - 1]; H[s] = hl[s][T - 1]; }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float Pp[BATCH], Hp[BATCH];
#pragma unroll
        for (int s = 0; s < BATCH; ++s) {
#if SS2D_KNOCK == 1 || SS2D_KNOCK == 8
            Pp[s] = P[s] * 0.5f; Hp[s] = H[s] + 1.f;
#else
            Pp[s] = __shfl_up_sync(0xffffffffu, P[s], d);
            Hp[s] = __shfl_up_sync(0xffffffffu, H[s], d);
#endif
        }
        if (lane >= d) {
#pragma unroll
            for (int s = 0; s < BATCH; ++s) { H[s] = fmaf(P[s], Hp[s], H[s]); P[s] *= Pp[s]; }
        }
    }
    float hin[BATCH];
#pragma unroll
    for (int s = 0; s < BATCH; ++s) {
        float Pe = __shfl_up_sync(0xffffffffu, P[s], 1), He = __shfl_up_sync(0xffffffffu, H[s], 1);
        if (lane == 0) { Pe = 1.f; He = 0.f; }
        hin[s] = fmaf(Pe, ck[(SLOTS - 1) * ck_stride + s], He);  // carry = h at the end of the previous chunk
    }
    constexpr int lanes_per_slot = kWarp / SLOTS;
#pragma unroll
    for (int s = 0; s < BATCH; ++s) {
        float Cv[T];
#if SS2D_KNOCK == 3 || SS2D_KNOCK == 8
#pragma unroll
        for (int i = 0; i < T; ++i) Cv[i] = du[i] + s;
#else
        lds_block<in_t, T>(Crow + s * RL::row_bytes, lane, Cv);
#endif
        float h = hin[s];
#pragma unroll
        for (int i = 0; i < T; ++i) { h = fmaf(a[s][i], hin[s], hl[s][i]); y[i] = fmaf(Cv[i], h, y[i]); }
        // lanes whose block ends on a checkpoint boundary publish h there (the last slot is the chunk carry)
        if ((lane + 1) % lanes_per_slot == 0) ck[((lane + 1) / lanes_per_slot - 1) * ck_stride + s] = h;
        if (lane == kWarp - 1) sP[s] *= P[s];
    }
}

// ---- software-pipelined state block --------------------------------------------------------------------
// The warp scan of state r (5 dependent shuffle rounds, ~200 cycles of pure latency) is interleaved, in
// program order, with the exp / pass-1 work of state r+1, so a warp always has independent instructions
// to issue while its shuffles are in flight.  (The warps of a CTA run in lock-step between barriers, so
// leaving the overlap to other warps does not work: they all sit in the same shuffle round.)
template <typename in_t, int T> struct StatePrep {
    float a[T], hl[T];  // after pass 1: running product of a / local state at every step
    // slice 0..4 of the preparation of one state; each slice is independent of the scan in flight
    template <int SLICE>
    __device__ __forceinline__ void run(const unsigned char *Brow, int lane, const float (&dl)[T], const float (&du)[T],
                                        float A2) {
        constexpr int H1 = T / 2;
        if constexpr (SLICE == 0 || SLICE == 1) {  // exps and drive terms, half the block each
            constexpr int lo = SLICE * H1;
            float Bv[H1];
            using RL = RowLayout<in_t, T>;
            const uint4 *src = reinterpret_cast<const uint4 *>(Brow) + RL::lane_unit(lane) + lo / RL::per;
#pragma unroll
            for (int i = 0; i < H1 / RL::per; ++i) unpack16<in_t>(src[i], &Bv[i * RL::per]);
#pragma unroll
            for (int i = 0; i < H1; ++i) { a[lo + i] = ex2(dl[lo + i] * A2); hl[lo + i] = du[lo + i] * Bv[i]; }
        } else {  // pass 1 in three slices
            constexpr int third = (T - 1 + 2) / 3;
            constexpr int lo = 1 + (SLICE - 2) * third, hi = (lo + third < T) ? lo + third : T;
#pragma unroll
            for (int i = lo; i < hi; ++i) { hl[i] = fmaf(a[i], hl[i - 1], hl[i]); a[i] *= a[i - 1]; }
        }
    }
};

template <typename in_t, int T, int SLOTS, int SB>
__device__ __forceinline__ void fwd_block_pipelined(const unsigned char *buf, int lane, const float (&dl)[T],
                                                    const float (&du)[T], float (&y)[T], const float *sA2, float *ck,
                                                    int ck_stride, float *sP) {
    using RL = RowLayout<in_t, T>;
    static_assert(T % (2 * RL::per) == 0, "half a lane block must be whole 16-byte pieces");
    StatePrep<in_t, T> st[2];
    {
        const float A2 = sA2[0];
        st[0].template run<0>(buf, lane, dl, du, A2);
        st[0].template run<1>(buf, lane, dl, du, A2);
        st[0].template run<2>(buf, lane, dl, du, A2);
        st[0].template run<3>(buf, lane, dl, du, A2);
        st[0].template run<4>(buf, lane, dl, du, A2);
    }
    constexpr int lanes_per_slot = kWarp / SLOTS;
#pragma unroll
    for (int r = 0; r < SB; ++r) {
        StatePrep<in_t, T> &cur = st[r & 1], &nxt = st[(r + 1) & 1];
        const unsigned char *Bnext = buf + (r + 1) * RL::row_bytes;
        const float A2n = r + 1 < SB ? sA2[r + 1] : 0.f;
        float P = cur.a[T - 1], H = cur.hl[T - 1];
#define SS2D_SCAN_STAGE(D, SLICE)                                                   \
        {                                                                           \
            const float Pp = __shfl_up_sync(0xffffffffu, P, D);                     \
            const float Hp = __shfl_up_sync(0xffffffffu, H, D);                     \
            if (r + 1 < SB) nxt.template run<SLICE>(Bnext, lane, dl, du, A2n);      \
            if (lane >= D) { H = fmaf(P, Hp, H); P *= Pp; }                         \
        }
        SS2D_SCAN_STAGE(1, 0)
        SS2D_SCAN_STAGE(2, 1)
        SS2D_SCAN_STAGE(4, 2)
        SS2D_SCAN_STAGE(8, 3)
        SS2D_SCAN_STAGE(16, 4)
#undef SS2D_SCAN_STAGE
        float Pe = __shfl_up_sync(0xffffffffu, P, 1), He = __shfl_up_sync(0xffffffffu, H, 1);
        float Cv[T];
        lds_block<in_t, T>(buf + (SB + r) * RL::row_bytes, lane, Cv);
        if (lane == 0) { Pe = 1.f; He = 0.f; }
        const float hin = fmaf(Pe, ck[(SLOTS - 1) * ck_stride + r], He);
        float h = hin;
#pragma unroll
        for (int i = 0; i < T; ++i) { h = fmaf(cur.a[i], hin, cur.hl[i]); y[i] = fmaf(Cv[i], h, y[i]); }
        if ((lane + 1) % lanes_per_slot == 0) ck[((lane + 1) / lanes_per_slot - 1) * ck_stride + r] = h;
        if (lane == kWarp - 1) sP[r] *= P;
    }
}

// CROSS: fused seam S3.  The groups are the 4 scan directions, `u` is the spatial-order plane x[b, d] shared by
// the 4 directions (u_dstride indexes d, not k*D+d) and `out` is the merged spatial-order fp32 plane y[b, d],
// accumulated with red.global.add (zero-filled by the caller).  delta / B / C stay in scan order.
template <typename in_t, typename out_t, int T, int NW, int SB, int BATCH, int MINB, bool CROSS>
__global__ void __launch_bounds__(NW * kWarp, MINB)
scan_fwd_kernel(const ss2d_scan_fwd_params p, const int tiles_per_group, const FwdFlags fl, const CrossInfo xinfo) {
    using FT = BCTile<in_t, T, SB>;
    using RL = typename FT::RL;
    constexpr int chunk = FT::chunk;
    constexpr int NT = NW * kWarp;
    constexpr int SLOTS = chunk / SS2D_CKPT_STEPS;  // fine checkpoints per chunk
    static_assert(chunk % SS2D_CKPT_STEPS == 0 && SS2D_REF_CHUNK % chunk == 0, "chunk must tile the checkpoint grids");
    extern __shared__ __align__(16) unsigned char smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = (int)p.dstate;
    const int Npad = (N + 3) & ~3;
    const int64_t L = p.seqlen;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tile = blockIdx.x % tiles_per_group;
    const int bg = blockIdx.x / tiles_per_group;
    const int g = bg % (int)p.ngroups, b = bg / (int)p.ngroups;
    const int c_local = tile * NW + warp;
    const bool active = c_local < per_g;
    const int64_t c = (int64_t)g * per_g + (active ? c_local : per_g - 1);

    unsigned char *tiles = smem;
    float *sA2 = reinterpret_cast<float *>(smem + 2 * FT::tile_bytes) + warp * Npad;  // A * log2(e)
    float *sP = reinterpret_cast<float *>(smem + 2 * FT::tile_bytes) + NW * Npad + warp * Npad;  // running prod a (for x)
    float *sCk = reinterpret_cast<float *>(smem + 2 * FT::tile_bytes) + 2 * NW * Npad + warp * SLOTS * Npad;  // [SLOTS][Npad]

    const int64_t cu = CROSS ? (active ? c_local : per_g - 1) : c;  // row of u / out: d in fused mode
    const in_t *u_row = reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + cu * p.u_dstride;
    const in_t *d_row = reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + c * p.delta_dstride;
    const in_t *z_row = p.z ? reinterpret_cast<const in_t *>(p.z) + b * p.z_bstride + c * p.z_dstride : nullptr;
    const in_t *Bg = reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride;
    const in_t *Cg = reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride;
    out_t *o_row = p.out ? reinterpret_cast<out_t *>(p.out) + b * p.out_bstride + cu * p.out_dstride : nullptr;
    out_t *oz_row = p.out_z ? reinterpret_cast<out_t *>(p.out_z) + b * p.out_bstride + c * p.out_dstride : nullptr;
    const float Dv = p.D ? p.D[c] : 0.f;
    const float bias = p.delta_bias ? p.delta_bias[c] : 0.f;

    for (int n = lane; n < N; n += kWarp) {
        sA2[n] = p.A[c * N + n] * kLog2e;
        sP[n] = 1.f;
        sCk[(SLOTS - 1) * Npad + n] = 0.f;
    }

    const int n_sb = (N + SB - 1) / SB;
    const int n_chunks = (int)((L + chunk - 1) / chunk);
    const int Q = n_chunks * n_sb;
    const int n_fine = (int)((L + SS2D_CKPT_STEPS - 1) / SS2D_CKPT_STEPS);
    const int n_ref = (int)((L + SS2D_REF_CHUNK - 1) / SS2D_REF_CHUNK);

    stage_bc<in_t, T, SB, NT>(tiles, Bg, Cg, p.B_nstride, p.C_nstride, 0, N, 0, L, fl.vec_bc);
    cp_async_commit();

    float dl[T], du[T], y[T];
    for (int q = 0; q < Q; ++q) {
        const int ci = q / n_sb, sb = q % n_sb;
        const int64_t t0 = (int64_t)ci * chunk;
        cp_async_wait<0>();
        __syncthreads();  // tile q visible to everyone; everyone is done with the buffer tile q+1 will overwrite
        if (q + 1 < Q) {
            const int ci1 = (q + 1) / n_sb, sb1 = (q + 1) % n_sb;
            stage_bc<in_t, T, SB, NT>(tiles + ((q + 1) & 1) * FT::tile_bytes, Bg, Cg, p.B_nstride, p.C_nstride,
                                      sb1 * SB, N, (int64_t)ci1 * chunk, L, fl.vec_bc);
            cp_async_commit();
        }
        const int64_t tl = t0 + lane * T;                 // first timestep of this lane's block
        const int valid = (int)min((int64_t)T, L - tl);   // may be <= 0
        if (sb == 0) {
            float uv[T];
#if SS2D_KNOCK == 6
#pragma unroll
            for (int i = 0; i < T; ++i) { uv[i] = 0.01f * (lane + i); dl[i] = 0.02f * (lane - i) + ci; }
#else
            if constexpr (CROSS) load_block_cross<in_t, T>(u_row, uv, g, tl, L, xinfo, fl.vec_u);
            else load_block<in_t, T>(u_row + tl, uv, valid, fl.vec_u);
            load_block<in_t, T>(d_row + tl, dl, valid, fl.vec_delta);
#endif
            if (tl + chunk < L) {  // pull the next chunk's u / delta lines into L2 while this chunk computes
                if constexpr (!CROSS) prefetch_l2(u_row + tl + chunk);
                prefetch_l2(d_row + tl + chunk);
            }
#pragma unroll
            for (int i = 0; i < T; ++i) {
                float d = dl[i] + bias;
                if (p.delta_softplus) d = softplus_ref(d);
                d = i < valid ? d : 0.f;  // identity element past the end (fwd_kernel_oflex.cuh:146-150)
                dl[i] = d;
                du[i] = d * uv[i];
                y[i] = Dv * uv[i];
            }
        }
        const unsigned char *buf = tiles + (q & 1) * FT::tile_bytes;
#if SS2D_KNOCK != 5
        {
            const int n_here = min(SB, N - sb * SB);
            int r = 0;
            if (n_here == SB && BATCH == 0) {
                fwd_block_pipelined<in_t, T, SLOTS, SB>(buf, lane, dl, du, y, sA2 + sb * SB, sCk + sb * SB, Npad,
                                                        sP + sb * SB);
                r = SB;
            } else if (n_here == SB) {
                constexpr int BT = BATCH > 0 ? BATCH : 1;
#pragma unroll
                for (int rr = 0; rr < SB; rr += BT)
                    fwd_states<in_t, T, SLOTS, BT>(buf + rr * RL::row_bytes, buf + (SB + rr) * RL::row_bytes, lane, dl,
                                                      du, y, sA2 + sb * SB + rr, sCk + sb * SB + rr, Npad,
                                                      sP + sb * SB + rr);
                r = SB;
            }
#pragma unroll 1
            for (; r < n_here; ++r)
                fwd_states<in_t, T, SLOTS, 1>(buf + r * RL::row_bytes, buf + (SB + r) * RL::row_bytes, lane, dl, du, y,
                                              sA2 + sb * SB + r, sCk + sb * SB + r, Npad, sP + sb * SB + r);
        }
#endif
        if (sb == n_sb - 1) {  // chunk finished
            __syncwarp();
            if (active) {
                if constexpr (CROSS) {
                    red_block_cross<T>(reinterpret_cast<float *>(o_row), y, g, tl, L, xinfo, fl.vec_out);
                } else if (o_row) {
#if SS2D_KNOCK == 7
                    if (y[0] == 123.4f)
#endif
                    store_block<out_t, T>(o_row + tl, y, valid, fl.vec_out);
                    if (z_row) {
                        float zv[T];
                        load_block<in_t, T>(z_row + tl, zv, valid, fl.vec_z);
#pragma unroll
                        for (int i = 0; i < T; ++i) y[i] *= zv[i] * sigmoidf_fast(zv[i]);
                        store_block<out_t, T>(oz_row + tl, y, valid, fl.vec_out);
                    }
                }
                const int64_t t_end = min(L, t0 + chunk);
                const bool last = ci == n_chunks - 1;
                if (p.ckpt) {
                    float *dst = p.ckpt + ((int64_t)b * p.dim + c) * n_fine * N;
#pragma unroll
                    for (int sl = 0; sl < SLOTS; ++sl) {
                        const int f = ci * SLOTS + sl;
                        if (f < n_fine)
                            for (int n = lane; n < N; n += kWarp) dst[(int64_t)f * N + n] = sCk[sl * Npad + n];
                    }
                }
                if (p.x && (t_end % SS2D_REF_CHUNK == 0 || last)) {
                    float2 *dst = reinterpret_cast<float2 *>(p.x) +
                                  (((int64_t)b * p.dim + c) * n_ref + (t_end - 1) / SS2D_REF_CHUNK) * N;
                    for (int n = lane; n < N; n += kWarp) dst[n] = make_float2(sP[n], sCk[(SLOTS - 1) * Npad + n]);
                }
            }
            __syncwarp();
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename in_t, typename out_t, int T = 16, int NW = 8, int BATCH = 1, int MINB = 2, int SB = 8, bool CROSS = false>
static int launch_fwd(const ss2d_scan_fwd_params &p, cudaStream_t stream, CrossInfo ci = CrossInfo{0, 0}) {
    using FT = BCTile<in_t, T, SB>;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tiles = (per_g + NW - 1) / NW;
    const int Npad = ((int)p.dstate + 3) & ~3;
    const size_t smem = 2 * FT::tile_bytes + (2 + FT::chunk / SS2D_CKPT_STEPS) * NW * Npad * sizeof(float);
    const int64_t ei = sizeof(in_t), eo = sizeof(out_t);
    FwdFlags fl;
    fl.vec_u = aligned16(p.u) && (p.u_bstride * ei) % 16 == 0 && (p.u_dstride * ei) % 16 == 0;
    fl.vec_delta = aligned16(p.delta) && (p.delta_bstride * ei) % 16 == 0 && (p.delta_dstride * ei) % 16 == 0;
    fl.vec_bc = aligned16(p.B) && aligned16(p.C) && (p.B_bstride * ei) % 16 == 0 && (p.B_gstride * ei) % 16 == 0 &&
                (p.B_nstride * ei) % 16 == 0 && (p.C_bstride * ei) % 16 == 0 && (p.C_gstride * ei) % 16 == 0 &&
                (p.C_nstride * ei) % 16 == 0;
    fl.vec_out = aligned16(p.out) && (!p.out_z || aligned16(p.out_z)) && (p.out_bstride * eo) % 16 == 0 &&
                 (p.out_dstride * eo) % 16 == 0;
    fl.vec_z = p.z && aligned16(p.z) && (p.z_bstride * ei) % 16 == 0 && (p.z_dstride * ei) % 16 == 0;
    static_assert(BATCH == 0 || SB % BATCH == 0, "BATCH must divide the state block (0 = software-pipelined)");
    auto kern = scan_fwd_kernel<in_t, out_t, T, NW, SB, BATCH, MINB, CROSS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t grid = p.batch * p.ngroups * tiles;
    kern<<<(unsigned)grid, NW * kWarp, smem, stream>>>(p, tiles, fl, ci);
    return (int)cudaGetLastError();
}

}  // namespace ss2d

extern "C" int ss2d_selective_scan_fwd(const ss2d_scan_fwd_params *pp, void *stream) {
    if (!pp) return SS2D_EINVAL;
    const ss2d_scan_fwd_params &p = *pp;
    if (!p.u || !p.delta || !p.A || !p.B || !p.C) return SS2D_EINVAL;  // out == NULL: states-only sweep (x / ckpt)
    if (p.batch <= 0 || p.dim <= 0 || p.seqlen <= 0 || p.dstate <= 0 || p.ngroups <= 0) return SS2D_EINVAL;
    if (p.dim % p.ngroups != 0 || p.dstate > SS2D_MAX_DSTATE) return SS2D_EINVAL;
    if (p.z && !p.out_z) return SS2D_EINVAL;
    if (p.batch * p.ngroups * (p.dim / p.ngroups) > 0x7fffffffLL) return SS2D_EINVAL;
    if (p.out_dtype != SS2D_F32 && p.out_dtype != p.in_dtype) return SS2D_EDTYPE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    using namespace ss2d;
#ifdef SS2D_TUNE  // development knob: alternative tilings for fp32, selected by SS2D_FWD_CFG=TxNWxBATCHxMINB
    if (p.in_dtype == SS2D_F32) {
        const char *cfg = getenv("SS2D_FWD_CFG");
        if (cfg) {
            // T x NW x BATCH x MINB
            if (!strcmp(cfg, "16x8x1x2")) return launch_fwd<float, float, 16, 8, 1, 2>(p, s);
            if (!strcmp(cfg, "8x8x1x3")) return launch_fwd<float, float, 8, 8, 1, 3>(p, s);
            if (!strcmp(cfg, "8x8x2x3")) return launch_fwd<float, float, 8, 8, 2, 3>(p, s);
            if (!strcmp(cfg, "8x8x2x2")) return launch_fwd<float, float, 8, 8, 2, 2>(p, s);
            if (!strcmp(cfg, "8x8x4x2")) return launch_fwd<float, float, 8, 8, 4, 2>(p, s);
            if (!strcmp(cfg, "8x16x2x1")) return launch_fwd<float, float, 8, 16, 2, 1>(p, s);
            if (!strcmp(cfg, "8x16x4x1")) return launch_fwd<float, float, 8, 16, 4, 1>(p, s);
            if (!strcmp(cfg, "16x8x2x1")) return launch_fwd<float, float, 16, 8, 2, 1>(p, s);
            if (!strcmp(cfg, "16x16x1x1")) return launch_fwd<float, float, 16, 16, 1, 1>(p, s);
            if (!strcmp(cfg, "16x4x1x4s4")) return launch_fwd<float, float, 16, 4, 1, 4, 4>(p, s);
            if (!strcmp(cfg, "16x8x1x2s4")) return launch_fwd<float, float, 16, 8, 1, 2, 4>(p, s);
            if (!strcmp(cfg, "16x4x1x2s8")) return launch_fwd<float, float, 16, 4, 1, 2, 8>(p, s);
            if (!strcmp(cfg, "8x4x1x4s8")) return launch_fwd<float, float, 8, 4, 1, 4, 8>(p, s);
            if (!strcmp(cfg, "8x4x2x4s8")) return launch_fwd<float, float, 8, 4, 2, 4, 8>(p, s);
            if (!strcmp(cfg, "8x4x1x6s4")) return launch_fwd<float, float, 8, 4, 1, 6, 4>(p, s);
            if (!strcmp(cfg, "8x2x1x8s4")) return launch_fwd<float, float, 8, 2, 1, 8, 4>(p, s);
            if (!strcmp(cfg, "8x8x0x3")) return launch_fwd<float, float, 8, 8, 0, 3>(p, s);
            if (!strcmp(cfg, "8x8x0x2")) return launch_fwd<float, float, 8, 8, 0, 2>(p, s);
            if (!strcmp(cfg, "8x16x0x1")) return launch_fwd<float, float, 8, 16, 0, 1>(p, s);
            if (!strcmp(cfg, "16x8x0x1")) return launch_fwd<float, float, 16, 8, 0, 1>(p, s);
            if (!strcmp(cfg, "16x8x0x2")) return launch_fwd<float, float, 16, 8, 0, 2>(p, s);
        }
    }
#endif
    switch (p.in_dtype) {
        case SS2D_F32: return launch_fwd<float, float>(p, s);
        case SS2D_F16:
            return p.out_dtype == SS2D_F32 ? launch_fwd<__half, float>(p, s) : launch_fwd<__half, __half>(p, s);
        case SS2D_BF16:
            return p.out_dtype == SS2D_F32 ? launch_fwd<__nv_bfloat16, float>(p, s)
                                           : launch_fwd<__nv_bfloat16, __nv_bfloat16>(p, s);
        default: return SS2D_EDTYPE;
    }
}

// Fused SS2D core forward (seam S3): see ss2d_cross_fwd_params in include/ss2d_b200.h.
extern "C" int ss2d_cross_scan_fwd(const ss2d_cross_fwd_params *pp, void *stream) {
    if (!pp) return SS2D_EINVAL;
    const ss2d_cross_fwd_params &c = *pp;
    if (!c.x || !c.delta || !c.B || !c.C || !c.A || !c.y) return SS2D_EINVAL;
    if (c.batch <= 0 || c.D <= 0 || c.H <= 0 || c.W <= 0 || c.dstate <= 0 || c.dstate > SS2D_MAX_DSTATE) return SS2D_EINVAL;
    if (c.H * c.W > 0x7fffffffLL || c.batch * 4 * c.D > 0x7fffffffLL) return SS2D_EINVAL;
    const int64_t L = c.H * c.W;
    ss2d_scan_fwd_params p{};
    p.batch = c.batch; p.dim = 4 * c.D; p.seqlen = L; p.dstate = c.dstate; p.ngroups = 4;
    p.in_dtype = c.in_dtype; p.out_dtype = SS2D_F32; p.delta_softplus = c.delta_softplus;
    p.u = c.x; p.delta = c.delta; p.A = c.A; p.B = c.B; p.C = c.C; p.D = c.Dskip; p.delta_bias = c.delta_bias;
    p.u_bstride = c.D * L; p.u_dstride = L;
    p.delta_bstride = 4 * c.D * L; p.delta_dstride = L;
    p.B_bstride = p.C_bstride = c.bc_bstride ? c.bc_bstride : 4 * c.dstate * L;
    p.B_gstride = p.C_gstride = c.bc_gstride ? c.bc_gstride : c.dstate * L;
    p.B_nstride = p.C_nstride = L;
    p.out = c.y; p.out_bstride = c.D * L; p.out_dstride = L;
    p.ckpt = c.ckpt;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    using namespace ss2d;
    const CrossInfo ci{(int)c.H, (int)c.W};
    switch (c.in_dtype) {
        case SS2D_F32: return launch_fwd<float, float, 16, 8, 1, 2, 8, true>(p, s, ci);
        case SS2D_F16: return launch_fwd<__half, float, 16, 8, 1, 2, 8, true>(p, s, ci);
        case SS2D_BF16: return launch_fwd<__nv_bfloat16, float, 16, 8, 1, 2, 8, true>(p, s, ci);
        default: return SS2D_EDTYPE;
    }
}
