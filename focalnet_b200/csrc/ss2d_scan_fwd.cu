// ss2d_scan_fwd.cu — selective-scan forward for sm_100a (seam S1, and the fused seam S3 with CROSS=true).
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
//
// The kernel is bound by instruction issue, so the state loop is written to cost no integer
// instructions: every shared-memory operand is [one base register + compile-time immediate].
#include "ss2d_common.cuh"
#include "ss2d_scan_tile.cuh"
#include "ss2d_scan_sl.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

int plane_transpose(const float *src, float *dst, int64_t planes, int H, int W, bool acc, cudaStream_t stream);  // ss2d_cross.cu

struct FwdFlags {
    bool vec_u, vec_delta, vec_bc, vec_out, vec_z;
};

// One state (index R inside the staged block) of one chunk for this lane's T steps.
//   pass 1 keeps the running product and the local state of every step (Pcum_i, hloc_i) in the registers that
//   held a_i and b_i, so the apply pass h_i = Pcum_i * h_in + hloc_i has no serial dependency.
//   bc : shared address of this lane's block in row 0 of the staged B tile (C rows follow SB rows later)
//   wa : shared address of this warp's per-state arrays at the block's first state:
//        [A*log2e | running prod a | ckpt slot 0 | .. | slot SLOTS-1], NSB bytes apart (last slot = chunk carry)
template <typename in_t, int T, int SLOTS, int SB, int R, int NSB_CT>
__device__ __forceinline__ void fwd_state(uint32_t bc, uint32_t wa, int nsb_rt, int lane, bool pub, uint32_t pub_addr,
                                          const float (&dl)[T], const float (&du)[T], float (&y)[T]) {
    using RL = RowLayout<in_t, T>;
    const float A2 = lds_f32<R * 4>(wa);
    float a[T], hl[T];
    {
        float Bv[T];
        lds_row<in_t, T, R * RL::row_bytes>(bc, Bv);
#pragma unroll
        for (int i = 0; i < T; ++i) { a[i] = ex2(dl[i] * A2); hl[i] = du[i] * Bv[i]; }
    }
#pragma unroll
    for (int i = 1; i < T; ++i) { hl[i] = fmaf(a[i], hl[i - 1], hl[i]); a[i] *= a[i - 1]; }
    float P = a[T - 1], H = hl[T - 1];
    warp_scan_inclusive(P, H, lane);
    float Pe, He;
    shift_up1(P, H, Pe, He);
    const uint32_t carry_addr = wa + (NSB_CT > 0 ? (1 + SLOTS) * NSB_CT : (1 + SLOTS) * nsb_rt);
    const float hin = fmaf(Pe, lds_f32<R * 4>(carry_addr), He);  // carry = h at the end of the previous chunk
    float Cv[T];
    lds_row<in_t, T, (SB + R) * RL::row_bytes>(bc, Cv);
    float h = hin;
#pragma unroll
    for (int i = 0; i < T; ++i) { h = fmaf(a[i], hin, hl[i]); y[i] = fmaf(Cv[i], h, y[i]); }
    // lanes whose block ends on a checkpoint boundary publish h there (the last slot is the chunk carry)
    sts_f32_if<R * 4>(pub, pub_addr, h);
    const uint32_t p_addr = wa + (NSB_CT > 0 ? NSB_CT : nsb_rt);
    sts_f32_if<R * 4>(lane == kWarp - 1, p_addr, lds_f32<R * 4>(p_addr) * P);  // lane 31 holds the chunk's prod a
}

template <typename in_t, int T, int SLOTS, int SB, int NSB_CT, int R = 0>
__device__ __forceinline__ void fwd_block(uint32_t bc, uint32_t wa, int nsb_rt, int lane, bool pub, uint32_t pub_addr,
                                          const float (&dl)[T], const float (&du)[T], float (&y)[T], int n_here) {
    if constexpr (R < SB) {
        if (R < n_here) {
            fwd_state<in_t, T, SLOTS, SB, R, NSB_CT>(bc, wa, nsb_rt, lane, pub, pub_addr, dl, du, y);
            fwd_block<in_t, T, SLOTS, SB, NSB_CT, R + 1>(bc, wa, nsb_rt, lane, pub, pub_addr, dl, du, y, n_here);
        }
    }
}

// CROSS: fused seam S3.  The groups are the 4 scan directions, `u` is the spatial-order plane x[b, d] shared by
// the 4 directions (u_dstride indexes d, not k*D+d) and `out` is the merged spatial-order fp32 plane y[b, d],
// accumulated with red.global.add (zero-filled by the caller).  delta / B / C stay in scan order.
template <typename in_t, typename out_t, int T, int NW, int SB, int MINB, bool CROSS, int NSB_CT>
__global__ void __launch_bounds__(NW * kWarp, MINB)
scan_fwd_kernel(const ss2d_scan_fwd_params p, const int tiles_per_group, const FwdFlags fl, const CrossInfo xinfo) {
    using FT = BCTile<in_t, T, SB>;
    using RL = typename FT::RL;
    constexpr int chunk = FT::chunk;
    constexpr int NT = NW * kWarp;
    constexpr int SLOTS = chunk / SS2D_CKPT_STEPS;  // fine checkpoints per chunk
    constexpr int lanes_per_slot = kWarp / SLOTS;
    static_assert(chunk % SS2D_CKPT_STEPS == 0 && SS2D_REF_CHUNK % chunk == 0, "chunk must tile the checkpoint grids");
    extern __shared__ __align__(16) unsigned char smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = (int)p.dstate;
    const int NS = NSB_CT > 0 ? NSB_CT / 4 : ((N + 3) & ~3);  // floats per per-warp array
    const int nsb_rt = NS * 4;
    const int64_t L = p.seqlen;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tile = blockIdx.x % tiles_per_group;
    const int bg = blockIdx.x / tiles_per_group;
    int g, b;
    cta_group_batch(xinfo, bg, (int)p.ngroups, g, b);
    const int c_local = tile * NW + warp;
    const bool active = c_local < per_g;
    const int64_t c = (int64_t)g * per_g + (active ? c_local : per_g - 1);

    unsigned char *tiles = smem;
    float *wsm = reinterpret_cast<float *>(smem + 2 * FT::tile_bytes) + warp * (2 + SLOTS) * NS;  // this warp's arrays
    float *sA2 = wsm, *sP = wsm + NS, *sCk = wsm + 2 * NS;
    const uint32_t tiles_addr = smem_u32(tiles) + RL::lane_unit(lane) * 16;
    const uint32_t wsm_addr = smem_u32(wsm);
    const bool pub = (lane + 1) % lanes_per_slot == 0;
    const uint32_t pub_base = wsm_addr + (2 + ((lane + 1) / lanes_per_slot - 1)) * nsb_rt;

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
        sCk[(SLOTS - 1) * NS + n] = 0.f;
    }

    const int n_sb = (N + SB - 1) / SB;
    const int n_chunks = (int)((L + chunk - 1) / chunk);
    const int Q = n_chunks * n_sb;
    const int n_fine = (int)((L + SS2D_CKPT_STEPS - 1) / SS2D_CKPT_STEPS);
    const int n_ref = (int)((L + SS2D_REF_CHUNK - 1) / SS2D_REF_CHUNK);

    stage_bc<in_t, T, SB, NT>(tiles, Bg, Cg, p.B_nstride, p.C_nstride, 0, N, 0, L, fl.vec_bc);
    cp_async_commit();

    float dl[T], du[T], y[T];
    int ci = 0, sb = 0;
    for (int q = 0; q < Q; ++q) {
        const int64_t t0 = (int64_t)ci * chunk;
        cp_async_wait<0>();
        __syncthreads();  // tile q visible to everyone; everyone is done with the buffer tile q+1 will overwrite
        const int64_t tl = t0 + lane * T;                 // first timestep of this lane's block
        const int valid = (int)min((int64_t)T, L - tl);   // may be <= 0
        float uv[T];
        if (sb == 0) {  // this chunk's u / delta loads are issued first, ahead of the cp.async burst of the next tile
            if constexpr (CROSS) load_block_cross<in_t, T>(u_row, uv, g, tl, L, xinfo, fl.vec_u);
            else load_block<in_t, T>(u_row + tl, uv, valid, fl.vec_u);
            load_block<in_t, T>(d_row + tl, dl, valid, fl.vec_delta);
            if (tl + chunk < L) {  // pull the next chunk's u / delta lines into L2 while this chunk computes
                if constexpr (!CROSS) prefetch_l2(u_row + tl + chunk);
                prefetch_l2(d_row + tl + chunk);
            }
        }
        if (q + 1 < Q) {
            const int sb1 = sb + 1 == n_sb ? 0 : sb + 1, ci1 = sb + 1 == n_sb ? ci + 1 : ci;
            stage_bc<in_t, T, SB, NT>(tiles + ((q + 1) & 1) * FT::tile_bytes, Bg, Cg, p.B_nstride, p.C_nstride,
                                      sb1 * SB, N, (int64_t)ci1 * chunk, L, fl.vec_bc);
            cp_async_commit();
        }
        if (sb == 0) {
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
        {
            const uint32_t bc = tiles_addr + (q & 1) * FT::tile_bytes;
            const uint32_t wa = wsm_addr + sb * SB * 4;
            fwd_block<in_t, T, SLOTS, SB, NSB_CT>(bc, wa, nsb_rt, lane, pub, pub_base + sb * SB * 4, dl, du, y,
                                                  min(SB, N - sb * SB));
        }
        if (sb == n_sb - 1) {  // chunk finished
            __syncwarp();
            if (active) {
                if constexpr (CROSS) {
                    if (o_row) red_block_cross<T>(reinterpret_cast<float *>(o_row), y, g, tl, L, xinfo, fl.vec_out);
                } else if (o_row) {
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
                            for (int n = lane; n < N; n += kWarp) dst[(int64_t)f * N + n] = sCk[sl * NS + n];
                    }
                }
                if (p.x && (t_end % SS2D_REF_CHUNK == 0 || last)) {
                    float2 *dst = reinterpret_cast<float2 *>(p.x) +
                                  (((int64_t)b * p.dim + c) * n_ref + (t_end - 1) / SS2D_REF_CHUNK) * N;
                    for (int n = lane; n < N; n += kWarp) dst[n] = make_float2(sP[n], sCk[(SLOTS - 1) * NS + n]);
                }
            }
            __syncwarp();
        }
        if (++sb == n_sb) { sb = 0; ++ci; }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename in_t, typename out_t, bool CROSS, int T = 16, int NW = 8, int MINB = 2, int SB = 8>
static int launch_fwd(const ss2d_scan_fwd_params &p, cudaStream_t stream, CrossInfo xinfo = CrossInfo{0, 0, -1}) {
    using FT = BCTile<in_t, T, SB>;
    constexpr int SLOTS = FT::chunk / SS2D_CKPT_STEPS;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tiles = (per_g + NW - 1) / NW;
    const bool small_n = p.dstate <= 16;
    const int NS = small_n ? 16 : (((int)p.dstate + 3) & ~3);
    const size_t smem = 2 * FT::tile_bytes + (size_t)(2 + SLOTS) * NW * NS * sizeof(float);
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
    const int64_t grid = p.batch * (xinfo.g_only >= 0 ? 1 : p.ngroups) * tiles;
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)smem);
        if (rc != 0) return rc;
        kern<<<(unsigned)grid, NW * kWarp, smem, stream>>>(p, tiles, fl, xinfo);
        return (int)cudaGetLastError();
    };
    if (small_n) return go(scan_fwd_kernel<in_t, out_t, T, NW, SB, MINB, CROSS, 64>);
    return go(scan_fwd_kernel<in_t, out_t, T, NW, SB, MINB, CROSS, 0>);
}

template <bool CROSS> static int dispatch_fwd(const ss2d_scan_fwd_params &p, cudaStream_t s, CrossInfo xi) {
    switch (p.in_dtype) {
        case SS2D_F32: return launch_fwd<float, float, CROSS>(p, s, xi);
        case SS2D_F16:
            if (CROSS || p.out_dtype == SS2D_F32) return launch_fwd<__half, float, CROSS>(p, s, xi);
            return launch_fwd<__half, __half, false>(p, s, xi);
        case SS2D_BF16:
            if (CROSS || p.out_dtype == SS2D_F32) return launch_fwd<__nv_bfloat16, float, CROSS>(p, s, xi);
            return launch_fwd<__nv_bfloat16, __nv_bfloat16, false>(p, s, xi);
        default: return SS2D_EDTYPE;
    }
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
    if (ss2d::sl::supported(p)) return ss2d::sl::launch_fwd(p, reinterpret_cast<cudaStream_t>(stream));
    return ss2d::dispatch_fwd<false>(p, reinterpret_cast<cudaStream_t>(stream), ss2d::CrossInfo{0, 0, -1});
}

// Fused SS2D core forward (seam S3): see ss2d_cross_fwd_params in include/ss2d_b200.h.
extern "C" int ss2d_cross_scan_fwd(const ss2d_cross_fwd_params *pp, void *stream) {
    if (!pp) return SS2D_EINVAL;
    const ss2d_cross_fwd_params &c = *pp;
    if (!c.x || !c.delta || !c.B || !c.C || !c.A) return SS2D_EINVAL;
    if (!c.y && !c.ckpt) return SS2D_EINVAL;  // y == NULL: states-only sweep that rebuilds ckpt for the backward
    if (c.batch <= 0 || c.D <= 0 || c.H <= 0 || c.W <= 0 || c.dstate <= 0 || c.dstate > SS2D_MAX_DSTATE) return SS2D_EINVAL;
    if (c.H * c.W > 0x7fffffffLL || c.batch * 4 * c.D > 0x7fffffffLL) return SS2D_EINVAL;
    const int64_t L = c.H * c.W;
    ss2d_scan_fwd_params p{};
    p.batch = c.batch; p.dim = 4 * c.D; p.seqlen = L; p.dstate = c.dstate; p.ngroups = 4;
    p.in_dtype = c.in_dtype; p.out_dtype = SS2D_F32; p.delta_softplus = c.delta_softplus; p.family = c.family;
    p.u = c.x; p.delta = c.delta; p.A = c.A; p.B = c.B; p.C = c.C; p.D = c.Dskip; p.delta_bias = c.delta_bias;
    p.u_bstride = c.D * L; p.u_dstride = L;
    p.delta_bstride = 4 * c.D * L; p.delta_dstride = L;
    p.B_bstride = p.C_bstride = c.bc_bstride ? c.bc_bstride : 4 * c.dstate * L;
    p.B_gstride = p.C_gstride = c.bc_gstride ? c.bc_gstride : c.dstate * L;
    p.B_nstride = p.C_nstride = L;
    p.out = c.y; p.out_bstride = c.D * L; p.out_dstride = L;
    p.ckpt = c.ckpt;
    if (c.work && ss2d::sl::cross_covered(p)) {
        // state-lanes kernels: directions 1 / 3 over x^T, their outputs into y^T, folded back into y at the end
        cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
        const int64_t n = c.batch * c.D * L;
        float *xT = c.work, *yT = c.work + n;
        int rc = ss2d::plane_transpose(reinterpret_cast<const float *>(c.x), xT, c.batch * c.D, (int)c.H, (int)c.W, false, s);
        if (rc != 0) return rc;
        if (!c.y) return ss2d::sl::launch_cross_fwd(p, ss2d::sl::CrossAux{xT, nullptr, nullptr}, s);
        cudaError_t e = cudaMemsetAsync(yT, 0, (size_t)n * sizeof(float), s);
        if (e != cudaSuccess) return (int)e;
        rc = ss2d::sl::launch_cross_fwd(p, ss2d::sl::CrossAux{xT, nullptr, yT}, s);
        if (rc != 0) return rc;
        return ss2d::plane_transpose(yT, c.y, c.batch * c.D, (int)c.W, (int)c.H, true, s);
    }
    // warp-scan kernels: all four directions accumulate into y.  One launch (sum in arrival order), or — deterministic — one
    // launch per direction in the order k = 0, 1, 2, 3 (every y element then receives its four terms in that order).
    // The state-lanes branch above is bit-reproducible as it stands: y and y^T each take exactly two commutative adds onto 0.
    if (!c.deterministic || !c.y)
        return ss2d::dispatch_fwd<true>(p, reinterpret_cast<cudaStream_t>(stream), ss2d::CrossInfo{(int)c.H, (int)c.W, -1});
    for (int k = 0; k < 4; ++k) {
        const int rc = ss2d::dispatch_fwd<true>(p, reinterpret_cast<cudaStream_t>(stream), ss2d::CrossInfo{(int)c.H, (int)c.W, k});
        if (rc != 0) return rc;
    }
    return 0;
}
