// ss2d_scan_bwd.cu — selective-scan backward for sm_100a (seam S1, scan-order operands).
//
// Replaces selective_scan_bwd_kernel + launcher/host code (reference:
// kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_bwd_kernel_oflex.cuh:73-322,
// selective_scan_oflex.cpp:245-358).  Gradient formulas (same as the reference):
//     dx_t   = dout_t C_t + a_{t+1} dx_{t+1}                (suffix scan of (a_{t+1}, dout_t C_t))
//     dC_t  += dout_t h_t                dB_t += dx_t dl_t u_t               (summed over the group's channels)
//     du_t   = D dout_t + dl_t sum_n dx_t B_t
//     ddl_t  = u_t sum_n dx_t B_t + sum_n A_n dx_t (a_t h_{t-1})             a_t h_{t-1} = h_t - dl_t u_t B_t
//     dA_n  += sum_t dl_t dx_t (a_t h_{t-1})      dD += sum_t dout_t u_t     ddelta_t = ddl_t * softplus'(delta_t+bias)
// Organisation: same warp-per-channel / lanes-over-time mapping as the forward (ss2d_common.cuh), chunks
// walked right-to-left.  Chunk interiors are RECOMPUTED from the forward's fine checkpoints (h at every
// 256 steps) — no per-step state is ever stored.  The prefix (h) and suffix (dx) recurrences are each one
// thread-local pass + one warp-shuffle scan + one thread-local pass.  dB/dC are summed over the CTA's NW
// channels in shared memory before ONE red.global.add.v4.f32 per 4 timesteps leaves the CTA (the
// reference issues one scalar atomic per channel per element, 805 M at the microbench).
#include "ss2d_common.cuh"
#include "ss2d_scan_tile.cuh"
#include "ss2d_scan_sl.cuh"
#include "../../include/ss2d_b200.h"
#include <cstdlib>
#include <cstring>

#ifndef SS2D_BWD_T
#define SS2D_BWD_T 8
#define SS2D_BWD_NW 8
#define SS2D_BWD_MINB 2
#endif

namespace ss2d {

int plane_transpose(const float *src, float *dst, int64_t planes, int H, int W, bool acc, cudaStream_t stream);  // ss2d_cross.cu

struct BwdFlags {
    bool vec_u, vec_delta, vec_bc, vec_dout, vec_z, vec_out, vec_dbc, vec_grad;
};

// One state (index R of the staged block) of one chunk for this lane's T steps.
//   bc   : shared address of this lane's block in row 0 of the staged B tile (C rows SB rows later)
//   wa   : shared address of this warp's [A*log2e | h entering the chunk | dx carry] arrays (NSB bytes apart)
//          at the block's first state;  da: this lane's dA partial sums (128 bytes per state)
//   st   : this lane's slot in the fp32 staging rows of its warp (dB row, then dC row `stage_row` bytes later)
//   The staging area is double buffered (`par` selects the half): state r's partial sums are reduced by half of the
//   warps (alternating halves) while everybody already computes state r+1, so there is ONE block barrier per state;
//   the barrier of state r+1 also proves that every reader of buffer `par` is done before state r+2 refills it.
template <typename in_t, int T, int SB, int R, int NSB_CT, int NW>
__device__ __forceinline__ void bwd_state(uint32_t bc, uint32_t wa, int nsb_rt, uint32_t da, uint32_t st, uint32_t rd,
                                          int lane, bool active, bool reducer, float *red_dst, bool red_vec, int red_valid,
                                          const float (&dl)[T], const float (&du)[T], const float (&go)[T], float (&s)[T],
                                          float (&w)[T], float dsum, float qsum, float dlnext) {
    using RL = RowLayout<in_t, T>;
    using RLf = RowLayout<float, T>;
    constexpr int stage_row = RLf::row_bytes;
    const float A2 = lds_f32<R * 4>(wa);
    const float An = A2 * (1.f / kLog2e);
    float a[T], hv[T], Bv[T], Cv[T], bb[T];
    lds_row<in_t, T, R * RL::row_bytes>(bc, Bv);
#pragma unroll
    for (int i = 0; i < T; ++i) { a[i] = ex2(dl[i] * A2); bb[i] = du[i] * Bv[i]; }
    // ---- prefix recurrence: pass 1, warp scan, pass 2 (materialise h) ----
    float H = bb[0];
#pragma unroll
    for (int i = 1; i < T; ++i) H = fmaf(a[i], H, bb[i]);
    float P = ex2(A2 * dsum);
    warp_scan_inclusive(P, H, lane);
    float Pe, He;
    shift_up1(P, H, Pe, He);
    const uint32_t hin_addr = wa + (NSB_CT > 0 ? NSB_CT : nsb_rt);
    float h = fmaf(Pe, lds_f32<R * 4>(hin_addr), He);
#pragma unroll
    for (int i = 0; i < T; ++i) { h = fmaf(a[i], h, bb[i]); hv[i] = h; }
    // ---- suffix recurrence on (a_{t+1}, dout_t C_t) ----
    lds_row<in_t, T, (SB + R) * RL::row_bytes>(bc, Cv);
#pragma unroll
    for (int i = 0; i < T; ++i) Cv[i] *= go[i];  // g_t = dout_t C_t
    const float a_last = ex2(A2 * dlnext);        // a at the step following this lane's block
    float Rr = Cv[T - 1];
#pragma unroll
    for (int i = T - 2; i >= 0; --i) Rr = fmaf(a[i + 1], Rr, Cv[i]);
    float Qp = ex2(A2 * qsum);
    warp_rscan_inclusive(Qp, Rr, lane);
    float Qe, Re;
    shift_down1(Qp, Rr, Qe, Re);
    const uint32_t dx_addr = wa + (NSB_CT > 0 ? 2 * NSB_CT : 2 * nsb_rt);
    float dx = fmaf(Qe, lds_f32<R * 4>(dx_addr), Re);  // dx at the first step right of this lane's block
    // ---- gradients, walking the block right to left ----
    float dBv[T], dCv[T];
    float dA_acc = 0.f;
#pragma unroll
    for (int i = T - 1; i >= 0; --i) {
        const float an = i == T - 1 ? a_last : a[i + 1];
        dx = fmaf(an, dx, Cv[i]);
        dCv[i] = go[i] * hv[i];
        dBv[i] = dx * du[i];
        s[i] = fmaf(dx, Bv[i], s[i]);
        const float pq = dx * (hv[i] - bb[i]);  // dx_t * a_t h_{t-1}
        w[i] = fmaf(An, pq, w[i]);
        dA_acc = fmaf(dl[i], pq, dA_acc);
    }
    sts_f32_if<R * 4>(lane == 0, dx_addr, dx);  // dx at this chunk's first step -> carry for the left chunk
    dA_acc += __shfl_xor_sync(0xffffffffu, dA_acc, 1);  // lane pairs share one slot: 16 partials per (warp, state)
    sts_f32_if<R * 64>((lane & 1) == 0, da, lds_f32<R * 64>(da) + dA_acc);
    // ---- dB/dC: reduce over the CTA's channels in shared memory, then one vector reduction per 4 steps ----
    if (!active) {
#pragma unroll
        for (int i = 0; i < T; ++i) { dBv[i] = 0.f; dCv[i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < T / 4; ++i) {
        if (i == 0) { sts_v4<0>(st, pack16<float>(&dBv[0])); sts_v4<stage_row>(st, pack16<float>(&dCv[0])); }
        if (i == 1) { sts_v4<16>(st, pack16<float>(&dBv[4])); sts_v4<stage_row + 16>(st, pack16<float>(&dCv[4])); }
        if (i == 2) { sts_v4<32>(st, pack16<float>(&dBv[8])); sts_v4<stage_row + 32>(st, pack16<float>(&dCv[8])); }
        if (i == 3) { sts_v4<48>(st, pack16<float>(&dBv[12])); sts_v4<stage_row + 48>(st, pack16<float>(&dCv[12])); }
    }
    __syncthreads();  // the only block barrier of this state
    if (reducer) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#define SS2D_ACC(WW)                                                                  \
        if constexpr (WW < NW) {                                                      \
            const uint4 q = lds_v4<WW * 2 * stage_row>(rd);                           \
            acc.x += __uint_as_float(q.x); acc.y += __uint_as_float(q.y);            \
            acc.z += __uint_as_float(q.z); acc.w += __uint_as_float(q.w);            \
        }
        SS2D_ACC(0) SS2D_ACC(1) SS2D_ACC(2) SS2D_ACC(3) SS2D_ACC(4) SS2D_ACC(5) SS2D_ACC(6) SS2D_ACC(7)
        SS2D_ACC(8) SS2D_ACC(9) SS2D_ACC(10) SS2D_ACC(11) SS2D_ACC(12) SS2D_ACC(13) SS2D_ACC(14) SS2D_ACC(15)
#undef SS2D_ACC
        static_assert(NW <= 16, "reduction unrolled for at most 16 warps");
        const bool fast = red_vec && red_valid >= 4;
        red_add_v4_if(fast, red_dst, acc.x, acc.y, acc.z, acc.w);
        if (!fast) {  // ragged tail / unaligned rows only
            if (red_valid > 0) atomicAdd(red_dst + 0, acc.x);
            if (red_valid > 1) atomicAdd(red_dst + 1, acc.y);
            if (red_valid > 2) atomicAdd(red_dst + 2, acc.z);
            if (red_valid > 3) atomicAdd(red_dst + 3, acc.w);
        }
    }
}

template <typename in_t, int T, int SB, int NSB_CT, int NW, int R = 0>
__device__ __forceinline__ void bwd_block(uint32_t bc, uint32_t wa, int nsb_rt, uint32_t da, uint32_t st, uint32_t st_other,
                                          uint32_t rd, uint32_t rd_other, int lane,
                                          bool active, bool reducer, float *red_dst, int64_t L, bool red_vec, int red_valid,
                                          const float (&dl)[T], const float (&du)[T], const float (&go)[T], float (&s)[T],
                                          float (&w)[T], float dsum, float qsum, float dlnext, int n_here) {
    if constexpr (R < SB) {
        if (R < n_here) {  // uniform across the CTA: the barrier inside bwd_state is safe
            bwd_state<in_t, T, SB, R, NSB_CT, NW>(bc, wa, nsb_rt, da, st, rd, lane, active, reducer, red_dst, red_vec,
                                                 red_valid, dl, du, go, s, w, dsum, qsum, dlnext);
            // next state: other staging half, other half of the warps reduces
            bwd_block<in_t, T, SB, NSB_CT, NW, R + 1>(bc, wa, nsb_rt, da, st_other, st, rd_other, rd, lane, active,
                                                     !reducer, red_dst + L, L, red_vec, red_valid, dl, du, go, s, w, dsum,
                                                     qsum, dlnext, n_here);
        }
    }
}

// CROSS (fused seam S3): u and dout are gathered from the spatial-order planes x[b,d] / dy[b,d] with the
// direction's addressing, and du is accumulated (red.global.add) into the spatial-order fp32 plane dx[b,d]
// — CrossMerge.backward and CrossScan.backward as load / store addressing.  ddelta, dB, dC stay in scan order.
template <typename in_t, typename out_t, int T, int NW, int SB, int MINB, bool CROSS, int NSB_CT>
__global__ void __launch_bounds__(NW * kWarp, MINB)
scan_bwd_kernel(const ss2d_scan_bwd_params pb, const int tiles_per_group, const BwdFlags fl, const CrossInfo xinfo) {
    using FT = BCTile<in_t, T, SB>;
    using RL = typename FT::RL;
    using RLf = RowLayout<float, T>;  // staging rows for the dB/dC reduction are fp32
    constexpr int chunk = FT::chunk;
    constexpr int NT = NW * kWarp;
    constexpr int ckpt_per_chunk = chunk / SS2D_CKPT_STEPS;
    constexpr int stage_row = RLf::row_bytes;
    static_assert(chunk % SS2D_CKPT_STEPS == 0, "chunk must be a multiple of the checkpoint spacing");
    static_assert(2 * (chunk / 4) == NT / 2, "half of the CTA's threads reduce one state: needs T == NW");
    constexpr int stage_half = NW * 2 * stage_row;  // one staging buffer: [NW][dB row | dC row]
    extern __shared__ __align__(16) unsigned char smem[];
    const ss2d_scan_fwd_params &p = pb.f;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = (int)p.dstate;
    const int NS = NSB_CT > 0 ? NSB_CT / 4 : ((N + 3) & ~3);
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

    // ---- shared memory carve-up ----
    unsigned char *tiles = smem;                                             // 2 x B/C tile (ping-pong)
    unsigned char *stage = smem + 2 * FT::tile_bytes;                        // [NW][2][padded chunk] fp32
    float *wsm = reinterpret_cast<float *>(stage + 2 * stage_half) + warp * 3 * NS;  // [A2 | Hin | Dx] of this warp
    float *sA2 = wsm, *sHin = wsm + NS, *sDx = wsm + 2 * NS;
    float *sdA_all = reinterpret_cast<float *>(stage + 2 * stage_half) + NW * 3 * NS;
    float *sdA = sdA_all + warp * N * 16;                                    // [N][16] lane-pair dA partials
    const uint32_t tiles_addr = smem_u32(tiles) + RL::lane_unit(lane) * 16;
    const uint32_t wsm_addr = smem_u32(wsm);
    const uint32_t da_addr = smem_u32(sdA) + (lane >> 1) * 4;
    const uint32_t st_addr = smem_u32(stage) + warp * 2 * stage_row + RLf::lane_unit(lane) * 16;
    // cross-channel reduction: thread `task` sums one 16-byte piece of dB (tasks 0..chunk/4-1) or dC over the NW warps
    // (the reducing half of the CTA alternates from state to state)
    const int task = threadIdx.x % (NT / 2);
    const int which = task / (chunk / 4), piece = task % (chunk / 4);
    int par = 0;  // staging buffer / reducing half of the next state
    const uint32_t rd_addr = smem_u32(stage) + which * stage_row + RLf::unit_of_piece(piece) * 16;

    const int64_t cu = CROSS ? (active ? c_local : per_g - 1) : c;  // row of u / dout / du: d in fused mode
    const in_t *u_row = reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + cu * p.u_dstride;
    const in_t *d_row = reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + c * p.delta_dstride;
    const in_t *z_row = p.z ? reinterpret_cast<const in_t *>(p.z) + b * p.z_bstride + c * p.z_dstride : nullptr;
    const out_t *pre_row = p.z ? reinterpret_cast<const out_t *>(p.out) + b * p.out_bstride + c * p.out_dstride : nullptr;
    const out_t *g_row = reinterpret_cast<const out_t *>(pb.dout) + b * pb.dout_bstride + cu * pb.dout_dstride;
    const in_t *Bg = reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride;
    const in_t *Cg = reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride;
    const int64_t row = ((int64_t)b * p.dim + c) * L;
    in_t *du_row = reinterpret_cast<in_t *>(pb.du) + row;
    float *dx_plane = CROSS ? reinterpret_cast<float *>(pb.du) + ((int64_t)b * per_g + cu) * L : nullptr;
    in_t *dd_row = reinterpret_cast<in_t *>(pb.ddelta) + row;
    in_t *dz_row = pb.dz ? reinterpret_cast<in_t *>(pb.dz) + row : nullptr;
    float *red_base = (which ? pb.dC : pb.dB) + ((int64_t)b * p.ngroups + g) * N * L + (int64_t)piece * 4;
    const float Dv = p.D ? p.D[c] : 0.f;
    const float bias = p.delta_bias ? p.delta_bias[c] : 0.f;
    const int n_fine = (int)((L + SS2D_CKPT_STEPS - 1) / SS2D_CKPT_STEPS);
    const float *ck_row = p.ckpt ? p.ckpt + ((int64_t)b * p.dim + c) * n_fine * N : nullptr;

    for (int n = lane; n < N; n += kWarp) {
        sA2[n] = p.A[c * N + n] * kLog2e;
        sDx[n] = 0.f;
    }
    for (int i = lane; i < N * 16; i += kWarp) sdA[i] = 0.f;

    const int n_sb = (N + SB - 1) / SB;
    const int n_chunks = (int)((L + chunk - 1) / chunk);
    const int Q = n_chunks * n_sb;

    stage_bc<in_t, T, SB, NT>(tiles, Bg, Cg, p.B_nstride, p.C_nstride, 0, N, (int64_t)(n_chunks - 1) * chunk, L, fl.vec_bc);
    cp_async_commit();

    float dl[T], uv[T], du[T], go[T], s[T], w[T];
    float dsum = 0.f, qsum = 0.f, dlnext = 0.f;
    float dl_first_right = 0.f;  // dl at the first step of the chunk to the right (0 past the end: a = 1)
    float dD_acc = 0.f, dbias_acc = 0.f;

    int ci = n_chunks - 1, sb = 0;
    float hin_next = (ci > 0 && ck_row && lane < N && N <= kWarp) ? ck_row[((int64_t)ci * ckpt_per_chunk - 1) * N + lane] : 0.f;
    for (int q = 0; q < Q; ++q) {
        const int64_t t0 = (int64_t)ci * chunk;
        cp_async_wait<0>();
        __syncthreads();
        const int64_t tl = t0 + lane * T;
        const int valid = (int)min((int64_t)T, L - tl);
        if (sb == 0) {  // this chunk's global loads are issued first, ahead of the cp.async burst of the next tile
            if constexpr (CROSS) {
                load_block_cross<in_t, T>(u_row, uv, g, tl, L, xinfo, fl.vec_u);
                load_block_cross<out_t, T>(g_row, go, g, tl, L, xinfo, fl.vec_dout);
            } else {
                load_block<in_t, T>(u_row + tl, uv, valid, fl.vec_u);
                load_block<out_t, T>(g_row + tl, go, valid, fl.vec_dout);
            }
            load_block<in_t, T>(d_row + tl, dl, valid, fl.vec_delta);
            if (ci > 0) {  // pull the next (left) chunk's lines into L2 while this chunk computes
                prefetch_l2(d_row + tl - chunk);
                if constexpr (!CROSS) { prefetch_l2(u_row + tl - chunk); prefetch_l2(g_row + tl - chunk); }
            }
        }
        if (q + 1 < Q) {
            const int sb1 = sb + 1 == n_sb ? 0 : sb + 1, ci1 = sb + 1 == n_sb ? ci - 1 : ci;
            stage_bc<in_t, T, SB, NT>(tiles + ((q + 1) & 1) * FT::tile_bytes, Bg, Cg, p.B_nstride, p.C_nstride,
                                      sb1 * SB, N, (int64_t)ci1 * chunk, L, fl.vec_bc);
            cp_async_commit();
        }
        if (sb == 0) {
            if (z_row) {  // out = pre * silu(z): dz and the gated upstream gradient
                float zv[T], pre[T];
                load_block<in_t, T>(z_row + tl, zv, valid, fl.vec_z);
                load_block<out_t, T>(pre_row + tl, pre, valid, fl.vec_out);
#pragma unroll
                for (int i = 0; i < T; ++i) {
                    const float sg = sigmoidf_fast(zv[i]);
                    pre[i] = go[i] * pre[i] * sg * (1.f + zv[i] * (1.f - sg));
                    go[i] *= zv[i] * sg;
                }
                if (active) store_block<in_t, T>(dz_row + tl, pre, valid, fl.vec_grad);
            }
            dsum = 0.f;
#pragma unroll
            for (int i = 0; i < T; ++i) {
                float d = dl[i] + bias;
                if (p.delta_softplus) d = softplus_ref(d);
                d = i < valid ? d : 0.f;
                dl[i] = d;
                du[i] = d * uv[i];
                s[i] = 0.f;
                w[i] = 0.f;
                dsum += d;
                dD_acc = fmaf(go[i], uv[i], dD_acc);
            }
            // dl of the step right after this lane's block: lane+1's first, or the right chunk's first
            dlnext = __shfl_down_sync(0xffffffffu, dl[0], 1);
            if (lane == 31) dlnext = dl_first_right;
            qsum = dsum - dl[0] + dlnext;
            // h entering this chunk, from the forward's fine checkpoints; the value for the NEXT (left) chunk is
            // fetched now into a register so that its global-load latency hides behind this chunk's 16 states
            __syncwarp();
            if (N <= kWarp) {
                if (lane < N) sHin[lane] = hin_next;
                hin_next = (ci > 1 && ck_row && lane < N) ? ck_row[((int64_t)(ci - 1) * ckpt_per_chunk - 1) * N + lane] : 0.f;
            } else {
                for (int n = lane; n < N; n += kWarp)
                    sHin[n] = (ci > 0 && ck_row) ? ck_row[((int64_t)ci * ckpt_per_chunk - 1) * N + n] : 0.f;
            }
            __syncwarp();
        }
        {
            const uint32_t bc = tiles_addr + (q & 1) * FT::tile_bytes;
            const int64_t tp = t0 + (int64_t)piece * 4;
            const int n_here = min(SB, N - sb * SB);
            const uint32_t st0 = st_addr + par * stage_half, st1 = st_addr + (par ^ 1) * stage_half;
            const uint32_t rd0 = rd_addr + par * stage_half, rd1 = rd_addr + (par ^ 1) * stage_half;
            bwd_block<in_t, T, SB, NSB_CT, NW>(bc, wsm_addr + sb * SB * 4, nsb_rt, da_addr + sb * SB * 64, st0, st1, rd0, rd1,
                                               lane, active, (warp < NW / 2) == (par == 0),
                                               red_base + (int64_t)sb * SB * L + t0, L, fl.vec_dbc,
                                               (int)min((int64_t)4, L - tp), dl, du, go, s, w, dsum, qsum, dlnext, n_here);
            par ^= n_here & 1;
        }
        if (sb == n_sb - 1) {  // chunk finished: du, ddelta
            float ddl[T];
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const float v = fmaf(uv[i], s[i], w[i]);
                // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x)); exact 1 beyond the x > 20 cut-off
                const float sg = p.delta_softplus ? -expm1f(-dl[i]) : 1.f;
                ddl[i] = v * sg;
                dbias_acc += i < valid ? ddl[i] : 0.f;
                s[i] = fmaf(dl[i], s[i], Dv * go[i]);  // du
            }
            if (active) {
                if constexpr (CROSS) red_block_cross<T>(dx_plane, s, g, tl, L, xinfo, fl.vec_dbc);
                else store_block<in_t, T>(du_row + tl, s, valid, fl.vec_grad);
                store_block<in_t, T>(dd_row + tl, ddl, valid, fl.vec_grad);
            }
            dl_first_right = __shfl_sync(0xffffffffu, dl[0], 0);
        }
        if (++sb == n_sb) { sb = 0; --ci; }
    }
    // ---- per-channel reductions over time (and atomically over batch) ----
    __syncwarp();
    if (active) {
        for (int n = 0; n < N; ++n) {
            float v = lane < 16 ? sdA[n * 16 + lane] : 0.f;
#pragma unroll
            for (int d = 8; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) atomicAdd(pb.dA + c * N + n, v);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, d);
            dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, d);
        }
        if (lane == 0) {
            if (pb.dD) atomicAdd(pb.dD + c, dD_acc);
            if (pb.ddelta_bias) atomicAdd(pb.ddelta_bias + c, dbias_acc);
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename in_t, typename out_t, int T, int NW, int MINB, bool CROSS = false, int SB = 8>
static int launch_bwd(const ss2d_scan_bwd_params &pb, cudaStream_t stream, CrossInfo xinfo = CrossInfo{0, 0, -1}) {
    using FT = BCTile<in_t, T, SB>;
    const ss2d_scan_fwd_params &p = pb.f;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tiles = (per_g + NW - 1) / NW;
    const int N = (int)p.dstate;
    const bool small_n = N <= 16;
    const int NS = small_n ? 16 : ((N + 3) & ~3);
    const size_t smem = 2 * FT::tile_bytes + 2 * (size_t)NW * 2 * RowLayout<float, T>::row_bytes +
                        ((size_t)3 * NW * NS + (size_t)NW * N * 16) * sizeof(float);
    const int64_t ei = sizeof(in_t), eo = sizeof(out_t);
    BwdFlags fl;
    fl.vec_u = aligned16(p.u) && (p.u_bstride * ei) % 16 == 0 && (p.u_dstride * ei) % 16 == 0;
    fl.vec_delta = aligned16(p.delta) && (p.delta_bstride * ei) % 16 == 0 && (p.delta_dstride * ei) % 16 == 0;
    fl.vec_bc = aligned16(p.B) && aligned16(p.C) && (p.B_bstride * ei) % 16 == 0 && (p.B_gstride * ei) % 16 == 0 &&
                (p.B_nstride * ei) % 16 == 0 && (p.C_bstride * ei) % 16 == 0 && (p.C_gstride * ei) % 16 == 0 &&
                (p.C_nstride * ei) % 16 == 0;
    fl.vec_dout = aligned16(pb.dout) && (pb.dout_bstride * eo) % 16 == 0 && (pb.dout_dstride * eo) % 16 == 0;
    fl.vec_z = p.z && aligned16(p.z) && (p.z_bstride * ei) % 16 == 0 && (p.z_dstride * ei) % 16 == 0;
    fl.vec_out = p.out && aligned16(p.out) && (p.out_bstride * eo) % 16 == 0 && (p.out_dstride * eo) % 16 == 0;
    fl.vec_dbc = aligned16(pb.dB) && aligned16(pb.dC) && p.seqlen % 4 == 0;
    // du / ddelta / dz rows are contiguous (batch, dim, L)
    fl.vec_grad = aligned16(pb.du) && aligned16(pb.ddelta) && (!pb.dz || aligned16(pb.dz)) && (p.seqlen * ei) % 16 == 0;
    if (CROSS) fl.vec_dbc = fl.vec_dbc && aligned16(pb.du);  // dx plane rows are L floats: 16-byte aligned iff L % 4 == 0
    const int64_t grid = p.batch * (xinfo.g_only >= 0 ? 1 : p.ngroups) * tiles;
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)smem);
        if (rc != 0) return rc;
        kern<<<(unsigned)grid, NW * kWarp, smem, stream>>>(pb, tiles, fl, xinfo);
        return (int)cudaGetLastError();
    };
    if (small_n) return go(scan_bwd_kernel<in_t, out_t, T, NW, SB, MINB, CROSS, 64>);
    return go(scan_bwd_kernel<in_t, out_t, T, NW, SB, MINB, CROSS, 0>);
}

}  // namespace ss2d

extern "C" int ss2d_selective_scan_bwd(const ss2d_scan_bwd_params *pp, void *stream) {
    if (!pp) return SS2D_EINVAL;
    ss2d_scan_bwd_params pb = *pp;
    const ss2d_scan_fwd_params &p = pb.f;
    if (!p.u || !p.delta || !p.A || !p.B || !p.C || !pb.dout || !pb.du || !pb.ddelta || !pb.dA || !pb.dB || !pb.dC)
        return SS2D_EINVAL;
    if (p.batch <= 0 || p.dim <= 0 || p.seqlen <= 0 || p.dstate <= 0 || p.ngroups <= 0) return SS2D_EINVAL;
    if (p.dim % p.ngroups != 0 || p.dstate > SS2D_MAX_DSTATE) return SS2D_EINVAL;
    if ((p.D && !pb.dD) || (p.delta_bias && !pb.ddelta_bias)) return SS2D_EINVAL;
    if (p.z && (!pb.dz || !p.out)) return SS2D_EINVAL;
    if (p.out_dtype != SS2D_F32 && p.out_dtype != p.in_dtype) return SS2D_EDTYPE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool state_lanes = ss2d::sl::supported(p);
    if (!p.ckpt && p.seqlen > (state_lanes ? SS2D_SL_BLOCK : SS2D_CKPT_STEPS)) {
        // Foreign caller that only kept the reference's coarse x: rebuild the fine checkpoints with a
        // states-only forward sweep (out == NULL) into the caller's scratch buffer.
        if (!pb.ckpt_scratch) return SS2D_EINVAL;
        ss2d_scan_fwd_params f = p;
        f.out = nullptr; f.out_z = nullptr; f.z = nullptr; f.x = nullptr; f.ckpt = pb.ckpt_scratch;
        f.out_dtype = SS2D_F32;
        const int rc = ss2d_selective_scan_fwd(&f, stream);
        if (rc != 0) return rc;
        pb.f.ckpt = pb.ckpt_scratch;
    }
    if (state_lanes) return ss2d::sl::launch_bwd(pb, s);
    using namespace ss2d;
    constexpr int T = SS2D_BWD_T, NW = SS2D_BWD_NW, MINB = SS2D_BWD_MINB;
    switch (p.in_dtype) {
        case SS2D_F32: return launch_bwd<float, float, T, NW, MINB>(pb, s);
        case SS2D_F16:
            return p.out_dtype == SS2D_F32 ? launch_bwd<__half, float, T, NW, MINB>(pb, s)
                                           : launch_bwd<__half, __half, T, NW, MINB>(pb, s);
        case SS2D_BF16:
            return p.out_dtype == SS2D_F32 ? launch_bwd<__nv_bfloat16, float, T, NW, MINB>(pb, s)
                                           : launch_bwd<__nv_bfloat16, __nv_bfloat16, T, NW, MINB>(pb, s);
        default: return SS2D_EDTYPE;
    }
}

// Fused SS2D core backward (seam S3): see ss2d_cross_bwd_params in include/ss2d_b200.h.
extern "C" int ss2d_cross_scan_bwd(const ss2d_cross_bwd_params *pp, void *stream) {
    if (!pp) return SS2D_EINVAL;
    const ss2d_cross_fwd_params &c = pp->f;
    if (!c.x || !c.delta || !c.B || !c.C || !c.A || !pp->dy || !pp->dx || !pp->ddelta || !pp->dA || !pp->dB || !pp->dC)
        return SS2D_EINVAL;
    if (c.batch <= 0 || c.D <= 0 || c.H <= 0 || c.W <= 0 || c.dstate <= 0 || c.dstate > SS2D_MAX_DSTATE) return SS2D_EINVAL;
    if ((c.Dskip && !pp->dDskip) || (c.delta_bias && !pp->ddelta_bias)) return SS2D_EINVAL;
    const int64_t L = c.H * c.W;
    ss2d_scan_bwd_params pb{};
    ss2d_scan_fwd_params &p = pb.f;
    p.batch = c.batch; p.dim = 4 * c.D; p.seqlen = L; p.dstate = c.dstate; p.ngroups = 4;
    p.in_dtype = c.in_dtype; p.out_dtype = SS2D_F32; p.delta_softplus = c.delta_softplus; p.family = c.family;
    p.u = c.x; p.delta = c.delta; p.A = c.A; p.B = c.B; p.C = c.C; p.D = c.Dskip; p.delta_bias = c.delta_bias;
    p.u_bstride = c.D * L; p.u_dstride = L;
    p.delta_bstride = 4 * c.D * L; p.delta_dstride = L;
    p.B_bstride = p.C_bstride = c.bc_bstride ? c.bc_bstride : 4 * c.dstate * L;
    p.B_gstride = p.C_gstride = c.bc_gstride ? c.bc_gstride : c.dstate * L;
    p.B_nstride = p.C_nstride = L;
    p.ckpt = c.ckpt;
    pb.dout = pp->dy; pb.dout_bstride = c.D * L; pb.dout_dstride = L;
    pb.du = pp->dx; pb.ddelta = pp->ddelta;
    pb.dA = pp->dA; pb.dB = pp->dB; pb.dC = pp->dC; pb.dD = pp->dDskip; pb.ddelta_bias = pp->ddelta_bias;
    using namespace ss2d;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool det = c.deterministic != 0;
    {
        ss2d_scan_fwd_params q = p;  // the forward's view of the problem decides the kernel family (checkpoint layout)
        q.out_bstride = c.D * L; q.out_dstride = L;
        if (c.work && ss2d::sl::cross_covered(q)) {
            if (!p.ckpt && L > SS2D_SL_BLOCK) return SS2D_EINVAL;
            const int64_t n = c.batch * c.D * L;
            float *xT = c.work, *dyT = c.work + n, *dxT = c.work + 2 * n;
            int rc = ss2d::plane_transpose(reinterpret_cast<const float *>(c.x), xT, c.batch * c.D, (int)c.H, (int)c.W, false, s);
            if (rc == 0) rc = ss2d::plane_transpose(pp->dy, dyT, c.batch * c.D, (int)c.H, (int)c.W, false, s);
            if (rc != 0) return rc;
            cudaError_t e = cudaMemsetAsync(dxT, 0, (size_t)n * sizeof(float), s);
            if (e != cudaSuccess) return (int)e;
            rc = ss2d::sl::launch_cross_bwd(pb, ss2d::sl::CrossAux{xT, dyT, dxT}, s);
            if (rc != 0) return rc;
            return ss2d::plane_transpose(dxT, pp->dx, c.batch * c.D, (int)c.W, (int)c.H, true, s);
        }
    }
    if (!p.ckpt && L > SS2D_CKPT_STEPS) return SS2D_EINVAL;  // the fused backward needs the forward's checkpoints
    constexpr int T = SS2D_BWD_T, NW = SS2D_BWD_NW, MINB = SS2D_BWD_MINB;
    // deterministic: dx receives the four directions in the order k = 0..3 (dA / dB / dC / dD / dbias stay sums of
    // atomics over the channel tiles, like the reference's backward, selective_scan_bwd_kernel_oflex.cuh:259-273)
    for (int k = det ? 0 : -1; k < (det ? 4 : 0); ++k) {
        const CrossInfo xi{(int)c.H, (int)c.W, k};
        int rc;
        switch (c.in_dtype) {
            case SS2D_F32: rc = launch_bwd<float, float, T, NW, MINB, true>(pb, s, xi); break;
            case SS2D_F16: rc = launch_bwd<__half, float, T, NW, MINB, true>(pb, s, xi); break;
            case SS2D_BF16: rc = launch_bwd<__nv_bfloat16, float, T, NW, MINB, true>(pb, s, xi); break;
            default: return SS2D_EDTYPE;
        }
        if (rc != 0) return rc;
    }
    return 0;
}
