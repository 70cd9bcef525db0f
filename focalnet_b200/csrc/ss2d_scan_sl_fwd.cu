// ss2d_scan_sl_fwd.cu — selective-scan forward, "state-lanes" organisation (dstate == 16), sm_100a.
//
// Replaces selective_scan_fwd_kernel + launcher (reference: kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_fwd_kernel_oflex.cuh:67-211).  Same maths as ss2d_scan_fwd.cu; the
// organisation is described in ss2d_scan_sl.cuh.  Per (element, state) the steady-state loop issues
//   0.5 FMUL2 (delta*A2) + 1 MUFU.EX2 + 0.5 FMUL2 (delta*u*B) + 1 FFMA (h) + 0.5 FFMA2 (y += C h) + 0.5 LDS.128
// against the 14.5 warp-instructions of the warp-scan kernel.
//
// HBM traffic per launch: s_in*(2*B*Dm*L + 2*B*G*N*L) + s_out*B*Dm*L + 4*B*Dm*(ceil(L/16)*N) [block checkpoints]
//                         + 4*B*Dm*ceil(L/2048)*2N [reference x]
#include "ss2d_scan_sl.cuh"
#include <atomic>

namespace ss2d {
namespace sl {

template <typename in_t, int SN, int NW, int TT> struct FwdSmem {
    using M = Map<SN>;
    static constexpr int NSTAGE = 3;  // tile ring: block k+1's softplus reads one stage ahead of block k's B/C
    static constexpr int CPC = NW * M::CPW;
    static constexpr int RSU = TT + 64 / (int)sizeof(in_t);  // channel rows: +64 B, two rows of an LDS phase differ by 16 banks
    static constexpr int RSB = TT + 16 / (int)sizeof(in_t);  // state rows: +16 B, consecutive rows tile the banks
    static constexpr int u_off = 0;
    static constexpr int d_off = u_off + CPC * RSU * (int)sizeof(in_t);
    static constexpr int B_off = d_off + CPC * RSU * (int)sizeof(in_t);
    static constexpr int C_off = B_off + kN * RSB * (int)sizeof(in_t);
    static constexpr int stage_bytes = C_off + kN * RSB * (int)sizeof(in_t);
    static_assert(stage_bytes % 16 == 0 && d_off % 16 == 0 && B_off % 16 == 0 && C_off % 16 == 0, "alignment");
    static constexpr int xch_off = NSTAGE * stage_bytes;            // [warp][buf][dl|du][q][cw] float4
    static constexpr int xch_warp = 2 * 2 * (BK / 4) * M::CPW * 16;
    static constexpr int total = xch_off + NW * xch_warp;
};

// The block loop is software pipelined by hand so that ONE warp always has independent work in flight (the
// microbench puts only ~2.6 warps on an SM sub-partition): iteration k runs the 16 x SN recurrence steps of block
// k, the lane reduction + stores of block k-1 and the softplus of block k+1 in one basic block.
// FAST: every tensor 16-byte aligned and L % 16 == 0 — no tail masks, one predicated vector store per output.
// CROSS (fused seam S3, FAST only): the groups are the 4 scan directions, `u` is the spatial-order fp32 plane x[b, d]
// shared by the 4 directions and `out` the merged fp32 plane y[b, d], accumulated with red.global.add (zero-filled).
// Directions 0 / 2 walk x forwards / backwards; directions 1 / 3 walk the transposed copy x^T (aux.uT) and accumulate
// into y^T (aux.accT), so every direction moves contiguous runs.  delta / B / C stay in scan order.
template <typename in_t, typename out_t, int SN, int NW, int TT, bool FAST, bool CROSS>
__device__ __forceinline__ void sl_fwd_body(const ss2d_scan_fwd_params &p, const int tiles_per_group, const Flags &fl, const CrossAux &aux) {
    static_assert(!CROSS || (FAST && sizeof(in_t) == 4 && sizeof(out_t) == 4), "fused seam: fp32, aligned, L % 16 == 0");
    using M = Map<SN>;
    using SM = FwdSmem<in_t, SN, NW, TT>;
    constexpr int NT = NW * kWarp, CPC = SM::CPC, LPC = M::LPC, CPW = M::CPW, OWN = M::OWN, NQ = BK / 4, BPS = TT / BK;
    constexpr int NSTAGE = SM::NSTAGE;
    extern __shared__ __align__(16) unsigned char smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cw = lane / LPC, ng = lane % LPC;
    const int L = (int)p.seqlen;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tile = blockIdx.x % tiles_per_group;
    const int bg = blockIdx.x / tiles_per_group;
    const int g = bg % (int)p.ngroups, b = bg / (int)p.ngroups;
    const int ch0 = tile * CPC;
    const int rows_valid = min(CPC, per_g - ch0);
    const int c_local = ch0 + warp * CPW + cw;
    const bool active = c_local < per_g;
    const int64_t c = (int64_t)g * per_g + (active ? c_local : per_g - 1);

    // per-warp exchange area for softplus(delta) and delta*u of a block (double buffered)
    float *xch = reinterpret_cast<float *>(smem + SM::xch_off + warp * SM::xch_warp);
    // element (kind, step j, channel) of buffer f lives at float index (((f*2 + kind)*NQ + j/4)*CPW + cw)*4 + j%4
    float *xpub = xch + (((OWN * ng) >> 2) * CPW + cw) * 4 + ((OWN * ng) & 3);
    const float4 *xq = reinterpret_cast<const float4 *>(xch) + cw;
    constexpr int XBUF = 2 * NQ * CPW;  // float4 per buffer

    const int64_t row0 = (int64_t)g * per_g + ch0;
    const int64_t urow0 = CROSS ? ch0 : row0;  // row of u / out: d in fused mode
    RowStager<in_t, TT, NT> st_u, st_d, st_B, st_C;
    st_u.init(reinterpret_cast<in_t *>(smem + SM::u_off), SM::RSU, reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + urow0 * p.u_dstride,
              p.u_dstride, CPC, rows_valid, fl.vec_u, 0);
    st_d.init(reinterpret_cast<in_t *>(smem + SM::d_off), SM::RSU,
              reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + row0 * p.delta_dstride, p.delta_dstride, CPC, rows_valid,
              fl.vec_delta, 0);
    st_B.init(reinterpret_cast<in_t *>(smem + SM::B_off), SM::RSB, reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride,
              p.B_nstride, kN, kN, fl.vec_bc, SN);
    st_C.init(reinterpret_cast<in_t *>(smem + SM::C_off), SM::RSB, reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride,
              p.C_nstride, kN, kN, fl.vec_bc, SN);

    // FAST: one flattened copy list instead of the four stagers.  Fused seam: the u rows are rows of x (direction 0), of
    // x^T (1), or the same rows walked backwards (2 / 3: mirrored 16-byte pieces, un-mirrored by lds_own) — contiguous
    // 16-byte copies in every direction; the element-wise gather only serves a short last stage (L % TT != 0)
    constexpr int NPIECE = (2 * CPC + 2 * kN) * (TT * (int)sizeof(in_t) / 16);
    constexpr int NCOPY = (NPIECE + NT - 1) / NT;
    CopyList<NT, NCOPY> cl;
    const bool rev = CROSS && (g & 2);
    if constexpr (CROSS) {
        const float *plane = ((g & 1) ? aux.uT : reinterpret_cast<const float *>(p.u)) + b * p.u_bstride + urow0 * p.u_dstride;
        st_u.init(reinterpret_cast<in_t *>(smem + SM::u_off), SM::RSU, reinterpret_cast<const in_t *>(plane), p.u_dstride, CPC, rows_valid,
                  false, 0);  // short last stage only (element-wise, bounds-checked)
        cl.clear();
        cl.add(reinterpret_cast<float *>(smem + SM::u_off), SM::RSU, plane, p.u_dstride, CPC, rows_valid, TT, 0, 0, TT, rev ? (int)p.seqlen : 0);
    }
    if constexpr (FAST) {
        if constexpr (!CROSS) {
            cl.clear();
            cl.add(reinterpret_cast<in_t *>(smem + SM::u_off), SM::RSU, reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + row0 * p.u_dstride,
                   p.u_dstride, CPC, rows_valid, TT, 0, 0, TT);
        }
        cl.add(reinterpret_cast<in_t *>(smem + SM::d_off), SM::RSU,
               reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + row0 * p.delta_dstride, p.delta_dstride, CPC, rows_valid, TT, 0, 0,
               TT);
        cl.add(reinterpret_cast<in_t *>(smem + SM::B_off), SM::RSB, reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride,
               p.B_nstride, kN, kN, TT, SN, 0, TT);
        cl.add(reinterpret_cast<in_t *>(smem + SM::C_off), SM::RSB, reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride,
               p.C_nstride, kN, kN, TT, SN, 0, TT);
    }

    const in_t *z_row = p.z ? reinterpret_cast<const in_t *>(p.z) + b * p.z_bstride + c * p.z_dstride : nullptr;
    const int64_t c_out = CROSS ? (active ? c_local : per_g - 1) : c;
    out_t *o_base = reinterpret_cast<out_t *>(p.out);
    if constexpr (CROSS) { if (g & 1) o_base = reinterpret_cast<out_t *>(aux.accT); }
    out_t *o_row = o_base ? o_base + b * p.out_bstride + c_out * p.out_dstride : nullptr;
    out_t *oz_row = p.out_z ? reinterpret_cast<out_t *>(p.out_z) + b * p.out_bstride + c * p.out_dstride : nullptr;
    const float Dv = p.D ? p.D[c] : 0.f;
    const float bias = p.delta_bias ? p.delta_bias[c] : 0.f;
    const bool softplus = p.delta_softplus != 0;

    float A2[SN], h[SN];
#pragma unroll
    for (int s = 0; s < SN; ++s) {
        A2[s] = p.A[c * kN + ng * SN + s] * kLog2e;
        h[s] = 0.f;
    }
    float dsum = 0.f;  // running sum of softplus(delta) over this lane's steps: prod a = exp2(A2 * sum)

    const int nblk = (L + BK - 1) / BK;
    const int nst = (L + TT - 1) / TT;
    const int n_ref = (L + SS2D_REF_CHUNK - 1) / SS2D_REF_CHUNK;
    float *ck = p.ckpt && active ? p.ckpt + ((int64_t)b * nblk * p.dim + c) * kN + ng * SN : nullptr;
    const int64_t ck_step = p.dim * kN;

    // rows of a ragged last tile are never copied: clear the stage buffers once
    for (int i = threadIdx.x; i < NSTAGE * SM::stage_bytes / 16; i += NT) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto issue = [&](int st) {
        const int off = (st % NSTAGE) * SM::stage_bytes, t0 = st * TT;
        if constexpr (FAST) {
            // the last stage may be short (L % TT != 0): whole 16-byte pieces past L read the next row / batch, or run
            // off the tensor — copy it with the bounds-checked stagers instead
            if (t0 + TT <= L) cl.issue(off);
            else {
                if constexpr (CROSS) st_u.issue_elems(t0, L, off, rev ? L : 0);
                else st_u.issue(t0, L, off);
                st_d.issue(t0, L, off); st_B.issue(t0, L, off); st_C.issue(t0, L, off);
            }
        } else {
            st_u.issue(t0, L, off);
            st_d.issue(t0, L, off);
            st_B.issue(t0, L, off);
            st_C.issue(t0, L, off);
        }
        cp_async_commit();
    };
    // this lane's OWN steps of block `blk` (tile of stage `sbuf`, block kb inside it): softplus, delta*u -> exchange buffer
    const int own_row = (warp * CPW + cw) * SM::RSU, own_off = own_row + OWN * ng;
    auto prepare = [&](const unsigned char *sbuf, int kb, int blk, float (&uv)[OWN], float &dl_sum) {
        float dv[OWN], dl[OWN], du[OWN];
        lds_k<in_t, OWN>(reinterpret_cast<const in_t *>(sbuf + SM::d_off) + own_off + kb * BK, dv);
        if constexpr (CROSS) lds_own<in_t, OWN>(reinterpret_cast<const in_t *>(sbuf + SM::u_off) + own_row, OWN * ng + kb * BK, rev, uv);
        else lds_k<in_t, OWN>(reinterpret_cast<const in_t *>(sbuf + SM::u_off) + own_off + kb * BK, uv);
        const int valid = L - (blk * BK + OWN * ng);
        dl_sum = 0.f;
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            float d = dv[i] + bias;
            const float sp = softplus_ref(d);  // evaluated unconditionally: a select, not a branch, keeps one basic block
            d = softplus ? sp : d;
            dl[i] = (FAST || i < valid) ? d : 0.f;  // identity element past the end (fwd_kernel_oflex.cuh:146-150)
            du[i] = dl[i] * uv[i];
            dl_sum += dl[i];
        }
        float *dst = xpub + (blk & 1) * (XBUF * 4);
        if constexpr (OWN == 4) {
            *reinterpret_cast<float4 *>(dst) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            *reinterpret_cast<float4 *>(dst + NQ * CPW * 4) = make_float4(du[0], du[1], du[2], du[3]);
        } else {
            *reinterpret_cast<float2 *>(dst) = make_float2(dl[0], dl[1]);
            *reinterpret_cast<float2 *>(dst + NQ * CPW * 4) = make_float2(du[0], du[1]);
        }
    };
    DirWalk ywalk;  // pixel of this lane's first own step of the NEXT block to be finished (fused seam only)
    if constexpr (CROSS) ywalk.init(g & 2, 1, 1, L, OWN * ng);
    // lane reduction + stores of a finished block
    auto finish = [&](const float2 (&y2)[BK / 2], const float (&uv)[OWN], int blk, bool store) {
        float y[BK], o[OWN];
#pragma unroll
        for (int j = 0; j < BK / 2; ++j) { y[2 * j] = y2[j].x; y[2 * j + 1] = y2[j].y; }
        reduce_lanes<LPC>(y, o, ng);
#pragma unroll
        for (int i = 0; i < OWN; ++i) o[i] = fmaf(Dv, uv[i], o[i]);
        const int t_own = blk * BK + OWN * ng;
        if constexpr (CROSS) {
            if (store && o_row) red_own_cross<OWN>(reinterpret_cast<float *>(o_row), ywalk, o);
            ywalk.advance(BK);
        } else if constexpr (FAST) {
            if (store && o_row) stg_k<out_t, OWN>(o_row + t_own, o, OWN, true);
        } else {
            if (store && o_row) stg_k<out_t, OWN>(o_row + t_own, o, L - t_own, fl.vec_out);
        }
        if (z_row && store) {
            float zv[OWN];
            ldg_k<in_t, OWN>(z_row + t_own, zv, L - t_own, fl.vec_z);
#pragma unroll
            for (int i = 0; i < OWN; ++i) o[i] *= zv[i] * sigmoidf_fast(zv[i]);
            stg_k<out_t, OWN>(oz_row + t_own, o, L - t_own, fl.vec_out);
        }
    };

    issue(0);
    if (nst > 1) { issue(1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    float uv_cur[OWN], uv_prev[OWN], dls_cur;
    float2 y2_prev[BK / 2];
#pragma unroll
    for (int i = 0; i < OWN; ++i) uv_prev[i] = 0.f;
#pragma unroll
    for (int j = 0; j < BK / 2; ++j) y2_prev[j] = make_float2(0.f, 0.f);
    prepare(smem, 0, 0, uv_cur, dls_cur);
    __syncwarp();

    int blk = 0;
    for (int st = 0; st < nst; ++st) {
        const unsigned char *buf = smem + (st % NSTAGE) * SM::stage_bytes;
        const unsigned char *buf_next = smem + ((st + 1) % NSTAGE) * SM::stage_bytes;
        const in_t *sB = reinterpret_cast<const in_t *>(buf + SM::B_off) + ng * SM::RSB;  // + s*LPC*RSB: state ng*SN+s
        const in_t *sC = reinterpret_cast<const in_t *>(buf + SM::C_off) + ng * SM::RSB;
        const int nkb = min(BPS, nblk - st * BPS);
#pragma unroll
        for (int kb = 0; kb < BPS; ++kb) {
            if (kb >= nkb) break;
            if (kb == BPS - 1 && blk + 1 < nblk) {
                // the next block opens stage st+1: it must have landed, and everybody must be done with stage st-1
                // before its buffer is refilled with stage st+2
                cp_async_wait<0>();
                __syncthreads();
                if (st + 2 < nst) issue(st + 2);
            }
            dsum += dls_cur;
            // ---- 16 steps x SN states of block blk ----
            float2 y2[BK / 2];
#pragma unroll
            for (int j = 0; j < BK / 2; ++j) y2[j] = make_float2(0.f, 0.f);
            const float4 *xc = xq + (blk & 1) * XBUF;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 dlq = xc[q * CPW], duq = xc[(NQ + q) * CPW];
                const float2 dl0 = make_float2(dlq.x, dlq.y), dl1 = make_float2(dlq.z, dlq.w);
                const float2 du0 = make_float2(duq.x, duq.y), du1 = make_float2(duq.z, duq.w);
#pragma unroll
                for (int s = 0; s < SN; ++s) {
                    float Bv[4], Cv[4];
                    lds_k<in_t, 4>(sB + s * LPC * SM::RSB + kb * BK + 4 * q, Bv);
                    lds_k<in_t, 4>(sC + s * LPC * SM::RSB + kb * BK + 4 * q, Cv);
                    const float2 A2d = make_float2(A2[s], A2[s]);
                    const float2 e0 = __fmul2_rn(dl0, A2d), e1 = __fmul2_rn(dl1, A2d);
                    const float2 a0 = make_float2(ex2(e0.x), ex2(e0.y)), a1 = make_float2(ex2(e1.x), ex2(e1.y));
                    const float2 b0 = __fmul2_rn(du0, make_float2(Bv[0], Bv[1])), b1 = __fmul2_rn(du1, make_float2(Bv[2], Bv[3]));
                    float2 h0, h1;
                    h0.x = fmaf(a0.x, h[s], b0.x);
                    h0.y = fmaf(a0.y, h0.x, b0.y);
                    h1.x = fmaf(a1.x, h0.y, b1.x);
                    h1.y = fmaf(a1.y, h1.x, b1.y);
                    h[s] = h1.y;
                    y2[2 * q] = __ffma2_rn(make_float2(Cv[0], Cv[1]), h0, y2[2 * q]);
                    y2[2 * q + 1] = __ffma2_rn(make_float2(Cv[2], Cv[3]), h1, y2[2 * q + 1]);
                }
            }
            if (ck) {
                if constexpr (SN == 4) *reinterpret_cast<float4 *>(ck) = make_float4(h[0], h[1], h[2], h[3]);
                else *reinterpret_cast<float2 *>(ck) = make_float2(h[0], h[1]);
                ck += ck_step;
            }
            // ---- block blk-1: sum over the 16 states, stores (independent of the recurrence above) ----
            if constexpr (CROSS) { if (blk == 0) ywalk.advance(-BK); }
            finish(y2_prev, uv_prev, blk - 1, active && blk > 0);
            // ---- block blk+1: softplus of this lane's own steps (the values of a block past the end are never used) ----
            float uv_next[OWN], dls_next;
            prepare(kb + 1 < BPS ? buf : buf_next, kb + 1 < BPS ? kb + 1 : 0, blk + 1, uv_next, dls_next);
            // ---- the reference's checkpoint tensor x: (running prod a, h) at the end of every 2048-step chunk ----
            const int t_end = min(L, blk * BK + BK);
            if (p.x && (t_end % SS2D_REF_CHUNK == 0 || t_end == L)) {
                float tot = dsum;
#pragma unroll
                for (int m = 1; m < LPC; m <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, m);
                if (active) {
                    float *dst = p.x + ((((int64_t)b * p.dim + c) * n_ref + (t_end - 1) / SS2D_REF_CHUNK) * kN + ng * SN) * 2;
#pragma unroll
                    for (int s = 0; s < SN; s += 2)
                        *reinterpret_cast<float4 *>(dst + 2 * s) = make_float4(ex2(A2[s] * tot), h[s], ex2(A2[s + 1] * tot), h[s + 1]);
                }
            }
#pragma unroll
            for (int j = 0; j < BK / 2; ++j) y2_prev[j] = y2[j];
#pragma unroll
            for (int i = 0; i < OWN; ++i) { uv_prev[i] = uv_cur[i]; uv_cur[i] = uv_next[i]; }
            dls_cur = dls_next;
            ++blk;
            __syncwarp();  // block blk+1's exchange buffer is complete; everyone is done reading block blk's
        }
    }
    finish(y2_prev, uv_prev, nblk - 1, active);
}

// two entry points over one body: the fused variant carries the gather state and is capped at 168 registers (3 CTAs per
// SM); the plain one is left to ptxas (137 registers) — an explicit minimum-blocks bound makes it spend all it is given
template <typename in_t, typename out_t, int SN, int NW, int TT, bool FAST>
__global__ void __launch_bounds__(NW *kWarp, sizeof(in_t) == 2 ? 3 : 0)  // 16-bit inputs: cap at 168 registers (3 CTAs/SM)
sl_fwd_kernel(const ss2d_scan_fwd_params p, const int tiles_per_group, const Flags fl) {
    sl_fwd_body<in_t, out_t, SN, NW, TT, FAST, false>(p, tiles_per_group, fl, CrossAux{nullptr, nullptr, nullptr});
}
template <int SN, int NW, int TT>
__global__ void __launch_bounds__(NW *kWarp, 3)
sl_fwd_cross_kernel(const ss2d_scan_fwd_params p, const int tiles_per_group, const Flags fl, const CrossAux aux) {
    sl_fwd_body<float, float, SN, NW, TT, true, true>(p, tiles_per_group, fl, aux);
}

// process-wide meaning of SS2D_FAMILY_AUTO (test hook ss2d_set_default_family; 0 = by problem size)
static std::atomic<int> g_default_family{SS2D_FAMILY_AUTO};
int set_default_family(int f) { return g_default_family.exchange(f >= 0 && f <= 2 ? f : 0); }

bool supported(const ss2d_scan_fwd_params &p) {
    if (p.dstate != kN || p.seqlen > (1LL << 28)) return false;  // step positions and tail byte counts are 32-bit
    int fam = p.family;
    if (fam != SS2D_FAMILY_STATELANES && fam != SS2D_FAMILY_WARPSCAN) fam = g_default_family.load(std::memory_order_relaxed);
    if (fam == SS2D_FAMILY_WARPSCAN) return false;
    if (fam == SS2D_FAMILY_STATELANES) return true;
    // a state-lanes warp walks its whole sequence serially, so the launch takes ~L/4096 * 240 us however few channels
    // there are; the warp-scan kernels scale with the work and win below ~4.6 k channels (measured: B=1 x 768 channels,
    // L=19200: 369 us vs 646 us forward; B=8 x 768, L=4096: 305 us vs 237 us)
    return p.batch * p.dim >= 4608;
}

int states_per_lane(const ss2d_scan_fwd_params &p) {
    static const int sms = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    // 4 states per lane costs the fewest instructions; 2 states per lane doubles the warps when the batch is too small
    // to put ~4 warps on every SM sub-partition
    const int64_t warps4 = p.batch * p.dim / 8;
    return warps4 >= (int64_t)sms * 16 ? 4 : 2;
}

static void fill_flags(const ss2d_scan_fwd_params &p, int64_t ei, int64_t eo, Flags &fl) {
    fl.vec_u = aligned16(p.u) && (p.u_bstride * ei) % 16 == 0 && (p.u_dstride * ei) % 16 == 0;
    fl.vec_delta = aligned16(p.delta) && (p.delta_bstride * ei) % 16 == 0 && (p.delta_dstride * ei) % 16 == 0;
    fl.vec_bc = aligned16(p.B) && aligned16(p.C) && (p.B_bstride * ei) % 16 == 0 && (p.B_gstride * ei) % 16 == 0 &&
                (p.B_nstride * ei) % 16 == 0 && (p.C_bstride * ei) % 16 == 0 && (p.C_gstride * ei) % 16 == 0 &&
                (p.C_nstride * ei) % 16 == 0;
    fl.vec_out = (!p.out || aligned16(p.out)) && (!p.out_z || aligned16(p.out_z)) && (p.out_bstride * eo) % 16 == 0 &&
                 (p.out_dstride * eo) % 16 == 0;
    fl.vec_z = p.z && aligned16(p.z) && (p.z_bstride * ei) % 16 == 0 && (p.z_dstride * ei) % 16 == 0;
}

template <typename in_t, typename out_t, int SN, int NW = 4, int TT = 64, bool CROSS = false>
static int launch_fwd_t(const ss2d_scan_fwd_params &p, cudaStream_t stream, CrossAux xi = CrossAux{nullptr, nullptr, nullptr}) {
    using SM = FwdSmem<in_t, SN, NW, TT>;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tiles = (per_g + SM::CPC - 1) / SM::CPC;
    Flags fl{};
    fill_flags(p, sizeof(in_t), sizeof(out_t), fl);
    if ((p.x && !aligned16(p.x)) || (p.ckpt && !aligned16(p.ckpt))) return SS2D_ESTRIDE;  // written with 16-byte stores
    const int64_t grid = p.batch * p.ngroups * tiles;
    const bool fast = fl.vec_u && fl.vec_delta && fl.vec_bc && fl.vec_out && p.seqlen % BK == 0 && !p.z;
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)SM::total);
        if (rc != 0) return rc;
        kern<<<(unsigned)grid, NW * kWarp, SM::total, stream>>>(p, tiles, fl);
        return (int)cudaGetLastError();
    };
    if constexpr (CROSS) {
        if (!fast) return SS2D_ESTRIDE;  // covered problem (cross_covered) whose y is not 16-byte aligned
        auto kern = sl_fwd_cross_kernel<SN, NW, TT>;
        const int rc = smem_optin(kern, (int)SM::total);
        if (rc != 0) return rc;
        kern<<<(unsigned)grid, NW * kWarp, SM::total, stream>>>(p, tiles, fl, xi);
        return (int)cudaGetLastError();
    } else {
        return fast ? go(sl_fwd_kernel<in_t, out_t, SN, NW, TT, true>) : go(sl_fwd_kernel<in_t, out_t, SN, NW, TT, false>);
    }
}

template <typename in_t, typename out_t> static int launch_fwd_sn(const ss2d_scan_fwd_params &p, cudaStream_t s) {
    // tile of 32 steps with 4 states per lane (3 CTAs/SM by shared memory), 64 steps with 2 (fewer stage barriers)
    if (states_per_lane(p) == 4) return launch_fwd_t<in_t, out_t, 4, 4, 32>(p, s);
    return launch_fwd_t<in_t, out_t, 2, 4, 64>(p, s);
}

int launch_fwd(const ss2d_scan_fwd_params &p, cudaStream_t s) {
    switch (p.in_dtype) {
        case SS2D_F32: return launch_fwd_sn<float, float>(p, s);
        case SS2D_F16:
            return p.out_dtype == SS2D_F32 ? launch_fwd_sn<__half, float>(p, s) : launch_fwd_sn<__half, __half>(p, s);
        case SS2D_BF16:
            return p.out_dtype == SS2D_F32 ? launch_fwd_sn<__nv_bfloat16, float>(p, s)
                                           : launch_fwd_sn<__nv_bfloat16, __nv_bfloat16>(p, s);
        default: return SS2D_EDTYPE;
    }
}

bool cross_covered(const ss2d_scan_fwd_params &p) {
    if (!supported(p) || p.in_dtype != SS2D_F32 || p.out_dtype != SS2D_F32 || p.z || p.seqlen % BK != 0) return false;
    Flags fl{};
    ss2d_scan_fwd_params q = p;
    q.out = nullptr; q.out_z = nullptr;  // the rule must not depend on buffers only one of the two passes sees
    fill_flags(q, 4, 4, fl);
    return fl.vec_u && fl.vec_delta && fl.vec_bc && (p.out_bstride * 4) % 16 == 0 && (p.out_dstride * 4) % 16 == 0;
}

int launch_cross_fwd(const ss2d_scan_fwd_params &p, const CrossAux &xi, cudaStream_t s) {
    if (!aligned16(xi.uT) || (xi.accT && !aligned16(xi.accT))) return SS2D_ESTRIDE;
    return states_per_lane(p) == 4 ? launch_fwd_t<float, float, 4, 4, 32, true>(p, s, xi) : launch_fwd_t<float, float, 2, 4, 64, true>(p, s, xi);
}

}  // namespace sl
}  // namespace ss2d
