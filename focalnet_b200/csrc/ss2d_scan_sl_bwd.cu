// ss2d_scan_sl_bwd.cu — selective-scan backward, "state-lanes" organisation (dstate == 16), sm_100a.
//
// Replaces selective_scan_bwd_kernel + launcher (reference: kernels/selective_scan/csrc/selective_scan/cusoflex/
// selective_scan_bwd_kernel_oflex.cuh:73-322).  Gradient formulas as in ss2d_scan_bwd.cu:
//     dx_t   = dout_t C_t + a_{t+1} dx_{t+1}
//     dC_t  += dout_t h_t                dB_t += dx_t dl_t u_t               (summed over the group's channels)
//     du_t   = D dout_t + dl_t sum_n dx_t B_t
//     ddl_t  = u_t sum_n dx_t B_t + sum_n A_n dx_t (a_t h_{t-1})             a_t h_{t-1} = h_t - dl_t u_t B_t
//     dA_n  += sum_t dl_t dx_t (a_t h_{t-1})      dD += sum_t dout_t u_t     ddelta_t = ddl_t * softplus'(delta_t+bias)
// Organisation (ss2d_scan_sl.cuh): a lane owns SN states of one channel and walks the sequence right-to-left in
// blocks of 16 steps.  Per block: (F) recompute a_t, h_t of the 16 steps from the forward's checkpoint at the
// block's left edge, keeping them in registers; (R) walk the block backwards — the dx recurrence is a serial
// FFMA chain, everything else is evaluated on step PAIRS with FMUL2 / FFMA2.  One ex2 per (element, state) in the
// whole backward; no warp scan.  The per-channel dB / dC products are summed over the warp's channels by a transposed
// shuffle reduction, the warp sums are staged in shared memory (double buffered, one block barrier per 16 steps) and
// summed over the CTA's warps before ONE red.global.add.v4.f32 per (state, 4 steps) leaves the CTA (the reference
// issues one scalar atomic per channel per element).  Sums over the 16 states (du, ddelta) are transposed reductions
// over the channel's lanes.
#include "ss2d_scan_sl.cuh"

namespace ss2d {
namespace sl {

// shared-memory carve-up of one backward CTA.  A pipeline stage is ONE block of 16 steps (3-deep ring: block k in
// use, block k-1 being pre-processed, block k-2 in flight).
template <typename in_t, typename out_t, int SN, int NW> struct BwdSmem {
    using M = Map<SN>;
    static constexpr int NSTAGE = 3;
    static constexpr int CPC = NW * M::CPW;
    static constexpr int NQ = BK / 4;
    static constexpr int RSU = BK + 64 / (int)sizeof(in_t);
    static constexpr int RSG = BK + 64 / (int)sizeof(out_t);
    static constexpr int RSB = BK + 16 / (int)sizeof(in_t);
    static constexpr int u_off = 0;
    static constexpr int d_off = u_off + CPC * RSU * (int)sizeof(in_t);
    static constexpr int g_off = d_off + CPC * RSU * (int)sizeof(in_t);
    static constexpr int B_off = g_off + CPC * RSG * (int)sizeof(out_t);
    static constexpr int C_off = B_off + kN * RSB * (int)sizeof(in_t);
    static constexpr int stage_bytes = C_off + kN * RSB * (int)sizeof(in_t);
    static_assert(stage_bytes % 16 == 0 && d_off % 16 == 0 && g_off % 16 == 0 && B_off % 16 == 0 && C_off % 16 == 0, "alignment");
    static constexpr int xch_off = NSTAGE * stage_bytes;            // [warp][buf][dl|du|go][q][cw] float4
    static constexpr int XBUF = 3 * NQ * M::CPW;                    // float4 per buffer
    static constexpr int xch_warp = 2 * XBUF * 16;
    static constexpr int red_off = xch_off + NW * xch_warp;        // [buf][warp][q][lane] float4: dB/dC summed over the warp's channels
    static constexpr int NIDX = 2 * SN * NQ;
    static constexpr int red_warp = NQ * kWarp * 16;
    static constexpr int red_buf = NW * red_warp;
    static constexpr int total = red_off + 2 * red_buf;
    static_assert(NIDX * M::LPC == 128, "one reduction output per (which, state, q)");
    static_assert(2 * SN == M::CPW, "the lane reduction over a warp's channels leaves one (which, state) per channel slot");
};

// Iteration k of the block loop (k = nblk-1 .. 0), software pipelined by hand:
//   [ sum + red.global of block k+1's staged dB/dC | du, ddelta of block k+1 | softplus of block k-1 | F(k) ]
//   [ R(k): dx chain, gradient products; dB/dC summed over the warp's channels by shuffles, staged (double buffered) ]
//   barrier  (staging complete; the tile of block k-1 has landed)
// CROSS (fused seam S3, FAST only): u and dout are read from the fp32 planes x[b,d] / dy[b,d] (directions 0 / 2,
// forwards / backwards) or from their transposed copies aux.uT / aux.doutT (directions 1 / 3); du is accumulated
// (red.global.add) into dx[b,d] resp. dx^T (aux.accT) — CrossMerge.backward and CrossScan.backward as load / store
// addressing over contiguous runs.  ddelta, dB, dC stay in scan order.
template <typename in_t, typename out_t, int SN, int NW, bool FAST, bool CROSS = false>
__global__ void __launch_bounds__(NW *kWarp, 3)
sl_bwd_kernel(const ss2d_scan_bwd_params pb, const int tiles_per_group, const Flags fl, const CrossAux aux) {
    static_assert(!CROSS || (FAST && sizeof(in_t) == 4 && sizeof(out_t) == 4), "fused seam: fp32, aligned, L % 16 == 0");
    using M = Map<SN>;
    using SM = BwdSmem<in_t, out_t, SN, NW>;
    constexpr int NT = NW * kWarp, CPC = SM::CPC, LPC = M::LPC, CPW = M::CPW, OWN = M::OWN, NQ = SM::NQ, NIDX = SM::NIDX;
    constexpr int NSTAGE = SM::NSTAGE, XBUF = SM::XBUF;
    constexpr int OPT = 128 / NT;  // reduction outputs per thread (128 per block and CTA)
    static_assert(NT <= 128 && 128 % NT == 0, "at most 4 warps per CTA");
    extern __shared__ __align__(16) unsigned char smem[];
    const ss2d_scan_fwd_params &p = pb.f;

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

    float *xch = reinterpret_cast<float *>(smem + SM::xch_off + warp * SM::xch_warp);
    // element (kind, step j, channel) of buffer f lives at float index (((f*3 + kind)*NQ + j/4)*CPW + cw)*4 + j%4
    float *xpub = xch + (((OWN * ng) >> 2) * CPW + cw) * 4 + ((OWN * ng) & 3);
    const float4 *xq = reinterpret_cast<const float4 *>(xch) + cw;
    float4 *red_w = reinterpret_cast<float4 *>(smem + SM::red_off + warp * SM::red_warp) + lane;  // + q*32 (+ buffer)
    const float4 *red_all = reinterpret_cast<const float4 *>(smem + SM::red_off);
    constexpr int RBUF = SM::red_buf / 16;  // float4 per staging buffer

    const int64_t row0 = (int64_t)g * per_g + ch0;
    const int64_t urow0 = CROSS ? ch0 : row0;  // row of u / dout / du: d in fused mode
    RowStager<in_t, BK, NT> st_u, st_d, st_B, st_C;
    RowStager<out_t, BK, NT> st_g;
    st_u.init(reinterpret_cast<in_t *>(smem + SM::u_off), SM::RSU, reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + urow0 * p.u_dstride,
              p.u_dstride, CPC, rows_valid, fl.vec_u, 0);
    st_d.init(reinterpret_cast<in_t *>(smem + SM::d_off), SM::RSU,
              reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + row0 * p.delta_dstride, p.delta_dstride, CPC, rows_valid,
              fl.vec_delta, 0);
    st_g.init(reinterpret_cast<out_t *>(smem + SM::g_off), SM::RSG,
              reinterpret_cast<const out_t *>(pb.dout) + b * pb.dout_bstride + urow0 * pb.dout_dstride, pb.dout_dstride, CPC, rows_valid,
              fl.vec_dout, 0);
    st_B.init(reinterpret_cast<in_t *>(smem + SM::B_off), SM::RSB, reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride,
              p.B_nstride, kN, kN, fl.vec_bc, SN);
    st_C.init(reinterpret_cast<in_t *>(smem + SM::C_off), SM::RSB, reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride,
              p.C_nstride, kN, kN, fl.vec_bc, SN);

    // FAST: one flattened copy list instead of the five stagers
    // (fused seam: u / dout rows are rows of x / dy, of their transposed copies, or the same rows walked backwards — the
    // mirrored 16-byte pieces of CopyList::add, un-mirrored by lds_own; L % 16 == 0 there, so no stage is ever short)
    constexpr int NPIECE = (2 * CPC + 2 * kN) * (BK * (int)sizeof(in_t) / 16) + CPC * (BK * (int)sizeof(out_t) / 16);
    constexpr int NCOPY = (NPIECE + NT - 1) / NT;
    CopyList<NT, NCOPY> cl;
    const bool rev = CROSS && (g & 2);
    if constexpr (FAST) {
        const int t_first = ((int)((p.seqlen + BK - 1) / BK) - 1) * BK;
        const int mirror_L = rev ? (int)p.seqlen : 0;
        cl.clear();
        if constexpr (CROSS)
            cl.add(reinterpret_cast<float *>(smem + SM::u_off), SM::RSU,
                   ((g & 1) ? aux.uT : reinterpret_cast<const float *>(p.u)) + b * p.u_bstride + urow0 * p.u_dstride, p.u_dstride, CPC,
                   rows_valid, BK, 0, t_first, -BK, mirror_L);
        else
            cl.add(reinterpret_cast<in_t *>(smem + SM::u_off), SM::RSU, reinterpret_cast<const in_t *>(p.u) + b * p.u_bstride + row0 * p.u_dstride,
                   p.u_dstride, CPC, rows_valid, BK, 0, t_first, -BK);
        cl.add(reinterpret_cast<in_t *>(smem + SM::d_off), SM::RSU,
               reinterpret_cast<const in_t *>(p.delta) + b * p.delta_bstride + row0 * p.delta_dstride, p.delta_dstride, CPC, rows_valid, BK, 0,
               t_first, -BK);
        if constexpr (CROSS)
            cl.add(reinterpret_cast<float *>(smem + SM::g_off), SM::RSG,
                   ((g & 1) ? aux.doutT : reinterpret_cast<const float *>(pb.dout)) + b * pb.dout_bstride + urow0 * pb.dout_dstride,
                   pb.dout_dstride, CPC, rows_valid, BK, 0, t_first, -BK, mirror_L);
        else
            cl.add(reinterpret_cast<out_t *>(smem + SM::g_off), SM::RSG,
                   reinterpret_cast<const out_t *>(pb.dout) + b * pb.dout_bstride + row0 * pb.dout_dstride, pb.dout_dstride, CPC, rows_valid, BK, 0,
                   t_first, -BK);
        cl.add(reinterpret_cast<in_t *>(smem + SM::B_off), SM::RSB, reinterpret_cast<const in_t *>(p.B) + b * p.B_bstride + g * p.B_gstride,
               p.B_nstride, kN, kN, BK, SN, t_first, -BK);
        cl.add(reinterpret_cast<in_t *>(smem + SM::C_off), SM::RSB, reinterpret_cast<const in_t *>(p.C) + b * p.C_bstride + g * p.C_gstride,
               p.C_nstride, kN, kN, BK, SN, t_first, -BK);
    }

    const in_t *z_row = p.z ? reinterpret_cast<const in_t *>(p.z) + b * p.z_bstride + c * p.z_dstride : nullptr;
    const out_t *pre_row = p.z ? reinterpret_cast<const out_t *>(p.out) + b * p.out_bstride + c * p.out_dstride : nullptr;
    const int64_t row = ((int64_t)b * p.dim + c) * L;
    in_t *du_base = reinterpret_cast<in_t *>(pb.du);
    if constexpr (CROSS) { if (g & 1) du_base = reinterpret_cast<in_t *>(aux.accT); }
    in_t *du_row = du_base + (CROSS ? ((int64_t)b * per_g + (active ? c_local : per_g - 1)) * L : row);
    in_t *dd_row = reinterpret_cast<in_t *>(pb.ddelta) + row;
    in_t *dz_row = pb.dz ? reinterpret_cast<in_t *>(pb.dz) + row : nullptr;
    const float Dv = p.D ? p.D[c] : 0.f;
    const float bias = p.delta_bias ? p.delta_bias[c] : 0.f;
    const bool softplus = p.delta_softplus != 0;

    float A2[SN], An[SN], dx[SN], anext[SN];
    float2 dA2[SN];
#pragma unroll
    for (int s = 0; s < SN; ++s) {
        An[s] = p.A[c * kN + ng * SN + s];
        A2[s] = An[s] * kLog2e;
        dx[s] = 0.f; anext[s] = 1.f; dA2[s] = make_float2(0.f, 0.f);
    }
    float dD_acc = 0.f, dbias_acc = 0.f;

    const int nblk = (L + BK - 1) / BK;
    const int64_t ck_step = p.dim * kN;
    // checkpoint j = h at the right edge of block j; block k restarts from checkpoint k-1
    const float *ck = p.ckpt + ((int64_t)b * nblk * p.dim + c) * kN + ng * SN;

    // this thread's dB / dC reduction outputs: o = idx * LPC + ngo with idx = (which*SN + s)*NQ + q
    float *red_dst[OPT];
    int red_src[OPT], red_q[OPT];
#pragma unroll
    for (int i = 0; i < OPT; ++i) {
        const int o = i * NT + threadIdx.x;
        const int idx = o / LPC, ngo = o % LPC;
        const int which = idx / (SN * NQ), s = (idx / NQ) % SN, q = idx % NQ;
        // after the lane reduction the sum of (which, s) over a warp's channels sits in channel slot which*SN + s
        red_src[i] = q * kWarp + (which * SN + s) * LPC + ngo;
        red_q[i] = q;
        red_dst[i] = (which ? pb.dC : pb.dB) + (((int64_t)b * p.ngroups + g) * kN + ngo * SN + s) * (int64_t)L + 4 * q;
    }

    for (int i = threadIdx.x; i < NSTAGE * SM::stage_bytes / 16; i += NT) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto issue = [&](int k, int slot_off) {  // tile of block k -> ring slot at slot_off (an empty group when k < 0)
        if (k >= 0) {
            if constexpr (FAST) {
                cl.issue(slot_off);
            } else {
                const int t0 = k * BK;
                st_u.issue(t0, L, slot_off);
                st_d.issue(t0, L, slot_off);
                st_g.issue(t0, L, slot_off);
                st_B.issue(t0, L, slot_off);
                st_C.issue(t0, L, slot_off);
            }
        }
        cp_async_commit();
    };
    const float *ck_ptr = ck + (int64_t)(nblk - 2) * ck_step;  // walks left one block per load
    auto load_ck = [&](bool exists) {  // next checkpoint to the left (zero left of the sequence)
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (exists) {
            if constexpr (SN == 4) v = __ldg(reinterpret_cast<const float4 *>(ck_ptr));
            else { const float2 t = __ldg(reinterpret_cast<const float2 *>(ck_ptr)); v.x = t.x; v.y = t.y; }
        }
        ck_ptr -= ck_step;
        return v;
    };
    // this lane's OWN steps of block k: softplus, delta*u, (gated) dout -> exchange buffer k & 1
    const int row_u = (warp * CPW + cw) * SM::RSU, row_g = (warp * CPW + cw) * SM::RSG, own_u = row_u + OWN * ng;
    auto prepare = [&](int k, int slot_off, int xoff, float (&uv)[OWN], float (&dl)[OWN], float (&gv)[OWN]) {
        const unsigned char *sbuf = smem + slot_off;
        float dv[OWN], du[OWN];
        lds_k<in_t, OWN>(reinterpret_cast<const in_t *>(sbuf + SM::d_off) + own_u, dv);
        lds_own<in_t, OWN>(reinterpret_cast<const in_t *>(sbuf + SM::u_off) + row_u, OWN * ng, rev, uv);
        lds_own<out_t, OWN>(reinterpret_cast<const out_t *>(sbuf + SM::g_off) + row_g, OWN * ng, rev, gv);
        const int t_own = k * BK + OWN * ng;
        const int valid = active ? L - t_own : 0;  // may be <= 0 or > OWN
        if (z_row && k >= 0) {  // out = pre * silu(z): dz and the gated upstream gradient
            float zv[OWN], pre[OWN];
            ldg_k<in_t, OWN>(z_row + t_own, zv, valid, fl.vec_z);
            ldg_k<out_t, OWN>(pre_row + t_own, pre, valid, fl.vec_out);
#pragma unroll
            for (int i = 0; i < OWN; ++i) {
                const float sgm = sigmoidf_fast(zv[i]);
                pre[i] = gv[i] * pre[i] * sgm * (1.f + zv[i] * (1.f - sgm));
                gv[i] *= zv[i] * sgm;
            }
            stg_k<in_t, OWN>(dz_row + t_own, pre, valid, fl.vec_grad);
        }
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            float d = dv[i] + bias;
            const float sp = softplus_ref(d);  // evaluated unconditionally: a select, not a branch
            d = softplus ? sp : d;
            const bool in = FAST ? active : i < valid;
            dl[i] = in ? d : 0.f;
            gv[i] = in ? gv[i] : 0.f;
            du[i] = dl[i] * uv[i];
        }
        float *dst = xpub + xoff * 4;
        if constexpr (OWN == 4) {
            *reinterpret_cast<float4 *>(dst) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            *reinterpret_cast<float4 *>(dst + NQ * CPW * 4) = make_float4(du[0], du[1], du[2], du[3]);
            *reinterpret_cast<float4 *>(dst + 2 * NQ * CPW * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        } else {
            *reinterpret_cast<float2 *>(dst) = make_float2(dl[0], dl[1]);
            *reinterpret_cast<float2 *>(dst + NQ * CPW * 4) = make_float2(du[0], du[1]);
            *reinterpret_cast<float2 *>(dst + 2 * NQ * CPW * 4) = make_float2(gv[0], gv[1]);
        }
    };
    // dB / dC of a finished block: sum the staged products over the CTA's channels, one vector reduction per (state, 4 steps)
#pragma unroll
    for (int i = 0; i < OPT; ++i) red_dst[i] += (int64_t)nblk * BK;  // one block right of the last: walks left per call
    auto reduce_bc = [&](int k, int buf) {
#pragma unroll
        for (int i = 0; i < OPT; ++i) {
            float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float4 v = red_all[buf + w * (NQ * kWarp) + red_src[i]];
                lo = __fadd2_rn(lo, make_float2(v.x, v.y));
                hi = __fadd2_rn(hi, make_float2(v.z, v.w));
            }
            red_dst[i] -= BK;
            float *dst = red_dst[i];
            if (FAST) {
                red_add_v4(dst, lo.x, lo.y, hi.x, hi.y);
            } else {
                const int rem = L - (k * BK + 4 * red_q[i]);
                if (fl.vec_dbc && rem >= 4) {
                    red_add_v4(dst, lo.x, lo.y, hi.x, hi.y);
                } else {
                    if (rem > 0) atomicAdd(dst + 0, lo.x);
                    if (rem > 1) atomicAdd(dst + 1, lo.y);
                    if (rem > 2) atomicAdd(dst + 2, hi.x);
                    if (rem > 3) atomicAdd(dst + 3, hi.y);
                }
            }
        }
    };
    // du, ddelta of a finished block: sums over the 16 states, every lane finishes its own steps
    DirWalk xwalk;  // pixel of this lane's first own step of the next block to be finalised (fused seam only)
    if constexpr (CROSS) xwalk.init(g & 2, 1, 1, L, nblk * BK + OWN * ng);
    in_t *du_ptr = du_row + (int64_t)nblk * BK + OWN * ng, *dd_ptr = dd_row + (int64_t)nblk * BK + OWN * ng;  // walk left per call
    auto finalize = [&](int k, const float2 (&sacc)[BK / 2], const float2 (&wacc)[BK / 2], const float (&uv)[OWN], const float (&dl)[OWN],
                        const float (&gv)[OWN], bool store) {
        float sv[BK], wv[BK], s4[OWN], w4[OWN];
#pragma unroll
        for (int j = 0; j < BK / 2; ++j) {
            sv[2 * j] = sacc[j].x; sv[2 * j + 1] = sacc[j].y;
            wv[2 * j] = wacc[j].x; wv[2 * j + 1] = wacc[j].y;
        }
        reduce_lanes<LPC>(sv, s4, ng);
        reduce_lanes<LPC>(wv, w4, ng);
        float duo[OWN], ddo[OWN];
        const int t_own = k * BK + OWN * ng;
        const int valid = L - t_own;
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            const float v = fmaf(uv[i], s4[i], w4[i]);
            // softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x)); exact 1 beyond the x > 20 cut-off
            const float sgm = softplus ? 1.f - ex2(-kLog2e * dl[i]) : 1.f;
            ddo[i] = v * sgm;
            dbias_acc += (FAST || i < valid) ? ddo[i] : 0.f;
            duo[i] = fmaf(dl[i], s4[i], Dv * gv[i]);
            dD_acc = fmaf(gv[i], uv[i], dD_acc);
        }
        if (store) {
            if constexpr (CROSS) red_own_cross<OWN>(reinterpret_cast<float *>(du_row), xwalk, duo);
            else stg_k<in_t, OWN>(du_ptr, duo, FAST ? OWN : valid, FAST ? true : fl.vec_grad);
            stg_k<in_t, OWN>(dd_ptr, ddo, FAST ? OWN : valid, FAST ? true : fl.vec_grad);
        }
        if constexpr (CROSS) xwalk.advance(-BK);
        du_ptr -= BK;
        dd_ptr -= BK;
    };

    // ---- prologue: tiles of the two rightmost blocks, softplus of the rightmost ----
    int slot0 = ((nblk - 1) % NSTAGE) * SM::stage_bytes;  // ring slots of blocks k, k-1, k-2 (byte offsets), rotated per iteration
    int slot1 = ((nblk + 1) % NSTAGE) * SM::stage_bytes;  // (k-1) mod 3 == (k+2) mod 3
    int slot2 = (nblk % NSTAGE) * SM::stage_bytes;        // (k-2) mod 3 == (k+1) mod 3
    int rcur = 0;                                         // staging buffer (float4 index) block k writes; block k+1 wrote the other
    int xcur = ((nblk - 1) & 1) * XBUF;                   // exchange buffer (float4 index) of block k; block k-1 uses the other
    issue(nblk - 1, slot0);
    issue(nblk - 2, slot1);
    cp_async_wait<0>();
    __syncthreads();
    float uv_c[OWN], dl_c[OWN], gv_c[OWN];  // own values of the block whose F/R runs in the current iteration
    float uv_p[OWN], dl_p[OWN], gv_p[OWN];  // ... of the block finished in the previous iteration
    float2 sacc[BK / 2], wacc[BK / 2];
#pragma unroll
    for (int j = 0; j < BK / 2; ++j) { sacc[j] = make_float2(0.f, 0.f); wacc[j] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < OWN; ++i) { uv_p[i] = 0.f; dl_p[i] = 0.f; gv_p[i] = 0.f; }
    prepare(nblk - 1, slot0, xcur, uv_c, dl_c, gv_c);
    float4 hin = load_ck(nblk >= 2);
    __syncthreads();

    for (int k = nblk - 1; k >= 0; --k) {
        issue(k - 2, slot2);  // ring slot of block k+1, free since the barrier that ended iteration k+1
        const float4 hin_next = load_ck(k >= 2);
        // ---- block k+1: dB/dC sums, du / ddelta (first iteration: all-zero dummies, nothing stored) ----
        if (k + 1 < nblk) reduce_bc(k + 1, rcur ^ RBUF);
        finalize(k + 1, sacc, wacc, uv_p, dl_p, gv_p, active && k + 1 < nblk);
        // ---- block k-1: softplus of this lane's own steps ----
        float uv_n[OWN], dl_n[OWN], gv_n[OWN];
        prepare(k - 1, slot1, xcur ^ XBUF, uv_n, dl_n, gv_n);
        // ---- F(k): recompute a_t and h_t of the block ----
        const unsigned char *buf = smem + slot0;
        const in_t *sB = reinterpret_cast<const in_t *>(buf + SM::B_off) + ng * SM::RSB;
        const in_t *sC = reinterpret_cast<const in_t *>(buf + SM::C_off) + ng * SM::RSB;
        const float4 *xc = xq + xcur;
        float hcur[SN];
        hcur[0] = hin.x; hcur[1] = hin.y;
        if constexpr (SN == 4) { hcur[2] = hin.z; hcur[3] = hin.w; }
        float2 a2[SN][BK / 2], H2[SN][BK / 2];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 dlq = xc[q * CPW], duq = xc[(NQ + q) * CPW];
            const float2 dl0 = make_float2(dlq.x, dlq.y), dl1 = make_float2(dlq.z, dlq.w);
            const float2 du0 = make_float2(duq.x, duq.y), du1 = make_float2(duq.z, duq.w);
#pragma unroll
            for (int s = 0; s < SN; ++s) {
                float Bv[4];
                lds_k<in_t, 4>(sB + s * LPC * SM::RSB + 4 * q, Bv);
                const float2 A2d = make_float2(A2[s], A2[s]);
                const float2 e0 = __fmul2_rn(dl0, A2d), e1 = __fmul2_rn(dl1, A2d);
                const float2 a0 = make_float2(ex2(e0.x), ex2(e0.y)), a1 = make_float2(ex2(e1.x), ex2(e1.y));
                const float2 b0 = __fmul2_rn(du0, make_float2(Bv[0], Bv[1])), b1 = __fmul2_rn(du1, make_float2(Bv[2], Bv[3]));
                float2 h0, h1;
                h0.x = fmaf(a0.x, hcur[s], b0.x);
                h0.y = fmaf(a0.y, h0.x, b0.y);
                h1.x = fmaf(a1.x, h0.y, b1.x);
                h1.y = fmaf(a1.y, h1.x, b1.y);
                hcur[s] = h1.y;
                a2[s][2 * q] = a0; a2[s][2 * q + 1] = a1;
                H2[s][2 * q] = h0; H2[s][2 * q + 1] = h1;
            }
        }
        // ---- R(k): walk the block right to left (stages into the buffer block k+1 did not use: no barrier needed here) ----
#pragma unroll
        for (int j = 0; j < BK / 2; ++j) { sacc[j] = make_float2(0.f, 0.f); wacc[j] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int q = NQ - 1; q >= 0; --q) {
            const float4 dlq = xc[q * CPW], duq = xc[(NQ + q) * CPW], goq = xc[(2 * NQ + q) * CPW];
            const float2 dl0 = make_float2(dlq.x, dlq.y), dl1 = make_float2(dlq.z, dlq.w);
            const float2 du0 = make_float2(duq.x, duq.y), du1 = make_float2(duq.z, duq.w);
            const float2 nu0 = make_float2(-duq.x, -duq.y), nu1 = make_float2(-duq.z, -duq.w);
            const float2 go0 = make_float2(goq.x, goq.y), go1 = make_float2(goq.z, goq.w);
            float V[CPW][4];  // this lane's dB (slots 0..SN-1) and dC (slots SN..2SN-1) products of the 4 steps
#pragma unroll
            for (int s = 0; s < SN; ++s) {
                float Bv[4], Cv[4];
                lds_k<in_t, 4>(sB + s * LPC * SM::RSB + 4 * q, Bv);
                lds_k<in_t, 4>(sC + s * LPC * SM::RSB + 4 * q, Cv);
                const float2 B0 = make_float2(Bv[0], Bv[1]), B1 = make_float2(Bv[2], Bv[3]);
                const float2 gc0 = __fmul2_rn(go0, make_float2(Cv[0], Cv[1])), gc1 = __fmul2_rn(go1, make_float2(Cv[2], Cv[3]));
                const float2 h0 = H2[s][2 * q], h1 = H2[s][2 * q + 1];
                const float2 a0 = a2[s][2 * q], a1 = a2[s][2 * q + 1];
                float2 x0, x1;  // dx of the 4 steps
                x1.y = fmaf(anext[s], dx[s], gc1.y);
                x1.x = fmaf(a1.y, x1.y, gc1.x);
                x0.y = fmaf(a1.x, x1.x, gc0.y);
                x0.x = fmaf(a0.y, x0.y, gc0.x);
                dx[s] = x0.x;
                anext[s] = a0.x;
                const float2 dC0 = __fmul2_rn(go0, h0), dC1 = __fmul2_rn(go1, h1);
                const float2 dB0 = __fmul2_rn(x0, du0), dB1 = __fmul2_rn(x1, du1);
                sacc[2 * q] = __ffma2_rn(x0, B0, sacc[2 * q]);
                sacc[2 * q + 1] = __ffma2_rn(x1, B1, sacc[2 * q + 1]);
                const float2 g0 = __ffma2_rn(nu0, B0, h0), g1 = __ffma2_rn(nu1, B1, h1);  // a_t h_{t-1}
                const float2 pq0 = __fmul2_rn(x0, g0), pq1 = __fmul2_rn(x1, g1);
                const float2 And = make_float2(An[s], An[s]);
                wacc[2 * q] = __ffma2_rn(pq0, And, wacc[2 * q]);
                wacc[2 * q + 1] = __ffma2_rn(pq1, And, wacc[2 * q + 1]);
                dA2[s] = __ffma2_rn(pq0, dl0, dA2[s]);
                dA2[s] = __ffma2_rn(pq1, dl1, dA2[s]);
                V[s][0] = dB0.x; V[s][1] = dB0.y; V[s][2] = dB1.x; V[s][3] = dB1.y;
                V[SN + s][0] = dC0.x; V[SN + s][1] = dC0.y; V[SN + s][2] = dC1.x; V[SN + s][3] = dC1.y;
            }
            // sum over the warp's CPW channels: transposed reduction across the channel bits of the lane id — CPW float4
            // in, ONE out (slot == this lane's channel index), CPW-1 float4 exchanged instead of CPW-1 staged and re-read
#pragma unroll
            for (int m = CPW / 2; m >= 1; m >>= 1) {
                const bool up = (cw & m) != 0;
#pragma unroll
                for (int j = 0; j < m; ++j) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float snd = up ? V[j][e] : V[j + m][e];
                        const float kp = up ? V[j + m][e] : V[j][e];
                        V[j][e] = kp + __shfl_xor_sync(0xffffffffu, snd, m * LPC);
                    }
                }
            }
            red_w[rcur + q * kWarp] = make_float4(V[0][0], V[0][1], V[0][2], V[0][3]);
        }
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            uv_p[i] = uv_c[i]; dl_p[i] = dl_c[i]; gv_p[i] = gv_c[i];
            uv_c[i] = uv_n[i]; dl_c[i] = dl_n[i]; gv_c[i] = gv_n[i];
        }
        hin = hin_next;
        { const int t = slot0; slot0 = slot1; slot1 = slot2; slot2 = t; }
        xcur ^= XBUF;
        rcur ^= RBUF;
        cp_async_wait<0>();  // the tile of block k-2 (issued at the top) has landed for this thread ...
        __syncthreads();     // ... and for everyone; block k's staged dB / dC products are complete
    }
    reduce_bc(0, rcur ^ RBUF);
    finalize(0, sacc, wacc, uv_p, dl_p, gv_p, active);

    // ---- per-channel sums over time (atomically over batch) ----
#pragma unroll
    for (int m = 1; m < LPC; m <<= 1) {
        dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, m);
        dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, m);
    }
    if (active) {
#pragma unroll
        for (int s = 0; s < SN; ++s) atomicAdd(pb.dA + c * kN + ng * SN + s, dA2[s].x + dA2[s].y);
        if (ng == 0) {
            if (pb.dD) atomicAdd(pb.dD + c, dD_acc);
            if (pb.ddelta_bias) atomicAdd(pb.ddelta_bias + c, dbias_acc);
        }
    }
}

template <typename in_t, typename out_t, int SN, int NW = 4, bool CROSS = false>
static int launch_bwd_t(const ss2d_scan_bwd_params &pb, cudaStream_t stream, CrossAux xi = CrossAux{nullptr, nullptr, nullptr}) {
    using SM = BwdSmem<in_t, out_t, SN, NW>;
    const ss2d_scan_fwd_params &p = pb.f;
    const int per_g = (int)(p.dim / p.ngroups);
    const int tiles = (per_g + SM::CPC - 1) / SM::CPC;
    const int64_t ei = sizeof(in_t), eo = sizeof(out_t);
    Flags fl{};
    fl.vec_u = aligned16(p.u) && (p.u_bstride * ei) % 16 == 0 && (p.u_dstride * ei) % 16 == 0;
    fl.vec_delta = aligned16(p.delta) && (p.delta_bstride * ei) % 16 == 0 && (p.delta_dstride * ei) % 16 == 0;
    fl.vec_bc = aligned16(p.B) && aligned16(p.C) && (p.B_bstride * ei) % 16 == 0 && (p.B_gstride * ei) % 16 == 0 &&
                (p.B_nstride * ei) % 16 == 0 && (p.C_bstride * ei) % 16 == 0 && (p.C_gstride * ei) % 16 == 0 &&
                (p.C_nstride * ei) % 16 == 0;
    fl.vec_dout = aligned16(pb.dout) && (pb.dout_bstride * eo) % 16 == 0 && (pb.dout_dstride * eo) % 16 == 0;
    fl.vec_z = p.z && aligned16(p.z) && (p.z_bstride * ei) % 16 == 0 && (p.z_dstride * ei) % 16 == 0;
    fl.vec_out = p.out && aligned16(p.out) && (p.out_bstride * eo) % 16 == 0 && (p.out_dstride * eo) % 16 == 0;
    fl.vec_dbc = aligned16(pb.dB) && aligned16(pb.dC) && p.seqlen % 4 == 0;
    fl.vec_grad = aligned16(pb.du) && aligned16(pb.ddelta) && (!pb.dz || aligned16(pb.dz)) && (p.seqlen * ei) % 16 == 0;
    if (p.ckpt && !aligned16(p.ckpt)) return SS2D_ESTRIDE;  // read with 8/16-byte loads
    const int64_t grid = p.batch * p.ngroups * tiles;
    const bool fast = fl.vec_u && fl.vec_delta && fl.vec_bc && fl.vec_dout && fl.vec_dbc && fl.vec_grad && p.seqlen % BK == 0 && !p.z;
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)SM::total);
        if (rc != 0) return rc;
        kern<<<(unsigned)grid, NW * kWarp, SM::total, stream>>>(pb, tiles, fl, xi);
        return (int)cudaGetLastError();
    };
    if constexpr (CROSS) {
        if (!fast) return SS2D_ESTRIDE;  // covered problem (cross_covered) with a gradient buffer that is not 16-byte aligned
        return go(sl_bwd_kernel<in_t, out_t, SN, NW, true, true>);
    } else {
        return fast ? go(sl_bwd_kernel<in_t, out_t, SN, NW, true>) : go(sl_bwd_kernel<in_t, out_t, SN, NW, false>);
    }
}

template <typename in_t, typename out_t> static int launch_bwd_sn(const ss2d_scan_bwd_params &pb, cudaStream_t s) {
    // 2 states per lane at every size: with 4 the a_t / h_t history of a block needs 128 registers (255 per thread) and
    // the dB/dC staging area 16 KB per warp; measured at B=32, L=16384: 4 warps/CTA 15.7 ms (one CTA per SM), 2 warps/CTA
    // 9.73 ms, against 9.34 ms for 2 states per lane
    return launch_bwd_t<in_t, out_t, 2>(pb, s);
}

int launch_bwd(const ss2d_scan_bwd_params &pb, cudaStream_t s) {
    const ss2d_scan_fwd_params &p = pb.f;
    switch (p.in_dtype) {
        case SS2D_F32: return launch_bwd_sn<float, float>(pb, s);
        case SS2D_F16:
            return p.out_dtype == SS2D_F32 ? launch_bwd_sn<__half, float>(pb, s) : launch_bwd_sn<__half, __half>(pb, s);
        case SS2D_BF16:
            return p.out_dtype == SS2D_F32 ? launch_bwd_sn<__nv_bfloat16, float>(pb, s)
                                           : launch_bwd_sn<__nv_bfloat16, __nv_bfloat16>(pb, s);
        default: return SS2D_EDTYPE;
    }
}

int launch_cross_bwd(const ss2d_scan_bwd_params &pb, const CrossAux &aux, cudaStream_t s) {
    if (!aligned16(aux.uT) || !aligned16(aux.doutT) || !aligned16(aux.accT)) return SS2D_ESTRIDE;
    return launch_bwd_t<float, float, 2, 4, true>(pb, s, aux);
}

}  // namespace sl
}  // namespace ss2d
