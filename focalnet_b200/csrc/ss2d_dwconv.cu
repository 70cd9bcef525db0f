// ss2d_dwconv.cu — depthwise 3x3 conv (pad 1) + bias + SiLU, channels-last in -> channels-first out (sm_100a).
//
// Replaces three separate passes of SS2D.forwardv2 (reference: ITS/models/vmamba_layers.py:591-594 with the conv
// of :460-469): `x.permute(0,3,1,2).contiguous()` (ATen copy), `self.conv2d(x)` (cuDNN depthwise) and
// `self.act(x)` (SiLU) — one kernel reads the in_proj output's x-half in its native (B,H,W,cstride) layout and
// writes (B,d_inner,H,W).  HBM-bound: algorithmic bytes = 4*B*H*W*C read + 4*B*C*H*W written.
//
// Organisation (round 2; the round-1 kernels staged 32x32 tiles through shared memory and spent their time in 9 LDS per
// output): **lanes = 32 consecutive channels, a thread owns a strip of 4 rows x 8 columns of ONE channel**.  Every
// channels-last access (input pixels, dpre, dxin) is then one 128-byte line per warp instruction, the 6 x 10 input
// window of a strip is loaded straight into registers (60 independent loads in flight per thread, halos hit L1/L2), the
// 3x3 window slides in registers (no shared memory, no barrier), and the channels-first side moves 32 contiguous bytes
// (one full sector) per lane and row.
//
// Backward: pass 1 recomputes s = conv + bias and writes dpre = dout * silu'(s) channels-last (scratch);
// pass 2 produces dxin (correlation with the flipped taps) and accumulates dweight / dbias in registers while a warp
// walks down its column of strips — one atomicAdd per (channel, tap) per warp.
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kDwR = 4;      // rows per strip
constexpr int kDwC = 8;      // columns per strip (one 32-byte sector of a channels-first row)
constexpr int kDwWarps = 8;  // warps per CTA: 8 strips that are neighbours along W (their halos share L1 lines)

struct DwGeom {
    int B, C, H, W, strips_w, strips_h, cgroups;
    int64_t cstride;
    bool vec;      // channels-first rows can be moved as float4 (W % 4 == 0, 16-byte aligned base)
    int hb_per_warp;  // grad kernel: row-blocks one warp walks (register accumulation of dweight / dbias)
};

// strip -> (b, hb, cg, ws): strips that are neighbours along W are neighbours in the index
__device__ __forceinline__ bool dw_strip(const DwGeom &g, int64_t sid, int &b, int &hb, int &cg, int &ws) {
    ws = (int)(sid % g.strips_w); sid /= g.strips_w;
    cg = (int)(sid % g.cgroups); sid /= g.cgroups;
    hb = (int)(sid % g.strips_h); sid /= g.strips_h;
    b = (int)sid;
    return b < g.B;
}

// the (kDwR+2) x (kDwC+2) window of a channels-last tensor around strip (hb, ws) for channel c (zero outside the image)
__device__ __forceinline__ void dw_window(float (&v)[kDwR + 2][kDwC + 2], const float *__restrict__ src, int64_t cstride,
                                          const DwGeom &g, int b, int h0, int w0, int c, bool active) {
#pragma unroll
    for (int r = 0; r < kDwR + 2; ++r) {
        const int hh = h0 - 1 + r;
        const bool rok = active && hh >= 0 && hh < g.H;
        const float *row = src + ((int64_t)b * g.H + (rok ? hh : 0)) * g.W * cstride + c;
#pragma unroll
        for (int j = 0; j < kDwC + 2; ++j) {
            const int ww = w0 - 1 + j;
            v[r][j] = (rok && ww >= 0 && ww < g.W) ? __ldg(row + (int64_t)ww * cstride) : 0.f;
        }
    }
}

// kDwC consecutive values of a channels-first row <-> registers
__device__ __forceinline__ void dw_row_load(float (&o)[kDwC], const float *__restrict__ p, int nvalid, bool vec) {
    if (vec && nvalid >= kDwC) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < kDwC; ++j) o[j] = j < nvalid ? __ldg(p + j) : 0.f;
    }
}
__device__ __forceinline__ void dw_row_store(float *__restrict__ p, const float (&o)[kDwC], int nvalid, bool vec) {
    if (vec && nvalid >= kDwC) {
        reinterpret_cast<float4 *>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int j = 0; j < kDwC; ++j)
            if (j < nvalid) p[j] = o[j];
    }
}

__global__ void __launch_bounds__(kDwWarps *kWarp) dwconv_silu_fwd_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                         const float *__restrict__ bias, float *__restrict__ out,
                                                                         const DwGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b, hb, cg, ws;
    if (!dw_strip(g, (int64_t)blockIdx.x * kDwWarps + warp, b, hb, cg, ws)) return;
    const int c = cg * kWarp + lane, h0 = hb * kDwR, w0 = ws * kDwC;
    const bool active = c < g.C;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = active ? __ldg(weight + c * 9 + q) : 0.f;
    const float bv = (bias && active) ? __ldg(bias + c) : 0.f;
    float v[kDwR + 2][kDwC + 2];
    dw_window(v, xin, g.cstride, g, b, h0, w0, c, active);
    if (!active) return;
#pragma unroll
    for (int r = 0; r < kDwR; ++r) {
        if (h0 + r >= g.H) break;
        float o[kDwC];
#pragma unroll
        for (int j = 0; j < kDwC; ++j) {
            float s = bv;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int e = 0; e < 3; ++e) s = fmaf(wgt[a * 3 + e], v[r + a][j + e], s);
            o[j] = s * sigmoidf_fast(s);
        }
        dw_row_store(out + (((int64_t)b * g.C + c) * g.H + h0 + r) * g.W + w0, o, g.W - w0, g.vec);
    }
}

// pass 1 of the backward: dpre (channels-last, dense C) = dout * silu'(conv + bias)
__global__ void __launch_bounds__(kDwWarps *kWarp) dwconv_silu_dpre_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                          const float *__restrict__ bias, const float *__restrict__ dout,
                                                                          float *__restrict__ dpre, const DwGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b, hb, cg, ws;
    if (!dw_strip(g, (int64_t)blockIdx.x * kDwWarps + warp, b, hb, cg, ws)) return;
    const int c = cg * kWarp + lane, h0 = hb * kDwR, w0 = ws * kDwC;
    const bool active = c < g.C;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = active ? __ldg(weight + c * 9 + q) : 0.f;
    const float bv = (bias && active) ? __ldg(bias + c) : 0.f;
    float v[kDwR + 2][kDwC + 2];
    dw_window(v, xin, g.cstride, g, b, h0, w0, c, active);
    if (!active) return;
#pragma unroll
    for (int r = 0; r < kDwR; ++r) {
        if (h0 + r >= g.H) break;
        float go[kDwC];
        dw_row_load(go, dout + (((int64_t)b * g.C + c) * g.H + h0 + r) * g.W + w0, g.W - w0, g.vec);
        float *drow = dpre + (((int64_t)b * g.H + h0 + r) * g.W + w0) * g.C + c;
#pragma unroll
        for (int j = 0; j < kDwC; ++j) {
            float s = bv;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int e = 0; e < 3; ++e) s = fmaf(wgt[a * 3 + e], v[r + a][j + e], s);
            const float sg = sigmoidf_fast(s);
            if (w0 + j < g.W) drow[(int64_t)j * g.C] = go[j] * sg * (1.f + s * (1.f - sg));
        }
    }
}

// pass 2: dxin = corr(dpre, flipped taps); dweight[c][tap] += sum dpre * xin(shifted); dbias[c] += sum dpre.
// A warp walks hb_per_warp row-blocks of its column of strips, so the dweight / dbias partial sums stay in registers and
// one atomicAdd per (channel, tap) per warp reaches L2.
__global__ void __launch_bounds__(kDwWarps *kWarp) dwconv_silu_grad_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                          const float *__restrict__ dpre, float *__restrict__ dxin,
                                                                          int64_t dx_cstride, float *__restrict__ dweight,
                                                                          float *__restrict__ dbias, const DwGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // walk index -> (b, hchunk, cg, ws)
    int64_t sid = (int64_t)blockIdx.x * kDwWarps + warp;
    const int hchunks = (g.strips_h + g.hb_per_warp - 1) / g.hb_per_warp;
    const int ws = (int)(sid % g.strips_w); sid /= g.strips_w;
    const int cg = (int)(sid % g.cgroups); sid /= g.cgroups;
    const int hc = (int)(sid % hchunks); sid /= hchunks;
    const int b = (int)sid;
    if (b >= g.B) return;
    const int c = cg * kWarp + lane, w0 = ws * kDwC;
    const bool active = c < g.C;
    float wgt[9], acc[10];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = active ? __ldg(weight + c * 9 + q) : 0.f;
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q] = 0.f;
    const int hb_end = min(g.strips_h, (hc + 1) * g.hb_per_warp);
    for (int hb = hc * g.hb_per_warp; hb < hb_end; ++hb) {
        const int h0 = hb * kDwR;
        float dp[kDwR + 2][kDwC + 2], xv[kDwR + 2][kDwC + 2];
        dw_window(dp, dpre, g.C, g, b, h0, w0, c, active);
        dw_window(xv, xin, g.cstride, g, b, h0, w0, c, active);
        if (!active) continue;
#pragma unroll
        for (int r = 0; r < kDwR; ++r) {
            if (h0 + r >= g.H) break;
            float *xrow = dxin + (((int64_t)b * g.H + h0 + r) * g.W + w0) * dx_cstride + c;
#pragma unroll
            for (int j = 0; j < kDwC; ++j) {
                // input pixel (h, w) was read by output pixel (h+1-a, w+1-e) through tap (a, e)
                float s = 0.f;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int e = 0; e < 3; ++e) s = fmaf(wgt[a * 3 + e], dp[r + 2 - a][j + 2 - e], s);
                if (w0 + j < g.W) xrow[(int64_t)j * dx_cstride] = s;
                const float gp = dp[r + 1][j + 1];  // dpre at output pixel (h, w): zero outside the image
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int e = 0; e < 3; ++e) acc[a * 3 + e] = fmaf(gp, xv[r + a][j + e], acc[a * 3 + e]);
                acc[9] += gp;
            }
        }
    }
    if (active) {
#pragma unroll
        for (int q = 0; q < 9; ++q) atomicAdd(dweight + c * 9 + q, acc[q]);
        if (dbias) atomicAdd(dbias + c, acc[9]);
    }
}

static int dw_geom(DwGeom &g, int64_t cstride, int64_t B, int64_t C, int64_t H, int64_t W, const void *cf_ptr) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || cstride < C) return SS2D_EINVAL;
    g.B = (int)B; g.C = (int)C; g.H = (int)H; g.W = (int)W; g.cstride = cstride;
    g.strips_w = (int)((W + kDwC - 1) / kDwC); g.strips_h = (int)((H + kDwR - 1) / kDwR); g.cgroups = (int)((C + kWarp - 1) / kWarp);
    g.vec = W % 4 == 0 && (reinterpret_cast<uintptr_t>(cf_ptr) & 15) == 0;
    g.hb_per_warp = 1;
    if ((int64_t)g.strips_w * g.strips_h * g.cgroups * B / kDwWarps + 1 > 0x7fffffffLL) return SS2D_EINVAL;
    return 0;
}

}  // namespace ss2d

extern "C" int ss2d_dwconv_silu_fwd(const float *xin, int64_t cstride, const float *weight, const float *bias, float *out,
                                    int64_t batch, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace ss2d;
    if (!xin || !weight || !out) return SS2D_EINVAL;
    DwGeom g;
    if (int rc = dw_geom(g, cstride, batch, C, H, W, out)) return rc;
    const int64_t strips = (int64_t)g.strips_w * g.strips_h * g.cgroups * batch;
    dwconv_silu_fwd_kernel<<<(unsigned)((strips + kDwWarps - 1) / kDwWarps), kDwWarps * kWarp, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xin, weight, bias, out, g);
    return (int)cudaGetLastError();
}

extern "C" int ss2d_dwconv_silu_bwd(const float *xin, int64_t cstride, const float *weight, const float *bias, const float *dout,
                                    float *dpre_scratch, float *dxin, int64_t dx_cstride, float *dweight, float *dbias,
                                    int64_t batch, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace ss2d;
    if (!xin || !weight || !dout || !dpre_scratch || !dxin || !dweight || dx_cstride < C) return SS2D_EINVAL;
    DwGeom g;
    if (int rc = dw_geom(g, cstride, batch, C, H, W, dout)) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int64_t strips = (int64_t)g.strips_w * g.strips_h * g.cgroups * batch;
    dwconv_silu_dpre_kernel<<<(unsigned)((strips + kDwWarps - 1) / kDwWarps), kDwWarps * kWarp, 0, s>>>(xin, weight, bias, dout, dpre_scratch, g);
    // row-blocks per warp in pass 2: as many as keeps >= ~8 warps per SM sub-partition in flight (fewer atomics per tap)
    const int64_t want_warps = 148 * 4 * 8;
    int64_t hbpw = strips / want_warps;
    hbpw = hbpw < 1 ? 1 : (hbpw > g.strips_h ? g.strips_h : hbpw);
    g.hb_per_warp = (int)hbpw;
    const int64_t hchunks = (g.strips_h + hbpw - 1) / hbpw;
    const int64_t walks = (int64_t)g.strips_w * g.cgroups * hchunks * batch;
    dwconv_silu_grad_kernel<<<(unsigned)((walks + kDwWarps - 1) / kDwWarps), kDwWarps * kWarp, 0, s>>>(xin, weight, dpre_scratch, dxin, dx_cstride,
                                                                                                       dweight, dbias, g);
    return (int)cudaGetLastError();
}
