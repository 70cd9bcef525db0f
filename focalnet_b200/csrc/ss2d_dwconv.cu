// ss2d_dwconv.cu — depthwise 3x3 conv (pad 1) + bias + SiLU, channels-last in -> channels-first out (sm_100a).
//
// Replaces three separate passes of SS2D.forwardv2 (reference: ITS/models/vmamba_layers.py:591-594 with the conv
// of :460-469): `x.permute(0,3,1,2).contiguous()` (ATen copy), `self.conv2d(x)` (cuDNN depthwise) and
// `self.act(x)` (SiLU) — one kernel reads the in_proj output's x-half in its native (B,H,W,cstride) layout and
// writes (B,d_inner,H,W).  HBM-bound: algorithmic bytes = 4*B*H*W*C read + 4*B*C*H*W written.
//
// Organisation (round 2; the round-1 kernels staged 32x32 tiles through shared memory and spent their time in 9 LDS per
// output): **lanes = 32 consecutive channels; a warp walks DOWN a strip of 8 columns** with the 3-row x 10-column input
// window of its channel sliding through registers (4 row buffers: the row two below the one being computed is in flight
// while the current row is evaluated).  Every channels-last access (input pixels, dpre, dxin) is one 128-byte line per
// warp instruction, each input value is loaded 1.25 times (column halo only), there is no shared memory and no barrier,
// and the channels-first side moves 32 contiguous bytes (one full sector) per lane and row.  Strips whose 10 columns lie
// inside the image take a path without column predicates.
//
// Backward: pass 1 recomputes s = conv + bias and writes dpre = dout * silu'(s) channels-last (scratch);
// pass 2 produces dxin (correlation with the flipped taps) and accumulates dweight / dbias in registers while the warp
// walks its chunk of rows — one atomicAdd per (channel, tap) per warp.
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kDwC = 8;      // columns per strip (one 32-byte sector of a channels-first row)
constexpr int kDwWin = kDwC + 2;
constexpr int kDwWarps = 4;  // warps per CTA: strips that are neighbours along W (their column halos share L1 lines)

struct DwGeom {
    int B, C, H, W, strips_w, cgroups, rows_per_walk, hchunks;
    int64_t cstride;
    bool vec;  // channels-first rows can be moved as float4 (W % 4 == 0, 16-byte aligned base)
};

// walk -> (b, hc, cg, ws): walks that are neighbours along W are neighbours in the index
__device__ __forceinline__ bool dw_walk(const DwGeom &g, int64_t id, int &b, int &hc, int &cg, int &ws) {
    ws = (int)(id % g.strips_w); id /= g.strips_w;
    cg = (int)(id % g.cgroups); id /= g.cgroups;
    hc = (int)(id % g.hchunks); id /= g.hchunks;
    b = (int)id;
    return b < g.B;
}

// one row (10 columns around the strip) of a channels-last tensor for this lane's channel; zero outside the image.
// `img` points at pixel (b, 0, 0), channel c.
template <bool INTERIOR>
__device__ __forceinline__ void dw_load_row(float (&r)[kDwWin], const float *__restrict__ img, int64_t cstride, int hh, int H, int W, int w0) {
    if (hh < 0 || hh >= H) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < kDwWin; ++j) r[j] = 0.f;
        return;
    }
    const float *p = img + ((int64_t)hh * W + (w0 - 1)) * cstride;
#pragma unroll
    for (int j = 0; j < kDwWin; ++j) {
        r[j] = (INTERIOR || (w0 - 1 + j >= 0 && w0 - 1 + j < W)) ? __ldg(p) : 0.f;
        p += cstride;
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

__device__ __forceinline__ float dw_conv_at(const float (&wgt)[9], float bv, const float (&ra)[kDwWin], const float (&rb)[kDwWin],
                                            const float (&rc)[kDwWin], int j) {
    float s = bv;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        s = fmaf(wgt[e], ra[j + e], s);
        s = fmaf(wgt[3 + e], rb[j + e], s);
        s = fmaf(wgt[6 + e], rc[j + e], s);
    }
    return s;
}

enum { kDwFwd = 0, kDwDpre = 1 };

// forward (MODE kDwFwd: out = silu(conv + bias), channels-first) and pass 1 of the backward (MODE kDwDpre: dpre = dout *
// silu'(conv + bias), channels-last dense C).  `cf` is the channels-first tensor of the mode (out resp. dout).
template <int MODE, bool INTERIOR>
__device__ __forceinline__ void dw_walk_rows(const float *__restrict__ xin, const float (&wgt)[9], float bv, float *__restrict__ out,
                                             const float *__restrict__ dout, float *__restrict__ dpre, const DwGeom &g, int b, int c,
                                             int h_begin, int h_end, int w0) {
    const float *img = xin + (int64_t)b * g.H * g.W * g.cstride + c;
    float win[4][kDwWin];
    // h_begin is a multiple of 4 (rows_per_walk is): row h lives in buffer h & 3, all buffer indices are compile-time constants
    dw_load_row<INTERIOR>(win[3], img, g.cstride, h_begin - 1, g.H, g.W, w0);
    dw_load_row<INTERIOR>(win[0], img, g.cstride, h_begin, g.H, g.W, w0);
    dw_load_row<INTERIOR>(win[1], img, g.cstride, h_begin + 1, g.H, g.W, w0);
#pragma unroll 1
    for (int h4 = h_begin; h4 < h_end; h4 += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = h4 + k;
            if (h >= h_end) break;
            dw_load_row<INTERIOR>(win[(k + 2) & 3], img, g.cstride, h + 2, g.H, g.W, w0);  // in flight while row h is evaluated
            float (&ra)[kDwWin] = win[(k + 3) & 3], (&rb)[kDwWin] = win[k], (&rc)[kDwWin] = win[(k + 1) & 3];
            const int64_t cf_off = (((int64_t)b * g.C + c) * g.H + h) * g.W + w0;
            if constexpr (MODE == kDwFwd) {
                float o[kDwC];
#pragma unroll
                for (int j = 0; j < kDwC; ++j) {
                    const float s = dw_conv_at(wgt, bv, ra, rb, rc, j);
                    o[j] = s * sigmoidf_fast(s);
                }
                dw_row_store(out + cf_off, o, g.W - w0, g.vec);
            } else {
                float go[kDwC];
                dw_row_load(go, dout + cf_off, g.W - w0, g.vec);
                float *drow = dpre + (((int64_t)b * g.H + h) * g.W + w0) * g.C + c;
#pragma unroll
                for (int j = 0; j < kDwC; ++j) {
                    const float s = dw_conv_at(wgt, bv, ra, rb, rc, j);
                    const float sg = sigmoidf_fast(s);
                    if (INTERIOR || w0 + j < g.W) drow[(int64_t)j * g.C] = go[j] * sg * (1.f + s * (1.f - sg));
                }
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kDwWarps *kWarp, 5) dwconv_silu_walk_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                             const float *__restrict__ bias, float *__restrict__ out,
                                                                             const float *__restrict__ dout, float *__restrict__ dpre,
                                                                             const DwGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b, hc, cg, ws;
    if (!dw_walk(g, (int64_t)blockIdx.x * kDwWarps + warp, b, hc, cg, ws)) return;
    const int c = cg * kWarp + lane, w0 = ws * kDwC;
    if (c >= g.C) return;
    float wgt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = __ldg(weight + c * 9 + q);
    const float bv = bias ? __ldg(bias + c) : 0.f;
    const int h_begin = hc * g.rows_per_walk, h_end = min(g.H, h_begin + g.rows_per_walk);
    if (w0 >= 1 && w0 + kDwC + 1 <= g.W) dw_walk_rows<MODE, true>(xin, wgt, bv, out, dout, dpre, g, b, c, h_begin, h_end, w0);
    else dw_walk_rows<MODE, false>(xin, wgt, bv, out, dout, dpre, g, b, c, h_begin, h_end, w0);
}

// pass 2: dxin = corr(dpre, flipped taps); dweight[c][tap] += sum dpre * xin(shifted); dbias[c] += sum dpre.
template <bool INTERIOR>
__device__ __forceinline__ void dw_grad_rows(const float *__restrict__ xin, const float *__restrict__ dpre, const float (&wgt)[9],
                                             float *__restrict__ dxin, int64_t dx_cstride, float (&acc)[10], const DwGeom &g, int b, int c,
                                             int h_begin, int h_end, int w0) {
    const float *ximg = xin + (int64_t)b * g.H * g.W * g.cstride + c;
    const float *dimg = dpre + (int64_t)b * g.H * g.W * g.C + c;
    float xw[4][kDwWin], dw[4][kDwWin];
#pragma unroll
    for (int k = -1; k <= 1; ++k) {  // h_begin % 4 == 0: row h lives in buffer h & 3
        dw_load_row<INTERIOR>(xw[k & 3], ximg, g.cstride, h_begin + k, g.H, g.W, w0);
        dw_load_row<INTERIOR>(dw[k & 3], dimg, g.C, h_begin + k, g.H, g.W, w0);
    }
#pragma unroll 1
    for (int h4 = h_begin; h4 < h_end; h4 += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = h4 + k;
            if (h >= h_end) break;
            dw_load_row<INTERIOR>(xw[(k + 2) & 3], ximg, g.cstride, h + 2, g.H, g.W, w0);
            dw_load_row<INTERIOR>(dw[(k + 2) & 3], dimg, g.C, h + 2, g.H, g.W, w0);
            float (&xa)[kDwWin] = xw[(k + 3) & 3], (&xb)[kDwWin] = xw[k], (&xc)[kDwWin] = xw[(k + 1) & 3];
            float (&da)[kDwWin] = dw[(k + 3) & 3], (&db)[kDwWin] = dw[k], (&dc)[kDwWin] = dw[(k + 1) & 3];
            float *xrow = dxin + (((int64_t)b * g.H + h) * g.W + w0) * dx_cstride + c;
#pragma unroll
            for (int j = 0; j < kDwC; ++j) {
                // input pixel (h, w) was read by output pixel (h+1-a, w+1-e) through tap (a, e): rows h+1, h, h-1 of dpre
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    s = fmaf(wgt[e], dc[j + 2 - e], s);
                    s = fmaf(wgt[3 + e], db[j + 2 - e], s);
                    s = fmaf(wgt[6 + e], da[j + 2 - e], s);
                }
                if (INTERIOR || w0 + j < g.W) xrow[(int64_t)j * dx_cstride] = s;
                const float gp = db[j + 1];  // dpre at output pixel (h, w): zero outside the image
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    acc[e] = fmaf(gp, xa[j + e], acc[e]);
                    acc[3 + e] = fmaf(gp, xb[j + e], acc[3 + e]);
                    acc[6 + e] = fmaf(gp, xc[j + e], acc[6 + e]);
                }
                acc[9] += gp;
            }
        }
    }
}

__global__ void __launch_bounds__(kDwWarps *kWarp, 3) dwconv_silu_grad_kernel(const float *__restrict__ xin, const float *__restrict__ weight,
                                                                             const float *__restrict__ dpre, float *__restrict__ dxin,
                                                                             int64_t dx_cstride, float *__restrict__ dweight,
                                                                             float *__restrict__ dbias, const DwGeom g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b, hc, cg, ws;
    if (!dw_walk(g, (int64_t)blockIdx.x * kDwWarps + warp, b, hc, cg, ws)) return;
    const int c = cg * kWarp + lane, w0 = ws * kDwC;
    if (c >= g.C) return;
    float wgt[9], acc[10];
#pragma unroll
    for (int q = 0; q < 9; ++q) wgt[q] = __ldg(weight + c * 9 + q);
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q] = 0.f;
    const int h_begin = hc * g.rows_per_walk, h_end = min(g.H, h_begin + g.rows_per_walk);
    if (w0 >= 1 && w0 + kDwC + 1 <= g.W) dw_grad_rows<true>(xin, dpre, wgt, dxin, dx_cstride, acc, g, b, c, h_begin, h_end, w0);
    else dw_grad_rows<false>(xin, dpre, wgt, dxin, dx_cstride, acc, g, b, c, h_begin, h_end, w0);
#pragma unroll
    for (int q = 0; q < 9; ++q) atomicAdd(dweight + c * 9 + q, acc[q]);
    if (dbias) atomicAdd(dbias + c, acc[9]);
}

// rows_per_walk: a multiple of 4, as long as possible (the 2-row halo of a walk is re-read) while >= ~48 warps per SM exist
static int dw_geom(DwGeom &g, int64_t cstride, int64_t B, int64_t C, int64_t H, int64_t W, const void *cf_ptr) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || cstride < C) return SS2D_EINVAL;
    g.B = (int)B; g.C = (int)C; g.H = (int)H; g.W = (int)W; g.cstride = cstride;
    g.strips_w = (int)((W + kDwC - 1) / kDwC); g.cgroups = (int)((C + kWarp - 1) / kWarp);
    g.vec = W % 4 == 0 && (reinterpret_cast<uintptr_t>(cf_ptr) & 15) == 0;
    const int64_t cols = (int64_t)g.strips_w * g.cgroups * B;
    int64_t chunks = (148LL * 48 + cols - 1) / cols;
    const int64_t max_chunks = (H + 3) / 4;
    chunks = chunks < 1 ? 1 : (chunks > max_chunks ? max_chunks : chunks);
    g.rows_per_walk = (int)(((H + chunks - 1) / chunks + 3) / 4 * 4);
    g.hchunks = (int)((H + g.rows_per_walk - 1) / g.rows_per_walk);
    if (cols * g.hchunks / kDwWarps + 1 > 0x7fffffffLL) return SS2D_EINVAL;
    return 0;
}
static unsigned dw_grid(const DwGeom &g) {
    return (unsigned)(((int64_t)g.strips_w * g.cgroups * g.hchunks * g.B + kDwWarps - 1) / kDwWarps);
}

}  // namespace ss2d

extern "C" int ss2d_dwconv_silu_fwd(const float *xin, int64_t cstride, const float *weight, const float *bias, float *out,
                                    int64_t batch, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace ss2d;
    if (!xin || !weight || !out) return SS2D_EINVAL;
    DwGeom g;
    if (int rc = dw_geom(g, cstride, batch, C, H, W, out)) return rc;
    dwconv_silu_walk_kernel<kDwFwd><<<dw_grid(g), kDwWarps * kWarp, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xin, weight, bias, out, nullptr,
                                                                                                                 nullptr, g);
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
    dwconv_silu_walk_kernel<kDwDpre><<<dw_grid(g), kDwWarps * kWarp, 0, s>>>(xin, weight, bias, nullptr, dout, dpre_scratch, g);
    dwconv_silu_grad_kernel<<<dw_grid(g), kDwWarps * kWarp, 0, s>>>(xin, weight, dpre_scratch, dxin, dx_cstride, dweight, dbias, g);
    return (int)cudaGetLastError();
}
