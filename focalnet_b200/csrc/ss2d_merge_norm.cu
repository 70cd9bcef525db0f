// ss2d_merge_norm.cu — epilogue of the SS2D core: transpose + LayerNorm (+ SiLU gate) in one pass (sm_100a).
//
// Replaces, for out_norm = LayerNorm / out_norm_shape "v0" (the ITS model's configuration):
//     y = y.transpose(1, 2).contiguous(); y = out_norm(y)      ITS/models/vmamba_layers.py:296-297
//     z = act(z) ... y = y * z                                  ITS/models/vmamba_layers.py:588-589,599
// i.e. three full passes over (B, L, d_inner) (ATen transpose copy, LayerNorm, elementwise multiply) by one
// HBM-bound kernel that reads the merged scan output y:(B,D,L) (spatial order, fp32) once, normalises each pixel over
// its D channels, optionally multiplies by silu(z) read straight from the z-half of the in_proj output (channels-last,
// arbitrary pixel stride) and writes (B,L,D) channels-last, ready for out_proj.
// Algorithmic bytes: fwd 4*B*D*L * (2 + gate); bwd 4*B*D*L * (3 + 2*gate).  Lanes walk pixels on the (B,D,L) side and
// channels on the (B,L,D) side; a 32-pixel x D tile is turned through shared memory so both sides are coalesced.
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kMnPix = 32;    // pixels per tile
constexpr int kMnWarps = 8;
constexpr int kMnMaxJ = 16;   // supports D <= 512 (kernels are instantiated for ceil(D/32) <= 2, 4, 6, 8, 16)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// y:(B,D,L) -> out:(B,L,D);  out = (LN_D(y) * w + b) [* silu(z)]
template <int NJ>
__global__ void __launch_bounds__(kMnWarps *kWarp, NJ <= 8 ? 3 : 1) merge_norm_fwd_kernel(const float *__restrict__ y, const float *__restrict__ w,
                                                                         const float *__restrict__ bvec, const float *__restrict__ z,
                                                                         int64_t z_pstride, float *__restrict__ out, int D,
                                                                         int64_t L, float eps, int tiles_per_image) {
    extern __shared__ float tile[];  // [D][kMnPix + 1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x / tiles_per_image;
    const int64_t l0 = (int64_t)(blockIdx.x % tiles_per_image) * kMnPix;
    const int nj = (D + kWarp - 1) / kWarp;
    constexpr int RPW = NJ * kWarp / kMnWarps;  // y rows per warp
    constexpr int PPW = kMnPix / kMnWarps;      // pixels per warp
    // every global load of the tile is issued before the first use: the RPW y rows of this warp, then (behind the barrier's
    // back) the z values and the affine parameters of the PPW pixels this warp will normalise
    {
        float r[RPW];
        const int64_t l = l0 + lane;
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int d = warp + i * kMnWarps;
            r[i] = (d < D && l < L) ? __ldg(y + ((int64_t)b * D + d) * L + l) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int d = warp + i * kMnWarps;
            if (d < D) tile[d * (kMnPix + 1) + lane] = r[i];
        }
    }
    float zv[PPW][NJ], wv[NJ], bv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = lane + j * kWarp;
        const bool ok = j < nj && d < D;
        wv[j] = ok ? __ldg(w + d) : 0.f;
        bv[j] = ok ? __ldg(bvec + d) : 0.f;
#pragma unroll
        for (int k = 0; k < PPW; ++k) {
            const int64_t l = l0 + warp + k * kMnWarps;
            zv[k][j] = (z && ok && l < L) ? __ldg(z + ((int64_t)b * L + l) * z_pstride + d) : 0.f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PPW; ++k) {
        const int p = warp + k * kMnWarps;
        const int64_t l = l0 + p;
        if (l >= L) break;
        float v[NJ];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane + j * kWarp;
            v[j] = (j < nj && d < D) ? tile[d * (kMnPix + 1) + p] : 0.f;
            s += v[j];
        }
        const float mean = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane + j * kWarp;
            const float c = (j < nj && d < D) ? v[j] - mean : 0.f;
            q = fmaf(c, c, q);
        }
        const float rstd = rsqrtf(warp_sum(q) / D + eps);
        float *o = out + ((int64_t)b * L + l) * D;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int d = lane + j * kWarp;
            if (j < nj && d < D) {
                float r = (v[j] - mean) * rstd * wv[j] + bv[j];
                if (z) r *= zv[k][j] * sigmoidf_fast(zv[k][j]);
                o[d] = r;
            }
        }
    }
}

// dout:(B,L,D) -> dy:(B,D,L), dz:(B,L,*) (z_pstride), dw, db (accumulated with atomics; zeroed by the caller)
template <int NJ>
__global__ void __launch_bounds__(kMnWarps *kWarp, NJ <= 8 ? 2 : 1) merge_norm_bwd_kernel(const float *__restrict__ y, const float *__restrict__ w,
                                                                         const float *__restrict__ bvec, const float *__restrict__ z,
                                                                         int64_t z_pstride, const float *__restrict__ dout,
                                                                         float *__restrict__ dy, float *__restrict__ dz,
                                                                         int64_t dz_pstride, float *__restrict__ dw,
                                                                         float *__restrict__ db, int D, int64_t L, float eps,
                                                                         int tiles_per_image, int tiles_per_block, int batch) {
    extern __shared__ float tile[];  // [D][kMnPix + 1], reused for y then dy
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nj = (D + kWarp - 1) / kWarp;
    float dw_acc[NJ], db_acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { dw_acc[j] = 0.f; db_acc[j] = 0.f; }
    for (int tb = 0; tb < tiles_per_block; ++tb) {
        const int64_t tix = (int64_t)blockIdx.x * tiles_per_block + tb;
        const int b = (int)(tix / tiles_per_image);
        const int64_t l0 = (tix % tiles_per_image) * kMnPix;
        const bool live = b < batch;  // the last block may run past the last tile
        __syncthreads();
        constexpr int RPW = NJ * kWarp / kMnWarps, PPW = kMnPix / kMnWarps;
        float gv[PPW][NJ], zv[PPW][NJ];  // dout / z of the pixels this warp owns: in flight across the barrier
        if (live) {
            float r[RPW];
            const int64_t l = l0 + lane;
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int d = warp + i * kMnWarps;
                r[i] = (d < D && l < L) ? __ldg(y + ((int64_t)b * D + d) * L + l) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < PPW; ++k) {
                const int64_t lp = l0 + warp + k * kMnWarps;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane + j * kWarp;
                    const bool ok = j < nj && d < D && lp < L;
                    gv[k][j] = ok ? __ldg(dout + ((int64_t)b * L + lp) * D + d) : 0.f;
                    zv[k][j] = (z && ok) ? __ldg(z + ((int64_t)b * L + lp) * z_pstride + d) : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int d = warp + i * kMnWarps;
                if (d < D) tile[d * (kMnPix + 1) + lane] = r[i];
            }
        }
        __syncthreads();
        if (live)
#pragma unroll
            for (int k = 0; k < PPW; ++k) {
                const int p = warp + k * kMnWarps;
                const int64_t l = l0 + p;
                if (l >= L) break;
                float v[NJ];
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane + j * kWarp;
                    v[j] = (j < nj && d < D) ? tile[d * (kMnPix + 1) + p] : 0.f;
                    s += v[j];
                }
                const float mean = warp_sum(s) / D;
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane + j * kWarp;
                    const float c = (j < nj && d < D) ? v[j] - mean : 0.f;
                    q = fmaf(c, c, q);
                }
                const float rstd = rsqrtf(warp_sum(q) / D + eps);
                float *dzp = (z && dz) ? dz + ((int64_t)b * L + l) * dz_pstride : nullptr;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane + j * kWarp;
                    if (j < nj && d < D) {
                        const float xh = (v[j] - mean) * rstd;
                        float g = gv[k][j];
                        if (z) {
                            const float zz = zv[k][j], sg = sigmoidf_fast(zz);
                            if (dzp) dzp[d] = g * (xh * w[d] + bvec[d]) * sg * (1.f + zz * (1.f - sg));
                            g *= zz * sg;
                        }
                        dw_acc[j] = fmaf(g, xh, dw_acc[j]);
                        db_acc[j] += g;
                        const float dxh = g * w[d];
                        s1 += dxh;
                        s2 = fmaf(dxh, xh, s2);
                        v[j] = xh;             // keep xhat
                        tile[d * (kMnPix + 1) + p] = dxh;  // this warp owns column p: stash dxhat
                    }
                }
                s1 = warp_sum(s1) / D;
                s2 = warp_sum(s2) / D;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int d = lane + j * kWarp;
                    if (j < nj && d < D) tile[d * (kMnPix + 1) + p] = rstd * (tile[d * (kMnPix + 1) + p] - s1 - v[j] * s2);
                }
            }
        __syncthreads();
        if (live)
            for (int d = warp; d < D; d += kMnWarps) {
                const int64_t l = l0 + lane;
                if (l < L) dy[((int64_t)b * D + d) * L + l] = tile[d * (kMnPix + 1) + lane];
            }
    }
    // block reduction of dw / db over the 8 warps, then one atomic per channel per block
    __syncthreads();
    float *red = tile;  // [kMnWarps][2][D]
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = lane + j * kWarp;
        if (j < nj && d < D) { red[(warp * 2 + 0) * D + d] = dw_acc[j]; red[(warp * 2 + 1) * D + d] = db_acc[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += kMnWarps * kWarp) {
        const int which = i / D, d = i % D;
        float sacc = 0.f;
#pragma unroll
        for (int ww = 0; ww < kMnWarps; ++ww) sacc += red[(ww * 2 + which) * D + d];
        atomicAdd((which ? db : dw) + d, sacc);
    }
}

static size_t mn_smem(int D) {
    const size_t a = (size_t)D * (kMnPix + 1) * sizeof(float), b = (size_t)kMnWarps * 2 * D * sizeof(float);
    return a > b ? a : b;
}

}  // namespace ss2d

extern "C" int ss2d_merge_norm_gate_fwd(const float *y, const float *weight, const float *bias, float eps, const float *z,
                                        int64_t z_pstride, float *out, int64_t batch, int64_t D, int64_t L, void *stream) {
    using namespace ss2d;
    if (!y || !weight || !bias || !out || batch <= 0 || D <= 0 || L <= 0 || D > kMnMaxJ * kWarp) return SS2D_EINVAL;
    if (z && z_pstride < D) return SS2D_EINVAL;
    const int64_t tiles = (L + kMnPix - 1) / kMnPix;
    if (tiles * batch > 0x7fffffffLL) return SS2D_EINVAL;
    const size_t smem = mn_smem((int)D);
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)smem);
        if (rc != 0) return rc;
        kern<<<(unsigned)(tiles * batch), kMnWarps * kWarp, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
            y, weight, bias, z, z_pstride, out, (int)D, L, eps, (int)tiles);
        return (int)cudaGetLastError();
    };
    const int nj = (int)((D + kWarp - 1) / kWarp);
    if (nj <= 2) return go(merge_norm_fwd_kernel<2>);
    if (nj <= 4) return go(merge_norm_fwd_kernel<4>);
    if (nj <= 6) return go(merge_norm_fwd_kernel<6>);
    if (nj <= 8) return go(merge_norm_fwd_kernel<8>);
    return go(merge_norm_fwd_kernel<16>);
}

extern "C" int ss2d_merge_norm_gate_bwd(const float *y, const float *weight, const float *bias, float eps, const float *z,
                                        int64_t z_pstride, const float *dout, float *dy, float *dz, int64_t dz_pstride,
                                        float *dweight, float *dbias, int64_t batch, int64_t D, int64_t L, void *stream) {
    using namespace ss2d;
    if (!y || !weight || !bias || !dout || !dy || !dweight || !dbias || batch <= 0 || D <= 0 || L <= 0 || D > kMnMaxJ * kWarp)
        return SS2D_EINVAL;
    if (z && (z_pstride < D || (dz && dz_pstride < D))) return SS2D_EINVAL;
    const int64_t tiles = (L + kMnPix - 1) / kMnPix;
    // tiles per block: as many as keep >= ~4 blocks per SM (the dweight / dbias partial sums stay in registers across them)
    int tpb = (int)(tiles * batch / 600);
    tpb = tpb < 1 ? 1 : (tpb > 8 ? 8 : tpb);
    const int64_t blocks = (tiles * batch + tpb - 1) / tpb;
    if (blocks > 0x7fffffffLL) return SS2D_EINVAL;
    const size_t smem = mn_smem((int)D);
    auto go = [&](auto kern) -> int {
        const int rc = smem_optin(kern, (int)smem);
        if (rc != 0) return rc;
        kern<<<(unsigned)blocks, kMnWarps * kWarp, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
            y, weight, bias, z, z_pstride, dout, dy, dz, dz_pstride, dweight, dbias, (int)D, L, eps, (int)tiles, tpb, (int)batch);
        return (int)cudaGetLastError();
    };
    const int nj = (int)((D + kWarp - 1) / kWarp);
    if (nj <= 2) return go(merge_norm_bwd_kernel<2>);
    if (nj <= 4) return go(merge_norm_bwd_kernel<4>);
    if (nj <= 6) return go(merge_norm_bwd_kernel<6>);
    if (nj <= 8) return go(merge_norm_bwd_kernel<8>);
    return go(merge_norm_bwd_kernel<16>);
}
