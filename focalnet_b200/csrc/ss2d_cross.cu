// ss2d_cross.cu — stand-alone CrossScan / CrossMerge for sm_100a (seam S2).
//
// Replaces the Triton kernels triton_cross_scan / triton_cross_merge (reference: ITS/models/csm_triton.py:7-80)
// and their torch twins (ITS/models/vmamba_layers.py:29-71).  Scan-order position l of direction k maps to the
// spatial pixel (vmamba_layers.py:35-37):
//     k=0: l = h*W + w      k=1: l = w*H + h      k=2: l = L-1-(h*W+w)      k=3: l = L-1-(w*H+h)
// Pure data movement, HBM-bound: every 32x32 pixel tile of a (b,c) plane goes through shared memory once, so
// the row-major AND the column-major images are both read/written with full 128-byte coalescing; the flipped
// directions are the same segments walked backwards.  Bytes per launch: (1 + 4) * B*C*L * sizeof(T) — the
// algorithmic minimum for a materialised (B,4,C,L) tensor.  (The fused path, ss2d_scan_*.cu with
// ss2d_cross_*_params, never materialises it.)
#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"

namespace ss2d {

constexpr int kTile = 32;
constexpr int kRows = 8;  // blockDim = (32, 8); each thread handles 4 rows of the tile

// x:(BC,H,W) -> xs:(B,4,C,L)
template <typename T>
__global__ void __launch_bounds__(kTile *kRows) cross_scan_kernel(const T *__restrict__ x, T *__restrict__ xs, int C,
                                                                   int H, int W) {
    __shared__ T tile[kTile][kTile + 1];
    const int64_t L = (int64_t)H * W;
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    const int bc = blockIdx.x / (tiles_w * tiles_h), tile_id = blockIdx.x % (tiles_w * tiles_h);
    const int b = bc / C, c = bc % C;
    const int h0 = (tile_id / tiles_w) * kTile, w0 = (tile_id % tiles_w) * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T *src = x + (int64_t)bc * L;
    T *d0 = xs + (((int64_t)b * 4 + 0) * C + c) * L;
    T *d1 = d0 + (int64_t)C * L, *d2 = d1 + (int64_t)C * L, *d3 = d2 + (int64_t)C * L;
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int h = h0 + ty + r, w = w0 + tx;
        if (h < H && w < W) {
            const T v = src[(int64_t)h * W + w];
            tile[ty + r][tx] = v;
            const int64_t l = (int64_t)h * W + w;
            d0[l] = v;
            d2[L - 1 - l] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int w = w0 + ty + r, h = h0 + tx;  // consecutive threads walk h: contiguous in the column-major image
        if (h < H && w < W) {
            const T v = tile[tx][ty + r];
            const int64_t l = (int64_t)w * H + h;
            d1[l] = v;
            d3[L - 1 - l] = v;
        }
    }
}

// ys:(B,4,C,L) -> y:(B,C,L) spatial;  y = ((ys0 + ys1) + ys2) + ys3 un-permuted, fp32 accumulation (the order of
// triton_cross_merge, csm_triton.py:72-73).
template <typename T>
__global__ void __launch_bounds__(kTile *kRows) cross_merge_kernel(const T *__restrict__ ys, T *__restrict__ y, int C,
                                                                    int H, int W) {
    __shared__ float tile[kTile][kTile + 1];
    const int64_t L = (int64_t)H * W;
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    const int bc = blockIdx.x / (tiles_w * tiles_h), tile_id = blockIdx.x % (tiles_w * tiles_h);
    const int b = bc / C, c = bc % C;
    const int h0 = (tile_id / tiles_w) * kTile, w0 = (tile_id % tiles_w) * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T *s0 = ys + (((int64_t)b * 4 + 0) * C + c) * L;
    const T *s1 = s0 + (int64_t)C * L, *s2 = s1 + (int64_t)C * L, *s3 = s2 + (int64_t)C * L;
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int w = w0 + ty + r, h = h0 + tx;
        if (h < H && w < W) {
            const int64_t l = (int64_t)w * H + h;
            tile[tx][ty + r] = to_f32<T>(s1[l]);                 // direction 1 at pixel (h, w)
        }
    }
    __syncthreads();
    float acc[kTile / kRows];
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int h = h0 + ty + r, w = w0 + tx;
        if (h < H && w < W) {
            const int64_t l = (int64_t)h * W + w;
            acc[r / kRows] = (to_f32<T>(s0[l]) + tile[ty + r][tx]) + to_f32<T>(s2[L - 1 - l]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int w = w0 + ty + r, h = h0 + tx;
        if (h < H && w < W) tile[tx][ty + r] = to_f32<T>(s3[L - 1 - ((int64_t)w * H + h)]);  // direction 3
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int h = h0 + ty + r, w = w0 + tx;
        if (h < H && w < W) y[(int64_t)bc * L + (int64_t)h * W + w] = from_f32<T>(acc[r / kRows] + tile[ty + r][tx]);
    }
}

// Per-direction permutation ("1b1", cf. triton_cross_scan_1b1, csm_triton.py:83-120): plane (b,k,c) of src holds
// direction k's rows in SPATIAL order and is rewritten in direction k's SCAN order (INVERSE: the other way round).
// Used by the fused path on the small x_dbl tensor (R+2N = 38 rows per direction) and on its gradient.
template <typename T, bool INVERSE>
__global__ void __launch_bounds__(kTile *kRows) cross_permute_kernel(const T *__restrict__ src, T *__restrict__ dst, int C, int H,
                                                                      int W) {
    __shared__ T tile[kTile][kTile + 1];
    const int64_t L = (int64_t)H * W;
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    const int plane = blockIdx.x / (tiles_w * tiles_h), tile_id = blockIdx.x % (tiles_w * tiles_h);
    const int k = (plane / C) % 4;
    const int h0 = (tile_id / tiles_w) * kTile, w0 = (tile_id % tiles_w) * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T *s = src + (int64_t)plane * L;
    T *d = dst + (int64_t)plane * L;
    const bool flip = k >= 2;
    if ((k & 1) == 0) {  // row-major both sides: straight or mirrored copy
#pragma unroll
        for (int r = 0; r < kTile; r += kRows) {
            const int h = h0 + ty + r, w = w0 + tx;
            if (h < H && w < W) {
                const int64_t l = (int64_t)h * W + w, lp = flip ? L - 1 - l : l;
                if (INVERSE) d[l] = s[lp]; else d[lp] = s[l];
            }
        }
        return;
    }
    // column-major scan order on one side: turn the 32x32 tile through shared memory
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        if (!INVERSE) {
            const int h = h0 + ty + r, w = w0 + tx;
            if (h < H && w < W) tile[ty + r][tx] = s[(int64_t)h * W + w];
        } else {
            const int w = w0 + ty + r, h = h0 + tx;
            if (h < H && w < W) {
                const int64_t l = (int64_t)w * H + h;
                tile[tx][ty + r] = s[flip ? L - 1 - l : l];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        if (!INVERSE) {
            const int w = w0 + ty + r, h = h0 + tx;
            if (h < H && w < W) {
                const int64_t l = (int64_t)w * H + h;
                d[flip ? L - 1 - l : l] = tile[tx][ty + r];
            }
        } else {
            const int h = h0 + ty + r, w = w0 + tx;
            if (h < H && w < W) d[(int64_t)h * W + w] = tile[ty + r][tx];
        }
    }
}

template <typename T>
static int launch_permute(bool inverse, const void *in, void *out, int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    dim3 block(kTile, kRows);
    dim3 grid((unsigned)(((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * B * 4 * C));
    if (inverse)
        cross_permute_kernel<T, true><<<grid, block, 0, s>>>(static_cast<const T *>(in), static_cast<T *>(out), (int)C, (int)H, (int)W);
    else
        cross_permute_kernel<T, false><<<grid, block, 0, s>>>(static_cast<const T *>(in), static_cast<T *>(out), (int)C, (int)H, (int)W);
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_cross(bool merge, const void *in, void *out, int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t s) {
    dim3 block(kTile, kRows);
    dim3 grid((unsigned)(((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * B * C));
    if (merge)
        cross_merge_kernel<T><<<grid, block, 0, s>>>(static_cast<const T *>(in), static_cast<T *>(out), (int)C, (int)H, (int)W);
    else
        cross_scan_kernel<T><<<grid, block, 0, s>>>(static_cast<const T *>(in), static_cast<T *>(out), (int)C, (int)H, (int)W);
    return (int)cudaGetLastError();
}

static int cross_dispatch(bool merge, const void *in, void *out, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                          void *stream) {
    if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return SS2D_EINVAL;
    if (((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * B * C > 0x7fffffffLL) return SS2D_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SS2D_F32: return launch_cross<float>(merge, in, out, B, C, H, W, s);
        case SS2D_F16: return launch_cross<__half>(merge, in, out, B, C, H, W, s);
        case SS2D_BF16: return launch_cross<__nv_bfloat16>(merge, in, out, B, C, H, W, s);
        default: return SS2D_EDTYPE;
    }
}

// src: planes of (H, W) -> dst: planes of (W, H), dst = src^T or dst += src^T.  The fused seam on the state-lanes
// kernels walks directions 1 / 3 over x^T (and accumulates their outputs into y^T), so that every direction reads and
// writes CONTIGUOUS runs: one transposed copy of the 192-row x, not the reference's 4x (B,4,D,L) tensor.
template <bool ACC, bool VEC>
__global__ void __launch_bounds__(kTile *kRows) plane_transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int H, int W) {
    __shared__ float tile[kTile][kTile + 1];
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    const int64_t plane = blockIdx.x / (tiles_w * tiles_h);
    const int tile_id = blockIdx.x % (tiles_w * tiles_h);
    const int h0 = (tile_id / tiles_w) * kTile, w0 = (tile_id % tiles_w) * kTile;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const float *s = src + plane * (int64_t)H * W;
    float *d = dst + plane * (int64_t)H * W;
#pragma unroll
    for (int r = 0; r < kTile; r += kRows) {
        const int h = h0 + ty + r, w = w0 + tx;
        tile[ty + r][tx] = (h < H && w < W) ? s[(int64_t)h * W + w] : 0.f;
    }
    __syncthreads();
    if constexpr (VEC) {
        // H % 4 == 0: every thread moves 4 consecutive h of one destination row as one 128-bit access; 8 lanes cover a
        // 128-byte line, a warp 4 destination rows.  tile[h4 + i][wl]: bank = (h4 + i + wl) % 32 — conflict-free.
        const int t = ty * kTile + tx, wl = t / 8, h4 = (t % 8) * 4;
        const int w = w0 + wl, h = h0 + h4;
        if (w < W && h < H) {
            float4 v = make_float4(tile[h4][wl], tile[h4 + 1][wl], tile[h4 + 2][wl], tile[h4 + 3][wl]);
            float4 *q = reinterpret_cast<float4 *>(d + (int64_t)w * H + h);
            if (ACC) { const float4 o = *q; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *q = v;
        }
    } else {
#pragma unroll
        for (int r = 0; r < kTile; r += kRows) {
            const int w = w0 + ty + r, h = h0 + tx;
            if (h < H && w < W) {
                float *q = d + (int64_t)w * H + h;
                if (ACC) *q += tile[tx][ty + r];
                else *q = tile[tx][ty + r];
            }
        }
    }
}

int plane_transpose(const float *src, float *dst, int64_t planes, int H, int W, bool acc, cudaStream_t stream) {
    const int64_t tiles = (int64_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    if (planes * tiles > 0x7fffffffLL) return SS2D_EINVAL;
    const dim3 blk(kTile, kRows);
    const bool vec = H % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const unsigned grid = (unsigned)(planes * tiles);
    if (acc) {
        if (vec) plane_transpose_kernel<true, true><<<grid, blk, 0, stream>>>(src, dst, H, W);
        else plane_transpose_kernel<true, false><<<grid, blk, 0, stream>>>(src, dst, H, W);
    } else {
        if (vec) plane_transpose_kernel<false, true><<<grid, blk, 0, stream>>>(src, dst, H, W);
        else plane_transpose_kernel<false, false><<<grid, blk, 0, stream>>>(src, dst, H, W);
    }
    return (int)cudaGetLastError();
}

}  // namespace ss2d

extern "C" int ss2d_plane_transpose(const float *src, float *dst, int64_t planes, int64_t H, int64_t W, int32_t accumulate, void *stream) {
    if (!src || !dst || planes <= 0 || H <= 0 || W <= 0 || H > 0x7fffffffLL || W > 0x7fffffffLL) return SS2D_EINVAL;
    return ss2d::plane_transpose(src, dst, planes, (int)H, (int)W, accumulate != 0, reinterpret_cast<cudaStream_t>(stream));
}

namespace ss2d {
}  // namespace ss2d

extern "C" int ss2d_cross_scan(const void *x, void *xs, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype, void *stream) {
    return ss2d::cross_dispatch(false, x, xs, B, C, H, W, dtype, stream);
}
extern "C" int ss2d_cross_merge(const void *ys, void *y, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype, void *stream) {
    return ss2d::cross_dispatch(true, ys, y, B, C, H, W, dtype, stream);
}

extern "C" int ss2d_cross_permute(const void *src, void *dst, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                                  int32_t inverse, void *stream) {
    using namespace ss2d;
    if (!src || !dst || B <= 0 || C <= 0 || H <= 0 || W <= 0) return SS2D_EINVAL;
    if (((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile) * B * 4 * C > 0x7fffffffLL) return SS2D_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SS2D_F32: return launch_permute<float>(inverse != 0, src, dst, B, C, H, W, s);
        case SS2D_F16: return launch_permute<__half>(inverse != 0, src, dst, B, C, H, W, s);
        case SS2D_BF16: return launch_permute<__nv_bfloat16>(inverse != 0, src, dst, B, C, H, W, s);
        default: return SS2D_EDTYPE;
    }
}
