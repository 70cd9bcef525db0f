// ss2d_common.cuh — device helpers shared by the SS2D kernels (sm_100a).
//
// Thread mapping used by every scan kernel in this library ("lanes = time, warps = channels"):
//   * one warp owns one channel sequence (b, c); its 32 lanes own 32 consecutive blocks of T
//     timesteps, so a warp covers a chunk of 32*T steps and walks the sequence chunk by chunk
//     carrying the N-state running prefix;
//   * the warps of a CTA are NW channels of the SAME (batch, group), marching over the chunks in
//     lock-step, so the group's B/C rows are staged once per CTA in shared memory (cp.async,
//     ping-pong over blocks of 8 states) instead of being re-read from L2 by every channel.
// The recurrence h_t = a_t h_{t-1} + b_t is resolved with a thread-local pass, one Kogge-Stone
// warp-shuffle scan of the (prod a, h) block aggregates, and a second thread-local pass that
// produces the outputs — the reference does the same with cub::BlockScan plus a block-wide
// barrier per state (selective_scan_fwd_kernel_oflex.cuh:132-171); here nothing crosses a warp.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ss2d {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;

// ---- numerics (same functions the reference evaluates, selective_scan_fwd_kernel_oflex.cuh:123-145) ----
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with the reference's cut-off (x <= 20 ? log1p(exp x) : x, fwd_kernel_oflex.cuh:124-126).
// log1p(t), t = e^x: for t < 1/4 the series 2*atanh(t/(2+t)) (|error| < 3e-7 relative, where lg2.approx
// would lose relative accuracy near 1); otherwise ln2*lg2(1+t).  ~14 instructions instead of log1pf's ~30.
__device__ __forceinline__ float softplus_ref(float x) {
    const float t = ex2(x * kLog2e);
    // s = t / (2 + t), only used for t < 1/4: 1/(2+t) on [2, 2.25] from a linear guess (0.17 % off) and two Newton
    // steps on the FMA pipe (error 5e-12) — the scan kernels are MUFU-bound, a MUFU.RCP here costs as much as an ex2
    const float d = 2.f + t;
    float rc = fmaf(t, -0.2222222f, 0.4991830f);
    rc = fmaf(rc, fmaf(-d, rc, 1.f), rc);
    rc = fmaf(rc, fmaf(-d, rc, 1.f), rc);
    const float s = t * rc, s2 = s * s;
    const float small = 2.f * s * fmaf(s2, fmaf(s2, fmaf(s2, 1.f / 7.f, 0.2f), 1.f / 3.f), 1.f);
    const float big = 0.69314718055994531f * lg2(1.f + t);
    const float r = t < 0.25f ? small : big;
    return x <= 20.f ? r : x;
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ---- dtype conversion -------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of elements of type T
template <typename T> struct Vec16 { static constexpr int n = 16 / sizeof(T); };

// Unpack a 16-byte register quad into 16/sizeof(T) floats.
template <typename T> __device__ __forceinline__ void unpack16(const uint4 &q, float *dst);
template <> __device__ __forceinline__ void unpack16<float>(const uint4 &q, float *dst) {
    dst[0] = __uint_as_float(q.x); dst[1] = __uint_as_float(q.y);
    dst[2] = __uint_as_float(q.z); dst[3] = __uint_as_float(q.w);
}
template <> __device__ __forceinline__ void unpack16<__half>(const uint4 &q, float *dst) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); dst[2 * i] = f.x; dst[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4 &q, float *dst) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { dst[2 * i] = __uint_as_float(w[i] << 16); dst[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <typename T> __device__ __forceinline__ uint4 pack16(const float *src);
template <> __device__ __forceinline__ uint4 pack16<float>(const float *src) {
    return make_uint4(__float_as_uint(src[0]), __float_as_uint(src[1]), __float_as_uint(src[2]), __float_as_uint(src[3]));
}
template <> __device__ __forceinline__ uint4 pack16<__half>(const float *src) {
    uint4 q; __half2 *h = reinterpret_cast<__half2 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(src[2 * i], src[2 * i + 1]);
    return q;
}
template <> __device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float *src) {
    uint4 q; __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(src[2 * i], src[2 * i + 1]);
    return q;
}

// ---- global <-> register block I/O: each lane owns T consecutive elements -------------------------
// `valid` = number of elements of this lane's block that lie inside the sequence (<=0: none).
// vec: the row base is 16-byte aligned and the block start is a multiple of 16 bytes.
template <typename T, int N>
__device__ __forceinline__ void load_block(const T *__restrict__ p, float (&v)[N], int valid, bool vec, float fill = 0.f) {
    constexpr int per = Vec16<T>::n;
    if (vec && valid >= N) {
#pragma unroll
        for (int i = 0; i < N / per; ++i) {
            uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            unpack16<T>(q, &v[i * per]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = i < valid ? to_f32<T>(p[i]) : fill;
    }
}
template <typename T, int N>
__device__ __forceinline__ void store_block(T *__restrict__ p, const float (&v)[N], int valid, bool vec) {
    constexpr int per = Vec16<T>::n;
    if (vec && valid >= N) {
#pragma unroll
        for (int i = 0; i < N / per; ++i) reinterpret_cast<uint4 *>(p)[i] = pack16<T>(&v[i * per]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) if (i < valid) p[i] = from_f32<T>(v[i]);
    }
}

// ---- CrossScan / CrossMerge folded into the scan's addressing (fused seam S3) -----------------------
// Direction k maps scan step l to a pixel of the (H, W) plane (vmamba_layers.py:35-37):
//   k=0: l = h*W+w   k=1: l = w*H+h   k=2: l = L-1-(h*W+w)   k=3: l = L-1-(w*H+h)
// A lane's T consecutive scan steps are therefore T contiguous pixels (k=0), the same run mirrored (k=2), or a
// column walk with stride W (k=1,3).  For the column walk, lanes H/T apart sit on adjacent columns, so one
// warp-wide scalar access still touches whole 32-byte sectors (4 rows x 8 pixels for H=64, T=16).
struct CrossInfo {
    int H, W;
    int g_only;  // < 0: the grid covers all groups; else one launch serves direction g_only alone (deterministic merge order)
};
// (group, batch) of a CTA: blockIdx.x / tiles_per_group enumerates (batch, group) pairs, or batches when one group is pinned
__device__ __forceinline__ void cta_group_batch(const CrossInfo &ci, int bg, int ngroups, int &g, int &b) {
    if (ci.g_only >= 0) { g = ci.g_only; b = bg; }
    else { g = bg % ngroups; b = bg / ngroups; }
}
struct CrossWalk {  // pixel offset of scan step l for k in {1,3}, advanced one step at a time
    int h, w, W, H, dir;
    __device__ __forceinline__ CrossWalk(int k, int64_t l, int64_t L, int H_, int W_) : W(W_), H(H_) {
        const int64_t ls = k == 3 ? L - 1 - l : l;  // column-major index w*H + h
        w = (int)(ls / H_);
        h = (int)(ls - (int64_t)w * H_);
        dir = k == 3 ? -1 : 1;
    }
    __device__ __forceinline__ int64_t offset() const { return (int64_t)h * W + w; }
    __device__ __forceinline__ void next() {
        h += dir;
        if (h == H) { h = 0; ++w; }
        if (h < 0) { h = H - 1; --w; }
    }
};

// Gather this lane's T scan steps [tl, tl+T) of direction k from a spatial-order plane.
template <typename T_, int N>
__device__ __forceinline__ void load_block_cross(const T_ *__restrict__ plane, float (&v)[N], int k, int64_t tl, int64_t L,
                                                 const CrossInfo &ci, bool vec) {
    const int valid = (int)min((int64_t)N, L - tl);
    if (k == 0) {
        load_block<T_, N>(plane + tl, v, valid, vec);
    } else if (k == 2) {
        float t[N];
        const int64_t start = L - tl - N;  // the mirrored run [start, start+N)
        if (vec && valid >= N && (start * (int64_t)sizeof(T_)) % 16 == 0) {
            load_block<T_, N>(plane + start, t, N, true);
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = t[N - 1 - i];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = i < valid ? to_f32<T_>(plane[L - 1 - (tl + i)]) : 0.f;
        }
    } else {
        CrossWalk cw(k, min(tl, L - 1), L, ci.H, ci.W);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            v[i] = i < valid ? to_f32<T_>(__ldg(plane + cw.offset())) : 0.f;
            cw.next();
        }
    }
}

__device__ __forceinline__ void red_add_f32(float *addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_if(bool p, float *addr, float a, float b, float c, float d) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "@q red.global.add.v4.f32 [%1], {%2, %3, %4, %5};\n\t"
        "}" ::"r"((int)p), "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
        : "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Accumulate this lane's T scan steps of direction k into a spatial-order fp32 plane (CrossMerge as a store).
template <int N>
__device__ __forceinline__ void red_block_cross(float *__restrict__ plane, const float (&v)[N], int k, int64_t tl, int64_t L,
                                                const CrossInfo &ci, bool vec) {
    const int valid = (int)min((int64_t)N, L - tl);
    if (k == 0 || k == 2) {
        const int64_t start = k == 0 ? tl : L - tl - N;
        if (vec && valid >= N && (start & 3) == 0) {
#pragma unroll
            for (int i = 0; i < N; i += 4) {
                if (k == 0) red_add_v4(plane + start + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                else red_add_v4(plane + start + i, v[N - 1 - i], v[N - 2 - i], v[N - 3 - i], v[N - 4 - i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (i < valid) red_add_f32(plane + (k == 0 ? tl + i : L - 1 - (tl + i)), v[i]);
        }
    } else {
        CrossWalk cw(k, min(tl, L - 1), L, ci.H, ci.W);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i < valid) red_add_f32(plane + cw.offset(), v[i]);
            cw.next();
        }
    }
}

// ---- cp.async (LDGSTS) -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// 16-byte async copy, `bytes` (0..16) taken from global, remainder zero-filled
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- shared-memory row layout for lane-blocked reads ----------------------------------------------
// A row holds `chunk = 32*T` elements.  Lane j reads elements [j*T, (j+1)*T) with 16-byte LDS.  To keep
// the eight lanes of an LDS.128 phase on distinct bank groups the stride between lane blocks (in
// 16-byte units) must be odd, so every block is followed by one 16-byte pad when needed.
template <typename T, int TT> struct RowLayout {
    static constexpr int per = Vec16<T>::n;                 // elements per 16 B
    static constexpr int units = TT / per;                  // 16-B units per lane block
    static constexpr int stride_units = (units % 2 == 0) ? units + 1 : units;
    static constexpr int row_units = 32 * stride_units;
    static constexpr int row_bytes = row_units * 16;
    // 16-B unit index inside the padded row of the q-th unpadded 16-B piece
    __device__ __forceinline__ static int unit_of_piece(int q) { return (q / units) * stride_units + (q % units); }
    __device__ __forceinline__ static int lane_unit(int lane) { return lane * stride_units; }
};

// Read this lane's T elements of a padded smem row into floats.
template <typename T, int TT>
__device__ __forceinline__ void lds_block(const unsigned char *row, int lane, float (&v)[TT]) {
    using RL = RowLayout<T, TT>;
    const uint4 *src = reinterpret_cast<const uint4 *>(row) + RL::lane_unit(lane);
#pragma unroll
    for (int i = 0; i < RL::units; ++i) unpack16<T>(src[i], &v[i * RL::per]);
}

// ---- shared memory by 32-bit address + compile-time immediate ------------------------------------------
// The scan kernels are bound by instruction issue; addressing every shared operand as
// [one base register + immediate] keeps integer address arithmetic out of the state loop.
template <int OFF> __device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
// Predicated store: inline asm inside an `if` forces a real branch (BRA + BSSY/BSYNC reconvergence) around one
// instruction; carrying the predicate into the asm keeps the state loop branch-free.
template <int OFF> __device__ __forceinline__ void sts_f32_if(bool p, uint32_t a, float v) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "@q st.shared.f32 [%1+%2], %3;\n\t"
        "}" ::"r"((int)p), "r"(a), "n"(OFF), "f"(v)
        : "memory");
}
template <int OFF> __device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 q;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(a), "n"(OFF));
    return q;
}
template <int OFF> __device__ __forceinline__ void sts_v4(uint32_t a, const uint4 &q) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2,%3,%4,%5};" ::"r"(a), "n"(OFF), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
}
// this lane's TT elements of the padded row that starts OFF bytes after `a` (a already includes the lane offset)
template <typename T, int TT, int OFF> __device__ __forceinline__ void lds_row(uint32_t a, float (&v)[TT]) {
    using RL = RowLayout<T, TT>;
    static_assert(RL::units <= 4, "lane block larger than 64 bytes");
    if constexpr (RL::units >= 1) unpack16<T>(lds_v4<OFF + 0>(a), &v[0 * RL::per]);
    if constexpr (RL::units >= 2) unpack16<T>(lds_v4<OFF + 16>(a), &v[1 * RL::per]);
    if constexpr (RL::units >= 3) unpack16<T>(lds_v4<OFF + 32>(a), &v[2 * RL::per]);
    if constexpr (RL::units >= 4) unpack16<T>(lds_v4<OFF + 48>(a), &v[3 * RL::per]);
}

// ---- the scan combine (a0,b0) o (a1,b1) = (a1 a0, a1 b0 + b1)  (selective_scan_common.h:93-95) ----
// Inclusive Kogge-Stone scan over the 32 lanes of the block aggregates (P = prod a, H = local h_end).
// shfl.sync returns an "in range" predicate together with the data, so each round is 2 SHFL + 2 predicated
// FP instructions with no persistent predicate registers (the backward runs two scans per state and was
// spilling its ten `lane >= d` predicates into a general register).
template <int D> __device__ __forceinline__ void scan_round_up(float &P, float &H) {
    asm(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .f32 hp, pp;\n\t"
        "shfl.sync.up.b32 hp|q, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 pp, %1, %2, 0, 0xffffffff;\n\t"
        "@q fma.rn.f32 %0, %1, hp, %0;\n\t"
        "@q mul.f32 %1, %1, pp;\n\t"
        "}"
        : "+f"(H), "+f"(P)
        : "n"(D));
}
template <int D> __device__ __forceinline__ void scan_round_down(float &P, float &H) {
    asm(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .f32 hp, pp;\n\t"
        "shfl.sync.down.b32 hp|q, %0, %2, 0x1f, 0xffffffff;\n\t"
        "shfl.sync.down.b32 pp, %1, %2, 0x1f, 0xffffffff;\n\t"
        "@q fma.rn.f32 %0, %1, hp, %0;\n\t"
        "@q mul.f32 %1, %1, pp;\n\t"
        "}"
        : "+f"(H), "+f"(P)
        : "n"(D));
}
__device__ __forceinline__ void warp_scan_inclusive(float &P, float &H, int /*lane*/) {
    scan_round_up<1>(P, H); scan_round_up<2>(P, H); scan_round_up<4>(P, H); scan_round_up<8>(P, H); scan_round_up<16>(P, H);
}
// Mirror image for the backward's suffix scan: lane j combines with lanes j+d.
__device__ __forceinline__ void warp_rscan_inclusive(float &P, float &H, int /*lane*/) {
    scan_round_down<1>(P, H); scan_round_down<2>(P, H); scan_round_down<4>(P, H); scan_round_down<8>(P, H); scan_round_down<16>(P, H);
}
// Exclusive neighbour (identity (1, 0) at the warp edge), again with the shuffle's own predicate.
__device__ __forceinline__ void shift_up1(float P, float H, float &Pe, float &He) {
    asm(
        "{\n\t"
        ".reg .pred q;\n\t"
        "shfl.sync.up.b32 %0|q, %2, 1, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 %1, %3, 1, 0, 0xffffffff;\n\t"
        "@!q mov.f32 %0, 0f00000000;\n\t"
        "@!q mov.f32 %1, 0f3f800000;\n\t"
        "}"
        : "=&f"(He), "=&f"(Pe)
        : "f"(H), "f"(P));
}
__device__ __forceinline__ void shift_down1(float P, float H, float &Pe, float &He) {
    asm(
        "{\n\t"
        ".reg .pred q;\n\t"
        "shfl.sync.down.b32 %0|q, %2, 1, 0x1f, 0xffffffff;\n\t"
        "shfl.sync.down.b32 %1, %3, 1, 0x1f, 0xffffffff;\n\t"
        "@!q mov.f32 %0, 0f00000000;\n\t"
        "@!q mov.f32 %1, 0f3f800000;\n\t"
        "}"
        : "=&f"(He), "=&f"(Pe)
        : "f"(H), "f"(P));
}

// Dynamic shared memory opt-in, raised at most once per (kernel, device, size): cudaFuncSetAttribute is a driver round trip
// that has no business on every launch.  Keyed by the kernel's ADDRESS (kernels of equal signature share the pointer TYPE,
// so a static per template instantiation would be shared between them).
int smem_optin_impl(const void *kern, int bytes);
template <typename K> static inline int smem_optin(K kern, int bytes) { return smem_optin_impl(reinterpret_cast<const void *>(kern), bytes); }

}  // namespace ss2d

