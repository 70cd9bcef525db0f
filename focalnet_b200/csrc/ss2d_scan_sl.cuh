// ss2d_scan_sl.cuh — the "state-lanes" organisation of the selective scan (dstate == 16), shared by
// ss2d_scan_sl_fwd.cu and ss2d_scan_sl_bwd.cu.
//
// Thread mapping ("lanes = channel x state group, time is serial"):
//   * a lane owns SN (4 or 2) of the 16 states of ONE channel and walks the whole sequence step by step, so the
//     recurrence h_t = a_t h_{t-1} + b_t is a plain serial FFMA chain — no scan, no second pass, one ex2 per
//     (element, state);  LPC = 16/SN lanes form a channel, a warp holds CPW = 32/LPC channels, a CTA NW warps of
//     the same (batch, group) so the group's B/C rows are staged once per CTA.  SN = 2 doubles the number of
//     warps for small batches (the microbench has only 6144 channels for 592 SM sub-partitions);
//   * time advances in blocks of BK = 16 steps.  Everything that is not the recurrence itself is evaluated on
//     PAIRS of consecutive steps with the packed fp32 instructions of sm_100 (FMUL2 / FFMA2 / FADD2), which
//     halves the issue slots of the discretisation (delta*A, delta*u*B) and of the C*h read-out;
//   * sums over the 16 states (y; du / ddelta in the backward) are transposed reductions over the LPC lanes of
//     a channel: 16 per-step partials enter, every lane leaves with its OWN = 16/LPC finished steps
//     [OWN*ng, OWN*ng+OWN) — one 16- or 8-byte global store per lane;
//   * u / delta (/ dout) and B / C reach the lanes through shared memory: 16-byte cp.async copies of
//     [rows x TT steps] tiles, double buffered, padded so that every LDS pattern below is conflict-free.
// Checkpoints: the forward stores h at the end of EVERY block (layout (batch, nblk, dim, 16), one coalesced
// line per warp and block); the backward recomputes a block's 16 steps from the checkpoint before it,
// keeping a_t and h_t of the block in registers, and walks the block right-to-left.  No per-step state is stored.
#pragma once

#include "ss2d_common.cuh"
#include "../../include/ss2d_b200.h"
#include <type_traits>

namespace ss2d {
namespace sl {

constexpr int kN = 16;  // dstate this organisation is compiled for
constexpr int BK = 16;  // steps per block
static_assert(BK == SS2D_SL_BLOCK, "header constant out of sync");

template <int SN> struct Map {
    static constexpr int LPC = kN / SN;      // lanes per channel
    static constexpr int CPW = kWarp / LPC;  // channels per warp
    static constexpr int OWN = BK / LPC;     // steps each lane finishes per block
    static_assert(SN == 2 || SN == 4, "2 or 4 states per lane");
    // state n lives in shared row (n % SN) * LPC + n / SN: the LPC lanes of a channel then read LPC CONSECUTIVE rows
    // for the same s, which the row padding (+16 bytes) spreads over distinct banks
    __host__ __device__ static constexpr int bc_row(int n) { return (n % SN) * LPC + n / SN; }
};

struct Flags {
    bool vec_u, vec_delta, vec_bc, vec_out, vec_z, vec_dout, vec_grad, vec_dbc;
};

// ---- shared <-> register helpers ------------------------------------------------------------------------------
// K (2 or 4) consecutive elements from shared memory -> floats
template <typename T, int K> __device__ __forceinline__ void lds_k(const T *p, float (&v)[K]) {
    static_assert(K == 2 || K == 4, "");
    if constexpr (sizeof(T) == 4) {
        if constexpr (K == 4) {
            const float4 q = *reinterpret_cast<const float4 *>(p);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
            const float2 q = *reinterpret_cast<const float2 *>(p);
            v[0] = q.x; v[1] = q.y;
        }
    } else {
        uint32_t w[K / 2];
        if constexpr (K == 4) {
            const uint2 q = *reinterpret_cast<const uint2 *>(p);
            w[0] = q.x; w[1] = q.y;
        } else {
            w[0] = *reinterpret_cast<const uint32_t *>(p);
        }
#pragma unroll
        for (int i = 0; i < K / 2; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {  // bf16: shift into the fp32 exponent position
                v[2 * i] = __uint_as_float(w[i] << 16);
                v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
                v[2 * i] = f.x; v[2 * i + 1] = f.y;
            }
        }
    }
}
// a lane's OWN (2 or 4) consecutive steps starting at the even / multiple-of-4 step offset `s0` of a tile row; `mirrored`
// rows keep step e at offset e ^ 3 (16-byte pieces copied from a sequence that runs backwards in memory)
template <typename T, int K> __device__ __forceinline__ void lds_own(const T *row, int s0, bool mirrored, float (&v)[K]) {
    float t[K];
    lds_k<T, K>(row + ((mirrored && K == 2) ? (s0 ^ 2) : s0), t);  // selects, not branches: the callers' loops stay one basic block
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = mirrored ? t[K - 1 - i] : t[i];
}
// K consecutive elements global <-> registers; `valid` = how many lie inside the sequence (<= 0: none)
template <typename T, int K> __device__ __forceinline__ void ldg_k(const T *p, float (&v)[K], int valid, bool vec) {
    if (vec && valid >= K) {
        struct alignas(K * sizeof(T)) Pack { T e[K]; };
        const Pack q = *reinterpret_cast<const Pack *>(p);
#pragma unroll
        for (int i = 0; i < K; ++i) v[i] = to_f32<T>(q.e[i]);
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) v[i] = i < valid ? to_f32<T>(p[i]) : 0.f;
    }
}
template <typename T, int K> __device__ __forceinline__ void stg_k(T *p, const float (&v)[K], int valid, bool vec) {
    if (vec && valid >= K) {
        struct alignas(K * sizeof(T)) Pack { T e[K]; };
        Pack q;
#pragma unroll
        for (int i = 0; i < K; ++i) q.e[i] = from_f32<T>(v[i]);
        *reinterpret_cast<Pack *>(p) = q;
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (i < valid) p[i] = from_f32<T>(v[i]);
    }
}

// ---- tile staging ---------------------------------------------------------------------------------------------
// Copies `nrows` rows x TT steps per pipeline stage into shared memory with 16-byte cp.async.  Every thread owns
// ONE 16-byte column piece of rows r0, r0+RPP, ...; the global pointer / shared address of its first row are
// computed once, a stage costs one add per copy.  Rows >= rows_valid (ragged last tile) are never copied — the
// kernels zero the tile buffers once at start.  Steps past L are zero-filled by the copy itself (src-size < 16).
template <typename T, int TT, int NT> struct RowStager {
    static constexpr int per = 16 / (int)sizeof(T);
    static constexpr int PPR = TT / per;  // 16-byte pieces per row
    static_assert(NT % PPR == 0, "threads must tile the row pieces");
    static constexpr int RPP = NT / PPR;  // rows per pass
    const T *src;       // row r0, this thread's piece, step 0
    int64_t src_step;   // elements between two passes
    uint32_t dst;       // shared address of (mapped row r0, piece) inside stage buffer 0
    int dst_step;       // bytes between two passes
    int npass;          // passes this thread takes part in
    int tq;             // first step of the piece inside the tile
    bool vec;
    // generic (slow) path operands for unaligned tensors
    const T *base; int64_t rstride; int nrows_valid; T *sdst; int rs; int perm_sn;

    // perm_sn: 0 = identity row map, else Map<perm_sn>::bc_row
    __device__ __forceinline__ void init(T *smem_rows, int rs_elems, const T *g, int64_t g_rstride, int nrows, int rows_valid,
                                         bool vec_ok, int perm_sn_) {
        const int r0 = threadIdx.x / PPR, q = threadIdx.x % PPR;
        vec = vec_ok; base = g; rstride = g_rstride; nrows_valid = rows_valid < nrows ? rows_valid : nrows; sdst = smem_rows;
        rs = rs_elems; perm_sn = perm_sn_;
        tq = q * per;
        src = g + (int64_t)r0 * g_rstride + tq;
        src_step = (int64_t)RPP * g_rstride;
        int row0 = r0, rstep = RPP;
        if (perm_sn_) {  // (n % SN) * LPC + n / SN ; RPP % SN == 0 keeps it affine in the pass index
            row0 = (r0 % perm_sn_) * (kN / perm_sn_) + r0 / perm_sn_;
            rstep = RPP / perm_sn_;
        }
        dst = smem_u32(smem_rows + row0 * rs_elems + tq);
        dst_step = rstep * rs_elems * (int)sizeof(T);
        npass = r0 < nrows_valid ? (nrows_valid - r0 + RPP - 1) / RPP : 0;
    }
    // issue the copies of the tile starting at step t0 into the stage buffer `buf_off` bytes after buffer 0
    __device__ __forceinline__ void issue(int t0, int L, int buf_off) const {
        if (vec) {
            const int rem = (L - (t0 + tq)) * (int)sizeof(T);
            const int bytes = rem >= 16 ? 16 : (rem > 0 ? rem : 0);
            const T *s = src + (bytes > 0 ? t0 : -tq);
            uint32_t d = dst + buf_off;
            for (int k = 0; k < npass; ++k) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(s), "r"(bytes) : "memory");
                s += src_step;
                d += dst_step;
            }
        } else {
            T *out = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(sdst) + buf_off);
            for (int idx = threadIdx.x; idx < nrows_valid * TT; idx += NT) {
                const int r = idx / TT, e = idx % TT;
                const int row = perm_sn ? (r % perm_sn) * (kN / perm_sn) + r / perm_sn : r;
                out[row * rs + e] = t0 + e < L ? base[(int64_t)r * rstride + t0 + e] : from_f32<T>(0.f);
            }
        }
    }
    // element-wise, bounds-checked fill of one stage (unaligned tensors; the short last stage of the fused seam).
    // mirror_L > 0: scan step l is element mirror_L-1-l of the row and step e of the tile is kept at offset e ^ 3 — the
    // layout CopyList::add(mirror_L) produces for the full stages
    __device__ __forceinline__ void issue_elems(int t0, int L, int buf_off, int mirror_L) const {
        T *out = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(sdst) + buf_off);
        for (int idx = threadIdx.x; idx < nrows_valid * TT; idx += NT) {
            const int r = idx / TT, e = idx % TT;
            const int row = perm_sn ? (r % perm_sn) * (kN / perm_sn) + r / perm_sn : r;
            const int l = t0 + e;
            const int64_t src_e = mirror_L > 0 ? mirror_L - 1 - l : l;
            out[row * rs + (mirror_L > 0 ? (e ^ 3) : e)] = l < L ? base[(int64_t)r * rstride + src_e] : from_f32<T>(0.f);
        }
    }
};

// Flattened per-thread copy list of one pipeline stage for the FAST kernels (every row 16-byte aligned and
// L % 16 == 0: a 16-byte piece is never ragged).  All arrays of a stage are laid end to end as 16-byte pieces and
// dealt round-robin to the CTA's threads; a stage then costs NCOPY predicated cp.async + NCOPY pointer bumps per
// thread, with no per-array loop and no address arithmetic.
template <int NT, int NCOPY> struct CopyList {
    const unsigned char *src[NCOPY];  // global address of this thread's j-th piece in the NEXT stage to issue
    uint32_t dst[NCOPY];              // shared address inside ring slot 0; 0 = no copy
    int adv[NCOPY];                   // bytes the source moves from one stage to the next (signed)
    int base;
    __device__ __forceinline__ void clear() {
        base = 0;
#pragma unroll
        for (int j = 0; j < NCOPY; ++j) { src[j] = nullptr; dst[j] = 0; adv[j] = 0; }
    }
    // rows x TTs steps of `g` (row stride g_rstride), first stage at step t_first, following stages stage_step steps on.
    // mirror_L > 0 (fused seam, directions 2 / 3): scan step l is element mirror_L-1-l of the row.  The stage's elements are
    // still copied as ascending 16-byte pieces, the piece that holds steps 4m..4m+3 lands at tile offset 4m with its four
    // values in descending step order — step e of the tile lives at offset e ^ 3 (readers un-mirror, see lds_own)
    template <typename T>
    __device__ __forceinline__ void add(T *smem_rows, int rs_elems, const T *g, int64_t g_rstride, int nrows, int rows_valid, int TTs,
                                        int perm_sn, int t_first, int stage_step, int mirror_L = 0) {
        constexpr int per = 16 / (int)sizeof(T);
        const int PPR = TTs / per;
#pragma unroll
        for (int j = 0; j < NCOPY; ++j) {
            const int local = (int)threadIdx.x + j * NT - base;
            if (local >= 0 && local < nrows * PPR) {
                const int r = local / PPR, q = local % PPR;
                if (r < rows_valid) {
                    const int row = perm_sn ? (r % perm_sn) * (kN / perm_sn) + r / perm_sn : r;
                    if (mirror_L > 0) {
                        src[j] = reinterpret_cast<const unsigned char *>(g + (int64_t)r * g_rstride + (mirror_L - t_first - TTs) + q * per);
                        dst[j] = smem_u32(smem_rows + row * rs_elems + (TTs - per - q * per));
                        adv[j] = -stage_step * (int)sizeof(T);
                    } else {
                        src[j] = reinterpret_cast<const unsigned char *>(g + (int64_t)r * g_rstride + q * per + t_first);
                        dst[j] = smem_u32(smem_rows + row * rs_elems + q * per);
                        adv[j] = stage_step * (int)sizeof(T);
                    }
                }
            }
        }
        base += nrows * PPR;
    }
    __device__ __forceinline__ void issue(int slot_off) {
#pragma unroll
        for (int j = 0; j < NCOPY; ++j) {
            if (dst[j]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst[j] + slot_off), "l"(src[j]) : "memory");
            src[j] += adv[j];
        }
    }
};

// ---- fused seam S3: CrossScan / CrossMerge as addressing ----------------------------------------------------
// Direction k maps scan step l to a pixel of the (H, W) plane (vmamba_layers.py:35-37):
//   k=0: l   k=1: (l % H) * W + l / H   k=2: L-1-l   k=3: with m = L-1-l: (m % H) * W + m / H
// DirWalk keeps the pixel offset of one scan step and moves it by whole strides without divisions.
struct DirWalk {
    int k, H, W, h, w, lin;
    __device__ __forceinline__ void init(int k_, int H_, int W_, int L, int l) {
        k = k_; H = H_; W = W_;
        const int m = (k & 2) ? L - 1 - l : l;
        lin = m;
        w = m / H_;  // floor also for negative m (never dereferenced there)
        h = m - w * H_;
        if (h < 0) { h += H_; --w; }
    }
    __device__ __forceinline__ void advance(int d) {  // l += d
        const int dm = (k & 2) ? -d : d;
        lin += dm;
        if (k & 1) {  // column walks only (the state-lanes fused path uses the linear directions on x / x^T)
            h += dm;
            while (h >= H) { h -= H; ++w; }
            while (h < 0) { h += H; --w; }
        }
    }
    __device__ __forceinline__ int off() const { return (k & 1) ? h * W + w : lin; }
};

// Accumulate OWN consecutive scan steps starting at `walk` into a spatial-order fp32 plane (CrossMerge / CrossScan^T
// as a store).  k = 0 / 2: the steps are adjacent pixels -> one vector reduction; k = 1 / 3: a column walk.
template <int OWN> __device__ __forceinline__ void red_own_cross(float *plane, const DirWalk &walk, const float (&v)[OWN]) {
    if (!(walk.k & 1)) {
        float *dst = plane + (walk.k == 0 ? walk.lin : walk.lin - (OWN - 1));
        if constexpr (OWN == 4) {
            if (walk.k == 0) red_add_v4(dst, v[0], v[1], v[2], v[3]);
            else red_add_v4(dst, v[3], v[2], v[1], v[0]);
        } else {
            const float a = walk.k == 0 ? v[0] : v[1], b = walk.k == 0 ? v[1] : v[0];
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(a), "f"(b) : "memory");
        }
    } else {
        DirWalk t = walk;
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            red_add_f32(plane + t.off(), v[i]);
            t.advance(1);
        }
    }
}

// Transposed reduction over the LPC lanes of a channel: y[16] per-step partial sums in, the lane's OWN finished
// steps [OWN*ng, OWN*ng+OWN) out.  16 - OWN shuffles for 16 sums (a butterfly per value would need 16 * log2 LPC).
template <int LPC> __device__ __forceinline__ void reduce_lanes(const float (&y)[BK], float (&r)[BK / LPC], int ng) {
    float t[BK];
#pragma unroll
    for (int j = 0; j < BK; ++j) t[j] = y[j];
#pragma unroll
    for (int m = LPC / 2, w = BK / 2; m >= 1; m >>= 1, w >>= 1) {
        const bool up = (ng & m) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
            const float send = up ? t[j] : t[j + w];
            const float keep = up ? t[j + w] : t[j];
            t[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
#pragma unroll
    for (int j = 0; j < BK / LPC; ++j) r[j] = t[j];
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// does the state-lanes organisation cover this problem?  (everything else runs on the warp-scan kernels)
bool supported(const ss2d_scan_fwd_params &p);
// test hook behind ss2d_set_default_family(): what SS2D_FAMILY_AUTO means for this process; returns the old value
int set_default_family(int family);
// states per lane for this problem size (the forward and the backward must agree: it fixes nothing in the
// checkpoint layout, but keeps one place that decides)
int states_per_lane(const ss2d_scan_fwd_params &p);
int launch_fwd(const ss2d_scan_fwd_params &p, cudaStream_t stream);
int launch_bwd(const ss2d_scan_bwd_params &p, cudaStream_t stream);
// fused seam S3 (u / out resp. u / dout / du are spatial-order fp32 planes, groups = the 4 directions).
// cross_covered(): fp32, dstate 16, L % 16 == 0 and 16-byte aligned x / delta / B / C — a rule both the forward and the
// backward can evaluate (they must pick the same kernel family: the checkpoint layouts differ); everything else runs on
// the warp-scan kernels.  The output / gradient buffers of a covered problem must be 16-byte aligned (SS2D_ESTRIDE).
// Directions 1 / 3 use the TRANSPOSED planes: uT / doutT = x^T / dy^T (inputs), accT = y^T (forward) or dx^T (backward).
struct CrossAux {
    const float *uT, *doutT;
    float *accT;
};
bool cross_covered(const ss2d_scan_fwd_params &p);
int launch_cross_fwd(const ss2d_scan_fwd_params &p, const CrossAux &aux, cudaStream_t stream);
int launch_cross_bwd(const ss2d_scan_bwd_params &p, const CrossAux &aux, cudaStream_t stream);

}  // namespace sl
}  // namespace ss2d
