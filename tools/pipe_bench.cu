// pipe_bench.cu — measures the per-SM issue rates the scan kernels are bounded by on this B200:
// MUFU.EX2, FFMA, FFMA+MUFU mix, SHFL, LDS.128, and global RED.ADD.F32 (scalar / v4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 2048;
constexpr int ILP = 8;

__global__ void k_mufu(float *out, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = ex2(v[i]);
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma(float *out, float seed) {
    float v[ILP]; float a = seed, b = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
// FFMA with three DISTINCT source registers per instruction (no operand reuse between neighbours)
__global__ void k_ffma3(float *out, float seed) {
    float v[ILP], a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = seed + i; a[i] = seed * 0.5f + i * 1e-3f + threadIdx.x * 1e-7f; b[i] = 1.f - 1e-3f * i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(a[i], v[(i + 3) % ILP], b[(i + 5) % ILP]);
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
// FMUL with two distinct registers
__global__ void k_fmul2(float *out, float seed) {
    float v[ILP], a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = seed + i; a[i] = 1.f - 1e-4f * i - threadIdx.x * 1e-9f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[(i + 3) % ILP] * a[i];
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
// R FFMA per MUFU, independent chains
template <int R> __global__ void k_mix(float *out, float seed) {
    float v[ILP], w[ILP]; float a = seed, b = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = seed + i; w[i] = seed * 0.1f + i * 1e-3f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            w[i] = ex2(w[i]);
#pragma unroll
            for (int r = 0; r < R; ++r) v[i] = fmaf(v[i], a, b);
        }
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i] + w[i];
    if (s == 123.456f) out[0] = s;
}
// packed fp32 (sm_100): FFMA2 / FMUL2 on register pairs
__global__ void k_ffma2(float *out, float seed) {
    float2 v[ILP]; float2 a = make_float2(seed, seed * 0.9f), b = make_float2(seed * 0.5f, seed * 0.4f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __ffma2_rn(v[i], a, b);
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}
// 1 FFMA2 + 1 FFMA + 1 MUFU interleaved (the state-lanes inner loop's mix)
__global__ void k_mix2(float *out, float seed) {
    float2 v[ILP]; float w[ILP], u[ILP]; float2 a = make_float2(seed, seed * 0.9f), b = make_float2(seed * 0.5f, seed * 0.4f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = make_float2(seed + i, seed - i); w[i] = seed * 0.1f + i * 1e-3f; u[i] = seed + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { v[i] = __ffma2_rn(v[i], a, b); w[i] = ex2(w[i]); u[i] = fmaf(u[i], a.x, b.x); }
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i].x + v[i].y + w[i] + u[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_shfl(float *out, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_lds128(float *out, float seed) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { float4 t = sm[(idx + i * 32) & 1023]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        idx += 1;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}
// global reductions: every thread adds into its own float (spread, coalesced per warp), repeated over a window
__global__ void k_red(float *buf, size_t n, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        size_t j = (i + (size_t)r * 977 * 32) % n;
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(buf + j), "f"(1.0f) : "memory");
    }
}
__global__ void k_red4(float *buf, size_t n4, int reps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        size_t j = (i + (size_t)r * 977 * 32) % n4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(buf + 4 * j), "f"(1.0f) : "memory");
    }
}

// LDS.128 and SHFL interleaved (ILP LDS.128 + 4*ILP SHFL per iteration): do shuffles share the shared-memory data pipe?
__global__ void k_lds_shfl(float *out, float seed) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    float v[4 * ILP];
#pragma unroll
    for (int i = 0; i < 4 * ILP; ++i) v[i] = seed + i + threadIdx.x;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float4 t = sm[(idx + i * 32) & 1023]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * i + j] = __shfl_up_sync(0xffffffffu, v[4 * i + j], 1);
        }
        idx += 1;
    }
    float s = acc.x + acc.y + acc.z + acc.w; for (int i = 0; i < 4 * ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

// LDS address-pattern study (round 2): does a broadcast (few distinct addresses per warp) cost fewer data-pipe cycles?
//   MODE 0: every lane its own 16 bytes (512 B per instruction)        MODE 1: all lanes the same 16 bytes
//   MODE 2: lane/8 -> 4 distinct addresses (xch pattern)                MODE 3: lane%8 -> 8 distinct addresses (B/C pattern)
//   MODE 4: lane%2 -> 2 distinct addresses                              MODE 5: lane/16 -> 2 distinct addresses
template <int MODE, typename V> __global__ void k_lds_pat(float *out, float seed) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = MODE == 0 ? lane : MODE == 1 ? 0 : MODE == 2 ? lane / 8 : MODE == 3 ? lane % 8 : MODE == 4 ? lane % 2 : lane / 16;
    idx += (threadIdx.x >> 5) * 37;
    float acc = 0.f;
    const V *base = reinterpret_cast<const V *>(sm);
    constexpr int PER = sizeof(float4) / sizeof(V);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            V t = base[((idx + i * 32) & 1023) * PER];
            if constexpr (sizeof(V) == 16) acc += (t.x + t.y) + (t.z + t.w);
            else if constexpr (sizeof(V) == 8) acc += t.x + t.y;
            else acc += t;
        }
        idx += 1;
    }
    if (acc == 123.456f) out[0] = acc;
}
// STS.128, every lane its own 16 bytes
__global__ void k_sts128(float *out, float seed) {
    __shared__ float4 sm[2048];
    float4 v = make_float4(seed, seed + threadIdx.x, seed, seed);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) sm[(idx + i * 256) & 2047] = v;
        v.x += 1.f;
    }
    __syncthreads();
    if (sm[threadIdx.x].x == 123.456f) out[0] = 1.f;
}
// packed half-precision exponentials: one MUFU instruction, two results
__global__ void k_mufu_h2(float *out, float seed) {
    unsigned v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 0x38003800u + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
    }
    unsigned s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123456u) out[0] = 1.f;
}
__global__ void k_mufu_bf2(float *out, float seed) {
    unsigned v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 0x3f003f00u + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
    }
    unsigned s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123456u) out[0] = 1.f;
}
// FSEL (ALU pipe) rate, and FSEL interleaved with FFMA (do they dual-issue on separate pipes?)
__global__ void k_fsel(float *out, float seed) {
    float v[ILP]; const bool up = (threadIdx.x & 4) != 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { const float a = v[i], b = v[(i + 3) % ILP]; v[i] = up ? a : b; }
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
template <typename F> float time_ms(F f, int n = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < n; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", pr.name, sms, clk_khz / 1000);
    float *out; CK(cudaMalloc(&out, 1 << 20));
    const int blocks = sms * 8, threads = 256;  // 64 warps/SM
    const double warp_instrs = (double)blocks * (threads / 32) * ITERS * ILP;
    auto rep = [&](const char *name, float ms, double instr_mult) {
        double wi = warp_instrs * instr_mult;
        double per_sm_per_ns = wi / sms / (ms * 1e6);
        printf("%-28s %8.3f ms  %7.3f warp-instr/ns/SM  = %6.2f lanes/clk/SM @%d MHz\n", name, ms, per_sm_per_ns,
               per_sm_per_ns * 32 / (clk_khz / 1e6), clk_khz / 1000);
    };
    rep("MUFU.EX2", time_ms([&] { k_mufu<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("FFMA (reg,reg,reg)", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("FFMA (3 distinct regs)", time_ms([&] { k_ffma3<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("FMUL (2 distinct regs)", time_ms([&] { k_fmul2<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("mix 4 FFMA : 1 MUFU (total)", time_ms([&] { k_mix<4><<<blocks, threads>>>(out, 0.5f); }), 5);
    rep("mix 6 FFMA : 1 MUFU (total)", time_ms([&] { k_mix<6><<<blocks, threads>>>(out, 0.5f); }), 7);
    rep("mix 8 FFMA : 1 MUFU (total)", time_ms([&] { k_mix<8><<<blocks, threads>>>(out, 0.5f); }), 9);
    rep("FFMA2 (packed, per instr)", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("mix FFMA2+FFMA+MUFU (total)", time_ms([&] { k_mix2<<<blocks, threads>>>(out, 0.5f); }), 3);
    rep("SHFL.UP", time_ms([&] { k_shfl<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 (conflict-free)", time_ms([&] { k_lds128<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("1 LDS.128 + 4 SHFL (total)", time_ms([&] { k_lds_shfl<<<blocks, threads>>>(out, 0.5f); }), 5);

    rep("LDS.128 lane-distinct", time_ms([&] { k_lds_pat<0, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 uniform address", time_ms([&] { k_lds_pat<1, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 4 addr (lane/8)", time_ms([&] { k_lds_pat<2, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 8 addr (lane%8)", time_ms([&] { k_lds_pat<3, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 2 addr (lane%2)", time_ms([&] { k_lds_pat<4, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.128 2 addr (lane/16)", time_ms([&] { k_lds_pat<5, float4><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.64 lane-distinct", time_ms([&] { k_lds_pat<0, float2><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.64 uniform address", time_ms([&] { k_lds_pat<1, float2><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.64 8 addr (lane%8)", time_ms([&] { k_lds_pat<3, float2><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.32 lane-distinct", time_ms([&] { k_lds_pat<0, float><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("LDS.32 uniform address", time_ms([&] { k_lds_pat<1, float><<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("STS.128 lane-distinct", time_ms([&] { k_sts128<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("MUFU.EX2 f16x2 (per instr)", time_ms([&] { k_mufu_h2<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("MUFU.EX2 bf16x2 (per instr)", time_ms([&] { k_mufu_bf2<<<blocks, threads>>>(out, 0.5f); }), 1);
    rep("FSEL", time_ms([&] { k_fsel<<<blocks, threads>>>(out, 0.5f); }), 1);
    // reductions: 64 MB window (fits L2), 16 M threads x reps
    size_t n = 16u << 20; float *buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
    for (int reps : {1, 8}) {
        int rb = 4096 * 4, rt = 256; double ops = (double)rb * rt * reps;
        float ms = time_ms([&] { k_red<<<rb, rt>>>(buf, n, reps); });
        printf("RED.ADD.F32 spread  reps=%d  %8.3f ms  %7.2f G red/s  (%6.1f GB/s payload)\n", reps, ms, ops / ms / 1e6, ops * 4 / ms / 1e6);
        ms = time_ms([&] { k_red4<<<rb, rt>>>(buf, n / 4, reps); });
        printf("RED.ADD.V4.F32      reps=%d  %8.3f ms  %7.2f G red/s  (%6.1f GB/s payload)\n", reps, ms, ops / ms / 1e6, ops * 16 / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
