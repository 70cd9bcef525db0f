// ss2d_abi.cu — version / diagnostics entry points of libss2d_b200.so (see include/ss2d_b200.h).
#include <cuda_runtime.h>
#include "../../include/ss2d_b200.h"
#include "ss2d_scan_sl.cuh"
#define SS2D_STR_(x) #x
#define SS2D_STR(x) SS2D_STR_(x)

#include <map>
#include <mutex>
#include <utility>

namespace ss2d {
int smem_optin_impl(const void *kern, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> done;  // (kernel, device) -> bytes already opted in
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    int &have = done[{kern, dev}];
    if (bytes > have) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return (int)e;
        have = bytes;
    }
    return 0;
}
}  // namespace ss2d

extern "C" int ss2d_abi_version(void) { return SS2D_ABI_VERSION; }

extern "C" const char *ss2d_build_info(void) {
    return "libss2d_b200 abi=3 arch=sm_100a kernels=scan_sl_fwd,scan_sl_bwd,scan_fwd,scan_bwd,cross_scan,cross_merge,cross_scan_fused,dwconv_silu,merge_norm_gate,dt_proj,optim_clip_adam "
           "cuda=" SS2D_STR(__CUDACC_VER_MAJOR__) "." SS2D_STR(__CUDACC_VER_MINOR__);
}

extern "C" int64_t ss2d_scan_ckpt_floats(int64_t batch, int64_t dim, int64_t seqlen, int64_t dstate) {
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
    const int64_t coarse = batch * dim * ((seqlen + SS2D_CKPT_STEPS - 1) / SS2D_CKPT_STEPS) * dstate;
    const int64_t fine = dstate == 16 ? batch * dim * ((seqlen + SS2D_SL_BLOCK - 1) / SS2D_SL_BLOCK) * dstate : 0;
    return fine > coarse ? fine : coarse;  // one size serves whichever kernel family takes the shape
}

extern "C" int ss2d_scan_family(const ss2d_scan_fwd_params *p) {
    if (!p) return SS2D_EINVAL;
    return ss2d::sl::supported(*p) ? SS2D_FAMILY_STATELANES : SS2D_FAMILY_WARPSCAN;
}

extern "C" int ss2d_set_default_family(int family) { return ss2d::sl::set_default_family(family); }

// the fused seam runs on the state-lanes kernels for fp32 / dstate 16 / L % 16 == 0 problems that `supported` takes
static bool cross_on_statelanes(int64_t batch, int64_t D, int64_t H, int64_t W, int64_t dstate, int32_t in_dtype, int32_t family) {
    if (batch <= 0 || D <= 0 || H <= 0 || W <= 0 || in_dtype != SS2D_F32 || (H * W) % SS2D_SL_BLOCK != 0) return false;
    ss2d_scan_fwd_params p{};
    p.batch = batch; p.dim = 4 * D; p.seqlen = H * W; p.dstate = dstate; p.ngroups = 4; p.family = family;
    return ss2d::sl::supported(p);
}

extern "C" int ss2d_cross_family(int64_t batch, int64_t D, int64_t H, int64_t W, int64_t dstate, int32_t in_dtype, int32_t family) {
    return cross_on_statelanes(batch, D, H, W, dstate, in_dtype, family) ? SS2D_FAMILY_STATELANES : SS2D_FAMILY_WARPSCAN;
}

extern "C" int64_t ss2d_cross_work_floats(int64_t batch, int64_t D, int64_t H, int64_t W, int64_t dstate, int32_t in_dtype,
                                          int32_t backward, int32_t family) {
    if (!cross_on_statelanes(batch, D, H, W, dstate, in_dtype, family)) return 0;
    return (backward ? 3 : 2) * batch * D * H * W;
}

extern "C" const char *ss2d_error_string(int code) {
    switch (code) {
        case SS2D_OK: return "ok";
        case SS2D_EINVAL: return "invalid argument (null pointer, non-positive size, dim % ngroups != 0, dstate > 256)";
        case SS2D_EDTYPE: return "unsupported dtype combination (out/dout must be f32 or equal to the input dtype)";
        case SS2D_ESTRIDE: return "last-dimension stride must be 1 / x, ckpt, work and the fused seam's buffers must be 16-byte aligned";
        case SS2D_EDEVICE: return "no sm_100 device";
        default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown ss2d error";
    }
}
