// Temporary: entry points not implemented yet return SS2D_EDEVICE (removed as kernels land).
#include "../../include/ss2d_b200.h"
extern "C" int ss2d_cross_scan(const void *, void *, int64_t, int64_t, int64_t, int64_t, int32_t, void *) { return SS2D_EDEVICE; }
extern "C" int ss2d_cross_merge(const void *, void *, int64_t, int64_t, int64_t, int64_t, int32_t, void *) { return SS2D_EDEVICE; }
extern "C" int ss2d_cross_scan_fwd(const ss2d_cross_fwd_params *, void *) { return SS2D_EDEVICE; }
extern "C" int ss2d_cross_scan_bwd(const ss2d_cross_bwd_params *, void *) { return SS2D_EDEVICE; }
extern "C" int ss2d_dwconv_silu_fwd(const float *, int64_t, const float *, const float *, float *, int64_t, int64_t, int64_t, int64_t, void *) { return SS2D_EDEVICE; }
extern "C" int ss2d_dwconv_silu_bwd(const float *, int64_t, const float *, const float *, const float *, float *, int64_t, float *, float *, int64_t, int64_t, int64_t, int64_t, void *) { return SS2D_EDEVICE; }
