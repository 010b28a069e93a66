// fic_device.cuh -- device-side helpers shared by the kernels of libfic_b200.
//
// Geometry follows the reference exactly (file:line refer to
// src/bvk_ss19/FractalCompression.java = FC); float helpers pin Java semantics:
// binary32 with round-to-nearest on every operation, never contracted into FMA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fic_internal.h"

namespace fic {

// FC:516-545 getDomainBlockIndex: index of the domain block "above" range (xr, yr).
__host__ __device__ inline int domain_block_index(int xr, int yr, int rpw, int rph, int dpw)
{
    if (yr == 0) yr = 1;
    if (xr == 0) xr = 1;
    if (yr == rph - 1) yr = yr - 1;
    if (xr == rpw - 1) xr = xr - 1;
    int i = 0;
    if (xr > 1) {
        i = (yr == 0) ? xr : (xr * 2) - 2 + (yr + yr - 1) * dpw;
    } else if (xr == 1) {
        i = (yr == 0) ? xr : xr + (yr + yr - 1) * dpw;
    }
    return i;
}

// FC:84-100 generateKernel (= FC:868-879): clamped origin of the wk x wk window.
__host__ __device__ inline void window_origin(int index, int dpw, int dph, int wk, int *dy, int *dx)
{
    int y = index / dpw - wk / 2;
    int x = index % dpw - wk / 2;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    if (x + wk >= dpw) x = dpw - wk;
    if (y + wk >= dph) y = dph - wk;
    *dy = y;
    *dx = x;
}

// Window origin of range block j (raster order).
__host__ __device__ inline void range_window(const Geom &g, int64_t j, int *dy, int *dx)
{
    int xr = (int)(j % g.rpw), yr = (int)(j / g.rpw);
    int i = domain_block_index(xr, yr, g.rpw, g.rph, g.dpw);
    window_origin(i, g.dpw, g.dph, g.wk, dy, dx);
}

// Isometry extension (not in the reference): T_k maps range pixel (ry, rx) of a B x B block to the domain
// pixel (sy, sx) it is compared with / reconstructed from.  0 identity, 1-3 rotations, 4 mirror x, 5 mirror y,
// 6 transpose, 7 anti-transpose.
__host__ __device__ inline void iso_map(int k, int B, int ry, int rx, int *sy, int *sx)
{
    const int m = B - 1;
    switch (k & 7) {
    case 0: *sy = ry;     *sx = rx;     break;
    case 1: *sy = m - rx; *sx = ry;     break;
    case 2: *sy = m - ry; *sx = m - rx; break;
    case 3: *sy = rx;     *sx = m - ry; break;
    case 4: *sy = ry;     *sx = m - rx; break;
    case 5: *sy = m - ry; *sx = rx;     break;
    case 6: *sy = rx;     *sx = ry;     break;
    default: *sy = m - rx; *sx = m - ry; break;
    }
}
// iso_map(iso_inverse(k)) undoes iso_map(k): the rotations by 90 and 270 degrees swap, the rest are involutions.
__host__ __device__ inline int iso_inverse(int k) { return k == 1 ? 3 : (k == 3 ? 1 : k); }

#ifdef __CUDACC__

// Java (int)float: truncate toward zero, saturate, NaN -> 0 == cvt.rzi.s32.f32.
__device__ __forceinline__ int j_f2i(float f) { return __float2int_rz(f); }

__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// The reference's grey "error" of one candidate (FC:677-683), from exact integers:
// kov = sum (r-rmean)(d-dmean), vR = sum (r-rmean), varD = sum (d-dmean)^2.
// sqd must be the correctly rounded double sqrt of varD.
__device__ __forceinline__ float grey_error(int kov, int vR, double sqd)
{
    float fvR = (float)vR;
    float r = 0.0f;
    if (!(fvR == 0.0f || sqd == 0.0)) {
        double den = __dmul_rn((double)fvR, sqd);
        r = __double2float_rn(__ddiv_rn((double)(float)kov, den));
    }
    r = __fmul_rn(r, r);
    return __fmul_rn(__fmul_rn(fvR, fvR), __fsub_rn(1.0f, r));
}

// varD = sum d^2 - 2*dmean*sum d + n*dmean^2 with dmean = floor(sum d / n) (DB:92-115).
__device__ __forceinline__ int dom_var(int dsum, int dsq, int n, int *dmean)
{
    int m = dsum / n;
    *dmean = m;
    return dsq - m * (2 * dsum - n * m);
}

#endif  // __CUDACC__

}  // namespace fic
