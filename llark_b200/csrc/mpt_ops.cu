// HBM-bound kernels of the CLAP + MPT-1B variant that the Llama path does not have:
//   bias-free LayerNorm forward / backward over bf16 rows with a row stride (so the q and k thirds of the packed qkv
//   buffer are normalised in place of a copy): LPLayerNorm of m2t/llava/model/mpt/norm.py:26-52 (norm_1, norm_2, norm_f)
//   and the q_ln / k_ln LayerNorms of attention.py:248-251,289-292 -- both are "fp32 statistics, one rounding of the
//   output to bf16", so one kernel serves both;
//   exact GELU forward / backward (blocks.py:14, nn.GELU(approximate="none")).
// Wired through llark_b200.ops / engine_mpt.py; per-op parity against fp32 references and the whole engine step against the
// pinned oracle run on a B200 in tests/test_mpt_gpu.py (round 2).
#include "common.cuh"
#include <math.h>

namespace llark {
extern thread_local int g_last_cuda_error;
namespace mpt {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = o;
}
// sum over the CTA (<= 1024 threads); every thread gets the result. `red` is reused: callers sync through this function.
__device__ __forceinline__ float cta_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();  // previous use of `red` is over
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (threadIdx.x < (unsigned)nw) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// y[row] = bf16(((x[row] - mean) * rstd) * w), W == blockDim.x * 8 * VPT; one CTA per row.
template <int VPT>
__global__ void ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                              __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out, int W, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const __nv_bfloat16* xr = x + row * ldx;
    float v[VPT][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        ld8(xr + (i * blockDim.x + threadIdx.x) * 8, v[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    const float mean = cta_sum(s, red) / (float)W;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float dlt = v[i][e] - mean;
            ss += dlt * dlt;
        }
    const float rstd = rsqrtf(cta_sum(ss, red) / (float)W + eps);  // two-pass variance: no cancellation
    if (threadIdx.x == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int c = (i * blockDim.x + threadIdx.x) * 8;
        float wv[8], o[8];
        ld8(w + c, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ((v[i][e] - mean) * rstd) * wv[e];
        st8(y + row * ldy + c, o);
    }
}

// dx = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w,  xhat = (x - mean) * rstd
// dw_partial[cta] += dy * xhat  (each CTA walks rows cta, cta + grid, ...; a second kernel sums the partials)
template <int VPT>
__global__ void ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                              const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ mean_in,
                              const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dres,
                              __nv_bfloat16* __restrict__ dx, long long lddx, float* __restrict__ dw_partial,
                              long long rows, int W) {
    __shared__ float red[32];
    float dwacc[VPT][8], wv[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        ld8(w + (i * blockDim.x + threadIdx.x) * 8, wv[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) dwacc[i][e] = 0.f;
    }
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float mean = mean_in[row], rstd = rstd_in[row];
        float xh[VPT][8], g[VPT][8];
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = (i * blockDim.x + threadIdx.x) * 8;
            float xv[8], dyv[8];
            ld8(x + row * ldx + c, xv);
            ld8(dy + row * lddy + c, dyv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                xh[i][e] = (xv[e] - mean) * rstd;
                g[i][e] = dyv[e] * wv[i][e];
                sg += g[i][e];
                sgx += g[i][e] * xh[i][e];
                dwacc[i][e] += dyv[e] * xh[i][e];
            }
        }
        const float mg = cta_sum(sg, red) / (float)W;
        const float mgx = cta_sum(sgx, red) / (float)W;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = (i * blockDim.x + threadIdx.x) * 8;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = rstd * (g[i][e] - mg - xh[i][e] * mgx);
            if (dres) {
                float r[8];
                ld8(dres + row * (long long)W + c, r);  // the residual-stream gradient is always a dense [rows, W] tensor
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += r[e];
            }
            st8(dx + row * lddx + c, o);
        }
    }
    if (dw_partial) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            float* dst = dw_partial + (long long)blockIdx.x * W + (i * blockDim.x + threadIdx.x) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[e] = dwacc[i][e];
        }
    }
}

// dw[c] (+)= sum_p partial[p][c]
__global__ void ln_dw_reduce_kernel(const float* __restrict__ partial, int n_partial, int W,
                                    __nv_bfloat16* __restrict__ dw, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= W) return;
    float s = accumulate ? __bfloat162float(dw[c]) : 0.f;
    for (int p = 0; p < n_partial; ++p) s += partial[(long long)p * W + c];
    dw[c] = __float2bfloat16_rn(s);
}

__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ h, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float v[8], o[8];
        ld8(u + i * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.5f * v[e] * (1.f + erff(v[e] * 0.70710678118654752f));
        st8(h + i * 8, o);
    }
}
// du = dh * (Phi(u) + u * phi(u))
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ dh,
                                __nv_bfloat16* __restrict__ du, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float v[8], g[8], o[8];
        ld8(u + i * 8, v);
        ld8(dh + i * 8, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float cdf = 0.5f * (1.f + erff(v[e] * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * __expf(-0.5f * v[e] * v[e]);
            o[e] = g[e] * (cdf + v[e] * pdf);
        }
        st8(du + i * 8, o);
    }
}

static bool cfg_for(int W, int* threads, int* vpt) {
    for (int v = 1; v <= 4; ++v) {
        if (W % (8 * v)) continue;
        const int t = W / (8 * v);
        if (t % 32 == 0 && t >= 32 && t <= 512) {
            *threads = t;
            *vpt = v;
            return true;
        }
    }
    return false;
}
static int launch_ok() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}
constexpr int kBwdCtas = 296;  // two per SM: rows are streamed, the dw partials stay in registers

}  // namespace mpt
}  // namespace llark

using namespace llark;
using namespace llark::mpt;

extern "C" int llark_ln_fwd(const void* x, long long ldx, const void* w, void* y, long long ldy, float* mean, float* rstd,
                            long long rows, int W, float eps, cudaStream_t st) {
    int t, v;
    if (rows <= 0 || !cfg_for(W, &t, &v) || (ldx % 8) || (ldy % 8) || ldx < W || ldy < W) return LLARK_EINVAL;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15)
        return LLARK_EINVAL;
    const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w;
    __nv_bfloat16* yp = (__nv_bfloat16*)y;
    switch (v) {
        case 1: ln_fwd_kernel<1><<<(unsigned)rows, t, 0, st>>>(xp, ldx, wp, yp, ldy, mean, rstd, W, eps); break;
        case 2: ln_fwd_kernel<2><<<(unsigned)rows, t, 0, st>>>(xp, ldx, wp, yp, ldy, mean, rstd, W, eps); break;
        case 3: ln_fwd_kernel<3><<<(unsigned)rows, t, 0, st>>>(xp, ldx, wp, yp, ldy, mean, rstd, W, eps); break;
        default: ln_fwd_kernel<4><<<(unsigned)rows, t, 0, st>>>(xp, ldx, wp, yp, ldy, mean, rstd, W, eps); break;
    }
    return launch_ok();
}

extern "C" long long llark_ln_bwd_workspace(long long rows, int W) {
    const long long nb = rows < kBwdCtas ? rows : kBwdCtas;
    return nb * W * (long long)sizeof(float);
}

extern "C" int llark_ln_bwd(const void* x, long long ldx, const void* w, const void* dy, long long lddy, const float* mean,
                            const float* rstd, const void* dres, void* dx, long long lddx, void* dw, int accumulate_dw,
                            long long rows, int W, void* workspace, cudaStream_t st) {
    int t, v;
    if (rows <= 0 || !cfg_for(W, &t, &v) || !mean || !rstd || (dw && !workspace)) return LLARK_EINVAL;
    if ((ldx % 8) || (lddy % 8) || (lddx % 8) || ldx < W || lddy < W || lddx < W) return LLARK_EINVAL;
    const int nb = (int)(rows < kBwdCtas ? rows : kBwdCtas);
    const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w, *dyp = (const __nv_bfloat16*)dy,
                        *drp = (const __nv_bfloat16*)dres;
    __nv_bfloat16* dxp = (__nv_bfloat16*)dx;
    float* part = dw ? (float*)workspace : nullptr;
    switch (v) {
        case 1: ln_bwd_kernel<1><<<nb, t, 0, st>>>(xp, ldx, wp, dyp, lddy, mean, rstd, drp, dxp, lddx, part, rows, W); break;
        case 2: ln_bwd_kernel<2><<<nb, t, 0, st>>>(xp, ldx, wp, dyp, lddy, mean, rstd, drp, dxp, lddx, part, rows, W); break;
        case 3: ln_bwd_kernel<3><<<nb, t, 0, st>>>(xp, ldx, wp, dyp, lddy, mean, rstd, drp, dxp, lddx, part, rows, W); break;
        default: ln_bwd_kernel<4><<<nb, t, 0, st>>>(xp, ldx, wp, dyp, lddy, mean, rstd, drp, dxp, lddx, part, rows, W); break;
    }
    int rc = launch_ok();
    if (rc) return rc;
    if (dw) {
        ln_dw_reduce_kernel<<<(W + 255) / 256, 256, 0, st>>>(part, nb, W, (__nv_bfloat16*)dw, accumulate_dw);
        rc = launch_ok();
    }
    return rc;
}

extern "C" int llark_gelu_fwd(const void* u, void* h, long long n, cudaStream_t st) {
    if (n <= 0 || (n % 8)) return LLARK_EINVAL;
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gelu_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)u, (__nv_bfloat16*)h, n / 8);
    return launch_ok();
}
extern "C" int llark_gelu_bwd(const void* u, const void* dh, void* du, long long n, cudaStream_t st) {
    if (n <= 0 || (n % 8)) return LLARK_EINVAL;
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gelu_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)u, (const __nv_bfloat16*)dh,
                                                      (__nv_bfloat16*)du, n / 8);
    return launch_ok();
}
