// Jukebox front end kernels (everything of the 5b VQ-VAE encoder + prior that is not a GEMM).
// Reference call sites: jukebox/main.py:54-68 (get_z -> vqvae.encode), :101-110 (prior.forward, only_encode),
// :113-130 (windowed_average). Arithmetic follows openai/jukebox@08efbbc as restated in SURVEY Appendix A
// (the upstream source is not vendored; oracle/jukebox_oracle.py is the CPU restatement these are tested against).
//
// All activations fp32 (the reference runs the prior with fp16=False). The conv stack is written time-tiled with
// all channels of a tile resident in shared memory, so each ResConv1DBlock is ONE pass over HBM
// (read x, write x'): the per-res-block fusion SURVEY §8(d) counts 1.42 GB/clip against.
#include "common.cuh"
#include <math.h>

namespace llark {

extern thread_local int g_last_cuda_error;
static inline int jb_check() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

constexpr int TT = 128;  // output time positions per CTA

// ---------------------------------------------------------------------------------------------- down conv k4 s2 p1
// y[n,co,t] = b[co] + sum_ci sum_k w[co,ci,k] * x[n,ci,2t-1+k].  256 threads = 128 t x 2 halves of c_out.
template <int COUT>
__global__ void __launch_bounds__(256)
conv_down_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, int c_in, int T) {
    extern __shared__ float sm[];
    const int To = T / 2;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int in_w = 2 * TT + 2;
    float* sx = sm;                        // [c_in][in_w]
    float* sw = sm + c_in * in_w;          // [COUT][c_in][4]
    for (int i = threadIdx.x; i < c_in * in_w; i += blockDim.x) {
        const int ci = i / in_w, j = i % in_w;
        const int ti = 2 * t0 - 1 + j;
        sx[i] = (ti >= 0 && ti < T) ? x[((long long)n * c_in + ci) * T + ti] : 0.f;
    }
    for (int i = threadIdx.x; i < COUT * c_in * 4; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int tl = threadIdx.x % TT, half = threadIdx.x / TT;
    constexpr int CPT = COUT / 2;
    float acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = bias[half * CPT + c];
    for (int ci = 0; ci < c_in; ++ci) {
        const float* xr = sx + ci * in_w + 2 * tl;
        const float x0 = xr[0], x1 = xr[1], x2 = xr[2], x3 = xr[3];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float* wr = sw + ((half * CPT + c) * c_in + ci) * 4;
            acc[c] = fmaf(wr[0], x0, acc[c]);
            acc[c] = fmaf(wr[1], x1, acc[c]);
            acc[c] = fmaf(wr[2], x2, acc[c]);
            acc[c] = fmaf(wr[3], x3, acc[c]);
        }
    }
    const int t = t0 + tl;
    if (t < To) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) y[((long long)n * COUT + half * CPT + c) * To + t] = acc[c];
    }
}

// ---------------------------------------------------------------------------------------------- fused res block
// y = x + W2 * relu(W1 (*dilation d, k=3) relu(x) + b1) + b2,   32 channels.
__global__ void __launch_bounds__(256)
resblock_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y, int T, int d) {
    extern __shared__ float sm[];
    constexpr int C = 32;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int in_w = TT + 2 * d;
    float* sx = sm;                  // relu(x) tile [C][in_w]
    float* sh = sx + C * in_w;       // relu(h) tile [C][TT]
    float* sw1 = sh + C * TT;        // [C][C][3]
    float* sw2 = sw1 + C * C * 3;    // [C][C]
    for (int i = threadIdx.x; i < C * in_w; i += blockDim.x) {
        const int ci = i / in_w, j = i % in_w;
        const int ti = t0 - d + j;
        const float v = (ti >= 0 && ti < T) ? x[((long long)n * C + ci) * T + ti] : 0.f;
        sx[i] = fmaxf(v, 0.f);
    }
    for (int i = threadIdx.x; i < C * C * 3; i += blockDim.x) sw1[i] = w1[i];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) sw2[i] = w2[i];
    __syncthreads();
    const int tl = threadIdx.x % TT, half = threadIdx.x / TT;
    constexpr int CPT = C / 2;
    float acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = b1[half * CPT + c];
    for (int ci = 0; ci < C; ++ci) {
        const float* xr = sx + ci * in_w + tl;
        const float x0 = xr[0], x1 = xr[d], x2 = xr[2 * d];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float* wr = sw1 + ((half * CPT + c) * C + ci) * 3;
            acc[c] = fmaf(wr[0], x0, acc[c]);
            acc[c] = fmaf(wr[1], x1, acc[c]);
            acc[c] = fmaf(wr[2], x2, acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) sh[(half * CPT + c) * TT + tl] = fmaxf(acc[c], 0.f);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = b2[half * CPT + c];
    for (int ci = 0; ci < C; ++ci) {
        const float hv = sh[ci * TT + tl];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = fmaf(sw2[(half * CPT + c) * C + ci], hv, acc[c]);
    }
    const int t = t0 + tl;
    if (t < T) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const long long off = ((long long)n * C + half * CPT + c) * T + t;
            y[off] = x[off] + acc[c];
        }
    }
}

// ---------------------------------------------------------------------------------------------- conv k3 s1 p1
// 256 threads = 64 t x 4 groups of c_out/4 channels.
template <int COUT>
__global__ void __launch_bounds__(256)
conv_k3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, int c_in, int T) {
    extern __shared__ float sm[];
    constexpr int TK = 64;
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * TK;
    const int in_w = TK + 2;
    float* sx = sm;
    float* sw = sm + c_in * in_w;  // [COUT][c_in][3]
    for (int i = threadIdx.x; i < c_in * in_w; i += blockDim.x) {
        const int ci = i / in_w, j = i % in_w;
        const int ti = t0 - 1 + j;
        sx[i] = (ti >= 0 && ti < T) ? x[((long long)n * c_in + ci) * T + ti] : 0.f;
    }
    for (int i = threadIdx.x; i < COUT * c_in * 3; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int tl = threadIdx.x % TK, grp = threadIdx.x / TK;
    constexpr int CPT = COUT / 4;
    float acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = bias[grp * CPT + c];
    for (int ci = 0; ci < c_in; ++ci) {
        const float* xr = sx + ci * in_w + tl;
        const float x0 = xr[0], x1 = xr[1], x2 = xr[2];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const float* wr = sw + ((grp * CPT + c) * c_in + ci) * 3;
            acc[c] = fmaf(wr[0], x0, acc[c]);
            acc[c] = fmaf(wr[1], x1, acc[c]);
            acc[c] = fmaf(wr[2], x2, acc[c]);
        }
    }
    const int t = t0 + tl;
    if (t < T) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) y[((long long)n * COUT + grp * CPT + c) * T + t] = acc[c];
    }
}

// ---------------------------------------------------------------------------------------------- VQ nearest code
// One thread per (n,t); x vector in registers, codebook streamed through shared memory (broadcast reads).
// dist = (|x|^2 - 2 x.k) + |k|^2, first minimum wins.
__global__ void __launch_bounds__(128)
vq_encode_kernel(const float* __restrict__ x, const float* __restrict__ cb, long long* __restrict__ codes, int T,
                 int n_codes) {
    constexpr int E = 64, CHUNK = 128;
    __shared__ float scb[CHUNK * E];
    __shared__ float skk[CHUNK];
    const int n = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float xv[E];
    float xx = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        xv[e] = (t < T) ? x[((long long)n * E + e) * T + t] : 0.f;
        xx = fmaf(xv[e], xv[e], xx);
    }
    float best = INFINITY;
    int best_j = 0;
    for (int c0 = 0; c0 < n_codes; c0 += CHUNK) {
        __syncthreads();
        for (int i = threadIdx.x; i < CHUNK * E; i += blockDim.x) {
            const int j = c0 + i / E;
            scb[i] = j < n_codes ? cb[(long long)c0 * E + i] : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < CHUNK) {
            float kk = 0.f;
            for (int e = 0; e < E; ++e) kk = fmaf(scb[threadIdx.x * E + e], scb[threadIdx.x * E + e], kk);
            skk[threadIdx.x] = kk;
        }
        __syncthreads();
        const int lim = min(CHUNK, n_codes - c0);
        for (int j = 0; j < lim; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) dot = fmaf(xv[e], scb[j * E + e], dot);
            const float dist = (xx - 2.f * dot) + skk[j];
            if (dist < best) {
                best = dist;
                best_j = c0 + j;
            }
        }
    }
    if (t < T) codes[(long long)n * T + t] = best_j;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ y, int y_bf16, int W, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const float* xr = x + row * W;
    float s = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        s += v.x + v.y + v.z + v.w;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float tot = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    tot = warp_sum(tot);
    __shared__ float s_mean, s_rstd;
    if (threadIdx.x == 0) s_mean = tot / (float)W;
    __syncthreads();
    const float mean = s_mean;
    float q = 0.f;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += a * a + b * b + c * c + d * d;
    }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    float qt = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    qt = warp_sum(qt);
    if (threadIdx.x == 0) s_rstd = rsqrtf(qt / (float)W + eps);
    __syncthreads();
    const float rstd = s_rstd;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + i);
        const float4 g = *reinterpret_cast<const float4*>(gamma + i);
        const float4 b = *reinterpret_cast<const float4*>(beta + i);
        const float o0 = (v.x - mean) * rstd * g.x + b.x, o1 = (v.y - mean) * rstd * g.y + b.y,
                    o2 = (v.z - mean) * rstd * g.z + b.z, o3 = (v.w - mean) * rstd * g.w + b.w;
        if (y_bf16) {
            uint2 o;
            o.x = pack_bf16x2(o0, o1);
            o.y = pack_bf16x2(o2, o3);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * W + i) = o;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * W + i) = make_float4(o0, o1, o2, o3);
        }
    }
}

// ---------------------------------------------------------------------------------------------- prior input embedding
__global__ void jb_embed_kernel(const long long* __restrict__ codes, const float* __restrict__ emb,
                                const float* __restrict__ y_cond, const float* __restrict__ pos_emb,
                                const float* __restrict__ x_cond, float* __restrict__ out, int T, int W) {
    const int t = blockIdx.x, n = blockIdx.y;
    const float* src = (t == 0) ? y_cond : emb + codes[(long long)n * T + t - 1] * W;
    const float* pe = pos_emb + (long long)t * W;
    const float* xc = x_cond + (long long)t * W;
    float* dst = out + ((long long)n * T + t) * W;
    for (int i = threadIdx.x * 4; i < W; i += blockDim.x * 4) {
        const float4 a = *reinterpret_cast<const float4*>(src + i);
        const float4 b = *reinterpret_cast<const float4*>(pe + i);
        const float4 c = *reinterpret_cast<const float4*>(xc + i);
        // reference order: (x_emb + pos_emb) + x_cond
        *reinterpret_cast<float4*>(dst + i) =
            make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    }
}

// ---------------------------------------------------------------------------------------------- factored attention
// One CTA per (group, head, n). A group is <= 128 queries and <= 128 keys gathered with a stride:
//   pattern 0 (block):            group g = block g; queries = keys = positions g*blk + i;           causal
//   pattern 1 (transposed block): group g = offset g; queries = keys = positions i*blk + g;          causal in i
//   pattern 2 (previous block):   group g = block g; queries g*blk + i, keys (g-1)*blk + i (none for g=0); full
// Each warp processes one query at a time: lanes over keys for QK^T, lanes over channels for PV (fp32 SIMT).
__global__ void __launch_bounds__(256)
jb_sparse_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads, int dh, int blk,
                      int pattern, float scale) {
    extern __shared__ float sm[];
    const int g = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    const int nblk = T / blk;
    int nq, nk, qstride, q0, k0;
    bool causal;
    if (pattern == 0) { nq = nk = blk; qstride = 1; q0 = k0 = g * blk; causal = true; }
    else if (pattern == 1) { nq = nk = nblk; qstride = blk; q0 = k0 = g; causal = true; }
    else { nq = blk; nk = (g == 0) ? 0 : blk; qstride = 1; q0 = g * blk; k0 = (g - 1) * blk; causal = false; }
    const int ldk = dh + 1;
    float* sK = sm;               // [nk][ldk]
    float* sV = sK + 128 * ldk;   // [nk][ldk]
    float* sQ = sV + 128 * ldk;   // [8 warps][dh]
    float* sP = sQ + 8 * dh;      // [8 warps][128]
    const long long rs = 3LL * heads * dh;
    const float* base = qkv + (long long)n * T * rs;
    for (int i = threadIdx.x; i < nk * dh; i += blockDim.x) {
        const int r = i / dh, c = i % dh;
        const long long pos = k0 + (long long)r * qstride;
        sK[r * ldk + c] = base[pos * rs + (long long)heads * dh + h * dh + c];
        sV[r * ldk + c] = base[pos * rs + 2LL * heads * dh + h * dh + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* myq = sQ + warp * dh;
    float* myp = sP + warp * 128;
    for (int qi = warp; qi < nq; qi += 8) {
        const long long qpos = q0 + (long long)qi * qstride;
        for (int c = lane; c < dh; c += 32) myq[c] = base[qpos * rs + h * dh + c];
        __syncwarp();
        const int kmax = causal ? qi + 1 : nk;  // keys [0, kmax)
        float sc[4];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kj = lane + r * 32;
            float dot = -INFINITY;
            if (kj < kmax) {
                dot = 0.f;
                const float* kr = sK + kj * ldk;
                for (int c = 0; c < dh; ++c) dot = fmaf(myq[c], kr[c], dot);
                dot *= scale;
            }
            sc[r] = dot;
            mx = fmaxf(mx, dot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float p = (sc[r] == -INFINITY) ? 0.f : __expf(sc[r] - mx);
            sc[r] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) myp[lane + r * 32] = sc[r] * inv;
        __syncwarp();
        float* orow = out + ((long long)n * T + qpos) * ((long long)heads * dh) + h * dh;
        for (int c = lane; c < dh; c += 32) {
            float acc = 0.f;
            for (int kj = 0; kj < kmax; ++kj) acc = fmaf(myp[kj], sV[kj * ldk + c], acc);
            orow[c] = acc;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- elementwise
__global__ void quick_gelu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (; i + 4 <= n; i += stride) {
        float4 v = *reinterpret_cast<const float4*>(x + i);
        v.x = v.x / (1.f + __expf(-1.702f * v.x));
        v.y = v.y / (1.f + __expf(-1.702f * v.y));
        v.z = v.z / (1.f + __expf(-1.702f * v.z));
        v.w = v.w / (1.f + __expf(-1.702f * v.w));
        *reinterpret_cast<float4*>(y + i) = v;
    }
    if (i < n)
        for (long long j = i; j < n; ++j) y[j] = x[j] / (1.f + __expf(-1.702f * x[j]));
}
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = c ? (a[i] + b[i]) + c[i] : a[i] + b[i];
}
// out[f, c] = mean_{j<frame} acts[f*frame + j, c]
__global__ void jb_pool_kernel(const float* __restrict__ acts, float* __restrict__ out, int W, int frame) {
    const int f = blockIdx.x;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < frame; ++j) s += acts[((long long)f * frame + j) * W + c];
        out[(long long)f * W + c] = s / (float)frame;
    }
}
// colsum over rows of a bf16 [rows, H] matrix -> bf16 [H] (bias gradient), fp32 accumulate.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows,
                                   int H, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    float s = 0.f;
    for (long long r = 0; r < rows; ++r) s += __bfloat162float(x[r * H + c]);
    if (accumulate) s += __bfloat162float(out[c]);
    out[c] = __float2bfloat16_rn(s);
}

static int set_dyn_smem(const void* fn, int bytes) {
    if (bytes <= 48 * 1024) return LLARK_OK;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

}  // namespace llark

using namespace llark;

extern "C" int llark_jb_conv_down(const float* x, const float* w, const float* bias, float* y, int N, int c_in,
                                  int c_out, int T, cudaStream_t st) {
    if (N <= 0 || T <= 0 || (T & 1) || c_in <= 0 || (c_out != 32 && c_out != 64)) return LLARK_EINVAL;
    const int smem = (c_in * (2 * TT + 2) + c_out * c_in * 4) * (int)sizeof(float);
    dim3 grid((T / 2 + TT - 1) / TT, N);
    int rc;
    if (c_out == 32) {
        rc = set_dyn_smem((const void*)conv_down_kernel<32>, smem);
        if (rc) return rc;
        conv_down_kernel<32><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    } else {
        rc = set_dyn_smem((const void*)conv_down_kernel<64>, smem);
        if (rc) return rc;
        conv_down_kernel<64><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    }
    return jb_check();
}

extern "C" int llark_jb_resblock(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                 float* y, int N, int T, int dilation, cudaStream_t st) {
    if (N <= 0 || T <= 0 || dilation < 1 || dilation > 243) return LLARK_EINVAL;
    const int smem = (32 * (TT + 2 * dilation) + 32 * TT + 32 * 32 * 3 + 32 * 32) * (int)sizeof(float);
    int rc = set_dyn_smem((const void*)resblock_kernel, smem);
    if (rc) return rc;
    dim3 grid((T + TT - 1) / TT, N);
    resblock_kernel<<<grid, 256, smem, st>>>(x, w1, b1, w2, b2, y, T, dilation);
    return jb_check();
}

extern "C" int llark_jb_conv_k3(const float* x, const float* w, const float* bias, float* y, int N, int c_in, int c_out,
                                int T, cudaStream_t st) {
    if (N <= 0 || T <= 0 || c_in <= 0 || (c_out != 32 && c_out != 64)) return LLARK_EINVAL;
    const int smem = (c_in * 66 + c_out * c_in * 3) * (int)sizeof(float);
    dim3 grid((T + 63) / 64, N);
    int rc;
    if (c_out == 64) {
        rc = set_dyn_smem((const void*)conv_k3_kernel<64>, smem);
        if (rc) return rc;
        conv_k3_kernel<64><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    } else {
        rc = set_dyn_smem((const void*)conv_k3_kernel<32>, smem);
        if (rc) return rc;
        conv_k3_kernel<32><<<grid, 256, smem, st>>>(x, w, bias, y, c_in, T);
    }
    return jb_check();
}

extern "C" int llark_jb_vq_encode(const float* x, const float* codebook, long long* codes, int N, int T, int n_codes,
                                  cudaStream_t st) {
    if (N <= 0 || T <= 0 || n_codes <= 0) return LLARK_EINVAL;
    dim3 grid((T + 127) / 128, N);
    vq_encode_kernel<<<grid, 128, 0, st>>>(x, codebook, codes, T, n_codes);
    return jb_check();
}

extern "C" int llark_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16,
                                   long long rows, int W, float eps, cudaStream_t st) {
    if (rows <= 0 || W <= 0 || (W & 3)) return LLARK_EINVAL;
    layernorm_kernel<<<(unsigned)rows, 256, 0, st>>>(x, gamma, beta, y, y_bf16, W, eps);
    return jb_check();
}

extern "C" int llark_jb_embed_tokens(const long long* codes, const float* emb, const float* y_cond,
                                     const float* pos_emb, const float* x_cond, float* out, int N, int T, int W,
                                     cudaStream_t st) {
    if (N <= 0 || T <= 0 || (W & 3)) return LLARK_EINVAL;
    dim3 grid(T, N);
    jb_embed_kernel<<<grid, 256, 0, st>>>(codes, emb, y_cond, pos_emb, x_cond, out, T, W);
    return jb_check();
}

extern "C" int llark_jb_sparse_attn(const float* qkv, float* out, int N, int T, int heads, int dh, int block,
                                    int pattern, cudaStream_t st) {
    if (N <= 0 || T <= 0 || block <= 0 || T % block || pattern < 0 || pattern > 2) return LLARK_EINVAL;
    const int nblk = T / block;
    if (block > 128 || nblk > 128 || dh > 256) return LLARK_EINVAL;
    const int groups = (pattern == 1) ? block : nblk;
    const int smem = (2 * 128 * (dh + 1) + 8 * dh + 8 * 128) * (int)sizeof(float);
    int rc = set_dyn_smem((const void*)jb_sparse_attn_kernel, smem);
    if (rc) return rc;
    dim3 grid(groups, heads, N);
    jb_sparse_attn_kernel<<<grid, 256, smem, st>>>(qkv, out, T, heads, dh, block, pattern, 1.f / sqrtf((float)dh));
    return jb_check();
}

static inline int ew_grid(long long n, int per_thread) {
    long long b = (n / per_thread + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (int)b;
}
extern "C" int llark_quick_gelu(const float* x, float* y, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    quick_gelu_kernel<<<ew_grid(n, 4), 256, 0, st>>>(x, y, n);
    return jb_check();
}
extern "C" int llark_add3(const float* a, const float* b, const float* c, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    add3_kernel<<<ew_grid(n, 1), 256, 0, st>>>(a, b, c, out, n);
    return jb_check();
}
extern "C" int llark_jb_pool(const float* acts, float* out, int T_valid, int W, int frame, cudaStream_t st) {
    if (frame <= 0 || W <= 0) return LLARK_EINVAL;
    const int frames = T_valid / frame;
    if (frames <= 0) return LLARK_OK;
    jb_pool_kernel<<<frames, 256, 0, st>>>(acts, out, W, frame);
    return jb_check();
}
extern "C" int llark_colsum(const void* x, void* out, long long rows, int H, int accumulate, cudaStream_t st) {
    if (rows <= 0 || H <= 0) return LLARK_EINVAL;
    colsum_bf16_kernel<<<(H + 127) / 128, 128, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, H, accumulate);
    return jb_check();
}
