// HBM-bound kernels of the LLark/Llama training step (everything that is not a GEMM or attention):
// embedding gather + audio splice, RMSNorm, RoPE, SwiGLU, shifted cross-entropy, gradient norm, AdamW.
// All are coalesced 16-byte-vector streaming kernels; none touches tensor cores (HBM roof).
//
// Reference call sites (file:line under /root/reference unless noted):
//   embed gather + splice     m2t/models/llamav2.py:124,141-222
//   RMSNorm / RoPE / SwiGLU   transformers LlamaRMSNorm / apply_rotary_pos_emb / LlamaMLP (reached at llamav2.py:224)
//   shifted CE                m2t/models/llamav2.py:314-325
//   AdamW + clip              HF Trainer (optim="adamw_torch", m2t/arguments.py:84; max_grad_norm 1.0)
#include "common.cuh"
#include "../../include/llark_b200.h"
#include <math.h>

namespace llark {

extern thread_local int g_last_cuda_error;

static inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}

struct alignas(16) bf16x8 {
    uint32_t u[4];
};
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = o;
}

// block-wide sum; blockDim.x multiple of 32, <= 1024. Result broadcast to all threads.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : -INFINITY;
    t = warp_max(t);
    return t;
}

// ------------------------------------------------------------------------------------------------ cast
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (; i + 8 <= n; i += stride) {
        const float4 a = *reinterpret_cast<const float4*>(in + i);
        const float4 b = *reinterpret_cast<const float4*>(in + i + 4);
        float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store8(out + i, f);
    }
    if (i < n)
        for (long long j = i; j < n && j < i + 8; ++j) out[j] = __float2bfloat16_rn(in[j]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (; i + 8 <= n; i += stride) {
        float f[8];
        load8(in + i, f);
        *reinterpret_cast<float4*>(out + i) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(out + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (i < n)
        for (long long j = i; j < n && j < i + 8; ++j) out[j] = __bfloat162float(in[j]);
}

// ------------------------------------------------------------------------------------------------ splice
// One warp per sample: count <audio_start>/<audio_end>, record the (single) start position, validate the
// reference's two ValueError conditions (llamav2.py:146-152, :169-175) into a flag word instead of a host sync.
// flags bits: 1 = start/end count mismatch, 2 = end token not at start+F+1, 4 = more than one audio segment.
__global__ void splice_scan_kernel(const long long* __restrict__ ids, int B, int S, long long start_id,
                                   long long end_id, const int* __restrict__ frames, int* __restrict__ start_pos,
                                   int* __restrict__ flags) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const long long* row = ids + (long long)b * S;
    int n_start = 0, n_end = 0, first = S;
    for (int s = lane; s < S; s += 32) {
        const long long t = row[s];
        if (t == start_id) {
            ++n_start;
            first = min(first, s);
        }
        if (t == end_id) ++n_end;
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_start += __shfl_xor_sync(0xffffffffu, n_start, o);
        n_end += __shfl_xor_sync(0xffffffffu, n_end, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    }
    if (lane == 0) {
        int f = 0;
        if (n_start != n_end) f |= 1;
        if (n_start > 1) f |= 4;
        int pos = -1;
        if (n_start >= 1) {
            pos = first;
            const int F = frames[b];
            const int e = pos + F + 1;
            if (e >= S || row[e] != end_id) f |= 2;
        }
        start_pos[b] = pos;
        if (f) atomicOr(flags, f);
    }
}

// out[b,s,:] = audio[b, s-pos-1, :] if pos < s <= pos+F else table[ids[b,s], :]. One CTA per token row.
__global__ void embed_splice_fwd_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                        const __nv_bfloat16* __restrict__ audio, const long long* __restrict__ audio_off,
                                        const int* __restrict__ frames, const int* __restrict__ start_pos,
                                        __nv_bfloat16* __restrict__ out, int S, int H, long long vocab) {
    const int b = blockIdx.y, s = blockIdx.x;
    const int pos = start_pos ? start_pos[b] : -1;
    const int F = frames ? frames[b] : 0;
    const __nv_bfloat16* src;
    if (pos >= 0 && s > pos && s <= pos + F) {
        src = audio + (audio_off[b] + (long long)(s - pos - 1)) * H;
    } else {
        long long t = ids[(long long)b * S + s];
        if (t < 0 || t >= vocab) t = 0;
        src = table + t * H;
    }
    __nv_bfloat16* dst = out + ((long long)b * S + s) * H;
    for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8)
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

// Backward of the splice. d_audio rows <- d_out audio block; d_table rows: with detach_text (the reference's
// orig_embeds_params path, llamav2.py:176-198) only the two marker rows receive gradient, otherwise every text row.
__global__ void embed_splice_bwd_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ dout,
                                        __nv_bfloat16* __restrict__ daudio, const long long* __restrict__ audio_off,
                                        const int* __restrict__ frames, const int* __restrict__ start_pos,
                                        __nv_bfloat16* __restrict__ dtable, int S, int H, long long vocab,
                                        int detach_text) {
    const int b = blockIdx.y, s = blockIdx.x;
    const int pos = start_pos ? start_pos[b] : -1;
    const int F = frames ? frames[b] : 0;
    const __nv_bfloat16* src = dout + ((long long)b * S + s) * H;
    if (pos >= 0 && s > pos && s <= pos + F) {
        if (!daudio) return;
        __nv_bfloat16* dst = daudio + (audio_off[b] + (long long)(s - pos - 1)) * H;
        for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8)
            *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
        return;
    }
    if (!dtable) return;
    const bool marker = pos >= 0 && (s == pos || s == pos + F + 1);
    // a sample without <audio_start> keeps its undetached embeddings in the reference (the loop body that detaches,
    // llamav2.py:162-208, never runs for it), so every one of its text rows receives gradient
    if (detach_text && !marker && pos >= 0) return;
    long long t = ids[(long long)b * S + s];
    if (t < 0 || t >= vocab) return;
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dtable + t * H);
    const __nv_bfloat162* s2 = reinterpret_cast<const __nv_bfloat162*>(src);
    for (int i = threadIdx.x; i < H / 2; i += blockDim.x) atomicAdd(dst + i, s2[i]);
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// HF LlamaRMSNorm: var = mean(x_f32^2); y = w * bf16(x_f32 * rsqrt(var + eps))   (bf16 * bf16 -> bf16)
template <int VPT>  // 8-element vectors per thread; H == blockDim.x * 8 * VPT
__global__ void rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                   __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out, int H, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const __nv_bfloat16* xr = x + row * H;
    float v[VPT][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        load8(xr + (i * blockDim.x + threadIdx.x) * 8, v[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[i][e] * v[i][e];
    }
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / (float)H + eps);
    if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int c = (i * blockDim.x + threadIdx.x) * 8;
        float wv[8], o[8];
        load8(w + c, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = wv[e] * bf16_round(v[i][e] * rstd);
        store8(y + row * H + c, o);
    }
}

// dx = [dres +] rstd * (g - xhat * mean(g * xhat)), g = dy * w, xhat = x * rstd;  dw_partial[blk] += dy * bf16(xhat)
// Each CTA walks rows blockIdx.x, +gridDim.x, ... and keeps its dw partial in registers (deterministic 2-stage).
template <int VPT>
__global__ void rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                   const __nv_bfloat16* __restrict__ dy, const float* __restrict__ rstd_in,
                                   const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx,
                                   float* __restrict__ dw_partial, long long rows, int H) {
    __shared__ float red[32];
    float dwacc[VPT][8];
    float wv[VPT][8];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        load8(w + (i * blockDim.x + threadIdx.x) * 8, wv[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) dwacc[i][e] = 0.f;
    }
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float rstd = rstd_in[row];
        float xh[VPT][8], g[VPT][8];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = (i * blockDim.x + threadIdx.x) * 8;
            float xv[8], dyv[8];
            load8(x + row * H + c, xv);
            load8(dy + row * H + c, dyv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                xh[i][e] = xv[e] * rstd;
                g[i][e] = dyv[e] * wv[i][e];
                dot += g[i][e] * xh[i][e];
                dwacc[i][e] += dyv[e] * bf16_round(xh[i][e]);
            }
        }
        dot = block_sum(dot, red) / (float)H;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = (i * blockDim.x + threadIdx.x) * 8;
            float o[8];
            if (dres) load8(dres + row * H + c, o);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += rstd * (g[i][e] - xh[i][e] * dot);
            store8(dx + row * H + c, o);
        }
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int c = (i * blockDim.x + threadIdx.x) * 8;
        float* dst = dw_partial + (long long)blockIdx.x * H + c;
        *reinterpret_cast<float4*>(dst) = make_float4(dwacc[i][0], dwacc[i][1], dwacc[i][2], dwacc[i][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(dwacc[i][4], dwacc[i][5], dwacc[i][6], dwacc[i][7]);
    }
}

// dw[c] (+)= sum_p partial[p][c]  -> bf16
__global__ void colsum_partials_kernel(const float* __restrict__ partial, int nparts, int H,
                                       __nv_bfloat16* __restrict__ dw, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(long long)p * H + c];
    if (accumulate) s += __bfloat162float(dw[c]);
    dw[c] = __float2bfloat16_rn(s);
}

// ------------------------------------------------------------------------------------------------ RoPE
// In place on the packed qkv buffer [rows, 3*H] (q heads | k heads | v heads), head_dim D, positions s = row % S.
// HF apply_rotary_pos_emb in bf16: out = bf16(bf16(x*cos) + bf16(rotate_half(x)*sin)), cos/sin pre-rounded to bf16.
// inverse != 0 applies the transpose (backward): dx = bf16(bf16(dy*cos) + bf16(rotate_half^T(dy*sin))).
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ cos_t,
                            const __nv_bfloat16* __restrict__ sin_t, long long rows, int S, int n_heads, int D,
                            long long row_stride, int inverse) {
    // thread handles 8 consecutive i in [0, D/2) of one (row, q-or-k head)
    const int half = D / 2;
    const int vec_per_head = half / 8;
    const long long total = rows * 2 * n_heads * vec_per_head;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int vi = (int)(idx % vec_per_head);
        long long t = idx / vec_per_head;
        const int head = (int)(t % (2 * n_heads));  // q heads then k heads are contiguous in the row
        const long long row = t / (2 * n_heads);
        const int s = (int)(row % S);
        __nv_bfloat16* base = qkv + row * row_stride + (long long)head * D + vi * 8;
        float x1[8], x2[8], c1[8], s1[8], c2[8], s2[8];
        load8(base, x1);
        load8(base + half, x2);
        load8(cos_t + (long long)s * D + vi * 8, c1);
        load8(sin_t + (long long)s * D + vi * 8, s1);
        load8(cos_t + (long long)s * D + half + vi * 8, c2);
        load8(sin_t + (long long)s * D + half + vi * 8, s2);
        float o1[8], o2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (!inverse) {
                // rotate_half(x) = cat(-x2, x1)
                o1[e] = bf16_round(x1[e] * c1[e]) + bf16_round(-x2[e] * s1[e]);
                o2[e] = bf16_round(x2[e] * c2[e]) + bf16_round(x1[e] * s2[e]);
            } else {
                // transpose: dx1 = dy1*c1 + dy2*s2 ; dx2 = dy2*c2 - dy1*s1
                o1[e] = bf16_round(x1[e] * c1[e]) + bf16_round(x2[e] * s2[e]);
                o2[e] = bf16_round(x2[e] * c2[e]) + bf16_round(-x1[e] * s1[e]);
            }
        }
        store8(base, o1);
        store8(base + half, o2);
    }
}

// ------------------------------------------------------------------------------------------------ SwiGLU
// gu [rows, 2I] = (gate | up);  act = bf16(silu(gate)) * up
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ act,
                                  long long rows, int I) {
    const int vec_per_row = I / 8;
    const long long total = rows * vec_per_row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / vec_per_row;
        const int c = (int)(idx % vec_per_row) * 8;
        float g[8], u[8], o[8];
        load8(gu + row * 2 * I + c, g);
        load8(gu + row * 2 * I + I + c, u);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = bf16_round(__fdividef(g[e], 1.f + __expf(-g[e]))) * u[e];
        store8(act + row * I + c, o);
    }
}
// dgu = (dact * up * silu'(gate) | dact * silu(gate))
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dact,
                                  __nv_bfloat16* __restrict__ dgu, long long rows, int I) {
    const int vec_per_row = I / 8;
    const long long total = rows * vec_per_row;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / vec_per_row;
        const int c = (int)(idx % vec_per_row) * 8;
        float g[8], u[8], d[8], dg[8], du[8];
        load8(gu + row * 2 * I + c, g);
        load8(gu + row * 2 * I + I + c, u);
        load8(dact + row * I + c, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float sg = __fdividef(1.f, 1.f + __expf(-g[e]));
            const float silu = g[e] * sg;
            du[e] = d[e] * bf16_round(silu);
            dg[e] = d[e] * u[e] * (sg * (1.f + g[e] * (1.f - sg)));
        }
        store8(dgu + row * 2 * I + c, dg);
        store8(dgu + row * 2 * I + I + c, du);
    }
}

// ------------------------------------------------------------------------------------------------ shifted CE
// One CTA per (b, s) row of bf16 logits [B*S, ldv]. target = labels[b, s+1] (ignore_index -100; the last
// position of every sample has no target). Writes per-row loss and lse; a second kernel reduces to the mean.
__global__ void ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ldv, int V,
                              const long long* __restrict__ labels, int S, float* __restrict__ row_loss,
                              float* __restrict__ row_lse) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const int s = (int)(row % S);
    long long target = -100;
    if (s < S - 1) target = labels[row + 1];
    if (target < 0 || target >= V) {
        if (threadIdx.x == 0) {
            row_loss[row] = 0.f;
            row_lse[row] = INFINITY;  // marks "no target"
        }
        return;
    }
    const __nv_bfloat16* lr = logits + row * ldv;
    float mx = -INFINITY;
    const int V8 = V & ~7;
    for (int c = threadIdx.x * 8; c < V8; c += blockDim.x * 8) {
        float f[8];
        load8(lr + c, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, f[e]);
    }
    for (int c = V8 + threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, __bfloat162float(lr[c]));
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int c = threadIdx.x * 8; c < V8; c += blockDim.x * 8) {
        float f[8];
        load8(lr + c, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += __expf(f[e] - mx);
    }
    for (int c = V8 + threadIdx.x; c < V; c += blockDim.x) sum += __expf(__bfloat162float(lr[c]) - mx);
    sum = block_sum(sum, red);
    if (threadIdx.x == 0) {
        const float lse = mx + logf(sum);
        row_lse[row] = lse;
        row_loss[row] = lse - __bfloat162float(lr[target]);
    }
}
// loss_out[0] = sum(row_loss over valid) / n_valid ; loss_out[1] = n_valid   (single CTA, deterministic)
__global__ void ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_lse,
                                 long long rows, float* __restrict__ loss_out) {
    __shared__ float red[32];
    float s = 0.f, n = 0.f;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
        if (row_lse[r] != INFINITY) {
            s += row_loss[r];
            n += 1.f;
        }
    }
    s = block_sum(s, red);
    n = block_sum(n, red);
    if (threadIdx.x == 0) {
        loss_out[0] = s / n;  // n == 0 -> NaN, like torch CrossEntropyLoss(mean) with no valid target
        loss_out[1] = n;
    }
}
// dlogits[row, v] = gscale/n_valid * (softmax - onehot) for rows with a target, else 0; pad columns [V, ldv) = 0.
__global__ void ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ldv, int V,
                              const long long* __restrict__ labels, int S, const float* __restrict__ row_lse,
                              const float* __restrict__ loss_out, const float* __restrict__ gscale_ptr, float gscale,
                              __nv_bfloat16* __restrict__ dlogits) {
    const long long row = blockIdx.x;
    const int s = (int)(row % S);
    const float lse = row_lse[row];
    __nv_bfloat16* dr = dlogits + row * ldv;
    const int ld = (int)ldv;
    if (lse == INFINITY) {
        for (int c = threadIdx.x * 8; c < ld; c += blockDim.x * 8) *reinterpret_cast<uint4*>(dr + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    const long long target = labels[row + 1];
    (void)s;
    const float scale = (gscale_ptr ? gscale_ptr[0] : gscale) / loss_out[1];
    const __nv_bfloat16* lr = logits + row * ldv;
    for (int c = threadIdx.x * 8; c < ld; c += blockDim.x * 8) {
        float f[8], o[8];
        load8(lr + c, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int col = c + e;
            float p = (col < V) ? __expf(f[e] - lse) : 0.f;
            if (col == target) p -= 1.f;
            o[e] = p * scale;
        }
        store8(dr + c, o);
    }
}

// ------------------------------------------------------------------------------------------------ grad norm / AdamW
__global__ void sumsq_partial_kernel(const __nv_bfloat16* __restrict__ g, long long n, float* __restrict__ partial) {
    __shared__ float red[32];
    float s = 0.f;
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (; i + 8 <= n; i += stride) {
        float f[8];
        load8(g + i, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += f[e] * f[e];
    }
    if (i < n)
        for (long long j = i; j < n && j < i + 8; ++j) {
            const float f = __bfloat162float(g[j]);
            s += f * f;
        }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// out[0] = (accumulate ? out[0] : 0) + sum(partial)
__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out, int accumulate) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s;
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, max_norm;
};
// torch.optim.AdamW (foreach) on bf16 tensors rounds to bf16 after every elementwise op; reproduce that order:
//   g  = bf16(g * grad_scale * clip_coef)                (1/(world*accum) scale and clip_grad_norm_)
//   p  = bf16(p * (1 - lr*wd)); m = bf16(lerp(m, g, 1-b1)); v = bf16(v*b2); v = bf16(v + (1-b2) g g)
//   den = bf16(bf16(bf16(sqrt(v)) / bc2_sqrt) + eps);  p = bf16(p - (lr/bc1) * m / den)
// The update costs ~100 instructions per element when written naively (IEEE sqrt and two IEEE divisions plus a
// cvt pair per bf16 rounding): at 6.6 G elements that is ISSUE-bound (measured 31 ms), not HBM-bound. Every
// intermediate is rounded to bf16 right after it is produced, so MUFU-based reciprocal / rsqrt (2 ulp in fp32) and
// packed cvt.rn.bf16x2 conversions change nothing after rounding (tests/test_kernels_gpu.py compares with torch).
__device__ __forceinline__ float2 round2(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);  // one cvt.rn.bf16x2.f32
    return __bfloat1622float2(t);
}
__device__ __forceinline__ void adamw_update2(float& p0, float& p1, float g0, float g1, float& m0, float& m1,
                                              float& v0, float& v1, const AdamArgs& a, float gmul, float w1,
                                              float one_m_b2, float inv_bc2, float step_size, float decay) {
    float2 g = make_float2(g0, g1);
    if (gmul != 1.f) g = round2(g.x * gmul, g.y * gmul);
    if (decay != 1.f) {
        const float2 pd = round2(p0 * decay, p1 * decay);
        p0 = pd.x; p1 = pd.y;
    }
    const float2 m = round2(fmaf(w1, g.x - m0, m0), fmaf(w1, g.y - m1, m1));  // lerp(m, g, 1 - beta1), weight < 0.5
    float2 v = round2(v0 * a.beta2, v1 * a.beta2);
    v = round2(fmaf(one_m_b2 * g.x, g.x, v.x), fmaf(one_m_b2 * g.y, g.y, v.y));
    float2 d = round2(v.x * rsqrtf(fmaxf(v.x, 1e-38f)), v.y * rsqrtf(fmaxf(v.y, 1e-38f)));  // sqrt(v)
    d = round2(d.x * inv_bc2, d.y * inv_bc2);
    d = round2(d.x + a.eps, d.y + a.eps);
    const float2 pn = round2(fmaf(-step_size, __fdividef(m.x, d.x), p0), fmaf(-step_size, __fdividef(m.y, d.y), p1));
    p0 = pn.x; p1 = pn.y; m0 = m.x; m1 = m.y; v0 = v.x; v1 = v.y;
}

__global__ void __launch_bounds__(256)
adamw_kernel(__nv_bfloat16* __restrict__ p, __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ m,
             __nv_bfloat16* __restrict__ v, long long n, const float* __restrict__ sumsq, AdamArgs a, int zero_grad) {
    float clip = 1.f;
    if (sumsq && a.max_norm > 0.f) {
        const float total = sqrtf(sumsq[0]) * a.grad_scale;
        clip = fminf(1.f, a.max_norm / (total + 1e-6f));
    }
    // torch scales the gradient twice (1/world by DDP, then clip_grad_norm_), rounding in between; when only one
    // of the two factors is active this is one multiply, otherwise the product differs from two roundings by at
    // most one bf16 ulp of the gradient.
    const float gmul = a.grad_scale * clip;
    const float w1 = 1.f - a.beta1, one_m_b2 = 1.f - a.beta2;
    const float inv_bc2 = 1.f / a.bc2_sqrt, step_size = a.lr / a.bc1;
    const float decay = 1.f - a.lr * a.weight_decay;
    const long long nvec = n / 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec; vi += stride) {
        const long long i = vi * 8;
        const uint4 pu = *reinterpret_cast<const uint4*>(p + i);
        const uint4 gu = *reinterpret_cast<const uint4*>(g + i);
        const uint4 mu = *reinterpret_cast<const uint4*>(m + i);
        const uint4 vu = *reinterpret_cast<const uint4*>(v + i);
        uint32_t pw[4] = {pu.x, pu.y, pu.z, pu.w}, gw[4] = {gu.x, gu.y, gu.z, gu.w};
        uint32_t mw[4] = {mu.x, mu.y, mu.z, mu.w}, vw[4] = {vu.x, vu.y, vu.z, vu.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 pf = unpack_bf16x2(pw[e]), gf = unpack_bf16x2(gw[e]), mf = unpack_bf16x2(mw[e]),
                   vf = unpack_bf16x2(vw[e]);
            adamw_update2(pf.x, pf.y, gf.x, gf.y, mf.x, mf.y, vf.x, vf.y, a, gmul, w1, one_m_b2, inv_bc2, step_size,
                          decay);
            pw[e] = pack_bf16x2(pf.x, pf.y);
            mw[e] = pack_bf16x2(mf.x, mf.y);
            vw[e] = pack_bf16x2(vf.x, vf.y);
        }
        *reinterpret_cast<uint4*>(p + i) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        *reinterpret_cast<uint4*>(m + i) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        *reinterpret_cast<uint4*>(v + i) = make_uint4(vw[0], vw[1], vw[2], vw[3]);
        if (zero_grad) *reinterpret_cast<uint4*>(g + i) = make_uint4(0, 0, 0, 0);
    }
}

static inline int grid_for(long long work_items, int block, int max_blocks = 148 * 16) {
    long long b = (work_items + block - 1) / block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

}  // namespace llark

using namespace llark;

extern "C" int llark_cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return LLARK_EINVAL;
    cast_f32_bf16_kernel<<<grid_for((n + 7) / 8, 256), 256, 0, st>>>(in, (__nv_bfloat16*)out, n);
    return check_launch();
}
extern "C" int llark_cast_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return LLARK_OK;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return LLARK_EINVAL;
    cast_bf16_f32_kernel<<<grid_for((n + 7) / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)in, out, n);
    return check_launch();
}

extern "C" int llark_splice_scan(const long long* ids, int B, int S, long long start_id, long long end_id,
                                 const int* frames, int* start_pos, int* flags, cudaStream_t st) {
    if (B <= 0 || S <= 0) return LLARK_EINVAL;
    splice_scan_kernel<<<B, 32, 0, st>>>(ids, B, S, start_id, end_id, frames, start_pos, flags);
    return check_launch();
}
extern "C" int llark_embed_splice_fwd(const long long* ids, const void* table, const void* audio,
                                      const long long* audio_off, const int* frames, const int* start_pos, void* out,
                                      int B, int S, int H, long long vocab, cudaStream_t st) {
    if (B <= 0 || S <= 0 || H % 8) return LLARK_EINVAL;
    dim3 grid(S, B);
    embed_splice_fwd_kernel<<<grid, 128, 0, st>>>(ids, (const __nv_bfloat16*)table, (const __nv_bfloat16*)audio,
                                                  audio_off, frames, start_pos, (__nv_bfloat16*)out, S, H, vocab);
    return check_launch();
}
extern "C" int llark_embed_splice_bwd(const long long* ids, const void* dout, void* daudio, const long long* audio_off,
                                      const int* frames, const int* start_pos, void* dtable, int B, int S, int H,
                                      long long vocab, int detach_text, cudaStream_t st) {
    if (B <= 0 || S <= 0 || H % 8) return LLARK_EINVAL;
    dim3 grid(S, B);
    embed_splice_bwd_kernel<<<grid, 128, 0, st>>>(ids, (const __nv_bfloat16*)dout, (__nv_bfloat16*)daudio, audio_off,
                                                  frames, start_pos, (__nv_bfloat16*)dtable, S, H, vocab, detach_text);
    return check_launch();
}

// H must be 8 * threads * VPT with threads a multiple of 32: pick (threads, VPT) for the supported widths.
// small_cta: prefer the SMALLEST CTA (largest VPT). A row is then handled by e.g. 128 threads x 4 vectors instead of 512 x 1,
// four times as many rows are resident per SM and four times the bytes are in flight: the one-row-per-CTA forward was
// latency-bound at 59 % of the HBM peak with 512-thread CTAs (profiles/r2f_prof_misc_hbm.md).
static bool norm_cfg(int H, int* threads, int* vpt, bool small_cta = false) {
    for (int k = 1; k <= 4; ++k) {
        const int v = small_cta ? 5 - k : k;
        if (H % (8 * v)) continue;
        const int t = H / (8 * v);
        if (t % 32 == 0 && t >= 32 && t <= 512) {
            *threads = t;
            *vpt = v;
            return true;
        }
    }
    return false;
}

extern "C" int llark_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, int H, float eps,
                                 cudaStream_t st) {
    int t, v;
    if (rows <= 0 || !norm_cfg(H, &t, &v, rows >= 1024)) return LLARK_EINVAL;
    const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w;
    __nv_bfloat16* yp = (__nv_bfloat16*)y;
    switch (v) {
        case 1: rmsnorm_fwd_kernel<1><<<(unsigned)rows, t, 0, st>>>(xp, wp, yp, rstd, H, eps); break;
        case 2: rmsnorm_fwd_kernel<2><<<(unsigned)rows, t, 0, st>>>(xp, wp, yp, rstd, H, eps); break;
        case 3: rmsnorm_fwd_kernel<3><<<(unsigned)rows, t, 0, st>>>(xp, wp, yp, rstd, H, eps); break;
        default: rmsnorm_fwd_kernel<4><<<(unsigned)rows, t, 0, st>>>(xp, wp, yp, rstd, H, eps); break;
    }
    return check_launch();
}

extern "C" long long llark_rmsnorm_bwd_workspace(long long rows, int H) {
    const long long nb = rows < 296 ? rows : 296;
    return nb * H * (long long)sizeof(float);
}
extern "C" int llark_rmsnorm_bwd(const void* x, const void* w, const void* dy, const float* rstd, const void* dres,
                                 void* dx, void* dw, int accumulate_dw, long long rows, int H, void* workspace,
                                 cudaStream_t st) {
    int t, v;
    // (measured, profiles/r2h_prof_opt_hbm.md: a two-rows-per-iteration variant on 128-thread CTAs ran at 2.5 TB/s against
    // 3.7-4.5 TB/s for this one-row, 512-thread form -- 255 registers left 8 warps per SM -- so the backward keeps it)
    if (rows <= 0 || !norm_cfg(H, &t, &v) || !workspace) return LLARK_EINVAL;
    const int nb = (int)(rows < 296 ? rows : 296);
    const __nv_bfloat16 *xp = (const __nv_bfloat16*)x, *wp = (const __nv_bfloat16*)w, *dyp = (const __nv_bfloat16*)dy,
                        *drp = (const __nv_bfloat16*)dres;
    __nv_bfloat16* dxp = (__nv_bfloat16*)dx;
    float* part = (float*)workspace;
    switch (v) {
        case 1: rmsnorm_bwd_kernel<1><<<nb, t, 0, st>>>(xp, wp, dyp, rstd, drp, dxp, part, rows, H); break;
        case 2: rmsnorm_bwd_kernel<2><<<nb, t, 0, st>>>(xp, wp, dyp, rstd, drp, dxp, part, rows, H); break;
        case 3: rmsnorm_bwd_kernel<3><<<nb, t, 0, st>>>(xp, wp, dyp, rstd, drp, dxp, part, rows, H); break;
        default: rmsnorm_bwd_kernel<4><<<nb, t, 0, st>>>(xp, wp, dyp, rstd, drp, dxp, part, rows, H); break;
    }
    int rc = check_launch();
    if (rc) return rc;
    if (dw) {
        colsum_partials_kernel<<<(H + 255) / 256, 256, 0, st>>>(part, nb, H, (__nv_bfloat16*)dw, accumulate_dw);
        rc = check_launch();
    }
    return rc;
}

extern "C" int llark_rope(void* qkv, const void* cos_t, const void* sin_t, long long rows, int S, int n_heads, int D,
                          long long row_stride, int inverse, cudaStream_t st) {
    if (rows <= 0 || D % 16 || S <= 0) return LLARK_EINVAL;
    const long long total = rows * 2 * n_heads * (D / 16);
    rope_kernel<<<grid_for(total, 256), 256, 0, st>>>((__nv_bfloat16*)qkv, (const __nv_bfloat16*)cos_t,
                                                       (const __nv_bfloat16*)sin_t, rows, S, n_heads, D, row_stride,
                                                       inverse);
    return check_launch();
}

extern "C" int llark_swiglu_fwd(const void* gu, void* act, long long rows, int I, cudaStream_t st) {
    if (rows <= 0 || I % 8) return LLARK_EINVAL;
    swiglu_fwd_kernel<<<grid_for(rows * (I / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)gu, (__nv_bfloat16*)act, rows, I);
    return check_launch();
}
extern "C" int llark_swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int I, cudaStream_t st) {
    if (rows <= 0 || I % 8) return LLARK_EINVAL;
    swiglu_bwd_kernel<<<grid_for(rows * (I / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)gu, (const __nv_bfloat16*)dact,
                                                                     (__nv_bfloat16*)dgu, rows, I);
    return check_launch();
}

extern "C" int llark_ce_fwd(const void* logits, long long ldv, int V, const long long* labels, int B, int S,
                            float* row_loss, float* row_lse, float* loss_out, cudaStream_t st) {
    if (B <= 0 || S <= 0 || V <= 0 || ldv % 8 || ldv < V) return LLARK_EINVAL;
    const long long rows = (long long)B * S;
    ce_fwd_kernel<<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)logits, ldv, V, labels, S, row_loss, row_lse);
    int rc = check_launch();
    if (rc) return rc;
    ce_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, row_lse, rows, loss_out);
    return check_launch();
}
extern "C" int llark_ce_bwd(const void* logits, long long ldv, int V, const long long* labels, int B, int S,
                            const float* row_lse, const float* loss_out, const float* gscale_ptr, float gscale,
                            void* dlogits, cudaStream_t st) {
    if (B <= 0 || S <= 0 || V <= 0 || ldv % 8 || ldv < V) return LLARK_EINVAL;
    const long long rows = (long long)B * S;
    ce_bwd_kernel<<<(unsigned)rows, 256, 0, st>>>((const __nv_bfloat16*)logits, ldv, V, labels, S, row_lse, loss_out,
                                                  gscale_ptr, gscale, (__nv_bfloat16*)dlogits);
    return check_launch();
}

#define LLARK_SUMSQ_BLOCKS 1184
extern "C" long long llark_sumsq_workspace(void) { return LLARK_SUMSQ_BLOCKS * (long long)sizeof(float); }
extern "C" int llark_grad_sumsq(const void* g, long long n, float* out, int accumulate, void* workspace,
                                cudaStream_t st) {
    if (n <= 0 || !workspace || (reinterpret_cast<uintptr_t>(g) & 15)) return LLARK_EINVAL;
    const int nb = grid_for((n + 7) / 8, 256, LLARK_SUMSQ_BLOCKS);
    sumsq_partial_kernel<<<nb, 256, 0, st>>>((const __nv_bfloat16*)g, n, (float*)workspace);
    int rc = check_launch();
    if (rc) return rc;
    sum_partials_kernel<<<1, 1024, 0, st>>>((const float*)workspace, nb, out, accumulate);
    return check_launch();
}

extern "C" int llark_adamw_step(void* p, void* g, void* m, void* v, long long n, const float* sumsq, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                float max_norm, int zero_grad, cudaStream_t st) {
    if (n <= 0 || n % 8 || step < 1) return LLARK_EINVAL;
    if ((reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) ||
        (reinterpret_cast<uintptr_t>(m) & 15) || (reinterpret_cast<uintptr_t>(v) & 15))
        return LLARK_EINVAL;
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.grad_scale = grad_scale; a.max_norm = max_norm;
    adamw_kernel<<<grid_for(n / 8, 256, 148 * 8), 256, 0, st>>>((__nv_bfloat16*)p, (__nv_bfloat16*)g, (__nv_bfloat16*)m,
                                                                (__nv_bfloat16*)v, n, sumsq, a, zero_grad);
    return check_launch();
}
