// Skinny linear layer for the single-token decode step of the generation path:
//   C[M, N] = A[M, K] . W[N, K]^T (+ bias[N]) (+ R[M, N]),  M <= 8 rows, bf16 in / out, fp32 accumulation.
// With M = batch (1 for the reference's inference scripts) the tcgen05 GEMM would run 128-row tiles that are >99 % padding
// and, for N = 4096, only 16 CTA pairs: the op is a pure stream of the weight matrix (N * K * 2 bytes, 13.5 GB per token over
// a 7B model), so it is written as an HBM-bound SIMT kernel instead: one warp per NCOL weight rows, every lane streams
// 16-byte vectors of W (coalesced 512 B per warp and row), the M activation rows come from L1 / L2 (22 KB for K = 11008),
// warp-shuffle reduction, deterministic. Algorithmic bytes per launch: 2 * N * K (+ 2 * M * (K + N)).
// Used by LlamaEngine.decode_step (the default path of generate()); first run on a B200 in round 2 (tests/test_kernels_gpu.py).
#include "common.cuh"

namespace llark {
extern thread_local int g_last_cuda_error;
namespace gemv {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

constexpr int kWarps = 8;

// grid: ceil(N / (kWarps * NCOL)) CTAs of kWarps warps; warp w of CTA c owns weight rows n0 .. n0 + NCOL - 1.
template <int M, int NCOL>
__global__ void __launch_bounds__(kWarps * 32)
gemv_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ W, long long ldw,
            __nv_bfloat16* __restrict__ C, long long ldc, const __nv_bfloat16* __restrict__ bias,
            const __nv_bfloat16* __restrict__ R, long long ldr, int N, int K) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n0 = ((long long)blockIdx.x * kWarps + warp) * NCOL;
    if (n0 >= N) return;
    float acc[M][NCOL];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int c = 0; c < NCOL; ++c) acc[m][c] = 0.f;
    const int kvec = K >> 3;  // 16-byte vectors per row
    const uint4* wrow[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
        const long long n = (n0 + c < N) ? n0 + c : (long long)N - 1;  // tail rows re-read the last row, never stored
        wrow[c] = reinterpret_cast<const uint4*>(W + n * ldw);
    }
    for (int v = lane; v < kvec; v += 32) {
        uint4 wv[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) wv[c] = __ldcs(wrow[c] + v);  // streamed once: evict-first
        float wf[NCOL][8];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) unpack8(wv[c], wf[c]);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float xf[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(A + (long long)m * lda) + v), xf);
#pragma unroll
            for (int c = 0; c < NCOL; ++c)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[m][c] = fmaf(xf[e], wf[c][e], acc[m][c]);
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            float s = warp_sum(acc[m][c]);
            if (lane == 0 && n0 + c < N) {
                const long long n = n0 + c;
                if (bias) s += __bfloat162float(bias[n]);
                if (R) s += __bfloat162float(R[(long long)m * ldr + n]);
                C[(long long)m * ldc + n] = __float2bfloat16_rn(s);
            }
        }
}

template <int M>
static void launch(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, __nv_bfloat16* C,
                   long long ldc, const __nv_bfloat16* bias, const __nv_bfloat16* R, long long ldr, int N, int K,
                   cudaStream_t st) {
    // enough warps to cover 148 SMs several times over: 4 rows per warp for the wide projections, 2 for N <= 8192
    if ((long long)N >= 4LL * kWarps * 148 * 4) {
        const unsigned grid = (unsigned)((N + kWarps * 4 - 1) / (kWarps * 4));
        gemv_kernel<M, 4><<<grid, kWarps * 32, 0, st>>>(A, lda, W, ldw, C, ldc, bias, R, ldr, N, K);
    } else {
        const unsigned grid = (unsigned)((N + kWarps * 2 - 1) / (kWarps * 2));
        gemv_kernel<M, 2><<<grid, kWarps * 32, 0, st>>>(A, lda, W, ldw, C, ldc, bias, R, ldr, N, K);
    }
}

}  // namespace gemv
}  // namespace llark

using namespace llark;

extern "C" int llark_gemv(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                          const void* bias, const void* R, long long ldr, int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || M > 8 || N <= 0 || K <= 0 || (K & 7) || (lda & 7) || (ldw & 7) || lda < K || ldw < K || ldc < N ||
        (R && ldr < N))
        return LLARK_EINVAL;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) return LLARK_EINVAL;
    const __nv_bfloat16 *a = (const __nv_bfloat16*)A, *w = (const __nv_bfloat16*)W, *b = (const __nv_bfloat16*)bias,
                        *r = (const __nv_bfloat16*)R;
    __nv_bfloat16* c = (__nv_bfloat16*)C;
    switch (M) {
        case 1: gemv::launch<1>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 2: gemv::launch<2>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 3: gemv::launch<3>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 4: gemv::launch<4>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 5: gemv::launch<5>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 6: gemv::launch<6>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        case 7: gemv::launch<7>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
        default: gemv::launch<8>(a, lda, w, ldw, c, ldc, b, r, ldr, N, K, st); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return LLARK_ECUDA;
    }
    return LLARK_OK;
}
