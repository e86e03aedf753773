// Front end (encoder) and back end (overlap-add of decoder frames, crop,
// mixture consistency).
//
// Reference semantics:
//   pad_to_appropriate_length   improved_sudormrf.py:303-314 (zeros to Tp; folded into the
//                               load predicate here: no padded copy is ever materialised)
//   encoder                     improved_sudormrf.py:247-251,286  Conv1d(A,N,K,stride=K/2,pad=K/2)
//                               (original model, sudormrf.py:212-218,269: the same Conv1d with a bias, then ReLU;
//                                its ConvTranspose1d decoder, :245-252, has one bias per source)
//   decoder                     improved_sudormrf.py:272-279,300  ConvTranspose1d(...,stride=K/2,
//                               padding=K/2, output_padding=K/2-1)  -> length hop*L
//   remove_trailing_zeros       improved_sudormrf.py:316-318      crop to T
//   mixture_consistency.apply   mixture_consistency.py:14-36
#include "common.cuh"

namespace sdr {

// ---------------------------------------------------------------------------
// encoder: enc[b,n,t] = sum_a sum_j w[n,a,j] * wav[b,a, hop*t + j - pad]
// CTA = 128 positions x kEncNB basis functions; the waveform chunk and the
// weight slab sit in shared memory; each thread owns one position and walks the
// basis functions four at a time (weights read as broadcast float4).
// ---------------------------------------------------------------------------
constexpr int kEncThreads = 128;
constexpr int kEncNB = 64;

__global__ void __launch_bounds__(kEncThreads)
encoder_kernel(const float* __restrict__ wav, const float* __restrict__ weight, const float* __restrict__ bias,
               float* __restrict__ enc, double* __restrict__ stats,
               int A, long long T, int N, int K, int L, int t_tiles, int pad, int relu) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_red[64];
    const int hop = K / 2;
    const int span = hop * (kEncThreads - 1) + K;        // samples needed by 128 positions
    float* s_x = smem;                                     // [A][span]
    float* s_w = smem + ((A * span + 3) & ~3);             // [A*K][kEncNB]  (n fastest)

    const int b = blockIdx.x / t_tiles;
    const int t0 = (blockIdx.x - b * t_tiles) * kEncThreads;
    const int n0 = blockIdx.y * kEncNB;
    const int tid = threadIdx.x;

    const long long base = (long long)hop * t0 - pad;      // first sample index of the chunk (pad = hop; 2 * hop for the causal model)
    for (int i = tid; i < A * span; i += kEncThreads) {
        const int a = i / span, p = i - a * span;
        const long long g = base + p;
        s_x[i] = (g >= 0 && g < T) ? __ldg(wav + ((size_t)b * A + a) * T + g) : 0.f;
    }
    for (int i = tid; i < A * K * kEncNB; i += kEncThreads) {
        const int n = i % kEncNB, aj = i / kEncNB;         // aj = a*K + j
        s_w[i] = (n0 + n < N) ? __ldg(weight + (size_t)(n0 + n) * A * K + aj) : 0.f;
    }
    __syncthreads();

    const int t = t0 + tid;
    float st_s = 0.f, st_q = 0.f;
    const int AK = A * K;
    for (int nn = 0; nn < kEncNB; nn += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int a = 0; a < A; ++a) {
            const float* xr = s_x + a * span + hop * tid;
            const float* wr = s_w + (size_t)a * K * kEncNB + nn;
            for (int j = 0; j < K; ++j) {
                const float xv = xr[j];
                const float4 w = *reinterpret_cast<const float4*>(wr + j * kEncNB);
                a0 = fmaf(w.x, xv, a0); a1 = fmaf(w.y, xv, a1);
                a2 = fmaf(w.z, xv, a2); a3 = fmaf(w.w, xv, a3);
            }
        }
        (void)AK;
        if (t < L) {
            float o[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = n0 + nn + e;
                if (n < N) {
                    if (bias) o[e] += __ldg(bias + n);
                    if (relu) o[e] = fmaxf(o[e], 0.f);
                    enc[((size_t)b * N + n) * L + t] = o[e];
                    st_s += o[e]; st_q = fmaf(o[e], o[e], st_q);
                }
            }
        }
    }
    if (stats) block_stats_atomic(st_s, st_q, stats, b, s_red);
}

int launch_encoder(const float* wav, const float* weight, const float* bias, int relu, float* enc, double* stats,
                   int B, int A, long long T, int N, int K, int L, int pad, cudaStream_t st) {
    if (B <= 0 || A <= 0 || T <= 0 || N <= 0 || K < 3 || L <= 0) return SDR_ERR_BAD_ARGUMENT;
    const int hop = K / 2;
    const int span = hop * (kEncThreads - 1) + K;
    const size_t smem = (size_t)(((A * span + 3) & ~3) + A * K * kEncNB) * sizeof(float);
    if (smem > 200 * 1024) return SDR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return SDR_ERR_CUDA;
    }
    const int t_tiles = (L + kEncThreads - 1) / kEncThreads;
    const long long gx = (long long)t_tiles * B;
    if (gx > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)((N + kEncNB - 1) / kEncNB));
    encoder_kernel<<<grid, kEncThreads, smem, st>>>(wav, weight, bias, enc, stats, A, T, N, K, L, t_tiles, pad, relu);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// overlap-add: frames[b, sa*K + j, t] (= sum_c Wd[c,sa,j] masked[b,c,t]) ->
// out[b, sa, tau] = sum over (t, j) with hop*t + j - hop == tau, tau < T;
// optional uniform mixture consistency (needs all sources of a tau in one thread).
// ---------------------------------------------------------------------------
constexpr int kMaxSrc = 16;

__global__ void __launch_bounds__(256)
overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ mix, const float* __restrict__ bias,
                   const float2* __restrict__ rescale, float* __restrict__ out, int SA, int K, int L, long long T) {
    const int hop = K / 2;
    const long long tau = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (tau >= T) return;
    // j = tau + hop - hop*t in [0, K)  ->  t in [ceil((tau+hop-K+1)/hop), floor((tau+hop)/hop)]
    const long long thi = (tau + hop) / hop;
    long long tlo = tau + hop - (K - 1);
    tlo = tlo <= 0 ? 0 : (tlo + hop - 1) / hop;
    float est[kMaxSrc];
    float sum = 0.f;
    // separate(): undo the per-utterance input normalisation, est * std + mean (README.md:109), before the
    // mixture-consistency projection, which the README applies to the rescaled estimates (README.md:113-114)
    const float2 rs = rescale ? rescale[b] : make_float2(0.f, 1.f);
    for (int s = 0; s < SA; ++s) {
        float acc = 0.f;
        for (long long t = tlo; t <= thi && t < L; ++t) {
            const int j = (int)(tau + hop - hop * t);
            acc += __ldg(frames + ((size_t)b * SA * K + (size_t)s * K + j) * L + t);
        }
        if (bias) acc += __ldg(bias + s);            // decoder bias of the original model (one per source)
        if (rescale) acc = __fadd_rn(__fmul_rn(acc, rs.y), rs.x);
        est[s] = acc;
        sum += acc;
    }
    float corr = 0.f;
    if (mix) corr = (__ldg(mix + (size_t)b * T + tau) - sum) * (1.0f / SA);   // mixture_consistency.py:29-35
    for (int s = 0; s < SA; ++s) out[((size_t)b * SA + s) * T + tau] = est[s] + corr;
}

int launch_overlap_add(const float* frames, const float* mix, const float* bias, const float2* rescale, float* out,
                       int B, int SA, int K, int L, long long T, cudaStream_t st) {
    if (B <= 0 || SA <= 0 || K < 3 || L <= 0 || T <= 0) return SDR_ERR_BAD_ARGUMENT;
    if (SA > kMaxSrc || B > 65535) return SDR_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((T + 255) / 256), (unsigned)B);
    overlap_add_kernel<<<grid, 256, 0, st>>>(frames, mix, bias, rescale, out, SA, K, L, T);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// standalone mixture consistency (mixture_consistency.py:14-36)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mc_power_kernel(const float* __restrict__ est, double* __restrict__ power, long long T) {
    // power[b*S+s] += sum_t est^2   (grid.y = B*S)
    __shared__ float red[32];
    const size_t row = blockIdx.y;
    float q = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T;
         t += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(est + row * T + t);
        q = fmaf(v, v, q);
    }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x < 32) {
        double d = threadIdx.x < (blockDim.x >> 5) ? (double)red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (threadIdx.x == 0) atomicAdd(power + row, d);
    }
}

__global__ void __launch_bounds__(256)
mc_apply_kernel(const float* __restrict__ est, const float* __restrict__ mix,
                const double* __restrict__ power, float* __restrict__ out, int S, long long T) {
    const long long tau = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (tau >= T) return;
    float sum = 0.f;
    for (int s = 0; s < S; ++s) sum += __ldg(est + ((size_t)b * S + s) * T + tau);
    const float resid = __ldg(mix + (size_t)b * T + tau) - sum;
    float wsum = 0.f;
    if (power) {
        for (int s = 0; s < S; ++s) wsum += (float)(power[(size_t)b * S + s] / (double)T);
    }
    for (int s = 0; s < S; ++s) {
        float w;
        if (power) {
            const float mw = (float)(power[(size_t)b * S + s] / (double)T);  // mean(est^2, -1)
            w = mw / (wsum + 1e-9f);                                          // mixture_consistency.py:27-28
        } else {
            w = 1.0f / S;
        }
        const size_t i = ((size_t)b * S + s) * T + tau;
        out[i] = __ldg(est + i) + w * resid;
    }
}

int launch_mixture_consistency(const float* est, const float* mix, float* out, int B, int S,
                               long long T, int weights_type, void* scratch, cudaStream_t st) {
    if (B <= 0 || S <= 0 || T <= 0 || !est || !mix || !out) return SDR_ERR_BAD_ARGUMENT;
    if (B > 65535 || (long long)B * S > 65535) return SDR_ERR_UNSUPPORTED;
    double* power = nullptr;
    if (weights_type == 1) {
        if (!scratch) return SDR_ERR_BAD_ARGUMENT;
        power = static_cast<double*>(scratch);
        if (cudaMemsetAsync(power, 0, sizeof(double) * B * S, st) != cudaSuccess) return SDR_ERR_CUDA;
        int gx = (int)((T + 256 * 8 - 1) / (256 * 8));
        if (gx < 1) gx = 1;
        mc_power_kernel<<<dim3((unsigned)gx, (unsigned)(B * S)), 256, 0, st>>>(est, power, T);
    } else if (weights_type != 0) {
        return SDR_ERR_BAD_ARGUMENT;
    }
    dim3 grid((unsigned)((T + 255) / 256), (unsigned)B);
    mc_apply_kernel<<<grid, 256, 0, st>>>(est, mix, power, out, S, T);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

}  // namespace sdr
