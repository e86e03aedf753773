// U-ConvBlock level kernels (HBM-bound): depthwise k=5 stride 1|2 with the
// producer's GlobLN(+PReLU) applied on load, and the multi-resolution merge.
//
// Reference semantics:
//   DilatedConvNorm  improved_sudormrf.py:138-159  (Conv1d k=5, pad 2, groups=C, then GlobLN)
//   UConvBlock.forward level loops  improved_sudormrf.py:206-216
// Each kernel stores its RAW result once and accumulates the (sum, sumsq) of
// that result per sample in fp64 so the consumer can normalise while loading.
#include "common.cuh"

namespace sdr {

constexpr int kDwThreads = 256;
constexpr int kDwItems = 2;        // output quads per thread

// ---------------------------------------------------------------------------
// depthwise, vector path: one item = 4 consecutive outputs of one (sample, c) row
// requires Lout % 4 == 0 (then Lin % 4 == 0 as well).
// ---------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(kDwThreads)
dw5_vec_kernel(const float* __restrict__ x, NormIn nin,
               const float* __restrict__ w5, const float* __restrict__ bias,
               float* __restrict__ y, double* __restrict__ stats_out,
               int C, int Lin, int Lout, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    __syncthreads();
    const SampleNorm sn = s_norm;

    const int QR = Lout >> 2;                 // quads per row
    const int items = C * QR;                 // per sample
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;

    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kDwItems; ++it) {
        const int item = (chunk * kDwItems + it) * kDwThreads + threadIdx.x;
        if (item < items) {
            const int c = item / QR;
            const int q = item - c * QR;
            const ChanNorm cn = chan_norm(nin, sn, c);
            const float* xr = xs + (size_t)c * Lin;
            float w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = __ldg(w5 + c * 5 + j);
            const float b = __ldg(bias + c);
            float o[4];
            if (STRIDE == 1) {
                // window v[0..7] = positions 4q-2 .. 4q+5
                float v[8];
                const float4 m = ldg4(xr + 4 * q);
                v[2] = apply_norm(cn, m.x); v[3] = apply_norm(cn, m.y);
                v[4] = apply_norm(cn, m.z); v[5] = apply_norm(cn, m.w);
                if (q > 0) {
                    const float2 l = __ldg(reinterpret_cast<const float2*>(xr + 4 * q - 2));
                    v[0] = apply_norm(cn, l.x); v[1] = apply_norm(cn, l.y);
                } else { v[0] = 0.f; v[1] = 0.f; }
                if (q < QR - 1) {
                    const float2 r = __ldg(reinterpret_cast<const float2*>(xr + 4 * q + 4));
                    v[6] = apply_norm(cn, r.x); v[7] = apply_norm(cn, r.y);
                } else { v[6] = 0.f; v[7] = 0.f; }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[i + j], a);
                    o[i] = a;
                }
            } else {
                // outputs 4q..4q+3 read inputs 8q-2 .. 8q+8 ; window v[0..10]
                float v[11];
                const float4 m0 = ldg4(xr + 8 * q);
                const float4 m1 = ldg4(xr + 8 * q + 4);
                v[2] = apply_norm(cn, m0.x); v[3] = apply_norm(cn, m0.y);
                v[4] = apply_norm(cn, m0.z); v[5] = apply_norm(cn, m0.w);
                v[6] = apply_norm(cn, m1.x); v[7] = apply_norm(cn, m1.y);
                v[8] = apply_norm(cn, m1.z); v[9] = apply_norm(cn, m1.w);
                if (q > 0) {
                    const float2 l = __ldg(reinterpret_cast<const float2*>(xr + 8 * q - 2));
                    v[0] = apply_norm(cn, l.x); v[1] = apply_norm(cn, l.y);
                } else { v[0] = 0.f; v[1] = 0.f; }
                v[10] = (8 * q + 8 < Lin) ? apply_norm(cn, __ldg(xr + 8 * q + 8)) : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[2 * i + j], a);
                    o[i] = a;
                }
            }
            *reinterpret_cast<float4*>(ys + (size_t)c * Lout + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

// scalar fallback (any Lin/Lout): one output element per thread-iteration
__global__ void __launch_bounds__(kDwThreads)
dw5_scalar_kernel(const float* __restrict__ x, NormIn nin,
                  const float* __restrict__ w5, const float* __restrict__ bias,
                  float* __restrict__ y, double* __restrict__ stats_out,
                  int C, int Lin, int Lout, int stride, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    __syncthreads();
    const SampleNorm sn = s_norm;
    const int items = C * Lout;
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;
    float acc_s = 0.f, acc_q = 0.f;
    for (int it = 0; it < 4; ++it) {
        const int item = (chunk * 4 + it) * kDwThreads + threadIdx.x;
        if (item < items) {
            const int c = item / Lout;
            const int t = item - c * Lout;
            const ChanNorm cn = chan_norm(nin, sn, c);
            float a = __ldg(bias + c);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int p = t * stride + j - 2;
                if (p >= 0 && p < Lin)
                    a = fmaf(__ldg(w5 + c * 5 + j), apply_norm(cn, __ldg(xs + (size_t)c * Lin + p)), a);
            }
            ys[(size_t)c * Lout + t] = a;
            acc_s += a; acc_q = fmaf(a, a, acc_q);
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

// ---------------------------------------------------------------------------
// merge: m[c,t] = sum_d norm_d(z_d)[c, t >> d]   (closed form of the
// upsample(scale 2, nearest)+add chain, improved_sudormrf.py:214-216)
// ---------------------------------------------------------------------------
constexpr int kMaxDepth = kMaxDepthApi;
struct MergeArgs {
    const float* z[kMaxDepth];
    NormIn n[kMaxDepth];
    int depth;
};

constexpr int kMgThreads = 256;
constexpr int kMgItems = 2;

// vector path: L % 4 == 0 and L % 2^(depth-1) == 0
__global__ void __launch_bounds__(kMgThreads)
merge_vec_kernel(MergeArgs a, float* __restrict__ m, double* __restrict__ stats_out,
                 int C, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm[kMaxDepth];
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x < a.depth) s_norm[threadIdx.x] = sample_norm(a.n[threadIdx.x], sample);
    __syncthreads();

    const int QR = L >> 2;
    const int items = C * QR;
    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kMgItems; ++it) {
        const int item = (chunk * kMgItems + it) * kMgThreads + threadIdx.x;
        if (item < items) {
            const int c = item / QR;
            const int q = item - c * QR;
            const size_t row = (size_t)sample * C + c;
            float o[4];
            {
                const ChanNorm cn = chan_norm(a.n[0], s_norm[0], c);
                const float4 v = ldg4(a.z[0] + row * L + 4 * q);
                o[0] = apply_norm(cn, v.x); o[1] = apply_norm(cn, v.y);
                o[2] = apply_norm(cn, v.z); o[3] = apply_norm(cn, v.w);
            }
            if (a.depth > 1) {
                const ChanNorm cn = chan_norm(a.n[1], s_norm[1], c);
                const float2 v = __ldg(reinterpret_cast<const float2*>(a.z[1] + row * (L >> 1) + 2 * q));
                const float v0 = apply_norm(cn, v.x), v1 = apply_norm(cn, v.y);
                o[0] += v0; o[1] += v0; o[2] += v1; o[3] += v1;
            }
            for (int d = 2; d < a.depth; ++d) {
                const ChanNorm cn = chan_norm(a.n[d], s_norm[d], c);
                const float v = apply_norm(cn, __ldg(a.z[d] + row * (L >> d) + (q >> (d - 2))));
                o[0] += v; o[1] += v; o[2] += v; o[3] += v;
            }
            *reinterpret_cast<float4*>(m + row * L + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

__global__ void __launch_bounds__(kMgThreads)
merge_scalar_kernel(MergeArgs a, float* __restrict__ m, double* __restrict__ stats_out,
                    int C, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm[kMaxDepth];
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x < a.depth) s_norm[threadIdx.x] = sample_norm(a.n[threadIdx.x], sample);
    __syncthreads();
    const int items = C * L;
    float acc_s = 0.f, acc_q = 0.f;
    for (int it = 0; it < 4; ++it) {
        const int item = (chunk * 4 + it) * kMgThreads + threadIdx.x;
        if (item < items) {
            const int c = item / L;
            const int t = item - c * L;
            const size_t row = (size_t)sample * C + c;
            float o = 0.f;
            for (int d = 0; d < a.depth; ++d) {
                const ChanNorm cn = chan_norm(a.n[d], s_norm[d], c);
                o += apply_norm(cn, __ldg(a.z[d] + row * (L >> d) + (t >> d)));
            }
            m[row * L + t] = o;
            acc_s += o; acc_q = fmaf(o, o, acc_q);
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}


// ---------------------------------------------------------------------------
// Wide paths (the ones the benchmark shapes take).  The narrow kernels above were
// instruction-issue bound (ncu: issue-active 70-80 %, DRAM 40-58 %): per output
// they paid ~30 instructions of parameter loads, index division and a 5-op
// normalise+PReLU.  Here a thread owns a run of 8 (depthwise) or 16 (merge)
// consecutive outputs of one row, the GlobLN mean is folded into the shift
// (y = x*a + b, a = gamma*rstd, b = beta - mean*a) and PReLU is 2 ops
// (t = y*slope; y = slope <= 1 ? max(y,t) : min(y,t)), so the kernels sit on the
// HBM roofline instead of the issue roofline.
// ---------------------------------------------------------------------------
struct FoldedNorm { float a, b; };
__device__ __forceinline__ FoldedNorm fold_norm(const NormIn& n, const SampleNorm& s, int c) {
    FoldedNorm f{1.f, 0.f};
    if (n.stats) { f.a = __ldg(n.gamma + c) * s.rstd; f.b = fmaf(-s.mean, f.a, __ldg(n.beta + c)); }
    return f;
}
template <bool ACT>
__device__ __forceinline__ float norm_act(float x, const FoldedNorm& f, float slope, bool slope_le1) {
    float y = fmaf(x, f.a, f.b);
    if (ACT) { const float t = y * slope; y = slope_le1 ? fmaxf(y, t) : fminf(y, t); }
    return y;
}

// tuning knobs, swept on the B200 (profiles/r01_level_kernel_sweep.md): 128 threads x 4 runs (2 for the
// merge) beat 256x2 / 512x1 / 64x8; ld.global.nc.L1::no_allocate + st.global.cs made stride-2 levels ~25 % slower
#ifndef SDR_DW_THREADS
#define SDR_DW_THREADS 128
#endif
#ifndef SDR_DW_ITEMS
#define SDR_DW_ITEMS 4
#endif
#ifndef SDR_MG_THREADS
#define SDR_MG_THREADS 128
#endif
#ifndef SDR_MG_ITEMS
#define SDR_MG_ITEMS 2
#endif
#ifndef SDR_STREAM_HINTS
#define SDR_STREAM_HINTS 0
#endif
#ifndef SDR_DW_PIPELINE
#define SDR_DW_PIPELINE 0               // 1: wide depthwise kernel with the next run's loads ahead of the arithmetic (experiment)
#endif
// SDR_DW_CHAIN (common.cuh): 1 = stride-2 levels of a block in one persistent kernel (experiment, see dw5_chain_kernel)
// streaming accesses: every byte of these kernels is touched once, so (optionally) keep it out of L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
#if SDR_STREAM_HINTS
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return ldg4(p);
#endif
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
#if SDR_STREAM_HINTS
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
    *reinterpret_cast<float4*>(p) = v;
#endif
}
constexpr int kDw8Threads = SDR_DW_THREADS;
constexpr int kDw8Items = SDR_DW_ITEMS;       // runs of 8 outputs per thread

// requires Lout % 8 == 0
template <int STRIDE, bool ACT>
__global__ void __launch_bounds__(kDw8Threads)
dw5_wide_kernel(const float* __restrict__ x, NormIn nin,
                const float* __restrict__ w5, const float* __restrict__ bias,
                float* __restrict__ y, double* __restrict__ stats_out,
                int C, int Lin, int Lout, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    const float slope = ACT ? __ldg(nin.prelu) : 1.f;
    const bool sle1 = slope <= 1.f;
    __syncthreads();
    const SampleNorm sn = s_norm;

    const int QR = Lout >> 3;                 // runs per row
    const int items = C * QR;                 // per sample
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;
    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kDw8Items; ++it) {
        const int item = (chunk * kDw8Items + it) * kDw8Threads + threadIdx.x;
        if (item < items) {
            const int c = item / QR;
            const int q = item - c * QR;
            const FoldedNorm f = fold_norm(nin, sn, c);
            float w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = __ldg(w5 + c * 5 + j);
            const float b = __ldg(bias + c);
            float o[8];
            if (STRIDE == 1) {
                const float* xr = xs + (size_t)c * Lin + 8 * q;       // window v[0..11] = positions 8q-2 .. 8q+9
                float v[12];
                const float4 m0 = ld_stream4(xr), m1 = ld_stream4(xr + 4);
                float2 l = make_float2(0.f, 0.f), r = make_float2(0.f, 0.f);
                const bool hl = q > 0, hr = q < QR - 1;
                if (hl) l = __ldg(reinterpret_cast<const float2*>(xr - 2));
                if (hr) r = __ldg(reinterpret_cast<const float2*>(xr + 8));
                v[0] = hl ? norm_act<ACT>(l.x, f, slope, sle1) : 0.f;
                v[1] = hl ? norm_act<ACT>(l.y, f, slope, sle1) : 0.f;
                v[2] = norm_act<ACT>(m0.x, f, slope, sle1); v[3] = norm_act<ACT>(m0.y, f, slope, sle1);
                v[4] = norm_act<ACT>(m0.z, f, slope, sle1); v[5] = norm_act<ACT>(m0.w, f, slope, sle1);
                v[6] = norm_act<ACT>(m1.x, f, slope, sle1); v[7] = norm_act<ACT>(m1.y, f, slope, sle1);
                v[8] = norm_act<ACT>(m1.z, f, slope, sle1); v[9] = norm_act<ACT>(m1.w, f, slope, sle1);
                v[10] = hr ? norm_act<ACT>(r.x, f, slope, sle1) : 0.f;
                v[11] = hr ? norm_act<ACT>(r.y, f, slope, sle1) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[i + j], a);
                    o[i] = a;
                }
            } else {
                const float* xr = xs + (size_t)c * Lin + 16 * q;      // window v[0..18] = positions 16q-2 .. 16q+16
                float v[19];
                const float4 m0 = ld_stream4(xr), m1 = ld_stream4(xr + 4), m2 = ld_stream4(xr + 8), m3 = ld_stream4(xr + 12);
                float2 l = make_float2(0.f, 0.f);
                float r = 0.f;
                const bool hl = q > 0, hr = 16 * q + 16 < Lin;
                if (hl) l = __ldg(reinterpret_cast<const float2*>(xr - 2));
                if (hr) r = __ldg(xr + 16);
                v[0] = hl ? norm_act<ACT>(l.x, f, slope, sle1) : 0.f;
                v[1] = hl ? norm_act<ACT>(l.y, f, slope, sle1) : 0.f;
                const float mm[16] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w,
                                      m2.x, m2.y, m2.z, m2.w, m3.x, m3.y, m3.z, m3.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) v[2 + i] = norm_act<ACT>(mm[i], f, slope, sle1);
                v[18] = hr ? norm_act<ACT>(r, f, slope, sle1) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[2 * i + j], a);
                    o[i] = a;
                }
            }
            float* yr = ys + (size_t)c * Lout + 8 * q;
            st_stream4(yr, make_float4(o[0], o[1], o[2], o[3]));
            st_stream4(yr + 4, make_float4(o[4], o[5], o[6], o[7]));
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

#if SDR_DW_PIPELINE
// ---------------------------------------------------------------------------
// EXPERIMENT (round-2 candidate, compiled only with -DSDR_DW_PIPELINE=1, not measured yet): the same kernel with the
// loads of run i+1 issued before the arithmetic of run i.  The ncu source view of dw5_wide_kernel<2,0> put 62 % of the
// stall samples on the first FFMA of each run (four serial load -> compute -> store round trips per thread).
// ---------------------------------------------------------------------------
template <int STRIDE>
struct DwRun {
    float4 m[STRIDE == 1 ? 2 : 4];
    float2 l;
    float2 r;                 // stride 1: two right-halo values; stride 2: r.x only
    int c, q;
    bool ok, hl, hr;
};

template <int STRIDE>
__device__ __forceinline__ DwRun<STRIDE> dw_load_run(const float* __restrict__ xs, int item, int items, int QR, int Lin) {
    DwRun<STRIDE> d;
    d.ok = item < items;
    d.c = 0; d.q = 0; d.hl = false; d.hr = false;
    d.l = make_float2(0.f, 0.f); d.r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < (STRIDE == 1 ? 2 : 4); ++i) d.m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.ok) {
        d.c = item / QR;
        d.q = item - d.c * QR;
        const float* xr = xs + (size_t)d.c * Lin + (STRIDE == 1 ? 8 : 16) * d.q;
#pragma unroll
        for (int i = 0; i < (STRIDE == 1 ? 2 : 4); ++i) d.m[i] = ld_stream4(xr + 4 * i);
        d.hl = d.q > 0;
        if (d.hl) d.l = __ldg(reinterpret_cast<const float2*>(xr - 2));
        if (STRIDE == 1) {
            d.hr = d.q < QR - 1;
            if (d.hr) d.r = __ldg(reinterpret_cast<const float2*>(xr + 8));
        } else {
            d.hr = 16 * d.q + 16 < Lin;
            if (d.hr) d.r.x = __ldg(xr + 16);
        }
    }
    return d;
}

template <int STRIDE, bool ACT>
__global__ void __launch_bounds__(kDw8Threads)
dw5_wide_pipe_kernel(const float* __restrict__ x, NormIn nin,
                     const float* __restrict__ w5, const float* __restrict__ bias,
                     float* __restrict__ y, double* __restrict__ stats_out,
                     int C, int Lin, int Lout, int chunks_per_sample) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    const int QR = Lout >> 3;                 // runs per row
    const int items = C * QR;                 // per sample
    const float* xs = x + (size_t)sample * C * Lin;
    float* ys = y + (size_t)sample * C * Lout;
    // the first run's loads do not depend on the sample statistics: issue them before the barrier
    DwRun<STRIDE> cur = dw_load_run<STRIDE>(xs, (chunk * kDw8Items + 0) * kDw8Threads + threadIdx.x, items, QR, Lin);
    if (threadIdx.x == 0) s_norm = sample_norm(nin, sample);
    const float slope = ACT ? __ldg(nin.prelu) : 1.f;
    const bool sle1 = slope <= 1.f;
    __syncthreads();
    const SampleNorm sn = s_norm;
    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int it = 0; it < kDw8Items; ++it) {
        DwRun<STRIDE> nxt;
        if (it + 1 < kDw8Items)
            nxt = dw_load_run<STRIDE>(xs, (chunk * kDw8Items + it + 1) * kDw8Threads + threadIdx.x, items, QR, Lin);
        if (cur.ok) {
            const FoldedNorm f = fold_norm(nin, sn, cur.c);
            float w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = __ldg(w5 + cur.c * 5 + j);
            const float b = __ldg(bias + cur.c);
            float o[8];
            if (STRIDE == 1) {
                float v[12];
                v[0] = cur.hl ? norm_act<ACT>(cur.l.x, f, slope, sle1) : 0.f;
                v[1] = cur.hl ? norm_act<ACT>(cur.l.y, f, slope, sle1) : 0.f;
                const float mm[8] = {cur.m[0].x, cur.m[0].y, cur.m[0].z, cur.m[0].w, cur.m[1].x, cur.m[1].y, cur.m[1].z, cur.m[1].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) v[2 + i] = norm_act<ACT>(mm[i], f, slope, sle1);
                v[10] = cur.hr ? norm_act<ACT>(cur.r.x, f, slope, sle1) : 0.f;
                v[11] = cur.hr ? norm_act<ACT>(cur.r.y, f, slope, sle1) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[i + j], a);
                    o[i] = a;
                }
            } else {
                float v[19];
                v[0] = cur.hl ? norm_act<ACT>(cur.l.x, f, slope, sle1) : 0.f;
                v[1] = cur.hl ? norm_act<ACT>(cur.l.y, f, slope, sle1) : 0.f;
                const float mm[16] = {cur.m[0].x, cur.m[0].y, cur.m[0].z, cur.m[0].w, cur.m[1].x, cur.m[1].y, cur.m[1].z, cur.m[1].w,
                                      cur.m[2 % (STRIDE == 1 ? 2 : 4)].x, cur.m[2 % (STRIDE == 1 ? 2 : 4)].y,
                                      cur.m[2 % (STRIDE == 1 ? 2 : 4)].z, cur.m[2 % (STRIDE == 1 ? 2 : 4)].w,
                                      cur.m[3 % (STRIDE == 1 ? 2 : 4)].x, cur.m[3 % (STRIDE == 1 ? 2 : 4)].y,
                                      cur.m[3 % (STRIDE == 1 ? 2 : 4)].z, cur.m[3 % (STRIDE == 1 ? 2 : 4)].w};
#pragma unroll
                for (int i = 0; i < 16; ++i) v[2 + i] = norm_act<ACT>(mm[i], f, slope, sle1);
                v[18] = cur.hr ? norm_act<ACT>(cur.r.x, f, slope, sle1) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) a = fmaf(w[j], v[2 * i + j], a);
                    o[i] = a;
                }
            }
            float* yr = ys + (size_t)cur.c * Lout + 8 * cur.q;
            st_stream4(yr, make_float4(o[0], o[1], o[2], o[3]));
            st_stream4(yr + 4, make_float4(o[4], o[5], o[6], o[7]));
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
        }
        if (it + 1 < kDw8Items) cur = nxt;
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}
#endif  // SDR_DW_PIPELINE

// merge, 16 outputs per thread, coarse-to-fine: s_d[i] = z_d[i]*a_d + (b_d + s_{d+1}[i>>1])
// requires depth >= 4 and L % 16 == 0
constexpr int kMg16Threads = SDR_MG_THREADS;
constexpr int kMg16Items = SDR_MG_ITEMS;
__global__ void __launch_bounds__(kMg16Threads)
merge_wide_kernel(MergeArgs a, float* __restrict__ m, double* __restrict__ stats_out,
                  int C, int L, int chunks_per_sample) {
    __shared__ SampleNorm s_norm[kMaxDepth];
    __shared__ float s_red[64];
    const int sample = blockIdx.x / chunks_per_sample;
    const int chunk = blockIdx.x - sample * chunks_per_sample;
    if (threadIdx.x < a.depth) s_norm[threadIdx.x] = sample_norm(a.n[threadIdx.x], sample);
    __syncthreads();
    const int QR = L >> 4;
    const int items = C * QR;
    float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
    for (int iti = 0; iti < kMg16Items; ++iti) {
      const int item = (chunk * kMg16Items + iti) * kMg16Threads + threadIdx.x;
      if (item < items) {
        const int c = item / QR;
        const int q = item - c * QR;
        const size_t row = (size_t)sample * C + c;
        // issue every load of this run first
        const float* z0 = a.z[0] + row * L + 16 * q;
        const float4 v00 = ld_stream4(z0), v01 = ld_stream4(z0 + 4), v02 = ld_stream4(z0 + 8), v03 = ld_stream4(z0 + 12);
        const float* z1 = a.z[1] + row * (L >> 1) + 8 * q;
        const float4 v10 = ld_stream4(z1), v11 = ld_stream4(z1 + 4);
        const float4 v2 = ld_stream4(a.z[2] + row * (L >> 2) + 4 * q);
        const float2 v3 = __ldg(reinterpret_cast<const float2*>(a.z[3] + row * (L >> 3) + 2 * q));
        float base = 0.f;                      // levels >= 4 are constant over the run
        for (int d = 4; d < a.depth; ++d) {
            const FoldedNorm f = fold_norm(a.n[d], s_norm[d], c);
            base += fmaf(__ldg(a.z[d] + row * (L >> d) + (q >> (d - 4))), f.a, f.b);
        }
        const FoldedNorm f3 = fold_norm(a.n[3], s_norm[3], c);
        const FoldedNorm f2 = fold_norm(a.n[2], s_norm[2], c);
        const FoldedNorm f1 = fold_norm(a.n[1], s_norm[1], c);
        const FoldedNorm f0 = fold_norm(a.n[0], s_norm[0], c);
        float s3[2], s2[4], s1[8], o[16];
        const float c3 = f3.b + base;
        s3[0] = fmaf(v3.x, f3.a, c3); s3[1] = fmaf(v3.y, f3.a, c3);
        const float z2v[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s2[i] = fmaf(z2v[i], f2.a, f2.b + s3[i >> 1]);
        const float z1v[8] = {v10.x, v10.y, v10.z, v10.w, v11.x, v11.y, v11.z, v11.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) s1[i] = fmaf(z1v[i], f1.a, f1.b + s2[i >> 1]);
        const float z0v[16] = {v00.x, v00.y, v00.z, v00.w, v01.x, v01.y, v01.z, v01.w,
                               v02.x, v02.y, v02.z, v02.w, v03.x, v03.y, v03.z, v03.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = fmaf(z0v[i], f0.a, f0.b + s1[i >> 1]);
        float* mr = m + row * L + 16 * q;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st_stream4(mr + 4 * i, make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
      }
    }
    block_stats_atomic(acc_s, acc_q, stats_out, sample, s_red);
}

#if SDR_DW_CHAIN
// ---------------------------------------------------------------------------
// EXPERIMENT (round-2 candidate, compiled only with -DSDR_DW_CHAIN=1, not measured yet):
// all stride-2 depthwise levels of a U-ConvBlock in ONE persistent kernel.
//
// Levels 1..D-1 are 315 / 157 / 79 / 39 MB kernels (cfg 2); the small ones lose half of their time to launch,
// ramp and tail, and level d+1 of a sample only needs level d of THAT sample (its GlobLN statistics).  Work items
// (level, sample, chunk) are claimed in level-major order from an atomic counter by whatever CTA is free; an item of
// level d >= 2 first waits until the per-sample completion counter of level d-1 reaches that level's chunk count.
// Items are claimed in increasing order by running CTAs only, so every item another CTA waits for is owned by a
// resident CTA: no co-residency requirement, no deadlock.  Data and statistics produced inside this kernel are read
// with ld.global.cg (L2): .nc / L1-cached loads are only legal for data that is read-only for the whole kernel.
// ---------------------------------------------------------------------------
struct DwChainArgs {
    int first, last;                              // levels handled: first..last (first >= 1)
    const float* x[kMaxDepthApi];                 // x[d]: input of level d  (= z[d-1])
    float* y[kMaxDepthApi];                       // y[d]: output of level d (= z[d])
    const double* stats_in[kMaxDepthApi];         // statistics of x[d]
    double* stats_out[kMaxDepthApi];              // statistics of y[d]
    const float* gamma[kMaxDepthApi];             // GlobLN affine applied to x[d] on load
    const float* beta[kMaxDepthApi];
    const float* w5[kMaxDepthApi];
    const float* bias[kMaxDepthApi];
    int Lin[kMaxDepthApi], chunks[kMaxDepthApi], item_base[kMaxDepthApi + 1];
    int C, samples;
    int* work;                                    // [1] next item
    int* done;                                    // [kMaxDepthApi][samples] finished chunks per (level, sample)
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(kDw8Threads)
dw5_chain_kernel(const DwChainArgs a) {
    __shared__ SampleNorm s_norm;
    __shared__ float s_red[64];
    __shared__ int s_item;
    const int total = a.item_base[a.last + 1];
    while (true) {
        __syncthreads();                          // s_item / s_norm / s_red of the previous item are no longer in use
        if (threadIdx.x == 0) s_item = atomicAdd(a.work, 1);
        __syncthreads();
        const int item_id = s_item;
        if (item_id >= total) break;
        int d = a.first;
        while (item_id >= a.item_base[d + 1]) ++d;
        const int local = item_id - a.item_base[d];
        const int sample = local / a.chunks[d];
        const int chunk = local - sample * a.chunks[d];
        const int Lin = a.Lin[d], Lout = Lin >> 1, C = a.C;
        if (threadIdx.x == 0) {
            if (d > a.first) {                    // producer = level d-1 of this kernel: wait for its last chunk
                const int need = a.chunks[d - 1];
                const int* flag = a.done + (size_t)(d - 1) * a.samples + sample;
                int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                    if (seen < need) __nanosleep(200);
                } while (seen < need);
            }
            SampleNorm sn{0.f, 1.f};
            const double cnt = (double)C * Lin;
            const double sum = __ldcg(a.stats_in[d] + 2 * (size_t)sample);
            const double sq = __ldcg(a.stats_in[d] + 2 * (size_t)sample + 1);
            const double mu = sum / cnt;
            double var = sq / cnt - mu * mu;
            var = var < 0.0 ? 0.0 : var;
            sn.mean = (float)mu;
            sn.rstd = (float)(1.0 / sqrt(var + (double)kGlnEps));
            s_norm = sn;
        }
        __syncthreads();
        const SampleNorm sn = s_norm;
        const int QR = Lout >> 3;
        const int items = C * QR;
        const float* xs = a.x[d] + (size_t)sample * C * Lin;
        float* ys = a.y[d] + (size_t)sample * C * Lout;
        float acc_s = 0.f, acc_q = 0.f;
#pragma unroll
        for (int it = 0; it < kDw8Items; ++it) {
            const int item = (chunk * kDw8Items + it) * kDw8Threads + threadIdx.x;
            if (item < items) {
                const int c = item / QR;
                const int q = item - c * QR;
                FoldedNorm f;
                f.a = __ldg(a.gamma[d] + c) * sn.rstd;
                f.b = fmaf(-sn.mean, f.a, __ldg(a.beta[d] + c));
                float w[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) w[j] = __ldg(a.w5[d] + c * 5 + j);
                const float b = __ldg(a.bias[d] + c);
                const float* xr = xs + (size_t)c * Lin + 16 * q;      // window v[0..18] = positions 16q-2 .. 16q+16
                float v[19];
                const float4 m0 = ldcg4(xr), m1 = ldcg4(xr + 4), m2 = ldcg4(xr + 8), m3 = ldcg4(xr + 12);
                float2 l = make_float2(0.f, 0.f);
                float r = 0.f;
                const bool hl = q > 0, hr = 16 * q + 16 < Lin;
                if (hl) l = __ldcg(reinterpret_cast<const float2*>(xr - 2));
                if (hr) r = __ldcg(xr + 16);
                v[0] = hl ? fmaf(l.x, f.a, f.b) : 0.f;
                v[1] = hl ? fmaf(l.y, f.a, f.b) : 0.f;
                const float mm[16] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w,
                                      m2.x, m2.y, m2.z, m2.w, m3.x, m3.y, m3.z, m3.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) v[2 + i] = fmaf(mm[i], f.a, f.b);
                v[18] = hr ? fmaf(r, f.a, f.b) : 0.f;
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float acc = b;
#pragma unroll
                    for (int j = 0; j < 5; ++j) acc = fmaf(w[j], v[2 * i + j], acc);
                    o[i] = acc;
                }
                float* yr = ys + (size_t)c * Lout + 8 * q;
                *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(yr + 4) = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc_s += o[i]; acc_q = fmaf(o[i], o[i], acc_q); }
            }
        }
        block_stats_atomic(acc_s, acc_q, a.stats_out[d], sample, s_red);
        // publish: every thread's stores (and the statistics atomics) before the completion count
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && d < a.last) atomicAdd(a.done + (size_t)d * a.samples + sample, 1);
    }
}

// levels first..last (all stride 2, no activation on load), outputs of 8-aligned length; `counters` = 1 + kMaxDepthApi *
// samples zeroed ints.  Returns SDR_ERR_UNSUPPORTED when a level does not fit the wide kernel (caller falls back).
int launch_depthwise_chain(const float* const* x, float* const* y, const double* const* stats_in,
                           double* const* stats_out, const float* const* gamma, const float* const* beta,
                           const float* const* w5, const float* const* bias, const int* Lin,
                           int first, int last, int samples, int C, int* counters, cudaStream_t st) {
    if (first < 1 || last < first || last >= kMaxDepthApi || samples <= 0 || C <= 0 || !counters) return SDR_ERR_BAD_ARGUMENT;
    DwChainArgs a;
    a.first = first; a.last = last; a.C = C; a.samples = samples;
    a.work = counters; a.done = counters + 1;
    long long base = 0;
    for (int d = 0; d < kMaxDepthApi; ++d) { a.x[d] = nullptr; a.y[d] = nullptr; a.chunks[d] = 0; a.Lin[d] = 0; a.item_base[d] = 0; }
    for (int d = first; d <= last; ++d) {
        const int Lout = Lin[d] / 2;
        if (Lin[d] != 2 * Lout || (Lout % 8) != 0) return SDR_ERR_UNSUPPORTED;
        if ((reinterpret_cast<uintptr_t>(x[d]) | reinterpret_cast<uintptr_t>(y[d])) % 16) return SDR_ERR_UNSUPPORTED;
        a.x[d] = x[d]; a.y[d] = y[d]; a.stats_in[d] = stats_in[d]; a.stats_out[d] = stats_out[d];
        a.gamma[d] = gamma[d]; a.beta[d] = beta[d]; a.w5[d] = w5[d]; a.bias[d] = bias[d];
        a.Lin[d] = Lin[d];
        const long long items = (long long)C * (Lout / 8);
        const int per_cta = kDw8Threads * kDw8Items;
        a.chunks[d] = (int)((items + per_cta - 1) / per_cta);
        a.item_base[d] = (int)base;
        base += (long long)a.chunks[d] * samples;
        if (base > 0x3fffffffLL) return SDR_ERR_UNSUPPORTED;
    }
    a.item_base[last + 1] = (int)base;
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dw5_chain_kernel, kDw8Threads, 0) != cudaSuccess || occ < 1)
        return SDR_ERR_CUDA;
    long long grid = (long long)sms * occ;
    if (grid > base) grid = base;
    dw5_chain_kernel<<<(unsigned)grid, kDw8Threads, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}
#endif  // SDR_DW_CHAIN

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
int launch_depthwise(const float* x, const NormIn& nin, const float* w5, const float* bias,
                     float* y, double* stats_out, int samples, int C, int Lin, int stride,
                     cudaStream_t st) {
    if (samples <= 0 || C <= 0 || Lin <= 0 || (stride != 1 && stride != 2)) return SDR_ERR_BAD_ARGUMENT;
    const int Lout = (Lin + 4 - 5) / stride + 1;
    const bool vec = (Lout % 4 == 0) && (stride == 1 || Lin == 2 * Lout) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0);
    if (vec && (Lout % 8 == 0)) {
        const long long items = (long long)C * (Lout / 8);
        const int per_cta = kDw8Threads * kDw8Items;
        const int chunks = (int)((items + per_cta - 1) / per_cta);
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        const bool act = nin.prelu != nullptr;
#if SDR_DW_PIPELINE
#define SDR_DW(S, A) dw5_wide_pipe_kernel<S, A><<<(unsigned)grid, kDw8Threads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks)
#else
#define SDR_DW(S, A) dw5_wide_kernel<S, A><<<(unsigned)grid, kDw8Threads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks)
#endif
        if (stride == 1) { if (act) SDR_DW(1, true); else SDR_DW(1, false); }
        else             { if (act) SDR_DW(2, true); else SDR_DW(2, false); }
#undef SDR_DW
    } else if (vec) {
        const long long items = (long long)C * (Lout / 4);
        const int chunks = (int)((items + kDwThreads * kDwItems - 1) / (kDwThreads * kDwItems));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        if (stride == 1)
            dw5_vec_kernel<1><<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks);
        else
            dw5_vec_kernel<2><<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, chunks);
    } else {
        const long long items = (long long)C * Lout;
        const int chunks = (int)((items + kDwThreads * 4 - 1) / (kDwThreads * 4));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        dw5_scalar_kernel<<<(unsigned)grid, kDwThreads, 0, st>>>(x, nin, w5, bias, y, stats_out, C, Lin, Lout, stride, chunks);
    }
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

int launch_merge(const float* const* z, const NormIn* nins, int depth, float* m, double* stats_out,
                 int samples, int C, int L, cudaStream_t st) {
    if (depth < 1 || depth > kMaxDepth) return SDR_ERR_UNSUPPORTED;
    if (samples <= 0 || C <= 0 || L <= 0 || (L % (1 << (depth - 1))) != 0) return SDR_ERR_BAD_ARGUMENT;
    MergeArgs a;
    a.depth = depth;
    bool aligned = reinterpret_cast<uintptr_t>(m) % 16 == 0;
    for (int d = 0; d < kMaxDepth; ++d) {
        a.z[d] = d < depth ? z[d] : nullptr;
        a.n[d] = d < depth ? nins[d] : NormIn{nullptr, nullptr, nullptr, nullptr, 1.0};
        if (d < depth) aligned = aligned && reinterpret_cast<uintptr_t>(z[d]) % 16 == 0;
    }
    // vector path: rows of level 0 are float4-aligned, rows of level 1 float2-aligned
    const bool vec = aligned && (L % 4 == 0);
    if (vec && depth >= 4 && (L % 16 == 0)) {
        const long long items = (long long)C * (L / 16);
        const int per_cta = kMg16Threads * kMg16Items;
        const int chunks = (int)((items + per_cta - 1) / per_cta);
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        merge_wide_kernel<<<(unsigned)grid, kMg16Threads, 0, st>>>(a, m, stats_out, C, L, chunks);
    } else if (vec) {
        const long long items = (long long)C * (L / 4);
        const int chunks = (int)((items + kMgThreads * kMgItems - 1) / (kMgThreads * kMgItems));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        merge_vec_kernel<<<(unsigned)grid, kMgThreads, 0, st>>>(a, m, stats_out, C, L, chunks);
    } else {
        const long long items = (long long)C * L;
        const int chunks = (int)((items + kMgThreads * 4 - 1) / (kMgThreads * 4));
        const long long grid = (long long)chunks * samples;
        if (grid > 0x7fffffffLL) return SDR_ERR_UNSUPPORTED;
        merge_scalar_kernel<<<(unsigned)grid, kMgThreads, 0, st>>>(a, m, stats_out, C, L, chunks);
    }
    return cudaGetLastError() == cudaSuccess ? SDR_OK : SDR_ERR_CUDA;
}

}  // namespace sdr
