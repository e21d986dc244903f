// Non-GEMM kernels of the evaluator: bit-plane encoding, the fp32 check path, and the policy/value tails.
// Reference semantics restated on the device:
//   encode_*            planes_to_tensor                      engine/src/net/mod.rs:121-156
//   policy_tail_kernel  non-finite clamp + calc_moves_probs   engine/src/net/mod.rs:57-61, :106-119
//   legal derivation    HexPosition::legal_moves              engine/src/hex/core.rs:297-305
//   value_tail_kernel   Linear(128,1) + Tanh                  training/cattus_train/net_utils.py:73-74
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace cb2 {

// One input record per position in the (pinned and device) batch block:
//   [u32 offset of its probabilities in the compact output][u32 #legal]   <- 8-byte prefix, written by the host
//   [planes * wpp u64 words][legal bitmap padded to a multiple of 8 bytes (absent when derived)]
// Kernels receive a pointer to record 0's PLANES; rec_bytes (the stride) includes the prefix.
struct RecLayout {
    int rec_bytes;     // multiple of 8
    int planes;        // C_in
    int wpp;           // u64 words per plane
    int s;             // board size
    int moves;         // M
    int legal_off;     // byte offset of the bitmap inside the record, or -1 when legality is derived
    int legal_words;   // ceil(M / 32)
};

__device__ __forceinline__ uint32_t plane_bit(const uint64_t* planes, int wpp, int c, int cell) {
    return static_cast<uint32_t>((planes[c * wpp + (cell >> 6)] >> (cell & 63)) & 1ull);
}

// 32 legality bits [32*j, 32*j+32) of position `rec`.
__device__ __forceinline__ uint32_t legal_word(const uint8_t* rec, const RecLayout& L, int j) {
    uint32_t w;
    if (L.legal_off >= 0) {
        w = reinterpret_cast<const uint32_t*>(rec + L.legal_off)[j];
    } else {
        // empty cells: ones-plane & ~(plane0 | plane1)   (hex/core.rs:297-305; planes per hex/net.rs:14-24)
        const uint64_t* pl = reinterpret_cast<const uint64_t*>(rec);
        const int k = j >> 1;
        const uint64_t e = pl[2 * L.wpp + k] & ~(pl[k] | pl[L.wpp + k]);
        w = static_cast<uint32_t>(e >> (32 * (j & 1)));
    }
    const int rem = L.moves - 32 * j;
    if (rem < 32) w &= (rem <= 0) ? 0u : ((1u << rem) - 1u);
    return w;
}

// ---------------------------------------------------------------------------------------------- encode
// Dense f32 NCHW, exactly planes_to_tensor: one thread per output element, rows >= n zero-filled.
__global__ void encode_nchw_f32_kernel(const uint8_t* __restrict__ recs, RecLayout L, const uint32_t* __restrict__ n_ptr, int batch,
                                       float* __restrict__ out) {
    const int n = static_cast<int>(*n_ptr);
    const int s2 = L.s * L.s;
    const long long total = static_cast<long long>(batch) * L.planes * s2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cell = static_cast<int>(i % s2);
        const int c = static_cast<int>((i / s2) % L.planes);
        const int b = static_cast<int>(i / (static_cast<long long>(s2) * L.planes));
        float v = 0.0f;
        if (b < n) v = plane_bit(reinterpret_cast<const uint64_t*>(recs + static_cast<size_t>(b) * L.rec_bytes), L.wpp, c, cell) ? 1.0f : 0.0f;
        out[i] = v;
    }
}

// NHWC bf16 with the channel dimension padded to 64 (one 128-byte row per cell = one TMA / UMMA swizzle row).
// One thread writes 8 channels (16 B); consecutive threads write consecutive 16-byte chunks -> fully coalesced.
__global__ void encode_nhwc_bf16_kernel(const uint8_t* __restrict__ recs, RecLayout L, const uint32_t* __restrict__ n_ptr,
                                        int rows_total, __nv_bfloat16* __restrict__ out) {
    const int n = static_cast<int>(*n_ptr);
    const int s2 = L.s * L.s;
    const long long total = static_cast<long long>(rows_total) * 8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i & 7);
        const long long row = i >> 3;
        const int b = static_cast<int>(row / s2);
        const int cell = static_cast<int>(row - static_cast<long long>(b) * s2);
        uint32_t w[4] = {0, 0, 0, 0};
        if (b < n && chunk * 8 < L.planes) {
            const uint64_t* pl = reinterpret_cast<const uint64_t*>(recs + static_cast<size_t>(b) * L.rec_bytes);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = chunk * 8 + j;
                if (c < L.planes && plane_bit(pl, L.wpp, c, cell)) w[j >> 1] |= (j & 1) ? 0x3F800000u : 0x00003F80u;  // bf16 1.0
            }
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// run_dense entry for the bf16 path: user f32 NCHW -> NHWC bf16 [rows][64]
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, int n, int c_in, int s2, int rows_total,
                                             __nv_bfloat16* __restrict__ out) {
    const long long total = static_cast<long long>(rows_total) * 64;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i & 63);
        const long long row = i >> 6;
        const int b = static_cast<int>(row / s2);
        const int cell = static_cast<int>(row - static_cast<long long>(b) * s2);
        float v = 0.0f;
        if (b < n && c < c_in) v = in[(static_cast<long long>(b) * c_in + c) * s2 + cell];
        out[i] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------------------------------------- SimpleTwoHeadedModel
// nn.Flatten of the NCHW planes (net_utils.py:101,113): out[b][j] = bit(plane j / S^2, cell j % S^2), j < features; columns up
// to k_pad and rows >= n are zero.  One thread writes 8 columns (16 B).
__global__ void encode_flat_bf16_kernel(const uint8_t* __restrict__ recs, RecLayout L, const uint32_t* __restrict__ n_ptr, int batch,
                                        int features, int k_pad, int ld, __nv_bfloat16* __restrict__ out) {
    const int n = static_cast<int>(*n_ptr);
    const int s2 = L.s * L.s;
    const int chunks = k_pad / 8;
    const long long total = static_cast<long long>(batch) * chunks;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i % chunks);
        const int b = static_cast<int>(i / chunks);
        uint32_t w[4] = {0, 0, 0, 0};
        if (b < n) {
            const uint64_t* pl = reinterpret_cast<const uint64_t*>(recs + static_cast<size_t>(b) * L.rec_bytes);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int f = chunk * 8 + j;
                if (f < features && plane_bit(pl, L.wpp, f / s2, f % s2)) w[j >> 1] |= (j & 1) ? 0x3F800000u : 0x00003F80u;  // bf16 1.0
            }
        }
        *reinterpret_cast<uint4*>(out + static_cast<long long>(b) * ld + chunk * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// run_dense entry: user f32 NCHW (already the flatten order) -> bf16 rows [batch][ld], columns >= features zero
__global__ void nchw_f32_to_flat_bf16_kernel(const float* __restrict__ in, int batch, int features, int k_pad, int ld,
                                             __nv_bfloat16* __restrict__ out) {
    const int chunks = k_pad / 8;
    const long long total = static_cast<long long>(batch) * chunks;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i % chunks);
        const int b = static_cast<int>(i / chunks);
        __nv_bfloat16* o = out + static_cast<long long>(b) * ld + chunk * 8;
        for (int j = 0; j < 8; ++j) {
            const int f = chunk * 8 + j;
            o[j] = __float2bfloat16(f < features ? in[static_cast<long long>(b) * features + f] : 0.0f);
        }
    }
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// value = tanh(b2 + sum_j x[b][j] * w2[j]) over k columns (net_utils.py:118-119); one warp per position; rows >= n zero
template <class T>
__global__ void dot_tanh_kernel(const T* __restrict__ x, int ld, int k, const float* __restrict__ w2, float b2,
                                const uint32_t* __restrict__ n_ptr, int batch, float* __restrict__ values) {
    const int n = n_ptr ? static_cast<int>(*n_ptr) : batch;
    const int lane = threadIdx.x & 31;
    const int b = static_cast<int>((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
    if (b >= batch || b >= n) return;
    const T* row = x + static_cast<long long>(b) * ld;
    float acc = 0.0f;
    for (int j = lane; j < k; j += 32) acc = fmaf(to_f32(row[j]), w2[j], acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) values[b] = tanhf(acc + b2);
}

// ---------------------------------------------------------------------------------------------- fp32 check path
// Direct NCHW convolution, ksize 1 or 3, "same" zero padding, folded-BN bias, optional residual and ReLU.
__global__ void conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                const float* __restrict__ resid, float* __restrict__ out, int n, int ci, int co, int s,
                                int ksize, int relu) {
    const int s2 = s * s;
    const long long total = static_cast<long long>(n) * co * s2;
    const int r = ksize / 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cell = static_cast<int>(i % s2);
        const int o = static_cast<int>((i / s2) % co);
        const int b = static_cast<int>(i / (static_cast<long long>(s2) * co));
        const int h = cell / s, x = cell % s;
        float acc = 0.0f;
        const float* inb = in + static_cast<long long>(b) * ci * s2;
        const float* wo = w + static_cast<long long>(o) * ci * ksize * ksize;
        for (int c = 0; c < ci; ++c) {
            for (int dy = -r; dy <= r; ++dy) {
                const int hh = h + dy;
                if (hh < 0 || hh >= s) continue;
                for (int dx = -r; dx <= r; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= s) continue;
                    acc = fmaf(inb[c * s2 + hh * s + xx], wo[(c * ksize + (dy + r)) * ksize + (dx + r)], acc);
                }
            }
        }
        acc += bias[o];
        if (resid != nullptr) acc += resid[i];
        if (relu) acc = fmaxf(acc, 0.0f);
        out[i] = acc;
    }
}

// out[b][o] = bias[o] + sum_k in[b][k] * w[o][k]; one warp per output, shuffle reduction.
__global__ void fc_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                              float* __restrict__ out, int n, int k, int no, int ld_out, int relu) {
    const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = static_cast<long long>(n) * no;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long i = warp_global; i < total; i += nwarps) {
        const int o = static_cast<int>(i % no);
        const int b = static_cast<int>(i / no);
        const float* x = in + static_cast<long long>(b) * k;
        const float* ww = w + static_cast<long long>(o) * k;
        float acc = 0.0f;
        for (int j = lane; j < k; j += 32) acc = fmaf(x[j], ww[j], acc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
        if (lane == 0) {
            acc += bias[o];
            if (relu) acc = fmaxf(acc, 0.0f);
            out[static_cast<long long>(b) * ld_out + o] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------- tails
// One warp per position: clamp non-finite logits to -FLT_MAX, softmax over the legal moves only, write the
// probabilities compactly (ascending nn index) at the offset the host stored in the record's prefix (it counts the
// legal moves while packing, so the exclusive scan is free there).  Three passes over <= M logits held in L1/L2.
__global__ void policy_tail_kernel(const float* __restrict__ logits, int ld_logits, const uint8_t* __restrict__ recs,
                                   RecLayout L, const uint32_t* __restrict__ n_ptr, float* __restrict__ probs) {
    const int n = static_cast<int>(*n_ptr);
    const int lane = threadIdx.x & 31;
    const int b = static_cast<int>((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
    if (b >= n) return;
    const uint8_t* rec = recs + static_cast<size_t>(b) * L.rec_bytes;
    const float* row = logits + static_cast<long long>(b) * ld_logits;
    float mx = -FLT_MAX;
    for (int j = 0; j < L.legal_words; ++j) {
        const uint32_t w = legal_word(rec, L, j);
        if ((w >> lane) & 1u) {
            float x = row[32 * j + lane];
            if (!isfinite(x)) x = -FLT_MAX;
            mx = fmaxf(mx, x);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    float sum = 0.0f;
    for (int j = 0; j < L.legal_words; ++j) {
        const uint32_t w = legal_word(rec, L, j);
        if ((w >> lane) & 1u) {
            float x = row[32 * j + lane];
            if (!isfinite(x)) x = -FLT_MAX;
            sum += expf(x - mx);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    uint32_t pos = *reinterpret_cast<const uint32_t*>(rec - 8);
    for (int j = 0; j < L.legal_words; ++j) {
        const uint32_t w = legal_word(rec, L, j);
        if ((w >> lane) & 1u) {
            float x = row[32 * j + lane];
            if (!isfinite(x)) x = -FLT_MAX;
            probs[pos + __popc(w & ((1u << lane) - 1u))] = expf(x - mx) / sum;
        }
        pos += __popc(w);
    }
}

// Over the compact legal logits the policy FC epilogue wrote (tc_gemm.cuh, epi 3): probs = softmax(logits) per position
// (calc_moves_probs, engine/src/net/mod.rs:106-119).  One warp per position; offset and count come from the record
// prefix.  `logits` and `probs` use the same offsets and may be the same buffer (large batches: in place in HBM) or differ
// (small batches: logits in HBM, probabilities straight into the caller's zero-copy host block).
__global__ void softmax_compact_kernel(const uint8_t* __restrict__ recs, RecLayout L, const uint32_t* __restrict__ n_ptr,
                                       const float* logits, float* probs) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched beside the tail of the policy FC (ptx.cuh: grid_dep_wait)
    const int n = static_cast<int>(*n_ptr);
    const int lane = threadIdx.x & 31;
    const int b = static_cast<int>((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
    if (b >= n) return;
    const uint32_t* prefix = reinterpret_cast<const uint32_t*>(recs + static_cast<size_t>(b) * L.rec_bytes - 8);
    const uint32_t off = prefix[0], cnt = prefix[1];
    const float* in = logits + off;
    float* row = probs + off;
    float x[8];  // <= 256 legal moves per position (chess: <= 218)
    float mx = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t i = lane + 32u * k;
        x[k] = i < cnt ? in[i] : -FLT_MAX;
        mx = fmaxf(mx, x[k]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t i = lane + 32u * k;
        x[k] = i < cnt ? expf(x[k] - mx) : 0.0f;
        sum += x[k];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t i = lane + 32u * k;
        if (i < cnt) row[i] = x[k] / sum;
    }
}

// value = tanh(b2 + sum_j hidden[b][j] * w2[j]); one warp per position (hidden width fixed 128: net_utils.py:71-73).
__global__ void value_tail_kernel(const float* __restrict__ hidden, int ld_hidden, const float* __restrict__ w2, float b2,
                                  const uint32_t* __restrict__ n_ptr, float* __restrict__ values) {
    const int n = static_cast<int>(*n_ptr);
    const int lane = threadIdx.x & 31;
    const int b = static_cast<int>((blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5);
    if (b >= n) return;
    const float* hrow = hidden + static_cast<long long>(b) * ld_hidden;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = fmaf(hrow[lane + 32 * j], w2[lane + 32 * j], acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) values[b] = tanhf(acc + b2);
}

}  // namespace cb2
