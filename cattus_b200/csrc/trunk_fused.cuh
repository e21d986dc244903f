// trunk_fused: the whole conv trunk of ConvNetV1 (stem + R residual blocks, training/cattus_train/net_utils.py:60-65,
// :23-42) for 8x8 boards and 128 filters in ONE persistent kernel, including the board -> feature-plane encoding
// (engine/src/net/mod.rs:121-156) of its input.
//
// Why it can be one kernel: a 3x3 "same" convolution never mixes two boards, so a CTA that owns a set of boards can
// carry them through all 1 + 2R layers without ever talking to another CTA.  Activations stay in shared memory as
// bf16 for the whole trunk; only the packed bitboards come in (144 B per board) and the final activations go out.
// The only streamed operand is the weights (295 KB per layer, identical for every CTA -> L2 hits).
//
// Work unit: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns 8 boards per round: each CTA holds two 128-row
// tiles (2 boards each).  One tcgen05.mma covers M = 256 (one tile of each CTA), N = 128 output channels, K = 16
// input channels; each CTA stores only ITS 64 output channels of the weights (B operand is split across the pair),
// which halves both the weight bytes streamed per SM and the shared-memory operand reads per MMA.
//
// Shared-memory activation layout ("padded chunk planes", no swizzle, K-major UMMA core matrices):
//   for each 8-channel chunk c8 (16 B per cell) a plane of cells; cell(g, x) = (g + 2) * 9 + 1 + x where
//   g = 2 * y + j is the board row y of board j in {0,1} and x the file.  Every 8-cell segment is followed by one
//   zero cell and there are two zero segments above and below, so the input of tap (dy, dx) for output row
//   r = 8 g + x is simply cell(g, x) + 18 dy + dx: a 3x3 tap is the SAME descriptor with the start address moved
//   by (18 dy + dx) * 16 bytes.  8-row groups are 144 B apart (SBO), the two 8-channel halves of a K = 16 step are
//   one plane (162 cells, 2592 B) apart (LBO); consecutive planes share their zero margins.  Halo cells are zeroed
//   once and never written.  (Descriptor behaviour probed on hardware: tools/umma_probe.cu.)
//
// Pipeline per layer l, k-chunk kc (16 input channels), tile t:
//   producer (1 thread / CTA) : TMA of this CTA's half of the (l, kc) weight stage (9 taps x 2 KB) into a 3-slot ring;
//                               both CTAs' bytes are credited to the leader's "full" barrier
//   MMA issuer (1 thread, leader CTA): waits weights(l, kc) and "input chunk kc of tile t is written in both CTAs",
//                               issues 9 MMAs (one per tap) into accumulator (t, l & 1); commits free the ring slot
//                               (multicast to both CTAs) and, after the last k-chunk, publish the accumulator
//   epilogue (4 warps per tile): TMEM -> registers, + folded-BN bias (+ residual from smem) -> ReLU -> bf16 ->
//                               the other activation buffer, 16 channels at a time, signalling the leader's
//                               "input chunk ready" barrier after every 16 channels -- the k-chunk-outer MMA order
//                               lets layer l+1 start while layer l's epilogue is still draining.
// With two tiles per CTA the issuer alternates tiles, so one tile's epilogue hides behind the other tile's MMAs.
//
// The two 1x1 head convolutions (net_utils.py:69, :79) ride along as one extra "layer": a single weight stage holding
// all 8 k-chunks of the [VH + PH][128] matrix (N = VH + PH <= 64 split across the pair like every other layer), no
// taps, whose epilogue writes ReLU(bf16) straight into the FC kernels' A operands.  The 67 MB trunk output tensor
// (and the two launches that re-read it) of the unfused arrangement never exists.
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"

namespace cb2 {

constexpr int kFtThreads = 384;  // warp 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue
constexpr int kFtSegCells = 9;
constexpr int kFtPlaneCells = 162;                             // (16 + 2) segments
constexpr int kFtCell0 = 19;                                   // cell(g = 0, x = 0)
constexpr int kFtLbo = kFtPlaneCells * 16;                     // 2592
constexpr int kFtSbo = kFtSegCells * 16;                       // 144
constexpr int kFtBufBytes = 15 * kFtLbo + 181 * 16;            // 16 chunk planes, margins shared: 41776
constexpr int kFtActBytes = 4 * kFtBufBytes;                   // tile0.P, tile0.Q, tile1.P, tile1.Q
constexpr int kFtWOff = (kFtActBytes + 127) / 128 * 128;       // 167168
constexpr int kFtTapBytes = 2 * 64 * 16;                       // one tap of one k-chunk: [2 halves][64 oc][8 ic] bf16
constexpr int kFtWStage = 9 * kFtTapBytes;                     // 18432
constexpr int kFtWStages = 3;
constexpr int kFtBarOff = kFtWOff + kFtWStages * kFtWStage;    // 222464
constexpr int kFtNumBars = 3 + 3 + 16 + 4;                     // w_full, w_empty, act_full[2][8], acc_full[2][2]
constexpr int kFtSmemBytes = kFtBarOff + kFtNumBars * 8 + 16 + 128;
constexpr int kFtWRowsPerStage = kFtWStage / 256;              // weight image is a u8 [rows][256] tensor: 72 rows per stage

struct alignas(64) TrunkFusedParams {
    CUtensorMap tma_w;       // weight image, u8 [rows][256], box {256, 72}
    const uint8_t* recs;     // packed records (planes first)
    const uint32_t* n_ptr;   // number of valid positions
    const float* bias;       // [layers][128] folded-BN bias
    __nv_bfloat16* out_v;    // [boards * 64][vhp]  ReLU(value head conv)
    __nv_bfloat16* out_p;    // [boards * 64][php]  ReLU(policy head conv)
    uint32_t* err;
    int rec_bytes;
    int planes;      // C_in <= 32
    int layers;      // 1 + 2R
    int num_rounds;  // ceil(boards / (4 * tiles))
    int tiles;       // 128-row tiles per CTA in use: 2 (8 boards per pair and round) or, for small batches, 1 (4 boards: half the MMAs per round)
    int vhp, php;    // padded head widths (multiples of 16, vhp + php <= 64)
};

// first weight stage of layer l (stem has 2 k-chunks, every other layer 8)
__device__ __forceinline__ int ft_stage_base(int l) { return l == 0 ? 0 : 2 + (l - 1) * 8; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFtThreads, 1) trunk_fused_kernel(const __grid_constant__ TrunkFusedParams p) {
    extern __shared__ uint8_t ft_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    uint8_t* act = smem;
    uint8_t* wring = smem + kFtWOff;
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + kFtBarOff);
    uint64_t* w_empty = w_full + 3;
    uint64_t* act_full = w_empty + 3;   // [tile * 8 + kc]
    uint64_t* acc_full = act_full + 16;  // [tile * 2 + parity]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 4);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int n_valid = static_cast<int>(*p.n_ptr);
    // Rounds that hold at least one valid position: the launch is sized for the batch bucket, but a round of nothing but
    // padding is skipped (every thread of both CTAs derives the same bound, so the barrier phases stay in step).
    const int rounds = min(p.num_rounds, (n_valid + 4 * p.tiles - 1) / (4 * p.tiles));

    // ---- one-time setup: zero the activation planes (halo cells stay zero for the whole kernel), barriers, TMEM
    for (int i = threadIdx.x; i < kFtActBytes / 16; i += kFtThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&p.tma_w);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 3; ++s) {
            ptx::mbar_init(&w_full[s], 1);
            ptx::mbar_init(&w_empty[s], 1);
        }
        for (int i = 0; i < 16; ++i) ptx::mbar_init(&act_full[i], 8);  // 4 epilogue warps x 2 CTAs
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&acc_full[i], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_pair(tmem_ptr, 512);
        ptx::tmem_relinquish_pair();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================================================================== weight producer (both CTAs)
        if (lane == 0) {
            uint32_t it = 0;
            for (int rd = pair; rd < rounds; rd += num_pairs) {
                for (int l = 0; l <= p.layers; ++l) {             // l == layers: the head convs, one stage for all k-chunks
                    const int nkc = l == 0 ? 2 : (l == p.layers ? 1 : 8);
                    for (int kc = 0; kc < nkc; ++kc, ++it) {
                        const uint32_t slot = it % 3, ph = (it / 3) & 1;
                        ptx::mbar_wait(&w_empty[slot], ph ^ 1, p.err, 0x1100 + slot);
                        if (rank == 0) ptx::mbar_arrive_expect_tx(&w_full[slot], 2 * kFtWStage);
                        const int block = (ft_stage_base(l) + kc) * 2 + static_cast<int>(rank);
                        ptx::tma_load_2d_pair(wring + slot * kFtWStage, &p.tma_w, ptx::mapa(ptx::smem_u32(&w_full[slot]), 0), 0,
                                              block * kFtWRowsPerStage);
                    }
                }
            }
            // tail: do not leave while the leader's commits may still arrive on our barriers
            for (uint32_t j = it; j < it + 3; ++j) ptx::mbar_wait(&w_empty[j % 3], ((j / 3) & 1) ^ 1, p.err, 0x1200 + (j % 3));
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================== MMA issuer (leader CTA only)
        // The WHOLE warp runs this loop (warp-uniform control flow and addresses, so the compiler keeps descriptors in
        // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.  An earlier
        // version that ran the loop inside `if (lane == 0)` spent ~25 SASS instructions per MMA converting registers
        // to uniform registers and was issue-bound at half the tensor-pipe rate (profiles/r01c).
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(256, 128);
            const uint64_t a_hi64 = ptx::umma_desc_none_hi(kFtLbo, kFtSbo);
            const uint64_t b_hi64 = ptx::umma_desc_none_hi(64 * 16, 128);
            const uint32_t a_hi = static_cast<uint32_t>(a_hi64 >> 32), a_lo_fixed = static_cast<uint32_t>(a_hi64);
            const uint32_t b_hi = static_cast<uint32_t>(b_hi64 >> 32), b_lo_fixed = static_cast<uint32_t>(b_hi64);
            const uint32_t act_addr = ptx::smem_u32(act);
            const uint32_t w_addr = ptx::smem_u32(wring);
            const bool leader_lane = ptx::elect_one();
            uint32_t act_par = 0;
            uint32_t it = 0;
            const int nh = p.vhp + p.php;
            const uint32_t idesc_head = ptx::umma_idesc_bf16(256, static_cast<uint32_t>(nh));
            const uint64_t bh_hi64 = ptx::umma_desc_none_hi(static_cast<uint32_t>(nh / 2) * 16, 128);
            const uint32_t bh_hi = static_cast<uint32_t>(bh_hi64 >> 32), bh_lo_fixed = static_cast<uint32_t>(bh_hi64);
            for (int rd = pair; rd < rounds; rd += num_pairs) {
                for (int l = 0; l < p.layers; ++l) {
                    const int nkc = l == 0 ? 2 : 8;
                    const int in_buf = (l & 1) ? 1 : 0;  // stem and conv2 read P (0), conv1 reads Q (1)
                    for (int kc = 0; kc < nkc; ++kc, ++it) {
                        const uint32_t slot = it % 3, ph = (it / 3) & 1;
                        ptx::mbar_wait(&w_full[slot], ph, p.err, 0x2100 + slot);
                        const uint32_t b_lo0 = b_lo_fixed | ((w_addr + slot * kFtWStage) >> 4);
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (t >= p.tiles) break;
                            const int bi = t * 8 + kc;
                            ptx::mbar_wait(&act_full[bi], (act_par >> bi) & 1u, p.err, 0x2200 + bi);
                            act_par ^= 1u << bi;
                            ptx::tc_fence_after();
                            const uint32_t a_lo0 = a_lo_fixed | ((act_addr + (t * 2 + in_buf) * kFtBufBytes + (2 * kc) * kFtLbo + kFtCell0 * 16) >> 4);
                            const uint32_t d = tmem_base + static_cast<uint32_t>((t * 2 + (l & 1)) * 128);
                            if (leader_lane) {
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const int shift16 = 18 * (tap / 3 - 1) + (tap % 3 - 1);  // in 16-byte units
                                    ptx::umma_bf16_ss_pair_lohi(d, a_lo0 + shift16, a_hi, b_lo0 + tap * (kFtTapBytes / 16), b_hi, idesc,
                                                                (kc | tap) != 0);
                                }
                                if (kc == nkc - 1) ptx::umma_commit_pair_multicast(&acc_full[t * 2 + (l & 1)], 3);
                            }
                        }
                        if (leader_lane) ptx::umma_commit_pair_multicast(&w_empty[slot], 3);
                        __syncwarp();
                    }
                }
                // ---- head convs: 8 k-chunks, no taps, N = vhp + php, input = Q (the last conv2 wrote it), accumulator parity 1
                {
                    const uint32_t slot = it % 3, ph = (it / 3) & 1;
                    ++it;
                    ptx::mbar_wait(&w_full[slot], ph, p.err, 0x2100 + slot);
                    const uint32_t b_lo0 = bh_lo_fixed | ((w_addr + slot * kFtWStage) >> 4);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (t >= p.tiles) break;
                        const uint32_t d = tmem_base + static_cast<uint32_t>((t * 2 + 1) * 128);
                        for (int kc = 0; kc < 8; ++kc) {
                            const int bi = t * 8 + kc;
                            ptx::mbar_wait(&act_full[bi], (act_par >> bi) & 1u, p.err, 0x2300 + bi);
                            act_par ^= 1u << bi;
                            ptx::tc_fence_after();
                            const uint32_t a_lo = a_lo_fixed | ((act_addr + (t * 2 + 1) * kFtBufBytes + (2 * kc) * kFtLbo + kFtCell0 * 16) >> 4);
                            if (leader_lane) ptx::umma_bf16_ss_pair_lohi(d, a_lo, a_hi, b_lo0 + kc * (nh / 2) * 2, bh_hi, idesc_head, kc != 0);
                        }
                        if (leader_lane) ptx::umma_commit_pair_multicast(&acc_full[t * 2 + 1], 3);
                    }
                    if (leader_lane) ptx::umma_commit_pair_multicast(&w_empty[slot], 3);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4 && static_cast<int>(warp - 4) >> 2 < p.tiles) {
        // ================================================================== encode + epilogue (4 warps per tile)
        const int t = static_cast<int>(warp - 4) >> 2;
        const uint32_t q = warp & 3;  // TMEM lane quarter this warp may access
        const int r = static_cast<int>(q * 32 + lane);
        const int g = r >> 3, x = r & 7;
        const int j = g & 1, y = g >> 1;
        const int cell = y * 8 + x;
        const int pcell = kFtCell0 + g * kFtSegCells + x;
        uint8_t* bufP = act + (t * 2 + 0) * kFtBufBytes + pcell * 16;
        uint8_t* bufQ = act + (t * 2 + 1) * kFtBufBytes + pcell * 16;
        const uint32_t leader_act = ptx::mapa(ptx::smem_u32(&act_full[t * 8]), 0);
        const uint32_t tmem_row = tmem_base + ((q * 32u) << 16) + static_cast<uint32_t>(t * 2 * 128);
        uint32_t acc_par = 0;
        for (int rd = pair; rd < rounds; rd += num_pairs) {
            const int board = ((rd * p.tiles + t) * 2 + static_cast<int>(rank)) * 2 + j;  // tile-major: a 1-tile round is boards 4 rd .. 4 rd + 3
            const bool valid = board < n_valid;
            // ---- planes_to_tensor for my cell: channels 0..31 of the stem input (planes >= C_in are zero)
            {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = 0;
                if (valid) {
                    const uint64_t* pl = reinterpret_cast<const uint64_t*>(p.recs + static_cast<size_t>(board) * p.rec_bytes);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (c < p.planes && ((__ldg(pl + c) >> cell) & 1ull)) w[c >> 1] |= (c & 1) ? 0x3F800000u : 0x00003F80u;  // bf16 1.0
                    }
                }
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                    *reinterpret_cast<uint4*>(bufP + c8 * kFtLbo) = make_uint4(w[4 * c8], w[4 * c8 + 1], w[4 * c8 + 2], w[4 * c8 + 3]);
                ptx::tc_fence_before();  // orders my earlier tcgen05.ld of accumulator (t, 0) before the stem MMAs that overwrite it
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive_cluster(leader_act + 0 * 8);
                    ptx::mbar_arrive_cluster(leader_act + 1 * 8);
                }
            }
            for (int l = 0; l < p.layers; ++l) {
                const int par = l & 1;
                const bool has_resid = l >= 2 && par == 0;  // conv2 of a block: + block input (Q), result back into Q
                uint8_t* outb = par ? bufP : bufQ;          // stem -> Q, conv1 -> P, conv2 -> Q
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias + l * 128);
                float4 bn[4];  // bias of the chunk about to be processed, fetched one chunk ahead
#pragma unroll
                for (int i = 0; i < 4; ++i) bn[i] = __ldg(bias4 + i);
                ptx::mbar_wait(&acc_full[t * 2 + par], (acc_par >> par) & 1u, p.err, 0x3100 + t * 2 + par);
                acc_par ^= 1u << par;
                ptx::tc_fence_after();
#pragma unroll 1
                for (int kc = 0; kc < 8; ++kc) {
                    uint32_t raw[16];
                    ptx::tmem_ld_x16_issue(tmem_row + static_cast<uint32_t>(par * 128 + kc * 16), raw);
                    float4 bc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) bc[i] = bn[i];
                    if (kc < 7) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) bn[i] = __ldg(bias4 + (kc + 1) * 4 + i);
                    }
                    ptx::tmem_ld_wait(raw);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[4 * i + 0] = __uint_as_float(raw[4 * i + 0]) + bc[i].x;
                        v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + bc[i].y;
                        v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + bc[i].z;
                        v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + bc[i].w;
                    }
                    if (has_resid) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint4 rr = *reinterpret_cast<const uint4*>(bufQ + (2 * kc + h) * kFtLbo);
                            const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
                                v[8 * h + 2 * i + 0] += __bfloat162float(b2.x);
                                v[8 * h + 2 * i + 1] += __bfloat162float(b2.y);
                            }
                        }
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f));
                        o[i] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(outb + (2 * kc) * kFtLbo) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(outb + (2 * kc + 1) * kFtLbo) = make_uint4(o[4], o[5], o[6], o[7]);
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(leader_act + kc * 8);
                }
            }
            // ---- both 1x1 head convolutions: accumulator (t, parity 1), columns [0, vhp) value, [vhp, vhp + php) policy
            {
                const float4* hb4 = reinterpret_cast<const float4*>(p.bias + p.layers * 128);
                ptx::mbar_wait(&acc_full[t * 2 + 1], (acc_par >> 1) & 1u, p.err, 0x3200 + t);
                acc_par ^= 1u << 1;
                ptx::tc_fence_after();
                const size_t row = static_cast<size_t>(board) * 64 + cell;
                const int nh = p.vhp + p.php;
                for (int cb = 0; cb < nh; cb += 16) {
                    uint32_t raw[16];
                    ptx::tmem_ld_x16_issue(tmem_row + static_cast<uint32_t>(128 + cb), raw);
                    float4 bc[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) bc[i] = __ldg(hb4 + cb / 4 + i);
                    ptx::tmem_ld_wait(raw);
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __nv_bfloat162 a = __floats2bfloat162_rn(fmaxf(__uint_as_float(raw[4 * i + 0]) + bc[i].x, 0.0f),
                                                                       fmaxf(__uint_as_float(raw[4 * i + 1]) + bc[i].y, 0.0f));
                        const __nv_bfloat162 b = __floats2bfloat162_rn(fmaxf(__uint_as_float(raw[4 * i + 2]) + bc[i].z, 0.0f),
                                                                       fmaxf(__uint_as_float(raw[4 * i + 3]) + bc[i].w, 0.0f));
                        o[2 * i] = *reinterpret_cast<const uint32_t*>(&a);
                        o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b);
                    }
                    if (valid) {
                        uint4* og = cb < p.vhp ? reinterpret_cast<uint4*>(p.out_v + row * p.vhp + cb) : reinterpret_cast<uint4*>(p.out_p + row * p.php + (cb - p.vhp));
                        og[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        og[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                ptx::tc_fence_before();
            }
        }
    }

    // ---- teardown: nobody leaves (or frees TMEM) before both CTAs are done
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace cb2
