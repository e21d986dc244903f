// trunk_small: the whole conv trunk of a 16-filter ConvNetV1 (stem + R residual blocks, training/cattus_train/
// net_utils.py:60-65, :23-42) PLUS the two 1x1 head convolutions (net_utils.py:69, :79) and the board -> feature-plane
// encoding of its input (engine/src/net/mod.rs:121-156) in ONE launch, for any board size 3..11 (tic-tac-toe, hex4..11,
// chess_dev).  This is every shipped training configuration (residual_filter_num: 16); the 128-filter chess net has its
// own kernel (trunk_fused.cuh).
//
// With 16 filters all weights of the net's convolutions fit in shared memory (1 + 2R layers x 9 taps x 512 B = 69 KB
// for R = 7), so nothing streams: weights are copied in once per CTA, activations live in shared memory as bf16 for the
// whole trunk, packed bitboards come in and the two head-conv outputs go out.
//
// Geometry.  A CTA owns `boards_per_round` boards per round, laid out in shared memory as ONE 1-D strip of 16-byte
// cells (8 channels each) per 8-channel plane: every board row of S cells is followed by one zero cell, every board by
// one zero row, so with row pitch W = S + 1 and board pitch BP = W * W the cell of (board b, y, x) is
// q = b * BP + y * W + x, and the input of tap (dy, dx) for output cell q is simply cell q + dy * W + dx -- all
// "same"-padding zeros are real zero cells of the strip.  The strip is cut into T tiles of 128 consecutive cells; one
// tcgen05.mma covers M = 128 cells x N = 16 output channels x K = 16 input channels, its A operand being the strip
// itself (K-major, no swizzle: 8 consecutive cells = one core-matrix row group, SBO = 128 B, LBO = plane size) with the
// start address moved by (dy * W + dx) * 16 bytes.  No im2col, no per-tap copies.  Pad cells are computed too (and
// forced back to zero by the epilogue): useful fraction S^2 / (S + 1)^2.
//
// Tiles of one strip are not independent (a tile's halo lies in its neighbours): MMA(l + 1, t) needs the epilogues of
// (l, t - 1), (l, t), (l, t + 1).  To keep the tensor core busy across that dependency the T tiles are split into
// GROUPS of `group_tiles` tiles that hold whole boards each (no board straddles a group, and the cells either side of a
// group boundary are always-zero pad cells), so groups are independent: the issuer walks layer l over group 0, 1, ...
// and by the time it returns to group 0 for layer l + 1 that group's epilogue has long finished.  TMEM: 64 columns per
// tile = [layer parity 2][even-tap / odd-tap accumulator 2][16 channels]; the head convolution (32 columns) reuses the
// tile's parity-0 half, which its own input wait proves free.
//
// Warps: 0..15 encode + epilogue (TMEM lane quarter = warp & 3, tiles t = (warp >> 2) mod 4); warps 16 and 17 issue the
// MMAs of the even / odd tile groups (whole warp runs the loop so descriptors stay in uniform registers, one elected
// lane issues): a `tcgen05.mma` issue blocks for about the MMA's own duration, so a single issuer leaves the tensor
// core idle during its per-group barrier wait + commit (≈480 of 1680 cycles, profiles/r01g); two issuers alternate.
// Per-cell geometry (local board, cell index, pad or not) and all biases are tabulated in shared memory once.
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"

// Build with -DCB2_TRUNK_TRACE to compile the clock64 trace points in (tools/trace_trunk.py); they cost ~10 % so the
// production build leaves them out.
#ifdef CB2_TRUNK_TRACE
#define CB2_TS_TRACE(...) __VA_ARGS__
#else
#define CB2_TS_TRACE(...)
#endif

namespace cb2 {

constexpr int kTsThreads = 576;
constexpr int kTsIssuerWarp = 16;  // warps 16 and 17 issue MMAs (even / odd tile groups)
constexpr int kTsMaxTiles = 8;
constexpr int kTsTapBytes = 2 * 16 * 16;  // [k-half 2][oc 16][ic 8] bf16

struct TrunkSmallParams {
    const uint8_t* recs;      // packed records (planes first)
    const uint32_t* n_ptr;    // number of valid positions
    const uint4* wimg;        // weight image, see Engine::upload_weights
    const float* bias;        // [layers][16] folded-BN bias, then [nh] head bias (value channels first)
    __nv_bfloat16* out_v;     // [positions * S^2][vhp]  ReLU(value head conv)
    __nv_bfloat16* out_p;     // [positions * S^2][php]  ReLU(policy head conv)
    uint32_t* err;
    unsigned long long* dbg;  // optional [64 layers][8] clock64 trace of CTA 0 (CATTUS_B200_TRACE_TRUNK=1), else nullptr
    int rec_bytes, planes, wpp, s;
    int layers;   // 1 + 2R
    int stem_kc;  // 16-channel k-chunks of the stem input: ceil(C_in / 16) (1 or 2)
    int vhp, php; // padded head widths, vhp + php <= 32
    int tiles;    // T in {1, 2, 4, 8}
    int group_tiles;       // tiles per independent group (divides T)
    int boards_per_group;  // group_tiles * 128 / BP
    int boards_per_round;  // boards_per_group * (T / group_tiles)
    int num_rounds;
    int w_bytes;  // bytes of wimg (multiple of 16)
    int margin;   // zero cells before / after the strip, >= S + 2, multiple of 8
};

// byte offsets inside dynamic shared memory
struct TsSmemLayout {
    int plane_bytes, p_off, q_off, w_off, bias_off, info_off, bar_off, total;
};
__host__ __device__ inline TsSmemLayout ts_smem_layout(int tiles, int stem_kc, int margin, int w_bytes) {
    TsSmemLayout L;
    L.plane_bytes = (tiles * 128 + 2 * margin) * 16;
    const int p_planes = 2 * (stem_kc > 1 ? stem_kc : 1);
    L.p_off = 0;
    L.q_off = L.p_off + p_planes * L.plane_bytes;
    L.w_off = (L.q_off + 2 * L.plane_bytes + 127) / 128 * 128;
    L.bias_off = (L.w_off + w_bytes + 15) / 16 * 16;
    L.info_off = L.bias_off + (64 * 16 + 32) * 4;  // up to 64 layers of 16 + 32 head biases
    L.bar_off = (L.info_off + tiles * 128 * 2 + 15) / 16 * 16;
    L.total = L.bar_off + (2 * kTsMaxTiles) * 8 + 16 + 128;
    return L;
}
constexpr uint16_t kTsPad = 0xFFFF;  // info[] entry of a pad cell; otherwise (local board << 7) | cell

__global__ void __launch_bounds__(kTsThreads, 1) trunk_small_kernel(const __grid_constant__ TrunkSmallParams p) {
    extern __shared__ uint8_t ts_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ts_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    const TsSmemLayout L = ts_smem_layout(p.tiles, p.stem_kc, p.margin, p.w_bytes);
    uint8_t* bufP = smem + L.p_off;
    uint8_t* bufQ = smem + L.q_off;
    uint8_t* wsm = smem + L.w_off;
    float* bias_sm = reinterpret_cast<float*>(smem + L.bias_off);
    uint16_t* info = reinterpret_cast<uint16_t*>(smem + L.info_off);
    uint64_t* ready = reinterpret_cast<uint64_t*>(smem + L.bar_off);  // [group]: inputs of the next layer are written
    uint64_t* acc_full = ready + kTsMaxTiles;                          // [tile]: accumulator of the current layer is complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + kTsMaxTiles);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int T = p.tiles;
    const int W = p.s + 1;
    const int BP = W * W;
    const int n_valid = static_cast<int>(*p.n_ptr);
    // rounds holding at least one valid position (the launch is sized for the batch bucket; all-padding rounds are skipped)
    const int rounds = min(p.num_rounds, (n_valid + p.boards_per_round - 1) / p.boards_per_round);
    const int nh = p.vhp + p.php;
    const uint32_t tmem_cols = static_cast<uint32_t>(64 * T);

    // ---- one-time setup: zero both activation buffers (margins and pads stay zero), weights -> smem, barriers, TMEM
    for (int i = threadIdx.x; i < L.w_off / 16; i += kTsThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < p.w_bytes / 16; i += kTsThreads) reinterpret_cast<uint4*>(wsm)[i] = __ldg(p.wimg + i);
    for (int i = threadIdx.x; i < p.layers * 16 + nh; i += kTsThreads) bias_sm[i] = __ldg(p.bias + i);
    for (int q = threadIdx.x; q < T * 128; q += kTsThreads) {
        const int g = q / (p.group_tiles * 128), ql = q - g * p.group_tiles * 128;
        const int bg = ql / BP, rem = ql - bg * BP;
        const int y = rem / W, x = rem - y * W;
        info[q] = (bg < p.boards_per_group && y < p.s && x < p.s) ? static_cast<uint16_t>(((g * p.boards_per_group + bg) << 7) | (y * p.s + x)) : kTsPad;
    }
    if (warp == 0 && lane == 0) {
        for (int t = 0; t < kTsMaxTiles; ++t) {
            ptx::mbar_init(&ready[t], 4 * p.group_tiles);  // [group]: the four quarter-warps of each of its tiles
            ptx::mbar_init(&acc_full[t], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kTsIssuerWarp) {
        ptx::tmem_alloc(tmem_ptr, tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= kTsIssuerWarp) {
        // ================================================================== MMA issuers (warp 16: even groups, 17: odd)
        const int issuer = static_cast<int>(warp) - kTsIssuerWarp;
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 16);
        const uint32_t idesc_head = ptx::umma_idesc_bf16(128, static_cast<uint32_t>(nh));
        const uint64_t a_hi64 = ptx::umma_desc_none_hi(static_cast<uint32_t>(L.plane_bytes), 128);
        const uint64_t b_hi64 = ptx::umma_desc_none_hi(16 * 16, 128);
        const uint64_t bh_hi64 = ptx::umma_desc_none_hi(static_cast<uint32_t>(nh) * 16, 128);
        const uint32_t a_hi = static_cast<uint32_t>(a_hi64 >> 32), a_lo_fixed = static_cast<uint32_t>(a_hi64);
        const uint32_t b_hi = static_cast<uint32_t>(b_hi64 >> 32), b_lo_fixed = static_cast<uint32_t>(b_hi64);
        const uint32_t bh_hi = static_cast<uint32_t>(bh_hi64 >> 32), bh_lo_fixed = static_cast<uint32_t>(bh_hi64);
        const uint32_t p_addr = ptx::smem_u32(bufP) + static_cast<uint32_t>(p.margin) * 16;
        const uint32_t q_addr = ptx::smem_u32(bufQ) + static_cast<uint32_t>(p.margin) * 16;
        const uint32_t w_addr = ptx::smem_u32(wsm);
        const bool leader_lane = ptx::elect_one();
        for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
            uint32_t w_off = 0;
            for (int l = 0; l <= p.layers; ++l) {
                const bool head = l == p.layers;
                const int nkc = l == 0 ? p.stem_kc : 1;
                const uint32_t in_addr = (l & 1) ? q_addr : p_addr;  // stem and conv2 read P, conv1 and the heads read Q
                const uint32_t par = static_cast<uint32_t>(l & 1);   // (layers + 1) is even: stage parity = l & 1 in every round
                for (int t0 = issuer * p.group_tiles; t0 < T; t0 += 2 * p.group_tiles) {
                    // inputs: the epilogues (or the encode) of every tile of this group
                    CB2_TS_TRACE(const bool trace = p.dbg != nullptr && blockIdx.x == 0 && rd == 0 && t0 == 0 && lane == 0;)
                    CB2_TS_TRACE(const bool trace2 = p.dbg != nullptr && blockIdx.x == 0 && rd == 0 && lane == 0 && l < 8;)
                    CB2_TS_TRACE(if (trace2) p.dbg[128 + (l * 8 + t0) * 3 + 0] = clock64();)
                    ptx::mbar_wait(&ready[t0 / p.group_tiles], par, p.err, 0x5100 + t0);
                    ptx::tc_fence_after();
                    CB2_TS_TRACE(if (trace2) p.dbg[128 + (l * 8 + t0) * 3 + 1] = clock64();)
                    CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 0] = clock64();)
                    const uint32_t a_lo0 = a_lo_fixed | ((in_addr + static_cast<uint32_t>(t0) * 2048u) >> 4);
                    const uint32_t d0 = tmem_base + static_cast<uint32_t>(t0 * 64);
                    if (leader_lane) {
                        if (head) {
                            for (int t = 0; t < p.group_tiles; ++t)
                                ptx::umma_bf16_ss_lohi(d0 + t * 64, a_lo0 + t * 128, a_hi, bh_lo_fixed | ((w_addr + w_off) >> 4), bh_hi, idesc_head, 0u);
                        } else {
                            // Tap-major across the group's tiles and two accumulators per tile (even / odd taps, summed by
                            // the epilogue): consecutive MMAs into ONE accumulator are ~100 cycles apart however small
                            // they are (profiles/r01f), so the issuer keeps 2 x group_tiles independent chains in flight.
                            for (int kc = 0; kc < nkc; ++kc) {
                                const uint32_t a_kc = a_lo0 + static_cast<uint32_t>(kc * 2 * (L.plane_bytes >> 4));
                                const uint32_t b_lo0 = b_lo_fixed | ((w_addr + w_off + static_cast<uint32_t>(kc * 9 * kTsTapBytes)) >> 4);
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const int shift16 = (tap / 3 - 1) * W + (tap % 3 - 1);  // in 16-byte cells
                                    const uint32_t which = static_cast<uint32_t>((kc * 9 + tap) & 1);
                                    const uint32_t acc = (kc * 9 + tap) >= 2 ? 1u : 0u;
                                    for (int t = 0; t < p.group_tiles; ++t)
                                        ptx::umma_bf16_ss_lohi(d0 + t * 64 + par * 32 + which * 16, a_kc + shift16 + t * 128, a_hi, b_lo0 + tap * (kTsTapBytes / 16),
                                                               b_hi, idesc, acc);
                                }
                            }
                        }
                        for (int t = t0; t < t0 + p.group_tiles; ++t) ptx::umma_commit(&acc_full[t]);
                    }
                    __syncwarp();
                    CB2_TS_TRACE(if (trace2) p.dbg[128 + (l * 8 + t0) * 3 + 2] = clock64();)
                    CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 1] = clock64();)
                }
                w_off += head ? 0u : static_cast<uint32_t>(nkc * 9 * kTsTapBytes);
            }
        }
    } else {
        // ================================================================== encode + epilogue
        const uint32_t q4 = warp & 3;  // TMEM lane quarter this warp may access
        const int slot = static_cast<int>(warp >> 2);  // tiles slot, slot + 4
        const int r = static_cast<int>(q4 * 32 + lane);
        const int s2 = p.s * p.s;
        for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
            const int board0 = rd * p.boards_per_round;
            // ---- planes_to_tensor into the stem input (buffer P), zeros on pads / beyond the batch
            for (int t = slot; t < T; t += 4) {
                const int q = t * 128 + r;
                const uint32_t inf = info[q];
                const int board = board0 + static_cast<int>(inf >> 7);
                const bool valid = inf != kTsPad && board < n_valid;
                uint8_t* dst = bufP + (p.margin + q) * 16;
                const uint64_t* pl = reinterpret_cast<const uint64_t*>(p.recs + static_cast<size_t>(valid ? board : 0) * p.rec_bytes);
                const int cell = static_cast<int>(inf & 127u);
                for (int c8 = 0; c8 < 2 * p.stem_kc; ++c8) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c = c8 * 8 + j;
                            if (c < p.planes && plane_bit(pl, p.wpp, c, cell)) w[j >> 1] |= (j & 1) ? 0x3F800000u : 0x00003F80u;  // bf16 1.0
                        }
                    }
                    *reinterpret_cast<uint4*>(dst + c8 * L.plane_bytes) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                ptx::tc_fence_before();  // my earlier tcgen05.ld of this tile's accumulators precede the MMAs that overwrite them
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&ready[t / p.group_tiles]);
            }
            for (int l = 0; l <= p.layers; ++l) {
                const bool head = l == p.layers;
                // Last layer of this CTA's last round: the head FC kernel may be staged now (it runs its prologue and then waits
                // for this whole grid).  Not earlier: staged CTAs hold their shared memory while they wait, and with many lanes
                // in flight they would take SMs away from the other lanes' trunks for the length of this kernel.
                if (l == p.layers - 1 && rd + static_cast<int>(gridDim.x) >= rounds && warp == 0 && lane == 0) ptx::grid_dep_launch();
                const uint32_t par = static_cast<uint32_t>(l & 1);
                const bool has_resid = l >= 2 && par == 0;  // conv2 of a block: + block input (Q), result back into Q
                uint8_t* outb = par ? bufP : bufQ;          // stem -> Q, conv1 -> P, conv2 -> Q
                const float4* bias4 = reinterpret_cast<const float4*>(bias_sm + l * 16);
                for (int t = slot; t < T; t += 4) {
                    const int q = t * 128 + r;
                    const uint32_t inf = info[q];
                    const int board = board0 + static_cast<int>(inf >> 7);
                    const bool valid = inf != kTsPad && board < n_valid;
                    ptx::mbar_wait(&acc_full[t], par, p.err, 0x6100 + t);
                    ptx::tc_fence_after();
                    CB2_TS_TRACE(const bool trace = p.dbg != nullptr && blockIdx.x == 0 && rd == 0 && warp == 0 && lane == 0 && t == 0;)
                    CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 2] = clock64();)
                    const uint32_t lane_addr = tmem_base + ((q4 * 32u) << 16);
                    if (!head) {
                        uint32_t raw[16], raw2[16];
                        ptx::tmem_ld_x16_issue(lane_addr + static_cast<uint32_t>(t * 64) + par * 32, raw);
                        ptx::tmem_ld_x16_issue(lane_addr + static_cast<uint32_t>(t * 64) + par * 32 + 16, raw2);
                        float4 bc[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) bc[i] = bias4[i];
                        uint8_t* cellp = outb + (p.margin + q) * 16;
                        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
                        if (has_resid) {
                            r0 = *reinterpret_cast<const uint4*>(cellp);
                            r1 = *reinterpret_cast<const uint4*>(cellp + L.plane_bytes);
                        }
                        ptx::tmem_ld_wait2(raw, raw2);
                        CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 3] = clock64();)
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            v[4 * i + 0] = (__uint_as_float(raw[4 * i + 0]) + __uint_as_float(raw2[4 * i + 0])) + bc[i].x;
                            v[4 * i + 1] = (__uint_as_float(raw[4 * i + 1]) + __uint_as_float(raw2[4 * i + 1])) + bc[i].y;
                            v[4 * i + 2] = (__uint_as_float(raw[4 * i + 2]) + __uint_as_float(raw2[4 * i + 2])) + bc[i].z;
                            v[4 * i + 3] = (__uint_as_float(raw[4 * i + 3]) + __uint_as_float(raw2[4 * i + 3])) + bc[i].w;
                        }
                        if (has_resid) {
                            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[i]);
                                v[2 * i + 0] += __bfloat162float(b2.x);
                                v[2 * i + 1] += __bfloat162float(b2.y);
                            }
                        }
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f));
                            o[i] = valid ? *reinterpret_cast<const uint32_t*>(&b2) : 0u;  // pad cells must stay zero
                        }
                        *reinterpret_cast<uint4*>(cellp) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(cellp + L.plane_bytes) = make_uint4(o[4], o[5], o[6], o[7]);
                        CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 4] = clock64();)
                        ptx::tc_fence_before();
                        ptx::fence_proxy_async_smem();
                        CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 5] = clock64();)
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&ready[t / p.group_tiles]);
                        CB2_TS_TRACE(if (trace) p.dbg[l * 8 + 6] = clock64();)
                    } else {
                        // ---- both 1x1 head convolutions: columns [0, vhp) value, [vhp, vhp + php) policy
                        const float4* hb4 = reinterpret_cast<const float4*>(bias_sm + p.layers * 16);
                        const size_t row = static_cast<size_t>(board) * s2 + (inf & 127u);
                        for (int cb = 0; cb < nh; cb += 16) {
                            uint32_t raw[16];
                            ptx::tmem_ld_x16_issue(lane_addr + static_cast<uint32_t>(t * 64 + cb), raw);
                            float4 bc[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) bc[i] = hb4[cb / 4 + i];
                            ptx::tmem_ld_wait(raw);
                            uint32_t o[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float v0 = fmaxf(__uint_as_float(raw[4 * i + 0]) + bc[i].x, 0.0f), v1 = fmaxf(__uint_as_float(raw[4 * i + 1]) + bc[i].y, 0.0f);
                                const float v2 = fmaxf(__uint_as_float(raw[4 * i + 2]) + bc[i].z, 0.0f), v3 = fmaxf(__uint_as_float(raw[4 * i + 3]) + bc[i].w, 0.0f);
                                const __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), b = __floats2bfloat162_rn(v2, v3);
                                o[2 * i] = *reinterpret_cast<const uint32_t*>(&a);
                                o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b);
                            }
                            if (valid) {
                                // 16-column block cb lies entirely in the value part or entirely in the policy part (vhp % 16 == 0)
                                uint4* og = cb < p.vhp ? reinterpret_cast<uint4*>(p.out_v + row * p.vhp + cb) : reinterpret_cast<uint4*>(p.out_p + row * p.php + (cb - p.vhp));
                                og[0] = make_uint4(o[0], o[1], o[2], o[3]);
                                og[1] = make_uint4(o[4], o[5], o[6], o[7]);
                            }
                        }
                        ptx::tc_fence_before();
                    }
                }
            }
        }
    }

    // ---- teardown
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kTsIssuerWarp) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace cb2
