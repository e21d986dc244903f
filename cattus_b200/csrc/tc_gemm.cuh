// tc_gemm: the generic tcgen05/TMEM bf16 GEMM used for every dense contraction of ConvNetV1
// (reference network: training/cattus_train/net_utils.py:45-89; engines it replaces: engine/src/net/model.rs:146-218).
//
//   D[128 x N] (fp32, TMEM) = sum over k-blocks of A[128 x 64] * B[N x 64]^T      (bf16 operands, 64 = one 128-byte row)
// B lives in memory as the sequence of its smem stages ([n_tile][k-block][N rows][64 k], rows already 128-byte swizzled,
// engine.cu tile_b()): one stage is one linear bulk copy.  A mode-0 A box holds min(128, rows) rows: the accumulator rows of
// the stale smem rows below it are never read.
//
// mode 0 (matrix):  A is a row-major [rows][K] bf16 matrix (FC layers, 1x1 head convs).
// mode 1 (conv3x3): A is the NHWC activation tensor [B][S][S][C]; the k-loop walks (tap, 64-channel slice) and each
//                   A tile is ONE 4-D TMA box {64ch, S, S, nb boards} whose start coordinate is shifted by the tap
//                   (dx, dy) in {-1,0,1}^2 -- TMA zero-fills the out-of-bounds halo, which is exactly `padding="same"`
//                   (net_utils.py:13,29,32).  A tile always holds whole boards, so no tile straddles two positions.
// Epilogue variants (`epi`): 0 = +bias (+residual) -> ReLU -> bf16 / f32 store.
//   1 = value head tail fused (N = 128 hidden units in one tile): relu(acc + b1) . w2 + b2 -> tanh -> values[row]
//       (net_utils.py:71-74); the hidden activations never leave the SM.
//   2 = policy tail fused (all M <= 128 moves in one tile: tic-tac-toe, hex): + bias, non-finite -> f32::MIN, softmax
//       over the legal moves only, compact write at the offset stored in the record prefix (engine/src/net/mod.rs:57-61,
//       :106-119).  One thread owns one position's whole logits row in TMEM; the sum runs in ascending move order like
//       the reference's sequential `iter().sum()`.
//   3 = policy mask fused for M > 128 (chess, 15 N tiles): + bias, non-finite -> f32::MIN, and only the logits of LEGAL
//       moves are written, compactly, at (record offset + number of legal moves in earlier columns) -- 0.5 MB instead of
//       the 31.5 MB dense f32 logits tensor per 4096 chess positions; softmax_compact_kernel then normalises.
// Warp roles (224 threads): warp 0 = A producer (TMA), warp 6 = B producer (bulk copy), warp 1 = MMA issuer, warps 2..5 = epilogue
// (TMEM -> registers -> +bias (+residual) -> ReLU -> bf16/f32 -> global).  Smem ring of 3 slots of one k-block (two CTAs per SM) or of two (one CTA per SM; template parameter), full/empty mbarriers.
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"

namespace cb2 {

constexpr int kTcStages = 3;      // ring depth of large grids (two CTAs per SM)
constexpr int kTcStagesDeep = 6;  // tile pairs of the deep form (3 slots of 2 k-blocks, one CTA per SM), see tc_gemm_body
constexpr int kTcTileBytes = 128 * 128;  // 128 rows x 128 B
constexpr int kTcThreads = 224;
constexpr int kTcTmemCols = 128;
constexpr int kTcSmemBytes = 2 * kTcStages * kTcTileBytes + 256 + 1024;  // tiles + barriers + alignment slack
constexpr int tc_smem_bytes(int stages) { return 2 * stages * kTcTileBytes + 256 + 1024; }

struct alignas(64) TcGemmParams {
    CUtensorMap tma_a;
    const uint8_t* b_img;         // B as pre-swizzled smem stages [n_tile][k-block][n_umma rows][128 B] (engine.cu tile_b())
    const float* bias;            // [n_tiles * n_umma], zero padded
    const __nv_bfloat16* resid;   // same layout as out (bf16) or nullptr
    void* out;
    uint32_t* err;
    int mode;
    int num_kb;         // number of 64-wide k-blocks
    int kh;             // mode 1: 64-channel slices per tap
    int n_umma;         // UMMA N: multiple of 16, <= 128
    int n_store;        // columns written per row (bf16: multiple of 8; f32: n_umma)
    int m_valid;        // mode 0: rows; mode 1: boards
    int rows_per_tile;  // mode 0: 128; mode 1: nb * S * S
    int s2;             // S * S
    int nb;             // boards per tile
    int ld_out;         // elements
    int out_f32;
    int relu;
    uint32_t tx_bytes;  // bytes one stage's two TMA boxes deliver
    int epi;            // epilogue variant, see above
    int fault;          // 1: fault injection, see the MMA issuer
    // epi 1
    const float* w2;    // [128]
    float b2;
    float* values;      // [positions]
    // epi 2
    const uint8_t* recs;  // record 0's planes (kernels.cuh: RecLayout)
    RecLayout rl;
    float* probs;         // compact output
    // epi 1, 2
    const uint32_t* n_ptr;
    unsigned long long* dbg;  // optional clock64 trace of tile (0, 1) (CATTUS_B200_TRACE_HEADS=1, printed by time_stage), else nullptr
};

// kStages ring slots of kKPer k-blocks each.  <3, 1>: the general form, two CTAs per SM.  <3, 2> ("deep", one CTA per SM,
// the same 192 KB as six single slots would take): half the barrier traffic per k -- the issuer's chain per slot is one
// wait + 8 MMAs + one commit, which fits under the 8 MMAs' 512 cycles of tensor work, where one wait + 4 MMAs + one commit
// (~430 cycles) did not fit under 256.  Needs an even number of k-blocks.
template <int kStages, int kKPer>
__device__ __forceinline__ void tc_gemm_body(const TcGemmParams& p, const int m_tile, const int n_tile) {
    // An M tile that holds nothing but padding positions (the launch is sized for the batch bucket) has no live output.
    // Mode 1 tiles are whole boards; mode 0 rows are positions only for the head FCs (the epilogues 1-3), elsewhere rows
    // are board cells and the tile is always computed.
    const bool trace = p.dbg != nullptr && m_tile == 0 && n_tile == 1 && p.epi != 1;
    const int trace_slot = 256 + 4 * (p.epi == 1 ? 0 : n_tile + 1);  // every CTA of the first M tile: entry / exit in ns and cycles
    if (p.dbg != nullptr && m_tile == 0 && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.dbg[trace_slot] = ns;
        p.dbg[trace_slot + 1] = clock64();
    }
    if (trace && threadIdx.x == 0) p.dbg[0] = clock64();
    if (p.n_ptr != nullptr && (p.mode == 1 || p.epi != 0) && m_tile * (p.mode == 0 ? 128 : p.nb) >= static_cast<int>(*p.n_ptr)) return;
    if (trace && threadIdx.x == 0) p.dbg[1] = clock64();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    constexpr int stages = kStages;  // compile time: the loops below unroll over one trip around the ring
    constexpr int kSlotBytes = kKPer * kTcTileBytes;  // per operand
    uint8_t* smem_b = smem + stages * kSlotBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * stages * kSlotBytes);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* tmem_full_bar = empty_bar + stages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&p.tma_a);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&full_bar[s], 2);  // the A and the B producer each arrive with their own byte count
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tmem_full_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr, kTcTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // Everything above ran beside the tail of the kernel before this one (the trunk); its output is read from here on.
    ptx::grid_dep_wait();
    ptx::grid_dep_launch();  // the next kernel (softmax_compact) may stage itself; it waits for this whole grid before it reads
    if (trace && threadIdx.x == 0) p.dbg[2] = clock64();

    // Producer and issuer: the WHOLE warp runs the loop (warp-uniform control flow and addresses, so descriptors and
    // coordinates stay in uniform registers) and one elected lane issues the TMA / tcgen05 instructions.  With the loops
    // inside `if (lane == 0)` the issuer needed ~600 cycles per k-block of four MMAs (256 cycles of tensor-pipe work) and the
    // producer ~680 per pair of loads: every GEMM of this file was issue-bound (clock64 trace).
    // Both loops are unrolled over one trip around the ring (the ring depth is a template parameter), so slot, barrier and
    // tile addresses are immediates, and the rest of the bookkeeping is incremental: one warp runs a dependent chain, and the
    // `kb % stages` / `kb / kh` divisions alone were ~60 of the ~100 instructions per k-block.
    if (warp == 0) {
        // ---- A producer
        const bool leader = ptx::elect_one();
        const uint32_t a_bytes = p.tx_bytes - static_cast<uint32_t>(p.n_umma) * 128;
        const int num_kb = p.num_kb, kh = p.kh, mode = p.mode;
        const int a_c1 = mode == 0 ? m_tile * 128 : m_tile * p.nb;
        int slice = 0, dx = -1, dy = -1;
        uint32_t ph = 1;  // parity of the `empty` phase to wait for: the first trip passes on fresh barriers
        for (int kb0 = 0; kb0 < num_kb; kb0 += stages * kKPer, ph ^= 1) {
#pragma unroll
            for (int s = 0; s < stages; ++s) {
                const int kb = kb0 + s * kKPer;
                if (kb >= num_kb) break;
                ptx::mbar_wait(&empty_bar[s], ph, p.err, 0x100 + s);
                if (trace && lane == 0 && kb < 40) p.dbg[48 + kb] = clock64();
                if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], a_bytes * kKPer);
#pragma unroll
                for (int i = 0; i < kKPer; ++i) {
                    if (leader) {
                        if (mode == 0)
                            ptx::tma_load_2d(smem_a + (s * kKPer + i) * kTcTileBytes, &p.tma_a, &full_bar[s], (kb + i) * 64, a_c1);
                        else
                            ptx::tma_load_4d(smem_a + (s * kKPer + i) * kTcTileBytes, &p.tma_a, &full_bar[s], slice * 64, dx, dy, a_c1);
                    }
                    if (++slice == kh) {  // mode 1: k-blocks walk (tap, 64-channel slice), taps row-major over (dy, dx)
                        slice = 0;
                        if (++dx == 2) {
                            dx = -1;
                            ++dy;
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 6) {
        // ---- B producer (its own warp: try_wait + expect_tx + one load is ~250 cycles of dependent issue, the MMAs of one
        // k-block take 256; one warp issuing both loads needed ~470 and paced the whole k-loop)
        const bool leader = ptx::elect_one();
        const uint8_t* b_src = p.b_img + static_cast<size_t>(n_tile) * p.num_kb * p.n_umma * 128;
        const uint32_t b_bytes = static_cast<uint32_t>(p.n_umma) * 128;
        const int num_kb = p.num_kb;
        uint32_t ph = 1;
        for (int kb0 = 0; kb0 < num_kb; kb0 += stages * kKPer, ph ^= 1) {
#pragma unroll
            for (int s = 0; s < stages; ++s) {
                if (kb0 + s * kKPer >= num_kb) break;
                ptx::mbar_wait(&empty_bar[s], ph, p.err, 0x180 + s);
                if (leader) {
                    ptx::mbar_arrive_expect_tx(&full_bar[s], b_bytes * kKPer);
#pragma unroll
                    for (int i = 0; i < kKPer; ++i)  // a k-block's rows start a whole tile apart in smem whatever n_umma is
                        ptx::bulk_load(smem_b + (s * kKPer + i) * kTcTileBytes, b_src + i * b_bytes, b_bytes, &full_bar[s]);
                }
                b_src += b_bytes * kKPer;
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, p.n_umma);
        const uint64_t desc0 = ptx::umma_desc_sw128(0, 1024);  // stage tiles are 1024-byte aligned: only the start address varies
        const uint32_t d_hi = static_cast<uint32_t>(desc0 >> 32), d_lo_fixed = static_cast<uint32_t>(desc0);
        const uint32_t a_lo0 = d_lo_fixed | (ptx::smem_u32(smem_a) >> 4), b_lo0 = d_lo_fixed | (ptx::smem_u32(smem_b) >> 4);
        const bool leader = ptx::elect_one();
        const int num_kb = p.num_kb;
        uint32_t ph = 0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += stages * kKPer, ph ^= 1) {
#pragma unroll
            for (int s = 0; s < stages; ++s) {
                const int kb = kb0 + s * kKPer;
                if (kb >= num_kb) break;
                if (trace && lane == 0 && kb < 40) p.dbg[88 + kb] = clock64();
                ptx::mbar_wait(&full_bar[s], ph, p.err, 0x200 + s);
                ptx::tc_fence_after();
                if (trace && lane == 0 && kb < 40) p.dbg[8 + kb] = clock64();
                if (leader) {
#pragma unroll
                    for (int i = 0; i < kKPer; ++i) {
                        const uint32_t a_lo = a_lo0 + (s * kKPer + i) * (kTcTileBytes >> 4), b_lo = b_lo0 + (s * kKPer + i) * (kTcTileBytes >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle row
                            ptx::umma_bf16_ss_lohi(tmem_base, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, (kb | i | k) != 0);
                    }
                    ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
                }
                __syncwarp();
            }
        }
        // fault injection (cattus_b200_desc.flags bit 1, tests only): tile (0, 0) never publishes its accumulator, so the
        // epilogue's bounded wait below expires, records its code and traps -- the path a pipeline bug would take
        if (leader && !(p.fault && m_tile == 0 && n_tile == 0)) ptx::umma_commit(tmem_full_bar);  // accumulator complete
    } else {
        // Epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); warps 2,3,4,5 cover all four quarters.
        const uint32_t q = warp & 3;
        const int row = static_cast<int>(q * 32 + lane);
        bool ok;
        long long grow;
        if (p.mode == 0) {
            grow = static_cast<long long>(m_tile) * 128 + row;
            ok = grow < p.m_valid;
        } else {
            ok = row < p.rows_per_tile && (m_tile * p.nb + row / p.s2) < p.m_valid;
            grow = static_cast<long long>(m_tile) * p.rows_per_tile + row;
        }
        // epi 2 and 3: what depends only on the records -- this tile's legal words, the number of legal moves in the columns before
        // it -- and the tile's bias lines are fetched NOW, while the k-loop runs; after the accumulator is complete the epilogue
        // touches global memory only to store.  (Fetched after the wait, these latencies were most of its 8 500 cycles.)
        int n = 0;
        bool live = false;
        uint32_t pos = 0;
        uint32_t legal[4] = {0, 0, 0, 0};
        if (p.epi == 2) {
            n = static_cast<int>(*p.n_ptr);
            live = grow < n;
            if (static_cast<int>(lane) * 32 < p.n_umma) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + lane * 32));
            if (live) {
                const uint8_t* rec = p.recs + static_cast<size_t>(grow) * p.rl.rec_bytes;
                pos = *reinterpret_cast<const uint32_t*>(rec - 8);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < p.rl.legal_words) legal[j] = legal_word(rec, p.rl, j);
            }
        }
        if (p.epi == 3) {
            if (lane < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + n_tile * p.n_umma + lane * 32));
            n = static_cast<int>(*p.n_ptr);
            live = grow < n;
            const uint8_t* rec = p.recs + static_cast<size_t>(live ? grow : 0) * p.rl.rec_bytes;
            const int w0 = n_tile * (p.n_umma >> 5);  // first 32-bit legal word of this tile
            if (p.rl.legal_off >= 0 && (w0 & 1) == 0 && w0 <= 64) {
                // Legal moves in the columns before this tile, for the warp's 32 rows: lane k reads u64 word k of a row's
                // bitmap (one coalesced request per row, 16 rows in flight) and the popcounts are added across the warp.
                const int nw64 = w0 >> 1;
                const long long g0 = static_cast<long long>(m_tile) * 128 + q * 32;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    unsigned long long wv[16];
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const long long g = g0 + 16 * h + r;
                        const unsigned long long* rr =
                            reinterpret_cast<const unsigned long long*>(p.recs + static_cast<size_t>(g < n ? g : 0) * p.rl.rec_bytes + p.rl.legal_off);
                        wv[r] = static_cast<int>(lane) < nw64 ? __ldg(rr + lane) : 0ull;
                    }
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, static_cast<uint32_t>(__popcll(wv[r])));
                        if (static_cast<int>(lane) == 16 * h + r) pos = c;
                    }
                }
                if (live) pos += *reinterpret_cast<const uint32_t*>(rec - 8);
            } else if (live) {
                pos = *reinterpret_cast<const uint32_t*>(rec - 8);
                for (int j = 0; j < w0; ++j) pos += __popc(legal_word(rec, p.rl, j));
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w0 + j < p.rl.legal_words) legal[j] = legal_word(rec, p.rl, w0 + j);
            }
        }
        ptx::mbar_wait(tmem_full_bar, 0, p.err, 0x300, p.fault ? (1u << 14) : (1u << 24));
        ptx::tc_fence_after();
        if (trace && threadIdx.x == 64) p.dbg[3] = clock64();
        const uint32_t taddr = tmem_base + ((q * 32u) << 16);
        if (p.epi == 1) {
            // ---- value head: tanh(b2 + sum_j relu(acc_j + b1_j) * w2_j)
            const int n = static_cast<int>(*p.n_ptr);
            float dot = 0.0f;
            for (int c0 = 0; c0 < 128; c0 += 16) {
                float v[16];
                ptx::tmem_ld_x16(taddr + c0, v);
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + c0);
                const float4* w4 = reinterpret_cast<const float4*>(p.w2 + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b = __ldg(b4 + j), w = __ldg(w4 + j);
                    dot = fmaf(fmaxf(v[4 * j + 0] + b.x, 0.0f), w.x, dot);
                    dot = fmaf(fmaxf(v[4 * j + 1] + b.y, 0.0f), w.y, dot);
                    dot = fmaf(fmaxf(v[4 * j + 2] + b.z, 0.0f), w.z, dot);
                    dot = fmaf(fmaxf(v[4 * j + 3] + b.w, 0.0f), w.w, dot);
                }
            }
            if (grow < n) p.values[grow] = tanhf(dot + p.b2);
        } else if (p.epi == 2) {
            // ---- policy: masked softmax over this position's row, three passes over TMEM (max, sum, write); the legal words and
            // the output offset were fetched before the accumulator wait
            float mx = -FLT_MAX;
            for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
                float v[16];
                ptx::tmem_ld_x16(taddr + c0, v);
                const uint32_t bits = (legal[c0 >> 5] >> (c0 & 31)) & 0xFFFFu;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = v[j] + __ldg(p.bias + c0 + j);
                    if (!isfinite(x)) x = -FLT_MAX;
                    if ((bits >> j) & 1u) mx = fmaxf(mx, x);
                }
            }
            float sum = 0.0f;
            for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
                float v[16];
                ptx::tmem_ld_x16(taddr + c0, v);
                const uint32_t bits = (legal[c0 >> 5] >> (c0 & 31)) & 0xFFFFu;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = v[j] + __ldg(p.bias + c0 + j);
                    if (!isfinite(x)) x = -FLT_MAX;
                    if ((bits >> j) & 1u) sum += expf(x - mx);
                }
            }
            for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
                float v[16];
                ptx::tmem_ld_x16(taddr + c0, v);
                const uint32_t bits = (legal[c0 >> 5] >> (c0 & 31)) & 0xFFFFu;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = v[j] + __ldg(p.bias + c0 + j);
                    if (!isfinite(x)) x = -FLT_MAX;
                    if (live && ((bits >> j) & 1u)) p.probs[pos++] = expf(x - mx) / sum;
                }
            }
        } else if (p.epi == 3) {
            // ---- policy, many N tiles: masked compact write of this tile's 128 columns (softmax runs afterwards)
            // Two 16-column chunks per trip (one TMEM wait for both); their 32 biases come as vector loads from lines prefetched
            // above, under the TMEM load.  Per-bit branches, not predicated stores and not a staging pass through shared
            // memory: a warp skips the columns none of its 32 rows may play, which is most of them (tried: 16 predicated STG per
            // chunk 13 200 cycles per full tile, shared-memory staging + ffs loop 5 800, branches 5 200).
            for (int c0 = 0; c0 < p.n_umma; c0 += 32) {
                uint32_t raw[2][16];
                ptx::tmem_ld_x16_issue(taddr + c0, raw[0]);
                ptx::tmem_ld_x16_issue(taddr + c0 + 16, raw[1]);
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + n_tile * p.n_umma + c0);
                float4 bb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) bb[i] = __ldg(b4 + i);
                ptx::tmem_ld_wait2(raw[0], raw[1]);
                const uint32_t bits32 = legal[c0 >> 5];
                if (bits32 == 0) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t bits = (bits32 >> (16 * h)) & 0xFFFFu;
                    if (bits == 0) continue;
                    const float bias16[16] = {bb[4 * h].x,     bb[4 * h].y,     bb[4 * h].z,     bb[4 * h].w,     bb[4 * h + 1].x, bb[4 * h + 1].y,
                                              bb[4 * h + 1].z, bb[4 * h + 1].w, bb[4 * h + 2].x, bb[4 * h + 2].y, bb[4 * h + 2].z, bb[4 * h + 2].w,
                                              bb[4 * h + 3].x, bb[4 * h + 3].y, bb[4 * h + 3].z, bb[4 * h + 3].w};
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if ((bits >> j) & 1u) {
                            float x = __uint_as_float(raw[h][j]) + bias16[j];
                            if (!isfinite(x)) x = -FLT_MAX;
                            p.probs[pos++] = x;
                        }
                    }
                }
            }
        } else
        for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
            float v[16];
            ptx::tmem_ld_x16(taddr + c0, v);  // warp-collective: executed by all lanes regardless of `ok`
            if (!ok || c0 >= p.n_store) continue;
            const int col = n_tile * p.n_umma + c0;
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b = __ldg(b4 + j);
                v[4 * j + 0] += b.x;
                v[4 * j + 1] += b.y;
                v[4 * j + 2] += b.z;
                v[4 * j + 3] += b.w;
            }
            const long long off = grow * p.ld_out + col;
            if (p.resid != nullptr) {
                const uint4* r4 = reinterpret_cast<const uint4*>(p.resid + off);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c0 + 8 * (h + 1) > p.n_store) break;
                    const uint4 r = __ldg(r4 + h);
                    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                        v[8 * h + 2 * j + 0] += __bfloat162float(b2.x);
                        v[8 * h + 2 * j + 1] += __bfloat162float(b2.y);
                    }
                }
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            if (p.out_f32) {
                float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
                uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c0 + 8 * (h + 1) > p.n_store) break;
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * h + 2 * j], v[8 * h + 2 * j + 1]);
                        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    o4[h] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }

    if (trace && threadIdx.x == 64) p.dbg[4] = clock64();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, kTcTmemCols);
    if (trace && threadIdx.x == 64) p.dbg[5] = clock64();
    if (p.dbg != nullptr && m_tile == 0 && threadIdx.x == 64) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.dbg[trace_slot + 2] = ns;
        p.dbg[trace_slot + 3] = clock64();
    }
}

template <int kKPer>
__global__ void __launch_bounds__(kTcThreads, kKPer > 1 ? 1 : 2) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
    tc_gemm_body<kTcStages, kKPer>(p, blockIdx.x, blockIdx.y);
}

// Two independent GEMMs in one launch: blockIdx.y == 0 runs problem `a` (one N tile), blockIdx.y >= 1 runs N tile
// blockIdx.y - 1 of problem `b`.  Used for the two head FCs (value FC1 + tanh tail | policy FC + mask / softmax): the
// value problem alone has only ceil(B / 128) CTAs with a 32-step latency-bound k-loop, so it hides behind the policy's.
struct alignas(64) TcGemmDualParams {
    TcGemmParams a;
    TcGemmParams b;
};
template <int kKPer>
__global__ void __launch_bounds__(kTcThreads, kKPer > 1 ? 1 : 2) tc_gemm_dual_kernel(const __grid_constant__ TcGemmDualParams d) {
    if (blockIdx.y == 0)
        tc_gemm_body<kTcStages, kKPer>(d.a, blockIdx.x, 0);
    else
        tc_gemm_body<kTcStages, kKPer>(d.b, blockIdx.x, static_cast<int>(blockIdx.y) - 1);
}

}  // namespace cb2
