/* CPU oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the scalar host loops the reference runs
 * around Model::run.  Built by oracle/Makefile into oracle/_build/liboracle.so; loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.
 *
 *   oracle_planes_to_tensor   engine/src/net/mod.rs:121-156  (per-element bit test, zero-filled padding rows)
 *   oracle_clamp_non_finite   engine/src/net/mod.rs:57-61    (non-finite logit -> f32::MIN)
 *   oracle_calc_moves_probs   engine/src/net/mod.rs:106-119  (gather legal, max fold from f32::MIN, exp, sum, divide)
 *
 * Pinned by tests/test_oracle_golden.py against the numpy restatement (oracle/games.py) and, through it, against
 * the reference's DataSet.unpack_planes fixtures in tests/golden/encode_ref.npz.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

/* words: [n][planes][wpp] u64, little word first (hex u128 = lo, hi: training/self-play/src/serialize/hex.rs:16-28) */
void oracle_planes_to_tensor(const uint64_t* words, uint32_t n, uint32_t batch_size, uint32_t planes, uint32_t board,
                             float* out /* [batch_size][planes][board][board] */) {
    const uint32_t s2 = board * board, wpp = (s2 + 63) / 64;
    for (uint32_t b = 0; b < n; ++b)
        for (uint32_t c = 0; c < planes; ++c) {
            const uint64_t* plane = words + ((size_t)b * planes + c) * wpp;
            for (uint32_t h = 0; h < board; ++h)
                for (uint32_t w = 0; w < board; ++w) {
                    const uint32_t idx = h * board + w;
                    out[(((size_t)b * planes + c) * board + h) * board + w] = ((plane[idx >> 6] >> (idx & 63)) & 1) ? 1.0f : 0.0f;
                }
        }
    if (batch_size > n) memset(out + (size_t)n * planes * s2, 0, (size_t)(batch_size - n) * planes * s2 * sizeof(float));
}

void oracle_clamp_non_finite(float* scores, uint32_t len) {
    for (uint32_t i = 0; i < len; ++i)
        if (!isfinite(scores[i])) scores[i] = -FLT_MAX;
}

/* legal: ascending nn indices; probs_out[k] belongs to legal[k].  Returns the number written. */
uint32_t oracle_calc_moves_probs(const uint32_t* legal, uint32_t n_legal, const float* move_scores, float* probs_out) {
    float max_p = -FLT_MAX;
    for (uint32_t k = 0; k < n_legal; ++k) max_p = fmaxf(max_p, move_scores[legal[k]]);
    float sum = 0.0f;
    for (uint32_t k = 0; k < n_legal; ++k) {
        probs_out[k] = expf(move_scores[legal[k]] - max_p);
        sum += probs_out[k];
    }
    for (uint32_t k = 0; k < n_legal; ++k) probs_out[k] /= sum;
    return n_legal;
}

/* Whole host tail for a batch given a 1-bit-per-move legal bitmap (training/self-play/src/serialize/chess.rs:34-41
 * layout): clamp, gather, softmax; compact output, offsets[n+1]. */
void oracle_policy_batch(const float* logits, uint32_t n, uint32_t moves, const uint8_t* bitmaps, uint32_t bitmap_stride,
                         float* probs_out, uint32_t* offsets) {
    uint32_t legal[4096];
    uint32_t off = 0;
    float row[4096];
    for (uint32_t b = 0; b < n; ++b) {
        memcpy(row, logits + (size_t)b * moves, moves * sizeof(float));
        oracle_clamp_non_finite(row, moves);
        uint32_t nl = 0;
        for (uint32_t i = 0; i < moves; ++i)
            if ((bitmaps[(size_t)b * bitmap_stride + (i >> 3)] >> (i & 7)) & 1) legal[nl++] = i;
        offsets[b] = off;
        off += oracle_calc_moves_probs(legal, nl, row, probs_out + off);
    }
    offsets[n] = off;
}
