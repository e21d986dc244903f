// TEST INFRASTRUCTURE -- never part of the product library.
//
// The device-resident search (cattus_b200/csrc/dsearch_core.hpp) compiled for the HOST with a one-lane "warp", behind the
// same host driver the GPU uses (dsearch_host.hpp), with a callback evaluator.  The CPU tests use it to replay whole games
// against the host driver (csrc/selfplay.cpp) and the oracle: same moves, same .traindata bytes.  What it cannot check is
// the cross-lane part (shuffles, __syncwarp); the -m gpu tests cover that on the device.
#include <memory>
#include <string>
#include <vector>

#include "../../cattus_b200/csrc/dsearch_host.hpp"

namespace {

template <class Rules>
struct EmulBackend {
    using Pos = typename Rules::Pos;
    ds::Params<Rules> p{};
    const Rules& R;
    uint32_t depth_;
    cattus_b200_eval_fn fn[2];
    void* ctx[2];
    std::vector<ds::SlotState> slots;
    std::vector<uint32_t> pools;
    std::vector<ds::PathStep> paths;
    std::vector<float> noise;
    std::vector<Pos> hist;
    std::vector<uint8_t> block[2], block_b[2];          // _b: the second population's evaluator blocks and caches
    std::vector<float> values[2], probs[2], values_b[2], probs_b[2];
    std::vector<uint32_t> cache_meta_b[2];
    std::vector<uint8_t> cache_entries_b[2];
    ds::Params<Rules> p2{};
    uint32_t n_pops = 1, split = 0, total_slots = 0;
    std::vector<uint8_t> cmds, results;
    std::vector<uint32_t> done_per_buf;
    uint32_t done_count = 0, error = 0, max_used = 0;
    unsigned long long counters[4] = {0, 0, 0, 0};
    uint32_t n_bufs;
    bool chess;

    std::vector<uint32_t> cache_meta[2];
    std::vector<uint8_t> cache_entries[2];

    EmulBackend(const Rules& rules, const void* rules_blob, uint32_t n_slots, uint32_t pool_words, uint32_t depth, const sp::Params params[2],
                cattus_b200_eval_fn f1, void* c1, cattus_b200_eval_fn f2, void* c2, uint32_t cache_size)
        : R(rules), depth_(depth & 0xFFu) {
        chess = Rules::kChess;
        fn[0] = f1;
        ctx[0] = c1;
        fn[1] = f2;
        ctx[1] = c2;
        const uint32_t maxc = Rules::kChess ? 224u : static_cast<uint32_t>(rules.max_children());
        n_bufs = (depth & 0xFFu) + 1;
        slots.assign(n_slots, ds::SlotState{});
        pools.assign(static_cast<size_t>(n_slots) * 3 * pool_words + 256, 0xDEADBEEFu);  // stale memory must not matter; + slack for select's early child fetch
        const uint32_t path_cap = Rules::kChess ? 256u : maxc + 2u;
        paths.resize(static_cast<size_t>(n_slots) * path_cap);
        noise.assign(static_cast<size_t>(n_slots) * maxc, 0.0f);
        const uint32_t hist_cap = Rules::kChess ? 128u : 1u;
        hist.resize(static_cast<size_t>(n_slots) * hist_cap);
        const uint32_t plane_words = Rules::kChess ? 18u : 3u * static_cast<uint32_t>((rules.moves_num() + 63) / 64);
        const uint32_t rec_bytes = 8u + plane_words * 8u + (Rules::kChess ? 240u : 0u);
        p.rules = rules_blob;
        p.slots = slots.data();
        p.n_slots = n_slots;
        p.pools = pools.data();
        p.pool_words = pool_words;
        p.paths = paths.data();
        p.path_cap = path_cap;
        p.noise = noise.data();
        p.max_children = maxc;
        p.hist = hist.data();
        p.hist_cap = hist_cap;
        p.n_evals = f2 ? 2 : 1;
        for (int e = 0; e < 2; ++e) {
            block[e].assign(16 + static_cast<size_t>(n_slots) * rec_bytes, 0);
            values[e].assign(n_slots, 0.0f);
            probs[e].assign(static_cast<size_t>(n_slots) * maxc, 0.0f);
            p.eval[e].n_ptr = reinterpret_cast<uint32_t*>(block[e].data());
            p.eval[e].recs = block[e].data() + 16 + 8;
            p.eval[e].values = values[e].data();
            p.eval[e].probs = probs[e].data();
            p.eval[e].rec_bytes = rec_bytes;
            p.eval[e].prob_stride = maxc;
            p.eval[e].max_rows = n_slots;
            p.eval[e].plane_words = plane_words;
            p.sim_num[e] = params[e].sim_num;
            p.explore[e] = params[e].explore_factor;
            p.noise_eps[e] = params[e].noise_eps;
        }
        for (int e = 0; e < 2; ++e) {
            p.cache[e] = ds::CacheIo{};
            if (cache_size && (e == 0 || f2)) {
                uint32_t buckets = 1;
                while (buckets * ds::kCacheWays < cache_size) buckets <<= 1;
                p.cache[e].entry_bytes = (40u + 4u * maxc + 15u) & ~15u;
                cache_meta[e].assign(static_cast<size_t>(buckets) * 8, 0u);
                cache_entries[e].assign(static_cast<size_t>(buckets) * ds::kCacheWays * p.cache[e].entry_bytes, 0xAB);
                p.cache[e].meta = cache_meta[e].data();
                p.cache[e].entries = cache_entries[e].data();
                p.cache[e].bucket_mask = buckets - 1;
                p.cache[e].enabled = 1;
            }
        }
        p.cmd_stride = ds::cmd_stride_for(maxc);
        p.result_stride = ds::result_stride_for(maxc);
        cmds.assign(16 + static_cast<size_t>(n_slots) * p.cmd_stride, 0);
        p.cmds = cmds.data();
        p.n_result_bufs = n_bufs;
        p.result_buf_bytes = static_cast<unsigned long long>(n_slots) * p.result_stride;
        results.assign(static_cast<size_t>(n_bufs) * p.result_buf_bytes, 0);
        p.results = results.data();
        p.done_count = &done_count;
        p.counters = counters;
        p.error = &error;
        p.max_used = &max_used;
        p.begin_lead = (depth >> 8) & 1u;  // test knobs: bit 8 of `depth`: begin overlapped; bit 9: two populations; bits 16.. = visit budget (0 = 24)
        p.visit_budget = (depth >> 16) ? (depth >> 16) : 24u;
        total_slots = n_slots;
        split = n_slots;
        if (((depth >> 9) & 1u) && n_slots >= 2) {
            // the GPU's arrangement with two evaluator lanes: populations [0, split) and [split, n) take waves in turn, each with
            // its own evaluator batch and its own cache
            n_pops = 2;
            split = (n_slots + 1) / 2;
            p2 = p;
            p.n_slots = split;
            p2.n_slots = n_slots - split;
            p2.slot_base = split;
            p2.slots = p.slots + split;
            p2.pools = p.pools + static_cast<size_t>(split) * 3 * pool_words;
            p2.paths = p.paths + static_cast<size_t>(split) * path_cap;
            p2.noise = p.noise + static_cast<size_t>(split) * maxc;
            p2.hist = p.hist + static_cast<size_t>(split) * hist_cap;
            for (int e = 0; e < 2; ++e) {
                block_b[e].assign(block[e].size(), 0);
                values_b[e].assign(n_slots, 0.0f);
                probs_b[e].assign(static_cast<size_t>(n_slots) * maxc, 0.0f);
                p2.eval[e].n_ptr = reinterpret_cast<uint32_t*>(block_b[e].data());
                p2.eval[e].recs = block_b[e].data() + 16 + 8;
                p2.eval[e].values = values_b[e].data();
                p2.eval[e].probs = probs_b[e].data();
                if (p.cache[e].enabled) {
                    cache_meta_b[e].assign(cache_meta[e].size(), 0u);
                    cache_entries_b[e].assign(cache_entries[e].size(), 0xCD);
                    p2.cache[e].meta = cache_meta_b[e].data();
                    p2.cache[e].entries = cache_entries_b[e].data();
                }
            }
        }
        done_per_buf.assign(n_bufs, 0);
    }
    uint32_t n_slots() const { return total_slots; }
    uint32_t populations() const { return n_pops; }
    uint32_t population_of(uint32_t slot) const { return slot >= split ? 1u : 0u; }
    uint32_t max_children() const { return p.max_children; }
    uint32_t depth() const { return depth_; }
    uint8_t* cmd_block(uint32_t) { return cmds.data(); }

    void run_eval(ds::Params<Rules>& p, std::vector<float>* probs, std::vector<float>* values, int e) {
        const uint32_t n = *p.eval[e].n_ptr;
        if (n == 0) return;
        const uint32_t pw = p.eval[e].plane_words, rb = p.eval[e].rec_bytes;
        std::vector<uint64_t> planes(static_cast<size_t>(n) * pw);
        std::vector<uint8_t> legal(chess ? static_cast<size_t>(n) * 235 : 0);
        size_t total = 0;
        for (uint32_t r = 0; r < n; ++r) {
            const uint8_t* rec = p.eval[e].recs + static_cast<size_t>(r) * rb;
            std::memcpy(planes.data() + static_cast<size_t>(r) * pw, rec, pw * 8);
            if (chess) std::memcpy(legal.data() + static_cast<size_t>(r) * 235, rec + pw * 8, 235);
            const uint32_t* prefix = reinterpret_cast<const uint32_t*>(rec - 8);
            if (prefix[0] != r * p.eval[e].prob_stride) throw sp::SpError{CATTUS_B200_EINVAL, "emul: record prefix offset is wrong"};
            total += prefix[1];
        }
        std::vector<float> pr(total + 1), va(n);
        std::vector<uint32_t> off(n + 1);
        const int rc = fn[e](ctx[e], planes.data(), chess ? legal.data() : nullptr, n, pr.data(), total, off.data(), va.data());
        if (rc != 0) throw sp::SpError{rc, "emul: evaluator callback failed"};
        for (uint32_t r = 0; r < n; ++r) {
            const uint32_t cnt = reinterpret_cast<const uint32_t*>(p.eval[e].recs + static_cast<size_t>(r) * rb - 8)[1];
            if (off[r + 1] - off[r] != cnt) throw sp::SpError{CATTUS_B200_EINVAL, "emul: evaluator returned a wrong number of probabilities"};
            std::memcpy(probs[e].data() + static_cast<size_t>(r) * p.eval[e].prob_stride, pr.data() + off[r], sizeof(float) * cnt);
            values[e][r] = va[r];
        }
    }

    void submit(uint32_t wave, uint32_t pop, uint32_t n_cmds) {
        ds::Params<Rules>& p = pop ? p2 : this->p;
        std::vector<float>* probs = pop ? probs_b : this->probs;
        std::vector<float>* values = pop ? values_b : this->values;
        done_count = 0;
        *p.eval[0].n_ptr = 0;
        *p.eval[1].n_ptr = 0;
        decltype(auto) rules = ds::RulesRef<Rules>::get(p.rules);
        // begin_lead = 1 is the GPU's arrangement: begin runs beside select / evaluator / expand of the same wave and takes
        // effect in the next one; here it runs in the middle of them
        if (!p.begin_lead)
            for (uint32_t ci = 0; ci < n_cmds; ++ci) ds::Core<Rules>::begin_slot(rules, p, ci, wave);
        for (uint32_t si = 0; si < p.n_slots; ++si) ds::Core<Rules>::select_slot(rules, p, si, wave);
        if (p.begin_lead)
            for (uint32_t ci = 0; ci < n_cmds; ++ci) ds::Core<Rules>::begin_slot(rules, p, ci, wave);
        for (uint32_t e = 0; e < p.n_evals; ++e) run_eval(p, probs, values, static_cast<int>(e));
        for (uint32_t si = 0; si < p.n_slots; ++si) ds::Core<Rules>::expand_slot(rules, p, si, wave);
        done_per_buf[wave % n_bufs] = done_count;
    }
    const uint8_t* wait(uint32_t wave, uint32_t* n_done) {
        if (error) throw sp::SpError{CATTUS_B200_ERANGE, "emul: device search error bits " + std::to_string(error)};
        *n_done = done_per_buf[wave % n_bufs];
        return results.data() + static_cast<size_t>(wave % n_bufs) * p.result_buf_bytes;
    }
    void read_counters(unsigned long long out[4]) {
        for (int i = 0; i < 4; ++i) out[i] = counters[i];
    }
};

struct EmulResult {
    sp::Shared sh;
    std::string error;
};

sp::Params params_from(const cattus_b200_selfplay_cfg* cfg) {
    sp::Params p;
    p.sim_num = cfg->sim_num;
    p.explore_factor = cfg->explore_factor;
    p.noise_alpha = cfg->prior_noise_alpha;
    p.noise_eps = cfg->prior_noise_epsilon;
    p.last_temperature = 1.0f;
    if (cfg->n_temperatures) {
        for (uint32_t i = 0; i + 1 < cfg->n_temperatures; ++i) p.temperatures.emplace_back(cfg->temperature_moves[i], cfg->temperature_values[i]);
        p.last_temperature = cfg->temperature_values[cfg->n_temperatures - 1];
    }
    return p;
}

template <class Rules>
void run_emul(const Rules& rules, const void* blob, const cattus_b200_selfplay_cfg* cfg, uint32_t n_slots, uint32_t pool_words, uint32_t depth,
              cattus_b200_eval_fn f1, void* c1, cattus_b200_eval_fn f2, void* c2, sp::Shared& sh) {
    const sp::Params p = params_from(cfg);
    sp::Params params[2] = {p, p};
    EmulBackend<Rules> be(rules, blob, n_slots, pool_words, depth, params, f1, c1, f2, c2, cfg->cache_size);
    ds::Driver<Rules, EmulBackend<Rules>> drv(rules, *cfg, params, be, sh);
    drv.run();
}

}  // namespace

extern "C" {

void* dsearch_emul_run(cattus_b200_eval_fn f1, void* c1, cattus_b200_eval_fn f2, void* c2, const cattus_b200_selfplay_cfg* cfg, uint32_t n_slots,
                       uint32_t pool_words, uint32_t depth) {
    std::unique_ptr<EmulResult> r(new EmulResult());
    try {
        if (cfg->game == CATTUS_B200_GAME_HEX) {
            if (cfg->board_size <= 8) {
                sp::HexRulesT<uint64_t> rules(static_cast<int>(cfg->board_size));
                run_emul(rules, &rules, cfg, n_slots, pool_words, depth, f1, c1, f2, c2, r->sh);
            } else {
                sp::HexRulesT<sp::u128> rules(static_cast<int>(cfg->board_size));
                run_emul(rules, &rules, cfg, n_slots, pool_words, depth, f1, c1, f2, c2, r->sh);
            }
        } else if (cfg->game == CATTUS_B200_GAME_CHESS) {
            sp::ChessRules rules;
            run_emul(rules, &sp::chess_tables(), cfg, n_slots, pool_words, depth, f1, c1, f2, c2, r->sh);
        } else {
            sp::TttRules rules;
            run_emul(rules, &rules, cfg, n_slots, pool_words, depth, f1, c1, f2, c2, r->sh);
        }
    } catch (const sp::SpError& e) {
        r->error = e.msg.empty() ? "error" : e.msg;
    }
    std::sort(r->sh.records.begin(), r->sh.records.end(), [](const sp::GameRecord& a, const sp::GameRecord& b) { return a.game_idx < b.game_idx; });
    return r.release();
}
const char* dsearch_emul_error(void* h) { return static_cast<EmulResult*>(h)->error.c_str(); }
uint32_t dsearch_emul_game_count(void* h) { return static_cast<uint32_t>(static_cast<EmulResult*>(h)->sh.records.size()); }
void dsearch_emul_counters(void* h, uint64_t out[8]) {
    const sp::Shared& s = static_cast<EmulResult*>(h)->sh;
    out[0] = s.simulations;
    out[1] = s.evaluations;
    out[2] = s.terminal;
    out[3] = s.searches;
    out[4] = s.cache_hits;
    out[5] = s.w1;
    out[6] = s.w2;
    out[7] = s.d;
}
void dsearch_emul_game_info(void* h, uint32_t k, uint32_t* game_idx, uint32_t* winner, uint32_t* n_moves) {
    const sp::GameRecord& g = static_cast<EmulResult*>(h)->sh.records[k];
    *game_idx = g.game_idx;
    *winner = g.winner;
    *n_moves = static_cast<uint32_t>(g.moves.size());
}
void dsearch_emul_game_moves(void* h, uint32_t k, uint16_t* out) {
    const sp::GameRecord& g = static_cast<EmulResult*>(h)->sh.records[k];
    std::memcpy(out, g.moves.data(), g.moves.size() * sizeof(uint16_t));
}
uint32_t dsearch_emul_entry(void* h, uint32_t k, uint32_t pos_idx, uint8_t* out, uint32_t cap, uint32_t* dir) {
    const sp::GameRecord& g = static_cast<EmulResult*>(h)->sh.records[k];
    const auto& b = g.entries[pos_idx];
    *dir = g.entry_dir[pos_idx];
    if (out && cap >= b.size()) std::memcpy(out, b.data(), b.size());
    return static_cast<uint32_t>(b.size());
}
void dsearch_emul_free(void* h) { delete static_cast<EmulResult*>(h); }
}
