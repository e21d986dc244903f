// Host-only throughput of the self-play driver: csrc/selfplay.cpp bound to a trivial evaluator (hash-derived priors), so
// the time measured is the MCTS / rules / cache work of the worker threads alone.  Exploration tool, not part of the
// library:
//   g++ -O3 -std=c++17 -ffp-contract=off -fno-trapping-math -fno-strict-aliasing -pthread tools/selfplay_host_bench.cpp
//       cattus_b200/csrc/selfplay.cpp -o /tmp/selfplay_host_bench
//   /tmp/selfplay_host_bench chess 600 1 256 16   (game, sim_num, threads, games per thread, max_moves[, games]; CACHE=n entries)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../include/cattus_b200_selfplay.h"

// the evaluator entry points selfplay.cpp links against; unused here (the driver is bound to fake_eval)
extern "C" {
int cattus_b200_eval(cattus_b200_t*, const uint64_t*, const uint8_t*, float*, uint32_t, uint32_t*, float*) { return CATTUS_B200_ENODEV; }
int cattus_b200_eval_batch(cattus_b200_t*, const uint64_t*, const uint8_t*, uint32_t, float*, size_t, uint32_t*, float*) { return CATTUS_B200_ENODEV; }
int cattus_b200_eval_batch_submit(cattus_b200_t*, const uint64_t*, const uint8_t*, uint32_t, int, int32_t*) { return CATTUS_B200_ENODEV; }
int cattus_b200_eval_batch_wait(cattus_b200_t*, int32_t, float*, size_t, uint32_t*, float*) { return CATTUS_B200_ENODEV; }
int cattus_b200_get_info(const cattus_b200_t*, cattus_b200_info*) { return CATTUS_B200_ENODEV; }
const char* cattus_b200_last_error(void) { return "host bench: no evaluator"; }
}

struct Ctx {
    int words;       // u64 words per position
    int legal_bytes; // 0: hex / ttt (legal = empty cells of planes 0 | 1 within plane 2)
    int wpp;
};

static inline uint64_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    return x;
}

static int fake_eval(void* vctx, const uint64_t* planes, const uint8_t* legal, uint32_t n, float* probs, size_t cap, uint32_t* offsets, float* values) {
    const Ctx& c = *static_cast<const Ctx*>(vctx);
    uint32_t off = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t* w = planes + static_cast<size_t>(i) * c.words;
        uint64_t h = 0;
        for (int k = 0; k < c.words; ++k) h = mix(h + w[k] + 0x9E3779B97F4A7C15ull);
        int cnt = 0;
        if (c.legal_bytes) {
            const uint8_t* bm = legal + static_cast<size_t>(i) * c.legal_bytes;
            for (int k = 0; k < c.legal_bytes; ++k) cnt += __builtin_popcount(bm[k]);
        } else {
            for (int k = 0; k < c.wpp; ++k) cnt += __builtin_popcountll(w[2 * c.wpp + k] & ~(w[k] | w[c.wpp + k]));
        }
        if (off + cnt > cap) return CATTUS_B200_ERANGE;
        offsets[i] = off;
        float tot = 0.0f;
        for (int k = 0; k < cnt; ++k) {
            const float x = 1.0f + static_cast<float>(mix(h + k) & 1023) * (1.0f / 256.0f);
            probs[off + k] = x;
            tot += x;
        }
        for (int k = 0; k < cnt; ++k) probs[off + k] /= tot;
        values[i] = static_cast<float>(static_cast<int>(h >> 40 & 2047) - 1024) * (1.0f / 1024.0f);
        off += cnt;
    }
    offsets[n] = off;
    return 0;
}

int main(int argc, char** argv) {
    const std::string game = argc > 1 ? argv[1] : "chess";
    cattus_b200_selfplay_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    Ctx ctx{};
    if (game == "chess") {
        cfg.game = CATTUS_B200_GAME_CHESS;
        cfg.board_size = 8;
        ctx = Ctx{18, 235, 1};
    } else {
        cfg.game = CATTUS_B200_GAME_HEX;
        cfg.board_size = static_cast<uint32_t>(std::atoi(game.c_str() + 3));
        const int wpp = static_cast<int>((cfg.board_size * cfg.board_size + 63) / 64);
        ctx = Ctx{3 * wpp, 0, wpp};
    }
    cfg.sim_num = argc > 2 ? std::atoi(argv[2]) : 600;
    cfg.threads = argc > 3 ? std::atoi(argv[3]) : 1;
    cfg.games_per_thread = argc > 4 ? std::atoi(argv[4]) : 256;
    cfg.max_moves = argc > 5 ? std::atoi(argv[5]) : 16;
    cfg.games_num = argc > 6 ? std::atoi(argv[6]) : cfg.threads * cfg.games_per_thread;
    cfg.explore_factor = 1.41421f;
    const uint32_t tm[2] = {30, 9999};
    const float tv[2] = {1.0f, 0.0f};
    cfg.temperature_moves = tm;
    cfg.temperature_values = tv;
    cfg.n_temperatures = 2;
    cfg.prior_noise_alpha = 0.03f;
    cfg.prior_noise_epsilon = 0.25f;
    cfg.cache_size = std::getenv("CACHE") ? std::atoi(std::getenv("CACHE")) : 1000000;
    cfg.game_stride = 1;
    cfg.seed = 1;
    cattus_b200_selfplay_t* res = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = cattus_b200_selfplay_run_with(fake_eval, &ctx, nullptr, nullptr, &cfg, &res);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != 0) {
        std::fprintf(stderr, "error %d: %s\n", rc, cattus_b200_selfplay_last_error());
        return 1;
    }
    cattus_b200_selfplay_summary s;
    cattus_b200_selfplay_summary_get(res, &s);
    std::printf("%s sim_num %u threads %u gpt %u: %.3f s, %.3f M sims/s (%.0f ns per simulation and thread), evaluations %llu, cache hits %llu, terminal %llu, eval wait %.1f %%\n",
                game.c_str(), cfg.sim_num, cfg.threads, cfg.games_per_thread, secs, s.simulations / secs * 1e-6, secs * cfg.threads / s.simulations * 1e9,
                static_cast<unsigned long long>(s.evaluations), static_cast<unsigned long long>(s.cache_hits), static_cast<unsigned long long>(s.terminal_leaves),
                100.0 * s.eval_wait_seconds / (secs * cfg.threads));
    cattus_b200_selfplay_free(res);
    return 0;
}
