// CUDA side of the device-resident self-play: the three kernels that wrap dsearch_core.hpp (one warp per game), the
// captured "wave" graph  [zero counters] -> begin -> select -> evaluator kernels -> expand,  and the transport the host
// driver (dsearch_host.hpp) talks to.  Included at the end of engine.cu (single translation unit with the evaluator).
//
// Data path of one simulation, all in HBM: select writes the leaf's record into the evaluator lane's DEVICE input block
// (Engine::ResidentIo) and bumps its row counter; the evaluator's kernels read the row count from that block (as they
// always do) and leave values / probabilities in d_values / d_probs; expand reads them.  The host sees one small command
// block per wave going in (new games / chosen moves + Dirichlet samples) and the finished searches' root visit counts
// coming out through mapped pinned memory.
#pragma once

#include "dsearch_api.hpp"
#include "dsearch_host.hpp"
#include "engine.hpp"

namespace cb2 {

template <class Rules>
__global__ void __launch_bounds__(128) ds_begin_kernel(ds::Params<Rules> p) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(p.cmds);
    const uint32_t n_cmds = hdr[0], wave = hdr[1];
    if (warp >= n_cmds) return;
    decltype(auto) R = ds::RulesRef<Rules>::get(p.rules);
    ds::Core<Rules>::begin_slot(R, p, warp, wave);
}

template <class Rules>
__global__ void __launch_bounds__(128) ds_select_kernel(ds::Params<Rules> p) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= p.n_slots) return;
    const uint32_t wave = reinterpret_cast<const uint32_t*>(p.cmds)[1];
    decltype(auto) R = ds::RulesRef<Rules>::get(p.rules);
    ds::Core<Rules>::select_slot(R, p, warp, wave);
}

template <class Rules>
__global__ void __launch_bounds__(128) ds_expand_kernel(ds::Params<Rules> p) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= p.n_slots) return;
    const uint32_t wave = reinterpret_cast<const uint32_t*>(p.cmds)[1];
    decltype(auto) R = ds::RulesRef<Rules>::get(p.rules);
    ds::Core<Rules>::expand_slot(R, p, warp, wave);
}

template <class Rules>
class DsCudaBackend {
  public:
    DsCudaBackend(const void* rules_blob, size_t rules_bytes, uint32_t max_children, Engine* e1, Engine* e2, const cattus_b200_selfplay_cfg& cfg,
                  const sp::Params params[2], uint32_t n_slots)
        : max_children_(max_children) {
        eng_[0] = e1;
        eng_[1] = e2;
        n_evals_ = e2 ? 2 : 1;
        if (e2 && e2->device() != e1->device()) throw Error(CATTUS_B200_EINVAL, "device search: both models must live on the same device");
        CB2_CUDA(cudaSetDevice(e1->device()));
        depth_ = cfg.device_waves_in_flight ? std::min<uint32_t>(cfg.device_waves_in_flight, 8) : 2;
        n_bufs_ = depth_ + 1;
        std::memset(&p_, 0, sizeof(p_));
        try {
            for (uint32_t e = 0; e < n_evals_; ++e) {
                io_[e] = eng_[e]->resident_acquire();
                if (io_[e].max_batch < n_slots)
                    throw Error(CATTUS_B200_ERANGE, "device search: device_games (" + std::to_string(n_slots) + ") exceeds the evaluator's max_batch (" +
                                                        std::to_string(io_[e].max_batch) + ")");
                if (io_[e].moves < max_children && !Rules::kChess) throw Error(CATTUS_B200_EINVAL, "device search: the model's move count does not fit the game");
            }
            using T = ds::TreeOps<Rules>;
            const uint32_t typ = Rules::kChess ? 48u : max_children;
            const uint64_t per_search = (static_cast<uint64_t>(std::max(params[0].sim_num, params[1].sim_num)) + 8u) * T::block_words(static_cast<int>(typ));
            uint64_t want = cfg.device_tree_kwords ? static_cast<uint64_t>(cfg.device_tree_kwords) * 1024u : 3u * per_search;
            size_t free_b = 0, total_b = 0;
            CB2_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const uint64_t budget = static_cast<uint64_t>(static_cast<double>(free_b) * 0.7) / (static_cast<uint64_t>(n_slots) * 3u * 4u);
            if (!cfg.device_tree_kwords) want = std::min(want, budget);
            want &= ~static_cast<uint64_t>(3);
            if (want > budget || want < per_search + per_search / 4 || want >= (static_cast<uint64_t>(0xFFFFFE) << 2))
                throw Error(CATTUS_B200_ENOMEM, "device search: " + std::to_string(n_slots) + " games x 3 tree buffers of " + std::to_string(want) +
                                                    " words do not fit the device (or its 2^26-word tree address space); lower device_games");
            pool_words_ = static_cast<uint32_t>(want);

            alloc(d_rules_, std::max<size_t>(rules_bytes, 16));
            CB2_CUDA(cudaMemcpy(d_rules_, rules_blob, rules_bytes, cudaMemcpyHostToDevice));
            alloc(d_slots_, sizeof(ds::SlotState) * n_slots);
            CB2_CUDA(cudaMemset(d_slots_, 0, sizeof(ds::SlotState) * n_slots));
            alloc(d_pools_, static_cast<size_t>(n_slots) * 3u * pool_words_ * 4u + 1024u);  // + slack: select fetches a node's first 32 child rows before it knows the count
            const uint32_t path_cap = Rules::kChess ? 256u : max_children + 2u;
            alloc(d_paths_, sizeof(ds::PathStep) * static_cast<size_t>(n_slots) * path_cap);
            alloc(d_noise_, sizeof(float) * static_cast<size_t>(n_slots) * max_children);
            const uint32_t hist_cap = Rules::kChess ? 128u : 1u;
            alloc(d_hist_, sizeof(typename Rules::Pos) * static_cast<size_t>(n_slots) * hist_cap);
            // ValueFuncCache in HBM, one per evaluator (cfg.cache_size entries, rounded up to whole buckets)
            for (uint32_t e = 0; e < 2; ++e) p_.cache[e] = ds::CacheIo{};
            if (cfg.cache_size) {
                uint32_t buckets = 1;
                while (static_cast<uint64_t>(buckets) * ds::kCacheWays < cfg.cache_size) buckets <<= 1;
                for (uint32_t e = 0; e < n_evals_; ++e) {
                    ds::CacheIo& c = p_.cache[e];
                    c.entry_bytes = (40u + 4u * max_children + 15u) & ~15u;
                    alloc(d_cache_meta_[e], static_cast<size_t>(buckets) * 32u);
                    CB2_CUDA(cudaMemset(d_cache_meta_[e], 0, static_cast<size_t>(buckets) * 32u));
                    alloc(d_cache_entries_[e], static_cast<size_t>(buckets) * ds::kCacheWays * c.entry_bytes);
                    c.meta = static_cast<uint32_t*>(d_cache_meta_[e]);
                    c.entries = static_cast<uint8_t*>(d_cache_entries_[e]);
                    c.bucket_mask = buckets - 1;
                    c.enabled = 1;
                }
            }
            cmd_stride_ = ds::cmd_stride_for(max_children);
            result_stride_ = ds::result_stride_for(max_children);
            cmd_bytes_ = 16 + static_cast<size_t>(n_slots) * cmd_stride_;
            alloc(d_cmds_, cmd_bytes_);
            CB2_CUDA(cudaMemset(d_cmds_, 0, 16));
            alloc(d_status_, 64);
            CB2_CUDA(cudaMemset(d_status_, 0, 64));
            result_buf_bytes_ = static_cast<size_t>(n_slots) * result_stride_;
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_results_), result_buf_bytes_ * n_bufs_, cudaHostAllocMapped));
            uint8_t* d_results = nullptr;
            CB2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_results), h_results_, 0));
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_cmds_), cmd_bytes_ * n_bufs_, cudaHostAllocDefault));
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_status_), 64 * n_bufs_, cudaHostAllocDefault));
            std::memset(h_status_, 0, 64 * n_bufs_);
            CB2_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
            events_.resize(n_bufs_);
            for (auto& ev : events_) CB2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

            p_.rules = d_rules_;
            p_.slots = static_cast<ds::SlotState*>(d_slots_);
            p_.n_slots = n_slots;
            p_.pools = static_cast<uint32_t*>(d_pools_);
            p_.pool_words = pool_words_;
            p_.paths = static_cast<ds::PathStep*>(d_paths_);
            p_.path_cap = path_cap;
            p_.noise = static_cast<float*>(d_noise_);
            p_.max_children = max_children;
            p_.hist = static_cast<typename Rules::Pos*>(d_hist_);
            p_.hist_cap = hist_cap;
            p_.n_evals = n_evals_;
            for (uint32_t e = 0; e < 2; ++e) {
                const Engine::ResidentIo& io = io_[e < n_evals_ ? e : 0];
                p_.eval[e].n_ptr = reinterpret_cast<uint32_t*>(io.d_block);
                p_.eval[e].recs = io.d_block + 16 + 8;
                p_.eval[e].values = io.d_values;
                p_.eval[e].probs = io.d_probs;
                p_.eval[e].rec_bytes = io.rec_bytes;
                p_.eval[e].prob_stride = io.moves < max_children ? io.moves : max_children;
                p_.eval[e].max_rows = io.max_batch;
                p_.eval[e].plane_words = io.plane_words;
                p_.sim_num[e] = params[e].sim_num;
                p_.explore[e] = params[e].explore_factor;
                p_.noise_eps[e] = params[e].noise_eps;
            }
            p_.cmds = static_cast<const uint8_t*>(d_cmds_);
            p_.cmd_stride = cmd_stride_;
            p_.results = d_results;
            p_.result_stride = result_stride_;
            p_.n_result_bufs = n_bufs_;
            p_.result_buf_bytes = result_buf_bytes_;
            p_.done_count = static_cast<uint32_t*>(d_status_);
            p_.error = static_cast<uint32_t*>(d_status_) + 1;
            p_.counters = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(d_status_) + 16);
            // begin (per-move commands: tree reuse copies subtrees) runs on a side branch of the wave graph, beside select /
            // evaluator / expand; its slots take part from the next wave on
            p_.begin_lead = std::getenv("CATTUS_B200_DSEARCH_SERIAL_BEGIN") ? 0u : 1u;
            p_.visit_budget = 24;  // measured on hex5 (sim_num 1400, whole games): 24 -> 57.1 M sims/s, 48 -> 54.4 M, 96 -> 49.4 M
            if (const char* vb = std::getenv("CATTUS_B200_DSEARCH_VISIT_BUDGET")) p_.visit_budget = std::max(1, std::atoi(vb));
            if (p_.begin_lead) {
                CB2_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
                CB2_CUDA(cudaEventCreateWithFlags(&fork_, cudaEventDisableTiming));
                CB2_CUDA(cudaEventCreateWithFlags(&join_, cudaEventDisableTiming));
            }
            capture();
        } catch (...) {
            destroy();
            throw;
        }
    }
    ~DsCudaBackend() { destroy(); }
    DsCudaBackend(const DsCudaBackend&) = delete;
    DsCudaBackend& operator=(const DsCudaBackend&) = delete;

    uint32_t n_slots() const { return p_.n_slots; }
    uint32_t max_children() const { return max_children_; }
    uint32_t depth() const { return depth_; }
    uint32_t pool_words() const { return pool_words_; }
    uint8_t* cmd_block(uint32_t wave) { return h_cmds_ + static_cast<size_t>(wave % n_bufs_) * cmd_bytes_; }

    void submit(uint32_t wave, uint32_t n_cmds) {
        const uint32_t b = wave % n_bufs_;
        CB2_CUDA(cudaMemcpyAsync(d_cmds_, cmd_block(wave), 16 + static_cast<size_t>(n_cmds) * cmd_stride_, cudaMemcpyHostToDevice, stream_));
        CB2_CUDA(cudaGraphLaunch(graph_, stream_));
        CB2_CUDA(cudaMemcpyAsync(h_status_ + 64 * b, d_status_, 64, cudaMemcpyDeviceToHost, stream_));
        CB2_CUDA(cudaEventRecord(events_[b], stream_));
        waves_ += 1;
    }
    const uint8_t* wait(uint32_t wave, uint32_t* n_done) {
        const uint32_t b = wave % n_bufs_;
        const cudaError_t e = cudaEventSynchronize(events_[b]);
        if (e != cudaSuccess) throw Error(CATTUS_B200_ECUDA, std::string("device search wave failed: ") + cudaGetErrorString(e));
        const uint32_t* st = reinterpret_cast<const uint32_t*>(h_status_ + 64 * b);
        if (st[1]) {
            std::string what;
            if (st[1] & ds::kErrPool) what += " tree pool exhausted (raise device_tree_kwords or lower device_games);";
            if (st[1] & ds::kErrPath) what += " search path deeper than the path buffer;";
            if (st[1] & ds::kErrRows) what += " evaluator batch overflow;";
            if (st[1] & ds::kErrNoise) what += " noise sample size does not match the root;";
            throw Error(CATTUS_B200_ERANGE, "device search:" + what);
        }
        *n_done = st[0];
        std::memcpy(last_counters_, h_status_ + 64 * b + 16, sizeof(last_counters_));
        return h_results_ + static_cast<size_t>(b) * result_buf_bytes_;
    }
    void read_counters(unsigned long long out[4]) {
        CB2_CUDA(cudaStreamSynchronize(stream_));
        CB2_CUDA(cudaMemcpy(out, static_cast<uint8_t*>(d_status_) + 16, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    uint64_t waves() const { return waves_; }
    uint32_t kernels_per_wave() const { return kernels_per_wave_; }

  private:
    void alloc(void*& ptr, size_t bytes) {
        const cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e != cudaSuccess) {
            ptr = nullptr;
            throw Error(CATTUS_B200_ENOMEM, "device search: cudaMalloc(" + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        }
    }
    void capture() {
        const uint32_t blocks = (p_.n_slots + 3u) / 4u;
        cudaGraph_t g = nullptr;
        CB2_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        try {
            for (uint32_t e = 0; e < n_evals_; ++e) CB2_CUDA(cudaMemsetAsync(io_[e].d_block, 0, 4, stream_));
            CB2_CUDA(cudaMemsetAsync(d_status_, 0, 4, stream_));
            if (p_.begin_lead) {
                CB2_CUDA(cudaEventRecord(fork_, stream_));
                CB2_CUDA(cudaStreamWaitEvent(side_, fork_, 0));
                ds_begin_kernel<Rules><<<blocks, 128, 0, side_>>>(p_);
                CB2_CUDA(cudaEventRecord(join_, side_));
            } else {
                ds_begin_kernel<Rules><<<blocks, 128, 0, stream_>>>(p_);
            }
            ds_select_kernel<Rules><<<blocks, 128, 0, stream_>>>(p_);
            for (uint32_t e = 0; e < n_evals_; ++e) eng_[e]->resident_enqueue(io_[e].lane, stream_);
            ds_expand_kernel<Rules><<<blocks, 128, 0, stream_>>>(p_);
            if (p_.begin_lead) CB2_CUDA(cudaStreamWaitEvent(stream_, join_, 0));
            CB2_CUDA(cudaGetLastError());
        } catch (...) {
            cudaStreamEndCapture(stream_, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        cudaError_t ce = cudaStreamEndCapture(stream_, &g);
        if (ce != cudaSuccess) throw Error(CATTUS_B200_ECUDA, std::string("device search: graph capture failed: ") + cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&graph_, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) throw Error(CATTUS_B200_ECUDA, std::string("device search: graph instantiate failed: ") + cudaGetErrorString(ce));
        kernels_per_wave_ = 3;
        for (uint32_t e = 0; e < n_evals_; ++e) kernels_per_wave_ += io_[e].kernels;
    }
    void destroy() {
        if (stream_) cudaStreamSynchronize(stream_);
        if (graph_) cudaGraphExecDestroy(graph_);
        graph_ = nullptr;
        for (auto& ev : events_)
            if (ev) cudaEventDestroy(ev);
        events_.clear();
        if (stream_) cudaStreamDestroy(stream_);
        stream_ = nullptr;
        if (side_) cudaStreamDestroy(side_);
        side_ = nullptr;
        if (fork_) cudaEventDestroy(fork_);
        if (join_) cudaEventDestroy(join_);
        fork_ = join_ = nullptr;
        for (void** q : {&d_rules_, &d_slots_, &d_pools_, &d_paths_, &d_noise_, &d_hist_, &d_cmds_, &d_status_, &d_cache_meta_[0], &d_cache_meta_[1],
                         &d_cache_entries_[0], &d_cache_entries_[1]}) {
            if (*q) cudaFree(*q);
            *q = nullptr;
        }
        if (h_results_) cudaFreeHost(h_results_);
        if (h_cmds_) cudaFreeHost(h_cmds_);
        if (h_status_) cudaFreeHost(h_status_);
        h_results_ = h_cmds_ = h_status_ = nullptr;
        for (uint32_t e = 0; e < n_evals_; ++e)
            if (io_[e].lane >= 0) {
                eng_[e]->resident_release(io_[e].lane);
                io_[e].lane = -1;
            }
    }

    ds::Params<Rules> p_;
    Engine* eng_[2] = {nullptr, nullptr};
    Engine::ResidentIo io_[2];
    uint32_t n_evals_ = 1, depth_ = 2, n_bufs_ = 3, max_children_ = 0, pool_words_ = 0, cmd_stride_ = 0, result_stride_ = 0, kernels_per_wave_ = 0;
    size_t cmd_bytes_ = 0, result_buf_bytes_ = 0;
    void *d_rules_ = nullptr, *d_slots_ = nullptr, *d_pools_ = nullptr, *d_paths_ = nullptr, *d_noise_ = nullptr, *d_hist_ = nullptr, *d_cmds_ = nullptr,
         *d_status_ = nullptr, *d_cache_meta_[2] = {nullptr, nullptr}, *d_cache_entries_[2] = {nullptr, nullptr};
    uint8_t *h_results_ = nullptr, *h_cmds_ = nullptr, *h_status_ = nullptr;
    cudaStream_t stream_ = nullptr, side_ = nullptr;
    cudaEvent_t fork_ = nullptr, join_ = nullptr;
    std::vector<cudaEvent_t> events_;
    cudaGraphExec_t graph_ = nullptr;
    uint64_t waves_ = 0;
    unsigned long long last_counters_[4] = {0, 0, 0, 0};
};

template <class Rules>
static void dsearch_run_rules(const Rules& rules, const void* blob, size_t blob_bytes, uint32_t max_children, Engine* e1, Engine* e2,
                              const cattus_b200_selfplay_cfg& cfg, const sp::Params params[2], sp::Shared& sh) {
    const uint32_t stride = std::max<uint32_t>(1, cfg.game_stride);
    const uint32_t my_games = cfg.games_num > cfg.first_game ? (cfg.games_num - cfg.first_game + stride - 1) / stride : 0;
    if (my_games == 0) return;
    const uint32_t n_slots = std::max<uint32_t>(1, std::min<uint32_t>(cfg.device_games, my_games));
    const auto t0 = std::chrono::steady_clock::now();
    DsCudaBackend<Rules> be(blob, blob_bytes, max_children, e1, e2, cfg, params, n_slots);
    ds::Driver<Rules, DsCudaBackend<Rules>> drv(rules, cfg, params, be, sh);
    drv.run();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const uint64_t waves = be.waves();
    e1->note_resident(waves, sh.evaluations, waves * be.kernels_per_wave(), waves ? secs / static_cast<double>(waves) : 0.0);
}

void dsearch_run(cattus_b200_t* m1, cattus_b200_t* m2, const cattus_b200_selfplay_cfg& cfg, const sp::Params params[2], sp::Shared& sh) {
    try {
        if (!m1) throw Error(CATTUS_B200_EINVAL, "null evaluator handle (there is no CPU fallback)");
        Engine* e1 = m1->engine;
        Engine* e2 = (m2 && m2 != m1) ? m2->engine : nullptr;
        if (cfg.game == CATTUS_B200_GAME_HEX) {
            const int s = static_cast<int>(cfg.board_size);
            if (s <= 8) {
                sp::HexRulesT<uint64_t> rules(s);
                dsearch_run_rules(rules, &rules, sizeof(rules), static_cast<uint32_t>(s * s), e1, e2, cfg, params, sh);
            } else {
                sp::HexRulesT<sp::u128> rules(s);
                dsearch_run_rules(rules, &rules, sizeof(rules), static_cast<uint32_t>(s * s), e1, e2, cfg, params, sh);
            }
        } else if (cfg.game == CATTUS_B200_GAME_CHESS) {
            sp::ChessRules rules;
            dsearch_run_rules(rules, &sp::chess_tables(), sizeof(sp::ChessTables), static_cast<uint32_t>(sp::ChessRules::kMaxMoves), e1, e2, cfg, params, sh);
        } else {
            sp::TttRules rules;
            dsearch_run_rules(rules, &rules, sizeof(rules), 9u, e1, e2, cfg, params, sh);
        }
    } catch (const Error& e) {
        throw sp::SpError{e.code, e.what()};
    }
}

}  // namespace cb2
