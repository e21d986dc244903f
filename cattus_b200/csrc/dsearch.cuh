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

#ifndef DS_SELECT_MIN_BLOCKS
#define DS_SELECT_MIN_BLOCKS 1
#endif
#ifndef DS_EXPAND_MIN_BLOCKS
#define DS_EXPAND_MIN_BLOCKS 1
#endif

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
__global__ void __launch_bounds__(128, DS_SELECT_MIN_BLOCKS) ds_select_kernel(ds::Params<Rules> p) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= p.n_slots) return;
    const uint32_t wave = reinterpret_cast<const uint32_t*>(p.cmds)[1];
    decltype(auto) R = ds::RulesRef<Rules>::get(p.rules);
    ds::Core<Rules>::select_slot(R, p, warp, wave);
}

template <class Rules>
__global__ void __launch_bounds__(128, DS_EXPAND_MIN_BLOCKS) ds_expand_kernel(ds::Params<Rules> p) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= p.n_slots) return;
    const uint32_t wave = reinterpret_cast<const uint32_t*>(p.cmds)[1];
    decltype(auto) R = ds::RulesRef<Rules>::get(p.rules);
    ds::Core<Rules>::expand_slot(R, p, warp, wave);
}

// One population of slots: its kernel parameters, streams, captured wave graph, command and status blocks, evaluator lanes
// and evaluation caches.  Two populations alternate waves on two streams (dsearch_host.hpp), so that one population's
// select / expand kernels run beside the other's evaluator kernels.
template <class Rules>
struct DsPopulation {
    ds::Params<Rules> p;
    Engine::ResidentIo io[2];
    cudaStream_t stream = nullptr, side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaGraphExec_t graph = nullptr;
    void *d_cmds = nullptr, *d_status = nullptr, *d_cache_meta[2] = {nullptr, nullptr}, *d_cache_entries[2] = {nullptr, nullptr};
};

template <class Rules>
class DsCudaBackend {
  public:
    DsCudaBackend(const void* rules_blob, size_t rules_bytes, uint32_t max_children, Engine* e1, Engine* e2, const cattus_b200_selfplay_cfg& cfg,
                  const sp::Params params[2], uint32_t n_slots)
        : max_children_(max_children), n_slots_(n_slots) {
        eng_[0] = e1;
        eng_[1] = e2;
        n_evals_ = e2 ? 2 : 1;
        if (e2 && e2->device() != e1->device()) throw Error(CATTUS_B200_EINVAL, "device search: both models must live on the same device");
        CB2_CUDA(cudaSetDevice(e1->device()));
        depth_ = cfg.device_waves_in_flight ? std::min<uint32_t>(cfg.device_waves_in_flight, 8) : 2;
        n_bufs_ = depth_ + 1;
        wave_pop_.assign(n_bufs_, 0);
        const auto t0 = std::chrono::steady_clock::now();
        try {
            // evaluator lanes: one per population and evaluator; a second population needs a second free stream in every evaluator
            for (uint32_t e = 0; e < n_evals_; ++e) {
                pop_[0].io[e] = eng_[e]->resident_acquire(true);
                if (pop_[0].io[e].max_batch < n_slots)
                    throw Error(CATTUS_B200_ERANGE, "device search: device_games (" + std::to_string(n_slots) + ") exceeds the evaluator's max_batch (" +
                                                        std::to_string(pop_[0].io[e].max_batch) + ")");
                if (pop_[0].io[e].moves < max_children && !Rules::kChess) throw Error(CATTUS_B200_EINVAL, "device search: the model's move count does not fit the game");
            }
            n_pops_ = 1;
            // Two populations taking waves in turn on two streams were measured and do NOT pay on a B200: the persistent trunk
            // kernels leave no room for the other population's select / expand blocks to run beside them, and half-size batches
            // are less efficient (hex5 64.1 M vs 64.4 M sims/s, hex7 30.0 M vs 37.1 M, chess 10x128 4.7 M vs 5.0 M).  Kept as an
            // experiment knob; the games are the same either way (tests/test_gpu_dsearch.py).
            if (n_slots >= 64 && depth_ >= 2 && std::getenv("CATTUS_B200_DSEARCH_TWO_POPULATIONS")) {
                bool ok = true;
                for (uint32_t e = 0; e < n_evals_; ++e) {
                    pop_[1].io[e] = eng_[e]->resident_acquire(false);
                    ok = ok && pop_[1].io[e].lane >= 0;
                }
                if (ok) {
                    n_pops_ = 2;
                } else {
                    for (uint32_t e = 0; e < n_evals_; ++e)
                        if (pop_[1].io[e].lane >= 0) {
                            eng_[e]->resident_release(pop_[1].io[e].lane);
                            pop_[1].io[e].lane = -1;
                        }
                }
            }
            split_ = n_pops_ == 2 ? (n_slots + 1) / 2 : n_slots;

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
            cmd_stride_ = ds::cmd_stride_for(max_children);
            result_stride_ = ds::result_stride_for(max_children);
            cmd_bytes_ = 16 + static_cast<size_t>(n_slots) * cmd_stride_;
            result_buf_bytes_ = static_cast<size_t>(n_slots) * result_stride_;
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_results_), result_buf_bytes_ * n_bufs_, cudaHostAllocMapped));
            uint8_t* d_results = nullptr;
            CB2_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_results), h_results_, 0));
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_cmds_), cmd_bytes_ * n_bufs_, cudaHostAllocDefault));
            CB2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_status_), 64 * n_bufs_, cudaHostAllocDefault));
            std::memset(h_status_, 0, 64 * n_bufs_);
            events_.resize(n_bufs_);
            for (auto& ev : events_) CB2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            const uint32_t begin_lead = std::getenv("CATTUS_B200_DSEARCH_SERIAL_BEGIN") ? 0u : 1u;
            uint32_t visit_budget = 24;  // measured on hex5 (sim_num 1400, whole games): 24 -> 57.1 M sims/s, 48 -> 54.4 M, 96 -> 49.4 M
            if (const char* vb = std::getenv("CATTUS_B200_DSEARCH_VISIT_BUDGET")) visit_budget = static_cast<uint32_t>(std::max(1, std::atoi(vb)));

            for (uint32_t k = 0; k < n_pops_; ++k) {
                DsPopulation<Rules>& P = pop_[k];
                const uint32_t base = k == 0 ? 0 : split_, cnt = k == 0 ? split_ : n_slots - split_;
                ds::Params<Rules>& p = P.p;
                std::memset(&p, 0, sizeof(p));
                CB2_CUDA(cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking));
                alloc(P.d_cmds, cmd_bytes_);
                CB2_CUDA(cudaMemset(P.d_cmds, 0, 16));
                alloc(P.d_status, 64);
                CB2_CUDA(cudaMemset(P.d_status, 0, 64));
                p.rules = d_rules_;
                p.slots = static_cast<ds::SlotState*>(d_slots_) + base;
                p.n_slots = cnt;
                p.slot_base = base;
                p.pools = static_cast<uint32_t*>(d_pools_) + static_cast<size_t>(base) * 3u * pool_words_;
                p.pool_words = pool_words_;
                p.paths = static_cast<ds::PathStep*>(d_paths_) + static_cast<size_t>(base) * path_cap;
                p.path_cap = path_cap;
                p.noise = static_cast<float*>(d_noise_) + static_cast<size_t>(base) * max_children;
                p.max_children = max_children;
                p.hist = static_cast<typename Rules::Pos*>(d_hist_) + static_cast<size_t>(base) * hist_cap;
                p.hist_cap = hist_cap;
                p.n_evals = n_evals_;
                for (uint32_t e = 0; e < 2; ++e) {
                    const Engine::ResidentIo& io = P.io[e < n_evals_ ? e : 0];
                    p.eval[e].n_ptr = reinterpret_cast<uint32_t*>(io.d_block);
                    p.eval[e].recs = io.d_block + 16 + 8;
                    p.eval[e].values = io.d_values;
                    p.eval[e].probs = io.d_probs;
                    p.eval[e].rec_bytes = io.rec_bytes;
                    p.eval[e].prob_stride = io.moves < max_children ? io.moves : max_children;
                    p.eval[e].max_rows = io.max_batch;
                    p.eval[e].plane_words = io.plane_words;
                    p.sim_num[e] = params[e].sim_num;
                    p.explore[e] = params[e].explore_factor;
                    p.noise_eps[e] = params[e].noise_eps;
                    p.cache[e] = ds::CacheIo{};
                }
                // ValueFuncCache in HBM, one per evaluator and population (cfg.cache_size entries over the populations, rounded up
                // to whole buckets): a population's probes (select) and inserts (expand) never overlap, two populations' would
                if (cfg.cache_size) {
                    uint32_t buckets = 1;
                    while (static_cast<uint64_t>(buckets) * ds::kCacheWays * n_pops_ < cfg.cache_size) buckets <<= 1;
                    for (uint32_t e = 0; e < n_evals_; ++e) {
                        ds::CacheIo& c = p.cache[e];
                        c.entry_bytes = (40u + 4u * max_children + 15u) & ~15u;
                        alloc(P.d_cache_meta[e], static_cast<size_t>(buckets) * 32u);
                        CB2_CUDA(cudaMemset(P.d_cache_meta[e], 0, static_cast<size_t>(buckets) * 32u));
                        alloc(P.d_cache_entries[e], static_cast<size_t>(buckets) * ds::kCacheWays * c.entry_bytes);
                        c.meta = static_cast<uint32_t*>(P.d_cache_meta[e]);
                        c.entries = static_cast<uint8_t*>(P.d_cache_entries[e]);
                        c.bucket_mask = buckets - 1;
                        c.enabled = 1;
                    }
                }
                p.cmds = static_cast<const uint8_t*>(P.d_cmds);
                p.cmd_stride = cmd_stride_;
                p.results = d_results;
                p.result_stride = result_stride_;
                p.n_result_bufs = n_bufs_;
                p.result_buf_bytes = result_buf_bytes_;
                p.done_count = static_cast<uint32_t*>(P.d_status);
                p.error = static_cast<uint32_t*>(P.d_status) + 1;
                p.max_used = static_cast<uint32_t*>(P.d_status) + 2;
                p.counters = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(P.d_status) + 16);
                // begin (per-move commands: tree reuse copies subtrees) runs on a side branch of the wave graph, beside select /
                // evaluator / expand; its slots take part from the population's next wave on
                p.begin_lead = begin_lead;
                p.visit_budget = visit_budget;
                if (begin_lead) {
                    CB2_CUDA(cudaStreamCreateWithFlags(&P.side, cudaStreamNonBlocking));
                    CB2_CUDA(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
                    CB2_CUDA(cudaEventCreateWithFlags(&P.join, cudaEventDisableTiming));
                }
                capture(P);
            }
            setup_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } catch (...) {
            destroy();
            throw;
        }
    }
    ~DsCudaBackend() { destroy(); }
    DsCudaBackend(const DsCudaBackend&) = delete;
    DsCudaBackend& operator=(const DsCudaBackend&) = delete;

    uint32_t n_slots() const { return n_slots_; }
    uint32_t max_children() const { return max_children_; }
    uint32_t depth() const { return depth_; }
    uint32_t pool_words() const { return pool_words_; }
    uint32_t populations() const { return n_pops_; }
    uint32_t population_of(uint32_t slot) const { return slot >= split_ ? 1u : 0u; }
    double setup_seconds() const { return setup_seconds_; }
    uint8_t* cmd_block(uint32_t wave) { return h_cmds_ + static_cast<size_t>(wave % n_bufs_) * cmd_bytes_; }

    void submit(uint32_t wave, uint32_t pop, uint32_t n_cmds) {
        const uint32_t b = wave % n_bufs_;
        DsPopulation<Rules>& P = pop_[pop];
        wave_pop_[b] = pop;
        CB2_CUDA(cudaMemcpyAsync(P.d_cmds, cmd_block(wave), 16 + static_cast<size_t>(n_cmds) * cmd_stride_, cudaMemcpyHostToDevice, P.stream));
        CB2_CUDA(cudaGraphLaunch(P.graph, P.stream));
        CB2_CUDA(cudaMemcpyAsync(h_status_ + 64 * b, P.d_status, 64, cudaMemcpyDeviceToHost, P.stream));
        CB2_CUDA(cudaEventRecord(events_[b], P.stream));
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
        return h_results_ + static_cast<size_t>(b) * result_buf_bytes_;
    }
    void read_counters(unsigned long long out[4]) {
        for (int i = 0; i < 4; ++i) out[i] = 0;
        for (uint32_t k = 0; k < n_pops_; ++k) {
            unsigned long long c[4];
            uint32_t hw = 0;
            CB2_CUDA(cudaStreamSynchronize(pop_[k].stream));
            CB2_CUDA(cudaMemcpy(c, static_cast<uint8_t*>(pop_[k].d_status) + 16, sizeof(c), cudaMemcpyDeviceToHost));
            CB2_CUDA(cudaMemcpy(&hw, static_cast<uint8_t*>(pop_[k].d_status) + 8, sizeof(hw), cudaMemcpyDeviceToHost));
            for (int i = 0; i < 4; ++i) out[i] += c[i];
            max_used_ = std::max(max_used_, hw);
        }
    }
    uint32_t max_used() const { return max_used_; }
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
    void capture(DsPopulation<Rules>& P) {
        const uint32_t blocks = (P.p.n_slots + 3u) / 4u;
        cudaGraph_t g = nullptr;
        CB2_CUDA(cudaStreamBeginCapture(P.stream, cudaStreamCaptureModeThreadLocal));
        try {
            for (uint32_t e = 0; e < n_evals_; ++e) CB2_CUDA(cudaMemsetAsync(P.io[e].d_block, 0, 4, P.stream));
            CB2_CUDA(cudaMemsetAsync(P.d_status, 0, 4, P.stream));
            if (P.p.begin_lead) {
                CB2_CUDA(cudaEventRecord(P.fork, P.stream));
                CB2_CUDA(cudaStreamWaitEvent(P.side, P.fork, 0));
                ds_begin_kernel<Rules><<<blocks, 128, 0, P.side>>>(P.p);
                CB2_CUDA(cudaEventRecord(P.join, P.side));
            } else {
                ds_begin_kernel<Rules><<<blocks, 128, 0, P.stream>>>(P.p);
            }
            ds_select_kernel<Rules><<<blocks, 128, 0, P.stream>>>(P.p);
            for (uint32_t e = 0; e < n_evals_; ++e) eng_[e]->resident_enqueue(P.io[e].lane, P.stream);
            ds_expand_kernel<Rules><<<blocks, 128, 0, P.stream>>>(P.p);
            if (P.p.begin_lead) CB2_CUDA(cudaStreamWaitEvent(P.stream, P.join, 0));
            CB2_CUDA(cudaGetLastError());
        } catch (...) {
            cudaStreamEndCapture(P.stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        cudaError_t ce = cudaStreamEndCapture(P.stream, &g);
        if (ce != cudaSuccess) throw Error(CATTUS_B200_ECUDA, std::string("device search: graph capture failed: ") + cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&P.graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) throw Error(CATTUS_B200_ECUDA, std::string("device search: graph instantiate failed: ") + cudaGetErrorString(ce));
        kernels_per_wave_ = 3;
        for (uint32_t e = 0; e < n_evals_; ++e) kernels_per_wave_ += P.io[e].kernels;
    }
    void destroy() {
        for (DsPopulation<Rules>& P : pop_) {
            if (P.stream) cudaStreamSynchronize(P.stream);
            if (P.graph) cudaGraphExecDestroy(P.graph);
            P.graph = nullptr;
            if (P.stream) cudaStreamDestroy(P.stream);
            if (P.side) cudaStreamDestroy(P.side);
            P.stream = P.side = nullptr;
            if (P.fork) cudaEventDestroy(P.fork);
            if (P.join) cudaEventDestroy(P.join);
            P.fork = P.join = nullptr;
            for (void** q : {&P.d_cmds, &P.d_status, &P.d_cache_meta[0], &P.d_cache_meta[1], &P.d_cache_entries[0], &P.d_cache_entries[1]}) {
                if (*q) cudaFree(*q);
                *q = nullptr;
            }
            for (uint32_t e = 0; e < 2; ++e)
                if (P.io[e].lane >= 0 && eng_[e]) {
                    eng_[e]->resident_release(P.io[e].lane);
                    P.io[e].lane = -1;
                }
        }
        for (auto& ev : events_)
            if (ev) cudaEventDestroy(ev);
        events_.clear();
        for (void** q : {&d_rules_, &d_slots_, &d_pools_, &d_paths_, &d_noise_, &d_hist_}) {
            if (*q) cudaFree(*q);
            *q = nullptr;
        }
        if (h_results_) cudaFreeHost(h_results_);
        if (h_cmds_) cudaFreeHost(h_cmds_);
        if (h_status_) cudaFreeHost(h_status_);
        h_results_ = h_cmds_ = h_status_ = nullptr;
    }

    DsPopulation<Rules> pop_[2];
    Engine* eng_[2] = {nullptr, nullptr};
    uint32_t n_evals_ = 1, n_pops_ = 1, split_ = 0, depth_ = 2, n_bufs_ = 3, max_children_ = 0, n_slots_ = 0, pool_words_ = 0, cmd_stride_ = 0,
             result_stride_ = 0, kernels_per_wave_ = 0;
    size_t cmd_bytes_ = 0, result_buf_bytes_ = 0;
    void *d_rules_ = nullptr, *d_slots_ = nullptr, *d_pools_ = nullptr, *d_paths_ = nullptr, *d_noise_ = nullptr, *d_hist_ = nullptr;
    uint8_t *h_results_ = nullptr, *h_cmds_ = nullptr, *h_status_ = nullptr;
    std::vector<cudaEvent_t> events_;
    std::vector<uint32_t> wave_pop_;
    uint64_t waves_ = 0;
    uint32_t max_used_ = 0;
    double setup_seconds_ = 0.0;
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
    if (std::getenv("CATTUS_B200_DSEARCH_PROFILE"))
        std::fprintf(stderr, "device search: %u slots in %u population(s), %u words per tree buffer, set up in %.3f s\n", n_slots, be.populations(),
                     be.pool_words(), be.setup_seconds());
    ds::Driver<Rules, DsCudaBackend<Rules>> drv(rules, cfg, params, be, sh);
    drv.run();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const uint64_t waves = be.waves();
    if (std::getenv("CATTUS_B200_DSEARCH_PROFILE"))
        std::fprintf(stderr, "device search: fullest tree buffer used %u of %u words (%.0f %%)\n", be.max_used(), be.pool_words(),
                     100.0 * be.max_used() / std::max(1u, be.pool_words()));
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
