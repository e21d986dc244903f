// Host runtime of the evaluator: weight blob -> device layouts, per-stream batch buffers ("lanes"), the launch
// sequence for one batch (captured into a CUDA graph per batch-size bucket), the pinned batch queue that the
// blocking per-leaf `eval` feeds, and metrics.  Everything here sits below the C ABI in include/cattus_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cattus_b200.h"
#include "kernels.cuh"
#include "tc_gemm.cuh"

namespace cb2 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CB2_CUDA(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            throw ::cb2::Error(CATTUS_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" +   \
                                                      __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
    } while (0)

struct NetDims {
    uint32_t game = 0, s = 0, c_in = 0, moves = 0, f = 0, r = 0, vh = 0, ph = 0, hidden = 128;
    uint32_t s2() const { return s * s; }
    uint32_t wpp() const { return (s * s + 63) / 64; }
    uint32_t bitmap_bytes() const { return (moves + 7) / 8; }
};

// .cb2 weight blob (written by cattus_b200/export.py): 64-byte header + BN-folded fp32 tensors in PyTorch layouts.
struct Blob {
    NetDims d;
    std::vector<float> data;
    // offsets (in floats) into data
    struct Conv { size_t w, b; uint32_t co, ci, k; };
    Conv stem;
    std::vector<Conv> block_conv;  // 2 per block
    Conv vconv, pconv;
    size_t vfc1_w, vfc1_b, vfc2_w, vfc2_b, pfc_w, pfc_b;
    static Blob parse(const void* bytes, size_t n);
};

enum class OpKind : int { EncodeNchw, EncodeNhwc, DenseToNhwc, ConvF32, FcF32, TcGemm, LegalOffsets, PolicyTail, ValueTail };

struct Op {
    OpKind kind;
    int stage;  // 0 encode, 1 trunk, 2 heads, 3 tail
    dim3 grid, block;
    size_t smem = 0;
    TcGemmParams tc;  // TcGemm
    // small-kernel arguments
    const void* in0 = nullptr;
    const void* in1 = nullptr;
    const void* in2 = nullptr;
    const void* in3 = nullptr;
    void* out0 = nullptr;
    int i[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float f0 = 0.0f;
};

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void alloc(size_t n);
    void free_();
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// One evaluator stream with everything a batch needs: pinned I/O blocks, device I/O blocks, activations, graphs.
struct Lane {
    int index = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // pinned host
    uint8_t* h_in = nullptr;    // [16 B header: n][records]
    float* h_values = nullptr;  // [max_batch]
    float* h_probs = nullptr;   // [max_batch * moves]
    // device
    DeviceBuf d_in;                    // same layout as h_in
    DeviceBuf d_values, d_offsets, d_probs;
    DeviceBuf d_x;                     // bf16: encoded input NHWC [rows][64]; fp32: NCHW f32
    DeviceBuf d_act[3];                // trunk activations
    DeviceBuf d_hv, d_hp;              // head conv outputs
    DeviceBuf d_hidden;                // value FC1 output [max_batch][128] f32
    DeviceBuf d_logits;                // [max_batch][ld_logits] f32
    DeviceBuf d_dense;                 // run_dense / encode staging (f32 NCHW)
    std::map<uint32_t, std::vector<Op>> ops;     // per bucket
    std::map<uint32_t, cudaGraphExec_t> graphs;  // per bucket
    uint32_t resident_n = 0;
};

struct LeafRequest {
    const uint64_t* planes;
    const uint8_t* legal;
    float* probs_out;
    uint32_t probs_cap;
    uint32_t* n_probs;
    float* value_out;
    int status = 1;  // 1 = pending, 0 = ok, <0 = error
    std::string error;
};

class Engine {
  public:
    Engine(const cattus_b200_desc& desc, const void* blob, size_t blob_bytes);
    ~Engine();

    void get_info(cattus_b200_info* info) const;
    void get_metrics(cattus_b200_metrics* m) const;

    void eval_leaf(LeafRequest* req);  // blocking
    void eval_batch(const uint64_t* planes, const uint8_t* legal, uint32_t n, float* probs_out, size_t probs_cap,
                    uint32_t* prob_offsets, float* values_out);
    void encode(const uint64_t* planes, uint32_t n, uint32_t batch, float* nchw_out);
    void run_dense(const float* nchw, uint32_t n, float* logits_out, float* values_out);
    void resident_upload(const uint64_t* planes, const uint8_t* legal, uint32_t n);
    void eval_resident(uint32_t n, cudaStream_t stream);
    void resident_download(uint32_t n, float* probs_out, size_t probs_cap, uint32_t* prob_offsets, float* values_out);
    void time_stage(uint32_t stage, uint32_t n, uint32_t iters, float* ms_out);

  private:
    // setup
    void upload_weights(const Blob& blob);
    void init_lane(Lane& lane);
    uint32_t bucket_for(uint32_t n) const;
    std::vector<Op>& ops_for(Lane& lane, uint32_t bucket);
    void build_ops_bf16(Lane& lane, uint32_t bucket, std::vector<Op>& ops, bool dense_input);
    void build_ops_fp32(Lane& lane, uint32_t bucket, std::vector<Op>& ops, bool dense_input);
    void add_tail_ops(Lane& lane, uint32_t bucket, std::vector<Op>& ops, const float* logits, int ld_logits);
    void launch_op(const Op& op, cudaStream_t stream);
    void run_bucket(Lane& lane, uint32_t bucket, cudaStream_t stream, bool use_graph);
    void check_device_error(const char* where);

    // batch plumbing
    uint32_t pack_records(Lane& lane, const uint64_t* planes, const uint8_t* legal, uint32_t n,
                          std::vector<uint32_t>& counts);  // returns total legal moves
    void submit(Lane& lane, uint32_t n, uint32_t total_probs);  // H2D + graph + D2H + event (async)
    Lane& acquire_lane(int want = -1);
    void release_lane(Lane& lane);
    void evaluator_loop();
    void note_batch(uint32_t n, double seconds);

    // tensor maps
    CUtensorMap make_map_2d(const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_rows);
    CUtensorMap make_map_conv(const void* base, uint32_t channels, uint32_t boards, uint32_t nb);
    Op make_tc_op(int stage, int mode, const CUtensorMap& a, const CUtensorMap& b, int num_kb, int kh, int n_umma,
                  int n_tiles, int n_store, int m_valid, int rows_per_tile, int nb, int ld_out, bool out_f32, bool relu,
                  const float* bias, const void* resid, void* out, uint32_t a_box_rows);

    NetDims d_;
    cattus_b200_desc desc_;
    int device_ = 0;
    int sm_count_ = 0;
    uint32_t max_batch_ = 0;
    uint32_t precision_ = 0;
    RecLayout rec_;
    bool derive_legal_ = false;

    // derived layout constants (bf16 path)
    uint32_t ca_ = 64;       // trunk activation channels (multiple of 64)
    uint32_t nb_ = 1;        // boards per 128-row tile
    uint32_t fp_ = 16;       // UMMA N for trunk convs
    uint32_t vhp_ = 8, php_ = 8;
    uint32_t n_pol_ = 16, pol_tiles_ = 1, ld_logits_ = 16;

    // weights on device
    struct ConvDev { DeviceBuf w, b; uint32_t co, ci, k; CUtensorMap map; uint32_t n_umma; uint32_t k_total; };
    std::vector<ConvDev> convs_;  // stem, then 2 per block (bf16: [Np][9*Cin_pad] bf16; fp32: torch layout f32)
    ConvDev vconv_, pconv_, vfc1_, pfc_;
    DeviceBuf vfc2_w_;
    float vfc2_b_ = 0.0f;

    std::vector<std::unique_ptr<Lane>> lanes_;
    std::vector<char> lane_busy_;
    std::mutex lane_mu_;
    std::condition_variable lane_cv_;

    // leaf queue
    std::mutex q_mu_;
    std::condition_variable q_cv_, done_cv_;
    std::deque<LeafRequest*> queue_;
    std::vector<std::thread> evaluators_;
    bool stopping_ = false;

    // error word (mapped pinned) written by kernels before a trap
    uint32_t* h_err_ = nullptr;
    uint32_t* d_err_ = nullptr;

    // metrics
    mutable std::mutex m_mu_;
    cattus_b200_metrics metrics_{};
    uint32_t kernels_per_batch_ = 0;

    // driver entry point
    void* encode_tiled_ = nullptr;
};

}  // namespace cb2
