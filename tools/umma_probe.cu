// Hardware probe (development tool, not product): does a K-major SWIZZLE_NONE UMMA descriptor accept an 8-row-group stride
// (SBO) of 144 B and a start address that is only 16-byte aligned?  That is what trunk_fused.cuh's "tap = start-address
// shift" trick needs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../cattus_b200/csrc/ptx.cuh"
using namespace cb2;

constexpr int kCells = 181, kPlaneBytes = 162 * 16;  // chunk planes overlap like in the real kernel
constexpr int kABytes = kPlaneBytes + kCells * 16;
constexpr int kN = 16;
constexpr int kBBytes = 2 * kN * 16;

__device__ __forceinline__ uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;  // layout type 0 = SWIZZLE_NONE
}

__global__ void probe(const uint8_t* a_img, const uint8_t* b_img, int start_cell, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, float* out) {
    __shared__ __align__(128) uint8_t sa[kABytes];
    __shared__ __align__(128) uint8_t sb[kBBytes];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < kABytes; i += blockDim.x) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < kBBytes; i += blockDim.x) sb[i] = b_img[i];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(&tmem_ptr, 32); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, kN);
        ptx::umma_bf16_ss(tm, desc_none(ptx::smem_u32(sa) + start_cell * 16, a_lbo, a_sbo), desc_none(ptx::smem_u32(sb), b_lbo, b_sbo), idesc, 0);
        ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0, nullptr, 0);
    ptx::tc_fence_after();
    float v[16];
    ptx::tmem_ld_x16(tm + ((warp * 32u) << 16), v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + (threadIdx.x & 31)) * 16 + j] = v[j];
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 32);
}

static float bf(uint16_t h) { uint32_t u = static_cast<uint32_t>(h) << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<uint16_t> a(kABytes / 2), b(kBBytes / 2);
    srand(1);
    auto rnd = [] { int v = rand() % 7 - 3; float f = static_cast<float>(v); uint32_t u; memcpy(&u, &f, 4); return static_cast<uint16_t>(u >> 16); };
    for (auto& x : a) x = rnd();
    for (auto& x : b) x = rnd();
    uint8_t *da, *db; float* dout;
    cudaMalloc(&da, kABytes); cudaMalloc(&db, kBBytes); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(da, a.data(), kABytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), kBBytes, cudaMemcpyHostToDevice);
    int bad_total = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int start_cell : {19, 0, 38, 24, 1}) {
            const uint32_t a_lbo = variant == 0 ? kPlaneBytes : 144, a_sbo = variant == 0 ? 144 : kPlaneBytes;
            const uint32_t b_lbo = variant == 0 ? kN * 16 : 128, b_sbo = variant == 0 ? 128 : kN * 16;
            cudaMemset(dout, 0, 128 * 16 * 4);
            probe<<<1, 128>>>(da, db, start_cell, a_lbo, a_sbo, b_lbo, b_sbo, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d start %d: CUDA error %s\n", variant, start_cell, cudaGetErrorString(e)); return 1; }
            std::vector<float> out(128 * 16);
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < kN; ++n) {
                    float acc = 0;
                    for (int k = 0; k < 16; ++k) {
                        const int g = r / 8, x = r % 8;
                        const float av = bf(a[((k / 8) * 162 + start_cell + g * 9 + x) * 8 + k % 8]);
                        const float bv = bf(b[((k / 8) * kN + n) * 8 + k % 8]);
                        acc += av * bv;
                    }
                    if (acc != out[r * 16 + n]) ++bad;
                }
            printf("variant %d (%s) start_cell %2d: %d / %d mismatches\n", variant, variant == 0 ? "LBO=K-chunk stride, SBO=8-row-group stride" : "swapped", start_cell, bad, 128 * kN);
            if (variant == 0) bad_total += bad;
        }
    }
    printf(bad_total == 0 ? "PROBE OK: design assumption holds\n" : "PROBE FAILED for the assumed semantics\n");
    return 0;
}
