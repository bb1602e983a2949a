// Hardware-semantics probe (tests/test_kernels_gpu.py::test_umma_row_shifted_descriptor): does a K-major SWIZZLE_128B
// UMMA operand descriptor whose start address is shifted by whole 128-byte rows, and whose 8-row groups are spaced by a
// stride that is not a multiple of the 1024-byte swizzle atom, read the rows TMA wrote?  This is what a halo-tile
// implicit-GEMM convolution needs: the nine taps become row-shifted views of one tile.
//   D[m][n] = sum_k A[row(m)][k] * B[n][k],   row(m) = (m / 8) * (sbo_bytes / 128) + m % 8 + shift_rows
#include "gemm_tc.cuh"

using namespace dsb;

namespace {

struct __align__(16) ProbeBars {
    uint64_t full, done;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__global__ void __launch_bounds__(128) umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
                                                              int shift_rows, int sbo_bytes, int base_offset_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                         // 512 rows x 128 B
    uint8_t* sB = smem + 512 * 128;             // 32 rows x 128 B
    ProbeBars* bars = reinterpret_cast<ProbeBars*>(sB + 32 * 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars->full, 1);
        mbar_init(&bars->done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&bars->tmem_base, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars->full, 512 * 128 + 32 * 128);
        tma_load_2d(sA, &tmA, &bars->full, 0, 0);
        tma_load_2d(sA + 256 * 128, &tmA, &bars->full, 0, 256);
        tma_load_2d(sB, &tmB, &bars->full, 0, 0);
        mbar_wait(&bars->full, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift_rows * 128u;
        uint64_t da = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
        if (base_offset_mode) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t db = umma_smem_desc(smem_u32(sB), 128);
        const uint32_t idesc = umma_idesc_bf16(32);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(&bars->done);
    }
    __syncwarp();
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    uint32_t raw[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), raw);
    tmem_ld_wait();
    const int m = warp * 32 + lane;
    for (int n = 0; n < 32; ++n) out[m * 32 + n] = __uint_as_float(raw[n]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

}  // namespace

// A: device bf16 [512][64], B: device bf16 [32][64], out: device fp32 [128][32]
extern "C" int dsb_test_umma_shift(const void* A, const void* B, float* out, int shift_rows, int sbo_bytes, int base_offset_mode,
                                   void* stream) {
    if (!A || !B || !out || shift_rows < 0 || sbo_bytes < 1024 || sbo_bytes % 16) return -1;
    if ((15 * (sbo_bytes / 128) + 7 + shift_rows) >= 512) return -1;
    CUtensorMap tmA, tmB;
    const uint64_t dA[2] = {64, 512}, dB[2] = {64, 32}, st[1] = {128};
    const uint32_t bA[2] = {64, 256}, bB[2] = {64, 32};
    if (int r = make_tensor_map(&tmA, A, 2, dA, st, bA)) return r;
    if (int r = make_tensor_map(&tmB, B, 2, dB, st, bB)) return r;
    const size_t smem = 512 * 128 + 32 * 128 + sizeof(ProbeBars) + 1024;
    cudaError_t e = cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    umma_shift_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, out, shift_rows, sbo_bytes, base_offset_mode);
    return (int)cudaGetLastError();
}
