// Host side of lsthm_gemm3 (include/lsthm_b200.h): tile/split planning and launch of the tcgen05 split-bf16 GEMM.
#include <algorithm>

#include "../../include/lsthm_b200.h"
#include "gemm3_kernels.cuh"

namespace lsthm {
int set_error(const char *what, cudaError_t e);
int fail_msg(const char *msg);

template <int AMN, int BMN, int PREC>
static int launch_gemm(const GemmArgs &g, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = GemmCfg<PREC>::kSmem;
    cudaError_t e = cudaFuncSetAttribute(gemm3_kernel<AMN, BMN, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lsthm_gemm3 shared-memory opt-in", e);
    gemm3_kernel<AMN, BMN, PREC><<<grid, kGemmThreads, smem, st>>>(g);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_gemm3 launch", e);
}

template <int BMN, bool BF16>
static int launch_gemm_w(const GemmWArgs &g, const float *W, int ldw, uint8_t *img, dim3 grid, cudaStream_t st) {
    const int nch = (g.N + kWNC - 1) / kWNC;
    const size_t chunks = (size_t)nch * g.nkb * 1024;
    const int pblocks = (int)std::min<size_t>((chunks + 255) / 256, 148 * 8);
    gemm3w_pack_kernel<BMN><<<pblocks, 256, 0, st>>>(W, ldw, g.N, g.K, g.nkb, img, chunks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("lsthm_gemm3w pack launch", e);
    e = cudaFuncSetAttribute(gemm3w_kernel<BMN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWSmemBytes);
    if (e != cudaSuccess) return set_error("lsthm_gemm3w shared-memory opt-in", e);
    gemm3w_kernel<BMN, BF16><<<grid, kGemmThreads, kWSmemBytes, st>>>(g);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : set_error("lsthm_gemm3w launch", e);
}
}  // namespace lsthm

using namespace lsthm;

extern "C" {

size_t lsthm_gemm3_workspace_floats(int32_t mode, int32_t M, int32_t N, int32_t K) {
    mode &= ~(LSTHM_GEMM_BF16 | LSTHM_GEMM_X6);
    if (mode == 3) return 0;
    const long tiles = (long)((M + kGemmBM - 1) / kGemmBM) * ((N + kGemmBN - 1) / kGemmBN);
    if (tiles >= 148 || K < 4 * kGemmBK * 8) return 0;
    const int splits = (int)std::min<long>((K + 255) / 256, std::max<long>(1, 296 / tiles));   // one full wave: 2 CTAs x 148 SMs
    return splits > 1 ? (size_t)splits * M * N : 0;
}

int lsthm_gemm3(int32_t mode, int32_t M, int32_t N, int32_t K, const float *A, int32_t lda, const float *B, int32_t ldb,
                const float *bias, float *C, int32_t ldc, float *workspace, size_t workspace_floats, void *stream) {
    const bool bf16 = (mode & LSTHM_GEMM_BF16) != 0, x6 = (mode & LSTHM_GEMM_X6) != 0;
    mode &= ~(LSTHM_GEMM_BF16 | LSTHM_GEMM_X6);
    if (mode < 0 || mode > 3) return fail_msg("lsthm_gemm3: mode must be 0 (NT), 1 (NN), 2 (TN) or 3 (NT + ReLU)");
    if (bf16 && x6) return fail_msg("lsthm_gemm3: LSTHM_GEMM_BF16 and LSTHM_GEMM_X6 exclude each other");
    if (M < 1 || N < 1 || K < 1 || !A || !B || !C) return fail_msg("lsthm_gemm3: bad shape or null pointer");
    if ((lda & 3) || (ldb & 3)) return fail_msg("lsthm_gemm3: lda and ldb must be multiples of 4 floats");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15)
        return fail_msg("lsthm_gemm3: operands must be 16-byte aligned");
    GemmArgs g;
    g.relu = mode == 3 ? 1 : 0;
    g.A = A; g.B = B; g.bias = bias; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    const int tm = (M + kGemmBM - 1) / kGemmBM, tn = (N + kGemmBN - 1) / kGemmBN;
    const long tiles = (long)tm * tn;
    int splits = 1;
    if (tiles < 148 && K >= 4 * kGemmBK * 8 && workspace && mode != 3) {
        splits = (int)std::min<long>((K + 255) / 256, std::max<long>(1, 296 / tiles));   // one full wave
        while (splits > 1 && (size_t)splits * M * N > workspace_floats) --splits;
    }
    int kps = (K + splits - 1) / splits;
    kps = (kps + kGemmBK - 1) / kGemmBK * kGemmBK;
    splits = (K + kps - 1) / kps;
    g.k_per_split = kps; g.splits = splits;
    if (splits > 1) { g.C = workspace; g.ldc = N; }
    const dim3 grid(tn, tm, splits);
    int rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0 || mode == 3) rc = x6 ? launch_gemm<0, 0, 2>(g, grid, st) : bf16 ? launch_gemm<0, 0, 1>(g, grid, st) : launch_gemm<0, 0, 0>(g, grid, st);
    else if (mode == 1) rc = x6 ? launch_gemm<0, 1, 2>(g, grid, st) : bf16 ? launch_gemm<0, 1, 1>(g, grid, st) : launch_gemm<0, 1, 0>(g, grid, st);
    else rc = x6 ? launch_gemm<1, 1, 2>(g, grid, st) : bf16 ? launch_gemm<1, 1, 1>(g, grid, st) : launch_gemm<1, 1, 0>(g, grid, st);
    if (rc) return rc;
    if (splits > 1) {
        const size_t total = (size_t)M * N;
        if (splits >= 32 && total <= 65536) {
            gemm3_reduce_wide_kernel<<<(int)((total + 31) / 32), 256, 0, (cudaStream_t)stream>>>(workspace, bias, C, M, N, ldc, splits);
        } else {
            const int blocks = (int)std::min<size_t>((total + 255) / 256, 1184);
            gemm3_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(workspace, bias, C, M, N, ldc, splits);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_error("lsthm_gemm3 reduce launch", e);
    }
    return 0;
}

size_t lsthm_gemm3w_pack_bytes(int32_t N, int32_t K) {
    if (N < 1 || K < 1) return 0;
    return (size_t)((N + kWNC - 1) / kWNC) * ((K + kGemmBK - 1) / kGemmBK) * 2 * (size_t)kWImgBytes;
}

int lsthm_gemm3w(int32_t mode, int32_t M, int32_t N, int32_t K, const float *A, int32_t lda, const float *W, int32_t ldw,
                 const float *bias, float *C, int32_t ldc, void *pack, size_t pack_bytes, void *stream) {
    const bool bf16 = (mode & LSTHM_GEMM_BF16) != 0;
    mode &= ~LSTHM_GEMM_BF16;
    if (mode != 0 && mode != 1 && mode != 3) return fail_msg("lsthm_gemm3w: mode must be 0 (NT), 1 (NN) or 3 (NT + ReLU)");
    if (M < 1 || N < 1 || K < 1 || !A || !W || !C || !pack) return fail_msg("lsthm_gemm3w: bad shape or null pointer");
    if ((lda & 3) || (ldw & 3)) return fail_msg("lsthm_gemm3w: lda and ldw must be multiples of 4 floats");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(pack)) & 15)
        return fail_msg("lsthm_gemm3w: operands must be 16-byte aligned");
    if (pack_bytes < lsthm_gemm3w_pack_bytes(N, K)) return fail_msg("lsthm_gemm3w: pack buffer too small");
    GemmWArgs g;
    g.A = A; g.bias = bias; g.img = static_cast<const uint8_t *>(pack); g.C = C;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldc = ldc; g.nkb = (K + kGemmBK - 1) / kGemmBK; g.relu = mode == 3 ? 1 : 0;
    const dim3 grid((N + kWNC - 1) / kWNC, (M + kGemmBM - 1) / kGemmBM, 1);
    uint8_t *img = static_cast<uint8_t *>(pack);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) return bf16 ? launch_gemm_w<1, true>(g, W, ldw, img, grid, st) : launch_gemm_w<1, false>(g, W, ldw, img, grid, st);
    return bf16 ? launch_gemm_w<0, true>(g, W, ldw, img, grid, st) : launch_gemm_w<0, false>(g, W, ldw, img, grid, st);
}

}  // extern "C"
